"""Whole-path parity on the B200 (-m gpu): GeneralizedRCNNWSL through the C ABI versus
(a) the golden vectors of the unmodified reference and (b) the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): pseudo-GT argmax ROI indices, Matcher labels and matched indices
BIT-EXACT; losses / proposal scores within 1e-3 relative in fp32 mode."""
import numpy as np
import pytest
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from oracle import wsl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-3  # the north_star tolerance for fp32 scores

CASES = ["wsddn_v16_300", "oicr_r18_small", "oicr_v16_small", "oicr_r18_reg", "oicr_r18_batch2", "oicr_r50_small",
         "oicr_r101_coco_small"]


def _build(case, precision="fp32"):
    cfg = helpers.case_config(case, device=DEV, precision=precision)
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    return cfg, model, weights


def _score_err(a, b):
    # relative error on the scores that matter (>= 1e-3 of the column max), absolute floor for the tiny tail
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    floor = 1e-3 * np.abs(b).max(axis=0, keepdims=True) + 1e-30
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor)))


# "fp32": SIMT fp32 kernels; "fp32_tc": the same layers as split-bf16 GEMMs on the tensor cores (csrc/drn_split.cu) --
# both are held to the fp32 bar (bit-exact indices, 1e-3 scores)
F32_PRECISIONS = ["fp32", "fp32_tc"]


@pytest.mark.parametrize("precision", F32_PRECISIONS)
@pytest.mark.parametrize("case", CASES)
def test_train_forward_matches_reference_golden(case, precision):
    g = helpers.load_golden(case)
    cfg, model, _ = _build(case, precision)
    model.train()
    model.roi_heads.box_head.eval()  # dropout off, as in the golden run
    inputs = helpers.case_inputs(case)
    losses = model(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV))
    assert {"loss/" + k for k in losses} == {k for k in g if k.startswith("loss/")}
    for k, v in losses.items():
        assert helpers.rel_err(v.item(), g["loss/" + k]) < RTOL, (k, v.item(), float(g["loss/" + k]))
    for i, tr in enumerate(model.roi_heads.last_trace):
        assert _score_err(tr["scores"].cpu().numpy(), g[f"img{i}/scores"]) < RTOL
        np.testing.assert_array_equal(tr["labels_gt"].cpu().numpy(), g[f"img{i}/labels_gt"])
        for k, st in enumerate(tr["stages"]):
            p = f"img{i}/stage{k}/"
            np.testing.assert_array_equal(st["pgt_idx"].cpu().numpy(), g[p + "pgt_idx"])  # bit-exact argmax ROI indices
            np.testing.assert_array_equal(st["labels"].cpu().numpy(), g[p + "labels"])
            np.testing.assert_array_equal(st["matched"].cpu().numpy(), g[p + "matched"])
            np.testing.assert_array_equal(st["pgt_boxes"].cpu().numpy(), g[p + "pgt_boxes"])
            assert _score_err(st["probs"].cpu().numpy(), g[p + "probs"]) < RTOL
            assert helpers.rel_err(st["pgt_weights"].cpu().numpy(), g[p + "pgt_weights"]) < RTOL


@pytest.mark.parametrize("precision", F32_PRECISIONS)
@pytest.mark.parametrize("case", CASES)
def test_eval_forward_matches_reference_golden(case, precision):
    g = helpers.load_golden(case)
    cfg, model, _ = _build(case, precision)
    model.eval()
    inputs = helpers.case_inputs(case)
    batched = helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV, train=False)
    results, all_scores, all_boxes = model.inference(batched, do_postprocess=False)
    for i in range(len(inputs)):
        assert _score_err(all_scores[i][0].cpu().numpy(), g[f"img{i}/eval/all_scores"]) < RTOL
        np.testing.assert_allclose(all_boxes[i][0].cpu().numpy(), g[f"img{i}/eval/all_boxes"], rtol=1e-6, atol=1e-4)
        # detections: same classes in the same order unless two scores are within the tolerance of each other
        ds, dc = results[i].scores.cpu().numpy(), results[i].pred_classes.cpu().numpy()
        gs, gc = g[f"img{i}/eval/det_scores"], g[f"img{i}/eval/det_classes"]
        assert len(ds) == len(gs)
        np.testing.assert_allclose(ds, gs, rtol=RTOL)
        gaps = np.abs(np.diff(gs)) / gs[:-1]
        if len(gs) > 1 and gaps.min() > 10 * RTOL:
            np.testing.assert_array_equal(dc, gc)
    out = model(batched)  # post-processed public API (rcnn.py:199-240)
    assert len(out) == len(inputs) and "instances" in out[0] and out[0]["instances"].has("pred_boxes")


def test_full_size_config1_matches_oracle():
    """BASELINE.json configs[1]: R18-WS, 600x1000, 2000 proposals, fp32 -- against the CPU oracle."""
    cfg = drn.builtin_config("oicr_WSR_18_DC5_1x", ["MODEL.DEVICE", DEV])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    model.train()
    model.roi_heads.box_head.eval()
    inputs = [helpers.synth.make_inputs(600, 1000, 2000, seed=0, num_gt=2)]
    losses = model(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV))
    with torch.no_grad():
        ref_losses, ref_tr = O.forward_train(inputs, dict(weights), O.spec_from_cfg(cfg))
    for k, v in ref_losses.items():
        assert helpers.rel_err(losses[k].item(), v.item()) < RTOL, (k, losses[k].item(), v.item())
    tr = model.roi_heads.last_trace[0]
    assert _score_err(tr["scores"].cpu().numpy(), ref_tr[0]["scores"].numpy()) < RTOL
    for k, st in enumerate(ref_tr[0]["stages"]):
        assert torch.equal(tr["stages"][k]["pgt_idx"].cpu(), st["pgt_idx"])
        assert torch.equal(tr["stages"][k]["labels"].cpu(), st["labels"])
    # size-independent properties: every proposal's WSDDN column sums to the image score, probs rows sum to 1
    assert torch.allclose(tr["scores"].sum(0).clamp(1e-6, 1 - 1e-6), tr["img_score"], rtol=1e-5)
    assert torch.allclose(tr["stages"][-1]["probs"].sum(1), torch.ones(2000, device=DEV), rtol=1e-5)


@pytest.mark.parametrize("case", ["oicr_r18_small", "oicr_r50_small", "oicr_v16_small"])
def test_bf16_tensor_core_path_close_to_fp32_oracle(case):
    """bf16 tcgen05 mode (BASELINE configs[2]): operands are rounded to bf16 layer by layer, so the
    1e-3 fp32 bar does not apply (SURVEY.md §7 'hard parts'); tolerance 6e-2 on losses, and the
    pseudo-GT argmax must agree wherever the golden run's top-2 margin exceeds that tolerance."""
    g = helpers.load_golden(case)
    cfg, model, _ = _build(case, precision="bf16")
    model.train()
    model.roi_heads.box_head.eval()
    inputs = helpers.case_inputs(case)
    losses = model(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV))
    tol = 6e-2
    assert helpers.rel_err(losses["loss_cls"].item(), g["loss/loss_cls"]) < tol
    tr = model.roi_heads.last_trace[0]
    margin = g["img0/stage0/argmax_margin_rel"]
    got, want = tr["stages"][0]["pgt_idx"].cpu().numpy(), g["img0/stage0/pgt_idx"]
    assert np.array_equal(got[margin > 2 * tol], want[margin > 2 * tol])
    s, gs = tr["scores"].cpu().numpy(), g["img0/scores"]
    big = gs > 0.05 * gs.max()
    assert np.max(np.abs(s[big] - gs[big]) / gs[big]) < 0.25


def test_full_size_config2_bf16_heads_consistent_with_oracle_building_blocks():
    """BASELINE.json configs[2] at FULL size (R50-WS, 600x1000, R=4000, bf16 tensor-core mode).  The fp32 oracle of
    the whole net would take ~50 s and bf16 is not comparable at 1e-3 anyway, so the full-size check is
    property-based: every head stage must be EXACTLY what the oracle's building blocks (the reference's functions)
    produce from the GPU's own upstream tensors -- argmax indices, labels and matched indices bit-exact,
    scores / probabilities / losses within 1e-4 -- and the GEMMs must agree with a torch matmul of the same bf16
    operands on sampled rows."""
    import torch.nn.functional as F

    cfg = drn.builtin_config("oicr_WSR_50_DC5_1x", ["MODEL.DEVICE", DEV, "B200.PRECISION", "bf16"])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    model.train()
    model.roi_heads.box_head.eval()
    spec = O.spec_from_cfg(cfg)
    inp = helpers.synth.make_inputs(600, 1000, 4000, seed=0, num_gt=2)
    losses = model(helpers.to_batched([inp], drn.Instances, drn.Boxes, device=DEV))
    tr = model.roi_heads.last_trace[0]
    K = spec.num_classes
    logits = tr["logits"].cpu()
    feat = tr["feat"]
    assert torch.isfinite(logits).all() and torch.isfinite(feat.float()).all()
    # heads GEMM vs torch on the same bf16 fc7 activations (sampled rows)
    rows = torch.arange(0, 4000, 97, device=DEV)
    heads = model.roi_heads._heads_packed()
    ref_logits = feat[rows].float() @ heads["w"].float().t() + heads["bias"]
    torch.testing.assert_close(tr["logits"][rows], ref_logits, rtol=2e-3, atol=2e-3)
    # WSDDN dual softmax + image score + BCE from the GPU logits (fast_rcnn.py:493-527, 689-700, 317-329)
    scores = F.softmax(logits[:, :K], dim=1) * F.softmax(logits[:, K:2 * K], dim=0)
    assert _score_err(tr["scores"].cpu().numpy(), scores.numpy()) < 1e-4
    img = O.image_scores(tr["scores"].cpu())
    torch.testing.assert_close(tr["img_score"].cpu()[None], img, rtol=1e-5, atol=1e-9)
    gt_int = torch.unique(inp["gt_classes"], sorted=True)
    oh = torch.zeros(1, K).scatter_(1, gt_int[None], 1)
    assert helpers.rel_err(losses["loss_cls"].item(), F.binary_cross_entropy(img, oh, reduction="mean").item()) < 1e-4
    # refinement stages: get_pgt / label_proposals / weighted CE of the reference, fed with the GPU's own tensors
    prev_scores, prev_boxes = tr["scores"].cpu(), inp["boxes"]
    for k, st in enumerate(tr["stages"]):
        pgt_idx, pgt_scores, pgt_boxes, pgt_w = O.get_pgt(prev_scores, prev_boxes, gt_int, tr["img_score"].cpu()[None], k, spec)
        assert torch.equal(st["pgt_idx"].cpu(), pgt_idx)                       # bit-exact argmax ROI indices
        assert torch.equal(st["pgt_boxes"].cpu(), pgt_boxes)
        labels, midx = O.label_proposals(inp["boxes"], pgt_boxes, gt_int, spec)
        assert torch.equal(st["labels"].cpu(), labels) and torch.equal(st["matched"].cpu(), midx)
        off = 2 * K + k * (K + 1)
        lg = logits[:, off:off + K + 1]
        loss, _ = O.oicr_stage_loss(lg, labels, torch.index_select(pgt_w, 0, midx))
        assert helpers.rel_err(losses[f"loss_cls_r{k}"].item(), loss.item()) < 1e-4
        probs = F.softmax(lg, dim=-1)
        assert _score_err(st["probs"].cpu().numpy(), probs.numpy()) < 1e-4
        prev_scores = st["probs"].cpu()
        prev_boxes = O.apply_deltas(torch.zeros(4000, 4 * K), inp["boxes"], spec.bbox_reg_weights)


# ------------------------------------------------------------------ CUDA-graph plans (the default launch path)
def test_graph_plan_replay_tracks_new_inputs_and_matches_eager():
    """A captured plan must (a) give the eager result, (b) follow NEW input values of the same shape on
    replay (static buffers are refilled), (c) accept host tensors, (d) re-capture on a new shape."""
    case = "oicr_r18_small"
    cfg, model, _ = _build(case)
    model.train()
    model.roi_heads.box_head.eval()
    inp_a = helpers.case_inputs(case)
    (H, W, R, G, seed) = helpers.CASES[case][3][0]
    from drn_wsod_pytorch_b200 import synth
    inp_b = [synth.make_inputs(H, W, R, seed=seed + 100, num_gt=G)]
    inp_c = [synth.make_inputs(H + 16, W, R + 5, seed=seed + 200, num_gt=G)]

    def run(inputs, graph, device=DEV):
        model.use_cuda_graph = graph
        out = model(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=device))
        idx = [st["pgt_idx"].clone() for st in model.roi_heads.last_trace[0]["stages"]]
        return {k: v.item() for k, v in out.items()}, idx

    eager = {n: run(i, False) for n, i in (("a", inp_a), ("b", inp_b), ("c", inp_c))}
    assert not model._plans
    run(inp_a, True)
    run(inp_c, True)
    assert not model._plans  # a signature is captured the second time it is seen (first sighting runs eagerly)
    for name, inputs in (("a", inp_a), ("b", inp_b), ("a", inp_a), ("c", inp_c), ("b", inp_b)):
        got, idx = run(inputs, True)
        for k, v in eager[name][0].items():
            assert got[k] == v, (name, k, got[k], v)  # same kernels, same order: bit-identical
        for x, y in zip(idx, eager[name][1]):
            assert torch.equal(x, y)
    assert len(model._plans) == 1  # (a, b) share one plan; c was 2 of the 5 calls since its first sighting: not dominant, stays eager
    for _ in range(3):             # ... until it makes up half of the calls since it was first seen
        got, _ = run(inp_c, True)
        assert got == eager["c"][0]
    assert len(model._plans) == 2
    got, _ = run(inp_b, True, device="cpu")  # host tensors: copied H2D into the plan's static buffers
    assert got == eager["b"][0]
    # prefetch: inputs staged on a side stream ahead of the call, picked up by identity; unrelated inputs ignore the stage
    model.use_cuda_graph = True
    host_a = helpers.to_batched(inp_a, drn.Instances, drn.Boxes, device="cpu")
    host_b = helpers.to_batched(inp_b, drn.Instances, drn.Boxes, device="cpu")
    model.prefetch(host_a)
    assert {k: v.item() for k, v in model(host_a).items()} == eager["a"][0]
    model.prefetch(host_a)
    assert {k: v.item() for k, v in model(host_b).items()} == eager["b"][0]
    assert {k: v.item() for k, v in model(host_a).items()} == eager["a"][0]
    # stale-weight protection: an in-place parameter update must not replay the old packed weights
    with torch.no_grad():
        model.roi_heads.box_predictor.cls.bias[3] += 2.0  # (a uniform shift would cancel in the softmax)
    changed, _ = run(inp_a, True)
    model.use_cuda_graph = False
    ref, _ = run(inp_a, False)
    assert changed == ref and changed["loss_cls"] != eager["a"][0]["loss_cls"]


def test_graph_plan_eval_matches_eager_and_dropout_mask_advances():
    case = "oicr_r18_small"
    cfg, model, _ = _build(case)
    inputs = helpers.case_inputs(case)
    model.eval()
    outs = []
    for graph in (False, True, True):
        model.use_cuda_graph = graph
        res, sc, bx = model.inference(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV, train=False),
                                      do_postprocess=False)
        outs.append((sc[0].clone(), bx[0].clone(), res[0].scores.clone()))
    for o in outs[1:]:
        assert all(torch.equal(a, b) for a, b in zip(o, outs[0]))
    # train mode with dropout ON: consecutive replays of one plan must draw different masks
    model.train()
    model.use_cuda_graph = True
    batched = helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV)
    l1 = model(batched)["loss_cls_r0"].item()
    l2 = model(batched)["loss_cls_r0"].item()
    l3 = model(batched)["loss_cls_r0"].item()
    assert len({l1, l2, l3}) == 3


def test_pool_fc6_row_block_overlap_is_bit_identical_to_one_piece():
    """The ROIPool -> fc6 pipeline in row blocks on two streams (opt-in, DRN_B200_OVERLAP_POOL=1) must give
    exactly the tensors of the single-launch path, eagerly and under graph capture, and the dropout masks of
    the row blocks must not repeat."""
    cfg = drn.builtin_config("oicr_WSR_18_DC5_1x", ["MODEL.DEVICE", DEV, "B200.PRECISION", "bf16"])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    rh = model.roi_heads
    rh.keep_trace = True
    R = 2048
    assert len(rh._row_blocks(R, rh.box_head.fcs[0].out_features)) > 1
    assert rh._row_blocks(300, 4096) == [(0, 300)] and rh._row_blocks(2048, 2048) == [(0, 2048)]
    for blocks in (rh._row_blocks(4000, 2048), rh._row_blocks(R, 4096), rh._row_blocks(8000, 2048)):
        assert blocks[0][0] == 0 and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:])) and all(b[0] % 256 == 0 for b in blocks)
    inp = [helpers.synth.make_inputs(192, 256, R, seed=3, num_gt=2)]
    batched = helpers.to_batched(inp, drn.Instances, drn.Boxes, device=DEV)
    model.train()
    rh.box_head.eval()
    from drn_wsod_pytorch_b200 import lib
    prev = lib.load().drn_gemm_set_tail_split(0)  # the one-piece fc6 would otherwise take the tail split-K schedule (other rounding)
    try:
        out = {}
        for overlap in (False, True):
            for graph in (False, True):
                rh.overlap_pool = overlap
                model.use_cuda_graph = graph
                model.invalidate_plans()
                for _ in range(3 if graph else 1):
                    losses = model(batched)
                tr = rh.last_trace[0]
                out[(overlap, graph)] = ({k: v.item() for k, v in losses.items()}, tr["feat"].clone(), tr["logits"].clone())
                assert bool(model._plans) == graph
    finally:
        lib.load().drn_gemm_set_tail_split(prev)
    ref = out[(False, False)]
    for key, got in out.items():
        assert got[0] == ref[0], key
        assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2]), key
    # dropout on: the two row blocks of fc6 draw different masks, and the keep rate is 1/2
    rh.overlap_pool = True
    model.use_cuda_graph = False
    rh.box_head.train()
    model(batched)
    feat = rh.last_trace[0]["feat"]
    assert 0.45 < (feat == 0).float().mean().item() < 0.75  # relu zeros + dropped half
