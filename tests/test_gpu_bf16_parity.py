"""Parity of the BENCHMARKED precision (-m gpu): the bf16 tensor-core mode against the oracle run with bf16-rounded
layer operands and fp32 accumulation (`oracle.wsl_oracle.bf16_operands`, the same restatement that is pinned to the
reference's goldens in fp32), at the small golden cases AND at the full size of BASELINE.json configs[2] and [4];
plus full-size fp32-accurate (`fp32_tc`) runs of configs[3] and [4] against the fp32 oracle.

Why a bf16-operand oracle: a 20-100-layer net whose activations are stored in bf16 cannot meet the north_star's 1e-3
against an fp32 run (SURVEY.md §7 "hard parts"); what CAN be held tight is that the B200 path computes exactly the
bf16-operand / fp32-accumulate arithmetic it claims.  The two sides then differ only by fp32 summation order, which
moves a stored activation by at most one bf16 ulp (2^-8 relative) where the fp32 value sits on a rounding boundary.

Tolerances (stated per check): intermediate tensors 1e-2 of the tensor's scale (a few bf16 ulps); losses and the
proposal scores that matter 1e-2 relative; pseudo-GT argmax indices / labels equal wherever the oracle's top-2 margin
exceeds 1e-2; pooled ROI features BIT-EXACT given the same feature map (max-pool has no rounding, the objectness
multiply is one fp32 product rounded once)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers
import drn_wsod_pytorch_b200 as drn
from oracle import wsl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2


def _build(cfg_name, precision, extra=()):
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", DEV, "B200.PRECISION", precision] + list(extra))
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    model.train()
    model.roi_heads.box_head.eval()  # dropout off for the comparison
    return cfg, model, weights


def _scale_err(a, b):
    """max |a - b| relative to the reference tensor's scale (its largest magnitude)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def _score_err(a, b, keep=1e-2):
    """relative error over the entries that matter (>= `keep` of their column's max)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    big = b >= keep * b.max(axis=0, keepdims=True)
    return float(np.max(np.abs(a[big] - b[big]) / b[big]))


def _margins(prev_scores, gt_int):
    sel = torch.index_select(prev_scores, 1, gt_int)
    top2 = torch.topk(sel, 2, dim=0)[0]
    return (top2[0] - top2[1]) / top2[0].abs().clamp(min=1e-30)


def _check_heads_chain(tr, ref_tr, losses, ref_losses, gt_int, tol=TOL):
    for k, v in ref_losses.items():
        assert helpers.rel_err(losses[k].item(), v.item()) < tol, (k, losses[k].item(), v.item())
    assert _score_err(tr["scores"].cpu().numpy(), ref_tr["scores"].cpu().numpy()) < tol
    assert helpers.rel_err(tr["img_score"].cpu().numpy(), ref_tr["img_score"].cpu().numpy()[0]) < tol
    prev = ref_tr["scores"].cpu()
    for k, st in enumerate(ref_tr["stages"]):
        sure = (_margins(prev, gt_int) > tol).numpy()
        got, want = tr["stages"][k]["pgt_idx"].cpu().numpy(), st["pgt_idx"].cpu().numpy()
        assert np.array_equal(got[sure], want[sure]), (k, got, want)
        if sure.all() and np.array_equal(got, want):  # same pseudo GT -> the labels must agree bit for bit
            assert torch.equal(tr["stages"][k]["labels"].cpu(), st["labels"].cpu())
            assert torch.equal(tr["stages"][k]["matched"].cpu(), st["matched"].cpu())
        assert _score_err(tr["stages"][k]["probs"].cpu().numpy(), st["probs"].cpu().numpy()) < tol
        prev = st["probs"].cpu()


@pytest.mark.parametrize("case", ["oicr_r18_small", "oicr_r50_small", "oicr_v16_small", "oicr_r101_coco_small", "wsddn_v16_300"])
def test_bf16_mode_matches_the_bf16_operand_oracle_small(case):
    cfg = helpers.case_config(case, device=DEV, precision="bf16")
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    model.train()
    model.roi_heads.box_head.eval()
    inputs = helpers.case_inputs(case)
    losses = model(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV))
    spec = O.spec_from_cfg(cfg)
    with torch.no_grad(), O.bf16_operands():
        ref_losses, ref_tr = O.forward_train(inputs, dict(weights), spec)
    tr = model.roi_heads.last_trace[0]
    # fc7 features (bf16 on both sides) within a few ulps of the tensor's scale
    assert _scale_err(tr["feat"].float().cpu(), ref_tr[0]["feat"]) < TOL
    if spec.heads == "oicr":
        _check_heads_chain(tr, ref_tr[0], losses, ref_losses, torch.unique(inputs[0]["gt_classes"], sorted=True))
    else:
        assert helpers.rel_err(losses["loss_cls"].item(), ref_losses["loss_cls"].item()) < TOL
        assert _score_err(tr["scores"].cpu().numpy(), ref_tr[0]["scores"].numpy()) < TOL


FULL = {
    # BASELINE.json configs[2]: the benchmarked configuration
    "config2_r50": ("oicr_WSR_50_DC5_1x", 600, 1000, 4000, 20),
    # BASELINE.json configs[4]: R101-WS, COCO shape, 80 classes
    "config4_r101_coco": ("oicr_WSR_101_DC5_1x_coco", 800, 1333, 4000, 80),
}


@pytest.mark.parametrize("name", list(FULL))
def test_full_size_bf16_matches_the_bf16_operand_oracle(name):
    """FULL-size bf16 run, stage by stage against the bf16-operand oracle executed by torch's fp32 CUDA kernels (TF32 off):
    (1) the res5 map of the whole conv stack, (2) pooled ROI features bit-exact given the GPU's own map, (3) fc6 / fc7 on
    sampled rows given the GPU's own pooled rows, (4) head logits given the GPU's own fc7, (5) the independent end-to-end
    chain image -> losses / scores / pseudo-GT indices."""
    import torchvision

    cfg_name, H, W, R, K = FULL[name]
    cfg, model, weights = _build(cfg_name, "bf16")
    spec = O.spec_from_cfg(cfg)
    inp = helpers.synth.make_inputs(H, W, R, seed=0, num_gt=2, num_classes=K)
    losses = model(helpers.to_batched([inp], drn.Instances, drn.Boxes, device=DEV))
    tr = model.roi_heads.last_trace[0]
    state = {k: v.to(DEV) for k, v in weights.items()}
    dinp = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in inp.items()}
    rh = model.roi_heads
    with torch.no_grad(), torch.device(DEV), O.bf16_operands():
        # (1) conv stack
        fmap = O.backbone_forward(O.preprocess_image(dinp["image"], spec), state, spec)        # [1, C, h, w] fp32 (bf16 values)
        ours_map = model._features([dinp["image"]], (H, W))[spec.out_feature]                  # [1, C, h, w] bf16, NHWC memory
        assert ours_map.shape == fmap.shape
        d = (ours_map.float() - fmap).abs()
        scale = fmap.abs().max()
        assert float(d.max() / scale) < TOL, float(d.max() / scale)
        # almost every element is identical; the rest moved by a rounding-boundary flip somewhere upstream
        assert float((d > 0.004 * fmap.abs().clamp(min=1e-3 * float(scale))).float().mean()) < 0.02
        # (2) ROIPool x (objectness + 1): bit-exact given the same map
        pooled_ref = torchvision.ops.roi_pool(ours_map.float().contiguous(),
                                              torch.cat([torch.zeros(R, 1), dinp["boxes"]], 1), (7, 7), 1.0 / spec.stride)
        pooled_ref = (pooled_ref * (dinp["objectness"] + 1).view(-1, 1, 1, 1)).to(torch.bfloat16)  # [R, C, 7, 7]
        C = fmap.shape[1]
        ours_pooled = tr["acts"][0].view(R, 49, C)                                             # bin-major [R][49][C]
        assert torch.equal(ours_pooled, pooled_ref.permute(0, 2, 3, 1).reshape(R, 49, C))
        # (3) fc6 / fc7 on sampled rows, from the GPU's own inputs of each layer (reference (c, ph, pw) flattening)
        rows = torch.arange(0, R, 61)[:66]
        x = pooled_ref[rows].float().flatten(1)
        y6 = F.relu(F.linear(x, O._qw(state, "roi_heads.box_head.fc1.weight"), state["roi_heads.box_head.fc1.bias"]))
        assert _scale_err(tr["acts"][1][rows].float(), y6.to(torch.bfloat16).float()) < TOL
        y7 = F.relu(F.linear(tr["acts"][1][rows].float(), O._qw(state, "roi_heads.box_head.fc2.weight"), state["roi_heads.box_head.fc2.bias"]))
        assert _scale_err(tr["acts"][2][rows].float(), y7.to(torch.bfloat16).float()) < TOL
        # (4) head logits (fp32 out) from the GPU's own fc7 rows
        heads = rh._heads_packed()
        n_real = 2 * K + rh.refine_K * (K + 1)
        ref_logits = tr["feat"][rows].float() @ heads["w"][:n_real].float().t() + heads["bias"][:n_real]
        torch.testing.assert_close(tr["logits"][rows, :n_real], ref_logits, rtol=1e-3, atol=1e-3)
        del pooled_ref, x
        # (5) the independent chain
        ref_losses, ref_tr = O.forward_train([dinp], state, spec)
    _check_heads_chain(tr, ref_tr[0], losses, ref_losses, torch.unique(inp["gt_classes"], sorted=True))
    assert _scale_err(tr["feat"].float(), ref_tr[0]["feat"]) < TOL


FULL_F32 = {
    # BASELINE.json configs[3]: OICR-VGG16, VOC shape, 2000 proposals
    "config3_vgg16": ("oicr_V_16_DC5_1x", 600, 1000, 2000, 20),
    # BASELINE.json configs[4] at fp32 accuracy
    "config4_r101_coco": ("oicr_WSR_101_DC5_1x_coco", 800, 1333, 4000, 80),
}


@pytest.mark.parametrize("name", list(FULL_F32))
def test_full_size_fp32_tc_matches_cpu_oracle(name):
    """configs[3] / [4] at full size in the fp32-accurate tensor-core mode against the fp32 CPU oracle at the north_star
    bar: pseudo-GT argmax indices and labels bit-exact, losses / scores within 1e-3."""
    cfg_name, H, W, R, K = FULL_F32[name]
    cfg, model, weights = _build(cfg_name, "fp32_tc")
    inp = helpers.synth.make_inputs(H, W, R, seed=0, num_gt=2, num_classes=K)
    losses = model(helpers.to_batched([inp], drn.Instances, drn.Boxes, device=DEV))
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        ref_losses, ref_tr = O.forward_train([inp], dict(weights), O.spec_from_cfg(cfg))
    for k, v in ref_losses.items():
        assert helpers.rel_err(losses[k].item(), v.item()) < 1e-3, (k, losses[k].item(), v.item())
    tr = model.roi_heads.last_trace[0]
    assert _score_err(tr["scores"].cpu().numpy(), ref_tr[0]["scores"].numpy(), keep=1e-3) < 1e-3
    for k, st in enumerate(ref_tr[0]["stages"]):
        assert torch.equal(tr["stages"][k]["pgt_idx"].cpu(), st["pgt_idx"])
        assert torch.equal(tr["stages"][k]["labels"].cpu(), st["labels"])
        assert torch.equal(tr["stages"][k]["matched"].cpu(), st["matched"])


def test_get_pgt_nan_wins_with_its_index():
    """CPU torch.max(dim=0) lets a NaN win, with the index of the (first) NaN; columns without a NaN keep
    the usual argmax with ties resolved to the lowest index (roi_heads_oicr.py:491-567)."""
    from drn_wsod_pytorch_b200 import ops

    R, K = 700, 20
    g = torch.Generator().manual_seed(5)
    scores = torch.rand(R, K, generator=g)
    gt_int = torch.tensor([2, 7, 11], dtype=torch.int64)
    scores[123, 7] = float("nan")
    scores[400, 7] = float("nan")      # a later NaN: the first one keeps the index
    scores[650, 11] = float("inf")     # inf is an ordinary maximum
    scores[10, 2] = scores[600, 2] = 5.0  # exact tie -> lowest index
    boxes = helpers.synth.make_inputs(600, 1000, R, seed=4)["boxes"]
    img = torch.clamp(torch.nan_to_num(scores, nan=0.0, posinf=1.0).sum(0), 1e-6, 1 - 1e-6)
    spec = O.Spec(num_classes=K)
    idx_r, sc_r, bx_r, w_r = O.get_pgt(scores, boxes, gt_int, img[None], 0, spec)
    assert idx_r.tolist() == [10, 123, 650] and torch.isnan(sc_r[1])
    idx, sc, bx, w = ops.oicr_pgt(scores.to(DEV), boxes.to(DEV), gt_int.to(DEV), img.to(DEV), False, None, 0, False,
                                  spec.bbox_reg_weights)
    assert torch.equal(idx.cpu(), idx_r)
    assert torch.equal(torch.isnan(sc.cpu()), torch.isnan(sc_r)) and torch.equal(sc.cpu()[[0, 2]], sc_r[[0, 2]])
    assert torch.equal(bx.cpu(), bx_r) and torch.equal(w.cpu(), w_r)
    # the fused MIL + first pseudo-GT kernel follows the same rule
    logits = torch.randn(R, 2 * K + 3 * (K + 1), generator=g)
    logits[77, 7] = float("nan")       # a NaN class logit makes that row's whole softmax NaN -> NaN scores in every column
    gt_oh = torch.zeros(K)
    gt_oh[gt_int] = 1
    loss = torch.zeros(1, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    s_dev, img_dev, pgt = ops.wsddn_mil_pgt(logits.to(DEV), K, 0, K, gt_oh.to(DEV), True, 1.0, loss, boxes.to(DEV),
                                            gt_int.to(DEV), counter)
    ref_scores = F.softmax(logits[:, :K], 1) * F.softmax(logits[:, K:2 * K], 0)
    ref_idx = torch.max(torch.index_select(ref_scores, 1, gt_int), dim=0)[1]
    assert ref_idx.tolist() == [77, 77, 77]
    assert torch.equal(pgt[0].cpu(), ref_idx)
