"""PCL head: the CPU restatement pinned to the golden of the UNMODIFIED reference PCL model (its own
compiled pcl_loss op, tests/golden/make_golden_pcl.py), the reference op itself where oracle/_ref holds it, and (-m gpu) the B200
head -- cluster mining on the host as in the reference, assignment / loss / gradient kernels on the device -- against the golden."""
import numpy as np
import pytest
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from oracle import pcl_oracle as P
from oracle import wsl_oracle as O

CASE = "oicr_r18_small"
DEV = "cuda:0"


def _setup(device="cpu", precision=None):
    name, _, ov, _ = helpers.CASES[CASE]
    ovr = list(ov) + ["MODEL.DEVICE", device, "MODEL.ROI_HEADS.NAME", "PCLROIHeads"]
    if precision:
        ovr += ["B200.PRECISION", precision]
    cfg = drn.builtin_config(name, ovr)
    return cfg


def _state(cfg):
    shapes = helpers.synth.state_shapes(drn.build_model(_setup())) if cfg.MODEL.DEVICE != "cpu" else helpers.synth.state_shapes(drn.build_model(cfg))
    return dict(helpers.case_weights(cfg, shapes))


def test_pcl_oracle_matches_reference_golden():
    g = helpers.load_golden("pcl_r18_small")
    cfg = _setup()
    state = _state(cfg)
    spec = O.spec_from_cfg(cfg)
    inputs = helpers.case_inputs(CASE)[:1]
    params = {k: v.clone().requires_grad_(True) for k, v in state.items() if k.startswith(O.TRAINABLE_PREFIXES)}
    losses, tr = P.forward_train(inputs, {**state, **params}, spec)
    for k, v in losses.items():
        assert helpers.rel_err(v.item(), g["loss/" + k]) < 1e-4, (k, v.item(), g["loss/" + k])
    for k, st in enumerate(tr["stages"]):
        for name in ("labels", "gt_assignment", "pc_labels", "pc_count"):
            np.testing.assert_array_equal(np.asarray(st[name]).astype(np.float32), g[f"stage{k}/{name}"], err_msg=f"stage {k} {name}")
        for name in ("cls_loss_weights", "pc_probs", "img_cls_loss_weights"):
            np.testing.assert_allclose(st[name], g[f"stage{k}/{name}"], rtol=1e-4, atol=1e-8, err_msg=f"stage {k} {name}")
        np.testing.assert_allclose(st["probs"].numpy(), g[f"stage{k}/probs"], rtol=1e-4, atol=1e-9)
    sum(losses.values()).backward()
    for k in (str(k) for k in g["trainable"]):
        if f"grad/{k}/none" in g:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
            continue
        helpers.check_grad(params[k].grad, g, f"grad/{k}/", 2e-4, "pcl")
    with torch.no_grad():
        ev = P.forward_eval_scores(inputs[0], state, spec)
    np.testing.assert_allclose(ev["all_scores"].numpy(), g["eval/all_scores"], rtol=1e-4, atol=1e-9)
    b, s, c, _ = O.inference_single_image(ev["all_boxes"], ev["all_scores"], (inputs[0]["height"], inputs[0]["width"]), spec)
    np.testing.assert_array_equal(c.numpy(), g["eval/det_classes"])
    np.testing.assert_allclose(s.numpy(), g["eval/det_scores"], rtol=1e-4)


def test_pcl_loss_restatement_matches_the_compiled_reference_op():
    """oracle._PCLLoss against the reference's own pcl_loss_cpu.cpp (oracle/_ref, built by oracle/build_ref.py) on random
    cluster assignments incl. an empty cluster, probabilities below the eps clamps and an absent image label."""
    from oracle.build_ref import load_pcl_ref

    ref = load_pcl_ref(build=True)
    if ref is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    g = torch.Generator().manual_seed(3)
    R, K, Pn = 300, 20, 6
    probs = torch.softmax(torch.randn(R, K + 1, generator=g) * 3, 1)
    probs[5, 0] = 1e-8
    m = {"labels": np.zeros(R, np.int32), "cls_loss_weights": torch.rand(R, generator=g).numpy(), "gt_assignment": -np.ones(R, np.int64),
         "pc_labels": np.array([3, 3, 8, 8, 12, 15], np.int32), "pc_probs": np.array([0.3, 1e-9, 0.5, 0.2, 0.9, np.nan], np.float32),
         "pc_count": np.zeros(Pn, np.int32), "img_cls_loss_weights": torch.rand(Pn, generator=g).numpy()}
    rows = torch.randperm(R, generator=g)[:120].numpy()
    m["pc_labels"][5] = 12  # the empty cluster (NaN mean, as np.average gives) carries a class that IS in the image: fmaxf drops the NaN
    for j, r in enumerate(rows):
        c = j % 5  # cluster 5 stays empty
        m["gt_assignment"][r], m["labels"][r] = c, m["pc_labels"][c]
        m["pc_count"][c] += 1
    real = np.zeros(K + 1, np.float32)
    real[[0, 3, 8, 12]] = 1  # class 15 has a cluster but no image label: ignored
    p = probs.clone().requires_grad_(True)
    loss = P._PCLLoss.apply(p, m, real)
    loss.backward()
    f32 = lambda a: torch.from_numpy(np.asarray(a, np.float32).reshape(1, -1).copy())
    out = torch.zeros(1, K + 1)
    ref.pcl_loss_forward(probs.clone(), f32(m["labels"]), f32(m["cls_loss_weights"]), f32(m["pc_labels"]), f32(m["pc_probs"]),
                         f32(m["img_cls_loss_weights"]), f32(real), out)
    assert helpers.rel_err(loss.item(), (out.sum() / R).item()) < 1e-5
    grad = torch.zeros(R, K + 1)
    ref.pcl_loss_backward(probs.clone(), f32(m["labels"]), f32(m["cls_loss_weights"]), f32(m["gt_assignment"]), f32(m["pc_labels"]),
                          f32(m["pc_probs"]), f32(m["pc_count"]), f32(m["img_cls_loss_weights"]), f32(real), torch.ones(1), grad)
    torch.testing.assert_close(p.grad, grad / R, rtol=1e-5, atol=1e-9)


def _gpu_model(precision="fp32"):
    cfg = _setup(device=DEV, precision=precision)
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    return cfg, model, weights


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp32_tc"])
def test_gpu_pcl_head_matches_reference_golden(precision):
    """PCLROIHeads on the B200 against the golden of the unmodified reference: losses (1e-3), every stage's mined clusters and
    proposal assignment (exact), cluster statistics (1e-4), the gradients of the reference's loss.backward() (fp32), eval scores."""
    g = helpers.load_golden("pcl_r18_small")
    cfg, model, _ = _gpu_model(precision)
    assert type(model.roi_heads).__name__ == "PCLROIHeads" and not model.roi_heads.train_capturable
    model.train()
    model.roi_heads.box_head.eval()
    inputs = helpers.case_inputs(CASE)[:1]
    batched = helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV)
    for rep in range(3):  # the PCL train step never becomes a captured plan (host mining between the stages)
        model.zero_grad(set_to_none=True)
        losses = model(batched)
        assert not any(k[0] == "train" for k in model._plans)
    assert list(losses) == ["loss_cls", "loss_cls_r0", "loss_cls_r1", "loss_cls_r2"]
    for k, v in losses.items():
        assert helpers.rel_err(v.item(), g["loss/" + k]) < 1e-3, (k, v.item(), g["loss/" + k])
    tr = model.roi_heads.last_trace[0]
    for k, st in enumerate(tr["stages"]):
        np.testing.assert_array_equal(st["labels"].cpu().numpy().astype(np.float32), g[f"stage{k}/labels"])
        np.testing.assert_array_equal(st["assignment"].cpu().numpy().astype(np.float32), g[f"stage{k}/gt_assignment"])
        np.testing.assert_array_equal(st["pc_count"].cpu().numpy(), g[f"stage{k}/pc_count"])
        np.testing.assert_array_equal(st["center_classes"].astype(np.float32), g[f"stage{k}/pc_labels"])
        np.testing.assert_allclose(st["weights"].cpu().numpy(), g[f"stage{k}/cls_loss_weights"], rtol=1e-3, atol=1e-8)
        np.testing.assert_allclose(st["pc_probs"].cpu().numpy(), g[f"stage{k}/pc_probs"], rtol=1e-3, atol=1e-8)
        np.testing.assert_allclose(st["img_w"].cpu().numpy(), g[f"stage{k}/img_cls_loss_weights"], rtol=1e-3, atol=1e-8)
    if precision == "fp32":
        sum(losses.values()).backward()
        named = dict(model.named_parameters())
        for k in (str(k) for k in g["trainable"]):
            if f"grad/{k}/none" in g:
                assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
                continue
            helpers.check_grad(named[k].grad, g, f"grad/{k}/", 1e-3, "pcl")
    model.eval()
    res, sc, bx = model.inference(helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV, train=False), do_postprocess=False)
    err = np.abs(sc[0][0].cpu().numpy().astype(np.float64) - g["eval/all_scores"]) / (np.abs(g["eval/all_scores"]) + 1e-3 * g["eval/all_scores"].max(0, keepdims=True) + 1e-30)
    assert err.max() < 1e-3
    ds, gs = res[0].scores.cpu().numpy(), g["eval/det_scores"]
    assert len(ds) == len(gs)
    np.testing.assert_allclose(ds, gs, rtol=1e-3)


@pytest.mark.gpu
def test_gpu_pcl_kernels_match_oracle_on_random_clusters():
    """drn_pcl_stage_fwd / bwd against the oracle's proposal_clusters + _PCLLoss on random logits and hand-placed centres
    (duplicate centres -> an empty cluster, proposals below both IoU thresholds)."""
    from drn_wsod_pytorch_b200 import ops

    R, K = 900, 20
    g = torch.Generator().manual_seed(11)
    inp = helpers.synth.make_inputs(400, 600, R, seed=5)
    boxes = inp["boxes"]
    logits = torch.randn(R, 2 * K + 3 * (K + 1), generator=g) * 2
    col = 2 * K + (K + 1)
    cb = torch.cat([boxes[[10, 200, 200, 555]], torch.tensor([[5000.0, 5000.0, 5100.0, 5100.0]])]).numpy()  # a duplicate and a far-away centre
    cc = np.array([4, 4, 9, 9, 13], np.int32)
    cs = np.array([0.9, 0.3, 0.5, 0.7, 0.2], np.float32)
    probs = torch.softmax(logits[:, col:col + K + 1], 1)
    m = P.proposal_clusters(boxes.numpy(), cb, cc, cs, np.clip(probs.numpy(), 1e-9, 1 - 1e-9))
    real = np.zeros(K + 1, np.float32)
    real[[0, 4, 9, 13]] = 1
    p = probs.clone().requires_grad_(True)
    with np.errstate(all="ignore"):
        ref_loss = P._PCLLoss.apply(p, m, real)
    ref_loss.backward()
    ref_dlog = probs * (p.grad - (p.grad * probs).sum(1, keepdim=True))
    loss = torch.zeros(1, device=DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = ops.pcl_stage(logits.to(DEV), col, K, boxes.to(DEV), torch.from_numpy(cb).to(DEV), torch.from_numpy(cc).to(DEV),
                       torch.from_numpy(cs).to(DEV), 1.0, loss, counter)
    assert counter.item() == 0
    np.testing.assert_array_equal(st["labels"].cpu().numpy(), m["labels"])
    np.testing.assert_array_equal(st["assignment"].cpu().numpy(), m["gt_assignment"])
    np.testing.assert_array_equal(st["pc_count"].cpu().numpy(), m["pc_count"].astype(np.float32))
    assert (m["pc_count"] == 0).any() and np.isnan(st["pc_probs"].cpu().numpy()[m["pc_count"] == 0]).all()  # empty cluster: NaN mean, as numpy
    ok = m["pc_count"] > 0
    np.testing.assert_allclose(st["pc_probs"].cpu().numpy()[ok], m["pc_probs"][ok], rtol=1e-5)
    np.testing.assert_allclose(st["img_w"].cpu().numpy(), m["img_cls_loss_weights"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(st["probs"].cpu(), probs, rtol=1e-5, atol=1e-9)
    assert helpers.rel_err(loss.item(), ref_loss.item()) < 1e-5
    dlog = torch.zeros(R, logits.shape[1], device=DEV)
    ops.pcl_stage_bwd(st, K, 1.0, torch.ones(1, device=DEV), col, dlog)
    torch.testing.assert_close(dlog[:, col:col + K + 1].cpu(), ref_dlog, rtol=1e-4, atol=1e-9)
    assert float(dlog[:, :col].abs().max()) == 0.0 and float(dlog[:, col + K + 1:].abs().max()) == 0.0
