"""Parity of the BENCHMARKED precision (-m gpu): the bf16 tensor-core mode against the oracle run with bf16-rounded
layer operands and fp32 accumulation (`oracle.wsl_oracle.bf16_operands`, the same restatement that is pinned to the
reference's goldens in fp32), at the small golden cases AND at the full size of BASELINE.json configs[2] and [4];
plus full-size fp32-accurate (`fp32_tc`) runs of configs[3] and [4] against the fp32 oracle.

Two kinds of check, because of what bf16 storage does to a deep net:

* LAYER BY LAYER ("teacher forced"), tight: every dense layer of the B200 path -- first conv, each 3x3 / 1x1 conv with
  its folded FrozenBN, shortcut add and ReLU, ROIPool x (objectness+1), fc6, fc7, the concatenated heads -- is recomputed
  by the oracle's arithmetic FROM THE GPU'S OWN INPUT of that layer (torch fp32 CUDA kernels, TF32 off, bf16-rounded
  operands).  The two sides then differ only by fp32 summation order (and tcgen05's truncating accumulator, DESIGN.md
  3.7), which can move a stored bf16 value by at most ONE bf16 ulp where the fp32 sum sits on a rounding boundary:
  tolerance |diff| <= 2^-7 |ref| (+ a floor of 1e-3 of the tensor's rms for values near zero), and only a small stated
  fraction of the elements may differ at all.  ROIPool must be bit-exact (max-pool has no rounding; the objectness
  multiply is one fp32 product rounded once).
* END TO END, loose by necessity: one flipped ulp changes ~1/sqrt(K) ulp of every output in its receptive field, which
  flips ~2 % of THOSE roundings -- the flips multiply layer by layer until every activation carries independent rounding
  noise of ~2^-9 per layer.  Two correct bf16 implementations with different summation orders therefore differ end to end
  as much as either differs from fp32 (measured here: up to ~12 % on individual proposal scores through R18, 0.1-5 % on
  the losses, 28 % on a refinement-stage loss of 4e-3 through R101).  The chain check keeps the tolerances of the
  fp32-vs-bf16 comparison -- MIL loss 6e-2, the proposal scores that matter 40 % (measured: 12 % through R18, 29 % through R101 with 80 classes), pseudo-GT argmax equal wherever the
  oracle's top-2 margin exceeds 12 %, refinement-stage losses 35 % and only while both sides mined the same pseudo GT."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers
import drn_wsod_pytorch_b200 as drn
from drn_wsod_pytorch_b200 import modeling as M, ops
from oracle import wsl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ULP = 2.0 ** -7       # one bf16 ulp is 2^-8 .. 2^-7 of the value
E2E_LOSS, E2E_STAGE_LOSS, E2E_SCORE, E2E_MARGIN = 6e-2, 3.5e-1, 0.40, 0.12


def _build(cfg_name, precision, extra=()):
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", DEV, "B200.PRECISION", precision] + list(extra))
    return _finish(cfg)


def _finish(cfg):
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.roi_heads.keep_trace = True
    model.train()
    model.roi_heads.box_head.eval()  # dropout off for the comparison
    model.use_cuda_graph = False     # eager launches: the per-layer hooks below see every call
    return cfg, model, weights


def _one_ulp(ours, ref, what, max_changed):
    """`ours` (bf16) equals `ref` (the oracle's fp32 result rounded to bf16) up to rounding-boundary flips."""
    a, b = ours.float(), ref.to(torch.bfloat16).float()
    diff = (a - b).abs()
    floor = 1e-3 * float(b.pow(2).mean().sqrt())
    worst = float((diff - ULP * b.abs()).max())
    assert worst <= floor, f"{what}: off by more than one bf16 ulp (excess {worst:.3e}, floor {floor:.3e})"
    changed = float((diff > 0).float().mean())
    assert changed <= max_changed, f"{what}: {changed:.4f} of the elements differ (allowed {max_changed})"
    return changed


def _score_err(a, b, keep=5e-2):
    """relative error over the entries that matter (>= `keep` of their column's max)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    big = b >= keep * b.max(axis=0, keepdims=True)
    return float(np.max(np.abs(a[big] - b[big]) / b[big]))


def _margins(prev_scores, gt_int):
    sel = torch.index_select(prev_scores, 1, gt_int)
    top2 = torch.topk(sel, 2, dim=0)[0]
    return (top2[0] - top2[1]) / top2[0].abs().clamp(min=1e-30)


def _run_recorded(model, batched, monkeypatch):
    """One eager forward with every backbone layer call recorded: (kind, module / args, input, residual, relu, output)."""
    rec = []
    orig_conv, orig_first, orig_pool = M.run_conv, ops.first_conv, ops.maxpool2x2

    def run_conv(conv, x, precision, relu, residual=None):
        y = orig_conv(conv, x, precision, relu, residual)
        rec.append(("conv", conv, x, residual, relu, y))
        return y

    def first_conv(img, canvas, mean, std, packed, stride, *a, **k):
        y = orig_first(img, canvas, mean, std, packed, stride, *a, **k)
        rec.append(("first", (canvas, mean, std, stride), img, None, True, y))
        return y

    def maxpool(x, stride):
        y = orig_pool(x, stride)
        rec.append(("pool", stride, x, None, False, y))
        return y

    monkeypatch.setattr(M, "run_conv", run_conv)
    monkeypatch.setattr(ops, "first_conv", first_conv)
    monkeypatch.setattr(ops, "maxpool2x2", maxpool)
    losses = model(batched)
    monkeypatch.undo()
    return losses, rec


def _check_backbone_layers(model, rec):
    """Every recorded layer against the oracle's bf16-operand arithmetic applied to the GPU's own input of that layer."""
    first = model.backbone.stem.conv1 if hasattr(model.backbone, "stem") else model.backbone.plain1[0].conv1
    stats = {"layers": 0, "max_changed": 0.0}
    with torch.no_grad(), O.bf16_operands():
        for kind, what, x, res, relu, y in rec:
            if kind == "pool":  # exact
                ref = F.max_pool2d(x.permute(0, 3, 1, 2).float(), 2, what).permute(0, 2, 3, 1)
                assert torch.equal(y.float(), ref), "max-pool must be exact"
                continue
            if kind == "first":  # fp32 conv of the normalised image (fp32 weights, FrozenBN affine), output rounded to bf16
                canvas, mean, std, stride = what
                conv = first
                img = (x.float() - torch.tensor(mean, device=x.device).view(3, 1, 1)) / torch.tensor(std, device=x.device).view(3, 1, 1)
                img = F.pad(img, (0, canvas[1] - img.shape[2], 0, canvas[0] - img.shape[1]))[None]
                t = F.conv2d(img, conv.weight.float(), None if conv.bias is None else conv.bias.float(), stride=stride, padding=1)
                if conv.norm is not None:
                    n = conv.norm
                    t = F.batch_norm(t, n.running_mean, n.running_var, n.weight, n.bias, training=False, eps=n.eps)
                ref = F.relu(t).permute(0, 2, 3, 1)
                name = "first conv"
            else:
                conv = what
                w = conv.weight.float()
                if conv.norm is not None:
                    n = conv.norm
                    scale = n.weight * (n.running_var + n.eps).rsqrt()
                    bias = n.bias - n.running_mean * scale
                    w = w * scale.view(-1, 1, 1, 1)
                else:
                    bias = conv.bias.float()
                wq = w.to(torch.bfloat16).float()
                k, d = conv.kernel_size, conv.dilation
                t = F.conv2d(x.permute(0, 3, 1, 2).float(), wq, None, padding=d * (k // 2), dilation=d) + bias.view(1, -1, 1, 1)
                if res is not None:
                    t = t + res.permute(0, 3, 1, 2).float()
                ref = (F.relu(t) if relu else t).permute(0, 2, 3, 1)
                name = f"conv {tuple(conv.weight.shape)} dil {d} @ {tuple(x.shape[1:3])}"
            assert y.shape == ref.shape, (name, y.shape, ref.shape)
            changed = _one_ulp(y, ref, name, max_changed=0.02)
            stats["layers"] += 1
            stats["max_changed"] = max(stats["max_changed"], changed)
    return stats


def _check_roi_stage(model, tr, feat_map, boxes, obj, stride, rows):
    """ROIPool bit-exact from the GPU's own map; fc6 / fc7 / head logits on `rows` from the GPU's own layer inputs."""
    import torchvision

    rh = model.roi_heads
    R, C = boxes.shape[0], feat_map.shape[1]
    with torch.no_grad(), O.bf16_operands():
        pooled_ref = torchvision.ops.roi_pool(feat_map.float().contiguous(), torch.cat([boxes.new_zeros(R, 1), boxes], 1), (7, 7), 1.0 / stride)
        pooled_ref = (pooled_ref * (obj + 1).view(-1, 1, 1, 1)).to(torch.bfloat16)  # [R, C, 7, 7]
        assert torch.equal(tr["acts"][0].view(R, 49, C), pooled_ref.permute(0, 2, 3, 1).reshape(R, 49, C)), "ROIPool must be bit-exact"
        fc1, fc2 = rh.box_head.fc1, rh.box_head.fc2
        x = pooled_ref[rows].float().flatten(1)  # the reference's (c, ph, pw) flattening against the parameter's own layout
        y6 = F.relu(F.linear(x, fc1.weight.to(torch.bfloat16).float(), fc1.bias.float()))
        _one_ulp(tr["acts"][1][rows], y6, "fc6", max_changed=0.15)  # K up to 100 352: the truncating accumulator moves more sums across a boundary
        y7 = F.relu(F.linear(tr["acts"][1][rows].float(), fc2.weight.to(torch.bfloat16).float(), fc2.bias.float()))
        _one_ulp(tr["acts"][2][rows], y7, "fc7", max_changed=0.05)
        heads = rh._heads_packed()
        ref_logits = tr["feat"][rows].float() @ heads["w"].float().t() + heads["bias"]
        torch.testing.assert_close(tr["logits"][rows], ref_logits, rtol=1e-3, atol=1e-3)  # fp32 out: summation order only


def _check_chain(tr, ref_tr, losses, ref_losses, gt_int):
    assert helpers.rel_err(losses["loss_cls"].item(), ref_losses["loss_cls"].item()) < E2E_LOSS, (losses["loss_cls"].item(), ref_losses["loss_cls"].item())
    assert _score_err(tr["scores"].cpu().numpy(), ref_tr["scores"].cpu().numpy()) < E2E_SCORE
    if "stages" not in ref_tr:
        return
    prev = ref_tr["scores"].cpu()
    same_pgt = True
    for k, st in enumerate(ref_tr["stages"]):
        sure = (_margins(prev, gt_int) > E2E_MARGIN).numpy()
        got, want = tr["stages"][k]["pgt_idx"].cpu().numpy(), st["pgt_idx"].cpu().numpy()
        assert np.array_equal(got[sure], want[sure]), (k, got, want)
        same_pgt = same_pgt and np.array_equal(got, want)
        # a refinement-stage loss is a weighted CE over the proposals the pseudo GT labels: comparable only while both
        # sides mined the same pseudo GT, and then still a sum of ~R small terms each carrying the logits' bf16 noise
        if same_pgt and sure.all():
            key = f"loss_cls_r{k}"
            assert helpers.rel_err(losses[key].item(), ref_losses[key].item()) < E2E_STAGE_LOSS, (key, losses[key].item(), ref_losses[key].item())
        prev = st["probs"].cpu()


SMALL = ["oicr_r18_small", "oicr_r50_small", "oicr_v16_small", "oicr_r101_coco_small", "wsddn_v16_300"]
FULL = {
    # BASELINE.json configs[2]: the benchmarked configuration
    "config2_r50": ("oicr_WSR_50_DC5_1x", 600, 1000, 4000, 20),
    # BASELINE.json configs[4]: R101-WS, COCO shape, 80 classes
    "config4_r101_coco": ("oicr_WSR_101_DC5_1x_coco", 800, 1333, 4000, 80),
}


@pytest.mark.parametrize("name", SMALL + list(FULL))
def test_bf16_mode_layer_by_layer_and_end_to_end(name, monkeypatch):
    if name in FULL:
        cfg_name, H, W, R, K = FULL[name]
        cfg, model, weights = _build(cfg_name, "bf16")
        inputs = [helpers.synth.make_inputs(H, W, R, seed=0, num_gt=2, num_classes=K)]
    else:
        cfg, model, weights = _finish(helpers.case_config(name, device=DEV, precision="bf16"))
        inputs = helpers.case_inputs(name)[:1]
    spec = O.spec_from_cfg(cfg)
    inp = inputs[0]
    R = inp["boxes"].shape[0]
    losses, rec = _run_recorded(model, helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV), monkeypatch)
    tr = model.roi_heads.last_trace[0]
    # ---- layer by layer, from the GPU's own layer inputs
    stats = _check_backbone_layers(model, rec)
    nconv = sum(1 for m in model.backbone.modules() if isinstance(m, M.Conv2d))
    assert stats["layers"] == nconv, (stats, nconv)  # every conv of the backbone was seen and checked
    fmap = rec[-1][5].permute(0, 3, 1, 2)  # the map the ROI stage pooled from
    rows = torch.arange(0, R, max(1, R // 66), device=DEV)[:66]
    _check_roi_stage(model, tr, fmap, inp["boxes"].to(DEV), inp["objectness"].to(DEV), spec.stride, rows)
    del rec
    # ---- end to end: the oracle's independent bf16-operand chain (torch fp32 kernels on the GPU for the full sizes)
    dev = DEV if name in FULL else "cpu"
    state = {k: v.to(dev) for k, v in weights.items()}
    dinp = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    with torch.no_grad(), torch.device(dev), O.bf16_operands():
        ref_losses, ref_tr = O.forward_train([dinp], state, spec)
    _check_chain(tr, ref_tr[0], losses, ref_losses, torch.unique(inp["gt_classes"], sorted=True))


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
