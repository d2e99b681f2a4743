"""Test-time augmentation driver (SURVEY.md §8f row 3;
projects/WSL/wsl/modeling/test_time_augmentation_avg.py).

CPU: `oracle/tta_oracle.py` pinned (a) bit-exact against Pillow itself for the 8-bit bilinear resample, (b) against
golden vectors of the UNMODIFIED reference TTA classes (tests/golden/tta_*.npz, tests/golden/make_golden_tta.py);
host logic of drn_wsod_pytorch_b200/tta.py (coefficient tables, shapes, proposal transforms) against the oracle.
GPU (-m gpu): the CUDA resample / merge kernels and the whole driver through the C ABI against the oracle + goldens.
"""
import zlib

import numpy as np
import pytest
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from drn_wsod_pytorch_b200 import tta
from oracle import tta_oracle as T
from oracle import wsl_oracle as O

RESIZE_CASES = [(60, 100, 96, 160), (60, 100, 48, 80), (60, 100, 60, 77), (60, 100, 33, 100), (375, 500, 688, 917),
                (600, 1000, 480, 800), (37, 53, 111, 29), (100, 100, 100, 100), (7, 5, 300, 2), (333, 500, 1152, 1730)]


def _rand_u8(H, W, C, seed):
    return np.random.Generator(np.random.PCG64(seed)).integers(0, 256, (H, W, C), dtype=np.uint8)


# ------------------------------------------------------------------ oracle pins (CPU)
@pytest.mark.parametrize("H,W,nh,nw", RESIZE_CASES)
def test_resample_restatement_bit_exact_vs_pillow(H, W, nh, nw):
    Image = pytest.importorskip("PIL.Image")
    for C in (3, 1):
        img = _rand_u8(H, W, C, H * 7 + nw)
        arr = img if C == 3 else img[:, :, 0]
        ref = np.asarray(Image.fromarray(arr).resize((nw, nh), Image.BILINEAR))
        np.testing.assert_array_equal(T.pil_resize_bilinear_u8(arr, nw, nh), ref)


def _tta_setup(name):
    case, min_sizes, max_size, flip, dataset_hw = helpers.TTA_CASES[name]
    cfg = helpers.case_config(case)
    model = drn.build_model(cfg)
    state = helpers.case_weights(cfg, model)
    return cfg, state, helpers.tta_input(name), (min_sizes, max_size, flip, dataset_hw)


def _expand(b, cols):
    return np.tile(b, (1, cols // b.shape[1])) if b.shape[1] != cols else b


@pytest.mark.parametrize("name", list(helpers.TTA_CASES))
def test_tta_oracle_matches_reference_golden(name):
    g = helpers.load_golden(name)
    cfg, state, inp, (min_sizes, max_size, flip, dataset_hw) = _tta_setup(name)
    spec = O.spec_from_cfg(cfg)
    out = T.tta_forward(inp, state, spec, min_sizes, max_size, flip, cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST, dataset_hw)
    assert len(out["views"]) == int(g["n_views"])
    np.testing.assert_array_equal(out["views"][0]["image"].numpy(), g["view0/image"])
    for i, v in enumerate(out["views"]):
        im = np.ascontiguousarray(v["image"].numpy())
        assert tuple(im.shape) == tuple(g[f"view{i}/image_shape"])
        assert zlib.crc32(im.tobytes()) == int(g[f"view{i}/image_crc32"]), f"view {i}: resampled image differs from Pillow's"
        np.testing.assert_array_equal(v["boxes"].numpy(), g[f"view{i}/boxes"])  # float32 transform arithmetic, bit-exact
        np.testing.assert_array_equal(v["objectness"].numpy(), g[f"view{i}/objectness"])
        sc, bx = out["per_view"][i]
        assert helpers.rel_err(sc.numpy(), g[f"view{i}/all_scores"], floor=1e-9) < 1e-3
        np.testing.assert_allclose(bx.numpy(), _expand(g[f"view{i}/all_boxes"], bx.shape[1]), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(out["mean_boxes"].numpy(), _expand(g["mean_boxes"], out["mean_boxes"].shape[1]), rtol=1e-6, atol=1e-4)
    assert helpers.rel_err(out["mean_scores"].numpy(), g["mean_scores"], floor=1e-9) < 1e-3
    boxes, scores, classes, _ = out["det"]
    np.testing.assert_array_equal(classes.numpy(), g["det_classes"])
    np.testing.assert_allclose(boxes.numpy(), g["det_boxes"], rtol=1e-6, atol=1e-4)
    assert helpers.rel_err(scores.numpy(), g["det_scores"], floor=1e-9) < 1e-3


def test_tta_union_oracle_matches_reference_golden():
    """oracle.tta_forward_union against the UNMODIFIED reference GeneralizedRCNNWithTTAUNION (tests/golden/
    make_golden_tta_union.py): per-view detections, their union on the original image, the final detections."""
    name = "tta_r18_small"
    g = helpers.load_golden("tta_union_r18_small")
    cfg, state, inp, (min_sizes, max_size, flip, dataset_hw) = _tta_setup(name)
    spec = O.spec_from_cfg(cfg)
    out = T.tta_forward_union(inp, state, spec, min_sizes, max_size, flip, dataset_hw)
    assert len(out["per_view"]) == int(g["n_views"])
    for i, (b, s, c) in enumerate(out["per_view"]):
        np.testing.assert_array_equal(c.numpy(), g[f"view{i}/det_classes"])
        np.testing.assert_allclose(s.numpy(), g[f"view{i}/det_scores"], rtol=1e-4, atol=1e-9)
        np.testing.assert_allclose(b.numpy(), g[f"view{i}/det_boxes"], rtol=1e-6, atol=1e-4)
    ub, us, uc = out["union"]
    np.testing.assert_allclose(ub.numpy(), g["union_boxes"], rtol=1e-6, atol=1e-4)
    np.testing.assert_array_equal(uc.numpy(), g["union_classes"])
    db, ds, dc = out["det"]
    np.testing.assert_array_equal(dc.numpy(), g["det_classes"])
    np.testing.assert_allclose(ds.numpy(), g["det_scores"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(db.numpy(), g["det_boxes"], rtol=1e-6, atol=1e-4)


# ------------------------------------------------------------------ host logic of tta.py (CPU, no kernels)
@pytest.mark.parametrize("in_size,out_size", [(100, 160), (100, 48), (1000, 1920), (1000, 800), (5, 2), (7, 300), (600, 600)])
def test_coefficient_tables_match_oracle(in_size, out_size):
    b0, k0 = T.pil_bilinear_coeffs(in_size, out_size)
    b1, k1 = tta.resample_tables(in_size, out_size)
    np.testing.assert_array_equal(b0, b1)
    np.testing.assert_array_equal(k0, k1)
    assert int(k1.sum(1).min()) > 0 and abs(int(k1.sum(1).max()) - (1 << 22)) <= k1.shape[1]


def test_view_shapes_and_proposal_transforms_match_oracle():
    for (h, w, size, mx) in [(600, 1000, 480, 4000), (600, 1000, 1152, 4000), (160, 224, 224, 300), (500, 375, 688, 4000),
                             (333, 500, 1152, 1200)]:
        assert tta.ResizeShortestEdge(size, mx).get_shape(h, w) == T.resize_shortest_edge_shape(h, w, size, mx)
    inp = helpers.tta_input("tta_r18_small")
    H, W = inp["height"], inp["width"]
    boxes = torch.cat([inp["boxes"], torch.tensor([[5.0, 5.0, 5.0, 40.0], [-3.0, 2.0, 500.0, 400.0]])])  # empty + out of bounds
    obj = torch.cat([inp["objectness"], torch.tensor([0.5, 0.25])])
    for size, do_flip in [(128, False), (224, True)]:
        nh, nw = T.resize_shortest_edge_shape(H, W, size, 300)
        rec = [("resize", H, W, nh, nw)] + ([("hflip", nw)] if do_flip else [])
        tl = tta.TransformList([tta.ResizeTransform(H, W, nh, nw)] + ([tta.HFlipTransform(nw)] if do_flip else []))
        b0, o0 = T.transform_proposals(boxes, obj, (nh, nw), rec, 60)
        d = {"proposals": drn.Instances((H, W), proposal_boxes=drn.Boxes(boxes.clone()), objectness_logits=obj.clone())}
        tta.transform_proposals(d, (nh, nw), tl, proposal_topk=60)
        np.testing.assert_array_equal(d["proposals"].proposal_boxes.tensor.numpy(), b0.numpy())
        np.testing.assert_array_equal(d["proposals"].objectness_logits.numpy(), o0.numpy())
        assert d["proposals"].image_size == (nh, nw)
        back0 = T.apply_box(b0.numpy().copy(), T.inverse(rec))
        back1 = tl.inverse().apply_box(b0.numpy().copy())
        np.testing.assert_array_equal(back0, back1)
        assert tl.inverse().device_params() == tta.TransformList(tl.inverse().transforms).device_params()


def test_tta_wrapper_rejects_other_models_and_float_images():
    cfg = helpers.case_config("oicr_r18_small")
    with pytest.raises(AssertionError):
        tta.GeneralizedRCNNWithTTAAVG(cfg, torch.nn.Linear(2, 2))
    mapper = tta.DatasetMapperTTAAVG(cfg)
    assert mapper.proposal_topk == cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST
    with pytest.raises((RuntimeError, TypeError)):  # float images take F.interpolate in the reference; not on the B200 path
        mapper({"image": torch.rand(3, 64, 64), "height": 64, "width": 64})


# ------------------------------------------------------------------ GPU: kernels + driver through the C ABI
DEV = "cuda:0"
RTOL = 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,nh,nw", RESIZE_CASES + [(600, 1000, 1152, 1920)])  # last: the largest TTA view of the bench image
def test_gpu_resample_bit_exact(H, W, nh, nw):
    img = _rand_u8(H, W, 3, H + 3 * nw)
    want = T.pil_resize_bilinear_u8(img, nw, nh).transpose(2, 0, 1)  # oracle (itself bit-exact vs Pillow)
    src = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).to(DEV)
    for flip in (False, True):
        ref = want[:, :, ::-1] if flip else want
        got_u8 = tta.resize_u8(src, nh, nw, flip=flip, out_dtype=torch.uint8)
        got_f32 = tta.resize_u8(src, nh, nw, flip=flip, out_dtype=torch.float32)
        assert got_u8.dtype == torch.uint8 and got_f32.dtype == torch.float32
        np.testing.assert_array_equal(got_u8.cpu().numpy(), ref)
        np.testing.assert_array_equal(got_f32.cpu().numpy(), ref.astype(np.float32))
    # property: flipping twice is the identity; a same-size "resize" is a copy
    same = tta.resize_u8(tta.resize_u8(src, H, W, flip=True), H, W, flip=True)
    assert torch.equal(same, src)


@pytest.mark.gpu
def test_gpu_resample_rejects_bad_inputs():
    with pytest.raises(TypeError):
        tta.resize_u8(torch.zeros(3, 8, 8, device=DEV), 4, 4)
    with pytest.raises(RuntimeError):
        tta.resize_u8(torch.zeros(3, 8, 8, dtype=torch.uint8), 4, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("R,cols", [(1, 4), (97, 80), (4000, 80), (513, 4)])
def test_gpu_tta_accumulate_matches_oracle(R, cols):
    g = torch.Generator().manual_seed(R + cols)
    H, W = 375, 500
    views = []
    for size, do_flip, pre in [(480, False, False), (480, True, False), (688, True, True), (1152, False, True), (333, True, False)]:
        nh, nw = T.resize_shortest_edge_shape(H, W, size, 1200)
        rec = ([("resize", 750, 1000, H, W)] if pre else [("noop",)]) + [("resize", H, W, nh, nw)] + ([("hflip", nw)] if do_flip else [])
        tl = tta.TransformList(([tta.ResizeTransform(750, 1000, H, W)] if pre else [tta.NoOpTransform()])
                               + [tta.ResizeTransform(H, W, nh, nw)] + ([tta.HFlipTransform(nw)] if do_flip else []))
        xy = torch.rand(R, cols // 4, 2, generator=g) * torch.tensor([nw * 0.8, nh * 0.8])
        wh = torch.rand(R, cols // 4, 2, generator=g) * torch.tensor([nw * 0.2, nh * 0.2])
        boxes = torch.cat([xy, xy + wh], dim=2).reshape(R, cols).contiguous()
        if R > 1:
            boxes[1, :4] = torch.tensor([30.0, 20.0, 10.0, 5.0])  # corners out of order: apply_box re-sorts them
        scores = torch.softmax(torch.randn(R, 21, generator=g), 1)
        views.append((rec, tl, boxes, scores))
    # one view: the inverse transform alone, bit-exact (x / 1 is exact)
    for rec, tl, boxes, scores in views:
        ab, asc = torch.empty(R, cols, device=DEV), torch.empty(R, 21, device=DEV)
        drn.ops.tta_accumulate(boxes.to(DEV), scores.to(DEV), tl.inverse().device_params(), ab, asc, 0, 1)
        want = T.apply_box(boxes.reshape(-1, 4).numpy().copy(), T.inverse(rec)).reshape(R, cols)
        np.testing.assert_array_equal(ab.cpu().numpy(), want)
        assert torch.equal(asc.cpu(), scores)
    # all views: the mean (summed in view order; torch.mean may associate differently -> 1 ulp-level tolerance)
    ab, asc = torch.empty(R, cols, device=DEV), torch.empty(R, 21, device=DEV)
    for i, (rec, tl, boxes, scores) in enumerate(views):
        drn.ops.tta_accumulate(boxes.to(DEV), scores.to(DEV), tl.inverse().device_params(), ab, asc, i, len(views))
    mb, ms = T.tta_merge([v[2] for v in views], [v[3] for v in views], [v[0] for v in views])
    np.testing.assert_allclose(ab.cpu().numpy(), mb.numpy(), rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(asc.cpu().numpy(), ms.numpy(), rtol=1e-6, atol=1e-9)


def _build_gpu(name, precision="fp32", use_graph=True):
    case, min_sizes, max_size, flip, dataset_hw = helpers.TTA_CASES[name]
    cfg = helpers.case_config(case, device=DEV, precision=precision)
    cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, cfg.TEST.AUG.FLIP = list(min_sizes), max_size, flip
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.eval()
    model.use_cuda_graph = use_graph
    inp = helpers.tta_input(name)
    d = helpers.to_batched([inp], drn.Instances, drn.Boxes, device="cpu", train=False)[0]
    d["image"] = inp["image_u8"]
    if dataset_hw is not None:
        d["height"], d["width"] = dataset_hw
    return cfg, model, d


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(helpers.TTA_CASES))
def test_gpu_tta_driver_matches_reference_golden(name):
    g = helpers.load_golden(name)
    cfg, model, d = _build_gpu(name)
    # the mapper: resampled views bit-exact vs the reference's Pillow output, proposals bit-exact
    views = tta.DatasetMapperTTAAVG(cfg)(dict(d))
    assert len(views) == int(g["n_views"])
    np.testing.assert_array_equal(views[0]["image"].cpu().numpy(), g["view0/image"])
    for i, v in enumerate(views):
        im = np.ascontiguousarray(v["image"].cpu().numpy())
        assert im.dtype == np.uint8 and tuple(im.shape) == tuple(g[f"view{i}/image_shape"])
        assert zlib.crc32(im.tobytes()) == int(g[f"view{i}/image_crc32"]), f"view {i}"
        np.testing.assert_array_equal(v["proposals"].proposal_boxes.tensor.cpu().numpy(), g[f"view{i}/boxes"])
        np.testing.assert_array_equal(v["proposals"].objectness_logits.cpu().numpy(), g[f"view{i}/objectness"])
    # the driver, twice: first pass eager (signatures seen once), second pass through the captured plans
    wrapper = tta.GeneralizedRCNNWithTTAAVG(cfg, model)
    for rep in range(2):
        mean_boxes, mean_scores = wrapper.merged_views(dict(d))
        np.testing.assert_allclose(mean_boxes.cpu().numpy(), _expand(g["mean_boxes"], mean_boxes.shape[1]), rtol=1e-5, atol=2e-3)
        err = np.abs(mean_scores.cpu().numpy().astype(np.float64) - g["mean_scores"]) / (np.abs(g["mean_scores"]) + 1e-3 * g["mean_scores"].max(0, keepdims=True) + 1e-30)
        assert err.max() < RTOL, (rep, err.max())
        res = wrapper([dict(d)])[0]["instances"]
        ds, dc = res.scores.cpu().numpy(), res.pred_classes.cpu().numpy()
        gs, gc = g["det_scores"], g["det_classes"]
        assert len(ds) == len(gs)
        np.testing.assert_allclose(ds, gs, rtol=RTOL)
        gaps = np.abs(np.diff(gs)) / gs[:-1]
        if len(gs) > 1 and gaps.min() > 10 * RTOL:
            np.testing.assert_array_equal(dc, gc)
            np.testing.assert_allclose(res.pred_boxes.tensor.cpu().numpy(), g["det_boxes"], rtol=1e-5, atol=2e-3)


@pytest.mark.gpu
def test_gpu_tta_union_driver_matches_reference_golden():
    """GeneralizedRCNNWithTTAUNION on the B200 (views resampled on the device, per-view detections by the on-device tail,
    inverse transforms by drn_tta_accumulate, final tail at threshold 1e-8) against the reference's golden."""
    name = "tta_r18_small"
    g = helpers.load_golden("tta_union_r18_small")
    cfg, model, d = _build_gpu(name)
    views = tta.DatasetMapperTTAUNION(cfg)(dict(d))
    assert len(views) == int(g["n_views"])
    for i, v in enumerate(views):  # the UNION mapper hands the proposals over untouched
        assert tuple(v["image"].shape) == tuple(g[f"view{i}/image_shape"])
        np.testing.assert_array_equal(v["proposals"].proposal_boxes.tensor.cpu().numpy(), g[f"view{i}/boxes"])
    wrapper = tta.GeneralizedRCNNWithTTAUNION(cfg, model)
    ub, us, uc = wrapper.union_of_views(dict(d))
    assert ub.shape[0] == g["union_boxes"].shape[0]
    # same candidates per view (order inside a view follows the scores: compare as sorted sets per view)
    off = 0
    for i in range(int(g["n_views"])):
        n = len(g[f"view{i}/det_scores"])
        gs, gc = g["union_scores"][off:off + n], g["union_classes"][off:off + n]
        np.testing.assert_allclose(np.sort(us[off:off + n].cpu().numpy()), np.sort(gs), rtol=RTOL)
        assert sorted(uc[off:off + n].cpu().tolist()) == sorted(gc.tolist())
        off += n
    res = wrapper([dict(d)])[0]["instances"]
    ds, dc = res.scores.cpu().numpy(), res.pred_classes.cpu().numpy()
    gs, gc = g["det_scores"], g["det_classes"]
    assert len(ds) == len(gs)
    np.testing.assert_allclose(ds, gs, rtol=RTOL)
    gaps = np.abs(np.diff(gs)) / gs[:-1]
    if len(gs) > 1 and gaps.min() > 10 * RTOL:
        np.testing.assert_array_equal(dc, gc)
        np.testing.assert_allclose(res.pred_boxes.tensor.cpu().numpy(), g["det_boxes"], rtol=1e-5, atol=2e-3)


@pytest.mark.gpu
def test_gpu_tta_batched_views_and_skipped_detections_agree():
    """batch_size > 1 pads the views of one chunk onto a common canvas (ImageList.from_tensors) -- same scores as one
    view per call within fp32 noise; with_detections=False returns the same (all_scores, all_boxes)."""
    cfg, model, d = _build_gpu("tta_r18_small", use_graph=False)
    one = tta.GeneralizedRCNNWithTTAAVG(cfg, model, batch_size=1)
    two = tta.GeneralizedRCNNWithTTAAVG(cfg, model, batch_size=2)  # (normal, flipped) pairs share their size
    b1, s1 = one.merged_views(dict(d))
    b2, s2 = two.merged_views(dict(d))
    np.testing.assert_allclose(b1.cpu().numpy(), b2.cpu().numpy(), rtol=1e-6, atol=1e-4)
    np.testing.assert_allclose(s1.cpu().numpy(), s2.cpu().numpy(), rtol=1e-4, atol=1e-8)
    view = one.tta_mapper(dict(d))[0]
    view.pop("transforms")
    r_a, sc_a, bx_a = model.inference([view], do_postprocess=False)
    r_b, sc_b, bx_b = model.inference([view], do_postprocess=False, with_detections=False)
    assert r_b == [None] and len(r_a[0]) > 0
    assert torch.equal(sc_a[0], sc_b[0]) and torch.equal(bx_a[0], bx_b[0])


# ------------------------------------------------------------------ wrapper orchestration on the CPU (kernels emulated)
def _emulate_device_kernels(monkeypatch, state, spec):
    """Replace the three device entry points the wrapper reaches (resample, merge, detection tail) and the model's eval
    pipeline by CPU emulations built from the oracle, so that the HOST logic of tta.py -- view streaming, batching, the
    pairing of every view with its inverse transform parameters, the view count, the final call -- runs without a GPU."""
    from drn_wsod_pytorch_b200 import ops

    def resize_u8(img, nh, nw, flip=False, out_dtype=torch.uint8):
        out = T.pil_resize_bilinear_u8(np.ascontiguousarray(img.permute(1, 2, 0).numpy()), nw, nh)
        out = np.flip(out, axis=1) if flip else out
        return torch.from_numpy(np.ascontiguousarray(out.transpose(2, 0, 1))).to(out_dtype)

    def tta_accumulate(bx, sc, params, acc_b, acc_s, idx, n):
        b = bx.numpy().reshape(-1, 4).copy()
        for kind, a, c in params:  # the (kind, a, b) records drn_tta_accumulate receives
            x0, y0, x1, y1 = b[:, 0].copy(), b[:, 1].copy(), b[:, 2].copy(), b[:, 3].copy()
            if kind == ops.TTA_OP_RESIZE:
                x0, x1, y0, y1 = x0 * np.float32(a), x1 * np.float32(a), y0 * np.float32(c), y1 * np.float32(c)
            elif kind == ops.TTA_OP_HFLIP:
                x0, x1 = np.float32(a) - x0, np.float32(a) - x1
            b = np.stack([np.minimum(x0, x1), np.minimum(y0, y1), np.maximum(x0, x1), np.maximum(y0, y1)], axis=1)
        b = torch.from_numpy(b.reshape(bx.shape))
        if idx == 0:
            acc_b.copy_(b), acc_s.copy_(sc)
        else:
            acc_b.add_(b), acc_s.add_(sc)
        if idx == n - 1:
            acc_b.div_(n), acc_s.div_(n)

    def tail(all_boxes, all_scores, image_shape, score_thresh, nms_thresh, topk, inst_cls, box_cls, dets=None):
        boxes, scores, classes, rows = O.inference_single_image(all_boxes, all_scores, image_shape, spec)
        res = inst_cls(image_shape)
        res.pred_boxes, res.scores, res.pred_classes = box_cls(boxes), scores, classes
        return res, rows

    calls = []

    def inference(batched, detected=None, do_postprocess=True, with_detections=True):
        assert detected is None and not do_postprocess and not with_detections
        calls.append(len(batched))
        canvas = (max(b["image"].shape[1] for b in batched), max(b["image"].shape[2] for b in batched))
        scs, bxs = [], []
        for b in batched:
            assert "transforms" not in b  # popped before the model sees the view
            with torch.no_grad():
                t = O.forward_eval_scores({"image": b["image"].float(), "boxes": b["proposals"].proposal_boxes.tensor,
                                           "objectness": b["proposals"].objectness_logits}, state, spec, canvas)
            scs.append(t["all_scores"][None]), bxs.append(t["all_boxes"][None])
        return [None] * len(batched), scs, bxs

    monkeypatch.setattr(tta, "resize_u8", resize_u8)
    monkeypatch.setattr(ops, "tta_accumulate", tta_accumulate)
    monkeypatch.setattr(tta, "fast_rcnn_inference_single_image", tail)
    return inference, calls


@pytest.mark.parametrize("batch_size", [1, 2, 4])
def test_tta_wrapper_host_flow_matches_reference_golden(monkeypatch, batch_size):
    name = "tta_r18_small"
    g = helpers.load_golden(name)
    cfg, state, inp, (min_sizes, max_size, flip, dataset_hw) = _tta_setup(name)
    cfg.TEST.AUG.MIN_SIZES, cfg.TEST.AUG.MAX_SIZE, cfg.TEST.AUG.FLIP = list(min_sizes), max_size, flip
    model = drn.build_model(cfg)
    model.eval()
    spec = O.spec_from_cfg(cfg)
    inference, calls = _emulate_device_kernels(monkeypatch, state, spec)
    monkeypatch.setattr(model, "inference", inference)
    d = helpers.to_batched([inp], drn.Instances, drn.Boxes, device="cpu", train=False)[0]
    d["image"] = inp["image_u8"]
    wrapper = tta.GeneralizedRCNNWithTTAAVG(cfg, model, batch_size=batch_size)
    # streamed path (what __call__ uses) and the reference's list API give the same merge
    mean_boxes, mean_scores = wrapper.merged_views(dict(d))
    n_views = int(g["n_views"])
    assert calls == [batch_size] * (n_views // batch_size) + ([n_views % batch_size] if n_views % batch_size else [])
    np.testing.assert_allclose(mean_boxes.numpy(), _expand(g["mean_boxes"], mean_boxes.shape[1]), rtol=1e-6, atol=1e-4)
    if batch_size == 1:  # larger batches pad the views of a chunk to a common canvas, as the reference would with batch_size > 1
        assert helpers.rel_err(mean_scores.numpy(), g["mean_scores"], floor=1e-9) < 1e-3
    del calls[:]
    res = wrapper([dict(d)])[0]["instances"]
    assert sum(calls) == n_views
    assert res.image_size == (inp["height"], inp["width"])
    if batch_size == 1:
        np.testing.assert_array_equal(res.pred_classes.numpy(), g["det_classes"])
        np.testing.assert_allclose(res.pred_boxes.tensor.numpy(), g["det_boxes"], rtol=1e-6, atol=1e-4)
        assert helpers.rel_err(res.scores.numpy(), g["det_scores"], floor=1e-9) < 1e-3
    assert "transforms" not in d and d["image"] is inp["image_u8"]  # the caller's dict is left alone


def test_resample_restatement_random_sizes_vs_pillow():
    """60 seeded random size pairs (1..90 px, up- and down-scaling by up to ~20x, single rows / columns): the oracle's
    resampler and the product's coefficient tables against Pillow itself."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.Generator(np.random.PCG64(2024))
    for _ in range(60):
        H, W, nh, nw = (int(v) for v in rng.integers(1, 91, 4))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
        np.testing.assert_array_equal(T.pil_resize_bilinear_u8(img, nw, nh), ref, err_msg=f"{H}x{W} -> {nh}x{nw}")
        for a, b in ((W, nw), (H, nh)):
            b0, k0 = T.pil_bilinear_coeffs(a, b)
            b1, k1 = tta.resample_tables(a, b)
            np.testing.assert_array_equal(b0, b1)
            np.testing.assert_array_equal(k0, k1)
            assert (b1[:, 0] >= 0).all() and (b1[:, 0] + b1[:, 1] <= a).all() and (b1[:, 1] >= 1).all()
