"""Input formats either side of the hot path, CPU only: proposal files, transform_proposals and
weight files, each against the UNMODIFIED reference function imported under oracle/refstub.py when /root/reference
is present (the comparisons are skipped on a box without it; the property checks always run)."""
import os
import pickle

import numpy as np
import pytest
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from drn_wsod_pytorch_b200 import data as D
from oracle import refstub

HAVE_REF = refstub.reference_available()


def _ref():
    refstub.install()
    if not hasattr(np, "int"):
        np.int = int  # the reference's Boxes.unique_boxes spells the cast np.int (removed in NumPy 1.24)
    from detectron2.data import detection_utils as du
    from detectron2.data.build import load_proposals_into_dataset
    return du, load_proposals_into_dataset


def _raw_proposals(n, seed, W=500, H=375):
    g = np.random.RandomState(seed)
    x0, y0 = g.randint(-5, W - 10, n), g.randint(-5, H - 10, n)
    w, h = g.randint(1, 200, n), g.randint(1, 200, n)
    boxes = np.stack([x0, y0, x0 + w, y0 + h], 1).astype(np.int16)
    boxes[n // 3] = boxes[0]           # exact duplicates
    boxes[n // 2] = boxes[1]
    boxes[5] = [10, 10, 12, 300]       # thinner than MIN_SIZE
    boxes[6] = [W + 20, 5, W + 80, 60]  # outside the image: clips to zero width
    scores = g.rand(n).astype(np.float32)
    return boxes, scores


def test_transform_proposals_properties_and_reference():
    H, W = 375, 500
    for seed, mode in ((0, D.XYXY_ABS), (1, D.XYWH_ABS), (2, D.XYXY_ABS)):
        boxes, scores = _raw_proposals(300, seed)
        if mode == D.XYWH_ABS:
            boxes = np.stack([boxes[:, 0], boxes[:, 1], boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]], 1)
        order = scores.argsort()[::-1]
        tfm = D.TransformList([D.ResizeTransform(H, W, 600, 800), D.HFlipTransform(800)]) if seed == 2 else None
        shape = (600, 800) if seed == 2 else (H, W)
        rec = {"proposal_boxes": boxes[order].copy(), "proposal_objectness_logits": scores[order].copy(), "proposal_bbox_mode": mode}
        ours = dict(rec)
        D.transform_proposals(ours, shape, tfm, proposal_topk=120, min_box_size=20)
        p = ours["proposals"]
        assert "proposal_boxes" not in ours and len(p) <= 120
        b = p.proposal_boxes.tensor
        assert (b[:, 0] >= 0).all() and (b[:, 2] <= shape[1]).all() and (b[:, 3] <= shape[0]).all()
        assert ((b[:, 2] - b[:, 0]) > 20).all() and ((b[:, 3] - b[:, 1]) > 20).all()
        assert len({tuple(np.round(r).tolist()) for r in b.numpy()}) == len(b)           # no duplicates survive
        assert (p.objectness_logits[:-1] >= p.objectness_logits[1:]).all()               # score order preserved
        if HAVE_REF:
            du, _ = _ref()
            from detectron2.structures import BoxMode

            class T:  # the reference only calls transforms.apply_box
                def apply_box(self, x):
                    return x if tfm is None else tfm.apply_box(x)

            theirs = {"proposal_boxes": rec["proposal_boxes"].copy(), "proposal_objectness_logits": rec["proposal_objectness_logits"].copy(),
                      "proposal_bbox_mode": BoxMode(mode)}
            du.transform_proposals(theirs, shape, T(), proposal_topk=120, min_box_size=20)
            q = theirs["proposals"]
            assert torch.equal(q.proposal_boxes.tensor, b) and torch.equal(q.objectness_logits, p.objectness_logits)


def test_proposal_file_roundtrip(tmp_path):
    ids = ["000012", "000017", "000023"]
    raw = [_raw_proposals(50 + 10 * i, 10 + i) for i in range(3)]
    # MCG-style arrays: 1-based (y1, x1, y2, x2)
    mcg_boxes = [b[:, (1, 0, 3, 2)].astype(np.int32) + 1 for b, _ in raw]
    conv = D.convert_proposals(ids, mcg_boxes, [s[:, None] for _, s in raw])
    for (b, s), cb, cs in zip(raw, conv["boxes"], conv["scores"]):
        assert cb.dtype == np.int16 and np.array_equal(cb, b) and np.array_equal(cs, s)
    ss = D.convert_proposals(ids, mcg_boxes)
    assert all((s == 1.0).all() and s.dtype == np.float32 for s in ss["scores"])
    path = str(tmp_path / "mcg_proposals.pkl")
    with open(path, "wb") as f:
        pickle.dump(conv, f, pickle.HIGHEST_PROTOCOL)
    records = [{"image_id": "000023"}, {"image_id": "000012"}]
    ours = D.load_proposals_into_dataset([dict(r) for r in records], path)
    for r in ours:
        i = ids.index(r["image_id"])
        order = raw[i][1].argsort()[::-1]
        assert np.array_equal(r["proposal_boxes"], raw[i][0][order]) and np.array_equal(r["proposal_objectness_logits"], raw[i][1][order])
        assert r["proposal_bbox_mode"] == D.XYXY_ABS
    if HAVE_REF:
        _, ref_load = _ref()
        theirs = ref_load([dict(r) for r in records], path)
        for a, b in zip(ours, theirs):
            assert np.array_equal(a["proposal_boxes"], b["proposal_boxes"])
            assert np.array_equal(a["proposal_objectness_logits"], b["proposal_objectness_logits"])
            assert int(a["proposal_bbox_mode"]) == int(b["proposal_bbox_mode"].value)


def test_load_checkpoint_pth_pkl_and_suffix_matching(tmp_path):
    cfg = helpers.case_config("oicr_r18_small")
    src = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, src)
    # (a) native .pth with a "model" entry and DataParallel prefixes
    pth = str(tmp_path / "model_final.pth")
    torch.save({"model": {"module." + k: v for k, v in weights.items()}, "iteration": 7}, pth)
    m1 = drn.build_model(cfg)
    missing, unexpected = D.load_checkpoint(m1, pth)
    assert missing == [] and unexpected == []
    for k, v in m1.state_dict().items():
        if k in weights:
            assert torch.equal(v, weights[k]), k
    # (b) Detectron2-format .pkl of a backbone only (numpy arrays, no prefix) -> suffix matching, as resnet18_ws_model_120_d2.pkl
    bb = {k[len("backbone."):]: v.numpy() for k, v in weights.items() if k.startswith("backbone.")}
    pkl = str(tmp_path / "resnet18_ws_d2.pkl")
    with open(pkl, "wb") as f:
        pickle.dump({"model": bb, "__author__": "test", "matching_heuristics": True}, f)
    m2 = drn.build_model(cfg)
    before = {k: v.clone() for k, v in m2.state_dict().items()}
    missing, unexpected = D.load_checkpoint(m2, pkl)
    assert missing == [] and unexpected == []
    for k, v in m2.state_dict().items():
        if k.startswith("backbone."):
            assert torch.equal(v, weights[k]), k
        else:
            assert torch.equal(v, before[k]), k  # heads untouched
    # (c) a Caffe2 pickle is refused with a pointer to the reference's converter
    c2 = str(tmp_path / "R-50.pkl")
    with open(c2, "wb") as f:
        pickle.dump({"blobs": {"conv1_w": np.zeros((1,), np.float32)}}, f)
    with pytest.raises(NotImplementedError):
        D.load_checkpoint(m2, c2)
    if HAVE_REF:
        refstub.install()
        from detectron2.checkpoint.c2_model_loading import align_and_update_state_dicts as ref_align

        ms_ref = {k: v.clone() for k, v in before.items()}
        ms_ours = {k: v.clone() for k, v in before.items()}
        ck = {k: torch.from_numpy(v) for k, v in bb.items()}
        ref_align(ms_ref, {k: v.clone() for k, v in ck.items()}, c2_conversion=False)
        D.align_and_update_state_dicts(ms_ours, ck)
        assert all(torch.equal(ms_ref[k], ms_ours[k]) for k in ms_ref)
