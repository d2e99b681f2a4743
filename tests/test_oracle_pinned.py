"""Pins oracle/wsl_oracle.py (the CPU restatement used as the checker of the CUDA path):
  * against the golden vectors produced by the unmodified reference (tests/golden/*.npz);
  * against the reference's own known-answer tests for pairwise_iou / Matcher / Box2BoxTransform;
  * ROIPool restatement bit-exact against torchvision.ops.roi_pool (the un-vendored dependency
    the reference calls at detectron2/modeling/poolers.py:223-226).
CPU only."""
import numpy as np
import pytest
import torch

import helpers
from drn_wsod_pytorch_b200 import synth
from oracle import wsl_oracle as O

FAST_CASES = ["wsddn_v16_300", "oicr_r18_small", "oicr_v16_small", "oicr_r18_reg", "oicr_r18_batch2", "oicr_r50_small",
              "oicr_r101_coco_small"]


def _state(case):
    cfg = helpers.case_config(case)
    import drn_wsod_pytorch_b200 as drn

    model = drn.build_model(cfg)
    return cfg, helpers.case_weights(cfg, model)


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_matches_reference_golden(case):
    g = helpers.load_golden(case)
    cfg, state = _state(case)
    assert abs(synth.weights_checksum(state) - float(g["weights_checksum"])) <= 1e-6 * abs(float(g["weights_checksum"])), \
        "regenerated weights differ from the ones the golden vectors were made with"
    inputs = helpers.case_inputs(case)
    spec = O.spec_from_cfg(cfg)
    with torch.no_grad():
        losses, traces = O.forward_train(inputs, state, spec)
    for k, v in losses.items():
        assert helpers.rel_err(v.item(), g["loss/" + k]) < 2e-4, (k, v.item(), g["loss/" + k])
    assert {"loss/" + k for k in losses} == {k for k in g if k.startswith("loss/")}
    for i, t in enumerate(traces):
        assert helpers.rel_err(t["scores"].numpy(), g[f"img{i}/scores"], floor=1e-9) < 1e-3
        np.testing.assert_array_equal(t["labels_gt"].numpy(), g[f"img{i}/labels_gt"])
        for k, st in enumerate(t.get("stages", [])):
            p = f"img{i}/stage{k}/"
            np.testing.assert_array_equal(st["pgt_idx"].numpy(), g[p + "pgt_idx"])  # bit-exact argmax ROI indices
            np.testing.assert_array_equal(st["labels"].numpy(), g[p + "labels"])
            np.testing.assert_array_equal(st["matched"].numpy(), g[p + "matched"])
            np.testing.assert_allclose(st["pgt_boxes"].numpy(), g[p + "pgt_boxes"], rtol=0, atol=0)
            assert helpers.rel_err(st["probs"].numpy(), g[p + "probs"], floor=1e-9) < 1e-3
            assert helpers.rel_err(st["pgt_weights"].numpy(), g[p + "pgt_weights"]) < 1e-4
    # eval branch
    canvas = (max(i["height"] for i in inputs), max(i["width"] for i in inputs))
    for i, inp in enumerate(inputs):
        with torch.no_grad():
            t = O.forward_eval_scores(inp, state, spec, canvas)
            boxes, scores, classes, _ = O.inference_single_image(t["all_boxes"], t["all_scores"],
                                                                 (inp["height"], inp["width"]), spec)
        assert helpers.rel_err(t["all_scores"].numpy(), g[f"img{i}/eval/all_scores"], floor=1e-9) < 1e-3
        np.testing.assert_allclose(t["all_boxes"].numpy(), g[f"img{i}/eval/all_boxes"], rtol=1e-6, atol=1e-4)
        np.testing.assert_array_equal(classes.numpy(), g[f"img{i}/eval/det_classes"])
        np.testing.assert_allclose(scores.numpy(), g[f"img{i}/eval/det_scores"], rtol=1e-3)


def test_pairwise_iou_kat():
    # reference tests/structures/test_boxes.py:149-173
    b1 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    b2 = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.0, 0.0, 0.5, 1.0], [0.0, 0.0, 1.0, 0.5], [0.0, 0.0, 0.5, 0.5],
                       [0.5, 0.5, 1.0, 1.0], [0.5, 0.5, 1.5, 1.5]])
    exp = torch.tensor([[1.0, 0.5, 0.5, 0.25, 0.25, 0.25 / (2 - 0.25)]] * 2)
    assert torch.allclose(O.pairwise_iou(b1, b2), exp)


def test_matcher_kat():
    # reference tests/modeling/test_matcher.py:16-27 (RPN thresholds [0.3, 0.7], labels [0, -1, 1], low-quality on)
    q = torch.tensor([[0.15, 0.45, 0.2, 0.6], [0.3, 0.65, 0.05, 0.1], [0.05, 0.4, 0.25, 0.4]])
    m, l = O.matcher(q, [0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)
    assert m.tolist() == [1, 1, 2, 0]
    assert l.tolist() == [-1, 1, 0, 1]


def test_box2box_roundtrip_kat():
    # reference tests/modeling/test_box2box_transform.py:15-30
    g = torch.Generator().manual_seed(0)
    w = (5, 5, 10, 10)
    src = torch.rand(10, 4, generator=g) + torch.tensor([10.0, 10, 20, 20])
    dst = torch.rand(10, 4, generator=g) + torch.tensor([10.0, 10, 20, 20])
    d = O.get_deltas(src, dst, w)
    assert torch.allclose(dst, O.apply_deltas(d, src, w))


def test_roipool_restatement_bit_exact_vs_torchvision():
    rng = np.random.default_rng(3)
    C, h, w = 8, 19, 27
    feat = rng.standard_normal((C, h, w)).astype(np.float32)
    boxes = []
    for _ in range(200):
        x0, y0 = rng.uniform(-20, 200), rng.uniform(-20, 140)
        boxes.append([x0, y0, x0 + rng.uniform(0, 150), y0 + rng.uniform(0, 120)])
    # edge cases: exact .5 multiples after scaling, zero-size, fully outside, covering everything
    boxes += [[4, 4, 20, 20], [12, 12, 12, 12], [36, 28, 44, 36], [-100, -100, -50, -50], [500, 500, 600, 600],
              [0, 0, 216, 152], [4.0, 12.0, 4.0, 100.0], [100, 3.9999, 101, 4.0001]]
    boxes = np.asarray(boxes, dtype=np.float32)
    ours = O.roi_pool_numpy(feat, boxes, 1.0 / 8)
    ref = O.roi_pool(torch.from_numpy(feat)[None], torch.from_numpy(boxes), 1.0 / 8).numpy()
    assert np.array_equal(ours, ref)


_rand_dets = helpers.rand_dets


def test_nms_restatement_bit_exact_vs_torchvision():
    """oracle.nms_numpy restates torchvision.ops.nms (un-vendored): identical keep lists, incl. ties, zero-area
    boxes and thresholds that sit exactly on an fp32 IoU value."""
    import torchvision

    for seed in range(6):
        all_boxes, scores = _rand_dets(700, 1, seed, nreg_k=False, cluster=seed % 2 == 0)
        s = scores[:, 0].clone()
        if seed >= 4:  # ties and degenerate boxes
            s[::7] = s[3]
            all_boxes[::11, 2] = all_boxes[::11, 0]
        for thr in (0.3, 0.5, 0.7):
            ref = torchvision.ops.nms(all_boxes, s, thr).numpy()
            np.testing.assert_array_equal(O.nms_numpy(all_boxes.numpy(), s.numpy(), thr), ref)
    # threshold exactly equal to an attained fp32 IoU: the double comparison keeps/suppresses like torchvision
    b = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 3.0]])
    s = torch.tensor([0.9, 0.8])
    for thr in (0.3, float(np.float32(0.3)), 0.30000001):
        np.testing.assert_array_equal(O.nms_numpy(b.numpy(), s.numpy(), thr), torchvision.ops.nms(b, s, thr).numpy())


@pytest.mark.parametrize("R,K,nreg_k", [(1500, 20, True), (300, 5, False), (1200, 20, False)])
def test_exact_per_class_inference_tail_matches_reference_function(R, K, nreg_k):
    """oracle.inference_single_image_exact (per-class NMS, what the CUDA tail implements) against the restatement
    that calls torchvision.ops.batched_nms like the reference does (fast_rcnn.py:126 -> layers/nms.py:20)."""
    spec = O.Spec(num_classes=K)
    for seed in range(3):
        all_boxes, scores = _rand_dets(R, K, 100 + seed, nreg_k)
        if seed == 2:
            scores[5, 1] = float("nan")
            all_boxes[9, 0] = float("inf")
        a = O.inference_single_image(all_boxes, scores, (600, 1000), spec)
        b = O.inference_single_image_exact(all_boxes, scores, (600, 1000), spec)
        assert len(a[1]) == spec.detections_per_image or len(a[1]) == len(b[1])
        for x, y in zip(a[:3], b[:3]):
            assert torch.equal(x, y)


@pytest.mark.parametrize("case", ["wsddn_v16_300", "oicr_r18_reg", "oicr_r18_batch2"])
def test_oracle_backward_matches_reference_gradients(case):
    """oracle.backward_reference (autograd through the restated forward, with the reference's detach points) against
    gradient fingerprints of the unmodified reference's loss.backward() (tests/golden/make_golden_grads.py)."""
    g = helpers.load_golden(case + "_grads")
    cfg = helpers.case_config(case)
    model = __import__("drn_wsod_pytorch_b200").build_model(cfg)
    state = dict(helpers.case_weights(cfg, model))
    spec = O.spec_from_cfg(cfg)
    losses, grads = O.backward_reference(helpers.case_inputs(case), state, spec)
    for k, v in losses.items():
        assert helpers.rel_err(v.item(), g["loss/" + k]) < 1e-4
    trainable = [str(k) for k in g["trainable"]]
    assert sorted(trainable) == sorted(grads)
    for k in trainable:
        if f"grad/{k}/none" in g:
            assert grads[k] is None or float(grads[k].abs().max()) == 0.0, k  # unused parameter (bbox_pred without REFINE_REG)
            continue
        helpers.check_grad(grads[k], g, f"grad/{k}/", 2e-4, case)
