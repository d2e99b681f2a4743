"""Generate tests/golden/tta_*.npz by running the UNMODIFIED reference TTA driver
(projects/WSL/wsl/modeling/test_time_augmentation_avg.py: DatasetMapperTTAAVG + GeneralizedRCNNWithTTAAVG, imported
under oracle/refstub.py) on CPU with the seeded inputs / weights of tests/helpers.py.  Only runnable where
/root/reference exists.

    python tests/golden/make_golden_tta.py

Per case: every augmented view's image (shape + CRC-32 of the uint8 CHW bytes; view 0 in full), transformed
proposals and per-view all_scores / all_boxes, the averaged scores / boxes in original-image coordinates (boxes
stored as R x 4 when the K class copies are identical), and the final detections.
"""
import os
import sys
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import refstub  # noqa: E402


def compact_boxes(b):
    """R x 4K boxes whose K class copies are identical (no box regression in the case) are stored as R x 4."""
    first = b[:, :4]
    return first.copy() if np.array_equal(np.tile(first, (1, b.shape[1] // 4)), b) else b


def run_case(name):
    case, min_sizes, max_size, flip, dataset_hw = helpers.TTA_CASES[name]
    _, yaml_rel, ov, _ = helpers.CASES[case]
    ov = [str(x) if not isinstance(x, str) else x for x in ov] + [
        "TEST.AUG.MIN_SIZES", str(list(min_sizes)), "TEST.AUG.MAX_SIZE", str(max_size), "TEST.AUG.FLIP", str(flip)]
    cfg, model = refstub.build_reference_model(yaml_rel, ov)
    from detectron2.structures import Boxes, Instances
    from wsl.modeling import GeneralizedRCNNWithTTAAVG

    ours_cfg = helpers.case_config(case)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = dict(helpers.case_weights(ours_cfg, shapes))
    sd["pixel_mean"] = model.state_dict()["pixel_mean"]
    sd["pixel_std"] = model.state_dict()["pixel_std"]
    model.load_state_dict(sd, strict=True)
    model.eval()
    inp = helpers.tta_input(name)
    d = helpers.to_batched([inp], Instances, Boxes, train=False)[0]
    d["image"] = inp["image_u8"]
    if dataset_hw is not None:
        d["height"], d["width"] = dataset_hw
    tta = GeneralizedRCNNWithTTAAVG(cfg, model)
    out = {}
    with torch.no_grad():
        views, tfms = tta._get_augmented_inputs(dict(d))
        out["n_views"] = np.int64(len(views))
        for i, v in enumerate(views):
            im = np.ascontiguousarray(v["image"].numpy())
            out[f"view{i}/image_shape"] = np.array(im.shape, dtype=np.int64)
            out[f"view{i}/image_crc32"] = np.int64(zlib.crc32(im.tobytes()))
            if i == 0:
                out["view0/image"] = im
            out[f"view{i}/boxes"] = v["proposals"].proposal_boxes.tensor.numpy()
            out[f"view{i}/objectness"] = v["proposals"].objectness_logits.numpy()
        _, all_scores, all_boxes = tta._batch_inference(views)
        for i in range(len(views)):
            out[f"view{i}/all_scores"] = all_scores[i][0].numpy()
            out[f"view{i}/all_boxes"] = compact_boxes(all_boxes[i][0].numpy())
        mean_boxes, mean_scores, _ = tta._get_augmented_boxes(views, tfms)
        out["mean_boxes"] = compact_boxes(mean_boxes.numpy())
        out["mean_scores"] = mean_scores.numpy()
        res = tta([dict(d)])[0]["instances"]
    out["det_boxes"] = res.pred_boxes.tensor.numpy()
    out["det_scores"] = res.scores.numpy()
    out["det_classes"] = res.pred_classes.numpy()
    path = os.path.join(helpers.GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, "views", len(views), "dets", len(res), "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(helpers.TTA_CASES)):
        run_case(c)
