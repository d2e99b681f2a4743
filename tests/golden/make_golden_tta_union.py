"""Golden vectors of the UNMODIFIED reference `GeneralizedRCNNWithTTAUNION` (projects/WSL/wsl/modeling/
test_time_augmentation_union.py) run under oracle/refstub.py on CPU -- per-view detections, their union on the original
image, the final detections.  Only runnable where /root/reference exists.

    python tests/golden/make_golden_tta_union.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import refstub  # noqa: E402

NAME = "tta_r18_small"  # the views of the AVG golden of the same name, other merge


def main():
    case, min_sizes, max_size, flip, dataset_hw = helpers.TTA_CASES[NAME]
    _, yaml_rel, ov, _ = helpers.CASES[case]
    ov = [str(x) if not isinstance(x, str) else x for x in ov] + [
        "TEST.AUG.MIN_SIZES", str(list(min_sizes)), "TEST.AUG.MAX_SIZE", str(max_size), "TEST.AUG.FLIP", str(flip)]
    cfg, model = refstub.build_reference_model(yaml_rel, ov)
    from detectron2.structures import Boxes, Instances
    from wsl.modeling import GeneralizedRCNNWithTTAUNION

    ours_cfg = helpers.case_config(case)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = dict(helpers.case_weights(ours_cfg, shapes))
    sd["pixel_mean"] = model.state_dict()["pixel_mean"]
    sd["pixel_std"] = model.state_dict()["pixel_std"]
    model.load_state_dict(sd, strict=True)
    model.eval()
    inp = helpers.tta_input(NAME)
    d = helpers.to_batched([inp], Instances, Boxes, train=False)[0]
    d["image"] = inp["image_u8"]
    tta = GeneralizedRCNNWithTTAUNION(cfg, model)
    out = {}
    with torch.no_grad():
        views, tfms = tta._get_augmented_inputs(dict(d))
        out["n_views"] = np.int64(len(views))
        for i, v in enumerate(views):
            out[f"view{i}/image_shape"] = np.array(v["image"].shape, dtype=np.int64)
            out[f"view{i}/boxes"] = v["proposals"].proposal_boxes.tensor.numpy()  # UNCHANGED by the UNION mapper
        outputs = tta._batch_inference(views)
        for i, o in enumerate(outputs):
            out[f"view{i}/det_boxes"] = o.pred_boxes.tensor.numpy()
            out[f"view{i}/det_scores"] = o.scores.numpy()
            out[f"view{i}/det_classes"] = o.pred_classes.numpy()
        all_boxes, all_scores, all_classes = tta._get_augmented_boxes(views, tfms)
        out["union_boxes"] = all_boxes.numpy()
        out["union_scores"] = torch.stack(list(all_scores)).numpy() if len(all_scores) else np.zeros((0,), np.float32)
        out["union_classes"] = torch.stack(list(all_classes)).numpy() if len(all_classes) else np.zeros((0,), np.int64)
        res = tta([dict(d)])[0]["instances"]
    out["det_boxes"] = res.pred_boxes.tensor.numpy()
    out["det_scores"] = res.scores.numpy()
    out["det_classes"] = res.pred_classes.numpy()
    path = os.path.join(helpers.GOLDEN_DIR, "tta_union_r18_small.npz")
    np.savez_compressed(path, **out)
    print("views", len(views), "union", len(out["union_scores"]), "dets", len(res), "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
