"""Golden vectors of the UNMODIFIED reference PCL model (ROI_HEADS.NAME PCLROIHeads, projects/WSL/configs/PascalVOC-Detection/
pcl_WSR_18_DC5_1x.yaml) run under oracle/refstub.py on CPU with the reference's own pcl_loss op compiled by oracle/build_ref.py:
losses, per refinement stage the mined clusters (third_party/pcl.py outputs) and the stage softmax, gradient fingerprints of the
trainable parameters from the reference's loss.backward(), and the eval-mode all_scores / detections.

    python tests/golden/make_golden_pcl.py
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

CASE = "oicr_r18_small"   # inputs / weights of this case (the PCL head has the OICR head's parameters)
YAML = "PascalVOC-Detection/pcl_WSR_18_DC5_1x.yaml"


def main():
    cfg, model = refstub.build_reference_model(YAML)
    from detectron2.structures import Boxes, Instances
    from detectron2.utils.events import EventStorage
    import wsl.modeling.roi_heads.fast_rcnn as FR

    ours_cfg = helpers.case_config(CASE)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = dict(helpers.case_weights(ours_cfg, shapes))
    sd["pixel_mean"], sd["pixel_std"] = model.state_dict()["pixel_mean"], model.state_dict()["pixel_std"]
    model.load_state_dict(sd, strict=True)
    inputs = helpers.case_inputs(CASE)[:1]
    rec = []
    orig = FR.PCL

    def spy(boxes, cls_prob, im_labels, cls_prob_new):
        out = orig(boxes, cls_prob, im_labels, cls_prob_new)
        rec.append({k: np.asarray(v).copy() for k, v in out.items()})
        return out

    FR.PCL = spy
    probs = []
    for k in range(model.roi_heads.refine_K):
        model.roi_heads.box_refinery[k].register_forward_hook(lambda m, i, o: probs.append(torch.softmax(o[0].detach(), -1)))
    model.train()
    model.roi_heads.box_head.eval()
    with EventStorage():
        losses = model(helpers.to_batched(inputs, Instances, Boxes, train=True))
        sum(losses.values()).backward()
    out = {"loss/" + k: np.float32(v.item()) for k, v in losses.items()}
    for k, r in enumerate(rec):
        for name, v in r.items():
            out[f"stage{k}/{name}"] = v.reshape(-1)
        out[f"stage{k}/probs"] = probs[k].numpy()
    trainable = []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        trainable.append(name)
        if p.grad is None:
            out[f"grad/{name}/none"] = np.int64(1)
            continue
        s = helpers.grad_summary(p.grad)
        for kk, vv in s.items():
            out[f"grad/{name}/{kk}"] = vv
    out["trainable"] = np.array(trainable)
    probs.clear()
    model.eval()
    with torch.no_grad():
        results, all_scores, all_boxes = model.inference(helpers.to_batched(inputs, Instances, Boxes, train=False), do_postprocess=False)
    out["eval/all_scores"] = all_scores[0][0].numpy()
    out["eval/det_scores"] = results[0].scores.numpy()
    out["eval/det_classes"] = results[0].pred_classes.numpy()
    out["eval/det_boxes"] = results[0].pred_boxes.tensor.numpy()
    path = os.path.join(helpers.GOLDEN_DIR, "pcl_r18_small.npz")
    np.savez_compressed(path, **out)
    print({k: float(v) for k, v in out.items() if k.startswith("loss/")}, "clusters per stage", [len(r["pc_labels"].reshape(-1)) for r in rec],
          "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
