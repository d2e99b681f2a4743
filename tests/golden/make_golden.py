"""Generate the committed golden vectors by running the UNMODIFIED reference
(/root/reference, imported under oracle/refstub.py) on CPU, on the seeded synthetic inputs and
weights of tests/helpers.py.  Only runnable where /root/reference exists.

    python tests/golden/make_golden.py [case ...]

Per case, `<case>.npz` holds: the loss dict, per image the WSDDN scores / image scores, per OICR
stage the pseudo-GT argmax indices, boxes, weights, labels, matched indices and softmax probs,
and the eval-mode all_scores / all_boxes / final detections -- plus checksums of the regenerated
inputs and weights so a test can prove it rebuilt the same tensors.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from drn_wsod_pytorch_b200 import synth  # noqa: E402
from oracle import refstub  # noqa: E402


def run_case(case):
    name, yaml_rel, ov, _ = helpers.CASES[case]
    cfg, model = refstub.build_reference_model(yaml_rel, [] if case in helpers.OURS_ONLY else [str(x) if not isinstance(x, str) else x for x in ov])
    from detectron2.structures import Boxes, Instances
    from detectron2.utils.events import EventStorage

    ours_cfg = helpers.case_config(case)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    weights = helpers.case_weights(ours_cfg, shapes)
    sd = dict(weights)
    sd["pixel_mean"] = model.state_dict()["pixel_mean"]
    sd["pixel_std"] = model.state_dict()["pixel_std"]
    model.load_state_dict(sd, strict=True)
    inputs = helpers.case_inputs(case)
    out = {"weights_checksum": np.float64(synth.weights_checksum(weights)),
           "inputs_checksum": np.float64(sum(float(i["image"].double().sum() + i["boxes"].double().sum()) for i in inputs))}

    heads = model.roi_heads
    rec = {"pgt": [], "labels": []}
    if hasattr(heads, "get_pgt"):
        orig_pgt = heads.get_pgt

        def get_pgt(prev_boxes, prev_scores, proposals, suffix):
            if isinstance(prev_scores, torch.Tensor):
                ps = list(prev_scores.split([len(p) for p in proposals], dim=0))
            else:
                ps = list(prev_scores)
            idx = [torch.max(torch.index_select(s, 1, g), dim=0)[1] for s, g in zip(ps, heads.gt_classes_img_int)]
            targets, weights_ = orig_pgt(prev_boxes, prev_scores, proposals, suffix)
            rec["pgt"].append((idx, [t.gt_boxes.tensor.clone() for t in targets], [t.gt_scores.clone() for t in targets],
                               [w.clone() for w in weights_], [s.clone() for s in ps]))
            return targets, weights_

        heads.get_pgt = get_pgt
    orig_label = heads.label_and_sample_proposals

    def label(proposals, targets, ret_MI=False, suffix=""):
        res = orig_label(proposals, targets, ret_MI=True, suffix=suffix)
        rec["labels"].append(([p.gt_classes.clone() for p in res[0]], [m.clone() for m in res[1]]))
        return res if ret_MI else res[0]

    heads.label_and_sample_proposals = label
    wsddn_out = {}
    heads.box_predictor.register_forward_hook(lambda m, i, o: wsddn_out.__setitem__("scores", o[0].detach().clone()))
    ref_probs = []
    for k in range(getattr(heads, "refine_K", 0)):
        heads.box_refinery[k].register_forward_hook(lambda m, i, o: ref_probs.append(torch.softmax(o[0].detach(), -1)))

    # ---- train-mode forward with dropout off ----
    model.train()
    model.roi_heads.box_head.eval()
    with EventStorage() as storage, torch.no_grad():
        losses = model(helpers.to_batched(inputs, Instances, Boxes, train=True))
    for k, v in losses.items():
        out["loss/" + k] = np.float32(v.item())
    ns = [len(i["boxes"]) for i in inputs]
    for i, s in enumerate(wsddn_out["scores"].split(ns, 0)):
        out[f"img{i}/scores"] = s.numpy()
        out[f"img{i}/img_score"] = heads.pred_class_img_logits[i].numpy() if hasattr(heads, "pred_class_img_logits") else \
            torch.clamp(s.sum(0), 1e-6, 1 - 1e-6).numpy()
    out["labels_gt_n"] = np.int64(len(rec["labels"]))
    for i in range(len(inputs)):
        out[f"img{i}/labels_gt"] = rec["labels"][0][0][i].numpy()
    for k, (idx, boxes, scores, wts, prev) in enumerate(rec["pgt"]):
        labs, mis = rec["labels"][k + 1]
        probs_k = ref_probs[k].split(ns, 0)
        for i in range(len(inputs)):
            p = f"img{i}/stage{k}/"
            out[p + "pgt_idx"] = idx[i].numpy()
            out[p + "pgt_boxes"] = boxes[i].numpy()
            out[p + "pgt_scores"] = scores[i].numpy()
            out[p + "pgt_weights"] = wts[i].numpy()
            out[p + "labels"] = labs[i].numpy()
            out[p + "matched"] = mis[i].numpy()
            out[p + "probs"] = probs_k[i].numpy()
            # margin of each per-class argmax over the runner-up (for the parity report)
            sel = torch.index_select(prev[i], 1, heads.gt_classes_img_int[i])
            top2 = torch.topk(sel, 2, dim=0)[0]
            out[p + "argmax_margin_rel"] = ((top2[0] - top2[1]) / top2[0].abs().clamp(min=1e-30)).numpy()
    # ---- eval ----
    model.eval()
    with torch.no_grad():
        results, all_scores, all_boxes = model.inference(helpers.to_batched(inputs, Instances, Boxes, train=False),
                                                         do_postprocess=False)
    for i in range(len(inputs)):
        out[f"img{i}/eval/all_scores"] = all_scores[i][0].numpy()
        out[f"img{i}/eval/all_boxes"] = all_boxes[i][0].numpy()
        out[f"img{i}/eval/det_boxes"] = results[i].pred_boxes.tensor.numpy()
        out[f"img{i}/eval/det_scores"] = results[i].scores.numpy()
        out[f"img{i}/eval/det_classes"] = results[i].pred_classes.numpy()
    path = os.path.join(helpers.GOLDEN_DIR, case + ".npz")
    np.savez_compressed(path, **out)
    print(case, {k: float(v) for k, v in out.items() if k.startswith("loss/")}, "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(helpers.CASES)):
        run_case(c)
