"""Golden gradient fingerprints from the UNMODIFIED reference (imported under oracle/refstub.py, CPU, fp32):
train-mode forward with the box head's dropout off, `sum(loss_dict.values()).backward()` through torch autograd
exactly as detectron2/engine/train_loop.py:215-240 does, then for every trainable parameter (everything under
roi_heads; FREEZE_AT 5 freezes the backbone) the L2 norm, the sum and 509 evenly spaced elements of `.grad`.

    python tests/golden/make_golden_grads.py [case ...]          -> tests/golden/<case>_grads.npz
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


def run_case(case):
    name, yaml_rel, ov, _ = helpers.CASES[case]
    cfg, model = refstub.build_reference_model(yaml_rel, [] if case in helpers.OURS_ONLY else [str(x) if not isinstance(x, str) else x for x in ov])
    from detectron2.structures import Boxes, Instances
    from detectron2.utils.events import EventStorage

    ours_cfg = helpers.case_config(case)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    sd = dict(helpers.case_weights(ours_cfg, shapes))
    sd["pixel_mean"] = model.state_dict()["pixel_mean"]
    sd["pixel_std"] = model.state_dict()["pixel_std"]
    model.load_state_dict(sd, strict=True)
    inputs = helpers.case_inputs(case)
    model.train()
    model.roi_heads.box_head.eval()  # dropout off: masks of different RNGs are not comparable
    with EventStorage():
        losses = model(helpers.to_batched(inputs, Instances, Boxes, train=True))
        sum(losses.values()).backward()
    out = {"loss/" + k: np.float32(v.item()) for k, v in losses.items()}
    trainable = []
    for k, p in model.named_parameters():
        if p.requires_grad:
            trainable.append(k)
            if p.grad is None:
                out[f"grad/{k}/none"] = np.int64(1)
                continue
            for f, v in helpers.grad_summary(p.grad).items():
                out[f"grad/{k}/{f}"] = v
    assert all(k.startswith("roi_heads.") for k in trainable), trainable
    out["trainable"] = np.array(trainable)
    path = os.path.join(helpers.GOLDEN_DIR, case + "_grads.npz")
    np.savez_compressed(path, **out)
    print(case, len(trainable), "trainable tensors ->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    for c in (sys.argv[1:] or helpers.GRAD_CASES):
        run_case(c)
