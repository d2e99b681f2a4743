"""Shared fixtures for the parity tests: case table, seeded inputs/weights, structure conversion."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import drn_wsod_pytorch_b200 as drn  # noqa: E402
from drn_wsod_pytorch_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (builtin config, reference YAML (relative to projects/WSL/configs), overrides, [(H, W, R, G, seed)...])
CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable plumbing case
    "wsddn_v16_300": ("wsddn_V_16_DC5_1x", "PascalVOC-Detection/wsddn_V_16_DC5_1x.yaml", [], [(300, 300, 64, 2, 11)]),
    "oicr_r18_small": ("oicr_WSR_18_DC5_1x", "PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml", [], [(160, 224, 96, 2, 12)]),
    "oicr_r50_small": ("oicr_WSR_50_DC5_1x", "PascalVOC-Detection/oicr_WSR_50_DC5_1x.yaml", [], [(128, 192, 64, 1, 13)]),
    "oicr_v16_small": ("oicr_V_16_DC5_1x", "PascalVOC-Detection/oicr_V_16_DC5_1x.yaml", [], [(128, 160, 64, 3, 14)]),
    "oicr_r101_coco_small": ("oicr_WSR_101_DC5_1x_coco", "COCO-Detection/oicr_WSR_101_DC5_1x.yaml", [], [(128, 160, 64, 5, 15)]),
    "oicr_r18_reg": ("oicr_WSR_18_DC5_1x", "PascalVOC-Detection/reg/oicr_WSR_18_DC5_1x.yaml",
                     ["WSL.REFINE_NUM", 4, "WSL.REFINE_REG", [False, False, False, True]], [(128, 192, 64, 2, 16)]),
    "oicr_r18_batch2": ("oicr_WSR_18_DC5_1x", "PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml", [],
                        [(128, 192, 48, 2, 17), (160, 160, 80, 1, 18)]),
}
# overrides that exist only on our side (the reference YAML already carries them)
OURS_ONLY = {"oicr_r18_reg"}


arch_key = synth.arch_key


def case_inputs(case):
    _, _, _, imgs = CASES[case]
    K = 80 if "coco" in case else 20
    return [synth.make_inputs(H, W, R, seed=seed, num_gt=G, num_classes=K) for (H, W, R, G, seed) in imgs]


def case_config(case, device="cpu", precision=None):
    name, _, ov, _ = CASES[case]
    ov = list(ov) + ["MODEL.DEVICE", device]
    if precision:
        ov += ["B200.PRECISION", precision]
    return drn.builtin_config(name, ov)


case_weights = synth.calibrated_weights


def to_batched(inputs, inst_cls, box_cls, device="cpu", train=True):
    """Our synthetic dicts -> the reference's batched_inputs format (rcnn.py:138-160)."""
    out = []
    for inp in inputs:
        H, W = inp["height"], inp["width"]
        d = {"image": inp["image"].to(device), "height": H, "width": W}
        p = inst_cls((H, W))
        p.proposal_boxes = box_cls(inp["boxes"].clone().to(device))
        p.objectness_logits = inp["objectness"].clone().to(device)
        d["proposals"] = p
        if train:
            g = inst_cls((H, W))
            g.gt_boxes = box_cls(inp["gt_boxes"].clone().to(device))
            g.gt_classes = inp["gt_classes"].clone().to(device)
            d["instances"] = g
        out.append(d)
    return out


def load_golden(case):
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / (np.abs(b) + floor))) if a.size else 0.0


def rand_dets(R, K, seed, nreg_k=True, cluster=True):
    g = torch.Generator().manual_seed(seed)
    if cluster:  # heavily overlapping boxes around a few centres, like proposals around objects
        ctr = torch.rand(8, 2, generator=g) * torch.tensor([900.0, 500.0]) + 50
        which = torch.randint(0, 8, (R,), generator=g)
        c = ctr[which] + torch.randn(R, 2, generator=g) * 25
        wh = torch.rand(R, 2, generator=g) * 150 + 30
        boxes = torch.cat([c - wh / 2, c + wh / 2], dim=1)
    else:
        xy = torch.rand(R, 2, generator=g) * torch.tensor([900.0, 500.0])
        boxes = torch.cat([xy, xy + torch.rand(R, 2, generator=g) * 200 + 5], dim=1)
    scores = torch.softmax(torch.randn(R, K + 1, generator=g) * 3, dim=1)
    nreg = K if nreg_k else 1
    all_boxes = (boxes[:, None, :] + torch.randn(R, nreg, 4, generator=g) * (2.0 if nreg_k else 0.0)).reshape(R, 4 * nreg)
    return all_boxes.contiguous(), scores.contiguous()


GRAD_CASES = ["wsddn_v16_300", "oicr_r18_small", "oicr_r18_reg", "oicr_r18_batch2", "oicr_r50_small"]


def grad_summary(g, n=509):
    """Compact fingerprint of a gradient tensor for the committed goldens: L2 norm, sum, and `n` evenly spaced
    elements of the flattened tensor (all of it when it is small)."""
    g = g.detach().double().flatten().cpu()
    idx = torch.arange(g.numel()) if g.numel() <= 4096 else (torch.arange(n, dtype=torch.int64) * (g.numel() - 1)) // (n - 1)
    return {"norm": np.float64(g.norm().item()), "sum": np.float64(g.sum().item()), "sample": g[idx].float().numpy(),
            "index": idx.numpy()}


def check_grad(g, gold, prefix, rtol, what="", atol=1e-6, bad_frac=0.005):
    """Compare a gradient with its golden fingerprint: L2 norm within rtol, and all but a fraction `bad_frac` of the
    sampled elements within rtol of the tensor's scale (max |sample|).
    * The outlier allowance is for ReLU-boundary flips: a pre-activation within rounding noise of zero is positive
      in one implementation and negative in the other (different fp32 summation order in the GEMMs), which switches
      that unit's whole gradient on or off -- about one unit per image in fp32, a fraction of a percent in bf16.
    * Absolute floors (`atol` on the norm, 1e-4 on the scale) cover gradients that are analytically zero --
      `det.bias`: the softmax over proposals is invariant to a per-class shift -- where both sides hold rounding
      noise; every real gradient of the golden cases has a scale above 3e-3."""
    s = grad_summary(g)
    norm, gn = s["norm"], float(gold[prefix + "norm"])
    assert abs(norm - gn) <= rtol * gn + atol, (what, prefix, norm, gn)
    ref = gold[prefix + "sample"].astype(np.float64)
    scale = max(np.abs(ref).max(), 1e-4)
    err = np.abs(s["sample"].astype(np.float64) - ref) / scale
    nbad = int((err > rtol).sum())
    assert nbad <= max(1, int(bad_frac * err.size)), (what, prefix, nbad, err.size, float(err.max()))


# ---------------------------------------------------------------- test-time augmentation
# name -> (CASES entry, TEST.AUG.MIN_SIZES, TEST.AUG.MAX_SIZE, TEST.AUG.FLIP, dataset (height, width) or None)
TTA_CASES = {
    # three scales x flip; the largest scale hits MAX_SIZE (224 -> 224x314 -> capped to 214x300)
    "tta_r18_small": ("oicr_r18_small", (128, 160, 224), 300, True, None),
    # the dataset image is larger than the model input: pre-transform + its inverse on the way back
    "tta_r18_pre": ("oicr_r18_small", (160, 192), 4000, True, (320, 448)),
    # WSDDN head (zero background column, identity boxes), no flip
    "tta_wsddn_v16": ("wsddn_v16_300", (240, 300), 4000, False, None),
}


def tta_input(name):
    """The CASES entry's first image with a uint8 BGR image (the dataset mapper hands uint8 CHW tensors to the TTA
    driver): low-frequency pattern + noise, so that the resampling filter matters."""
    case = TTA_CASES[name][0]
    inp = dict(case_inputs(case)[0])
    H, W = inp["height"], inp["width"]
    rng = np.random.Generator(np.random.PCG64(1000 + len(name)))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    chans = [128 + 70 * np.sin(xx / (9.0 + 4 * c) + c) * np.cos(yy / (7.0 + 3 * c)) + 45 * (rng.random((H, W)) - 0.5) for c in range(3)]
    inp["image_u8"] = torch.from_numpy(np.clip(np.stack(chans), 0, 255).astype(np.uint8))
    inp["image"] = inp["image_u8"].float()
    return inp
