#!/usr/bin/env python
"""Achieved HBM bandwidth of the streaming kernels added for TTA and the fp32_tc precision, at the bench workloads' sizes:
algorithmic bytes / warm CUDA-event time (inputs larger than L2 where the workload's are; 20 launches back to back).
GPU box only.  One JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from drn_wsod_pytorch_b200 import ops, tta

DEV = "cuda:0"
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def row(name, nbytes, ms):
    return {"kernel": name, "algorithmic_MB": round(nbytes / 1e6, 1), "us": round(ms * 1e3, 1), "GBps": round(nbytes / ms / 1e6, 1),
            "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 3)}


out = []
# fp32_tc split of the pooled ROI features (R18: 2000 x 25088 fp32 -> big 2 B + small 10 B per element)
x = torch.randn(2000, 25088, device=DEV)
out.append(row("drn_f32tc_split 2000x25088", x.numel() * (4 + 2 + 10), timed(lambda: ops.f32tc_split(x, 896))))
del x
# fp32_tc reduction of fc6's partials (1 + 28 partials of 2000 x 4096) + bias + ReLU
p = torch.randn(29, 2000, 4096, device=DEV)
b = torch.randn(4096, device=DEV)
out.append(row("drn_f32tc_reduce 29x2000x4096", p.numel() * 4 + 2000 * 4096 * 4, timed(lambda: ops.f32tc_reduce(p, b, None, True))))
del p
# TTA resample of the bench image to the largest view (fp32 output, flip fused)
img = torch.randint(0, 256, (3, 600, 1000), dtype=torch.uint8, device=DEV)
ms = timed(lambda: tta.resize_u8(img, 1152, 1920, flip=True, out_dtype=torch.float32))
out.append(row("drn_resample_u8_fwd 600x1000 -> 1152x1920 fp32", 3 * (600 * 1000 + 2 * 600 * 1920 + 1152 * 1920 * 4), ms))
# TTA merge of one view (R = 4000, 80 box columns, 21 score columns): read + accumulate + write
bx, sc = torch.rand(4000, 80, device=DEV), torch.rand(4000, 21, device=DEV)
ab, asc = torch.zeros_like(bx), torch.zeros_like(sc)
prm = tta.TransformList([tta.ResizeTransform(600, 1000, 1152, 1920), tta.HFlipTransform(1920)]).inverse().device_params()
out.append(row("drn_tta_accumulate R=4000", 3 * (bx.numel() + sc.numel()) * 4, timed(lambda: ops.tta_accumulate(bx, sc, prm, ab, asc, 1, 16))))
print(json.dumps({"hbm_peak_GBps": PEAK, "kernels": out}))
