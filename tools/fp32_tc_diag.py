#!/usr/bin/env python
"""Accuracy diagnostics of the "fp32_tc" precision (split-bf16 GEMMs, csrc/drn_split.cu): error of a linear layer versus
float64 for K = 512 .. 25088, the signed bias on all-positive operands (exposes the tensor
core's accumulator rounding mode), the same for plain bf16-exact operands, and the whole-model error on one golden
case.  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from drn_wsod_pytorch_b200 import modeling

DEV = "cuda:0"


def linear(M, K, N, nterms, positive):
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    if positive:
        x, w = x.abs(), w.abs()
    b = torch.zeros(N)
    ref = x.double() @ w.double().t()
    scale = x.double().abs() @ w.double().abs().t()
    packed = modeling.pack_linear([w.to(DEV)], [b.to(DEV)], "fp32_tc")
    y = modeling.run_linear(x.to(DEV), packed, "fp32_tc", relu=False).cpu().double()
    pf = modeling.pack_linear([w.to(DEV)], [b.to(DEV)], "fp32")
    ys = modeling.run_linear(x.to(DEV), pf, "fp32", relu=False).cpu().double()[:, :N]
    yt = (x.to(DEV) @ w.to(DEV).t()).cpu().double()  # cuBLAS fp32 (TF32 off by default)
    out = {}
    for name, v in (("fp32_tc", y[:, :N]), ("simt_fp32", ys), ("cublas_fp32", yt)):
        out[name] = (float(((v - ref).abs() / scale).max()), float(((v - ref) / ref).mean()) if positive else float("nan"))
    return out


def bf16_exact_bias(M, K, N):
    """plain bf16 GEMM on bf16-exact positive operands: every product is exact, so the error is the accumulator's."""
    g = torch.Generator().manual_seed(K)
    x = torch.rand(M, K, generator=g).to(torch.bfloat16)
    w = (torch.rand(N, K, generator=g) / K ** 0.5).to(torch.bfloat16)
    ref = x.double() @ w.double().t()
    p = modeling.pack_linear([w.float().to(DEV)], [torch.zeros(N, device=DEV)], "bf16")
    y = modeling.run_linear(x.to(DEV), p, "bf16", relu=False, out_dtype=torch.float32).cpu().double()[:, :N]
    return float(((y - ref) / ref).mean()), float(((y - ref) / ref).abs().max())


def model_case(case, precision, nterms=6):
    g = helpers.load_golden(case)
    cfg = helpers.case_config(case, device=DEV, precision=precision)
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.eval()
    inputs = helpers.case_inputs(case)
    batched = helpers.to_batched(inputs, drn.Instances, drn.Boxes, device=DEV, train=False)
    with torch.no_grad():
        _, all_scores, _ = model.inference(batched, do_postprocess=False)
    a, b = all_scores[0][0].cpu().numpy().astype(np.float64), g["img0/eval/all_scores"].astype(np.float64)
    floor = 1e-3 * np.abs(b).max(axis=0, keepdims=True) + 1e-30
    rel = np.abs(a - b) / (np.abs(b) + floor)
    big = b > 0.1 * b.max()
    return float(rel.max()), float(((a - b) / b)[big].mean()), float(np.abs((a - b) / b)[big].max())


if __name__ == "__main__":
    for K in (512, 4096, 25088):
        print(f"bf16-exact positive operands, K={K}: mean rel err {bf16_exact_bias(256, K, 128)[0]:+.3e} "
              f"(K/16 steps x 2^-25 = {K / 16 * 2 ** -25:.3e})")
    for (M, K, N) in ((300, 512, 128), (300, 4096, 128), (129, 25088, 64)):
        for nterms in (6,):
            for positive in (False, True):
                r = linear(M, K, N, nterms, positive)
                print(f"linear M={M} K={K} N={N} planes={nterms} positive={positive}: " +
                      "  ".join(f"{k}: max|err|/sum|xw| {v[0]:.2e} bias {v[1]:+.2e}" for k, v in r.items()))
    for case in ("oicr_r18_small", "wsddn_v16_300"):
        for precision, nterms in (("fp32", 6), ("fp32_tc", 6), ("bf16", 6)):
            e = model_case(case, precision, nterms)
            print(f"model {case} {precision} planes={nterms}: eval all_scores score_err {e[0]:.3e}, on scores > 10% of max: mean rel {e[1]:+.3e} max rel {e[2]:.3e}")
