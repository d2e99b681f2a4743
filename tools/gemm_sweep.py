#!/usr/bin/env python
"""GEMM-shape sweep of drn_conv_igemm_bf16_tc (GPU box only): us/call vs (M, N, K), graph-replayed.

    python tools/gemm_sweep.py [--res]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from drn_wsod_pytorch_b200 import ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", action="store_true")
    ap.add_argument("--shapes", default="")
    ap.add_argument("--dropout", type=float, default=0.0)
    args = ap.parse_args()
    dev = "cuda:0"
    shapes = [(9176, n, k) for n in (256, 512, 2048) for k in (256, 512, 1024, 2048, 4096, 8192)]
    shapes += [(37500, 256, 64), (37500, 64, 256), (4000, 4096, 2048), (4000, 2048, 16384)]
    if args.shapes:
        shapes = [tuple(int(x) for x in s.split("x")) for s in args.shapes.split(",")]
    print(f"{'M':>6} {'N':>5} {'K':>6}   us/call  TFLOP/s  us/kblock-per-SM-tile")
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        packed = {"w": w, "scale": None, "bias": torch.zeros(N, device=dev), "cout": N}
        r = torch.randn(M, N, device=dev).bfloat16().view(1, M, 1, N) if args.res else None
        us = timed(lambda: ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True, r, dropout_p=args.dropout, dropout_seed=3))
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        print(f"{M:6d} {N:5d} {K:6d} {us:9.1f} {tf:8.1f}")


if __name__ == "__main__":
    main()
