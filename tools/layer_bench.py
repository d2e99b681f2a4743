#!/usr/bin/env python
"""Per-layer warm timing of the hot path (GPU box only).

Records every C-ABI GEMM/conv call of one forward of a workload, then replays each distinct
signature REPS times back to back (inputs L2-warm, launches overlapped) and prints us/call,
TFLOP/s and the ideal time at the measured bf16 peak.  Used to rank layers, not as a bench value.

    python tools/layer_bench.py [--workload r50_bf16] [--reps 20]
"""
import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import ops, synth

    cfg_name, H, W, R, precision, _ = bench.WORKLOADS[args.workload]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    model.use_cuda_graph = False  # record the individual C-ABI calls
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    calls = []
    orig = ops.conv_bf16_tc

    def rec(x, packed, ksize, dilation, relu, residual=None, **kw):
        calls.append((x, packed, ksize, dilation, relu, residual, kw))
        return orig(x, packed, ksize, dilation, relu, residual, **kw)

    ops.conv_bf16_tc = rec
    model(batched)
    model(batched)
    calls.clear()
    model(batched)
    torch.cuda.synchronize()
    ops.conv_bf16_tc = orig
    agg = collections.OrderedDict()
    for (x, packed, ksize, dil, relu, res, kw) in calls:
        N, Hh, Ww, Cin = x.shape
        key = (N * Hh * Ww if ksize == 1 else (Hh, Ww), Cin, packed["cout"], ksize, dil, res is not None)
        agg.setdefault(key, [0, (x, packed, ksize, dil, relu, res, kw)])[0] += 1
    tot = 0.0
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
    print(f"{'rows/HxW':>12} {'Cin':>6} {'Cout':>5} k d res  count   us/call  TFLOP/s  ideal_us   total_us")
    for key, (cnt, (x, packed, ksize, dil, relu, res, kw)) in agg.items():
        for _ in range(3):
            orig(x, packed, ksize, dil, relu, res, **kw)
        # capture REPS back-to-back launches in a CUDA graph so the number is GPU time, not Python launch time
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(args.reps):
                orig(x, packed, ksize, dil, relu, res, **kw)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.reps
        N, Hh, Ww, Cin = x.shape
        flops = 2.0 * N * Hh * Ww * Cin * ksize * ksize * packed["cout"]
        tf = flops / (us * 1e-6) / 1e12
        print(f"{str(key[0]):>12} {Cin:6d} {packed['cout']:5d} {ksize} {dil} {int(res is not None)}   {cnt:5d} {us:9.1f} {tf:8.1f} {flops / (peak * 1e12) * 1e6:9.1f} {us * cnt:10.1f}")
        tot += us * cnt
    print(f"sum over layers (warm, back-to-back): {tot:.1f} us")


if __name__ == "__main__":
    main()
