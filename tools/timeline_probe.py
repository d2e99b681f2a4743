#!/usr/bin/env python
"""Device-side timeline of the conv stack's GEMM launches inside ONE captured graph (GPU box only; run with DRN_TC_DEBUG=256).
Each gemm_tc_kernel launch records %globaltimer at: first CTA start, first CTA past griddepcontrol.wait, first operand tile
landed, last CTA end.  Prints per launch the gap to the previous launch's end and the phases, i.e. where a kernel boundary's
time goes.  Not a bench value (the probe adds a few atomics per CTA).

    DRN_TC_DEBUG=256 python tools/timeline_probe.py [--workload r50_bf16]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16")
    args = ap.parse_args()
    assert int(os.environ.get("DRN_TC_DEBUG", "0")) & 256, "run with DRN_TC_DEBUG=256"
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import lib, ops, synth

    cfg_name, H, W, R, precision, gmac = bench.WORKLOADS[args.workload]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    img = batched[0]["image"].float().contiguous()
    h = lib.load()
    h.drn_gemm_timeline_reset.argtypes = [ctypes.c_void_p]
    h.drn_gemm_timeline_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    with torch.no_grad():
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            model._features([img], (H, W))
            model._features([img], (H, W))
        torch.cuda.synchronize()
        ops.drop_scratch(side)
        assert h.drn_gemm_timeline_reset(None) == 0  # arms the slot counter: the capture below bakes slot i into launch i
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = model._features([img], (H, W))
        n_slots = 512
        buf = (ctypes.c_ulonglong * (4 * n_slots))()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        n = h.drn_gemm_timeline_read(buf, n_slots)
        # values are min / max over ALL replays so far; re-arm the values (not the slots) and replay once
        h.drn_gemm_timeline_reset(None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        h.drn_gemm_timeline_read(buf, n_slots)
    print(f"launches with a slot: {n}; replay {e0.elapsed_time(e1) * 1e3:.1f} us (with the probe)")
    print(f"{'#':>3} {'gap_prev_end->start':>20} {'start->past_wait':>17} {'wait->first_tile':>17} {'first_tile->end':>16} {'total':>8}  (us)")
    prev_end = None
    t0 = buf[0]
    sums = [0.0] * 5
    for i in range(n):
        s, w, f, e = (buf[4 * i + k] for k in range(4))
        gap = (s - prev_end) / 1e3 if prev_end is not None else 0.0
        row = [gap, (w - s) / 1e3, (f - w) / 1e3, (e - f) / 1e3, (e - s) / 1e3]
        sums = [a + b for a, b in zip(sums, row)]
        print(f"{i:3d} {row[0]:20.2f} {row[1]:17.2f} {row[2]:17.2f} {row[3]:16.2f} {row[4]:8.2f}")
        prev_end = e
    print("sum " + " ".join(f"{x:17.2f}" for x in sums))
    print(f"first start -> last end: {(prev_end - t0) / 1e3:.1f} us")


if __name__ == "__main__":
    main()
