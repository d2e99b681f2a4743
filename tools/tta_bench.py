#!/usr/bin/env python
"""Test-time-augmentation throughput: one synthetic uint8 600x1000 image + R proposals through
GeneralizedRCNNWithTTAAVG with the reference's TEST.AUG block (8 scales x flip = 16 views), end to end from PINNED HOST
inputs: H2D of the image, 8 on-device resamples (+ fused flips), 16 eval forwards (captured plans), on-device merge,
one threshold / NMS / top-k, host read of the detection count.  GPU box only.  Prints one JSON line.

    python tools/tta_bench.py [--workload r50_bf16] [--steps 5] [--warmup 3] [--cpu-sample]

--cpu-sample also times Pillow's 8 resizes (+ flips) of the same image on the host -- the part of the reference's TTA mapper
that runs on the CPU even when the model is on a GPU.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import helpers  # noqa: E402
import drn_wsod_pytorch_b200 as drn  # noqa: E402
from drn_wsod_pytorch_b200 import synth, tta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-sample", action="store_true")
    ap.add_argument("--per-view", action="store_true",
                    help="also report, for one image: host enqueue time, GPU time per view (CUDA events around each model call), "
                         "mapper GPU time -- tells host-bound gaps from slow kernels")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE image between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    cfg_name, H, W, R, precision, _ = bench.WORKLOADS[args.workload]
    dev = "cuda:0"
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", dev, "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    del weights
    model.eval()
    inp = synth.make_inputs(H, W, R, seed=0)
    image_u8 = inp["image"].to(torch.uint8).pin_memory()
    prop = drn.Instances((H, W), proposal_boxes=drn.Boxes(inp["boxes"].pin_memory()), objectness_logits=inp["objectness"].pin_memory())
    batched = [{"image": image_u8, "proposals": prop, "height": H, "width": W}]
    wrapper = tta.GeneralizedRCNNWithTTAAVG(cfg, model)
    n_views = len(cfg.TEST.AUG.MIN_SIZES) * (2 if cfg.TEST.AUG.FLIP else 1)

    def step():
        with torch.no_grad():
            inst = wrapper(batched)[0]["instances"]
        return len(inst)

    for _ in range(max(args.warmup, 3)):  # pass 1 eager, pass 2 captures one plan per scale, pass 3+ replays
        n_det = step()
    torch.cuda.synchronize()
    if args.profile_step:
        torch.cuda.cudart().cudaProfilerStart()
        step()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        n_det = step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    ms = e0.elapsed_time(e1) / args.steps
    shapes = [tta.ResizeShortestEdge(s, cfg.TEST.AUG.MAX_SIZE).get_shape(H, W) for s in cfg.TEST.AUG.MIN_SIZES]
    line = {"metric": "images/sec with test-time augmentation (TEST.AUG of the WSL configs), end to end from pinned host inputs",
            "value": 1.0 / wall, "unit": "images/sec", "ms_per_image": wall * 1e3, "ms_per_image_device": ms, "views": n_views,
            "views_per_sec": n_views / wall, "view_shapes": shapes, "detections": n_det, "steps": args.steps,
            "warmup": max(args.warmup, 3), "dtype": precision,
            "config": {"workload": f"{cfg_name} {H}x{W} R={R} {precision}, MIN_SIZES {list(cfg.TEST.AUG.MIN_SIZES)} FLIP {cfg.TEST.AUG.FLIP}"},
            "h2d_bytes_per_image": image_u8.numel() + n_views * R * 20, "mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    if args.per_view:
        orig_inf = model.inference
        evs = []

        def timed_inference(*a, **k):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            out = orig_inf(*a, **k)
            a1.record()
            evs.append((a0, a1))
            return out

        model.inference = timed_inference
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        b0.record()
        with torch.no_grad():
            t_map = 0.0  # views are streamed: the mapper's host work is interleaved with the launches
            wrapper([dict(batched[0])])
            t_enq = time.perf_counter() - t0
        b1.record()
        torch.cuda.synchronize()
        model.inference = orig_inf
        line["per_view"] = {"host_mapper_ms": t_map * 1e3, "host_enqueue_all_views_ms": t_enq * 1e3, "gpu_total_ms": b0.elapsed_time(b1),
                            "gpu_ms_per_view": [round(a.elapsed_time(b), 3) for a, b in evs],
                            "gpu_ms_views_sum": sum(a.elapsed_time(b) for a, b in evs)}
    if args.cpu_sample:
        try:
            from PIL import Image

            arr = np.ascontiguousarray(image_u8.permute(1, 2, 0).numpy())
            t0 = time.perf_counter()
            for nh, nw in shapes:
                r = np.asarray(Image.fromarray(arr).resize((nw, nh), Image.BILINEAR))
                np.ascontiguousarray(r.transpose(2, 0, 1)), np.ascontiguousarray(np.flip(r, axis=1).transpose(2, 0, 1))
            line["cpu_resize_ms_pillow"] = (time.perf_counter() - t0) * 1e3
        except ImportError:
            pass
    print(json.dumps(line))


if __name__ == "__main__":
    main()
