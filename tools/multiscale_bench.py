#!/usr/bin/env python
"""Dynamic-shape path (GPU box only): multi-scale training as the WSL configs run it -- every iteration a new (H, W):
`INPUT.MIN_SIZE_TRAIN` = 480, 512, ..., 1216 (24 scales), MAX_SIZE_TRAIN 2000
(projects/WSL/configs/PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml:32-38), VOC-like aspect ratios, R = 2000 proposals.

Reports ms/step (forward + loss, device-timed over the whole sequence, host launch time included) for
  * "eager":   CUDA graphs off -- every step is ~75 C-ABI launches from Python;
  * "plans":   the default (a signature is captured the 2nd time it is seen, LRU of MAX_PLANS graphs);
and the fixed-shape replay time of the largest view for comparison.

    python tools/multiscale_bench.py [--steps 120] [--workload r18|r50]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=120)
    ap.add_argument("--workload", default="r50")
    args = ap.parse_args()
    import bench
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth

    name = {"r18": "oicr_WSR_18_DC5_1x", "r50": "oicr_WSR_50_DC5_1x"}[args.workload]
    R = 2000
    dev = torch.device("cuda:0")
    cfg = drn.builtin_config(name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", "bf16"])
    model = drn.build_model(cfg)
    weights = synth.calibrated_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    rng = np.random.Generator(np.random.PCG64(7))
    scales = list(range(480, 1217, 32))
    aspects = [500 / 375, 500 / 333, 375 / 500, 500 / 400, 1.0, 500 / 281]  # common VOC shapes
    shapes = []
    for _ in range(args.steps):
        s = int(rng.choice(scales))
        a = float(rng.choice(aspects))
        h, w = (s, int(round(s * a))) if a >= 1 else (int(round(s / a)), s)
        if max(h, w) > 2000:
            f = 2000.0 / max(h, w)
            h, w = int(h * f + 0.5), int(w * f + 0.5)
        shapes.append((h, w))
    distinct = sorted(set(shapes))
    # pinned host batches, one per distinct shape (what a dataloader hands over)
    batches = {hw: bench.make_batched(synth.make_inputs(hw[0], hw[1], R, seed=hw[0] * 7 + hw[1]), None, drn, pinned=True) for hw in distinct}

    def run(mode):
        model.invalidate_plans()
        model.use_cuda_graph = mode != "eager"
        seq = [batches[hw] for hw in shapes]
        with torch.no_grad():
            for b in seq[:3]:
                model(b)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for b in seq:
                out = model(b)
            e1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        return {"ms_per_step_device": e0.elapsed_time(e1) / len(seq), "ms_per_step_wall": wall * 1e3 / len(seq), "plans": len(model._plans),
                "signatures": len({tuple(b[0]["image"].shape) for b in seq})}

    res = {"workload": f"{name} bf16 R={R}, {args.steps} steps, {len(distinct)} distinct (H, W)", "eager": run("eager"), "plans": run("plans")}
    # fixed shape for reference: the mean-size image replayed
    hw = distinct[len(distinct) // 2]
    model.invalidate_plans()
    model.use_cuda_graph = True
    with torch.no_grad():
        for _ in range(4):
            model(batches[hw])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            model(batches[hw])
        e1.record()
        torch.cuda.synchronize()
    res["fixed_shape_replay"] = {"shape": hw, "ms_per_step_device": e0.elapsed_time(e1) / 30}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
