#!/usr/bin/env python
"""Where the end-to-end loop's time goes (GPU box only): wall-clock ms/step of the public forward under five input / output
regimes -- device-resident inputs with and without the host read of the loss, pinned host inputs with and without
model.prefetch and the read.  Not a bench value.

    python tools/e2e_probe.py [--workload r50_bf16] [--steps 60]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16")
    ap.add_argument("--steps", type=int, default=60)
    args = ap.parse_args()
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth

    cfg_name, H, W, R, precision, gmac = bench.WORKLOADS[args.workload]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    dev = torch.device("cuda:0")
    inp = synth.make_inputs(H, W, R, seed=0)
    on_dev = bench.make_batched(inp, dev, drn)
    host = bench.make_batched(inp, None, drn, pinned=True)[0]

    def step(b):
        with torch.no_grad():
            losses = model(b)
        return torch.stack([losses[k] for k in sorted(losses)])

    for _ in range(5):
        v = step(on_dev)
        step([host])
    torch.cuda.synchronize()
    pinned = [torch.empty(v.shape, dtype=v.dtype).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def loop(inputs, prefetch, read):
        batches = [[host], [host]] if inputs == "host" else [on_dev, on_dev]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if prefetch:
            model.prefetch(batches[0])
        for i in range(args.steps):
            v = step(batches[i & 1])
            if read:
                pinned[i & 1].copy_(v, non_blocking=True)
                done[i & 1].record()
            if prefetch:
                model.prefetch(batches[(i + 1) & 1])
            if read and i > 0:
                done[(i - 1) & 1].synchronize()
                pinned[(i - 1) & 1].clone()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.steps * 1e3

    out = {}
    for name, cfg_ in [("A device inputs, no read", ("dev", False, False)), ("B device inputs, read one step late", ("dev", False, True)),
                       ("C host inputs + prefetch, no read", ("host", True, False)), ("D host inputs + prefetch + read (= e2e)", ("host", True, True)),
                       ("E host inputs, copy inside forward, read", ("host", False, True))]:
        loop(*cfg_)
        out[name] = round(min(loop(*cfg_) for _ in range(3)), 4)
        print(f"{name:45s} {out[name]:.4f} ms/step", flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
