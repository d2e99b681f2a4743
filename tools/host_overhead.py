#!/usr/bin/env python
"""Host-side cost of one public forward call (GPU box only): wall time of the call with an idle GPU (launch path
only, no sync inside), and a cProfile of 200 calls -- what separates the e2e number from the device-timed one."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth

    cfg_name, H, W, R, precision, _ = bench.WORKLOADS["r50_bf16"]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    inp = synth.make_inputs(H, W, R, seed=0)
    dev_b = bench.make_batched(inp, torch.device("cuda:0"), drn)
    host_b = bench.make_batched(inp, None, drn, pinned=True)
    with torch.no_grad():
        for b in (dev_b, host_b):
            for _ in range(5):
                model(b)
            torch.cuda.synchronize()
            ts = []
            for _ in range(50):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = model(b)
                ts.append(time.perf_counter() - t0)
            torch.cuda.synchronize()
            ts.sort()
            print(f"{'device' if b is dev_b else 'pinned host'} inputs: host time per call median {ts[25] * 1e6:.0f} us, min {ts[0] * 1e6:.0f} us")
            t0 = time.perf_counter()
            for _ in range(100):
                v = torch.stack(list(model(b).values())).cpu()
            print(f"  with a D2H loss read every step: {(time.perf_counter() - t0) * 10:.3f} ms/step")
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(200):
            model(host_b)
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
