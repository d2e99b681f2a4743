#!/usr/bin/env python
"""A/B of the ROIPool/fc6 row-block overlap on the bench workload (GPU box only): ms/step with
DRN_B200_OVERLAP_POOL off and on, graph-replayed, CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth

    wl = sys.argv[1] if len(sys.argv) > 1 else "r50_bf16"
    cfg_name, H, W, R, precision, _ = bench.WORKLOADS[wl]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    for overlap, k in ((False, 0), (True, 0), (False, 0), (True, 0)):
        model.roi_heads.overlap_pool = overlap
        model.roi_heads.pool_ctas_per_sm = k
        model.invalidate_plans()
        for _ in range(5):
            model(batched)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            model(batched)
        e1.record()
        torch.cuda.synchronize()
        print(f"overlap_pool={overlap} ctas_per_sm={k}: {e0.elapsed_time(e1) / 20:.3f} ms/step", flush=True)


if __name__ == "__main__":
    main()
