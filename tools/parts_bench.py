#!/usr/bin/env python
"""Warm per-op timing of ONE step of a workload (GPU box only): every `ops.*` call of a train-mode forward is recorded
with its arguments and replayed REPS times as its own CUDA graph (bench._graph_ms), next to the two parts the bench line
reports (conv stack, ROIPool).  Ranks the non-GEMM kernels (first conv, max-pools, ROIPool, heads tail), which
tools/layer_bench.py does not see.  Not a bench value.

    python tools/parts_bench.py [--workload r50_bf16] [--reps 20]
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

OPS = ("first_conv", "conv_bf16_tc", "maxpool2x2", "roipool", "wsddn_mil_pgt", "oicr_stage_fused", "oicr_stages", "wsddn_mil", "oicr_pgt",
       "label_proposals", "oicr_stage", "oicr_boxreg_loss")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="comma-separated op names to replay (default: all but conv_bf16_tc)")
    args = ap.parse_args()
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import ops, synth

    cfg_name, H, W, R, precision, gmac = bench.WORKLOADS[args.workload]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    model.use_cuda_graph = False
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    calls = []
    saved = {n: getattr(ops, n) for n in OPS}

    def wrap(name, fn):
        def rec(*a, **kw):
            calls.append((name, a, kw))
            return fn(*a, **kw)
        return rec

    for n, fn in saved.items():
        setattr(ops, n, wrap(n, fn))
    model(batched)
    model(batched)
    calls.clear()
    model(batched)
    torch.cuda.synchronize()
    for n, fn in saved.items():
        setattr(ops, n, fn)
    only = set(x for x in args.only.split(",") if x) or (set(OPS) - {"conv_bf16_tc"})
    agg = collections.OrderedDict()
    for name, a, kw in calls:
        if name not in only:
            continue
        shp = tuple(tuple(x.shape) for x in a if torch.is_tensor(x))[:2]
        agg.setdefault((name, shp), [0, a, kw])[0] += 1
    out = {"workload": args.workload, "ops": []}
    with torch.no_grad():
        for (name, shp), (cnt, a, kw) in agg.items():
            ms = bench._graph_ms(lambda: saved[name](*a, **kw), reps=args.reps)
            out["ops"].append({"op": name, "shapes": [list(s) for s in shp], "count": cnt, "us_per_call": round(ms * 1e3, 2)})
            print(f"{name:18s} x{cnt}  {ms * 1e3:8.2f} us/call  {shp}", flush=True)
    parts = bench.measure_parts(model, batched[0], H, W, R, gmac, ops)
    out["parts"] = parts
    print(json.dumps(out))


if __name__ == "__main__":
    main()
