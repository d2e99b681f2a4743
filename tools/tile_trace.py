#!/usr/bin/env python
"""Per-tile pipeline trace of CTA 0 of single gemm_tc_kernel launches (GPU box only; run with DRN_TC_DEBUG=2048).

Records every C-ABI conv / GEMM call of one forward of a workload (like tools/layer_bench.py), then launches each distinct
layer signature once more with the trace armed and prints, per tile of CTA 0 and in SM clocks since kernel start: when the
producer started the tile and had issued its last load, when the MMA warp got the free accumulator, the first operand and had
issued the last commit, when epilogue warp 0 saw the full accumulator and had issued its last store.  Shows which hand-over a
layer's tiles wait on.  Not a bench value.

    DRN_TC_DEBUG=2048 python tools/tile_trace.py [--workload r50_bf16] [--only 64,64,3]   # Cin,Cout,ksize filters
"""
import argparse
import collections
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

COLS = ["prod_start", "prod_issued", "mma_acc_free", "mma_first_op", "mma_committed", "epi_acc_full", "epi_stored"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="r50_bf16")
    ap.add_argument("--only", default="", help="semicolon-separated Cin,Cout,ksize[,res] filters (default: all layers)")
    ap.add_argument("--tiles", type=int, default=12)
    args = ap.parse_args()
    assert int(os.environ.get("DRN_TC_DEBUG", "0")) & 2048, "run with DRN_TC_DEBUG=2048"
    import bench
    import helpers
    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import lib, ops, synth

    cfg_name, H, W, R, precision, _ = bench.WORKLOADS[args.workload]
    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cuda:0", "B200.PRECISION", precision])
    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    model.train()
    model.use_cuda_graph = False
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    calls = []
    orig = ops.conv_bf16_tc

    def rec(x, packed, ksize, dilation, relu, residual=None, **kw):
        calls.append((x, packed, ksize, dilation, relu, residual, kw))
        return orig(x, packed, ksize, dilation, relu, residual, **kw)

    ops.conv_bf16_tc = rec
    model(batched)
    calls.clear()
    model(batched)
    torch.cuda.synchronize()
    ops.conv_bf16_tc = orig
    filters = [tuple(int(v) for v in f.split(",")) for f in args.only.split(";") if f]
    agg = collections.OrderedDict()
    for (x, packed, ksize, dil, relu, res, kw) in calls:
        N, Hh, Ww, Cin = x.shape
        key = (N * Hh * Ww if ksize == 1 else (Hh, Ww), Cin, packed["cout"], ksize, dil, res is not None)
        sig = (Cin, packed["cout"], ksize, int(res is not None))
        if filters and not any(sig[: len(f)] == f for f in filters):
            continue
        agg.setdefault(key, (x, packed, ksize, dil, relu, res, kw))
    h = lib.load()
    h.drn_gemm_trace_reset.argtypes = [ctypes.c_void_p]
    h.drn_gemm_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ntile = 48
    buf = (ctypes.c_longlong * (8 * ntile))()
    for key, (x, packed, ksize, dil, relu, res, kw) in agg.items():
        for _ in range(3):
            orig(x, packed, ksize, dil, relu, res, **kw)
        torch.cuda.synchronize()
        assert h.drn_gemm_trace_reset(None) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(x, packed, ksize, dil, relu, res, **kw)
        e1.record()
        torch.cuda.synchronize()
        n = h.drn_gemm_trace_read(buf, ntile)
        t = [[buf[8 * i + j] for j in range(8)] for i in range(n)]
        t0 = t[0][7]
        print(f"== layer rows/HxW={key[0]} Cin={key[1]} Cout={key[2]} k={key[3]} dil={key[4]} res={int(key[5])}: "
              f"{e0.elapsed_time(e1) * 1e3:.1f} us for this single launch (cold pipeline, launch overhead included)")
        print(f"prologue done (barriers, TMEM, griddepcontrol.wait) at {t[1][7] - t0} clocks")
        print("tile " + " ".join(f"{c:>13}" for c in COLS) + "   (SM clocks since kernel start; ~1.9 clocks per ns)")
        for i in range(min(n, args.tiles)):
            if not any(t[i][:7]):
                break
            print(f"{i:4d} " + " ".join(f"{(v - t0) if v else -1:13d}" for v in t[i][:7]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
