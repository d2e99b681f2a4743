#!/usr/bin/env python
"""Can the ROIPool gather and the fc6 GEMM share the SMs?  Times each alone and both on two streams (GPU box only)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from drn_wsod_pytorch_b200 import ops
    from drn_wsod_pytorch_b200.modeling import pack_linear
    from drn_wsod_pytorch_b200 import synth

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    h, w, C, R, N = 74, 124, 2048, 2048, 2048
    feat = torch.randn(h, w, C, device=dev).bfloat16()
    inp = synth.make_inputs(600, 1000, R, seed=0)
    boxes, obj = inp["boxes"].to(dev), inp["objectness"].to(dev)
    K = 49 * C
    wt = (torch.randn(N, K, device=dev) * 0.01)
    packed = pack_linear([wt], [torch.zeros(N, device=dev)], "bf16")
    del wt
    pooled = torch.empty(R, K, device=dev, dtype=torch.bfloat16)
    x = torch.randn(R, K, device=dev).bfloat16()
    y = torch.empty(R, N, device=dev, dtype=torch.bfloat16)
    tables = ops.roipool_tables(feat)
    sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
    torch.cuda.synchronize()

    def gather(k):
        ops.roipool_rows(feat, boxes, obj, 0.125, tables, pooled, max_ctas=148 * k)

    def gemm():
        ops.conv_bf16_tc(x.view(1, R, 1, K), packed, 1, 1, True, out=y.view(1, R, 1, N))

    def timed(fa, fb, reps=5):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cur = torch.cuda.current_stream()
            e0.record(cur)
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            if fa is not None:
                with torch.cuda.stream(sa):
                    fa()
            if fb is not None:
                with torch.cuda.stream(sb):
                    fb()
            cur.wait_stream(sa)
            cur.wait_stream(sb)
            e1.record(cur)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    gemm(); gather(0); gather(1)
    print(f"gemm alone: {timed(None, gemm):.3f} ms")
    for k in (0, 1, 2, 3, 4):
        g = timed(lambda: gather(k), None)
        both = timed(lambda: gather(k), gemm)
        both_rev = timed(gemm, lambda: gather(k))
        print(f"gather k={k}: alone {g:.3f} ms; gather then gemm launched: {both:.3f} ms; gemm then gather launched: {both_rev:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
