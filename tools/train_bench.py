#!/usr/bin/env python
"""Training-step timing on the bench workload (GPU box only): forward+loss (captured plan), backward of the
trainable tail, torch.optim.SGD step -- ms per phase with CUDA events."""
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
    del weights
    model.train()
    cfg.SOLVER.BASE_LR = 1e-5
    if os.environ.get("TORCH_SGD") == "1":
        opt = torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=1e-5, momentum=0.9, weight_decay=1e-4)
    else:
        opt = drn.build_optimizer(cfg, model)
    batched = bench.make_batched(synth.make_inputs(H, W, R, seed=0), torch.device("cuda:0"), drn)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tot = {"fwd": 0.0, "bwd": 0.0, "opt": 0.0}
    n = 0
    for it in range(8):
        e = [ev() for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        e[0].record()
        losses = model(batched)
        loss = sum(losses.values())
        e[1].record()
        loss.backward()
        e[2].record()
        opt.step()
        e[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            tot["fwd"] += e[0].elapsed_time(e[1]); tot["bwd"] += e[1].elapsed_time(e[2]); tot["opt"] += e[2].elapsed_time(e[3])
            n += 1
        print(f"it {it}: loss {loss.item():.4f} fwd {e[0].elapsed_time(e[1]):.2f} bwd {e[1].elapsed_time(e[2]):.2f} opt {e[2].elapsed_time(e[3]):.2f} ms", flush=True)
    print({k: round(v / n, 3) for k, v in tot.items()}, "ms; total", round(sum(tot.values()) / n, 3), "ms/step;", "plans:", len(model._plans))
    print("max memory GB:", torch.cuda.max_memory_allocated() / 1e9)


if __name__ == "__main__":
    main()
