#!/usr/bin/env python
"""Benchmark of the DRN-WSOD hot path (BASELINE.json metric: images/sec, WSOD forward+loss,
one 600x1000 synthetic image + 4 000 (2 000) random proposals per GPU, at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload r50_bf16|r18_fp32|...] [--impl ours|reference]

One step = one train-mode forward+loss of GeneralizedRCNNWSL over one image per GPU (dropout on, as in
the reference's train mode).  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# workload -> (builtin config, H, W, R, precision, GMAC per image {conv, fc6, fc7, heads} from BASELINE.md §3)
WORKLOADS = {
    "r18_fp32": ("oicr_WSR_18_DC5_1x", 600, 1000, 2000, "fp32", dict(conv=118.0, fc6=205.5, fc7=33.6, heads=0.85)),
    "r18_fp32_tc": ("oicr_WSR_18_DC5_1x", 600, 1000, 2000, "fp32_tc", dict(conv=118.0, fc6=205.5, fc7=33.6, heads=0.85)),
    "r18_bf16": ("oicr_WSR_18_DC5_1x", 600, 1000, 2000, "bf16", dict(conv=118.0, fc6=205.5, fc7=33.6, heads=0.85)),
    "r50_bf16": ("oicr_WSR_50_DC5_1x", 600, 1000, 4000, "bf16", dict(conv=232.7, fc6=822.1, fc7=33.6, heads=1.7)),
    "r50_fp32_tc": ("oicr_WSR_50_DC5_1x", 600, 1000, 4000, "fp32_tc", dict(conv=232.7, fc6=822.1, fc7=33.6, heads=1.7)),
    "r50_fp32": ("oicr_WSR_50_DC5_1x", 600, 1000, 4000, "fp32", dict(conv=232.7, fc6=822.1, fc7=33.6, heads=1.7)),
    "v16_bf16": ("oicr_V_16_DC5_1x", 600, 1000, 2000, "bf16", dict(conv=231.9, fc6=205.5, fc7=33.6, heads=0.85)),
    "r101_coco_bf16": ("oicr_WSR_101_DC5_1x_coco", 800, 1333, 4000, "bf16", dict(conv=723.5, fc6=822.1, fc7=33.6, heads=6.6)),
}
DEFAULT_WORKLOAD = os.environ.get("DRN_BENCH_WORKLOAD", "r50_bf16")  # BASELINE.json configs[2]: the 4k-proposal metric
METRIC = "images/sec (4k proposals/img) WSOD forward+loss"
# DRAM bytes of ONE fc6 launch from the ncu capture of a whole step (read 2587.9 MB + write 15.8 MB; algorithmic operand
# bytes 1.21 GB: the A operand is streamed once per wave of N tiles)
FC6_DRAM_BYTES = {"r50_bf16": 2587.9e6 + 15.8e6}
# kernels launched per C-ABI call (for the gpu_launches claim)
LAUNCHES = {"drn_wsddn_mil_fwd": 3, "drn_wsddn_mil_pgt_fwd": 2, "drn_label_proposals": 2, "drn_roipool_fwd": 2}  # (memset counted as a launch)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_batched(inp, device, drn, pinned=False):
    H, W = inp["height"], inp["width"]
    t = {k: (v.pin_memory() if pinned and torch.is_tensor(v) else v) for k, v in inp.items()}
    if device is not None:
        t = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in t.items()}
    p = drn.Instances((H, W), proposal_boxes=drn.Boxes(t["boxes"]), objectness_logits=t["objectness"])
    g = drn.Instances((H, W), gt_boxes=drn.Boxes(t["gt_boxes"]), gt_classes=t["gt_classes"])
    return [{"image": t["image"], "proposals": p, "instances": g, "height": H, "width": W}]


def cpu_reference_step(state, spec, inp, R_sample, threads):
    """One bounded sample of the reference CPU path (oracle port): full backbone on the full image,
    ROI stage on the first R_sample proposals; per-image time extrapolated linearly in R for the
    ROI stage (ROIPool + fc6/fc7 + heads are linear in R).  Returns (est. seconds per full image, parts)."""
    from oracle import wsl_oracle as O

    torch.set_num_threads(threads)
    R = inp["boxes"].shape[0]
    with torch.no_grad():
        t0 = time.perf_counter()
        x = O.preprocess_image(inp["image"], spec)
        fmap = O.backbone_forward(x, state, spec)
        t1 = time.perf_counter()
        sub = dict(inp, boxes=inp["boxes"][:R_sample], objectness=inp["objectness"][:R_sample])
        pooled = O.roi_pool(fmap, sub["boxes"], 1.0 / spec.stride) * (sub["objectness"] + 1).view(-1, 1, 1, 1)
        t2 = time.perf_counter()
        feat = O.dan_forward(pooled, state)
        scores = O.wsddn_scores(feat, state)
        for k in range(spec.refine_num):
            pre = f"roi_heads.box_refinery_{k}."
            torch.softmax(torch.nn.functional.linear(feat, state[pre + "cls_score.weight"], state[pre + "cls_score.bias"]), -1)
        t3 = time.perf_counter()
    tb, tp, th = t1 - t0, t2 - t1, t3 - t2
    est = tb + (tp + th) * (R / R_sample)
    return est, {"backbone_s": tb, "roipool_s": tp, "fc_heads_s": th, "R_sample": R_sample}


def _graph_ms(fn, reps=20):
    """Average duration of `fn`'s kernels captured as one CUDA graph and replayed back to back (warm)."""
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        fn()
        fn()
    cur.wait_stream(side)
    torch.cuda.synchronize()
    from drn_wsod_pytorch_b200 import ops as _ops
    _ops.drop_scratch(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del keep
    return e0.elapsed_time(e1) / reps


def measure_parts(model, batched0, H, W, R, gmac, ops):
    """Conv stack (whole backbone: first conv + every tcgen05 conv + max-pools) and ROIPool (tables + gather), each as
    its own graph: ms, achieved TFLOP/s over the conv GMACs, achieved GB/s over the ROIPool's algorithmic bytes
    (feature map once + rois + output)."""
    with torch.no_grad():
        img = batched0["image"].float().contiguous()
        boxes = batched0["proposals"].proposal_boxes.tensor.float().contiguous()
        obj = batched0["proposals"].objectness_logits.float().contiguous()
        bb_ms = _graph_ms(lambda: model._features([img], (H, W)))
        feats = model._features([img], (H, W))
        fh = model.roi_heads._features_hwc(feats, 0)
        scale = model.roi_heads.pooler_scale
        rp_ms = _graph_ms(lambda: ops.roipool(fh, boxes, obj, scale))
    h, w, C = fh.shape
    e = fh.element_size()
    rp_bytes = h * w * C * e + R * 20 + R * 49 * C * e
    return {"conv_stack_ms": bb_ms, "conv_stack_gmac": gmac["conv"], "conv_stack_tflops": 2 * gmac["conv"] * 1e9 / (bb_ms * 1e-3) / 1e12,
            "roipool_ms": rp_ms, "roipool_algorithmic_bytes": rp_bytes, "roipool_gbs": rp_bytes / (rp_ms * 1e-3) / 1e9,
            "how": "each part captured as its own CUDA graph, 20 back-to-back replays (warm), CUDA events"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="forward", choices=["forward", "train", "eval"],
                    help="forward: forward+loss (BASELINE.json's metric, default); train: forward+loss+backward of the trainable "
                         "tail+SGD step (+ gradient all-reduce at N>1) -- SURVEY.md §8f row 1; eval: inference forward + on-device "
                         "threshold/NMS/top-k (§8f row 2).  train / eval are extra lines, not the headline metric")
    ap.add_argument("--library-baseline", nargs="?", const="fp32", default=None, choices=["fp32", "bf16"],
                    help="N=1 only: also time the oracle port on torch's CUDA ops (cuDNN TF32 convolutions, cuBLAS fp32 GEMMs, "
                         "torchvision roi_pool) -- what the reference's own code runs when MODEL.DEVICE is a GPU (SURVEY.md 8d); "
                         "bf16: the same under torch.autocast(bfloat16), the closest library-only equivalent of this tree's bf16 mode; "
                         "adds `library_baseline` to the JSON line")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-kernel-family time breakdown to stderr")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name, H, W, R, precision, gmac = WORKLOADS[args.workload]
    config = {"workload": f"{cfg_name} {H}x{W} R={R} {precision} (BASELINE.json configs[{2 if 'r50' in args.workload else 1}])"
              if args.workload in ("r50_bf16", "r18_fp32") else f"{cfg_name} {H}x{W} R={R} {precision}",
              "images_per_gpu": 1, "proposals_per_image": R, "parallelism": f"dp{world}", "dropout": "on (train mode)",
              "launch": "one CUDA-graph replay per step (captured per input signature by the public forward)",
              "l2": "no flush: per-step working set (fc6 weights 411 MB + ROI features 0.2-0.8 GB) exceeds the 126 MB L2",
              "e2e_read": "loss vector D2H into pinned memory behind every step, read by the host one step later (event-synchronised)"}

    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers

    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cpu" if args.impl == "reference" else f"cuda:{local_rank}",
                                        "B200.PRECISION", precision])
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import wsl_oracle as O

        model = drn.build_model(cfg)  # parameter shapes only (CPU); arithmetic below is the oracle's
        state = dict(helpers.case_weights(cfg, model))
        spec = O.spec_from_cfg(cfg)
        inp = synth.make_inputs(H, W, R, seed=0)
        R_sample = min(R, 250)
        for _ in range(min(args.warmup, 1)):
            cpu_reference_step(state, spec, inp, min(R_sample, 50), threads)
        ests = [cpu_reference_step(state, spec, inp, R_sample, threads) for _ in range(max(1, min(args.steps, 3)))]
        est = sum(e for e, _ in ests) / len(ests)
        val = 1.0 / est
        sample = (f"full backbone on the {H}x{W} image + ROI stage on {R_sample} of {R} proposals, ROI-stage time "
                  f"scaled x{R / R_sample:.0f} (linear in R); {len(ests)} reps; torch {torch.__version__} CPU ops")
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "images/sec", "n_gpus": args.gpus,
                          "steps": len(ests), "warmup": min(args.warmup, 1), "ms_per_step": est * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "images/sec", "cores": threads, "kind": "port", "sample": sample,
                                           "parts": ests[-1][1]},
                          "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ------------------------------------------------------------------ our arm (B200)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist

        if args.mode == "train" and os.environ.get("DRN_B200_COMM_SMS", "0") != "0":
            # opt-in: bound the SMs NCCL takes while gradient all-reduces run next to the backward's GEMMs (whose grids
            # are then capped to the rest, modeling._wgrad / drn_gemm_set_max_sms)
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ["DRN_B200_COMM_SMS"])
        dist.init_process_group("nccl", device_id=dev)
    from drn_wsod_pytorch_b200 import lib as drn_lib, ops

    model = drn.build_model(cfg)
    weights = helpers.case_weights(cfg, model)
    model.load_state_dict({**weights, "pixel_mean": model.pixel_mean, "pixel_std": model.pixel_std}, strict=True)
    del weights
    model.train()
    eval_mode = args.mode == "eval"
    if eval_mode:
        model.eval()
        config["dropout"] = "off (eval mode)"
        config["mode"] = "eval: inference forward, mean of the refinement softmaxes, threshold 1e-5, per-class NMS 0.3, top-100 on the device"
    inp = synth.make_inputs(H, W, R, seed=rank)  # one distinct image per rank (weak scaling)
    batched_dev = make_batched(inp, dev, drn)
    host = make_batched(inp, None, drn, pinned=True)[0]
    loss_keys = None

    # count kernel launches through the C ABI (one eager step; the timed steps replay exactly these)
    counter = {"n": 0}
    orig_call = drn_lib.call

    def counting_call(name, *a):
        counter["n"] += LAUNCHES.get(name, 1)
        return orig_call(name, *a)

    from drn_wsod_pytorch_b200 import distributed as D

    train_mode = args.mode == "train"
    if train_mode:
        cfg.SOLVER.BASE_LR = 1e-6  # synthetic data: keep the random-init weights in a sane range over the timed steps
        optimizer = drn.build_optimizer(cfg, model)  # detectron2/solver/build.py mirrored, fused update kernel
        sync = D.GradientSynchronizer()
        if world > 1:
            model.roi_heads.grad_ready_hook = sync.ready

    def step(batched):
        if eval_mode:
            with torch.no_grad():
                out = model(batched)  # list of {"instances"}: scores / boxes / classes of the kept detections
            inst = out[0]["instances"]
            n = torch.tensor([float(len(inst))], device=dev)
            losses = {"num_detections": n[0], "top_score": inst.scores[0] if len(inst) else n[0] * 0}
        elif not train_mode:
            with torch.no_grad():  # the metric is forward + loss; the backward has its own line (--mode train)
                losses = model(batched)
        else:
            optimizer.zero_grad(set_to_none=True)
            losses = model(batched)
            sum(losses.values()).backward()  # gradient blocks are all-reduced (AVG) as they are produced
            sync.finish()
            optimizer.step()
            losses = {k: v.detach() for k, v in losses.items()}
        losses = D.reduce_dict(losses)  # one packed all-reduce (detectron2/utils/comm.py:234-263 equivalent); identity at N=1
        keys = sorted(losses)
        return torch.stack([losses[k] for k in keys]), keys

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- eager pass (CUDA graph off): launch count + per-launch duration of the dominant kernel (fc6 GEMM)
    # measured with CUDA events on the launching stream inside real steps of the same path
    graph_default = model.use_cuda_graph
    model.use_cuda_graph = False
    fc6_events = []
    fc6_K = 49 * model.roi_heads.in_channels
    orig_tc, orig_f32 = ops.conv_bf16_tc, ops.conv_f32
    record = {"on": False}

    def wrap(fn):
        def inner(x, packed, ksize, dilation, relu, *a, **k):
            if ksize == 1 and x.shape[-1] == fc6_K and record["on"]:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(x, packed, ksize, dilation, relu, *a, **k)
                e1.record()
                fc6_events.append((e0, e1))
                return out
            return fn(x, packed, ksize, dilation, relu, *a, **k)
        return inner

    ops.conv_bf16_tc, ops.conv_f32 = wrap(orig_tc), wrap(orig_f32)
    from drn_wsod_pytorch_b200 import modeling as _modeling
    orig_layer = _modeling._f32tc_layer

    def layer_wrap(x, pk, ksize, dilation, relu, *a, **k):  # fp32_tc: fc6 = 1 + G GEMM launches + the reduction, timed together
        if ksize == 1 and x.shape[-1] == fc6_K and record["on"]:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_layer(x, pk, ksize, dilation, relu, *a, **k)
            e1.record()
            fc6_events.append((e0, e1))
            return out
        return orig_layer(x, pk, ksize, dilation, relu, *a, **k)

    _modeling._f32tc_layer = layer_wrap
    # the roofline kernel is timed as ONE launch on its own (the default path runs it in row blocks that
    # overlap the ROI pooling of the next block on a second stream, which CUDA events cannot separate)
    overlap_default = model.roi_heads.overlap_pool
    model.roi_heads.overlap_pool = False
    for _ in range(3):
        step(batched_dev)
    sync_all()
    record["on"] = True
    eager_steps = max(3, min(args.steps, 10))
    for _ in range(eager_steps):
        vec, loss_keys = step(batched_dev)
    sync_all()
    record["on"] = False
    ops.conv_bf16_tc, ops.conv_f32 = orig_tc, orig_f32
    _modeling._f32tc_layer = orig_layer
    model.roi_heads.overlap_pool = overlap_default
    step(batched_dev)
    sync_all()
    drn_lib.call = counting_call
    ops.call = counting_call
    for _ in range(eager_steps):
        step(batched_dev)
    sync_all()
    drn_lib.call = orig_call
    ops.call = orig_call
    launches_per_step = counter["n"] // eager_steps
    fc6_ms = sum(a.elapsed_time(b) for a, b in fc6_events) / max(1, len(fc6_events))
    model.use_cuda_graph = graph_default

    if args.profile_step:
        for _ in range(3):
            step(batched_dev)
        sync_all()
        torch.cuda.cudart().cudaProfilerStart()
        step(batched_dev)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return

    # ---- timed region: the public forward (CUDA-graph plan per input signature), inputs resident in HBM
    for _ in range(max(args.warmup, 3)):
        vec, loss_keys = step(batched_dev)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        vec, _ = step(batched_dev)
    e1.record()
    sync_all()
    launches = launches_per_step * args.steps
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()

    # end-to-end: pinned host inputs -> H2D -> forward+loss -> D2H loss read, every step
    def e2e_step():
        # the user-facing call: pinned HOST tensors in (the model's plan copies them H2D into its static
        # buffers inside this call), loss vector read back to the host
        v, _ = step([host])
        return v.cpu()

    # input prefetch (model.prefetch): the H2D copy of step i+1's inputs is started right after step i is
    # launched, so it overlaps step i's kernels -- still one full H2D of every input per step, inside the timed region
    for _ in range(3):
        e2e_step()
    sync_all()
    # the loss vector of step i is copied to pinned host memory right behind step i on the stream and read by the host
    # one step later (after step i+1 has been launched), so the host-side launch work of the next step overlaps the
    # running one: every step's result is still read on the host inside the timed region
    pinned = [torch.empty(vec.shape, dtype=vec.dtype).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    def read(i):
        done[i & 1].synchronize()
        return pinned[i & 1].clone()

    t0 = time.perf_counter()
    batches = [[host], [host]]  # two batch objects over the same pinned tensors (prefetch is keyed on the batch object)
    model.prefetch(batches[0])
    for i in range(args.steps):
        v, _ = step(batches[i & 1])
        pinned[i & 1].copy_(v, non_blocking=True)
        done[i & 1].record()
        model.prefetch(batches[(i + 1) & 1])
        if i > 0:
            loss_host = read(i - 1)
    loss_host = read(args.steps - 1)
    sync_all()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(v.numel() * v.element_size() for v in (inp["image"], inp["boxes"], inp["objectness"], inp["gt_boxes"], inp["gt_classes"]))
    d2h = loss_host.numel() * loss_host.element_size()

    if args.breakdown and rank == 0:
        fams = {}

        def fam_wrap(name, fn):
            def inner(*a, **k):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); out = fn(*a, **k); a1.record()
                fams.setdefault(name, []).append((a0, a1))
                return out
            return inner

        saved = {n: getattr(ops, n) for n in ("first_conv", "conv_f32", "conv_bf16_tc", "maxpool2x2", "roipool", "wsddn_mil",
                                               "oicr_pgt", "label_proposals", "oicr_stage", "dropout_", "wsddn_mil_pgt", "oicr_stage_fused")}
        for n, f in saved.items():
            setattr(ops, n, fam_wrap(n, f))
        model.use_cuda_graph = False
        model.roi_heads.overlap_pool = False
        step(batched_dev)
        torch.cuda.synchronize()
        model.use_cuda_graph = graph_default
        model.roi_heads.overlap_pool = overlap_default
        for n, f in saved.items():
            setattr(ops, n, f)
        print("breakdown (ms, 1 step):", {n: round(sum(a.elapsed_time(b) for a, b in ev), 3) for n, ev in fams.items()},
              file=sys.stderr)

    # ---- parts (rank 0, secondary numbers): the conv stack and the ROIPool as their own CUDA graphs, replayed back to back
    # (warm), for the north_star's "fraction of the conv roofline" and the HBM-bound piece of SURVEY.md 8(d)
    parts = None
    if world == 1 and precision == "bf16":  # single-GPU runs only: nothing extra between the ranks' teardown at N > 1
        try:
            parts = measure_parts(model, batched_dev[0], H, W, R, gmac, ops)
        except Exception as e:  # never lose the bench line over a secondary measurement
            parts = {"error": f"{type(e).__name__}: {e}"[:200]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = _peaks()
    ms_step = ms_total / args.steps
    value = world * args.steps / (ms_total / 1e3)
    fc6_tflops = 2 * gmac["fc6"] * 1e9 / (fc6_ms * 1e-3) / 1e12 if fc6_ms > 0 else 0.0
    total_tflops = 2 * sum(gmac.values()) * 1e9 / (ms_step * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    out = {
        "metric": METRIC.replace("forward+loss", "inference (forward + NMS + top-k)") if eval_mode else METRIC if not train_mode
        else METRIC.replace("forward+loss", "training step (forward+loss+backward+SGD)"), "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic", "config": config,
        "e2e": {"value": world * args.steps / t_e2e.item(), "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "fc6 GEMM (gemm_tc_kernel)" if precision == "bf16" else "fc6 layer (1 + G gemm_tc_kernel launches over split-bf16 operands + fp32 reduction; fp32-equivalent FLOPs counted once, 6x that many run on the tensor cores)" if precision == "fp32_tc" else "fc6 GEMM (conv_igemm_f32_kernel, SIMT fp32)",
                     "achieved": fc6_tflops, "peak": peak, "unit": "TFLOP/s", "frac": fc6_tflops / peak,
                     "peak_source": f"{peaks['src']} cuBLAS bf16 sustained (kernel timed inside a long step)",
                     "traffic": FC6_DRAM_BYTES.get(args.workload), "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the fc6 launch, profiles/r1_ncu_step_v7_per_launch.txt (#62)" if args.workload in FC6_DRAM_BYTES else None,
                     "algorithmic_flops_per_launch": 2 * gmac["fc6"] * 1e9, "kernel_ms": fc6_ms, "share_of_step": fc6_ms / ms_step,
                     "whole_step_tflops": total_tflops, "whole_step_frac": total_tflops / peak},
        "losses": dict(zip(loss_keys, [round(float(x), 6) for x in vec.tolist()])),
    }
    if parts is not None:
        if "conv_stack_ms" in parts:
            parts["conv_stack_frac_of_bf16_peak"] = parts["conv_stack_tflops"] / peak
            parts["roipool_frac_of_hbm_peak"] = parts["roipool_gbs"] / peaks["hbm_gbs"]
        out["parts"] = parts
    if train_mode:
        out["config"]["mode"] = "train: backward of fc6/fc7/heads (backbone frozen, FREEZE_AT 5) + fused SGD; gradients averaged over ranks"
        out["grad_allreduce_bytes_per_step"] = sync.bytes // max(1, sync.steps)
    if world == 1 and not args.no_cpu_baseline:
        from oracle import wsl_oracle as O

        state = dict(helpers.case_weights(cfg, model))
        spec = O.spec_from_cfg(cfg)
        cinp = synth.make_inputs(H, W, R, seed=0)
        R_sample = min(R, 250)
        cpu_reference_step(state, spec, cinp, 50, threads)
        est, parts = cpu_reference_step(state, spec, cinp, R_sample, threads)
        out["cpu_baseline"] = {"value": 1.0 / est, "unit": "images/sec", "cores": threads, "kind": "port",
                               "sample": f"full backbone on the {H}x{W} image + ROI stage on {R_sample} of {R} proposals, ROI-stage time "
                                         f"scaled x{R / R_sample:.0f} (linear in R); fp32 torch CPU ops, {threads} threads", "parts": parts}
    if world == 1 and args.library_baseline:
        try:
            out["library_baseline"] = library_baseline(cfg, model, helpers, synth, H, W, R, dev, autocast=args.library_baseline == "bf16")
        except Exception as e:
            out["library_baseline"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def library_baseline(cfg, model, helpers, synth, H, W, R, dev, reps=5, autocast=False):
    """The baseline leg on the GPU: the oracle's restatement of the reference forward+loss (dropout off) executed by
    torch's own CUDA kernels in fp32, as the reference does on a GPU (no AMP in detectron2 v0.2; cuDNN may use TF32
    for the convolutions, matmuls stay fp32).  Baseline only -- never part of the product path."""
    from oracle import wsl_oracle as O

    state = {k: v.to(dev) for k, v in helpers.case_weights(cfg, model).items()}
    spec = O.spec_from_cfg(cfg)
    inp = synth.make_inputs(H, W, R, seed=0)
    b = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in inp.items()}
    orig_features = O.forward_features
    if autocast:  # backbone -> ROIPool -> fc6/fc7 under bf16 autocast; heads and losses stay fp32 (BCE refuses autocast)
        def features(*a, **k):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                t = orig_features(*a, **k)
            t["feat"] = t["feat"].float()
            return t

        O.forward_features = features
    try:
        return _library_baseline_run(O, b, state, spec, dev, reps, autocast)
    finally:
        O.forward_features = orig_features


def _library_baseline_run(O, b, state, spec, dev, reps, autocast):
    with torch.device(dev), torch.no_grad():
        for _ in range(2):
            O.forward_train([b], state, spec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            losses, _ = O.forward_train([b], state, spec)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"value": 1e3 / ms, "unit": "images/sec", "ms_per_step": ms, "kind": "oracle port on torch CUDA ops (cuDNN / cuBLAS / torchvision roi_pool), " + ("bf16 autocast" if autocast else "fp32") + ", dropout off",
            "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32), "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32),
            "losses": {k: round(float(v), 6) for k, v in losses.items()}}


if __name__ == "__main__":
    main()
