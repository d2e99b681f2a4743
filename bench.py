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
# the reference's own YAML (relative to projects/WSL/configs) of each builtin config, for the reference arm
REF_YAML = {"oicr_WSR_18_DC5_1x": "PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml", "oicr_WSR_50_DC5_1x": "PascalVOC-Detection/oicr_WSR_50_DC5_1x.yaml",
            "oicr_V_16_DC5_1x": "PascalVOC-Detection/oicr_V_16_DC5_1x.yaml", "oicr_WSR_101_DC5_1x_coco": "COCO-Detection/oicr_WSR_101_DC5_1x.yaml"}
BASELINE_CONFIG_INDEX = {"r18_fp32": 1, "r18_fp32_tc": 1, "r50_bf16": 2, "v16_bf16": 3, "r101_coco_bf16": 4}
# DRAM bytes of ONE launch of the roofline kernel: read from the ncu capture of the CURRENT tree that tools/ncu_traffic.py
# summarises into profiles/ (never a literal here); null when no capture of this workload is committed
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_fc6_traffic.json")
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
    """device given: image / proposals / objectness / GT boxes resident in HBM; the (two) GT class ids stay a host tensor, as a
    dataloader hands them over -- their count sizes kernel launches, so device-resident class ids cost a D2H sync per step."""
    H, W = inp["height"], inp["width"]
    t = {k: (v.pin_memory() if pinned and torch.is_tensor(v) else v) for k, v in inp.items()}
    if device is not None:
        t = {k: (v.to(device) if torch.is_tensor(v) and k != "gt_classes" else v) for k, v in t.items()}
    p = drn.Instances((H, W), proposal_boxes=drn.Boxes(t["boxes"]), objectness_logits=t["objectness"])
    g = drn.Instances((H, W), gt_boxes=drn.Boxes(t["gt_boxes"]), gt_classes=t["gt_classes"])
    return [{"image": t["image"], "proposals": p, "instances": g, "height": H, "width": W}]


def _roofline_traffic(workload):
    if not os.path.exists(TRAFFIC_FILE):
        return None, None
    with open(TRAFFIC_FILE) as f:
        t = json.load(f)
    e = t.get(workload)
    if not e:
        return None, None
    return e["dram_bytes_per_launch"], e.get("source")


class CpuReference:
    """The reference's CPU implementation of the path, timed on this box's host cores with every thread torch can use,
    on the FULL workload (whole image, all R proposals, every stage incl. pseudo-GT mining / labelling / weighted CE) --
    no sampling, no extrapolation.
    kind "reference": the UNMODIFIED reference model (projects/WSL GeneralizedRCNNWSL built by the reference's own
    build_model from its own YAML, imported under oracle/refstub.py's stub layer), `model(inputs)` in train mode inside an
    EventStorage -- used whenever /root/reference exists (this container).
    kind "port": oracle/wsl_oracle.forward_train, the restatement pinned to the reference's goldens -- the GPU box has no
    /root/reference.  Both run dropout-free (box_head.eval()), fp32, torch CPU ops + torchvision roi_pool."""

    def __init__(self, cfg_name, cfg, H, W, R, threads, num_classes):
        from drn_wsod_pytorch_b200 import synth
        from oracle import refstub

        torch.set_num_threads(threads)
        self.threads = threads
        self.inp = synth.make_inputs(H, W, R, seed=0, num_classes=num_classes)
        self.kind = "reference" if (refstub.reference_available() and cfg_name in REF_YAML and
                                    os.environ.get("DRN_BENCH_REFERENCE_KIND", "") != "port") else "port"
        if self.kind == "reference":
            _, self.model = refstub.build_reference_model(REF_YAML[cfg_name])
            sd = self.model.state_dict()
            w = dict(synth.calibrated_weights(cfg, {k: tuple(v.shape) for k, v in sd.items()}))
            w["pixel_mean"], w["pixel_std"] = sd["pixel_mean"], sd["pixel_std"]
            self.model.load_state_dict(w, strict=True)
            self.model.train()
            self.model.roi_heads.box_head.eval()
            from detectron2.structures import Boxes, Instances

            i = self.inp
            p = Instances((H, W))
            p.proposal_boxes, p.objectness_logits = Boxes(i["boxes"].clone()), i["objectness"].clone()
            g = Instances((H, W))
            g.gt_boxes, g.gt_classes = Boxes(i["gt_boxes"].clone()), i["gt_classes"].clone()
            self.batched = [{"image": i["image"], "height": H, "width": W, "proposals": p, "instances": g}]
        else:
            from oracle import wsl_oracle as O

            self.O, self.spec = O, O.spec_from_cfg(cfg)
            self.state = None  # set_state(): the same calibrated weights the B200 model runs

    def set_state(self, state):
        self.state = dict(state)

    def step(self):
        """One full image; returns (seconds, losses)."""
        t0 = time.perf_counter()
        with torch.no_grad():
            if self.kind == "reference":
                from detectron2.utils.events import EventStorage

                with EventStorage():
                    losses = self.model(self.batched)
            else:
                losses, _ = self.O.forward_train([self.inp], self.state, self.spec)
        dt = time.perf_counter() - t0
        return dt, {k: round(float(v), 6) for k, v in losses.items()}

    def describe(self, reps):
        what = ("the unmodified reference model (refstub import), model(inputs) in EventStorage" if self.kind == "reference"
                else "oracle/wsl_oracle.forward_train (restatement pinned to the reference's goldens; /root/reference is absent on this box)")
        i = self.inp
        return (f"{reps} full image(s): whole {i['height']}x{i['width']} image, all {i['boxes'].shape[0]} proposals, every stage, "
                f"no extrapolation; {what}; fp32 torch {torch.__version__} CPU ops, {self.threads} threads, dropout off")


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
    ap.add_argument("--library-baseline", nargs="?", const="both", default="both", choices=["both", "fp32", "bf16", "none"],
                    help="N=1 only (default: both): also time the oracle port on torch's CUDA ops (cuDNN TF32 convolutions, cuBLAS "
                         "fp32 GEMMs, torchvision roi_pool) -- what the reference's own code runs when MODEL.DEVICE is a GPU "
                         "(SURVEY.md 8d, BASELINE.md 4.5); bf16: the same under torch.autocast(bfloat16), the closest library-only "
                         "equivalent of this tree's bf16 mode; adds `library_baseline` {fp32, bf16} to the JSON line")
    ap.add_argument("--breakdown", action="store_true", help="also print a per-kernel-family time breakdown to stderr")
    ap.add_argument("--profile-step", action="store_true",
                    help="after warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`; prints no bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name, H, W, R, precision, gmac = WORKLOADS[args.workload]
    config = {"workload": f"{cfg_name} {H}x{W} R={R} {precision}" + (f" (BASELINE.json configs[{BASELINE_CONFIG_INDEX[args.workload]}])"
                                                                      if args.workload in BASELINE_CONFIG_INDEX else ""),
              "images_per_gpu": 1, "proposals_per_image": R, "parallelism": f"dp{world}", "dropout": "on (train mode)",
              "launch": "one CUDA-graph replay per step (captured per input signature by the public forward)",
              "l2": "no flush: per-step working set (fc6 weights 411 MB + ROI features 0.2-0.8 GB) exceeds the 126 MB L2",
              "e2e_read": "loss vector D2H into pinned memory behind every step, read by the host one step later (event-synchronised)"}

    import drn_wsod_pytorch_b200 as drn
    from drn_wsod_pytorch_b200 import synth

    cfg = drn.builtin_config(cfg_name, ["MODEL.DEVICE", "cpu" if args.impl == "reference" else f"cuda:{local_rank}",
                                        "B200.PRECISION", precision])
    threads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        ref = CpuReference(cfg_name, cfg, H, W, R, threads, cfg.MODEL.ROI_HEADS.NUM_CLASSES)
        if ref.kind == "port":
            ref.set_state(synth.calibrated_weights(cfg, drn.build_model(cfg)))  # parameter shapes only; arithmetic is the oracle's
        nwarm, nsteps = min(args.warmup, 1), max(1, min(args.steps, 2))
        for _ in range(nwarm):
            ref.step()
        times = []
        for _ in range(nsteps):
            dt, losses = ref.step()
            times.append(dt)
        est = sum(times) / len(times)
        val = 1.0 / est
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "images/sec", "n_gpus": args.gpus,
                          "steps": nsteps, "warmup": nwarm, "ms_per_step": est * 1e3, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": "images/sec", "cores": threads, "kind": ref.kind,
                                           "sample": ref.describe(nsteps), "step_seconds": [round(t, 3) for t in times]},
                          "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "losses": losses, "gpu_launches": 0}))
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
    weights = synth.calibrated_weights(cfg, model)
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
    reducer = D.LossReducer()
    if train_mode:
        cfg.SOLVER.BASE_LR = 1e-6  # synthetic data: keep the random-init weights in a sane range over the timed steps
        optimizer = drn.build_optimizer(cfg, model)  # detectron2/solver/build.py mirrored, fused update kernel
        sync = D.GradientSynchronizer().attach(model)  # gradients averaged across ranks; p.grad is written by sync.finish()
        sharder = None
        if world > 1 and precision == "bf16" and os.environ.get("DRN_B200_SHARDED", "1") != "0":
            # fc6.weight (95 % of the gradient bytes): reduce-scatter fused into the weight-gradient GEMM's epilogue (peer stores
            # over NVLink), sharded optimizer step, bf16 rows all-gathered by the update kernel; DRN_B200_SHARDED=0 = all-reduce
            fc6 = model.roi_heads.box_head.fc1
            sharder = D.ShardedLinearTrainer(fc6, precision, model.roi_heads.in_channels)
            model.roi_heads.fc6_sharder = sharder
            optimizer.attach_sharded(fc6.weight, sharder)

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
        # one packed all-reduce per step (detectron2/utils/comm.py:234-263 equivalent; identity at N=1), launched on a side
        # stream behind the step and handed out one step late, so the next step's kernels never wait for it
        losses = reducer.submit(losses)
        keys = sorted(losses)
        return torch.stack([losses[k] for k in keys]), keys

    def drain():
        """Join the last step's loss all-reduce (inside the timed region); returns its vector."""
        last = reducer.flush()
        return None if last is None else torch.stack([last[k] for k in sorted(last)])

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
    last = drain()
    vec = vec if last is None else last
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
    drain()
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
    traffic, traffic_src = _roofline_traffic(args.workload)
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
                     "traffic": traffic, "traffic_source": traffic_src,
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
        out["config"]["fc6_gradient"] = ("reduce-scatter fused into the weight-gradient GEMM epilogue (peer stores over NVLink) + sharded SGD + "
                                         "bf16 all-gather by the update kernel" if sharder is not None else "all-reduce (NCCL)" if world > 1 else "local")
        if sharder is not None:
            out["fc6_scatter_bytes_per_step"] = sharder.bytes_scattered // max(1, sync.steps)
    if world == 1 and not args.no_cpu_baseline:
        # one full image on the host cores (the bounded sample: ~20-50 s of CPU work), after a small warm-up that only
        # spins up the thread pool; the driver's reference arm (--impl reference) repeats it with a full warm-up step
        ref = CpuReference(cfg_name, cfg, H, W, R, threads, cfg.MODEL.ROI_HEADS.NUM_CLASSES)
        if ref.kind == "port":
            ref.set_state(synth.calibrated_weights(cfg, model))
        torch.nn.functional.conv2d(torch.zeros(1, 8, 64, 64), torch.zeros(8, 8, 3, 3))
        dt, closs = ref.step()
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "images/sec", "cores": threads, "kind": ref.kind,
                               "sample": ref.describe(1), "losses": closs}
        del ref
    if world == 1 and args.library_baseline != "none" and args.mode == "forward":
        out["library_baseline"] = {}
        for which in (("fp32", "bf16") if args.library_baseline == "both" else (args.library_baseline,)):
            try:
                out["library_baseline"][which] = library_baseline(cfg, model, synth, H, W, R, dev, autocast=which == "bf16")
            except Exception as e:
                out["library_baseline"][which] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def library_baseline(cfg, model, synth, H, W, R, dev, reps=5, autocast=False):
    """The baseline leg on the GPU: the oracle's restatement of the reference forward+loss (dropout off) executed by
    torch's own CUDA kernels in fp32, as the reference does on a GPU (no AMP in detectron2 v0.2; cuDNN may use TF32
    for the convolutions, matmuls stay fp32).  Baseline only -- never part of the product path."""
    from oracle import wsl_oracle as O

    state = {k: v.to(dev) for k, v in synth.calibrated_weights(cfg, model).items()}
    spec = O.spec_from_cfg(cfg)
    inp = synth.make_inputs(H, W, R, seed=0, num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES)
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
