"""CPU-side checks: C-ABI exports, config surface, state_dict compatibility, registry drop-in,
structures, synthetic-data determinism.  No GPU compute."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import helpers
import drn_wsod_pytorch_b200 as drn
from drn_wsod_pytorch_b200 import lib, synth
from oracle import refstub

HAVE_REF = refstub.reference_available()


def test_cabi_library_exports_every_header_symbol():
    if not os.path.exists(lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    handle = ctypes.CDLL(lib.LIB_PATH)
    syms = lib.header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/drn_b200.h but not exported"
    assert set(lib._PROTOS) | {"drn_last_error", "drn_roipool_workspace_bytes", "drn_gemm_workspace_bytes", "drn_gemm_set_tail_split", "drn_gemm_set_max_sms",
                                 "drn_detections_workspace_bytes"} == set(syms), "lib.py prototypes out of sync with the header"
    handle.drn_version.restype = ctypes.c_int
    assert handle.drn_version() >= 100


def test_cabi_prototypes_have_the_header_argument_counts():
    """Every ctypes prototype in lib.py takes exactly as many arguments as the header declares for that entry point."""
    import re

    with open(lib.HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    decls = {m.group(1): m.group(2) for m in re.finditer(r"\b(drn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)}
    for name, argtypes in lib._PROTOS.items():
        assert name in decls, name
        params = decls[name].strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), f"{name}: header declares {n} parameters, lib.py binds {len(argtypes)}"


def test_cabi_stage_pair_rejects_bad_arguments_without_a_gpu():
    handle = lib.load()
    f3 = (ctypes.c_float * 4)(1, 1, 1, 1)
    i1 = (ctypes.c_int * 1)(0)
    args = [None, 128, 10, 20, 1, i1, i1, f3, None, None, 1, None, 0, None, None, f3, i1, 1, ctypes.c_float(1.0), None, None, -1, None,
            None, None, None, None, None, None, None, None, None, None, None, None, None, i1, None, None]
    assert handle.drn_oicr_stages_fwd(*args, 0, None) != 0 and b"phases" in handle.drn_last_error()
    assert handle.drn_oicr_stages_fwd(*args, 3, None) != 0 and b"null pointer" in handle.drn_last_error()


def test_cabi_argument_errors_are_reported_not_crashed():
    handle = lib.load()
    rc = handle.drn_roipool_fwd(None, 4, 4, 64, None, None, 3, ctypes.c_float(0.125), 0, None, None, 0, None)
    assert rc != 0 and b"null pointer" in handle.drn_last_error()
    with pytest.raises(RuntimeError, match="Cin"):
        lib.call("drn_conv_igemm_f32", 1, 1, 4, 4, 3, 1, 3, 1, None, None, None, 0, 1, 64, 64, None)


def test_product_path_refuses_cpu_tensors():
    from drn_wsod_pytorch_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.roipool(torch.zeros(4, 4, 64), torch.zeros(1, 4), None, 0.125)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(helpers.ROOT, "drn_wsod_pytorch_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), f"{fn} references oracle/"


@pytest.mark.parametrize("case", list(helpers.CASES))
def test_state_dict_keys_and_builtin_config(case):
    cfg = helpers.case_config(case)
    model = drn.build_model(cfg)
    keys = list(model.state_dict().keys())
    assert keys[0] == "pixel_mean" and any(k.startswith("roi_heads.box_head.fc1") for k in keys)
    trainable = {n for n, p in model.named_parameters() if p.requires_grad}
    assert all(n.startswith("roi_heads.") for n in trainable), "FREEZE_AT: 5 freezes the whole backbone"


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference")
def test_reference_yaml_loads_unchanged_and_matches_builtin():
    base = os.path.join(refstub.WSL_ROOT, "configs")
    for case, (name, yaml_rel, ov, _) in helpers.CASES.items():
        cfg = drn.get_cfg().merge_from_file(os.path.join(base, yaml_rel))
        ours = drn.builtin_config(name, ov)
        for path in ["MODEL.BACKBONE.NAME", "MODEL.BACKBONE.FREEZE_AT", "MODEL.RESNETS.DEPTH", "MODEL.RESNETS.RES5_DILATION",
                     "MODEL.RESNETS.RES2_OUT_CHANNELS", "MODEL.VGG.CONV5_DILATION", "MODEL.ROI_HEADS.NAME",
                     "MODEL.ROI_HEADS.NUM_CLASSES", "MODEL.ROI_HEADS.IN_FEATURES", "MODEL.ROI_HEADS.SCORE_THRESH_TEST",
                     "MODEL.ROI_HEADS.NMS_THRESH_TEST", "MODEL.ROI_HEADS.PROPOSAL_APPEND_GT", "MODEL.ROI_BOX_HEAD.DAN_DIM",
                     "MODEL.ROI_BOX_HEAD.POOLER_TYPE", "MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", "MODEL.PIXEL_MEAN",
                     "MODEL.LOAD_PROPOSALS", "WSL.REFINE_NUM", "WSL.REFINE_REG", "WSL.MEAN_LOSS", "MODEL.META_ARCHITECTURE"]:
            a, b = cfg, ours
            for part in path.split("."):
                a, b = a[part], b[part]
            assert list(a) == list(b) if isinstance(a, (list, tuple)) else a == b, (case, path, a, b)


_DROPIN = r"""
import sys
sys.path.insert(0, {root!r})
from oracle import refstub
refstub.install(with_wsl=False)                      # detectron2 only: the reference's wsl.modeling is NOT imported
import drn_wsod_pytorch_b200 as drn
drn.register_into_detectron2()
from detectron2.config import get_cfg
from detectron2.modeling import build_model
import importlib.util, os
spec = importlib.util.spec_from_file_location("wsl_cfg", os.path.join(refstub.WSL_ROOT, "wsl", "config", "defaults.py"))
m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
cfg = get_cfg(); m.add_wsl_config(cfg)
cfg.merge_from_file(os.path.join(refstub.WSL_ROOT, "configs", "PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml"))
cfg.MODEL.DEVICE = "cpu"
model = build_model(cfg)                               # the reference's own builder + unchanged YAML
assert type(model).__module__.startswith("drn_wsod_pytorch_b200"), type(model)
print("DROPIN_OK", type(model).__name__, len(model.state_dict()))
"""


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference")
def test_registry_drop_in_behind_reference_build_model():
    out = subprocess.run([sys.executable, "-c", _DROPIN.format(root=helpers.ROOT)], capture_output=True, text=True, timeout=300)
    assert "DROPIN_OK GeneralizedRCNNWSL 132" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference")
def test_state_dict_identical_to_reference_model():
    cfg_r, ref = refstub.build_reference_model("PascalVOC-Detection/oicr_WSR_50_DC5_1x.yaml", ["MODEL.ROI_BOX_HEAD.DAN_DIM", "[64,64]"])
    ours = drn.build_model(drn.builtin_config("oicr_WSR_50_DC5_1x", ["MODEL.DEVICE", "cpu", "MODEL.ROI_BOX_HEAD.DAN_DIM", [64, 64]]))
    a = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    b = [(k, tuple(v.shape)) for k, v in ours.state_dict().items()]
    assert a == b
    ours.load_state_dict(ref.state_dict(), strict=True)  # a reference checkpoint loads as-is


def test_structures_surface():
    b = drn.Boxes(torch.tensor([[0.0, 0, 10, 10], [5, 5, 50, 60]]))
    assert torch.equal(b.area(), torch.tensor([100.0, 2475.0]))
    b.clip((40, 30))
    assert b.tensor[1].tolist() == [5, 5, 30, 40]
    inst = drn.Instances((40, 30), proposal_boxes=b, objectness_logits=torch.tensor([0.1, 0.2]))
    assert len(inst) == 2 and inst.has("proposal_boxes") and len(inst[torch.tensor([True, False])]) == 1
    with pytest.raises(AssertionError):
        inst.gt_classes = torch.zeros(3)
    from drn_wsod_pytorch_b200.structures import detector_postprocess

    r = drn.Instances((40, 30), pred_boxes=drn.Boxes(torch.tensor([[0.0, 0, 30, 40], [3, 3, 3, 9]])), scores=torch.tensor([0.9, 0.8]))
    out = detector_postprocess(r, 80, 60)
    assert len(out) == 1 and out.pred_boxes.tensor[0].tolist() == [0, 0, 60, 80]


def test_synthetic_data_is_deterministic_and_well_formed():
    a, b = synth.make_inputs(600, 1000, 2000, seed=0), synth.make_inputs(600, 1000, 2000, seed=0)
    assert torch.equal(a["image"], b["image"]) and torch.equal(a["boxes"], b["boxes"])
    bx = a["boxes"]
    assert (bx[:, 2] - bx[:, 0]).min() >= 20 and (bx[:, 3] - bx[:, 1]).min() >= 20
    assert bx[:, 2].max() <= 1000 and bx[:, 3].max() <= 600 and bx.min() >= 0
    shapes = {"backbone.stem.conv1.weight": (64, 3, 3, 3), "backbone.stem.conv1.norm.weight": (64,),
              "roi_heads.box_head.fc1.weight": (16, 128), "roi_heads.box_head.fc1.bias": (16,)}
    w1, w2 = synth.make_weights(shapes, 3), synth.make_weights(shapes, 3)
    assert all(torch.equal(w1[k], w2[k]) for k in w1)
    assert synth.load_calib("resnet_ws18_d2"), "data/calib.json missing"


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, produced
    without a GPU, timing one FULL image of the workload (no sampling / extrapolation).  Here (with /root/reference) it runs the
    unmodified reference model; forced to the oracle port (what the GPU box runs) it must report the same losses."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = {}
    for kind in (("reference", "port") if refstub.reference_available() else ("port",)):
        env = dict(os.environ, DRN_BENCH_REFERENCE_KIND="port" if kind == "port" else "")
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "r18_bf16",
                              "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=root, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
            assert k in line, k
        assert line["impl"] == "reference" and line["unit"] == "images/sec" and line["value"] > 0 and line["vs_baseline"] is None
        assert line["gpu_launches"] == 0 and "workload" in line["config"] and "model" not in line["config"]
        cb = line["cpu_baseline"]
        assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == line["value"]
        assert "no extrapolation" in cb["sample"] and "all 2000 proposals" in cb["sample"]
        assert abs(line["steps"] * line["ms_per_step"] / 1e3 - sum(cb["step_seconds"])) < 0.01  # claimed time = measured time
        assert line["e2e"] == {"value": line["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        lines[kind] = line
    if len(lines) == 2:  # the port the GPU box times computes what the unmodified reference computes, at the full size
        for k, v in lines["reference"]["losses"].items():
            assert abs(lines["port"]["losses"][k] - v) <= 1e-4 * abs(v) + 1e-6, (k, lines["port"]["losses"][k], v)


def test_fp32_tc_operand_packing_reconstructs_the_layer_on_cpu():
    """Host logic of the "fp32_tc" precision (modeling._f32tc_pack / _kgroup), checked without a GPU: emulate the 1 + G
    GEMMs in float64 from the packed bf16 operands -- the small planes x1|x2|x1|x2|x3 against w2|w1|w3|w2|w1 plus the
    K-grouped leading product -- and compare with the fp32 layer (3x3 filter taps and a linear layer)."""
    import torch
    from drn_wsod_pytorch_b200 import modeling as M, ops

    assert [M._kgroup(3, c) for c in (64, 128, 512)] == [64, 64, 64]
    for c, want in ((64, 64), (512, 512), (1024, 1024), (2048, 1024), (4096, 1024), (25088, 896), (100352, 1792)):
        cg = M._kgroup(1, c)
        assert cg == want and c % cg == 0 and cg % 64 == 0 and c // cg <= 64, (c, cg)
    g = torch.Generator().manual_seed(5)
    for rows, taps, C, cout in ((7, 9, 128, 8), (5, 1, 2048, 16)):
        w = torch.randn(cout, taps, C, generator=g)
        b = torch.randn(cout, generator=g)
        cg = M._kgroup(3 if taps == 9 else 1, C)
        p = M._f32tc_pack(w, b, cg)
        G = C // cg
        assert p["groups"] == G and p["big"].shape == (G, cout, taps * cg) and p["small"].shape == (cout, taps * 5 * C)
        assert p["big"].dtype == torch.bfloat16 and p["small"].dtype == torch.bfloat16
        x = torch.randn(rows, taps, C, generator=g)  # im2col'd activation: [rows][tap][channel]
        t = M._bf16_terms(x)
        assert torch.equal((t[0].double() + t[1].double() + t[2].double()).float(), x)  # exact three-way split
        small_x = torch.stack([t[0], t[1], t[0], t[1], t[2]], dim=2).reshape(rows, taps * 5 * C)  # what drn_f32tc_split writes
        big_x = t[0].view(rows, taps, G, cg).permute(2, 0, 1, 3).reshape(G, rows, taps * cg)
        y = small_x.double() @ p["small"].double().t()
        for k in range(G):
            y = y + big_x[k].double() @ p["big"][k].double().t()
        y = y + p["bias"].double()
        ref = x.reshape(rows, -1).double() @ w.reshape(cout, -1).double().t() + b.double()
        mag = x.reshape(rows, -1).double().abs() @ w.reshape(cout, -1).double().abs().t()
        assert ((y - ref).abs() / mag).max().item() < 2e-7  # the nine dropped products are <= 2^-24 each
    assert ops.F32TC_SMALL_W == (1, 0, 2, 1, 0)


@pytest.mark.parametrize("clip", [None, ("value", 0.05, 2.0), ("norm", 0.3, 2.0)])
def test_fused_sgd_host_branch_matches_torch_sgd_with_gradient_clipping(clip):
    """solver.FusedSGD on CPU parameters (torch's own arithmetic) incl. the per-parameter clipping of
    detectron2/solver/build.py:19-92, and build_optimizer's param groups (BIAS_LR_FACTOR, WEIGHT_DECAY_BIAS)."""
    from drn_wsod_pytorch_b200.solver import FusedSGD, build_optimizer

    torch.manual_seed(0)
    a = [torch.nn.Parameter(torch.randn(5, 4)), torch.nn.Parameter(torch.randn(5))]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    oa = FusedSGD(a, lr=0.1, momentum=0.9, weight_decay=1e-3, clip=clip)
    ob = torch.optim.SGD(b, lr=0.1, momentum=0.9, weight_decay=1e-3)
    for it in range(3):
        for pa, pb in zip(a, b):
            g = torch.randn_like(pa)
            pa.grad, pb.grad = g.clone(), g.clone()
            if clip is not None:
                if clip[0] == "value":
                    torch.nn.utils.clip_grad_value_(pb, clip[1])
                else:
                    torch.nn.utils.clip_grad_norm_(pb, clip[1], clip[2])
        v0 = a[0]._version
        oa.step()
        ob.step()
        assert a[0]._version > v0  # caches keyed on the version see the update
        for pa, pb in zip(a, b):
            assert torch.equal(pa, pb)
    with pytest.raises(ValueError):
        FusedSGD(a, lr=0.1, clip=("median", 1.0, 2.0))
    cfg = drn.builtin_config("oicr_WSR_18_DC5_1x", ["MODEL.DEVICE", "cpu"])
    cfg.SOLVER.CLIP_GRADIENTS.ENABLED = True
    cfg.SOLVER.CLIP_GRADIENTS.CLIP_TYPE = "norm"
    opt = build_optimizer(cfg, torch.nn.Linear(3, 2))
    assert opt.clip == ("norm", 1.0, 2.0)
    lrs = sorted(g["lr"] for g in opt.param_groups)
    assert lrs == [cfg.SOLVER.BASE_LR, cfg.SOLVER.BASE_LR * cfg.SOLVER.BIAS_LR_FACTOR]


def test_run_step_iter_size_schedule_on_cpu():
    """solver.run_step = projects/WSL/tools/train_net.py:65-117: zero_grad once at start_iter, loss / ITER_SIZE, the optimizer
    steps (and zeroes) only on iterations that are a multiple of ITER_SIZE, non-finite losses raise FloatingPointError."""
    from drn_wsod_pytorch_b200.solver import run_step

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor([1.0, -2.0]))

        def forward(self, x):
            return {"loss_a": (self.w * x).sum(), "loss_b": (self.w ** 2).sum() * 0.5}

    m = Toy().train()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    m.w.grad = torch.tensor([9.0, 9.0])  # stale gradient: must be cleared at start_iter
    xs = [torch.tensor([1.0, 0.0]), torch.tensor([0.0, 2.0]), torch.tensor([1.0, 1.0])]
    w0 = m.w.detach().clone()
    out = run_step(m, opt, xs[0], 1, iter_size=3, start_iter=1)
    assert set(out) == {"loss_a", "loss_b"} and not out["loss_a"].requires_grad
    run_step(m, opt, xs[1], 2, iter_size=3, start_iter=1)
    assert torch.equal(m.w.detach(), w0)  # no step yet
    assert torch.allclose(m.w.grad, (xs[0] + xs[1] + 2 * w0) / 3)
    run_step(m, opt, xs[2], 3, iter_size=3, start_iter=1)
    assert torch.allclose(m.w.detach(), w0 - 0.1 * (sum(xs) + 3 * w0) / 3)
    assert m.w.grad is None or float(m.w.grad.abs().max()) == 0.0
    with pytest.raises(FloatingPointError):
        run_step(m, opt, torch.tensor([float("inf"), 0.0]), 4, iter_size=3, start_iter=1)
    m.eval()
    with pytest.raises(AssertionError):
        run_step(m, opt, xs[0], 5)
