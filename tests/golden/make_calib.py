"""Measure the per-layer output scales that drn_wsod_pytorch_b200/synth.py:make_weights applies
('calibrate once on the synthetic image').  Runs the CPU oracle layer by layer on a
small synthetic image; each conv's / fc's pre-activation std is measured and the layer is rescaled
in place to unit std before the pass continues.  Output: drn_wsod_pytorch_b200/data/calib.json.

    python tests/golden/make_calib.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import drn_wsod_pytorch_b200 as drn  # noqa: E402
from drn_wsod_pytorch_b200 import synth  # noqa: E402
from oracle import wsl_oracle as O  # noqa: E402

ARCHS = {
    "resnet_ws18_d2": "oicr_WSR_18_DC5_1x",
    "resnet_ws50_d2": "oicr_WSR_50_DC5_1x",
    "resnet_ws101_d2": "oicr_WSR_101_DC5_1x_coco",
    "vgg16_d1": "wsddn_V_16_DC5_1x",
    "vgg16_d2": "oicr_V_16_DC5_1x",
}


def arch_key(cfg):
    m = cfg.MODEL
    if "vgg" in m.BACKBONE.NAME:
        return f"vgg16_d{m.VGG.CONV5_DILATION}"
    return f"resnet_ws{m.RESNETS.DEPTH}_d{m.RESNETS.RES5_DILATION}"


def calibrate(name):
    cfg = drn.builtin_config(name, ["MODEL.DEVICE", "cpu"])
    model = drn.build_model(cfg)
    spec = O.spec_from_cfg(cfg)
    state = synth.make_weights(synth.state_shapes(model), seed=0, calib={})
    inp = synth.make_inputs(192, 256, 96, seed=7)
    calib = {}

    def tap(prefix, y):
        s = 1.0 / float(y.std())
        calib[prefix] = s
        if prefix + ".norm.weight" in state:
            state[prefix + ".norm.weight"].mul_(s)
        else:  # VGG: scale the filter, fix up the already-computed output exactly
            state[prefix + ".weight"].mul_(s)
            b = state[prefix + ".bias"].view(1, -1, 1, 1)
            y.sub_(b).mul_(s).add_(b)

    with torch.no_grad():
        x = O.preprocess_image(inp["image"], spec)
        fmap = O.backbone_forward(x, state, spec, tap=tap)
        pooled = O.roi_pool(fmap, inp["boxes"], 1.0 / spec.stride) * (inp["objectness"] + 1).view(-1, 1, 1, 1)
        h = torch.flatten(pooled, 1)
        for fc in ("fc1", "fc2"):
            p = f"roi_heads.box_head.{fc}"
            y = torch.nn.functional.linear(h, state[p + ".weight"], None)
            s = 1.0 / float(y.std())
            calib[p] = s
            state[p + ".weight"].mul_(s)
            h = torch.relu(y * s + state[p + ".bias"])
    return calib


if __name__ == "__main__":
    out = {}
    for key, name in ARCHS.items():
        out[key] = calibrate(name)
        print(key, len(out[key]), "layers; first/last scale", list(out[key].values())[0], list(out[key].values())[-1])
    path = os.path.join(ROOT, "drn_wsod_pytorch_b200", "data", "calib.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path)
