"""Seeded synthetic inputs and weights.

There is no network for datasets or checkpoints, so benches and parity tests run on synthetic
600x1000-style images, random >=20 px proposal boxes and random-init weights.  Everything is
drawn from numpy's PCG64 (bit-reproducible across machines), never from torch's RNG, so the GPU
box regenerates exactly the tensors the golden vectors were made from.

The reference's default init saturates the MIL scores on 0..255 inputs; weights
here are kaiming/normal draws times per-layer scalars from `data/calib.json` (measured once on a
synthetic image by tests/golden/make_calib.py) so every layer's output is O(1), the softmaxes are
neither flat nor one-hot and per-class argmax margins are far above fp32 noise.
"""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

_CALIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "calib.json")


def load_calib(arch_key):
    if not os.path.exists(_CALIB_PATH):
        return {}
    with open(_CALIB_PATH) as f:
        return json.load(f).get(arch_key, {})


def make_inputs(H, W, R, seed=0, num_gt=2, num_classes=20):
    """One image's inputs: dict(image 3xHxW fp32 BGR 0..255, boxes Rx4 XYXY, objectness R,
    gt_boxes Gx4, gt_classes G int64)."""
    assert H >= 64 and W >= 64, "synthetic proposals need a >=64 px image"
    rng = np.random.Generator(np.random.PCG64(seed))
    image = (rng.random((3, H, W), dtype=np.float32) * np.float32(255.0)).astype(np.float32)
    b = rng.random((R, 4), dtype=np.float32)
    x0 = b[:, 0] * np.float32(W - 40)
    y0 = b[:, 1] * np.float32(H - 40)
    w = np.float32(20) + b[:, 2] * (np.float32(W) - x0 - np.float32(21))
    h = np.float32(20) + b[:, 3] * (np.float32(H) - y0 - np.float32(21))
    boxes = np.stack([x0, y0, x0 + w, y0 + h], axis=1).astype(np.float32)
    objectness = rng.random(R, dtype=np.float32)
    base = np.array([[10, 10, 100, 100], [50, 60, 200, 220], [120, 30, 260, 140], [15, 90, 90, 250], [200, 200, 330, 300]], dtype=np.float32)
    classes = np.array([3, 7, 11, 3, 15], dtype=np.int64) % num_classes
    G = num_gt
    gtb = base[np.arange(G) % len(base)].copy()
    gtb[:, 0::2] *= np.float32(min(1.0, W / 400.0))
    gtb[:, 1::2] *= np.float32(min(1.0, H / 400.0))
    return {
        "image": torch.from_numpy(image),
        "boxes": torch.from_numpy(boxes),
        "objectness": torch.from_numpy(objectness),
        "gt_boxes": torch.from_numpy(gtb),
        "gt_classes": torch.from_numpy(classes[np.arange(G) % len(classes)].copy()),
        "height": H,
        "width": W,
    }


def _normal(rng, shape, std):
    return (rng.standard_normal(size=shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def make_weights(shapes, seed=0, calib=None):
    """shapes: ordered mapping state_dict key -> shape.  Returns OrderedDict key -> fp32 tensor.
    `calib` maps a layer prefix (e.g. 'backbone.res2.0.conv1') to the scalar applied to that
    layer's output (through the FrozenBN weight, or the conv/fc weight when there is no norm)."""
    calib = calib or {}
    rng = np.random.Generator(np.random.PCG64(seed + 1000003))
    norm_prefixes = {k[: -len(".norm.weight")] for k in shapes if k.endswith(".norm.weight")}
    out = OrderedDict()
    for key, shape in shapes.items():
        shape = tuple(int(s) for s in shape)
        prefix = key.rsplit(".", 1)[0]
        if key in ("pixel_mean", "pixel_std"):
            continue
        if key.endswith(".norm.weight"):
            s = calib.get(prefix[: -len(".norm")], 1.0)
            v = (rng.random(shape, dtype=np.float32) + np.float32(0.5)) * np.float32(s)
        elif key.endswith(".norm.bias"):
            v = _normal(rng, shape, 0.1)
        elif key.endswith(".norm.running_mean"):
            v = _normal(rng, shape, 0.1)
        elif key.endswith(".norm.running_var"):
            v = rng.random(shape, dtype=np.float32) + np.float32(0.5)
        elif key.endswith(".weight") and len(shape) == 4:
            fan_out = shape[0] * shape[2] * shape[3]
            s = 1.0 if prefix in norm_prefixes else calib.get(prefix, 1.0)
            v = _normal(rng, shape, np.sqrt(2.0 / fan_out) * s)
        elif key.endswith(".weight") and len(shape) == 2:
            if ".box_head." in key:
                v = _normal(rng, shape, calib.get(prefix, 1.0) / np.sqrt(shape[1]))
            elif "bbox_pred" in key:
                v = _normal(rng, shape, 0.5 / np.sqrt(shape[1]))
            else:  # cls / det / cls_score: spread logits so softmaxes are neither flat nor one-hot
                v = _normal(rng, shape, 4.0 / np.sqrt(shape[1]))
        elif key.endswith(".bias"):
            if ".box_head." in key:
                v = np.full(shape, 0.1, dtype=np.float32)
            elif "backbone." in key:
                v = _normal(rng, shape, 0.1)
            else:
                v = np.zeros(shape, dtype=np.float32)
        else:
            raise KeyError(f"make_weights: don't know how to initialise '{key}' {shape}")
        out[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out


def arch_key(cfg):
    """Key of the calibration table (data/calib.json) for a config's backbone."""
    m = cfg.MODEL
    if "vgg" in m.BACKBONE.NAME:
        return f"vgg16_d{m.VGG.CONV5_DILATION}"
    return f"resnet_ws{m.RESNETS.DEPTH}_d{m.RESNETS.RES5_DILATION}"


def calibrated_weights(cfg, model_or_shapes, seed=0):
    """The seeded, calibrated state_dict every bench and parity case uses for `cfg`'s architecture."""
    shapes = model_or_shapes if isinstance(model_or_shapes, dict) else state_shapes(model_or_shapes)
    return make_weights(shapes, seed=seed, calib=load_calib(arch_key(cfg)))


def state_shapes(model):
    return OrderedDict((k, tuple(v.shape)) for k, v in model.state_dict().items())


def weights_checksum(weights):
    """Order-dependent checksum used by the golden fixtures to prove identical regeneration."""
    acc = 0.0
    for i, (k, v) in enumerate(weights.items()):
        a = v.numpy().astype(np.float64).ravel()
        acc += float(a.sum()) * (1 + (i % 7)) + float(np.abs(a[:: max(1, a.size // 997)]).sum())
    return acc
