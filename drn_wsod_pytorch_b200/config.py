"""Config surface of the hot path.

Mirrors the keys the reference reads for this path: detectron2/config/defaults.py
and projects/WSL/wsl/config/defaults.py:14-50.  The reference's own YAML files load unchanged
through `CfgNode.merge_from_file` (keys outside the hot path are accepted and carried along);
`builtin_config(name)` provides the BASELINE.json configurations without the reference tree
(the GPU box has no /root/reference).

One extra key: `B200.PRECISION` in {"fp32", "bf16"} selects exact-fp32 SIMT arithmetic or the
bf16 tcgen05 tensor-core path (fp32 accumulate).  Overridable by the DRN_B200_PRECISION env var.
"""
import copy
import os

import yaml


class CfgNode(dict):
    """Attribute-style nested dict with yacs-like merge semantics (list<->tuple coercion)."""

    def __init__(self, init=None):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def freeze(self):
        return self

    def defrost(self):
        return self

    @staticmethod
    def _coerce(new, old):
        if isinstance(old, tuple) and isinstance(new, list):
            return tuple(new)
        if isinstance(old, list) and isinstance(new, tuple):
            return list(new)
        if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
            return float(new)
        return new

    def merge_from_other_cfg(self, other):
        for k, v in other.items():
            if isinstance(v, dict):
                if k not in self or not isinstance(self[k], CfgNode):
                    self[k] = CfgNode()
                self[k].merge_from_other_cfg(v)
            else:
                self[k] = self._coerce(v, self[k]) if k in self else v
        return self

    def merge_from_list(self, kv):
        assert len(kv) % 2 == 0, "merge_from_list expects KEY VALUE pairs"
        for key, v in zip(kv[0::2], kv[1::2]):
            node = self
            parts = key.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except yaml.YAMLError:
                    pass
            node[parts[-1]] = self._coerce(v, node.get(parts[-1], v))
        return self

    def merge_from_file(self, filename):
        self.merge_from_other_cfg(CfgNode(_load_yaml_with_base(filename)))
        return self


class _Loader(yaml.SafeLoader):
    pass


_Loader.add_constructor(
    "tag:yaml.org,2002:python/tuple", lambda loader, node: tuple(loader.construct_sequence(node))
)


def _load_yaml_with_base(filename):
    with open(filename) as f:
        cfg = yaml.load(f, Loader=_Loader) or {}
    base = cfg.pop("_BASE_", None)
    if base is None:
        return cfg
    if not os.path.isabs(base):
        base = os.path.join(os.path.dirname(filename), base)
    out = _load_yaml_with_base(base)

    def merge(a, b):
        for k, v in a.items():
            if isinstance(v, dict) and isinstance(b.get(k), dict):
                merge(v, b[k])
            else:
                b[k] = v

    merge(cfg, out)
    return out


def get_cfg():
    """Defaults for every key the hot path reads (values = the reference's defaults)."""
    return CfgNode(
        {
            "VERSION": 2,
            "MODEL": {
                "META_ARCHITECTURE": "GeneralizedRCNNWSL",
                "DEVICE": "cuda",
                "LOAD_PROPOSALS": False,
                "MASK_ON": False,
                "KEYPOINT_ON": False,
                "WEIGHTS": "",
                "PIXEL_MEAN": [103.530, 116.280, 123.675],
                "PIXEL_STD": [1.0, 1.0, 1.0],
                "BACKBONE": {"NAME": "build_ws_resnet_backbone", "FREEZE_AT": 2},
                "RESNETS": {
                    "DEPTH": 50,
                    "OUT_FEATURES": ["res4"],
                    "NUM_GROUPS": 1,
                    "NORM": "FrozenBN",
                    "WIDTH_PER_GROUP": 64,
                    "STRIDE_IN_1X1": True,
                    "RES5_DILATION": 1,
                    "RES2_OUT_CHANNELS": 256,
                    "STEM_OUT_CHANNELS": 64,
                    "DEFORM_ON_PER_STAGE": [False, False, False, False],
                    "DEFORM_MODULATED": False,
                    "DEFORM_NUM_GROUPS": 1,
                },
                "VGG": {"DEPTH": 16, "OUT_FEATURES": ["plain5"], "CONV5_DILATION": 1},
                "ROI_HEADS": {
                    "NAME": "WSDDNROIHeads",
                    "NUM_CLASSES": 80,
                    "IN_FEATURES": ["res4"],
                    "IOU_THRESHOLDS": [0.5],
                    "IOU_LABELS": [0, 1],
                    "BATCH_SIZE_PER_IMAGE": 512,
                    "POSITIVE_FRACTION": 0.25,
                    "SCORE_THRESH_TEST": 0.05,
                    "NMS_THRESH_TEST": 0.5,
                    "PROPOSAL_APPEND_GT": True,
                },
                "ROI_BOX_HEAD": {
                    "NAME": "DiscriminativeAdaptionNeck",
                    "BBOX_REG_LOSS_TYPE": "smooth_l1",
                    "BBOX_REG_LOSS_WEIGHT": 1.0,
                    "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0),
                    "SMOOTH_L1_BETA": 0.0,
                    "POOLER_RESOLUTION": 14,
                    "POOLER_SAMPLING_RATIO": 0,
                    "POOLER_TYPE": "ROIAlignV2",
                    "NUM_FC": 0,
                    "FC_DIM": 1024,
                    "NUM_CONV": 0,
                    "CONV_DIM": 256,
                    "NORM": "",
                    "CLS_AGNOSTIC_BBOX_REG": False,
                    "TRAIN_ON_PRED_BOXES": False,
                    "DAN_DIM": [4096, 4096],
                },
            },
            "INPUT": {"FORMAT": "BGR"},
            # detectron2/config/defaults.py:97-104, :569-583 (read by the TTA driver, tta.py)
            "DATASETS": {"PRECOMPUTED_PROPOSAL_TOPK_TRAIN": 2000, "PRECOMPUTED_PROPOSAL_TOPK_TEST": 1000},
            "TEST": {"DETECTIONS_PER_IMAGE": 100,
                     "AUG": {"ENABLED": False, "MIN_SIZES": [400, 500, 600, 700, 800, 900, 1000, 1100, 1200], "MAX_SIZE": 4000,
                             "FLIP": True}},
            # detectron2/config/defaults.py SOLVER keys read by build_optimizer (solver/build.py:93-137)
            "SOLVER": {"BASE_LR": 0.001, "MOMENTUM": 0.9, "NESTEROV": False, "WEIGHT_DECAY": 0.0001, "WEIGHT_DECAY_NORM": 0.0,
                       "BIAS_LR_FACTOR": 1.0, "WEIGHT_DECAY_BIAS": 0.0001, "IMS_PER_BATCH": 16,
                       # detectron2/config/defaults.py:549-559 (solver/build.py:19-92 maybe_add_gradient_clipping)
                       "CLIP_GRADIENTS": {"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0}},
            "WSL": {
                "VIS_TEST": False,
                "ITER_SIZE": 1,
                "MEAN_LOSS": True,
                "REFINE_NUM": 3,
                "REFINE_REG": [False, False, False],
            },
            "B200": {"PRECISION": "fp32"},
            "VIS_PERIOD": 0,
            "OUTPUT_DIR": "./output",
        }
    )


# what projects/WSL/configs/Base-RCNN-DilatedC5.yaml + the per-model YAMLs set, restated
_VOC_BASE = {
    "MODEL": {
        "META_ARCHITECTURE": "GeneralizedRCNNWSL",
        "LOAD_PROPOSALS": True,
        "MASK_ON": False,
        "BACKBONE": {"FREEZE_AT": 5},
        "RESNETS": {"OUT_FEATURES": ["res5"], "RES5_DILATION": 2},
        "ROI_HEADS": {
            "NAME": "WSDDNROIHeads",
            "IN_FEATURES": ["res5"],
            "BATCH_SIZE_PER_IMAGE": 4096,
            "POSITIVE_FRACTION": 1.0,
            "NUM_CLASSES": 20,
            "SCORE_THRESH_TEST": 0.00001,
            "NMS_THRESH_TEST": 0.3,
            "PROPOSAL_APPEND_GT": False,
        },
        "ROI_BOX_HEAD": {
            "NAME": "DiscriminativeAdaptionNeck",
            "NUM_FC": 2,
            "DAN_DIM": [4096, 4096],
            "POOLER_RESOLUTION": 7,
            "POOLER_TYPE": "ROIPool",
            "NUM_CONV": 0,
        },
    },
    "WSL": {"ITER_SIZE": 1, "MEAN_LOSS": True},
    # projects/WSL/configs/PascalVOC-Detection/Base-RCNN-DilatedC5.yaml:2-8, oicr_WSR_18_DC5_1x.yaml:49-53
    "DATASETS": {"PRECOMPUTED_PROPOSAL_TOPK_TRAIN": 4000, "PRECOMPUTED_PROPOSAL_TOPK_TEST": 4000},
    "TEST": {"AUG": {"ENABLED": True, "MIN_SIZES": [480, 576, 672, 768, 864, 960, 1056, 1152], "MAX_SIZE": 4000, "FLIP": True}},
    # projects/WSL/configs/PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml:40-48 (the same block in every WSL model YAML)
    "SOLVER": {"IMS_PER_BATCH": 4, "BASE_LR": 0.01, "WEIGHT_DECAY": 0.0005, "BIAS_LR_FACTOR": 2.0, "WEIGHT_DECAY_BIAS": 0.0},
}


def _wsr(depth, heads="OICRROIHeads", dan=(4096, 4096), classes=20, res2=None):
    return {
        "MODEL": {
            "PIXEL_MEAN": [102.9801, 115.9465, 122.7717],
            "BACKBONE": {"NAME": "build_ws_resnet_backbone"},
            "RESNETS": {
                "DEPTH": depth,
                "RES2_OUT_CHANNELS": res2 if res2 is not None else (64 if depth in (18, 34) else 256),
            },
            "ROI_HEADS": {"NAME": heads, "NUM_CLASSES": classes},
            "ROI_BOX_HEAD": {"DAN_DIM": list(dan)},
        }
    }


def _vgg(heads, dilation, score_thresh, nms):
    return {
        "MODEL": {
            "PIXEL_MEAN": [103.939, 116.779, 123.68],
            "BACKBONE": {"NAME": "build_vgg_backbone"},
            "VGG": {"DEPTH": 16, "CONV5_DILATION": dilation},
            "ROI_HEADS": {
                "NAME": heads,
                "IN_FEATURES": ["plain5"],
                "SCORE_THRESH_TEST": score_thresh,
                "NMS_THRESH_TEST": nms,
            },
        }
    }


BUILTIN = {
    # BASELINE.json configs[0..4] (reference YAML each one restates, under projects/WSL/configs/)
    "wsddn_V_16_DC5_1x": _vgg("WSDDNROIHeads", 1, 1e-9, 0.5),  # PascalVOC-Detection/wsddn_V_16_DC5_1x.yaml
    "oicr_WSR_18_DC5_1x": _wsr(18),  # PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml
    "oicr_WSR_50_DC5_1x": _wsr(50, dan=(2048, 4096)),  # PascalVOC-Detection/oicr_WSR_50_DC5_1x.yaml
    "oicr_V_16_DC5_1x": _vgg("OICRROIHeads", 2, 1e-5, 0.3),  # PascalVOC-Detection/oicr_V_16_DC5_1x.yaml
    "oicr_WSR_101_DC5_1x_coco": _wsr(101, dan=(2048, 4096), classes=80),  # COCO-Detection/oicr_WSR_101_DC5_1x.yaml
    "wsddn_WSR_18_DC5_1x": _wsr(18, heads="WSDDNROIHeads"),
}


def builtin_config(name, overrides=()):
    if name not in BUILTIN:
        raise KeyError(f"unknown builtin config '{name}'; have {sorted(BUILTIN)}")
    cfg = get_cfg()
    cfg.merge_from_other_cfg(CfgNode(copy.deepcopy(_VOC_BASE)))
    cfg.merge_from_other_cfg(CfgNode(copy.deepcopy(BUILTIN[name])))
    cfg.merge_from_list(list(overrides))
    return cfg


def precision_of(cfg):
    env = os.environ.get("DRN_B200_PRECISION")
    if env:
        return env
    b = cfg.get("B200") if hasattr(cfg, "get") else None
    if b is not None and "PRECISION" in b:
        return b["PRECISION"]
    return "fp32"
