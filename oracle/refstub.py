"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

In-memory stub layer that lets the UNMODIFIED reference sources under
/root/reference (detectron2 v0.2 fork + projects/WSL) import and run on CPU in
this container, where fvcore / yacs / termcolor / pycocotools / pydensecrf and
the compiled `detectron2._C` / `wsl._C` extensions are absent (SURVEY.md §8c,
Appendix A).  Nothing here touches hot-path arithmetic: the stubs are config
plumbing, registries, file IO shims and weight-init helpers.

Only usable where /root/reference exists (this container).  It is used by
`tests/golden/make_golden.py` to generate the committed golden vectors, by the
`not gpu` test that pins `oracle/wsl_oracle.py` against the live reference, and
never on the GPU box.
"""
import copy
import importlib.util
import math
import os
import sys
import types

import torch
import yaml

REFERENCE_ROOT = os.environ.get("DRN_REFERENCE_ROOT", "/root/reference")
WSL_ROOT = os.path.join(REFERENCE_ROOT, "projects", "WSL")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "detectron2")) and os.path.isdir(
        os.path.join(WSL_ROOT, "wsl")
    )


class _Lenient(types.ModuleType):
    """Module whose unknown attributes resolve to inert dummy classes."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        dummy = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, dummy)
        return dummy


def _mod(name, lenient=True, **attrs):
    m = (_Lenient if lenient else types.ModuleType)(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


# ----------------------------------------------------------------------------
# fvcore.common.config.CfgNode (attr-dict with yacs-like merge semantics)
# ----------------------------------------------------------------------------
class CfgNode(dict):
    IMMUTABLE = "__immutable__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        super().__init__()
        self.__dict__[CfgNode.IMMUTABLE] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get(CfgNode.IMMUTABLE, False):
            raise AttributeError("immutable CfgNode")
        self[name] = value

    def clone(self):
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = type(self)()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def is_frozen(self):
        return self.__dict__.get(CfgNode.IMMUTABLE, False)

    def _set_immutable(self, flag):
        self.__dict__[CfgNode.IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

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
                    dict.__setitem__(self, k, CfgNode())
                self[k].merge_from_other_cfg(v)
            else:
                dict.__setitem__(self, k, self._coerce(v, self[k]) if k in self else v)

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            d = self
            keys = full_key.split(".")
            for sub in keys[:-1]:
                d = d[sub]
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except Exception:
                    pass
            dict.__setitem__(d, keys[-1], self._coerce(v, d.get(keys[-1], v)))

    def merge_from_file(self, cfg_filename, allow_unsafe=False):
        loaded = self.load_yaml_with_base(cfg_filename, allow_unsafe=allow_unsafe)
        self.merge_from_other_cfg(type(self)(loaded))

    @staticmethod
    def load_yaml_with_base(filename, allow_unsafe=False):
        with open(filename, "r") as f:
            cfg = yaml.load(f, Loader=_TupleLoader) or {}

        def merge_a_into_b(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and k in b and isinstance(b[k], dict):
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if "_BASE_" in cfg:
            base = cfg.pop("_BASE_")
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not base.startswith("/"):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = CfgNode.load_yaml_with_base(base, allow_unsafe)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def dump(self, **kwargs):
        def plain(n):
            return {k: plain(v) if isinstance(v, dict) else v for k, v in n.items()}

        return yaml.safe_dump(plain(self), **kwargs)


class _TupleLoader(yaml.SafeLoader):
    pass


def _construct_python_tuple(loader, node):
    return tuple(loader.construct_sequence(node))


_TupleLoader.add_constructor("tag:yaml.org,2002:python/tuple", _construct_python_tuple)


# ----------------------------------------------------------------------------
class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"{name} already registered in {self._name}"
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:

            def deco(fn):
                self._do_register(fn.__name__, fn)
                return fn

            return deco
        self._do_register(obj.__name__, obj)

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map


class HistoryBuffer:
    def __init__(self, max_length=1000000):
        self._data = []
        self._count = 0
        self._global_avg = 0

    def update(self, value, iteration=None):
        self._data.append((value, iteration))
        self._count += 1
        self._global_avg += (value - self._global_avg) / self._count

    def latest(self):
        return self._data[-1][0]

    def median(self, window_size):
        import numpy as np

        return np.median([x[0] for x in self._data[-window_size:]])

    def avg(self, window_size):
        import numpy as np

        return np.mean([x[0] for x in self._data[-window_size:]])

    def global_avg(self):
        return self._global_avg

    def values(self):
        return self._data


class _PathManager:
    @staticmethod
    def open(path, mode="r", **kw):
        return open(path, mode)

    @staticmethod
    def get_local_path(path, **kw):
        return path

    @staticmethod
    def isfile(path):
        return os.path.isfile(path)

    @staticmethod
    def isdir(path):
        return os.path.isdir(path)

    @staticmethod
    def exists(path):
        return os.path.exists(path)

    @staticmethod
    def mkdirs(path):
        os.makedirs(path, exist_ok=True)

    @staticmethod
    def ls(path):
        return os.listdir(path)

    @staticmethod
    def register_handler(handler):
        pass


def _c2_msra_fill(module):
    torch.nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        torch.nn.init.constant_(module.bias, 0)


def _c2_xavier_fill(module):
    torch.nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        torch.nn.init.constant_(module.bias, 0)


def _smooth_l1_loss(input, target, beta, reduction="none"):
    # fvcore.nn.smooth_l1_loss (published definition; beta<1e-5 -> pure L1)
    if beta < 1e-5:
        loss = torch.abs(input - target)
    else:
        n = torch.abs(input - target)
        loss = torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
    if reduction == "mean":
        loss = loss.mean()
    elif reduction == "sum":
        loss = loss.sum()
    return loss


_INSTALLED = False


def _fvcore_transforms(names):
    """Functional restatement of the fvcore (>= 0.1.1, absent here) transform classes the reference's TTA driver
    uses -- Transform.apply_box (four corners through apply_coords, then min / max), TransformList (sequential
    application, `+` concatenates, inverse = reversed inverses), HFlipTransform (x -> width - x, np.flip on
    axis 1), NoOpTransform -- fvcore/transforms/transform.py.  The other names stay inert placeholders."""
    import numpy as np

    class Transform:
        def __init__(self, *a, **k):
            pass

        def _set_attributes(self, params=None):
            if params:
                for k, v in params.items():
                    if k != "self" and not k.startswith("_"):
                        setattr(self, k, v)

        @classmethod
        def register_type(cls, *a, **k):
            pass

        def apply_box(self, box):
            idxs = np.array([(0, 1), (2, 1), (0, 3), (2, 3)]).flatten()
            coords = np.asarray(box).reshape(-1, 4)[:, idxs].reshape(-1, 2)
            coords = self.apply_coords(coords).reshape((-1, 4, 2))
            minxy = coords.min(axis=1)
            maxxy = coords.max(axis=1)
            return np.concatenate((minxy, maxxy), axis=1)

        def apply_segmentation(self, segmentation):
            return self.apply_image(segmentation)

        def inverse(self):
            raise NotImplementedError

    class TransformList(Transform):
        def __init__(self, transforms):
            super().__init__()
            self.transforms = list(transforms)

        def _apply(self, x, meth):
            for t in self.transforms:
                x = getattr(t, meth)(x)
            return x

        def __getattribute__(self, name):
            if name.startswith("apply_"):
                return lambda x: self._apply(x, name)
            return super().__getattribute__(name)

        def __add__(self, other):
            others = other.transforms if isinstance(other, TransformList) else [other]
            return TransformList(self.transforms + others)

        def __iadd__(self, other):
            others = other.transforms if isinstance(other, TransformList) else [other]
            self.transforms.extend(others)
            return self

        def __radd__(self, other):
            others = other.transforms if isinstance(other, TransformList) else [other]
            return TransformList(others + self.transforms)

        def __len__(self):
            return len(self.transforms)

        def inverse(self):
            return TransformList([t.inverse() for t in self.transforms[::-1]])

    class HFlipTransform(Transform):
        def __init__(self, width):
            super().__init__()
            self.width = width

        def apply_image(self, img):
            return np.flip(img, axis=1) if img.ndim <= 3 else np.flip(img, axis=-2)

        def apply_coords(self, coords):
            coords[:, 0] = self.width - coords[:, 0]
            return coords

        def inverse(self):
            return self

    class NoOpTransform(Transform):
        def apply_image(self, img):
            return img

        def apply_coords(self, coords):
            return coords

        def inverse(self):
            return self

        def __getattr__(self, name):
            if name.startswith("apply_"):
                return lambda x: x
            raise AttributeError(name)

    real = {"Transform": Transform, "TransformList": TransformList, "HFlipTransform": HFlipTransform,
            "NoOpTransform": NoOpTransform}
    return {n: real.get(n, type(n, (Transform,), {})) for n in names}


def install(with_wsl=True):
    """Install the stubs and import the reference.  Returns (detectron2, wsl).
    with_wsl=False imports detectron2 only (used to test the registry drop-in, where the B200
    package registers the WSL names instead of the reference's wsl.modeling)."""
    global _INSTALLED
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if _INSTALLED:
        return sys.modules["detectron2"], sys.modules.get("wsl")

    _mod("fvcore", __version__="0.1.2")
    _mod("fvcore.common")
    _mod("fvcore.common.config", CfgNode=CfgNode)
    _mod("fvcore.common.file_io", PathManager=_PathManager)
    _mod("fvcore.common.registry", Registry=Registry)
    _mod("fvcore.common.history_buffer", HistoryBuffer=HistoryBuffer)
    _mod("fvcore.common.timer")
    _mod("fvcore.common.checkpoint")
    _mod("fvcore.nn", smooth_l1_loss=_smooth_l1_loss)
    _mod("fvcore.nn.weight_init", c2_msra_fill=_c2_msra_fill, c2_xavier_fill=_c2_xavier_fill)
    _mod("fvcore.nn.precise_bn")

    names = [
        "BlendTransform", "CropTransform", "GridSampleTransform", "HFlipTransform",
        "VFlipTransform", "NoOpTransform", "ScaleTransform", "Transform", "TransformList",
    ]
    tattrs = _fvcore_transforms(names)
    _mod("fvcore.transforms", **tattrs)
    _mod("fvcore.transforms.transform", __all__=names, **tattrs)

    _mod("termcolor", colored=lambda s, *a, **k: s)
    _mod("pycocotools")
    _mod("pycocotools.mask")
    _mod("pycocotools.coco")
    _mod("pycocotools.cocoeval")
    _mod("pydensecrf")
    _mod("pydensecrf.densecrf")
    _mod("pydensecrf.utils")
    try:
        import PIL.Image

        if not hasattr(PIL.Image, "LINEAR"):
            PIL.Image.LINEAR = PIL.Image.BILINEAR
    except ImportError:
        pass

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import detectron2  # noqa

    c = _mod("detectron2._C")
    detectron2._C = c
    if not with_wsl:
        return detectron2, None

    spec = importlib.util.spec_from_file_location(
        "wsl",
        os.path.join(WSL_ROOT, "wsl", "__init__.py"),
        submodule_search_locations=[os.path.join(WSL_ROOT, "wsl")],
    )
    wsl = importlib.util.module_from_spec(spec)
    sys.modules["wsl"] = wsl
    wc = _mod("wsl._C")
    wsl._C = wc
    # the one compiled op on a path this repo covers: the PCL loss (wsl/layers/pcl_loss.py:41-50, :90-102), compiled from the
    # reference's own pcl_loss_cpu.cpp by oracle/build_ref.py into oracle/_ref/
    try:
        from oracle.build_ref import load_pcl_ref

        pcl = load_pcl_ref()
    except Exception:  # no compiler / no reference sources: the PCL configs then fail at the first loss call
        pcl = None
    if pcl is not None:
        wc.pcl_loss_forward = pcl.pcl_loss_forward
        wc.pcl_loss_backward = pcl.pcl_loss_backward
    if not torch.cuda.is_available():
        # wsl/layers/pcl_loss.py:52,121 moves the op's CPU result back with `.cuda(device_id)` (device_id = -1 for CPU
        # tensors): on a box without CUDA that call is made the identity -- a placement shim, no arithmetic
        torch.Tensor.cuda = lambda self, *a, **k: self
    spec.loader.exec_module(wsl)
    _INSTALLED = True
    return detectron2, wsl


def build_reference_model(yaml_rel, overrides=()):
    """Build the reference model from one of its own YAML configs on CPU.

    yaml_rel: path relative to projects/WSL/configs (e.g.
    'PascalVOC-Detection/oicr_WSR_18_DC5_1x.yaml').
    """
    install()
    from detectron2.config import get_cfg
    from detectron2.modeling import build_model
    from wsl.config import add_wsl_config

    cfg = get_cfg()
    add_wsl_config(cfg)
    cfg.merge_from_file(os.path.join(WSL_ROOT, "configs", yaml_rel))
    cfg.merge_from_list(list(overrides))
    cfg.MODEL.DEVICE = "cpu"
    cfg.freeze()
    model = build_model(cfg)
    return cfg, model
