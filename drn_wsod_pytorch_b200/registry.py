"""Registries with the reference's names.

`META_ARCH_REGISTRY`, `BACKBONE_REGISTRY`, `ROI_HEADS_REGISTRY`, `ROI_BOX_HEAD_REGISTRY` mirror
detectron2/modeling/{meta_arch/build.py:8, backbone/build.py:7, roi_heads/roi_heads.py:27,
roi_heads/box_head.py:16}.  `register_into_detectron2()` puts the B200 implementations into the
reference's own registries under the same names, so `detectron2.modeling.build_model(cfg)` with an
unchanged WSL YAML config returns the B200 model (call it INSTEAD of importing `wsl.modeling`:
fvcore's Registry asserts on duplicate names).
"""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._do_register(o.__name__, o)
                return o
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def _do_register(self, name, obj):
        assert name not in self._obj_map, f"An object named '{name}' was already registered in '{self._name}' registry!"
        self._obj_map[name] = obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    def names(self):
        return sorted(self._obj_map)


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
ROI_BOX_HEAD_REGISTRY = Registry("ROI_BOX_HEAD")


def register_into_detectron2():
    """Register the B200 modules in the reference's registries (drop-in for its configs)."""
    from detectron2.modeling import BACKBONE_REGISTRY as D2_BACKBONE
    from detectron2.modeling import META_ARCH_REGISTRY as D2_META
    from detectron2.modeling import ROI_HEADS_REGISTRY as D2_HEADS
    from detectron2.modeling.roi_heads.box_head import ROI_BOX_HEAD_REGISTRY as D2_BOX_HEAD

    from . import modeling  # noqa: F401  (populates our registries)

    for ours, theirs in ((META_ARCH_REGISTRY, D2_META), (BACKBONE_REGISTRY, D2_BACKBONE),
                         (ROI_HEADS_REGISTRY, D2_HEADS), (ROI_BOX_HEAD_REGISTRY, D2_BOX_HEAD)):
        for name in ours.names():
            theirs._do_register(name, ours.get(name))
