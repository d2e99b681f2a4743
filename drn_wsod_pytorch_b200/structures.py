"""Data carriers at the boundary of the hot path.

Minimal stand-ins for detectron2/structures/{boxes,instances,image_list}.py with the same
attribute surface the path touches (`.tensor`, `.area()`, `.clip()`, `Instances.has/get/set`,
`image_size`, `__len__`, indexing, `.to`).  The model code is duck-typed: when it runs inside a
detectron2 process it accepts the reference's own `Instances`/`Boxes` objects and builds its
outputs with the caller's classes, so the reference's post-processing keeps working unchanged.
"""
from typing import Any, Dict, List, Tuple

import torch


class Boxes:
    """N x 4 XYXY absolute boxes (detectron2/structures/boxes.py:130-300)."""

    def __init__(self, tensor: torch.Tensor):
        if not isinstance(tensor, torch.Tensor):
            tensor = torch.as_tensor(tensor, dtype=torch.float32)
        if tensor.numel() == 0:
            tensor = tensor.reshape(-1, 4).to(torch.float32)
        assert tensor.dim() == 2 and tensor.size(-1) == 4, tensor.size()
        self.tensor = tensor

    def to(self, *args, **kwargs):
        return Boxes(self.tensor.to(*args, **kwargs))

    def clone(self):
        return Boxes(self.tensor.clone())

    def area(self):
        t = self.tensor
        return (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])

    def clip(self, box_size: Tuple[int, int]):
        h, w = box_size
        self.tensor[:, 0].clamp_(min=0, max=w)
        self.tensor[:, 1].clamp_(min=0, max=h)
        self.tensor[:, 2].clamp_(min=0, max=w)
        self.tensor[:, 3].clamp_(min=0, max=h)

    def nonempty(self, threshold: float = 0.0):
        t = self.tensor
        return ((t[:, 2] - t[:, 0]) > threshold) & ((t[:, 3] - t[:, 1]) > threshold)

    def scale(self, sx: float, sy: float):
        self.tensor[:, 0::2] *= sx
        self.tensor[:, 1::2] *= sy

    def __getitem__(self, item):
        if isinstance(item, int):
            return Boxes(self.tensor[item].view(1, -1))
        return Boxes(self.tensor[item])

    def __len__(self):
        return self.tensor.shape[0]

    @property
    def device(self):
        return self.tensor.device

    @classmethod
    def cat(cls, boxes_list):
        if len(boxes_list) == 0:
            return cls(torch.empty(0, 4))
        return cls(torch.cat([b.tensor for b in boxes_list], dim=0))

    def __repr__(self):
        return f"Boxes({self.tensor})"


class Instances:
    """Per-image bag of equally long fields (detectron2/structures/instances.py:8-185)."""

    def __init__(self, image_size: Tuple[int, int], **kwargs: Any):
        object.__setattr__(self, "_image_size", image_size)
        object.__setattr__(self, "_fields", {})
        for k, v in kwargs.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def __setattr__(self, name, val):
        if name.startswith("_"):
            object.__setattr__(self, name, val)
        else:
            self.set(name, val)

    def __getattr__(self, name):
        if name == "_fields" or name not in self._fields:
            raise AttributeError(f"Cannot find field '{name}' in the given Instances!")
        return self._fields[name]

    def set(self, name, value):
        if len(self._fields):
            assert len(self) == len(value), f"field '{name}' has length {len(value)}, expected {len(self)}"
        self._fields[name] = value

    def has(self, name):
        return name in self._fields

    def get(self, name):
        return self._fields[name]

    def remove(self, name):
        del self._fields[name]

    def get_fields(self) -> Dict[str, Any]:
        return self._fields

    def to(self, *args, **kwargs):
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            if hasattr(v, "to"):
                v = v.to(*args, **kwargs)
            ret.set(k, v)
        return ret

    def __getitem__(self, item):
        if isinstance(item, int):
            item = slice(item, None, len(self)) if item >= 0 else slice(item, None, len(self))
        ret = Instances(self._image_size)
        for k, v in self._fields.items():
            ret.set(k, v[item])
        return ret

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        raise NotImplementedError("Empty Instances does not support __len__!")

    def __repr__(self):
        return f"Instances(image_size={self._image_size}, fields={list(self._fields)})"


class ImageList:
    """Batch of images padded to one size (detectron2/structures/image_list.py:11-119)."""

    def __init__(self, tensor: torch.Tensor, image_sizes: List[Tuple[int, int]]):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    @property
    def device(self):
        return self.tensor.device


def detector_postprocess(results, output_height: int, output_width: int):
    """Rescale detections to the requested output resolution, clip, drop empty boxes.
    Behaviour of projects/WSL/wsl/modeling/postprocessing.py:10-79 for box-only instances."""
    cls = type(results)
    sx = output_width / results.image_size[1]
    sy = output_height / results.image_size[0]
    out = cls((output_height, output_width), **results.get_fields())
    boxes = out.pred_boxes if out.has("pred_boxes") else out.proposal_boxes
    boxes.scale(sx, sy)
    boxes.clip(out.image_size)
    return out[boxes.nonempty()]
