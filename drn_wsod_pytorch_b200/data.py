"""Input formats either side of the hot path: precomputed-proposal files, the proposal
transform that produces the `proposals` Instances the model consumes, and weight files.

Host-side restatements of the reference's loaders (no kernels: this is file parsing and a few hundred boxes per
image on the dataloader's CPU workers), kept here so that a user of the reference finds the same entry points:

* `load_proposals_into_dataset`  -- detectron2/data/build.py:102-153
* `transform_proposals`          -- detectron2/data/detection_utils.py:209-254 (with the WSL fork's `unique_boxes` step,
                                    detectron2/structures/boxes.py:216-227)
* `convert_proposals`            -- projects/WSL/tools/proposal_convert.py:52-94 (MCG / selective-search .mat arrays -> the
                                    pickled dict; the .mat reading itself stays with scipy)
* `load_checkpoint`              -- detectron2/checkpoint/detection_checkpoint.py:26-73 (`.pth` / `.pkl` in the Detectron2
                                    format, incl. the suffix-matching heuristic of c2_model_loading.py:211-313 used by the
                                    `resnet*_ws_model_120_d2.pkl` backbones; Caffe2 name conversion is not provided)
"""
import pickle
from typing import Dict, List, Optional

import numpy as np
import torch

from .structures import Boxes, Instances

XYXY_ABS, XYWH_ABS = 0, 1  # detectron2/structures/boxes.py BoxMode values used by proposal files


def _to_xyxy(boxes: np.ndarray, mode) -> np.ndarray:
    mode = int(getattr(mode, "value", mode))
    boxes = np.asarray(boxes)
    if mode == XYXY_ABS:
        return boxes
    if mode == XYWH_ABS:
        out = boxes.astype(np.float64 if boxes.dtype == np.float64 else np.float32, copy=True)
        out[:, 2] += out[:, 0]
        out[:, 3] += out[:, 1]
        return out
    raise NotImplementedError(f"proposal bbox_mode {mode}: only XYXY_ABS (0) and XYWH_ABS (1) occur in proposal files")


def load_proposals_into_dataset(dataset_dicts: List[dict], proposal_file: str) -> List[dict]:
    """detectron2/data/build.py:102-153: attach `proposal_boxes`, `proposal_objectness_logits` (sorted by descending
    score) and `proposal_bbox_mode` to every record.  Accepts the Detectron1 key names (`indexes`, `scores`) that
    projects/WSL/tools/proposal_convert.py writes."""
    with open(proposal_file, "rb") as f:
        proposals = pickle.load(f, encoding="latin1")
    for old, new in (("indexes", "ids"), ("scores", "objectness_logits")):
        if old in proposals:
            proposals[new] = proposals.pop(old)
    img_ids = {str(r["image_id"]) for r in dataset_dicts}
    id_to_index = {str(i): n for n, i in enumerate(proposals["ids"]) if str(i) in img_ids}
    bbox_mode = int(getattr(proposals.get("bbox_mode", XYXY_ABS), "value", proposals.get("bbox_mode", XYXY_ABS)))
    for record in dataset_dicts:
        i = id_to_index[str(record["image_id"])]
        boxes, logits = proposals["boxes"][i], proposals["objectness_logits"][i]
        inds = logits.argsort()[::-1]
        record["proposal_boxes"] = boxes[inds]
        record["proposal_objectness_logits"] = logits[inds]
        record["proposal_bbox_mode"] = bbox_mode
    return dataset_dicts


def unique_boxes(boxes: torch.Tensor, scale: float = 1.0) -> np.ndarray:
    """detectron2/structures/boxes.py:216-227 (WSL fork): indices of the first occurrence of every distinct rounded box,
    ascending.  (The reference spells the cast `np.int`, removed in NumPy 1.24; int64 is what it meant.)"""
    b = boxes.detach().cpu().numpy()
    hashes = np.round(b * scale).dot(np.array([1, 1e3, 1e6, 1e9])).astype(np.int64)
    _, index = np.unique(hashes, return_index=True)
    return np.sort(index)


def transform_proposals(dataset_dict: dict, image_shape, transforms=None, *, proposal_topk: int, min_box_size: float = 0):
    """detectron2/data/detection_utils.py:209-254: boxes -> XYXY_ABS -> `transforms.apply_box` -> clip -> unique ->
    min-size filter -> top-k; pops the three `proposal_*` keys and adds `proposals` (Instances with `proposal_boxes`,
    `objectness_logits`).  `transforms`: any object with `apply_box(ndarray[N,4]) -> ndarray[N,4]` (the reference's
    TransformList), or None."""
    if "proposal_boxes" not in dataset_dict:
        return
    boxes = _to_xyxy(dataset_dict.pop("proposal_boxes"), dataset_dict.pop("proposal_bbox_mode"))
    if transforms is not None:
        boxes = transforms.apply_box(boxes)
    boxes = Boxes(torch.as_tensor(np.asarray(boxes).astype("float32")).reshape(-1, 4))
    logits = torch.as_tensor(dataset_dict.pop("proposal_objectness_logits").astype("float32"))
    boxes.clip(image_shape)
    keep = torch.from_numpy(unique_boxes(boxes.tensor))
    boxes, logits = boxes[keep], logits[keep]
    keep = boxes.nonempty(threshold=min_box_size)
    boxes, logits = boxes[keep], logits[keep]
    proposals = Instances(tuple(image_shape))
    proposals.proposal_boxes = boxes[:proposal_topk]
    proposals.objectness_logits = logits[:proposal_topk]
    dataset_dict["proposals"] = proposals


def convert_proposals(ids, boxes_per_image, scores_per_image=None, one_based_yxyx: bool = True) -> Dict[str, list]:
    """projects/WSL/tools/proposal_convert.py:52-94: per-image proposal arrays (MCG: `boxes` in 1-based (y1, x1, y2, x2)
    with `scores`; selective search: no scores -> 1.0) -> the dict the loaders read: `boxes` int16-representable
    XYXY_ABS 0-based, `scores` float32, `indexes`."""
    out = {"boxes": [], "scores": [], "indexes": []}
    for n, (i, b) in enumerate(zip(ids, boxes_per_image)):
        b = np.asarray(b)
        if one_based_yxyx:
            b = b[:, (1, 0, 3, 2)] - 1
        s = np.ones((b.shape[0],), np.float32) if scores_per_image is None else np.asarray(scores_per_image[n], np.float32).reshape(-1)
        out["boxes"].append(b.astype(np.int16))
        out["scores"].append(s)
        out["indexes"].append(i)
    return out


class HFlipTransform:
    """fvcore HFlipTransform.apply_box for XYXY boxes: x -> width - x, corners re-sorted."""

    def __init__(self, width: int):
        self.width = width

    def apply_box(self, box: np.ndarray) -> np.ndarray:
        box = np.asarray(box, dtype=np.float64)
        out = box.copy()
        out[:, 0] = self.width - box[:, 2]
        out[:, 2] = self.width - box[:, 0]
        return out


class ResizeTransform:
    """detectron2/data/transforms/transform.py ResizeTransform.apply_coords on box corners."""

    def __init__(self, h: int, w: int, new_h: int, new_w: int):
        self.h, self.w, self.new_h, self.new_w = h, w, new_h, new_w

    def apply_box(self, box: np.ndarray) -> np.ndarray:
        out = np.asarray(box, dtype=np.float64).copy()
        out[:, 0::2] *= self.new_w * 1.0 / self.w
        out[:, 1::2] *= self.new_h * 1.0 / self.h
        return out


class TransformList:
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def apply_box(self, box: np.ndarray) -> np.ndarray:
        for t in self.transforms:
            box = t.apply_box(box)
        return box


# ---------------------------------------------------------------------------------------------- weights
def _load_file(path: str) -> dict:
    """detection_checkpoint.py:26-46."""
    if path.endswith(".pkl"):
        with open(path, "rb") as f:
            data = pickle.load(f, encoding="latin1")
        if "model" in data and "__author__" in data:
            return data
        raise NotImplementedError("Caffe2 / Detectron1 .pkl weights need detectron2's name conversion "
                                  "(checkpoint/c2_model_loading.py:9-208); convert them with the reference first")
    loaded = torch.load(path, map_location="cpu")
    return loaded if "model" in loaded else {"model": loaded}


def align_and_update_state_dicts(model_state: Dict[str, torch.Tensor], ckpt_state: Dict[str, torch.Tensor]) -> Dict[str, str]:
    """c2_model_loading.py:211-313 without the Caffe2 renaming: every model key takes the checkpoint key that is its
    longest complete suffix (`a == b or a.endswith('.' + b)`), shapes must agree, one checkpoint key may feed one model
    key only.  Updates `model_state` in place, returns {checkpoint key: model key}."""
    model_keys, ckpt_keys = sorted(model_state), sorted(ckpt_state)
    matched = {}
    for mk in model_keys:
        best = None
        for ck in ckpt_keys:
            if (mk == ck or mk.endswith("." + ck)) and (best is None or len(ck) > len(best)):
                best = ck
        if best is None:
            continue
        v = ckpt_state[best]
        if tuple(model_state[mk].shape) != tuple(v.shape):
            continue  # the reference warns and skips
        if best in matched:
            raise ValueError(f"Cannot match one checkpoint key to multiple keys in the model: {best} -> {matched[best]}, {mk}")
        model_state[mk] = v.clone()
        matched[best] = mk
    return matched


def load_checkpoint(model: torch.nn.Module, path: str, strict: bool = False):
    """DetectionCheckpointer.load for the formats the WSL configs use (`MODEL.WEIGHTS: *.pkl | *.pth`).  Returns the
    (missing_keys, unexpected_keys) pair of `load_state_dict`; `pixel_mean` / `pixel_std` are never reported missing
    (detection_checkpoint.py:63-72).  Parameters are copied IN PLACE, so captured plans and derived layouts follow."""
    ckpt = _load_file(path)
    state = {}
    for k, v in ckpt["model"].items():
        if isinstance(v, np.ndarray):
            v = torch.from_numpy(v)
        if not torch.is_tensor(v):
            raise ValueError(f"Unsupported type found in checkpoint! {k}: {type(v)}")
        state[k[len("module."):] if k.startswith("module.") else k] = v
    if ckpt.get("matching_heuristics", False):
        model_state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        align_and_update_state_dicts(model_state, state)
        state = model_state
    model_state = model.state_dict()
    for k in list(state):  # shape mismatches are dropped with a warning in fvcore's Checkpointer
        if k in model_state and tuple(model_state[k].shape) != tuple(state[k].shape):
            state.pop(k)
    res = model.load_state_dict(state, strict=strict)
    missing = [k for k in res.missing_keys if k not in ("pixel_mean", "pixel_std")]
    return missing, list(res.unexpected_keys)
