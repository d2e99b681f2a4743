"""Test-time augmentation on the B200: host mirror of
projects/WSL/wsl/modeling/test_time_augmentation_avg.py -- `DatasetMapperTTAAVG`, `GeneralizedRCNNWithTTAAVG`,
`transform_proposals` -- with the same names, constructor arguments and call signatures, so that
`tools/train_net.py:170-201` (`Trainer.test_with_TTA`) can wrap the B200 model unchanged.

What runs where:
  * the image is uploaded ONCE as uint8; every scale is resampled on the device by `drn_resample_u8_fwd`
    (bit-exact restatement of the Pillow 8-bit bilinear filter the reference reaches through
    `ResizeTransform.apply_image`, detectron2/data/transforms/transform.py:105-109), the flipped copy comes from the
    same kernel with the flip fused into its last pass;
  * proposals (R x 4 floats) are transformed on the host in numpy float32 exactly as the reference does
    (`apply_box` -> clip -> nonempty -> top-k: data-dependent length, host-resident input);
  * every view runs the captured eval pipeline of `GeneralizedRCNNWSL.inference(..., do_postprocess=False)` without
    the per-view NMS the reference computes and throws away;
  * `drn_tta_accumulate` maps each view's boxes back through the inverse transforms and keeps the running mean of
    boxes and scores on the device; one `drn_detections_fwd` (threshold + per-class NMS + top-k) finishes the image.
There is no CPU path: a missing library, a CPU model or a float image raises.
"""
import copy

import numpy as np
import torch
from torch import nn

from . import ops
from .modeling import GeneralizedRCNNWSL, fast_rcnn_inference_single_image
from .structures import Boxes, Instances

__all__ = ["DatasetMapperTTAAVG", "GeneralizedRCNNWithTTAAVG", "DatasetMapperTTAUNION", "GeneralizedRCNNWithTTAUNION", "transform_proposals", "resample_tables", "resize_u8",
           "NoOpTransform", "HFlipTransform", "ResizeTransform", "TransformList", "ResizeShortestEdge"]

PRECISION_BITS = 32 - 8 - 2  # Pillow Resample.c


# ---------------------------------------------------------------------------------------------- transforms
class Transform:
    """fvcore.transforms.transform.Transform: `apply_box` = the four corners through `apply_coords`, then min / max."""

    def apply_coords(self, coords):
        raise NotImplementedError

    def apply_box(self, box):
        idxs = np.array([(0, 1), (2, 1), (0, 3), (2, 3)]).flatten()
        coords = np.asarray(box).reshape(-1, 4)[:, idxs].reshape(-1, 2)
        coords = self.apply_coords(coords).reshape((-1, 4, 2))
        # = coords.min(axis=1) / coords.max(axis=1) (same values, same NaN propagation); numpy's reduction over a middle
        # axis of length 4 costs 0.5 ms per call for 4000 boxes -- 25 ms per image over the 16 TTA views
        c0, c1, c2, c3 = coords[:, 0], coords[:, 1], coords[:, 2], coords[:, 3]
        minxy = np.minimum(np.minimum(c0, c1), np.minimum(c2, c3))
        maxxy = np.maximum(np.maximum(c0, c1), np.maximum(c2, c3))
        return np.concatenate((minxy, maxxy), axis=1)

    def inverse(self):
        raise NotImplementedError

    def device_params(self):
        """[(kind, a, b)] records for drn_tta_accumulate (fp32 scalars as numpy would round them)."""
        raise NotImplementedError

    def __add__(self, other):
        return TransformList([self]) + other

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v}' for k, v in vars(self).items())})"


class NoOpTransform(Transform):
    def apply_coords(self, coords):
        return coords

    def inverse(self):
        return self

    def device_params(self):
        return []


class HFlipTransform(Transform):
    """fvcore HFlipTransform: x -> width - x."""

    def __init__(self, width):
        self.width = width

    def apply_coords(self, coords):
        coords[:, 0] = self.width - coords[:, 0]
        return coords

    def inverse(self):
        return self

    def device_params(self):
        return [(ops.TTA_OP_HFLIP, float(np.float32(self.width)), 0.0)]


class ResizeTransform(Transform):
    """detectron2/data/transforms/transform.py:83-134 (coordinates; the image side is `resize_u8`)."""

    def __init__(self, h, w, new_h, new_w, interp=None):
        self.h, self.w, self.new_h, self.new_w = h, w, new_h, new_w

    def apply_coords(self, coords):
        coords[:, 0] = coords[:, 0] * (self.new_w * 1.0 / self.w)
        coords[:, 1] = coords[:, 1] * (self.new_h * 1.0 / self.h)
        return coords

    def apply_image(self, img_chw_u8, flip=False, out_dtype=torch.uint8):
        assert tuple(img_chw_u8.shape[-2:]) == (self.h, self.w)
        return resize_u8(img_chw_u8, self.new_h, self.new_w, flip=flip, out_dtype=out_dtype)

    def inverse(self):
        return ResizeTransform(self.new_h, self.new_w, self.h, self.w)

    def device_params(self):
        # numpy multiplies the float32 coordinates by the Python-float scale rounded to float32 (weak scalar)
        return [(ops.TTA_OP_RESIZE, float(np.float32(self.new_w * 1.0 / self.w)), float(np.float32(self.new_h * 1.0 / self.h)))]


class TransformList(Transform):
    """fvcore TransformList: sequential application; `+` concatenates; inverse = reversed inverses."""

    def __init__(self, transforms):
        self.transforms = []
        for t in transforms:
            self.transforms.extend(t.transforms if isinstance(t, TransformList) else [t])

    def apply_coords(self, coords):
        for t in self.transforms:
            coords = t.apply_coords(coords)
        return coords

    def apply_box(self, box):
        for t in self.transforms:
            box = t.apply_box(box)
        return box

    def inverse(self):
        return TransformList([t.inverse() for t in self.transforms[::-1]])

    def device_params(self):
        return [p for t in self.transforms for p in t.device_params()]

    def __add__(self, other):
        return TransformList(self.transforms + (other.transforms if isinstance(other, TransformList) else [other]))

    def __radd__(self, other):
        return TransformList((other.transforms if isinstance(other, TransformList) else [other]) + self.transforms)

    def __len__(self):
        return len(self.transforms)


class ResizeShortestEdge:
    """detectron2/data/transforms/augmentation_impl.py:125-175 with one short-edge length (what the TTA mapper builds)."""

    def __init__(self, short_edge_length, max_size):
        self.size, self.max_size = int(short_edge_length), max_size

    def get_shape(self, h, w):
        size = self.size
        scale = size * 1.0 / min(h, w)
        if h < w:
            newh, neww = size, scale * w
        else:
            newh, neww = scale * h, size
        if max(newh, neww) > self.max_size:
            scale = self.max_size * 1.0 / max(newh, neww)
            newh = newh * scale
            neww = neww * scale
        return int(newh + 0.5), int(neww + 0.5)

    def get_transform(self, h, w):
        if self.size == 0:
            return NoOpTransform()
        return ResizeTransform(h, w, *self.get_shape(h, w))


# ---------------------------------------------------------------------------------------------- image resample
def resample_tables(in_size, out_size):
    """Pillow Resample.c `precompute_coeffs` (bilinear filter, support 1, box = the whole axis) + `normalize_coeffs_8bpc`:
    (bounds int32 [out, 2] = (first source index, tap count), coeffs int32 [out, ksize], 22-bit fixed point).
    Double arithmetic in the C code's operation order, so the integers equal Pillow's."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    center = 0.0 + (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    inv = 1.0 / filterscale
    lo = np.maximum((center - support + 0.5).astype(np.int64), 0)
    hi = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    cnt = hi - lo
    taps = np.arange(ksize, dtype=np.int64)[None, :]
    live = taps < cnt[:, None]
    arg = np.abs((taps + lo[:, None] - center[:, None] + 0.5) * inv)
    w = np.where(live & (arg < 1.0), 1.0 - arg, 0.0)
    total = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):  # left-to-right accumulation like the C loop (np.sum is pairwise)
        total = total + w[:, x]
    w = np.where(total[:, None] != 0.0, w / np.where(total == 0.0, 1.0, total)[:, None], w)
    fixed = (0.5 + w * (1 << PRECISION_BITS)).astype(np.int64)  # bilinear weights are >= 0: the C `(int)` cast truncates
    fixed = np.where(live, fixed, 0)
    return np.stack([lo, cnt], axis=1).astype(np.int32), fixed.astype(np.int32)


_TABLES = {}


def _device_tables(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    t = _TABLES.get(key)
    if t is None:
        if len(_TABLES) > 512:
            _TABLES.clear()
        b, k = resample_tables(in_size, out_size)
        t = (torch.from_numpy(b).to(device), torch.from_numpy(k).to(device), k.shape[1])
        _TABLES[key] = t
    return t


def resize_u8(img_chw_u8, new_h, new_w, flip=False, out_dtype=torch.uint8):
    """PIL `Image.resize((new_w, new_h), BILINEAR)` of a uint8 C x H x W CUDA tensor (+ optional horizontal flip and
    uint8 -> fp32 conversion) through drn_resample_u8_fwd.  Bit-exact with Pillow."""
    if img_chw_u8.dtype != torch.uint8:
        raise TypeError("the B200 TTA path resamples uint8 images (the reference's PIL branch, transform.py:105-109); "
                        f"got {img_chw_u8.dtype} -- float images take F.interpolate in the reference and are not supported")
    if not img_chw_u8.is_cuda:
        raise RuntimeError("resize_u8 needs a CUDA tensor: the B200 path has no CPU fallback")
    C, H, W = img_chw_u8.shape
    dev = img_chw_u8.device
    xt = _device_tables(W, new_w, dev) if new_w != W else (None, None, 0)
    yt = _device_tables(H, new_h, dev) if new_h != H else (None, None, 0)
    return ops.resample_u8(img_chw_u8.contiguous(), new_h, new_w, xt, yt, flip, out_dtype)


# ---------------------------------------------------------------------------------------------- mapper
def transform_proposals(dataset_dict, image_shape, transforms, *, proposal_topk, min_box_size=0, pin_memory=False):
    """What test_time_augmentation_avg.py:27-64 does to a view's proposals, as array arithmetic: boxes through the view's
    transforms (numpy float32, like `apply_box` there), clamped to the view, boxes whose width or height does not exceed
    `min_box_size` dropped, the first `proposal_topk` survivors kept in their original order (no `unique` step, unlike
    detection_utils.transform_proposals).  dataset_dict["proposals"] is replaced by a new Instances of the view's size.
    pin_memory (an addition): page-lock the results, so the model's H2D copies of them are truly asynchronous
    (a pageable cudaMemcpyAsync first drains the stream: the host would wait for the previous view to finish)."""
    src = dataset_dict["proposals"]
    view_h, view_w = image_shape
    xyxy = torch.from_numpy(np.ascontiguousarray(transforms.apply_box(src.proposal_boxes.tensor.cpu().numpy())))
    xyxy[:, 0::2] = xyxy[:, 0::2].clamp(min=0, max=view_w)
    xyxy[:, 1::2] = xyxy[:, 1::2].clamp(min=0, max=view_h)
    alive = ((xyxy[:, 2] - xyxy[:, 0]) > min_box_size) & ((xyxy[:, 3] - xyxy[:, 1]) > min_box_size)
    rows = torch.nonzero(alive).flatten()[:proposal_topk]
    xyxy = xyxy[rows]
    logits = src.objectness_logits[rows.to(src.objectness_logits.device)]
    if pin_memory:
        xyxy = xyxy.contiguous().pin_memory()
        if not logits.is_cuda:
            logits = logits.contiguous().pin_memory()
    out = type(src)(tuple(image_shape))
    out.proposal_boxes = type(src.proposal_boxes)(xyxy)
    out.objectness_logits = logits
    dataset_dict["proposals"] = out


class DatasetMapperTTAAVG:
    """test_time_augmentation_avg.py:67-136.  Takes one dataset dict (image: uint8 C x H x W tensor, host or device)
    and returns `len(MIN_SIZES) * (2 if FLIP else 1)` dicts whose `image` is the resampled (and flipped) view ON THE
    DEVICE, `transforms` the TransformList that produced it and `proposals` the transformed proposals."""

    transforms_proposals = True

    def __init__(self, cfg, image_dtype=torch.uint8):
        self.min_sizes = cfg.TEST.AUG.MIN_SIZES
        self.max_size = cfg.TEST.AUG.MAX_SIZE
        self.flip = cfg.TEST.AUG.FLIP
        self.image_format = cfg.INPUT.FORMAT
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.image_dtype = image_dtype  # torch.float32: the conversion preprocess_image would do is fused into the resample
        self.proposal_topk = None
        if cfg.MODEL.LOAD_PROPOSALS and self.transforms_proposals:
            self.proposal_topk = cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST

    @property
    def num_views(self):
        return len(self.min_sizes) * (2 if self.flip else 1)

    def __call__(self, dataset_dict):
        return list(self.iter_views(dataset_dict))

    def iter_views(self, dataset_dict):
        """The views of `__call__`, produced one at a time: the wrapper launches view i on the device before the host
        prepares view i+1 (resample launch + proposal transform), so the mapper's host work is hidden behind the GPU."""
        image = dataset_dict["image"]
        if image.dtype != torch.uint8:
            raise TypeError(f"{type(self).__name__}: uint8 image expected (the dataset mapper's output), got {image.dtype}")
        image = image.to(self.device, non_blocking=True)  # one H2D per image, every view is made on the device
        shape = tuple(image.shape[-2:])
        orig_shape = (dataset_dict["height"], dataset_dict["width"])
        if shape != orig_shape:
            pre_tfm = ResizeTransform(orig_shape[0], orig_shape[1], shape[0], shape[1])
        else:
            pre_tfm = NoOpTransform()
        rest = {k: v for k, v in dataset_dict.items() if k not in ("image", "proposals")}
        for min_size in self.min_sizes:
            resize = ResizeShortestEdge(min_size, self.max_size).get_transform(*shape)
            new_shape = (resize.new_h, resize.new_w) if isinstance(resize, ResizeTransform) else shape
            for do_flip in ([False, True] if self.flip else [False]):
                tfms = TransformList([resize] + ([HFlipTransform(new_shape[1])] if do_flip else []))
                dic = copy.copy(rest)
                dic["transforms"] = pre_tfm + tfms
                dic["image"] = resize_u8(image, new_shape[0], new_shape[1], flip=do_flip, out_dtype=self.image_dtype)
                if "proposals" in dataset_dict:
                    dic["proposals"] = dataset_dict["proposals"]
                    if self.proposal_topk is not None:
                        transform_proposals(dic, new_shape, tfms, proposal_topk=self.proposal_topk, pin_memory=self.device.type == "cuda")
                yield dic


class DatasetMapperTTAUNION(DatasetMapperTTAAVG):
    """test_time_augmentation_union.py:27-82.  Same views as the AVG mapper, but -- exactly as in the reference -- the
    proposals are handed to every view UNCHANGED (in the coordinates of the mapper's input image; the reference deep-copies
    the dataset dict and never calls transform_proposals), so with precomputed proposals only the un-resized,
    un-flipped view pools the regions the proposals were computed for."""

    transforms_proposals = False


# ---------------------------------------------------------------------------------------------- wrappers
class _TTADriver(nn.Module):
    """Shared part of the two TTA wrappers: one image in, its views streamed through the model chunk by chunk.
    The host prepares view i+1 (resample launch, proposal transform) while the device runs view i; nothing is kept per
    view beyond what `consume` stores."""

    mapper_cls = DatasetMapperTTAAVG
    per_view_detections = False  # run the per-view threshold / NMS / top-k tail of the model?

    def __init__(self, cfg, model, tta_mapper=None, batch_size=1):
        super().__init__()
        model = getattr(model, "module", model) if isinstance(model, nn.parallel.DistributedDataParallel) else model
        if not isinstance(model, GeneralizedRCNNWSL):
            raise AssertionError(f"TTA is only supported on GeneralizedRCNNWSL. Got a model of type {type(model)}")
        if cfg.MODEL.KEYPOINT_ON or cfg.MODEL.MASK_ON:
            raise NotImplementedError("the B200 TTA driver covers the box branch (the WSL detection configs: no masks / keypoints)")
        self.cfg = cfg.clone()
        self.model = model
        self.tta_mapper = tta_mapper if tta_mapper is not None else self.mapper_cls(cfg, image_dtype=torch.float32)
        self.batch_size = max(1, int(batch_size))

    def __call__(self, batched_inputs):
        return [self.run_image(x) for x in batched_inputs]

    def _with_image(self, dataset_dict):
        d = dict(dataset_dict)  # the caller's dict is left alone
        if "image" not in d:
            d["image"] = read_image(d.pop("file_name"), self.tta_mapper.image_format)
        if "height" not in d and "width" not in d:
            d["height"], d["width"] = int(d["image"].shape[1]), int(d["image"].shape[2])
        return d

    def stream_views(self, dataset_dict, consume):
        """Run every view of `dataset_dict` through the model, `batch_size` views per call, and hand each view's
        (instances or None, all_scores [R, K+1], all_boxes [R, 4K], TransformList) to `consume(i, n, ...)` in view order
        (n = number of views).  Returns n."""
        mapper = self.tta_mapper
        if hasattr(mapper, "iter_views"):
            views, n = mapper.iter_views(dataset_dict), mapper.num_views
        else:  # a user-supplied mapper with the reference's list interface
            made = mapper(dataset_dict)
            views, n = iter(made), len(made)
        for first in range(0, n, self.batch_size):
            chunk = [next(views) for _ in range(min(self.batch_size, n - first))]
            tfms = [v.pop("transforms") for v in chunk]
            res, scores, boxes = self.model.inference(chunk, do_postprocess=False, with_detections=self.per_view_detections)
            for j, (r, sc, bx, tfm) in enumerate(zip(res, scores, boxes, tfms)):
                if bx.shape[0] != 1:
                    raise AssertionError("one image per view expected")
                consume(first + j, n, r, sc[0], bx[0], tfm)
        return n


class GeneralizedRCNNWithTTAAVG(_TTADriver):
    """test_time_augmentation_avg.py:139-325 (box branch).  `__call__` has the interface of `GeneralizedRCNNWSL.forward`
    in eval mode.  Every view's all_boxes go back to the original image through the view's inverse transforms and are
    averaged with the all_scores over the views by `drn_tta_accumulate` as the views arrive (running sums in two device
    buffers, divided on the last view); threshold / per-class NMS / top-k run once, on the means."""

    mapper_cls = DatasetMapperTTAAVG
    per_view_detections = False  # the reference computes per-view detections here and never reads them

    def merged_views(self, dataset_dict):
        """(mean all_boxes [R, 4K], mean all_scores [R, K+1]) over the views of one image."""
        acc = {}

        def consume(i, n, _res, scores, boxes, tfm):
            if not acc:
                acc["boxes"], acc["scores"] = torch.empty_like(boxes), torch.empty_like(scores)
            elif boxes.shape != acc["boxes"].shape:
                raise AssertionError("every view must keep the same proposals (the reference concatenates the views' outputs)")
            ops.tta_accumulate(boxes, scores, tfm.inverse().device_params(), acc["boxes"], acc["scores"], i, n)

        self.stream_views(dataset_dict, consume)
        return acc["boxes"], acc["scores"]

    def run_image(self, dataset_dict):
        d = self._with_image(dataset_dict)
        boxes, scores = self.merged_views(d)
        m = self.cfg.MODEL.ROI_HEADS
        inst, _ = fast_rcnn_inference_single_image(boxes, scores, (d["height"], d["width"]), m.SCORE_THRESH_TEST, m.NMS_THRESH_TEST,
                                                   self.cfg.TEST.DETECTIONS_PER_IMAGE, Instances, Boxes)
        return {"instances": inst}


class GeneralizedRCNNWithTTAUNION(_TTADriver):
    """test_time_augmentation_union.py:85-262 (box branch): every view keeps its OWN detections (threshold, per-class NMS,
    top-k inside the model), their boxes go back to the original image through the inverse transforms, and the union of all
    views' detections -- one row per detection, its score in its class column of an otherwise zero [n, K+1] matrix -- goes
    through threshold 1e-8 / per-class NMS / top-k once more (:246-262)."""

    mapper_cls = DatasetMapperTTAUNION
    per_view_detections = True
    UNION_SCORE_THRESH = 1e-8  # test_time_augmentation_union.py:259

    def union_of_views(self, dataset_dict):
        """(boxes [n, 4] on the original image, scores [n], classes [n]) of all views' detections, in view order."""
        parts = []

        def consume(i, n, res, _scores, _boxes, tfm):
            b = res.pred_boxes.tensor.contiguous()
            if b.shape[0]:
                sc = res.scores.contiguous().view(-1, 1)
                back, dummy = torch.empty_like(b), torch.empty_like(sc)
                ops.tta_accumulate(b, sc, tfm.inverse().device_params(), back, dummy, 0, 1)
                b = back
            parts.append((b, res.scores, res.pred_classes))

        self.stream_views(dataset_dict, consume)
        return tuple(torch.cat([p[j] for p in parts], dim=0) for j in range(3))

    def run_image(self, dataset_dict):
        d = self._with_image(dataset_dict)
        boxes, scores, classes = self.union_of_views(d)
        K = self.cfg.MODEL.ROI_HEADS.NUM_CLASSES
        if boxes.shape[0] == 0:  # no view detected anything
            return {"instances": Instances((d["height"], d["width"]), pred_boxes=Boxes(boxes), scores=scores, pred_classes=classes)}
        table = torch.zeros((boxes.shape[0], K + 1), dtype=torch.float32, device=boxes.device)
        table[torch.arange(boxes.shape[0], device=boxes.device), classes] = scores
        inst, _ = fast_rcnn_inference_single_image(boxes, table, (d["height"], d["width"]), self.UNION_SCORE_THRESH,
                                                   self.cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST, self.cfg.TEST.DETECTIONS_PER_IMAGE,
                                                   Instances, Boxes)
        return {"instances": inst}


def read_image(file_name, format=None):
    """detectron2/data/detection_utils.py read_image for the formats the WSL configs use ("BGR", "RGB"): uint8 C x H x W."""
    from PIL import Image

    with open(file_name, "rb") as f:
        image = np.asarray(Image.open(f).convert("RGB"))
    if format == "BGR":
        image = image[:, :, ::-1]
    elif format not in (None, "RGB"):
        raise ValueError(f"unsupported INPUT.FORMAT {format}")
    return torch.from_numpy(np.ascontiguousarray(image.transpose(2, 0, 1)))
