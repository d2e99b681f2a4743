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
from contextlib import contextmanager
from itertools import count

import numpy as np
import torch
from torch import nn

from . import ops
from .modeling import GeneralizedRCNNWSL, fast_rcnn_inference_single_image
from .structures import Boxes, Instances

__all__ = ["DatasetMapperTTAAVG", "GeneralizedRCNNWithTTAAVG", "transform_proposals", "resample_tables", "resize_u8",
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
    """test_time_augmentation_avg.py:27-64: `apply_box` -> clip -> nonempty(min_box_size) -> first top-k (no `unique`
    step, unlike detection_utils.transform_proposals).  Replaces dataset_dict["proposals"] in place.
    pin_memory (an addition): page-lock the results, so that the model's H2D copies of them are truly asynchronous
    (a pageable cudaMemcpyAsync first drains the stream: the host would wait for the previous view to finish)."""
    prop = dataset_dict["proposals"]
    boxes = prop.proposal_boxes.tensor.cpu().numpy()
    boxes = transforms.apply_box(boxes)
    boxes = type(prop.proposal_boxes)(torch.from_numpy(np.ascontiguousarray(boxes)))
    objectness_logits = prop.objectness_logits
    boxes.clip(image_shape)
    keep = boxes.nonempty(threshold=min_box_size)
    boxes = boxes[keep]
    objectness_logits = objectness_logits[keep.to(objectness_logits.device)]
    boxes, objectness_logits = boxes[:proposal_topk], objectness_logits[:proposal_topk]
    if pin_memory:
        boxes = type(boxes)(boxes.tensor.contiguous().pin_memory())
        if not objectness_logits.is_cuda:
            objectness_logits = objectness_logits.contiguous().pin_memory()
    proposals = type(prop)(tuple(image_shape))
    proposals.proposal_boxes = boxes
    proposals.objectness_logits = objectness_logits
    dataset_dict["proposals"] = proposals


class DatasetMapperTTAAVG:
    """test_time_augmentation_avg.py:67-136.  Takes one dataset dict (image: uint8 C x H x W tensor, host or device)
    and returns `len(MIN_SIZES) * (2 if FLIP else 1)` dicts whose `image` is the resampled (and flipped) view ON THE
    DEVICE, `transforms` the TransformList that produced it and `proposals` the transformed proposals."""

    def __init__(self, cfg, image_dtype=torch.uint8):
        self.min_sizes = cfg.TEST.AUG.MIN_SIZES
        self.max_size = cfg.TEST.AUG.MAX_SIZE
        self.flip = cfg.TEST.AUG.FLIP
        self.image_format = cfg.INPUT.FORMAT
        self.device = torch.device(cfg.MODEL.DEVICE)
        self.image_dtype = image_dtype  # torch.float32: the conversion preprocess_image would do is fused into the resample
        self.proposal_topk = None
        if cfg.MODEL.LOAD_PROPOSALS:
            self.proposal_topk = cfg.DATASETS.PRECOMPUTED_PROPOSAL_TOPK_TEST

    def __call__(self, dataset_dict):
        return list(self.iter_views(dataset_dict))

    def iter_views(self, dataset_dict):
        """The views of `__call__`, produced one at a time: the wrapper launches view i on the device before the host
        prepares view i+1 (resample launch + proposal transform), so the mapper's host work is hidden behind the GPU."""
        image = dataset_dict["image"]
        if image.dtype != torch.uint8:
            raise TypeError(f"DatasetMapperTTAAVG: uint8 image expected (the dataset mapper's output), got {image.dtype}")
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
                if self.proposal_topk is not None:
                    dic["proposals"] = dataset_dict["proposals"]
                    transform_proposals(dic, new_shape, tfms, proposal_topk=self.proposal_topk, pin_memory=self.device.type == "cuda")
                yield dic


# ---------------------------------------------------------------------------------------------- wrapper
class GeneralizedRCNNWithTTAAVG(nn.Module):
    """test_time_augmentation_avg.py:139-325 (box branch; the WSL detection configs have MASK_ON False).
    `__call__` has the interface of `GeneralizedRCNNWSL.forward` in eval mode."""

    def __init__(self, cfg, model, tta_mapper=None, batch_size=1):
        super().__init__()
        if isinstance(model, nn.parallel.DistributedDataParallel):
            model = model.module
        assert isinstance(model, GeneralizedRCNNWSL), \
            "TTA is only supported on GeneralizedRCNNWSL. Got a model of type {}".format(type(model))
        self.cfg = cfg.clone()
        assert not self.cfg.MODEL.KEYPOINT_ON, "TTA for keypoint is not supported yet"
        if self.cfg.MODEL.MASK_ON:
            raise NotImplementedError("the B200 TTA driver covers the box branch (the WSL detection configs)")
        self.model = model
        if tta_mapper is None:
            tta_mapper = DatasetMapperTTAAVG(cfg, image_dtype=torch.float32)
        self.tta_mapper = tta_mapper
        self.batch_size = batch_size

    @contextmanager
    def _turn_off_roi_heads(self, attrs):
        roi_heads = self.model.roi_heads
        old = {a: getattr(roi_heads, a) for a in attrs if hasattr(roi_heads, a)}
        for a in old:
            setattr(roi_heads, a, False)
        try:
            yield
        finally:
            for a, v in old.items():
                setattr(roi_heads, a, v)

    def _batch_inference(self, batched_inputs, detected_instances=None):
        """:200-225 -- `batch_size` views per model call; the per-view thresholding / NMS, whose result the box
        branch never reads, is skipped (outputs holds None per view)."""
        if detected_instances is None:
            detected_instances = [None] * len(batched_inputs)
        outputs, all_scores, all_boxes = [], [], []
        inputs, instances = [], []
        for idx, input, instance in zip(count(), batched_inputs, detected_instances):
            inputs.append(input)
            instances.append(instance)
            if len(inputs) == self.batch_size or idx == len(batched_inputs) - 1:
                output, all_score, all_box = self.model.inference(
                    inputs, instances if instances[0] is not None else None, do_postprocess=False,
                    with_detections=instances[0] is not None)
                outputs.extend(output)
                all_scores.extend(all_score)
                all_boxes.extend(all_box)
                inputs, instances = [], []
        return outputs, all_scores, all_boxes

    def __call__(self, batched_inputs):
        def _maybe_read_image(dataset_dict):
            ret = copy.copy(dataset_dict)
            if "image" not in ret:
                ret["image"] = read_image(ret.pop("file_name"), self.tta_mapper.image_format)
            if "height" not in ret and "width" not in ret:
                ret["height"] = ret["image"].shape[1]
                ret["width"] = ret["image"].shape[2]
            return ret

        return [self._inference_one_image(_maybe_read_image(x)) for x in batched_inputs]

    def _inference_one_image(self, input):
        orig_shape = (input["height"], input["width"])
        if hasattr(self.tta_mapper, "iter_views"):  # stream the views: host preparation of view i+1 overlaps view i on the GPU
            augmented_inputs, tfms = self.tta_mapper.iter_views(input), None
        else:
            augmented_inputs, tfms = self._get_augmented_inputs(input)
        with self._turn_off_roi_heads(["mask_on", "keypoint_on"]):
            all_boxes, all_scores, all_classes = self._get_augmented_boxes(augmented_inputs, tfms)
        merged_instances = self._merge_detections(all_boxes, all_scores, all_classes, orig_shape)
        return {"instances": merged_instances}

    def _get_augmented_inputs(self, input):
        augmented_inputs = self.tta_mapper(input)
        tfms = [x.pop("transforms") for x in augmented_inputs]
        return augmented_inputs, tfms

    def _get_augmented_boxes(self, augmented_inputs, tfms):
        """:286-309 -- boxes of every view back to the original image (inverse transforms), mean of boxes and scores over
        the views; the inverse transform, the sum and the division run in drn_tta_accumulate, view by view."""
        if tfms is None:  # an iterator of views that still carry their "transforms" (DatasetMapperTTAAVG.iter_views)
            n = len(self.tta_mapper.min_sizes) * (2 if self.tta_mapper.flip else 1)
            views = iter(augmented_inputs)
        else:
            n = len(augmented_inputs)
            views = iter([dict(v, transforms=t) for v, t in zip(augmented_inputs, tfms)])
        acc_boxes = acc_scores = None
        done = 0
        while done < n:  # merge as the views arrive: nothing but the accumulators is kept
            chunk = [next(views) for _ in range(min(self.batch_size, n - done))]
            chunk_tfms = [v.pop("transforms") for v in chunk]
            _, all_scores, all_boxes = self._batch_inference(chunk)
            for sc, bx, tfm in zip(all_scores, all_boxes, chunk_tfms):
                num_img, num_pred, num_col = bx.shape
                assert num_img == 1
                if acc_boxes is None:
                    acc_boxes, acc_scores = torch.empty_like(bx[0]), torch.empty_like(sc[0])
                assert bx[0].shape == acc_boxes.shape, "every view must keep the same proposals (torch.cat in the reference)"
                ops.tta_accumulate(bx[0], sc[0], tfm.inverse().device_params(), acc_boxes, acc_scores, done, n)
                done += 1
        return acc_boxes, acc_scores, None

    def _merge_detections(self, all_boxes, all_scores, all_classes, shape_hw):
        merged_instances, _ = fast_rcnn_inference_single_image(
            all_boxes, all_scores, shape_hw, self.cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST,
            self.cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST, self.cfg.TEST.DETECTIONS_PER_IMAGE, Instances, Boxes)
        return merged_instances


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
