"""Thin Python wrappers over the C ABI (one function per entry point of include/drn_b200.h).

Tensors are torch CUDA tensors used purely as device buffers; all arithmetic happens inside
libdrn_b200.so.  Activations are NHWC.  Nothing here falls back to torch ops.
"""
import os

import torch

from . import lib
from .lib import DRN_BF16, DRN_F32, DRN_U8, call, current_stream, fvec, ivec


def _dt(t):
    if t.dtype == torch.float32:
        return DRN_F32
    if t.dtype == torch.bfloat16:
        return DRN_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _chk(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 path has no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")


def first_conv(img_chw, canvas_hw, mean, std, packed, stride, relu=True, out_dtype=torch.float32):
    """Fused (img-mean)/std + 3x3 Cin=3 conv + affine + ReLU.  img_chw: 3xHxW fp32 -> 1xHoxWoxCout."""
    _chk(img_chw, "image")
    assert img_chw.dtype == torch.float32 and img_chw.dim() == 3 and img_chw.shape[0] == 3
    H, W = img_chw.shape[1:]
    Hp, Wp = canvas_hw
    Cout = packed["cout"]
    Ho, Wo = (Hp + 2 - 3) // stride + 1, (Wp + 2 - 3) // stride + 1
    out = torch.empty((1, Ho, Wo, Cout), device=img_chw.device, dtype=out_dtype)
    call("drn_conv3x3_c3_fwd", img_chw, H, W, Hp, Wp, fvec(mean), fvec(std), packed["w"], packed["scale"],
         packed["bias"], Cout, stride, int(relu), out, _dt(out), current_stream())
    return out


def conv_f32(x, packed, ksize, dilation, relu, residual=None, out=None, ldo=None):
    """SIMT fp32 implicit-GEMM conv / linear.  x: [N,H,W,Cin] fp32 NHWC."""
    _chk(x, "x")
    N, H, W, Cin = x.shape
    Cout = packed["cout"]
    if out is None:
        out = torch.empty((N, H, W, Cout), device=x.device, dtype=torch.float32)
        ldo = Cout
    call("drn_conv_igemm_f32", x, N, H, W, Cin, packed["w"], ksize, dilation, packed["scale"], packed["bias"],
         residual, int(relu), out, Cout, ldo, current_stream())
    return out


_GEMM_WS = {}
# the split-K schedules of the deep-K GEMMs are opt-in (both measured slower on fc6: profiles/r1_*_negative_result.txt);
# only then is their zero-initialised scratch allocated (inside a graph capture the zero-fill is replayed every step)
SPLIT_K_WORKSPACE = os.environ.get("DRN_TC_STREAMK") == "1" or os.environ.get("DRN_TC_TAILSPLIT") == "1"


def _gemm_workspace(device):
    """Zero-initialised split-K scratch, one per (device, stream): kernels on one stream are ordered, so
    they can share it; the kernel resets the flags it raises."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _GEMM_WS.get(key)
    if ws is None:
        ws = torch.zeros((lib.load().drn_gemm_workspace_bytes(),), device=device, dtype=torch.uint8)
        _GEMM_WS[key] = ws
    return ws


def conv_bf16_tc(x, packed, ksize, dilation, relu, residual=None, out_dtype=torch.bfloat16, dropout_p=0.0, dropout_seed=0,
                 dropout_seed_dev=None, out=None):
    """tcgen05 implicit-GEMM conv / linear (+ fused train-mode dropout).  x: [N,H,W,Cin] bf16 NHWC.
    out: optional preallocated [N,H,W,Cout] destination (e.g. a row block of a larger matrix)."""
    _chk(x, "x")
    N, H, W, Cin = x.shape
    Cout = packed["cout"]
    if out is None:
        out = torch.empty((N, H, W, Cout), device=x.device, dtype=out_dtype)
    else:
        _chk(out, "out")
        assert out.numel() == N * H * W * Cout
    ws = _gemm_workspace(x.device) if (SPLIT_K_WORKSPACE and ksize == 1 and Cin >= 1024) else None  # deep-K GEMMs only
    call("drn_conv_igemm_bf16_tc", x, N, H, W, Cin, packed["w"], ksize, dilation, packed["scale"], packed["bias"],
         residual, int(relu), out, _dt(out), Cout, Cout, float(dropout_p), int(dropout_seed), dropout_seed_dev,
         ws, 0 if ws is None else ws.numel(), current_stream())
    return out


def maxpool2x2(x, stride):
    _chk(x, "x")
    N, H, W, C = x.shape
    Ho, Wo = (H - 2) // stride + 1, (W - 2) // stride + 1
    out = torch.empty((N, Ho, Wo, C), device=x.device, dtype=x.dtype)
    call("drn_maxpool2x2_nhwc", x, N, H, W, C, stride, _dt(x), out, current_stream())
    return out


_ROIPOOL_WS = {}


def roipool(feat_hwc, boxes, objectness, spatial_scale, use_tables=None):
    """feat_hwc: [h,w,C]; boxes [R,4] fp32; objectness [R] fp32 or None -> [R, 49*C] (bin-major).
    use_tables: build the per-image range-max tables (default: when R is large enough to amortise them)."""
    _chk(feat_hwc, "features")
    _chk(boxes, "boxes")
    h, w, C = feat_hwc.shape
    R = boxes.shape[0]
    out = torch.empty((R, 49 * C), device=feat_hwc.device, dtype=feat_hwc.dtype)
    if use_tables is None:
        use_tables = R >= 256
    ws, ws_bytes = None, 0
    if use_tables:
        ws_bytes = lib.load().drn_roipool_workspace_bytes(h, w, C, _dt(feat_hwc))
        key = (feat_hwc.device, torch.cuda.current_stream().cuda_stream)
        ws = _ROIPOOL_WS.get(key)
        if ws is None or ws.numel() < ws_bytes:  # grow-only scratch, reused across calls on the same stream
            ws = torch.empty((ws_bytes,), device=feat_hwc.device, dtype=torch.uint8)
            _ROIPOOL_WS[key] = ws
    call("drn_roipool_fwd", feat_hwc, h, w, C, boxes, objectness, R, float(spatial_scale), _dt(feat_hwc), out,
         ws, ws_bytes, current_stream())
    return out


def roipool_tables(feat_hwc):
    """Build the per-image range-max tables of `feat_hwc` on the current stream; returns the scratch tensor
    (reused per stream) for roipool_rows, or None when the channel count does not fit the table layout."""
    _chk(feat_hwc, "features")
    h, w, C = feat_hwc.shape
    if not lib.load().drn_roipool_tables_supported(C, _dt(feat_hwc)):
        return None
    ws_bytes = lib.load().drn_roipool_workspace_bytes(h, w, C, _dt(feat_hwc))
    key = (feat_hwc.device, torch.cuda.current_stream().cuda_stream)
    ws = _ROIPOOL_WS.get(key)
    if ws is None or ws.numel() < ws_bytes:
        ws = torch.empty((ws_bytes,), device=feat_hwc.device, dtype=torch.uint8)
        _ROIPOOL_WS[key] = ws
    call("drn_roipool_build_tables", feat_hwc, h, w, C, _dt(feat_hwc), ws, ws.numel(), current_stream())
    return ws


def roipool_rows(feat_hwc, boxes, objectness, spatial_scale, tables, out, max_ctas=0):
    """Pool the given rows (boxes [r,4], objectness [r] or None) into out [r, 49*C] from prebuilt tables,
    on the current stream (the caller orders it after roipool_tables).  max_ctas > 0: persistent grid."""
    _chk(boxes, "boxes")
    _chk(out, "out")
    h, w, C = feat_hwc.shape
    call("drn_roipool_rows_fwd", feat_hwc, h, w, C, boxes, objectness, boxes.shape[0], float(spatial_scale), _dt(feat_hwc),
         out, tables, tables.numel(), int(max_ctas), current_stream())
    return out


def drop_scratch(stream):
    """Free the grow-only scratch buffers that were created for `stream` (a warm-up side stream)."""
    for d in (_ROIPOOL_WS, _GEMM_WS):
        for key in [k for k in d if k[1] == stream.cuda_stream]:
            del d[key]


def dropout_(x, p, seed, seed_dev=None):
    """In place.  Effective seed = seed + *seed_dev (seed_dev: 1-element int64 device tensor or None)."""
    call("drn_dropout_inplace", x, x.numel(), _dt(x), float(p), int(seed), seed_dev, current_stream())
    return x


def wsddn_mil(logits, K, cls_off, det_off, gt_onehot, mean_loss, loss_scale, loss_out):
    R, ld = logits.shape
    dev = logits.device
    scores = torch.empty((R, K), device=dev, dtype=torch.float32)
    img_score = torch.empty((K,), device=dev, dtype=torch.float32)
    ws = torch.empty((2 * R + K,), device=dev, dtype=torch.float32)
    call("drn_wsddn_mil_fwd", logits, ld, R, K, cls_off, det_off, gt_onehot, int(mean_loss), float(loss_scale),
         scores, img_score, loss_out, ws, current_stream())
    return scores, img_score


def oicr_pgt(prev_scores, boxes, gt_classes, img_score, rederive, deltas, ld_deltas, cls_agnostic, bbox_w):
    R, ld = prev_scores.shape
    G = gt_classes.numel()
    dev = prev_scores.device
    pgt_idx = torch.empty((G,), device=dev, dtype=torch.int64)
    pgt_score = torch.empty((G,), device=dev, dtype=torch.float32)
    pgt_box = torch.empty((G, 4), device=dev, dtype=torch.float32)
    pgt_weight = torch.empty((G,), device=dev, dtype=torch.float32)
    call("drn_oicr_pgt", prev_scores, ld, R, boxes, gt_classes, G, img_score, int(rederive), deltas, int(ld_deltas),
         int(cls_agnostic), fvec(bbox_w), pgt_idx, pgt_score, pgt_box, pgt_weight, current_stream())
    return pgt_idx, pgt_score, pgt_box, pgt_weight


def label_proposals(boxes, gt_boxes, gt_classes, K, thresholds, labels_cfg):
    R = boxes.shape[0]
    G = 0 if gt_classes is None else gt_classes.numel()
    dev = boxes.device
    labels = torch.empty((R,), device=dev, dtype=torch.int64)
    matched = torch.empty((R,), device=dev, dtype=torch.int64)
    counts = torch.empty((3,), device=dev, dtype=torch.int32)
    call("drn_label_proposals", boxes, R, gt_boxes if G else None, gt_classes if G else None, G, K,
         fvec(thresholds), ivec(labels_cfg), len(thresholds), labels, matched, counts, current_stream())
    return labels, matched, counts


def oicr_stage(logits, col_off, K, labels, matched, pgt_weight, loss_scale, loss_out, counter):
    R, ld = logits.shape
    dev = logits.device
    probs = torch.empty((R, K + 1), device=dev, dtype=torch.float32)
    stats = torch.empty((6,), device=dev, dtype=torch.float32)
    weights = torch.empty((R,), device=dev, dtype=torch.float32)
    nb = (R + 255) // 256
    part = torch.empty((6 * nb,), device=dev, dtype=torch.float32)
    call("drn_oicr_stage_fwd", logits, ld, col_off, R, K, labels, matched, pgt_weight, pgt_weight.numel(),
         float(loss_scale), probs, loss_out, stats, weights, part, counter, current_stream())
    return probs, stats, weights


def wsddn_mil_pgt(logits, K, cls_off, det_off, gt_onehot, mean_loss, loss_scale, loss_out, boxes, gt_int, counter):
    """Fused WSDDN MIL + stage-0 pseudo-GT mining -> scores, img_score, (pgt_idx, pgt_score, pgt_box, pgt_weight)."""
    R, ld = logits.shape
    dev = logits.device
    G = gt_int.numel()
    scores = torch.empty((R, K), device=dev, dtype=torch.float32)
    img_score = torch.empty((K,), device=dev, dtype=torch.float32)
    ws = torch.empty((2 * R + K,), device=dev, dtype=torch.float32)
    pgt = (torch.empty((G,), device=dev, dtype=torch.int64), torch.empty((G,), device=dev, dtype=torch.float32),
           torch.empty((G, 4), device=dev, dtype=torch.float32), torch.empty((G,), device=dev, dtype=torch.float32))
    call("drn_wsddn_mil_pgt_fwd", logits, ld, R, K, cls_off, det_off, gt_onehot, int(mean_loss), float(loss_scale), boxes,
         gt_int, G, scores, img_score, loss_out, pgt[0], pgt[1], pgt[2], pgt[3], ws, counter, current_stream())
    return scores, img_score, pgt


def oicr_stage_fused(logits, col_off, K, boxes, gt_int, pgt_box, pgt_weight, thresholds, labels_cfg, loss_scale, loss_out,
                     counter, first_gt=None, nxt=None):
    """One fused refinement stage (labelling + weighted CE + softmax [+ first labelling vs the real GT]
    [+ next stage's pseudo GT]).  first_gt = (gt_boxes, gt_classes); nxt = dict(img_score, deltas, ld_deltas,
    cls_agnostic, bbox_w).  Returns dict(labels, matched, counts, probs, stats, weights, first=(...)|None, next=(...)|None)."""
    R, ld = logits.shape
    dev = logits.device
    G = gt_int.numel()
    nb = (R + 255) // 256
    i64 = lambda *s: torch.empty(s, device=dev, dtype=torch.int64)
    f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    out = dict(labels=i64(R), matched=i64(R), counts=torch.empty((3,), device=dev, dtype=torch.int32), probs=f32(R, K + 1),
               stats=f32(6), weights=f32(R), first=None, next=None)
    part = f32((12 + 2 * G) * nb)
    if first_gt is not None:
        gtb, gtc = first_gt
        Gb = gtc.numel()
        out["first"] = (i64(R), i64(R), torch.empty((3,), device=dev, dtype=torch.int32))
        f_args = (gtb if Gb else None, gtc if Gb else None, Gb) + out["first"]
    else:
        f_args = (None, None, -1, None, None, None)
    if nxt is not None:
        out["next"] = (i64(G), f32(G), f32(G, 4), f32(G))
        n_args = (nxt["img_score"], nxt["deltas"], int(nxt["ld_deltas"]), int(nxt["cls_agnostic"]), fvec(nxt["bbox_w"])) + out["next"]
    else:
        n_args = (None, None, 0, 0, None, None, None, None, None)
    call("drn_oicr_stage_fused_fwd", logits, ld, col_off, R, K, boxes, gt_int, G, pgt_box, pgt_weight, fvec(thresholds),
         ivec(labels_cfg), len(thresholds), float(loss_scale), *f_args, out["labels"], out["matched"], out["counts"],
         out["probs"], loss_out, out["stats"], out["weights"], *n_args, part, counter, current_stream())
    return out


def oicr_stages_launch1(logits, col_offs, delta_offs, bbox_ws, K, boxes, gt_int, cls_agnostic, thresholds, labels_cfg, counters,
                        first_gt=None, stream=None):
    """Launch 1 of drn_oicr_stages_fwd (every stage's softmax + the next stage's pseudo GT): needs the logits only, so a
    caller may run it on `stream` (a torch.cuda.Stream that has waited for the logits) beside the MIL kernels.  Allocates every
    output of both launches (on the current stream) and returns the context for oicr_stages_launch2."""
    R, ld = logits.shape
    dev = logits.device
    S, G = len(col_offs), gt_int.numel()
    nb = (R + 255) // 256
    assert counters.numel() >= 2 * S
    i64 = lambda *s: torch.empty(s, device=dev, dtype=torch.int64)
    f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    c = dict(S=S, G=G, K=K, logits=logits, boxes=boxes, gt_int=gt_int, cls_agnostic=int(cls_agnostic), counters=counters,
             col_offs=ivec(col_offs), delta_offs=ivec(delta_offs), bw=fvec([float(v) for w in bbox_ws for v in w]),
             thr=fvec(thresholds), labs=ivec(labels_cfg), nthr=len(thresholds),
             probs=f32(S, R, K + 1), labels=i64(S, R), matched=i64(S, R), weights=f32(S, R),
             counts=torch.empty((S, 3), device=dev, dtype=torch.int32), stats=f32(S, 6),
             pgt_idx=i64(S, G), pgt_score=f32(S, G), pgt_box=f32(S, G, 4), pgt_w=f32(S, G), part=f32(S * (12 + 2 * G) * nb))
    if first_gt is not None:
        gtb, gtc = first_gt
        Gb = gtc.numel()
        c["first"] = (i64(R), i64(R), torch.empty((3,), device=dev, dtype=torch.int32))
        c["f_args"] = (gtb if Gb else None, gtc if Gb else None, Gb) + c["first"]
    else:
        c["first"] = None
        c["f_args"] = (None, None, -1, None, None, None)
    _oicr_stages_call(c, None, None, 1.0, None, ivec([0] * S), 1, current_stream() if stream is None else stream.cuda_stream)
    return c


def _oicr_stages_call(c, img_score, pgt0, loss_scale, loss_row, loss_cols, phases, stream):
    R, ld = c["logits"].shape
    call("drn_oicr_stages_fwd", c["logits"], ld, R, c["K"], c["S"], c["col_offs"], c["delta_offs"], c["bw"], c["boxes"], c["gt_int"],
         c["G"], img_score, c["cls_agnostic"], None if pgt0 is None else pgt0[2], None if pgt0 is None else pgt0[3], c["thr"],
         c["labs"], c["nthr"], float(loss_scale), *c["f_args"], c["probs"], c["pgt_idx"], c["pgt_score"], c["pgt_box"], c["pgt_w"],
         c["labels"], c["matched"], c["counts"], c["weights"], c["stats"], loss_row, loss_cols, c["part"], c["counters"],
         int(phases), stream)


def oicr_stages_launch2(c, img_score, pgt0, loss_scale, loss_row, loss_cols):
    """Launch 2 (every stage's labelling + weighted CE) on the current stream, which must have joined launch 1's.
    pgt0 = (idx, score, box, weight) of stage 0 (wsddn_mil_pgt); stage k's loss goes to loss_row[loss_cols[k]].
    Returns (list of S dicts: labels, matched, counts, probs, stats, weights, pgt=(idx, score, box, weight); first labelling)."""
    _oicr_stages_call(c, img_score, pgt0, loss_scale, loss_row, ivec(loss_cols), 2, current_stream())
    out = []
    for k in range(c["S"]):
        pgt = tuple(pgt0) if k == 0 else (c["pgt_idx"][k], c["pgt_score"][k], c["pgt_box"][k], c["pgt_w"][k])
        out.append(dict(labels=c["labels"][k], matched=c["matched"][k], counts=c["counts"][k], probs=c["probs"][k],
                        stats=c["stats"][k], weights=c["weights"][k], pgt=pgt))
    return out, c["first"]


def oicr_stages(logits, col_offs, delta_offs, bbox_ws, K, boxes, gt_int, img_score, cls_agnostic, pgt0, thresholds, labels_cfg,
                loss_scale, loss_row, loss_cols, counters, first_gt=None):
    """All S refinement stages of one image in two launches on the current stream (drn_oicr_stages_fwd)."""
    c = oicr_stages_launch1(logits, col_offs, delta_offs, bbox_ws, K, boxes, gt_int, cls_agnostic, thresholds, labels_cfg, counters,
                            first_gt=first_gt)
    return oicr_stages_launch2(c, img_score, pgt0, loss_scale, loss_row, loss_cols)


def oicr_boxreg_loss(logits, col_off, K, cls_agnostic, boxes, pgt_box, labels, matched, bbox_w, beta, loss_scale,
                     loss_out, counter):
    R, ld = logits.shape
    nb = (R + 255) // 256
    part = torch.empty((nb,), device=logits.device, dtype=torch.float32)
    call("drn_oicr_boxreg_loss", logits, ld, col_off, R, K, int(cls_agnostic), boxes, pgt_box, labels, matched,
         fvec(bbox_w), float(beta), float(loss_scale), loss_out, part, counter, current_stream())


def oicr_infer(logits, K, col_offs, delta_offs, boxes, bbox_w, nreg):
    R, ld = logits.shape
    dev = logits.device
    all_scores = torch.empty((R, K + 1), device=dev, dtype=torch.float32)
    all_boxes = torch.empty((R, 4 * nreg), device=dev, dtype=torch.float32)
    call("drn_oicr_infer", logits, ld, R, K, nreg, len(col_offs), ivec(col_offs), ivec(delta_offs), boxes, fvec(bbox_w),
         all_scores, all_boxes, current_stream())
    return all_scores, all_boxes


def detections(all_scores, all_boxes, image_hw, score_thresh, nms_thresh, cap):
    """Device-side fast_rcnn_inference_single_image: returns fixed-size (boxes [cap,4], scores [cap], classes [cap]
    int64, rows [cap] int64, count int32[1]); the first `count` entries are the detections, best first."""
    _chk(all_scores, "all_scores")
    _chk(all_boxes, "all_boxes")
    R, K1 = all_scores.shape
    K = K1 - 1
    nreg = all_boxes.shape[1] // 4
    dev = all_scores.device
    cap = int(cap)
    out_boxes = torch.empty((cap, 4), device=dev, dtype=torch.float32)
    out_scores = torch.empty((cap,), device=dev, dtype=torch.float32)
    out_classes = torch.empty((cap,), device=dev, dtype=torch.int64)
    out_rows = torch.empty((cap,), device=dev, dtype=torch.int64)
    count = torch.empty((1,), device=dev, dtype=torch.int32)
    nbytes = lib.load().drn_detections_workspace_bytes(R, K)
    ws = torch.empty((max(nbytes, 16),), device=dev, dtype=torch.uint8)
    call("drn_detections_fwd", all_scores, all_boxes, R, K, nreg, float(image_hw[0]), float(image_hw[1]), float(score_thresh),
         float(nms_thresh), cap, out_boxes, out_scores, out_classes, out_rows, count, ws, ws.numel(), current_stream())
    return out_boxes, out_scores, out_classes, out_rows, count


# ---------------------------------------------------------------- "fp32_tc": fp32-accurate layers on the bf16 tensor cores
# weight planes that meet the "small" activation planes x1 | x2 | x1 | x2 | x3 (include/drn_b200.h drn_f32tc_split)
F32TC_SMALL_W = (1, 0, 2, 1, 0)


def f32tc_split(x, Cg):
    """x: fp32 [..., C] -> (big bf16 [C/Cg, ..., Cg] = term x1 per K-group, small bf16 [..., 5C] = correction planes)."""
    _chk(x, "x")
    assert x.dtype == torch.float32
    C = x.shape[-1]
    assert C % Cg == 0
    rows = x.numel() // C
    big = torch.empty((C // Cg,) + tuple(x.shape[:-1]) + (Cg,), device=x.device, dtype=torch.bfloat16)
    small = torch.empty(tuple(x.shape[:-1]) + (5 * C,), device=x.device, dtype=torch.bfloat16)
    call("drn_f32tc_split", x, rows, C, Cg, big, small, current_stream())
    return big, small


def f32tc_reduce(partials, bias=None, residual=None, relu=False):
    """partials: fp32 [n, ..., C] -> relu?(sum_n partials + bias + residual) as fp32 [..., C] (round-to-nearest adds)."""
    _chk(partials, "partials")
    assert partials.dtype == torch.float32
    n, C = partials.shape[0], partials.shape[-1]
    rows = partials[0].numel() // C
    y = torch.empty(tuple(partials.shape[1:]), device=partials.device, dtype=torch.float32)
    if residual is not None:
        _chk(residual, "residual")
        assert residual.dtype == torch.float32 and residual.numel() == y.numel()
    call("drn_f32tc_reduce", partials, n, rows * C, C, bias, residual, int(bool(relu)), rows, C, y, current_stream())
    return y


# ---------------------------------------------------------------- test-time augmentation (tta.py)
TTA_OP_NOOP, TTA_OP_RESIZE, TTA_OP_HFLIP = 0, 1, 2  # include/drn_b200.h DRN_TTA_OP_*
TTA_MAX_OPS = 4


def resample_u8(img_chw, new_h, new_w, xtab, ytab, flip=False, out_dtype=torch.uint8):
    """Pillow 8-bit bilinear resize of a uint8 [C,H,W] image (+ flip, + uint8 -> fp32).  xtab / ytab:
    (bounds int32 [new,2], coeffs int32 [new,ksize], ksize) device tables, or (None, None, 0) for an unchanged axis."""
    _chk(img_chw, "image")
    assert img_chw.dtype == torch.uint8 and img_chw.dim() == 3
    C, H, W = img_chw.shape
    assert out_dtype in (torch.uint8, torch.float32)
    out = torch.empty((C, new_h, new_w), device=img_chw.device, dtype=out_dtype)
    tmp = torch.empty((C, H, new_w), device=img_chw.device, dtype=torch.uint8) if (new_w != W and new_h != H) else None
    call("drn_resample_u8_fwd", img_chw, C, H, W, xtab[0], xtab[1], int(xtab[2]), ytab[0], ytab[1], int(ytab[2]), int(new_h),
         int(new_w), tmp, int(bool(flip)), out, DRN_U8 if out_dtype == torch.uint8 else DRN_F32, current_stream())
    return out


def tta_accumulate(all_boxes, all_scores, inverse_ops, acc_boxes, acc_scores, view_index, n_views):
    """acc += (boxes through `inverse_ops` [(kind, a, b)], scores); overwrite on view 0, divide by n_views on the last."""
    for t, name in ((all_boxes, "all_boxes"), (all_scores, "all_scores"), (acc_boxes, "acc_boxes"), (acc_scores, "acc_scores")):
        _chk(t, name)
        assert t.dtype == torch.float32
    R, box_cols = all_boxes.shape
    assert all_scores.shape[0] == R and acc_boxes.shape == all_boxes.shape and acc_scores.shape == all_scores.shape
    if len(inverse_ops) > TTA_MAX_OPS:
        raise RuntimeError(f"tta_accumulate: {len(inverse_ops)} transforms per view (max {TTA_MAX_OPS})")
    kinds = ivec([o[0] for o in inverse_ops] or [0])
    a = fvec([o[1] for o in inverse_ops] or [0.0])
    b = fvec([o[2] for o in inverse_ops] or [0.0])
    call("drn_tta_accumulate", all_boxes, all_scores, R, box_cols, all_scores.shape[1], len(inverse_ops), kinds, a, b, acc_boxes,
         acc_scores, int(view_index), int(n_views), current_stream())


# ---------------------------------------------------------------- PCL refinement stage
def pcl_stage(logits, col_off, K, boxes, center_boxes, center_classes, center_scores, loss_scale, loss_out, counter):
    """One PCL stage on the device from host-mined cluster centres -> dict(probs, labels, weights, assignment, pc_probs,
    pc_count, img_w) (third_party/pcl.py:148-200 + the pcl_loss op forward)."""
    R, ld = logits.shape
    P = center_boxes.shape[0]
    dev = logits.device
    f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
    i32 = lambda *s: torch.empty(s, device=dev, dtype=torch.int32)
    o = dict(probs=f32(R, K + 1), labels=i32(R), weights=f32(R), assignment=i32(R), pc_probs=f32(P), pc_count=f32(P), img_w=f32(P))
    ws = f32(P + 1)
    assert center_classes.dtype == torch.int32 and center_boxes.dtype == torch.float32 and counter.dtype == torch.int32
    call("drn_pcl_stage_fwd", logits, ld, R, K, int(col_off), boxes, center_boxes, center_classes, center_scores, P, float(loss_scale),
         o["probs"], o["labels"], o["weights"], o["assignment"], o["pc_probs"], o["pc_count"], o["img_w"], loss_out, ws, counter,
         current_stream())
    return o


def pcl_stage_bwd(st, K, loss_scale, grad_loss, col_off, dlogits):
    R = st["probs"].shape[0]
    call("drn_pcl_stage_bwd", st["probs"], R, K, st["labels"], st["weights"], st["assignment"], st["pc_probs"], st["pc_count"], st["img_w"],
         float(loss_scale), grad_loss, int(col_off), dlogits.shape[1], dlogits, current_stream())


# ---------------------------------------------------------------- backward of the trainable tail
def wsddn_mil_bwd(logits, K, cls_off, det_off, scores, gt_onehot, mean_loss, loss_scale, grad_loss, dlogits):
    R, ld = logits.shape
    ws = torch.empty((K,), device=logits.device, dtype=torch.float32)
    call("drn_wsddn_mil_bwd", logits, ld, R, K, cls_off, det_off, scores, gt_onehot, int(mean_loss), float(loss_scale), grad_loss,
         dlogits, ws, current_stream())


def oicr_stage_bwd(probs, labels, weights, nvalid, loss_scale, grad_loss, K, col_off, dlogits):
    R, ld = dlogits.shape
    call("drn_oicr_stage_bwd", probs, labels, weights, nvalid, float(loss_scale), grad_loss, R, K, ld, col_off, dlogits,
         current_stream())


def oicr_boxreg_bwd(logits, col_off, K, cls_agnostic, boxes, pgt_box, labels, matched, bbox_w, beta, loss_scale, denom, grad_loss,
                    dlogits):
    R, ld = logits.shape
    call("drn_oicr_boxreg_bwd", logits, ld, col_off, R, K, int(cls_agnostic), boxes, pgt_box, labels, matched, fvec(bbox_w),
         float(beta), float(loss_scale), float(denom), grad_loss, dlogits, current_stream())


def masked_transpose(grad, mask=None, mul=1.0, c49=0, ld_out=None, out_dtype=None, want_masked=False):
    """grad [R, C] (* mask != 0) * mul -> (transposed [C, ld_out] zero-padded beyond R, masked [R, C] or None)."""
    _chk(grad, "grad")
    R, C = grad.shape
    ld_out = R if ld_out is None else ld_out
    out_dtype = out_dtype or grad.dtype
    out_t = torch.empty((C, ld_out), device=grad.device, dtype=out_dtype)
    masked = torch.empty((R, C), device=grad.device, dtype=out_dtype) if want_masked else None
    if mask is not None:
        _chk(mask, "mask")
        assert mask.shape == grad.shape
    if mask is not None:
        mask_dt = _dt(mask)
    else:  # no mask: pick the code of a supported (grad, mask, out) combination
        mask_dt = DRN_F32 if (grad.dtype == torch.float32 and out_dtype == torch.float32) else DRN_BF16
    call("drn_masked_transpose", grad, C, _dt(grad), mask, C, mask_dt, float(mul), R, C, int(c49), out_t, ld_out, masked, C,
         _dt(out_t), current_stream())
    return out_t, masked


def rowsum(x, cols=None):
    rows, ld = x.shape
    out = torch.empty((rows,), device=x.device, dtype=torch.float32)
    call("drn_rowsum", x, ld, rows, ld if cols is None else cols, _dt(x), out, current_stream())
    return out


def permute_cols49(x, c49):
    out = torch.empty_like(x)
    call("drn_permute_cols49", x, out, x.shape[0], int(c49), current_stream())
    return out


def sgd_step(w, grad, momentum_buf, packed, c49, lr, momentum, weight_decay, nesterov, first_step):
    """In place: torch.optim.SGD update of `w` (fp32) and, in the same pass, of its bf16 kernel-layout copy `packed`
    (None: no copy; c49 > 0: the fc6 column permutation)."""
    _chk(w, "param")
    _chk(grad, "grad")
    assert w.dtype == torch.float32 and grad.dtype == torch.float32 and grad.shape == w.shape
    rows = w.shape[0] if w.dim() > 1 else 1
    cols = w.numel() // max(rows, 1)
    call("drn_sgd_step", w, grad, momentum_buf, packed, rows, cols, int(c49 or 0), float(lr), float(momentum), float(weight_decay),
         int(bool(nesterov)), int(bool(first_step)), current_stream())


def gemm_scatter(a, b, peer_windows, my_rank):
    """dW tile-scattered: a [M, K] bf16 (dY^T), b [N, K] bf16 (X^T) -> row block m // (M / n) of a @ b^T stored into rank
    (m // (M / n))'s window `peer_windows[that rank]` (raw device pointers of mapped peer memory), slot `my_rank`."""
    _chk(a, "a")
    _chk(b, "b")
    M, K = a.shape
    N = b.shape[0]
    call("drn_gemm_bf16_tc_scatter", a, M, K, b, N, None, lib.pvec(peer_windows), len(peer_windows), int(my_rank), N, current_stream())


def sgd_step_sharded(w, momentum_shard, slots, packed_ptrs, row0, rows, c49, lr, momentum, weight_decay, nesterov, first_step):
    """Owner's optimizer step on rows [row0, row0 + rows) of `w` from the gradient slots [n_src, rows, cols]; refreshed bf16
    kernel-layout rows go to every rank's weight buffer (raw device pointers)."""
    _chk(w, "param")
    _chk(slots, "slots")
    assert w.dtype == torch.float32 and slots.dtype == torch.float32 and slots.dim() == 3 and slots.shape[1] == rows
    call("drn_sgd_step_sharded", w, momentum_shard, slots, slots.shape[0], lib.pvec(packed_ptrs), len(packed_ptrs), int(row0), int(rows),
         w.shape[1], int(c49), float(lr), float(momentum), float(weight_decay), int(bool(nesterov)), int(bool(first_step)),
         current_stream())


def pack_linear_bf16(w, out, c49=0):
    """fp32 [N, K] parameter -> bf16 kernel operand `out` [N, K] (c49 > 0: columns (c, ph, pw) -> (ph, pw, c))."""
    _chk(w, "weight")
    _chk(out, "packed")
    call("drn_pack_linear_bf16", w, out, w.shape[0], w.shape[1], int(c49 or 0), current_stream())
    return out


def to_bf16(x):
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    call("drn_cast_f32_to_bf16", x.contiguous(), out, x.numel(), current_stream())
    return out


def to_f32(x):
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    call("drn_cast_bf16_to_f32", x.contiguous(), out, x.numel(), current_stream())
    return out
