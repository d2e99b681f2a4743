"""Per-kernel parity tests of libdrn_b200.so against the CPU oracle (run on the B200: -m gpu).
Every call goes through the C ABI (drn_wsod_pytorch_b200.ops -> ctypes -> libdrn_b200.so)."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers
from drn_wsod_pytorch_b200 import ops
from oracle import wsl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _c_roipool(feat_chw, boxes, scale):
    path = os.path.join(helpers.ROOT, "oracle", "_build", "libroipool_ref.so")
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build()
    lib = ctypes.CDLL(path)
    C, h, w = feat_chw.shape
    R = boxes.shape[0]
    out = np.empty((R, C, 7, 7), dtype=np.float32)
    f = np.ascontiguousarray(feat_chw, dtype=np.float32)
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    lib.roipool_ref(f.ctypes.data_as(ctypes.c_void_p), C, h, w, b.ctypes.data_as(ctypes.c_void_p), R,
                    ctypes.c_float(scale), 7, out.ctypes.data_as(ctypes.c_void_p))
    return out


def _edge_boxes():
    return [[4, 4, 20, 20], [12, 12, 12, 12], [36, 28, 44, 36], [-100, -100, -50, -50], [5000, 5000, 6000, 6000],
            [0, 0, 10000, 10000], [4.0, 12.0, 4.0, 100.0], [100, 3.9999, 101, 4.0001], [-30, 10, 40, 60]]


@pytest.mark.parametrize("tables", [False, True])
@pytest.mark.parametrize("C,h,w,R", [(64, 19, 27, 300), (512, 74, 124, 64), (2048, 18, 18, 32), (32, 40, 250, 64)])
def test_roipool_fp32_bit_exact(C, h, w, R, tables):
    rng = np.random.default_rng(C + R)
    feat = rng.standard_normal((C, h, w)).astype(np.float32)
    H, W = h * 8, w * 8
    x0 = rng.uniform(-10, W - 20, R); y0 = rng.uniform(-10, H - 20, R)
    boxes = np.stack([x0, y0, x0 + rng.uniform(0, W, R), y0 + rng.uniform(0, H, R)], 1).astype(np.float32)
    boxes = np.concatenate([boxes, np.asarray(_edge_boxes(), dtype=np.float32)], 0)
    obj = rng.uniform(0, 1, len(boxes)).astype(np.float32)
    ref = _c_roipool(feat, boxes, 0.125) * (obj + np.float32(1.0))[:, None, None, None]
    tv = O.roi_pool(torch.from_numpy(feat)[None], torch.from_numpy(boxes), 0.125).numpy() * (obj + np.float32(1.0))[:, None, None, None]
    assert np.array_equal(ref, tv), "C restatement disagrees with torchvision"
    f_hwc = torch.from_numpy(feat).permute(1, 2, 0).contiguous().to(DEV)
    out = ops.roipool(f_hwc, torch.from_numpy(boxes).to(DEV), torch.from_numpy(obj).to(DEV), 0.125, use_tables=tables)
    got = out.view(len(boxes), 49, C).permute(0, 2, 1).reshape(len(boxes), C, 7, 7).cpu().numpy()
    assert np.array_equal(got, ref)  # max-pool is exact; one fp32 multiply by (objectness+1)


@pytest.mark.parametrize("tables", [False, True])
def test_roipool_bf16_bit_exact(tables):
    rng = np.random.default_rng(5)
    C, h, w, R = 256, 37, 50, 200
    feat = torch.from_numpy(rng.standard_normal((C, h, w)).astype(np.float32)).bfloat16()
    x0 = rng.uniform(0, 300, R); y0 = rng.uniform(0, 200, R)
    boxes = np.stack([x0, y0, x0 + rng.uniform(0, 300, R), y0 + rng.uniform(0, 200, R)], 1).astype(np.float32)
    obj = rng.uniform(0, 1, R).astype(np.float32)
    pooled = O.roi_pool(feat.float()[None], torch.from_numpy(boxes), 0.125)
    ref = (pooled * (torch.from_numpy(obj) + 1).view(-1, 1, 1, 1)).bfloat16()
    out = ops.roipool(feat.permute(1, 2, 0).contiguous().to(DEV), torch.from_numpy(boxes).to(DEV),
                      torch.from_numpy(obj).to(DEV), 0.125, use_tables=tables)
    got = out.view(R, 49, C).permute(0, 2, 1).reshape(R, C, 7, 7).cpu()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("tables", [False, True])
def test_roipool_empty_and_constant_properties(tables):
    f = torch.full((10, 12, 64), 3.5, device=DEV)
    boxes = torch.tensor([[0.0, 0, 50, 50], [20, 20, 90, 70]], device=DEV)
    out = ops.roipool(f, boxes, None, 0.125, use_tables=tables)
    assert torch.all(out == 3.5)  # max over a constant map is the constant (idempotence)
    out0 = ops.roipool(f, boxes[:0].contiguous(), None, 0.125, use_tables=tables)
    assert out0.shape == (0, 49 * 64)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_roipool_full_size_tables_equal_direct_scan(dtype):
    """BASELINE.json full size (74x124 map, C=2048 / 512, R=4000 / 2000): the range-max-table path and
    the direct per-cell scan agree bit for bit (size-independent property: max is order-free)."""
    from drn_wsod_pytorch_b200 import synth
    C, R = (2048, 4000) if dtype == torch.bfloat16 else (512, 2000)
    g = torch.Generator().manual_seed(1)
    f = torch.randn(74, 124, C, generator=g).to(dtype).to(DEV)
    inp = synth.make_inputs(600, 1000, R, seed=3)
    boxes, obj = inp["boxes"].to(DEV), inp["objectness"].to(DEV)
    a = ops.roipool(f, boxes, obj, 0.125, use_tables=False)
    b = ops.roipool(f, boxes, obj, 0.125, use_tables=True)
    assert torch.equal(a, b)
    assert torch.isfinite(b.float()).all()


@pytest.mark.parametrize("cin,cout,k,dil,res,relu", [(64, 64, 3, 1, False, True), (64, 128, 3, 2, True, True),
                                                      (128, 64, 1, 1, True, False), (16, 64, 3, 1, False, False)])
def test_conv_simt_fp32(cin, cout, k, dil, res, relu):
    g = torch.Generator().manual_seed(cin * 7 + cout + k + dil)
    N, H, W = 2, 21, 35
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g)
    r = torch.randn(N, cout, H, W, generator=g) if res else None
    ref = F.conv2d(x, w, None, padding=dil * (k // 2), dilation=dil) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    if res:
        ref = ref + r
    if relu:
        ref = F.relu(ref)
    packed = {"w": w.permute(2, 3, 1, 0).reshape(-1, cout).contiguous().to(DEV), "scale": scale.to(DEV), "bias": bias.to(DEV), "cout": cout}
    out = ops.conv_f32(x.permute(0, 2, 3, 1).contiguous().to(DEV), packed, k, dil, relu,
                       r.permute(0, 2, 3, 1).contiguous().to(DEV) if res else None)
    got = out.permute(0, 3, 1, 2).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("stride,canvas", [(2, None), (1, None), (2, (45, 70))])
def test_first_conv_fused_normalise(stride, canvas):
    g = torch.Generator().manual_seed(stride)
    H, W, cout = 37, 61, 64
    img = torch.rand(3, H, W, generator=g) * 255
    mean, std = [102.9801, 115.9465, 122.7717], [1.0, 1.0, 1.0]
    w = torch.randn(cout, 3, 3, 3, generator=g) * 0.1
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g)
    x = (img - torch.tensor(mean).view(3, 1, 1)) / torch.tensor(std).view(3, 1, 1)
    cv = canvas or (H, W)
    x = F.pad(x, (0, cv[1] - W, 0, cv[0] - H))
    ref = F.relu(F.conv2d(x[None], w, None, stride=stride, padding=1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    packed = {"w": w.permute(2, 3, 1, 0).reshape(-1, cout).contiguous().to(DEV), "scale": scale.to(DEV), "bias": bias.to(DEV), "cout": cout}
    out = ops.first_conv(img.to(DEV), cv, mean, std, packed, stride)
    torch.testing.assert_close(out.permute(0, 3, 1, 2).cpu(), ref, rtol=1e-4, atol=2e-3)


@pytest.mark.parametrize("stride,dtype", [(2, torch.float32), (1, torch.float32), (1, torch.bfloat16), (2, torch.bfloat16)])
def test_maxpool_exact(stride, dtype):
    x = torch.randn(2, 64, 23, 31, generator=torch.Generator().manual_seed(1)).to(dtype)
    ref = F.max_pool2d(x.float(), 2, stride).to(dtype)
    out = ops.maxpool2x2(x.permute(0, 2, 3, 1).contiguous().to(DEV), stride)
    assert torch.equal(out.permute(0, 3, 1, 2).cpu(), ref)


# ------------------------------------------------------------------ tcgen05 path
def _bf16_ref_gemm(a, w, scale, bias, res, relu):
    y = a.float() @ w.float().t()
    if scale is not None:
        y = y * scale
    y = y + bias
    if res is not None:
        y = y + res.float()
    return F.relu(y) if relu else y


@pytest.mark.parametrize("M,K,N,relu,res,f32out", [(128, 64, 64, False, False, True), (300, 256, 128, True, False, False),
                                                    (2000, 1024, 512, True, True, False), (77, 128, 104, False, False, True),
                                                    (1000, 4096, 128, False, False, True)])
def test_tc_gemm_bf16(M, K, N, relu, res, f32out):
    g = torch.Generator().manual_seed(M + K + N)
    a = (torch.randn(M, K, generator=g)).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    scale = torch.rand(N, generator=g) + 0.5
    bias = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g).bfloat16() if res else None
    if res:  # the shortcut is accumulated by the MMA: per-channel scale must be folded into the weights
        w, scale = (w.float() * scale[:, None]).bfloat16(), None
    ref = _bf16_ref_gemm(a, w, scale, bias, r, relu)
    packed = {"w": w.to(DEV), "scale": scale.to(DEV) if scale is not None else None, "bias": bias.to(DEV), "cout": N}
    out = ops.conv_bf16_tc(a.view(1, M, 1, K).to(DEV), packed, 1, 1, relu, r.view(1, M, 1, N).to(DEV) if res else None,
                           out_dtype=torch.float32 if f32out else torch.bfloat16)
    got = out.view(M, N).float().cpu()
    # bf16 operands, fp32 accumulation: only summation order (and the bf16 output rounding) differ
    torch.testing.assert_close(got, ref, rtol=1e-2 if not f32out else 2e-3, atol=1e-2 if not f32out else 2e-3)


@pytest.mark.parametrize("cin,cout,dil,H,W,res", [(64, 64, 1, 16, 32, False), (128, 256, 2, 21, 35, True), (64, 128, 1, 74, 124, False),
                                                   (256, 64, 2, 9, 17, True), (512, 512, 2, 74, 124, True), (64, 64, 1, 5, 300, True),
                                                   (64, 64, 1, 150, 40, False),
                                                   # rows that fill a 128-pixel tile and no shortcut: the HALO variant (three input rows
                                                   # staged once per channel block, nine taps through shifted UMMA descriptors)
                                                   (256, 256, 2, 74, 124, False), (128, 128, 1, 75, 125, False), (64, 64, 1, 12, 500, False),
                                                   (64, 192, 4, 20, 100, False), (256, 64, 2, 37, 250, False), (192, 320, 1, 3, 127, False)])
def test_tc_conv3x3_bf16(cin, cout, dil, H, W, res):
    g = torch.Generator().manual_seed(cin + cout + dil + H)
    N = 2
    x = torch.randn(N, cin, H, W, generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).bfloat16()
    scale = torch.rand(cout, generator=g) + 0.5
    bias = torch.randn(cout, generator=g)
    r = torch.randn(N, cout, H, W, generator=g).bfloat16() if res else None
    if res:  # shortcut accumulated by the MMA: scale folded into the filter
        w, scale = (w.float() * scale.view(-1, 1, 1, 1)).bfloat16(), None
    ref = F.conv2d(x.float(), w.float(), None, padding=dil, dilation=dil)
    if scale is not None:
        ref = ref * scale.view(1, -1, 1, 1)
    ref = ref + bias.view(1, -1, 1, 1)
    if res:
        ref = ref + r.float()
    ref = F.relu(ref)
    packed = {"w": w.permute(0, 2, 3, 1).reshape(cout, -1).contiguous().to(DEV),
              "scale": scale.to(DEV) if scale is not None else None, "bias": bias.to(DEV), "cout": cout}
    out = ops.conv_bf16_tc(x.permute(0, 2, 3, 1).contiguous().to(DEV), packed, 3, dil, True,
                           r.permute(0, 2, 3, 1).contiguous().to(DEV) if res else None)
    torch.testing.assert_close(out.permute(0, 3, 1, 2).float().cpu(), ref, rtol=1e-2, atol=1e-2)


def test_tc_gemm_fused_dropout_matches_standalone_kernel():
    """The epilogue-fused dropout draws the same counter-based mask as drn_dropout_inplace."""
    M, K, N = 520, 192, 320
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    packed = {"w": (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV), "scale": None,
              "bias": torch.randn(N, generator=g).to(DEV), "cout": N}
    plain = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True).view(M, N)
    fused = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True, dropout_p=0.5, dropout_seed=77).view(M, N)
    ref = ops.dropout_(plain.clone(), 0.5, 77)
    assert torch.equal(fused, ref)
    frac = (fused == 0).float().mean().item()
    assert 0.6 < frac < 0.9  # relu zeros (~50%) + dropped half of the rest


def test_tc_gemm_stream_k_matches_reference_and_is_deterministic(monkeypatch):
    """fc7-shaped GEMM (M=4000, K=2048, N=4096: 256 pair-tiles on 74 CTA pairs = 3.46 waves) takes the stream-K
    schedule: partial tiles are exchanged through the workspace.  Checks values, run-to-run bit-identity (fixed
    reduction order) and that the fused dropout still matches the standalone kernel."""
    import os
    if os.environ.get("DRN_TC_STREAMK") != "1":
        pytest.skip("stream-K is opt-in (DRN_TC_STREAMK=1, read once per process); whole-tile waves are faster for fc6/fc7")
    M, K, N = 4000, 2048, 4096
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    packed = {"w": w, "scale": None, "bias": bias, "cout": N}
    ref = F.relu(a.float() @ w.float().t() + bias)
    outs = [ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True).view(M, N) for _ in range(3)]
    torch.testing.assert_close(outs[0].float(), ref, rtol=1e-2, atol=2e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    fused = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True, dropout_p=0.5, dropout_seed=11).view(M, N)
    assert torch.equal(fused, ops.dropout_(outs[0].clone(), 0.5, 11))


@pytest.mark.parametrize("M,K,N", [(4000, 16384, 2048), (2000, 25088, 4096), (4000, 16448, 2048)])
def test_tc_gemm_tail_split_k(M, K, N):
    """fc6-shaped GEMMs (128 pair-tiles on 74 CTA pairs = 1.73 waves) take the tail split-K schedule: the 54 tiles
    of the partial wave are cut into 4 K-ranges whose fp32 partials meet in the workspace.  Checks values against a
    torch matmul of the same bf16 operands, agreement with the whole-tile schedule (summation order only), run-to-run
    bit-identity (fixed fold order, self-resetting flags) and the fused dropout mask."""
    from drn_wsod_pytorch_b200 import lib
    g = torch.Generator().manual_seed(K)
    a = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    packed = {"w": w, "scale": None, "bias": bias, "cout": N}
    ref = F.relu(a.float() @ w.float().t() + bias)
    L = lib.load()
    prev = L.drn_gemm_set_tail_split(1)
    prev_ws, ops.SPLIT_K_WORKSPACE = ops.SPLIT_K_WORKSPACE, True
    try:
        outs = [ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True).view(M, N) for _ in range(3)]
        fused = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True, dropout_p=0.5, dropout_seed=11).view(M, N)
        L.drn_gemm_set_tail_split(0)
        whole = ops.conv_bf16_tc(a.view(1, M, 1, K), packed, 1, 1, True).view(M, N)
    finally:
        L.drn_gemm_set_tail_split(prev)
        ops.SPLIT_K_WORKSPACE = prev_ws
    torch.testing.assert_close(outs[0].float(), ref, rtol=1e-2, atol=2e-2)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    torch.testing.assert_close(outs[0].float(), whole.float(), rtol=1e-2, atol=1e-2)
    assert not torch.equal(outs[0], whole) or M * N < 148 * 128 * 256  # the split really happened (different rounding somewhere)
    assert torch.equal(fused, ops.dropout_(outs[0].clone(), 0.5, 11))


@pytest.mark.parametrize("M,K,N", [(9176, 512, 2048), (37500, 64, 256), (4000, 4096, 4096)])
def test_tc_gemm_bf16_large_residual(M, K, N):
    """Backbone-sized 1x1 convs with residual: exercises multi-tile-per-CTA staging ring + residual prefetch."""
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g).bfloat16()
    ad, wd, rd = a.to(DEV), w.to(DEV), r.to(DEV)
    ref = F.relu(ad.float() @ wd.float().t() + bias.to(DEV) + rd.float()).cpu()
    packed = {"w": wd, "scale": None, "bias": bias.to(DEV), "cout": N}
    out = ops.conv_bf16_tc(ad.view(1, M, 1, K), packed, 1, 1, True, rd.view(1, M, 1, N))
    torch.testing.assert_close(out.view(M, N).float().cpu(), ref, rtol=1e-2, atol=2e-2)


# ------------------------------------------------------------------ heads
def _rand_logits(R, K, S, seed):
    g = torch.Generator().manual_seed(seed)
    ld = ((2 * K + S * (K + 1) + 63) // 64) * 64
    return torch.randn(R, ld, generator=g) * 2.0, ld


@pytest.mark.parametrize("R,K", [(64, 20), (2000, 20), (4000, 80), (1, 20)])
def test_wsddn_mil(R, K):
    logits, ld = _rand_logits(R, K, 3, R + K)
    oh = torch.zeros(K); oh[[3, 7]] = 1
    s_ref = F.softmax(logits[:, :K], 1) * F.softmax(logits[:, K:2 * K], 0)
    img_ref = torch.clamp(s_ref.sum(0), 1e-6, 1 - 1e-6)
    loss_ref = F.binary_cross_entropy(img_ref[None], oh[None], reduction="mean")
    loss = torch.zeros(1, device=DEV)
    scores, img = ops.wsddn_mil(logits.to(DEV), K, 0, K, oh.to(DEV), True, 1.0, loss)
    torch.testing.assert_close(scores.cpu(), s_ref, rtol=1e-4, atol=1e-12)
    torch.testing.assert_close(img.cpu(), img_ref, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(loss.cpu()[0], loss_ref, rtol=1e-5, atol=1e-7)


def test_oicr_stage_chain_exact_indices():
    """pgt argmax (ties -> lowest index), IoU/Matcher labels and the weighted CE against the oracle."""
    R, K, S = 1500, 20, 3
    spec = O.Spec(num_classes=K)
    inp = helpers.synth.make_inputs(600, 1000, R, seed=3, num_gt=3)
    logits, ld = _rand_logits(R, K, S, 9)
    boxes = inp["boxes"]
    gt_int = torch.unique(inp["gt_classes"])
    scores = F.softmax(logits[:, :K], 1) * F.softmax(logits[:, K:2 * K], 0)
    scores[5, gt_int[0]] = scores[:, gt_int[0]].max() * 2  # plant an exact tie: rows 5 and 900
    scores[900, gt_int[0]] = scores[5, gt_int[0]]
    img = torch.clamp(scores.sum(0, keepdim=True), 1e-6, 1 - 1e-6)
    d_logits, d_boxes = logits.to(DEV), boxes.to(DEV)
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    prev_ref, prev_boxes_ref = scores, boxes
    prev = scores.to(DEV)
    for k in range(S):
        idx_r, sc_r, bx_r, w_r = O.get_pgt(prev_ref, prev_boxes_ref, gt_int, img, k, spec)
        lab_r, mi_r = O.label_proposals(boxes, bx_r, gt_int, spec)
        col = 2 * K + k * (K + 1)
        lg = logits[:, col:col + K + 1]
        loss_r, wts_r = O.oicr_stage_loss(lg, lab_r, torch.index_select(w_r, 0, mi_r))
        idx, sc, bx, w = ops.oicr_pgt(prev, d_boxes, gt_int.to(DEV), img[0].to(DEV), k > 0, None, 0, False, spec.bbox_reg_weights)
        lab, mi, cnt = ops.label_proposals(d_boxes, bx, gt_int.to(DEV), K, spec.iou_thresholds, spec.iou_labels)
        loss = torch.zeros(1, device=DEV)
        probs, stats, wts = ops.oicr_stage(d_logits, col, K, lab, mi, w, 1.0, loss, counter)
        assert torch.equal(idx.cpu(), idx_r)
        if k == 0:
            assert idx_r[0].item() == 5  # the planted tie resolves to the lowest index
        assert torch.equal(bx.cpu(), bx_r)  # incl. the re-derived (apply_deltas with zero deltas) boxes of k >= 1
        assert torch.equal(lab.cpu(), lab_r) and torch.equal(mi.cpu(), mi_r)
        assert cnt.cpu().tolist() == [int((lab_r < K).sum()), int((lab_r == K).sum()), 0]
        torch.testing.assert_close(wts.cpu(), wts_r, rtol=0, atol=0)
        torch.testing.assert_close(loss.cpu()[0], loss_r, rtol=1e-5, atol=1e-8)
        torch.testing.assert_close(probs.cpu(), F.softmax(lg, -1), rtol=1e-5, atol=1e-9)
        assert counter.item() == 0
        prev_ref = F.softmax(lg, -1)
        prev_boxes_ref = O.apply_deltas(torch.zeros(R, 4 * K), boxes, spec.bbox_reg_weights)
        prev = probs


@pytest.mark.parametrize("R,K,G", [(4000, 20, 2), (1500, 80, 5), (300, 20, 1)])
def test_fused_tail_is_bit_identical_to_the_per_function_kernels(R, K, G):
    """drn_wsddn_mil_pgt_fwd + drn_oicr_stage_fused_fwd (4 launches) against the one-kernel-per-reference-function
    chain (13 launches): every output must be bit-identical (same device functions, same reduction order)."""
    S = 3
    inp = helpers.synth.make_inputs(600, 1000, R, seed=R + K, num_gt=G, num_classes=K)
    logits, ld = _rand_logits(R, K, S, R)
    d_logits, boxes = logits.to(DEV), inp["boxes"].to(DEV)
    gtb, gtc = inp["gt_boxes"].to(DEV), inp["gt_classes"].to(DEV)
    gt_int = torch.unique(inp["gt_classes"]).to(DEV)
    oh = torch.zeros(K, device=DEV); oh[gt_int] = 1
    counter = torch.zeros(1, dtype=torch.int32, device=DEV)
    thr, labs, bw = [0.5], [0, 1], (10.0, 10.0, 5.0, 5.0)
    # --- per-function chain
    la = torch.zeros(1 + S, device=DEV)
    lab0, midx0, cnt0 = ops.label_proposals(boxes, gtb, gtc, K, thr, labs)
    scores_a, img_a = ops.wsddn_mil(d_logits, K, 0, K, oh, True, 1.0, la[0:1])
    ref_stages, prev = [], scores_a
    for k in range(S):
        pgt = ops.oicr_pgt(prev, boxes, gt_int, img_a, k > 0, None, 0, False, bw)
        lab, mi, cnt = ops.label_proposals(boxes, pgt[2], gt_int, K, thr, labs)
        probs, stats, wts = ops.oicr_stage(d_logits, 2 * K + k * (K + 1), K, lab, mi, pgt[3], 1.0, la[k + 1:k + 2], counter)
        ref_stages.append((pgt, lab, mi, cnt, probs, stats, wts))
        prev = probs
    # --- fused chain
    lb = torch.zeros(1 + S, device=DEV)
    scores_b, img_b, pgt = ops.wsddn_mil_pgt(d_logits, K, 0, K, oh, True, 1.0, lb[0:1], boxes, gt_int, counter)
    assert torch.equal(scores_a, scores_b) and torch.equal(img_a, img_b)
    for k in range(S):
        nxt = None if k == S - 1 else dict(img_score=img_b, deltas=None, ld_deltas=0, cls_agnostic=False, bbox_w=bw)
        o = ops.oicr_stage_fused(d_logits, 2 * K + k * (K + 1), K, boxes, gt_int, pgt[2], pgt[3], thr, labs, 1.0, lb[k + 1:k + 2],
                                 counter, first_gt=(gtb, gtc) if k == 0 else None, nxt=nxt)
        rp, lab, mi, cnt, probs, stats, wts = ref_stages[k]
        for a, b in zip(pgt, rp):
            assert torch.equal(a, b), f"stage {k} pseudo GT differs"
        assert torch.equal(o["labels"], lab) and torch.equal(o["matched"], mi) and torch.equal(o["counts"], cnt)
        assert torch.equal(o["probs"], probs) and torch.equal(o["stats"], stats) and torch.equal(o["weights"], wts)
        if k == 0:
            assert torch.equal(o["first"][0], lab0) and torch.equal(o["first"][1], midx0) and torch.equal(o["first"][2], cnt0)
        pgt = o["next"]
    assert torch.equal(la, lb) and counter.item() == 0
    # --- all stages in two launches (drn_oicr_stages_fwd): the next pseudo GT depends on a stage's logits only
    lc = torch.zeros(1 + S, device=DEV)
    counters = torch.zeros(16, dtype=torch.int32, device=DEV)
    scores_c, img_c, pgt0 = ops.wsddn_mil_pgt(d_logits, K, 0, K, oh, True, 1.0, lc[0:1], boxes, gt_int, counter)
    for rep in range(2):  # twice: the counters reset themselves
        sts, first = ops.oicr_stages(d_logits, [2 * K + k * (K + 1) for k in range(S)], [-1] * S, [bw] * S, K, boxes, gt_int, img_c,
                                     False, pgt0, thr, labs, 1.0, lc, [1 + k for k in range(S)], counters, first_gt=(gtb, gtc))
        for k in range(S):
            rp, lab, mi, cnt, probs, stats, wts = ref_stages[k]
            o = sts[k]
            for a, b in zip(o["pgt"], rp):
                assert torch.equal(a, b), f"stage {k} pseudo GT differs (two-launch form)"
            assert torch.equal(o["labels"], lab) and torch.equal(o["matched"], mi) and torch.equal(o["counts"], cnt)
            assert torch.equal(o["probs"], probs) and torch.equal(o["stats"], stats) and torch.equal(o["weights"], wts)
        assert torch.equal(first[0], lab0) and torch.equal(first[1], midx0) and torch.equal(first[2], cnt0)
        assert torch.equal(la, lc) and int(counters.abs().sum()) == 0


def test_label_proposals_no_gt_and_ignore_band():
    boxes = helpers.synth.make_inputs(300, 400, 500, seed=1)["boxes"]
    lab, mi, cnt = ops.label_proposals(boxes.to(DEV), None, None, 20, [0.5], [0, 1])
    assert torch.all(lab == 20) and cnt.cpu().tolist() == [0, 500, 0]
    spec = O.Spec(iou_thresholds=[0.1, 0.5], iou_labels=[-1, 0, 1])
    gtb = torch.tensor([[10.0, 10, 200, 220], [150, 60, 390, 290]])
    gtc = torch.tensor([4, 9])
    lab_r, mi_r = O.label_proposals(boxes, gtb, gtc, spec)
    lab, mi, cnt = ops.label_proposals(boxes.to(DEV), gtb.to(DEV), gtc.to(DEV), 20, [0.1, 0.5], [-1, 0, 1])
    assert torch.equal(lab.cpu(), lab_r) and torch.equal(mi.cpu(), mi_r)
    assert cnt.cpu()[2].item() == int((lab_r == -1).sum()) > 0


def test_oicr_infer_mean_softmax():
    R, K, S = 700, 20, 3
    logits, ld = _rand_logits(R, K, S, 4)
    boxes = helpers.synth.make_inputs(300, 400, R, seed=2)["boxes"]
    cols = [2 * K + k * (K + 1) for k in range(S)]
    ref = sum(F.softmax(logits[:, c:c + K + 1], -1) for c in cols) / S
    sc, bx = ops.oicr_infer(logits.to(DEV), K, cols, [-1] * S, boxes.to(DEV), [10.0, 10.0, 5.0, 5.0], K)
    torch.testing.assert_close(sc.cpu(), ref, rtol=1e-5, atol=1e-9)
    ref_b = O.apply_deltas(torch.zeros(R, 4 * K), boxes, [10.0, 10.0, 5.0, 5.0])
    assert torch.equal(bx.cpu(), ref_b)


def test_dropout_statistics():
    x = torch.ones(1 << 20, device=DEV)
    ops.dropout_(x, 0.5, 123)
    kept = (x != 0).float().mean().item()
    assert abs(kept - 0.5) < 5e-3 and torch.all((x == 0) | (x == 2.0))


# ------------------------------------------------------------------ inference tail (threshold + per-class NMS + top-k)
def _check_detections(all_boxes, scores, hw, spec, cap=None):
    K = scores.shape[1] - 1
    topk = spec.detections_per_image
    cap = cap if cap is not None else (topk if topk >= 0 else scores.shape[0] * K)
    b, s, c, r, n = ops.detections(scores.to(DEV), all_boxes.to(DEV), hw, spec.score_thresh_test, spec.nms_thresh_test, cap)
    n = int(n.item())
    rb, rs, rc, rr = O.inference_single_image_exact(all_boxes, scores, hw, spec)
    assert n == len(rs)
    assert torch.equal(c[:n].cpu(), rc) and torch.equal(r[:n].cpu(), rr)       # bit-exact kept (row, class) pairs, in order
    assert torch.equal(s[:n].cpu(), rs) and torch.equal(b[:n].cpu(), rb)       # scores / clipped boxes are copies
    return n


@pytest.mark.parametrize("R,K,nreg_k,cluster", [(1500, 20, True, True), (300, 5, False, True), (4000, 20, False, True),
                                                 (4000, 80, True, False), (1, 3, True, True), (33, 1, False, True)])
def test_detections_tail_bit_exact_vs_oracle(R, K, nreg_k, cluster):
    spec = O.Spec(num_classes=K)
    for seed in range(2):
        all_boxes, scores = helpers.rand_dets(R, K, 7 * R + seed, nreg_k, cluster)
        if seed == 1 and R > 10:  # non-finite rows are dropped, ties keep candidate order, degenerate boxes never suppress
            scores[5, 1 % (K + 1)] = float("nan")
            all_boxes[9, 0] = float("inf")
            scores[20:40] = scores[20]
            all_boxes[50:60, 2] = all_boxes[50:60, 0]
        _check_detections(all_boxes, scores, (600, 1000), spec)


def test_detections_tail_edge_cases():
    spec = O.Spec(num_classes=4)
    all_boxes, scores = helpers.rand_dets(200, 4, 3, True, True)
    # nothing above the threshold -> zero detections
    spec_hi = O.Spec(num_classes=4, score_thresh_test=2.0)
    assert _check_detections(all_boxes, scores, (600, 1000), spec_hi) == 0
    # top-k disabled: every kept candidate comes back (capacity R*K)
    spec_all = O.Spec(num_classes=4, detections_per_image=-1)
    n = _check_detections(all_boxes, scores, (600, 1000), spec_all)
    assert n > 100
    # nms threshold 1.0 keeps everything above the score threshold; 0.0 keeps one box per overlapping cluster
    for thr in (1.0, 0.0, 0.5):
        _check_detections(all_boxes, scores, (600, 1000), O.Spec(num_classes=4, nms_thresh_test=thr, detections_per_image=-1))
    # boxes far outside the image clip to zero area and are never suppressed (0/0 -> NaN -> not greater)
    far = all_boxes.clone() + 5000.0
    _check_detections(far, scores, (600, 1000), spec_all)
    # identical boxes and identical scores: first candidate wins
    same_b = all_boxes[:1].repeat(200, 1)
    same_s = scores[:1].repeat(200, 1)
    assert _check_detections(same_b, same_s, (600, 1000), spec_all) == 4
    # empty proposal list
    b, s, c, r, cnt = ops.detections(torch.zeros(0, 5, device=DEV), torch.zeros(0, 16, device=DEV), (600, 1000), 1e-5, 0.3, 100)
    assert int(cnt.item()) == 0


@pytest.mark.parametrize("R,C,c49,ld", [(100, 64, 0, 128), (4000, 49 * 64, 64, 4032), (77, 4096, 0, 128), (333, 49 * 128, 128, 384), (64, 8, 0, 64)])
def test_bf16_transpose_fast_path_matches_torch(R, C, c49, ld):
    """drn_masked_transpose's register-transpose path (plain bf16, the pooled-feature operand of the fc6 weight gradient),
    incl. the bin-major -> (c, ph, pw) row permutation and the zero fill up to the padded K."""
    g = torch.Generator().manual_seed(R + C)
    x = torch.randn(R, C, generator=g).bfloat16().to(DEV)
    out, _ = ops.masked_transpose(x, c49=c49, ld_out=ld)
    ref = x.t()
    if c49:
        ref = x.view(R, 49, c49).permute(2, 1, 0).reshape(C, R)  # row (ch * 49 + bin) <- column (bin * c49 + ch)
    assert out.shape == (C, ld) and torch.equal(out[:, :R], ref) and bool((out[:, R:] == 0).all())
    # the general (masked) kernel agrees on the same input
    mask = torch.ones_like(x)
    out2, masked = ops.masked_transpose(x, mask=mask, c49=c49, ld_out=ld, want_masked=True)
    assert torch.equal(out2, out) and torch.equal(masked, x)


# ---------------------------------------------------------------- "fp32_tc": split-bf16 operands (csrc/drn_split.cu)
@pytest.mark.parametrize("rows,C,Cg", [(1, 64, 64), (37, 128, 64), (1000, 512, 512), (5, 25088, 512)])
def test_f32tc_split_exact(rows, C, Cg):
    """The three bf16 terms sum back to the fp32 value EXACTLY; the big operand is term 1 regrouped [C/Cg][rows][Cg]; the
    small operand holds the planes x1 | x2 | x1 | x2 | x3."""
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * torch.exp(torch.randn(rows, C, generator=g) * 3)).to(DEV)
    big, small = ops.f32tc_split(x, Cg)
    t0 = x.to(torch.bfloat16).float()
    t1 = (x - t0).to(torch.bfloat16).float()
    t2 = (x - t0 - t1).to(torch.bfloat16).float()
    assert torch.equal((t0.double() + t1.double() + t2.double()).float(), x)  # the split itself is exact
    assert big.shape == (C // Cg, rows, Cg) and small.shape == (rows, 5 * C)
    assert torch.equal(big.float().permute(1, 0, 2).reshape(rows, C), t0)
    pl = small.float().view(rows, 5, C)
    for p, t in enumerate((t0, t1, t0, t1, t2)):
        assert torch.equal(pl[:, p], t), p


@pytest.mark.parametrize("n,rows,C", [(1, 3, 8), (5, 1000, 64), (50, 129, 4096)])
def test_f32tc_reduce_matches_torch(n, rows, C):
    g = torch.Generator().manual_seed(n + rows)
    parts = torch.randn(n, rows, C, generator=g).to(DEV)
    bias, res = torch.randn(C, generator=g).to(DEV), torch.randn(rows, C, generator=g).to(DEV)
    want = parts[0].clone()
    for k in range(1, n):
        want = want + parts[k]  # same order, round-to-nearest
    assert torch.equal(ops.f32tc_reduce(parts), want)
    assert torch.equal(ops.f32tc_reduce(parts, bias, res, relu=True), torch.relu(want + bias + res))
    assert torch.equal(ops.f32tc_reduce(parts, bias, None, relu=False), want + bias)


@pytest.mark.parametrize("M,K,N", [(300, 512, 128), (2000, 4096, 256), (129, 25088, 64)])
def test_fp32_tc_linear_accuracy(M, K, N):
    """A linear layer through the split-bf16 GEMMs versus float64: error relative to sum_k |x_k w_k| below 4e-6 (the SIMT
    fp32 kernel measures 2e-7 .. 1.1e-5 on the same layers) and, on all-positive operands (where the tensor core's
    truncating accumulator would show as a bias of (K/16) 2^-25 -- 4.7e-5 for K = 25 088 -- if the leading product ran
    as ONE chain), a mean signed error below 2.5e-6 (<= 64 MMA steps per accumulator)."""
    from drn_wsod_pytorch_b200 import modeling

    g = torch.Generator().manual_seed(M + K)
    for positive in (False, True):
        x = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / K ** 0.5
        b = torch.randn(N, generator=g)
        if positive:
            x, w, b = x.abs(), w.abs(), b.abs()
        packed = modeling.pack_linear([w.to(DEV)], [b.to(DEV)], "fp32_tc")
        y = modeling.run_linear(x.to(DEV), packed, "fp32_tc", relu=False).cpu().double()
        ref = x.double() @ w.double().t() + b.double()
        scale = x.double().abs() @ w.double().abs().t() + b.double().abs()
        err = ((y - ref).abs() / scale).max().item()
        assert err < 4e-6, (positive, err)
        if positive:
            bias_rel = ((y - ref) / ref).mean().item()
            assert abs(bias_rel) < 2.5e-6, bias_rel
            # plain bf16 on the same layer is orders of magnitude further away
            pb = modeling.pack_linear([w.to(DEV)], [b.to(DEV)], "bf16")
            yb = modeling.run_linear(x.to(DEV).to(torch.bfloat16), pb, "bf16", relu=False, out_dtype=torch.float32).cpu().double()
            assert ((yb - ref).abs() / scale).max().item() > 100 * err


@pytest.mark.parametrize("cin,cout,dil,H,W,res", [(64, 64, 1, 40, 56, False), (128, 128, 1, 19, 31, True), (512, 512, 2, 20, 28, True)])
def test_fp32_tc_conv3x3_accuracy(cin, cout, dil, H, W, res):
    """3x3 conv (+ FrozenBN-style bias, shortcut, ReLU) through the split-bf16 GEMMs versus torch float64."""
    from drn_wsod_pytorch_b200 import modeling

    g = torch.Generator().manual_seed(cin + H)
    conv = modeling.Conv2d(cin, cout, 3, dilation=dil, bias=True).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g))
    x = torch.randn(1, cin, H, W, generator=g).abs()
    r = torch.randn(1, cout, H, W, generator=g)
    ref = F.conv2d(x.double(), conv.weight.detach().cpu().double(), conv.bias.detach().cpu().double(), padding=dil, dilation=dil)
    mag = F.conv2d(x.double().abs(), conv.weight.detach().cpu().double().abs(), conv.bias.detach().cpu().double().abs(), padding=dil, dilation=dil)
    if res:
        ref, mag = ref + r.double(), mag + r.double().abs()
    ref = torch.relu(ref)
    y = modeling.run_conv(conv, x.permute(0, 2, 3, 1).contiguous().to(DEV), "fp32_tc", relu=True,
                          residual=r.permute(0, 2, 3, 1).contiguous().to(DEV) if res else None)
    err = ((y.permute(0, 3, 1, 2).cpu().double() - ref).abs() / mag).max().item()
    assert err < 2.5e-6, err
