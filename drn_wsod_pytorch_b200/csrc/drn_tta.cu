// Test-time-augmentation kernels: the image resample and the view merge of
// projects/WSL/wsl/modeling/test_time_augmentation_avg.py, on the device.
//
//  * drn_resample_u8_fwd  = PIL Image.resize(BILINEAR) on 8-bit images (call site
//    detectron2/data/transforms/transform.py:105-109; Pillow src/libImaging/Resample.c): horizontal pass into a
//    uint8 intermediate, then the vertical pass, 22-bit fixed-point coefficients (tables built on the host in double,
//    tta.resample_tables).  Integer arithmetic: bit-exact.  Planar (CHW) in and out, optional horizontal flip
//    (fvcore HFlipTransform.apply_image) and uint8 -> fp32 conversion fused into the last pass.
//  * drn_tta_accumulate   = test_time_augmentation_avg.py:286-309: every view's boxes through the inverse transforms
//    (fvcore Transform.apply_box: corners through apply_coords in fp32, then min / max), running sum over the views,
//    division by the view count on the last one.
//
// Both are HBM-streaming kernels over a few MB per view: one thread per output element, coalesced along x.
#include "common.cuh"

namespace drn {
namespace tta {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

template <typename OutT>
__device__ __forceinline__ void put(OutT* p, uint8_t v);
template <>
__device__ __forceinline__ void put<uint8_t>(uint8_t* p, uint8_t v) { *p = v; }
template <>
__device__ __forceinline__ void put<float>(float* p, uint8_t v) { *p = (float)v; }

// rows x in_w -> rows x out_w (rows = C * H planes flattened): out[r][xx] = clip8(2^21 + sum_k in[r][xmin + k] * kk[xx][k])
template <typename OutT>
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, int rows, int in_w, int out_w, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, int ksize, int flip, OutT* __restrict__ out) {
  const int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (xx >= out_w) return;
  const int xmin = __ldg(bounds + 2 * xx), cnt = __ldg(bounds + 2 * xx + 1);
  const uint8_t* src = in + (size_t)r * in_w + xmin;
  const int* k = kk + (size_t)xx * ksize;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < cnt; ++x) ss += (int)__ldg(src + x) * __ldg(k + x);
  put<OutT>(out + (size_t)r * out_w + (flip ? out_w - 1 - xx : xx), clip8(ss));
}

// planes x in_h x w -> planes x out_h x w
template <typename OutT>
__global__ void resample_v_kernel(const uint8_t* __restrict__ in, int in_h, int out_h, int w, const int* __restrict__ bounds,
                                  const int* __restrict__ kk, int ksize, int flip, OutT* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y, c = blockIdx.z;
  if (x >= w) return;
  const int ymin = __ldg(bounds + 2 * yy), cnt = __ldg(bounds + 2 * yy + 1);
  const uint8_t* src = in + ((size_t)c * in_h + ymin) * w + x;
  const int* k = kk + (size_t)yy * ksize;
  int ss = 1 << (PRECISION_BITS - 1);
  for (int y = 0; y < cnt; ++y) ss += (int)__ldg(src + (size_t)y * w) * __ldg(k + y);
  put<OutT>(out + ((size_t)c * out_h + yy) * w + (flip ? w - 1 - x : x), clip8(ss));
}

// neither pass runs (same size): copy / convert / flip
template <typename OutT>
__global__ void copy_flip_kernel(const uint8_t* __restrict__ in, int rows, int w, int flip, OutT* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (x >= w) return;
  put<OutT>(out + (size_t)r * w + (flip ? w - 1 - x : x), in[(size_t)r * w + x]);
}

template <typename OutT>
int resample(const uint8_t* src, int C, int H, int W, const int* xb, const int* xk, int xks, const int* yb, const int* yk, int yks,
             int new_h, int new_w, uint8_t* tmp, int flip, OutT* out, cudaStream_t st) {
  const int T = 128;
  const bool horiz = new_w != W, vert = new_h != H;
  if (!horiz && !vert) {
    copy_flip_kernel<OutT><<<dim3(cdiv(W, T), C * H), T, 0, st>>>(src, C * H, W, flip, out);
    DRN_CHECK_LAUNCH("tta copy_flip");
    return 0;
  }
  const uint8_t* cur = src;
  if (horiz) {
    if (vert) {
      resample_h_kernel<uint8_t><<<dim3(cdiv(new_w, T), C * H), T, 0, st>>>(src, C * H, W, new_w, xb, xk, xks, 0, tmp);
      cur = tmp;
    } else {
      resample_h_kernel<OutT><<<dim3(cdiv(new_w, T), C * H), T, 0, st>>>(src, C * H, W, new_w, xb, xk, xks, flip, out);
    }
    DRN_CHECK_LAUNCH("tta resample_h");
  }
  if (vert) {
    resample_v_kernel<OutT><<<dim3(cdiv(new_w, T), new_h, C), T, 0, st>>>(cur, H, new_h, new_w, yb, yk, yks, flip, out);
    DRN_CHECK_LAUNCH("tta resample_v");
  }
  return 0;
}

struct Ops {
  int n;
  int kind[DRN_TTA_MAX_OPS];
  float a[DRN_TTA_MAX_OPS], b[DRN_TTA_MAX_OPS];
};

// numpy min / max (NaN propagates)
__device__ __forceinline__ float np_min(float a, float b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ float np_max(float a, float b) { return (a > b || a != a) ? a : b; }

__global__ void accumulate_kernel(const float4* __restrict__ boxes, long long nboxes, const float* __restrict__ scores,
                                  long long nscores, Ops ops, float4* __restrict__ acc_boxes, float* __restrict__ acc_scores,
                                  int first, int last, float n_views) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nboxes) {
    float4 b = boxes[i];
#pragma unroll
    for (int o = 0; o < DRN_TTA_MAX_OPS; ++o) {
      if (o >= ops.n) break;
      float x0 = b.x, y0 = b.y, x1 = b.z, y1 = b.w;
      if (ops.kind[o] == DRN_TTA_OP_RESIZE) {  // coords[:, 0] *= new_w / w ; coords[:, 1] *= new_h / h   (fp32)
        x0 = __fmul_rn(x0, ops.a[o]);
        x1 = __fmul_rn(x1, ops.a[o]);
        y0 = __fmul_rn(y0, ops.b[o]);
        y1 = __fmul_rn(y1, ops.b[o]);
      } else if (ops.kind[o] == DRN_TTA_OP_HFLIP) {  // coords[:, 0] = width - coords[:, 0]
        x0 = __fsub_rn(ops.a[o], x0);
        x1 = __fsub_rn(ops.a[o], x1);
      }
      // corners (x0,y0) (x1,y0) (x0,y1) (x1,y1): min / max over the four
      b.x = np_min(x0, x1);
      b.y = np_min(y0, y1);
      b.z = np_max(x0, x1);
      b.w = np_max(y0, y1);
    }
    if (!first) {
      const float4 a = acc_boxes[i];
      b.x = __fadd_rn(a.x, b.x);
      b.y = __fadd_rn(a.y, b.y);
      b.z = __fadd_rn(a.z, b.z);
      b.w = __fadd_rn(a.w, b.w);
    }
    if (last) {
      b.x = __fdiv_rn(b.x, n_views);
      b.y = __fdiv_rn(b.y, n_views);
      b.z = __fdiv_rn(b.z, n_views);
      b.w = __fdiv_rn(b.w, n_views);
    }
    acc_boxes[i] = b;
  }
  if (i < nscores) {
    float s = scores[i];
    if (!first) s = __fadd_rn(acc_scores[i], s);
    if (last) s = __fdiv_rn(s, n_views);
    acc_scores[i] = s;
  }
}

}  // namespace tta
}  // namespace drn

extern "C" int drn_resample_u8_fwd(const void* src_chw, int C, int H, int W, const int* xbounds, const int* xcoef, int xksize,
                                   const int* ybounds, const int* ycoef, int yksize, int new_h, int new_w, void* tmp, int flip,
                                   void* out_chw, int out_dtype, drn_stream_t stream) {
  using namespace drn::tta;
  DRN_CHECK_ARG(src_chw && out_chw, "resample: null pointer");
  DRN_CHECK_ARG(C > 0 && H > 0 && W > 0 && new_h > 0 && new_w > 0, "resample: bad shape %dx%dx%d -> %dx%d", C, H, W, new_h, new_w);
  DRN_CHECK_ARG((long long)C * H <= 65535 && new_h <= 65535 && C <= 65535, "resample: image too large for the launch grid");
  DRN_CHECK_ARG(new_w == W || (xbounds && xcoef && xksize > 0), "resample: horizontal tables missing");
  DRN_CHECK_ARG(new_h == H || (ybounds && ycoef && yksize > 0), "resample: vertical tables missing");
  DRN_CHECK_ARG(!(new_w != W && new_h != H) || tmp, "resample: two passes need the C x H x new_w intermediate");
  DRN_CHECK_ARG(out_dtype == DRN_F32 || out_dtype == DRN_U8, "resample: out dtype %d", out_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == DRN_F32)
    return resample<float>((const uint8_t*)src_chw, C, H, W, xbounds, xcoef, xksize, ybounds, ycoef, yksize, new_h, new_w,
                           (uint8_t*)tmp, flip, (float*)out_chw, st);
  return resample<uint8_t>((const uint8_t*)src_chw, C, H, W, xbounds, xcoef, xksize, ybounds, ycoef, yksize, new_h, new_w,
                           (uint8_t*)tmp, flip, (uint8_t*)out_chw, st);
}

extern "C" int drn_tta_accumulate(const void* all_boxes, const void* all_scores, int R, int box_cols, int score_cols, int n_ops,
                                  const int* op_kind, const float* op_a, const float* op_b, void* acc_boxes, void* acc_scores,
                                  int view_index, int n_views, drn_stream_t stream) {
  using namespace drn::tta;
  DRN_CHECK_ARG(all_boxes && all_scores && acc_boxes && acc_scores, "tta_accumulate: null pointer");
  DRN_CHECK_ARG(R >= 0 && box_cols > 0 && box_cols % 4 == 0 && score_cols > 0, "tta_accumulate: bad shape R=%d box_cols=%d score_cols=%d",
                R, box_cols, score_cols);
  DRN_CHECK_ARG(n_ops >= 0 && n_ops <= DRN_TTA_MAX_OPS, "tta_accumulate: %d transforms (max %d)", n_ops, DRN_TTA_MAX_OPS);
  DRN_CHECK_ARG(n_ops == 0 || (op_kind && op_a && op_b), "tta_accumulate: transform arrays missing");
  DRN_CHECK_ARG(n_views >= 1 && view_index >= 0 && view_index < n_views, "tta_accumulate: view %d of %d", view_index, n_views);
  Ops ops;
  ops.n = n_ops;
  for (int i = 0; i < DRN_TTA_MAX_OPS; ++i) {
    ops.kind[i] = i < n_ops ? op_kind[i] : DRN_TTA_OP_NOOP;
    ops.a[i] = i < n_ops ? op_a[i] : 0.f;
    ops.b[i] = i < n_ops ? op_b[i] : 0.f;
    DRN_CHECK_ARG(ops.kind[i] == DRN_TTA_OP_NOOP || ops.kind[i] == DRN_TTA_OP_RESIZE || ops.kind[i] == DRN_TTA_OP_HFLIP,
                  "tta_accumulate: transform kind %d", ops.kind[i]);
  }
  if (R == 0) return 0;
  const long long nb = (long long)R * (box_cols / 4), ns = (long long)R * score_cols;
  const long long n = nb > ns ? nb : ns;
  accumulate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)all_boxes, nb, (const float*)all_scores, ns, ops, (float4*)acc_boxes, (float*)acc_scores, view_index == 0,
      view_index == n_views - 1, (float)n_views);
  DRN_CHECK_LAUNCH("tta_accumulate");
  return 0;
}
