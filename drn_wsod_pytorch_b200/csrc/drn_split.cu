// fp32-accurate GEMMs on the bf16 tensor cores ("fp32_tc" precision): operand splitting.
//
// An fp32 value is the exact sum of three bf16 terms (8 + 8 + 8 mantissa bits): x1 = bf16(x), x2 = bf16(x - x1),
// x3 = bf16(x - x1 - x2).  A dot product sum_k x[k] w[k] is then the sum of the bf16 x bf16 products x_i w_j -- each
// exact in fp32 -- and the six leading ones (i + j <= 4: x1w1, x1w2, x2w1, x1w3, x2w2, x3w1) carry it to ~2^-24
// relative.  Laying the terms out as extra K columns,
//     A' = [x_XI[0] | x_XI[1] | ... ]  (rows x P*C),   W' = [w_WI[0] | w_WI[1] | ...]  (cout x P*C per filter tap),
// turns the fp32 layer into ONE bf16 GEMM with K' = P*K on the existing tcgen05 kernel (fp32 accumulation in TMEM),
// instead of the SIMT fp32 kernel the exact-fp32 parity mode uses (30 TFLOP/s).  This file holds the activation side:
// one streaming pass that (optionally) adds the fp32 shortcut, applies ReLU, writes the fp32 result and the P bf16
// planes.  Weights are split once per weight version on the host side (modeling.py).
#include "common.cuh"

namespace drn {
namespace split {

struct Terms {
  int n;
  int idx[DRN_SPLIT_MAX_TERMS];
};

__device__ __forceinline__ void split3(float x, __nv_bfloat16 t[3]) {
  t[0] = __float2bfloat16_rn(x);
  const float r1 = __fsub_rn(x, __bfloat162float(t[0]));  // exact
  t[1] = __float2bfloat16_rn(r1);
  const float r2 = __fsub_rn(r1, __bfloat162float(t[1]));  // exact
  t[2] = __float2bfloat16_rn(r2);
}

// one thread per 4 consecutive channels of one row
__global__ void split_kernel(const float4* __restrict__ x, const float4* __restrict__ residual, int relu, long long rows, int C4,
                             Terms terms, float4* __restrict__ y, uint2* __restrict__ planes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c4 = (int)(i - r * C4);
  float4 v = x[i];
  if (residual) {
    const float4 s = residual[i];
    v.x = __fadd_rn(v.x, s.x);
    v.y = __fadd_rn(v.y, s.y);
    v.z = __fadd_rn(v.z, s.z);
    v.w = __fadd_rn(v.w, s.w);
  }
  if (relu) {
    v.x = fmaxf(v.x, 0.f);
    v.y = fmaxf(v.y, 0.f);
    v.z = fmaxf(v.z, 0.f);
    v.w = fmaxf(v.w, 0.f);
  }
  if (y) y[i] = v;
  __nv_bfloat16 a[3], b[3], c[3], d[3];
  split3(v.x, a);
  split3(v.y, b);
  split3(v.z, c);
  split3(v.w, d);
  uint2* row = planes + r * (long long)terms.n * C4 + c4;
#pragma unroll
  for (int p = 0; p < DRN_SPLIT_MAX_TERMS; ++p) {
    if (p >= terms.n) break;
    const int t = terms.idx[p];
    const __nv_bfloat16 e0 = t == 0 ? a[0] : (t == 1 ? a[1] : a[2]);
    const __nv_bfloat16 e1 = t == 0 ? b[0] : (t == 1 ? b[1] : b[2]);
    const __nv_bfloat16 e2 = t == 0 ? c[0] : (t == 1 ? c[1] : c[2]);
    const __nv_bfloat16 e3 = t == 0 ? d[0] : (t == 1 ? d[1] : d[2]);
    uint2 o;
    o.x = (uint32_t)__bfloat16_as_ushort(e0) | ((uint32_t)__bfloat16_as_ushort(e1) << 16);
    o.y = (uint32_t)__bfloat16_as_ushort(e2) | ((uint32_t)__bfloat16_as_ushort(e3) << 16);
    row[(long long)p * C4] = o;
  }
}

}  // namespace split
}  // namespace drn

extern "C" int drn_split_bf16_terms(const float* x, const float* residual, int relu, int64_t rows, int C, int n_terms,
                                    const int* term_idx, float* y_f32, void* planes_bf16, drn_stream_t stream) {
  using namespace drn::split;
  DRN_CHECK_ARG(x && planes_bf16, "split_bf16_terms: null pointer");
  DRN_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0, "split_bf16_terms: rows=%lld C=%d (C must be a multiple of 4)", (long long)rows, C);
  DRN_CHECK_ARG(n_terms >= 1 && n_terms <= DRN_SPLIT_MAX_TERMS && term_idx, "split_bf16_terms: %d planes (max %d)", n_terms, DRN_SPLIT_MAX_TERMS);
  DRN_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)residual % 16 == 0) && ((uintptr_t)y_f32 % 16 == 0) &&
                    ((uintptr_t)planes_bf16 % 8 == 0), "split_bf16_terms: operands must be 16-byte aligned");
  Terms t;
  t.n = n_terms;
  for (int i = 0; i < DRN_SPLIT_MAX_TERMS; ++i) {
    t.idx[i] = i < n_terms ? term_idx[i] : 0;
    DRN_CHECK_ARG(t.idx[i] >= 0 && t.idx[i] <= 2, "split_bf16_terms: term index %d", t.idx[i]);
  }
  if (rows == 0) return 0;
  const long long n = rows * (C / 4);
  DRN_CHECK_ARG((n + 255) / 256 < (1ll << 31), "split_bf16_terms: tensor too large");
  split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (const float4*)residual, relu, rows, C / 4, t, (float4*)y_f32, (uint2*)planes_bf16);
  DRN_CHECK_LAUNCH("split_bf16_terms");
  return 0;
}
