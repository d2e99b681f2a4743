// fp32-accurate GEMMs on the bf16 tensor cores ("fp32_tc" precision): operand splitting and partial-sum reduction.
//
// An fp32 value is the exact sum of three bf16 terms (8 + 8 + 8 mantissa bits): x1 = bf16(x), x2 = bf16(x - x1),
// x3 = bf16(x - x1 - x2).  A dot product sum_k x[k] w[k] is then the sum of the bf16 x bf16 products x_i w_j -- each
// exact in fp32 -- and the six leading ones (i + j <= 4) carry it to ~2^-24 relative.  Laid out as extra K columns they
// turn an fp32 layer into bf16 GEMMs on the existing tcgen05 kernel (fp32 accumulation in TMEM).
//
// The accumulator is the catch: tcgen05.mma adds every K = 16 slice into the fp32 accumulator with TRUNCATION
// (measured: all-positive bf16-exact operands come out low by ~2^-25 per MMA step, profiles/r1_fp32_tc_accumulator_*):
// a chain of K/16 steps is biased by (K/16) 2^-25 relative to the accumulator's magnitude -- 4.7e-5 for fc6's K = 25 088,
// against 1e-7 for an fp32 FMA chain with round-to-nearest.  Hence two kinds of GEMM per layer:
//   * ONE "small" GEMM over the five correction products (x1w2, x2w1, x1w3, x2w2, x3w1; operand planes
//     [x1 | x2 | x1 | x2 | x3] against [w2 | w1 | w3 | w2 | w1]): its accumulator is ~2^-8 of the result, so its
//     truncation error is ~2^-33 of the result however long the chain;
//   * the leading product x1w1 in K-GROUPS of a few hundred elements (<= 36..64 MMA steps each, bias <= ~1e-6), each
//     group its own GEMM launch with its own accumulator;
// and the partial results are summed here in fp32 with round-to-nearest.  This file holds both streaming kernels:
//   drn_f32tc_reduce: y = relu?(sum_i partial_i + bias + residual)
//   drn_f32tc_split : x -> "big" operand [G][rows][Cg] (term x1, one dense matrix per K-group) and "small" operand
//                     [rows][5 C].
// Weights are split the same way once per weight version on the host side (modeling.py).
#include "common.cuh"

namespace drn {
namespace split {

__device__ __forceinline__ void split3(float x, __nv_bfloat16 t[3]) {
  t[0] = __float2bfloat16_rn(x);
  const float r1 = __fsub_rn(x, __bfloat162float(t[0]));  // exact
  t[1] = __float2bfloat16_rn(r1);
  const float r2 = __fsub_rn(r1, __bfloat162float(t[1]));  // exact
  t[2] = __float2bfloat16_rn(r2);
}

__device__ __forceinline__ uint2 pack4(__nv_bfloat16 a, __nv_bfloat16 b, __nv_bfloat16 c, __nv_bfloat16 d) {
  uint2 o;
  o.x = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
  o.y = (uint32_t)__bfloat16_as_ushort(c) | ((uint32_t)__bfloat16_as_ushort(d) << 16);
  return o;
}

// one thread per 4 consecutive columns of one row; partials: [n_in][rows][C] with row pitch ld (>= C)
__global__ void reduce_kernel(const float* __restrict__ partials, int n_in, long long part_stride, int ld,
                              const float* __restrict__ bias, const float4* __restrict__ residual, int relu, long long rows,
                              int C4, float4* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c4 = (int)(i - r * C4);
  const float* p = partials + r * ld + 4 * c4;
  float4 v = *reinterpret_cast<const float4*>(p);
  for (int k = 1; k < n_in; ++k) {  // fixed order, round-to-nearest adds
    const float4 s = *reinterpret_cast<const float4*>(p + (long long)k * part_stride);
    v.x = __fadd_rn(v.x, s.x);
    v.y = __fadd_rn(v.y, s.y);
    v.z = __fadd_rn(v.z, s.z);
    v.w = __fadd_rn(v.w, s.w);
  }
  if (bias) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    v.x = __fadd_rn(v.x, b.x);
    v.y = __fadd_rn(v.y, b.y);
    v.z = __fadd_rn(v.z, b.z);
    v.w = __fadd_rn(v.w, b.w);
  }
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
  y[i] = v;
}

// one thread per 4 consecutive channels of one row
__global__ void split_kernel(const float4* __restrict__ x, long long rows, int C4, int Cg4, uint2* __restrict__ big,
                             uint2* __restrict__ small) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C4) return;
  const long long r = i / C4;
  const int c4 = (int)(i - r * C4);
  const float4 v = x[i];
  __nv_bfloat16 a[3], b[3], c[3], d[3];
  split3(v.x, a);
  split3(v.y, b);
  split3(v.z, c);
  split3(v.w, d);
  // big operand: term x1, one dense [rows][Cg] matrix per K-group
  const int g = c4 / Cg4, cg = c4 - g * Cg4;
  big[((long long)g * rows + r) * Cg4 + cg] = pack4(a[0], b[0], c[0], d[0]);
  // small operand: planes [x1 | x2 | x1 | x2 | x3]
  uint2* row = small + r * 5ll * C4 + c4;
  const uint2 t0 = pack4(a[0], b[0], c[0], d[0]), t1 = pack4(a[1], b[1], c[1], d[1]), t2 = pack4(a[2], b[2], c[2], d[2]);
  row[0] = t0;
  row[(long long)C4] = t1;
  row[2ll * C4] = t0;
  row[3ll * C4] = t1;
  row[4ll * C4] = t2;
}

}  // namespace split
}  // namespace drn

extern "C" int drn_f32tc_reduce(const float* partials, int n_in, int64_t part_stride, int ld, const float* bias,
                                const float* residual, int relu, int64_t rows, int C, float* y, drn_stream_t stream) {
  using namespace drn::split;
  DRN_CHECK_ARG(partials && y, "f32tc_reduce: null pointer");
  DRN_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0 && ld >= C && ld % 4 == 0 && n_in >= 1 && part_stride % 4 == 0,
                "f32tc_reduce: rows=%lld C=%d ld=%d n_in=%d", (long long)rows, C, ld, n_in);
  DRN_CHECK_ARG(((uintptr_t)partials % 16 == 0) && ((uintptr_t)bias % 16 == 0) && ((uintptr_t)residual % 16 == 0) &&
                    ((uintptr_t)y % 16 == 0), "f32tc_reduce: operands must be 16-byte aligned");
  if (rows == 0) return 0;
  const long long n = rows * (C / 4);
  DRN_CHECK_ARG((n + 255) / 256 < (1ll << 31), "f32tc_reduce: tensor too large");
  reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(partials, n_in, part_stride, ld, bias,
                                                                                (const float4*)residual, relu, rows, C / 4, (float4*)y);
  DRN_CHECK_LAUNCH("f32tc_reduce");
  return 0;
}

extern "C" int drn_f32tc_split(const float* x, int64_t rows, int C, int Cg, void* big_bf16, void* small_bf16,
                               drn_stream_t stream) {
  using namespace drn::split;
  DRN_CHECK_ARG(x && big_bf16 && small_bf16, "f32tc_split: null pointer");
  DRN_CHECK_ARG(rows >= 0 && C > 0 && Cg > 0 && Cg % 4 == 0 && C % Cg == 0, "f32tc_split: rows=%lld C=%d Cg=%d", (long long)rows, C, Cg);
  DRN_CHECK_ARG(((uintptr_t)x % 16 == 0) && ((uintptr_t)big_bf16 % 8 == 0) && ((uintptr_t)small_bf16 % 8 == 0),
                "f32tc_split: operands must be 16-byte aligned");
  if (rows == 0) return 0;
  const long long n = rows * (C / 4);
  DRN_CHECK_ARG((n + 255) / 256 < (1ll << 31), "f32tc_split: tensor too large");
  split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)x, rows, C / 4, Cg / 4, (uint2*)big_bf16,
                                                                               (uint2*)small_bf16);
  DRN_CHECK_LAUNCH("f32tc_split");
  return 0;
}
