// Backward of the trainable tail: gradients of the WSDDN / OICR losses with respect to the
// head logits, and the data-movement kernels that turn the fc backward into plain K-major GEMMs for the existing
// tensor-core / SIMT kernels (masked transposes, column sums).  What torch autograd derives for the reference
// (projects/WSL/wsl/modeling/roi_heads/fast_rcnn.py:317-329 BCE of the clamped image score, :493-527 dual softmax,
// :1128-1144 weighted CE, :1146-1211 smooth-L1; roi_heads_oicr.py:359-394: image scores, pseudo GT, proposal
// weights and next-stage inputs are all detached, so each loss only reaches its own head's logits).
#include "common.cuh"

#include <math.h>

namespace drn {

// ---- WSDDN MIL -------------------------------------------------------------------------------------------
// s_rk = a_rk * b_rk, a = softmax_k(cls_r.), b = softmax_r(det_.k);  p_k = clamp(S_k = sum_r s_rk, 1e-6, 1-1e-6)
// L = up * scale * sum_k BCE(p_k, y_k)  (scale holds 1/K for MEAN_LOSS and 1/N, 1/N^2)
// g_k = dL/dS_k = up * scale * (p_k - y_k) / (p_k (1 - p_k)) if 1e-6 <= S_k <= 1-1e-6 else 0      (torch.clamp)
// d det_rk = g_k (s_rk - b_rk S_k);   d cls_rk = g_k s_rk - a_rk T_r,  T_r = sum_j g_j s_rj
//
// pass 1: one CTA per class: column softmax statistics, S_k, g_k; writes d det and g.
__global__ void __launch_bounds__(512)
mil_bwd_cols_kernel(const float* __restrict__ logits, int ld, int R, int K, int det_off, const float* __restrict__ scores,
                    const float* __restrict__ gt_onehot, float scale, const float* __restrict__ up,
                    float* __restrict__ dlogits, float* __restrict__ g_out) {
  __shared__ float sh[32];
  const int c = blockIdx.x;
  const float* det = logits + det_off + c;
  float m = -INFINITY;
  for (int r = threadIdx.x; r < R; r += blockDim.x) m = fmaxf(m, det[(long long)r * ld]);
  m = block_max(m, sh);
  float s = 0.f, tot = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    s += expf(det[(long long)r * ld] - m);
    tot += scores[(long long)r * K + c];
  }
  s = block_sum(s, sh);
  tot = block_sum(tot, sh);
  float g = 0.f;
  if (tot >= 1e-6f && tot <= 1.0f - 1e-6f) {
    const float y = gt_onehot[c];
    // d/dp [-(y log p + (1-y) log(1-p))] = -y/p + (1-y)/(1-p)
    g = (up ? up[0] : 1.f) * scale * (-y / tot + (1.f - y) / (1.f - tot));
  }
  if (threadIdx.x == 0) g_out[c] = g;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float b = expf(det[(long long)r * ld] - m) / s;
    dlogits[(long long)r * ld + det_off + c] = g * (scores[(long long)r * K + c] - b * tot);
  }
}

// pass 2: one thread per proposal row: row softmax a, T_r, d cls.
__global__ void mil_bwd_rows_kernel(const float* __restrict__ logits, int ld, int R, int K, int cls_off,
                                    const float* __restrict__ scores, const float* __restrict__ g,
                                    float* __restrict__ dlogits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* x = logits + (long long)r * ld + cls_off;
  const float* sc = scores + (long long)r * K;
  float m = -INFINITY;
  for (int k = 0; k < K; ++k) m = fmaxf(m, x[k]);
  float s = 0.f, T = 0.f;
  for (int k = 0; k < K; ++k) {
    s += expf(x[k] - m);
    T += g[k] * sc[k];
  }
  float* d = dlogits + (long long)r * ld + cls_off;
  for (int k = 0; k < K; ++k) d[k] = g[k] * sc[k] - (expf(x[k] - m) / s) * T;
}

// ---- OICR stage: L = up * scale * sum_r w_r CE_r / nvalid  ->  d logit_rj = up * scale * w_r (p_rj - [j == label_r]) / nvalid
__global__ void oicr_stage_bwd_kernel(const float* __restrict__ probs, const int64_t* __restrict__ labels,
                                      const float* __restrict__ weights, const float* __restrict__ nvalid, float scale,
                                      const float* __restrict__ up, int R, int K, int ld, int col_off,
                                      float* __restrict__ dlogits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int C1 = K + 1;
  const int lab = (int)labels[r];
  const float coef = (lab >= 0) ? (up ? up[0] : 1.f) * scale * weights[r] / nvalid[0] : 0.f;  // ignore_index = -1
  const float* p = probs + (long long)r * C1;
  float* d = dlogits + (long long)r * ld + col_off;
  for (int j = 0; j < C1; ++j) d[j] = coef * (p[j] - (j == lab ? 1.f : 0.f));
}

// ---- box regression: L = up * scale * sum_{fg rows} smooth_l1(d - t) / denom  ->  d d_j = up*scale/denom * smooth_l1'(d_j - t_j)
__global__ void oicr_boxreg_bwd_kernel(const float* __restrict__ deltas, int ld, int col_off, int R, int K, int agnostic,
                                       const float* __restrict__ boxes, const float* __restrict__ pgt_box,
                                       const int64_t* __restrict__ labels, const int64_t* __restrict__ matched, float wx,
                                       float wy, float ww, float wh, float beta, float scale_over_denom,
                                       const float* __restrict__ up, float* __restrict__ dlogits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int ncol = agnostic ? 4 : 4 * K;
  float* dd = dlogits + (long long)r * ld + col_off;
  for (int j = 0; j < ncol; ++j) dd[j] = 0.f;
  const int lab = (int)labels[r];
  if (lab < 0 || lab >= K) return;
  const float* pb = boxes + 4 * (long long)r;
  const float* gb = pgt_box + 4 * matched[r];
  const float sw = __fsub_rn(pb[2], pb[0]), shh = __fsub_rn(pb[3], pb[1]);
  const float scx = __fadd_rn(pb[0], __fmul_rn(0.5f, sw)), scy = __fadd_rn(pb[1], __fmul_rn(0.5f, shh));
  const float tw = __fsub_rn(gb[2], gb[0]), th = __fsub_rn(gb[3], gb[1]);
  const float tcx = __fadd_rn(gb[0], __fmul_rn(0.5f, tw)), tcy = __fadd_rn(gb[1], __fmul_rn(0.5f, th));
  float t[4];
  t[0] = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tcx, scx)), sw);
  t[1] = __fdiv_rn(__fmul_rn(wy, __fsub_rn(tcy, scy)), shh);
  t[2] = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
  t[3] = __fmul_rn(wh, logf(__fdiv_rn(th, shh)));
  const float coef = (up ? up[0] : 1.f) * scale_over_denom;
  const float* d = deltas + (long long)r * ld + col_off + (agnostic ? 0 : 4 * lab);
  float* o = dd + (agnostic ? 0 : 4 * lab);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float e = d[j] - t[j];
    const float sgn = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
    o[j] = coef * ((beta < 1e-5f || fabsf(e) >= beta) ? sgn : e / beta);
  }
}

// ---- masked transpose ------------------------------------------------------------------------------------
// out[c'][r] = g[r][c] * (mask == NULL || mask[r][c] != 0 ? mul : 0),  c' = perm49 ? (c % C49) * 49 + c / C49 : c
// (perm49 = C49 > 0: input columns are bin-major (bin * C49 + ch), output rows channel-major (ch * 49 + bin): the
// reference's flatten order of the pooled features).  Rows r >= R of the output (up to ldo) are zero-filled so the
// transposed matrix can be the K-major operand of a GEMM with K = ldo.  Optionally also writes the masked,
// untransposed gradient (same layout as g) for the next dgrad GEMM.
// grid = (ceil(ldo / 32), ceil(C / 32)), block = (32, 8)
template <typename TI, typename TM, typename TO>
__global__ void masked_transpose_kernel(const TI* __restrict__ g, int ldg, const TM* __restrict__ mask, int ldm, float mul,
                                        int R, int C, int C49, TO* __restrict__ out, int ldo, TO* __restrict__ masked,
                                        int ldmk) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = (float)g[(long long)r * ldg + c];
      if (mask) v = ((float)mask[(long long)r * ldm + c] != 0.f) ? v * mul : 0.f;
      else v *= mul;
      if (masked) masked[(long long)r * ldmk + c] = (TO)v;
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < ldo) {
      const int co = C49 > 0 ? (c % C49) * 49 + c / C49 : c;
      out[(long long)co * ldo + r] = (TO)tile[threadIdx.x][i];
    }
  }
}

// Plain bf16 transpose (no mask, no scaling) for the big operands (pooled features: 0.8 GB): every thread moves an
// 8 x 8 block through registers -- eight 16-byte loads along the columns of eight consecutive rows, byte-permute
// transpose, eight 16-byte stores along the rows of the output.  A warp reads 4 x 128 contiguous bytes per load
// instruction and writes 8 x 64; no shared memory.  grid = (ceil(ldo / 256), ceil(C / 64)), block = (8, 32).
__global__ void __launch_bounds__(256)
transpose_bf16_8x8_kernel(const uint4* __restrict__ g, int ldg8, int R, int C, int C49, uint4* __restrict__ out, int ldo8) {
  const int c = (blockIdx.y * 8 + threadIdx.x) * 8;   // first of this thread's 8 input columns
  const int r = (blockIdx.x * 32 + threadIdx.y) * 8;  // first of this thread's 8 input rows
  if (c >= C || r >= ldo8 * 8) return;
  uint32_t in[8][4];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (r + k < R) v = __ldg(g + (size_t)(r + k) * ldg8 + (c >> 3));
    in[k][0] = v.x; in[k][1] = v.y; in[k][2] = v.z; in[k][3] = v.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {  // output row = input column c + j; its 8 elements = column j of rows r .. r + 7
    const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
    uint4 o;
    o.x = __byte_perm(in[0][j >> 1], in[1][j >> 1], sel);
    o.y = __byte_perm(in[2][j >> 1], in[3][j >> 1], sel);
    o.z = __byte_perm(in[4][j >> 1], in[5][j >> 1], sel);
    o.w = __byte_perm(in[6][j >> 1], in[7][j >> 1], sel);
    const int cc = c + j;
    const int co = C49 > 0 ? (cc % C49) * 49 + cc / C49 : cc;
    out[(size_t)co * ldo8 + (r >> 3)] = o;
  }
}

// rowsum[c] = sum_r x[c][r]  (bias gradients from a transposed gradient matrix); one warp per row, fixed order
template <typename T>
__global__ void rowsum_kernel(const T* __restrict__ x, int ld, int rows, int cols, float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  float s = 0.f;
  for (int i = lane; i < cols; i += 32) s += (float)x[(long long)row * ld + i];
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}

// [rows][C49 * 49] bin-major columns -> channel-major columns (fp32 weight gradients of fc6 in the exact-fp32 mode)
__global__ void permute_cols49_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int C49) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncol = (long long)C49 * 49;
  if (i >= rows * ncol) return;
  const long long row = i / ncol;
  const int c = (int)(i - row * ncol);  // bin-major index: bin * C49 + ch
  out[row * ncol + (long long)(c % C49) * 49 + c / C49] = in[i];
}

// ---- SGD step fused with the refresh of the tensor-core weight copy ------------------------------------------
// torch.optim.SGD arithmetic (detectron2/solver/build.py builds exactly that optimizer): d = g + wd * w;
// buf = first ? d : momentum * buf + d;  d = nesterov ? d + momentum * buf : buf;  w -= lr * d.
// PERM: w is [N][C49 * 49] in the reference's (c, ph, pw) column order and `packed` the bf16 copy in the kernels'
// bin-major order; one CTA owns 64 channels x 49 bins of one row (contiguous in w), updates them in place and
// transposes the new values through shared memory so that both sides are coalesced.  grid = (N, C49 / 64).
__device__ __forceinline__ float sgd_update(float w, float g, float* mom, long long i, float lr, float momentum, float wd,
                                            int nesterov, int first) {
  float d = fmaf(wd, w, g);
  if (momentum != 0.f) {
    const float b = first ? d : fmaf(momentum, mom[i], d);
    mom[i] = b;
    d = nesterov ? fmaf(momentum, b, d) : b;
  }
  return fmaf(-lr, d, w);
}

__global__ void __launch_bounds__(256)
sgd_pack_perm_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ mom,
                     __nv_bfloat16* __restrict__ packed, int C49, float lr, float momentum, float wd, int nesterov, int first) {
  __shared__ float tile[64 * 49];
  const long long K = (long long)C49 * 49;
  const long long base = (long long)blockIdx.x * K + (long long)blockIdx.y * 64 * 49;
  for (int i = threadIdx.x; i < 64 * 49; i += 256) {
    const float nw = sgd_update(w[base + i], g[base + i], mom, base + i, lr, momentum, wd, nesterov, first);
    w[base + i] = nw;
    tile[i] = nw;  // i = cc * 49 + bin
  }
  __syncthreads();
  __nv_bfloat16* prow = packed + (long long)blockIdx.x * K + blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 49; i += 256) {
    const int bin = i >> 6, cc = i & 63;
    prow[(long long)bin * C49 + cc] = __float2bfloat16(tile[cc * 49 + bin]);
  }
}

__global__ void sgd_pack_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ mom,
                                __nv_bfloat16* __restrict__ packed, long long n, float lr, float momentum, float wd,
                                int nesterov, int first) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float nw = sgd_update(w[i], g[i], mom, i, lr, momentum, wd, nesterov, first);
  w[i] = nw;
  if (packed) packed[i] = __float2bfloat16(nw);
}

// bf16 kernel layout of a linear layer's weight without an optimizer step (weight load / first use)
__global__ void __launch_bounds__(256)
pack_perm_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed, int C49) {
  __shared__ float tile[64 * 49];
  const long long K = (long long)C49 * 49;
  const long long base = (long long)blockIdx.x * K + (long long)blockIdx.y * 64 * 49;
  for (int i = threadIdx.x; i < 64 * 49; i += 256) tile[i] = w[base + i];
  __syncthreads();
  __nv_bfloat16* prow = packed + (long long)blockIdx.x * K + blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 49; i += 256) {
    const int bin = i >> 6, cc = i & 63;
    prow[(long long)bin * C49 + cc] = __float2bfloat16(tile[cc * 49 + bin]);
  }
}

}  // namespace drn

using namespace drn;

template <typename TI, typename TM, typename TO>
static int launch_mt(const void* g, int ldg, const void* mask, int ldm, float mul, int R, int C, int C49, void* out, int ldo,
                     void* masked, int ldmk, cudaStream_t st) {
  dim3 grid(cdiv(ldo, 32), cdiv(C, 32)), block(32, 8);
  masked_transpose_kernel<TI, TM, TO><<<grid, block, 0, st>>>((const TI*)g, ldg, (const TM*)mask, ldm, mul, R, C, C49, (TO*)out, ldo,
                                                             (TO*)masked, ldmk);
  DRN_CHECK_LAUNCH("masked_transpose");
  return 0;
}

extern "C" {

int drn_wsddn_mil_bwd(const float* logits, int ld, int R, int K, int cls_off, int det_off, const float* scores,
                      const float* gt_onehot, int mean_loss, float loss_scale, const float* grad_loss, float* dlogits,
                      float* g_ws, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && scores && gt_onehot && dlogits && g_ws, "wsddn_mil_bwd: null pointer");
  DRN_CHECK_ARG(R > 0 && K > 0, "wsddn_mil_bwd: R=%d K=%d", R, K);
  DRN_CHECK_ARG(cls_off + K <= ld && det_off + K <= ld, "wsddn_mil_bwd: column ranges exceed ld=%d", ld);
  cudaStream_t st = (cudaStream_t)stream;
  const float scale = mean_loss ? loss_scale / (float)K : loss_scale;
  mil_bwd_cols_kernel<<<K, 512, 0, st>>>(logits, ld, R, K, det_off, scores, gt_onehot, scale, grad_loss, dlogits, g_ws);
  DRN_CHECK_LAUNCH("wsddn_mil_bwd cols");
  mil_bwd_rows_kernel<<<cdiv(R, 128), 128, 0, st>>>(logits, ld, R, K, cls_off, scores, g_ws, dlogits);
  DRN_CHECK_LAUNCH("wsddn_mil_bwd rows");
  return 0;
}

int drn_oicr_stage_bwd(const float* probs, const int64_t* labels, const float* weights, const float* nvalid, float loss_scale,
                       const float* grad_loss, int R, int K, int ld, int col_off, float* dlogits, drn_stream_t stream) {
  DRN_CHECK_ARG(probs && labels && weights && nvalid && dlogits, "oicr_stage_bwd: null pointer");
  DRN_CHECK_ARG(R > 0 && col_off + K + 1 <= ld, "oicr_stage_bwd: R=%d, columns exceed ld=%d", R, ld);
  oicr_stage_bwd_kernel<<<cdiv(R, 128), 128, 0, (cudaStream_t)stream>>>(probs, labels, weights, nvalid, loss_scale, grad_loss, R, K,
                                                                        ld, col_off, dlogits);
  DRN_CHECK_LAUNCH("oicr_stage_bwd");
  return 0;
}

int drn_oicr_boxreg_bwd(const float* deltas, int ld, int col_off, int R, int K, int cls_agnostic, const float* boxes,
                        const float* pgt_box, const int64_t* labels, const int64_t* matched_idx, const float* bbox_w_host,
                        float beta, float loss_scale, float denom, const float* grad_loss, float* dlogits,
                        drn_stream_t stream) {
  DRN_CHECK_ARG(deltas && boxes && pgt_box && labels && matched_idx && bbox_w_host && dlogits, "oicr_boxreg_bwd: null pointer");
  DRN_CHECK_ARG(R > 0 && denom > 0.f, "oicr_boxreg_bwd: R=%d denom=%f", R, denom);
  oicr_boxreg_bwd_kernel<<<cdiv(R, 128), 128, 0, (cudaStream_t)stream>>>(deltas, ld, col_off, R, K, cls_agnostic, boxes, pgt_box,
      labels, matched_idx, bbox_w_host[0], bbox_w_host[1], bbox_w_host[2], bbox_w_host[3], beta, loss_scale / denom, grad_loss,
      dlogits);
  DRN_CHECK_LAUNCH("oicr_boxreg_bwd");
  return 0;
}

int drn_masked_transpose(const void* grad, int ld_grad, int grad_dtype, const void* mask, int ld_mask, int mask_dtype, float mul,
                         int R, int C, int c49, void* out_t, int ld_out, void* out_masked, int ld_masked, int out_dtype,
                         drn_stream_t stream) {
  DRN_CHECK_ARG(grad && out_t, "masked_transpose: null pointer");
  DRN_CHECK_ARG(R >= 0 && C > 0 && ld_out >= R, "masked_transpose: R=%d C=%d ld_out=%d", R, C, ld_out);
  DRN_CHECK_ARG(c49 == 0 || C == c49 * 49, "masked_transpose: C=%d is not 49 x %d", C, c49);
  if (ld_out == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool gi = grad_dtype == DRN_BF16, mi = mask_dtype == DRN_BF16, oi = out_dtype == DRN_BF16;
  if (gi && oi && !mask && !out_masked && mul == 1.f && C % 8 == 0 && ld_grad % 8 == 0 && ld_out % 8 == 0 &&
      (uintptr_t)grad % 16 == 0 && (uintptr_t)out_t % 16 == 0) {
    transpose_bf16_8x8_kernel<<<dim3(cdiv(ld_out, 256), cdiv(C, 64)), dim3(8, 32), 0, st>>>((const uint4*)grad, ld_grad / 8, R, C, c49,
                                                                                          (uint4*)out_t, ld_out / 8);
    DRN_CHECK_LAUNCH("transpose_bf16");
    return 0;
  }
#define DRN_MT(TI, TM, TO) return launch_mt<TI, TM, TO>(grad, ld_grad, mask, ld_mask, mul, R, C, c49, out_t, ld_out, out_masked, ld_masked, st)
  if (gi && mi && oi) DRN_MT(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
  if (!gi && mi && oi) DRN_MT(float, __nv_bfloat16, __nv_bfloat16);
  if (!gi && !mi && !oi) DRN_MT(float, float, float);
  if (gi && mi && !oi) DRN_MT(__nv_bfloat16, __nv_bfloat16, float);
#undef DRN_MT
  return set_err("masked_transpose: unsupported dtype combination (%d, %d, %d)", grad_dtype, mask_dtype, out_dtype);
}

int drn_rowsum(const void* x, int ld, int rows, int cols, int dtype, float* out, drn_stream_t stream) {
  DRN_CHECK_ARG(x && out, "rowsum: null pointer");
  if (rows == 0) return 0;
  if (dtype == DRN_BF16)
    rowsum_kernel<__nv_bfloat16><<<cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, rows, cols, out);
  else
    rowsum_kernel<float><<<cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, rows, cols, out);
  DRN_CHECK_LAUNCH("rowsum");
  return 0;
}

int drn_permute_cols49(const float* in, float* out, int64_t rows, int c49, drn_stream_t stream) {
  DRN_CHECK_ARG(in && out && c49 > 0, "permute_cols49: bad arguments");
  const long long n = (long long)rows * c49 * 49;
  if (n == 0) return 0;
  permute_cols49_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, rows, c49);
  DRN_CHECK_LAUNCH("permute_cols49");
  return 0;
}

int drn_sgd_step(float* w, const float* grad, float* momentum_buf, void* packed_bf16, int64_t rows, int64_t cols, int c49,
                 float lr, float momentum, float weight_decay, int nesterov, int first_step, drn_stream_t stream) {
  DRN_CHECK_ARG(w && grad, "sgd_step: null pointer");
  DRN_CHECK_ARG(momentum == 0.f || momentum_buf, "sgd_step: momentum without a buffer");
  const long long n = (long long)rows * cols;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (c49 > 0 && packed_bf16) {
    DRN_CHECK_ARG(cols == (int64_t)c49 * 49 && c49 % 64 == 0, "sgd_step: cols=%lld is not 49 x %d (c49 %% 64 == 0)", (long long)cols, c49);
    DRN_CHECK_ARG(rows <= 0x7fffffff, "sgd_step: too many rows");
    sgd_pack_perm_kernel<<<dim3((unsigned)rows, c49 / 64), 256, 0, st>>>(w, grad, momentum_buf, (__nv_bfloat16*)packed_bf16, c49, lr,
                                                                          momentum, weight_decay, nesterov, first_step);
  } else {
    sgd_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, grad, momentum_buf, (__nv_bfloat16*)packed_bf16, n, lr, momentum,
                                                                 weight_decay, nesterov, first_step);
  }
  DRN_CHECK_LAUNCH("sgd_step");
  return 0;
}

int drn_pack_linear_bf16(const float* w, void* packed_bf16, int64_t rows, int64_t cols, int c49, drn_stream_t stream) {
  DRN_CHECK_ARG(w && packed_bf16, "pack_linear: null pointer");
  if (rows * cols == 0) return 0;
  if (c49 > 0) {
    DRN_CHECK_ARG(cols == (int64_t)c49 * 49 && c49 % 64 == 0, "pack_linear: cols=%lld is not 49 x %d (c49 %% 64 == 0)", (long long)cols, c49);
    pack_perm_kernel<<<dim3((unsigned)rows, c49 / 64), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)packed_bf16, c49);
    DRN_CHECK_LAUNCH("pack_linear");
    return 0;
  }
  return drn_cast_f32_to_bf16(w, packed_bf16, rows * cols, stream);
}

}  // extern "C"
