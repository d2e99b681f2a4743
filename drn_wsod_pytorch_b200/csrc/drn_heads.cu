// WSDDN dual-softmax MIL head, OICR pseudo-GT mining / labelling / weighted CE, inference scores.
// All fp32, latency-bound (R x K is a few hundred KB); written so that no step needs a host sync:
// every reduction the reference does with .item()/numpy happens on the device.
// Box arithmetic uses __fmul_rn/__fadd_rn (no FMA contraction) so it rounds exactly like the
// reference's separate torch ops -- labels hinge on IoU >= 0.5 comparisons.
#include "common.cuh"
#include <float.h>
#include <math.h>

namespace drn {

// ---------------------------------------------------------------- WSDDN MIL
// pass 1: per-row softmax statistics of the cls logits (softmax over classes, dim=1).
__global__ void __launch_bounds__(256)
mil_rowstats_kernel(const float* __restrict__ logits, int ld, int R, int K, int cls_off,
                    float* __restrict__ rowmax, float* __restrict__ rowsum) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* x = logits + (long long)r * ld + cls_off;
  float m = -INFINITY;
#pragma unroll 8
  for (int k = 0; k < K; ++k) m = fmaxf(m, __ldg(x + k));  // unrolled: 8 independent loads in flight per thread
  float s = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) s += expf(__ldg(x + k) - m);
  rowmax[r] = m;
  rowsum[r] = s;
}

// pass 2: one CTA per class: softmax over proposals (dim=0) of the det logits, product with the
// row softmax, image-level score = clamp(sum_r), BCE term of this class -> bce_terms[c].
__global__ void __launch_bounds__(512)
mil_cols_kernel(const float* __restrict__ logits, int ld, int R, int K, int cls_off, int det_off,
                const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                const float* __restrict__ gt_onehot, float* __restrict__ scores,
                float* __restrict__ img_score, float* __restrict__ bce_terms) {
  __shared__ float sh[32];
  const int c = blockIdx.x;
  const float* det = logits + det_off + c;
  const float* cls = logits + cls_off + c;
  float m = -INFINITY;
  for (int r = threadIdx.x; r < R; r += blockDim.x) m = fmaxf(m, det[(long long)r * ld]);
  m = block_max(m, sh);
  float s = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) s += expf(det[(long long)r * ld] - m);
  s = block_sum(s, sh);
  float tot = 0.f;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float pc = expf(cls[(long long)r * ld] - rowmax[r]) / rowsum[r];
    const float pd = expf(det[(long long)r * ld] - m) / s;
    const float sc = pc * pd;
    scores[(long long)r * K + c] = sc;
    tot += sc;
  }
  tot = block_sum(tot, sh);
  if (threadIdx.x == 0) {
    const float p = fminf(fmaxf(tot, 1e-6f), 1.0f - 1e-6f);
    img_score[c] = p;
    const float y = gt_onehot[c];
    // torch.nn.functional.binary_cross_entropy clamps the logs at -100
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    bce_terms[c] = -(y * lp + (1.f - y) * l1p);
  }
}

__global__ void mil_finalize_kernel(const float* __restrict__ bce_terms, int K, int mean_loss,
                                    float loss_scale, float* __restrict__ loss) {
  // fixed-order sum (deterministic)
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += bce_terms[k];
    if (mean_loss) s = s / (float)K;
    loss[0] = s * loss_scale;
  }
}

// ---------------------------------------------------------------- box helpers
struct Box4 { float x1, y1, x2, y2; };

__device__ __forceinline__ Box4 apply_deltas_rn(Box4 b, float d0, float d1, float d2, float d3,
                                                float wx, float wy, float ww, float wh) {
  // detectron2/modeling/box_regression.py:73-110, one rounding per torch op
  const float scale_clamp = 4.135166556742356f;  // log(1000/16)
  const float w = __fsub_rn(b.x2, b.x1), h = __fsub_rn(b.y2, b.y1);
  const float cx = __fadd_rn(b.x1, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y1, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d0, wx), dy = __fdiv_rn(d1, wy);
  const float dw = fminf(__fdiv_rn(d2, ww), scale_clamp), dh = fminf(__fdiv_rn(d3, wh), scale_clamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  Box4 o;
  o.x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o.y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  o.y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  return o;
}

__device__ __forceinline__ float iou_rn(Box4 g, float area_g, Box4 p, float area_p) {
  // detectron2/structures/boxes.py:329-361 (boxes1 = targets, boxes2 = proposals)
  const float iw = fmaxf(__fsub_rn(fminf(g.x2, p.x2), fmaxf(g.x1, p.x1)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(g.y2, p.y2), fmaxf(g.y1, p.y1)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  return inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_g, area_p), inter)) : 0.f;
}

// ---------------------------------------------------------------- pseudo-GT mining (get_pgt)
__device__ __forceinline__ bool better(float av, int ai, float bv, int bi) {
  // true if (av, ai) beats (bv, bi) under torch.max(dim=0) CPU semantics:
  // NaN beats numbers, larger beats smaller, ties -> lowest index.
  const bool an = av != av, bn = bv != bv;
  if (an || bn) {
    if (an && bn) return ai < bi;
    return an;
  }
  if (av > bv) return true;
  if (av < bv) return false;
  return ai < bi;
}

__global__ void __launch_bounds__(256)
oicr_pgt_kernel(const float* __restrict__ prev, int ld, int R, const float* __restrict__ boxes,
                const int64_t* __restrict__ gt_classes, const float* __restrict__ img_score,
                int rederive, const float* __restrict__ deltas, int ld_deltas, int agnostic, float wx,
                float wy, float ww, float wh, int64_t* __restrict__ pgt_idx, float* __restrict__ pgt_score,
                float* __restrict__ pgt_box, float* __restrict__ pgt_weight) {
  __shared__ float sv[8];
  __shared__ int si[8];
  const int g = blockIdx.x;
  const int c = (int)gt_classes[g];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int r = threadIdx.x; r < R; r += blockDim.x) {
    const float v = prev[(long long)r * ld + c];
    if (better(v, r, bv, bi)) { bv = v; bi = r; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sv[wid] = bv; si[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
      if (better(sv[i], si[i], bv, bi)) { bv = sv[i]; bi = si[i]; }
    if (bi == 0x7fffffff) bi = 0;  // R == 0 guard
    pgt_idx[g] = bi;
    pgt_score[g] = bv;
    pgt_weight[g] = img_score[c];
    Box4 b = {boxes[4 * bi + 0], boxes[4 * bi + 1], boxes[4 * bi + 2], boxes[4 * bi + 3]};
    if (rederive) {
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
      if (deltas) {
        const float* d = deltas + (long long)bi * ld_deltas + (agnostic ? 0 : 4 * c);
        d0 = d[0]; d1 = d[1]; d2 = d[2]; d3 = d[3];
      }
      b = apply_deltas_rn(b, d0, d1, d2, d3, wx, wy, ww, wh);
    }
    pgt_box[4 * g + 0] = b.x1; pgt_box[4 * g + 1] = b.y1;
    pgt_box[4 * g + 2] = b.x2; pgt_box[4 * g + 3] = b.y2;
  }
}

// ---------------------------------------------------------------- labelling (IoU + Matcher)
constexpr int MAX_G = 256;
struct MatcherCfg { int nthr; float thr[4]; int lab[5]; };

__global__ void __launch_bounds__(256)
label_proposals_kernel(const float* __restrict__ boxes, int R, const float* __restrict__ gt_boxes,
                       const int64_t* __restrict__ gt_classes, int G, int K, MatcherCfg mc,
                       int64_t* __restrict__ labels, int64_t* __restrict__ matched,
                       int32_t* __restrict__ counts) {
  __shared__ float sg[MAX_G][5];
  __shared__ int sc[MAX_G];
  __shared__ int scount[3];
  if (threadIdx.x < 3) scount[threadIdx.x] = 0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float x1 = gt_boxes[4 * g], y1 = gt_boxes[4 * g + 1], x2 = gt_boxes[4 * g + 2], y2 = gt_boxes[4 * g + 3];
    sg[g][0] = x1; sg[g][1] = y1; sg[g][2] = x2; sg[g][3] = y2;
    sg[g][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    sc[g] = (int)gt_classes[g];
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(boxes) + r);
    const Box4 p = {bb.x, bb.y, bb.z, bb.w};
    const float ap = __fmul_rn(__fsub_rn(p.x2, p.x1), __fsub_rn(p.y2, p.y1));
    int lab_out, mi = 0;
    if (G == 0) {
      lab_out = K;  // roi_heads.py:236 (no gt -> background)
    } else {
      float best = -INFINITY;
      bool best_nan = false;
      for (int g = 0; g < G; ++g) {
        const Box4 gb = {sg[g][0], sg[g][1], sg[g][2], sg[g][3]};
        const float v = iou_rn(gb, sg[g][4], p, ap);
        // torch.max(dim=0): first maximum wins, NaN propagates
        if (!best_nan && (v > best || v != v)) { best = v; mi = g; best_nan = (v != v); }
      }
      int ml = 1;  // matcher.py:88-96
      for (int i = 0; i <= mc.nthr; ++i) {
        const float lo = (i == 0) ? -INFINITY : mc.thr[i - 1];
        const float hi = (i == mc.nthr) ? INFINITY : mc.thr[i];
        if (best >= lo && best < hi) ml = mc.lab[i];
      }
      lab_out = sc[mi];
      if (ml == 0) lab_out = K;
      if (ml == -1) lab_out = -1;
    }
    labels[r] = lab_out;
    matched[r] = mi;
    atomicAdd(&scount[lab_out == -1 ? 2 : (lab_out == K ? 1 : 0)], 1);
  }
  __syncthreads();
  if (threadIdx.x < 3 && scount[threadIdx.x]) atomicAdd(counts + threadIdx.x, scount[threadIdx.x]);
}

// ---------------------------------------------------------------- OICR stage loss + softmax
constexpr int STAGE_THREADS = 256;

__global__ void __launch_bounds__(STAGE_THREADS)
oicr_stage_kernel(const float* __restrict__ logits, int ld, int col_off, int R, int K,
                  const int64_t* __restrict__ labels, const int64_t* __restrict__ matched,
                  const float* __restrict__ pgt_weight, float loss_scale, float* __restrict__ probs,
                  float* __restrict__ loss, float* __restrict__ stats, float* __restrict__ weights,
                  float* __restrict__ part, uint32_t* __restrict__ counter) {
  __shared__ float sh[32];
  __shared__ bool is_last;
  const int C1 = K + 1;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float lw = 0.f, valid = 0.f, acc = 0.f, nfg = 0.f, fgacc = 0.f, fneg = 0.f;
  if (r < R) {
    const float* x = logits + (long long)r * ld + col_off;
    float m = -INFINITY;
    int am = 0;
#pragma unroll 8
    for (int k = 0; k < C1; ++k) {
      const float v = __ldg(x + k);
      if (v > m) { m = v; am = k; }  // argmax: first maximum
    }
    float s = 0.f;
#pragma unroll 8
    for (int k = 0; k < C1; ++k) s += expf(__ldg(x + k) - m);
    float* p = probs + (long long)r * C1;
#pragma unroll 8
    for (int k = 0; k < C1; ++k) p[k] = expf(__ldg(x + k) - m) / s;
    const int lab = (int)labels[r];
    float w = pgt_weight[matched[r]];
    if (lab == -1) w = 0.f;                 // fast_rcnn.py:1090
    if (weights) weights[r] = w;
    if (w > 1e-12f) valid = 1.f;            // fast_rcnn.py:1092-1093
    if (lab >= 0) {
      const float ce = -((x[lab] - m) - logf(s));  // log_softmax + nll, ignore_index=-1 -> 0
      lw = ce * w;
    }
    // _log_accuracy counters (fast_rcnn.py:1098-1126)
    const bool fg = lab >= 0 && lab < K;
    if (am == lab) acc = 1.f;
    if (fg) {
      nfg = 1.f;
      if (am == lab) fgacc = 1.f;
      if (am == K) fneg = 1.f;
    }
  }
  const int nb = gridDim.x;
  float v;
  v = block_sum(lw, sh);    if (threadIdx.x == 0) part[0 * nb + blockIdx.x] = v;
  v = block_sum(valid, sh); if (threadIdx.x == 0) part[1 * nb + blockIdx.x] = v;
  v = block_sum(acc, sh);   if (threadIdx.x == 0) part[2 * nb + blockIdx.x] = v;
  v = block_sum(nfg, sh);   if (threadIdx.x == 0) part[3 * nb + blockIdx.x] = v;
  v = block_sum(fgacc, sh); if (threadIdx.x == 0) part[4 * nb + blockIdx.x] = v;
  v = block_sum(fneg, sh);  if (threadIdx.x == 0) part[5 * nb + blockIdx.x] = v;
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == (unsigned)(nb - 1));
  __syncthreads();
  if (is_last && threadIdx.x < 6) {
    float s = 0.f;
    const volatile float* pp = part + threadIdx.x * nb;
    for (int b = 0; b < nb; ++b) s += pp[b];  // fixed order
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    loss[0] = (sh[0] / sh[1]) * loss_scale;  // fast_rcnn.py:1144 (0/0 -> NaN like the reference)
    stats[0] = sh[2]; stats[1] = sh[3]; stats[2] = sh[4]; stats[3] = sh[5];
    stats[4] = sh[0]; stats[5] = sh[1];  // numerator / #valid, for multi-image batches
    *counter = 0u;
  }
}

// ---------------------------------------------------------------- fused tail (4 kernels instead of 13)
// The separate kernels above are one launch per reference function; at R = 4000 each is a few microseconds
// of work behind ~5-10 us of launch + cold-start latency (measured: 122 us per image for the whole tail).
// The fused variants keep the same device functions and reduction orders (results are bit-identical) and
// hand the serial dependencies -- image score -> pseudo GT -> labels -> loss -> next pseudo GT -- from one
// kernel to the next through "last block" epilogues instead of extra launches.

// WSDDN MIL + BCE + stage-0 pseudo-GT mining (after mil_rowstats_kernel).  grid = K CTAs (one per class), 512 threads.
__global__ void __launch_bounds__(512)
mil_pgt_fused_kernel(const float* __restrict__ logits, int ld, int R, int K, int cls_off, int det_off,
                     const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                     const float* __restrict__ gt_onehot, int mean_loss, float loss_scale,
                     const float* __restrict__ boxes, const int64_t* __restrict__ gt_classes, int G,
                     float* __restrict__ scores, float* __restrict__ img_score, float* __restrict__ loss,
                     int64_t* __restrict__ pgt_idx, float* __restrict__ pgt_score, float* __restrict__ pgt_box,
                     float* __restrict__ pgt_weight, float* __restrict__ bce_terms, uint32_t* __restrict__ counter) {
  __shared__ float sh[32];
  __shared__ float sv[16];
  __shared__ int si[16];
  __shared__ bool is_last;
  const int c = blockIdx.x;
  const float* det = logits + det_off + c;
  const float* cls = logits + cls_off + c;
  float m = -INFINITY, s = 0.f, tot = 0.f, bv = -INFINITY;
  int bi = 0x7fffffff;
  constexpr int MR = 8;  // rows per thread held in registers: R <= 8 * 512
  if (R <= MR * (int)blockDim.x) {
    // every operand of this thread's rows is loaded once, all loads in flight together (the generic form below walks the
    // column three times: 3 x R / 512 dependent L2 round trips); per-thread visiting order and the block trees are unchanged
    float dv[MR], cv[MR], rmx[MR], rsm[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) {
      const int r = threadIdx.x + i * (int)blockDim.x;
      const bool ok = r < R;
      const long long off = (long long)(ok ? r : 0) * ld;
      dv[i] = ok ? __ldg(det + off) : -INFINITY;
      cv[i] = __ldg(cls + off);
      rmx[i] = __ldg(rowmax + (ok ? r : 0));
      rsm[i] = __ldg(rowsum + (ok ? r : 0));
    }
#pragma unroll
    for (int i = 0; i < MR; ++i) m = fmaxf(m, dv[i]);
    m = block_max(m, sh);
#pragma unroll
    for (int i = 0; i < MR; ++i)
      if (threadIdx.x + i * (int)blockDim.x < R) { dv[i] = expf(dv[i] - m); s += dv[i]; }
    s = block_sum(s, sh);
#pragma unroll
    for (int i = 0; i < MR; ++i) {
      const int r = threadIdx.x + i * (int)blockDim.x;
      if (r < R) {
        const float pc = expf(cv[i] - rmx[i]) / rsm[i];
        const float pd = dv[i] / s;
        const float sc = pc * pd;
        scores[(long long)r * K + c] = sc;
        tot += sc;
        if (better(sc, r, bv, bi)) { bv = sc; bi = r; }
      }
    }
  } else {
    for (int r = threadIdx.x; r < R; r += blockDim.x) m = fmaxf(m, det[(long long)r * ld]);
    m = block_max(m, sh);
    for (int r = threadIdx.x; r < R; r += blockDim.x) s += expf(det[(long long)r * ld] - m);
    s = block_sum(s, sh);
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      const float pc = expf(cls[(long long)r * ld] - rowmax[r]) / rowsum[r];
      const float pd = expf(det[(long long)r * ld] - m) / s;
      const float sc = pc * pd;
      scores[(long long)r * K + c] = sc;
      tot += sc;
      if (better(sc, r, bv, bi)) { bv = sc; bi = r; }
    }
  }
  tot = block_sum(tot, sh);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sv[wid] = bv; si[wid] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
      if (better(sv[i], si[i], bv, bi)) { bv = sv[i]; bi = si[i]; }
    if (bi == 0x7fffffff) bi = 0;
    const float p = fminf(fmaxf(tot, 1e-6f), 1.0f - 1e-6f);
    img_score[c] = p;
    const float y = gt_onehot[c];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    bce_terms[c] = -(y * lp + (1.f - y) * l1p);
    for (int g = 0; g < G; ++g)
      if ((int)gt_classes[g] == c) {  // roi_heads_oicr.py:491-567, stage 0: the proposal itself is the pseudo-GT box
        pgt_idx[g] = bi;
        pgt_score[g] = bv;
        pgt_weight[g] = p;
        pgt_box[4 * g + 0] = boxes[4 * bi + 0]; pgt_box[4 * g + 1] = boxes[4 * bi + 1];
        pgt_box[4 * g + 2] = boxes[4 * bi + 2]; pgt_box[4 * g + 3] = boxes[4 * bi + 3];
      }
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == (unsigned)(gridDim.x - 1));
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    const volatile float* bt = bce_terms;
    float t = 0.f;
    for (int k = 0; k < K; ++k) t += bt[k];  // fixed order, as mil_finalize_kernel
    if (mean_loss) t = t / (float)K;
    loss[0] = t * loss_scale;
    *counter = 0u;
  }
}

// One refinement stage: [first labelling vs the real GT] + labelling vs this stage's pseudo GT + weighted CE +
// softmax + accuracy counters + block partial argmax of the new probabilities; the last block reduces everything
// and mines the NEXT stage's pseudo GT.  grid = ceil(R / 256), 256 threads.
struct StageFusedArgs {
  const float* logits; int ld, col_off, R, K;
  const float* boxes;
  const int64_t* gt_img; int G;            // image-level classes (sorted), G <= MAX_G
  const float* pgt_box; const float* pgt_weight;
  MatcherCfg mc;
  float loss_scale;
  const float* gt_boxes; const int64_t* gt_classes; int Gb;   // real GT (first labelling), Gb == -1: skip
  int64_t* labels0; int64_t* matched0; int32_t* counts0;
  int64_t* labels; int64_t* matched; int32_t* counts;
  float* probs; float* loss; float* stats; float* weights;
  int has_next;
  const float* img_score; const float* deltas; int ld_deltas, agnostic; float wx, wy, ww, wh;
  int64_t* next_idx; float* next_score; float* next_box; float* next_weight;
  float* part;       // [9][nb] sums + [G][nb] (value) ; indices in part_idx
  int* part_idx;     // [G][nb]
  uint32_t* counter;
};

__device__ __forceinline__ void match_row(const Box4& p, float ap, const float (*sg)[5], const int* sc, int G, int K,
                                          const MatcherCfg& mc, int& lab_out, int& mi) {
  mi = 0;
  if (G == 0) { lab_out = K; return; }
  float best = -INFINITY;
  bool best_nan = false;
  for (int g = 0; g < G; ++g) {
    const Box4 gb = {sg[g][0], sg[g][1], sg[g][2], sg[g][3]};
    const float v = iou_rn(gb, sg[g][4], p, ap);
    if (!best_nan && (v > best || v != v)) { best = v; mi = g; best_nan = (v != v); }
  }
  int ml = 1;
  for (int i = 0; i <= mc.nthr; ++i) {
    const float lo = (i == 0) ? -INFINITY : mc.thr[i - 1];
    const float hi = (i == mc.nthr) ? INFINITY : mc.thr[i];
    if (best >= lo && best < hi) ml = mc.lab[i];
  }
  lab_out = sc[mi];
  if (ml == 0) lab_out = K;
  if (ml == -1) lab_out = -1;
}

// CMAX = 32: rows of up to 32 logits are loaded ONCE into registers (all loads in flight together; the generic CMAX = 0 form
// walks the row three times).  Sums keep the generic form's order (k ascending per row; block_sum tree over the rows, blocks
// in order), the 0/1 counters are exact integer reductions: results are bit-identical to the per-function kernels.
template <int CMAX>
__global__ void __launch_bounds__(STAGE_THREADS)
oicr_stage_fused_kernel(const StageFusedArgs a) {
  __shared__ float sg[MAX_G][5];
  __shared__ int scl[MAX_G];
  __shared__ float sg0[MAX_G][5];
  __shared__ int scl0[MAX_G];
  __shared__ float sh[32];
  __shared__ float sv[8][STAGE_THREADS / 32];
  __shared__ int si[8][STAGE_THREADS / 32];
  __shared__ int scnt[11];
  __shared__ bool is_last;
  const int C1 = a.K + 1, K = a.K, G = a.G, nb = gridDim.x;
  if (threadIdx.x < 11) scnt[threadIdx.x] = 0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float x1 = a.pgt_box[4 * g], y1 = a.pgt_box[4 * g + 1], x2 = a.pgt_box[4 * g + 2], y2 = a.pgt_box[4 * g + 3];
    sg[g][0] = x1; sg[g][1] = y1; sg[g][2] = x2; sg[g][3] = y2;
    sg[g][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    scl[g] = (int)a.gt_img[g];
  }
  const int Gb = a.Gb;
  for (int g = threadIdx.x; g < Gb; g += blockDim.x) {
    const float x1 = a.gt_boxes[4 * g], y1 = a.gt_boxes[4 * g + 1], x2 = a.gt_boxes[4 * g + 2], y2 = a.gt_boxes[4 * g + 3];
    sg0[g][0] = x1; sg0[g][1] = y1; sg0[g][2] = x2; sg0[g][3] = y2;
    sg0[g][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    scl0[g] = (int)a.gt_classes[g];
  }
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const float* x = a.logits + (long long)min(r, a.R - 1) * a.ld + a.col_off;
  float xr[CMAX > 0 ? CMAX : 1];
  if constexpr (CMAX > 0) {   // issued before the barrier: the row's loads fly while the GT tables are staged
#pragma unroll
    for (int k = 0; k < CMAX; ++k) xr[k] = (k < C1) ? __ldg(x + k) : -INFINITY;
  }
  __syncthreads();
  float lw = 0.f;
  int f_valid = 0, f_acc = 0, f_nfg = 0, f_fgacc = 0, f_fneg = 0;
  int c_fg = 0, c_bg = 0, c_ig = 0, c0_fg = 0, c0_bg = 0, c0_ig = 0;
  float pbuf_m = 0.f, pbuf_s = 1.f;
  if (r < a.R) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.boxes) + r);
    const Box4 p = {bb.x, bb.y, bb.z, bb.w};
    const float ap = __fmul_rn(__fsub_rn(p.x2, p.x1), __fsub_rn(p.y2, p.y1));
    if (Gb >= 0) {  // roi_heads_oicr.py:266: labelling against the real GT (logging + proposals' gt fields)
      int l0, m0;
      match_row(p, ap, sg0, scl0, Gb, K, a.mc, l0, m0);
      a.labels0[r] = l0;
      a.matched0[r] = m0;
      if (l0 == -1) c0_ig = 1; else if (l0 == K) c0_bg = 1; else c0_fg = 1;
    }
    int lab, mi;
    match_row(p, ap, sg, scl, G, K, a.mc, lab, mi);
    a.labels[r] = lab;
    a.matched[r] = mi;
    if (lab == -1) c_ig = 1; else if (lab == K) c_bg = 1; else c_fg = 1;
    float m = -INFINITY;
    int am = 0;
    float s = 0.f;
    float* pr = a.probs + (long long)r * C1;
    if constexpr (CMAX > 0) {
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1 && xr[k] > m) { m = xr[k]; am = k; }
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1) { xr[k] = expf(xr[k] - m); s += xr[k]; }
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1) pr[k] = xr[k] / s;
    } else {
#pragma unroll 8
      for (int k = 0; k < C1; ++k) {
        const float v = __ldg(x + k);
        if (v > m) { m = v; am = k; }
      }
#pragma unroll 8
      for (int k = 0; k < C1; ++k) s += expf(__ldg(x + k) - m);
#pragma unroll 8
      for (int k = 0; k < C1; ++k) pr[k] = expf(__ldg(x + k) - m) / s;
    }
    pbuf_m = m; pbuf_s = s;
    float w = a.pgt_weight[mi];
    if (lab == -1) w = 0.f;
    if (a.weights) a.weights[r] = w;
    if (w > 1e-12f) f_valid = 1;
    if (lab >= 0) lw = (-((__ldg(x + lab) - m) - logf(s))) * w;
    const bool fg = lab >= 0 && lab < K;
    if (am == lab) f_acc = 1;
    if (fg) {
      f_nfg = 1;
      if (am == lab) f_fgacc = 1;
      if (am == K) f_fneg = 1;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  {
    // the ten 0/1 counters: warp-level integer reductions + shared-memory atomics (exact, order-free)
    const int fl[11] = {f_valid, f_acc, f_nfg, f_fgacc, f_fneg, c_fg, c_bg, c_ig, c0_fg, c0_bg, c0_ig};
    const int nfl = Gb >= 0 ? 11 : 8;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      if (j < nfl) {
        const int v = __reduce_add_sync(0xffffffffu, fl[j]);
        if (lane == 0 && v) atomicAdd(&scnt[j], v);
      }
    }
  }
  float v = block_sum(lw, sh);   // the only floating-point sum: same tree as oicr_stage_kernel (its barriers also publish scnt)
  if (threadIdx.x == 0) a.part[0 * nb + blockIdx.x] = v;
  if (threadIdx.x < (Gb >= 0 ? 11 : 8)) a.part[(1 + threadIdx.x) * nb + blockIdx.x] = (float)scnt[threadIdx.x];
  if (a.has_next) {
    // block-partial argmax of the new probabilities for every image-level class (input of the next get_pgt), 8 classes per barrier
    for (int g0 = 0; g0 < G; g0 += 8) {
      const int ng = min(8, G - g0);
      if (g0 > 0) __syncthreads();
      for (int gg = 0; gg < ng; ++gg) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        if (r < a.R) { bv = expf(__ldg(x + scl[g0 + gg]) - pbuf_m) / pbuf_s; bi = r; }  // == probs[r][class g], same expression
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[gg][wid] = bv; si[gg][wid] = bi; }
      }
      __syncthreads();
      if (threadIdx.x < ng) {
        const int gg = threadIdx.x;
        float bv = sv[gg][0];
        int bi = si[gg][0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
          if (better(sv[gg][i], si[gg][i], bv, bi)) { bv = sv[gg][i]; bi = si[gg][i]; }
        a.part[(12 + g0 + gg) * nb + blockIdx.x] = bv;
        a.part_idx[(g0 + gg) * nb + blockIdx.x] = bi;
      }
    }
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counter, 1u) == (unsigned)(nb - 1));
  __syncthreads();
  if (!is_last) return;
  const int nsum = Gb >= 0 ? 12 : 9;
  if (threadIdx.x < nsum) {
    float s = 0.f;
    const volatile float* pp = a.part + threadIdx.x * nb;
    for (int b = 0; b < nb; ++b) s += pp[b];  // fixed order
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.loss[0] = (sh[0] / sh[1]) * a.loss_scale;
    a.stats[0] = sh[2]; a.stats[1] = sh[3]; a.stats[2] = sh[4]; a.stats[3] = sh[5];
    a.stats[4] = sh[0]; a.stats[5] = sh[1];
    a.counts[0] = (int)sh[6]; a.counts[1] = (int)sh[7]; a.counts[2] = (int)sh[8];
    if (Gb >= 0) { a.counts0[0] = (int)sh[9]; a.counts0[1] = (int)sh[10]; a.counts0[2] = (int)sh[11]; }
    *a.counter = 0u;
  }
  if (a.has_next && threadIdx.x < G) {
    const int g = threadIdx.x;
    const volatile float* pv = a.part + (12 + g) * nb;
    const volatile int* pi = a.part_idx + g * nb;
    float bv = pv[0];
    int bi = pi[0];
    for (int b = 1; b < nb; ++b) {
      const float ov = pv[b];
      const int oi = pi[b];
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;
    const int c = scl[g];
    a.next_idx[g] = bi;
    a.next_score[g] = bv;
    a.next_weight[g] = a.img_score[c];
    Box4 b = {a.boxes[4 * bi + 0], a.boxes[4 * bi + 1], a.boxes[4 * bi + 2], a.boxes[4 * bi + 3]};
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (a.deltas) {
      const float* d = a.deltas + (long long)bi * a.ld_deltas + (a.agnostic ? 0 : 4 * c);
      d0 = d[0]; d1 = d[1]; d2 = d[2]; d3 = d[3];
    }
    b = apply_deltas_rn(b, d0, d1, d2, d3, a.wx, a.wy, a.ww, a.wh);  // fast_rcnn.py:1511-1532 re-derived boxes
    a.next_box[4 * g + 0] = b.x1; a.next_box[4 * g + 1] = b.y1;
    a.next_box[4 * g + 2] = b.x2; a.next_box[4 * g + 3] = b.y2;
  }
}

// ---------------------------------------------------------------- all refinement stages of one image in TWO launches
// The stage chain above runs S kernels one after the other because each one mines the NEXT stage's pseudo GT.  But the pseudo
// GT of stage k+1 is the per-class argmax of stage k's softmax (roi_heads_oicr.py:491-567 get_pgt on
// `prev_pred_scores = predict_probs(...)`, detached) -- it depends on the stage's LOGITS only, never on its labels or loss.
// So: kernel A computes, for all stages side by side (blockIdx.y = stage), the softmax and the next pseudo GT; kernel B, again
// for all stages side by side, labels the proposals against their stage's pseudo GT and reduces the weighted CE and the
// counters.  Per-row arithmetic, block reductions and cross-block orders are those of oicr_stage_fused_kernel: results are
// bit-identical; the tail's latency chain drops from S kernels to two (measured: profiles/r2_tail_stage_parallel.txt).
constexpr int MAX_STAGES = 8;
struct StagesArgs {
  const float* logits; int ld, R, K, S;
  int col_off[MAX_STAGES];      // class-logit columns of stage k
  int delta_off[MAX_STAGES];    // bbox_pred columns of stage k (-1: none): re-derive the boxes of stage k+1's pseudo GT
  float bw[MAX_STAGES][4];      // ... with the regression weights of refinery k+1
  const float* boxes;
  const int64_t* gt_img; int G;
  const float* img_score; int agnostic;
  float* probs;                 // [S][R][K+1]
  int64_t* pgt_idx; float* pgt_score; float* pgt_box; float* pgt_weight;  // [S][G(,4)], entry k = pseudo GT of stage k; entry 0 is not touched
  const float* pgt0_box; const float* pgt0_weight;                        // stage 0's pseudo GT (drn_wsddn_mil_pgt_fwd)
  MatcherCfg mc; float loss_scale;
  const float* gt_boxes; const int64_t* gt_classes; int Gb;   // real GT: first labelling, done by stage 0's blocks; Gb == -1: skip
  int64_t* labels0; int64_t* matched0; int32_t* counts0;
  int64_t* labels; int64_t* matched;   // [S][R]
  int32_t* counts;                     // [S][3]
  float* weights;                      // [S][R]
  float* stats;                        // [S][6]
  float* loss; int loss_col[MAX_STAGES];
  float* partA; int* partA_idx;        // [S][G][nb]
  float* partB;                        // [S][12][nb]
  uint32_t* counters;                  // [2 S], zero on entry, left at zero
};

template <int CMAX>
__global__ void __launch_bounds__(STAGE_THREADS)
oicr_stages_probs_kernel(const StagesArgs a) {
  __shared__ int scl[MAX_G];
  __shared__ float sv[8][STAGE_THREADS / 32];
  __shared__ int si[8][STAGE_THREADS / 32];
  __shared__ bool is_last;
  const int st = blockIdx.y, C1 = a.K + 1, G = a.G, nb = gridDim.x;
  const bool has_next = st + 1 < a.S;
  for (int g = threadIdx.x; g < G; g += blockDim.x) scl[g] = (int)a.gt_img[g];
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const float* x = a.logits + (long long)min(r, a.R - 1) * a.ld + a.col_off[st];
  float xr[CMAX > 0 ? CMAX : 1];
  if constexpr (CMAX > 0) {
#pragma unroll
    for (int k = 0; k < CMAX; ++k) xr[k] = (k < C1) ? __ldg(x + k) : -INFINITY;
  }
  __syncthreads();
  float pbuf_m = 0.f, pbuf_s = 1.f;
  if (r < a.R) {
    float m = -INFINITY;
    float s = 0.f;
    float* pr = a.probs + ((long long)st * a.R + r) * C1;
    if constexpr (CMAX > 0) {
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1 && xr[k] > m) m = xr[k];
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1) { xr[k] = expf(xr[k] - m); s += xr[k]; }
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1) pr[k] = xr[k] / s;
    } else {
#pragma unroll 8
      for (int k = 0; k < C1; ++k) {
        const float v = __ldg(x + k);
        if (v > m) m = v;
      }
#pragma unroll 8
      for (int k = 0; k < C1; ++k) s += expf(__ldg(x + k) - m);
#pragma unroll 8
      for (int k = 0; k < C1; ++k) pr[k] = expf(__ldg(x + k) - m) / s;
    }
    pbuf_m = m; pbuf_s = s;
  }
  if (!has_next) return;  // block-uniform
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float* part = a.partA + (long long)st * G * nb;
  int* part_idx = a.partA_idx + (long long)st * G * nb;
  // block-partial argmax of the new probabilities for every image-level class, 8 classes per barrier (as oicr_stage_fused_kernel)
  for (int g0 = 0; g0 < G; g0 += 8) {
    const int ng = min(8, G - g0);
    if (g0 > 0) __syncthreads();
    for (int gg = 0; gg < ng; ++gg) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      if (r < a.R) { bv = expf(__ldg(x + scl[g0 + gg]) - pbuf_m) / pbuf_s; bi = r; }  // == probs[r][class g], same expression
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { sv[gg][wid] = bv; si[gg][wid] = bi; }
    }
    __syncthreads();
    if (threadIdx.x < ng) {
      const int gg = threadIdx.x;
      float bv = sv[gg][0];
      int bi = si[gg][0];
      for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
        if (better(sv[gg][i], si[gg][i], bv, bi)) { bv = sv[gg][i]; bi = si[gg][i]; }
      part[(g0 + gg) * nb + blockIdx.x] = bv;
      part_idx[(g0 + gg) * nb + blockIdx.x] = bi;
    }
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counters + st, 1u) == (unsigned)(nb - 1));
  __syncthreads();
  if (!is_last) return;
  if (threadIdx.x == 0) a.counters[st] = 0u;
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    const volatile float* pv = part + g * nb;
    const volatile int* pi = part_idx + g * nb;
    float bv = pv[0];
    int bi = pi[0];
    for (int b = 1; b < nb; ++b) {
      const float ov = pv[b];
      const int oi = pi[b];
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;
    const int c = scl[g];
    const long long o = (long long)(st + 1) * G + g;   // the pseudo GT of the NEXT stage
    a.pgt_idx[o] = bi;
    a.pgt_score[o] = bv;
    Box4 b = {a.boxes[4 * bi + 0], a.boxes[4 * bi + 1], a.boxes[4 * bi + 2], a.boxes[4 * bi + 3]};
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (a.delta_off[st] >= 0) {
      const float* d = a.logits + (long long)bi * a.ld + a.delta_off[st] + (a.agnostic ? 0 : 4 * c);
      d0 = d[0]; d1 = d[1]; d2 = d[2]; d3 = d[3];
    }
    b = apply_deltas_rn(b, d0, d1, d2, d3, a.bw[st][0], a.bw[st][1], a.bw[st][2], a.bw[st][3]);  // fast_rcnn.py:1511-1532
    a.pgt_box[4 * o + 0] = b.x1; a.pgt_box[4 * o + 1] = b.y1;
    a.pgt_box[4 * o + 2] = b.x2; a.pgt_box[4 * o + 3] = b.y2;
  }
}

template <int CMAX>
__global__ void __launch_bounds__(STAGE_THREADS)
oicr_stages_label_ce_kernel(const StagesArgs a) {
  __shared__ float sg[MAX_G][5];
  __shared__ int scl[MAX_G];
  __shared__ float sg0[MAX_G][5];
  __shared__ int scl0[MAX_G];
  __shared__ float sh[32];
  __shared__ int scnt[11];
  __shared__ bool is_last;
  const int st = blockIdx.y, C1 = a.K + 1, K = a.K, G = a.G, nb = gridDim.x;
  __shared__ float spw[MAX_G];
  const float* pgt_box = st == 0 ? a.pgt0_box : a.pgt_box + (long long)st * G * 4;
  // pseudo-GT weight = the image score of its class (get_pgt): stage 0's comes with its pseudo GT, the later stages' are taken
  // from the MIL image scores here -- kernel A, which mined their pseudo GT, depends on the logits only and may run beside the
  // MIL kernels -- and published by block 0
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float w = st == 0 ? a.pgt0_weight[g] : a.img_score[(int)a.gt_img[g]];
    spw[g] = w;
    if (st > 0 && blockIdx.x == 0) a.pgt_weight[(long long)st * G + g] = w;
  }
  if (threadIdx.x < 11) scnt[threadIdx.x] = 0;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float x1 = pgt_box[4 * g], y1 = pgt_box[4 * g + 1], x2 = pgt_box[4 * g + 2], y2 = pgt_box[4 * g + 3];
    sg[g][0] = x1; sg[g][1] = y1; sg[g][2] = x2; sg[g][3] = y2;
    sg[g][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    scl[g] = (int)a.gt_img[g];
  }
  const int Gb = st == 0 ? a.Gb : -1;
  for (int g = threadIdx.x; g < Gb; g += blockDim.x) {
    const float x1 = a.gt_boxes[4 * g], y1 = a.gt_boxes[4 * g + 1], x2 = a.gt_boxes[4 * g + 2], y2 = a.gt_boxes[4 * g + 3];
    sg0[g][0] = x1; sg0[g][1] = y1; sg0[g][2] = x2; sg0[g][3] = y2;
    sg0[g][4] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    scl0[g] = (int)a.gt_classes[g];
  }
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const float* x = a.logits + (long long)min(r, a.R - 1) * a.ld + a.col_off[st];
  float xr[CMAX > 0 ? CMAX : 1];
  if constexpr (CMAX > 0) {
#pragma unroll
    for (int k = 0; k < CMAX; ++k) xr[k] = (k < C1) ? __ldg(x + k) : -INFINITY;
  }
  __syncthreads();
  float lw = 0.f;
  int f_valid = 0, f_acc = 0, f_nfg = 0, f_fgacc = 0, f_fneg = 0;
  int c_fg = 0, c_bg = 0, c_ig = 0, c0_fg = 0, c0_bg = 0, c0_ig = 0;
  if (r < a.R) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(a.boxes) + r);
    const Box4 p = {bb.x, bb.y, bb.z, bb.w};
    const float ap = __fmul_rn(__fsub_rn(p.x2, p.x1), __fsub_rn(p.y2, p.y1));
    if (Gb >= 0) {  // roi_heads_oicr.py:266: labelling against the real GT (logging + proposals' gt fields)
      int l0, m0;
      match_row(p, ap, sg0, scl0, Gb, K, a.mc, l0, m0);
      a.labels0[r] = l0;
      a.matched0[r] = m0;
      if (l0 == -1) c0_ig = 1; else if (l0 == K) c0_bg = 1; else c0_fg = 1;
    }
    int lab, mi;
    match_row(p, ap, sg, scl, G, K, a.mc, lab, mi);
    a.labels[(long long)st * a.R + r] = lab;
    a.matched[(long long)st * a.R + r] = mi;
    if (lab == -1) c_ig = 1; else if (lab == K) c_bg = 1; else c_fg = 1;
    float m = -INFINITY;
    int am = 0;
    float s = 0.f;
    if constexpr (CMAX > 0) {   // the row's softmax statistics, in the order of the kernel that stored the probabilities
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1 && xr[k] > m) { m = xr[k]; am = k; }
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < C1) s += expf(xr[k] - m);
    } else {
#pragma unroll 8
      for (int k = 0; k < C1; ++k) {
        const float v = __ldg(x + k);
        if (v > m) { m = v; am = k; }
      }
#pragma unroll 8
      for (int k = 0; k < C1; ++k) s += expf(__ldg(x + k) - m);
    }
    float w = spw[mi];
    if (lab == -1) w = 0.f;
    a.weights[(long long)st * a.R + r] = w;
    if (w > 1e-12f) f_valid = 1;
    if (lab >= 0) lw = (-((__ldg(x + lab) - m) - logf(s))) * w;
    const bool fg = lab >= 0 && lab < K;
    if (am == lab) f_acc = 1;
    if (fg) {
      f_nfg = 1;
      if (am == lab) f_fgacc = 1;
      if (am == K) f_fneg = 1;
    }
  }
  const int lane = threadIdx.x & 31;
  {
    const int fl[11] = {f_valid, f_acc, f_nfg, f_fgacc, f_fneg, c_fg, c_bg, c_ig, c0_fg, c0_bg, c0_ig};
    const int nfl = Gb >= 0 ? 11 : 8;
#pragma unroll
    for (int j = 0; j < 11; ++j) {
      if (j < nfl) {
        const int v = __reduce_add_sync(0xffffffffu, fl[j]);
        if (lane == 0 && v) atomicAdd(&scnt[j], v);
      }
    }
  }
  float* part = a.partB + (long long)st * 12 * nb;
  float v = block_sum(lw, sh);   // same tree as oicr_stage_kernel (its barriers also publish scnt)
  if (threadIdx.x == 0) part[0 * nb + blockIdx.x] = v;
  if (threadIdx.x < (Gb >= 0 ? 11 : 8)) part[(1 + threadIdx.x) * nb + blockIdx.x] = (float)scnt[threadIdx.x];
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.counters + a.S + st, 1u) == (unsigned)(nb - 1));
  __syncthreads();
  if (!is_last) return;
  const int nsum = Gb >= 0 ? 12 : 9;
  if (threadIdx.x < nsum) {
    float s = 0.f;
    const volatile float* pp = part + threadIdx.x * nb;
    for (int b = 0; b < nb; ++b) s += pp[b];  // fixed order
    sh[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.loss[a.loss_col[st]] = (sh[0] / sh[1]) * a.loss_scale;
    float* stats = a.stats + 6 * st;
    stats[0] = sh[2]; stats[1] = sh[3]; stats[2] = sh[4]; stats[3] = sh[5];
    stats[4] = sh[0]; stats[5] = sh[1];
    int32_t* counts = a.counts + 3 * st;
    counts[0] = (int)sh[6]; counts[1] = (int)sh[7]; counts[2] = (int)sh[8];
    if (Gb >= 0) { a.counts0[0] = (int)sh[9]; a.counts0[1] = (int)sh[10]; a.counts0[2] = (int)sh[11]; }
    a.counters[a.S + st] = 0u;
  }
}

// ---------------------------------------------------------------- box-regression loss (reg/ configs)
__global__ void __launch_bounds__(STAGE_THREADS)
oicr_boxreg_kernel(const float* __restrict__ deltas, int ld, int col_off, int R, int K, int agnostic,
                   const float* __restrict__ boxes, const float* __restrict__ pgt_box,
                   const int64_t* __restrict__ labels, const int64_t* __restrict__ matched, float wx,
                   float wy, float ww, float wh, float beta, float loss_scale, float* __restrict__ loss,
                   float* __restrict__ part, uint32_t* __restrict__ counter) {
  __shared__ float sh[32];
  __shared__ bool is_last;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float l = 0.f;
  if (r < R) {
    const int lab = (int)labels[r];
    if (lab >= 0 && lab < K) {
      const float* pb = boxes + 4 * (long long)r;
      const float* gb = pgt_box + 4 * matched[r];
      // get_deltas (box_regression.py:38-71)
      const float sw = __fsub_rn(pb[2], pb[0]), shh = __fsub_rn(pb[3], pb[1]);
      const float scx = __fadd_rn(pb[0], __fmul_rn(0.5f, sw)), scy = __fadd_rn(pb[1], __fmul_rn(0.5f, shh));
      const float tw = __fsub_rn(gb[2], gb[0]), th = __fsub_rn(gb[3], gb[1]);
      const float tcx = __fadd_rn(gb[0], __fmul_rn(0.5f, tw)), tcy = __fadd_rn(gb[1], __fmul_rn(0.5f, th));
      float t[4];
      t[0] = __fdiv_rn(__fmul_rn(wx, __fsub_rn(tcx, scx)), sw);
      t[1] = __fdiv_rn(__fmul_rn(wy, __fsub_rn(tcy, scy)), shh);
      t[2] = __fmul_rn(ww, logf(__fdiv_rn(tw, sw)));
      t[3] = __fmul_rn(wh, logf(__fdiv_rn(th, shh)));
      const float* d = deltas + (long long)r * ld + col_off + (agnostic ? 0 : 4 * lab);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float n = fabsf(d[j] - t[j]);
        l += (beta < 1e-5f) ? n : (n < beta ? 0.5f * n * n / beta : n - 0.5f * beta);
      }
    }
  }
  const int nb = gridDim.x;
  const float v = block_sum(l, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = v;
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == (unsigned)(nb - 1));
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    float s = 0.f;
    const volatile float* pp = part;
    for (int b = 0; b < nb; ++b) s += pp[b];
    loss[0] = s / (float)R * loss_scale;
    *counter = 0u;
  }
}

// ---------------------------------------------------------------- inference scores / boxes
struct InferCfg { int S; int col[8]; int dcol[8]; };

__global__ void __launch_bounds__(256)
oicr_infer_kernel(const float* __restrict__ logits, int ld, int R, int K, int nreg, InferCfg ic,
                  const float* __restrict__ boxes, float wx, float wy, float ww, float wh,
                  float* __restrict__ all_scores, float* __restrict__ all_boxes) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int C1 = K + 1;
  float* o = all_scores + (long long)r * C1;
  for (int k = 0; k < C1; ++k) o[k] = 0.f;
  for (int s = 0; s < ic.S; ++s) {
    const float* x = logits + (long long)r * ld + ic.col[s];
    float m = -INFINITY;
    for (int k = 0; k < C1; ++k) m = fmaxf(m, x[k]);
    float z = 0.f;
    for (int k = 0; k < C1; ++k) z += expf(x[k] - m);
    for (int k = 0; k < C1; ++k) o[k] += expf(x[k] - m) / z;   // probs += softmax_k
  }
  const float fS = (float)ic.S;
  for (int k = 0; k < C1; ++k) o[k] = o[k] / fS;
  const Box4 b = {boxes[4 * r], boxes[4 * r + 1], boxes[4 * r + 2], boxes[4 * r + 3]};
  float* ob = all_boxes + (long long)r * nreg * 4;
  for (int k = 0; k < nreg; ++k) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
    bool any = false;
    for (int s = 0; s < ic.S; ++s)
      if (ic.dcol[s] >= 0) {
        const float* dd = logits + (long long)r * ld + ic.dcol[s] + 4 * k;
        d[0] += dd[0]; d[1] += dd[1]; d[2] += dd[2]; d[3] += dd[3];
        any = true;
      }
    if (any) { d[0] /= fS; d[1] /= fS; d[2] /= fS; d[3] /= fS; }
    const Box4 q = apply_deltas_rn(b, d[0], d[1], d[2], d[3], wx, wy, ww, wh);
    ob[4 * k + 0] = q.x1; ob[4 * k + 1] = q.y1; ob[4 * k + 2] = q.x2; ob[4 * k + 3] = q.y2;
  }
}

// ---------------------------------------------------------------- dropout (train-mode fc6/fc7)
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  // splitmix64 finaliser: counter-based, stateless RNG (one draw per element)
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(T* __restrict__ x, long long n, uint32_t keep_thresh, float inv_keep, uint64_t seed,
               const uint64_t* __restrict__ seed_dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (seed_dev) seed += __ldg(seed_dev);
  const bool keep = mix32(seed * 0x100000001B3ull + (uint64_t)i) < keep_thresh;
  float v;
  if constexpr (sizeof(T) == 4) v = x[i]; else v = __bfloat162float(x[i]);
  v = keep ? v * inv_keep : 0.f;
  if constexpr (sizeof(T) == 4) x[i] = v; else x[i] = __float2bfloat16(v);
}

}  // namespace drn

using namespace drn;

extern "C" {

int drn_wsddn_mil_fwd(const float* logits, int ld, int R, int K, int cls_off, int det_off,
                      const float* gt_onehot, int mean_loss, float loss_scale, float* scores,
                      float* img_score, float* loss, float* row_ws, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && gt_onehot && scores && img_score && loss && row_ws, "wsddn_mil: null pointer");
  DRN_CHECK_ARG(R > 0 && K > 0, "wsddn_mil: R=%d K=%d", R, K);
  DRN_CHECK_ARG(cls_off + K <= ld && det_off + K <= ld, "wsddn_mil: column ranges exceed ld=%d", ld);
  cudaStream_t st = (cudaStream_t)stream;
  float* rowmax = row_ws;
  float* rowsum = row_ws + R;
  float* bce = row_ws + 2 * (long long)R;  // [K]
  mil_rowstats_kernel<<<cdiv(R, 256), 256, 0, st>>>(logits, ld, R, K, cls_off, rowmax, rowsum);
  mil_cols_kernel<<<K, 512, 0, st>>>(logits, ld, R, K, cls_off, det_off, rowmax, rowsum, gt_onehot,
                                     scores, img_score, bce);
  mil_finalize_kernel<<<1, 32, 0, st>>>(bce, K, mean_loss, loss_scale, loss);
  DRN_CHECK_LAUNCH("wsddn_mil");
  return 0;
}

int drn_oicr_pgt(const float* prev_scores, int ld_prev, int R, const float* boxes,
                 const int64_t* gt_classes, int G, const float* img_score, int rederive,
                 const float* deltas, int ld_deltas, int cls_agnostic, const float* bbox_w,
                 int64_t* pgt_idx, float* pgt_score, float* pgt_box, float* pgt_weight,
                 drn_stream_t stream) {
  DRN_CHECK_ARG(prev_scores && boxes && gt_classes && img_score && pgt_idx && pgt_score && pgt_box && pgt_weight,
                "oicr_pgt: null pointer");
  DRN_CHECK_ARG(G > 0 && R > 0, "oicr_pgt: needs at least one image-level class and one proposal (G=%d R=%d)", G, R);
  DRN_CHECK_ARG(bbox_w || !rederive, "oicr_pgt: bbox weights required");
  const float wx = bbox_w ? bbox_w[0] : 1.f, wy = bbox_w ? bbox_w[1] : 1.f;
  const float ww = bbox_w ? bbox_w[2] : 1.f, wh = bbox_w ? bbox_w[3] : 1.f;
  oicr_pgt_kernel<<<G, 256, 0, (cudaStream_t)stream>>>(prev_scores, ld_prev, R, boxes, gt_classes, img_score,
      rederive, deltas, ld_deltas, cls_agnostic, wx, wy, ww, wh, pgt_idx, pgt_score, pgt_box, pgt_weight);
  DRN_CHECK_LAUNCH("oicr_pgt");
  return 0;
}

int drn_label_proposals(const float* boxes, int R, const float* gt_boxes, const int64_t* gt_classes,
                        int G, int K, const float* thresholds, const int* labels_cfg, int nthr,
                        int64_t* labels, int64_t* matched_idx, int32_t* counts, drn_stream_t stream) {
  DRN_CHECK_ARG(boxes && labels && matched_idx && counts, "label_proposals: null pointer");
  DRN_CHECK_ARG(G == 0 || (gt_boxes && gt_classes), "label_proposals: null gt");
  DRN_CHECK_ARG(G <= MAX_G, "label_proposals: G=%d exceeds %d", G, MAX_G);
  DRN_CHECK_ARG(nthr >= 0 && nthr <= 4, "label_proposals: %d thresholds (max 4)", nthr);
  if (R == 0) return 0;
  MatcherCfg mc;
  mc.nthr = nthr;
  for (int i = 0; i < nthr; ++i) mc.thr[i] = thresholds[i];
  for (int i = 0; i <= nthr; ++i) mc.lab[i] = labels_cfg[i];
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(counts, 0, 3 * sizeof(int32_t), st);
  label_proposals_kernel<<<cdiv(R, 256), 256, 0, st>>>(boxes, R, gt_boxes, gt_classes, G, K, mc, labels,
                                                       matched_idx, counts);
  DRN_CHECK_LAUNCH("label_proposals");
  return 0;
}

int drn_oicr_stage_fwd(const float* logits, int ld, int col_off, int R, int K, const int64_t* labels,
                       const int64_t* matched_idx, const float* pgt_weight, int G, float loss_scale,
                       float* probs, float* loss, float* stats, float* weights, float* part_ws,
                       uint32_t* counter, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && labels && matched_idx && pgt_weight && probs && loss && stats && part_ws && counter,
                "oicr_stage: null pointer");
  DRN_CHECK_ARG(R > 0 && G > 0, "oicr_stage: R=%d G=%d", R, G);
  DRN_CHECK_ARG(col_off + K + 1 <= ld, "oicr_stage: columns exceed ld=%d", ld);
  oicr_stage_kernel<<<cdiv(R, STAGE_THREADS), STAGE_THREADS, 0, (cudaStream_t)stream>>>(logits, ld, col_off, R, K,
      labels, matched_idx, pgt_weight, loss_scale, probs, loss, stats, weights, part_ws, counter);
  DRN_CHECK_LAUNCH("oicr_stage");
  return 0;
}

int drn_oicr_boxreg_loss(const float* deltas, int ld, int col_off, int R, int K, int cls_agnostic,
                         const float* boxes, const float* pgt_box, const int64_t* labels,
                         const int64_t* matched_idx, const float* bbox_w, float beta, float loss_scale,
                         float* loss, float* part_ws, uint32_t* counter, drn_stream_t stream) {
  DRN_CHECK_ARG(deltas && boxes && pgt_box && labels && matched_idx && bbox_w && loss && part_ws && counter,
                "oicr_boxreg: null pointer");
  DRN_CHECK_ARG(R > 0, "oicr_boxreg: R=%d", R);
  oicr_boxreg_kernel<<<cdiv(R, STAGE_THREADS), STAGE_THREADS, 0, (cudaStream_t)stream>>>(deltas, ld, col_off, R, K,
      cls_agnostic, boxes, pgt_box, labels, matched_idx, bbox_w[0], bbox_w[1], bbox_w[2], bbox_w[3], beta,
      loss_scale, loss, part_ws, counter);
  DRN_CHECK_LAUNCH("oicr_boxreg");
  return 0;
}

int drn_dropout_inplace(void* x, int64_t n, int dtype, float p, uint64_t seed, const uint64_t* seed_dev,
                        drn_stream_t stream) {
  DRN_CHECK_ARG(x || n == 0, "dropout: null pointer");
  DRN_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p=%f", p);
  if (n == 0 || p == 0.f) return 0;
  const double keep = 1.0 - (double)p;
  const uint32_t thr = (uint32_t)(keep * 4294967295.0);
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == DRN_BF16)
    dropout_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, n, thr, (float)(1.0 / keep), seed, seed_dev);
  else
    dropout_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, n, thr, (float)(1.0 / keep), seed, seed_dev);
  DRN_CHECK_LAUNCH("dropout");
  return 0;
}

int drn_wsddn_mil_pgt_fwd(const float* logits, int ld, int R, int K, int cls_off, int det_off,
                          const float* gt_onehot, int mean_loss, float loss_scale, const float* boxes,
                          const int64_t* gt_classes, int G, float* scores, float* img_score, float* loss,
                          int64_t* pgt_idx, float* pgt_score, float* pgt_box, float* pgt_weight, float* ws,
                          uint32_t* counter, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && gt_onehot && boxes && scores && img_score && loss && ws && counter, "wsddn_mil_pgt: null pointer");
  DRN_CHECK_ARG(G == 0 || (gt_classes && pgt_idx && pgt_score && pgt_box && pgt_weight), "wsddn_mil_pgt: null pseudo-GT outputs");
  DRN_CHECK_ARG(R > 0 && K > 0, "wsddn_mil_pgt: R=%d K=%d", R, K);
  DRN_CHECK_ARG(cls_off + K <= ld && det_off + K <= ld, "wsddn_mil_pgt: column ranges exceed ld=%d", ld);
  cudaStream_t st = (cudaStream_t)stream;
  float* rowmax = ws;
  float* rowsum = ws + R;
  float* bce = ws + 2 * (long long)R;  // [K]
  mil_rowstats_kernel<<<cdiv(R, 256), 256, 0, st>>>(logits, ld, R, K, cls_off, rowmax, rowsum);
  mil_pgt_fused_kernel<<<K, 512, 0, st>>>(logits, ld, R, K, cls_off, det_off, rowmax, rowsum, gt_onehot, mean_loss,
      loss_scale, boxes, gt_classes, G, scores, img_score, loss, pgt_idx, pgt_score, pgt_box, pgt_weight, bce, counter);
  DRN_CHECK_LAUNCH("wsddn_mil_pgt");
  return 0;
}

int drn_oicr_stage_fused_fwd(const float* logits, int ld, int col_off, int R, int K, const float* boxes,
                             const int64_t* gt_classes_img, int G, const float* pgt_box, const float* pgt_weight,
                             const float* thresholds, const int* labels_cfg, int nthr, float loss_scale,
                             const float* gt_boxes, const int64_t* gt_classes, int Gb, int64_t* labels0,
                             int64_t* matched0, int32_t* counts0, int64_t* labels, int64_t* matched_idx,
                             int32_t* counts, float* probs, float* loss, float* stats, float* weights,
                             const float* img_score, const float* deltas, int ld_deltas, int cls_agnostic,
                             const float* bbox_w, int64_t* next_pgt_idx, float* next_pgt_score,
                             float* next_pgt_box, float* next_pgt_weight, float* part_ws, uint32_t* counter,
                             drn_stream_t stream) {
  DRN_CHECK_ARG(logits && boxes && gt_classes_img && pgt_box && pgt_weight && labels && matched_idx && counts && probs &&
                    loss && stats && part_ws && counter, "oicr_stage_fused: null pointer");
  DRN_CHECK_ARG(R > 0 && G > 0 && G <= MAX_G, "oicr_stage_fused: R=%d G=%d (max %d)", R, G, MAX_G);
  DRN_CHECK_ARG(col_off + K + 1 <= ld, "oicr_stage_fused: columns exceed ld=%d", ld);
  DRN_CHECK_ARG(nthr >= 0 && nthr <= 4, "oicr_stage_fused: %d thresholds (max 4)", nthr);
  DRN_CHECK_ARG(Gb <= MAX_G, "oicr_stage_fused: Gb=%d exceeds %d", Gb, MAX_G);
  DRN_CHECK_ARG(Gb < 0 || (labels0 && matched0 && counts0 && (Gb == 0 || (gt_boxes && gt_classes))),
                "oicr_stage_fused: first-labelling buffers missing");
  const int has_next = next_pgt_idx != nullptr;
  DRN_CHECK_ARG(!has_next || (img_score && bbox_w && next_pgt_score && next_pgt_box && next_pgt_weight),
                "oicr_stage_fused: next pseudo-GT buffers missing");
  StageFusedArgs a;
  a.logits = logits; a.ld = ld; a.col_off = col_off; a.R = R; a.K = K; a.boxes = boxes;
  a.gt_img = gt_classes_img; a.G = G; a.pgt_box = pgt_box; a.pgt_weight = pgt_weight;
  a.mc.nthr = nthr;
  for (int i = 0; i < nthr; ++i) a.mc.thr[i] = thresholds[i];
  for (int i = 0; i <= nthr; ++i) a.mc.lab[i] = labels_cfg[i];
  a.loss_scale = loss_scale;
  a.gt_boxes = gt_boxes; a.gt_classes = gt_classes; a.Gb = Gb;
  a.labels0 = labels0; a.matched0 = matched0; a.counts0 = counts0;
  a.labels = labels; a.matched = matched_idx; a.counts = counts;
  a.probs = probs; a.loss = loss; a.stats = stats; a.weights = weights;
  a.has_next = has_next; a.img_score = img_score; a.deltas = deltas; a.ld_deltas = ld_deltas; a.agnostic = cls_agnostic;
  a.wx = bbox_w ? bbox_w[0] : 1.f; a.wy = bbox_w ? bbox_w[1] : 1.f; a.ww = bbox_w ? bbox_w[2] : 1.f; a.wh = bbox_w ? bbox_w[3] : 1.f;
  a.next_idx = next_pgt_idx; a.next_score = next_pgt_score; a.next_box = next_pgt_box; a.next_weight = next_pgt_weight;
  const int nb = cdiv(R, STAGE_THREADS);
  a.part = part_ws;                                        // [(12 + G) * nb] floats ...
  a.part_idx = reinterpret_cast<int*>(part_ws + (size_t)(12 + G) * nb);  // ... followed by [G * nb] ints
  a.counter = counter;
  if (K + 1 <= 32) oicr_stage_fused_kernel<32><<<nb, STAGE_THREADS, 0, (cudaStream_t)stream>>>(a);
  else oicr_stage_fused_kernel<0><<<nb, STAGE_THREADS, 0, (cudaStream_t)stream>>>(a);
  DRN_CHECK_LAUNCH("oicr_stage_fused");
  return 0;
}

int drn_oicr_stages_fwd(const float* logits, int ld, int R, int K, int S, const int* col_offs, const int* delta_offs,
                        const float* bbox_w, const float* boxes, const int64_t* gt_classes_img, int G,
                        const float* img_score, int cls_agnostic, const float* pgt0_box, const float* pgt0_weight,
                        const float* thresholds, const int* labels_cfg, int nthr, float loss_scale,
                        const float* gt_boxes, const int64_t* gt_classes, int Gb, int64_t* labels0, int64_t* matched0,
                        int32_t* counts0, float* probs, int64_t* pgt_idx, float* pgt_score, float* pgt_box,
                        float* pgt_weight, int64_t* labels, int64_t* matched_idx, int32_t* counts, float* weights,
                        float* stats, float* loss, const int* loss_cols, float* part_ws, uint32_t* counters, int phases,
                        drn_stream_t stream) {
  DRN_CHECK_ARG(phases >= 1 && phases <= 3, "oicr_stages: phases=%d (1 = probabilities + pseudo GT, 2 = labelling + CE, 3 = both)", phases);
  DRN_CHECK_ARG(logits && col_offs && boxes && gt_classes_img && probs && labels && matched_idx && counts && weights && stats &&
                    loss_cols && part_ws && counters, "oicr_stages: null pointer");
  DRN_CHECK_ARG(!(phases & 2) || (img_score && pgt0_box && pgt0_weight && loss), "oicr_stages: launch 2 needs img_score, pgt0_*, loss");
  DRN_CHECK_ARG(S >= 1 && S <= MAX_STAGES, "oicr_stages: S=%d (1..%d)", S, MAX_STAGES);
  DRN_CHECK_ARG(R > 0 && G > 0 && G <= MAX_G, "oicr_stages: R=%d G=%d (max %d)", R, G, MAX_G);
  DRN_CHECK_ARG(nthr >= 0 && nthr <= 4, "oicr_stages: %d thresholds (max 4)", nthr);
  DRN_CHECK_ARG(Gb <= MAX_G, "oicr_stages: Gb=%d exceeds %d", Gb, MAX_G);
  DRN_CHECK_ARG(Gb < 0 || (labels0 && matched0 && counts0 && (Gb == 0 || (gt_boxes && gt_classes))),
                "oicr_stages: first-labelling buffers missing");
  DRN_CHECK_ARG(S == 1 || (pgt_idx && pgt_score && pgt_box && pgt_weight && bbox_w), "oicr_stages: pseudo-GT buffers missing");
  StagesArgs a;
  a.logits = logits; a.ld = ld; a.R = R; a.K = K; a.S = S;
  for (int k = 0; k < S; ++k) {
    DRN_CHECK_ARG(col_offs[k] >= 0 && col_offs[k] + K + 1 <= ld, "oicr_stages: columns of stage %d exceed ld=%d", k, ld);
    a.col_off[k] = col_offs[k];
    a.delta_off[k] = delta_offs ? delta_offs[k] : -1;
    a.loss_col[k] = loss_cols[k];
    for (int j = 0; j < 4; ++j) a.bw[k][j] = bbox_w ? bbox_w[4 * k + j] : 1.f;
  }
  a.boxes = boxes; a.gt_img = gt_classes_img; a.G = G; a.img_score = img_score; a.agnostic = cls_agnostic;
  a.probs = probs; a.pgt_idx = pgt_idx; a.pgt_score = pgt_score; a.pgt_box = pgt_box; a.pgt_weight = pgt_weight;
  a.pgt0_box = pgt0_box; a.pgt0_weight = pgt0_weight;
  a.mc.nthr = nthr;
  for (int i = 0; i < nthr; ++i) a.mc.thr[i] = thresholds[i];
  for (int i = 0; i <= nthr; ++i) a.mc.lab[i] = labels_cfg[i];
  a.loss_scale = loss_scale;
  a.gt_boxes = gt_boxes; a.gt_classes = gt_classes; a.Gb = Gb;
  a.labels0 = labels0; a.matched0 = matched0; a.counts0 = counts0;
  a.labels = labels; a.matched = matched_idx; a.counts = counts; a.weights = weights; a.stats = stats; a.loss = loss;
  const int nb = cdiv(R, STAGE_THREADS);
  a.partA = part_ws;                                                   // [S][G][nb] floats
  a.partA_idx = reinterpret_cast<int*>(part_ws + (size_t)S * G * nb);  // [S][G][nb] ints
  a.partB = part_ws + (size_t)2 * S * G * nb;                          // [S][12][nb] floats
  a.counters = counters;
  const dim3 grid(nb, S);
  cudaStream_t st = (cudaStream_t)stream;
  if (K + 1 <= 32) {
    if (phases & 1) oicr_stages_probs_kernel<32><<<grid, STAGE_THREADS, 0, st>>>(a);
    if (phases & 2) oicr_stages_label_ce_kernel<32><<<grid, STAGE_THREADS, 0, st>>>(a);
  } else {
    if (phases & 1) oicr_stages_probs_kernel<0><<<grid, STAGE_THREADS, 0, st>>>(a);
    if (phases & 2) oicr_stages_label_ce_kernel<0><<<grid, STAGE_THREADS, 0, st>>>(a);
  }
  DRN_CHECK_LAUNCH("oicr_stages");
  return 0;
}

int drn_oicr_infer(const float* logits, int ld, int R, int K, int nreg, int S, const int* col_offs,
                   const int* delta_offs, const float* boxes, const float* bbox_w, float* all_scores,
                   float* all_boxes, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && col_offs && boxes && bbox_w && all_scores && all_boxes, "oicr_infer: null pointer");
  DRN_CHECK_ARG(S >= 1 && S <= 8, "oicr_infer: S=%d (1..8)", S);
  if (R == 0) return 0;
  InferCfg ic;
  ic.S = S;
  for (int s = 0; s < S; ++s) { ic.col[s] = col_offs[s]; ic.dcol[s] = delta_offs ? delta_offs[s] : -1; }
  oicr_infer_kernel<<<cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(logits, ld, R, K, nreg, ic, boxes, bbox_w[0],
      bbox_w[1], bbox_w[2], bbox_w[3], all_scores, all_boxes);
  DRN_CHECK_LAUNCH("oicr_infer");
  return 0;
}

}  // extern "C"
