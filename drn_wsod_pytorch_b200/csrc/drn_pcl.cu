// PCL refinement stage on the device: everything of projects/WSL/wsl/modeling/roi_heads/third_party/pcl.py
// that is per-proposal arithmetic (`_get_proposal_clusters`, :148-200) and the reference's pcl_loss op
// (wsl/layers/csrc/pcl_loss/pcl_loss_cpu.cpp:8-62 forward, :64-115 backward; its CUDA kernel is dead code behind `&& false`,
// pcl_loss.h:64,101) with the scaling of wsl/layers/pcl_loss.py:52,119.  The cluster CENTRES (k-means over the class scores +
// a greedy IoU-graph cover, :62-145: data-dependent sizes, a handful of boxes) are mined on the host as in the reference and
// arrive here as P <= ~100 boxes / classes / scores.
//   pcl_assign_kernel   one thread per proposal: softmax of the stage's logits, IoU against the centres (detectron2
//                       pairwise_iou arithmetic), first-maximum assignment, label / loss weight / cluster index.
//   pcl_cluster_kernel  one CTA per cluster (+ one for the background term): member count, summed weights, mean clipped
//                       probability, the cluster's loss term -- fixed-order reductions, bit-reproducible; the last CTA to finish
//                       adds the terms up (in cluster order) into the stage loss.
//   pcl_bwd_kernel      d loss / d logits through the softmax (one non-zero d loss / d prob entry per proposal).
#include "common.cuh"

namespace drn {

constexpr int PCL_THREADS = 256;

__device__ __forceinline__ float pcl_block_sum(float v, float* sh) {  // fixed-order tree over 256 threads
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = PCL_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const float r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(PCL_THREADS)
pcl_assign_kernel(const float* __restrict__ logits, int ld, int col_off, int R, int K, const float* __restrict__ boxes,
                  const float* __restrict__ cbox, const int* __restrict__ ccls, const float* __restrict__ cscore, int P,
                  float* __restrict__ probs, int* __restrict__ labels, float* __restrict__ weights, int* __restrict__ assign) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int C1 = K + 1;
  const float* x = logits + (long long)r * ld + col_off;
  float m = -INFINITY;
  for (int k = 0; k < C1; ++k) m = fmaxf(m, __ldg(x + k));
  float s = 0.f;
  for (int k = 0; k < C1; ++k) s += expf(__ldg(x + k) - m);
  float* p = probs + (long long)r * C1;
  for (int k = 0; k < C1; ++k) p[k] = expf(__ldg(x + k) - m) / s;
  // pairwise_iou(rois, centres) (detectron2/structures/boxes.py:329-361), numpy argmax / max over the centres (:166-167)
  const float x1 = boxes[4 * r], y1 = boxes[4 * r + 1], x2 = boxes[4 * r + 2], y2 = boxes[4 * r + 3];
  const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  float best = -1.f;
  int arg = 0;
  for (int i = 0; i < P; ++i) {
    const float cx1 = cbox[4 * i], cy1 = cbox[4 * i + 1], cx2 = cbox[4 * i + 2], cy2 = cbox[4 * i + 3];
    const float carea = __fmul_rn(__fsub_rn(cx2, cx1), __fsub_rn(cy2, cy1));
    const float iw = fmaxf(__fsub_rn(fminf(x2, cx2), fmaxf(x1, cx1)), 0.f);
    const float ih = fmaxf(__fsub_rn(fminf(y2, cy2), fmaxf(y1, cy1)), 0.f);
    const float inter = __fmul_rn(iw, ih);
    const float iou = inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(area, carea), inter)) : 0.f;
    if (iou > best) { best = iou; arg = i; }  // first maximum
  }
  const bool bg = best < 0.5f;                      // cfg_TRAIN_FG_THRESH (:16, :175, :180-181)
  labels[r] = bg ? 0 : ccls[arg];
  weights[r] = best < 0.1f ? 0.f : cscore[arg];     // cfg_TRAIN_BG_THRESH (:17, :177-178)
  assign[r] = bg ? -1 : arg;
}

// grid = P + 1.  CTA i < P: cluster i; CTA P: the background term.  terms[i] = the CTA's loss term (un-normalised).
__global__ void __launch_bounds__(PCL_THREADS)
pcl_cluster_kernel(const float* __restrict__ probs, int R, int K, const int* __restrict__ labels, const float* __restrict__ weights,
                   const int* __restrict__ assign, const int* __restrict__ ccls, int P, float loss_scale,
                   float* __restrict__ pc_probs, float* __restrict__ pc_count, float* __restrict__ img_w,
                   float* __restrict__ terms, float* __restrict__ loss, unsigned int* __restrict__ counter) {
  __shared__ float sh[PCL_THREADS];
  __shared__ bool is_last;
  const int C1 = K + 1;
  const int i = blockIdx.x;
  if (i < P) {
    const int c = ccls[i];
    float n = 0.f, sw = 0.f, sp = 0.f;
    for (int r = threadIdx.x; r < R; r += PCL_THREADS) {
      if (assign[r] == i) {
        n += 1.f;
        sw += weights[r];
        sp += fminf(fmaxf(probs[(long long)r * C1 + c], 1e-9f), 1.f - 1e-9f);  // cls_prob_new is clipped before the mining (:40-41)
      }
    }
    n = pcl_block_sum(n, sh);
    sw = pcl_block_sum(sw, sh);
    sp = pcl_block_sum(sp, sh);
    if (threadIdx.x == 0) {
      const float mean = sp / n;  // np.average of an empty cluster is NaN in the reference too; fmaxf(NaN, eps) = eps below
      pc_count[i] = n;
      img_w[i] = sw;
      pc_probs[i] = mean;
      terms[i] = -sw * logf(fmaxf(mean, 1e-6f));  // pcl_loss_cpu.cpp:50-53
    }
  } else {
    float t = 0.f;
    for (int r = threadIdx.x; r < R; r += PCL_THREADS)
      if (labels[r] == 0) t -= weights[r] * logf(fmaxf(probs[(long long)r * C1], 1e-6f));  // :41-45
    t = pcl_block_sum(t, sh);
    if (threadIdx.x == 0) terms[P] = t;
  }
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == (unsigned)P);
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    const volatile float* tt = terms;
    float s = tt[P];
    for (int j = 0; j < P; ++j) s += tt[j];       // cluster order
    loss[0] = s / (float)R * loss_scale;          // wsl/layers/pcl_loss.py:52
    *counter = 0u;
  }
}

__global__ void __launch_bounds__(PCL_THREADS)
pcl_bwd_kernel(const float* __restrict__ probs, int R, int K, const int* __restrict__ labels, const float* __restrict__ weights,
               const int* __restrict__ assign, const float* __restrict__ pc_probs, const float* __restrict__ pc_count,
               const float* __restrict__ img_w, float loss_scale, const float* __restrict__ grad_loss, int col_off, int ld,
               float* __restrict__ dlogits) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int C1 = K + 1;
  const float* p = probs + (long long)r * C1;
  const int lab = labels[r];
  int j;
  float g;
  if (lab == 0) {  // pcl_loss_cpu.cpp:96-99
    j = 0;
    g = -weights[r] / fmaxf(p[0], 1e-5f);
  } else {         // :101-110
    const int a = assign[r];
    j = lab;
    g = -img_w[a] / fmaxf(pc_count[a] * pc_probs[a], 1e-5f);
  }
  g *= grad_loss[0] * loss_scale / (float)R;       // wsl/layers/pcl_loss.py:119 and the upstream gradient
  const float gp = g * p[j];
  float* d = dlogits + (long long)r * ld + col_off;
  for (int c = 0; c < C1; ++c) d[c] = p[c] * ((c == j ? g : 0.f) - gp);  // softmax backward with one non-zero dL/dp entry
}

}  // namespace drn

using namespace drn;

extern "C" {

int drn_pcl_stage_fwd(const float* logits, int ld, int R, int K, int col_off, const float* boxes, const float* center_boxes,
                      const int* center_classes, const float* center_scores, int P, float loss_scale, float* probs, int* labels,
                      float* weights, int* assignment, float* pc_probs, float* pc_count, float* img_cls_loss_weights, float* loss,
                      float* terms_ws, unsigned int* counter, drn_stream_t stream) {
  DRN_CHECK_ARG(logits && boxes && center_boxes && center_classes && center_scores && probs && labels && weights && assignment &&
                    pc_probs && pc_count && img_cls_loss_weights && loss && terms_ws && counter, "pcl_stage_fwd: null pointer");
  DRN_CHECK_ARG(R > 0 && K > 0 && P > 0 && col_off >= 0 && col_off + K + 1 <= ld, "pcl_stage_fwd: R=%d K=%d P=%d col_off=%d ld=%d", R, K, P, col_off, ld);
  cudaStream_t st = (cudaStream_t)stream;
  pcl_assign_kernel<<<cdiv(R, PCL_THREADS), PCL_THREADS, 0, st>>>(logits, ld, col_off, R, K, boxes, center_boxes, center_classes,
                                                                  center_scores, P, probs, labels, weights, assignment);
  DRN_CHECK_LAUNCH("pcl_assign");
  pcl_cluster_kernel<<<P + 1, PCL_THREADS, 0, st>>>(probs, R, K, labels, weights, assignment, center_classes, P, loss_scale, pc_probs,
                                                    pc_count, img_cls_loss_weights, terms_ws, loss, counter);
  DRN_CHECK_LAUNCH("pcl_cluster");
  return 0;
}

int drn_pcl_stage_bwd(const float* probs, int R, int K, const int* labels, const float* weights, const int* assignment,
                      const float* pc_probs, const float* pc_count, const float* img_cls_loss_weights, float loss_scale,
                      const float* grad_loss, int col_off, int ld, float* dlogits, drn_stream_t stream) {
  DRN_CHECK_ARG(probs && labels && weights && assignment && pc_probs && pc_count && img_cls_loss_weights && grad_loss && dlogits,
                "pcl_stage_bwd: null pointer");
  DRN_CHECK_ARG(R > 0 && K > 0 && col_off >= 0 && col_off + K + 1 <= ld, "pcl_stage_bwd: R=%d K=%d col_off=%d ld=%d", R, K, col_off, ld);
  pcl_bwd_kernel<<<cdiv(R, PCL_THREADS), PCL_THREADS, 0, (cudaStream_t)stream>>>(probs, R, K, labels, weights, assignment, pc_probs,
                                                                                 pc_count, img_cls_loss_weights, loss_scale, grad_loss,
                                                                                 col_off, ld, dlogits);
  DRN_CHECK_LAUNCH("pcl_bwd");
  return 0;
}

}  // extern "C"
