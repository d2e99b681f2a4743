// Inference tail of the WSL ROI heads on the device: score threshold -> per-class NMS -> top-k.
// Replaces projects/WSL/wsl/modeling/roi_heads/fast_rcnn.py:88-141 (fast_rcnn_inference_single_image) with
// detectron2/layers/nms.py:10-29 (batched_nms -> torchvision nms) and detectron2/structures/boxes.py clip.
//
// Reference semantics restated (fp32 unless noted):
//   rows with a non-finite box coordinate or score are dropped; the background column (last) is dropped;
//   boxes are clipped to [0, w] x [0, h]; candidates = (r, c) with score[r][c] > score_thresh in row-major order;
//   per class: candidates sorted by score descending (stable: ties keep candidate order), greedy NMS with
//     inter = max(0, xx2 - xx1) * max(0, yy2 - yy1);  ovr = inter / ((area_i + area_j) - inter);
//     suppressed iff (double)ovr > (double)nms_thresh            (torchvision cpu/nms_kernel.cpp: the threshold is a double)
//   kept candidates of all classes sorted by score descending (ties: candidate order), first topk returned.
// torchvision's "coordinate trick" (used below 1000 / 5000 candidates on CPU / CUDA) adds class * (max + 1) to the
// coordinates before one class-agnostic NMS; in fp32 that rounds the boxes to the offset's ulp, so this kernel follows
// the exact per-class form (torchvision `_batched_nms_vanilla`, detectron2's own loop for >= 40000 candidates).
//
// One CTA per class (1024 threads): ordered gather of the class's candidates, bitonic sort of 64-bit keys
// ((~score bits) << 32 | row) in shared memory, greedy suppression loop with the boxes resident in shared memory.
// A second kernel ranks the kept candidates of all classes (binary searches in the per-class sorted lists) and
// scatters the first topk into fixed-size outputs -- no data-dependent shapes, so the tail can live in a CUDA graph.
#include "common.cuh"

#include <math.h>

namespace drn {

constexpr int NMS_THREADS = 1024;
constexpr int NMS_MAX_R = 8192;

// row_ok[r] = all scores (K+1 columns) and all box coordinates (4 * nreg) of row r are finite
__global__ void nms_row_valid_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, int R, int K, int nreg,
                                     unsigned char* __restrict__ row_ok) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= R) return;
  bool ok = true;
  for (int i = lane; i < K + 1; i += 32) ok = ok && isfinite(scores[(size_t)warp * (K + 1) + i]);
  for (int i = lane; i < 4 * nreg; i += 32) ok = ok && isfinite(boxes[(size_t)warp * 4 * nreg + i]);
  ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) row_ok[warp] = ok ? 1 : 0;
}

// grid = K, block = 1024, dynamic smem = P * 8 (keys) + P * 16 (boxes) + P (flags) + 4 * 33, P = next pow2 >= R
__global__ void __launch_bounds__(NMS_THREADS)
nms_class_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const unsigned char* __restrict__ row_ok,
                 int R, int K, int nreg, int P, float img_h, float img_w, float score_thresh, double nms_thresh, int cap,
                 unsigned long long* __restrict__ kept_keys, int* __restrict__ kept_count) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(nms_smem);
  float4* sbox = reinterpret_cast<float4*>(nms_smem + (size_t)P * 8);
  unsigned char* dead = nms_smem + (size_t)P * 24;
  int* wsum = reinterpret_cast<int*>(nms_smem + (size_t)P * 25 + 16 - ((size_t)P * 25) % 16);
  __shared__ int s_base;
  const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();

  // ---- ordered gather: candidates of class c in row order (the reference's nonzero() order within a class)
  for (int r0 = 0; r0 < R; r0 += NMS_THREADS) {
    const int r = r0 + tid;
    float sc = 0.f;
    bool take = false;
    if (r < R && row_ok[r]) {
      sc = scores[(size_t)r * (K + 1) + c];
      take = sc > score_thresh;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, take);
    const int in_warp = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) wsum[wid] = __popc(ball);
    __syncthreads();
    if (wid == 0) {
      int v = wsum[lane], x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      wsum[lane] = x - v;          // exclusive prefix per warp
      if (lane == 31) wsum[32] = x;  // total of this chunk
    }
    __syncthreads();
    if (take) {
      const int pos = s_base + wsum[wid] + in_warp;
      // ascending key order == score descending, then row ascending (scores here are > thresh >= 0: bit order == value order)
      keys[pos] = ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(sc)) << 32) | (unsigned)r;
    }
    __syncthreads();
    if (tid == 0) s_base += wsum[32];
    __syncthreads();
  }
  const int n = s_base;
  if (n == 0) {
    if (tid == 0) kept_count[c] = 0;
    return;
  }
  // ---- bitonic sort of the first P2 = pow2 >= n keys (padding = max key)
  int P2 = 1;
  while (P2 < n) P2 <<= 1;
  for (int i = n + tid; i < P2; i += NMS_THREADS) keys[i] = ~0ull;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < P2; i += NMS_THREADS) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // ---- boxes in sorted order, clipped to the image (Boxes.clip: x in [0, w], y in [0, h])
  for (int i = tid; i < n; i += NMS_THREADS) {
    const int r = (int)(keys[i] & 0xFFFFFFFFull);
    const float* b = boxes + (size_t)r * 4 * nreg + (nreg == 1 ? 0 : 4 * c);
    float4 v;
    v.x = fminf(fmaxf(b[0], 0.f), img_w);
    v.y = fminf(fmaxf(b[1], 0.f), img_h);
    v.z = fminf(fmaxf(b[2], 0.f), img_w);
    v.w = fminf(fmaxf(b[3], 0.f), img_h);
    sbox[i] = v;
    dead[i] = 0;
  }
  __syncthreads();
  // ---- greedy suppression in score order: every thread tests box i against its share of the later boxes.
  // Only the first `cap` kept boxes of a class can reach the image's top-`cap` (they precede every later box of the
  // class in the final order), so the scan stops there: the cost is bounded by cap barriers, not by R.
  int stop = n, nkept = 0;
  for (int i = 0; i < n; ++i) {
    if (dead[i]) continue;  // uniform: written before the last barrier
    if (++nkept >= cap) { stop = i + 1; break; }  // uniform
    const float4 bi = sbox[i];
    const float ai = __fmul_rn(__fsub_rn(bi.z, bi.x), __fsub_rn(bi.w, bi.y));
    for (int j = i + 1 + tid; j < n; j += NMS_THREADS) {
      if (dead[j]) continue;
      const float4 bj = sbox[j];
      const float aj = __fmul_rn(__fsub_rn(bj.z, bj.x), __fsub_rn(bj.w, bj.y));
      const float w = fmaxf(0.f, __fsub_rn(fminf(bi.z, bj.z), fmaxf(bi.x, bj.x)));
      const float h = fmaxf(0.f, __fsub_rn(fminf(bi.w, bj.w), fmaxf(bi.y, bj.y)));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(ai, aj), inter));
      if ((double)ovr > nms_thresh) dead[j] = 1;
    }
    __syncthreads();
  }
  // ---- kept keys of this class, still in sorted order (ordered compaction)
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < stop; i0 += NMS_THREADS) {
    const int i = i0 + tid;
    const bool keep = i < stop && !dead[i];
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    const int in_warp = __popc(ball & ((1u << lane) - 1u));
    if (lane == 0) wsum[wid] = __popc(ball);
    __syncthreads();
    if (wid == 0) {
      int v = wsum[lane], x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      wsum[lane] = x - v;
      if (lane == 31) wsum[32] = x;
    }
    __syncthreads();
    if (keep) kept_keys[(size_t)c * R + s_base + wsum[wid] + in_warp] = keys[i];
    __syncthreads();
    if (tid == 0) s_base += wsum[32];
    __syncthreads();
  }
  if (tid == 0) kept_count[c] = s_base;
}

// One thread per (class, kept slot): global rank in (score desc, row asc, class asc) order by binary searches in the
// other classes' sorted kept lists; rank < cap -> write the detection.  grid = (ceil(R / 256), K).
__global__ void nms_rank_scatter_kernel(const unsigned long long* __restrict__ kept_keys, const int* __restrict__ kept_count,
                                        const float* __restrict__ boxes, int R, int K, int nreg, float img_h, float img_w,
                                        int cap, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                        long long* __restrict__ out_classes, long long* __restrict__ out_rows,
                                        int* __restrict__ num_out) {
  const int c = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && m == 0) {
    int tot = 0;
    for (int k = 0; k < K; ++k) tot += kept_count[k];
    *num_out = tot < cap ? tot : cap;
  }
  if (m >= kept_count[c]) return;
  const unsigned long long key = kept_keys[(size_t)c * R + m];
  int rank = m;
  for (int k = 0; k < K; ++k) {
    if (k == c) continue;
    const unsigned long long* lst = kept_keys + (size_t)k * R;
    // entries of class k that come first: key' < key, or key' == key (same score, same row) and k < c
    int lo = 0, hi = kept_count[k];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const unsigned long long v = lst[mid];
      if (v < key || (v == key && k < c)) lo = mid + 1; else hi = mid;
    }
    rank += lo;
  }
  if (rank >= cap) return;
  const int r = (int)(key & 0xFFFFFFFFull);
  const float* b = boxes + (size_t)r * 4 * nreg + (nreg == 1 ? 0 : 4 * c);
  out_boxes[4 * (size_t)rank + 0] = fminf(fmaxf(b[0], 0.f), img_w);
  out_boxes[4 * (size_t)rank + 1] = fminf(fmaxf(b[1], 0.f), img_h);
  out_boxes[4 * (size_t)rank + 2] = fminf(fmaxf(b[2], 0.f), img_w);
  out_boxes[4 * (size_t)rank + 3] = fminf(fmaxf(b[3], 0.f), img_h);
  out_scores[rank] = __uint_as_float(0xFFFFFFFFu - (unsigned)(key >> 32));
  out_classes[rank] = c;
  out_rows[rank] = r;
}

}  // namespace drn

using namespace drn;

static int nms_pow2(int R) {
  int P = 1;
  while (P < R) P <<= 1;
  return P;
}

extern "C" {

size_t drn_detections_workspace_bytes(int R, int K) {
  if (R <= 0 || K <= 0) return 0;
  // kept keys [K][R] u64 + kept counts [K] i32 (padded) + row validity [R]
  return (size_t)K * R * 8 + (((size_t)K * 4 + 15) / 16) * 16 + (((size_t)R + 15) / 16) * 16;
}

int drn_detections_fwd(const float* all_scores, const float* all_boxes, int R, int K, int nreg, float img_h, float img_w,
                       float score_thresh, double nms_thresh, int cap, float* out_boxes, float* out_scores,
                       int64_t* out_classes, int64_t* out_rows, int32_t* num_out, void* workspace, size_t workspace_bytes,
                       drn_stream_t stream) {
  DRN_CHECK_ARG(R >= 0 && K > 0, "detections: R=%d K=%d", R, K);
  DRN_CHECK_ARG(num_out, "detections: null num_out");
  cudaStream_t st = (cudaStream_t)stream;
  if (R == 0 || cap == 0) {
    cudaError_t e = cudaMemsetAsync(num_out, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return set_err("detections: memset: %s", cudaGetErrorString(e));
    return 0;
  }
  DRN_CHECK_ARG(all_scores && all_boxes && out_boxes && out_scores && out_classes && out_rows && workspace, "detections: null pointer");
  DRN_CHECK_ARG(nreg == 1 || nreg == K, "detections: nreg=%d must be 1 or K=%d", nreg, K);
  DRN_CHECK_ARG(R <= NMS_MAX_R, "detections: R=%d exceeds the %d rows the in-shared-memory sort handles", R, NMS_MAX_R);
  DRN_CHECK_ARG(cap > 0, "detections: cap=%d", cap);
  DRN_CHECK_ARG(workspace_bytes >= drn_detections_workspace_bytes(R, K), "detections: workspace too small");
  DRN_CHECK_ARG((uintptr_t)workspace % 16 == 0, "detections: workspace must be 16-byte aligned");
  unsigned long long* kept_keys = (unsigned long long*)workspace;
  int* kept_count = (int*)((char*)workspace + (size_t)K * R * 8);
  unsigned char* row_ok = (unsigned char*)kept_count + (((size_t)K * 4 + 15) / 16) * 16;
  const int P = nms_pow2(R) < 2 ? 2 : nms_pow2(R);
  const size_t smem = (size_t)P * 25 + 16 + 33 * sizeof(int) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(nms_class_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err("detections: cudaFuncSetAttribute(%zu B): %s", smem, cudaGetErrorString(e));
    configured = smem;
  }
  nms_row_valid_kernel<<<cdiv(R * 32, 256), 256, 0, st>>>(all_scores, all_boxes, R, K, nreg, row_ok);
  DRN_CHECK_LAUNCH("detections row validity");
  nms_class_kernel<<<K, NMS_THREADS, smem, st>>>(all_scores, all_boxes, row_ok, R, K, nreg, P, img_h, img_w, score_thresh,
                                                 nms_thresh, cap, kept_keys, kept_count);
  DRN_CHECK_LAUNCH("detections per-class nms");
  nms_rank_scatter_kernel<<<dim3(cdiv(R, 256), K), 256, 0, st>>>(kept_keys, kept_count, all_boxes, R, K, nreg, img_h, img_w, cap,
                                                                out_boxes, out_scores, (long long*)out_classes, (long long*)out_rows, num_out);
  DRN_CHECK_LAUNCH("detections rank/scatter");
  return 0;
}

}  // extern "C"
