// Data-parallel training of the one big parameter (fc6.weight, 205 M values) without an all-reduce: peer-memory plumbing and the
// sharded optimizer step.  SURVEY.md 8(e): the reference wraps the model in DistributedDataParallel
// (detectron2/engine/defaults.py:279-282), i.e. an 822 MB fp32 gradient all-reduce per step followed by the same full update on
// every rank.  Here
//   * the weight-gradient GEMM's epilogue (drn_gemm_bf16_tc_scatter, drn_tc.cu) stores every 128-row tile of THIS rank's gradient
//     straight into the accumulation window of the rank that owns those rows (slot = source rank) over NVLink -- the
//     reduce-scatter's data movement rides the GEMM, tile by tile;
//   * drn_sgd_step_sharded sums the slots of the rows it owns (fixed order -> bit-reproducible), applies torch.optim.SGD's
//     arithmetic to its shard of the fp32 master weights / momentum (1/n of the update traffic per rank) and stores the refreshed
//     bf16 kernel-layout rows into EVERY rank's weight buffer -- the all-gather rides the update.
// Windows are ordinary device allocations shared through CUDA IPC handles (drn_peer_open); ordering between the phases is the
// caller's (two tiny collectives per step, see drn_wsod_pytorch_b200/distributed.py).
#include "common.cuh"
#include <cuda.h>
#include <string.h>

namespace drn {

struct ShardArgs {
  const float* slots;          // [n_src][rows][K] fp32: slot s = rank s's gradient of the rows this rank owns
  __nv_bfloat16* packed[8];    // every rank's bf16 kernel-layout weight [N][K]
  int n_src, n_dst;
};

__device__ __forceinline__ float sgd_math(float w, float g, float* mom, long long i, float lr, float momentum, float wd,
                                          int nesterov, int first) {
  float d = fmaf(wd, w, g);
  if (momentum != 0.f) {
    const float b = first ? d : fmaf(momentum, mom[i], d);
    mom[i] = b;
    d = nesterov ? fmaf(momentum, b, d) : b;
  }
  return fmaf(-lr, d, w);
}

// grid = (rows of the shard, C49 / 64): one CTA owns 64 channels x 49 bins of one row (contiguous in the parameter's (c, ph, pw)
// order), sums the gradient slots, updates master + momentum in place and transposes the new values through shared memory into
// the kernels' bin-major column order for every destination (both sides coalesced).
__global__ void __launch_bounds__(256)
sgd_sharded_kernel(float* __restrict__ w, float* __restrict__ mom, ShardArgs a, long long row0, int rows, long long K, int C49,
                   float inv_n, float lr, float momentum, float wd, int nesterov, int first) {
  __shared__ float tile[64 * 49];
  constexpr int span = 64 * 49;  // values per CTA
  const long long col0 = (long long)blockIdx.y * span;
  const long long lrow = blockIdx.x;                       // row inside the shard
  const long long wbase = (row0 + lrow) * K + col0;        // in the full parameter
  const long long sbase = lrow * K + col0;                 // in a slot / the momentum shard
  const long long slot_stride = (long long)rows * K;
  const int nval = (int)min((long long)span, K - col0);
  for (int i = threadIdx.x; i < nval; i += 256) {
    float g = 0.f;
    for (int s = 0; s < a.n_src; ++s) g += __ldcs(a.slots + (long long)s * slot_stride + sbase + i);  // fixed order
    const float nw = sgd_math(w[wbase + i], g * inv_n, mom, sbase + i, lr, momentum, wd, nesterov, first);
    w[wbase + i] = nw;
    tile[i] = nw;
  }
  __syncthreads();
  const long long prow = (row0 + lrow) * K + (long long)blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 49; i += 256) {
    const int bin = i >> 6, cc = i & 63;
    const __nv_bfloat16 v = __float2bfloat16(tile[cc * 49 + bin]);
    for (int d = 0; d < a.n_dst; ++d) a.packed[d][prow + (long long)bin * C49 + cc] = v;
  }
}

}  // namespace drn

using namespace drn;

extern "C" {

int drn_peer_get_handle(const void* ptr, void* ipc_handle, uint64_t* offset) {
  DRN_CHECK_ARG(ptr && ipc_handle && offset, "peer_get_handle: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err("peer_get_handle: cudaIpcGetMemHandle: %s (the tensor must live in a cudaMalloc allocation)", cudaGetErrorString(e));
  }
  // the handle names the whole allocation: report where `ptr` sits inside it
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  static RangeFn range = nullptr;
  if (!range) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return set_err("peer_get_handle: cuMemGetAddressRange entry point not available");
    range = (RangeFn)fp;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = range(&base, &size, (CUdeviceptr)(uintptr_t)ptr);
  if (r != CUDA_SUCCESS) return set_err("peer_get_handle: cuMemGetAddressRange failed with CUresult %d", (int)r);
  memcpy(ipc_handle, &h, sizeof(h));
  *offset = (uint64_t)((uintptr_t)ptr - (uintptr_t)base);
  return 0;
}

int drn_peer_open(const void* ipc_handle, void** base) {
  DRN_CHECK_ARG(ipc_handle && base, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err("peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  return 0;
}

int drn_peer_close(void* base) {
  cudaError_t e = cudaIpcCloseMemHandle(base);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return set_err("peer_close: cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  }
  return 0;
}

int drn_peer_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int drn_sgd_step_sharded(float* w, float* momentum_shard, const float* slots, int n_src, void* const* packed_bf16, int n_dst,
                         int64_t row0, int64_t rows, int64_t cols, int c49, float lr, float momentum, float weight_decay,
                         int nesterov, int first_step, drn_stream_t stream) {
  DRN_CHECK_ARG(w && slots && packed_bf16, "sgd_step_sharded: null pointer");
  DRN_CHECK_ARG(n_src >= 1 && n_src <= 8 && n_dst >= 1 && n_dst <= 8, "sgd_step_sharded: %d sources, %d destinations", n_src, n_dst);
  DRN_CHECK_ARG(momentum == 0.f || momentum_shard, "sgd_step_sharded: momentum without a buffer");
  DRN_CHECK_ARG(rows > 0 && rows <= 0x7fffffff && cols > 0, "sgd_step_sharded: shape");
  DRN_CHECK_ARG(c49 > 0 && cols == (int64_t)c49 * 49 && c49 % 64 == 0, "sgd_step_sharded: cols=%lld is not 49 x %d (c49 %% 64 == 0)", (long long)cols, c49);
  ShardArgs a{};
  a.slots = slots; a.n_src = n_src; a.n_dst = n_dst;
  for (int d = 0; d < n_dst; ++d) {
    DRN_CHECK_ARG(packed_bf16[d], "sgd_step_sharded: destination %d is null", d);
    a.packed[d] = (__nv_bfloat16*)packed_bf16[d];
  }
  const unsigned gy = (unsigned)(c49 / 64);
  sgd_sharded_kernel<<<dim3((unsigned)rows, gy), 256, 0, (cudaStream_t)stream>>>(w, momentum_shard, a, row0, (int)rows, cols, c49,
                                                                                 1.f / n_src, lr, momentum, weight_decay, nesterov,
                                                                                 first_step);
  DRN_CHECK_LAUNCH("sgd_step_sharded");
  return 0;
}

}  // extern "C"
