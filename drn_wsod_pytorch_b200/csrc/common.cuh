// Shared helpers for libdrn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/drn_b200.h"

namespace drn {

// thread-local error message behind drn_last_error()
char* err_buf();
int set_err(const char* fmt, ...);

#define DRN_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) return drn::set_err(__VA_ARGS__); \
  } while (0)

#define DRN_CHECK_LAUNCH(name)                                                        \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) return drn::set_err("%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions for blockDim.x <= 1024 (multiple of 32). `sh` holds >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.f;
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) sh[0] = r;
  }
  __syncthreads();
  return sh[0];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? sh[threadIdx.x] : -INFINITY;
  if (wid == 0) {
    r = warp_max(r);
    if (lane == 0) sh[0] = r;
  }
  __syncthreads();
  return sh[0];
}

}  // namespace drn
