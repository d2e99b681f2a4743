// tcgen05 / TMEM / TMA implicit-GEMM for sm_100a: the dense contractions of the hot path
// (ResNet-WS / VGG 3x3 + 1x1 convs, fc6, fc7, concatenated heads) in bf16 with fp32 accumulation.
//
//   out[m][n] = act( (sum_k A[m][k] * W[n][k]) * scale[n] + bias[n] + residual[m][n] )
//
// A is either a row-major [M][K] matrix (linear layers, 1x1 convs: 2D TMA) or an NHWC activation
// tensor read through a 4D TMA box of 8x16 output pixels x 64 channels shifted by the filter tap
// (3x3 convs: implicit im2col, zero padding comes from TMA out-of-bounds fill).  W is [N][K] K-major.
//
// Persistent warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-
// thread tcgen05.mma issuer, warps 2..5 = epilogue (tcgen05.ld -> scale/bias/residual/ReLU -> global).
// smem ring of STAGES x (128x64 A + BNx64 B) bf16 tiles in the 128B-swizzled K-major canonical layout;
// two TMEM accumulator stages (2 x BN fp32 columns) so the epilogue of tile i overlaps the MMAs of i+1.
#include "common.cuh"
#include <cuda.h>

namespace drn {
namespace tc {

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 64;       // one 128-byte swizzle atom of bf16 along K
constexpr int UMMA_K = 16;
constexpr int TILE_W = 16, TILE_H = 8;  // conv mode: 8 x 16 output pixels = 128 rows
constexpr int NUM_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// K-major, 128B-swizzled canonical smem descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (ignored for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, K-major A/B,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
  // problem
  int M;        // GEMM mode: rows.  conv mode: unused
  int N;        // output channels
  int KB;       // number of 64-wide K blocks (= taps * Cin/64)
  int conv;     // 0 = 2D A, 1 = 4D NHWC A with 3x3 taps
  int NB, H, W, Cin, dil;  // conv mode geometry
  int tiles_h, tiles_w;
  int num_m_tiles, num_n_tiles;
  // epilogue
  const float* scale;
  const float* bias;
  const __nv_bfloat16* residual;  // row pitch N
  void* out;
  int out_f32;
  int ldo;
  int relu;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = p.conv ? (p.Cin / BK) : 1;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % p.num_m_tiles, nt = tile / p.num_m_tiles;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_h * p.tiles_w;
          img = mt / per_img;
          const int r = mt - img * per_img;
          h0 = (r / p.tiles_w) * TILE_H;
          w0 = (r % p.tiles_w) * TILE_W;
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if (p.conv) {
            const int tap = kb / cblocks, cb = kb - tap * cblocks;
            const int dh = (tap / 3 - 1) * p.dil, dw = (tap % 3 - 1) * p.dil;
            tma_load_4d(&map_a, &full_bar[stage], sa, cb * BK, w0 + dw, h0 + dh, img);
          } else {
            tma_load_2d(&map_a, &full_bar[stage], sa, kb * BK, mt * BM);
          }
          tma_load_2d(&map_b, &full_bar[stage], sb, kb * BK, nt * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in the >>4 encoded address
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (kb == p.KB - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile % p.num_m_tiles, nt = tile / p.num_m_tiles;
      long long m_global;
      bool row_ok;
      if (p.conv) {
        const int per_img = p.tiles_h * p.tiles_w;
        const int img = mt / per_img;
        const int r = mt - img * per_img;
        const int hh = (r / p.tiles_w) * TILE_H + row / TILE_W;
        const int ww = (r % p.tiles_w) * TILE_W + row % TILE_W;
        row_ok = hh < p.H && ww < p.W;
        m_global = ((long long)img * p.H + hh) * p.W + ww;
      } else {
        m_global = (long long)mt * BM + row;
        row_ok = m_global < p.M;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        const int n0 = nt * BN + c;
        if (row_ok && n0 < p.N) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          const int nvalid = min(32, p.N - n0);  // multiple of 8 (host checks N % 8 == 0)
          if (p.scale) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < nvalid) {
                const float4 s = __ldg(reinterpret_cast<const float4*>(p.scale + n0 + j));
                f[j] *= s.x; f[j + 1] *= s.y; f[j + 2] *= s.z; f[j + 3] *= s.w;
              }
          }
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < nvalid) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
          }
          if (p.residual) {
            const __nv_bfloat16* rp = p.residual + m_global * p.N + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < nvalid) {
                const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + j));
                const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 rf = __bfloat1622float2(r2[q]);
                  f[j + 2 * q] += rf.x; f[j + 2 * q + 1] += rf.y;
                }
              }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.out_f32) {
            float* op = reinterpret_cast<float*>(p.out) + m_global * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < nvalid) *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + m_global * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              if (j < nvalid) {
                uint4 o;
                __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int q = 0; q < 4; ++q) o2[q] = __floats2bfloat162_rn(f[j + 2 * q], f[j + 2 * q + 1]);
                *reinterpret_cast<uint4*>(op + j) = o;
              }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                    const cuuint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err("cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const Params& p, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err("gemm_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_tc_kernel<BN, STAGES><<<grid, NUM_THREADS, smem, st>>>(ma, mb, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_err("gemm_tc launch: %s", cudaGetErrorString(e));
  return 0;
}

static int pick_bn(int m_tiles, int N) {
  // minimise (waves x per-tile cost); per-tile cost ~ BN + fixed overhead (pipeline fill + epilogue tail)
  const int sms = num_sms();
  int best = 64;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N < bn) continue;
    const int tiles = m_tiles * ((N + bn - 1) / bn);
    const int waves = (tiles + sms - 1) / sms;
    const double cost = (double)waves * (bn + 24);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace tc
}  // namespace drn

using namespace drn;

extern "C" int drn_conv_igemm_bf16_tc(const void* in, int N, int H, int W, int Cin, const void* w, int ksize,
                                      int dilation, const float* scale, const float* bias, const void* residual,
                                      int relu, void* out, int out_dtype, int Cout, int ldo, drn_stream_t stream) {
  using namespace drn::tc;
  DRN_CHECK_ARG(in && w && out, "conv_igemm_bf16_tc: null pointer");
  DRN_CHECK_ARG(ksize == 1 || ksize == 3, "conv_igemm_bf16_tc: ksize %d", ksize);
  DRN_CHECK_ARG(Cin % 64 == 0, "conv_igemm_bf16_tc: Cin=%d must be a multiple of 64", Cin);
  DRN_CHECK_ARG(Cout % 8 == 0, "conv_igemm_bf16_tc: Cout=%d must be a multiple of 8", Cout);
  DRN_CHECK_ARG(ldo >= Cout && ldo % 8 == 0, "conv_igemm_bf16_tc: ldo=%d", ldo);
  DRN_CHECK_ARG(((uintptr_t)in % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0),
                "conv_igemm_bf16_tc: operands must be 16-byte aligned");
  const long long Mll = (long long)N * H * W;
  if (Mll == 0) return 0;
  DRN_CHECK_ARG(Mll < (1ll << 31), "conv_igemm_bf16_tc: too many rows");
  Params p{};
  p.N = Cout;
  p.scale = scale; p.bias = bias; p.residual = (const __nv_bfloat16*)residual;
  p.out = out; p.out_f32 = (out_dtype == DRN_F32); p.ldo = ldo; p.relu = relu;
  const int Ktot = ksize * ksize * Cin;
  p.KB = Ktot / BK;
  CUtensorMap ma, mb;
  if (ksize == 3) {
    p.conv = 1; p.NB = N; p.H = H; p.W = W; p.Cin = Cin; p.dil = dilation;
    p.tiles_h = (H + TILE_H - 1) / TILE_H; p.tiles_w = (W + TILE_W - 1) / TILE_W;
    p.num_m_tiles = N * p.tiles_h * p.tiles_w;
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    const cuuint32_t box[4] = {BK, TILE_W, TILE_H, 1};
    if (make_map(&ma, in, 4, dims, strides, box)) return 1;
  } else {
    p.conv = 0; p.M = (int)Mll;
    p.num_m_tiles = (int)((Mll + BM - 1) / BM);
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Mll};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {BK, BM};
    if (make_map(&ma, in, 2, dims, strides, box)) return 1;
  }
  const int bn = pick_bn(p.num_m_tiles, Cout);
  p.num_n_tiles = (Cout + bn - 1) / bn;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    if (make_map(&mb, w, 2, dims, strides, box)) return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (bn == 256) return launch<256, 4>(ma, mb, p, st);
  if (bn == 128) return launch<128, 6>(ma, mb, p, st);
  return launch<64, 8>(ma, mb, p, st);
}
