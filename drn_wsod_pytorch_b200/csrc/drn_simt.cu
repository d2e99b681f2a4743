// fp32 / bandwidth kernels of the DRN-WSOD hot path: fused normalise+first conv, SIMT implicit-GEMM
// conv/linear (exact-fp32 mode), 2x2 max-pool, ROIPool (+objectness scaling), dtype casts.
// sm_100a only.  See include/drn_b200.h for the contract of every entry point.
#include "common.cuh"
#include <float.h>

namespace drn {

static thread_local char g_err[512];
char* err_buf() { return g_err; }
int set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

// ------------------------------------------------------------------------------------------
// a1+a2: (img - mean)/std fused into the 3x3, Cin=3 first convolution (stride 1 or 2, pad 1).
// One thread = NPX adjacent output pixels x 16 output channels: every filter value fetched from smem
// (one LDS.128 per 4 channels) feeds NPX FMAs, so the inner loop is FMA-bound instead of LDS-bound
// (the first version, 1 pixel per thread, spent 4 LDS cycles per FMA cycle: 117 us at 600x1000).
// Adjacent threads cover the 64-channel row of a pixel group so NHWC stores are contiguous.
// ------------------------------------------------------------------------------------------
constexpr int C3_NPX = 4;

template <typename OutT, int STRIDE>
__global__ void __launch_bounds__(128)
conv3x3_c3_kernel(const float* __restrict__ img, int H, int W, float m0, float m1, float m2, float s0, float s1, float s2,
                  const float* __restrict__ wp, const float* __restrict__ scale, const float* __restrict__ bias,
                  int Cout, int relu, int Ho, int Wo, OutT* __restrict__ out) {
  extern __shared__ float sw[];  // [27][Cout]
  for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) sw[i] = wp[i];
  __syncthreads();
  constexpr int NCOL = (C3_NPX - 1) * STRIDE + 3;  // input columns under NPX adjacent outputs
  const int groups = Cout >> 4;
  const int wq = (Wo + C3_NPX - 1) / C3_NPX;       // pixel groups per output row
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pg = gid / groups;
  const int cg = (int)(gid % groups);
  if (pg >= (long long)Ho * wq) return;
  const int oh = (int)(pg / wq), ow0 = (int)(pg % wq) * C3_NPX;
  const float mean[3] = {m0, m1, m2};
  const float stdv[3] = {s0, s1, s2};
  float acc[C3_NPX][16];
#pragma unroll
  for (int p = 0; p < C3_NPX; ++p)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
  const float* wrow = sw + cg * 16;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int ih = oh * STRIDE - 1 + kh;
    const bool rowok = ih >= 0 && ih < H;  // rows/cols outside the valid H x W image are ImageList zero padding
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x[NCOL];
#pragma unroll
      for (int i = 0; i < NCOL; ++i) {
        const int iw = ow0 * STRIDE - 1 + i;
        float v = 0.f;
        if (rowok && iw >= 0 && iw < W) v = __fdiv_rn(__fsub_rn(__ldg(img + ((long long)c * H + ih) * W + iw), mean[c]), stdv[c]);
        x[i] = v;
      }
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const float4* w4 = reinterpret_cast<const float4*>(wrow + ((kh * 3 + kw) * 3 + c) * Cout);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wv = w4[q];
#pragma unroll
          for (int p = 0; p < C3_NPX; ++p) {
            const float xv = x[p * STRIDE + kw];
            acc[p][q * 4 + 0] = fmaf(xv, wv.x, acc[p][q * 4 + 0]);
            acc[p][q * 4 + 1] = fmaf(xv, wv.y, acc[p][q * 4 + 1]);
            acc[p][q * 4 + 2] = fmaf(xv, wv.z, acc[p][q * 4 + 2]);
            acc[p][q * 4 + 3] = fmaf(xv, wv.w, acc[p][q * 4 + 3]);
          }
        }
      }
    }
  }
  float sc[16], bi[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    sc[j] = scale ? __ldg(scale + cg * 16 + j) : 1.f;
    bi[j] = __ldg(bias + cg * 16 + j);
  }
#pragma unroll
  for (int p = 0; p < C3_NPX; ++p) {
    const int ow = ow0 + p;
    if (ow >= Wo) break;
    OutT* o = out + ((long long)oh * Wo + ow) * Cout + cg * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float t = scale ? acc[p][j] * sc[j] : acc[p][j];  // same rounding sequence as the reference: conv, *scale, +bias
      t += bi[j];
      v[j] = relu ? fmaxf(t, 0.f) : t;
    }
    if constexpr (sizeof(OutT) == 4) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
      uint4 pk[2];
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
      for (int j = 0; j < 8; ++j) h2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      *reinterpret_cast<uint4*>(o) = pk[0];
      *reinterpret_cast<uint4*>(o + 8) = pk[1];
    }
  }
}

// Tiled variant for Cout == 64 (every stem on the path): one CTA = th x C3_TW output pixels (th even, chosen by the host), 256
// threads = 16 pixel groups x 2 rows x 8 channel groups, one thread = 4 adjacent pixels x 8 channels.  The input patch under the
// tile is normalised ONCE into shared memory ((img - mean) / std with IEEE division, zero outside the valid image: the same
// values the kernel above computes per thread).  The FMA order per output (kh, c, kw) is unchanged: results are bit-identical
// to the kernel above.
// History (ncu, 600x1000 image): v1 (16 channels per thread, 128 threads, 64 x 4 tile) 59 us -- patch staged one dependent
// global load at a time, 168 registers -> 3 CTAs / SM, 600 CTAs = 1.35 waves; v2 (batched patch loads, <= 128 registers, tile
// rows chosen so that the CTAs spread evenly: 64 x 6) 35 us -- 11 warps / SM, IPC 0.44 per scheduler, stalled on the shared-
// memory loads of the FMA loop (short scoreboard 2.2 per issue; 47 % of the shared wavefronts were bank conflicts).  v3 (this):
// half the accumulators per thread -> twice the warps (24 / SM), filter stored so that a quarter-warp's 128-bit loads are
// contiguous, stride-2 patches stored as even | odd columns so that the pixel groups' 128-bit loads are contiguous.
constexpr int C3_TW = 64;
constexpr int C3_TH_MAX = 12;
constexpr int C3_TILE_THREADS = 256;

template <int STRIDE> struct C3Patch;
template <> struct C3Patch<2> {
  static constexpr int PC = (C3_TW - 1) * 2 + 3;  // 129 input columns under 64 outputs
  static constexpr int NE = 68;                   // even columns 0, 2, .., 128 (65, padded to a 16-byte multiple)
  static constexpr int PCP = NE + 64;             // then odd columns 1, 3, .., 127
  static __device__ __forceinline__ int col_of_slot(int slot) { return slot < NE ? 2 * slot : 2 * (slot - NE) + 1; }
  // input columns 2 p0 .. 2 p0 + 8 of the four pixels p0 .. p0 + 3 (p0 % 4 == 0): x[2 j + kw]
  static __device__ __forceinline__ void load(const float* row, int p0, float (&x)[9]) {
    const float4 e = *reinterpret_cast<const float4*>(row + p0);
    const float e4 = row[p0 + 4];
    const float4 o = *reinterpret_cast<const float4*>(row + NE + p0);
    x[0] = e.x; x[1] = o.x; x[2] = e.y; x[3] = o.y; x[4] = e.z; x[5] = o.z; x[6] = e.w; x[7] = o.w; x[8] = e4;
  }
};
template <> struct C3Patch<1> {
  static constexpr int PC = (C3_TW - 1) + 3;      // 66
  static constexpr int PCP = 68;
  static __device__ __forceinline__ int col_of_slot(int slot) { return slot; }
  static __device__ __forceinline__ void load(const float* row, int p0, float (&x)[6]) {
    const float4 a = *reinterpret_cast<const float4*>(row + p0);
    const float2 b = *reinterpret_cast<const float2*>(row + p0 + 4);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y;
  }
};

template <typename OutT, int STRIDE>
__global__ void __launch_bounds__(C3_TILE_THREADS, 3)
conv3x3_c3_tile_kernel(const float* __restrict__ img, int H, int W, float m0, float m1, float m2, float s0, float s1, float s2,
                       const float* __restrict__ wp, const float* __restrict__ scale, const float* __restrict__ bias,
                       int relu, int Ho, int Wo, int th, OutT* __restrict__ out) {
  typedef C3Patch<STRIDE> P;
  constexpr int Cout = 64, PC = P::PC, PCP = P::PCP;
  constexpr int NCOL = (C3_NPX - 1) * STRIDE + 3;
  extern __shared__ __align__(16) float c3_smem[];
  float* sw = c3_smem;             // [27 taps][2 halves][8 channel groups][4]: channel = 8 cg + 4 half + k
  float* sx = c3_smem + 27 * Cout; // [3][PR][PCP] normalised input patch
  const int PR = (th - 1) * STRIDE + 3;
  for (int i = threadIdx.x; i < 27 * Cout; i += C3_TILE_THREADS) {
    const int ch = i & 63;
    sw[(i & ~63) + ((ch >> 2) & 1) * 32 + (ch >> 3) * 4 + (ch & 3)] = __ldg(wp + i);
  }
  const int oh_t = blockIdx.y * th, ow_t = blockIdx.x * C3_TW;
  const int ih0 = oh_t * STRIDE - 1, iw0 = ow_t * STRIDE - 1;
  const float mean[3] = {m0, m1, m2};
  const float stdv[3] = {s0, s1, s2};
  const int plane = PR * PCP;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* src = img + (long long)c * H * W;
    float* dst = sx + c * plane;
    constexpr int U = 8;  // loads in flight per thread
#pragma unroll 1
    for (int base = threadIdx.x; base < plane; base += C3_TILE_THREADS * U) {
      float raw[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * C3_TILE_THREADS;
        const int pr = i / PCP, pc = P::col_of_slot(i - pr * PCP);
        const int ih = ih0 + pr, iw = iw0 + pc;
        ok[u] = i < plane && pc < PC && ih >= 0 && ih < H && iw >= 0 && iw < W;  // outside the valid H x W image: ImageList zero padding
        raw[u] = ok[u] ? __ldg(src + (long long)ih * W + iw) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = base + u * C3_TILE_THREADS;
        if (i < plane) dst[i] = ok[u] ? __fdiv_rn(__fsub_rn(raw[u], mean[c]), stdv[c]) : 0.f;
      }
    }
  }
  __syncthreads();
  const int cg = threadIdx.x & 7, gx = (threadIdx.x >> 3) & 15, gy = threadIdx.x >> 7;  // 8 channels; pixel group: 16 across x 2 rows
  const float* wrow = sw + cg * 4;
#pragma unroll 1
  for (int pass = 0; pass < th / 2; ++pass) {
    const int orow = pass * 2 + gy;                       // output row inside the tile
    const int oh = oh_t + orow, ow0 = ow_t + gx * C3_NPX;
    if (oh >= Ho || ow0 >= Wo) continue;
    float acc[C3_NPX][8];
#pragma unroll
    for (int p = 0; p < C3_NPX; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float x[NCOL];
        P::load(sx + c * plane + (orow * STRIDE + kh) * PCP, gx * C3_NPX, x);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* wt = wrow + ((kh * 3 + kw) * 3 + c) * Cout;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + q * 32);
#pragma unroll
            for (int p = 0; p < C3_NPX; ++p) {
              const float xv = x[p * STRIDE + kw];
              acc[p][q * 4 + 0] = fmaf(xv, wv.x, acc[p][q * 4 + 0]);
              acc[p][q * 4 + 1] = fmaf(xv, wv.y, acc[p][q * 4 + 1]);
              acc[p][q * 4 + 2] = fmaf(xv, wv.z, acc[p][q * 4 + 2]);
              acc[p][q * 4 + 3] = fmaf(xv, wv.w, acc[p][q * 4 + 3]);
            }
          }
        }
      }
    }
    float sc[8], bi[8];
#pragma unroll
    for (int j = 0; j < 8; j += 4) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8 + j));
      bi[j] = b4.x; bi[j + 1] = b4.y; bi[j + 2] = b4.z; bi[j + 3] = b4.w;
      if (scale) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(scale + cg * 8 + j));
        sc[j] = s4.x; sc[j + 1] = s4.y; sc[j + 2] = s4.z; sc[j + 3] = s4.w;
      } else {
        sc[j] = sc[j + 1] = sc[j + 2] = sc[j + 3] = 1.f;
      }
    }
#pragma unroll
    for (int p = 0; p < C3_NPX; ++p) {
      const int ow = ow0 + p;
      if (ow >= Wo) break;
      OutT* o = out + ((long long)oh * Wo + ow) * Cout + cg * 8;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = scale ? acc[p][j] * sc[j] : acc[p][j];  // same rounding sequence as the reference: conv, *scale, +bias
        t += bi[j];
        v[j] = relu ? fmaxf(t, 0.f) : t;
      }
      if constexpr (sizeof(OutT) == 4) {
#pragma unroll
        for (int j = 0; j < 8; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
        uint4 pk;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(o) = pk;
      }
    }
  }
}

// Rows per tile of the kernel above: the launch is FMA-bound, so its time follows the SM with the most CTAs:
// ceil(#CTAs / #SMs) x (th / 2 compute passes + ~half a pass of patch staging); smallest wins, ties to the smaller tile.
static int c3_pick_th(int Ho, int Wo, int num_sms) {
  int best = 2;
  double best_cost = 1e30;
  for (int th = 2; th <= C3_TH_MAX; th += 2) {
    const long ctas = (long)((Wo + C3_TW - 1) / C3_TW) * ((Ho + th - 1) / th);
    const double cost = (double)((ctas + num_sms - 1) / num_sms) * (th / 2 + 0.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = th; }
  }
  return best;
}

// ------------------------------------------------------------------------------------------
// SIMT fp32 implicit GEMM: out[m][n] = act(sum_k A[m][k] * Wt[k][n] * scale[n] + bias[n] + res[m][n])
// m = output pixel (NHWC, stride 1, "same" padding), k = (kh, kw, cin), n = cout.
// Tile 128(m) x 64(n) x 16(k), 128 threads, 8x8 register tile, register-staged double buffering.
// ------------------------------------------------------------------------------------------
constexpr int CBM = 128, CBN = 64, CBK = 16, CPAD = 4;

__global__ void __launch_bounds__(128)
conv_igemm_f32_kernel(const float* __restrict__ in, int N, int H, int W, int Cin,
                      const float* __restrict__ wt, int ksize, int dil,
                      const float* __restrict__ scale, const float* __restrict__ bias,
                      const float* __restrict__ residual, int relu, float* __restrict__ out, int Cout,
                      int ldo) {
  __shared__ __align__(16) float As[2][CBK][CBM + CPAD];
  __shared__ __align__(16) float Bs[2][CBK][CBN];
  const int tid = threadIdx.x;
  const long long M = (long long)N * H * W;
  const long long m0 = (long long)blockIdx.x * CBM;
  const int n0 = blockIdx.y * CBN;
  const int KT = (ksize * ksize * Cin) / CBK;
  const int half = ksize >> 1;

  // A-load assignment: this thread loads float4 #kq of the 16-wide k slab of 4 pixels.
  const int kq = tid & 3;
  int ph[4], pw[4];
  long long pbase[4];
  bool pval[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + (tid >> 2) + 32 * i;
    pval[i] = m < M;
    const long long mm = pval[i] ? m : 0;
    const int w_ = (int)(mm % W);
    const int h_ = (int)((mm / W) % H);
    ph[i] = h_;
    pw[i] = w_;
    pbase[i] = mm * Cin;  // offset of (n,h,w,0)
  }
  // B-load assignment: 2 float4 per thread.
  const int bk0 = tid >> 4, bn4 = (tid & 15) * 4;

  float4 ra[4], rb[2];
  auto load_tile = [&](int kt) {
    const int kk = kt * CBK;
    const int tap = kk / Cin;
    const int c0 = kk - tap * Cin + kq * 4;
    const int dh = (tap / ksize - half) * dil, dw = (tap % ksize - half) * dil;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ih = ph[i] + dh, iw = pw[i] + dw;
      const bool ok = pval[i] && ih >= 0 && ih < H && iw >= 0 && iw < W;
      ra[i] = ok ? __ldg(reinterpret_cast<const float4*>(in + pbase[i] + ((long long)dh * W + dw) * Cin + c0))
                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      rb[i] = __ldg(reinterpret_cast<const float4*>(wt + (long long)(kk + bk0 + 8 * i) * Cout + n0 + bn4));
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ml = (tid >> 2) + 32 * i;
      As[buf][kq * 4 + 0][ml] = ra[i].x;
      As[buf][kq * 4 + 1][ml] = ra[i].y;
      As[buf][kq * 4 + 2][ml] = ra[i].z;
      As[buf][kq * 4 + 3][ml] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&Bs[buf][bk0 + 8 * i][bn4]) = rb[i];
  };

  const int ty = tid >> 3, tx = tid & 7;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < CBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][32 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tile(buf ^ 1);
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 32 + tx * 4;
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (scale) {
        const float4 s = __ldg(reinterpret_cast<const float4*>(scale + n));
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
      }
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (residual) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(residual + m * Cout + n));
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      if (relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      *reinterpret_cast<float4*>(out + m * ldo + n) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------
// MaxPool2d(2, stride s, pad 0) on NHWC; one thread = one output pixel x 16-byte channel vector.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ uint4 max8bf(uint4 a, uint4 b) {
  uint4 r;
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  return r;
}

template <bool BF16>
__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const void* __restrict__ in_, int N, int H, int W, int C, int s, int Ho, int Wo,
                  void* __restrict__ out_) {
  const int vec = BF16 ? 8 : 4;
  const int CV = C / vec;
  const long long total = (long long)N * Ho * Wo * CV;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= total) return;
  const int cv = (int)(gid % CV);
  long long p = gid / CV;
  const int ow = (int)(p % Wo); p /= Wo;
  const int oh = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const long long base = (((long long)n * H + oh * s) * W + ow * s) * CV + cv;
  const long long rs = (long long)W * CV;
  if constexpr (BF16) {
    const uint4* in = reinterpret_cast<const uint4*>(in_);
    uint4 v = max8bf(max8bf(__ldg(in + base), __ldg(in + base + CV)),
                     max8bf(__ldg(in + base + rs), __ldg(in + base + rs + CV)));
    reinterpret_cast<uint4*>(out_)[gid] = v;
  } else {
    const float4* in = reinterpret_cast<const float4*>(in_);
    float4 v = max4(max4(__ldg(in + base), __ldg(in + base + CV)),
                    max4(__ldg(in + base + rs), __ldg(in + base + rs + CV)));
    reinterpret_cast<float4*>(out_)[gid] = v;
  }
}

// ------------------------------------------------------------------------------------------
// a7+a8: ROIPool 7x7 (torchvision semantics) x (objectness+1).
// grid = (R, 7): one CTA per (proposal, bin row); threads stride over 16-byte channel vectors so
// every load/store is a coalesced 512 B per warp; the conv5 map stays L2-resident (<= 134 MB bf16
// only at cfg 5; 19-75 MB otherwise) and the R x 49 x C output is the only HBM stream.
// ------------------------------------------------------------------------------------------
struct RoiBins {
  int hs, he;
  int ws[7], we[7];
};

__device__ __forceinline__ void roi_geometry(const float* __restrict__ box, float scale, int h, int w,
                                             int ph, RoiBins& g) {
  // torchvision roi_pool: C round() (half away from zero) on scaled coords, +1 extents, floor/ceil
  // bin edges, clamp to the map.  All arithmetic fp32, single roundings (no contraction).
  const int sw = (int)roundf(__fmul_rn(box[0], scale));
  const int sh = (int)roundf(__fmul_rn(box[1], scale));
  const int ew = (int)roundf(__fmul_rn(box[2], scale));
  const int eh = (int)roundf(__fmul_rn(box[3], scale));
  const int rw = max(ew - sw + 1, 1), rh = max(eh - sh + 1, 1);
  const float bh = __fdiv_rn((float)rh, 7.f), bw = __fdiv_rn((float)rw, 7.f);
  g.hs = min(max((int)floorf(__fmul_rn((float)ph, bh)) + sh, 0), h);
  g.he = min(max((int)ceilf(__fmul_rn((float)(ph + 1), bh)) + sh, 0), h);
#pragma unroll
  for (int p = 0; p < 7; ++p) {
    g.ws[p] = min(max((int)floorf(__fmul_rn((float)p, bw)) + sw, 0), w);
    g.we[p] = min(max((int)ceilf(__fmul_rn((float)(p + 1), bw)) + sw, 0), w);
  }
}

__global__ void __launch_bounds__(128)
roipool_f32_kernel(const float* __restrict__ feat, int h, int w, int C, const float* __restrict__ boxes,
                   const float* __restrict__ obj, float scale, float* __restrict__ out) {
  const int r = blockIdx.x, ph = blockIdx.y;
  RoiBins g;
  roi_geometry(boxes + 4 * (long long)r, scale, h, w, ph, g);
  const float mul = obj ? __fadd_rn(__ldg(obj + r), 1.f) : 1.f;
  const int CV = C >> 2;
  const float4* f4 = reinterpret_cast<const float4*>(feat);
  float4* o4 = reinterpret_cast<float4*>(out + ((long long)r * 49 + ph * 7) * C);
  for (int cv = threadIdx.x; cv < CV; cv += blockDim.x) {
#pragma unroll 1
    for (int pw = 0; pw < 7; ++pw) {
      const int ws = g.ws[pw], we = g.we[pw];
      const bool empty = (g.he <= g.hs) || (we <= ws);
      float4 m = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                       : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
      for (int y = g.hs; y < g.he; ++y) {
        const float4* row = f4 + ((long long)y * w) * CV + cv;
#pragma unroll 4
        for (int x = ws; x < we; ++x) {
          const float4 v = __ldg(row + (long long)x * CV);
          if (v.x > m.x) m.x = v.x;
          if (v.y > m.y) m.y = v.y;
          if (v.z > m.z) m.z = v.z;
          if (v.w > m.w) m.w = v.w;
        }
      }
      m.x = __fmul_rn(m.x, mul); m.y = __fmul_rn(m.y, mul);
      m.z = __fmul_rn(m.z, mul); m.w = __fmul_rn(m.w, mul);
      __stcs(o4 + (long long)pw * CV + cv, m);
    }
  }
}

__global__ void __launch_bounds__(256)
roipool_bf16_kernel(const uint4* __restrict__ f8, int h, int w, int C, const float* __restrict__ boxes,
                    const float* __restrict__ obj, float scale, uint4* __restrict__ out) {
  const int r = blockIdx.x, ph = blockIdx.y;
  RoiBins g;
  roi_geometry(boxes + 4 * (long long)r, scale, h, w, ph, g);
  const float mul = obj ? __fadd_rn(__ldg(obj + r), 1.f) : 1.f;
  const int CV = C >> 3;
  uint4* o8 = out + ((long long)r * 49 + ph * 7) * CV;
  const __nv_bfloat16 lowest = __ushort_as_bfloat16((unsigned short)0xFF7F);  // most negative finite bf16
  for (int cv = threadIdx.x; cv < CV; cv += blockDim.x) {
#pragma unroll 1
    for (int pw = 0; pw < 7; ++pw) {
      const int ws = g.ws[pw], we = g.we[pw];
      const bool empty = (g.he <= g.hs) || (we <= ws);
      __nv_bfloat162 m[4];
      const __nv_bfloat16 init = empty ? __float2bfloat16(0.f) : lowest;
#pragma unroll
      for (int i = 0; i < 4; ++i) m[i] = __halves2bfloat162(init, init);
      for (int y = g.hs; y < g.he; ++y) {
        const uint4* row = f8 + ((long long)y * w) * CV + cv;
#pragma unroll 4
        for (int x = ws; x < we; ++x) {
          const uint4 v = __ldg(row + (long long)x * CV);
          const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
          for (int i = 0; i < 4; ++i) m[i] = __hmax2(m[i], pv[i]);
        }
      }
      uint4 o;
      __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(m[i]);
        po[i] = __floats2bfloat162_rn(__fmul_rn(f.x, mul), __fmul_rn(f.y, mul));
      }
      __stcs(o8 + (long long)pw * CV + cv, o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// ROIPool v2: range-max (sparse) tables + per-bin gather.
//
// The direct kernels above re-read every feature cell under every ROI (measured: 19 GB of L2->SM
// traffic for R=4000 on the 74x124x2048 map, L2-bound at 1.07 ms).  max() is idempotent, so a window
// [ws, we) equals max(T_j[ws], T_j[we - 2^j]) with j = floor(log2(we - ws)) and T_j[x] = max F[x .. x+2^j)
// -- the classic sparse table, built once per image: x windows 2,4,8,16 (wider bins take
// ceil(width/16) lookups) crossed with y windows 1,2.  A bin of hh x width cells then costs
// ceil(hh/2) x 2 lookups instead of hh x width loads; results are bit-identical (max has no rounding).
// Kernel 1 builds the 9 tables (one CTA per feature row x 128-byte channel slice, rows y and y+1 staged
// in smem); kernel 2 is one CTA per (ROI, 512-byte channel chunk): 8 warps x 49 bins, 16 B per lane.
// CTAs are rasterised chunk-major so the tables of the chunk being gathered stay L2-resident.
// ------------------------------------------------------------------------------------------
constexpr int XT_LEVELS = 4;        // x windows 2,4,8,16 (level j = 1..4; j = 0 is the map itself)
constexpr int YT_LEVELS = 1;        // y windows 2 (level i = 1; i = 0 is a single row)
constexpr int XT_TABLES = (YT_LEVELS + 1) * (XT_LEVELS + 1) - 1;  // every (i, j) except (0, 0)
constexpr int XT_SLOTS = 8;         // 16-byte channel vectors per build CTA (128 B of channels)
__host__ __device__ constexpr int xt_index(int i, int j) { return i * (XT_LEVELS + 1) + j - 1; }

// 16-byte read-only load that is skipped (and leaves `v`) when `pred` is false: no branch, so a run of them is issued
// back to back and all their latencies overlap (the gather's bins are warp-uniformly present / absent)
__device__ __forceinline__ uint4 ldg128_if(const void* p, bool pred, uint4 v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "l"(p), "r"((unsigned)pred));
  return v;
}

struct VecF32 {
  typedef float4 T;
  static __device__ __forceinline__ T ldg_if(const T* p, bool pred, T init) {
    const uint4 u = ldg128_if(p, pred, make_uint4(__float_as_uint(init.x), __float_as_uint(init.y), __float_as_uint(init.z), __float_as_uint(init.w)));
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  }
  static __device__ __forceinline__ T vmax(T a, T b) {
    return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
  }
  static __device__ __forceinline__ T lowest() { return make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX); }
  static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ T scale(T m, float mul) {
    return make_float4(__fmul_rn(m.x, mul), __fmul_rn(m.y, mul), __fmul_rn(m.z, mul), __fmul_rn(m.w, mul));
  }
};
struct VecBF16 {
  typedef uint4 T;
  static __device__ __forceinline__ T ldg_if(const T* p, bool pred, T init) { return ldg128_if(p, pred, init); }
  static __device__ __forceinline__ T vmax(T a, T b) { return max8bf(a, b); }
  static __device__ __forceinline__ T lowest() { return make_uint4(0xFF7FFF7Fu, 0xFF7FFF7Fu, 0xFF7FFF7Fu, 0xFF7FFF7Fu); }
  static __device__ __forceinline__ T zero() { return make_uint4(0u, 0u, 0u, 0u); }
  static __device__ __forceinline__ T scale(T m, float mul) {
    T o;
    const __nv_bfloat162* pm = reinterpret_cast<const __nv_bfloat162*>(&m);
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(pm[i]);
      po[i] = __floats2bfloat162_rn(__fmul_rn(f.x, mul), __fmul_rn(f.y, mul));
    }
    return o;
  }
};

// tables: [XT_TABLES][h][w][CV] vectors, table (i, j) at xt_index(i, j):
//   T_ij[y][x] = max F[y .. min(y+2^i, h)) [x .. min(x+2^j, w))     (windows truncated at the map edge)
// grid = (h, CV / XT_SLOTS), block = 256, smem = 4 * w * XT_SLOTS vectors (rows y and y+1, ping-pong)
template <typename V>
__global__ void __launch_bounds__(256)
xtable_build_kernel(const typename V::T* __restrict__ feat, int h, int w, int CV, typename V::T* __restrict__ tables) {
  typedef typename V::T T;
  extern __shared__ __align__(16) unsigned char xt_smem[];
  const int n = w * XT_SLOTS;
  T* a0 = reinterpret_cast<T*>(xt_smem);  // row y, level j-1
  T* a1 = a0 + n;                         // row y, level j
  T* b0 = a1 + n;                         // row y+1 (or a copy of row y on the last row), level j-1
  T* b1 = b0 + n;
  const int y = blockIdx.x, cv0 = blockIdx.y * XT_SLOTS;
  const int y1 = min(y + 1, h - 1);
  const size_t plane = (size_t)h * w * CV;
  const T* row0 = feat + (size_t)y * w * CV + cv0;
  const T* row1 = feat + (size_t)y1 * w * CV + cv0;
  T* out0 = tables + (size_t)y * w * CV + cv0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t off = (size_t)(i / XT_SLOTS) * CV + (i % XT_SLOTS);
    const T u = __ldg(row0 + off), v = __ldg(row1 + off);
    a0[i] = u;
    b0[i] = v;
    out0[(size_t)xt_index(1, 0) * plane + off] = V::vmax(u, v);
  }
  __syncthreads();
#pragma unroll 1
  for (int j = 1; j <= XT_LEVELS; ++j) {
    const int half = 1 << (j - 1);
    T* t0 = out0 + (size_t)xt_index(0, j) * plane;
    T* t1 = out0 + (size_t)xt_index(1, j) * plane;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int x = i / XT_SLOTS, sl = i % XT_SLOTS;
      T u = a0[i], v = b0[i];
      if (x + half < w) {  // truncated at the right edge
        u = V::vmax(u, a0[i + half * XT_SLOTS]);
        v = V::vmax(v, b0[i + half * XT_SLOTS]);
      }
      a1[i] = u;
      b1[i] = v;
      const size_t off = (size_t)x * CV + sl;
      t0[off] = u;
      t1[off] = V::vmax(u, v);
    }
    __syncthreads();
    T* t = a0; a0 = a1; a1 = t;
    t = b0; b0 = b1; b1 = t;
  }
}

// One work item = (ROI r, 512-byte channel chunk): item = chunk * R + r (chunk-major, so the tables being
// gathered stay in L2).  grid = #items (one item per CTA) or a small persistent grid that strides over the
// items (used while a GEMM shares the SMs).  block = 224: warp ph (0..6) owns bin row ph, lane = 16-byte vector
// inside the chunk.  Bin geometry (torchvision roi_pool, see roi_geometry above) is computed once per
// CTA by 14 threads; the per-lookup work is one 32-bit multiply-add + one 16-byte load (the first version
// recomputed geometry and 64-bit addresses per lookup and was instruction-issue bound: 45 instr/load).
template <typename V, bool PERSIST>
__device__ __forceinline__ void
roipool_gather_body(const typename V::T* __restrict__ feat, const typename V::T* __restrict__ tables, int h, int w,
                      int CV, const float* __restrict__ boxes, const float* __restrict__ obj, float scale,
                      typename V::T* __restrict__ out, int R, int nitems) {
  typedef typename V::T T;
  __shared__ int s_y0[7], s_ylast[7], s_nrow[7], s_i[7];      // per bin row: first window row, last window row, #windows, y level
  __shared__ int s_x0[7], s_x1[7], s_j[7], s_wd[7];           // per bin col: first / right-aligned window col, x level, width
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int item = PERSIST ? (int)blockIdx.x : 0; item < (PERSIST ? nitems : 1); item += (int)gridDim.x) {
  const int r = PERSIST ? item % R : (int)blockIdx.x;
  const int cv = (PERSIST ? item / R : (int)blockIdx.y) * 32 + lane;
  if (PERSIST && item != (int)blockIdx.x) __syncthreads();  // the previous item's geometry has been consumed
  if (threadIdx.x < 14) {
    const float* box = boxes + 4 * (size_t)r;
    const bool is_y = threadIdx.x < 7;
    const int p = is_y ? threadIdx.x : threadIdx.x - 7;
    const int lim = is_y ? h : w;
    const int s0 = (int)roundf(__fmul_rn(box[is_y ? 1 : 0], scale));
    const int e0 = (int)roundf(__fmul_rn(box[is_y ? 3 : 2], scale));
    const int rl = max(e0 - s0 + 1, 1);
    const float bs = __fdiv_rn((float)rl, 7.f);
    const int lo = min(max((int)floorf(__fmul_rn((float)p, bs)) + s0, 0), lim);
    const int hi = min(max((int)ceilf(__fmul_rn((float)(p + 1), bs)) + s0, 0), lim);
    const int len = hi - lo;  // <= 0: empty bin
    if (is_y) {
      const int i = len >= 2 ? 1 : 0;
      s_i[p] = i;
      s_y0[p] = lo;
      s_ylast[p] = hi - (1 << i);
      s_nrow[p] = len > 0 ? (len + (1 << i) - 1) >> i : 0;
    } else {
      int j = len > 0 ? 31 - __clz(len) : 0;
      if (j > XT_LEVELS) j = XT_LEVELS;
      s_j[p] = j;
      s_x0[p] = lo;
      s_x1[p] = hi - (1 << j);
      s_wd[p] = len;
    }
  }
  __syncthreads();
  if (cv >= CV) continue;
  const int ph = warp;
  const float mul = obj ? __fadd_rn(__ldg(obj + r), 1.f) : 1.f;
  const int i = s_i[ph], y0 = s_y0[ph], ylast = s_ylast[ph], nrow = s_nrow[ph];
  const int ystep = 1 << i;
  const unsigned plane = (unsigned)h * w * CV;       // vectors per table (< 2^27 for any map that fits the builder)
  const unsigned rowpitch = (unsigned)w * CV;
  T* orow = out + ((size_t)r * 49 + ph * 7) * CV + cv;
#pragma unroll 1
  for (int pw = 0; pw < 7; ++pw) {
    const int j = s_j[pw], wd = s_wd[pw];
    T m = V::zero();
    if (nrow > 0 && wd > 0) {
      m = V::lowest();
      const T* tab = ((i | j) == 0 ? feat : tables + (size_t)xt_index(i, j) * plane) + cv;
      const unsigned xo0 = (unsigned)s_x0[pw] * CV, xo1 = (unsigned)s_x1[pw] * CV;
      const int win = 1 << j;
      if (wd == win) {
        // the bin is exactly one table window wide: ONE lookup per window row (a fifth of all lookups on the bench boxes;
        // the two-lookup form below would fetch the same 16 bytes twice -- the gather runs at the L2 bandwidth limit)
        int k = 0;
        for (; k + 1 < nrow; k += 2) {
          const unsigned ra = (unsigned)(y0 + k * ystep) * rowpitch;
          const unsigned rb = (unsigned)min(y0 + (k + 1) * ystep, ylast) * rowpitch;
          const T a0 = __ldg(tab + ra + xo0), b0 = __ldg(tab + rb + xo0);
          m = V::vmax(m, V::vmax(a0, b0));
        }
        if (k < nrow) {
          const unsigned ra = (unsigned)min(y0 + k * ystep, ylast) * rowpitch;
          m = V::vmax(m, __ldg(tab + ra + xo0));
        }
      } else if (wd <= 2 * win) {
        // two overlapping windows per window row, two rows in flight
        int k = 0;
        for (; k + 1 < nrow; k += 2) {
          const unsigned ra = (unsigned)(y0 + k * ystep) * rowpitch;
          const unsigned rb = (unsigned)min(y0 + (k + 1) * ystep, ylast) * rowpitch;
          const T a0 = __ldg(tab + ra + xo0), a1 = __ldg(tab + ra + xo1);
          const T b0 = __ldg(tab + rb + xo0), b1 = __ldg(tab + rb + xo1);
          m = V::vmax(m, V::vmax(V::vmax(a0, a1), V::vmax(b0, b1)));
        }
        if (k < nrow) {
          const unsigned ra = (unsigned)min(y0 + k * ystep, ylast) * rowpitch;
          m = V::vmax(m, V::vmax(__ldg(tab + ra + xo0), __ldg(tab + ra + xo1)));
        }
      } else {
        // very wide bins (> 32 cells): step full windows, finish right-aligned
        for (int k = 0; k < nrow; ++k) {
          const unsigned ra = (unsigned)min(y0 + k * ystep, ylast) * rowpitch;
          for (unsigned xo = xo0; xo < xo1; xo += (unsigned)win * CV) m = V::vmax(m, __ldg(tab + ra + xo));
          m = V::vmax(m, __ldg(tab + ra + xo1));
        }
      }
    }
    __stcs(orow + (size_t)pw * CV, V::scale(m, mul));
  }
  }  // items
}

// Per-ROI bin geometry (torchvision roi_pool, see roi_geometry above) into shared memory, by threads 0..13
struct RoiBinTable {
  int y0[7], ylast[7], nrow[7], i[7];   // per bin row: first window row, last window row, #windows, y level
  int x0[7], x1[7], j[7], wd[7];        // per bin col: first / right-aligned window col, x level, width
};
__device__ __forceinline__ void roi_bin_table(const float* __restrict__ box, float scale, int h, int w, RoiBinTable& s,
                                              int tid = threadIdx.x) {
  if (tid < 14) {
    const bool is_y = tid < 7;
    const int p = is_y ? tid : tid - 7;
    const int lim = is_y ? h : w;
    const int s0 = (int)roundf(__fmul_rn(box[is_y ? 1 : 0], scale));
    const int e0 = (int)roundf(__fmul_rn(box[is_y ? 3 : 2], scale));
    const int rl = max(e0 - s0 + 1, 1);
    const float bs = __fdiv_rn((float)rl, 7.f);
    const int lo = min(max((int)floorf(__fmul_rn((float)p, bs)) + s0, 0), lim);
    const int hi = min(max((int)ceilf(__fmul_rn((float)(p + 1), bs)) + s0, 0), lim);
    const int len = hi - lo;  // <= 0: empty bin
    if (is_y) {
      const int i = len >= 2 ? 1 : 0;
      s.i[p] = i;
      s.y0[p] = lo;
      s.ylast[p] = hi - (1 << i);
      s.nrow[p] = len > 0 ? (len + (1 << i) - 1) >> i : 0;
    } else {
      int j = len > 0 ? 31 - __clz(len) : 0;
      if (j > XT_LEVELS) j = XT_LEVELS;
      s.j[p] = j;
      s.x0[p] = lo;
      s.x1[p] = hi - (1 << j);
      s.wd[p] = len;
    }
  }
}

// Default gather: one CTA per (8 ROIs, 512-byte channel chunk), 14 warps: warp (ph, half) owns bins pw = 4 half .. of bin
// ROW ph and walks the row's table windows ONCE for all of them: the row offset, the bins' table pointers and the
// one-or-two-lookups decision are set up once per warp and ROI, the eight lookups of a window row are predicated loads with
// no branch in between (all in flight together), then folded in.  The first table version looped (bin, window row) the
// other way round and spent ~48 issued instructions per lookup (260 M warp instructions per launch at the bench workload;
// profiles/r2_ncu_step_per_launch.txt).  Measured steps (R50 bench workload, tables + gather, warm): bin-major 410 us;
// row-major with branches around the loads 608 us (the loads of a row serialise); predicated loads 482 us; + 8 ROIs per
// CTA 389 us; cp.async into shared memory (28 lookups per lane in flight) 439 us.  All variants move the same ~2.8 GB
// L2 -> SM and sit at 9-11 TB/s: the lookups, not the instructions, are the floor of this design.
// Bins wider than two table windows (> 32 cells: maps wider than ~230 cells) step full windows.
constexpr int GATHER_THREADS = 448;
constexpr int GATHER_NR = 8;   // ROIs per CTA: one geometry pass + barrier per 8 ROIs; a CTA that lives for one ROI spends ~40 % of
                               // its ~5 us on launch, box load and the barrier with no lookups in flight (measured: 2 CTAs / SM, 0.54 ms)
template <typename V, int MINB>
__global__ void __launch_bounds__(GATHER_THREADS, MINB)
roipool_gather_kernel(const typename V::T* __restrict__ feat, const typename V::T* __restrict__ tables, int h, int w,
                      int CV, const float* __restrict__ boxes, const float* __restrict__ obj, float scale,
                      typename V::T* __restrict__ out, int R, int nitems) {
  typedef typename V::T T;
  __shared__ RoiBinTable gs[GATHER_NR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ph = warp >> 1, pw0 = (warp & 1) * 4, nb = (warp & 1) ? 3 : 4;
  const int r0 = (int)blockIdx.x * GATHER_NR, cv = (int)blockIdx.y * 32 + lane;
  {
    const int q = threadIdx.x / 14;
    if (q < GATHER_NR && r0 + q < R) roi_bin_table(boxes + 4 * (size_t)(r0 + q), scale, h, w, gs[q], threadIdx.x - 14 * q);
  }
  __syncthreads();
  if (cv >= CV) return;
  const unsigned plane = (unsigned)h * w * CV;       // vectors per table (< 2^27 for any map that fits the builder)
  const unsigned rowpitch = (unsigned)w * CV;
  const T low = V::lowest();
#pragma unroll 1
  for (int q = 0; q < GATHER_NR && r0 + q < R; ++q) {
    const RoiBinTable& g = gs[q];
    const int r = r0 + q;
    const float mul = obj ? __fadd_rn(__ldg(obj + r), 1.f) : 1.f;
    const int i = g.i[ph], y0 = g.y0[ph], ylast = g.ylast[ph], nrow = g.nrow[ph];
    const int ystep = 1 << i;
    const T* p0[4];     // first window of the bin in its table, at row 0
    unsigned d1[4];     // distance to the right-aligned second window
    unsigned one = 0, two = 0, wide = 0;  // bit b: bin non-empty / needs the second lookup / wider than two windows
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int pw = min(pw0 + b, 6);
      const int j = g.j[pw], wd = (b < nb) ? g.wd[pw] : 0, x0 = g.x0[pw], x1 = g.x1[pw];
      p0[b] = ((i | j) == 0 ? feat : tables + (size_t)xt_index(i, j) * plane) + cv + (unsigned)x0 * CV;
      d1[b] = (unsigned)max(x1 - x0, 0) * CV;
      if (wd > 0) one |= 1u << b;
      if (wd > (1 << j)) two |= 1u << b;
      if (wd > (2 << j)) wide |= 1u << b;
    }
    T acc[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[b] = low;
    if (wide == 0) {
#pragma unroll 1
      for (int k = 0; k < nrow; ++k) {
        const unsigned ro = (unsigned)min(y0 + k * ystep, ylast) * rowpitch;
        T v0[4], v1[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {   // eight loads in flight, no branches in between
          v0[b] = V::ldg_if(p0[b] + ro, (one >> b) & 1u, low);
          v1[b] = V::ldg_if(p0[b] + ro + d1[b], (two >> b) & 1u, low);
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[b] = V::vmax(acc[b], V::vmax(v0[b], v1[b]));
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < nrow; ++k) {
        const unsigned ro = (unsigned)min(y0 + k * ystep, ylast) * rowpitch;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (one & (1u << b)) {
            const unsigned step = (unsigned)(1 << g.j[min(pw0 + b, 6)]) * CV;
            acc[b] = V::vmax(acc[b], __ldg(p0[b] + ro + d1[b]));
            for (unsigned xo = 0; xo < d1[b]; xo += step) acc[b] = V::vmax(acc[b], __ldg(p0[b] + ro + xo));
          }
        }
      }
    }
    T* orow = out + ((size_t)r * 49 + ph * 7 + pw0) * CV + cv;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      if (b < nb) {
        const T m = (nrow > 0 && (one & (1u << b))) ? acc[b] : V::zero();
        __stcs(orow + (size_t)b * CV, V::scale(m, mul));
      }
    }
  }
}

template <typename V>
__global__ void __launch_bounds__(224)
roipool_gather_binmajor_kernel(const typename V::T* __restrict__ feat, const typename V::T* __restrict__ tables, int h, int w,
                               int CV, const float* __restrict__ boxes, const float* __restrict__ obj, float scale,
                               typename V::T* __restrict__ out, int R, int nitems) {
  roipool_gather_body<V, false>(feat, tables, h, w, CV, boxes, obj, scale, out, R, nitems);
}
// persistent variant: capped at 40 registers so that one or two of its CTAs fit next to a resident GEMM CTA
template <typename V>
__global__ void __maxnreg__(40)
roipool_gather_persistent_kernel(const typename V::T* __restrict__ feat, const typename V::T* __restrict__ tables, int h,
                                 int w, int CV, const float* __restrict__ boxes, const float* __restrict__ obj,
                                 float scale, typename V::T* __restrict__ out, int R, int nitems) {
  roipool_gather_body<V, true>(feat, tables, h, w, CV, boxes, obj, scale, out, R, nitems);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16(in[i]);
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

}  // namespace drn

using namespace drn;

template <typename V>
static int roipool_v2(const void* feat, int h, int w, int CV, const float* boxes, const float* objectness, int R,
                      float spatial_scale, void* out, void* ws, bool build, bool gather, int max_ctas, cudaStream_t st) {
  typedef typename V::T T;
  const size_t smem = (size_t)4 * w * XT_SLOTS * sizeof(T);
  DRN_CHECK_ARG(smem <= 200 * 1024, "roipool: feature map too wide for the table builder (w=%d)", w);
  DRN_CHECK_ARG((unsigned long long)XT_TABLES * h * w * CV < (1ull << 31), "roipool: feature map too large for 32-bit table offsets");
  if (build) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(xtable_build_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      if (e != cudaSuccess) return set_err("roipool: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      configured = true;
    }
    xtable_build_kernel<V><<<dim3(h, CV / XT_SLOTS), 256, smem, st>>>((const T*)feat, h, w, CV, (T*)ws);
    DRN_CHECK_LAUNCH("roipool xtable build");
  }
  if (gather && R > 0) {
    const long long nitems = (long long)R * cdiv(CV, 32);
    DRN_CHECK_ARG(nitems < (1ll << 31), "roipool: too many (ROI, chunk) items");
    const int grid = (max_ctas > 0 && max_ctas < nitems) ? max_ctas : (int)nitems;
    if (grid < nitems) {
      // a GEMM CTA with ~210 KB of shared memory has to fit on the same SM: both kernels must run under the
      // same (maximum) shared-memory carve-out, or the SM drains before it switches configuration
      static bool carveout_set = false;
      if (!carveout_set) {
        cudaError_t e = cudaFuncSetAttribute(roipool_gather_persistent_kernel<V>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return set_err("roipool: cudaFuncSetAttribute(carveout): %s", cudaGetErrorString(e));
        carveout_set = true;
      }
      roipool_gather_persistent_kernel<V><<<grid, 224, 0, st>>>((const T*)feat, (const T*)ws, h, w, CV, boxes, objectness,
                                                           spatial_scale, (T*)out, R, (int)nitems);
    } else {
      // DRN_ROIPOOL_GATHER=0 (measurement switch): the first, bin-major table kernel
      static int variant = -1;
      if (variant < 0) {
        const char* e = getenv("DRN_ROIPOOL_GATHER");
        variant = e ? atoi(e) : 2;
      }
      const dim3 grid2(R, cdiv(CV, 32)), gridn(cdiv(R, GATHER_NR), cdiv(CV, 32));
      if (variant == 0)
        roipool_gather_binmajor_kernel<V><<<grid2, 224, 0, st>>>((const T*)feat, (const T*)ws, h, w, CV, boxes, objectness,
                                                                  spatial_scale, (T*)out, R, (int)nitems);
      else
        roipool_gather_kernel<V, 2><<<gridn, GATHER_THREADS, 0, st>>>((const T*)feat, (const T*)ws, h, w, CV, boxes, objectness,
                                                                       spatial_scale, (T*)out, R, (int)nitems);
    }
    DRN_CHECK_LAUNCH("roipool gather");
  }
  return 0;
}

static int roipool_tables_dispatch(const void* feat, int h, int w, int C, const float* boxes, const float* objectness,
                                   int R, float spatial_scale, int dtype, void* out, void* workspace,
                                   size_t workspace_bytes, bool build, bool gather, int max_ctas, cudaStream_t st) {
  const int vec = dtype == DRN_BF16 ? 8 : 4;
  DRN_CHECK_ARG(feat && workspace, "roipool: null pointer");
  DRN_CHECK_ARG(h > 0 && w > 0, "roipool: empty feature map");
  DRN_CHECK_ARG(C % vec == 0 && (C / vec) % XT_SLOTS == 0, "roipool: C=%d does not fit the table layout", C);
  DRN_CHECK_ARG(workspace_bytes >= drn_roipool_workspace_bytes(h, w, C, dtype),
                "roipool: workspace of %zu bytes is smaller than drn_roipool_workspace_bytes()", workspace_bytes);
  DRN_CHECK_ARG((uintptr_t)workspace % 16 == 0, "roipool: workspace must be 16-byte aligned");
  if (dtype == DRN_BF16)
    return roipool_v2<VecBF16>(feat, h, w, C / vec, boxes, objectness, R, spatial_scale, out, workspace, build, gather, max_ctas, st);
  return roipool_v2<VecF32>(feat, h, w, C / vec, boxes, objectness, R, spatial_scale, out, workspace, build, gather, max_ctas, st);
}

extern "C" {

int drn_version(void) { return 100; }
const char* drn_last_error(void) { return drn::err_buf(); }

int drn_conv3x3_c3_fwd(const float* img, int H, int W, int Hp, int Wp, const float* mean3, const float* std3,
                       const float* w_packed, const float* scale, const float* bias, int Cout,
                       int stride, int relu, void* out, int out_dtype, drn_stream_t stream) {
  DRN_CHECK_ARG(img && mean3 && std3 && w_packed && bias && out, "conv3x3_c3: null pointer");
  DRN_CHECK_ARG(Cout % 16 == 0 && Cout <= 256, "conv3x3_c3: Cout=%d must be a multiple of 16, <=256", Cout);
  DRN_CHECK_ARG(stride == 1 || stride == 2, "conv3x3_c3: stride %d", stride);
  DRN_CHECK_ARG(H > 0 && W > 0 && Hp >= H && Wp >= W, "conv3x3_c3: bad image/canvas size %dx%d in %dx%d", H, W, Hp, Wp);
  const int Ho = (Hp + 2 - 3) / stride + 1, Wo = (Wp + 2 - 3) / stride + 1;
  const long long threads = (long long)Ho * ((Wo + C3_NPX - 1) / C3_NPX) * (Cout / 16);
  const int grid = (int)((threads + 127) / 128);
  const size_t smem = 27 * Cout * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
// DRN_C3_TILED=0 (measurement switch) keeps the per-thread-normalising kernel for Cout == 64 too
  static int tiled_env = -1;
  if (tiled_env < 0) {
    const char* e = getenv("DRN_C3_TILED");
    tiled_env = (e && e[0] == '0') ? 0 : 1;
  }
  const bool tiled = tiled_env && Cout == 64;
  // DRN_C3_TH (measurement switch): rows per tile of the tiled kernel; unset = c3_pick_th
  static int th_env = -1, sms = 0;
  if (th_env < 0) {
    const char* e = getenv("DRN_C3_TH");
    th_env = e ? atoi(e) : 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int th = (th_env >= 2 && th_env <= C3_TH_MAX) ? (th_env & ~1) : c3_pick_th(Ho, Wo, sms);
  const dim3 tgrid((Wo + C3_TW - 1) / C3_TW, (Ho + th - 1) / th);
  const int pcp = stride == 2 ? C3Patch<2>::PCP : C3Patch<1>::PCP;
  const size_t tsmem = (size_t)(27 * 64 + 3 * ((th - 1) * stride + 3) * pcp) * sizeof(float);  // <= 46.5 KB at th = 12, stride 2
#define DRN_C3_LAUNCH(T, S)                                                                                          \
  do {                                                                                                               \
    if (tiled)                                                                                                       \
      conv3x3_c3_tile_kernel<T, S><<<tgrid, C3_TILE_THREADS, tsmem, st>>>(img, H, W, mean3[0], mean3[1], mean3[2], std3[0],      \
                                                              std3[1], std3[2], w_packed, scale, bias, relu, Ho, Wo, \
                                                              th, (T*)out);                                          \
    else                                                                                                             \
      conv3x3_c3_kernel<T, S><<<grid, 128, smem, st>>>(img, H, W, mean3[0], mean3[1], mean3[2], std3[0], std3[1],    \
                                                       std3[2], w_packed, scale, bias, Cout, relu, Ho, Wo, (T*)out); \
  } while (0)
  if (out_dtype == DRN_F32) {
    if (stride == 1) DRN_C3_LAUNCH(float, 1); else DRN_C3_LAUNCH(float, 2);
  } else {
    if (stride == 1) DRN_C3_LAUNCH(__nv_bfloat16, 1); else DRN_C3_LAUNCH(__nv_bfloat16, 2);
  }
#undef DRN_C3_LAUNCH
  DRN_CHECK_LAUNCH("conv3x3_c3");
  return 0;
}

int drn_conv_igemm_f32(const float* in, int N, int H, int W, int Cin, const float* w, int ksize,
                       int dilation, const float* scale, const float* bias, const float* residual,
                       int relu, float* out, int Cout, int ldo, drn_stream_t stream) {
  DRN_CHECK_ARG(in && w && out, "conv_igemm_f32: null pointer");
  DRN_CHECK_ARG(ksize == 1 || ksize == 3, "conv_igemm_f32: ksize %d", ksize);
  DRN_CHECK_ARG(Cin % 16 == 0, "conv_igemm_f32: Cin=%d must be a multiple of 16", Cin);
  DRN_CHECK_ARG(Cout % 64 == 0, "conv_igemm_f32: Cout=%d must be a multiple of 64", Cout);
  DRN_CHECK_ARG(ldo >= Cout && ldo % 4 == 0, "conv_igemm_f32: ldo=%d", ldo);
  const long long M = (long long)N * H * W;
  if (M == 0) return 0;
  dim3 grid((unsigned)((M + CBM - 1) / CBM), Cout / CBN);
  conv_igemm_f32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(in, N, H, W, Cin, w, ksize, dilation,
      scale, bias, residual, relu, out, Cout, ldo);
  DRN_CHECK_LAUNCH("conv_igemm_f32");
  return 0;
}

int drn_maxpool2x2_nhwc(const void* in, int N, int H, int W, int C, int stride, int dtype, void* out,
                        drn_stream_t stream) {
  DRN_CHECK_ARG(in && out, "maxpool: null pointer");
  DRN_CHECK_ARG(stride == 1 || stride == 2, "maxpool: stride %d", stride);
  DRN_CHECK_ARG(H >= 2 && W >= 2, "maxpool: map %dx%d too small", H, W);
  const int vec = dtype == DRN_BF16 ? 8 : 4;
  DRN_CHECK_ARG(C % vec == 0, "maxpool: C=%d not a multiple of %d", C, vec);
  const int Ho = (H - 2) / stride + 1, Wo = (W - 2) / stride + 1;
  const long long total = (long long)N * Ho * Wo * (C / vec);
  const int grid = (int)((total + 255) / 256);
  if (dtype == DRN_BF16)
    maxpool2x2_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(in, N, H, W, C, stride, Ho, Wo, out);
  else
    maxpool2x2_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(in, N, H, W, C, stride, Ho, Wo, out);
  DRN_CHECK_LAUNCH("maxpool2x2");
  return 0;
}

size_t drn_roipool_workspace_bytes(int h, int w, int C, int dtype) {
  if (h <= 0 || w <= 0 || C <= 0) return 0;
  return (size_t)XT_TABLES * h * w * C * (dtype == DRN_BF16 ? 2 : 4);
}

int drn_roipool_fwd(const void* feat, int h, int w, int C, const float* boxes, const float* objectness,
                    int R, float spatial_scale, int dtype, void* out, void* workspace, size_t workspace_bytes,
                    drn_stream_t stream) {
  DRN_CHECK_ARG(R >= 0, "roipool: R=%d", R);
  if (R == 0) return 0;  // empty proposal list: nothing to write (out may be a null, zero-size buffer)
  DRN_CHECK_ARG(feat && out && boxes, "roipool: null pointer");
  DRN_CHECK_ARG(h > 0 && w > 0, "roipool: empty feature map");
  const int vec = dtype == DRN_BF16 ? 8 : 4;
  DRN_CHECK_ARG(C % vec == 0, "roipool: C=%d not a multiple of %d", C, vec);
  cudaStream_t st = (cudaStream_t)stream;
  const int CV = C / vec;
  if (workspace && CV % XT_SLOTS == 0)
    return roipool_tables_dispatch(feat, h, w, C, boxes, objectness, R, spatial_scale, dtype, out, workspace, workspace_bytes,
                                   true, true, 0, st);
  // no workspace (or odd channel count): direct scan kernels
  dim3 grid(R, 7);
  if (dtype == DRN_BF16) {
    roipool_bf16_kernel<<<grid, 256, 0, st>>>((const uint4*)feat, h, w, C, boxes, objectness, spatial_scale, (uint4*)out);
  } else {
    roipool_f32_kernel<<<grid, 128, 0, st>>>((const float*)feat, h, w, C, boxes, objectness, spatial_scale, (float*)out);
  }
  DRN_CHECK_LAUNCH("roipool");
  return 0;
}

int drn_roipool_tables_supported(int C, int dtype) {
  const int vec = dtype == DRN_BF16 ? 8 : 4;
  return (C > 0 && C % vec == 0 && (C / vec) % XT_SLOTS == 0) ? 1 : 0;
}

int drn_roipool_build_tables(const void* feat, int h, int w, int C, int dtype, void* workspace, size_t workspace_bytes,
                             drn_stream_t stream) {
  return roipool_tables_dispatch(feat, h, w, C, nullptr, nullptr, 0, 0.f, dtype, nullptr, workspace, workspace_bytes, true,
                                 false, 0, (cudaStream_t)stream);
}

int drn_roipool_rows_fwd(const void* feat, int h, int w, int C, const float* boxes, const float* objectness, int R,
                         float spatial_scale, int dtype, void* out, const void* tables, size_t tables_bytes,
                         int max_ctas, drn_stream_t stream) {
  DRN_CHECK_ARG(R >= 0, "roipool: R=%d", R);
  if (R == 0) return 0;
  DRN_CHECK_ARG(out && boxes, "roipool: null pointer");
  return roipool_tables_dispatch(feat, h, w, C, boxes, objectness, R, spatial_scale, dtype, out, const_cast<void*>(tables),
                                 tables_bytes, false, true, max_ctas, (cudaStream_t)stream);
}

int drn_cast_f32_to_bf16(const float* in, void* out, int64_t n, drn_stream_t stream) {
  if (n == 0) return 0;
  cast_f32_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16*)out, n);
  DRN_CHECK_LAUNCH("cast_f32_to_bf16");
  return 0;
}
int drn_cast_bf16_to_f32(const void* in, float* out, int64_t n, drn_stream_t stream) {
  if (n == 0) return 0;
  cast_bf16_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, out, n);
  DRN_CHECK_LAUNCH("cast_bf16_to_f32");
  return 0;
}

}  // extern "C"
