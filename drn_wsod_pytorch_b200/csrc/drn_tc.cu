// tcgen05 / TMEM / TMA implicit-GEMM for sm_100a: the dense contractions of the hot path
// (ResNet-WS / VGG 3x3 + 1x1 convs, fc6, fc7, concatenated heads) in bf16 with fp32 accumulation.
//
//   out[m][n] = drop( act( (sum_k A[m][k] * W[n][k]) * scale[n] + bias[n] + residual[m][n] ) )
//
// A is either a row-major [M][K] matrix (linear layers, 1x1 convs: 2D TMA) or an NHWC activation
// tensor read through a 4D TMA box of TILE_H x TILE_W output pixels x 64 channels shifted by the filter
// tap (3x3 convs: implicit im2col, zero padding comes from TMA out-of-bounds fill).  W is [N][K] K-major.
//
// Persistent warp-specialised CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-
// thread tcgen05.mma issuer, warps 2..5 = epilogue.  smem ring of STAGES x (128x64 A + BNx64 B) bf16
// tiles in the 128B-swizzled K-major canonical layout; two TMEM accumulator stages (2 x BN fp32 columns)
// so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Residual add (ResNet shortcuts) rides the tensor core: after the K loop the producer streams the
// residual tile through the same smem ring as 64-column "k-blocks" and the MMA thread multiplies them
// with a 64x64 identity that lives in smem (D[:, 64c:64c+64] += R_c * I, exact in fp32), so the
// shortcut is prefetched STAGES deep like any operand and the epilogue never waits on a load
// (a TMA-loaded residual in the epilogue was latency-bound: 33 us -> see profiles/).  FrozenBN scale
// is folded into the weights by the host when a residual is present (the kernel rejects scale+residual).
//
// Epilogue (bf16 output): every warp owns the 32 accumulator rows of its TMEM lane quadrant and walks
// the tile in 64-column chunks: tcgen05.ld -> scale/bias (from smem) -> ReLU -> dropout -> bf16 ->
// 128B-swizzled 32x64 staging buffer -> one TMA store per chunk, so all global traffic of the kernel
// is TMA (full 128 B lines) and clipping of partial tiles is done by the TMA unit.  fp32 output (head
// logits, a few hundred KB) keeps a direct register->global path.
#include "common.cuh"
#pragma nv_diag_suppress 128  // "loop is not reachable": the `if constexpr (HALO) { ...; continue; }` branches of gemm_tc_kernel
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace drn {
namespace tc {

constexpr int BM = 128;      // UMMA M (cta_group::1)
constexpr int BK = 64;       // one 128-byte swizzle atom of bf16 along K
constexpr int EPI_CHUNK = 64;                    // columns per staging buffer (128 B of bf16)
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK * 2;  // 32 rows x 128 B
constexpr int IDENT_BYTES = 64 * 64 * 2;           // 64x64 bf16 identity (B operand of the residual MMAs)
// HALO mode (3x3 convs without shortcut): the output tile is one image row of 128 pixels; per 64-channel block the three
// input rows h-d, h, h+d (128 + 2d pixels each, zero-filled outside the map by TMA) are staged ONCE and all nine filter
// taps read them through UMMA descriptors whose start address is shifted by the tap (dy: whole row buffers, dx: d pixels
// = d * 128 B inside the 128B-swizzled row), so every input pixel crosses L2->SM ~3x instead of 9x.  Only the weights
// stream through the STAGES ring.
constexpr int HALO_MAX_DIL = 4;
constexpr int HALO_ROW_BYTES = ((BM + 2 * HALO_MAX_DIL) * 128 + 1023) / 1024 * 1024;  // one input row: (128 + 2d) px x 128 B, 1 KB aligned
constexpr int HALO_SLOT_BYTES = 3 * HALO_ROW_BYTES;
constexpr int HALO_SLOTS = 2;

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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// One lane of a converged warp (PTX elect.sync).  The TMA / MMA role warps run their loops with all 32 lanes in
// warp-uniform control flow and guard only the issuing instructions with this: ptxas then feeds the instructions'
// uniform-register operands straight from the elected lane.  Under a plain `if (lane == 0)` it cannot prove that one
// thread is active and wraps EVERY tcgen05.mma / cp.async.bulk.tensor in an ELECT + 5x R2UR.BROADCAST + BRA.U.ANY
// waterfall loop: ~127 SASS instructions per k-block in the single issuing thread, ~0.4 us per k-block whatever the tile
// (profiles/r2_conv_ablation.txt, r2_ncu_mma_issue_loop.txt) -- that, not loads or MMAs, bounded every conv layer.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants.  In a 2-CTA cluster the even CTA (rank 0) is the MMA leader; the
// shared::cluster address of "the same variable in the leader" is the local address with the peer bit
// cleared (cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask).
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
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

// The same descriptor as two 32-bit words: only the low word (start address, LBO) varies, the high word (SBO, version,
// swizzle mode) is a constant.  The issue loops do all descriptor arithmetic in 32 bits so that it stays on the uniform
// datapath (64-bit shifts / ors are vector-only and cost an R2UR pair per operand and MMA).
constexpr uint32_t DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }

// kind::f16 instruction descriptor: D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1, K-major A/B,
// N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int n, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(DESC_HI) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {  // arrive on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// The four K = 16 MMAs of one 64-wide k-block in ONE asm block: the descriptor low words advance by 2 (32 bytes along K inside
// the swizzle atom) in place, so the issuing thread marshals five operands per k-block instead of per MMA.
template <int CG>
__device__ __forceinline__ void umma_kblock(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t a_hi, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 2) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .b32 al, bl;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.u32 al, %1, 2;\n\t"
        "add.u32 bl, %2, 2;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, t;\n\t"
        "add.u32 al, %1, 4;\n\t"
        "add.u32 bl, %2, 4;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, t;\n\t"
        "add.u32 al, %1, 6;\n\t"
        "add.u32 bl, %2, 6;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, t;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(DESC_HI) : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p, t;\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .b32 al, bl;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 t, 0, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
        "add.u32 al, %1, 2;\n\t"
        "add.u32 bl, %2, 2;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
        "add.u32 al, %1, 4;\n\t"
        "add.u32 bl, %2, 4;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
        "add.u32 al, %1, 6;\n\t"
        "add.u32 bl, %2, 6;\n\t"
        "mov.b64 da, {al, %5};\n\t"
        "mov.b64 db, {bl, %6};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, t;\n\t"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(DESC_HI) : "memory");
  }
}
// wait / commit on a barrier given by its shared-space address (the issue loops keep running addresses instead of indexing)
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  if constexpr (CG == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
        "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
        "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// same counter-based RNG as drn_dropout_inplace (drn_heads.cu): one draw per output element
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}

// DRN_TC_DEBUG & 256 (tools/timeline_probe.py): per-launch timeline in nanoseconds (%globaltimer), slot = launch ordinal since
// drn_gemm_timeline_reset(): [0] earliest CTA start, [1] earliest "past griddepcontrol.wait", [2] earliest first-operand
// arrival (MMA warp), [3] latest CTA end.  (The two dies' timers are offset by ~1.6 ms: min and max stamps are only
// comparable among themselves.)
constexpr int TL_SLOTS = 512;
__device__ unsigned long long g_timeline[TL_SLOTS][4];
// DRN_TC_DEBUG & 2048 (tools/tile_trace.py): per-tile pipeline trace of CTA 0, SM clock stamps.  Row = tile ordinal of the CTA,
// column: 0 producer starts the tile, 1 producer has issued its last load, 2 MMA warp past the accumulator-free wait, 3 MMA warp
// past the first operand wait, 4 MMA warp has issued the tile's last commit, 5 epilogue warp 0 past the accumulator-full wait,
// 6 epilogue warp 0 has issued its last store of the tile, 7: row 0 = kernel start, row 1 = prologue done (past griddepcontrol.wait).
constexpr int TR_TILES = 48;
__device__ long long g_trace[TR_TILES][8];
__device__ __forceinline__ void trace(bool on, int tile_ord, int col) {
  if (on && tile_ord < TR_TILES && (threadIdx.x & 31) == 0) g_trace[tile_ord][col] = clock64();
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct Params {
  int tl_slot;  // -1: no timeline
  // problem
  int M;        // GEMM mode: rows.  conv mode: unused
  int N;        // output channels
  int KB;       // number of 64-wide K blocks (= taps * Cin/64)
  int conv;     // 0 = 2D A, 1 = 4D NHWC A with 3x3 taps
  int NB, H, W, Cin, dil;  // conv mode geometry
  int tile_w, tile_h;      // conv mode: output-pixel tile, tile_w * tile_h == 128, tile_w in {16,32,64,128}
  int tiles_h, tiles_w;
  int num_m_tiles, num_n_tiles;
  // epilogue
  const float* scale;
  const float* bias;
  int has_residual;        // residual arrives through map_r (bf16, same geometry as the output)
  void* out;               // used by the fp32 direct path only
  int out_f32;
  int ldo;
  int relu;
  // fused dropout (train-mode fc6/fc7): keep iff mix32(seed*P + m*N + n) < keep_thresh
  uint32_t drop_keep_thresh;  // 0 = no dropout
  float drop_inv_keep;
  unsigned long long drop_seed;
  const unsigned long long* drop_seed_dev;  // optional device-resident addend (fresh masks under graph replay)
  // reduce-scatter epilogue (fp32 output, data-parallel weight gradient): output row m belongs to rank m / rows_per_peer and
  // is stored straight into THAT rank's accumulation window, slot `my_rank` -- peer memory mapped over NVLink
  float* peer_out[8];
  int n_peers, my_rank, rows_per_peer;
  int bres;     // HALO mode, Cin = 64, one N tile: the nine weight tiles (72 KB) are loaded once per CTA and stay in the ring's nine slots
  int halo_bo;  // HALO mode: also set the descriptor's base-offset field to the start's 128 B row phase (hardware probe switch)
  int debug;  // DRN_TC_DEBUG bit mask (profiling experiments only; every bit but 256 / 2048 gives WRONG results or a slower path):
              //   1 skip the output stores, 2 skip the epilogue math, 32 skip the A-operand loads, 256 per-launch timeline
              //   (tools/timeline_probe.py), 512 scalar instead of packed epilogue math, 1024 per-tap issue loop for
              //   resident-weight HALO layers, 2048 per-tile pipeline trace of CTA 0 (tools/tile_trace.py)
  // stream-K (deep-K GEMMs whose tile count does not fill whole waves): every unit gets an equal share of the
  // tiles x K-blocks iteration space; a unit that starts inside a tile dumps that partial accumulator to sk_ws
  // and raises sk_flags[unit], the unit that began the tile (and reaches it last in time) adds it and runs the epilogue
  // stream_k == 2, "tail split-K": the whole waves of tiles run data-parallel; the tiles of the last, partial wave
  // are cut into sk_parts K-ranges that are scheduled K-range-major (so concurrently running units still sit at the
  // same K offset and share their operand tiles in L2 -- what plain stream-K loses); parts 0..sk_parts-2 dump their
  // fp32 partial, the last part (scheduled last) folds them in, in part order, and runs the epilogue.
  int stream_k;
  int sk_tail0, sk_tail, sk_parts;  // first tail tile, number of tail tiles, K-ranges per tail tile
  float* sk_ws;
  unsigned int* sk_flags;
};

// kind: 0 = whole tile, 1 = dumps its partial accumulator to workspace slot `slot`, 2 = owner: folds in the partials
// of slots [slot, slot + nslots) and runs the epilogue
struct Piece { int tile, kb0, kb1, kind, slot, nslots; };

struct Sched {
  int KB, num_tiles, num_units, sk, cur, unit;
  int tail0, tail, parts, v;
  long long g, g1;
  __device__ Sched(int unit_, int num_units_, int num_tiles_, const Params& p)
      : KB(p.KB), num_tiles(num_tiles_), num_units(num_units_), sk(p.stream_k), cur(unit_), unit(unit_),
        tail0(p.sk_tail0), tail(p.sk_tail), parts(p.sk_parts), v(unit_) {
    const long long total = (long long)num_tiles_ * KB;
    const long long W = (total + num_units_ - 1) / num_units_;
    g = (long long)unit_ * W;
    g1 = g + W < total ? g + W : total;
  }
  __device__ bool next(Piece& pc) {
    pc.slot = 0; pc.nslots = 0;
    if (sk == 0) {
      if (cur >= num_tiles) return false;
      pc.tile = cur; pc.kb0 = 0; pc.kb1 = KB; pc.kind = 0;
      cur += num_units;
      return true;
    }
    if (sk == 2) {
      if (cur < tail0) {  // whole waves, data-parallel
        pc.tile = cur; pc.kb0 = 0; pc.kb1 = KB; pc.kind = 0;
        cur += num_units;
        return true;
      }
      if (v >= tail * parts) return false;
      const int q = v / tail, t = v - q * tail;  // K-range-major order
      pc.tile = tail0 + t;
      pc.kb0 = (int)((long long)q * KB / parts);
      pc.kb1 = (int)((long long)(q + 1) * KB / parts);
      if (q == parts - 1) { pc.kind = 2; pc.slot = t * (parts - 1); pc.nslots = parts - 1; }
      else { pc.kind = 1; pc.slot = t * (parts - 1) + q; }
      v += num_units;
      return true;
    }
    if (g >= g1) return false;
    pc.tile = (int)(g / KB);
    pc.kb0 = (int)(g - (long long)pc.tile * KB);
    const long long left = g1 - g;
    pc.kb1 = (KB - pc.kb0 <= left) ? KB : pc.kb0 + (int)left;
    pc.kind = pc.kb0 > 0 ? 1 : (pc.kb1 < KB ? 2 : 0);
    if (pc.kind == 1) pc.slot = unit;
    if (pc.kind == 2) { pc.slot = unit + 1; pc.nslots = 1; }
    g += pc.kb1 - pc.kb0;
    return true;
  }
};

__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// One 32-row x 64-column epilogue chunk of one warp: TMEM -> (scale) + bias -> max(., floor) -> (dropout) ->
// bf16 -> this lane's 128-byte row of the swizzled staging slot.  All smem traffic uses explicit
// shared-space instructions (LDS/STS): generic LD.E/ST.E here cost ~4x (measured, profiles/).
template <bool SCALE, bool DROP>
__device__ __forceinline__ void epi_chunk(uint32_t taddr, uint32_t scale_s, uint32_t bias_s, uint32_t row_s, uint32_t sw_xor,
                                          float floor_v, unsigned long long drop_base, uint32_t keep_thresh, float inv_keep) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t v[32];
    tmem_ld32_nowait(taddr + half * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      const int cl = half * 32 + j;
      float f[8];
      const float4 b0 = lds128f(bias_s + cl * 4), b1 = lds128f(bias_s + cl * 4 + 16);
      if constexpr (SCALE) {
        const float4 s0 = lds128f(scale_s + cl * 4), s1 = lds128f(scale_s + cl * 4 + 16);
        f[0] = fmaf(__uint_as_float(v[j + 0]), s0.x, b0.x); f[1] = fmaf(__uint_as_float(v[j + 1]), s0.y, b0.y);
        f[2] = fmaf(__uint_as_float(v[j + 2]), s0.z, b0.z); f[3] = fmaf(__uint_as_float(v[j + 3]), s0.w, b0.w);
        f[4] = fmaf(__uint_as_float(v[j + 4]), s1.x, b1.x); f[5] = fmaf(__uint_as_float(v[j + 5]), s1.y, b1.y);
        f[6] = fmaf(__uint_as_float(v[j + 6]), s1.z, b1.z); f[7] = fmaf(__uint_as_float(v[j + 7]), s1.w, b1.w);
      } else {
        f[0] = __uint_as_float(v[j + 0]) + b0.x; f[1] = __uint_as_float(v[j + 1]) + b0.y;
        f[2] = __uint_as_float(v[j + 2]) + b0.z; f[3] = __uint_as_float(v[j + 3]) + b0.w;
        f[4] = __uint_as_float(v[j + 4]) + b1.x; f[5] = __uint_as_float(v[j + 5]) + b1.y;
        f[6] = __uint_as_float(v[j + 6]) + b1.z; f[7] = __uint_as_float(v[j + 7]) + b1.w;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) f[t] = fmaxf(f[t], floor_v);  // ReLU: floor 0; no activation: floor -inf
      if constexpr (DROP) {
#pragma unroll
        for (int t = 0; t < 8; ++t) f[t] = (mix32(drop_base + cl + t) < keep_thresh) ? f[t] * inv_keep : 0.f;
      }
      const uint32_t q = (uint32_t)(cl >> 3);
      sts128(row_s + ((q ^ sw_xor) << 4), pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
             pack_bf16x2(f[6], f[7]));
    }
  }
}

// The same chunk without dropout (every conv layer, eval-mode fc6 / fc7): both 32-column TMEM loads are issued before the one
// wait, scale / bias run as packed fp32 pairs (FFMA2 / FADD2: same IEEE results per lane as fmaf / +, half the issue slots) and
// ReLU rides the conversion (cvt.rn.relu.bf16x2.f32 clamps the rounded value at +0: the value max(x, 0) rounds to).  Per
// chunk and lane: 2 LDTM + 16 (32) LDS.64x2 + 32 FADD2 (FFMA2) + 32 F2FP + 8 STS, against 64 FADD + 64 FMNMX + 32 F2FP before.
// The epilogue is a latency chain run by one or two warps per scheduler; the last tile's chain is the drain every launch pays.
__device__ __forceinline__ void lds_2x64(uint32_t saddr, uint64_t& a, uint64_t& b) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(saddr));
}
__device__ __forceinline__ uint64_t pair_u32(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <bool RELU>
__device__ __forceinline__ uint32_t pair_to_bf16x2(uint64_t v) {
  uint32_t lo, hi, r;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
  if constexpr (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
template <bool SCALE, bool RELU>
__device__ __forceinline__ void epi_half_packed(const uint32_t (&v)[32], int c0, uint32_t scale_s, uint32_t bias_s, uint32_t row_s,
                                                uint32_t sw_xor) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int cl = c0 + j;
    uint64_t b[4], f[4];
    lds_2x64(bias_s + cl * 4, b[0], b[1]);
    lds_2x64(bias_s + cl * 4 + 16, b[2], b[3]);
    if constexpr (SCALE) {
      uint64_t sc[4];
      lds_2x64(scale_s + cl * 4, sc[0], sc[1]);
      lds_2x64(scale_s + cl * 4 + 16, sc[2], sc[3]);
#pragma unroll
      for (int t = 0; t < 4; ++t) f[t] = fma_f32x2(pair_u32(v[j + 2 * t], v[j + 2 * t + 1]), sc[t], b[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) f[t] = add_f32x2(pair_u32(v[j + 2 * t], v[j + 2 * t + 1]), b[t]);
    }
    const uint32_t q = (uint32_t)(cl >> 3);
    sts128(row_s + ((q ^ sw_xor) << 4), pair_to_bf16x2<RELU>(f[0]), pair_to_bf16x2<RELU>(f[1]), pair_to_bf16x2<RELU>(f[2]),
           pair_to_bf16x2<RELU>(f[3]));
  }
}
template <bool SCALE, bool RELU>
__device__ __forceinline__ void epi_chunk_packed(uint32_t taddr, uint32_t scale_s, uint32_t bias_s, uint32_t row_s, uint32_t sw_xor) {
  uint32_t v0[32], v1[32];
  tmem_ld32_nowait(taddr, v0);
  tmem_ld32_nowait(taddr + 32, v1);
  tmem_ld_wait();
  epi_half_packed<SCALE, RELU>(v0, 0, scale_s, bias_s, row_s, sw_xor);
  epi_half_packed<SCALE, RELU>(v1, 32, scale_s, bias_s, row_s, sw_xor);
}

// CG = 1: one CTA computes a 128 x BN tile.  CG = 2: a CTA pair (2-CTA cluster, tcgen05 cta_group::2)
// computes a 256 x BN tile: each CTA stages its own 128 A rows and HALF of the B rows, the leader's MMA
// reads both halves (the peer's through the pair's shared-memory path), so the per-SM operand feed drops
// from 16 KB + BN*128 B to 16 KB + BN*64 B per k-block -- the SM<-L2 ingest limit (~64 B/clk/SM) is what
// bounds every GEMM here whose K loop is fed from L2.
// EW = epilogue warps (4 or 8).  A warp may only read the TMEM lanes of its quadrant (warp id % 4), so with 8 warps
// two warps share a quadrant and split the tile's 64-column chunks between them.  The epilogue of a chunk is a
// latency chain (tcgen05.ld -> math -> st.shared -> TMA store -> wait for the staging slot); short-K, wide-N layers
// (1x1 expansion convs: 8 k-blocks of MMA per 128 x 256 tile) are bound by it, and a second set of warps hides it.
template <int BN, int STAGES, int NBUF, int CG, int EW, bool HALO = false>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_r, const Params p) {
  static_assert(!HALO || CG == 1, "HALO mode is single-CTA");
  constexpr uint32_t BROWS = BN / CG;                       // B rows staged by this CTA
  constexpr uint32_t A_BYTES = HALO ? 0 : BM * BK * 2, B_BYTES = BROWS * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t HALO_BYTES = HALO ? HALO_SLOTS * HALO_SLOT_BYTES : 0;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  constexpr uint32_t EPI_BYTES = EW * NBUF * EPI_BUF_BYTES;
  constexpr int ETHREADS = 32 * EW, ESPLIT = EW / 4;  // epilogue threads; warps per TMEM quadrant
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // offset arithmetic keeps the shared address space
  uint8_t* halo_smem = smem + STAGES * STAGE_BYTES;                       // 1024-aligned (stage sizes are multiples of 1 KB)
  uint8_t* epi_smem = halo_smem + HALO_BYTES;
  uint8_t* ident = epi_smem + EPI_BYTES;                                  // 8 KB, 1024-aligned
  float* sb_smem = reinterpret_cast<float*>(ident + (HALO ? 0 : IDENT_BYTES));  // [2 acc stages][scale BN | bias BN]; HALO has no shortcut, no identity
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sb_smem + 4 * BN);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* hfull_bar = tempty_bar + 2;   // HALO: input rows of a channel block staged / consumed
  uint64_t* hempty_bar = hfull_bar + HALO_SLOTS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty_bar + HALO_SLOTS);

  // warp-uniform by construction, and shuffled from lane 0 so that ptxas KNOWS it: branches on them become uniform branches
  // and the role loops below run on the uniform datapath (no divergence bookkeeping, operands in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int cta_rank = (CG == 2) ? __shfl_sync(0xffffffffu, (int)cluster_ctarank(), 0) : 0;
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG, num_units = gridDim.x / CG;   // a unit = one CTA (CG=1) or one CTA pair
  const int num_mp = (p.num_m_tiles + CG - 1) / CG;               // M tiles per unit step (pairs for CG=2)
  const int num_tiles = num_mp * p.num_n_tiles;
  if (p.tl_slot >= 0 && threadIdx.x == 0) atomicMin(&g_timeline[p.tl_slot][0], gtime());
  const bool tr_on = (p.debug & 2048) && blockIdx.x == 0;
  if (threadIdx.x == 0) trace(tr_on, 0, 7);

  if (warp == 0) {
    // the barriers are one contiguous array (full | empty | tmem-full | tmem-empty | halo-full | halo-empty): lane i initialises
    // barrier i (one lane doing all ~20-26 of them in turn sat on every launch's critical path)
    constexpr int NBAR = 2 * STAGES + 4 + 2 * HALO_SLOTS;
    static_assert(NBAR <= 32, "one barrier per lane");
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_b) : "memory");
      if (!p.out_f32) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_o) : "memory");
      if (p.has_residual) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_r) : "memory");
    }
    if (lane < NBAR) {
      const bool is_tempty = lane >= 2 * STAGES + 2 && lane < 2 * STAGES + 4;
      mbar_init(&full_bar[lane], is_tempty ? (uint32_t)(EW * CG) : 1u);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.has_residual && threadIdx.x >= 64 && threadIdx.x < 64 + 64 / CG) {
    // (this CTA's rows of) the 64x64 identity in the canonical K-major SWIZZLE_128B layout: local row j at
    // j*128, 16-byte chunk q at (q ^ (j & 7)); row j is output column n = rank*32 + j (CG=2) or j (CG=1)
    const int j = threadIdx.x - 64;
    const int n = cta_rank * (64 / CG) + j;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (q == (n >> 3)) {
        const uint32_t one = 0x3F80u << (16 * (n & 1));  // bf16 1.0 at element n & 7 of the chunk
        const int wsel = (n & 7) >> 1;
        v.x = wsel == 0 ? one : 0u; v.y = wsel == 1 ? one : 0u; v.z = wsel == 2 ? one : 0u; v.w = wsel == 3 ? one : 0u;
      }
      *reinterpret_cast<uint4*>(ident + j * 128 + ((q ^ (j & 7)) << 4)) = v;
    }
    fence_async_smem();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / 2-SM TMA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Programmatic dependent launch: everything above (barrier init, TMEM alloc, descriptor prefetch) may run
  // while the previous kernel of the stream is still draining; global memory is only touched after this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (p.tl_slot >= 0 && threadIdx.x == 0) atomicMin(&g_timeline[p.tl_slot][1], gtime());
  if (threadIdx.x == 0) trace(tr_on, 1, 7);  // prologue done (row 1, column 7)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (every CTA): all 32 lanes walk the
    // schedule in warp-uniform control flow, one elected lane issues (see elect_one)
    {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = p.conv ? (p.Cin / BK) : 1;
      Sched sched(unit, num_units, num_tiles, p);
      Piece pc;
      int hs = 0;
      uint32_t hphase = 0;
      bool b_loaded = false;
      int tr_k = -1;
      while (sched.next(pc)) {
        if (tr_k >= 0) trace(tr_on, tr_k, 1);  // the previous tile's loads are all issued
        ++tr_k;
        trace(tr_on, tr_k, 0);
        const int tile = pc.tile;
        const int mt = (tile % num_mp) * CG + cta_rank, nt = tile / num_mp;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int per_img = p.tiles_h * p.tiles_w;
          img = mt / per_img;  // a padding tile of an odd pair lands at img == NB: all loads zero-fill, all stores clip
          const int r = mt - img * per_img;
          h0 = (r / p.tiles_w) * p.tile_h;
          w0 = (r % p.tiles_w) * p.tile_w;
        }
        if constexpr (HALO) {
          // per channel block: three input rows into a halo slot, then the nine taps' weight tiles through the ring
          const uint32_t row_load = (uint32_t)(BM + 2 * p.dil) * 128u;
          if (p.bres && !b_loaded) {  // weights of all nine taps: one slot and one barrier each, loaded once per CTA
            b_loaded = true;
            if (elect_one()) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                mbar_arrive_expect_tx(&full_bar[tap], B_BYTES);
                tma_load_2d(&map_b, &full_bar[tap], smem + tap * STAGE_BYTES, tap * BK, nt * BN);
              }
            }
            __syncwarp();
          }
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&hempty_bar[hs], hphase ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&hfull_bar[hs], 3 * row_load);
              uint8_t* slot = halo_smem + hs * HALO_SLOT_BYTES;
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
                tma_load_4d(&map_a, &hfull_bar[hs], slot + dy * HALO_ROW_BYTES, cb * BK, w0 - p.dil, h0 + (dy - 1) * p.dil, img);
            }
            __syncwarp();
            if (++hs == HALO_SLOTS) { hs = 0; hphase ^= 1; }
            if (p.bres) continue;
            int kcol = cb * BK;                      // weight K order is (tap, channel): tap t of this block at (t * cblocks + cb) * 64
            const int kstep = cblocks * BK;
            for (int tap = 0; tap < 9; ++tap) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[stage], B_BYTES);
                tma_load_2d(&map_b, &full_bar[stage], smem + stage * STAGE_BYTES, kcol, nt * BN);
              }
              __syncwarp();
              kcol += kstep;
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        // The issuing warps are single instruction streams: every instruction in this loop is paid once per k-block by the
        // whole pipeline (measured: ~100 SASS instructions with an integer division per k-block held every conv layer at
        // ~0.3 us per k-block after the MMA loop had been fixed), so the (tap, channel block) walk is kept in counters.
        const int nkb = pc.kb1 - pc.kb0;
        const int brow = nt * BN + cta_rank * (int)BROWS;
        const bool skip_a = (p.debug & 32) != 0;
        const int a_row = mt * BM;
        int kcol = pc.kb0 * BK;                      // K offset of the weight tile (and of A in GEMM mode)
        int cb = 0, tx = 0, ty = 0;                  // conv mode: channel block, filter tap (tx, ty)
        if (p.conv && pc.kb0 != 0) { const int tap0 = pc.kb0 / cblocks; cb = pc.kb0 - tap0 * cblocks; ty = tap0 / 3; tx = tap0 - 3 * ty; }
        int dw = (tx - 1) * p.dil, dh = (ty - 1) * p.dil;
        for (int ki = 0; ki < nkb; ++ki) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * (skip_a ? B_BYTES : STAGE_BYTES));
            uint8_t* sa = smem + stage * STAGE_BYTES;
            uint8_t* sb = sa + A_BYTES;
            if constexpr (CG == 2) {
              if (!skip_a) {
                if (p.conv) tma2_load_4d(&map_a, &full_bar[stage], sa, cb * BK, w0 + dw, h0 + dh, img);
                else tma2_load_2d(&map_a, &full_bar[stage], sa, kcol, a_row);
              }
              tma2_load_2d(&map_b, &full_bar[stage], sb, kcol, brow);
            } else {
              if (!skip_a) {
                if (p.conv) tma_load_4d(&map_a, &full_bar[stage], sa, cb * BK, w0 + dw, h0 + dh, img);
                else tma_load_2d(&map_a, &full_bar[stage], sa, kcol, a_row);
              }
              tma_load_2d(&map_b, &full_bar[stage], sb, kcol, brow);
            }
          }
          __syncwarp();
          kcol += BK;
          if (++cb == cblocks) {                     // next filter tap (GEMM mode: cblocks == 1, the counters are unused)
            cb = 0;
            if (++tx == 3) { tx = 0; ++ty; dh += p.dil; }
            dw = (tx - 1) * p.dil;
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.has_residual && pc.kind != 1) {
          const int nchunk = (min(BN, p.N - nt * BN) + EPI_CHUNK - 1) / EPI_CHUNK;
          for (int c = 0; c < nchunk; ++c) {  // shortcut tile as extra "k-blocks": 128 rows x 64 output columns each
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * A_BYTES);
              uint8_t* sa = smem + stage * STAGE_BYTES;
              if constexpr (CG == 2) {
                if (p.conv) tma2_load_4d(&map_r, &full_bar[stage], sa, nt * BN + c * EPI_CHUNK, w0, h0, img);
                else tma2_load_2d(&map_r, &full_bar[stage], sa, nt * BN + c * EPI_CHUNK, mt * BM);
              } else {
                if (p.conv) tma_load_4d(&map_r, &full_bar[stage], sa, nt * BN + c * EPI_CHUNK, w0, h0, img);
                else tma_load_2d(&map_r, &full_bar[stage], sa, nt * BN + c * EPI_CHUNK, mt * BM);
              }
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (tr_k >= 0) trace(tr_on, tr_k, 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA): the whole warp walks the
    // schedule, one elected lane issues the tcgen05.mma / tcgen05.commit instructions
    if (leader) {
      constexpr uint32_t idesc = make_idesc(BN, BM * CG);
      constexpr uint32_t idesc64 = make_idesc(64, BM * CG);
      const uint32_t ident_lo = smem_desc_lo(smem_u32(ident));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int hs = 0;
      uint32_t hphase = 0;
      bool b_ready = false;
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full_base = smem_u32(full_bar);
      constexpr uint32_t EMPTY_OFF = STAGES * 8;  // empty_bar = full_bar + STAGES
      Sched sched(unit, num_units, num_tiles, p);
      Piece pc;
      int tr_k = -1;
      while (sched.next(pc)) {
        const int tile = pc.tile;
        const bool with_res = p.has_residual && pc.kind != 1;
        ++tr_k;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        trace(tr_on, tr_k, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * BN;
        if constexpr (HALO) {
          const int cblocks = p.Cin / BK;
          const uint32_t dx_bytes = (uint32_t)p.dil * 128u;
          if (p.bres && !(p.debug & 1024)) {
            // Resident weights (Cin = 64: one channel block, tap ti in ring slot ti for the whole kernel): the nine weight barriers
            // are waited for once per CTA, and a tile is ONE elected block of 9 x 4 MMAs + two commits.  The general loop below
            // spends ~55 issue-warp instructions per tap (barrier try_wait, election, descriptor rebuild, slot bookkeeping) for
            // 128 clocks of N = 64 tensor work (stem conv2 / conv3 at 300 x 500: 18.8 -> 16.9 us).  What is left
            // (tools/tile_trace.py, profiles/r2_tile_trace.txt): the 36 MMAs of a tile take ~2080 clocks = 58 per N = 64 MMA
            // against 32 of tensor work -- every MMA reads its 128 x 16 A slice (4 KB) + 64 x 16 B slice (2 KB) from shared
            // memory, ~48 clocks at 128 B/clk: narrow-N layers are bound by the operand read, not by issue or the tensor pipe.
            if (!b_ready) {
              b_ready = true;
#pragma unroll 1
              for (int ti = 0; ti < 9; ++ti) mbar_wait(&full_bar[ti], 0u);
            }
            mbar_wait(&hfull_bar[hs], hphase);
            trace(tr_on, tr_k, 3);
            if (p.tl_slot >= 0 && lane == 0) atomicMin(&g_timeline[p.tl_slot][2], gtime());
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t slot = smem_base + (uint32_t)(halo_smem - smem) + hs * HALO_SLOT_BYTES;
            if (elect_one()) {
#pragma unroll
              for (int ti = 0; ti < 9; ++ti) {
                const uint32_t a_start = slot + (uint32_t)(ti / 3) * HALO_ROW_BYTES + (uint32_t)(ti % 3) * dx_bytes;
                const uint32_t a_hi = p.halo_bo ? (DESC_HI | (((a_start >> 7) & 7u) << 17)) : DESC_HI;
                umma_kblock<CG>(tmem_d, smem_desc_lo(a_start), smem_desc_lo(smem_base + ti * STAGE_BYTES), a_hi, idesc, ti > 0 ? 1u : 0u);
              }
              umma_commit_addr<CG>(smem_u32(&hempty_bar[hs]));  // all nine taps have read the rows
              umma_commit_addr<CG>(smem_u32(&tfull_bar[acc]));
            }
            __syncwarp();
            trace(tr_on, tr_k, 4);
            if (++hs == HALO_SLOTS) { hs = 0; hphase ^= 1; }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            continue;
          }
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&hfull_bar[hs], hphase);
            if (cb == 0) trace(tr_on, tr_k, 3);
            if (p.tl_slot >= 0 && cb == 0 && lane == 0) atomicMin(&g_timeline[p.tl_slot][2], gtime());
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t slot = smem_base + (uint32_t)(halo_smem - smem) + hs * HALO_SLOT_BYTES;
            uint32_t a_start = slot;
            int tx = 0;
            for (int ti = 0; ti < 9; ++ti) {
              // resident weights: tap ti lives in slot ti for the whole kernel (its barrier completed once, phase 0)
              const int bslot = p.bres ? ti : stage;
              mbar_wait(&full_bar[bslot], p.bres ? 0u : phase);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              // output pixel m of the row tile reads input pixel m + (dx + 1) * d of input row dy: shifted start, SBO stays 1024
              const uint32_t adesc = smem_desc_lo(a_start);
              const uint32_t a_hi = p.halo_bo ? (DESC_HI | (((a_start >> 7) & 7u) << 17)) : DESC_HI;  // base-offset field [49,52)
              const uint32_t bdesc = smem_desc_lo(smem_base + bslot * STAGE_BYTES);
              if (elect_one()) {
                umma_kblock<CG>(tmem_d, adesc, bdesc, a_hi, idesc, (cb > 0 || ti > 0) ? 1u : 0u);
                if (!p.bres) umma_commit_addr<CG>(smem_u32(&empty_bar[stage]));  // frees the weight slot once these MMAs retire
                if (ti == 8) {
                  umma_commit_addr<CG>(smem_u32(&hempty_bar[hs]));  // all nine taps have read the rows
                  if (cb == cblocks - 1) umma_commit_addr<CG>(smem_u32(&tfull_bar[acc]));
                }
              }
              __syncwarp();
              a_start += dx_bytes;
              if (++tx == 3) { tx = 0; a_start += HALO_ROW_BYTES - 3 * dx_bytes; }  // next input row
              if (!p.bres && ++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++hs == HALO_SLOTS) { hs = 0; hphase ^= 1; }
          }
          trace(tr_on, tr_k, 4);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
        // One k-block per iteration; this warp is a single instruction stream, so every instruction here is paid per k-block
        // by the whole CTA: running descriptor / barrier addresses (no per-iteration multiplies), one asm block for the four
        // MMAs, the accumulator hand-over peeled out of the loop.
        const int nkb = pc.kb1 - pc.kb0;
        constexpr uint32_t STAGE_DESC = STAGE_BYTES >> 4, B_DESC_OFF = A_BYTES >> 4;
        uint32_t a_lo = smem_desc_lo(smem_base + stage * STAGE_BYTES);
        uint32_t full_addr = full_base + stage * 8;
        uint32_t accum = 0;
        for (int ki = 0; ki < nkb; ++ki) {
          mbar_wait_addr(full_addr, phase);
          if (ki == 0) trace(tr_on, tr_k, 3);
          if (p.tl_slot >= 0 && ki == 0 && lane == 0) atomicMin(&g_timeline[p.tl_slot][2], gtime());
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            umma_kblock<CG>(tmem_d, a_lo, a_lo + B_DESC_OFF, DESC_HI, idesc, accum);
            umma_commit_addr<CG>(full_addr + EMPTY_OFF);  // frees the smem slot (in both CTAs) once these MMAs retire
          }
          __syncwarp();
          accum = 1;
          a_lo += STAGE_DESC;
          full_addr += 8;
          if (++stage == STAGES) { stage = 0; phase ^= 1; a_lo -= STAGES * STAGE_DESC; full_addr -= STAGES * 8; }
        }
        if (!with_res) {
          if (elect_one()) umma_commit_addr<CG>(smem_u32(&tfull_bar[acc]));
          __syncwarp();
        }
        if (with_res) {
          const int nt = tile / num_mp;
          const int nchunk = (min(BN, p.N - nt * BN) + EPI_CHUNK - 1) / EPI_CHUNK;
          for (int c = 0; c < nchunk; ++c) {
            mbar_wait(&full_bar[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t adesc = smem_desc_lo(smem_base + stage * STAGE_BYTES);
            if (elect_one()) {
              umma_kblock<CG>(tmem_d + c * EPI_CHUNK, adesc, ident_lo, DESC_HI, idesc64, 1u);
              umma_commit_addr<CG>(smem_u32(&empty_bar[stage]));
              if (c == nchunk - 1) umma_commit_addr<CG>(smem_u32(&tfull_bar[acc]));
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        trace(tr_on, tr_k, 4);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5, every CTA: its own 128 rows)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const int etid = (warp - 2) * 32 + lane;  // 0..ETHREADS-1 within the epilogue group
    const int half = (warp - 2) >> 2;          // which share of the column chunks this warp takes (0 when EW == 4)
    uint8_t* my_bufs = epi_smem + (warp - 2) * (NBUF * EPI_BUF_BYTES);
    const uint32_t sw_xor = (uint32_t)(lane & 7);
    // warp's 32-row sub-box inside a conv tile
    const int sub_h = (quad * 32) / (p.conv ? p.tile_w : 32), sub_w = (quad * 32) % (p.conv ? p.tile_w : 32);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t gchunk = 0;  // running chunk counter of this warp -> staging ring slot (one bulk group per chunk)
    const unsigned long long drop_seed = p.drop_seed + ((p.drop_keep_thresh && p.drop_seed_dev) ? __ldg(p.drop_seed_dev) : 0ull);
    constexpr int NCHUNK = (BN + EPI_CHUNK - 1) / EPI_CHUNK;
    const float relu_floor = p.relu ? 0.f : -INFINITY;
    float pre_sc[2] = {1.f, 1.f}, pre_bi[2] = {0.f, 0.f};
    auto fetch_sb = [&](int nt_) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = etid + ETHREADS * i, n = nt_ * BN + idx;
        const bool ok = idx < BN && n < p.N;
        pre_sc[i] = (ok && p.scale) ? __ldg(p.scale + n) : 1.f;
        pre_bi[i] = (ok && p.bias) ? __ldg(p.bias + n) : 0.f;
      }
    };
    Sched sched(unit, num_units, num_tiles, p);
    Piece pc, pc_next;
    bool have = sched.next(pc);
    if (have) fetch_sb(pc.tile / num_mp);
    int tr_e = 0;

    while (have) {
      const bool have_next = sched.next(pc_next);
      const int tile = pc.tile;
      const int mt = (tile % num_mp) * CG + cta_rank, nt = tile / num_mp;
      int img = 0, hh0 = 0, ww0 = 0;
      long long m_global;
      bool row_ok;
      if (p.conv) {
        const int per_img = p.tiles_h * p.tiles_w;
        img = mt / per_img;
        const int r = mt - img * per_img;
        hh0 = (r / p.tiles_w) * p.tile_h;
        ww0 = (r % p.tiles_w) * p.tile_w;
        const int hh = hh0 + row / p.tile_w, ww = ww0 + row % p.tile_w;
        row_ok = img < p.NB && hh < p.H && ww < p.W;
        m_global = ((long long)img * p.H + hh) * p.W + ww;
      } else {
        m_global = (long long)mt * BM + row;
        row_ok = m_global < p.M;
      }
      const int ncols = min(BN, p.N - nt * BN);            // valid columns of this tile (multiple of 8)
      const int nchunk = (ncols + EPI_CHUNK - 1) / EPI_CHUNK;

      // per-tile scale / bias into smem (broadcast reads later); the values were fetched one tile ahead
      float* s_scale = sb_smem + acc * 2 * BN;
      float* s_bias = s_scale + BN;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = etid + ETHREADS * i;
        if (idx < BN) { s_scale[idx] = pre_sc[i]; s_bias[idx] = pre_bi[i]; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
      if (have_next) fetch_sb(pc_next.tile / num_mp);  // in flight while this tile is drained

      mbar_wait(&tfull_bar[acc], acc_phase);
      if (warp == 2) trace(tr_on, tr_e, 5);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;

      if (pc.kind == 1) {
        // ---------------------------------------------------------- stream-K tail part: dump the raw fp32 partial
        // layout [column][128 rows] so that a warp's 32 lanes (rows) write / read 128 contiguous bytes
        float* wsp = p.sk_ws + (size_t)(pc.slot * CG + cta_rank) * (BN * 128) + row;
#pragma unroll 1
        for (int c = 32 * half; c < BN; c += 32 * ESPLIT) {
          uint32_t v[32];
          tmem_ld32_nowait(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(wsp + (size_t)(c + j) * 128, __uint_as_float(v[j]));
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
        if (etid == 0) {
          unsigned int* flag = p.sk_flags + pc.slot * CG + cta_rank;
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
        }
      } else if (pc.kind == 2) {
        // ---------------------------------------------------------- owner: fold in the other parts' partials (they
        // were scheduled earlier, or run right now on another, resident CTA), in slot order, then fall through to the
        // normal epilogue
#pragma unroll 1
        for (int s = 0; s < pc.nslots; ++s) {
          unsigned int* flag = p.sk_flags + (pc.slot + s) * CG + cta_rank;
          if (etid == 0) {
            unsigned int f;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
            } while (f == 0u);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
          const float* wsp = p.sk_ws + (size_t)((pc.slot + s) * CG + cta_rank) * (BN * 128) + row;
#pragma unroll 1
          for (int c = 32 * half; c < BN; c += 32 * ESPLIT) {
            uint32_t v[32];
            tmem_ld32_nowait(taddr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(wsp + (size_t)(c + j) * 128));
            tmem_st32(taddr + c, v);
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ETHREADS) : "memory");
          if (etid == 0) *flag = 0u;  // self-resetting: ready for the next launch
        }
      }

      if (pc.kind == 1) {
        // nothing to store
      } else if (!p.out_f32) {
        // ---------------------------------------------------------- bf16 output through smem + TMA store
        uint32_t mine = 0;
#pragma unroll 1
        for (int c = half; c < NCHUNK; c += ESPLIT) {
          if (c >= nchunk) break;
          const uint32_t g = gchunk + mine;
          ++mine;
          const int b = g % NBUF;
          uint8_t* buf = my_bufs + b * EPI_BUF_BYTES;
          if (lane == 0) bulk_wait_read<NBUF - 1>();  // the store that last used this slot has drained
          __syncwarp();
          if (!(p.debug & 2)) {
            const uint32_t row_s = smem_u32(buf) + lane * 128;
            const uint32_t ta = taddr + c * EPI_CHUNK;
            const uint32_t sc_s = smem_u32(s_scale) + c * EPI_CHUNK * 4, bi_s = smem_u32(s_bias) + c * EPI_CHUNK * 4;
            const unsigned long long dbase = drop_seed * 0x100000001B3ull + (unsigned long long)m_global * (unsigned long long)p.N +
                                             (unsigned long long)(nt * BN + c * EPI_CHUNK);
            if (p.drop_keep_thresh) {
              if (p.scale) epi_chunk<true, true>(ta, sc_s, bi_s, row_s, sw_xor, relu_floor, dbase, p.drop_keep_thresh, p.drop_inv_keep);
              else epi_chunk<false, true>(ta, sc_s, bi_s, row_s, sw_xor, relu_floor, dbase, p.drop_keep_thresh, p.drop_inv_keep);
            } else if (p.debug & 512) {  // measurement switch: the scalar epilogue math
              if (p.scale) epi_chunk<true, false>(ta, sc_s, bi_s, row_s, sw_xor, relu_floor, 0ull, 0u, 1.f);
              else epi_chunk<false, false>(ta, sc_s, bi_s, row_s, sw_xor, relu_floor, 0ull, 0u, 1.f);
            } else if (p.relu) {
              if (p.scale) epi_chunk_packed<true, true>(ta, sc_s, bi_s, row_s, sw_xor);
              else epi_chunk_packed<false, true>(ta, sc_s, bi_s, row_s, sw_xor);
            } else {
              if (p.scale) epi_chunk_packed<true, false>(ta, sc_s, bi_s, row_s, sw_xor);
              else epi_chunk_packed<false, false>(ta, sc_s, bi_s, row_s, sw_xor);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && !(p.debug & 1)) {
            if (p.conv)
              tma_store_4d(&map_o, buf, nt * BN + c * EPI_CHUNK, ww0 + sub_w, hh0 + sub_h, img);
            else
              tma_store_2d(&map_o, buf, nt * BN + c * EPI_CHUNK, mt * BM + quad * 32);
            bulk_commit();
          }
          __syncwarp();
        }
        gchunk += mine;
      } else {
        // ---------------------------------------------------------- fp32 output (head logits, fp32_tc partials)
        // Each 32 x 32 block goes through this warp's staging slot (unused on this path: no TMA stores) so that one store
        // instruction writes 4 rows x 128 contiguous bytes; the first version stored 32 rows x 16 B per instruction.
        const uint32_t stage_s = smem_u32(my_bufs);
        const int q4 = lane & 7, rsub = lane >> 3;
        float* obase = reinterpret_cast<float*>(p.out);
        if (p.n_peers > 0) {  // the whole 128-row tile has one owner (rows_per_peer % 128 == 0)
          const int owner = (mt * BM) / p.rows_per_peer;
          obase = p.peer_out[owner] + ((long long)p.my_rank - owner) * (long long)p.rows_per_peer * p.ldo;
        }
        long long mrow[8];  // global row of staging row rr = 4 i + rsub (owned by lane rr), -1 when out of range
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + rsub;
          const unsigned lo = __shfl_sync(0xffffffffu, (unsigned)(m_global & 0xffffffffll), rr);
          const unsigned hi = __shfl_sync(0xffffffffu, (unsigned)((unsigned long long)m_global >> 32), rr);
          const int ok = __shfl_sync(0xffffffffu, (int)row_ok, rr);
          mrow[i] = ok ? (long long)(((unsigned long long)hi << 32) | lo) : -1ll;
        }
#pragma unroll 1
        for (int c = 32 * half; c < BN; c += 32 * ESPLIT) {
          uint32_t v[32];
          tmem_ld32_nowait(taddr + c, v);
          tmem_ld_wait();
          const int n0 = nt * BN + c;
          if (n0 < p.N) {  // warp-uniform
            const int nvalid = min(32, p.N - n0);  // multiple of 8 (host checks N % 8 == 0)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o;
              o.x = fmaf(__uint_as_float(v[j + 0]), s_scale[c + j + 0], s_bias[c + j + 0]);
              o.y = fmaf(__uint_as_float(v[j + 1]), s_scale[c + j + 1], s_bias[c + j + 1]);
              o.z = fmaf(__uint_as_float(v[j + 2]), s_scale[c + j + 2], s_bias[c + j + 2]);
              o.w = fmaf(__uint_as_float(v[j + 3]), s_scale[c + j + 3], s_bias[c + j + 3]);
              if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              sts128(stage_s + lane * 128 + ((((uint32_t)j >> 2) ^ sw_xor) << 4), __float_as_uint(o.x), __float_as_uint(o.y),
                     __float_as_uint(o.z), __float_as_uint(o.w));
            }
            __syncwarp();
            if (4 * q4 < nvalid) {
              float* ocol = obase + n0 + 4 * q4;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                if (mrow[i] >= 0) {
                  const int rr = 4 * i + rsub;
                  const float4 val = lds128f(stage_s + rr * 128 + (((uint32_t)q4 ^ (uint32_t)(rr & 7)) << 4));
                  *reinterpret_cast<float4*>(ocol + mrow[i] * p.ldo) = val;
                }
              }
            }
            __syncwarp();  // the slot is rewritten by the next column block
          }
        }
      }
      if (warp == 2) trace(tr_on, tr_e, 6);
      ++tr_e;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && !leader) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      pc = pc_next;
      have = have_next;
    }
    if (lane == 0) bulk_wait_read<0>();  // staging smem must outlive the last TMA store's read (global visibility comes with grid completion)
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.tl_slot >= 0 && threadIdx.x == 0) atomicMax(&g_timeline[p.tl_slot][3], gtime());
  if constexpr (CG == 2) cluster_sync_all();  // the leader's MMAs read the peer's smem: nobody leaves early
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else
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

static int g_tl_next = -1;  // next timeline slot (drn_gemm_timeline_reset arms it)
static int g_max_sms = 0;  // drn_gemm_set_max_sms: leave SMs to a concurrently running kernel (NCCL)
static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return (g_max_sms > 0 && g_max_sms < n) ? g_max_sms : n;
}

constexpr size_t SK_FLAG_BYTES = 4096;
constexpr size_t SK_WS_BYTES = SK_FLAG_BYTES + (size_t)148 * 128 * 256 * sizeof(float);  // stream-K: one 128 x 256 fp32 partial per SM
constexpr size_t SK_WS_BYTES_MAX = SK_FLAG_BYTES + (size_t)4 * 148 * 128 * 256 * sizeof(float);  // tail split-K: up to 4 per SM
static int g_tail_split = 0;  // opt-in (drn_gemm_set_tail_split / DRN_TC_TAILSPLIT=1): measured neutral-to-negative, see launch()

template <int BN, int STAGES, int NBUF, int CG, int EW = 4, bool HALO = false>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo, const CUtensorMap& mr, Params p,
                  cudaStream_t st, void* workspace, size_t workspace_bytes) {
  constexpr size_t smem = (size_t)STAGES * ((HALO ? 0 : BM * BK * 2) + (BN / CG) * BK * 2) + (HALO ? HALO_SLOTS * HALO_SLOT_BYTES : 0) +
                          EW * NBUF * EPI_BUF_BYTES + (HALO ? 0 : IDENT_BYTES) + 4 * BN * sizeof(float) +
                          (2 * STAGES + 4 + 2 * HALO_SLOTS) * sizeof(uint64_t) + 16 + 1024;
  static_assert(smem <= 232448, "gemm_tc: shared memory budget exceeded");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, NBUF, CG, EW, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_err("gemm_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    configured = true;
  }
  const int units = ((p.num_m_tiles + CG - 1) / CG) * p.num_n_tiles;
  const int max_units = num_sms() / CG;
  // stream-K (opt-in, DRN_TC_STREAMK=1) when whole-tile scheduling would leave > 6 % of the machine idle in the last
  // wave and K is deep enough that every unit's share spans at least one whole tile (see Sched).  Measured on
  // fc6 (M=4000, N=2048, K=100352: 1.73 waves): 1836 us vs 1212 us for whole-tile waves -- units then sit at
  // different K offsets, so the A/B tiles shared by co-scheduled CTAs no longer meet in L2 and every CTA streams
  // its operands from HBM (12.8 GB instead of ~2 GB).  Kept for shapes whose operands fit in L2; off by default.
  p.stream_k = 0;
  p.sk_tail0 = p.sk_tail = p.sk_parts = 0;
  if (workspace && workspace_bytes >= SK_WS_BYTES && !p.out_f32 && units > max_units && max_units <= 148 && p.KB >= 8) {
    const int waves = (units + max_units - 1) / max_units;
    const double ideal = (double)units / max_units;
    const long long share = ((long long)units * p.KB + max_units - 1) / max_units;
    static int sk_env = -1;
    if (sk_env < 0) {
      const char* e = getenv("DRN_TC_STREAMK");
      sk_env = (e && e[0] == '1') ? 1 : 0;
    }
    if (sk_env && waves / ideal > 1.06 && share >= p.KB) {
      p.stream_k = 1;
      p.sk_flags = (unsigned int*)workspace;
      p.sk_ws = (float*)((char*)workspace + SK_FLAG_BYTES);
    }
  }
  // tail split-K (opt-in): cut the tiles of the partial last wave into `parts` K-ranges when that shortens the wave
  // by >= 15 % -- fc6 (M=4000, N=2048, K=100352) is 128 pair-tiles on 74 pair slots: 1 whole wave + 54 tiles x 4
  // quarters = 216 quarter-units = 3 rounds of 1/4 wave, 1.75 waves instead of 2.  Measured: 1164 us vs 1153-1193 us
  // alone, +0.14 ms inside the step (partials add 85 MB of traffic).  The kernel is POWER bound, not wave bound: it
  // already averages 1430 TFLOP/s over both waves (cuBLAS sustained: 1362), the 40 SMs idle in wave 2 hand their
  // power budget to the busy ones (SM clock rises), so evening out the waves buys nothing
  // (profiles/r1_tailsplit_negative_result.txt).
  static int ts_env = -1;
  if (ts_env < 0) {
    const char* e = getenv("DRN_TC_TAILSPLIT");
    ts_env = (e && e[0] == '1') ? 1 : 0;
    if (ts_env) g_tail_split = 1;
  }
  if (!p.stream_k && g_tail_split && workspace && !p.out_f32 && p.KB >= 256) {
    const int full = (units / max_units) * max_units, tail = units - full;
    if (tail > 0 && full > 0) {
      int best_parts = 1;
      double best = 1.0;
      for (int parts = 2; parts <= 8; ++parts) {
        if (p.KB / parts < 64) break;
        const double t = (double)((tail * parts + max_units - 1) / max_units) / parts + 0.02 * parts;  // + fixup cost
        if (t < best - 1e-9) { best = t; best_parts = parts; }
      }
      const size_t need = SK_FLAG_BYTES + (size_t)tail * (best_parts - 1) * CG * BN * 128 * sizeof(float);
      if (best_parts > 1 && best <= 0.85 && (size_t)tail * (best_parts - 1) * CG * sizeof(unsigned int) <= SK_FLAG_BYTES &&
          need <= workspace_bytes) {
        p.stream_k = 2;
        p.sk_tail0 = full;
        p.sk_tail = tail;
        p.sk_parts = best_parts;
        p.sk_flags = (unsigned int*)workspace;
        p.sk_ws = (float*)((char*)workspace + SK_FLAG_BYTES);
      }
    }
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(CG * (units < max_units ? units : max_units)));
  cfg.blockDim = dim3(64 + 32 * EW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int pdl_env = -1;
  if (pdl_env < 0) {
    const char* e = getenv("DRN_TC_PDL");
    pdl_env = (e && e[0] == '0') ? 0 : 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = pdl_env ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, NBUF, CG, EW, HALO>, ma, mb, mo, mr, p);
  if (e != cudaSuccess) return set_err("gemm_tc launch: %s", cudaGetErrorString(e));
  return 0;
}

// tile-shape choice: minimise waves x per-tile time, per k-block max(MMA cycles, operand bytes / 64 B/clk SM<-L2 ingest)
static int pick_bn(int m_tiles, int N, int KB, int cg) {
  const int units_max = num_sms() / cg;
  int best = 64;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N < bn) continue;
    const int units = ((m_tiles + cg - 1) / cg) * ((N + bn - 1) / bn);
    const int waves = (units + units_max - 1) / units_max;
    const double mma = 2.0 * bn, feed = (16384.0 + bn * 128.0 / cg) / 64.0;
    const double per_tile = KB * (mma > feed ? mma : feed) + 6.0 * bn;
    const double cost = waves * per_tile + 1500.0;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// HALO mode tile width: per 64-channel block a tile moves 3 (128 + 2d) x 128 B of input rows + 9 BN x 128 B of weights
// into the SM (~64 B/clk) and issues 9 x 4 MMAs of BN / 2 clocks
static int pick_bn_halo(int m_tiles, int N, int cblocks, int dil) {
  const int units_max = num_sms();
  int best = 64;
  double best_cost = 1e30;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && N < bn) continue;
    const int units = m_tiles * ((N + bn - 1) / bn);
    const int waves = (units + units_max - 1) / units_max;
    const double mma = 18.0 * bn, feed = (3.0 * (BM + 2 * dil) * 128.0 + 9.0 * bn * 128.0) / 64.0;
    const double cost = waves * (cblocks * (mma > feed ? mma : feed) + 6.0 * bn) + 1500.0;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// conv-mode spatial tile: tile_w x tile_h = 128 output pixels; fewest tiles wins, squarer wins ties
static void pick_conv_tile(int H, int W, int* tw_out, int* th_out) {
  const int cand[4] = {16, 32, 64, 128};
  int best = 16;
  long best_tiles = -1;
  for (int i = 0; i < 4; ++i) {
    const int tw = cand[i], th = 128 / tw;
    const long tiles = (long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    if (best_tiles < 0 || tiles < best_tiles) { best_tiles = tiles; best = tw; }
  }
  *tw_out = best;
  *th_out = 128 / best;
}

}  // namespace tc
}  // namespace drn

using namespace drn;

extern "C" size_t drn_gemm_workspace_bytes(void) { return drn::tc::SK_WS_BYTES_MAX; }
extern "C" int drn_gemm_set_max_sms(int max_sms) {
  const int prev = drn::tc::g_max_sms;
  drn::tc::g_max_sms = max_sms > 0 ? (max_sms & ~1) : 0;  // even: CTA pairs
  return prev;
}
// Profiling aid (DRN_TC_DEBUG & 256, tools/timeline_probe.py; not part of the drop-in ABI): arm / read the per-launch
// device timeline of the GEMM kernel.
extern "C" int drn_gemm_timeline_reset(drn_stream_t stream) {
  using namespace drn::tc;
  static unsigned long long init[TL_SLOTS][4];
  for (int i = 0; i < TL_SLOTS; ++i) { init[i][0] = init[i][1] = init[i][2] = ~0ull; init[i][3] = 0ull; }
  cudaError_t e = cudaMemcpyToSymbolAsync(g_timeline, init, sizeof(init), 0, cudaMemcpyHostToDevice, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_err("timeline reset: %s", cudaGetErrorString(e));
  g_tl_next = 0;
  return 0;
}
extern "C" int drn_gemm_timeline_read(unsigned long long* out, int max_slots) {
  using namespace drn::tc;
  const int n = g_tl_next < 0 ? 0 : (g_tl_next < max_slots ? g_tl_next : max_slots);
  if (n > 0) {
    cudaError_t e = cudaMemcpyFromSymbol(out, g_timeline, (size_t)n * 4 * sizeof(unsigned long long), 0, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return -1;
  }
  return n;
}
// Profiling aid (DRN_TC_DEBUG & 2048, tools/tile_trace.py; not part of the drop-in ABI): clear / read CTA 0's per-tile trace.
extern "C" int drn_gemm_trace_reset(drn_stream_t stream) {
  using namespace drn::tc;
  void* addr = nullptr;
  cudaError_t e = cudaGetSymbolAddress(&addr, g_trace);
  if (e == cudaSuccess) e = cudaMemsetAsync(addr, 0, sizeof(long long) * TR_TILES * 8, (cudaStream_t)stream);
  if (e != cudaSuccess) return set_err("trace reset: %s", cudaGetErrorString(e));
  return 0;
}
extern "C" int drn_gemm_trace_read(long long* out, int max_tiles) {
  using namespace drn::tc;
  const int n = max_tiles < TR_TILES ? max_tiles : TR_TILES;
  cudaError_t e = cudaMemcpyFromSymbol(out, g_trace, (size_t)n * 8 * sizeof(long long), 0, cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? n : -1;
}
extern "C" int drn_gemm_set_tail_split(int enabled) {
  const int prev = drn::tc::g_tail_split;
  drn::tc::g_tail_split = enabled ? 1 : 0;
  return prev;
}

static int conv_igemm_bf16_tc_impl(const void* in, int N, int H, int W, int Cin, const void* w, int ksize,
                                   int dilation, const float* scale, const float* bias, const void* residual,
                                   int relu, void* out, int out_dtype, int Cout, int ldo, float dropout_p,
                                   uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* workspace,
                                   size_t workspace_bytes, drn_stream_t stream, void* const* peer_out, int n_peers,
                                   int my_rank, int rows_per_peer) {
  using namespace drn::tc;
  DRN_CHECK_ARG(in && w && (out || n_peers > 0), "conv_igemm_bf16_tc: null pointer");
  DRN_CHECK_ARG(ksize == 1 || ksize == 3, "conv_igemm_bf16_tc: ksize %d", ksize);
  DRN_CHECK_ARG(Cin % 64 == 0, "conv_igemm_bf16_tc: Cin=%d must be a multiple of 64", Cin);
  DRN_CHECK_ARG(Cout % 8 == 0, "conv_igemm_bf16_tc: Cout=%d must be a multiple of 8", Cout);
  DRN_CHECK_ARG(ldo >= Cout && ldo % 8 == 0, "conv_igemm_bf16_tc: ldo=%d", ldo);
  DRN_CHECK_ARG(((uintptr_t)in % 16 == 0) && ((uintptr_t)w % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
                    ((uintptr_t)residual % 16 == 0),
                "conv_igemm_bf16_tc: operands must be 16-byte aligned");
  if (n_peers > 0) {
    DRN_CHECK_ARG(n_peers <= 8 && my_rank >= 0 && my_rank < n_peers, "gemm_scatter: %d peers, rank %d", n_peers, my_rank);
    DRN_CHECK_ARG(ksize == 1 && out_dtype == DRN_F32 && !residual && !scale, "gemm_scatter: fp32-output linear layers only");
    DRN_CHECK_ARG(rows_per_peer % 128 == 0 && (long long)rows_per_peer * n_peers == (long long)N * H * W,
                  "gemm_scatter: %lld rows do not split into %d blocks of %d rows (multiple of 128)", (long long)N * H * W, n_peers, rows_per_peer);
    for (int i = 0; i < n_peers; ++i) DRN_CHECK_ARG(peer_out[i] && (uintptr_t)peer_out[i] % 16 == 0, "gemm_scatter: peer window %d", i);
  }
  DRN_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "conv_igemm_bf16_tc: dropout p=%f", dropout_p);
  DRN_CHECK_ARG(!(residual && scale), "conv_igemm_bf16_tc: fold the per-channel scale into the weights when a residual is given");
  DRN_CHECK_ARG(!(dropout_p > 0.f && out_dtype == DRN_F32), "conv_igemm_bf16_tc: fused dropout needs bf16 output");
  const long long Mll = (long long)N * H * W;
  if (Mll == 0) return 0;
  DRN_CHECK_ARG(Mll < (1ll << 31), "conv_igemm_bf16_tc: too many rows");
  Params p{};
  p.N = Cout;
  p.scale = scale; p.bias = bias; p.has_residual = residual != nullptr;
  p.out = out; p.out_f32 = (out_dtype == DRN_F32); p.ldo = ldo; p.relu = relu;
  p.n_peers = n_peers; p.my_rank = my_rank; p.rows_per_peer = rows_per_peer;
  for (int i = 0; i < 8; ++i) p.peer_out[i] = (i < n_peers) ? (float*)peer_out[i] : nullptr;
  if (dropout_p > 0.f) {
    const double keep = 1.0 - (double)dropout_p;
    p.drop_keep_thresh = (uint32_t)(keep * 4294967295.0);
    p.drop_inv_keep = (float)(1.0 / keep);
    p.drop_seed = dropout_seed;
    p.drop_seed_dev = (const unsigned long long*)dropout_seed_dev;
  }
  {
    const char* e = getenv("DRN_TC_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  p.tl_slot = -1;
  if ((p.debug & 256) && g_tl_next >= 0 && g_tl_next < TL_SLOTS) p.tl_slot = g_tl_next++;
  const int Ktot = ksize * ksize * Cin;
  p.KB = Ktot / BK;
  CUtensorMap ma, mb, mo, mr;
  memset(&mo, 0, sizeof(mo));
  memset(&mr, 0, sizeof(mr));
  // HALO mode (see HALO_ROW_BYTES): 3x3 convs without shortcut whose rows fill most of a 128-pixel tile.
  // DRN_TC_HALO: 0 = never, 1 = wherever it is legal, unset = the measured rule below; DRN_TC_HALO_BO=1 sets the
  // descriptor base-offset field (hardware probe).
  static int halo_env = -2, halo_bo_env = 0;
  if (halo_env == -2) {
    const char* e = getenv("DRN_TC_HALO");
    halo_env = e ? atoi(e) : -1;
    const char* b = getenv("DRN_TC_HALO_BO");
    halo_bo_env = (b && b[0] == '1') ? 1 : 0;
  }
  bool halo = false;
  if (ksize == 3 && !residual && out_dtype != DRN_F32 && dropout_p == 0.f && dilation >= 1 && dilation <= HALO_MAX_DIL && halo_env != 0) {
    const double fill = (double)W / (128.0 * ((W + 127) / 128));
    // measured (tools/layer_bench.py, R50-WS and VGG16 maps): wins up to 128 input channels, and at 256 while the map has no more
    // row tiles than SMs (res4: 74 x 124); wider / larger layers are better off as CTA pairs with 256-wide tiles
    const long row_tiles = (long)N * H * ((W + 127) / 128);
    halo = halo_env == 1 ? true : (fill >= 0.7 && (Cin <= 128 || (Cin <= 256 && row_tiles <= 148)));
  }
  p.halo_bo = halo_bo_env;
  if (ksize == 3) {
    p.conv = 1; p.NB = N; p.H = H; p.W = W; p.Cin = Cin; p.dil = dilation;
    if (halo) { p.tile_w = 128; p.tile_h = 1; }
    else pick_conv_tile(H, W, &p.tile_w, &p.tile_h);
    p.tiles_h = (H + p.tile_h - 1) / p.tile_h; p.tiles_w = (W + p.tile_w - 1) / p.tile_w;
    p.num_m_tiles = N * p.tiles_h * p.tiles_w;
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    const cuuint32_t box[4] = {BK, (cuuint32_t)(halo ? BM + 2 * dilation : p.tile_w), (cuuint32_t)p.tile_h, 1};
    if (make_map(&ma, in, 4, dims, strides, box)) return 1;
    if (!p.out_f32) {
      const int bw = p.tile_w < 32 ? p.tile_w : 32;
      const cuuint32_t obox[4] = {EPI_CHUNK, (cuuint32_t)bw, (cuuint32_t)(32 / bw), 1};
      const cuuint64_t odims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t ostr[3] = {(cuuint64_t)ldo * 2, (cuuint64_t)W * ldo * 2, (cuuint64_t)H * W * ldo * 2};
      if (make_map(&mo, out, 4, odims, ostr, obox)) return 1;
    }
    if (residual) {
      const cuuint64_t rdims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t rstr[3] = {(cuuint64_t)Cout * 2, (cuuint64_t)W * Cout * 2, (cuuint64_t)H * W * Cout * 2};
      const cuuint32_t rbox[4] = {EPI_CHUNK, (cuuint32_t)p.tile_w, (cuuint32_t)p.tile_h, 1};
      if (make_map(&mr, residual, 4, rdims, rstr, rbox)) return 1;
    }
  } else {
    p.conv = 0; p.M = (int)Mll; p.tile_w = 32; p.tile_h = 4;
    p.num_m_tiles = (int)((Mll + BM - 1) / BM);
    const cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Mll};
    const cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    const cuuint32_t box[2] = {BK, BM};
    if (make_map(&ma, in, 2, dims, strides, box)) return 1;
    if (!p.out_f32) {
      const cuuint32_t obox[2] = {EPI_CHUNK, 32};
      const cuuint64_t odims[2] = {(cuuint64_t)Cout, (cuuint64_t)Mll};
      const cuuint64_t ostr[1] = {(cuuint64_t)ldo * 2};
      if (make_map(&mo, out, 2, odims, ostr, obox)) return 1;
    }
    if (residual) {
      const cuuint64_t rdims[2] = {(cuuint64_t)Cout, (cuuint64_t)Mll};
      const cuuint64_t rstr[1] = {(cuuint64_t)Cout * 2};
      const cuuint32_t rbox[2] = {EPI_CHUNK, BM};
      if (make_map(&mr, residual, 2, rdims, rstr, rbox)) return 1;
    }
  }
  // CTA pairs (cta_group::2) pay ~1 us of cluster launch + two cluster barriers and win once the operand feed matters:
  // measured cross-over (profiles/r1_gemm_sweep_cg2_vs_cg1.txt) at ~9000 blocks of 128x256x64 MACs.
  // DRN_TC_CTA_GROUP=1 / 2 forces single CTAs / pairs.
  static int cg_env = -1;
  if (cg_env < 0) {
    const char* e = getenv("DRN_TC_CTA_GROUP");
    cg_env = (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 0;
  }
  const long long work = (long long)p.num_m_tiles * ((Cout + 255) / 256) * p.KB;
  const int cg = halo ? 1 : (p.num_m_tiles >= 2 && (cg_env == 2 || (cg_env == 0 && work >= 9000))) ? 2 : 1;
  const int bn = halo ? pick_bn_halo(p.num_m_tiles, Cout, Cin / BK, dilation) : pick_bn(p.num_m_tiles, Cout, p.KB, cg);
  p.num_n_tiles = (Cout + bn - 1) / bn;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
    const cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
    const cuuint32_t box[2] = {BK, (cuuint32_t)(bn / cg)};
    if (make_map(&mb, w, 2, dims, strides, box)) return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (halo) {
    // resident weights (DRN_TC_BRES=0 disables): every tile of such a layer would otherwise re-read the same 72 KB of weights,
    // more than the 50 KB of input rows it needs (stem conv2 / conv3, res2 3x3, VGG conv1_2)
    static int bres_env = -1;
    if (bres_env < 0) {
      const char* e = getenv("DRN_TC_BRES");
      bres_env = (e && e[0] == '0') ? 0 : 1;
    }
    p.bres = (bres_env && Cin == BK && bn == 64 && p.num_n_tiles == 1) ? 1 : 0;
    if (p.bres) return launch<64, 9, 2, 1, 4, true>(ma, mb, mo, mr, p, st, nullptr, 0);
    if (bn == 256) return launch<256, 3, 1, 1, 4, true>(ma, mb, mo, mr, p, st, nullptr, 0);
    if (bn == 128) return launch<128, 5, 2, 1, 4, true>(ma, mb, mo, mr, p, st, nullptr, 0);
    return launch<64, 8, 2, 1, 4, true>(ma, mb, mo, mr, p, st, nullptr, 0);
  }
  // 8 epilogue warps for short-K layers (the epilogue latency chain, not the MMAs, bounds them); the split-K schedules
  // keep the 4-warp variant.  DRN_TC_EPI8_KB: largest K-block count that takes it (0 = never).
  static int epi8_kb = -1;
  if (epi8_kb < 0) {
    const char* e = getenv("DRN_TC_EPI8_KB");
    epi8_kb = e ? atoi(e) : 64;
  }
  // measured (tools/layer_bench.py, us warm, 4 -> 8 warps): 64->256 9.4 -> 7.7, 256->1024+res 12.7 -> 11.8, fc7 86 -> 77,
  // heads 30.7 -> 27.7; but the single-CTA 256-wide tile only has smem for 3 pipeline stages next to 8 staging
  // slots, which costs more than it gains once a tile streams more than 8 k-blocks (512->2048+res 25.4 -> 27.8)
  const int eff_kb = p.KB + (p.has_residual ? bn / 64 : 0);
  const bool epi8 = p.KB <= epi8_kb && !workspace && !(cg == 1 && bn == 256 && eff_kb > 8);
  if (epi8) {
    if (cg == 2) {
      if (bn == 256) return launch<256, 4, 2, 2, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
      if (bn == 128) return launch<128, 5, 2, 2, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
      return launch<64, 6, 2, 2, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
    }
    if (bn == 256) return launch<256, 3, 2, 1, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);  // two staging slots: a chunk's TMA store drains while the next is computed (tools/tile_trace.py: these tiles are epilogue-paced; 64->256+res 8.7 -> 8.0, 128->512+res 7.0 -> 6.5 us; 16 epilogue warps with one chunk each measured SLOWER, 9.8 / 7.2 us: profiles/r2_tile_trace.txt)
    if (bn == 128) return launch<128, 4, 2, 1, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
    return launch<64, 5, 2, 1, 8>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
  }
  if (cg == 2) {
    if (bn == 256 && p.KB >= 48) return launch<256, 6, 1, 2>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);  // deep K: smem goes to pipeline stages
    if (bn == 256) return launch<256, 5, 2, 2>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
    if (bn == 128) return launch<128, 7, 2, 2>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
    return launch<64, 8, 2, 2>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
  }
  if (bn == 256) return launch<256, 4, 1, 1>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
  if (bn == 128) return launch<128, 5, 2, 1>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
  return launch<64, 7, 2, 1>(ma, mb, mo, mr, p, st, workspace, workspace_bytes);
}

extern "C" int drn_conv_igemm_bf16_tc(const void* in, int N, int H, int W, int Cin, const void* w, int ksize,
                                      int dilation, const float* scale, const float* bias, const void* residual,
                                      int relu, void* out, int out_dtype, int Cout, int ldo, float dropout_p,
                                      uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* workspace,
                                      size_t workspace_bytes, drn_stream_t stream) {
  return conv_igemm_bf16_tc_impl(in, N, H, W, Cin, w, ksize, dilation, scale, bias, residual, relu, out, out_dtype, Cout, ldo,
                                 dropout_p, dropout_seed, dropout_seed_dev, workspace, workspace_bytes, stream, nullptr, 0, 0, 0);
}

extern "C" int drn_gemm_bf16_tc_scatter(const void* a, int M, int K, const void* b, int Nout, const float* bias,
                                        void* const* peer_windows, int n_peers, int my_rank, int ldo, drn_stream_t stream) {
  DRN_CHECK_ARG(n_peers > 0 && M % n_peers == 0, "gemm_scatter: %d rows over %d peers", M, n_peers);
  return conv_igemm_bf16_tc_impl(a, 1, M, 1, K, b, 1, 1, nullptr, bias, nullptr, 0, nullptr, DRN_F32, Nout, ldo, 0.f, 0, nullptr,
                                 nullptr, 0, stream, peer_windows, n_peers, my_rank, M / n_peers);
}
