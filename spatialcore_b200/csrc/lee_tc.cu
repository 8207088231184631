// Lee's L contraction on 5th-generation tensor cores (impl 2): L = Aᵀ·B with 3xTF32 splitting.
//
//   A = A_hi + A_lo,  A_hi = fp32 with the low 13 mantissa bits cleared (exactly a TF32 value),
//   A_lo = A - A_hi (exact in FP32);  same for B.   AᵀB ~= A_hiᵀB_hi + A_hiᵀB_lo + A_loᵀB_hi
//   (the dropped lo·lo term is ~2^-22 relative).  Three tcgen05.mma.kind::tf32 per K step
//   accumulate into one FP32 accumulator tile in TMEM.
//
// Layout trick: both operands are cell-major [cells][genes], i.e. "MN-major" for this product.
// For 32-bit MN-major operands the tensor core accepts exactly one shared-memory layout,
// SWIZZLE_128B_BASE32B: rows of 128 B (32 genes), 32-byte chunks XOR-swizzled with (row mod 4),
// atoms of 4 K-rows.  A TMA box of (32 genes x 8 cells) with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// lands as one such atom: K groups (4 cells) SBO = 512 B apart, gene atoms LBO = 1024 B apart --
// no transpose of Z or lag is ever made, and genes beyond the matrix width are zero-filled by TMA.
//
// The hi/lo split is fused: TMA brings the RAW FP32 tiles of Z and lag.  The tensor core reads a raw tile as its
// TF32 hi part -- it ignores the low 13 mantissa bits (measured: bit-identical to rewriting the tile masked) --
// and the epilogue warps (idle between accumulator drains) write the lo parts into a second, shorter ring,
// fence the generic->async proxy and release the stage to the MMA lane.  No split copies exist in HBM.
//
// Work decomposition: a CTA PAIR (cta_group::2, lee_tc2_kernel) = one 256 x 256 output tile x a span of
// consecutive `chunk`-cell chunks; each CTA stages its 128 rows of A and half of the B tile, 16 cells per stage.
// (lee_tc_kernel is the single-CTA form, 128 x 256 per CTA: SC_LEE_TC_CTA2=0.)
// The tensor core truncates when it adds into its FP32 accumulator, so a chunk is kept short (128
// cells = 48 accumulate steps); chunks alternate between two 256-column TMEM accumulators, and while
// the MMA lane fills one the eight epilogue warps drain the other into FP32 REGISTER accumulators
// (round-to-nearest adds).  One FP32 tile per CTA goes to a small partial buffer ([splits] tiles,
// ~40 MB instead of one tile per chunk = 3.3 GB at C3) and the splits are summed in FP64.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = split + epilogue.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include "lee.cuh"

namespace sc {

constexpr int kTcM = 128;      // UMMA M (genes x)
constexpr int kTcN = 256;      // UMMA N (genes y)
constexpr int kTcK = 8;        // cells per stage = one tf32 UMMA K step
constexpr int kTcMaxStages = 16;
constexpr int kTcThreads = 320;
constexpr int kTcEpiWarps = 8;
constexpr uint32_t kTcABytes = kTcM * kTcK * 4;  // 4 KB
constexpr uint32_t kTcBBytes = kTcN * kTcK * 4;  // 8 KB
constexpr uint32_t kTcStageBytes = kTcABytes + kTcBBytes;  // 12 KB: a raw stage [A | B] or a lo stage [A_lo | B_lo]
constexpr int kLeeTcMaxChunks = 1024;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TCWAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TCDONE_%=;\n"
      "bra TCWAIT_%=;\n"
      "TCDONE_%=:\n"
      "}\n" ::"r"(tc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          tc_smem_u32(dst)),
      "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B_BASE32B (layout_type 1), version 1
// (Blackwell).  LBO = byte stride between 32-gene atoms, SBO = byte stride between 4-cell K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                       uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version
  d |= (uint64_t)1 << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

// kind::tf32 instruction descriptor: D = F32, A = B = TF32, both MN-major, M = 128, N = 256.
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                              ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   tc_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}

// R raw stages (TMA destinations, read by the tensor core as the hi parts) and LR lo stages.  The raw ring is
// what hides the TMA latency -- the kernel is bound by bytes in flight per SM, not by the tensor pipe or the
// split (profiles/r02_lee_tc_sensitivity.txt) -- and a lo stage only lives from the split to the retirement of
// the three MMAs that read it, so few of them are needed: 12 + 4 stages in the 192 KB the old 8 x (raw + lo)
// ring took.  Stage `it` uses raw slot it % R and lo slot it % LR; both are released by the commit of its MMAs
// (empty_bar[it % R]), which the split warps consult for iteration it - LR before they overwrite the lo slot.
template <bool RAW_HI, int R, int LR, int SG>
__global__ void __launch_bounds__(kTcThreads, 1)
lee_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
              int64_t n, int64_t chunk, int chunks_per_cta, float* __restrict__ partial, int64_t ldt) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  static_assert(LR >= 2 && LR <= R && R <= kTcMaxStages, "ring sizes");
  __shared__ uint64_t full_bar[R], split_bar[R], empty_bar[R];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kTcN, m0 = blockIdx.y * kTcM;
  const int64_t cta_begin = (int64_t)blockIdx.z * chunks_per_cta * chunk;
  const int64_t cta_end = min(n, cta_begin + (int64_t)chunks_per_cta * chunk);
  const int n_chunks = cta_end > cta_begin ? (int)((cta_end - cta_begin + chunk - 1) / chunk) : 0;

  // dynamic smem is only guaranteed 16-byte aligned by the runtime: align the ring to 1024 B by hand
  const uint32_t raw = tc_smem_u32(tc_smem);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  unsigned char* ring_ptr = tc_smem + (ring - raw);
  unsigned char* lo_ptr = ring_ptr + (size_t)R * kTcStageBytes;
  const uint32_t lo_ring = ring + (uint32_t)R * kTcStageBytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      tc_mbar_init(&full_bar[s], 1);
      tc_mbar_init(&split_bar[s], kTcEpiWarps / SG);
      tc_mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) { tc_mbar_init(&tmem_full_bar[b], 1); tc_mbar_init(&tmem_empty_bar[b], kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tc_smem_u32(&tmem_base_slot)),
                 "n"(2 * kTcN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer: raw FP32 tiles, one continuous stream of K steps over the CTA's chunks =====
    if (lane == 0) {
      const int total_steps = (int)((cta_end - cta_begin + kTcK - 1) / kTcK);
      for (int it = 0; it < total_steps; ++it) {
        const int s = it % R;
        const uint32_t ph = (uint32_t)(it / R) & 1u;
        tc_mbar_wait(&empty_bar[s], ph ^ 1u);
        tc_mbar_expect_tx(&full_bar[s], kTcABytes + kTcBBytes);
        unsigned char* st = ring_ptr + (size_t)s * kTcStageBytes;
        const int cell = (int)(cta_begin + (int64_t)it * kTcK);
#pragma unroll
        for (int a = 0; a < kTcM / 32; ++a) tma_load_2d(st + a * 1024, &map_a, m0 + 32 * a, cell, &full_bar[s]);
#pragma unroll
        for (int a = 0; a < kTcN / 32; ++a)
          tma_load_2d(st + kTcABytes + a * 1024, &map_b, n0 + 32 * a, cell, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        tc_mbar_wait(&tmem_empty_bar[buf], (((uint32_t)c >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t kb = cta_begin + (int64_t)c * chunk;
        const int steps = (int)((min(cta_end, kb + chunk) - kb + kTcK - 1) / kTcK);
        const uint32_t acc = tmem_d + (uint32_t)(buf * kTcN);
        for (int ks = 0; ks < steps; ++ks, ++it) {
          const int s = it % R;
          const uint32_t ph = (uint32_t)(it / R) & 1u;
          tc_mbar_wait(&split_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = ring + (uint32_t)s * kTcStageBytes;
          const uint32_t lbase = lo_ring + (uint32_t)(it % LR) * kTcStageBytes;
          const uint64_t ahi = umma_desc_mn_sw128(base, 1024, 512);
          const uint64_t alo = umma_desc_mn_sw128(lbase, 1024, 512);
          const uint64_t bhi = umma_desc_mn_sw128(base + kTcABytes, 1024, 512);
          const uint64_t blo = umma_desc_mn_sw128(lbase + kTcABytes, 1024, 512);
          umma_tf32(acc, alo, bhi, ks > 0 ? 1u : 0u);  // small terms first
          umma_tf32(acc, ahi, blo, 1u);
          umma_tf32(acc, ahi, bhi, 1u);
          umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
        }
        umma_commit(&tmem_full_bar[buf]);  // this chunk's accumulator is complete
      }
    }
  } else {
    // ===== split + epilogue warps =====
    const int ew = warp - 2, et = threadIdx.x - 64;   // 0..7, 0..255
    const int lane_grp = warp & 3;               // TMEM lanes this warp may touch: 32*(warp % 4) ..
    const int half = ew >> 2;                    // column half (128 of the 256 accumulator columns)
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;

    // drain chunk c's accumulator into the register tile (round-to-nearest FP32 adds)
    auto drain = [&](int c) {
      const int buf = c & 1;
      tc_mbar_wait(&tmem_full_bar[buf], ((uint32_t)c >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * kTcN + half * 128 + q * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
              "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
              "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
              "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive(&tmem_empty_bar[buf]);
    };

    int it = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t kb = cta_begin + (int64_t)c * chunk;
      const int steps = (int)((min(cta_end, kb + chunk) - kb + kTcK - 1) / kTcK);
      for (int ks = 0; ks < steps; ++ks, ++it) {
        // split stage `it`: 768 float4 (A 256, B 512), in place hi + separate lo.  The two column
        // halves of the epilogue (4 warps each) take alternate stages, so two stages are in flight.
        if (it % SG == ew / (kTcEpiWarps / SG)) {
        const int s = it % R;
        tc_mbar_wait(&full_bar[s], (uint32_t)(it / R) & 1u);
        if (it >= LR) {  // the lo slot is free once the MMAs of stage it - LR have retired
          const int j = it - LR;
          tc_mbar_wait(&empty_bar[j % R], (uint32_t)(j / R) & 1u);
        }
        unsigned char* st = ring_ptr + (size_t)s * kTcStageBytes;
        unsigned char* lst = lo_ptr + (size_t)(it % LR) * kTcStageBytes;
        constexpr int kSplitThreads = 32 * kTcEpiWarps / SG;  // threads sharing one stage
#pragma unroll
        for (int u = 0; u < 768 / kSplitThreads; ++u) {
          const int idx = (et % kSplitThreads) + u * kSplitThreads;  // float4 index in the 12 KB stage ([A | B] in both rings)
          unsigned char* hp = st + idx * 16;
          unsigned char* lp = lst + idx * 16;
          const float4 x = *reinterpret_cast<const float4*>(hp);
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
          if (!RAW_HI) *reinterpret_cast<float4*>(hp) = h;
          *reinterpret_cast<float4*>(lp) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to UMMA / TMA
        __syncwarp();
        if (lane == 0) tc_mbar_arrive(&split_bar[s]);
        }
        // the previous chunk's MMAs have retired by now: drain it while this chunk computes
        if (c > 0 && ks == (steps > 4 ? 4 : steps - 1)) drain(c - 1);
      }
    }
    if (n_chunks > 0) drain(n_chunks - 1);
    const int row = m0 + lane_grp * 32 + lane;   // output row (gene x)
    float* dst = partial + ((int64_t)blockIdx.z * ldt + row) * ldt + n0 + half * 128;
#pragma unroll
    for (int j = 0; j < 128; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * kTcN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of one TPC compute a 256 x 256 output tile, each holding its own 128
// rows of A and HALF of the B tile (128 genes).  Per SM and K step the shared memory then carries 8 KB of TMA
// writes, 8 + 8 KB of split traffic and 3 x 8 KB of operand reads instead of 12 / 12 + 12 / 3 x 12 KB -- the
// single-CTA kernel is bound by exactly that traffic and by the TMA box rate (profiles/r02_lee_tc_sensitivity.txt).
// The leader (cluster rank 0) issues the MMAs; split warps and epilogue warps of both CTAs arrive on the
// leader's barriers through the cluster window, and the MMA commits are multicast to both CTAs.
// ------------------------------------------------------------------------------------------------
constexpr int kTc2Raw = 8, kTc2Lo = 4;  // stages of the raw / lo rings (16 KB each: 192 KB)
constexpr int kTc2KS = 16;                                       // cells per stage
// kind::tf32, D = F32, both MN-major, M = 256 (pair), N = 256
constexpr uint32_t kTc2Idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)((2 * kTcM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `local` (a shared-memory object of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map(const void* local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(tc_smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ void tc_mbar_arrive_cluster(uint32_t cluster_addr) {
  // default (CTA-scope release) semantics: a cluster-scope release costs a full memory barrier per arrival
  // (43 % of the stall samples when it was tried); the data handed over lives in shared memory and is made
  // visible to the async proxy by the fence.proxy.async before the arrival
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma2_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kTc2Idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this CTA-relative address in both CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc_smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int R, int LR, int KS>
__global__ void __launch_bounds__(kTcThreads, 1)
lee_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               int64_t n, int64_t chunk, int chunks_per_cta, float* __restrict__ partial, int64_t ldt) {
  // KS cells per stage (KS / 8 tf32 K steps): the per-stage costs -- three barrier hand-overs, the MMA lane's
  // wait / descriptor / commit sequence -- are paid once per KS cells
  static_assert(KS == 8 || KS == 16 || KS == 32, "cells per stage");
  constexpr uint32_t kA = kTcM * KS * 4;           // this CTA's A rows
  constexpr uint32_t kStage = kA + (kTcN / 2) * KS * 4;  // + its half of the B tile
  constexpr uint32_t kAtom = 128 * KS;             // one 32-gene atom of a stage (TMA box of 32 genes x KS cells)
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ uint64_t full_bar[R], split_bar[R], empty_bar[R];
  __shared__ uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader; the pair is (blockIdx.x even, odd): a CTA pair lies along x
  const int n0 = blockIdx.y * kTcN, m0 = blockIdx.x * kTcM;
  const int nb0 = n0 + (int)rank * (kTcN / 2);  // this CTA's half of the B tile
  const int64_t cta_begin = (int64_t)blockIdx.z * chunks_per_cta * chunk;
  const int64_t cta_end = min(n, cta_begin + (int64_t)chunks_per_cta * chunk);
  const int n_chunks = cta_end > cta_begin ? (int)((cta_end - cta_begin + chunk - 1) / chunk) : 0;

  const uint32_t raw = tc_smem_u32(tc_smem);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  unsigned char* ring_ptr = tc_smem + (ring - raw);
  unsigned char* lo_ptr = ring_ptr + (size_t)R * kStage;
  const uint32_t lo_ring = ring + (uint32_t)R * kStage;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      tc_mbar_init(&full_bar[s], 1);
      tc_mbar_init(&split_bar[s], 2 * (kTcEpiWarps / 2));  // the split warps of both CTAs (leader's copy is used)
      tc_mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) { tc_mbar_init(&tmem_full_bar[b], 1); tc_mbar_init(&tmem_empty_bar[b], 2 * kTcEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tc_smem_u32(&tmem_base_slot)),
                 "n"(2 * kTcN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated on both SMs
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 rows of A and 128 columns of B =====
    if (lane == 0) {
      const int total_steps = (int)((cta_end - cta_begin + KS - 1) / KS);
      for (int it = 0; it < total_steps; ++it) {
        const int s = it % R;
        const uint32_t ph = (uint32_t)(it / R) & 1u;
        tc_mbar_wait(&empty_bar[s], ph ^ 1u);
        tc_mbar_expect_tx(&full_bar[s], kStage);
        unsigned char* st = ring_ptr + (size_t)s * kStage;
        const int cell = (int)(cta_begin + (int64_t)it * KS);
#pragma unroll
        for (int a = 0; a < kTcM / 32; ++a) tma_load_2d(st + a * kAtom, &map_a, m0 + 32 * a, cell, &full_bar[s]);
#pragma unroll
        for (int a = 0; a < kTcN / 64; ++a)
          tma_load_2d(st + kA + a * kAtom, &map_b, nb0 + 32 * a, cell, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane of the leader CTA =====
    if (lane == 0 && rank == 0) {
      int it = 0;
      for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        tc_mbar_wait(&tmem_empty_bar[buf], (((uint32_t)c >> 1) & 1u) ^ 1u);  // both epilogues drained it
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t kb = cta_begin + (int64_t)c * chunk;
        const int steps = (int)((min(cta_end, kb + chunk) - kb + KS - 1) / KS);
        const uint32_t acc = tmem_d + (uint32_t)(buf * kTcN);
        for (int ks = 0; ks < steps; ++ks, ++it) {
          const int s = it % R;
          const uint32_t ph = (uint32_t)(it / R) & 1u;
          tc_mbar_wait(&split_bar[s], ph);  // both CTAs have staged and split this K step
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t base = ring + (uint32_t)s * kStage;
          const uint32_t lbase = lo_ring + (uint32_t)(it % LR) * kStage;
          const uint64_t ahi = umma_desc_mn_sw128(base, kAtom, 512);
          const uint64_t alo = umma_desc_mn_sw128(lbase, kAtom, 512);
          const uint64_t bhi = umma_desc_mn_sw128(base + kA, kAtom, 512);
          const uint64_t blo = umma_desc_mn_sw128(lbase + kA, kAtom, 512);
#pragma unroll
          for (int j = 0; j < KS / 8; ++j) {  // K step j: two 4-cell groups, 1024 B further into every atom
            const uint64_t off = (uint64_t)(j * 1024 >> 4);
            umma2_tf32(acc, alo + off, bhi + off, (ks > 0 || j > 0) ? 1u : 0u);  // small terms first
            umma2_tf32(acc, ahi + off, blo + off, 1u);
            umma2_tf32(acc, ahi + off, bhi + off, 1u);
          }
          umma2_commit(&empty_bar[s]);  // frees the stage in both CTAs
        }
        umma2_commit(&tmem_full_bar[buf]);  // this chunk's accumulators (both CTAs' halves) are complete
      }
    }
  } else {
    // ===== split + epilogue warps =====
    const int ew = warp - 2, et = threadIdx.x - 64;   // 0..7, 0..255
    const int lane_grp = warp & 3;               // TMEM lanes this warp may touch: 32*(warp % 4) ..
    const int half = ew >> 2;                    // column half (128 of the 256 accumulator columns)
    float acc[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) acc[j] = 0.f;

    auto drain = [&](int c) {
      const int buf = c & 1;
      tc_mbar_wait(&tmem_full_bar[buf], ((uint32_t)c >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(buf * kTcN + half * 128 + q * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
              "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
              "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
              "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[q * 32 + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc_mbar_arrive_cluster(cluster_map(&tmem_empty_bar[buf], 0));
    };

    int it = 0;
    for (int c = 0; c < n_chunks; ++c) {
      const int64_t kb = cta_begin + (int64_t)c * chunk;
      const int steps = (int)((min(cta_end, kb + chunk) - kb + KS - 1) / KS);
      for (int ks = 0; ks < steps; ++ks, ++it) {
        // split stage `it` ([A | B half], kStage / 16 float4); the two groups of four warps take alternate stages
        if ((it & 1) == half) {
          const int s = it % R;
          tc_mbar_wait(&full_bar[s], (uint32_t)(it / R) & 1u);
          if (it >= LR) {  // the lo slot is free once the MMAs of stage it - LR have retired
            const int j = it - LR;
            tc_mbar_wait(&empty_bar[j % R], (uint32_t)(j / R) & 1u);
          }
          const unsigned char* st = ring_ptr + (size_t)s * kStage;
          unsigned char* lst = lo_ptr + (size_t)(it % LR) * kStage;
#pragma unroll
          for (int u = 0; u < (int)(kStage / 16 / 128); ++u) {
            const int idx = (et & 127) + u * 128;
            const float4 x = *reinterpret_cast<const float4*>(st + idx * 16);
            float4 l;  // the tensor core reads the raw tile as its TF32 hi part (it ignores the low 13 bits)
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            *reinterpret_cast<float4*>(lst + idx * 16) = l;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> visible to UMMA
          __syncwarp();
          if (lane == 0) tc_mbar_arrive_cluster(cluster_map(&split_bar[s], 0));
        }
        if (c > 0 && ks == (steps > 32 / KS ? 32 / KS : steps - 1)) drain(c - 1);  // ~32 cells into the next chunk
      }
    }
    if (n_chunks > 0) drain(n_chunks - 1);
    const int row = m0 + lane_grp * 32 + lane;   // output row (gene x)
    float* dst = partial + ((int64_t)blockIdx.z * ldt + row) * ldt + n0 + half * 128;
#pragma unroll
    for (int j = 0; j < 128; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
  }

  // no CTA of the pair may leave (or free its TMEM) while the other can still be read or signalled
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * kTcN) : "memory");
  }
}

__global__ void lee_tc_reduce_kernel(const float* __restrict__ partial, int chunks, int64_t ldt, int g,
                                     float* __restrict__ L, int64_t ldl) {
  int x = blockIdx.y * blockDim.y + threadIdx.y;
  int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g || y >= g) return;
  double s = 0;
  for (int z = 0; z < chunks; ++z) s += (double)partial[((int64_t)z * ldt + x) * ldt + y];
  L[(int64_t)x * ldl + y] = (float)s;
}

struct TcPlan {
  int64_t ldt;         // genes padded to a multiple of 256
  int64_t chunk;       // cells per TMEM accumulation (multiple of 8)
  int chunks;          // over all cells
  int chunks_per_cta;  // consecutive chunks one CTA folds in registers
  int splits;          // CTAs along the cell dimension = partial tiles
};

static TcPlan tc_plan(int64_t n, int g) {
  TcPlan p;
  p.ldt = (int64_t)align_up((size_t)g, kTcN);
  // The tensor core truncates (RZ) when it adds into the FP32 accumulator: a chunk of c cells makes
  // 3c/8 accumulate steps and leaves a relative bias of ~(3c/16)*2^-24 of the accumulator.  Shorter chunks
  // are more accurate and cost more drains.  Measured at C3 size on uncorrelated data with the CTA-pair kernel
  // (B200, scripts/lee_tc_chunk.py; error in units of 2^-24 * sum|terms|, median / p99 / max): 512 cells
  // 0.158 / 0.61 / 5.7 at 1.51 ms, 256: 0.086 / 0.33 / 3.1 at 1.51 ms, 128: 0.050 / 0.20 / 1.9 at 1.53 ms,
  // 64: 0.033 / 0.13 / 1.4 at 1.64 ms, 32: 0.028 / 0.12 / 1.2 at 1.83 ms -- 128 buys 1.7x the accuracy of 256
  // for 1 % of the time.  SC_LEE_TC_CHUNK overrides.
  int64_t chunk = 128;
  if (const char* e = getenv("SC_LEE_TC_CHUNK")) { long v = atol(e); if (v >= 8) chunk = (v + 31) / 32 * 32; }  // a multiple of every stage length
  p.chunk = chunk;
  p.chunks = (int)((n + chunk - 1) / chunk);
  // one CTA per SM (512 TMEM columns): ~2 waves of CTAs over the (tiles x splits) grid
  const int tiles = (int)((p.ldt / kTcN) * (p.ldt / kTcM));
  int splits = (2 * sm_count()) / tiles;  // floor: a partial third wave would leave most SMs idle at the end
  if (splits > p.chunks) splits = p.chunks;
  if (splits > kLeeTcMaxChunks) splits = kLeeTcMaxChunks;
  if (splits < 1) splits = 1;
  p.chunks_per_cta = (p.chunks + splits - 1) / splits;
  p.splits = (p.chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
  return p;
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D view of a cell-major [n][ld] FP32 matrix: (ld genes, n cells); a box of (32 genes, 8 cells)
// lands as one MN-major SW128_32B atom (1024 B, two 4-cell K groups).  Genes >= ld and cells >= n are
// out of bounds for the map and arrive as zeros.
static int make_map(CUtensorMap* map, const float* base, int64_t n, int64_t ld, int box_cells = kTcK) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SC_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)n};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_cells};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SC_ERR_CUDA; }
  return SC_OK;
}

bool lee_tc_supported(int64_t n, int g, int64_t lda, int64_t ldb) {
  // TMA: row pitch a multiple of 16 bytes (base alignment is checked at launch)
  return n >= 8 && g >= 1 && g <= 8192 && lda % 4 == 0 && ldb % 4 == 0 && get_encode() != nullptr;
}

size_t lee_tc_extra_workspace_bytes(int64_t n, int g) {
  if (n < 1 || g < 1) return 0;
  TcPlan p = tc_plan(n, g);
  return align_up(sizeof(float) * (size_t)p.splits * (size_t)p.ldt * (size_t)p.ldt, 1024) + 4096;
}

int lee_tc_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n, int g,
                  const LeePlan&, double*, void* extra_ws, float* L, int64_t ldl, cudaStream_t st) {
  TcPlan p = tc_plan(n, g);
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) {
    set_error("sc_lee_gemm: tcgen05 path needs 16-byte aligned operands");
    return SC_ERR_UNSUPPORTED;
  }
  char* w = static_cast<char*>(extra_ws);
  w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(w), 1024));
  float* partial = reinterpret_cast<float*>(w);
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, A, n, lda))) return rc;
  if ((rc = make_map(&mb, B, n, ldb))) return rc;

  dim3 grid((unsigned)(p.ldt / kTcN), (unsigned)(p.ldt / kTcM), (unsigned)p.splits);
  // SC_LEE_TC_MASK_HI=1 rewrites the raw tile as its TF32 hi part instead of letting the tensor core truncate it
  // (bit-identical results, measured); SC_LEE_TC_CTA2=0 forces the single-CTA kernel.
  const char* mask = getenv("SC_LEE_TC_MASK_HI");
  const bool raw_hi = !(mask && mask[0] == '1');
  const char* c2 = getenv("SC_LEE_TC_CTA2");
  const bool pair = raw_hi && !(c2 && c2[0] == '0');
  bool pair_failed = false;
#define SC_LEE_GO(RH, RR, LL, SG)                                                                               \
  do {                                                                                                          \
    const size_t dyn = (size_t)(RR + LL) * kTcStageBytes + 1024;                                                \
    SC_CUDA_OK(cudaFuncSetAttribute(lee_tc_kernel<RH, RR, LL, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    lee_tc_kernel<RH, RR, LL, SG><<<grid, kTcThreads, dyn, st>>>(ma, mb, n, p.chunk, p.chunks_per_cta, partial, p.ldt); \
  } while (0)
  if (pair) {
    int cfgv = 0;  // SC_LEE_TC2_CFG = KS * 10000 + R * 100 + LR (experiments)
    if (const char* e = getenv("SC_LEE_TC2_CFG")) cfgv = atoi(e);
    int ks = kTc2KS, rr = kTc2Raw, ll = kTc2Lo;
    void (*kern)(const CUtensorMap, const CUtensorMap, int64_t, int64_t, int, float*, int64_t) = lee_tc2_kernel<kTc2Raw, kTc2Lo, kTc2KS>;
#define SC_TC2_PICK(K, RR, LL) if (cfgv == K * 10000 + RR * 100 + LL) { kern = lee_tc2_kernel<RR, LL, K>; ks = K; rr = RR; ll = LL; }
    SC_TC2_PICK(16, 6, 4) SC_TC2_PICK(8, 12, 8) SC_TC2_PICK(32, 3, 2)
#undef SC_TC2_PICK
    if (p.chunk % ks != 0) { set_error("sc_lee_gemm: chunk must be a multiple of the stage length"); return SC_ERR_INVALID; }
    if (ks != kTcK) {
      if ((rc = make_map(&ma, A, n, lda, ks))) return rc;
      if ((rc = make_map(&mb, B, n, ldb, ks))) return rc;
    }
    const size_t dyn = (size_t)(rr + ll) * (size_t)((kTcM + kTcN / 2) * ks * 4) + 1024;
    SC_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid.y, grid.x, grid.z);  // x = 128-row tiles (the pair), y = 256-column blocks
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;  // the pair: two consecutive 128-row tiles of one column block
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ma, mb, n, p.chunk, p.chunks_per_cta, partial, p.ldt);
    if (le == cudaErrorInvalidClusterSize || le == cudaErrorLaunchOutOfResources) {
      // a device (partition) on which the pair cannot be co-scheduled: the single-CTA kernel computes the same bits
      (void)cudaGetLastError();
      pair_failed = true;
      if (ks != kTcK) {
        if ((rc = make_map(&ma, A, n, lda))) return rc;
        if ((rc = make_map(&mb, B, n, ldb))) return rc;
      }
    } else if (le != cudaSuccess) {
      int clusters = -1;
      cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg);
      set_error("lee_tc2_kernel launch failed: %s (grid %u x %u x %u, %zu B dynamic smem, max active clusters %d)",
                cudaGetErrorString(le), cfg.gridDim.x, cfg.gridDim.y, cfg.gridDim.z, dyn, clusters);
      (void)cudaGetLastError();
      return SC_ERR_CUDA;
    }
  }
  if (!pair || pair_failed) {
    if (raw_hi) SC_LEE_GO(true, 8, 8, 2);
    else SC_LEE_GO(false, 8, 8, 2);
  }
#undef SC_LEE_GO
  SC_LAUNCH_OK();
  dim3 blk(32, 8);
  dim3 grd((g + 31) / 32, (g + 7) / 8);
  lee_tc_reduce_kernel<<<grd, blk, 0, st>>>(partial, p.splits, p.ldt, g, L, ldl);
  SC_LAUNCH_OK();
  return SC_OK;
}

}  // namespace sc
