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
// atoms of 4 K-rows.  A TMA box of (32 genes x 8 cells x atoms) with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
// lands in that layout directly: K groups (4 cells) SBO = 512 B apart, gene atoms LBO = 1024 B
// apart — no transpose of Z or lag is ever made.
//
// Work decomposition: CTA = one 128 x 256 output tile x one chunk of kLeeTcChunk cells.  The FP32
// tile of each chunk is written to a partial buffer; the chunks are summed in FP64 afterwards.
// Warp roles: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..5 = epilogue (TMEM -> global).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include "lee.cuh"

namespace sc {

constexpr int kTcM = 128;      // UMMA M (genes x)
constexpr int kTcN = 256;      // UMMA N (genes y)
constexpr int kTcK = 8;        // cells per stage = one tf32 UMMA K step
constexpr int kTcStages = 4;
constexpr int kTcThreads = 192;
constexpr uint32_t kTcABytes = kTcM * kTcK * 4;  // 4 KB per hi / lo
constexpr uint32_t kTcBBytes = kTcN * kTcK * 4;  // 8 KB per hi / lo
constexpr uint32_t kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;  // 24 KB
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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          tc_smem_u32(dst)),
      "l"(map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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

__global__ void __launch_bounds__(kTcThreads)
lee_tc_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
              const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
              int64_t n, int64_t chunk, float* __restrict__ partial, int64_t ldt) {
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  __shared__ uint64_t full_bar[kTcStages], empty_bar[kTcStages], tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kTcN, m0 = blockIdx.y * kTcM;
  const int64_t k_begin = (int64_t)blockIdx.z * chunk;
  const int64_t k_end = min(n, k_begin + chunk);
  const int k_steps = (int)((k_end - k_begin + kTcK - 1) / kTcK);

  // dynamic smem is only guaranteed 16-byte aligned by the runtime: align the ring to 1024 B by hand
  const uint32_t raw = tc_smem_u32(tc_smem);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  unsigned char* ring_ptr = tc_smem + (ring - raw);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { tc_mbar_init(&full_bar[s], 1); tc_mbar_init(&empty_bar[s], 1); }
    tc_mbar_init(&tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tc_smem_u32(&tmem_base_slot)),
                 "n"(kTcN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < k_steps; ++it) {
        const int s = it % kTcStages;
        const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
        tc_mbar_wait(&empty_bar[s], ph ^ 1u);
        tc_mbar_expect_tx(&full_bar[s], kTcStageBytes);
        unsigned char* st = ring_ptr + (size_t)s * kTcStageBytes;
        const int cell = (int)(k_begin + (int64_t)it * kTcK);
        tma_load_3d(st, &map_ahi, 0, cell, m0 / 32, &full_bar[s]);
        tma_load_3d(st + kTcABytes, &map_alo, 0, cell, m0 / 32, &full_bar[s]);
        tma_load_3d(st + 2 * kTcABytes, &map_bhi, 0, cell, n0 / 32, &full_bar[s]);
        tma_load_3d(st + 2 * kTcABytes + kTcBBytes, &map_blo, 0, cell, n0 / 32, &full_bar[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      for (int it = 0; it < k_steps; ++it) {
        const int s = it % kTcStages;
        const uint32_t ph = (uint32_t)(it / kTcStages) & 1u;
        tc_mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = ring + (uint32_t)s * kTcStageBytes;
        const uint64_t ahi = umma_desc_mn_sw128(base, 1024, 512);
        const uint64_t alo = umma_desc_mn_sw128(base + kTcABytes, 1024, 512);
        const uint64_t bhi = umma_desc_mn_sw128(base + 2 * kTcABytes, 1024, 512);
        const uint64_t blo = umma_desc_mn_sw128(base + 2 * kTcABytes + kTcBBytes, 1024, 512);
        umma_tf32(tmem_d, alo, bhi, it > 0 ? 1u : 0u);  // small terms first
        umma_tf32(tmem_d, ahi, blo, 1u);
        umma_tf32(tmem_d, ahi, bhi, 1u);
        umma_commit(&empty_bar[s]);  // frees the stage when these MMAs have read it
      }
      umma_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> FP32 partial tile =====
    tc_mbar_wait(&tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int lane_grp = warp & 3;               // TMEM lanes this warp may touch: 32*lane_grp ..
    const int row = m0 + lane_grp * 32 + lane;   // output row (gene x)
    float* dst = partial + ((int64_t)blockIdx.z * ldt + row) * ldt + n0;
#pragma unroll 1
    for (int c = 0; c < kTcN / 32; ++c) {
      uint32_t v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(c * 32);
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
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + c * 32 + j) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                        __uint_as_float(v[j + 3]));
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(kTcN) : "memory");
  }
}

// hi = x with the low 13 mantissa bits cleared (a TF32 value), lo = x - hi; zero padding to ldt.
__global__ void __launch_bounds__(256)
lee_split_kernel(const float* __restrict__ X, int64_t ldx, int64_t n, int g, int64_t ldt,
                 float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t Q = ldt / 4;
  const int64_t total = n * Q;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / Q;
    const int c = (int)(t - r * Q) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 3 < ldx) x = ldg4(X + r * ldx + c);
    float xs[4] = {x.x, x.y, x.z, x.w}, h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = (c + j < g) ? xs[j] : 0.f;
      h[j] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      l[j] = v - h[j];
    }
    *reinterpret_cast<float4*>(hi + r * ldt + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + r * ldt + c) = make_float4(l[0], l[1], l[2], l[3]);
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
  int64_t ldt;    // genes padded to a multiple of 256
  int64_t chunk;  // cells per CTA (multiple of 8)
  int chunks;
};

static TcPlan tc_plan(int64_t n, int g) {
  TcPlan p;
  p.ldt = (int64_t)align_up((size_t)g, kTcN);
  // The tensor core truncates (RZ) when it adds into the FP32 accumulator: a chunk of c cells makes
  // 3c/8 accumulate steps and leaves a relative bias of ~(3c/16)*2^-24.  Shorter chunks are more
  // accurate but write more partial tiles.  SC_LEE_TC_CHUNK overrides (multiple of 8).
  int64_t chunk = 256;
  if (const char* e = getenv("SC_LEE_TC_CHUNK")) { long v = atol(e); if (v >= 8) chunk = (v + 7) / 8 * 8; }
  while ((n + chunk - 1) / chunk > kLeeTcMaxChunks) chunk *= 2;
  p.chunk = chunk;
  p.chunks = (int)((n + chunk - 1) / chunk);
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

// 3-D view of a cell-major [n][ldt] FP32 matrix: (32 genes, n cells, ldt/32 gene groups); box
// (32, 8, atoms) lands as `atoms` MN-major SW128_32B gene atoms, 1024 B apart (2 K groups each).
static int make_map(CUtensorMap* map, const float* base, int64_t n, int64_t ldt, int atoms) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return SC_ERR_CUDA; }
  cuuint64_t dims[3] = {32, (cuuint64_t)n, (cuuint64_t)(ldt / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ldt * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)kTcK, (cuuint32_t)atoms};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return SC_ERR_CUDA; }
  return SC_OK;
}

bool lee_tc_supported(int64_t n, int g, int64_t lda, int64_t ldb) {
  (void)lda; (void)ldb;
  return n >= 8 && g >= 1 && g <= 8192 && get_encode() != nullptr;
}

size_t lee_tc_extra_workspace_bytes(int64_t n, int g) {
  if (n < 1 || g < 1) return 0;
  TcPlan p = tc_plan(n, g);
  size_t split = 4 * align_up(sizeof(float) * (size_t)n * (size_t)p.ldt, 1024);
  size_t part = align_up(sizeof(float) * (size_t)p.chunks * (size_t)p.ldt * (size_t)p.ldt, 1024);
  return split + part + 4096;
}

int lee_tc_launch(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n, int g,
                  const LeePlan&, double*, void* extra_ws, float* L, int64_t ldl, cudaStream_t st) {
  TcPlan p = tc_plan(n, g);
  char* w = static_cast<char*>(extra_ws);
  w = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(w), 1024));
  const size_t mat = align_up(sizeof(float) * (size_t)n * (size_t)p.ldt, 1024);
  float* ahi = reinterpret_cast<float*>(w);
  float* alo = reinterpret_cast<float*>(w + mat);
  float* bhi = reinterpret_cast<float*>(w + 2 * mat);
  float* blo = reinterpret_cast<float*>(w + 3 * mat);
  float* partial = reinterpret_cast<float*>(w + 4 * mat);

  int64_t total = n * (p.ldt / 4);
  int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  lee_split_kernel<<<blocks, 256, 0, st>>>(A, lda, n, g, p.ldt, ahi, alo);
  SC_LAUNCH_OK();
  if (B == A) {
    bhi = ahi; blo = alo;
  } else {
    lee_split_kernel<<<blocks, 256, 0, st>>>(B, ldb, n, g, p.ldt, bhi, blo);
    SC_LAUNCH_OK();
  }
  CUtensorMap mah, mal, mbh, mbl;
  int rc;
  if ((rc = make_map(&mah, ahi, n, p.ldt, kTcM / 32))) return rc;
  if ((rc = make_map(&mal, alo, n, p.ldt, kTcM / 32))) return rc;
  if ((rc = make_map(&mbh, bhi, n, p.ldt, kTcN / 32))) return rc;
  if ((rc = make_map(&mbl, blo, n, p.ldt, kTcN / 32))) return rc;

  const size_t dyn = (size_t)kTcStages * kTcStageBytes + 1024;
  SC_CUDA_OK(cudaFuncSetAttribute(lee_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  dim3 grid((unsigned)(p.ldt / kTcN), (unsigned)(p.ldt / kTcM), (unsigned)p.chunks);
  lee_tc_kernel<<<grid, kTcThreads, dyn, st>>>(mah, mal, mbh, mbl, n, p.chunk, partial, p.ldt);
  SC_LAUNCH_OK();
  dim3 blk(32, 8);
  dim3 grd((g + 31) / 32, (g + 7) / 8);
  lee_tc_reduce_kernel<<<grd, blk, 0, st>>>(partial, p.chunks, p.ldt, g, L, ldl);
  SC_LAUNCH_OK();
  return SC_OK;
}

}  // namespace sc
