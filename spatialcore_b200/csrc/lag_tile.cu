// lag_tile.cu — spatial lag W·Z (+ Moran sums, local statistic, per-cell counters) through shared-memory
// tiles of each row chunk's neighbour UNION.
//
// Replaces the reference's `W @ Z` [R spatial/autocorrelation.py:307, :864, :881] and the Moran
// numerator / denominator of the squidpy call at [R :576-583] for row-standardised binary graphs held in
// spatial (Z-curve) order.
//
// Why a tile: the register/L1 gather kernel (stats.cu, lag_stat_kernel) moves 4·nnz·G bytes through the
// L1 load path.  Measured on B200 (profiles/r02_lag_experiments.json) that path delivers ~46 B/clk/SM for
// 128-byte row pieces no matter how rows are aligned or how many lanes share a row — a multi-line LDG
// costs ~2 cycles per 128-byte line — so 400 GB of gathers at C4 cost 33 ms while HBM moves 47 GB.  The
// shared-memory crossbar serves the same 128-byte row piece in one cycle.  In spatial order the rows a
// chunk of 256 consecutive cells needs (its own rows plus a halo) number ~1.76 x 256, so:
//
//   build (once per graph):  per chunk, the sorted union `urows` of all neighbour columns (+ own rows) and,
//                            per row, its neighbour list as 16-bit words (index into urows), packed per
//                            chunk, padded to quads, with the count of real entries in the last quad;
//   run (per column block):  stage the urows' 128-byte pieces AND the chunk's word lists, offsets, degrees
//                            with cp.async (L1 bypass, every byte read once per chunk; the chunk header and
//                            the union list themselves arrive one chunk ahead), then every (row, float4
//                            lane) walks its word list out of shared memory: all loads of the inner loop
//                            are LDS, nothing in it waits for L2.
//
// One row per list: sharing a merged list between 2 / 4 consecutive rows (the union of 4 lists holds 0.49 of
// their summed lengths) was measured and lost -- the predicated adds cost more issue slots than the saved LDS
// wavefronts return (29 / 32 ms against 24.7 ms at C4, profiles/r02_lag_tile_variants.json).  The kernel is
// bound by the LSU data pipe (shared-memory wavefronts of the gathers, the word loads and the cp.async
// writes: 75 % of peak in ncu at 20.2 ms), not by HBM.
//
// Arithmetic: every row's neighbours are added in ascending column order in FP32, then scaled by 1/deg —
// exactly the order of lag_stat_kernel, so the two kernels agree bit for bit.
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sc {
namespace {

constexpr int kBuildThreads = 256;
constexpr int kQuads = 8;       // float4 lanes per staged row piece (128 bytes = all 32 banks)
constexpr int kHash = 4096;     // open-addressing slots of the build kernel (> largest cap + kBuildThreads)
constexpr int kSortMax = 2048;  // power of two >= largest cap
constexpr int kMaxBlocksY = 148 * 8;

// Shared-memory budget per chunk: `cap` staged rows of 128 bytes and `wcap` 16-bit words.
// Expected union of a square patch of kChunk cells whose neighbourhoods reach rho = sqrt(deg / pi) cell
// spacings: (sqrt(kChunk) + 2 rho)^2 (443 at degree 20; measured mean 450, max 546 on uniform points).
// Tier 0 (mean degree <= 22): 576 rows + 7 168 words = 94 KB, two CTAs of 512 threads per SM.  (With 6 144 words
// 1.9 % of the C4 chunks -- degree-20 radius graph, lists padded to multiples of four: up to 6 692 words -- fell
// to the fallback and cost 1.5 ms of a 25 ms pass; the unions themselves stay below 561 rows: scripts/tile_stats.py.)
// Tier 1: 1 280 rows + 12 288 words = 202 KB, one CTA per SM (unions up to degree ~120, words to degree ~44
// at one row per group).  Chunks that exceed the budget are computed by direct gathers (lag_overflow_kernel).
struct TileTier { int chunk, cap, wcap; };
TileTier tile_tier(int64_t n, int64_t nnz) {
  const double deg = n > 0 ? (double)nnz / (double)n : 0.0;
  int tier = deg <= 22.0 ? 0 : 1;
  if (const char* e = getenv("SC_LAG_TILE_TIER")) { int v = atoi(e); if (v == 0 || v == 1) tier = v; }
  return tier == 0 ? TileTier{256, 576, 7168} : TileTier{256, 1280, 12288};
}

struct TileLayout {
  int64_t n_chunks;
  int chunk, cap, wcap, rows, groups_per_chunk;
  size_t off_ucount, off_wtotal, off_urows, off_self, off_inv, off_ginfo, off_words, bytes;
};

TileLayout tile_layout(int64_t n, int64_t nnz, int rows) {
  TileLayout L;
  const TileTier t = tile_tier(n, nnz);
  L.rows = rows;
  L.chunk = t.chunk;
  L.cap = t.cap;
  L.wcap = t.wcap;
  L.groups_per_chunk = t.chunk / rows;
  L.n_chunks = (n + t.chunk - 1) / t.chunk;
  size_t off = 0;
  L.off_ucount = off; off += align_up(sizeof(int32_t) * (size_t)L.n_chunks, 256);
  L.off_wtotal = off; off += align_up(sizeof(int32_t) * (size_t)L.n_chunks, 256);
  L.off_urows = off;  off += align_up(sizeof(int32_t) * (size_t)L.n_chunks * (size_t)L.cap, 256);
  L.off_self = off;   off += align_up(sizeof(uint32_t) * (size_t)L.n_chunks * (size_t)L.chunk, 256);
  L.off_inv = off;    off += align_up(sizeof(float) * (size_t)L.n_chunks * (size_t)L.chunk, 256);
  L.off_ginfo = off;  off += align_up(sizeof(uint32_t) * (size_t)L.n_chunks * (size_t)L.groups_per_chunk, 256);
  L.off_words = off;  off += align_up(sizeof(uint16_t) * ((size_t)nnz + 4 * (size_t)L.n_chunks * (size_t)L.groups_per_chunk + 64), 256);
  L.bytes = off;
  return L;
}

__host__ __device__ __forceinline__ void row_span(const int32_t* __restrict__ indptr, int k_fixed, int64_t row,
                                                  int64_t* beg, int* deg) {
  if (indptr) {
    const int b = indptr[row];
    *beg = b;
    *deg = indptr[row + 1] - b;
  } else {
    *beg = row * k_fixed;
    *deg = k_fixed;
  }
}

// First word of a chunk's packed word block: 16-byte aligned.  Group lists are padded to a multiple of four
// words, so a chunk with G groups and E edges needs at most E + 3 G words, and round_up(e0, 4) + 4 * (groups
// before it) leaves room for that without a scan over chunks (e0 = CSR offset of the chunk's first row).
// (Words are 16 bits -- the index of the row in the chunk's union -- so the base is kept a multiple of 8.)
__host__ __device__ __forceinline__ int64_t chunk_words_base(int64_t e0, int64_t groups_before) {
  return ((e0 + 7) & ~(int64_t)7) + 4 * groups_before;
}

// A word: the index of the staged row piece inside the tile (16 bits; the piece lives at index * 128 bytes).

__device__ __forceinline__ uint32_t hash_slot(int32_t c) { return mix32((uint32_t)c) & (kHash - 1); }

// ------------------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------------------

// One CTA per chunk: hash-set of the chunk's columns and own rows -> compact -> bitonic sort -> urows;
// ranks go back into the hash table; one thread per group merges its R column-sorted rows into words
// (first pass: length, block scan for the packed offsets, second pass: write).
template <int R, int kChunk>
__global__ void __launch_bounds__(kBuildThreads)
tile_build_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                  int k_fixed, int cap, int wcap, int64_t n_chunks, int32_t* __restrict__ ucount,
                  int32_t* __restrict__ wtotal, int32_t* __restrict__ urows,
                  uint32_t* __restrict__ selfoff, float* __restrict__ rinv,
                  uint32_t* __restrict__ ginfo, uint16_t* __restrict__ words) {
  static_assert(R == 1, "16-bit words carry no membership mask: one row per group");
  constexpr int G = kChunk / R;
  __shared__ int32_t hkeys[kHash];
  __shared__ uint16_t hrank[kHash];
  __shared__ int32_t sorted[kSortMax];
  __shared__ int32_t glen[G + 1];
  __shared__ unsigned char gtail[G];  // unpadded list length mod 4
  __shared__ int s_count, s_pos, s_bad;
  const int tid = threadIdx.x;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t r0 = chunk * kChunk;
    const int64_t r1 = r0 + kChunk < n ? r0 + kChunk : n;
    int64_t e0, e1;
    int d_unused;
    row_span(indptr, k_fixed, r0, &e0, &d_unused);
    { int64_t b; int d; row_span(indptr, k_fixed, r1 - 1, &b, &d); e1 = b + d; }
    __syncthreads();  // previous chunk's readers are done with the tables
    for (int h = tid; h < kHash; h += kBuildThreads) hkeys[h] = -1;
    if (tid == 0) { s_count = 0; s_pos = 0; s_bad = 0; }
    __syncthreads();
    // own rows first (there are at most kChunk <= cap of them), then the neighbour columns
    const int64_t n_items = (r1 - r0) + (e1 - e0);
    for (int64_t it = tid; it < n_items; it += kBuildThreads) {
      const int32_t c = it < (r1 - r0) ? (int32_t)(r0 + it) : indices[e0 + (it - (r1 - r0))];
      uint32_t h = hash_slot(c);
      for (;;) {
        if (*(volatile int*)&s_count > cap) break;  // overflow chunk: stop inserting (the table never fills)
        const int32_t old = atomicCAS(&hkeys[h], -1, c);
        if (old == -1) { atomicAdd(&s_count, 1); break; }
        if (old == c) break;
        h = (h + 1) & (kHash - 1);
      }
    }
    __syncthreads();
    const int cnt = s_count;
    if (cnt > cap) {  // this chunk's union does not fit the tile: the run kernel leaves it to the fallback
      if (tid == 0) { ucount[chunk] = -1; wtotal[chunk] = 0; }
      continue;
    }
    for (int h = tid; h < kHash; h += kBuildThreads) {
      const int32_t k = hkeys[h];
      if (k >= 0) sorted[atomicAdd(&s_pos, 1)] = k;
    }
    int p2 = 2;
    while (p2 < cnt) p2 <<= 1;
    __syncthreads();
    for (int i = cnt + tid; i < p2; i += kBuildThreads) sorted[i] = INT_MAX;
    __syncthreads();
    for (int k = 2; k <= p2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < p2; i += kBuildThreads) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int32_t a = sorted[i], b = sorted[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { sorted[i] = b; sorted[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
    int32_t* ur = urows + chunk * cap;
    for (int i = tid; i < cnt; i += kBuildThreads) {
      const int32_t c = sorted[i];
      ur[i] = c;
      uint32_t h = hash_slot(c);
      while (hkeys[h] != c) h = (h + 1) & (kHash - 1);
      hrank[h] = (uint16_t)i;
      if (c >= r0 && c < r1) selfoff[c] = (uint32_t)i * (kQuads * 16u);
    }
    __syncthreads();
    auto rank_of = [&](int32_t c) -> uint32_t {
      uint32_t h = hash_slot(c);
      while (hkeys[h] != c) h = (h + 1) & (kHash - 1);
      return hrank[h];
    };
    // merge of group gl's rows; `out` == nullptr only counts (and records the rows' inverse degrees)
    auto merge = [&](int gl, uint16_t* out) -> int {
      int64_t beg[R];
      int len[R], pos[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = r0 + (int64_t)gl * R + r;
        beg[r] = 0; len[r] = 0; pos[r] = 0;
        if (row < r1) {
          row_span(indptr, k_fixed, row, &beg[r], &len[r]);
          if (!out) rinv[row] = len[r] > 0 ? 1.f / (float)len[r] : 0.f;
        }
      }
      int t = 0;
      for (;;) {
        int32_t best = INT_MAX;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (pos[r] < len[r]) { const int32_t c = indices[beg[r] + pos[r]]; best = c < best ? c : best; }
        if (best == INT_MAX) break;
        uint32_t mask = 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (pos[r] < len[r] && indices[beg[r] + pos[r]] == best) { mask |= 1u << r; ++pos[r]; }
        (void)mask;
        if (out) out[t] = (uint16_t)rank_of(best);
        ++t;
      }
      if (!out) gtail[gl] = (unsigned char)(t & 3);
      if (out)  // pad to a multiple of four with the tile's zero row (never gathered: the run kernel knows the tail)
        for (; t & 3; ++t) out[t] = (uint16_t)cap;
      return (t + 3) & ~3;
    };
    const int groups_here = (int)((r1 - r0 + R - 1) / R);
    for (int gl = tid; gl < G; gl += kBuildThreads) {
      if (gl >= groups_here) gtail[gl] = 0;
      const int len = gl < groups_here ? merge(gl, nullptr) : 0;
      if (len > 4 * 63) s_bad = 1;  // ginfo holds the length in quads in six bits
      glen[gl] = len;
    }
    __syncthreads();
    if (tid < 32) {  // exclusive scan of glen[0..G) by one warp; glen[G] = total
      int carry = 0;
      for (int base = 0; base < G; base += 32) {
        const int v = glen[base + tid];
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(kFull, s, o); if (tid >= o) s += t; }
        glen[base + tid] = carry + s - v;
        carry += __shfl_sync(kFull, s, 31);
      }
      if (tid == 0) glen[G] = carry;
    }
    __syncthreads();
    const int total = glen[G];
    if (total > wcap || s_bad) {
      if (tid == 0) { ucount[chunk] = -1; wtotal[chunk] = 0; }
      continue;
    }
    if (tid == 0) { ucount[chunk] = cnt; wtotal[chunk] = total; }
    uint32_t* gi = ginfo + chunk * G;
    // ginfo: first word of the list << 8 | quads << 2 | entries of the last quad that are real (0 = all four)
    for (int gl = tid; gl < G; gl += kBuildThreads)
      gi[gl] = ((uint32_t)glen[gl] << 8) | ((uint32_t)((glen[gl + 1] - glen[gl]) >> 2) << 2) | (uint32_t)gtail[gl];
    uint16_t* wb = words + chunk_words_base(e0, chunk * G);
    for (int gl = tid; gl < groups_here; gl += kBuildThreads) merge(gl, wb + glen[gl]);
  }
}

// ------------------------------------------------------------------------------------------------
// run
// ------------------------------------------------------------------------------------------------

struct LagTileArgs {
  const int32_t* indptr;
  const int32_t* indices;  // only the overflow kernel reads the CSR columns
  int k_fixed;
  int chunk, cap, wcap;
  int64_t n, n_chunks;
  const int32_t* ucount;
  const int32_t* wtotal;
  const int32_t* urows;
  const uint32_t* selfoff;
  const float* rinv;
  const uint32_t* ginfo;
  const uint16_t* words;
  const float* Zself;   // or NULL: the row's own value comes from the staged tile
  const float* Z;       // operand of the lag
  const int32_t* perm;  // or NULL; row j of the operand is Z[perm[j]] (value-permuting null, never materialised)
  const int32_t* purows;  // or NULL: urows already composed with perm (perm[urows[.]]), same layout
  int64_t ldz;
  float* lag;    // or NULL
  float* local;  // or NULL
  int64_t ldl;
  const float* cell_obs;  // or NULL
  int32_t* cell_cnt;      // or NULL
  int64_t ldc;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Two FP32 lanes per register pair: Blackwell adds them with one instruction (FADD2, add.rn.f32x2) --
// two IEEE additions, bit-identical to two FADDs, half the issue slots.
struct F4 { unsigned long long lo, hi; };  // (x, y), (z, w)
__device__ __forceinline__ F4 f4_zero() { return F4{0ull, 0ull}; }
__device__ __forceinline__ F4 f4_load(const void* smem_ptr) {
  const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(smem_ptr);
  return F4{t.x, t.y};
}
__device__ __forceinline__ void f4_add(F4& a, const F4& v) {
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.lo) : "l"(v.lo));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a.hi) : "l"(v.hi));
}
// acc += the row piece at shared address `saddr` if `cond` (predicated load and adds: no wavefront when off)
__device__ __forceinline__ void f4_load_add_if(F4& a, uint32_t saddr, int cond) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 x, y;\n\tsetp.ne.b32 p, %3, 0;\n\t@p ld.shared.v2.b64 {x, y}, [%2];\n\t"
      "@p add.rn.f32x2 %0, %0, x;\n\t@p add.rn.f32x2 %1, %1, y;\n\t}"
      : "+l"(a.lo), "+l"(a.hi)
      : "r"(saddr), "r"(cond)
      : "memory");
}
__device__ __forceinline__ float4 f4_unpack(const F4& a) {
  return make_float4(__uint_as_float((uint32_t)a.lo), __uint_as_float((uint32_t)(a.lo >> 32)),
                     __uint_as_float((uint32_t)a.hi), __uint_as_float((uint32_t)(a.hi >> 32)));
}

// Per-row epilogue shared by the tile and the overflow kernel: scale, local statistic, counters, Moran sums.
// FLAGS: bit 0 = lag is written, bit 1 = local, bit 2 = per-cell counters; 8 = decide at run time.
template <int FLAGS>
__device__ __forceinline__ void finish_row(const LagTileArgs& A, int64_t out_off /* row * ldl + col */,
                                           int64_t cnt_off /* row * ldc + col */, float inv, float4 s,
                                           const float4& z, double (&num)[4], double (&den)[4]) {
  s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
  const float4 loc = make_float4(z.x * s.x, z.y * s.y, z.z * s.z, z.w * s.w);
  if ((FLAGS & 1) || (FLAGS == 8 && A.lag)) *reinterpret_cast<float4*>(A.lag + out_off) = s;
  if ((FLAGS & 2) || (FLAGS == 8 && A.local)) *reinterpret_cast<float4*>(A.local + out_off) = loc;
  if ((FLAGS & 4) || (FLAGS == 8 && A.cell_cnt)) {
    const float4 o = ldg4(A.cell_obs + cnt_off);
    int4* cp = reinterpret_cast<int4*>(A.cell_cnt + cnt_off);
    int4 cc = *cp;
    cc.x += fabsf(loc.x) >= fabsf(o.x); cc.y += fabsf(loc.y) >= fabsf(o.y);
    cc.z += fabsf(loc.z) >= fabsf(o.z); cc.w += fabsf(loc.w) >= fabsf(o.w);
    *cp = cc;
  }
  const double zx = z.x, zy = z.y, zz = z.z, zw = z.w;
  num[0] = fma(zx, (double)s.x, num[0]); den[0] = fma(zx, zx, den[0]);
  num[1] = fma(zy, (double)s.y, num[1]); den[1] = fma(zy, zy, den[1]);
  num[2] = fma(zz, (double)s.z, num[2]); den[2] = fma(zz, zz, den[2]);
  num[3] = fma(zw, (double)s.w, num[3]); den[3] = fma(zw, zw, den[3]);
}

// Fixed-order CTA reduction of the per-thread Moran sums into partial[partial_row][{num,den}][ldz].
template <int SLOTS>
__device__ __forceinline__ void reduce_cta(double (&num)[4], double (&den)[4], double* sh /* [2][SLOTS][8][4] */,
                                           int slot, int q, int64_t col, bool active, int64_t ldz,
                                           double* __restrict__ partial, int partial_row) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    sh[((0 * SLOTS + slot) * kQuads + q) * 4 + c] = num[c];
    sh[((1 * SLOTS + slot) * kQuads + q) * 4 + c] = den[c];
  }
  __syncthreads();
  if (slot == 0 && active) {
    double* p = partial + ((int64_t)partial_row * 2) * ldz + col;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double a = 0, d = 0;
#pragma unroll 4
      for (int r = 0; r < SLOTS; ++r) {
        a += sh[((0 * SLOTS + r) * kQuads + q) * 4 + c];
        d += sh[((1 * SLOTS + r) * kQuads + q) * 4 + c];
      }
      p[c] = a; p[ldz + c] = d;
    }
  }
}


// grid.x = column block of 32 genes (8 float4 lanes), grid.y strides over chunks.  Thread = (group slot
// tid / 8, lane tid % 8): a quarter warp reads one staged row piece per LDS.128 (conflict-free).
template <int R, int FLAGS, int kChunk, int kTileThreads>
__global__ void __launch_bounds__(kTileThreads, kTileThreads == 512 ? 2 : 4)
lag_tile_kernel(const __grid_constant__ LagTileArgs A, double* __restrict__ partial) {
  static_assert(R == 1, "one row per group");
  extern __shared__ __align__(128) unsigned char tile_smem[];
  constexpr int kSlots = kTileThreads / kQuads;  // groups in flight per CTA
  constexpr int G = kChunk / R;                  // groups per chunk
  constexpr int kStageBatch = 5;
  unsigned char* tile = tile_smem;  // [cap + 1][128 bytes]; row cap is the zero row
  uint16_t* swords = reinterpret_cast<uint16_t*>(tile_smem + (size_t)(A.cap + 1) * kQuads * 16);  // [wcap + 8]
  uint32_t* sginfo = reinterpret_cast<uint32_t*>(swords + A.wcap + 8);                              // [G]
  uint32_t* sself = sginfo + G;                                                                    // [kChunk]
  float* sinv = reinterpret_cast<float*>(sself + kChunk);                                          // [kChunk]
  const int tid = threadIdx.x;
  const int q = tid & (kQuads - 1);
  const int slot = tid >> 3;
  const int64_t col = ((int64_t)blockIdx.x * kQuads + q) * 4;
  const bool active = col < A.ldz;
  const float* zcol = A.Z + col;
  const unsigned char* tq = tile + q * 16;  // this lane's column of the tile
  const uint32_t tq_s = (uint32_t)__cvta_generic_to_shared(tq);
  double num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};
  if (tid < kQuads) reinterpret_cast<float4*>(tile)[A.cap * kQuads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);  // zero row (pads)

  // A chunk's header (union size, word count, first CSR entry) and its union row list are fetched into shared
  // memory one chunk AHEAD, together with the row pieces of the chunk before it: the staging loop then finds its
  // indices on chip instead of behind two dependent global loads (13 % of the stall samples before).
  int32_t* snext = reinterpret_cast<int32_t*>(sinv + kChunk);  // [2][4 + cap]
  const int nstride = A.cap + 4;
  const int32_t* __restrict__ ur_base = A.purows ? A.purows : A.urows;
  auto prefetch = [&](int64_t c, int buf) {
    int32_t* dst = snext + buf * nstride;
    const int32_t* src = ur_base + c * A.cap;
    for (int i = tid; i * 4 < A.cap; i += kTileThreads) cp_async16(dst + 4 + i * 4, src + i * 4);
    if (tid == 0) cp_async4(dst, A.ucount + c);
    if (tid == 32) cp_async4(dst + 1, A.wtotal + c);
    if (tid == 64) {
      if (A.indptr) cp_async4(dst + 2, A.indptr + c * kChunk);
      else dst[2] = (int32_t)(c * kChunk * A.k_fixed);
    }
  };
  if ((int64_t)blockIdx.y < A.n_chunks) prefetch(blockIdx.y, 0);
  cp_async_wait_all();
  __syncthreads();

  int buf = 0;
  for (int64_t chunk = blockIdx.y; chunk < A.n_chunks; chunk += gridDim.y, buf ^= 1) {
    const int32_t* hdr = snext + buf * nstride;
    const int U = hdr[0];  // < 0: left to lag_overflow_kernel (uniform over the CTA)
    const int nW = hdr[1];
    const int64_t e0 = hdr[2];
    const int64_t r0 = chunk * kChunk;
    const int32_t* ur = hdr + 4;
    __syncthreads();  // the previous chunk's readers are done with the tile and with the other header buffer
    if (chunk + gridDim.y < A.n_chunks) prefetch(chunk + gridDim.y, buf ^ 1);
    if (U >= 0) {  // ---- stage: row pieces, word lists, group info, self offsets, inverse degrees -----------
      const uint16_t* wsrc = A.words + chunk_words_base(e0, chunk * G);
      for (int i = tid; i * 8 < nW; i += kTileThreads) cp_async16(swords + i * 8, wsrc + i * 8);
      if (tid < G / 4) cp_async16(sginfo + tid * 4, A.ginfo + chunk * G + tid * 4);
      else if (tid >= 64 && tid < 64 + kChunk / 4) cp_async16(sself + (tid - 64) * 4, A.selfoff + r0 + (tid - 64) * 4);
      else if (tid >= 128 && tid < 128 + kChunk / 4) cp_async16(sinv + (tid - 128) * 4, A.rinv + r0 + (tid - 128) * 4);
      static_assert(G / 4 <= 64 && kChunk / 4 <= 64 && kTileThreads >= 192, "staging lanes");
      if (active) {
        // the row indices of a batch are loaded together (one exposed L2 latency per batch, not per row)
        for (int u0 = slot; u0 < U; u0 += kSlots * kStageBatch) {
          int32_t src[kStageBatch];
#pragma unroll
          for (int b = 0; b < kStageBatch; ++b) { const int u = u0 + b * kSlots; src[b] = u < U ? ur[u] : -1; }
          if (A.perm && !A.purows) {
#pragma unroll
            for (int b = 0; b < kStageBatch; ++b) if (src[b] >= 0) src[b] = A.perm[src[b]];
          }
#pragma unroll
          for (int b = 0; b < kStageBatch; ++b)
            if (src[b] >= 0) cp_async16(tile + (size_t)(u0 + b * kSlots) * (kQuads * 16) + q * 16, zcol + (int64_t)src[b] * A.ldz);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (U < 0 || !active) continue;
    const int64_t out0 = r0 * A.ldl + col;  // element offset of this lane in the chunk's first output row
    const int64_t cnt0 = r0 * A.ldc + col;
#pragma unroll 1
    for (int gl = slot; gl < G; gl += kSlots) {
      if (r0 + (int64_t)gl * R >= A.n) break;
      const uint32_t gi = sginfo[gl];
      const uint2* __restrict__ wp = reinterpret_cast<const uint2*>(swords + (gi >> 8));
      // lists are padded to a multiple of four words (one LDS.64 = four neighbours); the pads of the last quad
      // are not even loaded (7 % of the gathers at degree 20): `tail` of its entries are real and are added
      // under a predicate
      const int tail = (int)(gi & 3u);
      const int nq = (int)((gi >> 2) & 63u) - (tail != 0 ? 1 : 0);
      F4 acc[R];
      acc[0] = f4_zero();
      uint2 w2 = wp[0];  // (an empty list reads the next group's first quad or the slack behind the block: unused)
#pragma unroll 1
      for (int i = 0; i < nq; ++i) {
        const uint32_t w[4] = {(w2.x & 0xffffu) << 7, (w2.x >> 16) << 7, (w2.y & 0xffffu) << 7, (w2.y >> 16) << 7};
        w2 = wp[i + 1];  // next quad while this one is consumed (one quad of slack behind every block)
        F4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = f4_load(tq + w[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) f4_add(acc[0], v[u]);
      }
      // w2 holds the partial quad (or the slack behind a full list: all three predicates off)
      f4_load_add_if(acc[0], tq_s + ((w2.x & 0xffffu) << 7), tail > 0);
      f4_load_add_if(acc[0], tq_s + ((w2.x >> 16) << 7), tail > 1);
      f4_load_add_if(acc[0], tq_s + ((w2.y & 0xffffu) << 7), tail > 2);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int lrow = gl * R + r;
        if (r0 + lrow >= A.n) break;
        const float4 z = A.Zself ? ldg4(A.Zself + (r0 + lrow) * A.ldz + col)
                                 : *reinterpret_cast<const float4*>(tq + sself[lrow]);
        finish_row<FLAGS>(A, out0 + (int64_t)lrow * A.ldl, cnt0 + (int64_t)lrow * A.ldc, sinv[lrow], f4_unpack(acc[r]), z, num, den);
      }
    }
  }
  __syncthreads();
  reduce_cta<kSlots>(num, den, reinterpret_cast<double*>(tile_smem), slot, q, col, active, A.ldz, partial, blockIdx.y);
}

// Chunks whose union did not fit the tile (ucount < 0): direct gathers through L1, same arithmetic.
// One row per (slot, lane) as in lag_stat_kernel; normally there is nothing to do and the kernel only
// reads ucount.
__global__ void __launch_bounds__(256)
lag_overflow_kernel(const __grid_constant__ LagTileArgs A, double* __restrict__ partial, int partial_row0) {
  __shared__ double sh[2 * 32 * kQuads * 4];
  const int tid = threadIdx.x;
  const int q = tid & (kQuads - 1);
  const int slot = tid >> 3;
  const int64_t col = ((int64_t)blockIdx.x * kQuads + q) * 4;
  const bool active = col < A.ldz;
  const float* zcol = A.Z + col;
  double num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};
  // Normally no chunk is flagged: the flags are scanned 256 at a time (one coalesced load per thread and a
  // block-wide OR), so the empty case costs a few microseconds instead of one dependent L2 round trip per chunk.
  for (int64_t base = (int64_t)blockIdx.y * 256; base < A.n_chunks; base += (int64_t)gridDim.y * 256) {
    const int64_t mine = base + tid;
    const int flagged = (mine < A.n_chunks && A.ucount[mine] < 0) ? 1 : 0;
    if (!__syncthreads_or(flagged)) continue;
  for (int64_t chunk = base; chunk < A.n_chunks && chunk < base + 256; ++chunk) {
    if (A.ucount[chunk] >= 0 || !active) continue;
    const int64_t r0 = chunk * A.chunk;
#pragma unroll 1
    for (int pass = 0; pass < A.chunk; pass += 32) {
      const int64_t row = r0 + pass + slot;
      if (row >= A.n) continue;
      int64_t b;
      int deg;
      row_span(A.indptr, A.k_fixed, row, &b, &deg);
      const int32_t* __restrict__ ip = A.indices + b;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int t = 0; t < deg; ++t) {
        int32_t j = ip[t];
        if (A.perm) j = A.perm[j];
        const float4 v = ldg4(zcol + (int64_t)j * A.ldz);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      const int64_t self = A.perm ? (int64_t)A.perm[row] : row;
      const float4 z = A.Zself ? ldg4(A.Zself + row * A.ldz + col) : ldg4(zcol + self * A.ldz);
      finish_row<8>(A, row * A.ldl + col, row * A.ldc + col, deg > 0 ? 1.f / (float)deg : 0.f, acc, z, num, den);
    }
  }
  }
  reduce_cta<32>(num, den, sh, slot, q, col, active, A.ldz, partial, partial_row0 + blockIdx.y);
}

// purows[chunk][u] = perm[urows[chunk][u]]: done once per permutation instead of once per (chunk, column block)
// inside the staging loop, where it is a dependent load in front of every row piece.
__global__ void compose_urows_kernel(const int32_t* __restrict__ urows, const int32_t* __restrict__ ucount,
                                     const int32_t* __restrict__ perm, int64_t n_chunks, int cap,
                                     int32_t* __restrict__ out) {
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int U = ucount[chunk];
    for (int u = threadIdx.x; u < U; u += blockDim.x) out[chunk * cap + u] = perm[urows[chunk * cap + u]];
  }
}

// out[col] = sum over partial rows, fixed order (bitwise reproducible).
__global__ void tile_reduce_kernel(const double* __restrict__ partial, int nrows, int64_t ld, int g,
                                   double* __restrict__ num, double* __restrict__ den) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= g) return;
  double a = 0, d = 0;
  for (int b = 0; b < nrows; ++b) {
    a += partial[((int64_t)b * 2) * ld + col];
    d += partial[((int64_t)b * 2 + 1) * ld + col];
  }
  num[col] = a;
  den[col] = d;
}

size_t tile_smem_bytes(const LagTileArgs& A, int R) {
  const int G = A.chunk / R;
  return (size_t)(A.cap + 1) * kQuads * 16 + sizeof(uint16_t) * ((size_t)A.wcap + 8) +
         sizeof(uint32_t) * ((size_t)G + 2 * A.chunk + 2 * ((size_t)A.cap + 4));
}

template <int R, int FLAGS, int kChunk, int kTileThreads>
int launch_tile(const LagTileArgs& A, int g, double* num, double* den, double* partial, cudaStream_t st) {
  const size_t smem = tile_smem_bytes(A, R);
  static thread_local size_t configured = 0;  // per instantiation: the dynamic shared-memory size set so far
  if (configured < smem) {
    SC_CUDA_OK(cudaFuncSetAttribute(lag_tile_kernel<R, FLAGS, kChunk, kTileThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int per_sm = 1;
  SC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lag_tile_kernel<R, FLAGS, kChunk, kTileThreads>, kTileThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int bx = (int)((A.ldz + 4 * kQuads - 1) / (4 * kQuads));
  int64_t by = ((int64_t)sm_count() * per_sm) / bx;  // one resident wave, persistent over chunks
  if (by < 1) by = 1;
  if (by > A.n_chunks) by = A.n_chunks;
  if (by > kMaxBlocksY - 16) by = kMaxBlocksY - 16;
  lag_tile_kernel<R, FLAGS, kChunk, kTileThreads><<<dim3(bx, (unsigned)by), kTileThreads, smem, st>>>(A, partial);
  SC_LAUNCH_OK();
  int64_t by2 = A.n_chunks < 16 ? A.n_chunks : 16;
  lag_overflow_kernel<<<dim3(bx, (unsigned)by2), 256, 0, st>>>(A, partial, (int)by);
  SC_LAUNCH_OK();
  tile_reduce_kernel<<<(g + 127) / 128, 128, 0, st>>>(partial, (int)(by + by2), A.ldz, g, num, den);
  SC_LAUNCH_OK();
  return SC_OK;
}

template <int R, int kChunk, int kTileThreads>
int launch_tile_flags(const LagTileArgs& A, int g, double* num, double* den, double* partial, cudaStream_t st) {
  const int flags = (A.lag ? 1 : 0) | (A.local ? 2 : 0) | (A.cell_cnt ? 4 : 0);
  if (flags == 0) return launch_tile<R, 0, kChunk, kTileThreads>(A, g, num, den, partial, st);  // statistic only (value-permuting null)
  if (flags == 1) return launch_tile<R, 1, kChunk, kTileThreads>(A, g, num, den, partial, st);  // lag + statistic (morans_i)
  return launch_tile<R, 8, kChunk, kTileThreads>(A, g, num, den, partial, st);
}

}  // namespace
}  // namespace sc

using namespace sc;

extern "C" size_t sc_graph_tile_bytes(int64_t n, int64_t nnz) {
  if (n < 1 || nnz < 0) return 0;
  return tile_layout(n, nnz, 1).bytes;
}

extern "C" int sc_graph_tile_build(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                                   int64_t nnz, void* tiles, size_t tile_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && tiles, "sc_graph_tile_build: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_graph_tile_build: need indptr or k_fixed");
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "sc_graph_tile_build: n and nnz must be below 2^31");
  SC_CHECK_ARG(indptr || nnz == n * (int64_t)k_fixed, "sc_graph_tile_build: nnz must equal n * k_fixed");
  const TileLayout L = tile_layout(n, nnz, 1);
  if (tile_bytes < L.bytes) { set_error("sc_graph_tile_build: tile buffer too small (%zu < %zu)", tile_bytes, L.bytes); return SC_ERR_WORKSPACE; }
  char* base = static_cast<char*>(tiles);
  int32_t* ucount = reinterpret_cast<int32_t*>(base + L.off_ucount);
  int32_t* wtotal = reinterpret_cast<int32_t*>(base + L.off_wtotal);
  int32_t* urows = reinterpret_cast<int32_t*>(base + L.off_urows);
  uint32_t* selfoff = reinterpret_cast<uint32_t*>(base + L.off_self);
  float* rinv = reinterpret_cast<float*>(base + L.off_inv);
  uint32_t* ginfo = reinterpret_cast<uint32_t*>(base + L.off_ginfo);
  uint16_t* words = reinterpret_cast<uint16_t*>(base + L.off_words);
  const int blocks = (int)(L.n_chunks > 148 * 16 ? 148 * 16 : L.n_chunks);
tile_build_kernel<1, 256><<<blocks, kBuildThreads, 0, st>>>(indptr, indices, n, k_fixed, L.cap, L.wcap, L.n_chunks, ucount, wtotal, urows, selfoff, rinv, ginfo, words);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_csr_lag_moran_tiled(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                                      int64_t nnz, const void* tiles, size_t tile_bytes,
                                      const float* Zself, const float* Z, const int32_t* perm, int64_t ldz,
                                      int g, float* lag, float* local, int64_t ldl, double* num, double* den,
                                      const float* cell_obs, int32_t* cell_cnt, int64_t ldc, void* ws,
                                      size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && tiles && Z && num && den && ws, "sc_csr_lag_moran_tiled: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_csr_lag_moran_tiled: need indptr or k_fixed");
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31) && nnz >= 0 && nnz < (1ll << 31), "sc_csr_lag_moran_tiled: n and nnz must be below 2^31");
  SC_CHECK_ARG(ldz % 4 == 0 && ldz >= g && g >= 1 && (size_t)ldz <= align_up((size_t)g, 32),
               "sc_csr_lag_moran_tiled: ldz must be a multiple of 4 in [g, round_up(g,32)]");
  SC_CHECK_ARG((!lag && !local) || (ldl % 4 == 0 && ldl >= ldz), "sc_csr_lag_moran_tiled: ldl must be a multiple of 4 and >= ldz");
  SC_CHECK_ARG((cell_cnt == nullptr) == (cell_obs == nullptr) && (!cell_cnt || (ldc % 4 == 0 && ldc >= ldz)),
               "sc_csr_lag_moran_tiled: cell_obs and cell_cnt go together, ldc a multiple of 4 and >= ldz");
  const TileLayout L = tile_layout(n, nnz, 1);
  if (tile_bytes < L.bytes) { set_error("sc_csr_lag_moran_tiled: tile buffer too small (%zu < %zu)", tile_bytes, L.bytes); return SC_ERR_WORKSPACE; }
  if (ws_bytes < sc_csr_lag_moran_workspace_bytes(n, g)) { set_error("sc_csr_lag_moran_tiled: workspace too small"); return SC_ERR_WORKSPACE; }
  const char* base = static_cast<const char*>(tiles);
  LagTileArgs A;
  A.indptr = indptr; A.indices = indices; A.k_fixed = k_fixed; A.chunk = L.chunk; A.cap = L.cap; A.wcap = L.wcap; A.n = n; A.n_chunks = L.n_chunks;
  A.ucount = reinterpret_cast<const int32_t*>(base + L.off_ucount);
  A.wtotal = reinterpret_cast<const int32_t*>(base + L.off_wtotal);
  A.urows = reinterpret_cast<const int32_t*>(base + L.off_urows);
  A.selfoff = reinterpret_cast<const uint32_t*>(base + L.off_self);
  A.rinv = reinterpret_cast<const float*>(base + L.off_inv);
  A.ginfo = reinterpret_cast<const uint32_t*>(base + L.off_ginfo);
  A.words = reinterpret_cast<const uint16_t*>(base + L.off_words);
  A.Zself = Zself; A.Z = Z; A.perm = perm; A.purows = nullptr; A.ldz = ldz; A.lag = lag; A.local = local; A.ldl = ldl;
  A.cell_obs = cell_obs; A.cell_cnt = cell_cnt; A.ldc = ldc;
  double* partial = static_cast<double*>(ws);
  if (perm) {
    int32_t* purows = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + sc_csr_lag_moran_workspace_bytes(0, g));
    const int blocks = (int)(L.n_chunks > 148 * 8 ? 148 * 8 : L.n_chunks);
    compose_urows_kernel<<<blocks, 256, 0, st>>>(A.urows, A.ucount, perm, L.n_chunks, L.cap, purows);
    SC_LAUNCH_OK();
    A.purows = purows;
  }
  return launch_tile_flags<1, 256, 512>(A, g, num, den, partial, st);
}
