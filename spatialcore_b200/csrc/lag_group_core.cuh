// Core of the row-group lag (lag_group.cu), written __host__ __device__ so that the same code is
// exercised on the CPU by tests/native/lag_group_host_test.cu (no GPU needed).  Internal header.
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdint.h>

namespace sc {

__host__ __device__ __forceinline__ void row_span(const int32_t* __restrict__ indptr, int k_fixed,
                                                  int64_t row, int64_t* beg, int* deg) {
  if (indptr) {
    const int b = indptr[row];
    *beg = b;
    *deg = indptr[row + 1] - b;
  } else {
    *beg = row * k_fixed;
    *deg = k_fixed;
  }
}

// Word offset of group a's union list: 16-byte aligned (the kernel reads four words per load) with room for
// the union (never longer than the group's summed row lengths) plus padding to a multiple of four.
// b0 = CSR offset of the group's first row.  The buffer holds nnz + 8 * n_groups + 8 words.
__host__ __device__ __forceinline__ int64_t group_offset(int64_t b0, int64_t a) {
  return ((b0 + 3) & ~(int64_t)3) + 8 * a;
}

// R-way merge of the (column-sorted) neighbour lists of rows [R*a, R*a + R) of an n-row CSR graph into one
// ascending list of words (membership mask << (32 - R)) | column at group_offset(); the list is padded to a
// multiple of four with null words (mask 0, a valid column) so that it can be walked with 16-byte loads.
// Returns the union length (without padding).
template <int R>
__host__ __device__ __forceinline__ int group_union(const int32_t* __restrict__ indptr,
                                                    const int32_t* __restrict__ indices, int64_t n,
                                                    int k_fixed, int64_t a, uint32_t* __restrict__ uwords) {
  int64_t beg[R];
  int len[R], pos[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = a * R + r;
    beg[r] = 0; len[r] = 0; pos[r] = 0;
    if (row < n) row_span(indptr, k_fixed, row, &beg[r], &len[r]);
  }
  uint32_t* out = uwords + group_offset(beg[0], a);
  int cnt = 0;
  uint32_t last = 0;
  for (;;) {
    int best = INT_MAX;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (pos[r] < len[r]) { const int c = indices[beg[r] + pos[r]]; best = c < best ? c : best; }
    if (best == INT_MAX) break;
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (pos[r] < len[r] && indices[beg[r] + pos[r]] == best) { mask |= 1u << r; ++pos[r]; }
    last = (uint32_t)best;
    out[cnt++] = (mask << (32 - R)) | last;
  }
  for (int t = cnt; t & 3; ++t) out[t] = last;  // null words: no owner, a column that is in cache anyway
  return cnt;
}

// Add a gathered value to the accumulators of the rows whose membership bit is set.
template <int R>
__host__ __device__ __forceinline__ void scatter_add(float4 (&acc)[R], uint32_t word, const float4& v) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if ((word >> (32 - R + r)) & 1u) { acc[r].x += v.x; acc[r].y += v.y; acc[r].z += v.z; acc[r].w += v.w; }
  }
}

template <int R>
__host__ __device__ __forceinline__ uint32_t word_column(uint32_t word) {
  return word & ((1u << (32 - R)) - 1u);
}

// ---- the per-thread body of lag_group_kernel ------------------------------------------------------------
struct LagGroupArgs {
  const int32_t* indptr;   // CSR row pointers or NULL (k_fixed entries per row)
  int k_fixed;
  const uint32_t* uwords;  // union words, group a at group_offset(CSR offset of row R*a, a)
  const int32_t* ucnt;     // union length per group
  int64_t n, n_groups;
  const float* Z;
  int64_t ldz;
  float* lag;    // or NULL
  float* local;  // or NULL
  int64_t ldl;
  const float* cell_obs;  // or NULL
  int32_t* cell_cnt;      // or NULL
  int64_t ldc;
  int64_t n_chunks;
  int chunk_groups;  // multiple of THREADS / Q
};

__host__ __device__ __forceinline__ float4 load4(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldg(reinterpret_cast<const float4*>(p));
#else
  return *reinterpret_cast<const float4*>(p);
#endif
}

__host__ __device__ __forceinline__ uint4 load_words4(const uint4* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// Geometry as lag_stat_kernel (stats.cu): block_x = column block of Q float4 quads, block_y strides over
// chunks of `chunk_groups` groups; thread `tid` owns (group slot tid / Q, column quad tid % Q) and R float4
// accumulators.  Adds this thread's share of sum z*lag and sum z*z to num / den (FP64 sums of exact FP32
// products, as everywhere on this path) and writes lag / local / cell counters of the rows it owns.
template <int R, int Q, int THREADS>
__host__ __device__ __forceinline__ void lag_group_thread(const LagGroupArgs& A, int tid, int block_x, int block_y,
                                                          int grid_y, double (&num)[4], double (&den)[4]) {
  constexpr int kSlots = THREADS / Q;  // groups per pass
  const int lane = tid & 31, warp = tid >> 5;
  const int q = lane & (Q - 1);
  const int slot = warp * (32 / Q) + lane / Q;
  const int64_t col = ((int64_t)block_x * Q + q) * 4;
  if (col >= A.ldz) return;
  const float* zcol = A.Z + col;
  const char* zbytes = reinterpret_cast<const char*>(zcol);
  const uint32_t ldzb = (uint32_t)A.ldz * 4u;  // one IMAD.WIDE.U32 per gathered row
  for (int64_t chunk = block_y; chunk < A.n_chunks; chunk += grid_y) {
    const int64_t g0 = chunk * A.chunk_groups;
#pragma unroll 1
    for (int pass = 0; pass < A.chunk_groups; pass += kSlots) {
      const int64_t a = g0 + pass + slot;
      if (a >= A.n_groups) continue;
      const int64_t row0 = a * R;
      int64_t b0;
      int deg0;
      row_span(A.indptr, A.k_fixed, row0, &b0, &deg0);
      const uint4* __restrict__ up = reinterpret_cast<const uint4*>(A.uwords + group_offset(b0, a));
      const int cnt = A.ucnt[a];
      float4 acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int t = 0; t < cnt; t += 4) {  // the list is padded to a multiple of four with null words
        const uint4 w4 = load_words4(up + (t >> 2));
        const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = load4(reinterpret_cast<const float*>(zbytes + (uint64_t)word_column<R>(w[u]) * ldzb));
#pragma unroll
        for (int u = 0; u < 4; ++u) scatter_add<R>(acc, w[u], v[u]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        if (row >= A.n) break;
        int64_t b;
        int deg;
        row_span(A.indptr, A.k_fixed, row, &b, &deg);
        const float inv = (deg > 0) ? 1.f / (float)deg : 0.f;
        float4 s = acc[r];
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
        const float4 z = load4(zcol + row * A.ldz);
        const float4 loc = make_float4(z.x * s.x, z.y * s.y, z.z * s.z, z.w * s.w);
        if (A.lag) *reinterpret_cast<float4*>(A.lag + row * A.ldl + col) = s;
        if (A.local) *reinterpret_cast<float4*>(A.local + row * A.ldl + col) = loc;
        if (A.cell_cnt) {
          const float4 o = load4(A.cell_obs + row * A.ldc + col);
          int4* cp = reinterpret_cast<int4*>(A.cell_cnt + row * A.ldc + col);
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
    }
  }
}

}  // namespace sc
