// Core of the row-group lag (lag_group.cu), written __host__ __device__ so that the same code is
// exercised on the CPU by tests/native/lag_group_host_test.cu (no GPU needed).  Internal header.
#pragma once

#include <cuda_runtime.h>
#include <limits.h>
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

// R-way merge of the (column-sorted) neighbour lists of rows [R*a, R*a + R) of an n-row CSR graph.
// The union is written at the CSR offset of the group's first row -- the R rows are contiguous in the CSR
// and the union is never longer than their summed lengths, so `uwords` has the size of `indices` and no
// scan is needed.  word = (membership mask << (32 - R)) | column.  Returns the union length.
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
  uint32_t* out = uwords + beg[0];
  int cnt = 0;
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
    out[cnt++] = (mask << (32 - R)) | (uint32_t)best;
  }
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

}  // namespace sc
