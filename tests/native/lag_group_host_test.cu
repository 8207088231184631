// Host-side check of the row-group lag core (spatialcore_b200/csrc/lag_group_core.cuh): the same
// __host__ __device__ code the kernels run is executed on the CPU -- union build + masked accumulation must
// reproduce the plain per-row sums.  TEST INFRASTRUCTURE; compiled and run by tests/test_abi.py (no GPU).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../spatialcore_b200/csrc/lag_group_core.cuh"

static uint32_t rng_state = 12345u;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

template <int R>
static int check(int n, int max_deg, int k_fixed) {
  std::vector<int32_t> indptr(1, 0), indices;
  for (int i = 0; i < n; ++i) {
    int deg = k_fixed > 0 ? k_fixed : (int)(rnd() % (max_deg + 1));
    std::vector<char> seen(n, 0);
    int lo = i - 12 < 0 ? 0 : i - 12, hi = i + 12 >= n ? n - 1 : i + 12;  // near-diagonal: rows share neighbours
    if (deg > hi - lo + 1) deg = hi - lo + 1;
    for (int t = 0; t < deg;) { int j = lo + (int)(rnd() % (hi - lo + 1)); if (!seen[j]) { seen[j] = 1; ++t; } }
    for (int j = 0; j < n; ++j) if (seen[j]) indices.push_back(j);
    indptr.push_back((int32_t)indices.size());
  }
  std::vector<float4> z(n);
  for (int i = 0; i < n; ++i) z[i] = make_float4((float)(rnd() % 1000) / 37.f, (float)(rnd() % 1000) / 91.f, 1.f, (float)i);
  std::vector<uint32_t> words(indices.size() + 1, 0xdeadbeefu);
  const int32_t* ip = k_fixed > 0 ? nullptr : indptr.data();
  const int n_groups = (n + R - 1) / R;
  long total = 0;
  for (int a = 0; a < n_groups; ++a) {
    const int cnt = sc::group_union<R>(ip, indices.data(), n, k_fixed, a, words.data());
    total += cnt;
    int64_t b0; int d0;
    sc::row_span(ip, k_fixed, (int64_t)a * R, &b0, &d0);
    int64_t span_end = indptr[(a + 1) * R < n ? (a + 1) * R : n];
    if (b0 + cnt > span_end) { printf("R=%d group %d overflows its CSR span\n", R, a); return 1; }
    float4 acc[R];
    for (int r = 0; r < R; ++r) acc[r] = make_float4(0, 0, 0, 0);
    uint32_t prev = 0;
    for (int t = 0; t < cnt; ++t) {
      const uint32_t w = words[b0 + t], c = sc::word_column<R>(w);
      if (t > 0 && c < prev) { printf("R=%d group %d union not ascending\n", R, a); return 1; }
      if ((w >> (32 - R)) == 0) { printf("R=%d group %d empty mask\n", R, a); return 1; }
      prev = c;
      sc::scatter_add<R>(acc, w, z[c]);
    }
    for (int r = 0; r < R; ++r) {
      const int row = a * R + r;
      if (row >= n) { if (acc[r].z != 0.f) { printf("R=%d phantom row %d\n", R, row); return 1; } continue; }
      float4 want = make_float4(0, 0, 0, 0);
      for (int e = indptr[row]; e < indptr[row + 1]; ++e) {  // ascending column order, like the union walk
        const float4 v = z[indices[e]];
        want.x += v.x; want.y += v.y; want.z += v.z; want.w += v.w;
      }
      if (want.x != acc[r].x || want.y != acc[r].y || want.z != acc[r].z || want.w != acc[r].w) {
        printf("R=%d row %d: got (%g %g %g %g) want (%g %g %g %g)\n", R, row, acc[r].x, acc[r].y, acc[r].z, acc[r].w,
               want.x, want.y, want.z, want.w);
        return 1;
      }
    }
  }
  if (words[indices.size()] != 0xdeadbeefu) { printf("R=%d wrote past the end\n", R); return 1; }
  printf("R=%d n=%d k_fixed=%d: union/nnz = %.3f ok\n", R, n, k_fixed, (double)total / (double)indices.size());
  return 0;
}

int main() {
  int rc = 0;
  for (int n : {1, 7, 64, 1001}) {
    rc |= check<2>(n, 9, 0); rc |= check<4>(n, 9, 0); rc |= check<8>(n, 9, 0);
    if (n > 6) { rc |= check<2>(n, 0, 6); rc |= check<4>(n, 0, 6); rc |= check<8>(n, 0, 6); }
  }
  return rc;
}
