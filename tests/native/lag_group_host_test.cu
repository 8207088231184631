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
  const int32_t* ip = k_fixed > 0 ? nullptr : indptr.data();
  const int n_groups = (n + R - 1) / R;
  const size_t cap = indices.size() + 8 * (size_t)n_groups + 8;  // the documented buffer size
  std::vector<uint32_t> words(cap + 1, 0xdeadbeefu);
  long total = 0;
  for (int a = 0; a < n_groups; ++a) {
    const int cnt = sc::group_union<R>(ip, indices.data(), n, k_fixed, a, words.data());
    total += cnt;
    int64_t b0; int d0;
    sc::row_span(ip, k_fixed, (int64_t)a * R, &b0, &d0);
    const int64_t off = sc::group_offset(b0, a);
    int64_t next_b0 = indptr[(a + 1) * R < n ? (a + 1) * R : n];
    const int64_t limit = a + 1 < n_groups ? sc::group_offset(next_b0, a + 1) : (int64_t)cap;
    const int padded = (cnt + 3) & ~3;
    if ((off & 3) || off + padded > limit) { printf("R=%d group %d: list [%lld, +%d) misaligned or past %lld\n", R, a, (long long)off, padded, (long long)limit); return 1; }
    for (int t = cnt; t < padded; ++t)
      if ((words[off + t] >> (32 - R)) != 0 || sc::word_column<R>(words[off + t]) >= (uint32_t)n) { printf("R=%d group %d bad padding\n", R, a); return 1; }
    float4 acc[R];
    for (int r = 0; r < R; ++r) acc[r] = make_float4(0, 0, 0, 0);
    uint32_t prev = 0;
    for (int t = 0; t < cnt; ++t) {
      const uint32_t w = words[off + t], c = sc::word_column<R>(w);
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
  if (words[cap] != 0xdeadbeefu) { printf("R=%d wrote past the end\n", R); return 1; }
  printf("R=%d n=%d k_fixed=%d: union/nnz = %.3f ok\n", R, n, k_fixed, (double)total / (double)indices.size());
  return 0;
}

// The whole per-thread body of lag_group_kernel over an emulated launch grid: every (block_x, block_y, thread)
// is executed in turn on the host; lag / local / per-cell counters and the Moran sums must equal a direct
// evaluation, which also proves that the (chunk, pass, slot) geometry visits every (row, column) exactly once.
template <int R, int Q = 8>
static int check_kernel_body(int n, int g, int ld, int k_fixed, int grid_y) {
  constexpr int THREADS = 256;
  std::vector<int32_t> indptr(1, 0), indices;
  for (int i = 0; i < n; ++i) {
    int deg = k_fixed > 0 ? k_fixed : (int)(rnd() % 8);  // CSR case: some empty rows
    int lo = i - 9 < 0 ? 0 : i - 9, hi = i + 9 >= n ? n - 1 : i + 9;
    if (deg > hi - lo + 1) deg = hi - lo + 1;
    std::vector<char> seen(n, 0);
    for (int t = 0; t < deg;) { int j = lo + (int)(rnd() % (hi - lo + 1)); if (!seen[j]) { seen[j] = 1; ++t; } }
    for (int j = lo; j <= hi; ++j) if (seen[j]) indices.push_back(j);
    indptr.push_back((int32_t)indices.size());
  }
  const int32_t* ip = k_fixed > 0 ? nullptr : indptr.data();
  const int n_groups = (n + R - 1) / R;
  std::vector<uint4> word_store((indices.size() + 8 * (size_t)n_groups + 8) / 4 + 1);  // 16-byte aligned
  uint32_t* words_p = reinterpret_cast<uint32_t*>(word_store.data());
  struct { uint32_t* p; uint32_t* data() { return p; } } words{words_p};
  std::vector<int32_t> ucnt(n_groups);
  for (int a = 0; a < n_groups; ++a) ucnt[a] = sc::group_union<R>(ip, indices.data(), n, k_fixed, a, words.data());
  const size_t quads = (size_t)n * ld / 4;
  std::vector<float4> Zs(quads), lag(quads), loc(quads), obs(quads);
  std::vector<int4> cnt(quads);
  float* Z = reinterpret_cast<float*>(Zs.data());
  float* O = reinterpret_cast<float*>(obs.data());
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < ld; ++c) {
      Z[(size_t)i * ld + c] = c < g ? ((float)(rnd() % 2001) - 1000.f) / 512.f : 0.f;
      O[(size_t)i * ld + c] = (float)(rnd() % 1000) / 4096.f;
    }
  for (size_t e = 0; e < quads; ++e) { lag[e] = make_float4(-7, -7, -7, -7); loc[e] = lag[e]; cnt[e] = make_int4(5, 5, 5, 5); }
  sc::LagGroupArgs A;
  A.indptr = ip; A.k_fixed = k_fixed; A.uwords = words.data(); A.ucnt = ucnt.data(); A.n = n; A.n_groups = n_groups;
  A.Z = Z; A.ldz = ld; A.lag = reinterpret_cast<float*>(lag.data()); A.local = reinterpret_cast<float*>(loc.data()); A.ldl = ld;
  A.cell_obs = O; A.cell_cnt = reinterpret_cast<int32_t*>(cnt.data()); A.ldc = ld;
  A.chunk_groups = 64; A.n_chunks = (n_groups + A.chunk_groups - 1) / A.chunk_groups;
  const int grid_x = (ld + 4 * Q - 1) / (4 * Q);
  std::vector<double> num(ld, 0.0), den(ld, 0.0);
  for (int bx = 0; bx < grid_x; ++bx)
    for (int by = 0; by < grid_y; ++by)
      for (int tid = 0; tid < THREADS; ++tid) {
        double a[4] = {0, 0, 0, 0}, d[4] = {0, 0, 0, 0};
        sc::lag_group_thread<R, Q, THREADS>(A, tid, bx, by, grid_y, a, d);
        const int col = (bx * Q + (tid & (Q - 1))) * 4;
        for (int c = 0; c < 4; ++c) if (col + c < ld) { num[col + c] += a[c]; den[col + c] += d[c]; }
      }
  const float* L = reinterpret_cast<const float*>(lag.data());
  const float* I = reinterpret_cast<const float*>(loc.data());
  const int32_t* C = reinterpret_cast<const int32_t*>(cnt.data());
  for (int c = 0; c < ld; ++c) {
    double want_num = 0, want_den = 0;
    for (int i = 0; i < n; ++i) {
      float s = 0.f;
      for (int e = indptr[i]; e < indptr[i + 1]; ++e) s += Z[(size_t)indices[e] * ld + c];  // ascending columns
      const int deg = indptr[i + 1] - indptr[i];
      s *= deg > 0 ? 1.f / (float)deg : 0.f;
      const float z = Z[(size_t)i * ld + c], l = z * s;
      const size_t at = (size_t)i * ld + c;
      if (L[at] != s || I[at] != l || C[at] != 5 + (fabsf(l) >= fabsf(O[at]) ? 1 : 0)) {
        printf("R=%d n=%d ld=%d row %d col %d: lag %g/%g local %g/%g cnt %d\n", R, n, ld, i, c, L[at], s, I[at], l, C[at]);
        return 1;
      }
      want_num += (double)z * (double)s;
      want_den += (double)z * (double)z;
    }
    if (fabs(num[c] - want_num) > 1e-9 * (fabs(want_num) + 1.0) || fabs(den[c] - want_den) > 1e-9 * (want_den + 1.0)) {
      printf("R=%d col %d: num %.12g/%.12g den %.12g/%.12g\n", R, c, num[c], want_num, den[c], want_den);
      return 1;
    }
  }
  printf("R=%d Q=%d n=%d g=%d ld=%d k_fixed=%d grid_y=%d: kernel body ok\n", R, Q, n, g, ld, k_fixed, grid_y);
  return 0;
}

int main() {
  int rc = 0;
  rc |= check_kernel_body<2>(1003, 40, 40, 0, 3);
  rc |= check_kernel_body<4>(1003, 37, 40, 0, 5);   // padded columns, ragged rows, n % R != 0
  rc |= check_kernel_body<4>(515, 5, 8, 6, 1);      // narrow matrix, fixed degree, one CTA row
  rc |= check_kernel_body<8>(1003, 70, 96, 0, 2);   // three column blocks, ld = round_up(g, 32)
  rc |= check_kernel_body<8>(64, 33, 40, 6, 7);     // more CTA rows than chunks
  rc |= check_kernel_body<4, 16>(1003, 100, 128, 0, 3);  // wider column blocks
  rc |= check_kernel_body<8, 32>(777, 130, 160, 6, 2);
  for (int n : {1, 7, 64, 1001}) {
    rc |= check<2>(n, 9, 0); rc |= check<4>(n, 9, 0); rc |= check<8>(n, 9, 0);
    if (n > 6) { rc |= check<2>(n, 0, 6); rc |= check<4>(n, 0, 6); rc |= check<8>(n, 0, 6); }
  }
  return rc;
}
