// lag_group.cu — row-group spatial lag (EXPERIMENTAL, opt-in through SC_LAG_GROUP).
//
// The lag kernel of stats.cu costs one 128-byte L1 gather per (edge, 32-gene column block) and is bound
// by L1 wavefronts, not HBM.  In spatial order, R consecutive rows share most of their neighbours
// (measured on uniform 2-D points at degree 20: the union of 4 consecutive neighbour lists holds 0.49 of
// their summed lengths, of 2 lists 0.73), so the gathers can be shared: a group of R rows walks the UNION
// of its neighbour lists once, and every gathered float4 is added to the accumulators of the rows that
// own that neighbour (an R-bit membership mask travels in the top bits of the column index).
//
// Replaces the same reference step as sc_csr_lag_moran (`W @ Z`, autocorrelation.py:307, 864, and the
// Moran numerator / denominator of the squidpy call at :576-583) for row-standardised binary graphs.
// Arithmetic: each row's neighbours are summed in ascending column order in FP32 and scaled by 1/deg
// (the default kernel sums them four at a time), so the two kernels agree to FP32 rounding, not bit for bit.
#include <limits.h>

#include "common.cuh"
#include "lag_group_core.cuh"

namespace sc {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

// One thread per group (group_union in lag_group_core.cuh).
template <int R>
__global__ void __launch_bounds__(kThreads)
group_build_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                   int k_fixed, int64_t n_groups, uint32_t* __restrict__ uwords,
                   int32_t* __restrict__ ucnt) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < n_groups;
       a += (int64_t)gridDim.x * blockDim.x)
    ucnt[a] = group_union<R>(indptr, indices, n, k_fixed, a, uwords);
}

__device__ __forceinline__ float4 gather4(const char* base, uint32_t j, uint32_t ld_bytes) {
  return __ldg(reinterpret_cast<const float4*>(base + (uint64_t)j * ld_bytes));
}

// Geometry as lag_stat_kernel: blockIdx.x = column block of Q float4 quads, blockIdx.y strides over chunks
// of `chunk_groups` groups; a thread owns one (group, float4 column quad) and R float4 accumulators.
template <int R, int Q>
__global__ void __launch_bounds__(kThreads, R >= 8 ? 2 : 3)
lag_group_kernel(const int32_t* __restrict__ indptr, int k_fixed, const uint32_t* __restrict__ uwords,
                 const int32_t* __restrict__ ucnt, int64_t n, int64_t n_groups,
                 const float* __restrict__ Z, int64_t ldz, float* __restrict__ lag,
                 float* __restrict__ local, int64_t ldl, double* __restrict__ partial,
                 const float* __restrict__ cell_obs, int32_t* __restrict__ cell_cnt, int64_t ldc,
                 int64_t n_chunks, int chunk_groups) {
  constexpr int kSlots = kThreads / Q;  // groups per pass
  __shared__ double sh[2][kSlots][Q][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (Q - 1);
  const int slot = warp * (32 / Q) + lane / Q;
  const int64_t col = ((int64_t)blockIdx.x * Q + q) * 4;
  const bool active = col < ldz;
  const char* zbase = reinterpret_cast<const char*>(Z + col);
  const uint32_t ldzb = (uint32_t)ldz * 4u;
  double num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};

  for (int64_t chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
    const int64_t g0 = chunk * chunk_groups;
#pragma unroll 1
    for (int pass = 0; pass < chunk_groups; pass += kSlots) {
      const int64_t a = g0 + pass + slot;
      if (a >= n_groups || !active) continue;
      const int64_t row0 = a * R;
      int64_t b0;
      int deg0;
      row_span(indptr, k_fixed, row0, &b0, &deg0);
      const uint32_t* __restrict__ up = uwords + b0;
      const int cnt = ucnt[a];
      float4 acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      int t = 0;
#pragma unroll 1
      for (; t + 4 <= cnt; t += 4) {
        uint32_t w[4];
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = up[t + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = gather4(zbase, word_column<R>(w[u]), ldzb);
#pragma unroll
        for (int u = 0; u < 4; ++u) scatter_add<R>(acc, w[u], v[u]);
      }
#pragma unroll 1
      for (; t < cnt; ++t) {
        const uint32_t w = up[t];
        scatter_add<R>(acc, w, gather4(zbase, word_column<R>(w), ldzb));
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        if (row >= n) break;
        int64_t b;
        int deg;
        row_span(indptr, k_fixed, row, &b, &deg);
        const float inv = (deg > 0) ? 1.f / (float)deg : 0.f;
        float4 s = acc[r];
        s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
        const float4 z = ldg4(Z + row * ldz + col);
        const float4 loc = make_float4(z.x * s.x, z.y * s.y, z.z * s.z, z.w * s.w);
        if (lag) *reinterpret_cast<float4*>(lag + row * ldl + col) = s;
        if (local) *reinterpret_cast<float4*>(local + row * ldl + col) = loc;
        if (cell_cnt) {
          const float4 o = ldg4(cell_obs + row * ldc + col);
          int4* cp = reinterpret_cast<int4*>(cell_cnt + row * ldc + col);
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
#pragma unroll
  for (int c = 0; c < 4; ++c) { sh[0][slot][q][c] = num[c]; sh[1][slot][q][c] = den[c]; }
  __syncthreads();
  if (slot == 0 && active) {
    double* p = partial + ((int64_t)blockIdx.y * 2) * ldz + col;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double s0 = 0, s1 = 0;
#pragma unroll 4
      for (int r = 0; r < kSlots; ++r) { s0 += sh[0][r][q][c]; s1 += sh[1][r][q][c]; }
      p[c] = s0; p[ldz + c] = s1;
    }
  }
}

// out[col] = sum over CTA rows of partial[(b*2 + which)*ld + col], fixed order (bitwise reproducible).
__global__ void group_reduce_kernel(const double* __restrict__ partial, int nblocks, int64_t ld, int g,
                                    double* __restrict__ num, double* __restrict__ den) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= g) return;
  double a = 0, d = 0;
  for (int b = 0; b < nblocks; ++b) {
    a += partial[((int64_t)b * 2) * ld + col];
    d += partial[((int64_t)b * 2 + 1) * ld + col];
  }
  num[col] = a;
  den[col] = d;
}

template <int R>
int launch_group(const int32_t* indptr, int k_fixed, const uint32_t* uwords, const int32_t* ucnt, int64_t n,
                 const float* Z, int64_t ldz, int g, float* lag, float* local, int64_t ldl, double* num,
                 double* den, const float* cell_obs, int32_t* cell_cnt, int64_t ldc, double* partial,
                 cudaStream_t st) {
  constexpr int Q = 8;
  const int64_t n_groups = (n + R - 1) / R;
  constexpr int kSlots = kThreads / Q;  // groups per pass of the kernel: chunks must be multiples of it
  int chunk_groups = 512 / R;           // the default kernel's 512-row chunks
  const int bx = (int)((ldz + 4 * Q - 1) / (4 * Q));
  while (chunk_groups > kSlots && ((n_groups + chunk_groups - 1) / chunk_groups) * bx < 4 * (int64_t)sm_count())
    chunk_groups /= 2;
  static_assert((512 / R) % kSlots == 0, "a chunk must hold whole passes");
  const int64_t n_chunks = (n_groups + chunk_groups - 1) / chunk_groups;
  int64_t by = ((int64_t)sm_count() * 8 + bx - 1) / bx;
  if (by > n_chunks) by = n_chunks;
  if (by > kMaxBlocks) by = kMaxBlocks;
  if (by < 1) by = 1;
  lag_group_kernel<R, Q><<<dim3(bx, (unsigned)by), kThreads, 0, st>>>(
      indptr, k_fixed, uwords, ucnt, n, n_groups, Z, ldz, lag, local, ldl, partial, cell_obs, cell_cnt, ldc,
      n_chunks, chunk_groups);
  SC_LAUNCH_OK();
  group_reduce_kernel<<<(g + 127) / 128, 128, 0, st>>>(partial, (int)by, ldz, g, num, den);
  SC_LAUNCH_OK();
  return SC_OK;
}

}  // namespace
}  // namespace sc

using namespace sc;

extern "C" int sc_graph_group_build(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                                    int group_rows, uint32_t* uwords, int32_t* ucnt, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && uwords && ucnt, "sc_graph_group_build: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_graph_group_build: need indptr or k_fixed");
  SC_CHECK_ARG(group_rows == 2 || group_rows == 4 || group_rows == 8, "sc_graph_group_build: group_rows must be 2, 4 or 8");
  SC_CHECK_ARG(n >= 1 && n <= (1ll << (32 - group_rows)), "sc_graph_group_build: n must be in [1, 2^(32-group_rows)]");
  const int64_t n_groups = (n + group_rows - 1) / group_rows;
  const int64_t want = (n_groups + kThreads - 1) / kThreads;
  const int blocks = (int)(want > 148 * 16 ? 148 * 16 : want);
  if (group_rows == 2) group_build_kernel<2><<<blocks, kThreads, 0, st>>>(indptr, indices, n, k_fixed, n_groups, uwords, ucnt);
  else if (group_rows == 4) group_build_kernel<4><<<blocks, kThreads, 0, st>>>(indptr, indices, n, k_fixed, n_groups, uwords, ucnt);
  else group_build_kernel<8><<<blocks, kThreads, 0, st>>>(indptr, indices, n, k_fixed, n_groups, uwords, ucnt);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_csr_lag_moran_grouped(const int32_t* indptr, int64_t n, int k_fixed, int group_rows,
                                        const uint32_t* uwords, const int32_t* ucnt, const float* Z,
                                        int64_t ldz, int g, float* lag, float* local, int64_t ldl,
                                        double* num, double* den, const float* cell_obs,
                                        int32_t* cell_cnt, int64_t ldc, void* ws, size_t ws_bytes,
                                        sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(uwords && ucnt && Z && num && den && ws, "sc_csr_lag_moran_grouped: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_csr_lag_moran_grouped: need indptr or k_fixed");
  SC_CHECK_ARG(group_rows == 2 || group_rows == 4 || group_rows == 8, "sc_csr_lag_moran_grouped: group_rows must be 2, 4 or 8");
  SC_CHECK_ARG(n >= 1 && n <= (1ll << (32 - group_rows)), "sc_csr_lag_moran_grouped: n must be in [1, 2^(32-group_rows)]");
  SC_CHECK_ARG(ldz % 4 == 0 && ldz >= g && g >= 1 && (size_t)ldz <= align_up((size_t)g, 32),
               "sc_csr_lag_moran_grouped: ldz must be a multiple of 4 in [g, round_up(g,32)]");
  SC_CHECK_ARG((!lag && !local) || (ldl % 4 == 0 && ldl >= ldz), "sc_csr_lag_moran_grouped: ldl must be a multiple of 4 and >= ldz");
  SC_CHECK_ARG((cell_cnt == nullptr) == (cell_obs == nullptr) && (!cell_cnt || (ldc % 4 == 0 && ldc >= ldz)),
               "sc_csr_lag_moran_grouped: cell_obs and cell_cnt go together, ldc a multiple of 4 and >= ldz");
  if (ws_bytes < sc_csr_lag_moran_workspace_bytes(n, g)) { set_error("sc_csr_lag_moran_grouped: workspace too small"); return SC_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  if (group_rows == 2)
    return launch_group<2>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
  if (group_rows == 4)
    return launch_group<4>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
  return launch_group<8>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
}
