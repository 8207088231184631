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
#include <stdlib.h>

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

// Per-thread work in lag_group_core.cuh (shared with the host-side test); here: launch geometry and the
// fixed-order reduction of the per-thread Moran sums over the CTA.
template <int R, int Q>
__global__ void __launch_bounds__(kThreads, R >= 8 ? 2 : 3)
lag_group_kernel(const __grid_constant__ LagGroupArgs args, double* __restrict__ partial) {
  constexpr int kSlots = kThreads / Q;
  __shared__ double sh[2][kSlots][Q][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (Q - 1);
  const int slot = warp * (32 / Q) + lane / Q;
  const int64_t col = ((int64_t)blockIdx.x * Q + q) * 4;
  double num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};
  lag_group_thread<R, Q, kThreads>(args, threadIdx.x, blockIdx.x, blockIdx.y, gridDim.y, num, den);
#pragma unroll
  for (int c = 0; c < 4; ++c) { sh[0][slot][q][c] = num[c]; sh[1][slot][q][c] = den[c]; }
  __syncthreads();
  if (slot == 0 && col < args.ldz) {
    double* p = partial + ((int64_t)blockIdx.y * 2) * args.ldz + col;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double s0 = 0, s1 = 0;
#pragma unroll 4
      for (int r = 0; r < kSlots; ++r) { s0 += sh[0][r][q][c]; s1 += sh[1][r][q][c]; }
      p[c] = s0; p[args.ldz + c] = s1;
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

template <int R, int Q>
int launch_group_q(const int32_t* indptr, int k_fixed, const uint32_t* uwords, const int32_t* ucnt, int64_t n,
                 const float* Z, int64_t ldz, int g, float* lag, float* local, int64_t ldl, double* num,
                 double* den, const float* cell_obs, int32_t* cell_cnt, int64_t ldc, double* partial,
                 cudaStream_t st) {
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
  LagGroupArgs args;
  args.indptr = indptr; args.k_fixed = k_fixed; args.uwords = uwords; args.ucnt = ucnt;
  args.n = n; args.n_groups = n_groups; args.Z = Z; args.ldz = ldz; args.lag = lag; args.local = local; args.ldl = ldl;
  args.cell_obs = cell_obs; args.cell_cnt = cell_cnt; args.ldc = ldc; args.n_chunks = n_chunks; args.chunk_groups = chunk_groups;
  lag_group_kernel<R, Q><<<dim3(bx, (unsigned)by), kThreads, 0, st>>>(args, partial);
  SC_LAUNCH_OK();
  group_reduce_kernel<<<(g + 127) / 128, 128, 0, st>>>(partial, (int)by, ldz, g, num, den);
  SC_LAUNCH_OK();
  return SC_OK;
}

// Q = lanes (float4 column quads) per group: 8 = the default kernel's 128-byte column blocks; SC_LAG_GROUP_Q=16|32
// (experiment switch) widens them to 256 / 512 bytes per gathered row piece.
template <int R>
int launch_group(const int32_t* indptr, int k_fixed, const uint32_t* uwords, const int32_t* ucnt, int64_t n,
                 const float* Z, int64_t ldz, int g, float* lag, float* local, int64_t ldl, double* num,
                 double* den, const float* cell_obs, int32_t* cell_cnt, int64_t ldc, double* partial,
                 cudaStream_t st) {
  int q = 8;
  if (const char* e = getenv("SC_LAG_GROUP_Q")) { int v = atoi(e); if (v == 16 || v == 32) q = v; }
  if (ldz < 4 * q) q = 8;
  if (q == 32)
    return launch_group_q<R, 32>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
  if (q == 16)
    return launch_group_q<R, 16>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
  return launch_group_q<R, 8>(indptr, k_fixed, uwords, ucnt, n, Z, ldz, g, lag, local, ldl, num, den, cell_obs, cell_cnt, ldc, partial, st);
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
