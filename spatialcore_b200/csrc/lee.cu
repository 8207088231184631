// Lee's L over all ordered gene pairs: L = Aᵀ·B with A = Z, B = W·Z, both cell-major [n, ld]
// (the per-pair statistic of autocorrelation.py:307-315 evaluated for every (x, y) at once).
//
// K (= cells) is huge and M = N = genes is small, so the contraction is split along K: every CTA
// owns one K-range of one 128x128 output tile.  FP32 accumulation runs only over sub-chunks of
// kLeeSubChunk cells; each finished sub-chunk is folded into the CTA's FP64 partial tile, and a
// second kernel sums the partial tiles per output element in FP64 in fixed order.  The rounding
// error of the contraction then stays near the FP32 rounding of the inputs themselves
// (~1e-5 * sqrt(N) absolute), independent of N.
//
// This file holds the CUDA-core FP32 kernel (impl 1).  The tcgen05 3xTF32 kernel (impl 2) lives in
// lee_tc.cu and shares the split-K partial layout and the reduction kernel.
#include "common.cuh"
#include "lee.cuh"

namespace sc {

constexpr int kTile = 128;   // output tile edge
constexpr int kKStep = 16;   // cells per shared-memory stage

// 256 threads, each owns an 8x8 block of the 128x128 tile.
__global__ void __launch_bounds__(256)
lee_simt_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                int64_t n, int g, int64_t chunk, double* __restrict__ partial, int64_t ldt) {
  __shared__ __align__(16) float As[2][kKStep][kTile];
  __shared__ __align__(16) float Bs[2][kKStep][kTile];
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int64_t k_begin = (int64_t)blockIdx.z * chunk;
  const int64_t k_end = min(n, k_begin + chunk);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid

  // loader mapping: 16 rows x 128 cols = 512 float4 per operand, 2 per thread
  const int lrow = tid >> 5;          // 0..7 (+8)
  const int lcol = (tid & 31) * 4;    // 0..124

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // partial[z][m][n] (FP64), m/n padded to ldt; this CTA is the only writer of its tile
  double* P = partial + (int64_t)blockIdx.z * ldt * ldt;
  auto flush = [&](bool first) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        double* dst = P + (int64_t)m * ldt + n0 + jh * 64 + tx * 4;
        double2 lo = make_double2(0, 0), hi = make_double2(0, 0);
        if (!first) {
          lo = *reinterpret_cast<double2*>(dst);
          hi = *reinterpret_cast<double2*>(dst + 2);
        }
        lo.x += (double)acc[i][jh * 4 + 0]; lo.y += (double)acc[i][jh * 4 + 1];
        hi.x += (double)acc[i][jh * 4 + 2]; hi.y += (double)acc[i][jh * 4 + 3];
        *reinterpret_cast<double2*>(dst) = lo;
        *reinterpret_cast<double2*>(dst + 2) = hi;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][jh * 4 + c] = 0.f;
      }
    }
  };

  auto load_stage = [&](int buf, int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lrow + 8 * h;
      int64_t kk = k0 + r;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (kk < k_end) {
        if (m0 + lcol < lda) va = ldg4(A + kk * lda + m0 + lcol);
        if (n0 + lcol < ldb) vb = ldg4(B + kk * ldb + n0 + lcol);
      }
      *reinterpret_cast<float4*>(&As[buf][r][lcol]) = va;
      *reinterpret_cast<float4*>(&Bs[buf][r][lcol]) = vb;
    }
  };

  int buf = 0;
  bool flushed = false;
  if (k_begin < k_end) load_stage(0, k_begin);
  __syncthreads();
  for (int64_t k0 = k_begin; k0 < k_end; k0 += kKStep) {
    if (k0 + kKStep < k_end) load_stage(buf ^ 1, k0 + kKStep);
#pragma unroll
    for (int kk = 0; kk < kKStep; ++kk) {
      float a[8], b[8];
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    buf ^= 1;
    const int64_t done = k0 + kKStep - k_begin;
    if (done % kLeeSubChunk == 0 && k0 + kKStep < k_end) {
      flush(!flushed);
      flushed = true;
    }
  }
  flush(!flushed);
}

__global__ void lee_reduce_kernel(const double* __restrict__ partial, int splits, int64_t ldt, int g,
                                  float* __restrict__ L, int64_t ldl) {
  int x = blockIdx.y * blockDim.y + threadIdx.y;
  int y = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= g || y >= g) return;
  double s = 0;
  for (int z = 0; z < splits; ++z) s += partial[((int64_t)z * ldt + x) * ldt + y];
  L[(int64_t)x * ldl + y] = (float)s;
}

LeePlan lee_plan(int64_t n, int g) {
  LeePlan p;
  p.ldt = (int)align_up((size_t)g, kTile);
  // enough K-splits to fill the machine (tiles x splits >= ~4 CTAs per SM), each >= one sub-chunk
  int64_t tiles = (int64_t)(p.ldt / kTile) * (p.ldt / kTile);
  int64_t splits = (4 * 148 + tiles - 1) / tiles;
  int64_t max_by_n = (n + kLeeSubChunk - 1) / kLeeSubChunk;
  if (splits > max_by_n) splits = max_by_n;
  if (splits > kLeeMaxSplits) splits = kLeeMaxSplits;
  if (splits < 1) splits = 1;
  p.chunk = (n + splits - 1) / splits;
  p.chunk = (int64_t)align_up((size_t)p.chunk, 64);
  p.splits = (int)((n + p.chunk - 1) / p.chunk);
  return p;
}

int lee_reduce(const double* partial, const LeePlan& p, int g, float* L, int64_t ldl,
               cudaStream_t st) {
  dim3 blk(32, 8);
  dim3 grd((g + 31) / 32, (g + 7) / 8);
  lee_reduce_kernel<<<grd, blk, 0, st>>>(partial, p.splits, p.ldt, g, L, ldl);
  SC_LAUNCH_OK();
  return SC_OK;
}

}  // namespace sc

using namespace sc;

extern "C" size_t sc_lee_gemm_workspace_bytes(int64_t n, int g) {
  LeePlan p = lee_plan(n > 0 ? n : 1, g > 0 ? g : 1);
  return align_up(sizeof(double) * (size_t)p.splits * p.ldt * p.ldt, 256) + lee_tc_extra_workspace_bytes(n, g) + 512;
}

extern "C" int sc_lee_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n,
                           int g, float* L, int64_t ldl, int impl, void* ws, size_t ws_bytes,
                           sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(A && B && L && ws, "sc_lee_gemm: null argument");
  SC_CHECK_ARG(n >= 1 && g >= 1 && lda >= g && ldb >= g && ldl >= g, "sc_lee_gemm: bad shape");
  SC_CHECK_ARG(lda % 4 == 0 && ldb % 4 == 0, "sc_lee_gemm: lda/ldb must be multiples of 4");
  SC_CHECK_ARG(impl >= 0 && impl <= 2, "sc_lee_gemm: impl must be 0, 1 or 2");
  if (ws_bytes < sc_lee_gemm_workspace_bytes(n, g)) { set_error("sc_lee_gemm: workspace too small"); return SC_ERR_WORKSPACE; }
  LeePlan p = lee_plan(n, g);
  double* partial = static_cast<double*>(ws);
  if (impl == 0) impl = lee_tc_supported(n, g, lda, ldb) ? 2 : 1;
  if (impl == 2) {
    if (!lee_tc_supported(n, g, lda, ldb)) { set_error("sc_lee_gemm: tcgen05 path unsupported for this shape"); return SC_ERR_UNSUPPORTED; }
    char* extra = static_cast<char*>(ws) + align_up(sizeof(double) * (size_t)p.splits * p.ldt * p.ldt, 256);
    int rc = lee_tc_launch(A, lda, B, ldb, n, g, p, partial, extra, st);
    if (rc) return rc;
  } else {
    dim3 grid(p.ldt / kTile, p.ldt / kTile, p.splits);
    lee_simt_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, n, g, p.chunk, partial, p.ldt);
    SC_LAUNCH_OK();
  }
  return lee_reduce(partial, p, g, L, ldl, st);
}
