// Lee's L over all ordered gene pairs: L = Aᵀ·B with A = Z, B = W·Z, both cell-major [n, ld]
// (the per-pair statistic of autocorrelation.py:307-315 evaluated for every (x, y) at once).
//
// K (= cells) is huge and M = N = genes is small, so the contraction is split along K: every CTA
// owns one K-range of one 128x128 output tile and writes an FP64 partial tile; a second kernel sums
// the partial tiles per output element in FP64 in fixed order (bitwise reproducible).
//
// Accuracy note: the reference sums with numpy's pairwise reduction, good to ~1e-7 relative.  FP32
// accumulation over K (even in 512-cell sub-chunks folded into FP64) measured 4.7e-4 absolute error
// at N = 10k on entries of magnitude ~20, so the exact path accumulates in FP64.
//
// This file holds the CUDA-core FP64-accumulating kernel (impl 1).  The tcgen05 3xTF32 kernel (impl 2) lives in
// lee_tc.cu and shares the split-K partial layout and the reduction kernel.
#include "common.cuh"
#include "lee.cuh"

namespace sc {

constexpr int kTile = 128;   // output tile edge
constexpr int kKStep = 8;    // cells per shared-memory stage

// 256 threads, each owns an 8x8 block of the 128x128 tile.  Operands are widened to FP64 once, when
// they are staged in shared memory; the inner loop is pure DFMA (B200: ~37 TFLOP/s FP64), so the
// result carries only the FP32 rounding of the inputs.
__global__ void __launch_bounds__(256)
lee_f64_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
               int64_t n, int g, int64_t chunk, double* __restrict__ partial, int64_t ldt) {
  __shared__ __align__(16) double As[2][kKStep][kTile];
  __shared__ __align__(16) double Bs[2][kKStep][kTile];
  const int m0 = blockIdx.y * kTile, n0 = blockIdx.x * kTile;
  const int64_t k_begin = (int64_t)blockIdx.z * chunk;
  const int64_t k_end = min(n, k_begin + chunk);
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid

  // loader mapping: 8 rows x 128 cols = 256 float4 per operand, one per thread
  const int lrow = tid >> 5;          // 0..7
  const int lcol = (tid & 31) * 4;    // 0..124

  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  float4 va, vb;
  auto fetch = [&](int64_t k0) {
    const int64_t kk = k0 + lrow;
    va = make_float4(0.f, 0.f, 0.f, 0.f);
    vb = va;
    if (kk < k_end) {
      if (m0 + lcol < lda) va = ldg4(A + kk * lda + m0 + lcol);
      if (n0 + lcol < ldb) vb = ldg4(B + kk * ldb + n0 + lcol);
    }
  };
  auto stash = [&](int buf) {
    double2* pa = reinterpret_cast<double2*>(&As[buf][lrow][lcol]);
    double2* pb = reinterpret_cast<double2*>(&Bs[buf][lrow][lcol]);
    pa[0] = make_double2((double)va.x, (double)va.y);
    pa[1] = make_double2((double)va.z, (double)va.w);
    pb[0] = make_double2((double)vb.x, (double)vb.y);
    pb[1] = make_double2((double)vb.z, (double)vb.w);
  };

  int buf = 0;
  if (k_begin < k_end) { fetch(k_begin); stash(0); }
  __syncthreads();
  for (int64_t k0 = k_begin; k0 < k_end; k0 += kKStep) {
    const bool more = k0 + kKStep < k_end;
    if (more) fetch(k0 + kKStep);  // global loads in flight during the DFMA block
#pragma unroll
    for (int kk = 0; kk < kKStep; ++kk) {
      double a[8], b[8];
      const double2* pa0 = reinterpret_cast<const double2*>(&As[buf][kk][ty * 4]);
      const double2* pa1 = reinterpret_cast<const double2*>(&As[buf][kk][64 + ty * 4]);
      const double2* pb0 = reinterpret_cast<const double2*>(&Bs[buf][kk][tx * 4]);
      const double2* pb1 = reinterpret_cast<const double2*>(&Bs[buf][kk][64 + tx * 4]);
      double2 t;
      t = pa0[0]; a[0] = t.x; a[1] = t.y; t = pa0[1]; a[2] = t.x; a[3] = t.y;
      t = pa1[0]; a[4] = t.x; a[5] = t.y; t = pa1[1]; a[6] = t.x; a[7] = t.y;
      t = pb0[0]; b[0] = t.x; b[1] = t.y; t = pb0[1]; b[2] = t.x; b[3] = t.y;
      t = pb1[0]; b[4] = t.x; b[5] = t.y; t = pb1[1]; b[6] = t.x; b[7] = t.y;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // partial[z][m][n] (FP64), m/n padded to ldt; this CTA is the only writer of its tile
  double* P = partial + (int64_t)blockIdx.z * ldt * ldt;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      double* dst = P + (int64_t)m * ldt + n0 + jh * 64 + tx * 4;
      *reinterpret_cast<double2*>(dst) = make_double2(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
    }
  }
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
  p.chunk = (int64_t)align_up((size_t)p.chunk, 64);  // multiple of every kernel's K step
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

// cnt[x,y] += (|Lp[x,y]| >= |L[x,y]|): two-tailed exceedance counts of one permuted all-pairs matrix
// (autocorrelation.py:331-332 applied to every pair at once).
__global__ void lee_abs_ge_kernel(const float* __restrict__ Lp, int64_t ldp, const float* __restrict__ L,
                                  int64_t ldl, int g, int32_t* __restrict__ cnt, int64_t ldc) {
  const int y = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
  if (y >= g) return;
  cnt[(int64_t)x * ldc + y] += fabsf(Lp[(int64_t)x * ldp + y]) >= fabsf(L[(int64_t)x * ldl + y]);
}

extern "C" int sc_lee_abs_ge_accumulate(const float* Lp, int64_t ldp, const float* L, int64_t ldl, int g,
                                        int32_t* cnt, int64_t ldc, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(Lp && L && cnt && g >= 1 && ldp >= g && ldl >= g && ldc >= g, "sc_lee_abs_ge_accumulate: bad argument");
  lee_abs_ge_kernel<<<dim3((g + 127) / 128, g), 128, 0, st>>>(Lp, ldp, L, ldl, g, cnt, ldc);
  SC_LAUNCH_OK();
  return SC_OK;
}

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
  // impl 0 = auto: tensor cores (3xTF32, ~2e-6 relative) for contractions big enough to fill them,
  // the exact FP64 CUDA-core kernel otherwise
  if (impl == 0) impl = (n >= 8192 && g >= 64 && lee_tc_supported(n, g, lda, ldb) && !((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15)) ? 2 : 1;
  if (impl == 2) {
    if (!lee_tc_supported(n, g, lda, ldb)) { set_error("sc_lee_gemm: tcgen05 path unsupported for this shape"); return SC_ERR_UNSUPPORTED; }
    char* extra = static_cast<char*>(ws) + align_up(sizeof(double) * (size_t)p.splits * p.ldt * p.ldt, 256);
    return lee_tc_launch(A, lda, B, ldb, n, g, p, partial, extra, L, ldl, st);
  } else {
    dim3 grid(p.ldt / kTile, p.ldt / kTile, p.splits);
    lee_f64_kernel<<<grid, 256, 0, st>>>(A, lda, B, ldb, n, g, p.chunk, partial, p.ldt);
    SC_LAUNCH_OK();
  }
  return lee_reduce(partial, p, g, L, ldl, st);
}
