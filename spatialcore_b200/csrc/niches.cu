// K-means on the neighbourhood-profile matrix (identify_niches, [R spatial/neighborhoods.py:299-522];
// the reference calls sklearn.cluster.KMeans: k-means++ seeding + Lloyd iterations).
//
// X is the N x d profile matrix (d = number of cell types, <= 128), FP32 row-major; K <= 64 centres.
// One Lloyd iteration is ONE pass over X (HBM-bound: 4*N*d bytes read + 4*N label bytes):
//   - a CTA stages a tile of 256 rows in shared memory (coalesced, odd row stride: conflict-free),
//   - thread r finds the nearest centre of row r (FP32 differences, first minimum wins like argmin),
//   - thread (k, j) then sums column j of the tile's rows labelled k into an FP64 register that lives
//     across all tiles of the (persistent) CTA: no atomics, fixed summation order,
//   - per-CTA partials are reduced in block order by a second kernel.
// The k-means++ seeding uses three small kernels: candidate potentials, distance update, and
// D^2-sampling by an FP64 prefix sum (CUB) + binary search.
#include <cub/cub.cuh>
#include <float.h>

#include "common.cuh"

namespace sc {

constexpr int kKmThreads = 256;
constexpr int kKmMaxK = 64;
constexpr int kKmMaxD = 128;
constexpr int kKmMaxBlocks = 148 * 4;
constexpr int kKmMaxCand = 16;
constexpr int kKmMaxPairs = 8;    // K*d <= 2048 (sum registers per thread)

__device__ __forceinline__ double block_sum_f64(double v, double* sh /*[8]*/) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0;
  for (int w = 0; w < kKmThreads / 32; ++w) t += sh[w];
  return t;
}

// Stage rows [r0, r0+rows) of X into shared memory with row stride `sd` (odd).
__device__ __forceinline__ void stage_tile(const float* __restrict__ X, int64_t ldx, int d, int64_t r0,
                                           int rows, float* __restrict__ xs, int sd) {
  if (ldx == d) {  // the tile is one contiguous run
    const float* src = X + r0 * ldx;
    for (int t = threadIdx.x; t < rows * d; t += kKmThreads) {
      const int r = t / d, j = t - r * d;
      xs[r * sd + j] = src[t];
    }
  } else {
    for (int t = threadIdx.x; t < rows * d; t += kKmThreads) {
      const int r = t / d, j = t - r * d;
      xs[r * sd + j] = X[(r0 + r) * ldx + j];
    }
  }
}

template <int KMAX>
__global__ void __launch_bounds__(kKmThreads)
kmeans_assign_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int d,
                     const float* __restrict__ C, int K, int32_t* __restrict__ labels,
                     float* __restrict__ mind, double* __restrict__ partial /*[blocks][K*d + K + 2]*/) {
  extern __shared__ __align__(16) unsigned char km_smem[];
  const int sd = d | 1;
  float* cs = reinterpret_cast<float*>(km_smem);           // [K][d]
  float* xs = cs + K * d;                                   // [256][sd]
  int* ls = reinterpret_cast<int*>(xs + kKmThreads * sd);  // [256]
  __shared__ double sh_red[kKmThreads / 32];

  for (int t = threadIdx.x; t < K * d; t += kKmThreads) cs[t] = C[t];

  // sums owned by this thread: pairs p = threadIdx.x + i*256 < K*d  (p = k*d + j)
  constexpr int kPairsMax = kKmMaxPairs;
  const int n_pairs = (K * d + kKmThreads - 1) / kKmThreads;
  double acc[kPairsMax];
#pragma unroll
  for (int i = 0; i < kPairsMax; ++i) acc[i] = 0.0;
  double cnt = 0.0;        // thread k < K counts members of cluster k
  double inertia = 0.0, changed = 0.0;

  const int64_t n_tiles = (n + kKmThreads - 1) / kKmThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kKmThreads;
    const int rows = (int)min((int64_t)kKmThreads, n - r0);
    __syncthreads();
    stage_tile(X, ldx, d, r0, rows, xs, sd);
    __syncthreads();
    if (threadIdx.x < rows) {
      float dist[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) dist[k] = 0.f;
      const float* xr = xs + threadIdx.x * sd;
      for (int j = 0; j < d; ++j) {
        const float xj = xr[j];
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < K) { const float df = xj - cs[k * d + j]; dist[k] = fmaf(df, df, dist[k]); }
      }
      float best = dist[0];
      int bk = 0;
#pragma unroll
      for (int k = 1; k < KMAX; ++k)
        if (k < K && dist[k] < best) { best = dist[k]; bk = k; }
      const int64_t row = r0 + threadIdx.x;
      changed += (labels[row] != bk) ? 1.0 : 0.0;
      labels[row] = bk;
      if (mind) mind[row] = best;
      inertia += (double)best;
      ls[threadIdx.x] = bk;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPairsMax; ++i) {
      const int p = threadIdx.x + i * kKmThreads;
      if (i < n_pairs && p < K * d) {
        const int k = p / d, j = p - k * d;
        double a = 0.0;
        for (int r = 0; r < rows; ++r) a += (ls[r] == k) ? (double)xs[r * sd + j] : 0.0;
        acc[i] += a;
      }
    }
    if (threadIdx.x < K) {
      int c = 0;
      for (int r = 0; r < rows; ++r) c += ls[r] == threadIdx.x;
      cnt += (double)c;
    }
  }
  double* out = partial + (int64_t)blockIdx.x * (K * d + K + 2);
#pragma unroll
  for (int i = 0; i < kPairsMax; ++i) {
    const int p = threadIdx.x + i * kKmThreads;
    if (i < n_pairs && p < K * d) out[p] = acc[i];
  }
  if (threadIdx.x < K) out[K * d + threadIdx.x] = cnt;
  const double tin = block_sum_f64(inertia, sh_red);
  const double tch = block_sum_f64(changed, sh_red);
  if (threadIdx.x == 0) { out[K * d + K] = tin; out[K * d + K + 1] = tch; }
}

// out[t] = sum over blocks of partial[b][t], fixed order.
__global__ void kmeans_reduce_kernel(const double* __restrict__ partial, int blocks, int width,
                                     double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= width) return;
  double s = 0;
  for (int b = 0; b < blocks; ++b) s += partial[(int64_t)b * width + t];
  out[t] = s;
}

// pot[l] = sum_i min(mind[i], ||x_i - x_cand[l]||^2)   (mind == NULL: no min).  When `commit` >= 0 the
// distances to candidate `commit` are also folded into mind (mind[i] = min(mind[i], d)).
__global__ void __launch_bounds__(kKmThreads)
kmeans_pp_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int d,
                 const int32_t* __restrict__ cand, int L, const float* __restrict__ mind_in,
                 float* __restrict__ mind_out, int commit, double* __restrict__ partial /*[blocks][L]*/) {
  extern __shared__ __align__(16) unsigned char km_smem[];
  const int sd = d | 1;
  float* cs = reinterpret_cast<float*>(km_smem);  // [L][d] candidate rows
  float* xs = cs + L * d;                          // [256][sd]
  __shared__ double sh_red[kKmThreads / 32];
  for (int t = threadIdx.x; t < L * d; t += kKmThreads) {
    const int l = t / d, j = t - l * d;
    cs[t] = X[(int64_t)cand[l] * ldx + j];
  }
  double pot[kKmMaxCand];
#pragma unroll
  for (int l = 0; l < kKmMaxCand; ++l) pot[l] = 0.0;
  const int64_t n_tiles = (n + kKmThreads - 1) / kKmThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kKmThreads;
    const int rows = (int)min((int64_t)kKmThreads, n - r0);
    __syncthreads();
    stage_tile(X, ldx, d, r0, rows, xs, sd);
    __syncthreads();
    if (threadIdx.x < rows) {
      float dist[kKmMaxCand];
#pragma unroll
      for (int l = 0; l < kKmMaxCand; ++l) dist[l] = 0.f;
      const float* xr = xs + threadIdx.x * sd;
      for (int j = 0; j < d; ++j) {
        const float xj = xr[j];
#pragma unroll
        for (int l = 0; l < kKmMaxCand; ++l)
          if (l < L) { const float df = xj - cs[l * d + j]; dist[l] = fmaf(df, df, dist[l]); }
      }
      const int64_t row = r0 + threadIdx.x;
      const float m = mind_in ? mind_in[row] : FLT_MAX;
#pragma unroll
      for (int l = 0; l < kKmMaxCand; ++l)
        if (l < L) {
          const float v = fminf(m, dist[l]);
          pot[l] += (double)v;
          if (l == commit) mind_out[row] = v;
        }
    }
  }
  for (int l = 0; l < L; ++l) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < kKmMaxCand; ++q) if (q == l) v = pot[q];
    const double t = block_sum_f64(v, sh_red);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * L + l] = t;
  }
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, int64_t n, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
}

// numpy.searchsorted(cumsum, v, side="left"), clipped to n-1
__global__ void kmeans_search_kernel(const double* __restrict__ cum, int64_t n,
                                     const double* __restrict__ vals, int L, int32_t* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const double v = vals[l];
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (cum[mid] < v) lo = mid + 1; else hi = mid;
  }
  out[l] = (int32_t)(lo < n ? lo : n - 1);
}

static int km_blocks(int64_t n) {
  int64_t tiles = (n + kKmThreads - 1) / kKmThreads;
  int b = sm_count() * 4;
  if (b > kKmMaxBlocks) b = kKmMaxBlocks;
  if (b > tiles) b = (int)tiles;
  return b < 1 ? 1 : b;
}

}  // namespace sc

using namespace sc;

extern "C" size_t sc_kmeans_workspace_bytes(int64_t n, int d, int k) {
  (void)n;
  size_t width = (size_t)k * d + k + 2;
  if (width < (size_t)kKmMaxCand) width = kKmMaxCand;
  return align_up(sizeof(double) * (size_t)kKmMaxBlocks * width, 256) + 512;
}

extern "C" int sc_kmeans_assign(const float* X, int64_t n, int64_t ldx, int d, const float* centers, int k,
                                int32_t* labels, float* mind, double* out, void* ws, size_t ws_bytes,
                                sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(X && centers && labels && out && ws, "sc_kmeans_assign: null argument");
  SC_CHECK_ARG(n >= 1 && d >= 1 && d <= kKmMaxD && ldx >= d, "sc_kmeans_assign: need 1 <= d <= %d and ldx >= d", kKmMaxD);
  SC_CHECK_ARG(k >= 1 && k <= kKmMaxK, "sc_kmeans_assign: need 1 <= k <= %d", kKmMaxK);
  if ((int64_t)k * d > (int64_t)kKmMaxPairs * kKmThreads) { set_error("sc_kmeans_assign: k*d = %d exceeds the compiled limit %d", k * d, kKmMaxPairs * kKmThreads); return SC_ERR_UNSUPPORTED; }
  if (ws_bytes < sc_kmeans_workspace_bytes(n, d, k)) { set_error("sc_kmeans_assign: workspace too small"); return SC_ERR_WORKSPACE; }
  const int blocks = km_blocks(n);
  const int sd = d | 1;
  const size_t smem = sizeof(float) * ((size_t)k * d + (size_t)kKmThreads * sd) + sizeof(int) * kKmThreads;
  double* partial = static_cast<double*>(ws);
  const int width = k * d + k + 2;
#define SC_KM_LAUNCH(KM)                                                                                        \
  do {                                                                                                          \
    SC_CUDA_OK(cudaFuncSetAttribute(kmeans_assign_kernel<KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kmeans_assign_kernel<KM><<<blocks, kKmThreads, smem, st>>>(X, n, ldx, d, centers, k, labels, mind, partial);  \
  } while (0)
  if (k <= 8) SC_KM_LAUNCH(8);
  else if (k <= 16) SC_KM_LAUNCH(16);
  else if (k <= 32) SC_KM_LAUNCH(32);
  else SC_KM_LAUNCH(64);
#undef SC_KM_LAUNCH
  SC_LAUNCH_OK();
  kmeans_reduce_kernel<<<(width + 127) / 128, 128, 0, st>>>(partial, blocks, width, out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_kmeans_pp_potential(const float* X, int64_t n, int64_t ldx, int d, const int32_t* cand,
                                      int n_cand, const float* mind_in, float* mind_out, int commit,
                                      double* pot_out, void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(X && cand && pot_out && ws, "sc_kmeans_pp_potential: null argument");
  SC_CHECK_ARG(n >= 1 && d >= 1 && d <= kKmMaxD && ldx >= d, "sc_kmeans_pp_potential: need 1 <= d <= %d and ldx >= d", kKmMaxD);
  SC_CHECK_ARG(n_cand >= 1 && n_cand <= kKmMaxCand, "sc_kmeans_pp_potential: need 1 <= n_cand <= %d", kKmMaxCand);
  SC_CHECK_ARG(commit < n_cand && (commit < 0 || mind_out), "sc_kmeans_pp_potential: commit needs mind_out");
  if (ws_bytes < sc_kmeans_workspace_bytes(n, d, 1)) { set_error("sc_kmeans_pp_potential: workspace too small"); return SC_ERR_WORKSPACE; }
  const int blocks = km_blocks(n);
  const int sd = d | 1;
  const size_t smem = sizeof(float) * ((size_t)n_cand * d + (size_t)kKmThreads * sd);
  SC_CUDA_OK(cudaFuncSetAttribute(kmeans_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double* partial = static_cast<double*>(ws);
  kmeans_pp_kernel<<<blocks, kKmThreads, smem, st>>>(X, n, ldx, d, cand, n_cand, mind_in, mind_out, commit, partial);
  SC_LAUNCH_OK();
  kmeans_reduce_kernel<<<1, 128, 0, st>>>(partial, blocks, n_cand, pot_out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_kmeans_pp_sample_workspace_bytes(int64_t n) {
  size_t scan = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan, (const double*)nullptr, (double*)nullptr, (int)n);
  return align_up(scan, 256) + 2 * align_up(sizeof(double) * (size_t)(n > 0 ? n : 1), 256) + 512;
}

extern "C" int sc_kmeans_pp_sample(const float* mind, int64_t n, const double* vals, int n_vals,
                                   int32_t* idx_out, void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(mind && vals && idx_out && ws, "sc_kmeans_pp_sample: null argument");
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31) && n_vals >= 1 && n_vals <= kKmMaxCand, "sc_kmeans_pp_sample: bad sizes");
  if (ws_bytes < sc_kmeans_pp_sample_workspace_bytes(n)) { set_error("sc_kmeans_pp_sample: workspace too small"); return SC_ERR_WORKSPACE; }
  size_t scan = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan, (const double*)nullptr, (double*)nullptr, (int)n);
  char* w = static_cast<char*>(ws);
  void* tmp = w;
  double* wide = reinterpret_cast<double*>(w + align_up(scan, 256));
  double* cum = wide + align_up(sizeof(double) * (size_t)n, 256) / sizeof(double);
  int blocks = (int)((n + 255) / 256 > sm_count() * 8 ? sm_count() * 8 : (n + 255) / 256);
  f32_to_f64_kernel<<<blocks, 256, 0, st>>>(mind, n, wide);
  SC_LAUNCH_OK();
  SC_CUDA_OK(cub::DeviceScan::InclusiveSum(tmp, scan, wide, cum, (int)n, st));
  kmeans_search_kernel<<<1, 32, 0, st>>>(cum, n, vals, n_vals, idx_out);
  SC_LAUNCH_OK();
  return SC_OK;
}
