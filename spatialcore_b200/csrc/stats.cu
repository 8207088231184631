// Standardisation, spatial lag + Moran statistic, and the two permutation nulls on sm_100a.
// See include/sc_b200.h for the contract and the reference lines each export replaces.
//
// Layout: all N x G matrices are cell-major (a cell's genes contiguous, ld % 4 == 0).  A warp owns
// one row and 32 consecutive gene quads (128 genes, 512 contiguous bytes per request); a CTA of
// 8 warps covers `wpr` warps per row x 8/wpr rows per pass and strides over the rows persistently.
// Every N-long reduction is accumulated in FP64 per thread, reduced per CTA, written to a partial
// buffer and finished by a second kernel in fixed order (bitwise reproducible, no float atomics).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sc {

constexpr int kStatThreads = 256;
constexpr int kStatWarps = kStatThreads / 32;

// ------------------------------------------------------------------------------------------------
// column statistics and z-scoring
// ------------------------------------------------------------------------------------------------

// blockDim = (128, 2): 128 consecutive columns x 2 row phases; gridDim = (col tiles, row groups).
// Shifted one-pass moments in FP64 (shift = first row's value) -> exact zero variance for constant
// columns, no catastrophic cancellation.
template <typename T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ X, int64_t n, int64_t ldx, int g,
                const int32_t* __restrict__ cols, double* __restrict__ partial) {
  const int col = blockIdx.x * 128 + threadIdx.x;
  const bool active = col < g;
  const int64_t src = active ? (cols ? cols[col] : col) : 0;
  const double shift = active ? (double)X[src] : 0.0;
  double s = 0, ss = 0;
  if (active) {
    const int64_t step = (int64_t)gridDim.y * 2;
    int64_t r = (int64_t)blockIdx.y * 2 + threadIdx.y;
    for (; r + 3 * step < n; r += 4 * step) {
      double v0 = (double)X[r * ldx + src] - shift;
      double v1 = (double)X[(r + step) * ldx + src] - shift;
      double v2 = (double)X[(r + 2 * step) * ldx + src] - shift;
      double v3 = (double)X[(r + 3 * step) * ldx + src] - shift;
      s += v0; ss += v0 * v0;
      s += v1; ss += v1 * v1;
      s += v2; ss += v2 * v2;
      s += v3; ss += v3 * v3;
    }
    for (; r < n; r += step) {
      double v = (double)X[r * ldx + src] - shift;
      s += v; ss += v * v;
    }
  }
  __shared__ double sh[2][128];
  if (threadIdx.y == 1) { sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = ss; }
  __syncthreads();
  if (threadIdx.y == 0 && active) {
    s += sh[0][threadIdx.x];
    ss += sh[1][threadIdx.x];
    double* p = partial + ((int64_t)blockIdx.y * g + col) * 2;
    p[0] = s; p[1] = ss;
  }
}

template <typename T>
__global__ void colstats_final_kernel(const T* __restrict__ X, int64_t n, int g,
                                      const int32_t* __restrict__ cols,
                                      const double* __restrict__ partial, int ngroups,
                                      double* __restrict__ mean, double* __restrict__ std,
                                      uint8_t* __restrict__ zero_var) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= g) return;
  double s = 0, ss = 0;
  for (int j = 0; j < ngroups; ++j) {
    s += partial[((int64_t)j * g + col) * 2];
    ss += partial[((int64_t)j * g + col) * 2 + 1];
  }
  double shift = (double)X[cols ? cols[col] : col];
  double m = s / (double)n;
  double var = ss / (double)n - m * m;
  if (!(var > 0)) var = 0;
  mean[col] = shift + m;
  std[col] = sqrt(var);
  zero_var[col] = var == 0;
}

template <typename T>
__global__ void __launch_bounds__(256)
zscore_write_kernel(const T* __restrict__ X, int64_t n, int64_t ldx, int g,
                    const int32_t* __restrict__ cols, const int32_t* __restrict__ rows,
                    const double* __restrict__ mean, const double* __restrict__ std,
                    const uint8_t* __restrict__ zero_var, float* __restrict__ Z, int64_t ldz) {
  const int64_t Q = ldz / 4;
  const int64_t total = n * Q;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = t / Q;
    const int q = (int)(t - a * Q);
    const int64_t src_row = rows ? rows[a] : a;
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int col = 4 * q + c;
      float z = 0.f;
      if (col < g && !zero_var[col]) {
        double v = (double)X[src_row * ldx + (cols ? cols[col] : col)];
        z = (float)((v - mean[col]) / std[col]);
      }
      o[c] = z;
    }
    *reinterpret_cast<float4*>(Z + a * ldz + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---- fast path: dense FP32 input, all columns, 16-byte aligned rows --------------------------------
// A thread owns one column quad (float4) and walks the rows; a CTA is QW quads wide (QW a power of
// two <= 256) and 256/QW rows deep per pass, so every warp request is a contiguous >= 128-byte run
// and the per-column constants stay in registers.

__global__ void __launch_bounds__(256)
colstats4_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int g, int qw_log2,
                 double* __restrict__ partial) {
  __shared__ double sh[8][256];
  const int qw = 1 << qw_log2;
  const int qx = threadIdx.x & (qw - 1), ry = threadIdx.x >> qw_log2, rpc = 256 >> qw_log2;
  const int col = (blockIdx.x * qw + qx) * 4;
  const bool active = col < g;  // g % 4 may be non-zero: the caller guarantees ldx >= round_up(g,4)
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (active) {
    const float4 sh4 = ldg4(X + col);
    const double shift[4] = {(double)sh4.x, (double)sh4.y, (double)sh4.z, (double)sh4.w};
    const int64_t step = (int64_t)gridDim.y * rpc;
    int64_t r = (int64_t)blockIdx.y * rpc + ry;
    const float* p = X + col;
    for (; r + 3 * step < n; r += 4 * step) {
      const float4 a = ld_stream4(p + r * ldx);
      const float4 b = ld_stream4(p + (r + step) * ldx);
      const float4 c = ld_stream4(p + (r + 2 * step) * ldx);
      const float4 d = ld_stream4(p + (r + 3 * step) * ldx);
      const float4 q[4] = {a, b, c, d};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double v0 = (double)q[u].x - shift[0], v1 = (double)q[u].y - shift[1];
        const double v2 = (double)q[u].z - shift[2], v3 = (double)q[u].w - shift[3];
        s[0] += v0; ss[0] = fma(v0, v0, ss[0]);
        s[1] += v1; ss[1] = fma(v1, v1, ss[1]);
        s[2] += v2; ss[2] = fma(v2, v2, ss[2]);
        s[3] += v3; ss[3] = fma(v3, v3, ss[3]);
      }
    }
    for (; r < n; r += step) {
      const float4 a = ld_stream4(p + r * ldx);
      const double v0 = (double)a.x - shift[0], v1 = (double)a.y - shift[1];
      const double v2 = (double)a.z - shift[2], v3 = (double)a.w - shift[3];
      s[0] += v0; ss[0] = fma(v0, v0, ss[0]);
      s[1] += v1; ss[1] = fma(v1, v1, ss[1]);
      s[2] += v2; ss[2] = fma(v2, v2, ss[2]);
      s[3] += v3; ss[3] = fma(v3, v3, ss[3]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { sh[c][threadIdx.x] = s[c]; sh[4 + c][threadIdx.x] = ss[c]; }
  __syncthreads();
  if (ry == 0 && active) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (col + c >= g) break;
      double a = 0, b = 0;
      for (int r2 = 0; r2 < rpc; ++r2) { a += sh[c][r2 * qw + qx]; b += sh[4 + c][r2 * qw + qx]; }
      double* dst = partial + ((int64_t)blockIdx.y * g + col + c) * 2;
      dst[0] = a; dst[1] = b;
    }
  }
}

__global__ void __launch_bounds__(256)
zscore_write4_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int g,
                     const int32_t* __restrict__ rows, const double* __restrict__ mean,
                     const double* __restrict__ std, const uint8_t* __restrict__ zero_var,
                     float* __restrict__ Z, int64_t ldz, int qw_log2) {
  const int qw = 1 << qw_log2;
  const int qx = threadIdx.x & (qw - 1), ry = threadIdx.x >> qw_log2, rpc = 256 >> qw_log2;
  const int col = (blockIdx.x * qw + qx) * 4;
  if (col >= ldz) return;
  double m[4], inv[4];
  bool live[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    live[c] = col + c < g && !zero_var[col + c];  // padding / zero-variance columns -> exactly 0
    m[c] = live[c] ? mean[col + c] : 0.0;
    inv[c] = live[c] ? 1.0 / std[col + c] : 0.0;
  }
  const bool in_x = col < g;  // a quad entirely in the padding has nothing to read
  const int64_t step = (int64_t)gridDim.y * rpc;
  for (int64_t a = (int64_t)blockIdx.y * rpc + ry; a < n; a += step) {
    const int64_t src = rows ? rows[a] : a;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_x) v = ld_stream4(X + src * ldx + col);
    float4 o;
    o.x = live[0] ? (float)(((double)v.x - m[0]) * inv[0]) : 0.f;
    o.y = live[1] ? (float)(((double)v.y - m[1]) * inv[1]) : 0.f;
    o.z = live[2] ? (float)(((double)v.z - m[2]) * inv[2]) : 0.f;
    o.w = live[3] ? (float)(((double)v.w - m[3]) * inv[3]) : 0.f;
    *reinterpret_cast<float4*>(Z + a * ldz + col) = o;
  }
}

// Fused standardise + all-gather + spatial re-order over NVLink peer memory (row-sharded ingest): this
// GPU z-scores ITS block of cells once and stores every output row into the Z matrix of EVERY GPU
// (peers[] are the peer-mapped base addresses of the symmetric Z buffers, peers[self] the local one) at
// the row's position in the spatial order, dst_rows[a].  No staging copy, no separate collective, no
// re-order pass: the NVLink stores overlap the streaming read of X.
constexpr int kMaxPeers = 16;
struct PeerPtrs { float* p[kMaxPeers]; };

__global__ void __launch_bounds__(256)
zscore_scatter4_kernel(const float* __restrict__ X, int64_t n, int64_t ldx, int g,
                       const int32_t* __restrict__ dst_rows, const double* __restrict__ mean,
                       const double* __restrict__ std, const uint8_t* __restrict__ zero_var,
                       const __grid_constant__ PeerPtrs peers, int n_peers, int64_t ldz, int qw_log2) {
  const int qw = 1 << qw_log2;
  const int qx = threadIdx.x & (qw - 1), ry = threadIdx.x >> qw_log2, rpc = 256 >> qw_log2;
  const int col = (blockIdx.x * qw + qx) * 4;
  if (col >= ldz) return;
  double m[4], inv[4];
  bool live[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    live[c] = col + c < g && !zero_var[col + c];
    m[c] = live[c] ? mean[col + c] : 0.0;
    inv[c] = live[c] ? 1.0 / std[col + c] : 0.0;
  }
  const bool in_x = col < g;
  const int64_t step = (int64_t)gridDim.y * rpc;
  for (int64_t a = (int64_t)blockIdx.y * rpc + ry; a < n; a += step) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in_x) v = ld_stream4(X + a * ldx + col);
    float4 o;
    o.x = live[0] ? (float)(((double)v.x - m[0]) * inv[0]) : 0.f;
    o.y = live[1] ? (float)(((double)v.y - m[1]) * inv[1]) : 0.f;
    o.z = live[2] ? (float)(((double)v.z - m[2]) * inv[2]) : 0.f;
    o.w = live[3] ? (float)(((double)v.w - m[3]) * inv[3]) : 0.f;
    const int64_t off = (int64_t)dst_rows[a] * ldz + col;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p)
      if (p < n_peers) *reinterpret_cast<float4*>(peers.p[p] + off) = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
csr_densify_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const T* __restrict__ data, int64_t n, const int32_t* __restrict__ colmap,
                   int g_out, float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += warps) {
    const int64_t b = indptr[i], e = indptr[i + 1];
    for (int64_t t = b + lane; t < e; t += 32) {
      int c = indices[t];
      int oc = colmap ? colmap[c] : c;
      if (oc >= 0 && oc < g_out) out[i * ldo + oc] = (float)data[t];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// shared thread geometry of the row-streaming kernels
// ------------------------------------------------------------------------------------------------

struct RowGeom {
  int wpr;  // warps per row (1,2,4,8)
  int rpp;  // rows per CTA pass = 8 / wpr
};

static RowGeom row_geom(int64_t ld) {
  int64_t quads = ld / 4;
  RowGeom rg;
  rg.wpr = quads > 128 ? 8 : quads > 64 ? 4 : quads > 32 ? 2 : 1;
  rg.rpp = kStatWarps / rg.wpr;
  return rg;
}

// Sum per-thread FP64 values over the CTA's row phases: threads with the same (warp % wpr, lane)
// own the same columns.  Result valid in row phase 0.
__device__ __forceinline__ void reduce_row_phases(double (&v)[4], int wpr, double* sh /*[8*32*4]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rphase = warp / wpr;
  __syncthreads();
  if (rphase > 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) sh[(warp * 32 + lane) * 4 + c] = v[c];
  }
  __syncthreads();
  if (rphase == 0) {
    for (int w2 = warp + wpr; w2 < kStatWarps; w2 += wpr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] += sh[(w2 * 32 + lane) * 4 + c];
    }
  }
}

// out[r*ld_out + col] = Σ_b partial[(b*rows + r)*ldp + col], fixed order.
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nblocks, int rows,
                                       int64_t ldp, int g, double* __restrict__ out,
                                       int64_t ld_out) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (col >= g) return;
  double s = 0;
  for (int b = 0; b < nblocks; ++b) s += partial[((int64_t)b * rows + r) * ldp + col];
  out[(int64_t)r * ld_out + col] = s;
}

// ------------------------------------------------------------------------------------------------
// spatial lag + Moran numerator / denominator
// ------------------------------------------------------------------------------------------------

// Geometry: a CTA owns contiguous chunks of `chunk_rows` rows and one column block of 32 genes
// (8 lanes x float4 per row, 32 rows per pass).  With the cells in spatial order the neighbour
// rows of a chunk form a small working set (~1.7x the chunk) that stays in L1, so Z is read from HBM
// about once instead of once per edge.  blockIdx.x = column block (fastest: the column blocks of one
// chunk run together and share the CSR indices through L2), blockIdx.y = chunk group.
//
// The kernel is bound by L1 wavefronts and instruction issue, not HBM (one FADD/FFMA per 4 gathered
// bytes), so the inner loop is kept lean: 32-bit edge counters, one IMAD.WIDE per gathered row.
// The Moran sums stay FP64 sums of the exact
// FP32 products (identical arithmetic to the permutation kernels, so the identity permutation
// reproduces the observed statistic to round-off).
constexpr int kLagColQuads = 8;
constexpr int kLagDefaultUnr = 4;
constexpr int kLagRowsPerPass = kStatThreads / kLagColQuads;  // 32

__device__ __forceinline__ float4 ldg4_row(const char* base, int j, uint32_t ld_bytes) {
  return __ldg(reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)j * ld_bytes));
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& v) {
  a.x += w * v.x; a.y += w * v.y; a.z += w * v.z; a.w += w * v.w;
}

// Fallback lag kernel: one (row, float4 lane) per thread, neighbour rows gathered through L1.  It serves
// explicitly weighted graphs, graphs without the tile form and matrices narrower than 32 columns; the
// shared-memory tile kernel (lag_tile.cu) is the fast path.  Measured on B200 (C4: 33 ms, C2: 1.45 ms):
// bound by the L1 load path (~46 B/clk/SM for 128-byte row pieces), insensitive to row alignment, lanes per
// row (8/16/32), chunk size and unroll depth (profiles/r02_lag_experiments.json).
// A row's neighbours are added one by one in ascending column order (the order of lag_tile.cu, so the two
// kernels agree bit for bit); four gathers are in flight per thread.
template <bool HAS_W, int UNR, int Q>
__global__ void __launch_bounds__(kStatThreads, 4)
lag_stat_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                const float* __restrict__ weights, int64_t n, int k_fixed,
                const float* __restrict__ Zself, const float* __restrict__ Zlag, int64_t ldz,
                float* __restrict__ lag, float* __restrict__ local, int64_t ldl,
                double* __restrict__ partial, const float* __restrict__ cell_obs,
                int32_t* __restrict__ cell_cnt, int64_t ldc, int64_t n_chunks, int chunk_rows) {
  constexpr int kRows = kStatThreads / Q;  // rows per pass
  __shared__ double sh[2][kRows][Q][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & (Q - 1);
  const int rslot = warp * (32 / Q) + lane / Q;
  const int64_t col = ((int64_t)blockIdx.x * Q + q) * 4;
  const bool active = col < ldz;
  const float* Zs = Zself ? Zself : Zlag;
  const char* zbase = reinterpret_cast<const char*>(Zlag + col);
  const uint32_t ldzb = (uint32_t)ldz * 4u;
  double num[4] = {0, 0, 0, 0}, den[4] = {0, 0, 0, 0};  // FP64 sums of the exact FP32 products

  for (int64_t chunk = blockIdx.y; chunk < n_chunks; chunk += gridDim.y) {
    const int64_t r0 = chunk * chunk_rows;
#pragma unroll 1
    for (int pass = 0; pass < chunk_rows; pass += kRows) {
      const int64_t row = r0 + pass + rslot;
      if (row >= n || !active) continue;
      int64_t b;
      int deg;
      if (indptr) { b = indptr[row]; deg = indptr[row + 1] - (int)b; } else { b = row * k_fixed; deg = k_fixed; }
      const int32_t* __restrict__ ip = indices + b;
      const float* __restrict__ wp = HAS_W ? weights + b : nullptr;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int t = 0;
#pragma unroll 1
      for (; t + UNR <= deg; t += UNR) {
        int j[UNR];
        float w[UNR];
        float4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) { j[u] = ip[t + u]; w[u] = HAS_W ? wp[t + u] : 1.f; }
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = ldg4_row(zbase, j[u], ldzb);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (HAS_W) fma4(acc, w[u], v[u]);
          else { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
      }
#pragma unroll 1
      for (; t < deg; ++t) {
        const float4 v = ldg4_row(zbase, ip[t], ldzb);
        if (HAS_W) fma4(acc, wp[t], v);
        else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
      }
      if (!HAS_W) {
        const float inv = (deg > 0) ? 1.f / (float)deg : 0.f;
        acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      }
      const float4 z = ldg4(Zs + row * ldz + col);
      const float4 loc = make_float4(z.x * acc.x, z.y * acc.y, z.z * acc.z, z.w * acc.w);
      if (lag) *reinterpret_cast<float4*>(lag + row * ldl + col) = acc;
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
      num[0] = fma(zx, (double)acc.x, num[0]); den[0] = fma(zx, zx, den[0]);
      num[1] = fma(zy, (double)acc.y, num[1]); den[1] = fma(zy, zy, den[1]);
      num[2] = fma(zz, (double)acc.z, num[2]); den[2] = fma(zz, zz, den[2]);
      num[3] = fma(zw, (double)acc.w, num[3]); den[3] = fma(zw, zw, den[3]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) { sh[0][rslot][q][c] = num[c]; sh[1][rslot][q][c] = den[c]; }
  __syncthreads();
  if (rslot == 0 && active) {
    double* p = partial + ((int64_t)blockIdx.y * 2) * ldz + col;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double a = 0, d = 0;
#pragma unroll 4
      for (int r = 0; r < kRows; ++r) { a += sh[0][r][q][c]; d += sh[1][r][q][c]; }
      p[c] = a; p[ldz + c] = d;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// permutation sources
// ------------------------------------------------------------------------------------------------

constexpr int kMaxPermBatch = 16;

struct PermBatch {
  int source;               // sc_perm_source
  int count;                // permutations in this launch (<= template PB)
  const int32_t* replay;    // [count, n] when source == REPLAY (already offset to this batch)
  PermDomain dom;
  uint32_t keys[kMaxPermBatch][kFeistelRounds];
};

__device__ __forceinline__ int perm_lookup(const PermBatch& pb, const uint32_t (*s_keys)[kFeistelRounds],
                                           int p, int64_t i, int64_t n) {
  if (pb.source == SC_PERM_REPLAY) return pb.replay[(int64_t)p * n + i];
  return (int)perm_apply((uint32_t)i, pb.dom, s_keys[p]);
}

// ------------------------------------------------------------------------------------------------
// graph-row null: sims[p,c] = Σ_i A[i,c] · B[π_p(i),c]
// ------------------------------------------------------------------------------------------------

template <int PB, typename AccT>
__global__ void __launch_bounds__(kStatThreads)
perm_rows_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                 int64_t ldb, int64_t n, const __grid_constant__ PermBatch pb,
                 double* __restrict__ partial, int64_t ldp, int wpr) {
  __shared__ double sh[kStatWarps * 32 * 4];
  __shared__ uint32_t s_keys[kMaxPermBatch][kFeistelRounds];
  for (int t = threadIdx.x; t < kMaxPermBatch * kFeistelRounds; t += blockDim.x)
    s_keys[t / kFeistelRounds][t % kFeistelRounds] = pb.keys[t / kFeistelRounds][t % kFeistelRounds];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rphase = warp / wpr, rpp = kStatWarps / wpr;
  const int64_t col = ((int64_t)blockIdx.y * 256 + (warp % wpr) * 32 + lane) * 4;
  const bool active = col < lda;
  const int count = pb.count;

  AccT acc[PB][4];
#pragma unroll
  for (int p = 0; p < PB; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0;

  for (int64_t row = (int64_t)blockIdx.x * rpp + rphase; row < n; row += (int64_t)gridDim.x * rpp) {
    // lane p resolves π_p(row); broadcast by shuffle (one evaluation per warp, not per thread)
    int src_l = 0;
    if (lane < count) src_l = perm_lookup(pb, s_keys, lane, row, n);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) a = ld_stream4(A + row * lda + col);
    float4 v[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      int src = __shfl_sync(kFull, src_l, p);
      v[p] = (active && p < count) ? ldg4(B + (int64_t)src * ldb + col)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      acc[p][0] += (AccT)a.x * (AccT)v[p].x;
      acc[p][1] += (AccT)a.y * (AccT)v[p].y;
      acc[p][2] += (AccT)a.z * (AccT)v[p].z;
      acc[p][3] += (AccT)a.w * (AccT)v[p].w;
    }
  }
#pragma unroll
  for (int p = 0; p < PB; ++p) {
    double v4[4] = {(double)acc[p][0], (double)acc[p][1], (double)acc[p][2], (double)acc[p][3]};
    reduce_row_phases(v4, wpr, sh);
    if (rphase == 0 && active && p < count) {
      double* dst = partial + ((int64_t)blockIdx.x * PB + p) * ldp + col;
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = v4[c];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// graph-row null, bulk-async pipeline (the production kernel)
//
// The gather `B[π_p(i), :]` moves whole rows (ld*4 contiguous bytes) from random places in HBM.
// Holding those loads in registers caps the bytes in flight per SM well below what HBM3e needs
// (Little: ~6.5 TB/s x ~1 us = ~45 KB per SM just to break even).  Here one producer warp resolves
// the permutation indices and issues one `cp.async.bulk` (TMA 1-D bulk copy, SASS UBLKCP) per row
// into a shared-memory ring of `stages` stages, completion tracked by mbarrier transaction counts;
// eight consumer warps read the staged rows conflict-free and accumulate in FP64.  In-flight bytes
// per SM = (stages-1) x rows-per-stage x (PB+1) x ld x 4, i.e. 100-200 KB.
// ------------------------------------------------------------------------------------------------

// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

constexpr int kBulkConsumerWarps = 8;
constexpr int kBulkProducerWarps = 4;  // one warpgroup; producer w owns pipeline iterations it = w (mod 4)
constexpr int kBulkThreads = (kBulkConsumerWarps + kBulkProducerWarps) * 32;
constexpr int kBulkMaxStages = 8;

template <int PB>
__global__ void __launch_bounds__(kBulkThreads, 1)
perm_rows_bulk_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
                      int64_t ldb, int64_t n, int cols /*floats per staged row segment*/,
                      const __grid_constant__ PermBatch pb, double* __restrict__ partial,
                      int64_t ldp, int wpr, int stages, int n_producers) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  __shared__ uint64_t full_bar[kBulkMaxStages], empty_bar[kBulkMaxStages];
  __shared__ uint32_t s_keys[kMaxPermBatch][kFeistelRounds];
  __shared__ double sh[kStatWarps * 32 * 4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rpp = kBulkConsumerWarps / wpr;
  const int count = pb.count;
  const int64_t col0 = (int64_t)blockIdx.y * 1024;  // first column of this CTA's column block
  // the last column block of a matrix whose width is not a multiple of 1024 is narrower: it stages (and
  // expects) fewer bytes per row; the stage layout is per CTA, so nothing else changes
  if (lda - col0 < cols) cols = (int)(lda - col0);
  const uint32_t row_bytes = (uint32_t)cols * 4u;
  const uint32_t unit_bytes = (uint32_t)(PB + 1) * row_bytes;
  const uint32_t stage_bytes = (uint32_t)rpp * unit_bytes;
  const int64_t n_groups = (n + rpp - 1) / rpp;

  for (int t = threadIdx.x; t < kMaxPermBatch * kFeistelRounds; t += blockDim.x)
    s_keys[t / kFeistelRounds][t % kFeistelRounds] = pb.keys[t / kFeistelRounds][t % kFeistelRounds];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kBulkConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kBulkConsumerWarps) {
    // ===== producer warps: permutation lookup + bulk copies =====================================
    // The Feistel evaluation is a ~1000-cycle dependent chain per stage; up to four warps
    // interleave the pipeline iterations so index generation never paces the copies.
    // INVARIANT: n_producers divides stages, so a given stage is always refilled by the same warp,
    // strictly one phase after its own previous fill.  (With rotating owners a fast warp could reach
    // a stage two phases early and the 1-bit mbarrier parity test would let it through.)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    const uint64_t pol = policy_evict_first();
    const int pw = warp - kBulkConsumerWarps;
    int it = pw;
    for (int64_t grp = blockIdx.x + (int64_t)pw * gridDim.x; pw < n_producers && grp < n_groups;
         grp += (int64_t)n_producers * gridDim.x, it += n_producers) {
      const int stage = it % stages;
      const uint32_t phase = (uint32_t)(it / stages) & 1u;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      const int64_t row0 = grp * rpp;
      const int nslots = (int)min((int64_t)rpp, n - row0);
      const int per_slot = count + 1;
      if (lane == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(nslots * per_slot) * row_bytes);
      __syncwarp();
      unsigned char* sbase = dyn_smem + (size_t)stage * stage_bytes;
      for (int j = lane; j < nslots * per_slot; j += 32) {
        const int slot = j / per_slot;
        const int p = j - slot * per_slot - 1;  // -1: the streamed A row
        const int64_t row = row0 + slot;
        const float* src;
        if (p < 0) {
          src = A + row * lda + col0;
        } else {
          const int64_t srow = perm_lookup(pb, s_keys, p, row, n);
          src = B + srow * ldb + col0;
        }
        bulk_g2s(sbase + (size_t)slot * unit_bytes + (size_t)(p + 1) * row_bytes, src, row_bytes,
                 &full_bar[stage], pol);
      }
    }
  } else {
    // ===== consumer warps ========================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    const int slot = warp / wpr;
    const int c4 = ((warp % wpr) * 32 + lane) * 4;  // column within the staged row segment
    const bool active = c4 < cols;
    double acc[PB][4];
#pragma unroll
    for (int p = 0; p < PB; ++p)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[p][c] = 0.0;

    int it = 0;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
      const int stage = it % stages;
      const uint32_t phase = (uint32_t)(it / stages) & 1u;
      mbar_wait(&full_bar[stage], phase);
      const int64_t row = grp * rpp + slot;
      if (active && row < n) {
        const float* u = reinterpret_cast<const float*>(dyn_smem + (size_t)stage * stage_bytes +
                                                        (size_t)slot * unit_bytes) + c4;
        const float4 a = *reinterpret_cast<const float4*>(u);
        const double ax = a.x, ay = a.y, az = a.z, aw = a.w;
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          if (p < count) {
            const float4 v = *reinterpret_cast<const float4*>(u + (size_t)(p + 1) * cols);
            acc[p][0] = fma(ax, (double)v.x, acc[p][0]);
            acc[p][1] = fma(ay, (double)v.y, acc[p][1]);
            acc[p][2] = fma(az, (double)v.z, acc[p][2]);
            acc[p][3] = fma(aw, (double)v.w, acc[p][3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
    }

    // reduce over row slots among the 256 consumer threads (named barrier 1), write partials
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (slot > 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) sh[(warp * 32 + lane) * 4 + c] = acc[p][c];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (slot == 0 && active && p < count) {
        double v4[4] = {acc[p][0], acc[p][1], acc[p][2], acc[p][3]};
        for (int w2 = warp + wpr; w2 < kBulkConsumerWarps; w2 += wpr)
#pragma unroll
          for (int c = 0; c < 4; ++c) v4[c] += sh[(w2 * 32 + lane) * 4 + c];
        double* dst = partial + ((int64_t)blockIdx.x * PB + p) * ldp + col0 + c4;
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = v4[c];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// value-permuting null (gather-SpMM):
//   sims[p,c] = Σ_i S_p[i,c] · Σ_j w_ij Zy[π_p(j),c],  S_p[i] = Zx ? Zx[i] : Zy[π_p(i)]
// ------------------------------------------------------------------------------------------------

template <int PB, bool HAS_W>
__global__ void __launch_bounds__(kStatThreads)
perm_values_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   const float* __restrict__ weights, int64_t n, int k_fixed,
                   const float* __restrict__ Zx, const float* __restrict__ Zy, int64_t ldz,
                   const __grid_constant__ PermBatch pb, double* __restrict__ partial, int64_t ldp,
                   const float* __restrict__ cell_obs, int32_t* __restrict__ cell_cnt, int64_t ldc,
                   int wpr) {
  __shared__ double sh[kStatWarps * 32 * 4];
  __shared__ uint32_t s_keys[kMaxPermBatch][kFeistelRounds];
  for (int t = threadIdx.x; t < kMaxPermBatch * kFeistelRounds; t += blockDim.x)
    s_keys[t / kFeistelRounds][t % kFeistelRounds] = pb.keys[t / kFeistelRounds][t % kFeistelRounds];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rphase = warp / wpr, rpp = kStatWarps / wpr;
  const int64_t col = ((int64_t)blockIdx.y * 256 + (warp % wpr) * 32 + lane) * 4;
  const bool active = col < ldz;
  const int count = pb.count;

  double acc[PB][4];
#pragma unroll
  for (int p = 0; p < PB; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0;

  for (int64_t row = (int64_t)blockIdx.x * rpp + rphase; row < n; row += (int64_t)gridDim.x * rpp) {
    const int64_t b = indptr ? indptr[row] : row * k_fixed;
    const int64_t e = indptr ? indptr[row + 1] : b + k_fixed;
    int self_l = 0;
    if (!Zx && lane < count) self_l = perm_lookup(pb, s_keys, lane, row, n);
    float4 lagp[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) lagp[p] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t t0 = b; t0 < e; t0 += 32) {
      const int cnt = (int)min((int64_t)32, e - t0);
      // lane j resolves neighbour t0+j under every permutation of the batch
      int nb_l[PB];
      float w_l = 1.f;
#pragma unroll
      for (int p = 0; p < PB; ++p) nb_l[p] = 0;
      if (lane < cnt) {
        int j = indices[t0 + lane];
        if (HAS_W) w_l = weights[t0 + lane];
#pragma unroll
        for (int p = 0; p < PB; ++p)
          if (p < count) nb_l[p] = perm_lookup(pb, s_keys, p, j, n);
      }
      for (int jj = 0; jj < cnt; ++jj) {
        float w = HAS_W ? __shfl_sync(kFull, w_l, jj) : 1.f;
#pragma unroll
        for (int p = 0; p < PB; ++p) {
          int src = __shfl_sync(kFull, nb_l[p], jj);
          if (active && p < count) {
            float4 v = ldg4(Zy + (int64_t)src * ldz + col);
            lagp[p].x += w * v.x; lagp[p].y += w * v.y; lagp[p].z += w * v.z; lagp[p].w += w * v.w;
          }
        }
      }
    }
    const float inv = HAS_W ? 1.f : ((e > b) ? 1.f / (float)(e - b) : 0.f);
    int4 hits = make_int4(0, 0, 0, 0);
    float4 obs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cell_cnt && active) obs = ldg4(cell_obs + row * ldc + col);
#pragma unroll
    for (int p = 0; p < PB; ++p) {
      int self = __shfl_sync(kFull, self_l, p);
      if (active && p < count) {
        float4 s = Zx ? ldg4(Zx + row * ldz + col) : ldg4(Zy + (int64_t)self * ldz + col);
        float4 l = make_float4(lagp[p].x * inv, lagp[p].y * inv, lagp[p].z * inv, lagp[p].w * inv);
        acc[p][0] += (double)s.x * (double)l.x;
        acc[p][1] += (double)s.y * (double)l.y;
        acc[p][2] += (double)s.z * (double)l.z;
        acc[p][3] += (double)s.w * (double)l.w;
        if (cell_cnt) {
          hits.x += fabsf(s.x * l.x) >= fabsf(obs.x);
          hits.y += fabsf(s.y * l.y) >= fabsf(obs.y);
          hits.z += fabsf(s.z * l.z) >= fabsf(obs.z);
          hits.w += fabsf(s.w * l.w) >= fabsf(obs.w);
        }
      }
    }
    if (cell_cnt && active) {
      int4* cp = reinterpret_cast<int4*>(cell_cnt + row * ldc + col);
      int4 c = *cp;
      c.x += hits.x; c.y += hits.y; c.z += hits.z; c.w += hits.w;
      *cp = c;
    }
  }
#pragma unroll
  for (int p = 0; p < PB; ++p) {
    double v4[4] = {acc[p][0], acc[p][1], acc[p][2], acc[p][3]};
    reduce_row_phases(v4, wpr, sh);
    if (rphase == 0 && active && p < count) {
      double* dst = partial + ((int64_t)blockIdx.x * PB + p) * ldp + col;
#pragma unroll
      for (int c = 0; c < 4; ++c) dst[c] = v4[c];
    }
  }
}

// out[a, :] = Z[π(a), :]  (materialised permuted copy for the value-permuting null).  Warp per row;
// lane 0 resolves the index once.
__global__ void __launch_bounds__(256)
permute_rows_kernel(const float* __restrict__ Z, int64_t ld, int64_t n,
                    const __grid_constant__ PermBatch pb, float* __restrict__ out) {
  __shared__ uint32_t s_keys[kMaxPermBatch][kFeistelRounds];
  for (int t = threadIdx.x; t < kMaxPermBatch * kFeistelRounds; t += blockDim.x)
    s_keys[t / kFeistelRounds][t % kFeistelRounds] = pb.keys[t / kFeistelRounds][t % kFeistelRounds];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t Q = ld / 4;
  for (int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < n; a += warps) {
    int src = 0;
    if (lane == 0) src = perm_lookup(pb, s_keys, 0, a, n);
    src = __shfl_sync(kFull, src, 0);
    const float4* in = reinterpret_cast<const float4*>(Z + (int64_t)src * ld);
    float4* o = reinterpret_cast<float4*>(out + a * ld);
    for (int64_t qq = lane; qq < Q; qq += 32) o[qq] = __ldg(in + qq);
  }
}

// dst[a, 0:cols) = src[rows[a], 0:cols)  (cols % 4 == 0).  Warp per row.
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, int64_t lds, int64_t n, int64_t cols,
                   const int32_t* __restrict__ rows, float* __restrict__ dst, int64_t ldd) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t Q = cols / 4;
  for (int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < n; a += warps) {
    const float4* in = reinterpret_cast<const float4*>(src + (int64_t)rows[a] * lds);
    float4* o = reinterpret_cast<float4*>(dst + a * ldd);
    for (int64_t qq = lane; qq < Q; qq += 32) o[qq] = __ldg(in + qq);
  }
}

// out[p, a] = rank[perm[p, order[a]]]: a permutation of cell ids re-expressed on sorted positions.
__global__ void perm_conjugate_kernel(const int32_t* __restrict__ perm, int64_t n, int n_perms,
                                      const int32_t* __restrict__ order,
                                      const int32_t* __restrict__ rank, int32_t* __restrict__ out) {
  const int64_t total = n * n_perms;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = t / n, a = t - p * n;
    out[t] = rank[perm[p * n + order[a]]];
  }
}

// ------------------------------------------------------------------------------------------------
// local Moran epilogue: per-cell p-values, multiple-testing adjustment and LISA quadrants on the device
// (replaces the reference's N x G Python loop and the per-gene numpy sorts of
// [R autocorrelation.py:132-183, 219-265, 888-928]).
//
// A per-cell permutation p-value takes only P+1 values, (c+1)/(P+1), so the Benjamini-Hochberg
// step-up over the N cells of a gene needs no sort: with h[v] = number of cells at level v and
// cum[v] = sum_{u<=v} h[u], every cell at level v gets  min_{u>=v, h[u]>0} p_u*N/cum[u]  -- exactly what
// sorting, dividing by the rank and taking the running minimum from the end produces (tied cells share
// the value of the last of them).  The arithmetic follows numpy's: p*N in FP32, the division in FP64.
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
local_hist_kernel(const int32_t* __restrict__ cnt, int64_t ldc, int64_t n, int g, int n_perms,
                  const uint8_t* __restrict__ zero_var, int genes_per_block,
                  unsigned int* __restrict__ hist /*[g][P+1]*/) {
  extern __shared__ unsigned int sh_hist[];  // [genes_per_block][P+1]
  const int levels = n_perms + 1;
  const int g0 = blockIdx.x * genes_per_block;
  const int gb = min(genes_per_block, g - g0);
  for (int t = threadIdx.x; t < gb * levels; t += blockDim.x) sh_hist[t] = 0;
  __syncthreads();
  const int64_t total = n * gb;
  for (int64_t t = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.y * blockDim.x) {
    const int64_t row = t / gb;
    const int j = (int)(t - row * gb);
    int c = zero_var && zero_var[g0 + j] ? n_perms : cnt[row * ldc + g0 + j];
    c = min(max(c, 0), n_perms);
    atomicAdd(&sh_hist[j * levels + c], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < gb * levels; t += blockDim.x)
    if (sh_hist[t]) atomicAdd(&hist[(int64_t)g0 * levels + t], sh_hist[t]);
}

// One thread per gene: adjusted p-value of every level.  method 0 none, 1 bonferroni, 2 fdr_bh.
__global__ void local_adjust_table_kernel(const unsigned int* __restrict__ hist, int g, int n_perms,
                                          int64_t n, int method, float* __restrict__ table /*[g][P+1]*/) {
  const int gene = blockIdx.x * blockDim.x + threadIdx.x;
  if (gene >= g) return;
  const int levels = n_perms + 1;
  const unsigned int* h = hist + (int64_t)gene * levels;
  float* tb = table + (int64_t)gene * levels;
  const float nf = (float)n;
  if (method != 2) {
    for (int v = 0; v < levels; ++v) {
      const float p = (float)((double)(v + 1) / (double)levels);
      tb[v] = method == 0 ? p : fminf(fmaxf(__fmul_rn(p, nf), 0.f), 1.f);
    }
    return;
  }
  int64_t cum = n;  // walk the levels from the top: cum = number of cells at levels <= v
  double running = INFINITY;
  for (int v = levels - 1; v >= 0; --v) {
    if (h[v]) {
      const float p = (float)((double)(v + 1) / (double)levels);
      const double t = (double)__fmul_rn(p, nf) / (double)cum;
      running = fmin(running, t);
    }
    tb[v] = (float)fmin(fmax(running, 0.0), 1.0);
    cum -= h[v];
  }
}

// Per cell and gene: p, adjusted p, quadrant; every per-cell output is written at the cell's ORIGINAL
// row (order[a] = original id of stored row a), tightly packed [n][g].
__global__ void __launch_bounds__(256)
local_finish_kernel(const int32_t* __restrict__ cnt, int64_t ldc, const float* __restrict__ Z,
                    const float* __restrict__ lag, const float* __restrict__ loc, int64_t ldz,
                    const int32_t* __restrict__ order, int64_t n, int g, int n_perms,
                    const uint8_t* __restrict__ zero_var, const float* __restrict__ table, float alpha,
                    float* __restrict__ z_out, float* __restrict__ lag_out, float* __restrict__ loc_out,
                    float* __restrict__ p_out, float* __restrict__ padj_out, int8_t* __restrict__ quad_out) {
  const int levels = n_perms + 1;
  const int64_t total = n * g;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = t / g;
    const int j = (int)(t - a * g);
    const int64_t dst = (order ? (int64_t)order[a] : a) * g + j;
    const bool dead = zero_var && zero_var[j];
    const float z = dead ? 0.f : Z[a * ldz + j];
    const float lg = dead ? 0.f : lag[a * ldz + j];
    z_out[dst] = z;
    lag_out[dst] = lg;
    loc_out[dst] = dead ? 0.f : loc[a * ldz + j];
    int q = 0;
    if (z > 0.f && lg > 0.f) q = 1;
    else if (z < 0.f && lg < 0.f) q = 2;
    else if (z > 0.f && lg < 0.f) q = 3;
    else if (z < 0.f && lg > 0.f) q = 4;
    if (n_perms > 0) {
      int c = dead ? n_perms : cnt[a * ldc + j];
      c = min(max(c, 0), n_perms);
      const float p = (float)((double)(c + 1) / (double)levels);
      const float pa = table[(int64_t)j * levels + c];
      p_out[dst] = p;
      padj_out[dst] = pa;
      if (pa >= alpha) q = 0;
    } else {
      p_out[dst] = 1.f;
      padj_out[dst] = 1.f;
    }
    quad_out[dst] = (int8_t)q;
  }
}

__global__ void philox_permutation_kernel(PermDomain dom, const __grid_constant__ PermBatch pb,
                                          int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dom.n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (int32_t)perm_apply((uint32_t)i, dom, pb.keys[0]);
}

__global__ void null_accumulate_kernel(const double* __restrict__ sims, int n_perms, int g,
                                       const double* __restrict__ scale,
                                       const double* __restrict__ obs, int64_t* __restrict__ cnt_ge,
                                       int64_t* __restrict__ cnt_abs_ge, double* __restrict__ sum,
                                       double* __restrict__ sumsq) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g) return;
  const double sc_ = scale ? scale[c] : 1.0, o = obs[c];
  int64_t ge = 0, age = 0;
  double s = 0, ss = 0;
  for (int p = 0; p < n_perms; ++p) {
    double v = sims[(int64_t)p * g + c] * sc_;
    ge += v >= o;
    age += fabs(v) >= fabs(o);
    s += v; ss += v * v;
  }
  if (cnt_ge) cnt_ge[c] += ge;
  if (cnt_abs_ge) cnt_abs_ge[c] += age;
  if (sum) sum[c] += s;
  if (sumsq) sumsq[c] += ss;
}

static int persistent_blocks(const void* kernel, int threads, size_t smem) {
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
  if (per_sm < 1) per_sm = 1;
  return sm_count() * per_sm;
}

static void fill_batch(PermBatch* pb, int source, const int32_t* perm_idx, uint64_t seed,
                       int64_t first_perm, int64_t local_first, int count, int64_t n) {
  pb->source = source;
  pb->count = count;
  pb->replay = (source == SC_PERM_REPLAY) ? perm_idx + local_first * n : nullptr;
  pb->dom = make_perm_domain((uint32_t)n);
  for (int p = 0; p < kMaxPermBatch; ++p) {
    if (p < count && source == SC_PERM_PHILOX) {
      perm_round_keys(seed, (uint64_t)(first_perm + p), pb->keys[p]);
    } else {
      for (int r = 0; r < kFeistelRounds; ++r) pb->keys[p][r] = 0;
    }
  }
}

constexpr int kMaxStatBlocks = 148 * 8;

// Launch lag_stat_kernel; *by_out = number of partial rows written ([by][2][ldz] doubles).
static int launch_lag_stat(const int32_t* indptr, const int32_t* indices, const float* weights,
                           int64_t n, int k_fixed, const float* Zself, const float* Zlag, int64_t ldz,
                           float* lag, float* local, int64_t ldl, double* partial,
                           const float* cell_obs, int32_t* cell_cnt, int64_t ldc, int* by_out,
                           cudaStream_t st) {
  int chunk_rows = 512;
  while (chunk_rows > 64 && ((n + chunk_rows - 1) / chunk_rows) * ((ldz + 31) / 32) < 4 * (int64_t)sm_count()) chunk_rows /= 2;
  const int bx = (int)((ldz + 4 * kLagColQuads - 1) / (4 * kLagColQuads));
  const int64_t n_chunks = (n + chunk_rows - 1) / chunk_rows;
  int64_t by = ((int64_t)sm_count() * 8 + bx - 1) / bx;
  if (by > n_chunks) by = n_chunks;
  if (by > kMaxStatBlocks) by = kMaxStatBlocks;
  if (by < 1) by = 1;
  dim3 grid(bx, (unsigned)by);
  if (weights)
    lag_stat_kernel<true, kLagDefaultUnr, kLagColQuads><<<grid, kStatThreads, 0, st>>>(indptr, indices, weights, n, k_fixed, Zself, Zlag, ldz, lag, local, ldl, partial, cell_obs, cell_cnt, ldc, n_chunks, chunk_rows);
  else
    lag_stat_kernel<false, kLagDefaultUnr, kLagColQuads><<<grid, kStatThreads, 0, st>>>(indptr, indices, weights, n, k_fixed, Zself, Zlag, ldz, lag, local, ldl, partial, cell_obs, cell_cnt, ldc, n_chunks, chunk_rows);
  SC_LAUNCH_OK();
  *by_out = (int)by;
  return SC_OK;
}

}  // namespace sc

using namespace sc;

// ================================================================================================
// C ABI
// ================================================================================================

extern "C" size_t sc_zscore_workspace_bytes(int64_t n, int g) {
  (void)n;
  return align_up(sizeof(double) * 2 * (size_t)kMaxStatBlocks * (size_t)(g > 0 ? g : 1), 256) + 512;
}

template <typename T>
static int zscore_impl(const T* X, int64_t n, int64_t ldx, int g, const int32_t* cols,
                       const int32_t* rows, float* Z, int64_t ldz, double* mean, double* std,
                       uint8_t* zero_var, double* partial, cudaStream_t st) {
  // fast path: FP32, all columns, rows readable as float4 up to round_up(g,4)
  const bool fast = sizeof(T) == 4 && !cols && ldx % 4 == 0 && ldx >= (g + 3) / 4 * 4 &&
                    (reinterpret_cast<uintptr_t>(X) & 15) == 0 && !getenv("SC_ZSCORE_GENERIC");
  if (fast) {
    const int quads = (g + 3) / 4;
    int qw_log2 = 0;
    while ((1 << qw_log2) < quads && qw_log2 < 8) ++qw_log2;
    const int qw = 1 << qw_log2, rpc = 256 >> qw_log2;
    const int bx = (quads + qw - 1) / qw;
    int64_t by = ((int64_t)sm_count() * 8 + bx - 1) / bx;
    const int64_t max_by = (n + rpc - 1) / rpc;
    if (by > max_by) by = max_by;
    if (by * bx > kMaxStatBlocks) by = kMaxStatBlocks / bx;
    if (by < 1) by = 1;
    const float* Xf = reinterpret_cast<const float*>(X);
    colstats4_kernel<<<dim3(bx, (unsigned)by), 256, 0, st>>>(Xf, n, ldx, g, qw_log2, partial);
    SC_LAUNCH_OK();
    colstats_final_kernel<T><<<(g + 127) / 128, 128, 0, st>>>(X, n, g, cols, partial, (int)by, mean, std, zero_var);
    SC_LAUNCH_OK();
    if (Z) {
      const int zquads = (int)(ldz / 4);
      int zq_log2 = 0;
      while ((1 << zq_log2) < zquads && zq_log2 < 8) ++zq_log2;
      const int zqw = 1 << zq_log2, zrpc = 256 >> zq_log2;
      const int zbx = (zquads + zqw - 1) / zqw;
      int64_t zby = ((int64_t)sm_count() * 16 + zbx - 1) / zbx;
      const int64_t zmax = (n + zrpc - 1) / zrpc;
      if (zby > zmax) zby = zmax;
      if (zby > 65535) zby = 65535;
      zscore_write4_kernel<<<dim3(zbx, (unsigned)zby), 256, 0, st>>>(Xf, n, ldx, g, rows, mean, std, zero_var, Z, ldz, zq_log2);
      SC_LAUNCH_OK();
    }
    return SC_OK;
  }
  int tiles = (g + 127) / 128;
  int groups = (sm_count() * 8 + tiles - 1) / tiles;
  int64_t max_groups = (n + 1) / 2;
  if (groups > max_groups) groups = (int)max_groups;
  if (groups < 1) groups = 1;
  if ((int64_t)groups * tiles > kMaxStatBlocks) groups = kMaxStatBlocks / tiles;
  if (groups < 1) groups = 1;
  colstats_kernel<T><<<dim3(tiles, groups), dim3(128, 2), 0, st>>>(X, n, ldx, g, cols, partial);
  SC_LAUNCH_OK();
  colstats_final_kernel<T><<<(g + 127) / 128, 128, 0, st>>>(X, n, g, cols, partial, groups, mean, std,
                                                            zero_var);
  SC_LAUNCH_OK();
  if (Z) {
    int64_t total = n * (ldz / 4);
    int64_t want = (total + 255) / 256;
    int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : (want < 1 ? 1 : want));
    zscore_write_kernel<T><<<blocks, 256, 0, st>>>(X, n, ldx, g, cols, rows, mean, std, zero_var, Z,
                                                   ldz);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}

extern "C" int sc_zscore(const void* X, int dtype, int64_t n, int64_t ldx, int g,
                         const int32_t* cols, const int32_t* rows, float* Z, int64_t ldz,
                         double* mean, double* std, uint8_t* zero_var, void* ws, size_t ws_bytes,
                         sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(X && mean && std && zero_var && ws, "sc_zscore: null argument");
  SC_CHECK_ARG(n >= 1 && g >= 1, "sc_zscore: empty matrix (%lld x %d)", (long long)n, g);
  SC_CHECK_ARG(!Z || (ldz % 4 == 0 && ldz >= g), "sc_zscore: ldz=%lld must be a multiple of 4 and >= g", (long long)ldz);
  SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_F64, "sc_zscore: bad dtype %d", dtype);
  if (ws_bytes < sc_zscore_workspace_bytes(n, g)) { set_error("sc_zscore: workspace too small"); return SC_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  if (dtype == SC_F32)
    return zscore_impl<float>(static_cast<const float*>(X), n, ldx, g, cols, rows, Z, ldz, mean, std, zero_var, partial, st);
  return zscore_impl<double>(static_cast<const double*>(X), n, ldx, g, cols, rows, Z, ldz, mean, std, zero_var, partial, st);
}

extern "C" int sc_zscore_apply(const void* X, int dtype, int64_t n, int64_t ldx, int g,
                               const int32_t* cols, const int32_t* rows, const double* mean,
                               const double* std, const uint8_t* zero_var, float* Z, int64_t ldz,
                               sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(X && mean && std && zero_var && Z, "sc_zscore_apply: null argument");
  SC_CHECK_ARG(n >= 1 && g >= 1, "sc_zscore_apply: empty matrix (%lld x %d)", (long long)n, g);
  SC_CHECK_ARG(ldz % 4 == 0 && ldz >= g, "sc_zscore_apply: ldz=%lld must be a multiple of 4 and >= g", (long long)ldz);
  SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_F64, "sc_zscore_apply: bad dtype %d", dtype);
  const bool fast = dtype == SC_F32 && !cols && ldx % 4 == 0 && ldx >= (g + 3) / 4 * 4 &&
                    (reinterpret_cast<uintptr_t>(X) & 15) == 0;
  if (fast) {
    const int zquads = (int)(ldz / 4);
    int zq_log2 = 0;
    while ((1 << zq_log2) < zquads && zq_log2 < 8) ++zq_log2;
    const int zqw = 1 << zq_log2, zrpc = 256 >> zq_log2;
    const int zbx = (zquads + zqw - 1) / zqw;
    int64_t zby = ((int64_t)sm_count() * 16 + zbx - 1) / zbx;
    const int64_t zmax = (n + zrpc - 1) / zrpc;
    if (zby > zmax) zby = zmax;
    if (zby > 65535) zby = 65535;
    zscore_write4_kernel<<<dim3(zbx, (unsigned)zby), 256, 0, st>>>(static_cast<const float*>(X), n, ldx, g, rows, mean, std, zero_var, Z, ldz, zq_log2);
  } else {
    int64_t total = n * (ldz / 4);
    int64_t want = (total + 255) / 256;
    int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : (want < 1 ? 1 : want));
    if (dtype == SC_F32)
      zscore_write_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(X), n, ldx, g, cols, rows, mean, std, zero_var, Z, ldz);
    else
      zscore_write_kernel<double><<<blocks, 256, 0, st>>>(static_cast<const double*>(X), n, ldx, g, cols, rows, mean, std, zero_var, Z, ldz);
  }
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_zscore_scatter(const float* X, int64_t n, int64_t ldx, int g, const int32_t* dst_rows,
                                 const double* mean, const double* std, const uint8_t* zero_var,
                                 const uint64_t* peer_ptrs_host, int n_peers, int64_t ldz,
                                 sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(X && dst_rows && mean && std && zero_var && peer_ptrs_host, "sc_zscore_scatter: null argument");
  SC_CHECK_ARG(n >= 1 && g >= 1 && ldz % 4 == 0 && ldz >= g, "sc_zscore_scatter: bad shape");
  SC_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers, "sc_zscore_scatter: need 1 <= n_peers <= %d", kMaxPeers);
  SC_CHECK_ARG(ldx % 4 == 0 && ldx >= (g + 3) / 4 * 4 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
               "sc_zscore_scatter: X rows must be float4-readable (ldx %% 4 == 0, 16-byte aligned)");
  PeerPtrs pp;
  for (int p = 0; p < kMaxPeers; ++p) pp.p[p] = p < n_peers ? reinterpret_cast<float*>(peer_ptrs_host[p]) : nullptr;
  for (int p = 0; p < n_peers; ++p) SC_CHECK_ARG(pp.p[p] && (peer_ptrs_host[p] & 15) == 0, "sc_zscore_scatter: peer pointer %d null or misaligned", p);
  const int zquads = (int)(ldz / 4);
  int zq_log2 = 0;
  while ((1 << zq_log2) < zquads && zq_log2 < 8) ++zq_log2;
  const int zqw = 1 << zq_log2, zrpc = 256 >> zq_log2;
  const int zbx = (zquads + zqw - 1) / zqw;
  int64_t zby = ((int64_t)sm_count() * 16 + zbx - 1) / zbx;
  const int64_t zmax = (n + zrpc - 1) / zrpc;
  if (zby > zmax) zby = zmax;
  if (zby > 65535) zby = 65535;
  zscore_scatter4_kernel<<<dim3(zbx, (unsigned)zby), 256, 0, st>>>(X, n, ldx, g, dst_rows, mean, std, zero_var, pp, n_peers, ldz, zq_log2);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_csr_densify(const int64_t* indptr, const int32_t* indices, const void* data,
                              int dtype, int64_t n, const int32_t* colmap, int g_out, float* out,
                              int64_t ldo, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indptr && out && (indices || n == 0), "sc_csr_densify: null argument");
  SC_CHECK_ARG(ldo >= g_out, "sc_csr_densify: ldo < g_out");
  SC_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)n * (size_t)ldo, st));
  int64_t want = (n + 7) / 8;
  int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : (want < 1 ? 1 : want));
  if (dtype == SC_F32)
    csr_densify_kernel<float><<<blocks, 256, 0, st>>>(indptr, indices, static_cast<const float*>(data), n, colmap, g_out, out, ldo);
  else if (dtype == SC_F64)
    csr_densify_kernel<double><<<blocks, 256, 0, st>>>(indptr, indices, static_cast<const double*>(data), n, colmap, g_out, out, ldo);
  else { set_error("sc_csr_densify: bad dtype %d", dtype); return SC_ERR_INVALID; }
  SC_LAUNCH_OK();
  return SC_OK;
}

// Largest leading dimension the partial buffers are sized for.
static size_t max_ld(int g) { return align_up((size_t)(g > 0 ? g : 1), 32); }

// [partial sums][512 B slack][sc_csr_lag_moran_tiled with a permutation: the chunk unions composed with it,
// at most 1 280 rows per 256-row chunk]
extern "C" size_t sc_csr_lag_moran_workspace_bytes(int64_t n, int g) {
  size_t ld = max_ld(g);
  size_t bytes = align_up(sizeof(double) * 2 * (size_t)kMaxStatBlocks * ld, 256) + 512;
  if (n > 0) bytes += align_up(sizeof(int32_t) * (size_t)((n + 255) / 256) * 1280, 256);
  return bytes;
}

extern "C" int sc_csr_lag_moran(const int32_t* indptr, const int32_t* indices,
                                const float* weights, int64_t n, int k_fixed, const float* Z,
                                int64_t ldz, int g, float* lag, float* local, int64_t ldl,
                                double* num, double* den, void* ws, size_t ws_bytes,
                                sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && Z && num && den && ws, "sc_csr_lag_moran: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_csr_lag_moran: need indptr or k_fixed");
  SC_CHECK_ARG(ldz % 4 == 0 && ldz >= g && g >= 1 && (size_t)ldz <= max_ld(g),
               "sc_csr_lag_moran: ldz must be a multiple of 4 in [g, round_up(g,32)]");
  SC_CHECK_ARG((!lag && !local) || (ldl % 4 == 0 && ldl >= ldz), "sc_csr_lag_moran: ldl must be a multiple of 4 and >= ldz");
  if (ws_bytes < sc_csr_lag_moran_workspace_bytes(n, g)) { set_error("sc_csr_lag_moran: workspace too small"); return SC_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  int by = 0;
  int rc = launch_lag_stat(indptr, indices, weights, n, k_fixed, nullptr, Z, ldz, lag, local, ldl, partial,
                           nullptr, nullptr, 0, &by, st);
  if (rc) return rc;
  const int bx = by;
  // partial rows: [block][0]=num, [block][1]=den  -> view as (nblocks, rows=2, ldp=ldz)
  reduce_partials_kernel<<<dim3((g + 127) / 128, 1), 128, 0, st>>>(partial, bx, 2, ldz, g, num, g);
  SC_LAUNCH_OK();
  reduce_partials_kernel<<<dim3((g + 127) / 128, 1), 128, 0, st>>>(partial + ldz, bx, 2, ldz, g, den, g);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_perm_null_workspace_bytes(int64_t n, int g) {
  (void)n;
  size_t ld = max_ld(g);
  return align_up(sizeof(double) * (size_t)kMaxStatBlocks * kMaxPermBatch * ld, 256) + 512;
}

static int check_perm_args(const char* who, int source, const int32_t* perm_idx, int64_t n,
                           int n_perms, const double* sims) {
  SC_CHECK_ARG(source == SC_PERM_REPLAY || source == SC_PERM_PHILOX, "%s: bad permutation source %d", who, source);
  SC_CHECK_ARG(source != SC_PERM_REPLAY || perm_idx, "%s: replay source needs perm_idx", who);
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31), "%s: n out of range", who);
  SC_CHECK_ARG(n_perms >= 0 && sims, "%s: bad n_perms/sims", who);
  return SC_OK;
}

template <int PB, typename AccT>
static int launch_perm_rows(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n,
                            int g, int source, const int32_t* perm_idx, uint64_t seed,
                            int64_t perm_offset, int n_perms, double* sims, double* partial,
                            cudaStream_t st) {
  RowGeom rg = row_geom(lda);
  int bx = persistent_blocks((const void*)perm_rows_kernel<PB, AccT>, kStatThreads, 0);
  int64_t max_bx = (n + rg.rpp - 1) / rg.rpp;
  if (bx > max_bx) bx = (int)max_bx;
  if (bx > kMaxStatBlocks) bx = kMaxStatBlocks;
  int by = (int)((lda / 4 + 255) / 256);
  for (int p0 = 0; p0 < n_perms; p0 += PB) {
    int count = n_perms - p0 < PB ? n_perms - p0 : PB;
    PermBatch pb;
    fill_batch(&pb, source, perm_idx, seed, perm_offset + p0, p0, count, n);
    perm_rows_kernel<PB, AccT><<<dim3(bx, by), kStatThreads, 0, st>>>(A, lda, B, ldb, n, pb, partial, lda, rg.wpr);
    SC_LAUNCH_OK();
    reduce_partials_kernel<<<dim3((g + 127) / 128, count), 128, 0, st>>>(partial, bx, PB, lda, g, sims + (int64_t)p0 * g, g);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}


template <int PB>
static int launch_perm_rows_bulk(const float* A, int64_t lda, const float* B, int64_t ldb,
                                 int64_t n, int g, int source, const int32_t* perm_idx,
                                 uint64_t seed, int64_t perm_offset, int n_perms, double* sims,
                                 double* partial, cudaStream_t st) {
  RowGeom rg = row_geom(lda);
  const int by = (int)((lda + 1023) / 1024);
  // column blocks of 1024 floats; the last one may be narrower (the kernel trims its row segment)
  const int cols = (int)(by > 1 ? 1024 : lda);
  const size_t stage_bytes = (size_t)rg.rpp * (PB + 1) * cols * 4;
  int max_smem = 0, dev = 0;
  SC_CUDA_OK(cudaGetDevice(&dev));
  SC_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t budget = (size_t)max_smem - 12 * 1024;  // static smem (barriers, keys, reduction)
  int stages = (int)(budget / stage_bytes);
  if (stages > kBulkMaxStages) stages = kBulkMaxStages;
  if (stages < 2) return SC_ERR_UNSUPPORTED;
  // producers must divide stages (see the kernel's invariant): prefer 4 warps, then 3, then 2
  int n_producers = stages >= 4 ? 4 : stages;
  stages = stages / n_producers * n_producers;
  const size_t dyn = (size_t)stages * stage_bytes;
  SC_CUDA_OK(cudaFuncSetAttribute(perm_rows_bulk_kernel<PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int bx = sm_count();
  int64_t n_groups = (n + rg.rpp - 1) / rg.rpp;
  if (bx > n_groups) bx = (int)n_groups;
  if (by > 1) bx = bx / by > 0 ? bx / by : 1;
  for (int p0 = 0; p0 < n_perms; p0 += PB) {
    int count = n_perms - p0 < PB ? n_perms - p0 : PB;
    PermBatch pb;
    fill_batch(&pb, source, perm_idx, seed, perm_offset + p0, p0, count, n);
    perm_rows_bulk_kernel<PB><<<dim3(bx, by), kBulkThreads, dyn, st>>>(A, lda, B, ldb, n, cols, pb, partial, lda, rg.wpr, stages, n_producers);
    SC_LAUNCH_OK();
    reduce_partials_kernel<<<dim3((g + 127) / 128, count), 128, 0, st>>>(partial, bx, PB, lda, g, sims + (int64_t)p0 * g, g);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}

// Tunable from the host for experiments: SC_PERM_ROWS_VARIANT =
//   "bulk16" (default) | "bulk8" : bulk-async shared-memory pipeline, 16 / 8 permutations per pass
//   "8d" | "16d" | "16f" | "8f" | "4d" : register-staged gather kernel (PB, accumulator type)
static int perm_rows_variant() {
  const char* v = getenv("SC_PERM_ROWS_VARIANT");
  if (!v) return 11;
  if (!strcmp(v, "bulk8")) return 10;
  if (!strcmp(v, "bulk16")) return 11;
  if (!strcmp(v, "8d")) return 0;
  if (!strcmp(v, "16d")) return 1;
  if (!strcmp(v, "16f")) return 2;
  if (!strcmp(v, "8f")) return 3;
  if (!strcmp(v, "4d")) return 4;
  return 0;
}

extern "C" int sc_perm_null_graph_rows(const float* A, int64_t lda, const float* B, int64_t ldb,
                                       int64_t n, int g, int source, const int32_t* perm_idx,
                                       uint64_t seed, int64_t perm_offset, int n_perms,
                                       double* sims, void* ws, size_t ws_bytes,
                                       sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_perm_args("sc_perm_null_graph_rows", source, perm_idx, n, n_perms, sims);
  if (rc) return rc;
  SC_CHECK_ARG(A && B && ws, "sc_perm_null_graph_rows: null argument");
  SC_CHECK_ARG(lda % 4 == 0 && lda >= g && ldb >= lda && ldb % 4 == 0 && g >= 1 &&
                   (size_t)lda <= max_ld(g),
               "sc_perm_null_graph_rows: lda/ldb must be multiples of 4, g <= lda <= round_up(g,32), ldb >= lda");
  if (ws_bytes < sc_perm_null_workspace_bytes(n, g)) { set_error("sc_perm_null_graph_rows: workspace too small"); return SC_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  int variant = perm_rows_variant();
  // narrow rows (<= 1 KB per gathered row): 8 permutations per pass keep more, smaller stages in flight and
  // measured 5-12 % faster than 16 (profiles/r01c_perm_rows_width_sweep.txt); an explicit choice wins
  if (!getenv("SC_PERM_ROWS_VARIANT") && lda <= 256) variant = 10;
  if (variant >= 10) {
    rc = variant == 11
             ? launch_perm_rows_bulk<16>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st)
             : launch_perm_rows_bulk<8>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
    if (rc != SC_ERR_UNSUPPORTED) return rc;
    variant = 0;  // geometry the pipeline does not cover: register-staged kernel
  }
  switch (variant) {
    case 1: return launch_perm_rows<16, double>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
    case 2: return launch_perm_rows<16, float>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
    case 3: return launch_perm_rows<8, float>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
    case 4: return launch_perm_rows<4, double>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
    default: return launch_perm_rows<8, double>(A, lda, B, ldb, n, g, source, perm_idx, seed, perm_offset, n_perms, sims, partial, st);
  }
}

template <int PB, bool HAS_W>
static int launch_perm_values(const int32_t* indptr, const int32_t* indices, const float* weights,
                              int64_t n, int k_fixed, const float* Zx, const float* Zy,
                              int64_t ldz, int g, int source, const int32_t* perm_idx,
                              uint64_t seed, int64_t perm_offset, int n_perms, double* sims,
                              const float* cell_obs, int32_t* cell_cnt, int64_t ldc,
                              double* partial, cudaStream_t st) {
  RowGeom rg = row_geom(ldz);
  int bx = persistent_blocks((const void*)perm_values_kernel<PB, HAS_W>, kStatThreads, 0);
  int64_t max_bx = (n + rg.rpp - 1) / rg.rpp;
  if (bx > max_bx) bx = (int)max_bx;
  if (bx > kMaxStatBlocks) bx = kMaxStatBlocks;
  int by = (int)((ldz / 4 + 255) / 256);
  for (int p0 = 0; p0 < n_perms; p0 += PB) {
    int count = n_perms - p0 < PB ? n_perms - p0 : PB;
    PermBatch pb;
    fill_batch(&pb, source, perm_idx, seed, perm_offset + p0, p0, count, n);
    perm_values_kernel<PB, HAS_W><<<dim3(bx, by), kStatThreads, 0, st>>>(
        indptr, indices, weights, n, k_fixed, Zx, Zy, ldz, pb, partial, ldz, cell_obs, cell_cnt, ldc, rg.wpr);
    SC_LAUNCH_OK();
    reduce_partials_kernel<<<dim3((g + 127) / 128, count), 128, 0, st>>>(partial, bx, PB, ldz, g, sims + (int64_t)p0 * g, g);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}

// Wide matrices: materialise Zp = Zy[π_p(.)] (one streaming row gather) and run the lag kernel on it.
// With the cells in spatial order the lag kernel reads Zp about once, so a permutation costs ~12·N·ld
// bytes of HBM traffic instead of (k+1) random row gathers per cell.
constexpr int kValuesMaterializeMinLd = 32;

extern "C" size_t sc_perm_null_values_workspace_bytes(int64_t n, int g) {
  size_t base = sc_perm_null_workspace_bytes(n, g);
  if (max_ld(g) >= (size_t)kValuesMaterializeMinLd && n > 0)
    base += align_up(sizeof(float) * (size_t)n * max_ld(g), 256);
  return base;
}

static int launch_perm_values_materialised(const int32_t* indptr, const int32_t* indices,
                                           const float* weights, int64_t n, int k_fixed,
                                           const float* Zx, const float* Zy, int64_t ldz, int g,
                                           int source, const int32_t* perm_idx, uint64_t seed,
                                           int64_t perm_offset, int n_perms, double* sims,
                                           const float* cell_obs, int32_t* cell_cnt, int64_t ldc,
                                           double* partial, float* Zp, cudaStream_t st) {
  int64_t want = (n + 7) / 8;
  const int pblocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : (want < 1 ? 1 : want));
  for (int p = 0; p < n_perms; ++p) {
    PermBatch pb;
    fill_batch(&pb, source, perm_idx, seed, perm_offset + p, p, 1, n);
    permute_rows_kernel<<<pblocks, 256, 0, st>>>(Zy, ldz, n, pb, Zp);
    SC_LAUNCH_OK();
    int by = 0;
    int rc = launch_lag_stat(indptr, indices, weights, n, k_fixed, Zx, Zp, ldz, nullptr, nullptr, 0,
                             partial, cell_obs, cell_cnt, ldc, &by, st);
    if (rc) return rc;
    reduce_partials_kernel<<<dim3((g + 127) / 128, 1), 128, 0, st>>>(partial, by, 2, ldz, g, sims + (int64_t)p * g, g);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}

extern "C" int sc_perm_null_values(const int32_t* indptr, const int32_t* indices,
                                   const float* weights, int64_t n, int k_fixed, const float* Zx,
                                   const float* Zy, int64_t ldz, int g, int source,
                                   const int32_t* perm_idx, uint64_t seed, int64_t perm_offset,
                                   int n_perms, double* sims, const float* cell_obs,
                                   int32_t* cell_cnt, int64_t ldc, void* ws, size_t ws_bytes,
                                   sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = check_perm_args("sc_perm_null_values", source, perm_idx, n, n_perms, sims);
  if (rc) return rc;
  SC_CHECK_ARG(indices && Zy && ws, "sc_perm_null_values: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_perm_null_values: need indptr or k_fixed");
  SC_CHECK_ARG(ldz % 4 == 0 && ldz >= g && g >= 1 && (size_t)ldz <= max_ld(g),
               "sc_perm_null_values: ldz must be a multiple of 4 in [g, round_up(g,32)]");
  SC_CHECK_ARG((cell_cnt == nullptr) == (cell_obs == nullptr), "sc_perm_null_values: cell_obs and cell_cnt go together");
  SC_CHECK_ARG(!cell_cnt || (ldc % 4 == 0 && ldc >= ldz), "sc_perm_null_values: ldc must be a multiple of 4 and >= ldz");
  if (ws_bytes < sc_perm_null_workspace_bytes(n, g)) { set_error("sc_perm_null_values: workspace too small"); return SC_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  const size_t base = sc_perm_null_workspace_bytes(n, g);
  const size_t zp_bytes = align_up(sizeof(float) * (size_t)n * (size_t)ldz, 256);
  const char* force = getenv("SC_PERM_VALUES_VARIANT");  // "gather" forces the register-gather kernel
  const bool allow = !(force && !strcmp(force, "gather"));
  if (allow && ldz >= kValuesMaterializeMinLd && ws_bytes >= base + zp_bytes) {
    float* Zp = reinterpret_cast<float*>(static_cast<char*>(ws) + base);
    return launch_perm_values_materialised(indptr, indices, weights, n, k_fixed, Zx, Zy, ldz, g, source,
                                           perm_idx, seed, perm_offset, n_perms, sims, cell_obs,
                                           cell_cnt, ldc, partial, Zp, st);
  }
  if (weights)
    return launch_perm_values<4, true>(indptr, indices, weights, n, k_fixed, Zx, Zy, ldz, g, source, perm_idx, seed, perm_offset, n_perms, sims, cell_obs, cell_cnt, ldc, partial, st);
  return launch_perm_values<4, false>(indptr, indices, weights, n, k_fixed, Zx, Zy, ldz, g, source, perm_idx, seed, perm_offset, n_perms, sims, cell_obs, cell_cnt, ldc, partial, st);
}

extern "C" int sc_gather_rows(const float* src, int64_t lds, int64_t n, int64_t cols,
                              const int32_t* rows, float* dst, int64_t ldd, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(src && rows && dst, "sc_gather_rows: null argument");
  SC_CHECK_ARG(n >= 1 && cols >= 4 && cols % 4 == 0 && lds >= cols && ldd >= cols && lds % 4 == 0 && ldd % 4 == 0,
               "sc_gather_rows: cols, lds, ldd must be multiples of 4 with lds, ldd >= cols");
  int64_t want = (n + 7) / 8;
  int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : want);
  gather_rows_kernel<<<blocks, 256, 0, st>>>(src, lds, n, cols, rows, dst, ldd);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_perm_conjugate(const int32_t* perm_idx, int64_t n, int n_perms,
                                 const int32_t* order, const int32_t* rank, int32_t* out,
                                 sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(perm_idx && order && rank && out && perm_idx != out, "sc_perm_conjugate: null or aliased argument");
  SC_CHECK_ARG(n >= 1 && n_perms >= 1, "sc_perm_conjugate: empty input");
  int64_t want = (n * n_perms + 255) / 256;
  int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : want);
  perm_conjugate_kernel<<<blocks, 256, 0, st>>>(perm_idx, n, n_perms, order, rank, out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_local_moran_finish_workspace_bytes(int g, int n_perms) {
  const size_t levels = (size_t)(n_perms > 0 ? n_perms : 0) + 1;
  return align_up(sizeof(unsigned int) * (size_t)(g > 0 ? g : 1) * levels, 256) +
         align_up(sizeof(float) * (size_t)(g > 0 ? g : 1) * levels, 256) + 512;
}

extern "C" int sc_local_moran_finish(const int32_t* cnt, int64_t ldc, const float* Z, const float* lag,
                                     const float* local, int64_t ldz, const int32_t* order, int64_t n,
                                     int g, int n_perms, const uint8_t* zero_var, int method, float alpha,
                                     float* z_out, float* lag_out, float* local_out, float* p_out,
                                     float* padj_out, int8_t* quad_out, void* ws, size_t ws_bytes,
                                     sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(Z && lag && local && z_out && lag_out && local_out && p_out && padj_out && quad_out && ws,
               "sc_local_moran_finish: null argument");
  SC_CHECK_ARG(n >= 1 && g >= 1 && ldz >= g && n_perms >= 0, "sc_local_moran_finish: bad sizes");
  SC_CHECK_ARG(n_perms == 0 || (cnt && ldc >= g), "sc_local_moran_finish: permutation counts missing");
  SC_CHECK_ARG(method >= 0 && method <= 2, "sc_local_moran_finish: method must be 0 (none), 1 (bonferroni) or 2 (fdr_bh)");
  if (ws_bytes < sc_local_moran_finish_workspace_bytes(g, n_perms)) { set_error("sc_local_moran_finish: workspace too small"); return SC_ERR_WORKSPACE; }
  const int levels = n_perms + 1;
  unsigned int* hist = static_cast<unsigned int*>(ws);
  float* table = reinterpret_cast<float*>(static_cast<char*>(ws) + align_up(sizeof(unsigned int) * (size_t)g * levels, 256));
  if (n_perms > 0) {
    if (method == 2) {
      SC_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(unsigned int) * (size_t)g * levels, st));
      int gpb = (int)(48 * 1024 / (sizeof(unsigned int) * levels));
      if (gpb < 1) { set_error("sc_local_moran_finish: more than 12287 permutations are not supported"); return SC_ERR_UNSUPPORTED; }
      if (gpb > g) gpb = g;
      const int bx = (g + gpb - 1) / gpb;
      int by = (sm_count() * 8 + bx - 1) / bx;
      const int64_t max_by = (n * gpb + 255) / 256;
      if (by > max_by) by = (int)max_by;
      if (by < 1) by = 1;
      local_hist_kernel<<<dim3(bx, by), 256, sizeof(unsigned int) * (size_t)gpb * levels, st>>>(cnt, ldc, n, g, n_perms, zero_var, gpb, hist);
      SC_LAUNCH_OK();
    }
    local_adjust_table_kernel<<<(g + 63) / 64, 64, 0, st>>>(hist, g, n_perms, n, method, table);
    SC_LAUNCH_OK();
  }
  int64_t want = (n * g + 255) / 256;
  int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : want);
  local_finish_kernel<<<blocks, 256, 0, st>>>(cnt, ldc, Z, lag, local, ldz, order, n, g, n_perms, zero_var, table, alpha,
                                              z_out, lag_out, local_out, p_out, padj_out, quad_out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_philox_permutation(uint64_t seed, int64_t perm_index, int64_t n, int32_t* out,
                                     sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(out && n >= 1 && n < (1ll << 31), "sc_philox_permutation: bad argument");
  PermBatch pb;
  fill_batch(&pb, SC_PERM_PHILOX, nullptr, seed, perm_index, 0, 1, n);
  int64_t want = (n + 255) / 256;
  int blocks = (int)(want > sm_count() * 8 ? sm_count() * 8 : want);
  philox_permutation_kernel<<<blocks, 256, 0, st>>>(pb.dom, pb, out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_null_accumulate(const double* sims, int n_perms, int g, const double* scale,
                                  const double* obs, int64_t* cnt_ge, int64_t* cnt_abs_ge,
                                  double* sum, double* sumsq, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(sims && obs && g >= 1 && n_perms >= 0, "sc_null_accumulate: bad argument");
  null_accumulate_kernel<<<(g + 127) / 128, 128, 0, st>>>(sims, n_perms, g, scale, obs, cnt_ge, cnt_abs_ge, sum, sumsq);
  SC_LAUNCH_OK();
  return SC_OK;
}
