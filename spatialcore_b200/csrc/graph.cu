// Grid-hashed exact neighbour graphs on sm_100a: kNN, radius, neighbourhood composition, graph
// moments.  See include/sc_b200.h for the contract and the reference lines each export replaces.
//
// Data layout: points are counting-sorted into a uniform grid (row-major cells); xs/ys/order hold the
// coordinates and original ids in cell order, cell_start[c] the first sorted position of cell c.  A
// cell-row segment [cx0,cx1] is therefore one contiguous range of the sorted arrays, which is what
// the query warps stream through.
#include <cub/cub.cuh>
#include <float.h>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace sc {

struct GridParams {
  double x0, y0, inv_h, h, margin;
  int nx, ny, ncells, pad;
};

constexpr int kBboxBlocks = 256;

// ------------------------------------------------------------------------------------------------
// binning
// ------------------------------------------------------------------------------------------------

__global__ void bbox_partial_kernel(const double* __restrict__ coords, int64_t n,
                                    double* __restrict__ partial) {
  double xmin = DBL_MAX, xmax = -DBL_MAX, ymin = DBL_MAX, ymax = -DBL_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double2 p = reinterpret_cast<const double2*>(coords)[i];
    xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
    ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
  }
  __shared__ double s[4][32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fmin(xmin, shfl_f64(xmin, (threadIdx.x & 31) ^ o));
    xmax = fmax(xmax, shfl_f64(xmax, (threadIdx.x & 31) ^ o));
    ymin = fmin(ymin, shfl_f64(ymin, (threadIdx.x & 31) ^ o));
    ymax = fmax(ymax, shfl_f64(ymax, (threadIdx.x & 31) ^ o));
  }
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { s[0][w] = xmin; s[1][w] = xmax; s[2][w] = ymin; s[3][w] = ymax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nw = blockDim.x >> 5;
    for (int j = 1; j < nw; ++j) {
      xmin = fmin(xmin, s[0][j]); xmax = fmax(xmax, s[1][j]);
      ymin = fmin(ymin, s[2][j]); ymax = fmax(ymax, s[3][j]);
    }
    partial[4 * blockIdx.x + 0] = xmin; partial[4 * blockIdx.x + 1] = xmax;
    partial[4 * blockIdx.x + 2] = ymin; partial[4 * blockIdx.x + 3] = ymax;
  }
}

// One thread: final bbox + grid geometry.  pts_per_cell > 0: occupancy-driven cell edge (kNN);
// h_fixed > 0: cell edge >= h_fixed (radius).  The cell count is clamped to max_cells.
__global__ void grid_setup_kernel(const double* __restrict__ partial, int nblocks, int64_t n,
                                  double pts_per_cell, double h_fixed, int max_cells,
                                  GridParams* __restrict__ gp) {
  double xmin = DBL_MAX, xmax = -DBL_MAX, ymin = DBL_MAX, ymax = -DBL_MAX;
  for (int j = 0; j < nblocks; ++j) {
    xmin = fmin(xmin, partial[4 * j + 0]); xmax = fmax(xmax, partial[4 * j + 1]);
    ymin = fmin(ymin, partial[4 * j + 2]); ymax = fmax(ymax, partial[4 * j + 3]);
  }
  double w = xmax - xmin, hg = ymax - ymin;
  double h;
  if (h_fixed > 0) {
    h = h_fixed;
  } else {
    h = sqrt(pts_per_cell * w * hg / (double)n);
    if (!(h > 0)) h = pts_per_cell * fmax(w, hg) / (double)n;
    if (!(h > 0)) h = 1.0;
  }
  double fx, fy;
  for (int it = 0; it < 64; ++it) {
    fx = floor(w / h) + 1.0;
    fy = floor(hg / h) + 1.0;
    if (fx * fy <= (double)max_cells) break;
    h *= 1.05 * sqrt(fx * fy / (double)max_cells);
  }
  gp->x0 = xmin; gp->y0 = ymin; gp->h = h; gp->inv_h = 1.0 / h;
  gp->nx = (int)fx; gp->ny = (int)fy; gp->ncells = gp->nx * gp->ny;
  double ext = fmax(fmax(fabs(xmin), fabs(xmax)), fmax(fabs(ymin), fabs(ymax)));
  gp->margin = 1e-9 * (ext + h);
  gp->pad = 0;
}

__device__ __forceinline__ int cell_coord(double v, double v0, double inv_h, int nmax) {
  int c = (int)floor((v - v0) * inv_h);
  return min(max(c, 0), nmax - 1);
}

__device__ __forceinline__ uint32_t spread_bits16(uint32_t v) {
  v &= 0xFFFFu;
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

// morton: key = Z-order code of the cell (used only to produce a spatially compact ordering; the
// query kernels need the row-major key so that a cell-row segment is one contiguous range).
__global__ void assign_cells_kernel(const double* __restrict__ coords, int64_t n,
                                    const GridParams* __restrict__ gp, int32_t* __restrict__ keys,
                                    int32_t* __restrict__ vals, int morton) {
  GridParams g = *gp;
  const bool zorder = morton && g.nx <= 32768 && g.ny <= 32768;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double2 p = reinterpret_cast<const double2*>(coords)[i];
    int cx = cell_coord(p.x, g.x0, g.inv_h, g.nx);
    int cy = cell_coord(p.y, g.y0, g.inv_h, g.ny);
    keys[i] = zorder ? (int32_t)(spread_bits16((uint32_t)cx) | (spread_bits16((uint32_t)cy) << 1))
                     : cy * g.nx + cx;
    vals[i] = (int32_t)i;
  }
}

// cell_start[c] = first sorted position whose key >= c, for c in [0, ncells].
__global__ void cell_bounds_kernel(const int32_t* __restrict__ keys_sorted, int64_t n,
                                   const GridParams* __restrict__ gp,
                                   int32_t* __restrict__ cell_start) {
  int ncells = gp->ncells;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s <= n;
       s += (int64_t)gridDim.x * blockDim.x) {
    int prev = (s == 0) ? -1 : keys_sorted[s - 1];
    int cur = (s == n) ? ncells : keys_sorted[s];
    for (int c = prev + 1; c <= cur; ++c) cell_start[c] = (int32_t)s;
  }
}

__global__ void gather_coords_kernel(const double* __restrict__ coords,
                                     const int32_t* __restrict__ order, int64_t n,
                                     double* __restrict__ xs, double* __restrict__ ys) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < n;
       s += (int64_t)gridDim.x * blockDim.x) {
    double2 p = reinterpret_cast<const double2*>(coords)[order[s]];
    xs[s] = p.x; ys[s] = p.y;
  }
}

struct Binning {
  GridParams* gp;
  double* partial;
  int32_t *keys, *vals, *keys_sorted, *order, *cell_start;
  double *xs, *ys;
  void* cub_tmp;
  size_t cub_bytes;
  int max_cells;
};

static int key_bits(int max_cells) {
  int b = 1;
  while ((1ll << b) < (long long)max_cells + 1) ++b;
  return b;
}

static size_t binning_cub_bytes(int64_t n, int max_cells) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 31);
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr,
                                (int)(n + 1));
  return bytes > scan ? bytes : scan;
}

static int max_cells_for(int64_t n) { return (int)(n + 1024); }

static size_t binning_bytes(int64_t n) {
  int mc = max_cells_for(n);
  size_t b = 0;
  b += align_up(sizeof(GridParams), 256);
  b += align_up(sizeof(double) * 4 * kBboxBlocks, 256);
  b += 4 * align_up(sizeof(int32_t) * n, 256);
  b += align_up(sizeof(int32_t) * ((size_t)mc + 1), 256);
  b += 2 * align_up(sizeof(double) * n, 256);
  b += align_up(binning_cub_bytes(n, mc), 256);
  return b;
}

static bool carve_binning(Arena& a, int64_t n, Binning* b) {
  b->max_cells = max_cells_for(n);
  b->gp = a.take<GridParams>(1);
  b->partial = a.take<double>(4 * kBboxBlocks);
  b->keys = a.take<int32_t>(n);
  b->vals = a.take<int32_t>(n);
  b->keys_sorted = a.take<int32_t>(n);
  b->order = a.take<int32_t>(n);
  b->cell_start = a.take<int32_t>((size_t)b->max_cells + 1);
  b->xs = a.take<double>(n);
  b->ys = a.take<double>(n);
  b->cub_bytes = binning_cub_bytes(n, b->max_cells);
  b->cub_tmp = a.take<char>(b->cub_bytes);
  return b->cub_tmp != nullptr;
}

static int run_binning(const double* coords, int64_t n, double pts_per_cell, double h_fixed,
                       const Binning& b, cudaStream_t st, bool order_only = false) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  bbox_partial_kernel<<<kBboxBlocks, 256, 0, st>>>(coords, n, b.partial);
  SC_LAUNCH_OK();
  grid_setup_kernel<<<1, 1, 0, st>>>(b.partial, kBboxBlocks, n, pts_per_cell, h_fixed, b.max_cells,
                                     b.gp);
  SC_LAUNCH_OK();
  assign_cells_kernel<<<blocks, 256, 0, st>>>(coords, n, b.gp, b.keys, b.vals, order_only ? 1 : 0);
  SC_LAUNCH_OK();
  size_t bytes = b.cub_bytes;
  SC_CUDA_OK(cub::DeviceRadixSort::SortPairs(b.cub_tmp, bytes, b.keys, b.keys_sorted, b.vals,
                                             b.order, (int)n, 0, order_only ? 31 : key_bits(b.max_cells), st));
  if (order_only) return SC_OK;  // Z-order keys: no cell table
  cell_bounds_kernel<<<blocks, 256, 0, st>>>(b.keys_sorted, n, b.gp, b.cell_start);
  SC_LAUNCH_OK();
  gather_coords_kernel<<<blocks, 256, 0, st>>>(coords, b.order, n, b.xs, b.ys);
  SC_LAUNCH_OK();
  return SC_OK;
}

// ------------------------------------------------------------------------------------------------
// kNN: one warp per query point, sorted top-k list distributed over the warp's lanes
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ bool cand_less(double da, int ia, double db, int ib) {
  return da < db || (da == db && ia < ib);
}

// d² exactly as the CPU tree libraries evaluate it: separate multiply and add, no FMA contraction.
__device__ __forceinline__ double sq_dist(double qx, double qy, double px, double py) {
  double dx = __dsub_rn(qx, px), dy = __dsub_rn(qy, py);
  return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

template <int E>
struct WarpTopK {
  double d[E];
  int id[E];
  double thr_d;
  int thr_id;
  int k;

  __device__ __forceinline__ void init(int k_) {
    k = k_;
#pragma unroll
    for (int e = 0; e < E; ++e) { d[e] = INFINITY; id[e] = INT_MAX; }
    thr_d = INFINITY; thr_id = INT_MAX;
  }

  // Warp-uniform call: insert (cd, cid), known to be below the current threshold.
  __device__ __forceinline__ void insert(double cd, int cid, int lane) {
    int pos = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      bool less = (e * 32 + lane < k) && cand_less(d[e], id[e], cd, cid);
      pos += __popc(__ballot_sync(kFull, less));
    }
#pragma unroll
    for (int e = E - 1; e >= 0; --e) {
      double pd = shfl_up_f64(d[e], 1);
      int pi = __shfl_up_sync(kFull, id[e], 1);
      if (e > 0) {
        double cdn = shfl_f64(d[e - 1], 31);
        int cin = __shfl_sync(kFull, id[e - 1], 31);
        if (lane == 0) { pd = cdn; pi = cin; }
      }
      int gpos = e * 32 + lane;
      if (gpos > pos) { d[e] = pd; id[e] = pi; }
      else if (gpos == pos) { d[e] = cd; id[e] = cid; }
    }
    int te = (k - 1) >> 5, tl = (k - 1) & 31;
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e == te) { thr_d = shfl_f64(d[e], tl); thr_id = __shfl_sync(kFull, id[e], tl); }
  }
};

// Exchange step of a warp-wide bitonic network on (d, id) keys, one key per lane: the lane keeps the smaller of
// its own and its partner's key (lane ^ j) when `keep_min`, else the larger.
__device__ __forceinline__ void bitonic_step(double& d, int& id, int j, bool keep_min) {
  const double pd = shfl_xor_f64(d, j);
  const int pi = __shfl_xor_sync(kFull, id, j);
  const bool partner_less = cand_less(pd, pi, d, id);
  if (partner_less == keep_min) { d = pd; id = pi; }
}

// k <= 32 (one list entry per lane): fold a whole batch of 32 candidates into the sorted list at once -- sort the
// batch (15 exchange steps), take the element-wise minimum with the reversed list (the 32 smallest of the union, a
// bitonic sequence), merge (5 steps), cut at k.  ~260 warp instructions whatever the number of accepted candidates,
// against ~35 per accepted candidate for the one-by-one insertion: pays when more than a handful are accepted, i.e.
// in the first batches of a query, where nearly every candidate is.
__device__ __forceinline__ void merge_batch_k32(WarpTopK<1>& tk, double cd, int cid, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) bitonic_step(cd, cid, j, ((lane & j) == 0) == ((lane & k2) == 0));
  }
  const double rd = shfl_f64(cd, 31 - lane);  // candidates descending
  const int ri = __shfl_sync(kFull, cid, 31 - lane);
  if (cand_less(rd, ri, tk.d[0], tk.id[0])) { tk.d[0] = rd; tk.id[0] = ri; }
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) bitonic_step(tk.d[0], tk.id[0], j, (lane & j) == 0);
  if (lane >= tk.k) { tk.d[0] = INFINITY; tk.id[0] = INT_MAX; }
  tk.thr_d = shfl_f64(tk.d[0], tk.k - 1);
  tk.thr_id = __shfl_sync(kFull, tk.id[0], tk.k - 1);
}

constexpr int kMergeBatchMin = 5;  // accepted candidates per batch from which the batch merge is cheaper (measured flat between 3 and 8: C5 5.86-5.95 ms; never merging: 9.6 ms)

// Stream the sorted range [b, e): lane-strided candidates, ballot the ones beating the threshold,
// insert them one by one (warp-uniform) or, for k <= 32 and a batch with many of them, all at once.
template <int E>
__device__ __forceinline__ void scan_range_knn(WarpTopK<E>& tk, int b, int e, double qx, double qy,
                                               int self_id, const double* __restrict__ xs,
                                               const double* __restrict__ ys,
                                               const int32_t* __restrict__ order, int lane) {
  for (int base = b; base < e; base += 32) {
    int s = base + lane;
    bool valid = s < e;
    double cd = INFINITY;
    int cid = INT_MAX;
    if (valid) {
      cid = order[s];
      cd = sq_dist(qx, qy, xs[s], ys[s]);
      valid = cid != self_id;
    }
    unsigned m = __ballot_sync(kFull, valid && cand_less(cd, cid, tk.thr_d, tk.thr_id));
    if constexpr (E == 1) {
      if (__popc(m) >= kMergeBatchMin) {
        merge_batch_k32(tk, valid ? cd : (double)INFINITY, valid ? cid : INT_MAX, lane);
        continue;
      }
    }
    while (m) {
      int src = __ffs(m) - 1;
      m &= m - 1;
      double bd = shfl_f64(cd, src);
      int bi = __shfl_sync(kFull, cid, src);
      if (cand_less(bd, bi, tk.thr_d, tk.thr_id)) tk.insert(bd, bi, lane);
    }
  }
}

template <int E>
__global__ void __launch_bounds__(256)
knn_query_kernel(const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                 const double* __restrict__ xs, const double* __restrict__ ys,
                 const int32_t* __restrict__ order, int64_t n, int k, int include_self,
                 int32_t* __restrict__ idx_out, double* __restrict__ dist_out,
                 const int32_t* __restrict__ labels, int n_types, float* __restrict__ profile,
                 const int32_t* __restrict__ todo, const int* __restrict__ todo_count) {
  extern __shared__ int s_hist[];  // [warps][n_types] when the fused composition is requested
  const GridParams g = *gp;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  int* hist = s_hist + warp_in_block * n_types;
  if (profile)
    for (int t = lane; t < n_types; t += 32) hist[t] = 0;

  // todo == NULL: contiguous chunk of sorted positions per block (neighbouring queries share candidate
  // cells in L1).  todo != NULL: the queries the tile kernel could not settle within its 3x3 block.
  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  int64_t s_begin = blockIdx.x * per_block;
  int64_t s_end = min(n, s_begin + per_block);
  int64_t stride = warps_per_block;
  if (todo) {
    s_begin = (int64_t)blockIdx.x * warps_per_block;
    s_end = *todo_count;
    stride = (int64_t)gridDim.x * warps_per_block;
  }

  for (int64_t it = s_begin + warp_in_block; it < s_end; it += stride) {
    const int64_t s = todo ? todo[it] : it;
    const double qx = xs[s], qy = ys[s];
    const int self_id = order[s];
    const double ux = (qx - g.x0) * g.inv_h, uy = (qy - g.y0) * g.inv_h;
    const int cx = min(max((int)floor(ux), 0), g.nx - 1);
    const int cy = min(max((int)floor(uy), 0), g.ny - 1);

    WarpTopK<E> tk;
    tk.init(k);
    {
      int c = cy * g.nx + cx;
      scan_range_knn<E>(tk, cell_start[c], cell_start[c + 1], qx, qy, self_id, xs, ys, order, lane);
    }
    for (int r = 1;; ++r) {
      const int xlo = max(cx - r, 0), xhi = min(cx + r, g.nx - 1);
      for (int yy = max(cy - r, 0); yy <= min(cy + r, g.ny - 1); ++yy) {
        const int rowbase = yy * g.nx;
        if (yy == cy - r || yy == cy + r) {
          scan_range_knn<E>(tk, cell_start[rowbase + xlo], cell_start[rowbase + xhi + 1], qx, qy,
                            self_id, xs, ys, order, lane);
        } else {
          if (cx - r >= 0)
            scan_range_knn<E>(tk, cell_start[rowbase + cx - r], cell_start[rowbase + cx - r + 1],
                              qx, qy, self_id, xs, ys, order, lane);
          if (cx + r <= g.nx - 1)
            scan_range_knn<E>(tk, cell_start[rowbase + cx + r], cell_start[rowbase + cx + r + 1],
                              qx, qy, self_id, xs, ys, order, lane);
        }
      }
      // exactness: every unseen point lies outside the (2r+1)² block; stop when the k-th best is
      // closer than the nearest block side that still has cells beyond it.
      double gap = INFINITY;
      if (cx - r > 0) gap = fmin(gap, ux - (double)(cx - r));
      if (cx + r < g.nx - 1) gap = fmin(gap, (double)(cx + r + 1) - ux);
      if (cy - r > 0) gap = fmin(gap, uy - (double)(cy - r));
      if (cy + r < g.ny - 1) gap = fmin(gap, (double)(cy + r + 1) - uy);
      if (isinf(gap)) break;
      double safe = gap * g.h * (1.0 - 1e-6) - g.margin;
      if (safe > 0 && tk.thr_d <= safe * safe) break;
    }

    // ---- epilogue: rank by column index and write row `self_id` ------------------------------
    const int kk = k + (include_self ? 1 : 0);
    const int64_t row = (int64_t)self_id * kk;
    int rank[E];
    int self_rank = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) rank[e] = 0;
    for (int j = 0; j < k; ++j) {
      int idj = 0;
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (e == (j >> 5)) idj = __shfl_sync(kFull, tk.id[e], j & 31);
#pragma unroll
      for (int e = 0; e < E; ++e) rank[e] += idj < tk.id[e];
      self_rank += idj < self_id;
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (e * 32 + lane < k) {
        int rk = rank[e] + ((include_self && tk.id[e] > self_id) ? 1 : 0);
        if (idx_out) idx_out[row + rk] = tk.id[e];
        if (dist_out) dist_out[row + rk] = sqrt(tk.d[e]);
        if (profile) atomicAdd(&hist[labels[tk.id[e]]], 1);
      }
    }
    if (include_self && lane == 0) {
      if (idx_out) idx_out[row + self_rank] = self_id;
      if (dist_out) dist_out[row + self_rank] = 0.0;
      if (profile) atomicAdd(&hist[labels[self_id]], 1);
    }
    if (profile) {
      __syncwarp();
      float* prow = profile + (int64_t)self_id * n_types;
      for (int t = lane; t < n_types; t += 32) { prow[t] = (float)hist[t]; hist[t] = 0; }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kNN, k <= 32: one THREAD per query over a shared-memory staged cell-bin tile
//
// A CTA owns `tx` consecutive cells of one cell row.  The candidate points of the three cell rows
// around it (cells cx0-1 .. cx1) are three contiguous runs of the sorted arrays; they are staged in
// shared memory once (x, y, id = 20 B per point).  Every thread owns one query and keeps its k best
// in a max-heap in shared memory ([slot][thread]: conflict-free); a warp walks the UNION of its
// lanes' 3x3 neighbourhoods, so all lanes read the same candidate (a broadcast) and the loop trip
// count is warp-uniform -- only the heap update diverges.  Per query this costs ~250 warp
// instructions against ~2500 for the warp-per-query kernel (whose sorted-list insertion is ~40
// warp-wide instructions per accepted candidate).  A query whose k-th distance does not clear the
// edge of its own 3x3 block is appended to `todo` and finished by knn_query_kernel's ring search.
// ------------------------------------------------------------------------------------------------

constexpr int kTileThreads = 256;
constexpr int kTileCap = 1280;   // staged candidate points per tile (25 KB); larger tiles read global
constexpr int kTileMaxK = 16;   // measured: k=30 on clustered data is faster on the warp-per-query kernel (round 1: 9.4 vs 11.3 ms at 2 M; 5.8 ms with its batch merge)
constexpr int kTileMaxTypes = 64;
constexpr int kTileSplit = 4;

__global__ void __launch_bounds__(kTileThreads)
knn_tile_kernel(const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                const double* __restrict__ xs, const double* __restrict__ ys,
                const int32_t* __restrict__ order, int k, int include_self, int tx,
                int32_t* __restrict__ idx_out, double* __restrict__ dist_out,
                const int32_t* __restrict__ labels, int n_types, float* __restrict__ profile,
                int32_t* __restrict__ todo, int* __restrict__ todo_count, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  __shared__ int s_work;
  double* cand_x = reinterpret_cast<double*>(tile_smem);
  double* cand_y = cand_x + kTileCap;
  double* heap_d = cand_y + kTileCap;                                   // [k][T]
  int* cand_id = reinterpret_cast<int*>(heap_d + (size_t)k * kTileThreads);
  int* heap_i = cand_id + kTileCap;                                     // [k][T]
  int* hist = heap_i + (size_t)k * kTileThreads;                        // [n_types][T] (profile only)
  unsigned char* rank_b = reinterpret_cast<unsigned char*>(hist + (profile ? (size_t)n_types * kTileThreads : 0));  // [k][T]

  const GridParams g = *gp;
  const int tid = threadIdx.x, lane = tid & 31, wbase = tid & ~31;
  const int tiles_x = (g.nx + tx - 1) / tx;
  const int n_tiles = tiles_x * g.ny;
  const int kk = k + (include_self ? 1 : 0);
  double* hd = heap_d + tid;
  int* hi = heap_i + tid;

  // Work units are handed out dynamically (clustered data makes tiles very uneven): unit w = tile
  // w / kTileSplit, whose query chunks q0 = qs + (y + j*kTileSplit)*128 go to sub-unit y = w % kTileSplit.
  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = atomicAdd(work_counter, 1);
    __syncthreads();
    const int w = s_work;
    if (w >= n_tiles * kTileSplit) break;
    const int tile = w / kTileSplit, ysub = w - tile * kTileSplit;
    const int cy = tile / tiles_x;
    const int cx0 = (tile - cy * tiles_x) * tx, cx1 = min(cx0 + tx, g.nx);
    const int qs = cell_start[cy * g.nx + cx0], qe = cell_start[cy * g.nx + cx1];
    if (qs + ysub * kTileThreads >= qe) continue;  // block-uniform
    const int xlo = max(cx0 - 1, 0), xhi = min(cx1, g.nx - 1);
    int rb[3], rn[3], total = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int yy = cy + d - 1;
      rb[d] = 0; rn[d] = 0;
      if (yy >= 0 && yy < g.ny) {
        rb[d] = cell_start[yy * g.nx + xlo];
        rn[d] = cell_start[yy * g.nx + xhi + 1] - rb[d];
      }
      total += rn[d];
    }
    const bool staged = total <= kTileCap;
    __syncthreads();  // the previous tile's readers are done with the staging area
    if (staged) {
      int off = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        for (int t = tid; t < rn[d]; t += kTileThreads) {
          cand_x[off + t] = xs[rb[d] + t];
          cand_y[off + t] = ys[rb[d] + t];
          cand_id[off + t] = order[rb[d] + t];
        }
        off += rn[d];
      }
    }
    __syncthreads();

    for (int q0 = qs + ysub * kTileThreads; q0 < qe; q0 += kTileSplit * kTileThreads) {
      const int s = q0 + tid;
      const bool valid = s < qe;
      double qx = 0, qy = 0, ux = 0, uy = 0;
      int self_id = -1, cxq = cx0;
      if (valid) {
        qx = xs[s]; qy = ys[s]; self_id = order[s];
        ux = (qx - g.x0) * g.inv_h; uy = (qy - g.y0) * g.inv_h;
        cxq = min(max((int)floor(ux), 0), g.nx - 1);
      }
      const unsigned vmask = __ballot_sync(kFull, valid);
      if (vmask == 0) continue;  // warp-uniform
      const int wlo = max(__reduce_min_sync(kFull, valid ? cxq : INT_MAX) - 1, xlo);
      const int whi = min(__reduce_max_sync(kFull, valid ? cxq : INT_MIN) + 1, xhi);
      for (int e = 0; e < k; ++e) { hd[e * kTileThreads] = INFINITY; hi[e * kTileThreads] = INT_MAX; }
      double thr_d = INFINITY;
      int thr_id = INT_MAX;

      // replace the root of the max-heap by (cd, cid) and sift it down; refresh the threshold
      auto heap_replace = [&](double cd, int cid) {
        int pos = 0;
        for (;;) {
          const int l = 2 * pos + 1;
          if (l >= k) break;
          int ch = l;
          double chd = hd[l * kTileThreads];
          int chi = hi[l * kTileThreads];
          if (l + 1 < k) {
            const double rd = hd[(l + 1) * kTileThreads];
            const int ri = hi[(l + 1) * kTileThreads];
            if (cand_less(chd, chi, rd, ri)) { ch = l + 1; chd = rd; chi = ri; }
          }
          if (!cand_less(cd, cid, chd, chi)) break;
          hd[pos * kTileThreads] = chd; hi[pos * kTileThreads] = chi;
          pos = ch;
        }
        hd[pos * kTileThreads] = cd; hi[pos * kTileThreads] = cid;
        thr_d = hd[0]; thr_id = hi[0];
      };

      // Candidates are screened 16 at a time against the lane's current threshold (all lanes read the
      // same candidate: a broadcast, no divergence); the accepted ones are then popped from a bit mask
      // and pushed through the heap.  Screening in blocks keeps the divergent heap code to
      // max-over-lanes(accepted per block) executions instead of one per candidate.  The query's own
      // cell row is scanned first so the threshold tightens early.
#pragma unroll 1
      for (int dd = 0; dd < 3; ++dd) {
        const int d = dd == 0 ? 1 : (dd == 1 ? 0 : 2);
        const int yy = cy + d - 1;
        if (yy < 0 || yy >= g.ny) continue;
        const int soff = d == 0 ? 0 : (d == 1 ? rn[0] : rn[0] + rn[1]);
        const int rbd = d == 0 ? rb[0] : (d == 1 ? rb[1] : rb[2]);
        const int b = cell_start[yy * g.nx + wlo], e = cell_start[yy * g.nx + whi + 1];
        const double* px = staged ? cand_x + soff + (b - rbd) : xs + b;
        const double* py = staged ? cand_y + soff + (b - rbd) : ys + b;
        const int* pid = staged ? cand_id + soff + (b - rbd) : order + b;
        const int cnt = e - b;
        for (int c0 = 0; c0 < cnt; c0 += 16) {
          unsigned mask = 0;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int c = c0 + u;
            if (c < cnt) {
              const int cid = pid[c];
              const double cd = sq_dist(qx, qy, px[c], py[c]);
              if (cid != self_id && cand_less(cd, cid, thr_d, thr_id)) mask |= 1u << u;
            }
          }
          if (!valid) mask = 0;
          while (__any_sync(kFull, mask != 0)) {
            if (mask) {
              const int c = c0 + __ffs(mask) - 1;
              mask &= mask - 1;
              const int cid = pid[c];
              const double cd = sq_dist(qx, qy, px[c], py[c]);
              if (cand_less(cd, cid, thr_d, thr_id)) heap_replace(cd, cid);
            }
          }
        }
      }

      // exactness: every unseen point lies outside this query's own 3x3 block
      bool settled = valid;
      if (valid) {
        const int cyq = cy;
        double gap = INFINITY;
        if (cxq - 1 > 0) gap = fmin(gap, ux - (double)(cxq - 1));
        if (cxq + 1 < g.nx - 1) gap = fmin(gap, (double)(cxq + 2) - ux);
        if (cyq - 1 > 0) gap = fmin(gap, uy - (double)(cyq - 1));
        if (cyq + 1 < g.ny - 1) gap = fmin(gap, (double)(cyq + 2) - uy);
        if (!isinf(gap)) {
          const double safe = gap * g.h * (1.0 - 1e-6) - g.margin;
          settled = safe > 0 && thr_d <= safe * safe;
        }
        if (!settled) todo[atomicAdd(todo_count, 1)] = s;
      }

      // rank the k results by column index (canonical CSR order)
      int self_rank = 0;
      if (settled) {
        for (int a = 0; a < k; ++a) {
          const int ida = hi[a * kTileThreads];
          int rk = (include_self && self_id < ida) ? 1 : 0;
          for (int b2 = 0; b2 < k; ++b2) rk += hi[b2 * kTileThreads] < ida;
          rank_b[a * kTileThreads + tid] = (unsigned char)rk;
          self_rank += ida < self_id;
        }
        if (profile) {
          for (int t = 0; t < n_types; ++t) hist[t * kTileThreads + tid] = 0;
          for (int a = 0; a < k; ++a) hist[labels[hi[a * kTileThreads]] * kTileThreads + tid] += 1;
          if (include_self) hist[labels[self_id] * kTileThreads + tid] += 1;
        }
      }
      __syncwarp();
      // cooperative, row-contiguous writes: lane L emits slot L of query ql
      unsigned smask = __ballot_sync(kFull, settled);
      while (smask) {
        const int ql = __ffs(smask) - 1;
        smask &= smask - 1;
        const int64_t row = (int64_t)__shfl_sync(kFull, self_id, ql);
        const int srk = __shfl_sync(kFull, self_rank, ql);
        const int col = wbase + ql;
        if (lane < k) {
          const int rk = rank_b[lane * kTileThreads + col];
          if (idx_out) idx_out[row * kk + rk] = heap_i[lane * kTileThreads + col];
          if (dist_out) dist_out[row * kk + rk] = sqrt(heap_d[lane * kTileThreads + col]);
        }
        if (include_self && lane == 31) {
          if (idx_out) idx_out[row * kk + srk] = (int)row;
          if (dist_out) dist_out[row * kk + srk] = 0.0;
        }
        if (profile)
          for (int t = lane; t < n_types; t += 32) profile[row * n_types + t] = (float)hist[t * kTileThreads + col];
      }
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// radius graph: count pass, scan, fill pass with in-row rank sort
// ------------------------------------------------------------------------------------------------

constexpr int kFillCap = 64;

template <bool FILL>
__global__ void __launch_bounds__(256, 5)
radius_query_kernel(const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                    const double* __restrict__ xs, const double* __restrict__ ys,
                    const int32_t* __restrict__ order, int64_t n, double r2,
                    int32_t* __restrict__ deg_out,          // COUNT: degree per original row
                    const int32_t* __restrict__ indptr,     // FILL
                    int32_t* tmp_idx, double* tmp_dist,     // FILL scratch (read back: no restrict)
                    int32_t* __restrict__ indices, double* __restrict__ dist,
                    const int32_t* __restrict__ labels, int n_types, float* __restrict__ profile,
                    const int32_t* __restrict__ todo, const int* __restrict__ todo_count) {
  extern __shared__ int s_hist[];
  // FILL: a row's hits wait for their rank in shared memory (rows of up to kFillCap neighbours; longer rows go
  // through the global scratch).  Reading them back from the scratch costs an L2 round trip per row.
  __shared__ int s_fill_idx[FILL ? 8 * kFillCap : 1];
  __shared__ double s_fill_dist[FILL ? 8 * kFillCap : 1];
  const GridParams g = *gp;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  int* hist = s_hist + warp_in_block * n_types;
  int* fill_idx = s_fill_idx + (FILL ? warp_in_block * kFillCap : 0);
  double* fill_dist = s_fill_dist + (FILL ? warp_in_block * kFillCap : 0);
  if (!FILL && profile)
    for (int t = lane; t < n_types; t += 32) hist[t] = 0;

  const int64_t per_block = (n + gridDim.x - 1) / gridDim.x;
  int64_t s_begin = blockIdx.x * per_block;
  int64_t s_end = min(n, s_begin + per_block);
  int64_t stride = warps_per_block;
  if (todo) {  // only the rows the tile kernel could not hold in shared memory
    s_begin = (int64_t)blockIdx.x * warps_per_block;
    s_end = *todo_count;
    stride = (int64_t)gridDim.x * warps_per_block;
  }

  for (int64_t it = s_begin + warp_in_block; it < s_end; it += stride) {
    const int64_t s = todo ? todo[it] : it;
    const double qx = xs[s], qy = ys[s];
    const int self_id = order[s];
    const int cx = cell_coord(qx, g.x0, g.inv_h, g.nx);
    const int cy = cell_coord(qy, g.y0, g.inv_h, g.ny);
    const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.nx - 1);
    int count = 0;
    const int64_t start = FILL ? (int64_t)indptr[self_id] : 0;
    const bool in_smem = FILL && indptr[FILL ? self_id + 1 : 0] - (int)start <= kFillCap;
    // the three cell rows of the 3x3 block as one flattened candidate list: their six range lookups are issued
    // together (one exposed latency instead of three) and the 32-wide batches are fuller
    int rb0 = 0, rb1 = 0, rb2 = 0, rn0 = 0, rn1 = 0, rn2 = 0;
    if (cy > 0) { rb0 = cell_start[(cy - 1) * g.nx + xlo]; rn0 = cell_start[(cy - 1) * g.nx + xhi + 1] - rb0; }
    { rb1 = cell_start[cy * g.nx + xlo]; rn1 = cell_start[cy * g.nx + xhi + 1] - rb1; }
    if (cy + 1 < g.ny) { rb2 = cell_start[(cy + 1) * g.nx + xlo]; rn2 = cell_start[(cy + 1) * g.nx + xhi + 1] - rb2; }
    const int n01 = rn0 + rn1, n_cand = n01 + rn2;
    {
      for (int base = 0; base < n_cand; base += 32) {
        const int u = base + lane;
        bool hit = false;
        int cid = 0;
        double d2 = 0;
        if (u < n_cand) {
          const int t = u < rn0 ? rb0 + u : (u < n01 ? rb1 + (u - rn0) : rb2 + (u - n01));
          cid = order[t];
          d2 = sq_dist(qx, qy, xs[t], ys[t]);
          hit = (cid != self_id) && (d2 <= r2);
        }
        unsigned m = __ballot_sync(kFull, hit);
        if (FILL) {
          if (hit) {
            int off = count + __popc(m & ((1u << lane) - 1u));
            if (in_smem) {
              fill_idx[off] = cid;
              if (dist) fill_dist[off] = sqrt(d2);
            } else {
              tmp_idx[start + off] = cid;
              if (tmp_dist) tmp_dist[start + off] = sqrt(d2);
            }
          }
        } else if (profile && hit) {
          atomicAdd(&hist[labels[cid]], 1);
        }
        count += __popc(m);
      }
    }
    if (!FILL) {
      if (lane == 0 && deg_out) deg_out[self_id] = count;
      if (profile) {
        __syncwarp();
        float* prow = profile + (int64_t)self_id * n_types;
        for (int t = lane; t < n_types; t += 32) { prow[t] = (float)hist[t]; hist[t] = 0; }
        __syncwarp();
      }
    } else {
      __syncwarp();  // orders the warp's writes before the reads below
      if (in_smem) {
        for (int e = lane; e < count; e += 32) {
          const int my = fill_idx[e];
          int rk = 0;
          for (int j = 0; j < count; ++j) rk += fill_idx[j] < my;
          indices[start + rk] = my;
          if (dist) dist[start + rk] = fill_dist[e];
        }
        __syncwarp();  // the next row overwrites the buffer
      } else {
        for (int e = lane; e < count; e += 32) {
          int my = tmp_idx[start + e];
          int rk = 0;
          for (int j = 0; j < count; ++j) rk += tmp_idx[start + j] < my;
          indices[start + rk] = my;
          if (dist) dist[start + rk] = tmp_dist[start + e];
        }
      }
    }
  }
}

// Thread-per-query degree count of the radius graph over a shared-memory staged tile (same tiling as
// knn_tile_kernel): 0.64 ms at 5 M cells / 10^8 edges against 1.39 ms for the warp-per-query count.
// (The fill pass stays warp-per-query: collecting, ranking and writing a row is warp-cooperative work.
// Thread-per-query fills over the same tiles were measured twice -- round 1: the same 2.4-2.5 ms; round 2,
// private shared-memory lists + insertion sort + per-thread row writes: 2.8 ms against 2.1 ms.)
__global__ void __launch_bounds__(kTileThreads)
radius_count_tile_kernel(const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                   const double* __restrict__ xs, const double* __restrict__ ys,
                   const int32_t* __restrict__ order, double r2, int tx,
                   int32_t* __restrict__ deg_out, int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  __shared__ int s_work;
  double* cand_x = reinterpret_cast<double*>(tile_smem);
  double* cand_y = cand_x + kTileCap;
  int* cand_id = reinterpret_cast<int*>(cand_y + kTileCap);

  const GridParams g = *gp;
  const int tid = threadIdx.x;
  const int tiles_x = (g.nx + tx - 1) / tx;
  const int n_tiles = tiles_x * g.ny;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = atomicAdd(work_counter, 1);
    __syncthreads();
    const int w = s_work;
    if (w >= n_tiles * kTileSplit) break;
    const int tile = w / kTileSplit, ysub = w - tile * kTileSplit;
    const int cy = tile / tiles_x;
    const int cx0 = (tile - cy * tiles_x) * tx, cx1 = min(cx0 + tx, g.nx);
    const int qs = cell_start[cy * g.nx + cx0], qe = cell_start[cy * g.nx + cx1];
    if (qs + ysub * kTileThreads >= qe) continue;  // block-uniform
    const int xlo = max(cx0 - 1, 0), xhi = min(cx1, g.nx - 1);
    int rb[3], rn[3], total = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int yy = cy + d - 1;
      rb[d] = 0; rn[d] = 0;
      if (yy >= 0 && yy < g.ny) {
        rb[d] = cell_start[yy * g.nx + xlo];
        rn[d] = cell_start[yy * g.nx + xhi + 1] - rb[d];
      }
      total += rn[d];
    }
    const bool staged = total <= kTileCap;
    if (staged) {
      int off = 0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        for (int t = tid; t < rn[d]; t += kTileThreads) {
          cand_x[off + t] = xs[rb[d] + t];
          cand_y[off + t] = ys[rb[d] + t];
          cand_id[off + t] = order[rb[d] + t];
        }
        off += rn[d];
      }
    }
    __syncthreads();

    for (int q0 = qs + ysub * kTileThreads; q0 < qe; q0 += kTileSplit * kTileThreads) {
      const int s = q0 + tid;
      const bool valid = s < qe;
      double qx = 0, qy = 0;
      int self_id = -1, cxq = cx0;
      if (valid) {
        qx = xs[s]; qy = ys[s]; self_id = order[s];
        cxq = cell_coord(qx, g.x0, g.inv_h, g.nx);
      }
      if (__ballot_sync(kFull, valid) == 0) continue;  // warp-uniform
      const int wlo = max(__reduce_min_sync(kFull, valid ? cxq : INT_MAX) - 1, xlo);
      const int whi = min(__reduce_max_sync(kFull, valid ? cxq : INT_MIN) + 1, xhi);
      int count = 0;
#pragma unroll 1
      for (int d = 0; d < 3; ++d) {
        const int yy = cy + d - 1;
        if (yy < 0 || yy >= g.ny) continue;
        const int soff = d == 0 ? 0 : (d == 1 ? rn[0] : rn[0] + rn[1]);
        const int rbd = d == 0 ? rb[0] : (d == 1 ? rb[1] : rb[2]);
        const int b = cell_start[yy * g.nx + wlo], e = cell_start[yy * g.nx + whi + 1];
        const double* px = staged ? cand_x + soff + (b - rbd) : xs + b;
        const double* py = staged ? cand_y + soff + (b - rbd) : ys + b;
        const int* pid = staged ? cand_id + soff + (b - rbd) : order + b;
        const int cnt = e - b;
        for (int c = 0; c < cnt; ++c) {
          const int cid = pid[c];
          const double d2 = sq_dist(qx, qy, px[c], py[c]);
          count += (valid && cid != self_id && d2 <= r2) ? 1 : 0;
        }
      }
      if (valid) deg_out[self_id] = count;
    }
  }
}

__global__ void sum_degrees_kernel(const int32_t* __restrict__ deg, int64_t n,
                                   unsigned long long* __restrict__ total) {
  unsigned long long acc = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc += (unsigned long long)deg[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(total, acc);
}

// ------------------------------------------------------------------------------------------------
// neighbourhood composition from an existing graph; profile normalisation
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
nbhd_counts_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                   int64_t n, int k_fixed, const int32_t* __restrict__ labels, int n_types,
                   float* __restrict__ profile) {
  extern __shared__ int s_hist[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  int* hist = s_hist + w * n_types;
  for (int t = lane; t < n_types; t += 32) hist[t] = 0;
  __syncwarp();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + w; i < n; i += warps) {
    int64_t b = indptr ? indptr[i] : i * k_fixed;
    int64_t e = indptr ? indptr[i + 1] : b + k_fixed;
    for (int64_t j = b + lane; j < e; j += 32) atomicAdd(&hist[labels[indices[j]]], 1);
    __syncwarp();
    float* prow = profile + i * n_types;
    for (int t = lane; t < n_types; t += 32) { prow[t] = (float)hist[t]; hist[t] = 0; }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256)
profile_normalize_kernel(float* __restrict__ profile, int64_t n, int n_types, int normalize,
                         unsigned long long* __restrict__ n_empty) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  unsigned long long empty = 0;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n;
       i += warps) {
    float* prow = profile + i * n_types;
    float sum = 0.f;  // counts are small integers: exact in FP32 in any order
    for (int t = lane; t < n_types; t += 32) sum += prow[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
    if (sum == 0.f) { empty += (lane == 0); continue; }
    if (normalize)
      for (int t = lane; t < n_types; t += 32) prow[t] = __fdiv_rn(prow[t], sum);
  }
  if (lane == 0 && empty) atomicAdd(n_empty, empty);
}

// ------------------------------------------------------------------------------------------------
// graph moments
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ double edge_weight(const float* __restrict__ weights, int64_t e,
                                              int deg) {
  return weights ? (double)weights[e] : 1.0 / (double)deg;
}

// warp per row: s0/s1 partials (deterministic two-stage), rowsum[i], colsum scatter (FP64 atomics).
__global__ void __launch_bounds__(256)
moments_edges_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                     const float* __restrict__ weights, int64_t n, int k_fixed,
                     double* __restrict__ rowsum, double* __restrict__ colsum,
                     double* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  double s0 = 0, s1 = 0;
  for (int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + w; i < n; i += warps) {
    int64_t b = indptr ? indptr[i] : i * k_fixed;
    int64_t e = indptr ? indptr[i + 1] : b + k_fixed;
    int deg = (int)(e - b);
    double rs = 0;
    for (int64_t t = b + lane; t < e; t += 32) {
      int j = indices[t];
      double wij = edge_weight(weights, t, deg);
      // reverse edge (j -> i): binary search in row j (column-sorted)
      int64_t jb = indptr ? indptr[j] : (int64_t)j * k_fixed;
      int64_t je = indptr ? indptr[j + 1] : jb + k_fixed;
      int degj = (int)(je - jb);
      int64_t lo = jb, hi = je;
      while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (indices[mid] < (int)i) lo = mid + 1; else hi = mid;
      }
      bool rev = lo < je && indices[lo] == (int)i;
      if (rev) {
        double t2 = wij + edge_weight(weights, lo, degj);
        s1 += 0.5 * t2 * t2;
      } else {
        s1 += wij * wij;
      }
      rs += wij;
      atomicAdd(&colsum[j], wij);
    }
    rs = warp_sum(rs);
    if (lane == 0) rowsum[i] = rs;
    s0 += (lane == 0) ? rs : 0.0;
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  __shared__ double sh[2][8];
  if (lane == 0) { sh[0][w] = s0; sh[1][w] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b2 = 0;
    for (int j = 0; j < (blockDim.x >> 5); ++j) { a += sh[0][j]; b2 += sh[1][j]; }
    partial[2 * blockIdx.x] = a;
    partial[2 * blockIdx.x + 1] = b2;
  }
}

__global__ void __launch_bounds__(256)
moments_s2_kernel(const double* __restrict__ rowsum, const double* __restrict__ colsum, int64_t n,
                  double* __restrict__ partial) {
  double s2 = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double t = rowsum[i] + colsum[i];
    s2 += t * t;
  }
  s2 = warp_sum(s2);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0;
    for (int j = 0; j < (blockDim.x >> 5); ++j) a += sh[j];
    partial[blockIdx.x] = a;
  }
}

__global__ void moments_final_kernel(const double* __restrict__ p01, const double* __restrict__ p2,
                                     int nblocks, double* __restrict__ out) {
  double s0 = 0, s1 = 0, s2 = 0;
  for (int j = 0; j < nblocks; ++j) { s0 += p01[2 * j]; s1 += p01[2 * j + 1]; s2 += p2[j]; }
  out[0] = s0; out[1] = s1; out[2] = s2;
}


// ------------------------------------------------------------------------------------------------
// spatial (grid) order and graph relabelling
//
// The statistics are sums over cells, invariant under any relabelling of the cells.  Storing Z / lag
// in grid order makes a row's neighbours close in memory, so the lag SpMM reads them from L1/L2
// instead of HBM.  order[a] = original id of the cell at sorted position a; rank = its inverse.
// ------------------------------------------------------------------------------------------------

__global__ void invert_order_kernel(const int32_t* __restrict__ order, int64_t n,
                                    int32_t* __restrict__ rank) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a < n;
       a += (int64_t)gridDim.x * blockDim.x)
    rank[order[a]] = (int32_t)a;
}

__global__ void gather_degrees_kernel(const int32_t* __restrict__ indptr,
                                      const int32_t* __restrict__ order, int64_t n,
                                      int32_t* __restrict__ out) {
  for (int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; a <= n;
       a += (int64_t)gridDim.x * blockDim.x) {
    if (a == n) { out[a] = 0; continue; }
    int i = order[a];
    out[a] = indptr[i + 1] - indptr[i];
  }
}

// warp per output row a (= original row order[a]): map columns through rank, rank-sort within the
// row, write (weights follow their edge).
__global__ void __launch_bounds__(256)
relabel_rows_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                    const float* __restrict__ weights, int64_t n, int k_fixed,
                    const int32_t* __restrict__ order, const int32_t* __restrict__ rank,
                    const int32_t* __restrict__ out_indptr, int32_t* __restrict__ out_indices,
                    float* __restrict__ out_weights) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t a = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); a < n; a += warps) {
    const int64_t i = order[a];
    const int64_t sb = indptr ? indptr[i] : i * k_fixed;
    const int64_t se = indptr ? indptr[i + 1] : sb + k_fixed;
    const int64_t db = out_indptr ? out_indptr[a] : a * k_fixed;
    const int deg = (int)(se - sb);
    for (int e = lane; e < deg; e += 32) {
      const int mine = rank[indices[sb + e]];
      int rk = 0;
      for (int j = 0; j < deg; ++j) rk += rank[indices[sb + j]] < mine;
      out_indices[db + rk] = mine;
      if (out_weights) out_weights[db + rk] = weights[sb + e];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// cross-set nearest neighbour and pairwise distance reductions (calculate_domain_distances,
// [R spatial/distance.py:46-449]: cKDTree.query(k=1) between labelled subsets, cdist().min()/.mean())
// ------------------------------------------------------------------------------------------------

// One thread per query point (not a member of the binned target set): ring search over the target
// grid until the best distance clears the nearest unsearched cell.  Key (d2, target index).
__global__ void __launch_bounds__(256)
cross_nn_kernel(const GridParams* __restrict__ gp, const int32_t* __restrict__ cell_start,
                const double* __restrict__ xs, const double* __restrict__ ys,
                const int32_t* __restrict__ order, const double* __restrict__ queries, int64_t nq,
                int32_t* __restrict__ idx_out, double* __restrict__ dist_out) {
  const GridParams g = *gp;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq;
       q += (int64_t)gridDim.x * blockDim.x) {
    const double2 p = reinterpret_cast<const double2*>(queries)[q];
    const double ux = (p.x - g.x0) * g.inv_h, uy = (p.y - g.y0) * g.inv_h;
    const int cx = min(max((int)floor(fmin(fmax(ux, -1.0), (double)g.nx)), 0), g.nx - 1);
    const int cy = min(max((int)floor(fmin(fmax(uy, -1.0), (double)g.ny)), 0), g.ny - 1);
    double bd = INFINITY;
    int bi = INT_MAX;
    auto scan = [&](int b, int e) {
      for (int s = b; s < e; ++s) {
        const double d2 = sq_dist(p.x, p.y, xs[s], ys[s]);
        const int id = order[s];
        if (cand_less(d2, id, bd, bi)) { bd = d2; bi = id; }
      }
    };
    scan(cell_start[cy * g.nx + cx], cell_start[cy * g.nx + cx + 1]);
    for (int r = 1;; ++r) {
      const int xlo = max(cx - r, 0), xhi = min(cx + r, g.nx - 1);
      for (int yy = max(cy - r, 0); yy <= min(cy + r, g.ny - 1); ++yy) {
        const int rowbase = yy * g.nx;
        if (yy == cy - r || yy == cy + r) {
          scan(cell_start[rowbase + xlo], cell_start[rowbase + xhi + 1]);
        } else {
          if (cx - r >= 0) scan(cell_start[rowbase + cx - r], cell_start[rowbase + cx - r + 1]);
          if (cx + r <= g.nx - 1) scan(cell_start[rowbase + cx + r], cell_start[rowbase + cx + r + 1]);
        }
      }
      double gap = INFINITY;
      if (cx - r > 0) gap = fmin(gap, ux - (double)(cx - r));
      if (cx + r < g.nx - 1) gap = fmin(gap, (double)(cx + r + 1) - ux);
      if (cy - r > 0) gap = fmin(gap, uy - (double)(cy - r));
      if (cy + r < g.ny - 1) gap = fmin(gap, (double)(cy + r + 1) - uy);
      if (isinf(gap)) break;  // the block covers the whole grid
      const double safe = gap * g.h * (1.0 - 1e-6) - g.margin;
      if (safe > 0 && bd <= safe * safe) break;
    }
    idx_out[q] = bi;
    if (dist_out) dist_out[q] = sqrt(bd);
  }
}

// out partial[b] = {min over pairs of d, sum over pairs of d} for the A rows of block b against all
// of B (staged through shared memory in tiles).  FP64, fixed reduction order.
constexpr int kPairTile = 1024;
__global__ void __launch_bounds__(256)
pairwise_reduce_kernel(const double* __restrict__ A, int64_t na, const double* __restrict__ B,
                       int64_t nb, double* __restrict__ partial) {
  __shared__ double2 sb[kPairTile];
  __shared__ double sred[2][8];
  double vmin = INFINITY, vsum = 0.0;
  for (int64_t a0 = (int64_t)blockIdx.x * blockDim.x; a0 < na; a0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = a0 + threadIdx.x;
    const bool live = a < na;
    const double2 pa = live ? reinterpret_cast<const double2*>(A)[a] : make_double2(0, 0);
    for (int64_t b0 = 0; b0 < nb; b0 += kPairTile) {
      const int cnt = (int)min((int64_t)kPairTile, nb - b0);
      __syncthreads();
      for (int t = threadIdx.x; t < cnt; t += blockDim.x) sb[t] = reinterpret_cast<const double2*>(B)[b0 + t];
      __syncthreads();
      if (live) {
        double lmin = INFINITY, lsum = 0.0;
        for (int t = 0; t < cnt; ++t) {
          const double d2 = sq_dist(pa.x, pa.y, sb[t].x, sb[t].y);
          lmin = fmin(lmin, d2);
          lsum += sqrt(d2);
        }
        vmin = fmin(vmin, lmin);
        vsum += lsum;
      }
    }
  }
  // block reduction
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmin = fmin(vmin, shfl_f64(vmin, lane ^ o));
  }
  vsum = warp_sum(vsum);
  if (lane == 0) { sred[0][w] = vmin; sred[1][w] = vsum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = INFINITY, sm = 0.0;
    for (int j = 0; j < 8; ++j) { m = fmin(m, sred[0][j]); sm += sred[1][j]; }
    partial[2 * blockIdx.x] = m;
    partial[2 * blockIdx.x + 1] = sm;
  }
}

__global__ void pairwise_final_kernel(const double* __restrict__ partial, int blocks, double* __restrict__ out) {
  double m = INFINITY, sm = 0.0;
  for (int b = 0; b < blocks; ++b) { m = fmin(m, partial[2 * b]); sm += partial[2 * b + 1]; }
  out[0] = sqrt(m);
  out[1] = sm;
}

constexpr int kMomentBlocks = 592;

}  // namespace sc

using namespace sc;

// ================================================================================================
// C ABI
// ================================================================================================

extern "C" size_t sc_grid_knn_workspace_bytes(int64_t n, int k) {
  (void)k;
  if (n <= 0) return 256;
  return binning_bytes(n) + 1024;
}

extern "C" int sc_grid_knn(const double* coords, int64_t n, int k, int include_self, int32_t* idx,
                           double* dist, int32_t* order_out, const int32_t* labels, int n_types,
                           float* profile, void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(coords && ws, "sc_grid_knn: null coords/workspace");
  SC_CHECK_ARG(n >= 2 && n < (1ll << 31) - 2048, "sc_grid_knn: n=%lld out of range", (long long)n);
  SC_CHECK_ARG(k >= 1 && k <= SC_KNN_MAX_K, "sc_grid_knn: k=%d outside [1,%d]", k, SC_KNN_MAX_K);
  SC_CHECK_ARG(k < n, "k must be < number of cells (%lld), got %d", (long long)n, k);
  SC_CHECK_ARG(idx || profile, "sc_grid_knn: nothing to compute (idx and profile are NULL)");
  SC_CHECK_ARG(!profile || (labels && n_types >= 1 && n_types <= 2048),
               "sc_grid_knn: fused composition needs labels and 1 <= n_types <= 2048");
  if (ws_bytes < sc_grid_knn_workspace_bytes(n, k)) {
    set_error("sc_grid_knn: workspace %zu < %zu", ws_bytes, sc_grid_knn_workspace_bytes(n, k));
    return SC_ERR_WORKSPACE;
  }
  Arena arena(ws, ws_bytes);
  Binning b;
  if (!carve_binning(arena, n, &b)) { set_error("sc_grid_knn: workspace carve failed"); return SC_ERR_WORKSPACE; }
  // occupancy: a circle of radius h (guaranteed covered by the 3x3 block) holds ~pi*c points
  // (measured, B200: the thread-per-query tile kernel (k <= 16) likes 0.65 k -- k = 6 at 5 M cells 3.38 -> 2.97 ms,
  // k = 15 at 500 k 0.90 -> 0.83 ms, k = 15 at 5 M unchanged --, the warp-per-query kernel 0.5 k)
  double c = k <= kTileMaxK ? 0.65 * k : 0.5 * k;
  if (const char* e = getenv("SC_KNN_OCCUPANCY")) { double f = atof(e); if (f > 0.05 && f < 8.0) c = k * f; }  // experiments
  if (c < 1.0) c = 1.0;
  int rc = run_binning(coords, n, c, 0.0, b, st);
  if (rc) return rc;
  if (order_out)
    SC_CUDA_OK(cudaMemcpyAsync(order_out, b.order, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));

  const int threads = 256;
  int64_t want = (n + 255) / 256;  // ~256 queries per block
  int blocks = (int)(want < 1 ? 1 : want);
  size_t smem = profile ? sizeof(int) * (threads / 32) * (size_t)n_types : 0;
  // after binning, `keys` (n ints) and `partial` are free: todo list and its counter
  int32_t* todo = nullptr;
  int* todo_count = nullptr;
  const char* force = getenv("SC_KNN_VARIANT");  // "warp" forces the warp-per-query kernel
  const bool use_tile = k <= kTileMaxK && (!profile || n_types <= kTileMaxTypes) && !(force && !strcmp(force, "warp"));
  if (use_tile) {
    todo = b.keys;
    todo_count = reinterpret_cast<int*>(b.partial);
    int* work_counter = todo_count + 1;
    SC_CUDA_OK(cudaMemsetAsync(todo_count, 0, 2 * sizeof(int), st));
    int tx = (int)(kTileThreads / c + 0.5);
    if (tx < 2) tx = 2;
    if (tx > 64) tx = 64;
    size_t tsm = (size_t)kTileCap * 20 + (size_t)k * kTileThreads * 12 + (size_t)k * kTileThreads +
                 (profile ? sizeof(int) * (size_t)n_types * kTileThreads : 0) + 64;
    SC_CUDA_OK(cudaFuncSetAttribute(knn_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
    int per_sm = 1;
    SC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, knn_tile_kernel, kTileThreads, tsm));
    if (per_sm < 1) per_sm = 1;
    int tblocks = sm_count() * per_sm;  // persistent CTAs pulling work units from an atomic counter
    knn_tile_kernel<<<tblocks, kTileThreads, tsm, st>>>(b.gp, b.cell_start, b.xs, b.ys, b.order, k, include_self, tx,
                                                        idx, dist, labels, n_types, profile, todo, todo_count, work_counter);
    SC_LAUNCH_OK();
    blocks = sm_count() * 4;  // ring-search kernel over the unsettled queries only
  }
#define SC_KNN_LAUNCH(E)                                                                          \
  do {                                                                                            \
    if (smem > 48 * 1024)                                                                         \
      SC_CUDA_OK(cudaFuncSetAttribute(knn_query_kernel<E>,                                        \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    knn_query_kernel<E><<<blocks, threads, smem, st>>>(b.gp, b.cell_start, b.xs, b.ys, b.order, n, \
                                                      k, include_self, idx, dist, labels,         \
                                                      n_types, profile, todo, todo_count);        \
  } while (0)
  if (k <= 32) SC_KNN_LAUNCH(1);
  else if (k <= 64) SC_KNN_LAUNCH(2);
  else SC_KNN_LAUNCH(4);
#undef SC_KNN_LAUNCH
  SC_LAUNCH_OK();
  return SC_OK;
}


extern "C" size_t sc_spatial_order_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return binning_bytes(n) + 1024;
}

extern "C" int sc_spatial_order(const double* coords, int64_t n, int32_t* order_out,
                                int32_t* rank_out, void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(coords && order_out && ws, "sc_spatial_order: null argument");
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31) - 2048, "sc_spatial_order: n out of range");
  if (ws_bytes < sc_spatial_order_workspace_bytes(n)) { set_error("sc_spatial_order: workspace too small"); return SC_ERR_WORKSPACE; }
  Arena arena(ws, ws_bytes);
  Binning b;
  if (!carve_binning(arena, n, &b)) { set_error("sc_spatial_order: carve failed"); return SC_ERR_WORKSPACE; }
  // ~8 points per Z-curve cell is the measured default; SC_ORDER_POINTS_PER_CELL (1 .. 64) is an experiment
  // switch: a finer curve brings consecutive rows closer together (more shared neighbours for the row-group
  // lag kernel: simulated 5-8 % fewer gathers at 1-2 points per cell)
  double ppc = 8.0;
  if (const char* e = getenv("SC_ORDER_POINTS_PER_CELL")) { double v = atof(e); if (v >= 1.0 && v <= 64.0) ppc = v; }
  int rc = run_binning(coords, n, ppc, 0.0, b, st, /*order_only=*/true);
  if (rc) return rc;
  SC_CUDA_OK(cudaMemcpyAsync(order_out, b.order, sizeof(int32_t) * n, cudaMemcpyDeviceToDevice, st));
  if (rank_out) {
    int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
    invert_order_kernel<<<blocks, 256, 0, st>>>(order_out, n, rank_out);
    SC_LAUNCH_OK();
  }
  return SC_OK;
}

extern "C" size_t sc_graph_relabel_workspace_bytes(int64_t n) {
  size_t scan = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(n + 1));
  return align_up(scan, 256) + 512;
}

extern "C" int sc_graph_relabel(const int32_t* indptr, const int32_t* indices, const float* weights,
                                int64_t n, int k_fixed, const int32_t* order, const int32_t* rank,
                                int32_t* out_indptr, int32_t* out_indices, float* out_weights,
                                void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && order && rank && out_indices, "sc_graph_relabel: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_graph_relabel: need indptr or k_fixed");
  SC_CHECK_ARG(!indptr || out_indptr, "sc_graph_relabel: CSR input needs out_indptr");
  SC_CHECK_ARG((weights == nullptr) == (out_weights == nullptr), "sc_graph_relabel: weights and out_weights go together");
  int blocks = (int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256);
  if (indptr) {
    if (!ws || ws_bytes < sc_graph_relabel_workspace_bytes(n)) { set_error("sc_graph_relabel: workspace too small"); return SC_ERR_WORKSPACE; }
    gather_degrees_kernel<<<blocks, 256, 0, st>>>(indptr, order, n, out_indptr);
    SC_LAUNCH_OK();
    size_t bytes = ws_bytes;
    SC_CUDA_OK(cub::DeviceScan::ExclusiveSum(ws, bytes, out_indptr, out_indptr, (int)(n + 1), st));
  }
  int64_t want = (n + 7) / 8;
  int rb = (int)(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
  relabel_rows_kernel<<<rb, 256, 0, st>>>(indptr, indices, weights, n, k_fixed, order, rank,
                                          indptr ? out_indptr : nullptr, out_indices, out_weights);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_grid_radius_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return binning_bytes(n) + align_up(sizeof(unsigned long long), 256) + 1024;
}

// Cells per tile: the radius grid has cell edge ~r, i.e. an unknown occupancy; aim at ~256 queries
// per tile for a typical degree of 10-30 (pi r^2 density): 256 / (degree / pi) ~ 40 cells.
static int radius_tile_tx(int64_t, double) { return 40; }

static int radius_prepare(Arena& arena, int64_t n, Binning* b, unsigned long long** total) {
  if (!carve_binning(arena, n, b)) return SC_ERR_WORKSPACE;
  *total = arena.take<unsigned long long>(1);
  return *total ? SC_OK : SC_ERR_WORKSPACE;
}

extern "C" int sc_grid_radius_count(const double* coords, int64_t n, double r, int32_t* indptr,
                                    int64_t* nnz_out, const int32_t* labels, int n_types,
                                    float* profile, void* ws, size_t ws_bytes,
                                    sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(coords && ws, "sc_grid_radius_count: null coords/workspace");
  SC_CHECK_ARG(n >= 1 && n < (1ll << 31) - 2048, "sc_grid_radius_count: n out of range");
  SC_CHECK_ARG(r > 0, "radius must be > 0, got %g", r);
  SC_CHECK_ARG(indptr || profile, "sc_grid_radius_count: nothing to compute");
  SC_CHECK_ARG(!profile || (labels && n_types >= 1 && n_types <= 2048),
               "sc_grid_radius_count: fused composition needs labels and 1 <= n_types <= 2048");
  if (ws_bytes < sc_grid_radius_workspace_bytes(n)) {
    set_error("sc_grid_radius_count: workspace too small");
    return SC_ERR_WORKSPACE;
  }
  Arena arena(ws, ws_bytes);
  Binning b;
  unsigned long long* total;
  if (radius_prepare(arena, n, &b, &total)) { set_error("sc_grid_radius_count: carve failed"); return SC_ERR_WORKSPACE; }
  // cell edge slightly above r so that |dx| <= r implies a cell-index difference <= 1 under rounding
  int rc = run_binning(coords, n, 0.0, r * (1.0 + 1e-9), b, st);
  if (rc) return rc;
  const int threads = 256;
  int blocks = (int)((n + 255) / 256);
  const char* force = getenv("SC_RADIUS_VARIANT");  // "warp" forces the warp-per-query kernels
  if (!profile && !(force && !strcmp(force, "warp"))) {
    int* counters = reinterpret_cast<int*>(b.partial);  // [0] todo count, [1] work counter
    SC_CUDA_OK(cudaMemsetAsync(counters, 0, 2 * sizeof(int), st));
    const size_t tsm = (size_t)kTileCap * 20 + 64;
    int per_sm = 1;
    SC_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, radius_count_tile_kernel, kTileThreads, tsm));
    if (per_sm < 1) per_sm = 1;
    radius_count_tile_kernel<<<sm_count() * per_sm, kTileThreads, tsm, st>>>(
        b.gp, b.cell_start, b.xs, b.ys, b.order, r * r, radius_tile_tx(n, r), indptr, counters + 1);
    SC_LAUNCH_OK();
  } else {
    size_t smem = profile ? sizeof(int) * (threads / 32) * (size_t)n_types : 0;
    if (smem > 48 * 1024)
      SC_CUDA_OK(cudaFuncSetAttribute(radius_query_kernel<false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    radius_query_kernel<false><<<blocks, threads, smem, st>>>(
        b.gp, b.cell_start, b.xs, b.ys, b.order, n, r * r, indptr, nullptr, nullptr, nullptr, nullptr,
        nullptr, labels, n_types, profile, nullptr, nullptr);
    SC_LAUNCH_OK();
  }
  if (indptr) {
    SC_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(unsigned long long), st));
    sum_degrees_kernel<<<296, 256, 0, st>>>(indptr, n, total);
    SC_LAUNCH_OK();
    if (nnz_out)
      SC_CUDA_OK(cudaMemcpyAsync(nnz_out, total, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    SC_CUDA_OK(cudaMemsetAsync(indptr + n, 0, sizeof(int32_t), st));
    size_t bytes = b.cub_bytes;
    SC_CUDA_OK(cub::DeviceScan::ExclusiveSum(b.cub_tmp, bytes, indptr, indptr, (int)(n + 1), st));
  }
  return SC_OK;
}

extern "C" int sc_grid_radius_fill(const double* coords, int64_t n, double r,
                                   const int32_t* indptr, int32_t* indices, double* dist,
                                   void* scratch, size_t scratch_bytes, void* ws, size_t ws_bytes,
                                   sc_stream_t stream) {
  (void)coords;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indptr && indices && ws && scratch, "sc_grid_radius_fill: null argument");
  SC_CHECK_ARG(ws_bytes >= sc_grid_radius_workspace_bytes(n), "sc_grid_radius_fill: workspace too small");
  Arena arena(ws, ws_bytes);
  Binning b;
  unsigned long long* total;
  if (radius_prepare(arena, n, &b, &total)) { set_error("sc_grid_radius_fill: carve failed"); return SC_ERR_WORKSPACE; }
  // scratch = [tmp_idx i32[nnz]] [tmp_dist f64[nnz]]; the caller sized it from nnz
  size_t per = dist ? 12 : 4;
  size_t nnz_cap = scratch_bytes / per;
  int32_t* tmp_idx;
  double* tmp_dist = nullptr;
  if (dist) {
    tmp_dist = static_cast<double*>(scratch);
    tmp_idx = reinterpret_cast<int32_t*>(tmp_dist + nnz_cap);
  } else {
    tmp_idx = static_cast<int32_t*>(scratch);
  }
  const int threads = 256;
  int blocks = (int)((n + 255) / 256);
  radius_query_kernel<true><<<blocks, threads, 0, st>>>(b.gp, b.cell_start, b.xs, b.ys, b.order, n,
                                                       r * r, nullptr, indptr, tmp_idx, tmp_dist,
                                                       indices, dist, nullptr, 0, nullptr, nullptr, nullptr);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_cross_nn_workspace_bytes(int64_t n_targets) {
  if (n_targets <= 0) return 256;
  return binning_bytes(n_targets) + 1024;
}

extern "C" int sc_cross_nn(const double* targets, int64_t n_targets, const double* queries,
                           int64_t n_queries, int32_t* idx_out, double* dist_out, void* ws,
                           size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(targets && queries && idx_out && ws, "sc_cross_nn: null argument");
  SC_CHECK_ARG(n_targets >= 1 && n_targets < (1ll << 31) - 2048 && n_queries >= 1, "sc_cross_nn: sizes out of range");
  if (ws_bytes < sc_cross_nn_workspace_bytes(n_targets)) { set_error("sc_cross_nn: workspace too small"); return SC_ERR_WORKSPACE; }
  Arena arena(ws, ws_bytes);
  Binning b;
  if (!carve_binning(arena, n_targets, &b)) { set_error("sc_cross_nn: workspace carve failed"); return SC_ERR_WORKSPACE; }
  int rc = run_binning(targets, n_targets, 2.0, 0.0, b, st);
  if (rc) return rc;
  int64_t want = (n_queries + 255) / 256;
  int blocks = (int)(want > sm_count() * 16 ? sm_count() * 16 : want);
  cross_nn_kernel<<<blocks, 256, 0, st>>>(b.gp, b.cell_start, b.xs, b.ys, b.order, queries, n_queries, idx_out, dist_out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_pairwise_reduce_workspace_bytes(void) { return sizeof(double) * 2 * 148 * 8 + 256; }

extern "C" int sc_pairwise_reduce(const double* a, int64_t na, const double* b, int64_t nb, double* out,
                                  void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(a && b && out && ws, "sc_pairwise_reduce: null argument");
  SC_CHECK_ARG(na >= 1 && nb >= 1, "sc_pairwise_reduce: empty point set");
  if (ws_bytes < sc_pairwise_reduce_workspace_bytes()) { set_error("sc_pairwise_reduce: workspace too small"); return SC_ERR_WORKSPACE; }
  int64_t want = (na + 255) / 256;
  int blocks = (int)(want > 148 * 8 ? 148 * 8 : want);
  double* partial = static_cast<double*>(ws);
  pairwise_reduce_kernel<<<blocks, 256, 0, st>>>(a, na, b, nb, partial);
  SC_LAUNCH_OK();
  pairwise_final_kernel<<<1, 1, 0, st>>>(partial, blocks, out);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_nbhd_counts(const int32_t* indptr, const int32_t* indices, int64_t n,
                              int k_fixed, const int32_t* labels, int n_types, float* profile,
                              sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && labels && profile, "sc_nbhd_counts: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_nbhd_counts: need indptr or k_fixed");
  SC_CHECK_ARG(n_types >= 1 && n_types <= 2048, "sc_nbhd_counts: n_types outside [1,2048]");
  const int threads = 256;
  size_t smem = sizeof(int) * (threads / 32) * (size_t)n_types;
  if (smem > 48 * 1024)
    SC_CUDA_OK(cudaFuncSetAttribute(nbhd_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
  int64_t want = (n + 7) / 8;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
  nbhd_counts_kernel<<<blocks, threads, smem, st>>>(indptr, indices, n, k_fixed, labels, n_types,
                                                    profile);
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" int sc_profile_normalize(float* profile, int64_t n, int n_types, int normalize,
                                    int64_t* n_empty_out, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(profile && n_empty_out, "sc_profile_normalize: null argument");
  SC_CUDA_OK(cudaMemsetAsync(n_empty_out, 0, sizeof(int64_t), st));
  int64_t want = (n + 7) / 8;
  int blocks = (int)(want > 148 * 16 ? 148 * 16 : (want < 1 ? 1 : want));
  profile_normalize_kernel<<<blocks, 256, 0, st>>>(
      profile, n, n_types, normalize, reinterpret_cast<unsigned long long*>(n_empty_out));
  SC_LAUNCH_OK();
  return SC_OK;
}

extern "C" size_t sc_graph_moments_workspace_bytes(int64_t n) {
  return 2 * align_up(sizeof(double) * (size_t)(n > 0 ? n : 1), 256) +
         align_up(sizeof(double) * 3 * kMomentBlocks, 256) + 512;
}

extern "C" int sc_graph_moments(const int32_t* indptr, const int32_t* indices,
                                const float* weights, int64_t n, int k_fixed, double* out,
                                void* ws, size_t ws_bytes, sc_stream_t stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SC_CHECK_ARG(indices && out && ws, "sc_graph_moments: null argument");
  SC_CHECK_ARG(indptr || k_fixed > 0, "sc_graph_moments: need indptr or k_fixed");
  if (ws_bytes < sc_graph_moments_workspace_bytes(n)) { set_error("sc_graph_moments: workspace too small"); return SC_ERR_WORKSPACE; }
  Arena arena(ws, ws_bytes);
  double* rowsum = arena.take<double>(n);
  double* colsum = arena.take<double>(n);
  double* partial = arena.take<double>(3 * kMomentBlocks);
  SC_CUDA_OK(cudaMemsetAsync(colsum, 0, sizeof(double) * n, st));
  moments_edges_kernel<<<kMomentBlocks, 256, 0, st>>>(indptr, indices, weights, n, k_fixed, rowsum,
                                                     colsum, partial);
  SC_LAUNCH_OK();
  moments_s2_kernel<<<kMomentBlocks, 256, 0, st>>>(rowsum, colsum, n, partial + 2 * kMomentBlocks);
  SC_LAUNCH_OK();
  moments_final_kernel<<<1, 1, 0, st>>>(partial, partial + 2 * kMomentBlocks, kMomentBlocks, out);
  SC_LAUNCH_OK();
  return SC_OK;
}
