/*
 * sc_b200.h — C ABI of libsc_b200.so: sm_100a CUDA kernels for the spatial-statistics hot path
 * of mcap91/SpatialCore (spatialcore.spatial).
 *
 * The reference has no FFI: its boundary is the Python function API
 * (src/spatialcore/spatial/__init__.py:11-52).  Each export below replaces the third-party
 * compiled routine the reference calls at the cited line; the Python layer (the modules under
 * spatialcore_b200/spatial/) keeps the reference's signatures and calls these through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless marked host;
 *   - the caller owns every buffer; the library allocates nothing and keeps no global state:
 *     scratch comes from a caller-provided workspace sized by the matching *_workspace_bytes();
 *   - all work is enqueued on `stream` (a cudaStream_t) and is stream-ordered and re-entrant;
 *   - return 0 on success, a negative sc_status otherwise; sc_last_error() gives a thread-local
 *     message.  No exception crosses the ABI.
 *   - matrices are row-major "cell-major": element (cell i, gene g) at base[i*ld + g].
 */
#ifndef SC_B200_H
#define SC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SC_API __attribute__((visibility("default")))
#else
#define SC_API
#endif

typedef void* sc_stream_t; /* cudaStream_t */

enum sc_status {
  SC_OK = 0,
  SC_ERR_INVALID = -1,     /* bad argument */
  SC_ERR_WORKSPACE = -2,   /* workspace too small */
  SC_ERR_CUDA = -3,        /* CUDA runtime error (message in sc_last_error) */
  SC_ERR_UNSUPPORTED = -4  /* valid request outside the compiled limits */
};

enum sc_dtype { SC_F32 = 0, SC_F64 = 1 };
enum sc_perm_source { SC_PERM_REPLAY = 0, SC_PERM_PHILOX = 1 };

#define SC_KNN_MAX_K 128

SC_API int sc_version(void);
SC_API const char* sc_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports the delta over the
 * timed region as gpu_launches; CUB sorts / scans called inside are not counted). */
SC_API long long sc_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Neighbour graphs.  Replace sklearn BallTree kneighbors (autocorrelation.py:393-395), squidpy's
 * kd_tree kneighbors()/radius_neighbors() (autocorrelation.py:565-570) and scipy cKDTree
 * query / query_ball_point (neighborhoods.py:213-241).
 *
 * Exact: neighbours ranked by FP64 d² = dx*dx + dy*dy (no FMA contraction), ties by index, self
 * excluded by index; each output row sorted by column index (canonical CSR order).
 * ------------------------------------------------------------------------------------------- */

SC_API size_t sc_grid_knn_workspace_bytes(int64_t n, int k);

/* coords f64[n,2]; idx i32[n, k+include_self]; dist f64[n, k+include_self] or NULL (aligned with
 * idx); order_out i32[n] or NULL (grid-sorted position -> original cell id, a spatial ordering).
 * labels/profile: optional fused neighbourhood composition (neighborhoods.py:226-237):
 * labels i32[n] in [0,n_types), profile f32[n,n_types] raw counts; pass NULL/0 to skip; idx may be
 * NULL when only the profile is wanted.  Requires 1 <= k <= SC_KNN_MAX_K and k < n. */
SC_API int sc_grid_knn(const double* coords, int64_t n, int k, int include_self, int32_t* idx,
                double* dist, int32_t* order_out, const int32_t* labels, int n_types,
                float* profile, void* ws, size_t ws_bytes, sc_stream_t stream);

SC_API size_t sc_grid_radius_workspace_bytes(int64_t n);

/* Pass 1: degrees + exclusive scan.  indptr i32[n+1]; nnz_out i64[1] (device).  Inclusive
 * d² <= r*r, self excluded.  The workspace keeps the binning for pass 2 and must not be touched
 * in between.  Optional fused composition as above (raw counts; empty rows stay zero). */
SC_API int sc_grid_radius_count(const double* coords, int64_t n, double r, int32_t* indptr,
                         int64_t* nnz_out, const int32_t* labels, int n_types, float* profile,
                         void* ws, size_t ws_bytes, sc_stream_t stream);

/* Pass 2: indices i32[nnz] column-sorted per row; dist f64[nnz] or NULL;
 * scratch: nnz*4 bytes (+ nnz*8 when dist != NULL). */
SC_API int sc_grid_radius_fill(const double* coords, int64_t n, double r, const int32_t* indptr,
                        int32_t* indices, double* dist, void* scratch, size_t scratch_bytes,
                        void* ws, size_t ws_bytes, sc_stream_t stream);

/* Spatial (grid) ordering of the cells: order_out i32[n] = original id of the cell at sorted position
 * a; rank_out i32[n] or NULL = its inverse.  The statistics are sums over cells, so the library may
 * hold Z / lag / the graph in this order: a row's neighbours are then close in memory. */
SC_API size_t sc_spatial_order_workspace_bytes(int64_t n);
SC_API int sc_spatial_order(const double* coords, int64_t n, int32_t* order_out, int32_t* rank_out,
                            void* ws, size_t ws_bytes, sc_stream_t stream);

/* Relabel a column-sorted CSR graph into that order: output row a = input row order[a], columns
 * mapped through rank and re-sorted.  out_indptr i32[n+1] (CSR input only), out_indices i32[nnz],
 * out_weights f32[nnz] iff weights != NULL. */
SC_API size_t sc_graph_relabel_workspace_bytes(int64_t n);
SC_API int sc_graph_relabel(const int32_t* indptr, const int32_t* indices, const float* weights,
                            int64_t n, int k_fixed, const int32_t* order, const int32_t* rank,
                            int32_t* out_indptr, int32_t* out_indices, float* out_weights, void* ws,
                            size_t ws_bytes, sc_stream_t stream);

/* Cross-set nearest neighbour: for every query point the nearest of `targets` (exact FP64 d2, ties by
 * target index).  Replaces cKDTree(target).query(source, k=1) in calculate_domain_distances
 * (distance.py:222-233, 359-367).  idx_out i32[n_queries] (index into targets), dist_out f64 or NULL. */
SC_API size_t sc_cross_nn_workspace_bytes(int64_t n_targets);
SC_API int sc_cross_nn(const double* targets, int64_t n_targets, const double* queries, int64_t n_queries,
                       int32_t* idx_out, double* dist_out, void* ws, size_t ws_bytes, sc_stream_t stream);

/* out f64[2] = { min over all pairs of |a_i - b_j|, sum over all pairs of |a_i - b_j| } (FP64, brute
 * force).  Replaces scipy cdist(a, b).min() / .mean() (distance.py:268-270, 341-343, 392-393). */
SC_API size_t sc_pairwise_reduce_workspace_bytes(void);
SC_API int sc_pairwise_reduce(const double* a, int64_t na, const double* b, int64_t nb, double* out,
                              void* ws, size_t ws_bytes, sc_stream_t stream);

/* Neighbourhood composition from an existing CSR graph (k_fixed>0 and indptr==NULL: every row has
 * k_fixed entries).  profile f32[n,n_types] raw counts. */
SC_API int sc_nbhd_counts(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                   const int32_t* labels, int n_types, float* profile, sc_stream_t stream);

/* Row-normalise counts to proportions in place; n_empty_out i64[1] = rows with zero sum
 * (neighborhoods.py:253-264). */
SC_API int sc_profile_normalize(float* profile, int64_t n, int n_types, int normalize,
                         int64_t* n_empty_out, sc_stream_t stream);

/* Analytic graph moments s0,s1,s2 of the row-standardised (weights==NULL: w_ij = 1/deg_i) or
 * explicitly weighted graph; rows must be column-sorted.  out f64[3].  Feeds var_norm / z_score
 * (autocorrelation.py:599-608; squidpy analytic moments). */
SC_API size_t sc_graph_moments_workspace_bytes(int64_t n);
SC_API int sc_graph_moments(const int32_t* indptr, const int32_t* indices, const float* weights,
                     int64_t n, int k_fixed, double* out, void* ws, size_t ws_bytes,
                     sc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Standardisation.  Replaces _compute_{mean,variance,std}_sparse and the z-scoring at
 * autocorrelation.py:66-124, 853-858, 1126-1143 (ddof = 0, FP64 accumulation).
 * ------------------------------------------------------------------------------------------- */

SC_API size_t sc_zscore_workspace_bytes(int64_t n, int g);

/* X: dense [n, ldx] of `dtype`.  cols i32[g] or NULL selects/reorders columns of X.
 * rows i32[n] or NULL: output row a is built from input row rows[a].
 * Z f32[n, ldz] (ldz % 4 == 0, ldz >= g; padding columns are zeroed).
 * mean, std f64[g]; zero_var u8[g].  Zero-variance genes get Z = 0. */
SC_API int sc_zscore(const void* X, int dtype, int64_t n, int64_t ldx, int g, const int32_t* cols,
              const int32_t* rows, float* Z, int64_t ldz, double* mean, double* std,
              uint8_t* zero_var, void* ws, size_t ws_bytes, sc_stream_t stream);

/* Second half of sc_zscore on its own: write Z from GIVEN per-gene mean / std / zero_var (device
 * arrays).  Lets several devices standardise row blocks of one matrix with moments combined across
 * them (row-sharded ingest; spatialcore_b200/distributed.py). */
SC_API int sc_zscore_apply(const void* X, int dtype, int64_t n, int64_t ldx, int g, const int32_t* cols,
                           const int32_t* rows, const double* mean, const double* std,
                           const uint8_t* zero_var, float* Z, int64_t ldz, sc_stream_t stream);

/* Fused standardise + all-gather + re-order over NVLink peer memory (row-sharded ingest, multi-GPU):
 * z-score the n rows of this GPU's block of X (f32, float4-readable rows) with the pooled moments and
 * store output row a at row dst_rows[a] of the Z matrix of EVERY peer.  peer_ptrs_host: HOST array of
 * n_peers (<= 16) device addresses, the peer-mapped bases of the symmetric Z buffers [*, ldz] (the local
 * buffer included).  The caller synchronises the peers afterwards (a barrier over the same group). */
SC_API int sc_zscore_scatter(const float* X, int64_t n, int64_t ldx, int g, const int32_t* dst_rows,
                             const double* mean, const double* std, const uint8_t* zero_var,
                             const uint64_t* peer_ptrs_host, int n_peers, int64_t ldz, sc_stream_t stream);

/* Scatter CSR expression (indptr i64[n+1], indices i32, data of `dtype`) into dense f32 [n, ldo]
 * (zero-filled first).  colmap i32[n_cols_x] or NULL: source column -> output column, -1 = drop. */
SC_API int sc_csr_densify(const int64_t* indptr, const int32_t* indices, const void* data, int dtype,
                   int64_t n, const int32_t* colmap, int g_out, float* out, int64_t ldo,
                   sc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Spatial lag + Moran numerator/denominator in one pass.  Replaces scanpy's numba Moran kernel
 * (via squidpy, autocorrelation.py:576-583) and `W @ Z` (autocorrelation.py:307, 864).
 * lag f32[n, ldl] or NULL; local f32[n, ldl] or NULL (= Z∘lag, autocorrelation.py:870);
 * num[g] = Σ_i z·lag, den[g] = Σ_i z² (FP64).
 * ------------------------------------------------------------------------------------------- */

SC_API size_t sc_csr_lag_moran_workspace_bytes(int64_t n, int g);
SC_API int sc_csr_lag_moran(const int32_t* indptr, const int32_t* indices, const float* weights,
                     int64_t n, int k_fixed, const float* Z, int64_t ldz, int g, float* lag,
                     float* local, int64_t ldl, double* num, double* den, void* ws,
                     size_t ws_bytes, sc_stream_t stream);

/* The fast path of the same step for row-standardised binary graphs held in spatial order
 * (sc_spatial_order + sc_graph_relabel): the lag through SHARED-MEMORY TILES.  The L1 load path delivers
 * ~46 B/clk/SM for 128-byte row pieces (measured), the shared-memory crossbar 126 B/clk/SM, so:
 * sc_graph_tile_build (once per graph) finds, for every chunk of 256 consecutive rows, the sorted union of
 * its neighbour columns and own rows (`urows`: at most 576 rows for mean degrees up to 22, else 1280) and
 * turns every CSR entry into a 16-bit word, the index of that row inside the tile (lists padded to a
 * multiple of four); sc_csr_lag_moran_tiled stages the urows' 128-byte pieces and the chunk's word lists
 * per column block with cp.async and walks the lists out of shared memory (FADD2 accumulation).  Chunks
 * whose union or word count exceeds the tile are computed by direct gathers inside the same call.
 * A row's neighbours are added in ascending column order, as in sc_csr_lag_moran: the two agree bit for bit.
 *
 * `tiles` is one caller-owned device buffer of sc_graph_tile_bytes(n, nnz) bytes (0 = invalid arguments);
 * pass the same n / nnz to all three calls.  nnz = n * k_fixed when indptr is NULL.
 * Zself f32[n, ldz] or NULL (NULL: the row's own value is the operand's); perm i32[n] or NULL: the operand
 * is Z[perm[j]] for row j -- the value-permuting null (autocorrelation.py:877-884) without materialising
 * the permuted matrix.  cell_obs / cell_cnt as in sc_perm_null_values (or NULL).
 * Workspace: sc_csr_lag_moran_workspace_bytes (partial sums + room for the chunk unions composed with `perm`). */
SC_API size_t sc_graph_tile_bytes(int64_t n, int64_t nnz);
SC_API int sc_graph_tile_build(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                               int64_t nnz, void* tiles, size_t tile_bytes, sc_stream_t stream);
SC_API int sc_csr_lag_moran_tiled(const int32_t* indptr, const int32_t* indices, int64_t n, int k_fixed,
                                  int64_t nnz, const void* tiles, size_t tile_bytes, const float* Zself,
                                  const float* Z, const int32_t* perm, int64_t ldz, int g, float* lag,
                                  float* local, int64_t ldl, double* num, double* den,
                                  const float* cell_obs, int32_t* cell_cnt, int64_t ldc, void* ws,
                                  size_t ws_bytes, sc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Permutation nulls.
 *
 * graph_rows (squidpy semantics, the null behind morans_i; autocorrelation.py:576-583):
 *     sims[p, c] = Σ_i A[i, c] · B[π_p(i), c]          (A = Z, B = cached lag)
 * values (the reference's own null; autocorrelation.py:322-328, 877-884):
 *     sims[p, c] = Σ_i S_p[i, c] · Σ_j w_ij · Zy[π_p(j), c]
 *     with S_p[i] = Zy[π_p(i)] when Zx == NULL (Moran) or Zx[i] (Lee: only y is permuted).
 *
 * Permutation p (0-based, global index perm_offset + p):
 *     SC_PERM_REPLAY : π_p(i) = perm_idx[p*n + i]  (host-generated, e.g. numpy's stream)
 *     SC_PERM_PHILOX : on-the-fly bijection keyed by Philox4x32-10(seed, perm_offset + p)
 * sims f64[n_perms, g] receives raw sums (no N/S0/den scaling).
 * ------------------------------------------------------------------------------------------- */

SC_API size_t sc_perm_null_workspace_bytes(int64_t n, int g);

SC_API int sc_perm_null_graph_rows(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n,
                            int g, int source, const int32_t* perm_idx, uint64_t seed,
                            int64_t perm_offset, int n_perms, double* sims, void* ws,
                            size_t ws_bytes, sc_stream_t stream);

/* cell_obs f32[n, ldc] / cell_cnt i32[n, ldc] or NULL: when given, cell_cnt[i,c] is incremented
 * for every permutation with |local_p[i,c]| >= |cell_obs[i,c]| (autocorrelation.py:888-896,
 * 1411-1413). */
SC_API int sc_perm_null_values(const int32_t* indptr, const int32_t* indices, const float* weights,
                        int64_t n, int k_fixed, const float* Zx, const float* Zy, int64_t ldz,
                        int g, int source, const int32_t* perm_idx, uint64_t seed,
                        int64_t perm_offset, int n_perms, double* sims, const float* cell_obs,
                        int32_t* cell_cnt, int64_t ldc, void* ws, size_t ws_bytes,
                        sc_stream_t stream);

/* Workspace for sc_perm_null_values: sc_perm_null_workspace_bytes plus, for wide matrices, room for
 * one permuted copy of Zy (the materialised variant; a smaller workspace selects the register-gather
 * kernel). */
SC_API size_t sc_perm_null_values_workspace_bytes(int64_t n, int g);

/* dst[a, 0:cols) = src[rows[a], 0:cols): put per-cell matrices into / out of the spatial order.
 * cols, lds, ldd multiples of 4. */
SC_API int sc_gather_rows(const float* src, int64_t lds, int64_t n, int64_t cols, const int32_t* rows,
                          float* dst, int64_t ldd, sc_stream_t stream);

/* Re-express replayed permutations of cell ids on sorted positions:
 * out[p, a] = rank[perm_idx[p, order[a]]], so that Σ_a A_s[a]·B_s[out[p,a]] = Σ_i A[i]·B[perm_idx[p,i]]
 * for A_s[a] = A[order[a]].  perm_idx, out i32[n_perms, n] (must not alias). */
SC_API int sc_perm_conjugate(const int32_t* perm_idx, int64_t n, int n_perms, const int32_t* order,
                             const int32_t* rank, int32_t* out, sc_stream_t stream);

/* Local Moran epilogue (autocorrelation.py:888-928 with 132-183 and 219-265): from the per-cell
 * exceedance counts cnt i32[n, ldc] of n_perms permutations, the standardised values Z, their lag and
 * local I (all f32[n, ldz], stored in the order given by `order`, or NULL = user order) produce, at each
 * cell's ORIGINAL row and tightly packed [n, g]: z, lag, local I, p = (cnt+1)/(n_perms+1),
 * the adjusted p (method 0 none, 1 bonferroni, 2 Benjamini-Hochberg per gene over the n cells, computed
 * from a (n_perms+1)-bin histogram, no sort) and the LISA quadrant i8 (0 NS, 1 HH, 2 LL, 3 HL, 4 LH;
 * set to 0 where adjusted p >= alpha when n_perms > 0).  zero_var u8[g] or NULL: those genes get
 * z = lag = I = 0, p = 1. */
SC_API size_t sc_local_moran_finish_workspace_bytes(int g, int n_perms);
SC_API int sc_local_moran_finish(const int32_t* cnt, int64_t ldc, const float* Z, const float* lag,
                                 const float* local, int64_t ldz, const int32_t* order, int64_t n, int g,
                                 int n_perms, const uint8_t* zero_var, int method, float alpha,
                                 float* z_out, float* lag_out, float* local_out, float* p_out,
                                 float* padj_out, int8_t* quad_out, void* ws, size_t ws_bytes,
                                 sc_stream_t stream);

/* Materialise Philox permutation `perm_index` of [0,n) into out i32[n] (tests, replay export). */
SC_API int sc_philox_permutation(uint64_t seed, int64_t perm_index, int64_t n, int32_t* out,
                          sc_stream_t stream);

/* Same permutation evaluated on the HOST into host memory (no GPU needed): lets a Philox-mode run be
 * replayed through any CPU implementation. */
SC_API int sc_philox_permutation_host(uint64_t seed, int64_t perm_index, int64_t n, int32_t* out_host);

/* Fold a batch of simulated statistics into running null summaries:
 * s = sims[p,c]*scale[c]; cnt_ge += (s >= obs[c]); cnt_abs_ge += (|s| >= |obs[c]|);
 * sum += s; sumsq += s*s.  (squidpy pval_sim / var_sim; autocorrelation.py:331-332.) */
SC_API int sc_null_accumulate(const double* sims, int n_perms, int g, const double* scale,
                       const double* obs, int64_t* cnt_ge, int64_t* cnt_abs_ge, double* sum,
                       double* sumsq, sc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Lee's L for all ordered gene pairs: L[x, y] = Σ_i A[i, x] · B[i, y]  (A = Z, B = W Z), i.e. the
 * matrix of autocorrelation.py:307-315 over every pair.
 * impl 1: CUDA-core kernel, FP64 accumulation (exact up to input rounding).
 * impl 2: tensor cores -- tcgen05.mma kind::tf32 with 3xTF32 operand splitting (fused in-kernel), FP32
 *         accumulation in TMEM per 256-cell chunk, chunks folded in FP32 registers, CTA partials summed
 *         in FP64 (~2e-6 relative).
 * impl 0 (default): impl 2 when n >= 8192 and g >= 64, else impl 1.
 * L f32[g, ldl].
 * ------------------------------------------------------------------------------------------- */
SC_API size_t sc_lee_gemm_workspace_bytes(int64_t n, int g);

/* Two-tailed exceedance counts for an all-pairs permutation null: cnt[x,y] += |Lp[x,y]| >= |L[x,y]|
 * (autocorrelation.py:331-332 for every pair at once).  Lp, L f32[g, ld*]; cnt i32[g, ldc]. */
SC_API int sc_lee_abs_ge_accumulate(const float* Lp, int64_t ldp, const float* L, int64_t ldl, int g,
                                    int32_t* cnt, int64_t ldc, sc_stream_t stream);
SC_API int sc_lee_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t n, int g,
                float* L, int64_t ldl, int impl, void* ws, size_t ws_bytes, sc_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * K-means on the neighbourhood-profile matrix (identify_niches, neighborhoods.py:299-522; replaces
 * sklearn.cluster.KMeans: k-means++ seeding + Lloyd iterations).  X f32[n, ldx], d <= 128 features,
 * k <= 64 centres, k*d <= 2048.
 *
 * sc_kmeans_assign: one Lloyd pass.  labels i32[n] IN/OUT (the previous labels are compared to count
 * changes; fill with -1 before the first pass); mind f32[n] or NULL = squared distance to the chosen
 * centre; out f64[k*d + k + 2] = per-centre feature sums [k][d], member counts [k], inertia
 * (sum of min squared distances), number of rows whose label changed.  Nearest centre by FP32
 * squared differences, first minimum wins (numpy argmin).
 * ------------------------------------------------------------------------------------------- */
SC_API size_t sc_kmeans_workspace_bytes(int64_t n, int d, int k);
SC_API int sc_kmeans_assign(const float* X, int64_t n, int64_t ldx, int d, const float* centers, int k,
                            int32_t* labels, float* mind, double* out, void* ws, size_t ws_bytes,
                            sc_stream_t stream);

/* k-means++ seeding.  cand i32[n_cand] (device) are row indices of candidate centres;
 * pot_out f64[n_cand] = sum_i min(mind_in[i], ||x_i - x_cand||^2) (mind_in NULL: plain sum).
 * commit >= 0: also store min(mind_in, distance to candidate `commit`) into mind_out f32[n]. */
SC_API int sc_kmeans_pp_potential(const float* X, int64_t n, int64_t ldx, int d, const int32_t* cand,
                                  int n_cand, const float* mind_in, float* mind_out, int commit,
                                  double* pot_out, void* ws, size_t ws_bytes, sc_stream_t stream);

/* D^2 sampling: idx_out[l] = searchsorted(cumsum_f64(mind), vals[l], side="left") clipped to n-1
 * (vals f64[n_vals] on the device, n_vals <= 16). */
SC_API size_t sc_kmeans_pp_sample_workspace_bytes(int64_t n);
SC_API int sc_kmeans_pp_sample(const float* mind, int64_t n, const double* vals, int n_vals,
                               int32_t* idx_out, void* ws, size_t ws_bytes, sc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H */
