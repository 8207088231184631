/*
 * TEST INFRASTRUCTURE ONLY — CPU port of the Moran's I permutation loop, used as the timed CPU
 * baseline (bench.py cpu_baseline / --impl reference) and cross-checked against oracle/restate.py.
 *
 * What it ports [parity unpinned: squidpy/scanpy are not in /root/reference; see oracle/__init__.py]:
 *   squidpy.gr.spatial_autocorr(mode="moran") as called at
 *   /root/reference/src/spatialcore/spatial/autocorrelation.py:576-583 —
 *     for each permutation: g_p = g[idx, :]            (scipy CSR row fancy-index, single thread)
 *                           sims[p] = morans_i(g_p, vals)   (scanpy numba kernel, prange over genes)
 *   with morans_i(g, x) = N/S0 * sum_i z_i * sum_{j in row i} g_ij z_j / sum_i z_i^2, z = x - mean,
 *   FP64.  vals is gene-major (G, N) as squidpy densifies it.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the CPU arm asks for all host cores explicitly */
void moran_port_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int moran_port_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* one statistic per gene; prange over genes like scanpy's kernel */
static void moran_all_genes(const int32_t* indptr, const int32_t* indices, const double* data,
                            const double* z /*[G][N] centred*/, const double* den, int64_t n,
                            int64_t g, double s0, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < g; ++c) {
    const double* zc = z + c * n;
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double lag = 0.0;
      for (int32_t t = indptr[i]; t < indptr[i + 1]; ++t) lag += data[t] * zc[indices[t]];
      acc += zc[i] * lag;
    }
    out[c] = (double)n / s0 * acc / den[c];
  }
}

/*
 * vals: (G, N) float64 gene-major.  perms: (P, N) int32, or NULL for the observed statistic only.
 * score: (G) out.  sims: (P, G) out.  Returns 0.
 */
int moran_port_run(const int32_t* indptr, const int32_t* indices, const double* data, int64_t n,
                   const double* vals, int64_t g, const int32_t* perms, int64_t n_perms,
                   double* score, double* sims) {
  double* z = (double*)malloc(sizeof(double) * (size_t)n * (size_t)g);
  double* den = (double*)malloc(sizeof(double) * (size_t)g);
  if (!z || !den) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < g; ++c) {
    double m = 0.0;
    for (int64_t i = 0; i < n; ++i) m += vals[c * n + i];
    m /= (double)n;
    double d = 0.0;
    for (int64_t i = 0; i < n; ++i) {
      double v = vals[c * n + i] - m;
      z[c * n + i] = v;
      d += v * v;
    }
    den[c] = d;
  }
  const int64_t nnz = indptr[n];
  double s0 = 0.0;
  for (int64_t t = 0; t < nnz; ++t) s0 += data[t];
  moran_all_genes(indptr, indices, data, z, den, n, g, s0, score);

  if (perms && n_perms > 0) {
    int32_t* p_indptr = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n + 1));
    int32_t* p_indices = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    double* p_data = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    if (!p_indptr || !p_indices || !p_data) return -1;
    for (int64_t p = 0; p < n_perms; ++p) {
      const int32_t* idx = perms + p * n;
      /* g[idx, :] — row gather, serial like scipy's csr_row_index */
      p_indptr[0] = 0;
      for (int64_t i = 0; i < n; ++i) {
        int32_t r = idx[i];
        int32_t len = indptr[r + 1] - indptr[r];
        memcpy(p_indices + p_indptr[i], indices + indptr[r], sizeof(int32_t) * (size_t)len);
        memcpy(p_data + p_indptr[i], data + indptr[r], sizeof(double) * (size_t)len);
        p_indptr[i + 1] = p_indptr[i] + len;
      }
      moran_all_genes(p_indptr, p_indices, p_data, z, den, n, g, s0, sims + p * g);
    }
    free(p_indptr); free(p_indices); free(p_data);
  }
  free(z); free(den);
  return 0;
}
