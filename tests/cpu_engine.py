"""numpy / torch-CPU stand-in for ``spatialcore_b200.engine`` (TEST INFRASTRUCTURE ONLY).

Implements, with the oracle's arithmetic, the subset of the engine that the ``spatial`` entry points
(``morans_i``, ``local_morans_i``, ``lees_l*``, ``build_spatial_weights``, ``compute_neighborhood_profile``)
drive, so that the drop-in's HOST logic (validation, gene resolution, sharding arithmetic, spatial re-ordering and
permutation conjugation, p-value folding, z-scores, table / metadata / obsp assembly) can be exercised
without a GPU.  The CUDA path is compared with the oracle in tests/test_gpu_parity.py; nothing here is
imported by the product."""

from __future__ import annotations

import numpy as np
import torch
from scipy import sparse

from oracle import restate as R
from spatialcore_b200 import engine as real
from spatialcore_b200 import philox

DeviceGraph = real.DeviceGraph
CellOrder = real.CellOrder
Standardized = real.Standardized
padded_ld = real.padded_ld


def _csr(indptr, indices, n):
    return DeviceGraph(n=n, indices=torch.from_numpy(indices.astype(np.int32)), indptr=torch.from_numpy(indptr.astype(np.int32)))


def _composition(indptr, indices, labels, n_types):
    """Per-row label histogram of a CSR graph, FP32 [R neighborhoods.py:226-233, 246-250]."""
    n = len(indptr) - 1
    lab = np.asarray(labels.numpy() if hasattr(labels, "numpy") else labels, dtype=np.int64)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    prof = np.zeros((n, n_types), dtype=np.float32)
    np.add.at(prof, (rows, lab[indices]), 1.0)
    return torch.from_numpy(prof)


def knn_graph(coords, k, include_self=False, want_dist=False, want_order=False, labels=None, n_types=0, want_idx=True,
              device="cpu"):
    c = np.asarray(coords, dtype=np.float64)[:, :2]
    if k < 1:
        raise ValueError(f"n_neighbors must be >= 1, got {k}")
    if k >= c.shape[0]:
        raise ValueError(f"k must be < number of cells ({c.shape[0]}), got {k}")
    n = c.shape[0]
    if include_self:  # [R autocorrelation.py:398-401]: the k+1 nearest, the cell itself among them
        idx, dist = R.knn_canonical(c, k)
        idx = np.sort(np.concatenate([np.arange(n)[:, None], idx], axis=1), axis=1)
        g = DeviceGraph(n=n, indices=torch.from_numpy(idx.astype(np.int32)), k_fixed=k + 1)
        return g, None, None
    idx, dist = R.knn_canonical(c, k)
    g = DeviceGraph(n=n, indices=torch.from_numpy(idx.astype(np.int32)) if want_idx else None, k_fixed=k,
                    dist=torch.from_numpy(dist) if want_dist else None)
    prof = None
    if labels is not None:
        prof = _composition(np.arange(0, n * k + 1, k), idx.reshape(-1), labels, n_types)
    return g, None, prof


def radius_graph(coords, radius, want_dist=False, labels=None, n_types=0, want_graph=True, device="cpu"):
    c = np.asarray(coords, dtype=np.float64)[:, :2]
    indptr, indices, dist = R.radius_graph(c, radius)
    g = _csr(indptr, indices, c.shape[0])
    g.dist = torch.from_numpy(dist) if want_dist else None
    prof = _composition(indptr, indices, labels, n_types) if labels is not None else None
    return (g if want_graph else None), prof


def nbhd_counts(graph, labels, n_types):
    return _composition(graph.indptr_tensor().numpy(), graph.indices.reshape(-1).numpy(), labels, n_types)


def profile_normalize(profile, normalize):
    """In place; number of all-zero rows [R neighborhoods.py:253-264]."""
    sums = profile.sum(1, keepdim=True)
    n_empty = int((sums == 0).sum())
    if normalize and n_empty == 0:
        profile /= sums
    return n_empty


def graph_from_scipy(adj, device="cpu", use_weights=False):
    a = sparse.csr_matrix(adj)
    a.sort_indices()
    g = _csr(a.indptr, a.indices, a.shape[0])
    if use_weights:
        g.weights = torch.from_numpy(a.data.astype(np.float32))
    return g


def spatial_order(coords, device="cpu"):
    c = np.asarray(coords, dtype=np.float64)[:, :2]
    order = np.lexsort((c[:, 0], c[:, 1])).astype(np.int32)  # any bijection exercises the conjugation logic
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order), dtype=np.int32)
    return CellOrder(order=torch.from_numpy(order), rank=torch.from_numpy(rank))


def _to_scipy(graph):
    indptr = graph.indptr_tensor().numpy()
    idx = graph.indices.reshape(-1).numpy()
    if graph.weights is not None:
        vals = graph.weights.numpy().astype(np.float64)
    else:
        deg = np.diff(indptr)
        vals = np.repeat(1.0 / np.maximum(deg, 1), deg)
    return sparse.csr_matrix((vals, idx, indptr), shape=(graph.n, graph.n))


def relabel_graph(graph, co, tiles=True):
    o = co.order.numpy()
    A = _to_scipy(graph)[o][:, o].tocsr()
    A.sort_indices()
    out = _csr(A.indptr, A.indices, graph.n)
    if graph.weights is not None:
        out.weights = torch.from_numpy(A.data.astype(np.float32))
    return out


def expression_to_device(X, gene_idx, device="cpu"):
    Xd = np.asarray(X.todense()) if sparse.issparse(X) else np.asarray(X)
    if gene_idx is not None:
        Xd = Xd[:, np.asarray(gene_idx)]
    return torch.from_numpy(np.ascontiguousarray(Xd)), None


def zscore_dense(X, cols=None, rows=None, want_z=True):
    x = X.numpy()
    Z, mean, std, zero = R.zscore(x if cols is None else x[:, cols.numpy()])
    if rows is not None:
        Z = Z[rows.numpy()]
    n, g = Z.shape
    out = np.zeros((n, padded_ld(g)), dtype=np.float32)
    out[:, :g] = Z
    return Standardized(Z=torch.from_numpy(out) if want_z else None, g=g, mean=torch.from_numpy(mean),
                        std=torch.from_numpy(std), zero_var=torch.from_numpy(zero.astype(np.uint8)))


def zscore_apply(X, mean, std, zero_var, cols=None, rows=None, out=None):
    """z-score with GIVEN moments (row-sharded ingest) in the kernel's arithmetic: FP64, stored FP32."""
    x = X.numpy().astype(np.float64)
    x = x if cols is None else x[:, cols.numpy()]
    x = x if rows is None else x[rows.numpy()]
    zero = zero_var.numpy().astype(bool)
    z = np.where(zero, 0.0, (x - mean.numpy()) / np.where(zero, 1.0, std.numpy())).astype(np.float32)
    n, g = z.shape
    if out is None:
        out = torch.zeros((n, padded_ld(g)), dtype=torch.float32)
    out.zero_()
    out[:, :g] = torch.from_numpy(z)
    return out


def lag_moran(graph, Z, g, want_lag=True, want_local=False):
    W = _to_scipy(graph)
    z = Z.numpy().astype(np.float64)
    lag = (W @ z).astype(np.float32)
    num = (z[:, :g] * lag[:, :g].astype(np.float64)).sum(0)
    den = (z[:, :g] ** 2).sum(0)
    return torch.from_numpy(num), torch.from_numpy(den), torch.from_numpy(lag) if want_lag else None, None


def graph_moments(graph):
    return R.graph_moments(_to_scipy(graph))


def conjugate_perms(perm_idx, co):
    o, r = co.order.numpy(), co.rank.numpy()
    return torch.from_numpy(r[perm_idx.numpy()[:, o]].astype(np.int32))


def gather_rows(src, rows):
    return src[rows.long()]


def philox_permutation(seed, perm_index, n, device="cpu"):
    return torch.from_numpy(philox.permutation(seed, perm_index, n).astype(np.int32))


def lee_gemm(A, B, g, impl=0):
    """``L[x, y] = sum_i A[i, x] B[i, y]`` in FP64, stored FP32 like the device kernels."""
    return torch.from_numpy((A.numpy()[:, :g].astype(np.float64).T @ B.numpy()[:, :g].astype(np.float64)).astype(np.float32))


def lee_abs_ge_accumulate(Lp, L_obs, cnt):
    cnt += (Lp.abs() >= L_obs.abs()).to(cnt.dtype)


def perm_null_graph_rows(A, B, g, n_perms, perm_idx=None, seed=0, perm_offset=0, out=None, ws=None):
    a = A.numpy().astype(np.float64)[:, :g]
    b = B.numpy().astype(np.float64)[:, :g]
    n = a.shape[0]
    sims = np.empty((n_perms, g))
    for p in range(n_perms):
        pi = perm_idx[p].numpy() if perm_idx is not None else philox.permutation(seed, perm_offset + p, n)
        sims[p] = (a * b[pi]).sum(0)
    return torch.from_numpy(sims)


def null_accumulate(sims, scale, obs, cnt_ge, cnt_abs_ge, ssum, ssq):
    s = sims * scale if scale is not None else sims
    cnt_ge += (s >= obs).sum(0)
    cnt_abs_ge += (s.abs() >= obs.abs()).sum(0)
    ssum += s.sum(0)
    ssq += (s * s).sum(0)


def perm_null_values(graph, Zy, g, n_perms, Zx=None, perm_idx=None, seed=0, perm_offset=0, cell_obs=None, cell_cnt=None):
    """Value-permuting null [R autocorrelation.py:322-328, 877-896] in FP32 like the kernel: lag of the
    permuted matrix, local statistic, optional per-cell exceedance counts."""
    W = _to_scipy(graph).astype(np.float32)
    zy = Zy.numpy()
    n = zy.shape[0]
    sims = np.empty((n_perms, g))
    for p in range(n_perms):
        pi = perm_idx[p].numpy() if perm_idx is not None else philox.permutation(seed, perm_offset + p, n)
        zp = zy[pi]
        lag = (W @ zp).astype(np.float32)
        s = Zx.numpy() if Zx is not None else zp
        loc = s * lag
        sims[p] = loc[:, :g].astype(np.float64).sum(0)
        if cell_cnt is not None:
            cell_cnt += torch.from_numpy((np.abs(loc) >= np.abs(cell_obs.numpy())).astype(np.int32))
    return torch.from_numpy(sims)


def lag_moran(graph, Z, g, want_lag=True, want_local=False, perm=None):  # noqa: F811  (adds the local statistic)
    W = _to_scipy(graph).astype(np.float32)
    z32 = Z.numpy() if perm is None else Z.numpy()[perm.numpy()]
    lag = (W @ z32).astype(np.float32)
    z = z32.astype(np.float64)
    num = (z[:, :g] * lag[:, :g].astype(np.float64)).sum(0)
    den = (z[:, :g] ** 2).sum(0)
    return (torch.from_numpy(num), torch.from_numpy(den), torch.from_numpy(lag) if want_lag else None,
            torch.from_numpy(z32 * lag) if want_local else None)


def local_moran_finish(cnt, Z, lag, loc, g, n_perms, zero_var, method, alpha, order=None):
    """Per-cell p, adjusted p and quadrants with the reference's numpy recipe, un-sorted to user order."""
    from spatialcore_b200.spatial import autocorrelation as ac

    n = Z.shape[0]
    inv = np.arange(n) if order is None else np.argsort(order.numpy())  # original row i is stored at inv[i]
    zr, lr, ir = (t.numpy()[inv][:, :g].copy() for t in (Z, lag, loc))
    dead = zero_var.numpy().astype(bool) if zero_var is not None else np.zeros(g, bool)
    zr[:, dead] = 0; lr[:, dead] = 0; ir[:, dead] = 0
    p = np.ones((n, g), dtype=np.float32)
    pa = np.ones((n, g), dtype=np.float32)
    if n_perms > 0:
        p = ((cnt.numpy()[inv][:, :g] + 1) / (n_perms + 1)).astype(np.float32)
        p[:, dead] = 1.0
        for j in range(g):
            pa[:, j] = ac._fdr(p[:, j], method)
        q = ac._classify_quadrants(zr, lr, pa, alpha)
    else:
        q = ac._classify_quadrants(zr, lr, None, alpha)
    return tuple(torch.from_numpy(np.ascontiguousarray(x)) for x in (zr, lr, ir, p, pa, q))
