"""numpy/scipy restatement of the reference's spatial-statistics hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Each function cites the reference
lines it follows; paths are relative to /root/reference/src/spatialcore/.  Functions that
restate the squidpy/scanpy segment are tagged ``[unpinned]`` — they follow SURVEY.md
Appendix A (recalled upstream behaviour; those libraries are not in the tree).

Oracle policy: always evaluate in FP64.  The reference's own code follows
the dtype of ``X``; feeding it FP64 ``X`` evaluates the same algorithm without FP32 noise.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
from scipy import sparse
from scipy.spatial import cKDTree
from scipy.special import ndtr

# --------------------------------------------------------------------------------------
# neighbour graphs
# --------------------------------------------------------------------------------------


def sqdist(coords: np.ndarray, i, j) -> np.ndarray:
    """d² exactly as the CUDA kernel and the tree libraries evaluate it: FP64
    ``dx*dx + dy*dy`` with separate multiply and add (no FMA contraction)."""
    dx = coords[i, 0] - coords[j, 0]
    dy = coords[i, 1] - coords[j, 1]
    return dx * dx + dy * dy


def knn_canonical(coords: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Exact kNN, self excluded BY INDEX, ties broken by (d², index); each row returned
    sorted by column index (canonical CSR order, as scipy emits for
    ``build_spatial_weights`` [R spatial/autocorrelation.py:393-413]).

    Returns ``(idx int32 [N,k], dist float64 [N,k])``.  For unique coordinates this equals
    sklearn ``NearestNeighbors(k+1, 'ball_tree').kneighbors(coords)`` minus column 0
    [R autocorrelation.py:393-401], squidpy's ``kneighbors()`` [unpinned] and
    ``cKDTree.query(k+1)`` minus self [R spatial/neighborhoods.py:223-228].
    """
    coords = np.ascontiguousarray(coords[:, :2], dtype=np.float64)
    n = coords.shape[0]
    if k >= n:
        raise ValueError(f"k must be < number of cells ({n}), got {k}")
    kq = min(k + 2, n)
    tree = cKDTree(coords)
    _, nbr = tree.query(coords, k=kq)
    rows = np.arange(n)[:, None]
    d2 = sqdist(coords, rows, nbr)
    is_self = nbr == rows
    # order candidates by (d², idx) with self pushed to the end
    key_d = np.where(is_self, np.inf, d2)
    order = np.lexsort((nbr, key_d), axis=1)
    nbr_s = np.take_along_axis(nbr, order, axis=1)
    d2_s = np.take_along_axis(key_d, order, axis=1)
    out_idx = nbr_s[:, :k].copy()
    out_d2 = d2_s[:, :k].copy()
    # rows whose k-th and (k+1)-th candidates tie (or whose self was not returned because of
    # duplicates) are ambiguous for a tree: recompute them by brute force.
    ambiguous = np.zeros(n, dtype=bool)
    if kq > k:
        ambiguous |= d2_s[:, k - 1] == d2_s[:, k]
    ambiguous |= ~is_self.any(axis=1)
    for i in np.nonzero(ambiguous)[0]:
        dd = sqdist(coords, i, np.arange(n))
        dd[i] = np.inf
        o = np.lexsort((np.arange(n), dd))[:k]
        out_idx[i] = o
        out_d2[i] = dd[o]
    col = np.argsort(out_idx, axis=1, kind="stable")
    out_idx = np.take_along_axis(out_idx, col, axis=1).astype(np.int32)
    out_d2 = np.take_along_axis(out_d2, col, axis=1)
    return out_idx, np.sqrt(out_d2)


def knn_bruteforce(coords: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """O(N²) canonical kNN for small N (ties / duplicates / lattices)."""
    coords = np.ascontiguousarray(coords[:, :2], dtype=np.float64)
    n = coords.shape[0]
    ar = np.arange(n)
    d2 = sqdist(coords, ar[:, None], ar[None, :])
    d2[ar, ar] = np.inf
    idx = np.empty((n, k), dtype=np.int32)
    dist = np.empty((n, k), dtype=np.float64)
    for i in range(n):
        o = np.lexsort((ar, d2[i]))[:k]
        o = np.sort(o)
        idx[i] = o
        dist[i] = np.sqrt(d2[i, o])
    return idx, dist


def radius_graph(coords: np.ndarray, r: float) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Radius neighbours, inclusive (d² <= r²), self excluded by index, columns sorted.
    ``cKDTree.query_ball_point`` semantics [R spatial/neighborhoods.py:241-244] and sklearn
    ``radius_neighbors()`` as squidpy calls it [unpinned].  Returns CSR
    ``(indptr int32 [N+1], indices int32 [nnz], dist float64 [nnz])``."""
    coords = np.ascontiguousarray(coords[:, :2], dtype=np.float64)
    n = coords.shape[0]
    tree = cKDTree(coords)
    lists = tree.query_ball_point(coords, r=r, return_sorted=True)
    indptr = np.zeros(n + 1, dtype=np.int64)
    cols = []
    for i, l in enumerate(lists):
        a = np.asarray(l, dtype=np.int64)
        a = a[a != i]
        cols.append(a)
        indptr[i + 1] = indptr[i] + a.size
    indices = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    dist = np.sqrt(sqdist(coords, rows, indices)) if indices.size else np.zeros(0)
    return indptr.astype(np.int32), indices.astype(np.int32), dist


def build_spatial_weights(coords: np.ndarray, k: int, include_self: bool = False) -> sparse.csr_matrix:
    """Row-standardised kNN weights, FP32 data / int32 indices, columns sorted
    [R spatial/autocorrelation.py:342-413]."""
    n = coords.shape[0]
    idx, _ = knn_canonical(coords, k)
    if include_self:
        idx = np.sort(np.concatenate([np.arange(n, dtype=np.int32)[:, None], idx], axis=1), axis=1)
    kk = idx.shape[1]
    data = np.full(n * kk, np.float32(1.0) / np.float32(kk), dtype=np.float32)
    indptr = (np.arange(n + 1) * kk).astype(np.int32)
    return sparse.csr_matrix((data, idx.ravel(), indptr), shape=(n, n))


def spatial_neighbors(coords: np.ndarray, k: Optional[int] = None, radius: Optional[float] = None):
    """[unpinned] squidpy ``spatial_neighbors(coord_type='generic')`` as called at
    [R spatial/autocorrelation.py:565-570]: binary FP64 adjacency + FP64 distances, directed,
    diagonal unset."""
    n = coords.shape[0]
    if radius is None:
        idx, dist = knn_canonical(coords, k)
        indptr = (np.arange(n + 1) * k).astype(np.int32)
        indices = idx.ravel()
        dist = dist.ravel()
    else:
        indptr, indices, dist = radius_graph(coords, radius)
    adj = sparse.csr_matrix((np.ones(indices.size, dtype=np.float64), indices, indptr), shape=(n, n))
    dst = sparse.csr_matrix((dist.astype(np.float64), indices.copy(), indptr.copy()), shape=(n, n))
    return adj, dst


# --------------------------------------------------------------------------------------
# standardisation
# --------------------------------------------------------------------------------------


def zscore(X) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Per-gene mean / population std (ddof=0) / z = (x-mean)/std in FP64; zero-variance
    genes flagged and z set to 0 [R spatial/autocorrelation.py:820-830, 853-858, 902-906,
    1126-1143].  Returns ``(Z, mean, std, zero_var)``."""
    Xd = np.asarray(X.todense()) if sparse.issparse(X) else np.asarray(X)
    Xd = Xd.astype(np.float64)
    mean = Xd.mean(axis=0)
    std = Xd.std(axis=0)
    zero = std == 0
    safe = np.where(zero, 1.0, std)
    Z = (Xd - mean) / safe
    Z[:, zero] = 0.0
    return Z, mean, std, zero


# --------------------------------------------------------------------------------------
# global Moran's I (squidpy / scanpy segment)  [unpinned]
# --------------------------------------------------------------------------------------


def row_normalize(adj: sparse.csr_matrix) -> sparse.csr_matrix:
    """[unpinned] ``sklearn.preprocessing.normalize(g, 'l1', axis=1)``; empty rows stay 0."""
    g = adj.astype(np.float64).tocsr(copy=True)
    rs = np.asarray(np.abs(g).sum(axis=1)).ravel()
    rs[rs == 0] = 1.0
    g.data /= np.repeat(rs, np.diff(g.indptr))
    return g


def morans_i_stat(g: sparse.csr_matrix, X: np.ndarray) -> np.ndarray:
    """[unpinned] scanpy ``metrics.morans_i(g, vals)``: per gene ``z = x - mean``,
    ``I = N/S0 · Σ_i z_i (g z)_i / Σ_i z_i²`` in FP64 (Appendix A.2).  ``X`` is (N, G)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    z = X - X.mean(axis=0)
    lag = g @ z
    num = (z * lag).sum(axis=0)
    den = (z * z).sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        return n / g.data.sum() * num / den


def morans_i_perms_graph_rows(g: sparse.csr_matrix, X: np.ndarray, perms: np.ndarray) -> np.ndarray:
    """[unpinned] squidpy null: ``sims[p] = morans_i(g[idx_p, :], vals)`` — graph ROWS are
    permuted, columns are not (Appendix A.2).  Uses the exact identity
    ``Σ_i z_i·lag[idx[i]]``.  ``perms`` is (P, N) int."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    z = X - X.mean(axis=0)
    lag = g @ z
    den = (z * z).sum(axis=0)
    s0 = g.data.sum()
    sims = np.empty((perms.shape[0], X.shape[1]), dtype=np.float64)
    for p in range(perms.shape[0]):
        sims[p] = n / s0 * (z * lag[perms[p]]).sum(axis=0) / den
    return sims


def squidpy_perm_indices(n: int, n_perms: int, seed: int) -> np.ndarray:
    """[unpinned] permutation stream of ``spatial_autocorr(n_jobs=1, seed=s)``:
    ``default_rng(s + 0)``, one ``permutation(N)`` per permutation (Appendix A.2)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.permutation(n) for _ in range(n_perms)]).astype(np.int64)


def pval_sim_folded(score: np.ndarray, sims: np.ndarray) -> np.ndarray:
    """[unpinned] ``c = #{sims >= score}; c = min(c, P-c); p = (c+1)/(P+1)``."""
    P = sims.shape[0]
    c = (sims >= score[None, :]).sum(axis=0)
    c = np.where(P - c < c, P - c, c)
    return (c + 1) / (P + 1)


def graph_moments(g: sparse.csr_matrix) -> Tuple[float, float, float]:
    """[unpinned] ``s0 = Σg; t = g+gᵀ; s1 = Σ t∘t / 2; s2 = Σ_i (rowsum_i + colsum_i)²``."""
    g = g.astype(np.float64)
    s0 = float(g.sum())
    t = g + g.T
    s1 = float(t.multiply(t).sum() / 2.0)
    s2 = float((np.asarray(g.sum(axis=1)).ravel() + np.asarray(g.sum(axis=0)).ravel()) ** 2 @ np.ones(g.shape[0]))
    return s0, s1, s2


def var_norm(n: int, s0: float, s1: float, s2: float) -> float:
    """[unpinned] analytic variance of I under normality (Appendix A.2)."""
    return (n * n * s1 - n * s2 + 3.0 * s0 * s0) / ((n - 1.0) * (n + 1.0) * s0 * s0) - 1.0 / (n - 1.0) ** 2


def pval_norm(I: np.ndarray, n: int, vnorm: float) -> np.ndarray:
    z = (I - (-1.0 / (n - 1))) / np.sqrt(vnorm)
    return np.where(z > 0, 1.0 - ndtr(z), ndtr(z))


def pval_sim_two_tailed(score: np.ndarray, sims: np.ndarray) -> np.ndarray:
    """Two-tailed permutation p-value, the reference's own convention [R spatial/autocorrelation.py:330,
    :888-896]: ``(#{|sims| >= |score|} + 1)/(P + 1)`` (the ``two_tailed=True`` switch of the drop-in)."""
    P = sims.shape[0]
    return ((np.abs(sims) >= np.abs(score)[None, :]).sum(axis=0) + 1) / (P + 1)


def morans_i_perms_values(g: sparse.csr_matrix, X: np.ndarray, perms: np.ndarray) -> np.ndarray:
    """Global Moran's I under the VALUE-permuting null (the ``null_mode="values"`` switch of the drop-in):
    ``z_p = z[idx_p]; sims[p] = N/S0 · Σ_i z_p,i (g z_p)_i / Σ z²`` -- the permutation scheme of the reference's
    own ``local_morans_i`` [R spatial/autocorrelation.py:877-884] applied to the global statistic."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    z = X - X.mean(axis=0)
    den = (z * z).sum(axis=0)
    s0 = g.data.sum()
    sims = np.empty((perms.shape[0], X.shape[1]), dtype=np.float64)
    for p in range(perms.shape[0]):
        zp = z[perms[p]]
        sims[p] = n / s0 * (zp * (g @ zp)).sum(axis=0) / den
    return sims


def morans_i_table(coords, X, k=6, n_perms=10, seed=0, adj: Optional[sparse.csr_matrix] = None,
                   null_mode: str = "graph_rows", two_tailed: bool = False, transformation: bool = True):
    """End-to-end restatement of ``morans_i`` [R spatial/autocorrelation.py:421-648]:
    the squidpy segment [unpinned] + the reference's own result assembly (:589-616).
    Returns a dict of per-gene arrays in input gene order.

    The keyword defaults are the recalled squidpy behaviour (Appendix A); the alternatives restate the other
    reading of each convention so that the drop-in's switches have an oracle: ``null_mode="values"`` (values
    permuted instead of graph rows), ``two_tailed=True`` (|sims| >= |I| counts, normal p doubled),
    ``transformation=False`` (binary / stored weights instead of L1-row-normalised)."""
    n = X.shape[0]
    if adj is None:
        adj, _ = spatial_neighbors(coords, k=k)
    g = row_normalize(adj) if transformation else adj.astype(np.float64).tocsr()
    I = morans_i_stat(g, X)
    s0, s1, s2 = graph_moments(g)
    vn = var_norm(n, s0, s1, s2)
    expected = -1.0 / (n - 1)
    pn = pval_norm(I, n, vn)
    out = {"I": I, "expected_I": expected, "var_norm": vn, "pval_norm": pn * 2.0 if two_tailed else pn, "s": (s0, s1, s2)}
    out["z_score"] = (I - expected) / np.sqrt(vn) if vn > 0 else np.zeros_like(I)
    if n_perms and n_perms > 0:
        perms = squidpy_perm_indices(n, n_perms, seed)
        sims = morans_i_perms_graph_rows(g, X, perms) if null_mode == "graph_rows" else morans_i_perms_values(g, X, perms)
        out["sims"] = sims
        out["pval_sim"] = pval_sim_two_tailed(I, sims) if two_tailed else pval_sim_folded(I, sims)
        out["p_value"] = out["pval_sim"]
        out["count_ge"] = (sims >= I[None, :]).sum(axis=0)
    else:
        out["p_value"] = out["pval_norm"]
    return out


# --------------------------------------------------------------------------------------
# reference-own statistics (value-permuting null)
# --------------------------------------------------------------------------------------


def lees_l_pair(z_x, z_y, W, n_perms: int, rng: np.random.Generator):
    """[R spatial/autocorrelation.py:273-334]: ``L = Σ z_x·(W z_y)``; null permutes z_y only;
    two-tailed ``p = (#{|L_p| >= |L|}+1)/(P+1)``.  Returns ``(L_local, L, lag_y, p, L_perm)``."""
    lag = W @ z_y
    L_local = z_x * lag
    L = float(L_local.sum())
    p = 1.0
    Lp = np.zeros(n_perms)
    if n_perms > 0:
        for q in range(n_perms):
            zp = rng.permutation(z_y)
            Lp[q] = (z_x * (W @ zp)).sum()
        p = float(((np.abs(Lp) >= abs(L)).sum() + 1) / (n_perms + 1))
    return L_local, L, lag, p, Lp


def lees_l_all_pairs(Z: np.ndarray, W) -> np.ndarray:
    """All ordered pairs at once: ``L = Zᵀ (W Z)`` — the matrix whose (x,y) entry is
    [R spatial/autocorrelation.py:307-315] for the pair (x,y).  Not symmetric, not divided by N."""
    return Z.T @ (W @ Z)


def local_morans(Z: np.ndarray, W, perms: Optional[np.ndarray] = None):
    """[R spatial/autocorrelation.py:864-896]: ``lag = W Z``; ``I_loc = Z∘lag``; value-permuting
    null ``Zs = Z[perm]; I_p = Zs∘(W Zs)``; per-(cell,gene) two-tailed p."""
    lag = W @ Z
    I_loc = Z * lag
    if perms is None or len(perms) == 0:
        return I_loc, lag, np.ones_like(I_loc)
    cnt = np.zeros(I_loc.shape, dtype=np.int64)
    a = np.abs(I_loc)
    for pm in perms:
        Zs = Z[pm]
        cnt += np.abs(Zs * (W @ Zs)) >= a
    return I_loc, lag, (cnt + 1) / (len(perms) + 1)


def morans_values_null(Z: np.ndarray, W, perms: np.ndarray) -> np.ndarray:
    """Global statistic under the reference's own (value-permuting) null:
    ``sim_p,g = Σ_i Zs_i,g (W Zs)_i,g`` with ``Zs = Z[perm_p]`` [R autocorrelation.py:879-884]."""
    out = np.empty((len(perms), Z.shape[1]))
    for q, pm in enumerate(perms):
        Zs = Z[pm]
        out[q] = (Zs * (W @ Zs)).sum(axis=0)
    return out


def bh_adjust(p: np.ndarray) -> np.ndarray:
    """[R spatial/autocorrelation.py:132-164]."""
    n = len(p)
    if n == 0:
        return p.copy()
    o = np.argsort(p)
    adj = p[o] * n / np.arange(1, n + 1)
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    out = np.empty(n)
    out[o] = adj
    return np.clip(out, 0, 1)


def quadrants(z, lag, p=None, alpha=0.05) -> np.ndarray:
    """[R spatial/autocorrelation.py:219-265]: 0 NS, 1 HH, 2 LL, 3 HL, 4 LH."""
    q = np.zeros(z.shape, dtype=np.int8)
    q[(z > 0) & (lag > 0)] = 1
    q[(z < 0) & (lag < 0)] = 2
    q[(z > 0) & (lag < 0)] = 3
    q[(z < 0) & (lag > 0)] = 4
    if p is not None:
        q[p >= alpha] = 0
    return q


# --------------------------------------------------------------------------------------
# neighbourhood composition
# --------------------------------------------------------------------------------------


def neighborhood_profile(coords, labels: np.ndarray, n_types: int, k=None, radius=None, normalize=True):
    """[R spatial/neighborhoods.py:211-264]: per-cell histogram of neighbour cell types
    (kNN: k nearest excluding self by index; radius: d <= r excluding self), FP32, optional
    row normalisation.  ``labels`` are integer codes into ``sorted(unique)`` (:196-198)."""
    n = coords.shape[0]
    prof = np.zeros((n, n_types), dtype=np.float32)
    if radius is None:
        idx, _ = knn_canonical(coords, k)
        np.add.at(prof, (np.repeat(np.arange(n), k), labels[idx.ravel()]), 1.0)
    else:
        indptr, indices, _ = radius_graph(coords, radius)
        rows = np.repeat(np.arange(n), np.diff(indptr))
        np.add.at(prof, (rows, labels[indices]), 1.0)
    rs = prof.sum(axis=1)
    n_empty = int((rs == 0).sum())
    if n_empty:
        raise ValueError(f"{n_empty} cells have empty neighborhood profiles.")
    if normalize:
        prof = prof / rs[:, None]
    return prof


# --------------------------------------------------------------------------------------
# niches (k-means on the profile matrix)
# --------------------------------------------------------------------------------------


def identify_niches(profiles: np.ndarray, n_niches: int, random_state: int = 0, n_init: int = 10, max_iter: int = 300):
    """[R spatial/neighborhoods.py:441-466]: the reference's exact call,
    ``KMeans(n_clusters, init="k-means++", n_init, max_iter, random_state).fit_predict(profiles)``
    (scikit-learn is the third-party dependency behind this row and is installed on both the build
    container and the GPU box).  Returns ``(labels, centroids, inertia)``."""
    from sklearn.cluster import KMeans

    km = KMeans(n_clusters=n_niches, init="k-means++", n_init=n_init, max_iter=max_iter, random_state=random_state)
    labels = km.fit_predict(profiles)
    return labels.astype(np.int32), km.cluster_centers_, float(km.inertia_)


def kmeans_lloyd_from(profiles: np.ndarray, centers: np.ndarray, max_iter: int = 300):
    """Lloyd iterations from given centres (``KMeans(init=centers, n_init=1)``): isolates the
    iteration from the random seeding, so labels / centres are comparable exactly."""
    from sklearn.cluster import KMeans

    km = KMeans(n_clusters=centers.shape[0], init=np.asarray(centers), n_init=1, max_iter=max_iter)
    labels = km.fit_predict(profiles)
    return labels.astype(np.int32), km.cluster_centers_, float(km.inertia_), int(km.n_iter_)


def adjusted_rand_index(a: np.ndarray, b: np.ndarray) -> float:
    """Hubert-Arabie adjusted Rand index of two labelings (for clustering parity)."""
    a = np.asarray(a)
    b = np.asarray(b)
    _, ai = np.unique(a, return_inverse=True)
    _, bi = np.unique(b, return_inverse=True)
    cont = np.zeros((ai.max() + 1, bi.max() + 1), dtype=np.int64)
    np.add.at(cont, (ai, bi), 1)
    comb = lambda x: x * (x - 1) / 2.0  # noqa: E731
    s_ij = comb(cont).sum()
    s_a = comb(cont.sum(1)).sum()
    s_b = comb(cont.sum(0)).sum()
    tot = comb(a.size)
    expected = s_a * s_b / tot
    return float((s_ij - expected) / (0.5 * (s_a + s_b) - expected))


# --------------------------------------------------------------------------------------
# domain distances
# --------------------------------------------------------------------------------------


def domain_distances(coords, src_labels, tgt_labels, source_domains, target_domains, metric="minimum",
                     mode="both", same_column=False):
    """[R spatial/distance.py:196-400] restated with the same third-party calls (``cKDTree.query``,
    ``cdist``): returns ``(cell_dist float64[n] (NaN = not a source cell), cell_nearest object[n],
    matrix float64[len(source_domains), len(target_domains)] (NaN = not computed))``.
    ``src_labels`` / ``tgt_labels`` are object arrays with ``None`` for unlabelled cells."""
    from scipy.spatial import cKDTree
    from scipy.spatial.distance import cdist

    coords = np.asarray(coords, dtype=np.float64)[:, :2]
    n = coords.shape[0]
    src_labels = np.asarray(src_labels, dtype=object)
    tgt_labels = np.asarray(tgt_labels, dtype=object)
    cell_d = np.full(n, np.nan)
    cell_n = np.full(n, None, dtype=object)
    M = np.full((len(source_domains), len(target_domains)), np.nan)
    want_cell = mode in ("cell", "both")

    def nearest_cells():
        t_idx = np.where(np.isin(tgt_labels, target_domains))[0]
        s_idx = np.where(np.isin(src_labels, source_domains))[0]
        if len(t_idx) == 0 or len(s_idx) == 0:
            return None
        d, j = cKDTree(coords[t_idx]).query(coords[s_idx], k=1)
        cell_d[s_idx] = d
        cell_n[s_idx] = tgt_labels[t_idx][j]
        return s_idx, t_idx, d, tgt_labels[t_idx][j]

    if metric == "minimum" and want_cell:  # :214-271
        r = nearest_cells()
        if r is not None:
            s_idx, t_idx, d, near = r
            for a, src in enumerate(source_domains):
                m = src_labels[s_idx] == src
                if not m.any():
                    continue
                for b, tgt in enumerate(target_domains):
                    if src == tgt and same_column:
                        M[a, b] = 0.0
                    elif (near[m] == tgt).any():
                        M[a, b] = d[m][near[m] == tgt].min()
                    elif (tgt_labels[t_idx] == tgt).any():
                        M[a, b] = cdist(coords[s_idx][m], coords[t_idx][tgt_labels[t_idx] == tgt]).min()
    elif metric == "centroid":  # :273-327
        sc = {s: coords[src_labels == s].mean(0) for s in source_domains if (src_labels == s).any()}
        tc = {t: coords[tgt_labels == t].mean(0) for t in target_domains if (tgt_labels == t).any()}
        for a, src in enumerate(source_domains):
            for b, tgt in enumerate(target_domains):
                if src not in sc:
                    continue
                if src == tgt and same_column:
                    M[a, b] = 0.0
                elif tgt in tc:
                    M[a, b] = np.linalg.norm(sc[src] - tc[tgt])
        if want_cell:
            for i in np.where(np.isin(src_labels, source_domains))[0]:
                if src_labels[i] not in sc:
                    continue
                best, who = np.inf, None
                for tgt, c in tc.items():
                    if tgt == src_labels[i] and same_column:
                        continue
                    dd = np.linalg.norm(coords[i] - c)
                    if dd < best:
                        best, who = dd, tgt
                cell_d[i], cell_n[i] = best, who
    elif metric == "mean":  # :329-372
        for a, src in enumerate(source_domains):
            A = coords[src_labels == src]
            for b, tgt in enumerate(target_domains):
                B = coords[tgt_labels == tgt]
                if len(A) == 0:
                    continue
                if src == tgt and same_column:
                    M[a, b] = 0.0
                elif len(B):
                    M[a, b] = cdist(A, B).mean()
        if want_cell:
            nearest_cells()
    else:  # minimum, matrix only :374-398
        for a, src in enumerate(source_domains):
            A = coords[src_labels == src]
            for b, tgt in enumerate(target_domains):
                B = coords[tgt_labels == tgt]
                if len(A) == 0:
                    continue
                if src == tgt and same_column:
                    M[a, b] = 0.0
                elif len(B):
                    M[a, b] = cdist(A, B).min()
    return cell_d, cell_n, M
