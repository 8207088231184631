"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle and the reference goldens.

Bars: bit-exact for indices / counts / p-values under replayed permutations; FP32 statistics within
``|a-b| <= 1e-5*|b| + floor`` of the FP64 oracle (floor stated per test)."""

import os

import numpy as np
import pandas as pd
import pytest
import torch
from scipy import sparse

from oracle import restate as R
from tests.golden import inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spatialcore_b200 import engine

    return engine


@pytest.fixture(scope="module")
def api(eng):
    from spatialcore_b200 import spatial

    return spatial


@pytest.fixture(scope="module")
def g0(golden_dir):
    coords, X = inputs.g0()
    return coords, X, np.load(os.path.join(golden_dir, "ref_g0.npz"))


def _adata(X, coords, **kw):
    from spatialcore_b200 import AnnDataLite

    return AnnDataLite(X, obsm={"spatial": coords}, **kw)


# ---------------------------------------------------------------------------------------------
# graphs
# ---------------------------------------------------------------------------------------------


def test_knn_matches_reference_goldens(eng, g0, golden_dir):
    coords, _, ref = g0
    for k in (6, 15):
        graph, _, _ = eng.knn_graph(coords, k, want_dist=True)
        assert np.array_equal(graph.indices.cpu().numpy().ravel(), ref[f"W_k{k}_indices"])
        _, dist = R.knn_canonical(coords, k)
        assert np.array_equal(graph.dist.cpu().numpy(), dist)  # FP64 distances bit-identical
    graph, _, _ = eng.knn_graph(coords, 6, include_self=True)
    assert np.array_equal(graph.indices.cpu().numpy().ravel(), ref["W_k6_self_indices"])
    refc = np.load(os.path.join(golden_dir, "ref_knn_clustered.npz"))
    cc = inputs.clustered(4000, seed=11)
    for k in (6, 15, 30, 50):
        graph, _, _ = eng.knn_graph(cc, k)
        assert np.array_equal(graph.indices.cpu().numpy().ravel(), refc[f"W_k{k}_indices"]), k


@pytest.mark.parametrize("k", [1, 4, 9, 33, 64, 100])
def test_knn_ties_duplicates_and_wide_k(eng, k):
    for coords in (inputs.lattice(23, 17), inputs.with_duplicates(700, 3), inputs.clustered(1500, 4)):
        if k >= coords.shape[0]:
            continue
        graph, _, _ = eng.knn_graph(coords, k, want_dist=True)
        idx, dist = R.knn_bruteforce(coords, k)
        assert np.array_equal(graph.indices.cpu().numpy(), idx)
        assert np.array_equal(graph.dist.cpu().numpy(), dist)


def test_knn_degenerate_geometry(eng):
    rng = np.random.default_rng(0)
    line = np.stack([rng.uniform(0, 10, 500), np.full(500, 3.0)], axis=1)  # collinear
    same = np.concatenate([np.zeros((40, 2)), rng.uniform(0, 1, (60, 2))])  # 40 identical points
    far = np.concatenate([rng.uniform(0, 1, (300, 2)), rng.uniform(1e5, 1e5 + 1, (300, 2))])  # big empty gap
    for coords in (line, same, far):
        graph, _, _ = eng.knn_graph(coords, 7)
        idx, _ = R.knn_bruteforce(coords, 7)
        assert np.array_equal(graph.indices.cpu().numpy(), idx)


def test_knn_errors(eng):
    c = np.random.default_rng(0).uniform(0, 1, (10, 2))
    with pytest.raises(ValueError, match="k must be < number of cells"):
        eng.knn_graph(c, 10)
    with pytest.raises(ValueError, match="n_neighbors must be >= 1"):
        eng.knn_graph(c, 0)


def test_radius_graph_matches_oracle(eng):
    coords = inputs.clustered(5000, seed=2)
    for r in (3.0, 12.0, 40.0):
        graph, _ = eng.radius_graph(coords, r, want_dist=True)
        indptr, indices, dist = R.radius_graph(coords, r)
        assert np.array_equal(graph.indptr.cpu().numpy(), indptr)
        assert np.array_equal(graph.indices.cpu().numpy(), indices)
        assert np.array_equal(graph.dist.cpu().numpy(), dist)
    # inclusive boundary on a lattice: neighbours at distance exactly r are kept
    lat = inputs.lattice(20, 20)
    graph, _ = eng.radius_graph(lat, 1.0)
    indptr, indices, _ = R.radius_graph(lat, 1.0)
    assert np.array_equal(graph.indptr.cpu().numpy(), indptr)
    assert np.array_equal(graph.indices.cpu().numpy(), indices)
    assert int(graph.indptr[-1]) == 2 * (19 * 20) * 2


def test_graph_moments_golden(eng, g0):
    coords, _, _ = g0
    graph, _, _ = eng.knn_graph(coords, 6)
    s0, s1, s2 = eng.graph_moments(graph)
    adj, _ = R.spatial_neighbors(coords, k=6)
    e0, e1, e2 = R.graph_moments(R.row_normalize(adj))
    np.testing.assert_allclose([s0, s1, s2], [e0, e1, e2], rtol=1e-12)
    np.testing.assert_allclose([s0, s1, s2], [10000.0, 3038.6111111, 40903.1666667], rtol=1e-9)
    # radius graph (variable degree, empty rows) and explicit weights
    cc = inputs.clustered(3000, seed=8)
    graph, _ = eng.radius_graph(cc, 9.0)
    adj, _ = R.spatial_neighbors(cc, radius=9.0)
    np.testing.assert_allclose(eng.graph_moments(graph), R.graph_moments(R.row_normalize(adj)), rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# standardisation
# ---------------------------------------------------------------------------------------------


def test_zscore_dense_sparse_subset(eng):
    rng = np.random.default_rng(1)
    X = np.log1p(rng.poisson(0.6, (3001, 37))).astype(np.float64)
    X[:, 5] = 2.5  # zero variance, non-zero constant
    X[:, 11] = 0.0
    Zr, mean, std, zero = R.zscore(X)
    for arr in (X.astype(np.float32), X):
        s = eng.zscore_dense(torch.from_numpy(arr).cuda())
        np.testing.assert_allclose(s.mean.cpu().numpy(), arr.astype(np.float64).mean(0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(s.std.cpu().numpy(), arr.astype(np.float64).std(0), rtol=1e-10, atol=1e-14)
        assert np.array_equal(s.zero_var.cpu().numpy().astype(bool), zero)
        Zq, _, _, _ = R.zscore(arr)
        np.testing.assert_allclose(s.Z[:, :37].cpu().numpy(), Zq, rtol=2e-7, atol=1e-7)
        assert torch.all(s.Z[:, 37:] == 0)
    cols = np.array([30, 2, 5, 17, 2])
    Xd, c = eng.expression_to_device(X.astype(np.float32), cols)
    s = eng.zscore_dense(Xd, cols=c)
    np.testing.assert_allclose(s.Z[:, :5].cpu().numpy(), Zr[:, cols], rtol=2e-7, atol=1e-7)
    for fmt in (sparse.csr_matrix, sparse.csc_matrix):
        for sel in (None, cols, np.array([4, 9, 20])):
            Xd, c = eng.expression_to_device(fmt(X.astype(np.float32)), sel)
            s = eng.zscore_dense(Xd, cols=c)
            want = Zr if sel is None else Zr[:, sel]
            np.testing.assert_allclose(s.Z[:, : want.shape[1]].cpu().numpy(), want, rtol=2e-7, atol=1e-7)


@pytest.mark.parametrize("shape", [(4001, 1000), (777, 64), (2500, 20), (300, 2052), (9, 8)])
def test_zscore_vector_path_matches_generic_and_numpy(eng, shape, monkeypatch):
    """FP32 matrices whose rows are float4-readable take the vectorised kernels (thread per column
    quad); they must agree with the generic kernels and with numpy FP64, with and without a row map."""
    n, g = shape
    rng = np.random.default_rng(g)
    X = np.log1p(rng.poisson(0.7, (n, g))).astype(np.float32)
    X[:, g // 2] = 1.25  # zero variance
    Xd = torch.from_numpy(X).cuda()
    rows = torch.from_numpy(rng.permutation(n).astype(np.int32)).cuda()
    fast = eng.zscore_dense(Xd, rows=rows)
    monkeypatch.setenv("SC_ZSCORE_GENERIC", "1")
    slow = eng.zscore_dense(Xd, rows=rows)
    monkeypatch.delenv("SC_ZSCORE_GENERIC")
    Zq, mean, std, zero = R.zscore(X)
    for s in (fast, slow):
        np.testing.assert_allclose(s.mean.cpu().numpy(), mean, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(s.std.cpu().numpy(), std, rtol=1e-10, atol=1e-14)
        assert np.array_equal(s.zero_var.cpu().numpy().astype(bool), zero)
        np.testing.assert_allclose(s.Z[:, :g].cpu().numpy(), Zq[rows.cpu().numpy()], rtol=2e-7, atol=1e-7)
        assert torch.all(s.Z[:, g:] == 0) and torch.all(s.Z[:, g // 2] == 0)
    np.testing.assert_allclose(fast.Z.cpu().numpy(), slow.Z.cpu().numpy(), rtol=0, atol=1.2e-7)  # x/std vs x*(1/std)


# ---------------------------------------------------------------------------------------------
# Moran's I and the graph-row null
# ---------------------------------------------------------------------------------------------


def _moran_device(eng, coords, X, k):
    graph, _, _ = eng.knn_graph(coords, k)
    std = eng.zscore_dense(torch.from_numpy(X).cuda())
    num, den, lag, _ = eng.lag_moran(graph, std.Z, X.shape[1])
    return graph, std, num, den, lag


def test_morans_i_statistic(eng, g0):
    coords, X, _ = g0
    graph, std, num, den, lag = _moran_device(eng, coords, X, 6)
    I = (num / den).cpu().numpy()  # N/S0 = 1 for a row-standardised kNN graph
    t = R.morans_i_table(coords, X, k=6, n_perms=0)
    assert np.all(np.abs(I - t["I"]) <= 1e-5 * np.abs(t["I"]) + 1e-7)
    np.testing.assert_allclose(I[:3], [-0.003001858001485166, -0.0038586272813623985, 0.0010727504543280996], rtol=1e-5, atol=1e-7)
    # lag against scipy
    W = R.build_spatial_weights(coords, 6).astype(np.float64)
    Z, _, _, _ = R.zscore(X)
    np.testing.assert_allclose(lag[:, :50].cpu().numpy(), W @ Z, rtol=1e-5, atol=2e-6)


def test_graph_rows_null_replay_counts_identical(eng, g0):
    coords, X, _ = g0
    n, g = X.shape
    graph, std, num, den, lag = _moran_device(eng, coords, X, 6)
    perms = R.squidpy_perm_indices(n, 99, 0)
    t = R.morans_i_table(coords, X, k=6, n_perms=99, seed=0)
    sims = eng.perm_null_graph_rows(std.Z, lag, g, 99, perm_idx=torch.from_numpy(perms.astype(np.int32)).cuda())
    sims = (sims / den).cpu().numpy()
    np.testing.assert_allclose(sims, t["sims"], rtol=1e-5, atol=2e-7)
    I = (num / den).cpu().numpy()
    # Poisson counts make the statistic lattice-valued: a permutation can tie the observed value
    # EXACTLY in exact arithmetic (integer Σ x_i S_π(i) equal), and then `>=` is decided by
    # 1e-17 round-off in the FP64 oracle itself.  Everywhere else the decisions must be identical.
    exact_tie = np.abs(t["sims"] - t["I"][None, :]) < 1e-12
    assert exact_tie.sum() <= 5
    differs = (sims >= I[None, :]) != (t["sims"] >= t["I"][None, :])
    assert not (differs & ~exact_tie).any()
    tie_free = ~exact_tie.any(0)
    assert np.array_equal((sims >= I[None, :]).sum(0)[tie_free], t["count_ge"][tie_free])
    assert np.array_equal(R.pval_sim_folded(I, sims)[:3], [0.32, 0.21, 0.38])
    # continuous expression has no lattice ties: counts and p-values identical for every gene
    cc, Xc = inputs.g0_continuous()
    tc = R.morans_i_table(cc, Xc, k=6, n_perms=99, seed=0)
    graph, std, num, den, lag = _moran_device(eng, cc, Xc.astype(np.float32), 6)
    tc32 = R.morans_i_table(cc, Xc.astype(np.float32), k=6, n_perms=99, seed=0)
    sc_ = eng.perm_null_graph_rows(std.Z, lag, Xc.shape[1], 99, perm_idx=torch.from_numpy(perms.astype(np.int32)).cuda())
    sc_ = (sc_ / den).cpu().numpy()
    Ic = (num / den).cpu().numpy()
    assert np.array_equal((sc_ >= Ic[None, :]).sum(0), tc32["count_ge"])
    assert np.array_equal(R.pval_sim_folded(Ic, sc_), tc32["pval_sim"])
    n, g = X.shape
    graph, std, num, den, lag = _moran_device(eng, coords, X, 6)
    # identity permutation reproduces the observed statistic exactly
    ident = torch.arange(n, dtype=torch.int32, device="cuda").reshape(1, -1)
    s1 = eng.perm_null_graph_rows(std.Z, lag, g, 1, perm_idx=ident)
    np.testing.assert_allclose(s1[0].cpu().numpy(), num.cpu().numpy(), rtol=1e-13)


@pytest.mark.parametrize("shape", [(3000, 7), (20011, 50), (9000, 260), (4000, 1000), (1500, 1030), (900, 2048)])
def test_graph_rows_kernel_variants_agree(eng, shape, monkeypatch):
    """The bulk-async pipeline and the register-staged kernels compute the same sums (FP64
    accumulation of identical FP32 products; only the summation order differs), on every geometry:
    1/2/4/8 warps per row, several column blocks, ragged last permutation batch."""
    n, g = shape
    rng = np.random.default_rng(n + g)
    A = torch.zeros((n, eng.padded_ld(g)), dtype=torch.float32, device="cuda")
    B = torch.zeros_like(A)
    A[:, :g] = torch.from_numpy(rng.normal(size=(n, g)).astype(np.float32)).cuda()
    B[:, :g] = torch.from_numpy(rng.normal(size=(n, g)).astype(np.float32)).cuda()
    P = 19
    perms = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)
    want = np.stack([(A[:, :g].double().cpu().numpy() * B[:, :g].double().cpu().numpy()[pm]).sum(0) for pm in perms])
    pidx = torch.from_numpy(perms).cuda()
    outs = {}
    for variant in ("bulk8", "bulk16", "8d", "16f"):
        monkeypatch.setenv("SC_PERM_ROWS_VARIANT", variant)
        outs[variant] = eng.perm_null_graph_rows(A, B, g, P, perm_idx=pidx).cpu().numpy()
        tol = 2e-4 if variant.endswith("f") else 1e-9
        np.testing.assert_allclose(outs[variant], want, rtol=0, atol=tol * np.sqrt(n)), variant
    monkeypatch.setenv("SC_PERM_ROWS_VARIANT", "bulk8")
    s1 = eng.perm_null_graph_rows(A, B, g, P, seed=5, perm_offset=2)
    monkeypatch.setenv("SC_PERM_ROWS_VARIANT", "8d")
    s2 = eng.perm_null_graph_rows(A, B, g, P, seed=5, perm_offset=2)
    np.testing.assert_allclose(s1.cpu().numpy(), s2.cpu().numpy(), rtol=0, atol=1e-9 * np.sqrt(n))


def test_philox_device_matches_host_mirror(eng, g0):
    from spatialcore_b200 import philox

    coords, X, _ = g0
    n, g = X.shape
    for m in (1, 2, 7, 1000, 4097, 100_000):
        dev = eng.philox_permutation(99, 5, m).cpu().numpy()
        assert np.array_equal(dev, philox.permutation(99, 5, m))
        assert np.array_equal(np.sort(dev), np.arange(m))
    graph, std, num, den, lag = _moran_device(eng, coords, X, 6)
    P = 21
    sims_dev = eng.perm_null_graph_rows(std.Z, lag, g, P, seed=1234, perm_offset=3)
    host = np.stack([philox.permutation(1234, 3 + p, n) for p in range(P)])
    sims_rep = eng.perm_null_graph_rows(std.Z, lag, g, P, perm_idx=torch.from_numpy(host).cuda())
    assert torch.equal(sims_dev, sims_rep)  # same kernel arithmetic, same indices -> bitwise equal
    # and against the oracle driven by the mirrored permutations
    gN = R.row_normalize(R.spatial_neighbors(coords, k=6)[0])
    ref = R.morans_i_perms_graph_rows(gN, X, host)
    np.testing.assert_allclose((sims_dev / den).cpu().numpy(), ref, rtol=1e-5, atol=2e-7)


def test_morans_i_api_end_to_end(api, g0):
    coords, X, _ = g0
    names = [f"g{i}" for i in range(X.shape[1])]
    a = _adata(X, coords, var_names=names)
    sel = ["g7", "g0", "g1", "g2", "g49"]
    out = api.morans_i(a, genes=sel, n_neighbors=6, n_permutations=99, seed=0, perm_source="replay")
    assert out is a
    df = a.uns["morans_i"]
    assert list(df.columns) == ["gene", "I", "expected_I", "z_score", "p_value"]
    assert df["gene"].tolist() == sel
    idx = [names.index(s) for s in sel]
    t = R.morans_i_table(coords, X[:, idx], k=6, n_perms=99, seed=0)
    assert np.all(np.abs(df["I"].to_numpy() - t["I"]) <= 1e-5 * np.abs(t["I"]) + 1e-7)
    assert np.array_equal(df["p_value"].to_numpy(), t["p_value"])  # replayed permutations: identical
    np.testing.assert_allclose(df["z_score"].to_numpy(), t["z_score"], rtol=1e-4, atol=1e-5)
    assert df["expected_I"].iloc[0] == -1 / (X.shape[0] - 1)
    # squidpy-style side effects
    adj, dst = R.spatial_neighbors(coords, k=6)
    assert (a.obsp["spatial_connectivities"] != adj).nnz == 0 and a.obsp["spatial_connectivities"].dtype == np.float64
    assert (a.obsp["spatial_distances"] != dst).nnz == 0
    assert a.uns["spatial_neighbors"]["params"]["n_neighbors"] == 6
    op = a.uns["spatialcore_metadata"]["operations"][-1]
    assert op["function"] == "morans_i" and op["parameters"]["backend"] == "b200" and op["outputs"] == {"uns": "morans_i"}
    # no permutations -> analytic p-value; copy=True leaves the input untouched
    b = _adata(X, coords, var_names=names)
    c = api.morans_i(b, genes=sel, n_permutations=0, copy=True)
    assert "morans_i" not in b.uns and c is not b
    np.testing.assert_allclose(c.uns["morans_i"]["p_value"].to_numpy(), t["pval_norm"], rtol=1e-4, atol=1e-6)
    # existing graph (here: a radius graph, variable degree)
    adj_r, _ = R.spatial_neighbors(coords, radius=25.0)
    d = _adata(X, coords, var_names=names)
    d.obsp["spatial_connectivities"] = adj_r
    api.morans_i(d, genes=sel, n_permutations=19, seed=2, use_existing_graph=True, perm_source="replay")
    tr = R.morans_i_table(coords, X[:, idx], n_perms=19, seed=2, adj=adj_r)
    dfr = d.uns["morans_i"]
    assert np.all(np.abs(dfr["I"].to_numpy() - tr["I"]) <= 1e-5 * np.abs(tr["I"]) + 1e-7)
    assert np.array_equal(dfr["p_value"].to_numpy(), tr["p_value"])
    np.testing.assert_allclose(dfr["z_score"].to_numpy(), tr["z_score"], rtol=1e-4, atol=1e-5)
    # radius keyword builds the same graph on the device
    e = _adata(X, coords, var_names=names)
    api.morans_i(e, genes=sel, n_permutations=19, seed=2, radius=25.0, perm_source="replay")
    assert np.array_equal(e.uns["morans_i"]["p_value"].to_numpy(), tr["p_value"])
    assert (e.obsp["spatial_connectivities"] != adj_r).nnz == 0


def test_morans_i_api_errors(api, g0):
    coords, X, _ = g0
    a = _adata(X[:100], coords[:100])
    with pytest.raises(ValueError, match=r"adata.obsm\['nope'\] not found"):
        api.morans_i(a, spatial_key="nope")
    with pytest.raises(ValueError, match="n_neighbors must be >= 1, got 0"):
        api.morans_i(a, n_neighbors=0)
    with pytest.raises(ValueError, match="n_permutations must be >= 0, got -1"):
        api.morans_i(a, n_permutations=-1)
    with pytest.raises(ValueError, match="Genes not found in adata.var_names"):
        api.morans_i(a, genes=["zzz"])
    with pytest.raises(ValueError, match="Invalid fdr_correction"):
        api.local_morans_i(a, fdr_correction="bogus")
    with pytest.raises(ValueError, match="Must provide either 'gene_pairs' or 'genes'"):
        api.lees_l_local(a)
    with pytest.raises(ValueError, match="significance_filter=True requires compute_cell_pvalues=True"):
        api.lees_l_local(a, gene_pairs=("g0", "g1"), significance_filter=True)


# ---------------------------------------------------------------------------------------------
# value-permuting null: Lee's L and local statistics (reference-own code, pinned by goldens)
# ---------------------------------------------------------------------------------------------


def test_lees_l_matches_reference(api, g0):
    coords, X, ref = g0
    a = _adata(X.astype(np.float64), coords)
    pairs = [("g0", "g1"), ("g1", "g0"), ("g0", "g0"), ("g3", "g7")]
    res = api.lees_l(a, pairs, n_neighbors=6, n_permutations=99, seed=0, perm_source="replay")
    L = np.array([r["L"] for r in res])
    p = np.array([r["p_value"] for r in res])
    assert np.all(np.abs(L - ref["lee_f64_L"]) <= 1e-5 * np.abs(ref["lee_f64_L"]) + 1e-4)
    assert np.array_equal(p, ref["lee_f64_p"])
    assert [r["gene_x"] for r in res] == [x for x, _ in pairs]
    single = api.lees_l(a, ("g0", "g1"), n_permutations=0)
    assert isinstance(single, dict) and single["p_value"] == 1.0
    assert "spatialcore_metadata" not in a.uns  # lees_l is pure


def test_lees_l_zero_variance_and_continuous(api, golden_dir):
    ref = np.load(os.path.join(golden_dir, "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = _adata(X, coords)
    res = api.lees_l(a, [("g0", "g1"), ("g1", "g0"), ("g5", "g5")], n_permutations=99, seed=0, perm_source="replay")
    L = np.array([r["L"] for r in res])
    assert np.all(np.abs(L - ref["leec_L"]) <= 1e-5 * np.abs(ref["leec_L"]) + 1e-4)
    assert np.array_equal([r["p_value"] for r in res], ref["leec_p"])
    Xz = X.copy()
    Xz[:, 2] = 1.0
    r = api.lees_l(_adata(Xz, coords), [("g2", "g1"), ("g0", "g1")], n_permutations=9, seed=0, perm_source="replay")
    assert r[0] == {"gene_x": "g2", "gene_y": "g1", "L": 0.0, "p_value": 1.0}


def _local_flip_margins(coords, X, k, perms, mism):
    """FP64 oracle margins at the (cell, gene) entries where a per-cell permutation p-value differs: the smallest
    relative distance between |I_perm| and |I_obs| over the permutations (a flip is legitimate only at a near tie)."""
    cells, genes = np.nonzero(mism)
    if cells.size == 0:
        return np.zeros(0)
    Z = R.zscore(X.astype(np.float64))[0]
    W = R.build_spatial_weights(coords, k).astype(np.float64)
    obs = np.abs(Z * (W @ Z))[cells, genes]
    best = np.full(cells.size, np.inf)
    for pm in perms:
        Zs = Z[pm]
        best = np.minimum(best, np.abs(np.abs(Zs * (W @ Zs))[cells, genes] - obs))
    return best / np.maximum(obs, 1e-300)


def test_local_morans_i_matches_reference(api, golden_dir):
    ref = np.load(os.path.join(golden_dir, "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = _adata(X, coords)
    api.local_morans_i(a, genes=["g0", "g1", "g2"], n_neighbors=6, n_permutations=99, seed=0, perm_source="replay")
    I, z, lag, p = (a.obsm[f"local_morans_{s}"] for s in ("I", "z", "lag", "p"))
    assert I.dtype == np.float32 and a.obsm["local_morans_quadrant"].dtype == np.int8
    np.testing.assert_allclose(I, ref["lmc_I"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(z, ref["lmc_z"], rtol=5e-6, atol=1e-6)
    np.testing.assert_allclose(lag, ref["lmc_lag"], rtol=5e-5, atol=2e-6)
    mism = p != ref["lmc_p"]
    assert mism.mean() < 2e-4, f"per-cell p-value flips: {mism.sum()} of {mism.size}"
    # every flip must be a near tie: in an FP64 evaluation of the same permutations some |I_perm| lies within FP32
    # rounding of |I_obs| for that (cell, gene) -- the margins are printed
    margins = _local_flip_margins(coords, X[:, :3], 6, R.squidpy_perm_indices(len(coords), 99, 0), mism)
    print(f"local Moran: {int(mism.sum())} p-value flips of {mism.size}; oracle margins min|abs(I_p)-abs(I)|/abs(I) at the flips: "
          f"{np.sort(margins)[-5:] if margins.size else margins}")
    assert margins.size == 0 or margins.max() < 2e-5
    q = a.obsm["local_morans_quadrant"]
    assert (q != ref["lmc_quadrant"]).mean() < 2e-4
    prm = a.uns["local_morans_params"]
    assert prm["genes"] == ["g0", "g1", "g2"] and prm["n_cells"] == 10000 and prm["zero_variance_genes"] == []
    assert a.uns["spatialcore_metadata"]["operations"][-1]["function"] == "local_morans_i"
    # batch_size changes the permutation stream exactly like the reference (2nd batch continues it)
    b = _adata(X, coords)
    api.local_morans_i(b, genes=["g0", "g1", "g2"], n_permutations=99, seed=0, batch_size=2, perm_source="replay")
    assert np.array_equal(b.obsm["local_morans_p"][:, :2], p[:, :2])
    assert (b.obsm["local_morans_p"][:, 2] != p[:, 2]).mean() > 0.3
    # n_permutations = 0: sign-only quadrants
    c = _adata(X, coords)
    api.local_morans_i(c, genes=["g0"], n_permutations=0)
    assert np.array_equal(c.obsm["local_morans_quadrant"][:, 0], R.quadrants(z[:, 0], lag[:, 0]))
    assert np.all(c.obsm["local_morans_p"] == 1)


def test_lees_l_local_matches_reference(api, golden_dir):
    ref = np.load(os.path.join(golden_dir, "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = _adata(X, coords)
    api.lees_l_local(a, gene_pairs=[("g0", "g1"), ("g2", "g3")], n_neighbors=6, n_permutations=19,
                     compute_cell_pvalues=True, significance_filter=True, alpha=0.2, seed=0, perm_source="replay")
    for key in ("g0_g1", "g2_g3"):
        np.testing.assert_allclose(a.obs[f"{key}_lees_l"].to_numpy(), ref[f"llc_{key}_L"], rtol=5e-5, atol=2e-6)
        assert (a.obs[f"{key}_pvalue"].to_numpy() != ref[f"llc_{key}_p"]).mean() < 2e-4
        assert (a.obs[f"{key}_quadrant"].cat.codes.to_numpy() != ref[f"llc_{key}_q"]).mean() < 2e-4
        prm = a.uns[f"{key}_lees_l_params"]
        assert abs(prm["global_L"] - ref[f"llc_{key}_global"][0]) <= 1e-5 * abs(ref[f"llc_{key}_global"][0]) + 1e-4
        assert prm["global_pvalue"] == ref[f"llc_{key}_global"][1]
        assert sum(prm["quadrant_counts"].values()) == 10000
    assert a.obs["g0_g1_quadrant"].cat.categories.tolist() == ["NS", "HH", "LL", "HL", "LH"]


def test_lee_matrix_all_pairs(api, eng, g0):
    coords, X, ref = g0
    a = _adata(X, coords)
    df = api.lees_l_matrix(a, n_neighbors=6)
    W = R.build_spatial_weights(coords, 6).astype(np.float64)
    Z, _, _, _ = R.zscore(X)
    want = R.lees_l_all_pairs(Z, W)
    got = df.to_numpy()
    # entries are sums of N products of O(1) values: the natural absolute scale is sqrt(N)
    assert np.all(np.abs(got - want) <= 1e-5 * np.abs(want) + 4e-6 * np.sqrt(X.shape[0]))
    np.testing.assert_allclose([got[0, 1], got[1, 0], got[0, 0]], ref["lee_f64_L"][:3], rtol=1e-5)
    assert abs(got[0, 1] - got[1, 0]) > 1.0  # not symmetric
    t = R.morans_i_table(coords, X, k=6, n_perms=0)
    np.testing.assert_allclose(np.diag(got), X.shape[0] * t["I"], rtol=1e-4, atol=1e-3)  # L[x,x] = N·I[x]
    sym = api.lees_l_matrix(a, n_neighbors=6, variant="lee2001").to_numpy()
    lagZ = W @ Z
    np.testing.assert_allclose(sym, lagZ.T @ lagZ / X.shape[0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape", [(64, 32), (1000, 50), (5003, 300), (20000, 1000), (9000, 1300)])
def test_lee_gemm_tensor_core_path(eng, shape):
    """impl 2 = tcgen05 kind::tf32 with 3xTF32 splitting.  The tensor core truncates when it adds
    into its FP32 accumulator, which leaves a relative bias of up to ~5e-6 per 256-cell chunk (all-positive sums are the worst case); the bar is
    1e-5 of the matrix maximum and 1e-5 median relative error.  impl 1 (FP64 accumulate) is exact."""
    n, g = shape
    torch.manual_seed(n + g)
    ld = eng.padded_ld(g)
    A = torch.zeros((n, ld), device="cuda")
    B = torch.zeros((n, ld), device="cuda")
    A[:, :g] = torch.randn((n, g), device="cuda")
    B[:, :g] = 0.4 * torch.randn((n, g), device="cuda") + 0.1 * A[:, :g]
    ref = A[:, :g].double().T @ B[:, :g].double()
    exact = eng.lee_gemm(A, B, g, impl=1).double()
    assert (exact - ref).abs().max() <= 2e-7 * ref.abs().max() + 1e-6  # only the FP32 rounding of the output
    fast = eng.lee_gemm(A, B, g, impl=2).double()
    err = (fast - ref).abs()
    assert err.max() <= 1e-5 * ref.abs().max(), float(err.max())
    assert (err / ref.abs().clamp_min(1e-12)).median() <= 1e-5
    # symmetric variant (A == B): (WZ)ᵀ(WZ)
    sym = eng.lee_gemm(B, B, g, impl=2).double()
    refs = B[:, :g].double().T @ B[:, :g].double()
    assert (sym - refs).abs().max() <= 1e-5 * refs.abs().max()
    assert torch.equal(sym, sym.T)  or (sym - sym.T).abs().max() <= 1e-5 * refs.abs().max()


# ---------------------------------------------------------------------------------------------
# neighbourhood composition
# ---------------------------------------------------------------------------------------------


def test_neighborhood_profile_bit_identical(api, golden_dir):
    ref = np.load(os.path.join(golden_dir, "ref_nbhd.npz"))
    coords, labels = inputs.nbhd()

    def run(**kw):
        a = _adata(np.zeros((coords.shape[0], 1), np.float32), coords,
                   obs=pd.DataFrame({"ct": pd.Categorical([f"t{c:02d}" for c in labels])}))
        api.compute_neighborhood_profile(a, "ct", **kw)
        return a

    a = run(method="knn", k=5)
    assert a.obsm["neighborhood_profile"].dtype == np.float32
    assert np.array_equal(a.obsm["neighborhood_profile"], ref["knn5_norm"])
    assert a.uns["neighborhood_profile_celltypes"] == ref["celltypes"].tolist()
    assert np.array_equal(run(method="knn", k=30).obsm["neighborhood_profile"], ref["knn30_norm"])
    assert np.array_equal(run(method="knn", k=30, normalize=False).obsm["neighborhood_profile"], ref["knn30_raw"])
    assert np.array_equal(run(method="radius", radius=inputs.NBHD_RADIUS, normalize=False).obsm["neighborhood_profile"], ref["radius_raw"])
    assert np.array_equal(run(method="radius", radius=inputs.NBHD_RADIUS).obsm["neighborhood_profile"], ref["radius_norm"])
    with pytest.raises(ValueError, match="cells have empty neighborhood profiles"):
        run(method="radius", radius=0.5)
    with pytest.raises(ValueError, match="'radius' must be provided"):
        run(method="radius")
    with pytest.raises(ValueError, match="k must be < number of cells"):
        run(method="knn", k=6000)
    with pytest.raises(ValueError, match="Invalid method"):
        run(method="ball")


def test_nbhd_counts_from_graph(eng):
    coords, labels = inputs.nbhd()
    graph, _, fused = eng.knn_graph(coords, 12, labels=torch.from_numpy(labels.astype(np.int32)).cuda(), n_types=8)
    sep = eng.nbhd_counts(graph, torch.from_numpy(labels.astype(np.int32)).cuda(), 8)
    assert torch.equal(fused, sep)
    want = R.neighborhood_profile(coords, labels, 8, k=12, normalize=False)
    assert np.array_equal(sep.cpu().numpy(), want)


# ---------------------------------------------------------------------------------------------
# spatial order: relabelled graph / matrices give the same statistics
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("kind", ["knn", "radius", "weighted"])
def test_spatial_order_and_graph_relabel(eng, kind):
    rng = np.random.default_rng(11)
    n = 6011
    coords = np.concatenate([rng.uniform(0, 300, (n - 2000, 2)), rng.normal(150, 5, (2000, 2))])
    cd = torch.from_numpy(coords).cuda()
    co = eng.spatial_order(cd)
    order, rank = co.order.cpu().numpy(), co.rank.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(n))
    assert np.array_equal(rank[order], np.arange(n))
    # spatially compact: consecutive sorted cells are close (median step far below the extent)
    step = np.linalg.norm(np.diff(coords[order], axis=0), axis=1)
    assert np.median(step) < 10.0
    if kind == "knn":
        graph, _, _ = eng.knn_graph(cd, 7)
    else:
        graph, _ = eng.radius_graph(cd, 6.0)
        assert (np.diff(graph.indptr.cpu().numpy()) == 0).any()  # some empty rows survive relabelling
    A = graph.to_scipy("ones", np.float64)
    if kind == "weighted":
        A.data = rng.uniform(0.5, 2.0, A.nnz)
        graph = eng.graph_from_scipy(A, use_weights=True)
    gs = eng.relabel_graph(graph, co)
    B = sparse.csr_matrix(
        (np.ones(gs.nnz) if gs.weights is None else gs.weights.cpu().numpy().astype(np.float64),
         gs.indices.reshape(-1).cpu().numpy(), gs.indptr_tensor().cpu().numpy()), shape=(n, n))
    want = sparse.csr_matrix(A.astype(np.float32).astype(np.float64))[order][:, order]
    want.sort_indices()
    B2 = B.copy(); B2.sort_indices()
    assert np.array_equal(B.indices, B2.indices), "relabelled rows are not column-sorted"
    assert np.array_equal(want.indptr, B.indptr) and np.array_equal(want.indices, B.indices)
    np.testing.assert_array_equal(want.data, B.data)


def test_statistics_invariant_under_spatial_order(eng, g0):
    coords, X, _ = g0
    n, g = X.shape
    cd = torch.from_numpy(coords).cuda()
    Xd = torch.from_numpy(X).cuda()
    graph, _, _ = eng.knn_graph(cd, 6)
    co = eng.spatial_order(cd)
    gs = eng.relabel_graph(graph, co)
    su, ss = eng.zscore_dense(Xd), eng.zscore_dense(Xd, rows=co.order)
    assert torch.equal(ss.Z, su.Z[co.order.long()])
    assert torch.equal(eng.gather_rows(ss.Z, co.rank), su.Z)
    nu, du, lag_u, loc_u = eng.lag_moran(graph, su.Z, g, want_local=True)
    ns, ds, lag_s, loc_s = eng.lag_moran(gs, ss.Z, g, want_local=True)
    np.testing.assert_allclose(ns.cpu().numpy(), nu.cpu().numpy(), rtol=0, atol=5e-5)  # FP32 lag rounded in a different neighbour order
    np.testing.assert_allclose(ds.cpu().numpy(), du.cpu().numpy(), rtol=1e-13)
    # per-cell lag: same neighbours, summed in a different order (FP32)
    np.testing.assert_allclose(eng.gather_rows(lag_s, co.rank).cpu().numpy(), lag_u.cpu().numpy(), rtol=0, atol=2e-6)
    # replayed permutations of cell ids, conjugated onto sorted positions
    perms = np.stack([np.random.default_rng(5).permutation(n) for _ in range(1)] + [R.squidpy_perm_indices(n, 20, 0)[j] for j in range(20)]).astype(np.int32)
    pidx = torch.from_numpy(perms).cuda()
    sims_u = eng.perm_null_graph_rows(su.Z, lag_u, g, len(perms), perm_idx=pidx)
    sims_s = eng.perm_null_graph_rows(ss.Z, lag_s, g, len(perms), perm_idx=eng.conjugate_perms(pidx, co))
    np.testing.assert_allclose(sims_s.cpu().numpy(), sims_u.cpu().numpy(), rtol=0, atol=5e-4)  # sums of ~1e4 O(1) terms, FP32 lag rounding


@pytest.mark.parametrize("shape,k", [((5003, 40), 6), ((3001, 130), 9)])
def test_values_null_materialised_matches_gather_kernel(eng, shape, k, monkeypatch):
    """Wide matrices take the materialise-then-lag variant; it must agree with the register-gather
    kernel (identical FP32 products, different summation order) and with a numpy FP64 evaluation."""
    n, g = shape
    rng = np.random.default_rng(n)
    coords = rng.uniform(0, 100, (n, 2))
    Xc = rng.normal(size=(n, g)).astype(np.float32)
    cd = torch.from_numpy(coords).cuda()
    graph, _, _ = eng.knn_graph(cd, k)
    std = eng.zscore_dense(torch.from_numpy(Xc).cuda())
    P = 5
    perms = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)
    pidx = torch.from_numpy(perms).cuda()
    _, _, _, loc = eng.lag_moran(graph, std.Z, g, want_lag=False, want_local=True)
    res = {}
    for variant in ("gather", "materialised"):
        if variant == "gather":
            monkeypatch.setenv("SC_PERM_VALUES_VARIANT", "gather")
        else:
            monkeypatch.delenv("SC_PERM_VALUES_VARIANT", raising=False)
        cnt = torch.zeros(std.Z.shape, dtype=torch.int32, device="cuda")
        sims = eng.perm_null_values(graph, std.Z, g, P, perm_idx=pidx, cell_obs=loc, cell_cnt=cnt)
        sims_ph = eng.perm_null_values(graph, std.Z, g, 3, seed=9, perm_offset=4)
        res[variant] = (sims.cpu().numpy(), cnt[:, :g].cpu().numpy(), sims_ph.cpu().numpy())
    W = graph.to_scipy("weights", np.float64)
    Z = std.Z[:, :g].double().cpu().numpy()
    want = np.stack([(Z[pm] * (W @ Z[pm])).sum(0) for pm in perms])
    for variant, (sims, cnt, sims_ph) in res.items():
        np.testing.assert_allclose(sims, want, rtol=0, atol=2e-5 * np.sqrt(n)), variant
    np.testing.assert_allclose(res["gather"][2], res["materialised"][2], rtol=0, atol=2e-5 * np.sqrt(n))
    # per-cell exceedance counts: identical except where |local_p| ties |obs| to FP32 rounding
    diff = res["gather"][1] != res["materialised"][1]
    assert diff.mean() < 1e-4, diff.mean()
    # Lee form (x fixed, only y permuted) on the wide path
    sims_lee = eng.perm_null_values(graph, std.Z, g, P, Zx=std.Z, perm_idx=pidx).cpu().numpy()
    want_lee = np.stack([(Z * (W @ Z[pm])).sum(0) for pm in perms])
    np.testing.assert_allclose(sims_lee, want_lee, rtol=0, atol=2e-5 * np.sqrt(n))


# ---------------------------------------------------------------------------------------------
# niches: k-means on the profile matrix
# ---------------------------------------------------------------------------------------------


def test_kmeans_lloyd_matches_sklearn_from_same_centres(eng, golden_dir):
    """Same starting centres -> the device Lloyd iteration must land on sklearn's fixed point:
    labels identical up to near-ties (sklearn ranks by |x|^2 - 2x.c + |c|^2 in FP32), centres and
    inertia to FP32 accuracy."""
    from spatialcore_b200.spatial import niches

    P = inputs.niche_profiles()
    rng = np.random.default_rng(3)
    c0 = P[rng.choice(P.shape[0], 6, replace=False)].astype(np.float32)
    want_l, want_c, want_i, _ = R.kmeans_lloyd_from(P, c0)
    km = eng.KMeansDevice(torch.from_numpy(P).cuda(), 6)
    var = P.astype(np.float64).var(0).mean()
    cent, inertia, n_iter = niches.lloyd(km, c0, 300, float(var * 1e-4))
    labels = km.labels.cpu().numpy()
    assert (labels != want_l).mean() < 2e-3
    assert R.adjusted_rand_index(labels, want_l) > 0.995
    np.testing.assert_allclose(cent, want_c, rtol=0, atol=2e-3)
    np.testing.assert_allclose(inertia, want_i, rtol=1e-4)
    # one assignment pass against a numpy FP64 evaluation: sums, counts, inertia, labels
    sums, counts, inert, changed = km.assign(want_c.astype(np.float32), want_mind=True)
    d2 = ((P[:, None, :].astype(np.float64) - want_c[None].astype(np.float64)) ** 2).sum(-1)
    lab64 = d2.argmin(1)
    lab = km.labels.cpu().numpy()
    near_tie = np.sort(d2, 1)[:, 1] - np.sort(d2, 1)[:, 0] < 1e-6
    assert np.array_equal(lab[~near_tie], lab64[~near_tie])
    np.testing.assert_allclose(inert, d2[np.arange(len(lab)), lab].sum(), rtol=1e-6)
    np.testing.assert_allclose(km.mind.cpu().numpy(), d2[np.arange(len(lab)), lab], rtol=2e-5, atol=1e-7)
    for k in range(6):
        assert counts[k] == (lab == k).sum()
        np.testing.assert_allclose(sums[k], P[lab == k].astype(np.float64).sum(0), rtol=1e-12, atol=1e-12)


def test_kmeans_seeding_kernels(eng):
    """k-means++ building blocks against numpy: candidate potentials with and without commit,
    D^2 sampling = searchsorted(cumsum)."""
    rng = np.random.default_rng(8)
    n, d = 7013, 9
    X = rng.random((n, d)).astype(np.float32)
    km = eng.KMeansDevice(torch.from_numpy(X).cuda(), 4)
    X64 = X.astype(np.float64)
    dist = lambda i: ((X64 - X64[i]) ** 2).sum(1)  # noqa: E731
    pot = km.pp_potential(np.array([17]), first=True, commit=0)
    np.testing.assert_allclose(pot[0], dist(17).sum(), rtol=1e-6)
    np.testing.assert_allclose(km.mind.cpu().numpy(), dist(17), rtol=2e-5, atol=1e-7)
    mind = km.mind.cpu().numpy().astype(np.float64)
    cand = np.array([5, 999, 7012, 17, 3000])
    pots = km.pp_potential(cand, first=False)
    want = np.array([np.minimum(mind, dist(c)).sum() for c in cand])
    np.testing.assert_allclose(pots, want, rtol=1e-6)
    vals = np.array([0.0, 1e-9, 0.25, 0.5, 0.999999, 1.0, 1.5]) * mind.sum()
    got = km.pp_sample(vals)
    ref = np.minimum(np.searchsorted(np.cumsum(mind), vals), n - 1)
    assert np.abs(got - ref).max() <= 1  # cumsum order differs in the last FP64 bit
    assert got[-1] == n - 1 and got[0] == ref[0]
    km.pp_potential(cand[1:2], first=False, commit=0)
    np.testing.assert_allclose(km.mind.cpu().numpy(), np.minimum(mind, dist(999)), rtol=2e-5, atol=1e-7)


def test_identify_niches_matches_reference(api, golden_dir):
    """Full drop-in against the frozen output of the unmodified reference: same partition (ARI),
    inertia within 0.1 %, schema and error behaviour of [R neighborhoods.py:299-522]."""
    g = np.load(os.path.join(golden_dir, "ref_niches.npz"))
    prof = {"nbhd": np.load(os.path.join(golden_dir, "ref_nbhd.npz"))["knn30_norm"], "dirichlet": inputs.niche_profiles()}
    for tag, P in prof.items():
        k = int(g[f"{tag}_k"])
        a = _adata(np.zeros((P.shape[0], 1), np.float32), np.zeros((P.shape[0], 2)))
        a.obsm["neighborhood_profile"] = P
        api.identify_niches(a, n_niches=k, random_state=0)
        codes = a.obs["niche"].cat.codes.to_numpy()
        assert list(a.obs["niche"].cat.categories) == [f"niche_{i + 1}" for i in range(k)]
        inertia = a.uns["niche_params"]["inertia"]
        assert inertia <= float(g[f"{tag}_inertia"]) * 1.001, (tag, inertia, float(g[f"{tag}_inertia"]))
        assert R.adjusted_rand_index(codes, g[f"{tag}_labels"]) > 0.97, tag
        assert a.uns["niche_centroids"].shape == (k, P.shape[1]) and a.uns["niche_centroids"].dtype == np.float32
        assert a.uns["niche_params"]["n_niches"] == k and a.uns["niche_params"]["method"] == "kmeans"
        assert a.uns["spatialcore_metadata"]["operations"][-1]["function"] == "identify_niches"
        # deterministic in random_state
        b = _adata(np.zeros((P.shape[0], 1), np.float32), np.zeros((P.shape[0], 2)))
        b.obsm["neighborhood_profile"] = P
        api.identify_niches(b, n_niches=k, random_state=0)
        assert np.array_equal(b.obs["niche"].cat.codes.to_numpy(), codes)
    a = _adata(np.zeros((50, 1), np.float32), np.zeros((50, 2)))
    with pytest.raises(ValueError, match="not found. Run compute_neighborhood_profile"):
        api.identify_niches(a, n_niches=3)
    a.obsm["neighborhood_profile"] = np.abs(np.random.default_rng(0).normal(size=(50, 4))).astype(np.float32)
    with pytest.raises(ValueError, match="Invalid method"):
        api.identify_niches(a, n_niches=3, method="dbscan")
    with pytest.raises(ValueError, match="n_niches must be >= 2"):
        api.identify_niches(a, n_niches=1)
    with pytest.raises(ValueError, match="cannot exceed number of cells"):
        api.identify_niches(a, n_niches=51)
    a.obsm["neighborhood_profile"][7] = 0
    with pytest.raises(ValueError, match="1 cells have empty neighborhood profiles"):
        api.identify_niches(a, n_niches=3)


# ---------------------------------------------------------------------------------------------
# domain distances
# ---------------------------------------------------------------------------------------------


def test_cross_nn_and_pairwise_kernels(eng):
    from scipy.spatial import cKDTree
    from scipy.spatial.distance import cdist

    rng = np.random.default_rng(12)
    for nt, nq in ((1, 50), (7, 300), (5000, 20000)):
        T = rng.uniform(0, 100, (nt, 2))
        Q = np.concatenate([rng.uniform(-400, 500, (nq // 2, 2)), rng.uniform(0, 100, (nq - nq // 2, 2))])  # inside and far outside the target box
        d, j = eng.cross_nn(T, Q)
        wd, wj = cKDTree(T).query(Q, k=1)
        assert np.array_equal(j, wj)
        assert np.array_equal(d, wd)  # sqrt(dx*dx + dy*dy) in FP64 on both sides
    T = inputs.lattice(12, 9)  # exact ties: lowest target index wins
    d, j = eng.cross_nn(T, np.array([[0.5, 0.5], [3.5, 2.0]]))
    assert list(j) == [0, 27] and np.allclose(d, [np.sqrt(0.5), 0.5])
    A, B = rng.normal(size=(1300, 2)) * 50, rng.normal(size=(2700, 2)) * 50 + 20
    dmin, dsum = eng.pairwise_reduce(A, B)
    D = cdist(A, B)
    assert dmin == D.min()
    np.testing.assert_allclose(dsum, D.sum(), rtol=1e-13)


def test_calculate_domain_distances_matches_reference(api, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_distances.npz"))
    coords, src, tgt = inputs.domains()
    cases = [("min_both", "src", "tgt", "minimum", "both"), ("min_matrix", "src", "tgt", "minimum", "matrix"),
             ("centroid_both", "src", "tgt", "centroid", "both"), ("mean_both", "src", "tgt", "mean", "both"),
             ("same_min_both", "tgt", "tgt", "minimum", "both"), ("same_centroid_both", "tgt", "tgt", "centroid", "both")]
    for tag, sc, tc, metric, mode in cases:
        obs = pd.DataFrame({"src": pd.Series(src, dtype=object), "tgt": pd.Series(tgt, dtype=object)})
        a = _adata(np.zeros((coords.shape[0], 1), np.float32), coords, obs=obs)
        api.calculate_domain_distances(a, sc, tc, distance_metric=metric, output_mode=mode)
        M = api.get_distance_matrix(a)
        assert list(M.index) == list(g[f"{tag}_rows"]) and list(M.columns) == list(g[f"{tag}_cols"])
        np.testing.assert_allclose(M.to_numpy(dtype=np.float64), g[f"{tag}_matrix"], rtol=1e-12, equal_nan=True)
        if mode != "matrix":
            np.testing.assert_allclose(a.obs["distance_to_target"].to_numpy(dtype=np.float64), g[f"{tag}_dist"], rtol=1e-13, equal_nan=True)
            near = np.array(["" if v is None or v != v else str(v) for v in a.obs["nearest_target_domain"]])
            assert np.array_equal(near, g[f"{tag}_nearest"]), tag
        else:
            assert "distance_to_target" not in a.obs.columns
        meta = a.uns["domain_distances"]
        assert meta["distance_metric"] == metric and set(meta["summary_statistics"]) == {"min_distance", "max_distance", "mean_distance", "median_distance"}
        assert a.uns["spatialcore_metadata"]["operations"][-1]["function"] == "calculate_domain_distances"
    # subsets and error behaviour [R distance.py:142-196]
    obs = pd.DataFrame({"src": pd.Series(src, dtype=object), "tgt": pd.Series(tgt, dtype=object)})
    a = _adata(np.zeros((coords.shape[0], 1), np.float32), coords, obs=obs)
    api.calculate_domain_distances(a, "src", "tgt", source_domain_subset=["Bcell_2"], target_domain_subset=["Tumor_1", "Tumor_4"], output_mode="matrix")
    M = api.get_distance_matrix(a)
    assert list(M.index) == ["Bcell_2"] and sorted(M.columns) == ["Tumor_1", "Tumor_4"]
    d, _, Mref = R.domain_distances(coords, src, tgt, ["Bcell_2"], list(M.columns), "minimum", "matrix")
    np.testing.assert_allclose(M.to_numpy(dtype=np.float64), Mref, rtol=1e-12)
    with pytest.raises(ValueError, match="Source column 'nope' not found"):
        api.calculate_domain_distances(a, "nope", "tgt")
    with pytest.raises(ValueError, match="Invalid distance_metric"):
        api.calculate_domain_distances(a, "src", "tgt", distance_metric="manhattan")
    with pytest.raises(ValueError, match="Invalid output_mode"):
        api.calculate_domain_distances(a, "src", "tgt", output_mode="table")
    with pytest.raises(ValueError, match="No valid source domains"):
        api.calculate_domain_distances(a, "src", "tgt", source_domain_subset=["Bcell_9"])
    with pytest.raises(KeyError, match="not found in adata.uns"):
        api.get_distance_matrix(_adata(np.zeros((5, 1), np.float32), np.zeros((5, 2))))


@pytest.mark.parametrize("method", ["fdr_bh", "bonferroni", "none"])
@pytest.mark.parametrize("n_perms", [19, 999])
def test_local_moran_epilogue_matches_numpy(eng, method, n_perms):
    """p = (c+1)/(P+1), per-gene BH / Bonferroni over the cells, quadrants and the un-sort, against the
    reference's numpy recipe [R autocorrelation.py:132-183, 219-265] evaluated on the host."""
    from spatialcore_b200.spatial import autocorrelation as ac

    rng = np.random.default_rng(n_perms)
    n, g = 5003, 7
    ld = eng.padded_ld(g)
    cnt = rng.integers(0, n_perms + 1, (n, ld)).astype(np.int32)
    cnt[:, 1] = np.minimum(cnt[:, 1], 2)          # heavy ties at small p
    cnt[: n // 50, 2] = 0                          # a block of minimal p-values
    Z = rng.normal(size=(n, ld)).astype(np.float32)
    lag = rng.normal(size=(n, ld)).astype(np.float32)
    Z[rng.random((n, ld)) < 0.01] = 0.0
    zero = np.zeros(g, np.uint8); zero[4] = 1
    order = rng.permutation(n).astype(np.int32)
    alpha = 0.05
    outs = eng.local_moran_finish(torch.from_numpy(cnt).cuda(), torch.from_numpy(Z).cuda(), torch.from_numpy(lag).cuda(),
                                  torch.from_numpy(Z * lag).cuda(), g, n_perms, torch.from_numpy(zero).cuda(), method, alpha,
                                  order=torch.from_numpy(order).cuda())
    z_d, lag_d, loc_d, p_d, pa_d, q_d = [o.cpu().numpy() for o in outs]
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)     # original row i is stored at inv[i]
    zr, lr = Z[inv][:, :g].copy(), lag[inv][:, :g].copy()
    p = ((cnt[inv][:, :g] + 1) / (n_perms + 1)).astype(np.float32)
    zr[:, 4] = 0; lr[:, 4] = 0; p[:, 4] = 1.0
    assert np.array_equal(z_d, zr) and np.array_equal(lag_d, lr) and np.array_equal(p_d, p)
    loc_want = (Z * lag)[inv][:, :g].copy(); loc_want[:, 4] = 0
    assert np.array_equal(loc_d, loc_want)
    pa = np.ones_like(p)
    for j in range(g):
        pa[:, j] = ac._fdr(p[:, j], method)
    assert np.array_equal(pa_d, pa)
    assert np.array_equal(q_d, ac._classify_quadrants(zr, lr, pa, alpha))
    # no permutations: sign-only quadrants, p = p_adj = 1
    outs0 = eng.local_moran_finish(None, torch.from_numpy(Z).cuda(), torch.from_numpy(lag).cuda(), torch.from_numpy(Z * lag).cuda(),
                                   g, 0, torch.from_numpy(zero).cuda(), method, alpha, order=torch.from_numpy(order).cuda())
    assert np.array_equal(outs0[5].cpu().numpy(), ac._classify_quadrants(zr, lr, None, alpha))
    assert bool((outs0[3] == 1).all()) and bool((outs0[4] == 1).all())


def test_lee_matrix_permutation_pvalues(api):
    """All-pairs Lee's L with the reference's null (only y permuted) for every pair per permutation:
    p-values identical to a numpy FP64 evaluation driven by the same replayed permutations."""
    rng = np.random.default_rng(17)
    n, g, P, k = 3001, 12, 29, 6
    coords = rng.uniform(0, 300, (n, 2))
    X = (np.log1p(rng.poisson(1.0, (n, g))) + 0.3 * rng.normal(size=(n, g))).astype(np.float32)
    X[:, 3] += np.sin(coords[:, 0] / 25.0).astype(np.float32)
    X[:, 7] += np.sin(coords[:, 0] / 25.0 + 0.4).astype(np.float32)
    X[:, 9] = 2.0  # zero variance
    a = _adata(X, coords)
    L, pv = api.lees_l_matrix(a, n_neighbors=k, n_permutations=P, seed=5, perm_source="replay", impl=1, key_added="lee")
    W = R.build_spatial_weights(coords, k).astype(np.float64)
    Z, _, _, zero = R.zscore(X)
    want_L = R.lees_l_all_pairs(Z, W)
    prng = np.random.default_rng(5)
    cnt = np.zeros((g, g), dtype=np.int64)
    margin = np.full((g, g), np.inf)
    for _ in range(P):
        perm = prng.permutation(n)
        Lp = Z.T @ (W @ Z[perm])
        cnt += np.abs(Lp) >= np.abs(want_L)
        margin = np.minimum(margin, np.abs(np.abs(Lp) - np.abs(want_L)))
    want_p = (cnt + 1) / (P + 1)
    want_p[zero, :] = 1.0
    want_p[:, zero] = 1.0
    live = ~zero
    np.testing.assert_allclose(L.to_numpy()[np.ix_(live, live)], want_L[np.ix_(live, live)], rtol=1e-5, atol=4e-6 * np.sqrt(n))
    safe = margin > 1e-3  # a permuted value within FP32 noise of the observed one may fall either side
    safe[zero, :] = True; safe[:, zero] = True
    assert safe.mean() > 0.98
    assert np.array_equal(pv.to_numpy()[safe], want_p[safe])
    assert pv.loc["g3", "g7"] <= 2 / (P + 1) and pv.loc["g9", "g1"] == 1.0
    assert a.uns["lee_pvalues"] is not None and list(pv.index) == list(L.index)
    # Philox source: deterministic, and the tensor-core path gives the same decisions away from ties
    _, p1 = api.lees_l_matrix(a, n_neighbors=k, n_permutations=P, seed=5, perm_source="philox", impl=1)
    _, p2 = api.lees_l_matrix(a, n_neighbors=k, n_permutations=P, seed=5, perm_source="philox", impl=2)
    assert (p1.to_numpy() != p2.to_numpy()).mean() < 0.03
    with pytest.raises(ValueError, match="variant='reference'"):
        api.lees_l_matrix(a, n_permutations=3, variant="lee2001")


def test_philox_null_pvalues_are_calibrated(eng):
    """SURVEY E12 on the on-device Philox permutations.  Value-permuting null (a true permutation test):
    two-tailed p-values of pure-noise genes are uniform (Kolmogorov-Smirnov).  Graph-row null (squidpy's
    scheme pairs z_i with the lag of a random OTHER cell, which drops the mutual-edge covariance of the
    observed statistic, so its p-values are not uniform by construction): its simulated sums must have
    exactly the moments of a uniform random pairing, E = (sum a)(sum b)/n and
    Var = sum (a - mean a)^2 * sum (b - mean b)^2 / (n - 1).  Disjoint permutation ranges (what
    different ranks run) give independent draws addressed by global index."""
    from scipy import stats

    rng = np.random.default_rng(2)
    n, g, P, k = 4000, 600, 199, 6
    coords = rng.uniform(0, 400, (n, 2))
    X = rng.normal(size=(n, g)).astype(np.float32)
    cd = torch.from_numpy(coords).cuda()
    graph, _, _ = eng.knn_graph(cd, k)
    co = eng.spatial_order(cd)
    gs = eng.relabel_graph(graph, co)
    std = eng.zscore_dense(torch.from_numpy(X).cuda(), rows=co.order)
    num, den, lag, _ = eng.lag_moran(gs, std.Z, g)
    sims = eng.perm_null_values(gs, std.Z, g, P, seed=9)
    c = (sims.abs() >= num.abs()[None, :]).sum(0).cpu().numpy()
    p = (c + 1) / (P + 1)
    ks = stats.kstest(p, "uniform")
    assert ks.pvalue > 1e-3, ks
    assert abs(p.mean() - 0.5) < 0.04 and abs((p <= 0.05).mean() - 0.05) < 0.03
    rows = eng.perm_null_graph_rows(std.Z, lag, g, P, seed=9).cpu().numpy()
    A, B = std.Z[:, :g].double(), lag[:, :g].double()
    mean_th = (A.sum(0) * B.sum(0) / n).cpu().numpy()
    var_th = (((A - A.mean(0)) ** 2).sum(0) * ((B - B.mean(0)) ** 2).sum(0) / (n - 1)).cpu().numpy()
    zmean = (rows.mean(0) - mean_th) / np.sqrt(var_th / P)
    assert abs(zmean.mean()) < 0.2 and 0.85 < zmean.std() < 1.15       # per-gene means ~ N(theory, var/P)
    ratio = rows.var(0, ddof=1) / var_th
    assert abs(ratio.mean() - 1.0) < 0.02                               # chi-square(P-1)/(P-1) averaged over 600 genes
    a = eng.perm_null_graph_rows(std.Z, lag, g, 50, seed=9, perm_offset=0)
    b = eng.perm_null_graph_rows(std.Z, lag, g, 50, seed=9, perm_offset=50)
    assert not torch.equal(a, b)
    corr = np.corrcoef(a.cpu().numpy().ravel(), b.cpu().numpy().ravel())[0, 1]
    assert abs(corr) < 0.03  # different global permutation indices: independent draws
    again = eng.perm_null_graph_rows(std.Z, lag, g, 100, seed=9, perm_offset=0)
    assert torch.equal(again[:50], a) and torch.equal(again[50:], b)  # addressed by global index, not by batch


def test_moran_known_answers_on_a_torus(api):
    """Closed-form Moran's I, analytic moments and z-scores on a 4-regular torus through the public API
    (``use_existing_graph=True``): checkerboard -> -1, cosine waves -> (1 + cos(2 pi / n)) / 2."""
    nx, ny = 24, 18
    A, coords = inputs.torus_rook(nx, ny)
    n = nx * ny
    i, j = coords[:, 0], coords[:, 1]
    X = np.stack([(-1.0) ** (i + j), np.cos(2 * np.pi * i / nx), np.cos(2 * np.pi * j / ny) + 0.5 * np.cos(2 * np.pi * i / nx)], axis=1)
    a = _adata(X.astype(np.float32), coords)
    a.obsp["spatial_connectivities"] = A
    api.morans_i(a, n_permutations=0, use_existing_graph=True)
    df = a.uns["morans_i"]
    l1, l2 = (1 + np.cos(2 * np.pi / ny)) / 2, (1 + np.cos(2 * np.pi / nx)) / 2
    want = np.array([-1.0, l2, (0.5 * l1 + 0.125 * l2) / 0.625])
    np.testing.assert_allclose(df["I"].to_numpy(), want, rtol=1e-5, atol=1e-7)
    vn = (n * n * (n / 2) - n * 4 * n + 3 * n * n) / ((n - 1) * (n + 1) * n * n) - 1 / (n - 1) ** 2
    np.testing.assert_allclose(df["z_score"].to_numpy(), (want + 1 / (n - 1)) / np.sqrt(vn), rtol=1e-5)
    assert np.all(df["expected_I"].to_numpy() == -1 / (n - 1))
    # the checkerboard is the most extreme negative pattern: every permuted statistic is larger
    api.morans_i(a, n_permutations=99, use_existing_graph=True, perm_source="philox", key_added="perm")
    p = a.uns["perm"]["p_value"].to_numpy()
    assert p[0] == 0.01 and p[1] == 0.01


def test_calls_run_on_the_device_of_their_operands(api, g0):
    """``device='cuda:1'`` while ``cuda:0`` is current: every library call must switch to the device of its operands
    (the library launches on the CURRENT device and takes the current stream)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    coords, X, _ = g0
    torch.cuda.set_device(0)
    a0 = _adata(X[:, :8], coords)
    api.morans_i(a0, n_neighbors=6, n_permutations=19, seed=0, perm_source="replay", device="cuda:0")
    a1 = _adata(X[:, :8], coords)
    api.morans_i(a1, n_neighbors=6, n_permutations=19, seed=0, perm_source="replay", device="cuda:1")
    assert torch.cuda.current_device() == 0
    pd.testing.assert_frame_equal(a0.uns["morans_i"], a1.uns["morans_i"])
    assert (a0.obsp["spatial_connectivities"] != a1.obsp["spatial_connectivities"]).nnz == 0
    b1 = _adata(X[:, :3], coords)
    api.local_morans_i(b1, n_permutations=9, perm_source="philox", device="cuda:1")
    b0 = _adata(X[:, :3], coords)
    api.local_morans_i(b0, n_permutations=9, perm_source="philox", device="cuda:0")
    assert np.array_equal(b0.obsm["local_morans_p"], b1.obsm["local_morans_p"])
