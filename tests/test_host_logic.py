"""CPU-only tests of the host-side logic: AnnData duck type, metadata log, FDR / quadrant helpers,
permutation-source selection, replay chunking, the C port of the CPU baseline, and the multi-rank
sharding + reduction path on the gloo backend (world_size 2)."""

import os
import socket
import sys

import numpy as np
import pandas as pd
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restate as R
from tests.golden import inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_anndata_lite_protocol():
    from spatialcore_b200 import AnnDataLite

    X = np.arange(12, dtype=np.float32).reshape(4, 3)
    a = AnnDataLite(X, obsm={"spatial": np.zeros((4, 2))}, var_names=["a", "b", "c"])
    assert a.n_obs == 4 and a.n_vars == 3 and a.shape == (4, 3)
    sub = a[:, ["c", "a"]]
    assert np.array_equal(sub.X, X[:, [2, 0]]) and list(sub.var_names) == ["c", "a"]
    b = a.copy()
    b.uns["k"] = 1
    b.X[0, 0] = 99
    assert "k" not in a.uns and a.X[0, 0] == 0
    with pytest.raises(ValueError):
        AnnDataLite(X, var_names=["a"])


def test_metadata_log_matches_reference_schema():
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.core import update_metadata

    a = AnnDataLite(np.zeros((2, 2), np.float32))
    update_metadata(a, "morans_i", {"genes": ("g0", "g1"), "seed": 0, "obj": object(), "nested": {"x": 1.5}}, {"uns": "morans_i"})
    update_metadata(a, "lees_l_local", {"n_pairs": 2})
    meta = a.uns["spatialcore_metadata"]
    assert set(meta) == {"created", "operations"} and len(meta["operations"]) == 2
    op = meta["operations"][0]
    assert set(op) == {"timestamp", "function", "parameters", "outputs"}
    assert op["parameters"] == {"genes": ["g0", "g1"], "seed": 0, "obj": "object", "nested": {"x": 1.5}}
    assert "outputs" not in meta["operations"][1]


def test_fdr_and_quadrants_match_oracle():
    from spatialcore_b200.spatial import autocorrelation as ac

    rng = np.random.default_rng(0)
    p = rng.uniform(size=500)
    p[:20] = 0.001
    np.testing.assert_array_equal(ac._fdr(p, "fdr_bh"), R.bh_adjust(p))
    np.testing.assert_array_equal(ac._fdr(p, "bonferroni"), np.clip(p * 500, 0, 1))
    np.testing.assert_array_equal(ac._fdr(p, "none"), p)
    assert ac._fdr(np.zeros(0), "fdr_bh").size == 0
    with pytest.raises(ValueError, match="Unknown FDR method"):
        ac._fdr(p, "holm")
    z, lag = rng.normal(size=(50, 3)), rng.normal(size=(50, 3))
    z[0, 0] = 0.0
    pa = rng.uniform(size=(50, 3))
    np.testing.assert_array_equal(ac._classify_quadrants(z, lag, pa, 0.3), R.quadrants(z, lag, pa, 0.3))
    assert ac._classify_quadrants(z, lag)[0, 0] == 0


def test_perm_source_selection_and_replay_stream():
    from spatialcore_b200.spatial import autocorrelation as ac

    assert ac._pick_perm_source("auto", 10_000, 99) == "replay"
    assert ac._pick_perm_source("auto", 5_000_000, 999) == "philox"
    assert ac._pick_perm_source("philox", 10, 1) == "philox"
    with pytest.raises(ValueError):
        ac._pick_perm_source("numpy", 10, 1)
    # chunks reproduce numpy's draw order exactly, whatever the chunking
    old = ac._REPLAY_CHUNK_ELEMS
    try:
        ac._REPLAY_CHUNK_ELEMS = 2500  # 2 permutations of 1000 per chunk
        got = torch.cat([c for _, c in ac._replay_chunks(np.random.default_rng(5), 1000, 7, "cpu")]).numpy()
    finally:
        ac._REPLAY_CHUNK_ELEMS = old
    rng = np.random.default_rng(5)
    want = np.stack([rng.permutation(1000) for _ in range(7)])
    np.testing.assert_array_equal(got, want)


def test_block_slice_partitions():
    from spatialcore_b200.distributed import block_slice

    for total in (0, 1, 7, 999, 1000):
        for ws in (1, 2, 3, 8):
            parts = [block_slice(total, r, ws) for r in range(ws)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_cpu_port_matches_restatement():
    from oracle import port

    coords, X = inputs.g0()
    adj, _ = R.spatial_neighbors(coords[:3000], k=6)
    g = R.row_normalize(adj)
    perms = R.squidpy_perm_indices(3000, 5, 1)
    score, sims = port.morans_i(g, X[:3000, :7].T.astype(np.float64), perms.astype(np.int32))
    np.testing.assert_allclose(score, R.morans_i_stat(g, X[:3000, :7]), rtol=1e-12)
    np.testing.assert_allclose(sims, R.morans_i_perms_graph_rows(g, X[:3000, :7], perms), rtol=1e-11)
    # radius graph (ragged rows, empty rows)
    adj, _ = R.spatial_neighbors(coords[:3000], radius=12.0)
    g = R.row_normalize(adj)
    score, sims = port.morans_i(g, X[:3000, :4].T.astype(np.float64), perms.astype(np.int32))
    np.testing.assert_allclose(sims, R.morans_i_perms_graph_rows(g, X[:3000, :4], perms), rtol=1e-11)
    assert port.threads() >= 1


# ---------------------------------------------------------------------------------------------
# multi-rank path on gloo (world_size 2): the same sharding / reduction code the GPU ranks run
# ---------------------------------------------------------------------------------------------


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialcore_b200 import distributed as du
    from spatialcore_b200.spatial.autocorrelation import MoranNull

    G, P = 11, 37
    rng = np.random.default_rng(0)
    sims = rng.normal(size=(P, G))
    obs = rng.normal(size=G) * 0.5
    # --- permutation sharding: every rank folds its block, one all-reduce of the [4, G] summary
    lo, hi = du.my_slice(P)
    null = MoranNull(G, "cpu")
    blk = torch.from_numpy(sims[lo:hi])
    null.cnt_ge += (blk >= torch.from_numpy(obs)).sum(0)
    null.cnt_abs_ge += (blk.abs() >= torch.from_numpy(obs).abs()).sum(0)
    null.sum += blk.sum(0)
    null.sumsq += (blk * blk).sum(0)
    du.all_reduce_null(null)
    # --- gene sharding: every rank owns a gene block, one all-gather of per-gene columns
    glo, ghi = du.block_slice(G, rank, world)
    local = np.stack([obs[glo:ghi], obs[glo:ghi] * 2])
    gathered = du.all_gather_columns(local, G, "cpu")
    # --- row-sharded ingest: per-block moments all-gathered and pooled, blocks all-gathered into Z
    n, g = 1001, 5
    X = np.random.default_rng(3).normal(2.0, 3.0, (n, g))
    X[:, 2] = 0.75  # constant gene: must pool to exactly zero variance
    per, rlo, rhi = du.row_block(n, rank, world)
    stats = torch.zeros((3, g), dtype=torch.float64)
    stats[0] = float(rhi - rlo)
    stats[1] = torch.from_numpy(X[rlo:rhi].mean(0))
    stats[2] = torch.from_numpy(X[rlo:rhi].std(0))
    allst = torch.empty((world * 3, g), dtype=torch.float64)  # concatenation form: accepted by gloo and NCCL
    dist.all_gather_into_tensor(allst, stats)
    h = allst.view(world, 3, g).numpy()
    mean, std, zero = du.combine_moments(h[:, 0, 0], h[:, 1], h[:, 2])
    Zu = torch.zeros((world * per, g), dtype=torch.float64)
    Zu[rlo:rhi] = torch.from_numpy(np.where(zero, 0.0, (X[rlo:rhi] - mean) / np.where(zero, 1.0, std)))
    dist.all_gather_into_tensor(Zu, Zu[rank * per:(rank + 1) * per].clone())
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), cnt=null.cnt_ge.numpy(), cabs=null.cnt_abs_ge.numpy(),
             s=null.sum.numpy(), ss=null.sumsq.numpy(), gathered=gathered, slice=np.array([lo, hi]),
             mean=mean, std=std, zero=zero, Z=Zu[:n].numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(0)
    sims = rng.normal(size=(37, 11))
    obs = rng.normal(size=11) * 0.5
    r0, r1 = (np.load(tmp_path / f"r{r}.npz") for r in range(2))
    assert r0["slice"].tolist() == [0, 19] and r1["slice"].tolist() == [19, 37]
    for r in (r0, r1):
        np.testing.assert_array_equal(r["cnt"], (sims >= obs).sum(0))
        np.testing.assert_array_equal(r["cabs"], (np.abs(sims) >= np.abs(obs)).sum(0))
        np.testing.assert_allclose(r["s"], sims.sum(0), rtol=1e-12)
        np.testing.assert_allclose(r["ss"], (sims**2).sum(0), rtol=1e-12)
        np.testing.assert_array_equal(r["gathered"], np.stack([obs, obs * 2]))
    X = np.random.default_rng(3).normal(2.0, 3.0, (1001, 5))
    X[:, 2] = 0.75
    for r in (r0, r1):
        np.testing.assert_allclose(r["mean"], X.mean(0), rtol=1e-13)
        np.testing.assert_allclose(r["std"], X.std(0), rtol=1e-12, atol=0)
        assert r["zero"].tolist() == [False, False, True, False, False] and r["std"][2] == 0.0
        want = (X - X.mean(0)) / np.where(X.std(0) == 0, 1.0, X.std(0))
        want[:, 2] = 0.0
        np.testing.assert_allclose(r["Z"], want, rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------------------------------------
# identify_niches host control flow (k-means++ stream, Lloyd convergence, restarts) on a numpy stand-in
# for the device object: the same Python code the GPU path runs, without a GPU
# ---------------------------------------------------------------------------------------------


class _NumpyKM:
    """Duck type of ``engine.KMeansDevice`` evaluated with numpy (FP64 distances)."""

    def __init__(self, X, k):
        self.X64 = np.asarray(X, dtype=np.float64)
        self.n, self.d = self.X64.shape
        self.k = k
        self.labels = torch.full((self.n,), -1, dtype=torch.int32)
        self.mind = torch.zeros(self.n, dtype=torch.float32)

    def _d2(self, C):
        return ((self.X64[:, None, :] - np.asarray(C, dtype=np.float64)[None, :, :]) ** 2).sum(-1)

    def assign(self, centers, want_mind=False):
        d2 = self._d2(centers)
        lab = d2.argmin(1).astype(np.int32)
        changed = int((self.labels.numpy() != lab).sum())
        self.labels = torch.from_numpy(lab)
        best = d2[np.arange(self.n), lab]
        if want_mind:
            self.mind = torch.from_numpy(best.astype(np.float32))
        sums = np.zeros((self.k, self.d))
        np.add.at(sums, lab, self.X64)
        return sums, np.bincount(lab, minlength=self.k).astype(np.float64), float(best.sum()), changed

    def pp_potential(self, cand, first, commit=-1):
        d2 = self._d2(self.X64[np.asarray(cand)])
        m = d2 if first else np.minimum(self.mind.numpy().astype(np.float64)[:, None], d2)
        if commit >= 0:
            self.mind = torch.from_numpy(m[:, commit].astype(np.float32))
        return m.sum(0)

    def pp_sample(self, vals):
        cum = np.cumsum(self.mind.numpy().astype(np.float64))
        return np.minimum(np.searchsorted(cum, vals), self.n - 1).astype(np.int64)

    def rows(self, idx):
        return self.X64[np.asarray(idx, dtype=np.int64)].astype(np.float32)


def test_kmeans_host_control_flow_matches_sklearn():
    from oracle import restate as R
    from spatialcore_b200.spatial import niches
    from tests.golden import inputs

    P = inputs.niche_profiles(n=4000, n_types=9, n_niches=5, seed=4)
    k = 5
    # Lloyd from sklearn's own starting point lands on sklearn's fixed point
    c0 = P[np.random.default_rng(1).choice(len(P), k, replace=False)]
    want_l, want_c, want_i, _ = R.kmeans_lloyd_from(P, c0)
    km = _NumpyKM(P, k)
    cent, inertia, n_iter = niches.lloyd(km, c0, 300, float(P.astype(np.float64).var(0).mean() * 1e-4))
    assert R.adjusted_rand_index(km.labels.numpy(), want_l) > 0.995
    np.testing.assert_allclose(inertia, want_i, rtol=1e-4)
    np.testing.assert_allclose(cent, want_c, atol=2e-3)
    assert 1 <= n_iter <= 300
    # seeding: k distinct rows of X, first index drawn like RandomState.choice(n) with uniform p
    rs = np.random.RandomState(7)
    first_expected = min(int(np.random.RandomState(7).random_sample() * km.n), km.n - 1)
    init = niches._kmeans_plusplus(km, rs)
    assert init.shape == (k, P.shape[1]) and len({tuple(r) for r in init.round(6)}) == k
    np.testing.assert_array_equal(init[0], P[first_expected])
    # an empty cluster takes the farthest point (sklearn's relocation) and the loop still converges
    far = np.vstack([c0[:4], np.full((1, P.shape[1]), 50.0)]).astype(np.float32)
    km2 = _NumpyKM(P, k)
    cent2, inertia2, _ = niches.lloyd(km2, far, 300, 0.0)
    assert np.bincount(km2.labels.numpy(), minlength=k).min() >= 1 and np.isfinite(inertia2)
    assert inertia2 < 2.0 * want_i


def test_domain_distance_host_logic_matches_reference_golden(monkeypatch):
    """The pandas / numpy bookkeeping of ``calculate_domain_distances`` (domain lists, matrix assembly from
    per-cell results, diagonal and fallback rules, output slots) against the frozen reference output,
    with the two device kernels replaced by the scipy calls they stand for."""
    import pandas as pd
    from scipy.spatial import cKDTree
    from scipy.spatial.distance import cdist

    from spatialcore_b200 import AnnDataLite, engine
    from spatialcore_b200.spatial import distance
    from tests.golden import inputs

    def fake_cross_nn(targets, queries, device="cpu"):
        d, j = cKDTree(np.asarray(targets)).query(np.asarray(queries), k=1)
        return d, j.astype(np.int64)

    def fake_pairwise(a, b, device="cpu"):
        D = cdist(a.numpy() if hasattr(a, "numpy") else a, b.numpy() if hasattr(b, "numpy") else b)
        return float(D.min()), float(D.sum())

    monkeypatch.setattr(engine, "cross_nn", fake_cross_nn)
    monkeypatch.setattr(engine, "pairwise_reduce", fake_pairwise)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_distances.npz"))
    coords, src, tgt = inputs.domains()
    cases = [("min_both", "src", "tgt", "minimum", "both"), ("min_matrix", "src", "tgt", "minimum", "matrix"),
             ("centroid_both", "src", "tgt", "centroid", "both"), ("mean_both", "src", "tgt", "mean", "both"),
             ("same_min_both", "tgt", "tgt", "minimum", "both"), ("same_centroid_both", "tgt", "tgt", "centroid", "both")]
    for tag, sc, tc, metric, mode in cases:
        obs = pd.DataFrame({"src": pd.Series(src, dtype=object), "tgt": pd.Series(tgt, dtype=object)})
        a = AnnDataLite(np.zeros((coords.shape[0], 1), np.float32), obs=obs, obsm={"spatial": coords})
        distance.calculate_domain_distances(a, sc, tc, distance_metric=metric, output_mode=mode, device="cpu")
        M = distance.get_distance_matrix(a)
        assert list(M.index) == list(g[f"{tag}_rows"]) and list(M.columns) == list(g[f"{tag}_cols"])
        np.testing.assert_allclose(M.to_numpy(dtype=np.float64), g[f"{tag}_matrix"], rtol=1e-12, equal_nan=True)
        if mode != "matrix":
            np.testing.assert_allclose(a.obs["distance_to_target"].to_numpy(dtype=np.float64), g[f"{tag}_dist"], rtol=1e-12, equal_nan=True)
            near = np.array(["" if v is None or v != v else str(v) for v in a.obs["nearest_target_domain"]])
            assert np.array_equal(near, g[f"{tag}_nearest"]), tag


# ---------------------------------------------------------------------------------------------
# morans_i host logic on the numpy stand-in engine (tests/cpu_engine.py)
# ---------------------------------------------------------------------------------------------


def test_morans_i_host_logic_on_cpu_stand_in(monkeypatch):
    """Validation, gene resolution, spatial re-ordering + permutation conjugation, null folding,
    p-value / z-score / table / metadata / obsp assembly of ``spatial.morans_i`` against the oracle,
    with every device call replaced by its numpy equivalent."""
    from oracle import restate as R
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(ac, "engine", cpu_engine)
    coords, X = inputs.g0()
    coords, X = coords[:2500], X[:2500, :12]
    names = [f"g{i}" for i in range(12)]
    sel = ["g7", "g0", "g1", "g11"]
    idx = [names.index(s) for s in sel]
    a = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
    out = ac.morans_i(a, genes=sel, n_neighbors=6, n_permutations=49, seed=3, perm_source="replay", device="cpu")
    assert out is a
    df = a.uns["morans_i"]
    t = R.morans_i_table(coords, X[:, idx], k=6, n_perms=49, seed=3)
    assert list(df.columns) == ["gene", "I", "expected_I", "z_score", "p_value"] and df["gene"].tolist() == sel
    np.testing.assert_allclose(df["I"].to_numpy(), t["I"], rtol=1e-5, atol=1e-7)
    tie = np.abs(t["sims"] - t["I"][None, :]).min(0) < 1e-9     # Poisson counts: exact lattice ties are order dependent
    assert np.array_equal(df["p_value"].to_numpy()[~tie], t["p_value"][~tie])
    np.testing.assert_allclose(df["z_score"].to_numpy(), t["z_score"], rtol=1e-6)
    adj, dst = R.spatial_neighbors(coords, k=6)
    assert (a.obsp["spatial_connectivities"] != adj).nnz == 0 and (a.obsp["spatial_distances"] != dst).nnz == 0
    assert a.uns["spatial_neighbors"]["params"] == {"n_neighbors": 6, "coord_type": "generic", "radius": None, "transform": None}
    op = a.uns["spatialcore_metadata"]["operations"][-1]
    assert op["function"] == "morans_i" and op["parameters"]["n_permutations"] == 49
    # Philox source goes through the host mirror of the device bijection: deterministic
    p1 = ac.morans_i(AnnDataLite(X, obsm={"spatial": coords}, var_names=names), genes=sel, n_permutations=29, seed=5,
                     perm_source="philox", device="cpu").uns["morans_i"]["p_value"].to_numpy()
    p2 = ac.morans_i(AnnDataLite(X, obsm={"spatial": coords}, var_names=names), genes=sel, n_permutations=29, seed=5,
                     perm_source="philox", device="cpu").uns["morans_i"]["p_value"].to_numpy()
    assert np.array_equal(p1, p2) and np.all((p1 > 0) & (p1 <= 0.5))
    # analytic p-value without permutations; copy semantics; existing weighted graph; radius keyword
    b = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
    c = ac.morans_i(b, genes=sel, n_permutations=0, copy=True, device="cpu")
    assert "morans_i" not in b.uns and c is not b
    np.testing.assert_allclose(c.uns["morans_i"]["p_value"].to_numpy(), t["pval_norm"], rtol=1e-6)
    adj_r, _ = R.spatial_neighbors(coords, radius=40.0)
    d = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
    d.obsp["spatial_connectivities"] = adj_r
    ac.morans_i(d, genes=sel, n_permutations=0, use_existing_graph=True, device="cpu")
    tr = R.morans_i_table(coords, X[:, idx], n_perms=0, adj=adj_r)
    np.testing.assert_allclose(d.uns["morans_i"]["I"].to_numpy(), tr["I"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(d.uns["morans_i"]["z_score"].to_numpy(), tr["z_score"], rtol=1e-6)
    e = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
    ac.morans_i(e, genes=sel, n_permutations=0, radius=40.0, device="cpu")
    np.testing.assert_allclose(e.uns["morans_i"]["I"].to_numpy(), tr["I"], rtol=1e-5, atol=1e-7)
    assert (e.obsp["spatial_connectivities"] != adj_r).nnz == 0
    # errors are raised before any compute [R autocorrelation.py:517-547]
    import pytest as _pytest
    with _pytest.raises(ValueError, match="not found"):
        ac.morans_i(AnnDataLite(X, obsm={}, var_names=names), device="cpu")
    with _pytest.raises(ValueError, match="n_neighbors must be >= 1"):
        ac.morans_i(a, n_neighbors=0, device="cpu")
    with _pytest.raises(ValueError, match="n_permutations must be >= 0"):
        ac.morans_i(a, n_permutations=-1, device="cpu")
    with _pytest.raises(ValueError, match="Genes not found"):
        ac.morans_i(a, genes=["nope"], device="cpu")
    with _pytest.raises(ValueError, match="perm_source"):
        ac.morans_i(a, genes=sel, perm_source="sobol", device="cpu")


def test_local_morans_i_host_logic_on_cpu_stand_in(monkeypatch):
    """``local_morans_i`` end to end on the numpy stand-in engine against the frozen output of the
    unmodified reference: batching over one shared permutation stream, spatial re-ordering and un-sorting,
    output slots and parameters."""
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(ac, "engine", cpu_engine)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = AnnDataLite(X, obsm={"spatial": coords})
    ac.local_morans_i(a, genes=["g0", "g1", "g2"], n_neighbors=6, n_permutations=99, seed=0, perm_source="replay", device="cpu")
    I, z, lag, p = (a.obsm[f"local_morans_{s}"] for s in ("I", "z", "lag", "p"))
    assert I.dtype == np.float32 and a.obsm["local_morans_quadrant"].dtype == np.int8
    np.testing.assert_allclose(I, ref["lmc_I"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(z, ref["lmc_z"], rtol=5e-6, atol=1e-6)
    np.testing.assert_allclose(lag, ref["lmc_lag"], rtol=5e-5, atol=2e-6)
    assert (p != ref["lmc_p"]).mean() < 2e-4
    assert (a.obsm["local_morans_quadrant"] != ref["lmc_quadrant"]).mean() < 2e-4
    prm = a.uns["local_morans_params"]
    assert prm["genes"] == ["g0", "g1", "g2"] and prm["n_cells"] == 10000 and prm["zero_variance_genes"] == []
    b = AnnDataLite(X, obsm={"spatial": coords})
    ac.local_morans_i(b, genes=["g0", "g1", "g2"], n_permutations=99, seed=0, batch_size=2, perm_source="replay", device="cpu")
    assert np.array_equal(b.obsm["local_morans_p"][:, :2], p[:, :2])       # first batch: same draws
    assert (b.obsm["local_morans_p"][:, 2] != p[:, 2]).mean() > 0.3         # second batch continues the stream
