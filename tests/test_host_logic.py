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


def test_morans_i_semantic_switches_on_cpu_stand_in(monkeypatch):
    """``null_mode`` / ``two_tailed`` / ``transformation`` of ``morans_i`` (the named switches for every recalled
    squidpy convention) against the oracle's restatement of each alternative, replayed permutations,
    on a radius graph (where row-normalised and binary weights differ) and a kNN graph (where they do not)."""
    from oracle import restate as R
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine

    monkeypatch.setattr(ac, "engine", cpu_engine)
    rng = np.random.default_rng(21)
    n, g, P = 1800, 6, 39
    coords = rng.uniform(0, 300, (n, 2))
    X = rng.normal(size=(n, g)).astype(np.float32)  # continuous data: no lattice ties
    X[:, 0] += np.sin(coords[:, 0] / 25.0).astype(np.float32)
    names = [f"g{i}" for i in range(g)]
    adj_r, _ = R.spatial_neighbors(coords, radius=14.0)

    def run(**kw):
        a = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
        ac.morans_i(a, n_permutations=kw.pop("P", P), seed=4, perm_source="replay", device="cpu", **kw)
        return a.uns["morans_i"], a

    for null_mode in ("graph_rows", "values"):
        for two_tailed in (False, True):
            for transformation in (True, False):
                df, a = run(radius=14.0, null_mode=null_mode, two_tailed=two_tailed, transformation=transformation)
                t = R.morans_i_table(coords, X, n_perms=P, seed=4, adj=adj_r, null_mode=null_mode, two_tailed=two_tailed,
                                     transformation=transformation)
                np.testing.assert_allclose(df["I"].to_numpy(), t["I"], rtol=1e-5, atol=1e-7)
                np.testing.assert_allclose(df["z_score"].to_numpy(), t["z_score"], rtol=1e-6)
                assert np.array_equal(df["p_value"].to_numpy(), t["p_value"]), (null_mode, two_tailed, transformation)
                par = a.uns["spatialcore_metadata"]["operations"][-1]["parameters"]
                assert (par["null_mode"], par["two_tailed"], par["transformation"]) == (null_mode, two_tailed, transformation)
    # the two weightings differ on a radius graph and coincide on a kNN graph
    i_norm, i_bin = run(radius=14.0)[0]["I"].to_numpy(), run(radius=14.0, transformation=False)[0]["I"].to_numpy()
    assert np.abs(i_norm - i_bin).max() > 1e-4
    k_norm, k_bin = run(n_neighbors=6)[0]["I"].to_numpy(), run(n_neighbors=6, transformation=False)[0]["I"].to_numpy()
    np.testing.assert_allclose(k_norm, k_bin, rtol=1e-5, atol=1e-7)
    # analytic p-value: doubled by two_tailed
    p1, p2 = run(P=0, radius=14.0)[0]["p_value"].to_numpy(), run(P=0, radius=14.0, two_tailed=True)[0]["p_value"].to_numpy()
    np.testing.assert_allclose(p2, 2.0 * p1, rtol=1e-12)
    # existing weighted graph, stored weights kept
    d = AnnDataLite(X, obsm={"spatial": coords}, var_names=names)
    w = adj_r.copy().astype(np.float64)
    w.data = rng.uniform(0.5, 2.0, w.nnz)
    d.obsp["spatial_connectivities"] = w
    ac.morans_i(d, n_permutations=0, use_existing_graph=True, transformation=False, device="cpu")
    tw = R.morans_i_table(coords, X, n_perms=0, adj=w, transformation=False)
    np.testing.assert_allclose(d.uns["morans_i"]["I"].to_numpy(), tw["I"], rtol=1e-5, atol=1e-7)
    import pytest as _pytest
    with _pytest.raises(ValueError, match="null_mode"):
        ac.morans_i(AnnDataLite(X, obsm={"spatial": coords}, var_names=names), null_mode="rows", device="cpu")


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


# ---------------------------------------------------------------------------------------------
# Lee's L, spatial weights and neighbourhood composition host logic on the stand-in engine
# ---------------------------------------------------------------------------------------------


def test_lees_l_host_logic_matches_reference_golden(monkeypatch):
    """``lees_l`` [R autocorrelation.py:991-1163] on the stand-in engine against the frozen output of the
    unmodified reference: pair normalisation, one permutation stream shared by the pairs in order,
    zero-variance short-cut, purity (no adata mutation), errors."""
    import pytest as _pytest

    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(ac, "engine", cpu_engine)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = AnnDataLite(X, obsm={"spatial": coords})
    res = ac.lees_l(a, [("g0", "g1"), ("g1", "g0"), ("g5", "g5")], n_permutations=99, seed=0, perm_source="replay", device="cpu")
    L = np.array([r["L"] for r in res])
    assert np.all(np.abs(L - ref["leec_L"]) <= 1e-5 * np.abs(ref["leec_L"]) + 1e-4)
    assert np.array_equal([r["p_value"] for r in res], ref["leec_p"])
    assert [(r["gene_x"], r["gene_y"]) for r in res] == [("g0", "g1"), ("g1", "g0"), ("g5", "g5")]
    assert "spatialcore_metadata" not in a.uns and not a.obsp           # pure
    single = ac.lees_l(a, ("g0", "g1"), n_permutations=0, device="cpu")
    assert isinstance(single, dict) and single["p_value"] == 1.0 and abs(single["L"] - ref["leec_L"][0]) < 1e-3
    # the second pair of a list continues the stream of the first: evaluating it alone gives other draws,
    # evaluating the first alone gives the same p-value
    assert ac.lees_l(a, ("g0", "g1"), n_permutations=99, seed=0, perm_source="replay", device="cpu")["p_value"] == ref["leec_p"][0]
    Xz = X.copy()
    Xz[:, 2] = 1.0
    r = ac.lees_l(AnnDataLite(Xz, obsm={"spatial": coords}), [("g2", "g1"), ("g0", "g1")], n_permutations=9, seed=0,
                  perm_source="replay", device="cpu")
    assert r[0] == {"gene_x": "g2", "gene_y": "g1", "L": 0.0, "p_value": 1.0}
    # Philox source: deterministic, addressed by (pair index, permutation index)
    p1 = [r["p_value"] for r in ac.lees_l(a, [("g0", "g1"), ("g2", "g3")], n_permutations=19, seed=4, perm_source="philox", device="cpu")]
    p2 = [r["p_value"] for r in ac.lees_l(a, [("g0", "g1"), ("g2", "g3")], n_permutations=19, seed=4, perm_source="philox", device="cpu")]
    assert p1 == p2 and all(0 < p <= 1 for p in p1)
    with _pytest.raises(ValueError, match="Genes not found"):
        ac.lees_l(a, ("g0", "nope"), device="cpu")
    with _pytest.raises(ValueError, match="not found"):
        ac.lees_l(AnnDataLite(X, obsm={}), ("g0", "g1"), device="cpu")
    with _pytest.raises(ValueError, match="n_permutations must be >= 0"):
        ac.lees_l(a, ("g0", "g1"), n_permutations=-2, device="cpu")


def test_lees_l_local_host_logic_matches_reference_golden(monkeypatch):
    """``lees_l_local`` [R autocorrelation.py:1171-1479]: two draws of P permutations per pair from one
    stream (global test, then per-cell test), categorical quadrants, parameter dicts, all-pairs
    expansion, zero-variance pairs, copy semantics, errors."""
    import pytest as _pytest

    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(ac, "engine", cpu_engine)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_g0.npz"))
    coords, X = inputs.g0_continuous()
    a = AnnDataLite(X, obsm={"spatial": coords})
    out = ac.lees_l_local(a, gene_pairs=[("g0", "g1"), ("g2", "g3")], n_neighbors=6, n_permutations=19, compute_cell_pvalues=True,
                          significance_filter=True, alpha=0.2, seed=0, perm_source="replay", device="cpu")
    assert out is a
    for key in ("g0_g1", "g2_g3"):
        np.testing.assert_allclose(a.obs[f"{key}_lees_l"].to_numpy(), ref[f"llc_{key}_L"], rtol=5e-5, atol=2e-6)
        assert (a.obs[f"{key}_pvalue"].to_numpy() != ref[f"llc_{key}_p"]).mean() < 2e-4
        assert (a.obs[f"{key}_quadrant"].cat.codes.to_numpy() != ref[f"llc_{key}_q"]).mean() < 2e-4
        prm = a.uns[f"{key}_lees_l_params"]
        assert abs(prm["global_L"] - ref[f"llc_{key}_global"][0]) <= 1e-5 * abs(ref[f"llc_{key}_global"][0]) + 1e-4
        assert prm["global_pvalue"] == ref[f"llc_{key}_global"][1]
        assert sum(prm["quadrant_counts"].values()) == 10000 and prm["compute_cell_pvalues"] is True
    assert a.obs["g0_g1_quadrant"].cat.categories.tolist() == ["NS", "HH", "LL", "HL", "LH"]
    op = a.uns["spatialcore_metadata"]["operations"][-1]
    assert op["function"] == "lees_l_local" and op["parameters"]["n_pairs"] == 2
    # all-pairs expansion in combinations() order; copy=True leaves the input untouched; without per-cell
    # p-values every p is 1 and the quadrants are sign-only
    small = AnnDataLite(X[:1500], obsm={"spatial": coords[:1500]})
    c = ac.lees_l_local(small, genes=["g0", "g1", "g2"], n_permutations=5, seed=1, copy=True, perm_source="replay", device="cpu")
    assert c is not small and not [k for k in small.obs.columns if k.endswith("_lees_l")]
    assert [k for k in c.obs.columns if k.endswith("_lees_l")] == ["g0_g1_lees_l", "g0_g2_lees_l", "g1_g2_lees_l"]
    assert np.all(c.obs["g0_g2_pvalue"].to_numpy() == 1.0) and (c.obs["g0_g2_quadrant"] != "NS").mean() > 0.9
    Xz = X[:1500].copy()
    Xz[:, 1] = 3.0
    z = ac.lees_l_local(AnnDataLite(Xz, obsm={"spatial": coords[:1500]}), gene_pairs=("g0", "g1"), n_permutations=3, device="cpu",
                        perm_source="replay")
    assert z.uns["g0_g1_lees_l_params"]["zero_variance"] is True and np.all(z.obs["g0_g1_lees_l"].to_numpy() == 0)
    assert set(z.obs["g0_g1_quadrant"]) == {"NS"}
    with _pytest.raises(ValueError, match="Must provide either"):
        ac.lees_l_local(a, device="cpu")
    with _pytest.raises(ValueError, match="significance_filter=True requires"):
        ac.lees_l_local(a, gene_pairs=("g0", "g1"), significance_filter=True, device="cpu")
    with _pytest.raises(ValueError, match="Genes not found"):
        ac.lees_l_local(a, gene_pairs=("g0", "zz"), device="cpu")


def test_lees_l_matrix_host_logic_on_cpu_stand_in(monkeypatch):
    """All-pairs matrix: spatial re-ordering of Z / graph / replayed permutations leaves L and the
    permutation p-values equal to a direct FP64 evaluation in the user's labelling."""
    import pytest as _pytest

    from oracle import restate as R
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine

    monkeypatch.setattr(ac, "engine", cpu_engine)
    rng = np.random.default_rng(17)
    n, g, P, k = 1201, 7, 19, 6
    coords = rng.uniform(0, 200, (n, 2))
    X = (np.log1p(rng.poisson(1.0, (n, g))) + 0.3 * rng.normal(size=(n, g))).astype(np.float32)
    X[:, 2] += np.sin(coords[:, 0] / 20.0).astype(np.float32)
    X[:, 4] += np.sin(coords[:, 0] / 20.0 + 0.4).astype(np.float32)
    X[:, 5] = 2.0
    a = AnnDataLite(X, obsm={"spatial": coords})
    L, pv = ac.lees_l_matrix(a, n_neighbors=k, n_permutations=P, seed=5, perm_source="replay", key_added="lee", device="cpu")
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
    safe = margin > 1e-3
    safe[zero, :] = True
    safe[:, zero] = True
    assert safe.mean() > 0.95 and np.array_equal(pv.to_numpy()[safe], want_p[safe])
    assert pv.loc["g2", "g4"] <= 2 / (P + 1) and pv.loc["g5", "g1"] == 1.0
    assert list(a.uns["lee"].index) == [f"g{i}" for i in range(g)] and a.uns["lee_pvalues"] is pv
    sym = ac.lees_l_matrix(a, genes=["g0", "g2", "g4"], n_neighbors=k, variant="lee2001", device="cpu")
    lagZ = (W @ Z)[:, [0, 2, 4]]
    np.testing.assert_allclose(sym.to_numpy(), lagZ.T @ lagZ / n, rtol=1e-5, atol=1e-6)
    with _pytest.raises(ValueError, match="variant='reference'"):
        ac.lees_l_matrix(a, n_permutations=3, variant="lee2001", device="cpu")
    with _pytest.raises(ValueError, match="variant must be"):
        ac.lees_l_matrix(a, variant="moran", device="cpu")


def test_spatial_weights_and_neighbors_host_logic(monkeypatch):
    """``build_spatial_weights`` [R autocorrelation.py:342-413] (scipy CSR assembly, FP32 data, int32
    indices) against the reference golden; squidpy-style graph slots for kNN and radius graphs."""
    import pytest as _pytest

    from oracle import restate as R
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(ac, "engine", cpu_engine)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_g0.npz"))
    coords, X = inputs.g0()
    a = AnnDataLite(X, obsm={"spatial": coords})
    for k in (6, 15):
        W = ac.build_spatial_weights(a, n_neighbors=k, device="cpu")
        assert W.dtype == np.float32 and W.indices.dtype == np.int32 and W.shape == (10000, 10000)
        assert np.array_equal(W.indices, ref[f"W_k{k}_indices"]) and np.array_equal(W.indptr, ref[f"W_k{k}_indptr"])
        assert np.array_equal(W.data[:8], ref[f"W_k{k}_data0"]) and np.all(W.data == W.data[0])
    Ws = ac.build_spatial_weights(a, n_neighbors=6, include_self=True, device="cpu")
    assert np.array_equal(Ws.indices, ref["W_k6_self_indices"]) and np.array_equal(Ws.data[:8], ref["W_k6_self_data0"])
    assert not a.obsp and "spatialcore_metadata" not in a.uns             # returns W, writes nothing
    b = AnnDataLite(X, obsm={"xy": coords})
    ac.spatial_neighbors(b, radius=25.0, spatial_key="xy", device="cpu")
    adj, dst = R.spatial_neighbors(coords, radius=25.0)
    assert b.obsp["spatial_connectivities"].dtype == np.float64 and (b.obsp["spatial_connectivities"] != adj).nnz == 0
    assert (b.obsp["spatial_distances"] != dst).nnz == 0
    assert b.uns["spatial_neighbors"]["params"]["radius"] == 25.0
    c = AnnDataLite(X, obsm={"spatial": coords})
    ac.spatial_neighbors(c, n_neighs=4, write=False, device="cpu")
    assert not c.obsp and "spatial_neighbors" not in c.uns
    with _pytest.raises(ValueError, match="not found"):
        ac.build_spatial_weights(b, device="cpu")
    with _pytest.raises(ValueError, match="n_neighbors must be >= 1"):
        ac.build_spatial_weights(a, n_neighbors=0, device="cpu")


def test_neighborhood_profile_host_logic_matches_reference_golden(monkeypatch):
    """``compute_neighborhood_profile`` [R neighborhoods.py:48-296]: label coding in sorted order, kNN and
    radius composition, normalisation, output slots, metadata and every argument error, bit-identical to
    the frozen reference output."""
    import pytest as _pytest

    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import neighborhoods as nb
    from tests import cpu_engine
    from tests.golden import inputs

    monkeypatch.setattr(nb, "engine", cpu_engine)
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_nbhd.npz"))
    coords, labels = inputs.nbhd()
    n = coords.shape[0]

    def run(obs=None, **kw):
        obs = pd.DataFrame({"ct": pd.Categorical([f"t{c:02d}" for c in labels])}) if obs is None else obs
        a = AnnDataLite(np.zeros((n, 1), np.float32), obs=obs, obsm={"spatial": coords})
        return a, nb.compute_neighborhood_profile(a, "ct", device="cpu", **kw)

    a, out = run(method="knn", k=5)
    assert out is a and a.obsm["neighborhood_profile"].dtype == np.float32
    assert np.array_equal(a.obsm["neighborhood_profile"], ref["knn5_norm"])
    assert a.uns["neighborhood_profile_celltypes"] == ref["celltypes"].tolist()
    op = a.uns["spatialcore_metadata"]["operations"][-1]
    assert op["function"] == "compute_neighborhood_profile" and op["parameters"]["k"] == 5 and op["parameters"]["radius"] is None
    assert np.array_equal(run(method="knn", k=30)[0].obsm["neighborhood_profile"], ref["knn30_norm"])
    assert np.array_equal(run(method="knn", k=30, normalize=False)[0].obsm["neighborhood_profile"], ref["knn30_raw"])
    assert np.array_equal(run(method="radius", radius=inputs.NBHD_RADIUS, normalize=False)[0].obsm["neighborhood_profile"], ref["radius_raw"])
    assert np.array_equal(run(method="radius", radius=inputs.NBHD_RADIUS)[0].obsm["neighborhood_profile"], ref["radius_norm"])
    # string labels in arbitrary order are coded by sorted unique value; copy / key_added
    a2, out2 = run(obs=pd.DataFrame({"ct": pd.Series([f"t{c:02d}" for c in labels], dtype=object)}), method="knn", k=5,
                   key_added="nbhd", copy=True)
    assert out2 is not a2 and "nbhd" not in a2.obsm and np.array_equal(out2.obsm["nbhd"], ref["knn5_norm"])
    assert out2.uns["nbhd_celltypes"] == ref["celltypes"].tolist()
    with _pytest.raises(ValueError, match="cells have empty neighborhood profiles"):
        run(method="radius", radius=0.5)
    with _pytest.raises(ValueError, match="'radius' must be provided"):
        run(method="radius")
    with _pytest.raises(ValueError, match="radius must be > 0"):
        run(method="radius", radius=-1.0)
    with _pytest.raises(ValueError, match="k must be < number of cells"):
        run(method="knn", k=n)
    with _pytest.raises(ValueError, match="k must be >= 1"):
        run(method="knn", k=0)
    with _pytest.raises(ValueError, match="Invalid method"):
        run(method="ball")
    with _pytest.raises(ValueError, match="not found in adata.obs"):
        nb.compute_neighborhood_profile(AnnDataLite(np.zeros((n, 1), np.float32), obsm={"spatial": coords}), "ct", device="cpu")
    with _pytest.raises(ValueError, match="At least 2 unique cell types"):
        run(obs=pd.DataFrame({"ct": ["a"] * n}))
    holes = pd.Series([f"t{c:02d}" for c in labels], dtype=object)
    holes[:7] = None
    with _pytest.raises(ValueError, match="7 cells have missing labels"):
        run(obs=pd.DataFrame({"ct": holes}))


# ---------------------------------------------------------------------------------------------
# the public entry points on 2 gloo ranks (stand-in engine): results do not depend on the world size
# ---------------------------------------------------------------------------------------------


def _sharded_inputs():
    rng = np.random.default_rng(23)
    n, g = 1501, 9
    coords = rng.uniform(0, 220, (n, 2))
    X = (np.log1p(rng.poisson(1.0, (n, g))) + 0.3 * rng.normal(size=(n, g))).astype(np.float32)
    X[:, 1] += np.sin(coords[:, 0] / 18.0).astype(np.float32)
    X[:, 6] = 4.0  # zero variance
    return coords, X


def _run_entry_points(ac, AnnDataLite, **kw):
    """The calls whose outputs must not depend on how many ranks share the work."""
    coords, X = _sharded_inputs()
    out = {}
    for tag, extra in (("perms_replay", dict(shard="perms", perm_source="replay")),
                       ("perms_philox", dict(shard="perms", perm_source="philox")),
                       ("genes_replay", dict(shard="genes", perm_source="replay")),
                       ("rows_philox", dict(shard="perms", ingest="sharded", perm_source="philox"))):
        a = AnnDataLite(X, obsm={"spatial": coords})
        if not kw.get("distributed", True):
            extra = dict(extra, shard="none", ingest="replicated")
        ac.morans_i(a, n_neighbors=6, n_permutations=23, seed=2, device="cpu", **extra)
        df = a.uns["morans_i"]
        out[f"mi_{tag}"] = np.stack([df["I"].to_numpy(), df["z_score"].to_numpy(), df["p_value"].to_numpy()])
    for src in ("replay", "philox"):
        a = AnnDataLite(X, obsm={"spatial": coords})
        ac.local_morans_i(a, n_neighbors=6, n_permutations=19, seed=4, batch_size=2, perm_source=src, device="cpu",
                          shard="auto" if kw.get("distributed", True) else "none")
        for s in ("I", "z", "lag", "p", "p_adj", "quadrant"):
            out[f"lm_{src}_{s}"] = a.obsm[f"local_morans_{s}"]
        out[f"lm_{src}_zero"] = np.array(a.uns["local_morans_params"]["zero_variance_genes"])
        L, pv = ac.lees_l_matrix(AnnDataLite(X, obsm={"spatial": coords}), n_neighbors=6, n_permutations=13, seed=6,
                                 perm_source=src, device="cpu", shard="auto" if kw.get("distributed", True) else "none")
        out[f"lee_{src}_L"], out[f"lee_{src}_p"] = L.to_numpy(), pv.to_numpy()
    return out


def _entry_point_worker(rank, world, port_no, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["SC_INGEST_NCCL"] = "1"  # no peer-mapped memory on CPU: all-gather form of the row-sharded ingest
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine

    ac.engine = cpu_engine
    np.savez(os.path.join(out_dir, f"e{rank}.npz"), **_run_entry_points(ac, AnnDataLite))
    dist.destroy_process_group()


def test_entry_points_on_two_gloo_ranks_match_single_process(tmp_path, monkeypatch):
    """SURVEY E13 on CPU: ``morans_i`` (permutation / gene / row sharding), ``local_morans_i`` (gene-batch
    sharding + all-gather of the per-cell matrices) and ``lees_l_matrix`` (permutation sharding) give,
    on every rank of a 2-rank job, the single-process result: identical counts and p-values in both
    replay and Philox mode."""
    from spatialcore_b200 import AnnDataLite
    from spatialcore_b200.spatial import autocorrelation as ac
    from tests import cpu_engine

    world = 2
    mp.spawn(_entry_point_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    monkeypatch.setattr(ac, "engine", cpu_engine)
    want = _run_entry_points(ac, AnnDataLite, distributed=False)
    for r in range(world):
        got = np.load(tmp_path / f"e{r}.npz")
        assert sorted(got.files) == sorted(want)
        for key, ref in want.items():
            if key == "mi_rows_philox":  # pooled moments differ in the last FP64 bit: Z agrees to FP32 rounding
                np.testing.assert_allclose(got[key][:2], ref[:2], rtol=2e-5, atol=1e-7, equal_nan=True, err_msg=key)
                assert (got[key][2] != ref[2]).sum() <= 1, key
            elif ref.dtype.kind == "f":
                np.testing.assert_allclose(got[key], ref, rtol=1e-12, atol=0, equal_nan=True, err_msg=key)
            else:
                assert np.array_equal(got[key], ref), key


# ---------------------------------------------------------------------------------------------
# property tests of the small host-side building blocks
# ---------------------------------------------------------------------------------------------


def test_host_building_blocks_properties():
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from oracle import restate as R
    from spatialcore_b200 import distributed as du
    from spatialcore_b200 import philox
    from spatialcore_b200.spatial import autocorrelation as ac

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 4000), st.integers(0, 2**40), st.integers(0, 10**6))
    def philox_is_a_bijection(n, seed, p):
        perm = philox.permutation(seed, p, n)
        assert perm.shape == (n,) and np.array_equal(np.sort(perm), np.arange(n))
        if n > 50:
            assert not np.array_equal(perm, philox.permutation(seed, p + 1, n))

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 10**6), st.integers(1, 64))
    def block_slices_tile_the_range(total, world):
        spans = [du.block_slice(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
        per, lo, hi = du.row_block(total, world - 1, world)
        assert per * world >= total and 0 <= lo <= hi <= total

    @settings(max_examples=40, deadline=None)
    @given(st.integers(2, 400), st.integers(1, 6), st.integers(0, 2**31 - 1))
    def pooled_moments_equal_global_moments(n, blocks, seed):
        rng = np.random.default_rng(seed)
        X = rng.normal(3.0, 2.0, (n, 4))
        X[:, 1] = -1.25                      # constant everywhere: exactly zero pooled variance
        cuts = np.sort(rng.integers(1, n, size=min(blocks, n) - 1)) if min(blocks, n) > 1 else np.array([], int)
        parts = [p for p in np.split(X, cuts) if len(p)]
        mean, std, zero = du.combine_moments([len(p) for p in parts], np.stack([p.mean(0) for p in parts]),
                                             np.stack([p.std(0) for p in parts]))
        np.testing.assert_allclose(mean, X.mean(0), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(std, X.std(0), rtol=1e-9, atol=1e-12)
        assert zero.tolist() == [False, True, False, False] and std[1] == 0.0

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.floats(0.0, 1.0), min_size=1, max_size=200))
    def bh_matches_the_oracle_and_is_monotone(ps):
        p = np.asarray(ps)
        adj = ac._fdr(p, "fdr_bh")
        np.testing.assert_allclose(adj, R.bh_adjust(p), rtol=1e-12, atol=0)
        assert np.all(adj >= p - 1e-15) and np.all(adj <= 1.0)
        order = np.argsort(p, kind="stable")
        assert np.all(np.diff(adj[order]) >= -1e-15)       # adjusted p-values keep the order of the raw ones
        np.testing.assert_allclose(ac._fdr(p, "bonferroni"), np.minimum(p * len(p), 1.0))
        assert np.array_equal(ac._fdr(p, "none"), p)

    philox_is_a_bijection()
    block_slices_tile_the_range()
    pooled_moments_equal_global_moments()
    bh_matches_the_oracle_and_is_monotone()


def test_three_dimensional_coordinates_are_refused_not_truncated():
    """The kernels are 2-D; the reference hands every column of ``obsm['spatial']`` to its tree libraries, so a
    varying third coordinate must raise instead of being dropped silently; constant extra columns are harmless."""
    import pytest as _pytest
    import torch

    from spatialcore_b200 import engine

    rng = np.random.default_rng(0)
    xy = rng.uniform(0, 10, (50, 2))
    flat = np.concatenate([xy, np.zeros((50, 1))], axis=1)
    out = engine._coords_tensor(flat, "cpu")
    assert out.shape == (50, 2) and np.array_equal(out.numpy(), xy)
    assert engine._coords_tensor(torch.from_numpy(flat), "cpu").shape == (50, 2)
    xyz = np.concatenate([xy, rng.uniform(0, 1, (50, 1))], axis=1)
    with _pytest.raises(ValueError, match="only 2-D coordinates are supported"):
        engine._coords_tensor(xyz, "cpu")
    with _pytest.raises(ValueError, match="only 2-D coordinates are supported"):
        engine._coords_tensor(torch.from_numpy(xyz), "cpu")
    with _pytest.raises(ValueError, match="shape"):
        engine._coords_tensor(xy[:, :1], "cpu")
