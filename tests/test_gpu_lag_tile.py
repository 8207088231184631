"""GPU tests of the shared-memory tile lag (``csrc/lag_tile.cu``): it must reproduce the L1-gather kernel
(``lag_stat_kernel``) BIT FOR BIT -- both add a row's neighbours in ascending column order -- on kNN and
radius graphs (empty rows, ragged tails), on graphs whose chunk unions overflow the tile (random graphs:
the direct-gather fallback inside the same call), and with the permutation applied
while staging (value-permuting null) against an explicitly permuted copy."""

import numpy as np
import pytest
import torch
from scipy import sparse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spatialcore_b200 import engine

    return engine


def _graphs(eng, n, seed):
    rng = np.random.default_rng(seed)
    coords = np.concatenate([rng.uniform(0, 400, (n - n // 3, 2)), rng.normal(200, 12, (n // 3, 2))])
    cd = torch.from_numpy(coords).cuda()
    co = eng.spatial_order(cd)
    knn, _, _ = eng.knn_graph(cd, 7)
    knn30, _, _ = eng.knn_graph(cd, 30)  # mean degree above 22: the 1 280-row / one-CTA-per-SM tile budget
    rad, _ = eng.radius_graph(cd, float(np.sqrt(6.0 * 160000.0 / (np.pi * n))))  # mean degree ~6 on the uniform part: some empty rows
    assert (np.diff(rad.indptr.cpu().numpy()) == 0).any()  # empty rows
    # a graph with no spatial structure: every chunk's union overflows the tile -> direct-gather fallback
    cols = np.sort(rng.integers(0, n, (n, 6)), axis=1)
    A = sparse.csr_matrix((np.ones(n * 6), cols.reshape(-1), np.arange(0, 6 * n + 1, 6)), shape=(n, n))
    A.sum_duplicates()
    A.data[:] = 1.0
    rnd = eng.graph_from_scipy(A)
    out = {}
    for name, g in (("knn", knn), ("radius", rad), ("knn30", knn30)):
        out[name] = eng.relabel_graph(g, co)
        out[name].tiles = None
    out["random"] = rnd
    return out


@pytest.mark.parametrize("n,g", [(3001, 40), (5120, 130), (777, 32), (20011, 1000)])
def test_tiled_lag_bit_identical_to_gather_kernel(eng, n, g):
    graphs = _graphs(eng, n, seed=n)
    Z = eng.zscore_dense(torch.from_numpy(np.random.default_rng(g).normal(size=(n, g)).astype(np.float32)).cuda()).Z
    for kind, gr in graphs.items():
        gr.tiles = None
        num0, den0, lag0, loc0 = eng.lag_moran(gr, Z, g, want_lag=True, want_local=True)
        eng.tile_graph(gr)
        assert gr.tiles is not None
        num1, den1, lag1, loc1 = eng.lag_moran(gr, Z, g, want_lag=True, want_local=True)
        assert torch.equal(lag0[:, :g], lag1[:, :g]), (kind, (lag0 - lag1).abs().max().item())
        assert torch.equal(loc0[:, :g], loc1[:, :g]), kind
        np.testing.assert_allclose(num1.cpu().numpy(), num0.cpu().numpy(), rtol=1e-11, atol=1e-9)
        np.testing.assert_allclose(den1.cpu().numpy(), den0.cpu().numpy(), rtol=1e-12)
        # statistic only (no lag written)
        num2, den2, _, _ = eng.lag_moran(gr, Z, g, want_lag=False)
        assert torch.equal(num2, num1) and torch.equal(den2, den1)
        gr.tiles = None


def test_tiled_values_null_matches_permuted_copy(eng):
    """perm applied while staging == lag of an explicitly permuted copy (bitwise per cell), Moran and Lee form,
    per-cell exceedance counters included; replayed and Philox permutations."""
    n, g, P = 4099, 70, 4
    graphs = _graphs(eng, n, seed=3)
    rng = np.random.default_rng(8)
    Z = eng.zscore_dense(torch.from_numpy(rng.normal(size=(n, g)).astype(np.float32)).cuda()).Z
    perms = np.stack([rng.permutation(n) for _ in range(P)]).astype(np.int32)
    pidx = torch.from_numpy(perms).cuda()
    for kind, gr in graphs.items():
        gr.tiles = None
        _, _, _, loc = eng.lag_moran(gr, Z, g, want_lag=False, want_local=True)
        want_sims, want_lee = [], []
        want_cnt = torch.zeros(Z.shape, dtype=torch.int32, device="cuda")
        for p in range(P):
            Zp = eng.gather_rows(Z, pidx[p])
            num, _, _, locp = eng.lag_moran(gr, Zp, g, want_lag=False, want_local=True)
            want_sims.append(num)
            want_cnt += (locp.abs() >= loc.abs()).int()
            _, _, lagp, _ = eng.lag_moran(gr, Zp, g, want_lag=True)
            want_lee.append((Z.double() * lagp.double()).sum(0)[:g])
        eng.tile_graph(gr)
        cnt = torch.zeros(Z.shape, dtype=torch.int32, device="cuda")
        sims = eng.perm_null_values(gr, Z, g, P, perm_idx=pidx, cell_obs=loc, cell_cnt=cnt)
        np.testing.assert_allclose(sims.cpu().numpy(), torch.stack(want_sims).cpu().numpy(), rtol=1e-11, atol=1e-9)
        assert torch.equal(cnt[:, :g], want_cnt[:, :g]), kind
        lee = eng.perm_null_values(gr, Z, g, P, Zx=Z, perm_idx=pidx)
        np.testing.assert_allclose(lee.cpu().numpy(), torch.stack(want_lee).cpu().numpy(), rtol=1e-10, atol=1e-7)
        # Philox permutations: the same numbers as replaying the mirrored indices
        ph = eng.perm_null_values(gr, Z, g, 2, seed=5, perm_offset=7)
        idx = torch.stack([eng.philox_permutation(5, 7 + j, n) for j in range(2)])
        rp = eng.perm_null_values(gr, Z, g, 2, perm_idx=idx)
        assert torch.equal(ph, rp)
        gr.tiles = None


def test_relabel_graph_builds_tiles_by_default(eng, monkeypatch):
    rng = np.random.default_rng(0)
    cd = torch.from_numpy(rng.uniform(0, 100, (2000, 2))).cuda()
    graph, _, _ = eng.knn_graph(cd, 6)
    co = eng.spatial_order(cd)
    assert eng.relabel_graph(graph, co).tiles is not None
    monkeypatch.setenv("SC_LAG_TILE_ROWS", "0")
    assert eng.relabel_graph(graph, co).tiles is None
    A = graph.to_scipy("ones", np.float64)
    A.data = rng.uniform(0.5, 2.0, A.nnz)
    monkeypatch.delenv("SC_LAG_TILE_ROWS")
    assert eng.relabel_graph(eng.graph_from_scipy(A, use_weights=True), co).tiles is None  # weighted: gather kernel
