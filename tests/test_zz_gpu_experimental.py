"""GPU tests of the EXPERIMENTAL opt-in kernels (row-group lag, ``SC_LAG_GROUP``; row alignment,
``SC_ROW_ALIGN``).  They were written after round 1's GPU budget was spent and have not run on a GPU
yet, so they are skipped unless ``SC_TEST_EXPERIMENTAL=1``; the default path does not touch these kernels."""

import os

import numpy as np
import pytest
import torch

from oracle import restate as R
from tests.golden import inputs

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(os.environ.get("SC_TEST_EXPERIMENTAL") != "1",
                       reason="experimental kernels, not yet run on a GPU: set SC_TEST_EXPERIMENTAL=1"),
]


@pytest.fixture(scope="module")
def eng():
    from spatialcore_b200 import engine

    return engine


def _union_model(indptr, indices, n, rows):
    """numpy model of sc_graph_group_build: per group of ``rows`` consecutive rows the sorted union of the
    neighbour lists with membership masks, at word round_up(CSR offset of the group's first row, 4) + 8 * group."""
    n_groups = (n + rows - 1) // rows
    words = np.zeros(len(indices) + 8 * n_groups + 8, dtype=np.uint32)
    cnt = np.zeros(n_groups, dtype=np.int32)
    offs = np.zeros(n_groups, dtype=np.int64)
    for a in range(n_groups):
        member = {}
        for r, row in enumerate(range(a * rows, min(n, (a + 1) * rows))):
            for j in indices[indptr[row]:indptr[row + 1]]:
                member[int(j)] = member.get(int(j), 0) | (1 << r)
        offs[a] = (int(indptr[a * rows]) + 3) // 4 * 4 + 8 * a
        for t, j in enumerate(sorted(member)):
            words[offs[a] + t] = (member[j] << (32 - rows)) | j
        cnt[a] = len(member)
    return words, cnt, offs


def _graphs(eng):
    rng = np.random.default_rng(11)
    coords = rng.uniform(0, 400, (3001, 2))  # not a multiple of 2, 4 or 8
    co = eng.spatial_order(coords)
    knn, _, _ = eng.knn_graph(coords, 6)
    rad, _ = eng.radius_graph(coords, 11.0)  # mean degree ~7, some empty rows
    return coords, co, {"knn": eng.relabel_graph(knn, co), "radius": eng.relabel_graph(rad, co)}


@pytest.mark.parametrize("rows", [2, 4, 8])
def test_group_build_matches_numpy_model(eng, rows):
    _, _, graphs = _graphs(eng)
    for kind, g in graphs.items():
        assert g.groups is None
        indptr = g.indptr_tensor().cpu().numpy()
        indices = g.indices.reshape(-1).cpu().numpy()
        eng.group_graph(g, rows)
        r, uwords, ucnt = g.groups
        want_w, want_c, offs = _union_model(indptr, indices, g.n, rows)
        got_w = uwords.cpu().numpy().view(np.uint32)
        assert r == rows and np.array_equal(ucnt.cpu().numpy(), want_c), kind
        assert got_w.size == want_w.size
        for a in range(len(want_c)):
            b = offs[a]
            assert np.array_equal(got_w[b:b + want_c[a]], want_w[b:b + want_c[a]]), (kind, a)
            pad = got_w[b + want_c[a]:b + (want_c[a] + 3) // 4 * 4]
            assert np.all(pad >> (32 - rows) == 0) and np.all((pad & ((1 << (32 - rows)) - 1)) < g.n)  # owner-less
        assert want_c.sum() < 0.95 * len(indices)  # spatial order: consecutive rows do share neighbours


@pytest.mark.parametrize("rows", [2, 4, 8])
def test_grouped_lag_matches_default_kernel(eng, rows):
    coords, co, graphs = _graphs(eng)
    rng = np.random.default_rng(12)
    for g_cols in (5, 40, 100):
        X = torch.from_numpy(rng.normal(size=(3001, g_cols)).astype(np.float32)).cuda()
        std = eng.zscore_dense(X, rows=co.order)
        for kind, g in graphs.items():
            g.groups = None
            num0, den0, lag0, loc0 = eng.lag_moran(g, std.Z, g_cols, want_lag=True, want_local=True)
            eng.group_graph(g, rows)
            num1, den1, lag1, loc1 = eng.lag_moran(g, std.Z, g_cols, want_lag=True, want_local=True)
            g.groups = None
            torch.testing.assert_close(lag1[:, :g_cols], lag0[:, :g_cols], rtol=1e-5, atol=5e-6)
            torch.testing.assert_close(loc1[:, :g_cols], loc0[:, :g_cols], rtol=1e-5, atol=2e-5)
            torch.testing.assert_close(num1, num0, rtol=1e-5, atol=5e-5)  # sums of 3001 products of FP32-rounded lags
            assert torch.equal(den1, den0) or torch.allclose(den1, den0, rtol=1e-12)
            W = g.to_scipy("weights", np.float64)
            z = std.Z[:, :g_cols].double().cpu().numpy()
            np.testing.assert_allclose(num1.cpu().numpy(), (z * (W @ z)).sum(0), rtol=1e-5, atol=5e-5)


def test_grouped_values_null_matches_default(eng):
    coords, co, graphs = _graphs(eng)
    rng = np.random.default_rng(13)
    g_cols, P = 40, 7
    X = torch.from_numpy(rng.normal(size=(3001, g_cols)).astype(np.float32)).cuda()
    std = eng.zscore_dense(X, rows=co.order)
    for kind, g in graphs.items():
        g.groups = None
        _, _, _, loc = eng.lag_moran(g, std.Z, g_cols, want_lag=False, want_local=True)
        cnt0 = torch.zeros(std.Z.shape, dtype=torch.int32, device="cuda")
        sims0 = eng.perm_null_values(g, std.Z, g_cols, P, seed=9, perm_offset=3, cell_obs=loc, cell_cnt=cnt0)
        eng.group_graph(g, 4)
        cnt1 = torch.zeros_like(cnt0)
        sims1 = eng.perm_null_values(g, std.Z, g_cols, P, seed=9, perm_offset=3, cell_obs=loc, cell_cnt=cnt1)
        g.groups = None
        torch.testing.assert_close(sims1, sims0, rtol=1e-5, atol=5e-5)
        assert (cnt1[:, :g_cols] != cnt0[:, :g_cols]).float().mean() < 1e-4, kind


def test_morans_i_end_to_end_with_grouped_lag_and_aligned_rows(monkeypatch):
    from spatialcore_b200 import AnnDataLite, spatial

    coords, X = inputs.g0_continuous()

    def run():
        a = AnnDataLite(X, obsm={"spatial": coords})
        spatial.morans_i(a, n_neighbors=6, n_permutations=99, seed=0, perm_source="replay")
        return a.uns["morans_i"]

    base = run()
    for env in ({"SC_LAG_GROUP": "4"}, {"SC_ROW_ALIGN": "32"}, {"SC_LAG_GROUP": "8", "SC_ROW_ALIGN": "32"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        got = run()
        for k in env:
            monkeypatch.delenv(k)
        np.testing.assert_allclose(got["I"].to_numpy(), base["I"].to_numpy(), rtol=1e-5, atol=1e-7)
        assert np.array_equal(got["p_value"].to_numpy(), base["p_value"].to_numpy()), env
    t = R.morans_i_table(coords, X, k=6, n_perms=99, seed=0)
    np.testing.assert_allclose(base["I"].to_numpy(), t["I"], rtol=1e-5, atol=1e-7)
