"""Full-size property tests (BASELINE.json configurations C2 / C4 / C5): sizes the CPU oracle cannot
finish, checked through size-independent properties of the domain instead:

* exact kNN / radius rows against a brute-force FP64 evaluation on a random SAMPLE of queries;
* radius graph symmetry, canonical CSR form;
* standardisation: every column has mean 0 and population variance 1;
* W is row-standardised: the lag of a constant vector is that constant, wherever a row is non-empty;
* the identity permutation reproduces the observed statistic; a device (Philox) permutation run equals
  the replay of its host mirror bit for bit; linearity of the null in Z;
* a checksum of checksums: the sum over all cells of a permuted column equals the unpermuted sum.
"""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spatialcore_b200 import engine

    return engine


def _brute_knn(cd: torch.Tensor, q: torch.Tensor, k: int):
    """FP64 brute force for query rows ``q``: ranking key (d2, index), self excluded, no FMA."""
    x, y = cd[:, 0], cd[:, 1]
    out = []
    for qi in q.tolist():
        dx, dy = x - x[qi], y - y[qi]
        d2 = dx * dx + dy * dy  # torch evaluates mul and add separately in FP64
        d2[qi] = float("inf")
        # stable ordering by (d2, index): topk on d2 then fix ties by index
        vals, idx = torch.topk(d2, k + 8, largest=False, sorted=True)
        order = np.lexsort((idx.cpu().numpy(), vals.cpu().numpy()))[:k]
        out.append(np.sort(idx.cpu().numpy()[order]))
    return np.stack(out)


@pytest.mark.parametrize("n,k,gen", [(5_000_000, 15, "uniform"), (2_000_000, 30, "mixture"), (500_000, 6, "mixture")])
def test_knn_full_size_sampled_against_brute_force(eng, n, k, gen):
    from spatialcore_b200 import synthetic

    c = synthetic.coords_uniform(n, 1.2e5, 3) if gen == "uniform" else synthetic.coords_mixture(n, 2e4, 4)
    cd = torch.from_numpy(c).cuda()
    graph, _, _ = eng.knn_graph(cd, k, want_dist=True)
    idx = graph.indices
    assert idx.shape == (n, k)
    assert bool((idx[:, 1:] > idx[:, :-1]).all()), "rows must be strictly column-sorted"
    assert not bool((idx == torch.arange(n, device="cuda", dtype=torch.int32)[:, None]).any()), "self must be excluded"
    rng = np.random.default_rng(0)
    q = torch.from_numpy(rng.choice(n, 300, replace=False))
    want = _brute_knn(cd, q, k)
    got = idx[q.cuda()].cpu().numpy()
    assert np.array_equal(got, want)
    # distances are the FP64 distances of the listed neighbours
    nb = cd[idx[q.cuda()].long()]
    d = torch.sqrt(((nb - cd[q.cuda()][:, None, :]) ** 2).sum(-1))
    assert torch.allclose(graph.dist[q.cuda()], d, rtol=1e-14, atol=0)


def test_radius_graph_full_size_properties(eng):
    from spatialcore_b200 import synthetic

    n = 5_000_000
    c = synthetic.coords_uniform(n, 1.2e5, 3)
    r = synthetic.radius_for_mean_degree(n, 1.2e5, 20.0)
    cd = torch.from_numpy(c).cuda()
    g, _ = eng.radius_graph(cd, r)
    indptr, indices = g.indptr.long(), g.indices.long()
    deg = indptr[1:] - indptr[:-1]
    assert abs(g.nnz / n - 20.0) < 0.2
    rows = torch.repeat_interleave(torch.arange(n, device="cuda"), deg)
    # every edge is within r (FP64) and not a self loop
    d2 = ((cd[rows] - cd[indices]) ** 2).sum(-1)
    assert bool((d2 <= r * r).all()) and not bool((rows == indices).any())
    # symmetric: the multiset of (i, j) equals the multiset of (j, i)
    key_f = rows * n + indices
    key_b = indices * n + rows
    assert torch.equal(torch.sort(key_f).values, torch.sort(key_b).values)
    # canonical CSR: strictly increasing columns inside each row
    same_row = rows[1:] == rows[:-1]
    assert bool((indices[1:][same_row] > indices[:-1][same_row]).all())
    # sampled rows against brute force (inclusive d <= r)
    rng = np.random.default_rng(1)
    for qi in rng.choice(n, 100, replace=False).tolist():
        dq = ((cd - cd[qi]) ** 2).sum(-1)
        want = torch.nonzero(dq <= r * r).flatten()
        want = want[want != qi]
        got = indices[indptr[qi]:indptr[qi + 1]]
        assert torch.equal(got, want)


def test_moran_pipeline_full_size_properties(eng):
    """C2 (500 k x 400, kNN k=15) through the production path in spatial order."""
    from spatialcore_b200 import philox, synthetic

    n, g, k = 500_000, 400, 15
    c = synthetic.coords_mixture(n, 1e4, 1)
    cd = torch.from_numpy(c).cuda()
    X = synthetic.expression_device(c, g, seed=1000)
    graph, _, _ = eng.knn_graph(cd, k)
    co = eng.spatial_order(cd)
    assert torch.equal(torch.sort(co.order).values, torch.arange(n, device="cuda", dtype=torch.int32))
    gs = eng.relabel_graph(graph, co)
    std = eng.zscore_dense(X, rows=co.order)
    Z = std.Z[:, :g].double()
    assert float(Z.mean(0).abs().max()) < 1e-6 and float(((Z * Z).mean(0) - 1).abs().max()) < 1e-5
    # lag of a constant is the constant (row-standardised W, every kNN row non-empty)
    ones = torch.ones((n, eng.padded_ld(8)), dtype=torch.float32, device="cuda")
    _, _, lag1, _ = eng.lag_moran(gs, ones, 8)
    assert bool((lag1[:, :8] == 1.0).all())
    num, den, lag, _ = eng.lag_moran(gs, std.Z, g)
    assert torch.allclose(den, torch.full_like(den, float(n)), rtol=1e-5)
    # Moran's I does not depend on the labelling of the cells
    std_u = eng.zscore_dense(X)
    num_u, den_u, _, _ = eng.lag_moran(graph, std_u.Z, g)
    I_s, I_u = (num / den).cpu().numpy(), (num_u / den_u).cpu().numpy()
    np.testing.assert_allclose(I_s, I_u, rtol=1e-5, atol=1e-7)
    assert np.median(I_s[: g // 4]) > 0.02 and abs(I_s[g // 4:]).max() < 0.01  # smooth genes vs noise genes
    # identity permutation == observed statistic
    ident = torch.arange(n, dtype=torch.int32, device="cuda").reshape(1, -1)
    s1 = eng.perm_null_graph_rows(std.Z, lag, g, 1, perm_idx=ident)
    assert torch.allclose(s1[0], num, rtol=1e-12, atol=1e-9)
    # Philox run == replay of its host mirror (bitwise), ragged batch of 19
    P = 19
    dev = eng.perm_null_graph_rows(std.Z, lag, g, P, seed=77, perm_offset=5)
    host = np.stack([philox.permutation(77, 5 + p, n) for p in range(P)])
    rep = eng.perm_null_graph_rows(std.Z, lag, g, P, perm_idx=torch.from_numpy(host).cuda())
    assert torch.equal(dev, rep)
    # linearity in Z: scaling Z by 2 (exact in FP32) scales every simulated sum by 2 exactly
    dev2 = eng.perm_null_graph_rows(std.Z * 2.0, lag, g, P, seed=77, perm_offset=5)
    assert torch.equal(dev2, dev * 2.0)
    # null is centred: mean of the simulated I is ~ -1/(n-1) within 6 standard errors
    sims = (eng.perm_null_graph_rows(std.Z, lag, g, 64, seed=3) / den).cpu().numpy()
    se = sims.std(0, ddof=1) / np.sqrt(64)
    assert (np.abs(sims.mean(0) + 1.0 / (n - 1)) < 6 * se + 1e-9).mean() > 0.99
    # value-permuting null: checksum of checksums + identity
    sv = eng.perm_null_values(gs, std.Z, g, 1, perm_idx=ident)
    assert torch.allclose(sv[0], num, rtol=1e-6, atol=1e-3)
    perm = eng.philox_permutation(5, 0, n)
    Zp = eng.gather_rows(std.Z, perm)
    assert torch.allclose(Zp.double().sum(0), std.Z.double().sum(0), atol=1e-6)


def test_c4_scale_null_against_fp64_column_check(eng):
    """C4 geometry (5 M cells, 4 KB rows): for a handful of genes the simulated sums of the production
    kernel equal a torch FP64 evaluation of sum_i z_i * lag[pi(i)] on those columns."""
    from spatialcore_b200 import philox

    n, g = 5_000_000, 1000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    Z = torch.randn((n, g), generator=gen, device="cuda", dtype=torch.float32)
    Lg = torch.randn((n, g), generator=gen, device="cuda", dtype=torch.float32)
    P = 17
    sims = eng.perm_null_graph_rows(Z, Lg, g, P, seed=11, perm_offset=100)
    cols = [0, 1, 499, 998, 999]
    for p in (0, 7, 16):
        pi = torch.from_numpy(philox.permutation(11, 100 + p, n)).cuda().long()
        for cidx in cols:
            want = (Z[:, cidx].double() * Lg[pi, cidx].double()).sum()
            assert abs(float(sims[p, cidx]) - float(want)) <= 1e-7


def test_neighborhood_composition_full_size(eng):
    """C5: 2 M cells, 30 types, k = 30: the fused profile equals counting labels over the graph."""
    from spatialcore_b200 import synthetic

    n, T, k = 2_000_000, 30, 30
    c = synthetic.coords_mixture(n, 2e4, 4)
    lab = synthetic.patchy_labels(c, T, 5)
    cd, ld = torch.from_numpy(c).cuda(), torch.from_numpy(lab).cuda()
    _, _, prof = eng.knn_graph(cd, k, labels=ld, n_types=T, want_idx=False)
    graph, _, _ = eng.knn_graph(cd, k)
    assert bool((prof.sum(1) == k).all())
    onehot_counts = torch.zeros((n, T), dtype=torch.float32, device="cuda")
    onehot_counts.scatter_add_(1, ld[graph.indices.long()].long(), torch.ones((n, k), dtype=torch.float32, device="cuda"))
    assert torch.equal(prof, onehot_counts)
