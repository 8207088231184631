"""On-disk expression matrices: a read-only ``numpy.memmap`` (dense) and a scipy
CSR / CSC whose three arrays are memory-mapped are accepted by every entry point as they are -- the upload
pages them in, no in-memory copy of X is made by the library -- and give exactly the tables of the
in-memory matrix."""

import numpy as np
import pytest
import torch
from scipy import sparse

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from spatialcore_b200 import spatial

    return spatial


def _case(n=6000, g=24, seed=11):
    rng = np.random.default_rng(seed)
    coords = rng.uniform(0, 1000, (n, 2))
    X = rng.poisson(1.0 + (coords[:, :1] / 500.0), (n, g)).astype(np.float32)
    X[:, 3] = 0.0  # a constant gene
    return coords, X


def _adata(X, coords, names):
    from spatialcore_b200 import AnnDataLite

    return AnnDataLite(X, obsm={"spatial": coords}, var_names=names)


def _mapped(tmp_path, name, arr):
    path = tmp_path / f"{name}.npy"
    np.save(path, arr)
    m = np.load(path, mmap_mode="r")
    assert isinstance(m, np.memmap) and not m.flags.writeable
    return m


def test_memmapped_dense_and_sparse_matrices_give_the_in_memory_tables(api, tmp_path):
    coords, X = _case()
    names = [f"g{i}" for i in range(X.shape[1])]
    kw = dict(n_neighbors=6, n_permutations=49, seed=4, perm_source="replay")
    ref = _adata(X, coords, names)
    api.morans_i(ref, **kw)
    want = ref.uns["morans_i"]

    dense = _adata(_mapped(tmp_path, "dense", X), coords, names)
    api.morans_i(dense, **kw)
    assert dense.uns["morans_i"].equals(want)

    for fmt in ("csr", "csc"):
        S = sparse.csr_matrix(X) if fmt == "csr" else sparse.csc_matrix(X)
        cls = sparse.csr_matrix if fmt == "csr" else sparse.csc_matrix
        Sm = cls((_mapped(tmp_path, fmt + "_data", S.data), _mapped(tmp_path, fmt + "_indices", S.indices),
                  _mapped(tmp_path, fmt + "_indptr", S.indptr)), shape=S.shape, copy=False)
        assert not Sm.data.flags.writeable  # still the read-only mapping, not a copy
        a = _adata(Sm, coords, names)
        api.morans_i(a, **kw)
        got = a.uns["morans_i"]
        assert got["gene"].tolist() == want["gene"].tolist()
        np.testing.assert_allclose(got["I"].to_numpy(), want["I"].to_numpy(), rtol=1e-6, atol=1e-9)
        assert np.array_equal(got["p_value"].to_numpy(), want["p_value"].to_numpy())
        # a gene subset straight from the mapped matrix
        sub = _adata(Sm, coords, names)
        api.morans_i(sub, genes=["g5", "g0", "g17"], **kw)
        w = want.set_index("gene").loc[["g5", "g0", "g17"]]
        np.testing.assert_allclose(sub.uns["morans_i"]["I"].to_numpy(), w["I"].to_numpy(), rtol=1e-6, atol=1e-9)

    # local Moran and the all-pairs Lee matrix read the mapped matrix as well
    la, lb = _adata(X, coords, names), _adata(_mapped(tmp_path, "dense2", X), coords, names)
    api.local_morans_i(la, genes=names[:6], n_permutations=19, seed=1)
    api.local_morans_i(lb, genes=names[:6], n_permutations=19, seed=1)
    for key in la.obsm:
        if key != "spatial":
            assert np.array_equal(np.asarray(la.obsm[key]), np.asarray(lb.obsm[key]), equal_nan=True), key
