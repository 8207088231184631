"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/sc_b200.h declares, the ctypes table matches the header, argument validation and the host
permutation export work without a GPU, and the product never imports the oracle."""

import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r"SC_API\s+[\w\s\*]+?\b(sc_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    from spatialcore_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    names = _header_functions()
    assert len(names) >= 24
    handle = C.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in sc_b200.h but not exported"
    assert sorted(names) == sorted(_lib.SIGNATURES), "ctypes table and header disagree"


def test_ctypes_arity_matches_header():
    from spatialcore_b200 import _lib

    text = open(os.path.join(ROOT, "include", "sc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", text)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes {len(args)}"


def test_version_error_channel_and_argument_validation_without_gpu():
    from spatialcore_b200 import _lib

    L = _lib.lib()
    assert L.sc_version() >= 100
    # workspace queries are pure host arithmetic
    assert L.sc_grid_knn_workspace_bytes(1_000_000, 15) > 1_000_000 * 4 * 4
    assert L.sc_perm_null_workspace_bytes(5_000_000, 1000) > 0
    assert L.sc_lee_gemm_workspace_bytes(200_000, 1000) >= 1024 * 1024 * 8
    # invalid arguments are rejected before any CUDA call, with a message
    rc = L.sc_grid_knn(None, 10, 3, 0, None, None, None, None, 0, None, None, 0, None)
    assert rc == -1 and b"null" in L.sc_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc, "sc_grid_knn")
    rc = L.sc_philox_permutation_host(1, 0, 0, None)
    assert rc == -1


def test_philox_host_export_is_a_bijection_and_matches_python_mirror():
    from spatialcore_b200 import _lib, philox

    L = _lib.lib()
    for n in (1, 2, 3, 17, 1024, 1025, 65_537, 300_000):
        out = np.empty(n, dtype=np.int32)
        assert L.sc_philox_permutation_host(42, 9, n, out.ctypes.data) == 0
        assert np.array_equal(np.sort(out), np.arange(n))
        assert np.array_equal(out, philox.permutation(42, 9, n))
    a = philox.permutation(1, 0, 5000)
    b = philox.permutation(1, 1, 5000)
    c = philox.permutation(2, 0, 5000)
    assert (a != b).mean() > 0.99 and (a != c).mean() > 0.99


def test_philox_permutations_are_statistically_uniform():
    """position uniformity (chi-square), fixed-point count ~ Poisson(1), and the
    permutation statistic Σ x_i y_π(i) has the moments of a uniform random permutation."""
    from spatialcore_b200 import philox

    n, P = 2000, 400
    perms = np.stack([philox.permutation(7, p, n) for p in range(P)])
    # where does element 0..9 land? 20 equal bins, P draws each
    for i in range(10):
        hist = np.bincount(perms[:, i] * 20 // n, minlength=20)
        chi2 = ((hist - P / 20) ** 2 / (P / 20)).sum()
        assert chi2 < 50.0, (i, chi2)  # 19 dof: P(chi2 > 50) ~ 1e-4
    fixed = (perms == np.arange(n)[None, :]).sum(1)
    assert 0.7 < fixed.mean() < 1.3
    rng = np.random.default_rng(0)
    x, y = rng.normal(size=n), rng.normal(size=n)
    stat = (x[None, :] * y[perms]).sum(1)
    mu = n * x.mean() * y.mean()
    var = (n - 1) * x.var(ddof=1) * y.var(ddof=1) * (n - 1) / n  # exact permutation variance
    assert abs(stat.mean() - mu) < 4 * np.sqrt(var / P)
    assert 0.8 < stat.var() / var < 1.25
    # consecutive permutations are uncorrelated
    r = np.corrcoef(perms[:-1, :50].ravel(), perms[1:, :50].ravel())[0, 1]
    assert abs(r) < 0.02


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "spatialcore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "moran_port" not in src and "oracle.restate" not in src, f


def test_no_cpu_fallback_when_cuda_is_missing():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from spatialcore_b200 import AnnDataLite, spatial

    a = AnnDataLite(np.zeros((30, 3), np.float32), obsm={"spatial": np.random.default_rng(0).uniform(size=(30, 2))})
    with pytest.raises(Exception) as ei:
        spatial.morans_i(a, n_permutations=0)
    assert "nvidia" in str(ei.value).lower() or "cuda" in str(ei.value).lower()
    # argument errors still surface first, exactly like the reference
    with pytest.raises(ValueError, match="n_neighbors must be >= 1"):
        spatial.morans_i(a, n_neighbors=0)


def test_argument_validation_of_the_wider_abi_without_gpu():
    """Every export rejects malformed calls before touching CUDA, with a message in sc_last_error()."""
    from spatialcore_b200 import _lib

    L = _lib.lib()
    one = C.c_void_p(16)  # a non-null, 16-byte aligned dummy address: validation must fail before any dereference
    cases = [
        (L.sc_kmeans_assign(one, 10, 4, 4, one, 0, one, None, one, one, 1 << 20, None), b"k"),
        (L.sc_kmeans_assign(one, 10, 4, 200, one, 3, one, None, one, one, 1 << 20, None), b"d"),
        (L.sc_kmeans_pp_potential(one, 10, 4, 4, one, 99, None, None, -1, one, one, 1 << 20, None), b"n_cand"),
        (L.sc_kmeans_pp_sample(one, 10, one, 0, one, one, 1 << 20, None), b"bad sizes"),
        (L.sc_cross_nn(None, 5, one, 5, one, None, one, 1 << 20, None), b"null"),
        (L.sc_pairwise_reduce(one, 0, one, 5, one, one, 1 << 20, None), b"empty"),
        (L.sc_local_moran_finish(None, 0, one, one, one, 8, None, 10, 4, 0, None, 7, 0.05, one, one, one, one, one, one, one, 1 << 20, None), b"method"),
        (L.sc_zscore_scatter(one, 10, 8, 8, one, one, one, one, one, 0, 8, None), b"n_peers"),
        (L.sc_zscore_apply(one, 5, 10, 8, 8, None, None, one, one, one, one, 8, None), b"dtype"),
        (L.sc_gather_rows(one, 8, 10, 6, one, one, 8, None), b"multiples of 4"),
        (L.sc_perm_conjugate(one, 10, 1, one, one, one, None), b"aliased"),
        (L.sc_graph_relabel(None, one, None, 10, 0, one, one, None, one, None, None, 0, None), b"indptr or k_fixed"),
        (L.sc_lee_abs_ge_accumulate(one, 4, one, 4, 8, one, 8, None), b"bad argument"),
    ]
    for rc, needle in cases:
        assert rc == -1, (rc, needle)
    # the last message belongs to the last failing call
    assert b"sc_lee_abs_ge_accumulate" in L.sc_last_error()
    assert L.sc_kmeans_workspace_bytes(2_000_000, 30, 8) > 0 and L.sc_kmeans_pp_sample_workspace_bytes(2_000_000) > 16_000_000
    assert L.sc_local_moran_finish_workspace_bytes(100, 999) >= 100 * 1000 * 8
    assert L.sc_cross_nn_workspace_bytes(1_000_000) > 1_000_000 * 16
    assert L.sc_launch_count() >= 0


def test_tile_exports_validate_without_a_gpu():
    """sc_graph_tile_bytes / _build / sc_csr_lag_moran_tiled reject bad arguments before any launch."""
    from spatialcore_b200 import _lib

    L = _lib.lib()
    assert L.sc_graph_tile_bytes(0, 10) == 0 and L.sc_graph_tile_bytes(100, -1) == 0
    big = L.sc_graph_tile_bytes(5_000_000, 100_000_000)
    # 2 bytes per edge for the word lists (16-bit union indices) + the per-chunk union rows: less than the CSR itself
    assert 250_000_000 < big < 450_000_000
    assert L.sc_graph_tile_build(None, None, 100, 6, 600, None, 0, None) == -1
    assert "null argument" in L.sc_last_error().decode()
    assert L.sc_graph_tile_build(None, 1, 100, 6, 601, 1, 1 << 20, None) == -1  # nnz != n * k_fixed
    assert L.sc_graph_tile_build(None, 1, 100, 6, 600, 1, 16, None) == -2  # tile buffer too small
    assert "too small" in L.sc_last_error().decode()
    assert L.sc_csr_lag_moran_tiled(None, 1, 100, 6, 600, 1, 1 << 20, None, 1, None, 30, 32, None, None, 0, 1, 1, None, None, 0,
                                    1, 1 << 20, None) == -1  # ldz < g
    assert L.sc_csr_lag_moran_tiled(None, 1, 100, 6, 600, 1, 1 << 20, None, 1, None, 32, 32, None, None, 0, 1, 1, 1, None, 0,
                                    1, 1 << 20, None) == -1  # cell_obs without cell_cnt


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: ``include/sc_b200.h`` must compile as C99 (no C++ in the signatures) and a
    C program must link against the shared library and call it."""
    import shutil
    import subprocess

    from spatialcore_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "use_abi.c"
    src.write_text(
        '#include "sc_b200.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  size_t ws = sc_grid_knn_workspace_bytes(1000, 6);\n"
        "  int rc = sc_philox_permutation_host(1u, 0, 0, (int32_t*)0);\n"
        '  printf("%d %d %d %s\\n", sc_version(), ws > 0, rc, sc_last_error());\n'
        "  return 0;\n}\n")
    exe = str(tmp_path / "use_abi")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe,
                    "-L", libdir, "-l:libsc_b200.so", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split(None, 3)
    assert int(out[0]) >= 100 and out[1] == "1" and out[2] == "-1" and "bad argument" in out[3]
