"""Seeded input recipes shared by make_golden.py (reference run) and the parity tests.
numpy 2.3 Generator streams; the recipes are tiny so they are regenerated, not stored."""

import numpy as np

NBHD_RADIUS = 28.0


def g0():
    """SURVEY.md Appendix B recipe G0: 10k uniform cells x 50 Poisson(1) genes."""
    rng = np.random.default_rng(0)
    coords = rng.uniform(0, 1000, (10000, 2))
    X = rng.poisson(1.0, (10000, 50)).astype(np.float32)
    return coords, X


def g0_continuous():
    """G0 coordinates with continuous expression (log1p counts + Gaussian noise): no exact
    ties, so per-cell permutation p-values do not hinge on FP32 round-off."""
    rng = np.random.default_rng(0)
    coords = rng.uniform(0, 1000, (10000, 2))
    rng = np.random.default_rng(100)
    X = np.log1p(rng.poisson(1.0, (10000, 8))) + 0.25 * rng.normal(size=(10000, 8))
    return coords, X


def clustered(n: int, seed: int):
    """70 % uniform + 30 % Gaussian blobs (Xenium-like density contrast)."""
    rng = np.random.default_rng(seed)
    n_blob = int(0.3 * n)
    uni = rng.uniform(0, 1000, (n - n_blob, 2))
    centers = rng.uniform(100, 900, (6, 2))
    which = rng.integers(0, 6, n_blob)
    blob = centers[which] + rng.normal(0, 12.0, (n_blob, 2))
    c = np.concatenate([uni, blob])
    return c[rng.permutation(n)]


def nbhd():
    """6000 clustered cells with 8 spatially patchy cell types."""
    coords = clustered(6000, seed=5)
    rng = np.random.default_rng(6)
    seeds = rng.uniform(0, 1000, (40, 2))
    seed_type = rng.integers(0, 8, 40)
    d = ((coords[:, None, :] - seeds[None, :, :]) ** 2).sum(-1)
    labels = seed_type[d.argmin(1)]
    noise = rng.random(6000) < 0.2
    labels = np.where(noise, rng.integers(0, 8, 6000), labels)
    return coords, labels.astype(np.int64)


def lattice(nx: int, ny: int):
    """Integer lattice: every kNN query has exact distance ties."""
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    return np.stack([xs.ravel(), ys.ravel()], axis=1)


def with_duplicates(n: int, seed: int):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, 100, (n, 2))
    dup = rng.integers(0, n, n // 10)
    c[rng.integers(0, n, n // 10)] = c[dup]
    return c


def niche_profiles(n: int = 20000, n_types: int = 12, n_niches: int = 6, seed: int = 21):
    """Row-normalised neighbourhood profiles drawn from a mixture of ``n_niches`` Dirichlet
    archetypes (k = 15 neighbours): overlapping clusters, FP32 like the reference's profiles."""
    rng = np.random.default_rng(seed)
    arche = rng.dirichlet(np.full(n_types, 0.35), n_niches)
    which = rng.integers(0, n_niches, n)
    counts = np.stack([rng.multinomial(15, arche[w]) for w in which]).astype(np.float32)
    return counts / counts.sum(1, keepdims=True)


def domains(n: int = 6000, seed: int = 31):
    """Clustered cells with two partial labelings: ``src`` (3 "Bcell" domains around blob centres,
    most cells unlabelled) and ``tgt`` (4 "Tumor" domains by quadrant, 30 % unlabelled)."""
    coords = clustered(n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    centres = rng.uniform(150, 850, (3, 2))
    d = np.sqrt(((coords[:, None, :] - centres[None]) ** 2).sum(-1))
    src = np.full(n, None, dtype=object)
    for j in range(3):
        src[d[:, j] < 70.0] = f"Bcell_{j + 1}"
    quad = (coords[:, 0] > 500).astype(int) + 2 * (coords[:, 1] > 500).astype(int)
    tgt = np.array([f"Tumor_{q + 1}" for q in quad], dtype=object)
    tgt[rng.random(n) < 0.3] = None
    return coords, src, tgt


def torus_rook(nx: int, ny: int):
    """Rook (4-neighbour) adjacency of an nx x ny torus, binary CSR, plus lattice coordinates: a graph on
    which Moran's I has closed forms (Cliff & Ord): checkerboard -> -1, cos(2 pi i / nx) -> (1 + cos(2 pi / nx)) / 2,
    and for the row-standardised graph s0 = N, s1 = N/2, s2 = 4N."""
    from scipy import sparse

    idx = np.arange(nx * ny).reshape(ny, nx)
    rows = np.repeat(idx.ravel(), 4)
    cols = np.stack([np.roll(idx, 1, 0), np.roll(idx, -1, 0), np.roll(idx, 1, 1), np.roll(idx, -1, 1)], axis=-1).ravel()
    A = sparse.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(nx * ny, nx * ny))
    A.sort_indices()
    xs, ys = np.meshgrid(np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64))
    return A, np.stack([xs.ravel(), ys.ravel()], axis=1)
