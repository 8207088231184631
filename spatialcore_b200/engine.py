"""Device-side building blocks: thin wrappers that pass torch CUDA tensors to ``libsc_b200.so``.

PyTorch is plumbing here (device memory, streams, ``torch.distributed``); every numeric step is a
hand-written sm_100a kernel behind the C ABI (``include/sc_b200.h``).  No function in this module
has a CPU path: inputs must live on a CUDA device.
"""

from __future__ import annotations

import functools
import os
import warnings
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from spatialcore_b200 import _lib
from spatialcore_b200._lib import SC_F32, SC_F64, SC_PERM_PHILOX, SC_PERM_REPLAY, check


def launches() -> int:
    """Kernels launched by ``libsc_b200.so`` in this process so far (``sc_launch_count``: incremented at
    every launch site of the library; CUB sorts / scans it calls are not counted)."""
    return int(_lib.lib().sc_launch_count())


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _target_device(args, kwargs) -> Optional[torch.device]:
    """The CUDA device a call works on: its first CUDA tensor / device graph / cell order, else an
    explicitly indexed ``device=`` argument."""
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor):
            if a.is_cuda:
                return a.device
        elif isinstance(a, (DeviceGraph, CellOrder, Standardized)):
            t = a.indices if isinstance(a, DeviceGraph) else (a.order if isinstance(a, CellOrder) else a.Z)
            if t is not None and t.is_cuda:
                return t.device
        elif isinstance(a, KMeansDevice) and getattr(a, "X", None) is not None:
            return a.X.device
    dev = kwargs.get("device")
    if dev is not None:
        dev = torch.device(dev)
        if dev.type == "cuda" and dev.index is not None:
            return dev
    return None


def _on_device(fn):
    """Run ``fn`` with the CUDA device of its operands current.  The library launches on the CURRENT
    device (``cudaGetDevice`` for the SM count and function attributes, the stream handed over is the
    current stream), so a call on tensors of ``cuda:1`` while ``cuda:0`` is current must switch first."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _target_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"spatialcore_b200: {name} must be a CUDA tensor (there is no CPU path)")


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def padded_ld(g: int) -> int:
    """Leading dimension used for every N x G matrix: multiple of 8 floats (32-byte sectors)."""
    return (g + 7) // 8 * 8


@dataclass
class DeviceGraph:
    """CSR neighbour graph on the device.  ``indptr is None`` means every row has ``k_fixed``
    entries (kNN).  ``weights is None`` means row-standardised binary weights ``1/deg_i``."""

    n: int
    indices: torch.Tensor
    indptr: Optional[torch.Tensor] = None
    k_fixed: int = 0
    weights: Optional[torch.Tensor] = None
    dist: Optional[torch.Tensor] = None
    # shared-memory tile form for the lag kernel (opaque device buffer); see tile_graph()
    tiles: Optional[torch.Tensor] = None

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())

    def indptr_tensor(self) -> torch.Tensor:
        if self.indptr is not None:
            return self.indptr
        return torch.arange(0, (self.n + 1) * self.k_fixed, self.k_fixed, dtype=torch.int32, device=self.indices.device)

    def to_scipy(self, data: str = "weights", dtype=np.float32):
        """Host scipy CSR: ``data`` in {"weights" (row-standardised), "ones", "dist"}."""
        from scipy import sparse

        indptr = self.indptr_tensor().cpu().numpy()
        indices = self.indices.reshape(-1).cpu().numpy()
        if data == "dist":
            vals = self.dist.reshape(-1).cpu().numpy().astype(dtype)
        elif data == "ones":
            vals = np.ones(indices.size, dtype=dtype)
        elif self.weights is not None:
            vals = self.weights.reshape(-1).cpu().numpy().astype(dtype)
        else:
            deg = np.diff(indptr)
            with np.errstate(divide="ignore"):
                vals = np.repeat((np.ones(1, dtype) / np.maximum(deg, 1).astype(dtype)).astype(dtype), deg)
        return sparse.csr_matrix((vals, indices, indptr), shape=(self.n, self.n))


# --------------------------------------------------------------------------------------------------
# graphs
# --------------------------------------------------------------------------------------------------


def check_planar(extra_is_constant: bool, shape) -> None:
    """The kernels are 2-D.  The reference hands the whole ``obsm['spatial']`` array to its tree
    libraries [R autocorrelation.py:393-395, neighborhoods.py:213-223], so coordinates with a third
    column that VARIES would give different neighbours there: refuse them instead of silently dropping
    the column.  Constant extra columns (e.g. z = 0) do not change any distance and are accepted."""
    if not extra_is_constant:
        raise ValueError(
            f"spatial coordinates have shape {tuple(shape)}: only 2-D coordinates are supported "
            "(columns beyond the first two must be constant); project or slice obsm['spatial'] to (n_cells, 2)")


def _coords_tensor(coords, device) -> torch.Tensor:
    if isinstance(coords, torch.Tensor):
        if coords.dim() != 2 or coords.shape[1] < 2:
            raise ValueError(f"spatial coordinates must have shape (n_cells, >=2), got {tuple(coords.shape)}")
        if coords.shape[1] > 2 and coords.shape[0] > 0:
            check_planar(bool((coords[:, 2:] == coords[:1, 2:]).all().item()), coords.shape)
        c = coords.to(device=device, dtype=torch.float64)
    else:
        a = np.asarray(coords)
        if a.ndim != 2 or a.shape[1] < 2:
            raise ValueError(f"spatial coordinates must have shape (n_cells, >=2), got {a.shape}")
        if not np.isfinite(a[:, :2]).all():
            raise ValueError("spatial coordinates contain NaN or infinite values")
        if a.shape[1] > 2 and a.shape[0] > 0:
            check_planar(bool((a[:, 2:] == a[:1, 2:]).all()), a.shape)
        c = torch.from_numpy(np.ascontiguousarray(a[:, :2], dtype=np.float64)).to(device)
    return c[:, :2].contiguous()


@_on_device
def knn_graph(
    coords,
    k: int,
    include_self: bool = False,
    want_dist: bool = False,
    want_order: bool = False,
    labels: Optional[torch.Tensor] = None,
    n_types: int = 0,
    want_idx: bool = True,
    device="cuda",
):
    """Exact kNN graph (``sc_grid_knn``).  Returns ``(DeviceGraph|None, order|None, profile|None)``."""
    c = _coords_tensor(coords, device)
    n = c.shape[0]
    if k < 1:
        raise ValueError(f"n_neighbors must be >= 1, got {k}")
    if k >= n:
        raise ValueError(f"k must be < number of cells ({n}), got {k}")
    if k > _lib.SC_KNN_MAX_K:
        raise ValueError(f"k={k} exceeds the compiled limit SC_KNN_MAX_K={_lib.SC_KNN_MAX_K}")
    L = _lib.lib()
    kk = k + (1 if include_self else 0)
    idx = torch.empty((n, kk), dtype=torch.int32, device=c.device) if want_idx else None
    dist = torch.empty((n, kk), dtype=torch.float64, device=c.device) if (want_dist and want_idx) else None
    order = torch.empty(n, dtype=torch.int32, device=c.device) if want_order else None
    profile = None
    if labels is not None:
        _require_cuda(labels, "labels")
        profile = torch.empty((n, n_types), dtype=torch.float32, device=c.device)
    ws = _workspace(L.sc_grid_knn_workspace_bytes(n, k), c.device)
    check(
        L.sc_grid_knn(_ptr(c), n, k, int(include_self), _ptr(idx), _ptr(dist), _ptr(order), _ptr(labels),
                      int(n_types), _ptr(profile), _ptr(ws), ws.numel(), _stream()),
        "sc_grid_knn",
    )
    graph = DeviceGraph(n=n, indices=idx, k_fixed=kk, dist=dist) if want_idx else None
    return graph, order, profile


@_on_device
def radius_graph(
    coords,
    radius: float,
    want_dist: bool = False,
    labels: Optional[torch.Tensor] = None,
    n_types: int = 0,
    want_graph: bool = True,
    device="cuda",
):
    """Radius graph, inclusive ``d <= r``, self excluded (``sc_grid_radius_count`` / ``_fill``).
    Returns ``(DeviceGraph|None, profile|None)``."""
    if radius is None or not radius > 0:
        raise ValueError(f"radius must be > 0, got {radius}")
    c = _coords_tensor(coords, device)
    n = c.shape[0]
    L = _lib.lib()
    ws = _workspace(L.sc_grid_radius_workspace_bytes(n), c.device)
    indptr = torch.empty(n + 1, dtype=torch.int32, device=c.device) if want_graph else None
    nnz_t = torch.zeros(1, dtype=torch.int64, device=c.device)
    profile = None
    if labels is not None:
        _require_cuda(labels, "labels")
        profile = torch.empty((n, n_types), dtype=torch.float32, device=c.device)
    check(
        L.sc_grid_radius_count(_ptr(c), n, float(radius), _ptr(indptr), _ptr(nnz_t), _ptr(labels), int(n_types),
                               _ptr(profile), _ptr(ws), ws.numel(), _stream()),
        "sc_grid_radius_count",
    )
    if not want_graph:
        return None, profile
    nnz = int(nnz_t.item())  # host sync: the caller-owned index buffer must be sized
    if nnz >= 2**31 - 1:
        raise ValueError(f"radius graph has {nnz} edges, beyond int32 CSR capacity; reduce the radius")
    indices = torch.empty(nnz, dtype=torch.int32, device=c.device)
    dist = torch.empty(nnz, dtype=torch.float64, device=c.device) if want_dist else None
    scratch = _workspace(nnz * (12 if want_dist else 4), c.device)
    if nnz > 0:
        check(
            L.sc_grid_radius_fill(_ptr(c), n, float(radius), _ptr(indptr), _ptr(indices), _ptr(dist), _ptr(scratch),
                                  nnz * (12 if want_dist else 4), _ptr(ws), ws.numel(), _stream()),
            "sc_grid_radius_fill",
        )
    return DeviceGraph(n=n, indices=indices, indptr=indptr, dist=dist), profile


@dataclass
class CellOrder:
    """A spatially compact (Z-order) relabelling of the cells.  ``order[a]`` = original id of the cell
    at sorted position ``a``; ``rank`` is the inverse.  Every statistic on this path is a sum over
    cells, so Z / lag / the graph may be held in this order; a row's neighbours are then close in
    memory and the lag kernel reads Z about once instead of once per edge."""

    order: torch.Tensor  # int32 [n]
    rank: torch.Tensor  # int32 [n]


@_on_device
def spatial_order(coords, device="cuda") -> CellOrder:
    """``sc_spatial_order``: Z-order curve over a uniform grid of the coordinates."""
    c = _coords_tensor(coords, device)
    n = c.shape[0]
    L = _lib.lib()
    order = torch.empty(n, dtype=torch.int32, device=c.device)
    rank = torch.empty(n, dtype=torch.int32, device=c.device)
    ws = _workspace(L.sc_spatial_order_workspace_bytes(n), c.device)
    check(L.sc_spatial_order(_ptr(c), n, _ptr(order), _ptr(rank), _ptr(ws), ws.numel(), _stream()), "sc_spatial_order")
    return CellOrder(order=order, rank=rank)


@_on_device
def relabel_graph(graph: DeviceGraph, co: CellOrder, tiles: bool = True) -> DeviceGraph:
    """``sc_graph_relabel``: the same graph on sorted positions (rows permuted, columns mapped and
    re-sorted; weights follow their edges).  ``tiles``: also build the shared-memory tile form of the lag
    kernel (``tile_graph``; row-standardised binary graphs only)."""
    L = _lib.lib()
    dev = graph.indices.device
    out_idx = torch.empty_like(graph.indices)
    out_ptr = torch.empty_like(graph.indptr) if graph.indptr is not None else None
    out_w = torch.empty_like(graph.weights) if graph.weights is not None else None
    ws = _workspace(L.sc_graph_relabel_workspace_bytes(graph.n), dev)
    check(
        L.sc_graph_relabel(_ptr(graph.indptr), _ptr(graph.indices), _ptr(graph.weights), graph.n, int(graph.k_fixed),
                           _ptr(co.order), _ptr(co.rank), _ptr(out_ptr), _ptr(out_idx), _ptr(out_w), _ptr(ws), ws.numel(),
                           _stream()),
        "sc_graph_relabel",
    )
    out = DeviceGraph(n=graph.n, indices=out_idx, indptr=out_ptr, k_fixed=graph.k_fixed, weights=out_w)
    if tiles and out_w is None and tile_rows() > 0:
        tile_graph(out)
    return out


def tile_rows() -> int:
    """1 when graphs put into spatial order also get the shared-memory tile form of the lag kernel (the
    default); ``SC_LAG_TILE_ROWS=0`` leaves every lag to the L1-gather kernel (cross-check / comparison)."""
    v = os.environ.get("SC_LAG_TILE_ROWS", "1")
    if v not in ("0", "1"):
        raise ValueError(f"SC_LAG_TILE_ROWS must be 0 or 1, got '{v}'")
    return int(v)


@_on_device
def tile_graph(graph: DeviceGraph) -> DeviceGraph:
    """``sc_graph_tile_build``: the shared-memory tile form of a row-standardised binary graph in spatial
    order (per chunk of 256 rows the union of neighbour rows, per CSR entry the byte offset of its row inside
    the tile).  ``lag_moran`` and ``perm_null_values`` use it when present."""
    if graph.weights is not None:
        raise ValueError("tile_graph: explicitly weighted graphs are not supported")
    L = _lib.lib()
    nbytes = int(L.sc_graph_tile_bytes(graph.n, graph.nnz))
    if nbytes == 0 or graph.nnz == 0:
        graph.tiles = None
        return graph
    buf = torch.empty(nbytes, dtype=torch.uint8, device=graph.indices.device)
    check(
        L.sc_graph_tile_build(_ptr(graph.indptr), _ptr(graph.indices), graph.n, int(graph.k_fixed), graph.nnz,
                              _ptr(buf), nbytes, _stream()),
        "sc_graph_tile_build",
    )
    graph.tiles = buf
    return graph


@_on_device
def gather_rows(src: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """``sc_gather_rows``: ``dst[a] = src[rows[a]]`` for a float32 [n, ld] matrix (ld % 4 == 0)."""
    _require_cuda(src, "src")
    L = _lib.lib()
    n, ld = src.shape
    dst = torch.empty_like(src)
    check(L.sc_gather_rows(_ptr(src), src.stride(0), n, ld, _ptr(rows), _ptr(dst), ld, _stream()), "sc_gather_rows")
    return dst


@_on_device
def conjugate_perms(perm_idx: torch.Tensor, co: CellOrder) -> torch.Tensor:
    """``sc_perm_conjugate``: replayed permutations of cell ids -> permutations of sorted positions."""
    L = _lib.lib()
    P, n = perm_idx.shape
    out = torch.empty_like(perm_idx)
    check(L.sc_perm_conjugate(_ptr(perm_idx), n, P, _ptr(co.order), _ptr(co.rank), _ptr(out), _stream()), "sc_perm_conjugate")
    return out


@_on_device
def graph_from_scipy(adj, device="cuda", use_weights: bool = False) -> DeviceGraph:
    """Upload an existing scipy CSR connectivity matrix (``use_existing_graph`` path)."""
    from scipy import sparse

    a = sparse.csr_matrix(adj)
    a.sort_indices()
    a.sum_duplicates()
    n = a.shape[0]
    indptr = torch.from_numpy(a.indptr.astype(np.int32)).to(device)
    indices = torch.from_numpy(a.indices.astype(np.int32)).to(device)
    weights = torch.from_numpy(a.data.astype(np.float32)).to(device) if use_weights else None
    return DeviceGraph(n=n, indices=indices, indptr=indptr, weights=weights)


@_on_device
def nbhd_counts(graph: DeviceGraph, labels: torch.Tensor, n_types: int) -> torch.Tensor:
    L = _lib.lib()
    prof = torch.empty((graph.n, n_types), dtype=torch.float32, device=labels.device)
    check(
        L.sc_nbhd_counts(_ptr(graph.indptr), _ptr(graph.indices), graph.n, int(graph.k_fixed), _ptr(labels),
                         int(n_types), _ptr(prof), _stream()),
        "sc_nbhd_counts",
    )
    return prof


@_on_device
def profile_normalize(profile: torch.Tensor, normalize: bool) -> int:
    """Normalises in place; returns the number of empty rows (host sync)."""
    L = _lib.lib()
    n_empty = torch.zeros(1, dtype=torch.int64, device=profile.device)
    check(
        L.sc_profile_normalize(_ptr(profile), profile.shape[0], profile.shape[1], int(normalize), _ptr(n_empty), _stream()),
        "sc_profile_normalize",
    )
    return int(n_empty.item())


@_on_device
def graph_moments(graph: DeviceGraph) -> Tuple[float, float, float]:
    L = _lib.lib()
    dev = graph.indices.device
    out = torch.empty(3, dtype=torch.float64, device=dev)
    ws = _workspace(L.sc_graph_moments_workspace_bytes(graph.n), dev)
    check(
        L.sc_graph_moments(_ptr(graph.indptr), _ptr(graph.indices), _ptr(graph.weights), graph.n, int(graph.k_fixed),
                           _ptr(out), _ptr(ws), ws.numel(), _stream()),
        "sc_graph_moments",
    )
    s = out.cpu().numpy()
    return float(s[0]), float(s[1]), float(s[2])


# --------------------------------------------------------------------------------------------------
# standardisation
# --------------------------------------------------------------------------------------------------


@dataclass
class Standardized:
    Z: torch.Tensor  # [n, ld] float32, padding columns zero
    g: int
    mean: torch.Tensor  # [g] float64
    std: torch.Tensor  # [g] float64
    zero_var: torch.Tensor  # [g] uint8


@_on_device
def zscore_dense(X: torch.Tensor, cols: Optional[torch.Tensor] = None, rows: Optional[torch.Tensor] = None,
                 want_z: bool = True) -> Standardized:
    """``sc_zscore`` on a dense device matrix (float32 or float64, row-major, any row stride)."""
    _require_cuda(X, "X")
    if X.dim() != 2:
        raise ValueError("X must be 2-D")
    if X.dtype not in (torch.float32, torch.float64):
        X = X.to(torch.float32)
    if X.stride(1) != 1:
        X = X.contiguous()
    n = X.shape[0]
    g = int(cols.numel()) if cols is not None else X.shape[1]
    ld = padded_ld(g)
    L = _lib.lib()
    dev = X.device
    Z = torch.empty((n, ld), dtype=torch.float32, device=dev) if want_z else None
    mean = torch.empty(g, dtype=torch.float64, device=dev)
    std = torch.empty(g, dtype=torch.float64, device=dev)
    zero = torch.empty(g, dtype=torch.uint8, device=dev)
    ws = _workspace(L.sc_zscore_workspace_bytes(n, g), dev)
    check(
        L.sc_zscore(_ptr(X), SC_F32 if X.dtype == torch.float32 else SC_F64, n, X.stride(0), g, _ptr(cols), _ptr(rows),
                    _ptr(Z), ld, _ptr(mean), _ptr(std), _ptr(zero), _ptr(ws), ws.numel(), _stream()),
        "sc_zscore",
    )
    return Standardized(Z=Z, g=g, mean=mean, std=std, zero_var=zero)


@_on_device
def zscore_apply(X: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, zero_var: torch.Tensor,
                 cols: Optional[torch.Tensor] = None, rows: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sc_zscore_apply``: z-score ``X`` with GIVEN per-gene moments (row-sharded ingest: the moments
    were combined over the ranks).  ``out`` may be a row slice of a larger [*, ld] buffer."""
    _require_cuda(X, "X")
    if X.dtype not in (torch.float32, torch.float64):
        X = X.to(torch.float32)
    if X.stride(1) != 1:
        X = X.contiguous()
    n = int(rows.numel()) if rows is not None else X.shape[0]
    g = int(cols.numel()) if cols is not None else X.shape[1]
    ld = padded_ld(g)
    L = _lib.lib()
    if out is None:
        out = torch.empty((n, ld), dtype=torch.float32, device=X.device)
    if out.shape != (n, ld) or out.stride(1) != 1 or out.stride(0) != ld or out.dtype != torch.float32:
        raise ValueError("zscore_apply: out must be a float32 [n, padded_ld(g)] row-contiguous tensor")
    check(
        L.sc_zscore_apply(_ptr(X), SC_F32 if X.dtype == torch.float32 else SC_F64, n, X.stride(0), g, _ptr(cols),
                          _ptr(rows), _ptr(mean), _ptr(std), _ptr(zero_var), _ptr(out), ld, _stream()),
        "sc_zscore_apply",
    )
    return out


@_on_device
def zscore_scatter(X: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, zero_var: torch.Tensor,
                   dst_rows: torch.Tensor, peer_ptrs, ld: int) -> None:
    """``sc_zscore_scatter``: z-score this rank's row block and store each output row at
    ``dst_rows[a]`` of the Z matrix of every peer (``peer_ptrs``: peer-mapped base addresses of the
    symmetric Z buffers).  One kernel: standardise + all-gather + spatial re-order over NVLink."""
    import ctypes as C

    _require_cuda(X, "X")
    L = _lib.lib()
    n, g = X.shape
    arr = (C.c_uint64 * len(peer_ptrs))(*[int(q) for q in peer_ptrs])
    check(
        L.sc_zscore_scatter(_ptr(X), n, X.stride(0), g, _ptr(dst_rows), _ptr(mean), _ptr(std), _ptr(zero_var),
                            C.cast(arr, C.c_void_p), len(peer_ptrs), int(ld), _stream()),
        "sc_zscore_scatter",
    )


@_on_device
def densify_csr(indptr: torch.Tensor, indices: torch.Tensor, data: torch.Tensor, n: int, n_cols: int,
                colmap: Optional[torch.Tensor], g_out: int) -> torch.Tensor:
    """``sc_csr_densify``: CSR expression -> dense float32 [n, padded_ld(g_out)] on the device."""
    L = _lib.lib()
    ld = padded_ld(g_out)
    out = torch.empty((n, ld), dtype=torch.float32, device=data.device)
    if data.dtype not in (torch.float32, torch.float64):
        data = data.to(torch.float32)
    check(
        L.sc_csr_densify(_ptr(indptr), _ptr(indices), _ptr(data), SC_F32 if data.dtype == torch.float32 else SC_F64, n,
                         _ptr(colmap), g_out, _ptr(out), ld, _stream()),
        "sc_csr_densify",
    )
    return out


@_on_device
def expression_to_device(X, gene_idx: Optional[np.ndarray], device="cuda") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Bring an AnnData expression matrix to the device as a dense matrix plus an optional column
    selector.  Accepts numpy, scipy sparse (CSR/CSC/…), or torch tensors (host or device)."""
    from scipy import sparse

    cols = None
    if sparse.issparse(X):
        csr = X.tocsr()
        if not csr.has_canonical_format:
            csr = csr.copy()
            csr.sum_duplicates()
        n, n_cols = csr.shape
        if gene_idx is None:
            g_out, colmap = n_cols, None
        else:
            g_out = len(gene_idx)
            if len(np.unique(gene_idx)) != g_out:
                # duplicated genes: densify all requested source columns once, select afterwards
                uniq, inv = np.unique(gene_idx, return_inverse=True)
                cm = np.full(n_cols, -1, dtype=np.int32)
                cm[uniq] = np.arange(len(uniq), dtype=np.int32)
                dense = densify_csr(
                    torch.from_numpy(csr.indptr.astype(np.int64)).to(device),
                    torch.from_numpy(csr.indices.astype(np.int32)).to(device),
                    torch.from_numpy(csr.data).to(device), n, n_cols, torch.from_numpy(cm).to(device), len(uniq))
                return dense, torch.from_numpy(inv.astype(np.int32)).to(device)
            cm = np.full(n_cols, -1, dtype=np.int32)
            cm[gene_idx] = np.arange(g_out, dtype=np.int32)
            colmap = torch.from_numpy(cm).to(device)
        dense = densify_csr(
            torch.from_numpy(csr.indptr.astype(np.int64)).to(device),
            torch.from_numpy(csr.indices.astype(np.int32)).to(device),
            torch.from_numpy(csr.data).to(device), n, n_cols, colmap, g_out)
        return dense[:, :g_out], None
    if isinstance(X, torch.Tensor):
        Xd = X.to(device, non_blocking=True)
    else:
        a = np.asarray(X)  # a read-only numpy.memmap (on-disk X) stays a mapping: the upload pages it in
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", message="The given NumPy array is not writable")  # it is only read
            Xd = torch.from_numpy(a).to(device, non_blocking=True)
    if gene_idx is not None:
        cols = torch.from_numpy(np.asarray(gene_idx, dtype=np.int32)).to(device)
    return Xd, cols


# --------------------------------------------------------------------------------------------------
# lag, nulls, contraction
# --------------------------------------------------------------------------------------------------


def _use_tiles(graph: DeviceGraph, ld: int) -> bool:
    return graph.tiles is not None and graph.weights is None and ld >= 32


def _lag_tiled(graph: DeviceGraph, Z: torch.Tensor, Zself: Optional[torch.Tensor], perm: Optional[torch.Tensor], g: int,
               lag: Optional[torch.Tensor], local: Optional[torch.Tensor], num: torch.Tensor, den: torch.Tensor,
               cell_obs: Optional[torch.Tensor], cell_cnt: Optional[torch.Tensor], ws: torch.Tensor) -> None:
    """``sc_csr_lag_moran_tiled`` on the graph's tile form."""
    L = _lib.lib()
    n, ld = Z.shape
    buf = graph.tiles
    ldl = lag.shape[1] if lag is not None else (local.shape[1] if local is not None else ld)
    check(
        L.sc_csr_lag_moran_tiled(_ptr(graph.indptr), _ptr(graph.indices), n, int(graph.k_fixed), graph.nnz, _ptr(buf),
                                 buf.numel(), _ptr(Zself), _ptr(Z), _ptr(perm), ld, g, _ptr(lag), _ptr(local), ldl, _ptr(num),
                                 _ptr(den), _ptr(cell_obs), _ptr(cell_cnt), cell_cnt.shape[1] if cell_cnt is not None else 0,
                                 _ptr(ws), ws.numel(), _stream()),
        "sc_csr_lag_moran_tiled",
    )


@_on_device
def lag_moran(graph: DeviceGraph, Z: torch.Tensor, g: int, want_lag: bool = True, want_local: bool = False,
              perm: Optional[torch.Tensor] = None):
    """``sc_csr_lag_moran`` / ``sc_csr_lag_moran_tiled``: returns ``(num[g], den[g], lag|None, local|None)``.
    ``perm`` (int32 [n]): the operand is ``Z[perm]`` -- applied while staging on the tile path, through a permuted
    copy otherwise."""
    _require_cuda(Z, "Z")
    if perm is not None and not _use_tiles(graph, Z.shape[1]):
        Z, perm = gather_rows(Z, perm), None
    L = _lib.lib()
    n, ld = Z.shape
    dev = Z.device
    lag = torch.empty((n, ld), dtype=torch.float32, device=dev) if want_lag else None
    local = torch.empty((n, ld), dtype=torch.float32, device=dev) if want_local else None
    num = torch.empty(g, dtype=torch.float64, device=dev)
    den = torch.empty(g, dtype=torch.float64, device=dev)
    ws = _workspace(L.sc_csr_lag_moran_workspace_bytes(n, g), dev)
    if _use_tiles(graph, ld):
        _lag_tiled(graph, Z, None, perm, g, lag, local, num, den, None, None, ws)
        return num, den, lag, local
    check(
        L.sc_csr_lag_moran(_ptr(graph.indptr), _ptr(graph.indices), _ptr(graph.weights), n, int(graph.k_fixed), _ptr(Z),
                           ld, g, _ptr(lag), _ptr(local), ld, _ptr(num), _ptr(den), _ptr(ws), ws.numel(), _stream()),
        "sc_csr_lag_moran",
    )
    return num, den, lag, local


def _perm_source(perm_idx: Optional[torch.Tensor], n: int, n_perms: int):
    if perm_idx is None:
        return SC_PERM_PHILOX, None
    _require_cuda(perm_idx, "perm_idx")
    if perm_idx.dtype != torch.int32 or perm_idx.shape != (n_perms, n) or not perm_idx.is_contiguous():
        raise ValueError("perm_idx must be a contiguous int32 tensor of shape (n_perms, n)")
    return SC_PERM_REPLAY, perm_idx


@_on_device
def perm_null_graph_rows(A: torch.Tensor, B: torch.Tensor, g: int, n_perms: int,
                         perm_idx: Optional[torch.Tensor] = None, seed: int = 0, perm_offset: int = 0,
                         out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sc_perm_null_graph_rows``: ``sims[p,c] = Σ_i A[i,c]·B[π_p(i),c]`` (raw sums, float64)."""
    L = _lib.lib()
    n, ld = A.shape
    source, pidx = _perm_source(perm_idx, n, n_perms)
    sims = out if out is not None else torch.empty((n_perms, g), dtype=torch.float64, device=A.device)
    if ws is None:
        ws = _workspace(L.sc_perm_null_workspace_bytes(n, g), A.device)
    check(
        L.sc_perm_null_graph_rows(_ptr(A), ld, _ptr(B), B.shape[1], n, g, source, _ptr(pidx), int(seed) & (2**64 - 1),
                                  int(perm_offset), int(n_perms), _ptr(sims), _ptr(ws), ws.numel(), _stream()),
        "sc_perm_null_graph_rows",
    )
    return sims


@_on_device
def perm_null_values(graph: DeviceGraph, Zy: torch.Tensor, g: int, n_perms: int, Zx: Optional[torch.Tensor] = None,
                     perm_idx: Optional[torch.Tensor] = None, seed: int = 0, perm_offset: int = 0,
                     cell_obs: Optional[torch.Tensor] = None, cell_cnt: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``sc_perm_null_values``: the reference's own value-permuting null (gather-SpMM)."""
    L = _lib.lib()
    n, ld = Zy.shape
    source, pidx = _perm_source(perm_idx, n, n_perms)
    sims = torch.empty((n_perms, g), dtype=torch.float64, device=Zy.device)
    if _use_tiles(graph, ld):
        # the tile kernel applies the permutation while staging (operand row j = Zy[perm[j]]): no permuted copy
        # of Z is materialised; the statistic of permutation p is the Moran numerator of that pass
        ws = _workspace(L.sc_csr_lag_moran_workspace_bytes(n, g), Zy.device)
        den = torch.empty(g, dtype=torch.float64, device=Zy.device)
        for p in range(n_perms):
            perm = pidx[p] if pidx is not None else philox_permutation(seed, perm_offset + p, n, device=Zy.device)
            _lag_tiled(graph, Zy, Zx, perm, g, None, None, sims[p], den, cell_obs, cell_cnt, ws)
        return sims
    ws = _workspace(L.sc_perm_null_values_workspace_bytes(n, g), Zy.device)
    ldc = cell_cnt.shape[1] if cell_cnt is not None else 0
    check(
        L.sc_perm_null_values(_ptr(graph.indptr), _ptr(graph.indices), _ptr(graph.weights), n, int(graph.k_fixed),
                              _ptr(Zx), _ptr(Zy), ld, g, source, _ptr(pidx), int(seed) & (2**64 - 1), int(perm_offset),
                              int(n_perms), _ptr(sims), _ptr(cell_obs), _ptr(cell_cnt), ldc, _ptr(ws), ws.numel(),
                              _stream()),
        "sc_perm_null_values",
    )
    return sims


@_on_device
def philox_permutation(seed: int, perm_index: int, n: int, device="cuda") -> torch.Tensor:
    L = _lib.lib()
    out = torch.empty(n, dtype=torch.int32, device=device)
    check(L.sc_philox_permutation(int(seed) & (2**64 - 1), int(perm_index), n, _ptr(out), _stream()), "sc_philox_permutation")
    return out


@_on_device
def null_accumulate(sims: torch.Tensor, scale: Optional[torch.Tensor], obs: torch.Tensor, cnt_ge: torch.Tensor,
                    cnt_abs_ge: torch.Tensor, ssum: torch.Tensor, ssq: torch.Tensor) -> None:
    L = _lib.lib()
    n_perms, g = sims.shape
    check(
        L.sc_null_accumulate(_ptr(sims), n_perms, g, _ptr(scale), _ptr(obs), _ptr(cnt_ge), _ptr(cnt_abs_ge), _ptr(ssum),
                             _ptr(ssq), _stream()),
        "sc_null_accumulate",
    )


@_on_device
def lee_gemm(A: torch.Tensor, B: torch.Tensor, g: int, impl: int = 0) -> torch.Tensor:
    """``sc_lee_gemm``: ``L[x,y] = Σ_i A[i,x]·B[i,y]`` (float32 [g,g])."""
    L = _lib.lib()
    n = A.shape[0]
    out = torch.empty((g, g), dtype=torch.float32, device=A.device)
    ws = _workspace(L.sc_lee_gemm_workspace_bytes(n, g), A.device)
    check(
        L.sc_lee_gemm(_ptr(A), A.shape[1], _ptr(B), B.shape[1], n, g, _ptr(out), g, int(impl), _ptr(ws), ws.numel(), _stream()),
        "sc_lee_gemm",
    )
    return out


@_on_device
def lee_abs_ge_accumulate(Lp: torch.Tensor, L_obs: torch.Tensor, cnt: torch.Tensor) -> None:
    """``sc_lee_abs_ge_accumulate``: ``cnt += (|Lp| >= |L_obs|)`` for a permuted all-pairs matrix."""
    Lb = _lib.lib()
    g = L_obs.shape[0]
    check(Lb.sc_lee_abs_ge_accumulate(_ptr(Lp), Lp.stride(0), _ptr(L_obs), L_obs.stride(0), g, _ptr(cnt), cnt.stride(0), _stream()),
          "sc_lee_abs_ge_accumulate")


# --------------------------------------------------------------------------------------------------
# k-means on neighbourhood profiles (identify_niches)
# --------------------------------------------------------------------------------------------------


class KMeansDevice:
    """Device state of one k-means problem: the profile matrix, labels, distances and workspaces.
    Every numeric step is a kernel of ``csrc/niches.cu``; the host only draws random numbers, divides
    K x d sums by counts and tests convergence."""

    @_on_device
    def __init__(self, X: torch.Tensor, k: int) -> None:
        _require_cuda(X, "X")
        if X.dim() != 2:
            raise ValueError("X must be 2-D")
        if X.dtype != torch.float32:
            X = X.to(torch.float32)
        if X.stride(1) != 1:
            X = X.contiguous()
        self.X, self.k = X, int(k)
        self.n, self.d = X.shape
        dev = X.device
        L = _lib.lib()
        self.labels = torch.full((self.n,), -1, dtype=torch.int32, device=dev)
        self.mind = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.mind_tmp = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.out = torch.empty(self.k * self.d + self.k + 2, dtype=torch.float64, device=dev)
        self.ws = _workspace(L.sc_kmeans_workspace_bytes(self.n, self.d, self.k), dev)
        self.ws_sample = _workspace(L.sc_kmeans_pp_sample_workspace_bytes(self.n), dev)

    @_on_device
    def assign(self, centers: np.ndarray, want_mind: bool = False):
        """One Lloyd pass with ``centers`` (float32 [k, d]).  Returns (sums[k,d], counts[k], inertia,
        n_changed) as host float64 / ints; ``self.labels`` holds the new labels."""
        L = _lib.lib()
        c = torch.from_numpy(np.ascontiguousarray(centers, dtype=np.float32)).to(self.X.device)
        check(
            L.sc_kmeans_assign(_ptr(self.X), self.n, self.X.stride(0), self.d, _ptr(c), self.k, _ptr(self.labels),
                               _ptr(self.mind) if want_mind else None, _ptr(self.out), _ptr(self.ws), self.ws.numel(), _stream()),
            "sc_kmeans_assign",
        )
        o = self.out.cpu().numpy()
        kd = self.k * self.d
        return o[:kd].reshape(self.k, self.d).copy(), o[kd:kd + self.k].copy(), float(o[kd + self.k]), int(round(o[kd + self.k + 1]))

    @_on_device
    def pp_potential(self, cand: np.ndarray, first: bool, commit: int = -1) -> np.ndarray:
        """Potentials of candidate centres (row indices); ``commit`` folds that candidate into mind."""
        L = _lib.lib()
        c = torch.from_numpy(np.ascontiguousarray(cand, dtype=np.int32)).to(self.X.device)
        pot = torch.empty(len(cand), dtype=torch.float64, device=self.X.device)
        check(
            L.sc_kmeans_pp_potential(_ptr(self.X), self.n, self.X.stride(0), self.d, _ptr(c), len(cand),
                                     None if first else _ptr(self.mind), _ptr(self.mind_tmp), int(commit), _ptr(pot),
                                     _ptr(self.ws), self.ws.numel(), _stream()),
            "sc_kmeans_pp_potential",
        )
        if commit >= 0:
            self.mind, self.mind_tmp = self.mind_tmp, self.mind
        return pot.cpu().numpy()

    @_on_device
    def pp_sample(self, vals: np.ndarray) -> np.ndarray:
        """``searchsorted(cumsum(mind), vals)``: D^2 sampling of candidate rows."""
        L = _lib.lib()
        v = torch.from_numpy(np.ascontiguousarray(vals, dtype=np.float64)).to(self.X.device)
        idx = torch.empty(len(vals), dtype=torch.int32, device=self.X.device)
        check(
            L.sc_kmeans_pp_sample(_ptr(self.mind), self.n, _ptr(v), len(vals), _ptr(idx), _ptr(self.ws_sample),
                                  self.ws_sample.numel(), _stream()),
            "sc_kmeans_pp_sample",
        )
        return idx.cpu().numpy().astype(np.int64)

    def rows(self, idx) -> np.ndarray:
        return self.X[torch.as_tensor(np.asarray(idx, dtype=np.int64), device=self.X.device)].cpu().numpy()


# --------------------------------------------------------------------------------------------------
# domain distances
# --------------------------------------------------------------------------------------------------


@_on_device
def cross_nn(targets, queries, device="cuda") -> Tuple[np.ndarray, np.ndarray]:
    """``sc_cross_nn``: index (into ``targets``) and FP64 distance of the nearest target of every query
    point -- ``cKDTree(targets).query(queries, k=1)``.  Returns host arrays ``(dist, idx)``."""
    t = _coords_tensor(targets, device)
    q = _coords_tensor(queries, device)
    L = _lib.lib()
    idx = torch.empty(q.shape[0], dtype=torch.int32, device=t.device)
    dist = torch.empty(q.shape[0], dtype=torch.float64, device=t.device)
    ws = _workspace(L.sc_cross_nn_workspace_bytes(t.shape[0]), t.device)
    check(L.sc_cross_nn(_ptr(t), t.shape[0], _ptr(q), q.shape[0], _ptr(idx), _ptr(dist), _ptr(ws), ws.numel(), _stream()),
          "sc_cross_nn")
    return dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)


@_on_device
def pairwise_reduce(a, b, device="cuda") -> Tuple[float, float]:
    """``sc_pairwise_reduce``: ``(cdist(a, b).min(), cdist(a, b).sum())`` in FP64 on the device."""
    A = a if isinstance(a, torch.Tensor) else _coords_tensor(a, device)
    B = b if isinstance(b, torch.Tensor) else _coords_tensor(b, device)
    L = _lib.lib()
    out = torch.empty(2, dtype=torch.float64, device=A.device)
    ws = _workspace(L.sc_pairwise_reduce_workspace_bytes(), A.device)
    check(L.sc_pairwise_reduce(_ptr(A), A.shape[0], _ptr(B), B.shape[0], _ptr(out), _ptr(ws), ws.numel(), _stream()),
          "sc_pairwise_reduce")
    o = out.cpu().numpy()
    return float(o[0]), float(o[1])


_FDR_METHODS = {"none": 0, "bonferroni": 1, "fdr_bh": 2}


@_on_device
def local_moran_finish(cnt: Optional[torch.Tensor], Z: torch.Tensor, lag: torch.Tensor, loc: torch.Tensor, g: int,
                       n_perms: int, zero_var: Optional[torch.Tensor], method: str, alpha: float,
                       order: Optional[torch.Tensor] = None):
    """``sc_local_moran_finish``: per-cell p, adjusted p and LISA quadrant on the device, every output
    un-sorted to the user's cell order.  Returns device tensors ``(z, lag, I, p, p_adj, quadrant)`` of
    shape [n, g] (float32 x5, int8)."""
    L = _lib.lib()
    n, ld = Z.shape
    dev = Z.device
    outs = [torch.empty((n, g), dtype=torch.float32, device=dev) for _ in range(5)]
    quad = torch.empty((n, g), dtype=torch.int8, device=dev)
    ws = _workspace(L.sc_local_moran_finish_workspace_bytes(g, n_perms), dev)
    check(
        L.sc_local_moran_finish(_ptr(cnt), cnt.shape[1] if cnt is not None else 0, _ptr(Z), _ptr(lag), _ptr(loc), ld,
                                _ptr(order), n, g, int(n_perms), _ptr(zero_var), _FDR_METHODS[method], float(alpha),
                                _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), _ptr(outs[4]), _ptr(quad),
                                _ptr(ws), ws.numel(), _stream()),
        "sc_local_moran_finish",
    )
    return outs[0], outs[1], outs[2], outs[3], outs[4], quad
