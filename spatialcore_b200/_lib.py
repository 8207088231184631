"""ctypes binding of ``libsc_b200.so`` (C ABI declared in ``include/sc_b200.h``).

The product path has no CPU fallback: if the shared library is missing or fails to load, importing
this module's ``lib()`` raises ``RuntimeError`` telling the user how to build it.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsc_b200.so")

SC_F32, SC_F64 = 0, 1
SC_PERM_REPLAY, SC_PERM_PHILOX = 0, 1
SC_KNN_MAX_K = 128

_i32, _i64, _u64, _f64, _sz, _vp = C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_size_t, C.c_void_p

# name -> (restype, argtypes).  Pointers are passed as integers (device addresses) via c_void_p.
SIGNATURES = {
    "sc_version": (_i32, []),
    "sc_last_error": (C.c_char_p, []),
    "sc_launch_count": (C.c_longlong, []),
    "sc_grid_knn_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_grid_knn": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "sc_grid_radius_workspace_bytes": (_sz, [_i64]),
    "sc_grid_radius_count": (_i32, [_vp, _i64, _f64, _vp, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "sc_grid_radius_fill": (_i32, [_vp, _i64, _f64, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "sc_cross_nn_workspace_bytes": (_sz, [_i64]),
    "sc_cross_nn": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "sc_pairwise_reduce_workspace_bytes": (_sz, []),
    "sc_pairwise_reduce": (_i32, [_vp, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "sc_nbhd_counts": (_i32, [_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "sc_profile_normalize": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp]),
    "sc_graph_moments_workspace_bytes": (_sz, [_i64]),
    "sc_graph_moments": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _sz, _vp]),
    "sc_zscore_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_zscore": (_i32, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sc_zscore_apply": (_i32, [_vp, _i32, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sc_zscore_scatter": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "sc_csr_densify": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _i64, _vp]),
    "sc_csr_lag_moran_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_csr_lag_moran": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "sc_graph_tile_bytes": (_sz, [_i64, _i64]),
    "sc_graph_tile_build": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _sz, _vp]),
    "sc_csr_lag_moran_tiled": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _sz, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64,
                                      _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "sc_perm_null_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_perm_null_graph_rows": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _u64, _i64, _i32, _vp, _vp, _sz, _vp]),
    "sc_perm_null_values": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64, _i32, _i32, _vp, _u64, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "sc_perm_null_values_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_gather_rows": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _i64, _vp]),
    "sc_perm_conjugate": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "sc_spatial_order_workspace_bytes": (_sz, [_i64]),
    "sc_spatial_order": (_i32, [_vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "sc_graph_relabel_workspace_bytes": (_sz, [_i64]),
    "sc_graph_relabel": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sc_local_moran_finish_workspace_bytes": (_sz, [_i32, _i32]),
    "sc_local_moran_finish": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i32, _i32, _vp, _i32, C.c_float,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sc_philox_permutation": (_i32, [_u64, _i64, _i64, _vp, _vp]),
    "sc_philox_permutation_host": (_i32, [_u64, _i64, _i64, _vp]),
    "sc_null_accumulate": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sc_kmeans_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "sc_kmeans_assign": (_i32, [_vp, _i64, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "sc_kmeans_pp_potential": (_i32, [_vp, _i64, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _sz, _vp]),
    "sc_kmeans_pp_sample_workspace_bytes": (_sz, [_i64]),
    "sc_kmeans_pp_sample": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _sz, _vp]),
    "sc_lee_abs_ge_accumulate": (_i32, [_vp, _i64, _vp, _i64, _i32, _vp, _i64, _vp]),
    "sc_lee_gemm_workspace_bytes": (_sz, [_i64, _i32]),
    "sc_lee_gemm": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _i64, _i32, _vp, _sz, _vp]),
}

_lock = threading.Lock()
_lib: Optional[C.CDLL] = None


class SCError(RuntimeError):
    """A libsc_b200 call returned a negative status."""


def lib() -> C.CDLL:
    """Load (once) and return the shared library; fails loudly when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"spatialcore_b200: CUDA extension not found at {LIB_PATH}. There is no CPU fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C spatialcore_b200/csrc`)."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale: fail loudly
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().sc_last_error()
        text = msg.decode("utf-8", "replace") if msg else ""
        if status == -1:
            raise ValueError(text or f"{what}: invalid argument")
        raise SCError(f"{what} failed with status {status}: {text}")
