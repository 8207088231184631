"""ctypes wrapper of oracle/moran_port.c (TEST INFRASTRUCTURE ONLY: CPU baseline + cross-check)."""

import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libmoran_port.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError(f"{_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(_PATH)
        _lib.moran_port_run.restype = C.c_int
        _lib.moran_port_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.moran_port_threads.restype = C.c_int
        _lib.moran_port_set_threads.argtypes = [C.c_int]
        _lib.moran_port_set_threads.restype = None
    return _lib


def use_all_cores() -> int:
    """Use every core this process may run on (ignores OMP_NUM_THREADS, which torchrun sets to 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().moran_port_set_threads(n)
    return threads()


def threads() -> int:
    return int(lib().moran_port_threads())


def morans_i(g_csr, vals_gn: np.ndarray, perms: np.ndarray = None):
    """``g_csr``: row-normalised scipy CSR (float64); ``vals_gn``: (G, N) float64; ``perms``: (P, N)
    int32 or None.  Returns ``(score[G], sims[P, G])``."""
    L = lib()
    indptr = np.ascontiguousarray(g_csr.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(g_csr.indices, dtype=np.int32)
    data = np.ascontiguousarray(g_csr.data, dtype=np.float64)
    vals = np.ascontiguousarray(vals_gn, dtype=np.float64)
    g, n = vals.shape
    score = np.empty(g, dtype=np.float64)
    if perms is None:
        P, pp, sims = 0, None, np.empty((0, g))
    else:
        perms = np.ascontiguousarray(perms, dtype=np.int32)
        P, pp, sims = perms.shape[0], perms.ctypes.data, np.empty((perms.shape[0], g), dtype=np.float64)
    rc = L.moran_port_run(indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, n, vals.ctypes.data, g,
                          pp, P, score.ctypes.data, sims.ctypes.data if P else None)
    if rc != 0:
        raise MemoryError("moran_port_run failed")
    return score, sims
