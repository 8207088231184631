"""spatialcore_b200 — B200-native spatial statistics (drop-in for ``spatialcore.spatial``'s hot path).

Python host code over PyTorch tensors; all arithmetic in hand-written sm_100a CUDA kernels behind
the C ABI of ``libsc_b200.so`` (``include/sc_b200.h``).  No Triton, no CuPy, no CPU fallback.
"""

__version__ = "0.1.0"

from spatialcore_b200.anndata_lite import AnnDataLite  # noqa: F401
from spatialcore_b200 import core  # noqa: F401


def __getattr__(name):
    if name == "spatial":
        import importlib

        return importlib.import_module("spatialcore_b200.spatial")
    raise AttributeError(name)
