"""Load the UNMODIFIED reference source from /root/reference/src (this container only).

``anndata`` and ``squidpy`` are not installed, and the reference imports both at module
top [R src/spatialcore/spatial/autocorrelation.py:46,50; core/metadata.py:8].  Two stub
modules are placed in ``sys.modules`` (a duck-typed ``anndata.AnnData`` and an empty
``squidpy.gr``) and the reference is imported as-is.  Everything except ``morans_i``
(which needs real squidpy) then runs: ``build_spatial_weights``, ``local_morans_i``,
``lees_l``, ``lees_l_local``, ``compute_neighborhood_profile``.

TEST INFRASTRUCTURE ONLY.  Never imported by the product, never available on the GPU box.
"""

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "spatialcore"))


def load():
    """Returns ``(autocorrelation_module, neighborhoods_module, AnnData_class)``."""
    if not available():
        raise RuntimeError("reference source not present (expected only in the build container)")
    from spatialcore_b200.anndata_lite import AnnDataLite

    if "anndata" not in sys.modules:
        ad = types.ModuleType("anndata")
        ad.AnnData = AnnDataLite
        sys.modules["anndata"] = ad
    if "squidpy" not in sys.modules:
        sq = types.ModuleType("squidpy")
        sq.gr = types.SimpleNamespace()
        sys.modules["squidpy"] = sq
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import logging

    ac = importlib.import_module("spatialcore.spatial.autocorrelation")
    nb = importlib.import_module("spatialcore.spatial.neighborhoods")
    logging.getLogger("spatialcore").setLevel(logging.ERROR)
    return ac, nb, AnnDataLite
