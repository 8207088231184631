"""Boundary helpers the hot path calls at the end of every API function
(logger + provenance log in ``adata.uns``), mirroring the reference's
``spatialcore.core`` [R src/spatialcore/core/logging.py, core/metadata.py]."""

from spatialcore_b200.core.logging import get_logger, setup_logging
from spatialcore_b200.core.metadata import update_metadata

__all__ = ["get_logger", "setup_logging", "update_metadata"]
