"""Provenance log in ``adata.uns['spatialcore_metadata']['operations']``.

Observable output of every reference API function
[R src/spatialcore/core/metadata.py:11-110]; entries keep the reference's keys
(``timestamp, function, parameters, outputs``).
"""

from datetime import datetime
from pathlib import Path
from typing import Any, Dict, Optional

_KEY = "spatialcore_metadata"


def _jsonable(params: Dict[str, Any]) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for k, v in params.items():
        if isinstance(v, (str, int, float, bool, type(None))):
            out[k] = v
        elif isinstance(v, (list, tuple)):
            out[k] = list(v)
        elif isinstance(v, dict):
            out[k] = _jsonable(v)
        elif isinstance(v, Path):
            out[k] = str(v)
        else:
            out[k] = type(v).__name__
    return out


def update_metadata(
    adata: Any,
    function_name: str,
    parameters: Dict[str, Any],
    outputs: Optional[Dict[str, Any]] = None,
) -> None:
    meta = adata.uns.get(_KEY)
    if meta is None:
        meta = {"created": datetime.now().isoformat(), "operations": []}
        adata.uns[_KEY] = meta
    ops = meta.get("operations")
    if ops is None:
        meta["operations"] = ops = []
    elif not isinstance(ops, list):
        meta["operations"] = ops = list(ops)
    entry = {
        "timestamp": datetime.now().isoformat(),
        "function": function_name,
        "parameters": _jsonable(parameters),
    }
    if outputs:
        entry["outputs"] = outputs
    ops.append(entry)
