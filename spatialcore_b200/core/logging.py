"""``spatialcore.*`` logger, same names and default format as the reference
[R src/spatialcore/core/logging.py:37-59] so log scraping keeps working."""

import logging
import sys
from typing import Optional

_ROOT = "spatialcore"
_FORMAT = "[%(levelname)s] %(name)s: %(message)s"
_ready = False


def _ensure_handler() -> None:
    global _ready
    if _ready:
        return
    root = logging.getLogger(_ROOT)
    if not root.handlers:
        h = logging.StreamHandler(sys.stdout)
        h.setFormatter(logging.Formatter(_FORMAT))
        h.setLevel(logging.INFO)
        root.addHandler(h)
        root.setLevel(logging.INFO)
        root.propagate = False
    _ready = True


def get_logger(name: Optional[str] = None) -> logging.Logger:
    _ensure_handler()
    return logging.getLogger(f"{_ROOT}.{name}" if name else _ROOT)


def setup_logging(level: int = logging.INFO, format_string: Optional[str] = None) -> None:
    global _ready
    root = logging.getLogger(_ROOT)
    root.handlers.clear()
    h = logging.StreamHandler(sys.stdout)
    h.setFormatter(logging.Formatter(format_string or _FORMAT))
    h.setLevel(level)
    root.addHandler(h)
    root.setLevel(level)
    root.propagate = False
    _ready = True
