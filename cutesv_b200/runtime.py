"""Process-wide Engine used by the drop-in resolution_* entry points (one csv_ctx per GPU)."""
import os

_engine = None


def get_engine():
    """The shared Engine on CUTESV_B200_DEVICE (default 0).  Raises when the CUDA library or a
    B200 is missing -- the drop-in API never computes on the CPU."""
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(int(os.environ.get("CUTESV_B200_DEVICE", "0")))
    return _engine


def set_engine(e):
    global _engine
    _engine = e
