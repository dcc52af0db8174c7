"""
Device-time breakdown of a sweep by category (Lanczos runs, environment updates, QR, SVD, glue).

Disabled by default (`region` is then a no-op costing one attribute read).  When enabled, every region
records a pair of CUDA events on the current stream -- no host synchronisation, so the asynchronous
single-site TDVP sweep keeps running ahead -- and `report()` sums the elapsed times per category after one
final synchronise.  Regions may nest; a nested region's time is also part of its parent's, so only
leaf categories are used by the sweep drivers.  Used by bench.py's `sweeps` block.
"""
import contextlib

import torch

_enabled = False
_events = []          # (name, start event, end event)


def enable(flag=True):
    global _enabled
    _enabled = bool(flag)
    _events.clear()


@contextlib.contextmanager
def _timed(name):
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    try:
        yield
    finally:
        e1.record()
        _events.append((name, e0, e1))


_NULL = contextlib.nullcontext()


def region(name):
    """Context manager timing the device work enqueued inside it under category `name`."""
    return _timed(name) if _enabled else _NULL


def report():
    """{category: milliseconds of device time} of everything recorded since `enable()`; clears the log."""
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in _events:
        out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
    _events.clear()
    return out
