"""Input generators shared with the reference's seed semantics (pytenet/util.py:9-17)."""
import numpy as np

__all__ = ["crandn"]


def crandn(size=None, rng: np.random.Generator = None):
    """Standard complex normal samples (N(0,1) + i N(0,1)) / sqrt(2), drawn on the host
    with the same two `rng.normal` calls as the reference so seeds reproduce its tensors."""
    if rng is None:
        rng = np.random.default_rng()
    re = rng.normal(size=size)
    im = rng.normal(size=size)
    return (re + 1j * im) / np.sqrt(2)
