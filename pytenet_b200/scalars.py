"""
Small host-side helpers shared with the reference's conventions:

* packing of two conserved charges (e.g. particle number and spin of the Fermi-Hubbard model) into a
  single additive integer quantum number, `q = (qa << 16) + qb` with `qb` in [-2^15, 2^15)
  (pytenet/qnumber.py:8-25; used by hamiltonian/fermi_hubbard.py:34);
* `crandn`, the complex standard normal generator whose draw order defines what a seed means for
  `MPS(..., fill="random", rng=...)` (pytenet/util.py:9-17: real parts first, then imaginary parts).
"""
import numpy as np

__all__ = ["encode_quantum_number_pair", "decode_quantum_number_pair", "crandn"]

_PAIR_BITS = 16


def encode_quantum_number_pair(qa: int, qb: int):
    """Single integer carrying the pair `(qa, qb)`; additive in both components."""
    return qb + (qa << _PAIR_BITS)


def decode_quantum_number_pair(qnum: int):
    """Recover `(qa, qb)` from :func:`encode_quantum_number_pair` (`qb` is the signed low half-word)."""
    half = 1 << (_PAIR_BITS - 1)
    qb = ((qnum + half) & ((1 << _PAIR_BITS) - 1)) - half
    qa = (qnum - qb) >> _PAIR_BITS
    return qa, qb


def crandn(size=None, rng: np.random.Generator = None):
    """Samples of (x + i y) / sqrt(2) with x, y ~ N(0, 1).  All real parts are drawn before all imaginary
    parts -- two `rng.normal(size=size)` calls -- so a seeded generator yields the reference's numbers."""
    gen = np.random.default_rng() if rng is None else rng
    parts = [gen.normal(size=size) for _ in range(2)]
    return (parts[0] + 1j * parts[1]) / np.sqrt(2)      # division, not multiplication: bit-identical to the reference
