"""Packing of two conserved charges into one integer quantum number
(pytenet/qnumber.py:8-25): q = (qa << 16) + qb with qb in [-2^15, 2^15)."""

__all__ = ["encode_quantum_number_pair", "decode_quantum_number_pair"]

_SHIFT = 16
_HALF = 1 << (_SHIFT - 1)
_FULL = 1 << _SHIFT


def encode_quantum_number_pair(qa: int, qb: int):
    """Encode `(qa, qb)` as a single integer."""
    return (qa << _SHIFT) + qb


def decode_quantum_number_pair(qnum: int):
    """Inverse of :func:`encode_quantum_number_pair`."""
    qb = ((qnum + _HALF) % _FULL) - _HALF
    return (qnum - qb) >> _SHIFT, qb
