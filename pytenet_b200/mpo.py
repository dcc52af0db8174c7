"""
Matrix product operator container with device-resident tensors
(pytenet/mpo.py:15-185) and the two-site merge used by the two-site sweeps
(pytenet/mpo.py:314-322).  MPO tensors have shape `(b[i], d, d, b[i+1])`:
(left bond, physical out, physical in, right bond); float64 unless a local
operator is complex.
"""
import numpy as np
import torch

from . import _device as dev
from .block_sparse_util import qnumber_outer_sum

__all__ = ["MPO", "mpo_merge_tensor_pair"]


class MPO:
    """Matrix product operator; `a[i]` are CUDA torch tensors, quantum numbers host integers."""

    def __init__(self, qsite, qbonds, fill=0.0, rng=None, device=None):
        self.qsite = np.array(qsite)
        self.qbonds = [np.array(qb) for qb in qbonds]
        self.device = torch.device(device) if device is not None else dev.default_device()
        d = len(self.qsite)
        b = [len(qb) for qb in self.qbonds]
        nsites = len(b) - 1
        if isinstance(fill, (int, float, complex)) and not isinstance(fill, bool):
            host = [np.full((b[i], d, d, b[i + 1]), fill) for i in range(nsites)]
        elif fill in ("random", "random real"):
            from .scalars import crandn
            rng = np.random.default_rng() if rng is None else rng
            draw = (lambda s: crandn(s, rng)) if fill == "random" else (lambda s: rng.normal(size=s))
            host = [draw((b[i], d, d, b[i + 1])) / np.sqrt(b[i] * d * b[i + 1]) for i in range(nsites)]
        elif fill == "postpone":
            host = None
        else:
            raise ValueError(f'`fill` = {fill} invalid; must be a number, '
                             f'"random", "random real" or "postpone".')
        if host is None:
            self.a = nsites * [None]
        else:
            self.a = []
            for i, t in enumerate(host):
                t[qnumber_outer_sum([self.qbonds[i], self.qsite, -self.qsite, -self.qbonds[i + 1]]) != 0] = 0
                self.a.append(dev.to_device(t, self.device))

    @classmethod
    def from_tensors(cls, qsite, qbonds, tensors, device=None):
        """Wrap host (NumPy) or device tensors, e.g. the output of a reference Hamiltonian builder."""
        op = cls(qsite, qbonds, fill="postpone", device=device)
        op.a = [dev.dense(dev.to_device(t, op.device)) for t in tensors]
        assert len(op.a) == len(op.qbonds) - 1
        return op

    @property
    def nsites(self) -> int:
        return len(self.a)

    @property
    def bond_dims(self) -> list:
        if len(self.a) == 0:
            return []
        return [t.shape[0] for t in self.a] + [self.a[-1].shape[3]]

    def zero_qnumbers(self):
        self.qsite = np.zeros_like(self.qsite)
        self.qbonds = [np.zeros_like(qb) for qb in self.qbonds]
        return self

    def to_matrix(self) -> np.ndarray:
        """Dense matrix on the full Hilbert space (host; validation at small sizes only)."""
        t = dev.to_host(self.a[0])
        for nxt in self.a[1:]:
            t = dev.to_host(mpo_merge_tensor_pair(dev.to_device(t, self.device), nxt))
        assert t.shape[0] == 1 and t.shape[3] == 1
        return t[0, :, :, 0]


def mpo_merge_tensor_pair(a0, a1):
    """
    Merge two neighbouring MPO tensors into `(b0, d0*d1, d0*d1, b2)` (mpo.py:314-322).
    Runs once per sweep call (tdvp.py:163, dmrg.py:135); the bond contraction is a
    GEMM on the engine, the physical-axis interleave a device permute.
    """
    b0, p0, q0, b1 = a0.shape
    b1b, p1, q1, b2 = a1.shape
    assert b1 == b1b
    t = dev.gemm(a0.reshape(b0 * p0 * q0, b1), a1.reshape(b1, p1 * q1 * b2))
    t = dev.dense(t.reshape(b0, p0, q0, p1, q1, b2).permute(0, 1, 3, 2, 4, 5))
    return t.reshape(b0, p0 * p1, q0 * q1, b2)
