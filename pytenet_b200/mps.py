"""
Matrix product state container with device-resident tensors, plus the local
operations the TDVP / DMRG sweeps call (pytenet/mps.py).

`MPS.a[i]` is a CUDA torch tensor of shape `(b[i], d, b[i+1])` (complex128 or
float64, C-contiguous -- the layout the C ABI expects); `qsite` / `qbonds` are
host integer arrays.  Random fills are drawn on the host with the reference's
exact sequence of generator calls (mps.py:50-61, util.py:9-17), so a seed
produces the same tensors as the reference, then uploaded once.
"""
import numpy as np
import torch

from . import _device as dev
from .block_sparse_util import qnumber_outer_sum, qnumber_flatten, block_sparse_qr
from .bond_ops import split_block_sparse_matrix_svd
from .scalars import crandn

__all__ = ["MPS", "mps_vdot", "mps_norm", "mps_merge_tensor_pair", "mps_split_tensor_svd",
           "mps_local_orthonormalize_left_qr", "mps_local_orthonormalize_right_qr"]


def _scalar_block(device):
    return torch.ones((1, 1, 1), dtype=dev.F64, device=device)


class MPS:
    """
    Matrix product state; tensor i has shape `(b[i], d, b[i+1])`.

    Quantum numbers are additive integers: for every non-zero entry the left bond
    and physical quantum numbers sum to the right bond quantum number
    (mps.py:14-26).
    """

    def __init__(self, qsite, qbonds, fill=0.0, rng=None, device=None):
        self.qsite = np.asarray(qsite).copy() if isinstance(qsite, np.ndarray) else np.asarray(qsite)
        self.qbonds = [np.asarray(qb) for qb in qbonds]
        self.device = torch.device(device) if device is not None else dev.default_device()
        d = len(self.qsite)
        b = [len(qb) for qb in self.qbonds]
        assert b[0] == 1 and b[-1] == 1, "leading and trailing bond dimensions must be 1"
        nsites = len(b) - 1
        if isinstance(fill, (int, float, complex)) and not isinstance(fill, bool):
            host = [np.full((b[i], d, b[i + 1]), fill) for i in range(nsites)]
        elif fill == "random":
            rng = np.random.default_rng() if rng is None else rng
            host = [crandn((b[i], d, b[i + 1]), rng) / np.sqrt(b[i] * d * b[i + 1]) for i in range(nsites)]
        elif fill == "random real":
            rng = np.random.default_rng() if rng is None else rng
            host = [rng.normal(size=(b[i], d, b[i + 1])) / np.sqrt(b[i] * d * b[i + 1]) for i in range(nsites)]
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
                # block sparsity dictated by the quantum numbers (vectorised mps.py:68-71)
                t[qnumber_outer_sum([self.qbonds[i], self.qsite, -self.qbonds[i + 1]]) != 0] = 0
                self.a.append(dev.to_device(t, self.device))

    @classmethod
    def construct_random(cls, nsites: int, qsite, qnum_sector: int, max_vdim: int = 256,
                         dtype="complex", rng: np.random.Generator = None, device=None):
        """
        Random MPS in the overall quantum-number sector `qnum_sector` with bond
        dimensions capped at `max_vdim` (mps.py:73-113): bond quantum numbers grow
        from both ends towards the centre; when more than `max_vdim` combinations
        exist a random subset is drawn (`rng.choice`, without replacement).
        """
        assert nsites > 0
        qsite = np.asarray(qsite)
        rng = np.random.default_rng() if rng is None else rng
        qbonds = [None] * (nsites + 1)
        qbonds[0] = [0]
        qbonds[nsites] = [qnum_sector]
        half = (nsites + 1) // 2

        def cap(candidates):
            if len(candidates) <= max_vdim:
                return candidates
            pick = rng.choice(len(candidates), size=max_vdim, replace=False)
            return candidates[pick]

        for pos in range(1, half):
            qbonds[pos] = cap(qnumber_flatten([qbonds[pos - 1], qsite]))
        for pos in range(nsites - 1, half - 1, -1):
            qbonds[pos] = cap(qnumber_flatten([qbonds[pos + 1], -qsite]))
        kind = "random" if dtype in (complex, "complex") else "random real"
        return cls(qsite, qbonds, fill=kind, rng=rng, device=device)

    @classmethod
    def from_tensors(cls, qsite, qbonds, tensors, device=None):
        """Wrap existing tensors (NumPy or torch), uploading them to the device."""
        psi = cls(qsite, qbonds, fill="postpone", device=device)
        psi.a = [dev.dense(dev.to_device(t, psi.device)) for t in tensors]
        assert len(psi.a) == len(psi.qbonds) - 1
        return psi

    @property
    def nsites(self) -> int:
        return len(self.a)

    @property
    def bond_dims(self) -> list:
        if len(self.a) == 0:
            return []
        return [t.shape[0] for t in self.a] + [self.a[-1].shape[2]]

    def zero_qnumbers(self):
        """Set every quantum number to zero (disables block sparsity); chainable."""
        self.qsite = np.zeros_like(self.qsite)
        self.qbonds = [np.zeros_like(qb) for qb in self.qbonds]
        return self

    def copy(self):
        other = MPS(self.qsite.copy(), [q.copy() for q in self.qbonds], fill="postpone", device=self.device)
        other.a = [t.clone() for t in self.a]
        return other

    def __deepcopy__(self, memo):
        return self.copy()

    def orthonormalize(self, mode="left"):
        """Left- or right-orthonormalise by local QR steps; returns the norm (mps.py:141-178)."""
        if len(self.a) == 0:
            return 1
        n = len(self.a)
        if mode == "left":
            for i in range(n - 1):
                self.a[i], self.a[i + 1], self.qbonds[i + 1] = mps_local_orthonormalize_left_qr(
                    self.a[i], self.a[i + 1], self.qsite, self.qbonds[i:i + 2])
            self.a[-1], t, self.qbonds[-1] = mps_local_orthonormalize_left_qr(
                self.a[-1], _scalar_block(self.device), self.qsite, self.qbonds[-2:])
            edge = n - 1
        elif mode == "right":
            for i in reversed(range(1, n)):
                self.a[i], self.a[i - 1], self.qbonds[i] = mps_local_orthonormalize_right_qr(
                    self.a[i], self.a[i - 1], self.qsite, self.qbonds[i:i + 2])
            self.a[0], t, self.qbonds[0] = mps_local_orthonormalize_right_qr(
                self.a[0], _scalar_block(self.device), self.qsite, self.qbonds[:2])
            edge = 0
        else:
            raise ValueError(f'`mode` = {mode} invalid; must be "left" or "right".')
        assert tuple(t.shape) == (1, 1, 1)
        nrm = t.reshape(-1)[0].real.item() if t.dtype.is_complex else t.reshape(-1)[0].item()
        if nrm < 0:
            self.a[edge] = -self.a[edge]
            nrm = -nrm
        return nrm

    def to_vector(self) -> np.ndarray:
        """Full Hilbert-space vector as a NumPy array (validation at small sizes; mps.py:292-301)."""
        psi = self.a[0]
        for nxt in self.a[1:]:
            psi = mps_merge_tensor_pair(psi, nxt)
        assert psi.ndim == 3 and psi.shape[0] == 1 and psi.shape[2] == 1
        return dev.to_host(psi.reshape(-1))


def mps_vdot(chi: MPS, psi: MPS):
    """
    Scalar product `<chi | psi>` (complex conjugating `chi`), contracted from the right on the device
    (pytenet/mps.py:430-448): per site `t <- sum a[i,s,j] t[j,j'] conj(b[i',s,j'])`, two GEMMs on the engine.
    Returns a Python scalar.
    """
    assert psi.nsites == chi.nsites
    if psi.nsites == 0:
        return 0
    last = psi.a[-1]
    n = last.shape[2]
    assert chi.a[-1].shape[2] == n
    t = torch.eye(n, dtype=last.dtype, device=last.device)
    for a, b in zip(reversed(psi.a), reversed(chi.a)):
        dl, d, dr = a.shape
        dlp, _, drp = b.shape
        at = dev.gemm(a.reshape(dl * d, dr), t)                                   # [(i,s), j']
        t = dev.gemm(at.reshape(dl, d * drp), b.reshape(dlp, d * drp), trans_b=True, conj_b=True)   # [i, i']
    assert tuple(t.shape) == (1, 1)
    return t.reshape(-1)[0].item()


def mps_norm(psi: MPS):
    """Standard L2 norm of a matrix product state (pytenet/mps.py:451-457)."""
    return float(np.sqrt(np.real(mps_vdot(psi, psi))))


def _left_multiply(m, t):
    """m (p x q) times tensor t (q, ...) over its first axis, on the DMMA engine."""
    out = dev.gemm(m, t.reshape(t.shape[0], -1))
    return out.reshape((m.shape[0],) + tuple(t.shape[1:]))


def _right_multiply_t(t, m):
    """t (..., q) contracted with m (p x q) over the last axis / second axis of m."""
    out = dev.gemm(t.reshape(-1, t.shape[-1]), m, trans_b=True)
    return out.reshape(tuple(t.shape[:-1]) + (m.shape[0],))


def mps_local_orthonormalize_left_qr(a, a_next, qsite, qbonds):
    """Left-orthonormalise `a` by QR and absorb `r` into the next tensor (mps.py:460-473)."""
    s = a.shape
    assert len(s) == 3
    q, r, qbond = block_sparse_qr(a.reshape(s[0] * s[1], s[2]), qnumber_flatten((qbonds[0], qsite)), qbonds[1])
    a = dev.dense(q.reshape(s[0], s[1], q.shape[1]))
    return a, _left_multiply(r, a_next), qbond


def mps_local_orthonormalize_right_qr(a, a_prev, qsite, qbonds):
    """Right-orthonormalise `a` by QR of its bond-flipped matricisation (mps.py:476-491)."""
    at = dev.dense(a.permute(2, 1, 0))
    s = at.shape
    assert len(s) == 3
    q, r, qbond = block_sparse_qr(at.reshape(s[0] * s[1], s[2]), qnumber_flatten((-qbonds[1], qsite)), -qbonds[0])
    a = dev.dense(q.reshape(s[0], s[1], q.shape[1]).permute(2, 1, 0))
    return a, _right_multiply_t(a_prev, r), -qbond


def mps_merge_tensor_pair(a0, a1):
    """Contract two neighbouring MPS tensors into `(b0, d0*d1, b2)` (mps.py:528-535)."""
    b0, d0, b1 = a0.shape
    b1b, d1, b2 = a1.shape
    assert b1 == b1b
    out = dev.gemm(a0.reshape(b0 * d0, b1), a1.reshape(b1, d1 * b2))
    return out.reshape(b0, d0 * d1, b2)


def mps_split_tensor_svd(a, qsite0, qsite1, qbonds_outer, svd_distr: str, tol=0):
    """
    Split a two-site tensor `(b0, d0*d1, b2)` by a sector-wise SVD with truncation
    tolerance `tol`; the singular values go "left", "right" or as "sqrt" to both
    sides (mps.py:538-568).
    """
    assert a.ndim == 3
    qsite0 = np.asarray(qsite0); qsite1 = np.asarray(qsite1)
    d0, d1 = len(qsite0), len(qsite1)
    assert d0 * d1 == a.shape[1], "physical dimension of MPS tensor must be equal to d0 * d1"
    b0, b2 = a.shape[0], a.shape[2]
    q0 = qnumber_flatten([qbonds_outer[0], qsite0])
    q1 = qnumber_flatten([-qsite1, qbonds_outer[1]])
    u, sigma, v, qbond = split_block_sparse_matrix_svd(a.reshape(b0 * d0, d1 * b2), q0, q1, tol)
    nb = len(sigma)
    sig = torch.as_tensor(sigma, device=a.device)
    if svd_distr == "left":
        u = u * sig
    elif svd_distr == "right":
        v = v * sig[:, None]
    elif svd_distr == "sqrt":
        rt = torch.sqrt(sig)
        u = u * rt
        v = v * rt[:, None]
    else:
        raise ValueError('`svd_distr` parameter must be "left", "right" or "sqrt".')
    return dev.dense(u.reshape(b0, d0, nb)), dev.dense(v.reshape(nb, d1, b2)), qbond
