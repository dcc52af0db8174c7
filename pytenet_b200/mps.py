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
from .block_sparse_util import (qnumber_outer_sum, qnumber_flatten, block_sparse_qr, block_sparse_eigh, is_qsparse,
                                single_block_svd)
from .bond_ops import split_block_sparse_matrix_svd, retained_bond_indices
from .scalars import crandn

__all__ = ["MPS", "mps_vdot", "mps_norm", "mps_add", "mps_merge_tensor_pair", "mps_split_tensor_svd",
           "mps_local_orthonormalize_left_qr", "mps_local_orthonormalize_right_qr",
           "mps_local_orthonormalize_left_svd", "mps_local_orthonormalize_right_svd"]


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

    def compress(self, tol: float, mode="svd", direction="left"):
        """
        Compress and orthonormalise (mps.py:180-290): "svd" = site-local SVDs with singular-value
        truncation, "density" = rounding by the local density matrix (McCulloch, J. Stat. Mech. (2007) P10014).
        Returns `(original norm, scaling factor due to compression)`.
        """
        if mode == "svd":
            return self._compress_svd(tol, direction)
        if mode == "density":
            return self._compress_density(tol)
        raise ValueError(f'`mode` = {mode} invalid; must be "svd" or "density".')

    def _compress_svd(self, tol, direction):
        n = len(self.a)
        if direction == "left":
            nrm = self.orthonormalize(mode="right")
            for i in range(n - 1):
                self.a[i], self.a[i + 1], self.qbonds[i + 1] = mps_local_orthonormalize_left_svd(
                    self.a[i], self.a[i + 1], self.qsite, self.qbonds[i:i + 2], tol)
            self.a[-1], t, self.qbonds[-1] = mps_local_orthonormalize_left_svd(
                self.a[-1], _scalar_block(self.device), self.qsite, self.qbonds[-2:], tol)
            edge = n - 1
        elif direction == "right":
            nrm = self.orthonormalize(mode="left")
            for i in reversed(range(1, n)):
                self.a[i], self.a[i - 1], self.qbonds[i] = mps_local_orthonormalize_right_svd(
                    self.a[i], self.a[i - 1], self.qsite, self.qbonds[i:i + 2], tol)
            self.a[0], t, self.qbonds[0] = mps_local_orthonormalize_right_svd(
                self.a[0], _scalar_block(self.device), self.qsite, self.qbonds[:2], tol)
            edge = 0
        else:
            raise ValueError(f'`direction` = {direction} invalid; must be "left" or "right".')
        for i in range(n):
            assert is_qsparse(self.a[i], [self.qbonds[i], self.qsite, -self.qbonds[i + 1]]), \
                "sparsity pattern of MPS tensor does not match quantum numbers"
        assert tuple(t.shape) == (1, 1, 1)
        t0 = t.reshape(-1)[0].item()
        # absorb a potential phase factor into the edge tensor (mps.py:216,235)
        self.a[edge] = self.a[edge] * (t0 / abs(t0))
        return (nrm, abs(t0))

    def _compress_density(self, tol):
        lblocks = _mps_compute_left_blocks(self, self)
        top = lblocks[-1].reshape(-1)[0].item()
        assert tuple(lblocks[-1].shape) == (1, 1) and np.real(top) > 0
        nrm = float(np.sqrt(np.real(top)))
        dtype = self.a[-1].dtype
        b = torch.ones((1, 1, 1), dtype=dtype, device=self.device)
        u = torch.ones((1, 1, 1), dtype=dtype, device=self.device)
        for i in reversed(range(1, self.nsites)):
            # b[j, m] = sum_{s,j'} b_prev[j, s, j'] conj(u[m, s, j'])  then  b[i, s, m] = a[i, s, j] b[j, m]
            bm = dev.gemm(b.reshape(b.shape[0], -1), u.reshape(u.shape[0], -1), trans_b=True, conj_b=True)
            ai = self.a[i]
            b = dev.gemm(ai.reshape(-1, ai.shape[2]), bm).reshape(ai.shape[0], ai.shape[1], bm.shape[1])
            # rho[(s,m),(s',m')] = sum_{i,i'} b[i,s,m] L[i,i'] conj(b[i',s',m'])
            lb = dev.gemm(lblocks[i], b.reshape(b.shape[0], -1), trans_a=True)             # [i', (s,m)] = L^T b
            rho = dev.gemm(lb, b.reshape(b.shape[0], -1), trans_a=True, conj_b=True)       # [(s,m), (s',m')]
            qnums_rho = qnumber_flatten((self.qsite, -self.qbonds[i + 1]))
            assert is_qsparse(rho, (qnums_rho, -qnums_rho))
            uu, evals, qnums_eig = block_sparse_eigh(rho, qnums_rho)
            idx = retained_bond_indices(np.abs(evals), tol)       # eigenvalues are real but can be negative
            uu = uu.index_select(1, torch.as_tensor(idx, device=uu.device))
            qnums_eig = -qnums_eig[idx]
            u = dev.dense(uu.reshape(ai.shape[1], b.shape[2], len(idx)).permute(2, 0, 1))
            self.a[i] = u
            self.qbonds[i] = qnums_eig
        bm = dev.gemm(b.reshape(b.shape[0], -1), u.reshape(u.shape[0], -1), trans_b=True, conj_b=True)
        a0 = self.a[0]
        b = dev.gemm(a0.reshape(-1, a0.shape[2]), bm).reshape(a0.shape[0], a0.shape[1], bm.shape[1])
        s = float(torch.linalg.norm(b.reshape(-1)).item())
        self.a[0] = b / s
        return (nrm, s / nrm)

    @classmethod
    def from_vector(cls, d: int, nsites: int, v, tol: float = 0, device=None):
        """
        MPS representation of the vector `v` by the TT-SVD algorithm for local dimension `d`; all quantum
        numbers are zero (mps.py:304-337).  SVDs on the device (cuSOLVER), truncation rule on the host.
        """
        mps = cls(d * [0], [[0] for _ in range(nsites + 1)], fill="postpone", device=device)
        v = dev.to_device(v, mps.device)
        assert v.ndim == 1 and len(v) == d ** nsites, f"`v` has length {len(v)}, expecting {d**nsites}."
        v = v.reshape(1, -1)
        from .block_sparse_util import _SVD_DRIVER
        for i in range(nsites):
            bleft = v.shape[0]
            u, s, vh = torch.linalg.svd(dev.dense(v).reshape(bleft * d, d ** (nsites - i - 1)), full_matrices=False,
                                        driver=_SVD_DRIVER)
            idx = retained_bond_indices(s.cpu().numpy(), tol)
            it = torch.as_tensor(idx, device=v.device)
            u = u.index_select(1, it); s = s.index_select(0, it)
            v = dev.dense(vh).index_select(0, it) * s[:, None]
            mps.a[i] = dev.dense(u.reshape(bleft, d, len(idx)))
            mps.qbonds[i + 1] = np.zeros(len(idx), dtype=int)
        assert tuple(v.shape) == (1, 1)
        mps.a[-1] = mps.a[-1] * v.reshape(-1)[0]
        return mps

    def __add__(self, other):
        return mps_add(self, other)

    def __sub__(self, other):
        return mps_add(self, other, alpha=-1)

    def to_vector(self) -> np.ndarray:
        """Full Hilbert-space vector as a NumPy array (validation at small sizes; mps.py:292-301)."""
        psi = self.a[0]
        for nxt in self.a[1:]:
            psi = mps_merge_tensor_pair(psi, nxt)
        assert psi.ndim == 3 and psi.shape[0] == 1 and psi.shape[2] == 1
        return dev.to_host(psi.reshape(-1))


def _mps_compute_left_blocks(chi: MPS, psi: MPS):
    """All partial contractions of `<chi | psi>` from the left (mps.py:384-427):
    `l_next[j, j'] = sum a[i,s,j] l[i,i'] conj(b[i',s,j'])`, two GEMMs per site on the engine."""
    nsites = chi.nsites
    assert nsites == psi.nsites
    first = psi.a[0]
    blocks = [torch.eye(1, dtype=first.dtype, device=first.device)]
    for a, b in zip(psi.a, chi.a):
        dl, d, dr = a.shape
        dlp, _, drp = b.shape
        t = dev.gemm(blocks[-1], b.reshape(dlp, d * drp), conj_b=True)                    # [i, (s, j')]
        blocks.append(dev.gemm(a.reshape(dl * d, dr), t.reshape(dl * d, drp), trans_a=True))   # [j, j']
    return blocks


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


def mps_local_orthonormalize_left_svd(a, a_next, qsite, qbonds, tol: float):
    """Left-orthonormalise `a` by a truncated SVD and absorb `sigma v` into the next tensor (mps.py:494-508)."""
    s = a.shape
    assert len(s) == 3
    u, sigma, v, qbond, sig = split_block_sparse_matrix_svd(
        a.reshape(s[0] * s[1], s[2]), qnumber_flatten((qbonds[0], qsite)), qbonds[1], tol, with_device_sigma=True)
    sv = dev.dense(v) * sig[:, None]
    return dev.dense(u.reshape(s[0], s[1], u.shape[1])), _left_multiply(sv, a_next), qbond


def mps_local_orthonormalize_right_svd(a, a_prev, qsite, qbonds, tol: float):
    """Right-orthonormalise `a` by a truncated SVD and absorb `u sigma` into the previous tensor (mps.py:511-525)."""
    s = a.shape
    assert len(s) == 3
    u, sigma, v, qbond, sig = split_block_sparse_matrix_svd(
        a.reshape(s[0], s[1] * s[2]), qbonds[0], qnumber_flatten([-np.asarray(qsite), qbonds[1]]), tol,
        with_device_sigma=True)
    us = dev.dense(u) * sig
    prev = dev.gemm(a_prev.reshape(-1, a_prev.shape[-1]), us).reshape(tuple(a_prev.shape[:-1]) + (us.shape[1],))
    return dev.dense(v.reshape(v.shape[0], s[1], s[2])), prev, qbond


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
    if (isinstance(a, torch.Tensor) and a.is_cuda
            and not (qsite0.any() or qsite1.any() or np.any(qbonds_outer[0]) or np.any(qbonds_outer[1]))):
        # all quantum numbers zero: one dense block (block_sparse_util.single_block_svd); same results as the
        # general path below
        res = single_block_svd(a.reshape(b0 * d0, d1 * b2))
        if res is not None:
            u, sigma, v, sig = res
            keep = retained_bond_indices(sigma, tol)
            nb = len(keep)
            if nb != len(sigma):
                if nb > 0 and keep[-1] == nb - 1:
                    u, v, sig = u[:, :nb], v[:nb], sig[:nb]
                else:
                    kt = torch.as_tensor(keep, device=u.device)
                    u, v, sig = u.index_select(1, kt), v.index_select(0, kt), sig.index_select(0, kt)
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
            qbond = np.zeros(nb, dtype=np.result_type(np.asarray(qbonds_outer[0]).dtype, qsite0.dtype))
            return dev.dense(u.reshape(b0, d0, nb)), dev.dense(v.reshape(nb, d1, b2)), qbond
    q0 = qnumber_flatten([qbonds_outer[0], qsite0])
    q1 = qnumber_flatten([-qsite1, qbonds_outer[1]])
    u, sigma, v, qbond, sig = split_block_sparse_matrix_svd(a.reshape(b0 * d0, d1 * b2), q0, q1, tol,
                                                            with_device_sigma=True)
    nb = len(sigma)
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


def mps_add(mps0: MPS, mps1: MPS, alpha=1) -> MPS:
    """
    Logical sum `mps0 + alpha mps1`: virtual bond dimensions add, tensors become block diagonal
    (mps.py:571-617).  Pure data movement on the device.
    """
    assert mps0.nsites == mps1.nsites
    nsites = mps0.nsites
    assert np.array_equal(mps0.qsite, mps1.qsite)
    assert np.array_equal(mps0.qbonds[0], mps1.qbonds[0]) and np.array_equal(mps0.qbonds[-1], mps1.qbonds[-1])
    qbonds = [np.asarray(mps0.qbonds[0]).copy()]
    qbonds += [np.concatenate((mps0.qbonds[i], mps1.qbonds[i])) for i in range(1, nsites)]
    qbonds.append(np.asarray(mps0.qbonds[-1]).copy())
    out = MPS(mps0.qsite, qbonds, fill="postpone", device=mps0.device)
    cplx = (dev.any_complex(*mps0.a, *mps1.a) or isinstance(alpha, complex)) if nsites else False
    a0 = [dev.as_dtype(t, cplx) for t in mps0.a]
    a1 = [dev.as_dtype(t, cplx) for t in mps1.a]
    if nsites == 1:
        out.a[0] = a0[0] + alpha * a1[0]
    elif nsites > 1:
        out.a[0] = torch.cat((a0[0], alpha * a1[0]), dim=2)
        for i in range(1, nsites - 1):
            s0, s1 = a0[i].shape, a1[i].shape
            t = torch.zeros((s0[0] + s1[0], s0[1], s0[2] + s1[2]), dtype=a0[i].dtype, device=a0[i].device)
            t[:s0[0], :, :s0[2]] = a0[i]
            t[s0[0]:, :, s0[2]:] = a1[i]
            out.a[i] = t
        out.a[-1] = torch.cat((a0[-1], a1[-1]), dim=0)
    for i in range(nsites):
        out.a[i] = dev.dense(out.a[i])
        assert is_qsparse(out.a[i], (out.qbonds[i], out.qsite, -out.qbonds[i + 1])), \
            "sparsity pattern of MPS tensor does not match quantum numbers"
    return out
