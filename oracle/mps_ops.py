"""
NumPy restatement of the MPS-level operations next to the hot path (TEST
INFRASTRUCTURE, see oracle/__init__.py): apply_mpo, MPS.compress (svd / density), mps_add, MPS.from_vector
and the SVD-based local orthonormalisation, on ``oracle.sweeps.Chain`` objects.  Pinned against the imported
reference by tests/golden/make_golden.py (mps_ops_case).
"""
import numpy as np
from . import blocksparse as ob
from . import sweeps as osw


def apply_mpo(w_list, w_qbonds, psi):
    """Restates chain_ops.py:215-234."""
    qbonds = [ob.qnumber_flatten((w_qbonds[i], psi.qbonds[i])) for i in range(psi.nsites + 1)]
    tensors = []
    for w, a in zip(w_list, psi.a):
        cl, dout, din, cr = w.shape
        Dl, d, Dr = a.shape
        # t[k, i, s', kappa, j] = sum_s w[k, s', s, kappa] a[i, s, j]
        t = np.einsum("kpsc,isj->kipcj", w, a)
        tensors.append(t.reshape(cl * Dl, dout, cr * Dr))
    return osw.Chain(tensors, psi.qsite, qbonds)


def local_orthonormalize_left_svd(a, a_next, qsite, qbonds, tol):
    """Restates mps.py:494-508."""
    Dl, d, Dr = a.shape
    u, sigma, v, qb = ob.split_block_sparse_matrix_svd(a.reshape(Dl * d, Dr), ob.qnumber_flatten((qbonds[0], qsite)),
                                                       qbonds[1], tol)
    sv = sigma[:, None] * v
    a_next = (sv @ a_next.reshape(a_next.shape[0], -1)).reshape((sv.shape[0],) + a_next.shape[1:])
    return u.reshape(Dl, d, u.shape[1]), a_next, qb


def local_orthonormalize_right_svd(a, a_prev, qsite, qbonds, tol):
    """Restates mps.py:511-525."""
    Dl, d, Dr = a.shape
    u, sigma, v, qb = ob.split_block_sparse_matrix_svd(a.reshape(Dl, d * Dr), qbonds[0],
                                                       ob.qnumber_flatten([-np.asarray(qsite), qbonds[1]]), tol)
    a_prev = np.tensordot(a_prev, u * sigma, (2, 0))
    return v.reshape(v.shape[0], d, Dr), a_prev, qb


def compress_svd(psi, tol, direction="left"):
    """Restates mps.py:196-236; returns (norm, scale) and modifies `psi` in place."""
    n = psi.nsites
    one = np.array([[[1]]])
    if direction == "left":
        nrm = osw.orthonormalize_right(psi)
        for i in range(n - 1):
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = local_orthonormalize_left_svd(
                psi.a[i], psi.a[i + 1], psi.qsite, psi.qbonds[i:i + 2], tol)
        psi.a[-1], t, psi.qbonds[-1] = local_orthonormalize_left_svd(psi.a[-1], one, psi.qsite, psi.qbonds[-2:], tol)
        psi.a[-1] = psi.a[-1] * (t[0, 0, 0] / abs(t[0, 0, 0]))
    elif direction == "right":
        nrm = osw.orthonormalize_left(psi)
        for i in reversed(range(1, n)):
            psi.a[i], psi.a[i - 1], psi.qbonds[i] = local_orthonormalize_right_svd(
                psi.a[i], psi.a[i - 1], psi.qsite, psi.qbonds[i:i + 2], tol)
        psi.a[0], t, psi.qbonds[0] = local_orthonormalize_right_svd(psi.a[0], one, psi.qsite, psi.qbonds[:2], tol)
        psi.a[0] = psi.a[0] * (t[0, 0, 0] / abs(t[0, 0, 0]))
    else:
        raise ValueError(f'`direction` = {direction} invalid; must be "left" or "right".')
    return nrm, abs(t[0, 0, 0])


def block_sparse_eigh(a, q0):
    """Restates block_sparse_util.py:183-241 (sectors in the iteration order of set(q0))."""
    q0 = np.asarray(q0)
    n = a.shape[0]
    order = np.argsort(q0, kind="stable")
    u = np.zeros((n, n), dtype=a.dtype)
    evals = np.zeros(n)
    q = np.zeros(n, dtype=q0.dtype)
    pos = 0
    for qn in set(q0):
        idx = order[q0[order] == qn]
        ev, us = np.linalg.eigh(a[np.ix_(idx, idx)])
        u[idx, pos:pos + len(idx)] = us
        evals[pos:pos + len(idx)] = ev
        q[pos:pos + len(idx)] = qn
        pos += len(idx)
    return u, evals, q


def compress_density(psi, tol):
    """Restates mps.py:238-290."""
    n = psi.nsites
    lblocks = [np.identity(1, dtype=psi.a[0].dtype)]
    for a in psi.a:
        t = np.tensordot(lblocks[-1], a.conj(), axes=(1, 0))
        lblocks.append(np.tensordot(a, t, axes=((0, 1), (0, 1))))
    nrm = np.sqrt(lblocks[-1][0, 0].real)
    b = np.array([[[1]]], dtype=psi.a[-1].dtype)
    u = np.array([[[1]]], dtype=psi.a[-1].dtype)
    for i in reversed(range(1, n)):
        b = np.tensordot(b, u.conj(), axes=((1, 2), (1, 2)))
        b = np.tensordot(psi.a[i], b, axes=((2,), (0,)))
        rho = np.tensordot(b, lblocks[i], axes=((0,), (0,)))
        rho = np.tensordot(rho, b.conj(), axes=((2,), (0,)))
        shp = rho.shape[0:2]
        rho = rho.reshape(shp[0] * shp[1], -1)
        qr = ob.qnumber_flatten((psi.qsite, -psi.qbonds[i + 1]))
        u, evals, qe = block_sparse_eigh(rho, qr)
        idx = ob.retained_bond_indices(np.abs(evals), tol)
        u = u[:, idx]
        qe = -qe[idx]
        u = u.reshape(shp[0], shp[1], u.shape[1]).transpose(2, 0, 1)
        psi.a[i] = u
        psi.qbonds[i] = qe
    b = np.tensordot(b, u.conj(), axes=((1, 2), (1, 2)))
    b = np.tensordot(psi.a[0], b, axes=((2,), (0,)))
    s = np.linalg.norm(b.reshape(-1))
    psi.a[0] = b / s
    return nrm, s / nrm


def mps_add(p0, p1, alpha=1):
    """Restates mps.py:571-617."""
    n = p0.nsites
    qbonds = [p0.qbonds[0].copy()] + [np.concatenate((p0.qbonds[i], p1.qbonds[i])) for i in range(1, n)] \
        + [p0.qbonds[-1].copy()]
    if n == 1:
        return osw.Chain([p0.a[0] + alpha * p1.a[0]], p0.qsite, qbonds)
    tensors = [np.concatenate((p0.a[0], alpha * p1.a[0]), axis=2)]
    for i in range(1, n - 1):
        s0, s1 = p0.a[i].shape, p1.a[i].shape
        t = np.zeros((s0[0] + s1[0], s0[1], s0[2] + s1[2]), dtype=np.result_type(p0.a[i], p1.a[i]))
        t[:s0[0], :, :s0[2]] = p0.a[i]
        t[s0[0]:, :, s0[2]:] = p1.a[i]
        tensors.append(t)
    tensors.append(np.concatenate((p0.a[-1], p1.a[-1]), axis=0))
    return osw.Chain(tensors, p0.qsite, qbonds)


def from_vector(d, nsites, v, tol=0):
    """Restates mps.py:304-337 (TT-SVD, all quantum numbers zero)."""
    v = np.asarray(v).reshape(1, -1)
    tensors, qbonds = [], [np.zeros(1, dtype=int)]
    for i in range(nsites):
        bl = v.shape[0]
        u, s, v = np.linalg.svd(v.reshape(bl * d, d ** (nsites - i - 1)), full_matrices=False)
        idx = ob.retained_bond_indices(s, tol)
        u, v, s = u[:, idx], v[idx, :], s[idx]
        v = v * s[:, None]
        tensors.append(u.reshape(bl, d, len(s)))
        qbonds.append(np.zeros(len(s), dtype=int))
    tensors[-1] = tensors[-1] * v[0, 0]
    return osw.Chain(tensors, np.zeros(d, dtype=int), qbonds)
