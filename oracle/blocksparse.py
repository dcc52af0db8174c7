"""
NumPy restatement of the quantum-number (block-sparse) helpers that sit either
side of every matvec batch (TEST INFRASTRUCTURE, see oracle/__init__.py).

Sector conventions restated from pytenet/block_sparse_util.py:
  * sector order = ascending common quantum numbers (np.intersect1d, :33-37)
  * rows/cols grouped by a *stable* sort of their quantum numbers (:82-83)
  * each sector contributes min(rows, cols) bond indices, in sector order
"""
import numpy as np


def qnumber_outer_sum(qnums):
    """Tensor of all sums q0[i0] + q1[i1] + ...  Restates block_sparse_util.py:13-30."""
    if len(qnums) == 0:
        return np.array(0)
    acc = np.asarray(qnums[0])
    for q in qnums[1:]:
        acc = np.add.outer(acc, np.asarray(q))
    return acc


def qnumber_flatten(qnums):
    """Restates block_sparse_util.py:40-44."""
    return qnumber_outer_sum(qnums).reshape(-1)


def is_qsparse(a, qnums):
    """True iff a is zero wherever the quantum numbers do not sum to zero.
    Restates block_sparse_util.py:47-53."""
    forbidden = qnumber_outer_sum(qnums) != 0
    return not np.any(np.asarray(a)[forbidden]) if np.ndim(forbidden) else not (forbidden and np.any(a))


def enforce_qsparsity(a, qnums):
    """Vectorised equivalent of block_sparse_util.py:56-64 (in place)."""
    a[qnumber_outer_sum(qnums) != 0] = 0


def _sector_plan(q0, q1):
    """Stable sector grouping shared by QR and SVD (block_sparse_util.py:67-103, 122, 151-155).

    Returns (sectors, rows, cols) where rows[n]/cols[n] are the *original*
    indices of sector n in stable order."""
    q0 = np.asarray(q0); q1 = np.asarray(q1)
    sectors = np.intersect1d(q0, q1)
    order0 = np.argsort(q0, kind="stable")
    order1 = np.argsort(q1, kind="stable")
    rows = [order0[q0[order0] == q] for q in sectors]
    cols = [order1[q1[order1] == q] for q in sectors]
    return sectors, rows, cols


def block_sparse_qr(a, q0, q1):
    """Sector-wise reduced QR.  Restates block_sparse_util.py:106-180
    (incl. the no-common-sector special case :124-134)."""
    a = np.asarray(a)
    q0 = np.asarray(q0); q1 = np.asarray(q1)
    assert a.ndim == 2 and len(q0) == a.shape[0] and len(q1) == a.shape[1]
    assert is_qsparse(a, [q0, -q1])
    sectors, rows, cols = _sector_plan(q0, q1)
    if len(sectors) == 0:
        assert np.linalg.norm(a) == 0
        q = np.zeros((a.shape[0], 1), dtype=a.dtype)
        r = np.zeros((1, a.shape[1]), dtype=a.dtype)
        q[0, 0] = 1
        return q, r, q0[:1]
    sizes = [min(len(ri), len(ci)) for ri, ci in zip(rows, cols)]
    nb = sum(sizes)
    q = np.zeros((a.shape[0], nb), dtype=a.dtype)
    r = np.zeros((nb, a.shape[1]), dtype=a.dtype)
    qinterm = np.zeros(nb, dtype=q0.dtype)
    pos = 0
    for qn, ri, ci, sz in zip(sectors, rows, cols, sizes):
        qs, rs = np.linalg.qr(a[np.ix_(ri, ci)], mode="reduced")
        q[ri, pos:pos + sz] = qs
        r[pos:pos + sz, ci] = rs
        qinterm[pos:pos + sz] = qn
        pos += sz
    return q, r, qinterm


def block_sparse_svd(a, q0, q1):
    """Sector-wise thin SVD.  Restates block_sparse_util.py:244-319."""
    a = np.asarray(a)
    q0 = np.asarray(q0); q1 = np.asarray(q1)
    assert a.ndim == 2 and len(q0) == a.shape[0] and len(q1) == a.shape[1]
    assert is_qsparse(a, [q0, -q1])
    sectors, rows, cols = _sector_plan(q0, q1)
    if len(sectors) == 0:
        assert np.linalg.norm(a) == 0
        u = np.zeros((a.shape[0], 1), dtype=a.dtype)
        v = np.zeros((1, a.shape[1]), dtype=a.dtype)
        if a.shape[0] > 0:
            u[0, 0] = 1
        return u, np.zeros(1), v, q0[:1]
    sizes = [min(len(ri), len(ci)) for ri, ci in zip(rows, cols)]
    nb = sum(sizes)
    u = np.zeros((a.shape[0], nb), dtype=a.dtype)
    v = np.zeros((nb, a.shape[1]), dtype=a.dtype)
    s = np.zeros(nb)
    qb = np.zeros(nb, dtype=q0.dtype)
    pos = 0
    for qn, ri, ci, sz in zip(sectors, rows, cols, sizes):
        us, ss, vs = np.linalg.svd(a[np.ix_(ri, ci)], full_matrices=False)
        u[ri, pos:pos + sz] = us
        v[pos:pos + sz, ci] = vs
        s[pos:pos + sz] = ss
        qb[pos:pos + sz] = qn
        pos += sz
    return u, s, v, qb


def retained_bond_indices(s, tol):
    """Indices kept by the truncation rule.  Restates pytenet/bond_ops.py:23-38:
    normalise, square, accumulate ascending (default argsort), keep cumsum > tol,
    return indices in original order."""
    s = np.asarray(s, dtype=float)
    nrm = np.linalg.norm(s)
    if nrm == 0:
        return np.array([], dtype=int)
    p = (s / nrm) ** 2
    order = np.argsort(p)
    acc = np.empty_like(p)
    acc[order] = np.cumsum(p[order])
    return np.where(acc > tol)[0]


def split_block_sparse_matrix_svd(a, q0, q1, tol):
    """Restates pytenet/bond_ops.py:41-54."""
    u, s, v, q = block_sparse_svd(a, q0, q1)
    keep = retained_bond_indices(s, tol)
    return u[:, keep], s[keep], v[keep, :], q[keep]


def encode_quantum_number_pair(qa, qb):
    """Restates pytenet/qnumber.py:8-12."""
    return (qa << 16) + qb
