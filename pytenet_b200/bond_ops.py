"""
Virtual-bond truncation (pytenet/bond_ops.py:23-54).  The singular values are a
short host vector; the index rule is evaluated on the host exactly as the
reference does (normalise, square, accumulate ascending, keep cumsum > tol,
indices in original order), so retained indices match bit-for-bit given the
same singular values.
"""
import numpy as np
import torch

from .block_sparse_util import block_sparse_svd

__all__ = ["retained_bond_indices", "split_block_sparse_matrix_svd"]


def retained_bond_indices(s, tol):
    """Indices of the singular values kept for tolerance `tol` (:23-38)."""
    s = np.asarray(s, dtype=float)
    nrm = np.linalg.norm(s)
    if nrm == 0:
        return np.array([], dtype=int)
    weights = (s / nrm) ** 2
    order = np.argsort(weights)
    cum = np.empty_like(weights)
    cum[order] = np.cumsum(weights[order])
    return np.where(cum > tol)[0]


def split_block_sparse_matrix_svd(a, q0, q1, tol):
    """Sector-wise SVD followed by truncation -> `(u, s, v, q)` (:41-54); `u`, `v` stay on the device."""
    u, s, v, q = block_sparse_svd(a, q0, q1)
    keep = retained_bond_indices(s, tol)
    if len(keep) != len(s):
        kt = torch.as_tensor(keep, device=u.device)
        u = u.index_select(1, kt)
        v = v.index_select(0, kt)
        s = s[keep]
        q = q[keep]
    return u, s, v, q
