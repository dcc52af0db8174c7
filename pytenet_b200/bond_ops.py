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


def split_block_sparse_matrix_svd(a, q0, q1, tol, with_device_sigma=False):
    """Sector-wise SVD followed by truncation -> `(u, s, v, q)` (:41-54); `u`, `v` stay on the device.
    `with_device_sigma`: a fifth return value holds the kept singular values as a device vector."""
    u, s, v, q, s_dev = block_sparse_svd(a, q0, q1, with_device_sigma=True)
    keep = retained_bond_indices(s, tol)
    if len(keep) != len(s):
        nk = len(keep)
        if nk > 0 and keep[-1] == nk - 1:
            # the kept indices are a prefix (one sector, values descending): slices, no index upload
            u, v, s_dev = u[:, :nk], v[:nk], s_dev[:nk]
        else:
            kt = torch.as_tensor(keep, device=u.device)
            u = u.index_select(1, kt)
            v = v.index_select(0, kt)
            s_dev = s_dev.index_select(0, kt)
        s = s[keep]
        q = q[keep]
    if with_device_sigma:
        return u, s, v, q, s_dev
    return u, s, v, q
