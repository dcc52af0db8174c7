"""
Shared machinery of the four sweep drivers (tdvp.py, dmrg.py): environment
bookkeeping and the local Krylov problems, all device-resident.

The closures handed to the Krylov drivers are the integration point named in
SURVEY.md section 8(a10) (pytenet/tdvp.py:223-238, dmrg.py:181-189): they bind
(w, l, r), view the flat Lanczos vector as the local tensor (no copy) and call
the fused contraction chain.
"""
import os

import numpy as np
import torch

from . import _device as dev
from .sectors import HeffSectorPlan
from .block_sparse_util import is_qsparse
from .chain_ops import (apply_local_hamiltonian, apply_local_bond_contraction,
                        compute_right_operator_blocks)
from .krylov import eigh_krylov, expm_krylov


def prepare_environments(hamiltonian, psi):
    """Right-orthonormalise `psi`, build all right blocks and the dummy left block, and run the
    reference's sparsity consistency check (tdvp.py:51-63, dmrg.py:44-56).  Returns (nrm, lblocks, rblocks)."""
    nsites = hamiltonian.nsites
    assert nsites == psi.nsites
    nrm = psi.orthonormalize(mode="right")
    rblocks = compute_right_operator_blocks(psi, hamiltonian)
    lblocks = [None for _ in range(nsites)]
    lblocks[0] = torch.ones((1, 1, 1), dtype=rblocks[0].dtype, device=rblocks[0].device)
    for i, rb in enumerate(rblocks):
        assert is_qsparse(rb, [psi.qbonds[i + 1], hamiltonian.qbonds[i + 1], -psi.qbonds[i + 1]]), \
            "sparsity pattern of operator blocks must match quantum numbers"
    return nrm, lblocks, rblocks


# Sector-banded matvec (pytenet_b200/sectors.py): "auto" uses it when the quantum numbers are
# non-trivial and the bonds are large enough for whole tiles to be skipped; "1" forces it, "0" disables.
_SECTOR_MODE = os.environ.get("PYTENET_B200_SECTORS", "auto")
_SECTOR_MIN_BOND = 256


def sector_plan(ql, qs, qr, qwl, qwr, like):
    """HeffSectorPlan for a local problem, or None when the dense path is the right choice."""
    if _SECTOR_MODE == "0":
        return None
    if _SECTOR_MODE != "1" and max(len(ql), len(qr)) < _SECTOR_MIN_BOND:
        return None
    if not (np.any(ql) or np.any(qr) or np.any(qs) or np.any(qwl) or np.any(qwr)):
        return None
    return HeffSectorPlan(ql, qs, qr, qwl, qwr, cplx=True if like is None else like.dtype.is_complex)


def _heff(w, l, r, shape, plan):
    if plan is not None:
        def matvec(x):
            if x.dtype.is_complex != plan.cplx:          # a real state turned complex (or vice versa)
                return apply_local_hamiltonian(x.reshape(shape), w, l, r).reshape(-1)
            return plan.apply(x.reshape(shape), w, l, r).reshape(-1)
        return matvec
    return lambda x: apply_local_hamiltonian(x.reshape(shape), w, l, r).reshape(-1)


def local_hamiltonian_step(l, r, w, a, dt, numiter: int, plan=None):
    """exp(-dt H_eff) a for the one- or two-site effective Hamiltonian (tdvp.py:223-229)."""
    shape = tuple(a.shape)
    return expm_krylov(_heff(w, l, r, shape, plan), a.reshape(-1), -dt, numiter, hermitian=True).reshape(shape)


def local_bond_step(l, r, c, dt, numiter: int):
    """exp(-dt K_eff) c for the zero-site (bond) effective Hamiltonian (tdvp.py:232-238)."""
    shape = tuple(c.shape)
    return expm_krylov(
        lambda x: apply_local_bond_contraction(x.reshape(shape), l, r).reshape(-1),
        c.reshape(-1), -dt, numiter, hermitian=True).reshape(shape)


def minimize_local_energy(w, l, r, a_start, numiter: int, plan=None):
    """Lowest Ritz pair of the local effective Hamiltonian (dmrg.py:181-189)."""
    shape = tuple(a_start.shape)
    ev, u_ritz = eigh_krylov(_heff(w, l, r, shape, plan), a_start.reshape(-1), numiter, 1)
    return ev[0], dev.dense(u_ritz[:, 0]).reshape(shape)
