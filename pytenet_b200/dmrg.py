"""
DMRG ground-state search for device-resident MPS -- `dmrg_singlesite`,
`dmrg_twosite` with the signatures of pytenet/dmrg.py:22,96 (Schollwoeck,
Ann. Phys. 326, 96 (2011)).  Local eigenproblems are solved by the
device-resident Lanczos driver; environments, the state and the two-site MPO
tensors never leave the GPU.
"""
import numpy as np
import torch

from . import _device as dev
from .mps import (MPS, mps_local_orthonormalize_left_qr, mps_local_orthonormalize_right_qr,
                  mps_merge_tensor_pair, mps_split_tensor_svd)
from .mpo import MPO, mpo_merge_tensor_pair
from ._sweep import prepare_environments, minimize_local_energy, sector_plan, env_step_left, env_step_right
from .block_sparse_util import qnumber_flatten
from ._prof import region

__all__ = ["dmrg_singlesite", "dmrg_twosite"]


def _renormalize_first_site(psi):
    """Right-normalise the leftmost tensor so that `psi` has unit norm (dmrg.py:87-88, 172-173)."""
    unit = torch.ones((1, 1, 1), dtype=dev.F64, device=psi.a[0].device)
    psi.a[0], _, psi.qbonds[0] = mps_local_orthonormalize_right_qr(psi.a[0], unit, psi.qsite, psi.qbonds[:2])


def dmrg_singlesite(hamiltonian: MPO, psi: MPS, numsweeps: int, numiter_lanczos: int = 25):
    """
    Single-site DMRG: left and right sweeps of local single-site optimisations.
    `psi` is the starting state and is updated in place; its bond dimensions
    cannot increase.

    Returns:
        numpy.ndarray: approximate ground-state energy after each sweep
    """
    nsites = hamiltonian.nsites
    _, lblocks, rblocks = prepare_environments(hamiltonian, psi)
    ham, k = hamiltonian.a, numiter_lanczos
    qh = hamiltonian.qbonds
    en_min = np.zeros(numsweeps)

    def site_plan(i):
        return sector_plan(psi.qbonds[i], psi.qsite, psi.qbonds[i + 1], qh[i], qh[i + 1], psi.a[i],
                           lblocks[i], rblocks[i], ham[i])

    for n in range(numsweeps):
        en = 0
        for i in range(nsites - 1):                                          # dmrg.py:65-73
            en, psi.a[i] = minimize_local_energy(ham[i], lblocks[i], rblocks[i], psi.a[i], k, site_plan(i))
            with region("qr"):
                psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = mps_local_orthonormalize_left_qr(
                    psi.a[i], psi.a[i + 1], psi.qsite, psi.qbonds[i:i + 2])
            lblocks[i + 1] = env_step_left(psi, hamiltonian, i, lblocks[i])
        for i in reversed(range(1, nsites)):                                 # dmrg.py:76-84
            en, psi.a[i] = minimize_local_energy(ham[i], lblocks[i], rblocks[i], psi.a[i], k, site_plan(i))
            with region("qr"):
                psi.a[i], psi.a[i - 1], psi.qbonds[i] = mps_local_orthonormalize_right_qr(
                    psi.a[i], psi.a[i - 1], psi.qsite, psi.qbonds[i:i + 2])
            rblocks[i - 1] = env_step_right(psi, hamiltonian, i, rblocks[i])
        _renormalize_first_site(psi)
        en_min[n] = en                     # energy of the last local problem of the sweep (dmrg.py:91)

    return en_min


def dmrg_twosite(hamiltonian: MPO, psi: MPS, numsweeps: int, numiter_lanczos: int = 25, tol_split: float = 0):
    """
    Two-site DMRG: left and right sweeps of local two-site optimisations followed by
    SVD splits with truncation tolerance `tol_split`.  `psi` is updated in place.

    Returns:
        numpy.ndarray: approximate ground-state energy after each sweep
    """
    nsites = hamiltonian.nsites
    _, lblocks, rblocks = prepare_environments(hamiltonian, psi)
    ham, k, qs = hamiltonian.a, numiter_lanczos, psi.qsite
    en_min = np.zeros(numsweeps)
    h2 = [mpo_merge_tensor_pair(ham[i], ham[i + 1]) for i in range(nsites - 1)]       # dmrg.py:135

    qh = hamiltonian.qbonds
    qs2 = qnumber_flatten([qs, qs])                   # quantum numbers of the merged physical index

    def optimize_pair(i, distr):
        with region("glue"):
            merged = mps_merge_tensor_pair(psi.a[i], psi.a[i + 1])
        plan = sector_plan(psi.qbonds[i], qs2, psi.qbonds[i + 2], qh[i], qh[i + 2], merged,
                           lblocks[i], rblocks[i + 1], h2[i])
        en, merged = minimize_local_energy(h2[i], lblocks[i], rblocks[i + 1], merged, k, plan)
        with region("svd"):
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = mps_split_tensor_svd(
                merged, qs, qs, [psi.qbonds[i], psi.qbonds[i + 2]], distr, tol=tol_split)
        return en

    for n in range(numsweeps):
        en = 0
        for i in range(nsites - 2):                                          # dmrg.py:142-154
            en = optimize_pair(i, "right")
            lblocks[i + 1] = env_step_left(psi, hamiltonian, i, lblocks[i])
        for i in reversed(range(nsites - 1)):                                # dmrg.py:157-169
            en = optimize_pair(i, "left")
            rblocks[i] = env_step_right(psi, hamiltonian, i + 1, rblocks[i + 1])
        _renormalize_first_site(psi)
        en_min[n] = en

    return en_min
