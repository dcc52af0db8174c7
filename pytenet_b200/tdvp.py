"""
TDVP time integration for device-resident MPS -- `tdvp_singlesite`,
`tdvp_twosite` with the signatures of pytenet/tdvp.py:26,121.

Algorithm: symmetric (second-order) projector-splitting integrator of Haegeman,
Lubich, Oseledets, Vandereycken, Verstraete, Phys. Rev. B 94, 165116 (2016),
with Lanczos-based local exponentials.  Every tensor (state, environments,
two-site MPO tensors, Lanczos vectors) stays on the GPU; per local problem one
small device->host copy returns the Lanczos coefficients.
"""
from . import _device as dev
from .mps import MPS, mps_merge_tensor_pair, mps_split_tensor_svd
from .mpo import MPO, mpo_merge_tensor_pair
from .block_sparse_util import qnumber_flatten, block_sparse_qr
from ._sweep import (prepare_environments, local_hamiltonian_step, local_bond_step, sector_plan, bond_plan,
                     env_step_left, env_step_right)
from .krylov import defer_checks
from ._prof import region

__all__ = ["tdvp_singlesite", "tdvp_twosite"]


@defer_checks
def tdvp_singlesite(hamiltonian: MPO, psi: MPS, dt, numsteps: int, numiter_lanczos: int = 25):
    """
    Symmetric single-site TDVP integration; `psi` is overwritten in place.

    Args:
        hamiltonian: Hamiltonian as MPO
        psi: initial state as MPS
        dt: time step; purely imaginary `dt` gives real-time evolution
        numsteps: number of time steps
        numiter_lanczos: Lanczos iterations per local step

    Returns:
        float: norm of the initial `psi`
    """
    nsites = hamiltonian.nsites
    nrm, lblocks, rblocks = prepare_environments(hamiltonian, psi)
    ham, k = hamiltonian.a, numiter_lanczos
    qh = hamiltonian.qbonds

    def site_plan(i):
        return sector_plan(psi.qbonds[i], psi.qsite, psi.qbonds[i + 1], qh[i], qh[i + 1], psi.a[i],
                           lblocks[i], rblocks[i], ham[i])

    for _ in range(numsteps):
        # left -> right: half step on each site, backward half step on each bond (tdvp.py:68-84)
        for i in range(nsites - 1):
            psi.a[i] = local_hamiltonian_step(lblocks[i], rblocks[i], ham[i], psi.a[i], 0.5 * dt, k, site_plan(i))
            b0, d, b1 = psi.a[i].shape
            qold = psi.qbonds[i + 1]
            with region("qr"):
                q, c, psi.qbonds[i + 1] = block_sparse_qr(
                    psi.a[i].reshape(b0 * d, b1), qnumber_flatten((psi.qbonds[i], psi.qsite)), psi.qbonds[i + 1])
                psi.a[i] = dev.dense(q.reshape(b0, d, q.shape[1]))
            lblocks[i + 1] = env_step_left(psi, hamiltonian, i, lblocks[i])
            c = dev.dense(c)
            c = local_bond_step(lblocks[i + 1], rblocks[i], c, -0.5 * dt, k,
                                bond_plan(psi.qbonds[i + 1], qold, qh[i + 1], c, lblocks[i + 1], rblocks[i]))
            with region("glue"):
                nxt = psi.a[i + 1]
                psi.a[i + 1] = dev.gemm(c, nxt.reshape(nxt.shape[0], -1)).reshape(
                    (c.shape[0],) + tuple(nxt.shape[1:]))

        # full step on the last site (tdvp.py:87-89)
        i = nsites - 1
        psi.a[i] = local_hamiltonian_step(lblocks[i], rblocks[i], ham[i], psi.a[i], dt, k, site_plan(i))

        # right -> left (tdvp.py:92-115)
        for i in reversed(range(1, nsites)):
            with region("qr"):
                at = dev.dense(psi.a[i].permute(2, 1, 0))
                b1, d, b0 = at.shape
                q, c, qbond = block_sparse_qr(
                    at.reshape(b1 * d, b0), qnumber_flatten((-psi.qbonds[i + 1], psi.qsite)), -psi.qbonds[i])
                qold = psi.qbonds[i]
                psi.qbonds[i] = -qbond
                psi.a[i] = dev.dense(q.reshape(b1, d, q.shape[1]).permute(2, 1, 0))
            rblocks[i - 1] = env_step_right(psi, hamiltonian, i, rblocks[i])
            c = dev.dense(c.T)
            c = local_bond_step(lblocks[i], rblocks[i - 1], c, -0.5 * dt, k,
                                bond_plan(qold, psi.qbonds[i], qh[i], c, lblocks[i], rblocks[i - 1]))
            with region("glue"):
                prv = psi.a[i - 1]
                psi.a[i - 1] = dev.gemm(prv.reshape(-1, prv.shape[2]), c).reshape(
                    tuple(prv.shape[:2]) + (c.shape[1],))
            psi.a[i - 1] = local_hamiltonian_step(
                lblocks[i - 1], rblocks[i - 1], ham[i - 1], psi.a[i - 1], 0.5 * dt, k, site_plan(i - 1))

    return nrm


@defer_checks
def tdvp_twosite(hamiltonian: MPO, psi: MPS, dt, numsteps: int, numiter_lanczos: int = 25, tol_split=0):
    """
    Symmetric two-site TDVP integration; `psi` is overwritten in place.

    Args:
        hamiltonian: Hamiltonian as MPO
        psi: initial state as MPS
        dt: time step; purely imaginary `dt` gives real-time evolution
        numsteps: number of time steps
        numiter_lanczos: Lanczos iterations per local step
        tol_split: truncation tolerance of the SVD splits

    Returns:
        float: norm of the initial `psi`
    """
    nsites = hamiltonian.nsites
    assert nsites >= 2
    nrm, lblocks, rblocks = prepare_environments(hamiltonian, psi)
    ham, k, qs = hamiltonian.a, numiter_lanczos, psi.qsite
    h2 = [mpo_merge_tensor_pair(ham[i], ham[i + 1]) for i in range(nsites - 1)]       # tdvp.py:163

    qh = hamiltonian.qbonds
    qs2 = qnumber_flatten([qs, qs])                   # quantum numbers of the merged physical index

    def evolve_pair(i, tau, distr):
        with region("glue"):
            merged = mps_merge_tensor_pair(psi.a[i], psi.a[i + 1])
        plan = sector_plan(psi.qbonds[i], qs2, psi.qbonds[i + 2], qh[i], qh[i + 2], merged,
                           lblocks[i], rblocks[i + 1], h2[i])
        merged = local_hamiltonian_step(lblocks[i], rblocks[i + 1], h2[i], merged, tau, k, plan)
        with region("svd"):
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = mps_split_tensor_svd(
                merged, qs, qs, (psi.qbonds[i], psi.qbonds[i + 2]), distr, tol=tol_split)

    def backward_site(i):
        plan = sector_plan(psi.qbonds[i], qs, psi.qbonds[i + 1], qh[i], qh[i + 1], psi.a[i],
                           lblocks[i], rblocks[i], ham[i])
        psi.a[i] = local_hamiltonian_step(lblocks[i], rblocks[i], ham[i], psi.a[i], -0.5 * dt, k, plan)

    for _ in range(numsteps):
        # left -> right (tdvp.py:168-184)
        for i in range(nsites - 2):
            evolve_pair(i, 0.5 * dt, "right")
            lblocks[i + 1] = env_step_left(psi, hamiltonian, i, lblocks[i])
            backward_site(i + 1)
        # rightmost pair, full step (tdvp.py:187-198)
        i = nsites - 2
        evolve_pair(i, dt, "left")
        rblocks[i] = env_step_right(psi, hamiltonian, i + 1, rblocks[i + 1])
        # right -> left (tdvp.py:201-217)
        for i in reversed(range(nsites - 2)):
            backward_site(i + 1)
            evolve_pair(i, 0.5 * dt, "left")
            rblocks[i] = env_step_right(psi, hamiltonian, i + 1, rblocks[i + 1])

    return nrm
