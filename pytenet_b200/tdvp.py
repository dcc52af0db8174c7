"""
TDVP time integration for device-resident MPS -- `tdvp_singlesite`,
`tdvp_twosite` with the signatures of pytenet/tdvp.py:26,121.

Algorithm: symmetric (second-order) projector-splitting integrator of Haegeman,
Lubich, Oseledets, Vandereycken, Verstraete, Phys. Rev. B 94, 165116 (2016),
with Lanczos-based local exponentials.  Every tensor (state, environments,
two-site MPO tensors, Lanczos vectors) stays on the GPU; per local problem one
small device->host copy returns the Lanczos coefficients.
"""
import os

import numpy as np

from . import _device as dev
from .mps import MPS, mps_merge_tensor_pair, mps_split_tensor_svd
from .mpo import MPO, mpo_merge_tensor_pair
from .block_sparse_util import qnumber_flatten, block_sparse_qr
from ._sweep import (prepare_environments, local_hamiltonian_step, local_bond_step, sector_plan, bond_plan,
                     env_step_left, env_step_right, absorb_left, absorb_right)
from .krylov import defer_checks
from ._prof import region

__all__ = ["tdvp_singlesite", "tdvp_twosite"]


# CUDA graphs for the launch-latency regime (README config: D <= 28): a single-site TDVP
# time step is a fixed sequence of ~400 small kernels with no host decision in between (bond dimensions cannot
# change, the tridiagonal problems are solved on the device, breakdown checks are deferred), so from the third
# time step on the whole step is ONE graph launch.  "auto": when every bond is at most _GRAPH_MAX_BOND.
_GRAPHS = os.environ.get("PYTENET_B200_GRAPHS", "auto")
_GRAPH_MAX_BOND = 192
_GRAPH_MIN_STEPS = 4


class _StepGraph:
    """One symmetric single-site TDVP time step captured as a CUDA graph.

    The graph reads the state (site tensors, environment blocks) from the tensors the Python objects held at
    capture time, runs the step, and copies the results back into those same tensors as its last nodes -- so the
    state lives at fixed addresses and `replay()` advances it by one time step.  Scalars of the Lanczos runs land in
    page-locked slots owned by the graph and are checked after every replay (same warnings / assertions as the
    eager path)."""

    last_error = None        # repr of the exception that made the most recent capture fall back to eager steps
    _pending = None          # scalars of the most recent replay, not yet examined

    def __init__(self, psi, lblocks, rblocks, one_step):
        import torch
        from . import krylov
        from . import block_sparse_util as bsu
        self.ok = False
        a0, l0, r0 = list(psi.a), list(lblocks), list(rblocks)
        q0 = [np.array(q) for q in psi.qbonds]
        krylov._flush_deferred()             # examine what the eager steps left; the ring is then empty
        meta0 = len(krylov._Deferred.meta)
        graph = torch.cuda.CUDAGraph()
        try:
            bsu._CAPTURING = True
            with torch.cuda.graph(graph):
                one_step()
                for dst, src in zip(a0 + l0 + r0, list(psi.a) + list(lblocks) + list(rblocks)):
                    if dst is src:
                        continue
                    if dst.shape != src.shape or dst.dtype != src.dtype:
                        raise RuntimeError("state layout changed during the time step")
                    dst.copy_(src)
            same_q = all(np.array_equal(x, y) for x, y in zip(q0, psi.qbonds))
        except Exception as exc:             # anything not capturable: keep running eagerly
            same_q = False
            _StepGraph.last_error = repr(exc)
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
        finally:
            bsu._CAPTURING = False
        # the capture executed nothing: the state is still the one before the step, held by a0 / l0 / r0
        psi.a[:] = a0
        lblocks[:] = l0
        rblocks[:] = r0
        self.meta = krylov._Deferred.meta[meta0:]
        self.slot0 = meta0
        del krylov._Deferred.meta[meta0:]
        if not same_q:
            psi.qbonds[:] = q0
            return
        self.graph = graph
        self.ok = True

    def replay(self):
        """Enqueue one time step; the scalars of the PREVIOUS replay are examined while the device runs this one
        (the slots are overwritten by every replay, so they are copied out right after the synchronisation)."""
        import torch
        self.graph.replay()
        self.finish()
        if self.meta:
            from . import krylov
            torch.cuda.current_stream().synchronize()
            host = krylov._Deferred.ring.numpy()
            self._pending = host[self.slot0:self.slot0 + len(self.meta)].copy()

    def finish(self):
        """Examine the scalars of the last replay (same warnings / assertions as the eager path)."""
        pending, self._pending = self._pending, None
        if pending is not None:
            from . import krylov
            for row, (n, numiter) in zip(pending, self.meta):
                krylov._check_scalars(row, n, numiter)


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

    def one_step():
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
                psi.a[i + 1] = absorb_left(c, psi.a[i + 1], psi.qbonds[i + 1], qold, psi.qsite, psi.qbonds[i + 2])

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
                psi.a[i - 1] = absorb_right(psi.a[i - 1], c, qold, psi.qbonds[i], psi.qsite, psi.qbonds[i - 1])
            psi.a[i - 1] = local_hamiltonian_step(
                lblocks[i - 1], rblocks[i - 1], ham[i - 1], psi.a[i - 1], 0.5 * dt, k, site_plan(i - 1))

    def signature():
        return tuple((tuple(t.shape), t.dtype) for t in psi.a)

    use_graph = (_GRAPHS == "1" or (_GRAPHS == "auto" and numsteps >= _GRAPH_MIN_STEPS
                                     and max(psi.bond_dims) <= _GRAPH_MAX_BOND))
    use_graph = use_graph and numiter_lanczos <= 64 and all(t.is_cuda for t in psi.a)
    graph = None
    prev_sig = None
    step = 0
    while step < numsteps:
        if graph is not None:
            graph.replay()
            step += 1
            continue
        sig = signature()
        if use_graph and step >= 2 and sig == prev_sig and numsteps - step >= 2:
            # two eager steps have left shapes and dtypes unchanged: capture the next step and replay it
            cand = _StepGraph(psi, lblocks, rblocks, one_step)
            if cand.ok:
                graph = cand
                continue
            use_graph = False
        prev_sig = sig
        one_step()
        step += 1
    if graph is not None:
        graph.finish()

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
