"""
METTS sampling (minimally entangled typical thermal states; Stoudenmire & White, New J. Phys. 12,
055026 (2010)) on matrix product states -- BASELINE config 5.

The reference ships only a dense 7-site exact-diagonalisation script (experiments/metts_ising.py); its
sample loop (`:59-105`) is re-expressed here on MPS: classical product state -> imaginary-time evolution
exp(-beta H / 2) by two-site TDVP (the hot path of this package) -> measurement -> collapse onto a new
product state in random local Bloch bases (`collapse_random_cps`, `:27-46`, same sequence of random
draws as the reference so a seed reproduces its outcomes).  Samples are independent: with several GPUs
every rank runs its own stream of samples (one process per GPU, no communication until the final gather
of the scalar estimates).
"""
import numpy as np
import torch

from . import _device as dev
from .mps import MPS
from .tdvp import tdvp_twosite
from .chain_ops import mpo_average

__all__ = ["random_bloch_basis", "product_state_mps", "collapse_random_cps", "metts_energy_samples"]


def random_bloch_basis(rng: np.random.Generator):
    """Uniformly random orthonormal Bloch basis as the columns of a 2 x 2 unitary
    (experiments/metts_ising.py:17-24; two `rng.uniform()` draws: theta, then phi)."""
    theta = np.arccos(2 * rng.uniform() - 1)
    phi = 2 * np.pi * rng.uniform()
    c, s, ph = np.cos(theta / 2), np.sin(theta / 2), np.exp(1j * phi)
    return np.array([[c, -s], [ph * s, ph * c]])


def product_state_mps(local_states, device=None):
    """MPS with bond dimension 1 from a list of local state vectors (all quantum numbers zero)."""
    d = len(local_states[0])
    n = len(local_states)
    psi = MPS(np.zeros(d, dtype=int), [np.zeros(1, dtype=int) for _ in range(n + 1)], fill="postpone", device=device)
    psi.a = [dev.to_device(np.asarray(v, dtype=complex).reshape(1, d, 1), psi.device) for v in local_states]
    return psi


def collapse_random_cps(psi: MPS, rng: np.random.Generator):
    """
    Sequentially collapse the (normalised) state `psi` onto a classical product state, drawing a random
    Bloch basis per site (experiments/metts_ising.py:27-46, on MPS tensors instead of the dense vector).
    Returns the list of chosen local states.  `psi` is right-orthonormalised in place first, so the norm
    of each projected boundary vector is the conditional probability amplitude.
    """
    assert len(psi.qsite) == 2, "random Bloch bases are defined for local dimension 2"
    psi.orthonormalize(mode="right")
    left = torch.ones((1, 1), dtype=dev.C128, device=psi.device)
    states = []
    for a in psi.a:
        dl, d, dr = a.shape
        u = random_bloch_basis(rng)
        t = dev.gemm(left, dev.as_dtype(a, True).reshape(dl, d * dr)).reshape(d, dr)        # boundary . A
        chi = dev.gemm(torch.from_numpy(np.ascontiguousarray(u.conj().T)).to(psi.device), t)  # (2, dr)
        p = torch.linalg.norm(chi, dim=1).cpu().numpy()
        pick = 0 if rng.uniform() < p[0] ** 2 else 1
        states.append(u[:, pick].copy())
        left = (chi[pick] / p[pick]).reshape(1, dr)
    assert left.shape == (1, 1)
    return states


def metts_energy_samples(hamiltonian, beta, nsamples, rng, numsteps=20, numiter_lanczos=10, tol_split=1e-10,
                         start=None, observable=None, stats=None):
    """
    Run a chain of `nsamples` METTS and return the per-sample energies <phi|H|phi> (or
    <phi|observable|phi> when an MPO `observable` is given) as a NumPy array.

    Each sample: |phi> = exp(-beta H / 2)|cps> / norm  by `numsteps` two-site TDVP steps with real time
    step beta / (2 numsteps), then the next |cps> is drawn by `collapse_random_cps(phi)`.  A dict passed as
    `stats` collects the realised bond dimensions per sample.
    """
    nsites = hamiltonian.nsites
    if start is None:
        start = [random_bloch_basis(rng)[:, 0] for _ in range(nsites)]
    cps = start
    op = hamiltonian if observable is None else observable
    values = np.zeros(nsamples, dtype=complex)
    for n in range(nsamples):
        phi = product_state_mps(cps, device=hamiltonian.device)
        tdvp_twosite(hamiltonian, phi, 0.5 * beta / numsteps, numsteps, numiter_lanczos=numiter_lanczos,
                     tol_split=tol_split)
        if stats is not None:                    # realised bond dimensions of the thermal state |phi>
            stats.setdefault("max_bond", []).append(int(max(phi.bond_dims)))
            stats.setdefault("mean_bond", []).append(float(np.mean(phi.bond_dims)))
        phi.orthonormalize(mode="left")          # normalise: drop the norm accumulated by exp(-beta H / 2)
        values[n] = mpo_average(phi, op)
        cps = collapse_random_cps(phi, rng)
    return values
