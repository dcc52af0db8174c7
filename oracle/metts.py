"""Dense restatement of the reference's METTS helpers (TEST INFRASTRUCTURE, see oracle/__init__.py):
experiments/metts_ising.py:17-46 on full state vectors."""
import numpy as np


def random_bloch_basis(rng):
    """Restates experiments/metts_ising.py:17-24."""
    theta = np.arccos(2 * rng.uniform() - 1)
    phi = 2 * np.pi * rng.uniform()
    return np.array([[np.cos(theta / 2), -np.sin(theta / 2)],
                     [np.exp(1j * phi) * np.sin(theta / 2), np.exp(1j * phi) * np.cos(theta / 2)]])


def collapse_random_cps(nsites, psi, rng):
    """Restates experiments/metts_ising.py:27-46; returns the list of chosen local states."""
    states = []
    for _ in range(nsites):
        u = random_bloch_basis(rng)
        chi = u.conj().T @ np.reshape(psi, (2, -1))
        p = (np.linalg.norm(chi[0]), np.linalg.norm(chi[1]))
        if rng.uniform() < p[0] ** 2:
            states.append(u[:, 0]); psi = chi[0] / p[0]
        else:
            states.append(u[:, 1]); psi = chi[1] / p[1]
    assert len(psi) == 1
    return states
