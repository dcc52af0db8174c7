"""GPU: MPS METTS sampling (config 5) against the dense algorithm of the reference's experiment
(experiments/metts_ising.py, restated in oracle/metts.py) on the experiment's own 7-site Ising model."""
import numpy as np
import pytest
import torch
from scipy.linalg import expm

import oracle.metts as om

pytestmark = pytest.mark.gpu

NS, J, H, G, BETA = 7, 1.0, 0.8, -0.375, 1.2


def test_collapse_matches_reference_draws(cuda_lib):
    """Same seed, same state -> the MPS collapse picks the same product state as the dense reference."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(11)
    bonds = [1, 2, 4, 8, 8, 4, 2, 1]
    psi = ptb.MPS(np.zeros(2, int), [np.zeros(b, int) for b in bonds], fill="random", rng=rng)
    psi.orthonormalize(mode="left")
    vec = psi.to_vector()
    got = ptb.collapse_random_cps(psi, np.random.default_rng(857))
    want = om.collapse_random_cps(NS, vec, np.random.default_rng(857))
    for g, w in zip(got, want):
        assert np.allclose(g, w, atol=1e-12)


def test_imaginary_time_evolution_and_thermal_energy(cuda_lib):
    import pytenet_b200 as ptb
    h = ptb.ising_1d_mpo(NS, J, H, G)
    hm = h.to_matrix()
    rho = expm(-0.5 * BETA * hm)
    # (1) exp(-beta H / 2)|cps> by two-site TDVP vs the dense matrix exponential
    rng = np.random.default_rng(3)
    cps = [ptb.random_bloch_basis(rng)[:, 0] for _ in range(NS)]
    phi = ptb.product_state_mps(cps)
    ptb.tdvp_twosite(h, phi, 0.5 * BETA / 40, 40, numiter_lanczos=10, tol_split=1e-12)
    v = phi.to_vector(); v /= np.linalg.norm(v)
    dense = np.array([1.0 + 0j])
    for c in cps:
        dense = np.kron(dense, c)
    ref = rho @ dense; ref /= np.linalg.norm(ref)
    assert abs(abs(np.vdot(ref, v)) - 1) < 1e-6
    # (2) METTS estimate of the thermal energy vs the exact value (statistical: 5 sigma of the sample mean)
    vals = ptb.metts_energy_samples(h, BETA, 120, np.random.default_rng(5), numsteps=12, numiter_lanczos=8)
    e_exact = np.trace(rho @ hm @ rho).real / np.trace(rho @ rho).real
    mean, err = vals.real.mean(), vals.real.std() / np.sqrt(len(vals))
    assert abs(mean - e_exact) < 5 * err + 0.02, (mean, e_exact, err)
    assert np.max(np.abs(vals.imag)) < 1e-10
