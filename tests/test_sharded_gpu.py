"""GPU (single rank): the MPO-bond-sharded effective Hamiltonian on the CUDA engine reduces to the
plain matvec when there is one rank, and its shard packing is exact for ragged MPO bonds.
Multi-rank logic is covered on CPU/gloo in tests/test_sharded_cpu.py and on GPUs by
tools/sharded_check.py under torchrun."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def test_single_rank_sharded_matches_oracle(cuda_lib):
    from pytenet_b200.sharded import ShardedEffectiveHamiltonian
    rng = np.random.default_rng(4)
    for (Dl, d, Dr, cl, cr) in [(33, 2, 47, 19, 23), (64, 4, 64, 5, 5)]:
        a = rng.normal(size=(Dl, d, Dr)) + 1j * rng.normal(size=(Dl, d, Dr))
        l = rng.normal(size=(Dl, cl, Dl)) + 1j * rng.normal(size=(Dl, cl, Dl))
        r = rng.normal(size=(Dr, cr, Dr)) + 1j * rng.normal(size=(Dr, cr, Dr))
        w = rng.normal(size=(cl, d, d, cr)); w[rng.random(w.shape) < 0.8] = 0
        heff = ShardedEffectiveHamiltonian.from_full(w, torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
        out = heff.matvec(torch.from_numpy(a).cuda()).cpu().numpy()
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-12


def test_synthetic_shards_shapes(cuda_lib):
    from pytenet_b200.sharded import ShardedEffectiveHamiltonian
    h = ShardedEffectiveHamiltonian.synthetic(64, 2, 64, 21, 17, density=0.2, seed=3)
    x = torch.randn(64, 2, 64, dtype=torch.complex128, device="cuda")
    y = h.matvec(x)
    assert tuple(y.shape) == (64, 2, 64) and torch.isfinite(torch.view_as_real(y)).all()
    # linearity of the sharded operator
    z = h.matvec(3 * x)
    assert (torch.linalg.norm(z - 3 * y) / torch.linalg.norm(z)).item() < 1e-13


def test_precontracted_single_rank_matches_oracle(cuda_lib):
    """All-reduce-only variant: LW precontraction + two GEMMs (the second one with split-K)."""
    from pytenet_b200.sharded import PrecontractedShardedHamiltonian
    rng = np.random.default_rng(8)
    for (Dl, d, Dr, cl, cr, cw) in [(33, 2, 47, 19, 23, False), (64, 4, 64, 5, 5, False), (40, 2, 36, 7, 9, True)]:
        a = rng.normal(size=(Dl, d, Dr)) + 1j * rng.normal(size=(Dl, d, Dr))
        l = rng.normal(size=(Dl, cl, Dl)) + 1j * rng.normal(size=(Dl, cl, Dl))
        r = rng.normal(size=(Dr, cr, Dr)) + 1j * rng.normal(size=(Dr, cr, Dr))
        w = rng.normal(size=(cl, d, d, cr)) + (1j * rng.normal(size=(cl, d, d, cr)) if cw else 0)
        w[rng.random(w.shape) < 0.7] = 0
        heff = PrecontractedShardedHamiltonian.from_full(w, torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda())
        out = heff.matvec(torch.from_numpy(a).cuda()).cpu().numpy()
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-12, (Dl, d, Dr, cl, cr)
    h, setup_ms = PrecontractedShardedHamiltonian.synthetic(64, 2, 64, 21, 17, density=0.2, seed=3)
    x = torch.randn(64, 2, 64, dtype=torch.complex128, device="cuda")
    y = h.matvec(x)
    assert (torch.linalg.norm(h.matvec(3 * x) - 3 * y) / torch.linalg.norm(y)).item() < 1e-13
