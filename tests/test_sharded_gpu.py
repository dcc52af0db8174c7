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


def _load(ptb, z, tag, n):
    return ptb.MPS.from_tensors(z[f"{tag}/qsite"], [z[f"{tag}/qb{i}"] for i in range(n + 1)],
                                [z[f"{tag}/a{i}"] for i in range(n)])


def test_sharded_dmrg_singlesite_molecular_fixture(cuda_lib, golden_dir):
    """BASELINE config 4 at CPU scale: molecular_hamiltonian_mpo (8 orbitals, MPO bonds up to 46) built by the
    reference, dmrg_singlesite energies from the reference -- through the unsharded device driver and through
    the MPO-bond-sharded driver (one rank here; 2 and 8 ranks in tools/sharded_dmrg_check.py)."""
    import os
    import pytenet_b200 as ptb
    from pytenet_b200.sharded_dmrg import dmrg_singlesite_sharded
    z = np.load(os.path.join(golden_dir, "dmrg_molecular_N8.npz"))
    n = int(z["h/nsites"])
    h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
    assert h.bond_dims == list(z["mpo_bond_dims"])
    psi = _load(ptb, z, "psi0", n)
    en = ptb.dmrg_singlesite(h, psi, 3, numiter_lanczos=int(z["k"]))
    assert np.max(np.abs(en - z["single/en"])) < 1e-10
    assert abs(en[-1] - float(z["ed_e0_sector"])) < 1e-9
    psi = _load(ptb, z, "psi0", n)
    en_s = dmrg_singlesite_sharded(h, psi, 3, numiter_lanczos=int(z["k"]))
    assert np.max(np.abs(en_s - z["single/en"])) < 1e-10
    assert abs(np.linalg.norm(psi.to_vector()) - 1) < 1e-12
    for i in range(n + 1):
        assert len(psi.qbonds[i]) == psi.bond_dims[i]


def test_sharded_tdvp_singlesite_matches_reference_fixture(cuda_lib, golden_dir):
    """tdvp_singlesite_sharded (one rank) on the README config and on the quantum-number XXZ fixture:
    same state vector as the reference's tdvp_singlesite."""
    import os
    import pytenet_b200 as ptb
    from pytenet_b200.sharded_dmrg import tdvp_singlesite_sharded
    for name, steps_key in (("tdvp_xxz_L10.npz", "nsteps"), ("tdvp_xxz_qnum_L8.npz", "nsteps")):
        z = np.load(os.path.join(golden_dir, name))
        n = int(z["h/nsites"])
        h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
        psi = _load(ptb, z, "psi0", n)
        k = int(z["k"]) if "k" in z else 5
        nrm = tdvp_singlesite_sharded(h, psi, complex(z["dt"]), int(z[steps_key]), numiter_lanczos=k)
        v = psi.to_vector()
        ref = z["single/vec"]
        assert np.linalg.norm(v - ref) / np.linalg.norm(ref) < 1e-9, name
        if "single/nrm" in z:
            assert abs(nrm - float(z["single/nrm"])) < 1e-12


def test_sharded_dmrg_cached_molecular_mpo_N10(cuda_lib, golden_dir):
    """Config-4 pipeline at CPU scale: the sparse MPO cache (same generator and format as the 32-orbital file) ->
    `load_cached_mpo` -> `dmrg_singlesite` and `dmrg_singlesite_sharded`; energies after each sweep equal the
    reference's `dmrg_singlesite` (stored in the cache file by tests/golden/make_molecular_mpo.py) to 1e-10."""
    import os
    import pytenet_b200 as ptb
    from pytenet_b200.hamiltonian import load_cached_mpo
    from pytenet_b200.sharded_dmrg import dmrg_singlesite_sharded
    path = os.path.join(golden_dir, "molecular_mpo_N10.npz")
    z = np.load(path)
    h = load_cached_mpo(path)
    n = h.nsites
    assert h.bond_dims == list(z["bond_dims"]) and n == 10

    def start():
        return ptb.MPS.from_tensors(h.qsite, [z[f"psi0_qb{i}"] for i in range(n + 1)],
                                    [z[f"psi0_a{i}"] for i in range(n)])
    k = int(z["dmrg_k"])
    ref = z["dmrg_single_en"]
    psi = start()
    en = ptb.dmrg_singlesite(h, psi, len(ref), numiter_lanczos=k)
    assert np.max(np.abs(en - ref)) < 1e-10
    psi = start()
    en_s = dmrg_singlesite_sharded(h, psi, len(ref), numiter_lanczos=k)
    assert np.max(np.abs(en_s - ref)) < 1e-10


def test_multi_gpu_sharded_dmrg_under_torchrun(cuda_lib):
    """Two ranks over NCCL (only on boxes with >= 2 GPUs; the driver's single-GPU test box skips it and the
    multi-rank equality check travels in bench.py's `sharded` block instead): tools/config4_sweep.py on the
    10-orbital cache must reproduce the reference's energies on every rank with zero drift of psi."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(root, "tools", "config4_sweep.py"), "--norb", "10"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "max_abs_err_vs_reference" in res.stdout
