"""GPU: the four sweep drivers (device-resident) against the golden fixtures generated from
the reference and against exact results, mirroring the reference's own tests
(test/test_tdvp.py:7-119, test/test_dmrg.py:5-95).

Tolerances: energies / observables 1e-10 (north_star); sector layouts (`qbonds`) and bond
dimensions bit-exact; state vectors 1e-9 (gauge-invariant full vectors after many local steps)."""
import copy
import os

import numpy as np
import pytest
import torch
from scipy.linalg import expm

pytestmark = pytest.mark.gpu


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def load_mps(ptb, z, tag, n):
    return ptb.MPS.from_tensors(z[f"{tag}/qsite"], [z[f"{tag}/qb{i}"] for i in range(n + 1)],
                                [z[f"{tag}/a{i}"] for i in range(n)])


def load_mpo(ptb, z, tag="h"):
    n = int(z[f"{tag}/nsites"])
    return ptb.MPO.from_tensors(z[f"{tag}/qsite"], [z[f"{tag}/qb{i}"] for i in range(n + 1)],
                                [z[f"{tag}/w{i}"] for i in range(n)]), n


def test_tdvp_readme_config(cuda_lib, golden_dir):
    """BASELINE config 1 (README.rst:18-48): XXZ L=10, tdvp_singlesite dt=0.01-0.05j, k=5."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_L10.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    nrm = ptb.tdvp_singlesite(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=int(z["k"]))
    assert abs(nrm - float(z["single/nrm"])) < 1e-12
    assert all(t.is_cuda for t in psi.a)
    assert rel(psi.to_vector(), z["single/vec"]) < 1e-9
    en = ptb.mpo_average(psi, h)
    assert abs(en - complex(z["single/energy"])) < 1e-10
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_twosite(h, psi, complex(z["dt"]), int(z["two/nsteps"]), numiter_lanczos=int(z["two/k"]),
                     tol_split=float(z["two/tol"]))
    assert psi.bond_dims == list(z["two/bond_dims"])
    assert rel(psi.to_vector(), z["two/vec"]) < 1e-9


@pytest.mark.parametrize("fixture,steps_key", [("tdvp_xxz_L10.npz", "nsteps"), ("tdvp_xxz_qnum_L8.npz", "nsteps")])
def test_tdvp_singlesite_cuda_graph_equals_eager(cuda_lib, golden_dir, monkeypatch, fixture, steps_key):
    """Launch-latency regime: from the third time step on `tdvp_singlesite` replays one captured CUDA graph per time
    step.  Same kernels in the same order: the final state is bit-identical to the eager run, equals the reference
    fixture, and the graph really was used (with and without quantum numbers)."""
    import pytenet_b200 as ptb
    from pytenet_b200 import tdvp
    z = np.load(os.path.join(golden_dir, fixture))
    h, n = load_mpo(ptb, z)
    k = int(z["k"]) if "k" in z else 5
    nsteps = int(z[steps_key])
    used = []
    real_init = tdvp._StepGraph.__init__

    def spy(self, *a, **kw):
        real_init(self, *a, **kw)
        used.append(self.ok)
    monkeypatch.setattr(tdvp._StepGraph, "__init__", spy)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setattr(tdvp, "_GRAPHS", mode)
        psi = load_mps(ptb, z, "psi0", n)
        nrm = ptb.tdvp_singlesite(h, psi, complex(z["dt"]), nsteps, numiter_lanczos=k)
        out[mode] = (nrm, psi.to_vector(), [np.array(q) for q in psi.qbonds])
    assert used == [True], "the graph path must have been taken exactly once (mode 1)"
    assert out["0"][0] == out["1"][0]
    assert np.array_equal(out["0"][1], out["1"][1])
    assert all(np.array_equal(x, y) for x, y in zip(out["0"][2], out["1"][2]))
    assert rel(out["1"][1], z["single/vec"]) < 1e-9


def test_tdvp_quantum_numbers_bit_exact_sectors(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_qnum_L8.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_twosite(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=10)
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"two/qb{i}"]), f"sector layout of bond {i}"
    assert rel(psi.to_vector(), z["two/vec"]) < 1e-9
    for i in range(n):
        assert ptb.is_qsparse(psi.a[i], [psi.qbonds[i], psi.qsite, -psi.qbonds[i + 1]])
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_singlesite(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=5)
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"single/qb{i}"])
    assert rel(psi.to_vector(), z["single/vec"]) < 1e-9


def test_dmrg_notebook_known_answer(cuda_lib, golden_dir):
    """doc/dmrg.ipynb:130,140: e0 = -18.48435890403327, bond dims [1,4,16,30,16,4,1]."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "dmrg_fermi_hubbard_L6.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    assert psi.a[0].dtype == torch.float64          # dtype="real" start (test_dmrg.py:24 pattern)
    en = ptb.dmrg_twosite(h, psi, 4, tol_split=1e-8)
    assert abs(en[-1] - (-18.48435890403327)) < 1e-10
    assert np.max(np.abs(en - z["two/en"])) < 1e-10
    assert psi.bond_dims == [1, 4, 16, 30, 16, 4, 1]
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"two/qb{i}"])
    assert abs(np.linalg.norm(psi.to_vector()) - 1) < 1e-12
    psi = load_mps(ptb, z, "psi1", n)
    en1 = ptb.dmrg_singlesite(h, psi, 3)
    assert np.max(np.abs(en1 - z["single/en"])) < 1e-10


def test_basics_notebook_known_answer(cuda_lib, golden_dir):
    """doc/basics.ipynb:256,530."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "basics_notebook.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    nrm = psi.orthonormalize(mode="left")
    assert abs(nrm - 0.008359386283800499) < 1e-15
    avg = ptb.mpo_average(psi, h)
    assert abs(avg.real - 9.188269028617416) < 1e-12 and abs(avg.imag) < 1e-12


def test_same_seed_gives_reference_tensors(cuda_lib, golden_dir):
    """MPS(..., fill="random", rng) draws the reference's exact tensors for a seed (mps.py:50-71)."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "basics_notebook.npz"))
    rng = np.random.default_rng(42)
    b = [1, 4, 15, 13, 7, 1]
    psi = ptb.MPS(np.zeros(3, dtype=int), [np.zeros(bi, dtype=int) for bi in b], fill="random", rng=rng)
    for i in range(5):
        assert np.array_equal(psi.a[i].cpu().numpy(), z[f"psi0/a{i}"])
    z = np.load(os.path.join(golden_dir, "dmrg_fermi_hubbard_L6.npz"))
    rng = np.random.default_rng(42)
    psi = ptb.MPS.construct_random(6, z["h/qsite"], ptb.encode_quantum_number_pair(7, 1), max_vdim=18,
                                   dtype="real", rng=rng)
    for i in range(6):
        assert np.array_equal(psi.a[i].cpu().numpy(), z[f"psi0/a{i}"])
        assert np.array_equal(psi.qbonds[i], z[f"psi0/qb{i}"])


def test_tdvp_against_exact_evolution(cuda_lib):
    """reference test_tdvp.py:7-75 with a seed and this package's own input builder."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(17)
    L = 8
    dt = 0.02 - 0.05j
    nsteps = 8
    h = ptb.heisenberg_xxz_1d_mpo(L, 4.0 / 3, 5.0 / 13, -2.0 / 7)
    spin_tot = 1
    qbonds = [np.array([0])]
    for _ in range(L - 1):
        qbonds.append(np.sort(np.array([q + h.qsite for q in qbonds[-1]]).reshape(-1)))
    qbonds.append(np.array([2 * spin_tot]))
    psi = ptb.MPS(h.qsite, qbonds, fill="random", rng=rng)
    psi.orthonormalize(mode="left")
    psi.orthonormalize(mode="right")
    for i in range(L):
        psi.a[i][6:, :, :] = 0
        psi.a[i][:, :, 6:] = 0
    psi.orthonormalize(mode="left")
    assert psi.qbonds[-1][0] == 2 * spin_tot
    ref = expm(-dt * nsteps * h.to_matrix()) @ psi.to_vector()
    p1 = copy.deepcopy(psi); p2 = copy.deepcopy(psi)
    ptb.tdvp_singlesite(h, p1, dt, nsteps, numiter_lanczos=5)
    ptb.tdvp_twosite(h, p2, dt, nsteps, numiter_lanczos=10)
    assert np.allclose(p1.to_vector(), ref, atol=2e-5)
    assert np.allclose(p2.to_vector(), ref, atol=1e-10)


def test_tdvp_time_reversal_symmetry(cuda_lib):
    """reference test_tdvp.py:78-119."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(23)
    L = 8
    h = ptb.heisenberg_xxz_1d_mpo(L, 4.0 / 3, 5.0 / 13, -2.0 / 7).zero_qnumbers()
    qbonds = [np.array([0])] + [np.zeros(5, dtype=int) for _ in range(L - 1)] + [np.array([0])]
    psi = ptb.MPS(h.qsite, qbonds, fill="random", rng=rng)
    psi.orthonormalize(mode="left")
    ref = psi.to_vector()
    p1 = copy.deepcopy(psi)
    ptb.tdvp_singlesite(h, p1, 0.5j, 1, numiter_lanczos=10)
    ptb.tdvp_singlesite(h, p1, -0.5j, 1, numiter_lanczos=10)
    assert np.allclose(p1.to_vector(), ref, atol=1e-10)
    p2 = copy.deepcopy(psi)
    ptb.tdvp_twosite(h, p2, 0.5j, 1, numiter_lanczos=10, tol_split=1e-10)
    ptb.tdvp_twosite(h, p2, -0.5j, 1, numiter_lanczos=10, tol_split=1e-10)
    assert np.allclose(p2.to_vector(), ref, atol=1e-6)


def test_dmrg_against_exact_diagonalisation(cuda_lib):
    """reference test_dmrg.py:5-95 (real start vector, bond growth in the two-site variant)."""
    import pytenet_b200 as ptb
    L = 8
    h = ptb.heisenberg_xxz_1d_mpo(L, 4 / 5, 8 / 3, -2 / 7)
    hm = h.to_matrix()
    en_ref, v_ref = np.linalg.eigh(hm)
    rng = np.random.default_rng(31)
    psi = ptb.MPS.construct_random(L, h.qsite, 0, max_vdim=32, dtype="real", rng=rng)
    e0 = ptb.dmrg_singlesite(h, psi, 4)[-1]
    # lowest state within the Sz = 0 sector
    sz = np.array([sum(1 if (i >> b) & 1 == 0 else -1 for b in range(L)) for i in range(2 ** L)])
    sector = np.where(sz == 0)[0]
    e_sec = np.linalg.eigvalsh(hm[np.ix_(sector, sector)])[0]
    assert abs(e0 - e_sec) < 1e-12
    vec = psi.to_vector()
    assert np.allclose(hm @ vec, e0 * vec, atol=1e-7)
    rng = np.random.default_rng(32)
    psi = ptb.MPS.construct_random(L, h.qsite, 2, max_vdim=6, dtype="complex", rng=rng)
    e2 = ptb.dmrg_twosite(h, psi, 3)[-1]
    sector = np.where(sz == 2)[0]
    assert abs(e2 - np.linalg.eigvalsh(hm[np.ix_(sector, sector)])[0]) < 1e-12
    assert max(psi.bond_dims) > 6          # bonds must have grown


@pytest.mark.parametrize("segmented", [True, False])
def test_sector_banded_matvec_matches_dense(cuda_lib, monkeypatch, segmented):
    """Block-sparse path: HeffSectorPlan.apply == dense device matvec == oracle on block-sparse inputs
    (sorted and unsorted bonds), and it actually skips work for sorted sectors.  Step 3 both as one segmented
    launch (default) and as one accumulating banded launch per left MPO index."""
    import oracle
    import oracle.blocksparse as ob
    import pytenet_b200 as ptb
    from pytenet_b200 import sectors
    from pytenet_b200.sectors import HeffSectorPlan
    monkeypatch.setattr(sectors, "_SEGMENTED", segmented)
    rng = np.random.default_rng(77)
    for (Dl, d, Dr, cl, cr, sort) in [(300, 2, 260, 5, 5, True), (150, 4, 330, 6, 6, True), (200, 3, 129, 4, 5, False)]:
        qs = rng.integers(-1, 2, size=d)
        ql = rng.integers(-2, 3, size=Dl); qr = rng.integers(-2, 3, size=Dr)
        if sort:
            ql = np.sort(ql); qr = np.sort(qr)
        qwl = rng.integers(-1, 2, size=cl); qwr = rng.integers(-1, 2, size=cr)
        a = rng.normal(size=(Dl, d, Dr)) + 1j * rng.normal(size=(Dl, d, Dr)); ob.enforce_qsparsity(a, [ql, qs, -qr])
        l = rng.normal(size=(Dl, cl, Dl)) + 1j * rng.normal(size=(Dl, cl, Dl)); ob.enforce_qsparsity(l, [ql, qwl, -ql])
        r = rng.normal(size=(Dr, cr, Dr)) + 1j * rng.normal(size=(Dr, cr, Dr)); ob.enforce_qsparsity(r, [qr, qwr, -qr])
        w = rng.normal(size=(cl, d, d, cr)); ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr])
        plan = HeffSectorPlan(ql, qs, qr, qwl, qwr, cplx=True)
        cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
        got = plan.apply(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        dense = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
        assert rel(got, ref) < 1e-12 and rel(got, dense) < 1e-13
        # forbidden entries of the output stay exactly zero (SURVEY.md headline 3)
        mask = ob.qnumber_outer_sum([ql, qs, -qr]) != 0
        assert np.all(got[mask] == 0)


def test_sector_plan_reuse_keeps_structural_zeros(cuda_lib):
    """The lean sector path writes the structural zeros of t1 / t2 once per plan and afterwards only touches what
    the quantum numbers allow: repeated applications with new vectors, another plan in between (same workspace)
    and a different MPO tensor on the same plan must all still equal the dense device matvec."""
    import oracle.blocksparse as ob
    import pytenet_b200 as ptb
    from pytenet_b200.sectors import HeffSectorPlan
    rng = np.random.default_rng(5)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()

    def make(Dl, d, Dr, cl, cr, nsec):
        qs = rng.integers(-1, 2, size=d)
        ql = np.sort(rng.integers(-nsec, nsec + 1, size=Dl)); qr = np.sort(rng.integers(-nsec, nsec + 1, size=Dr))
        qwl = rng.integers(-1, 2, size=cl); qwr = rng.integers(-1, 2, size=cr)
        def tensor(shape, qn):
            t = rng.normal(size=shape) + 1j * rng.normal(size=shape); ob.enforce_qsparsity(t, qn); return t
        vecs = [tensor((Dl, d, Dr), [ql, qs, -qr]) for _ in range(3)]
        l = tensor((Dl, cl, Dl), [ql, qwl, -ql]); r = tensor((Dr, cr, Dr), [qr, qwr, -qr])
        ws = []
        for dens in (0.5, 0.15):
            w = rng.normal(size=(cl, d, d, cr)); w[rng.random(w.shape) > dens] = 0
            ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr]); ws.append(w)
        return HeffSectorPlan(ql, qs, qr, qwl, qwr, cplx=True), vecs, ws, l, r

    p1, v1, w1, l1, r1 = make(700, 3, 650, 5, 4, 6)
    p2, v2, w2, l2, r2 = make(520, 2, 900, 4, 6, 5)
    seq = [(p1, v1[0], w1[0], l1, r1), (p1, v1[1], w1[0], l1, r1), (p2, v2[0], w2[0], l2, r2),
           (p1, v1[2], w1[0], l1, r1), (p1, v1[0], w1[1], l1, r1), (p2, v2[1], w2[1], l2, r2),
           (p2, v2[2], w2[1], l2, r2)]
    for plan, a, w, l, r in seq:
        wd = cu(w)
        got = plan.apply(cu(a), wd, cu(l), cu(r))
        want = ptb.apply_local_hamiltonian(cu(a), wd, cu(l), cu(r))
        assert (torch.linalg.norm(got - want) / torch.linalg.norm(want)).item() < 1e-13


def test_sweeps_with_forced_sector_plans(cuda_lib, golden_dir, monkeypatch):
    """The sweeps give the same energies / sector layouts when every local problem goes through the
    sector-banded matvec (forced; by default it is only used for bonds >= 256)."""
    import pytenet_b200 as ptb
    from pytenet_b200 import _sweep
    monkeypatch.setattr(_sweep, "_SECTOR_MODE", "1")
    z = np.load(os.path.join(golden_dir, "dmrg_fermi_hubbard_L6.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    en = ptb.dmrg_twosite(h, psi, 4, tol_split=1e-8)
    assert abs(en[-1] - (-18.48435890403327)) < 1e-10
    assert psi.bond_dims == [1, 4, 16, 30, 16, 4, 1]
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"two/qb{i}"])
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_qnum_L8.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_twosite(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=10)
    assert rel(psi.to_vector(), z["two/vec"]) < 1e-9
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_singlesite(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=5)
    assert rel(psi.to_vector(), z["single/vec"]) < 1e-9


def test_mps_norm_and_vdot(cuda_lib, golden_dir):
    """doc/basics.ipynb:256 (norm of the seed-42 random MPS) and reference test_chain_ops.py:5-29 pattern."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "basics_notebook.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    assert abs(ptb.mps_norm(psi) - 0.008359386283800499) < 1e-16
    z = np.load(os.path.join(golden_dir, "mpo_inner.npz"))
    n = int(z["nsites"])
    rng = np.random.default_rng(2)
    qd = np.zeros(3, dtype=int)
    a = ptb.MPS(qd, [np.zeros(b, int) for b in [1, 4, 9, 7, 3, 1]], fill="random", rng=rng)
    b = ptb.MPS(qd, [np.zeros(b, int) for b in [1, 3, 8, 5, 2, 1]], fill="random", rng=rng)
    assert abs(ptb.mps_vdot(b, a) - np.vdot(b.to_vector(), a.to_vector())) < 1e-15


def test_sector_plans_for_environment_updates_and_bond_contraction(cuda_lib):
    """EnvSectorPlan / BondSectorPlan against the oracle on block-sparse inputs (sorted and unsorted bonds)."""
    import oracle
    import oracle.blocksparse as ob
    from pytenet_b200.sectors import EnvSectorPlan, BondSectorPlan
    rng = np.random.default_rng(91)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()      # noqa: E731
    for (Dl, d, Dr, cl, cr, sort) in [(260, 2, 300, 5, 5, True), (140, 4, 200, 6, 4, True), (150, 3, 131, 4, 5, False)]:
        qs = rng.integers(-1, 2, size=d)
        ql = rng.integers(-2, 3, size=Dl); qr = rng.integers(-2, 3, size=Dr)
        if sort:
            ql = np.sort(ql); qr = np.sort(qr)
        qwl = rng.integers(-1, 2, size=cl); qwr = rng.integers(-1, 2, size=cr)
        crand = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)      # noqa: E731
        a = crand(Dl, d, Dr); ob.enforce_qsparsity(a, [ql, qs, -qr])
        l = crand(Dl, cl, Dl); ob.enforce_qsparsity(l, [ql, qwl, -ql])
        r = crand(Dr, cr, Dr); ob.enforce_qsparsity(r, [qr, qwr, -qr])
        w = rng.normal(size=(cl, d, d, cr)); ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr])
        plan = EnvSectorPlan(ql, qs, qr, qwl, qwr, cplx=True)
        got = plan.step_left(cu(a), cu(w), cu(l)).cpu().numpy()
        assert rel(got, oracle.contraction_operator_step_left(a, a, w, l)) < 1e-12, ("left", Dl, d, Dr)
        got = plan.step_right(cu(a), cu(w), cu(r)).cpu().numpy()
        assert rel(got, oracle.contraction_operator_step_right(a, a, w, r)) < 1e-12, ("right", Dl, d, Dr)
        # zero-site contraction between a bond of dimension Dl and its copy of dimension Dr
        qbl = np.sort(rng.integers(-2, 3, size=Dl)) if sort else rng.integers(-2, 3, size=Dl)
        qbr = np.sort(rng.integers(-2, 3, size=Dr)) if sort else rng.integers(-2, 3, size=Dr)
        c = crand(Dl, Dr); ob.enforce_qsparsity(c, [qbl, -qbr])
        lb = crand(Dl, cl, Dl); ob.enforce_qsparsity(lb, [qbl, qwl, -qbl])
        rb = crand(Dr, cl, Dr); ob.enforce_qsparsity(rb, [qbr, qwl, -qbr])
        bplan = BondSectorPlan(qbl, qbr, qwl, cplx=True)
        got = bplan.apply(cu(c), cu(lb), cu(rb)).cpu().numpy()
        assert rel(got, oracle.apply_local_bond_contraction(c, lb, rb)) < 1e-12, ("bond", Dl, Dr)


def test_sector_path_at_config3_shape(cuda_lib):
    """BASELINE config-3 shape (two-site Fermi-Hubbard, a (2048,16,2048), h2 (6,16,16,6), (N,Sz) sectors): the sector
    path (banded step 1, masked W step, segmented step 3, work-sorted schedules) equals the dense device matvec,
    forbidden entries stay exactly zero, and a second application with a new vector on the same plan (structural
    zeros of the intermediates not rewritten) is still exact."""
    import pytenet_b200 as ptb
    from pytenet_b200 import hamiltonian as ham
    from pytenet_b200.sectors import HeffSectorPlan
    D = 2048
    qsite, qb, wbulk, _, _ = ham._fermi_hubbard_bulk(1.0, 4.0, 0.0)
    qsite = np.array(qsite); qb = np.array(qb)
    qs2 = np.add.outer(qsite, qsite).reshape(-1)
    w2 = np.einsum("kpqm,mrsn->kprqsn", wbulk, wbulk).reshape(6, 16, 16, 6)
    cand = [(dn, ds) for dn in range(-4, 5) for ds in range(-4, 5) if (dn + ds) % 2 == 0]
    wts = np.array([np.exp(-(dn ** 2 + ds ** 2) / (2 * 1.6 ** 2)) for dn, ds in cand])
    sizes = np.floor(wts / wts.sum() * D).astype(int)
    sizes[np.argmax(sizes)] += D - sizes.sum()
    q = np.sort(np.concatenate([np.full(sz, ptb.encode_quantum_number_pair(32 + dn, ds))
                                for (dn, ds), sz in zip(cand, sizes)]))
    g = torch.Generator(device="cuda").manual_seed(3)

    def tensor(shape, qn):
        t = torch.randn(*shape, dtype=torch.complex128, device="cuda", generator=g)
        ptb.enforce_qsparsity(t, qn)
        return t

    l = tensor((D, 6, D), [q, qb, -q]); r = tensor((D, 6, D), [q, qb, -q])
    w = torch.from_numpy(w2).cuda()
    plan = HeffSectorPlan(q, qs2, q, qb, qb, cplx=True)
    for _ in range(2):
        a = tensor((D, 16, D), [q, qs2, -q])
        got = plan.apply(a, w, l, r)
        want = ptb.apply_local_hamiltonian(a, w, l, r)
        assert (torch.linalg.norm(got - want) / torch.linalg.norm(want)).item() < 1e-13
        mask = (torch.from_numpy(q).cuda()[:, None, None] + torch.from_numpy(qs2).cuda()[None, :, None]
                - torch.from_numpy(q).cuda()[None, None, :]) != 0
        assert not bool(torch.any((got != 0) & mask).item())
        del a, got, want, mask


@pytest.mark.parametrize("sector_mode", ["auto", "1"])
def test_truncating_quantum_number_twosite_runs_match_reference(cuda_lib, golden_dir, monkeypatch, sector_mode):
    """A TRUNCATING quantum-number run pinned against the reference: Fermi-Hubbard L = 8, (N, Sz)
    sectors, `MPS.construct_random` start with bonds up to 48 -- `tdvp_twosite` with tol_split = 1e-6 and `dmrg_twosite`
    with tol_split = 1e-8.  Bond dimensions and sector layouts bit-exact, state 1e-8, energies 1e-10; once on the
    default path and once with every local problem forced through the sector plans (packed matvec, packed zero-site,
    banded environment updates)."""
    import pytenet_b200 as ptb
    from pytenet_b200 import _sweep
    monkeypatch.setattr(_sweep, "_SECTOR_MODE", sector_mode)
    z = np.load(os.path.join(golden_dir, "tdvp_fh_qnum_trunc_L8.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    ptb.tdvp_twosite(h, psi, complex(z["tdvp/dt"]), int(z["tdvp/nsteps"]), numiter_lanczos=int(z["tdvp/k"]),
                     tol_split=float(z["tdvp/tol"]))
    assert psi.bond_dims == list(z["tdvp/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"tdvp/qb{i}"]), f"sector layout of bond {i}"
    assert rel(psi.to_vector(), z["tdvp/vec"]) < 1e-8
    psi = load_mps(ptb, z, "psi0", n)
    en = ptb.dmrg_twosite(h, psi, len(z["dmrg/en"]), numiter_lanczos=int(z["dmrg/k"]), tol_split=float(z["dmrg/tol"]))
    assert np.max(np.abs(en - z["dmrg/en"])) < 1e-10
    assert psi.bond_dims == list(z["dmrg/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"dmrg/qb{i}"])


def test_rank_deficient_tol0_twosite_runs_match_reference(cuda_lib, golden_dir):
    """tol_split = 0 from a PRODUCT state (VERDICT r1 next #1d): every first split is rank deficient; the reference
    (LAPACK rounding noise > 0) keeps those bond indices and so must this path (batched Jacobi SVD + cuSOLVER
    refactorisation of rank-deficient blocks).  Bond dimensions and sector layouts bit-exact, state / energies to
    tolerance."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "twosite_rank_deficient_L6.npz"))
    h, n = load_mpo(ptb, z)
    psi = load_mps(ptb, z, "psi0", n)
    assert psi.bond_dims == [1] * (n + 1)
    ptb.tdvp_twosite(h, psi, complex(z["tdvp/dt"]), int(z["tdvp/nsteps"]), numiter_lanczos=int(z["tdvp/k"]), tol_split=0)
    assert psi.bond_dims == list(z["tdvp/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"tdvp/qb{i}"])
    assert rel(psi.to_vector(), z["tdvp/vec"]) < 1e-9
    psi = load_mps(ptb, z, "psi0", n)
    en = ptb.dmrg_twosite(h, psi, len(z["dmrg/en"]), numiter_lanczos=int(z["dmrg/k"]), tol_split=0)
    assert psi.bond_dims == list(z["dmrg/bond_dims"])
    assert np.max(np.abs(en - z["dmrg/en"])) < 1e-10
