"""
Generate the committed golden fixtures in tests/golden/*.npz from the REAL
reference (cmendl/pytenet v1.3.0, imported read-only from /root/reference) and,
while doing so, check the oracle/ restatement against it.

Run in the dev container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Nothing in tests/ (other than this script), smoke() or bench.py reads
/root/reference at run time -- they read the .npz files written here.
"""
import copy
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import pytenet as ptn          # noqa: E402  (the real reference)
import oracle                  # noqa: E402
import oracle.blocksparse as ob  # noqa: E402
import oracle.sweeps as osw    # noqa: E402


def rel(x, y):
    x = np.asarray(x); y = np.asarray(y)
    n = np.linalg.norm(y)
    return np.linalg.norm(x - y) / (n if n > 0 else 1.0)


def rand(shape, rng, cplx=True):
    if cplx:
        return ptn.crandn(shape, rng)
    return rng.normal(size=shape)


# --------------------------------------------------------------------------
# 1. function-level fixtures for the four contractions
# --------------------------------------------------------------------------
def chain_ops_cases():
    rng = np.random.default_rng(20261017)
    out = {}
    # (name, Dl, d, Dr, chil, chir, Dl', Dr', complex a/l/r, complex w)
    shapes = [
        ("c_small",   3, 2, 5, 4, 3,  6,  7, True,  False),
        ("c_square", 16, 2, 16, 5, 5, 16, 16, True, False),
        ("c_wcplx",   7, 3, 6, 4, 5,  5,  8, True,  True),
        ("r_real",    9, 2, 11, 5, 5, 9, 11, False, False),
        ("c_edge",    1, 2, 2, 1, 4,  1,  2, True,  False),
        ("c_odd",    17, 4, 9, 6, 6, 13, 11, True, False),
        ("c_ragged", 34, 2, 67, 5, 3, 35, 65, True, False),
    ]
    for name, Dl, d, Dr, cl, cr, Dlp, Drp, cz, wz in shapes:
        a = rand((Dl, d, Dr), rng, cz)
        b = rand((Dlp, d, Drp), rng, cz)
        w = rand((cl, d, d, cr), rng, wz)
        w[rng.random(w.shape) < 0.6] = 0          # MPO tensors are sparse
        l = rand((Dl, cl, Dlp), rng, cz)
        r = rand((Dr, cr, Drp), rng, cz)
        c = rand((Dl, Dr), rng, cz)
        rb = rand((Dr, cl, Drp), rng, cz)         # bond contraction: same chi both sides
        ref = {
            "hv": ptn.apply_local_hamiltonian(a, w, l, r),
            "bond": ptn.apply_local_bond_contraction(c, l, rb),
            "sr": ptn.contraction_operator_step_right(a, b, w, r),
            "sl": ptn.contraction_operator_step_left(a, b, w, l),
        }
        orc = {
            "hv": oracle.apply_local_hamiltonian(a, w, l, r),
            "bond": oracle.apply_local_bond_contraction(c, l, rb),
            "sr": oracle.contraction_operator_step_right(a, b, w, r),
            "sl": oracle.contraction_operator_step_left(a, b, w, l),
        }
        for key in ref:
            e = rel(orc[key], ref[key])
            assert e < 1e-14, (name, key, e)
            assert orc[key].shape == ref[key].shape and orc[key].dtype == ref[key].dtype
        for key, val in dict(a=a, b=b, w=w, l=l, r=r, c=c, rb=rb).items():
            out[f"{name}/{key}"] = val
        for key, val in ref.items():
            out[f"{name}/ref_{key}"] = val
    out["names"] = np.array([s[0] for s in shapes])
    np.savez_compressed(os.path.join(HERE, "chain_ops.npz"), **out)
    print("chain_ops.npz: oracle == reference on", len(shapes), "cases")


# --------------------------------------------------------------------------
# 2. block-sparse case: reference test_chain_ops.py:32-65 (test_mpo_average)
#    restated with a seed; pins contraction_operator_step_right with a != b
# --------------------------------------------------------------------------
def mpo_inner_case():
    for seed in range(7, 200):
        if _mpo_inner_case(seed):
            return
    raise RuntimeError("no seed gave a non-vanishing inner product")


def _mpo_inner_case(seed):
    rng = np.random.default_rng(seed)
    d = 3
    qd = rng.integers(-1, 2, size=d)
    D = [1, 7, 26, 19, 25, 8, 1]
    qD = [rng.integers(-1, 2, size=Di) for Di in D]
    psi = ptn.MPS(qd, qD, fill="random", rng=rng)
    # chi: different bond dimensions, same leading/trailing sectors (a != b path)
    D2 = [1, 6, 21, 23, 17, 5, 1]
    qD2 = [rng.integers(-1, 2, size=Di) for Di in D2]
    qD2[0] = qD[0].copy(); qD2[-1] = qD[-1].copy()
    chi = ptn.MPS(qd, qD2, fill="random", rng=rng)
    for i in range(psi.nsites):
        psi.a[i] *= 5
        chi.a[i] *= 5
    DO = [1, 5, 16, 14, 17, 4, 1]
    qO = [rng.integers(-1, 2, size=Di) for Di in DO]
    # zero leading/trailing operator sectors to avoid a vanishing value (test_chain_ops.py:47-49)
    qO[0] = np.array([0]); qO[-1] = np.array([0])
    op = ptn.MPO(qd, qO, fill="random", rng=rng)
    for i in range(op.nsites):
        op.a[i] *= 5
    val = ptn.mpo_inner_product(chi, op, psi)
    if abs(val) < 1e-6 or abs(ptn.mpo_average(psi, op)) < 1e-6:
        return False
    ref_dense = np.vdot(chi.to_vector(), op.to_matrix() @ psi.to_vector())
    assert abs(val - ref_dense) / abs(ref_dense) < 1e-12
    avg = ptn.mpo_average(psi, op)
    # oracle replay
    t = np.identity(psi.a[-1].shape[2], dtype=psi.a[-1].dtype).reshape(psi.a[-1].shape[2], 1, -1)
    for i in reversed(range(psi.nsites)):
        t = oracle.contraction_operator_step_right(psi.a[i], chi.a[i], op.a[i], t)
    assert abs(t[0, 0, 0] - val) <= 1e-13 * max(1, abs(val)), (t[0, 0, 0], val)
    out = {"value": np.array(val), "average": np.array(avg), "nsites": np.array(psi.nsites)}
    for i in range(psi.nsites):
        out[f"psi{i}"] = psi.a[i]; out[f"chi{i}"] = chi.a[i]; out[f"op{i}"] = op.a[i]
    out["seed"] = np.array(seed)
    np.savez_compressed(os.path.join(HERE, "mpo_inner.npz"), **out)
    print("mpo_inner.npz: seed", seed, "<chi|op|psi> =", val, "<psi|op|psi> =", avg)
    return True


# --------------------------------------------------------------------------
# 3. Krylov fixtures (reference test_krylov.py shapes, seeded)
# --------------------------------------------------------------------------
def krylov_cases():
    rng = np.random.default_rng(11)
    n, k = 96, 24
    m = ptn.crandn((n, n), rng); m = 0.5 * (m + m.conj().T)
    v0 = ptn.crandn(n, rng)
    al, be, V = ptn.lanczos_iteration(lambda x: m @ x, v0, k)
    al2, be2, V2 = oracle.lanczos_iteration(lambda x: m @ x, v0, k)
    assert np.array_equal(al, al2) and np.array_equal(be, be2) and np.array_equal(V, V2)
    dt = 0.4 + 0.2j
    ex = ptn.expm_krylov(lambda x: m @ x, v0, dt, 12, hermitian=True)
    ex2 = oracle.expm_krylov(lambda x: m @ x, v0, dt, 12)
    assert rel(ex2, ex) < 1e-14
    ew, eu = ptn.eigh_krylov(lambda x: m @ x, v0, 30, 2)
    ew2, eu2 = oracle.eigh_krylov(lambda x: m @ x, v0, 30, 2)
    assert rel(ew2, ew) < 1e-14 and rel(eu2, eu) < 1e-13
    # breakdown case: rank-3 operator, k = 8 (krylov.py:44-50)
    q = np.linalg.qr(ptn.crandn((32, 3), rng))[0]
    low = (q * np.array([1.0, -2.0, 0.5])) @ q.conj().T
    vb = q @ ptn.crandn(3, rng)
    with warnings.catch_warnings(record=True) as wrec:
        warnings.simplefilter("always")
        alb, beb, Vb = ptn.lanczos_iteration(lambda x: low @ x, vb, 8)
    assert len(wrec) == 1 and "beta[2]" in str(wrec[0].message), [str(x.message) for x in wrec]
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        alb2, beb2, Vb2 = oracle.lanczos_iteration(lambda x: low @ x, vb, 8)
    assert np.array_equal(alb, alb2) and np.array_equal(beb, beb2) and Vb.shape == Vb2.shape == (32, 3)
    np.savez_compressed(os.path.join(HERE, "krylov.npz"), m=m, v0=v0, alpha=al, beta=be, V=V,
                        dt=np.array(dt), expm=ex, eig_w=ew, eig_u=eu,
                        low=low, vb=vb, alpha_b=alb, beta_b=beb, V_b=Vb)
    print("krylov.npz: oracle == reference (Lanczos bit-identical)")


# --------------------------------------------------------------------------
# 4. sweep-level fixtures
# --------------------------------------------------------------------------
def save_mps(out, tag, psi):
    for i, t in enumerate(psi.a):
        out[f"{tag}/a{i}"] = t
    for i, q in enumerate(psi.qbonds):
        out[f"{tag}/qb{i}"] = np.asarray(q)
    out[f"{tag}/qsite"] = np.asarray(psi.qsite)


def save_mpo(out, tag, op):
    for i, t in enumerate(op.a):
        out[f"{tag}/w{i}"] = t
    for i, q in enumerate(op.qbonds):
        out[f"{tag}/qb{i}"] = np.asarray(q)
    out[f"{tag}/qsite"] = np.asarray(op.qsite)
    out[f"{tag}/nsites"] = np.array(op.nsites)


def to_chain(psi):
    return osw.Chain(psi.a, psi.qsite, psi.qbonds)


def readme_tdvp_case():
    """BASELINE config 1: README.rst:18-48 (XXZ L=10, J=1, D=0.8, h=-0.1), D<=28 clamped
    to 8, tdvp_singlesite dt=0.01-0.05j k=5; zero quantum numbers."""
    L = 10
    h = ptn.heisenberg_xxz_1d_mpo(L, 1.0, 0.8, -0.1)
    h.zero_qnumbers()
    D = [1, 2, 4, 8, 16, 28, 16, 8, 4, 2, 1]
    rng = np.random.default_rng(42)
    psi = ptn.MPS(h.qsite, [np.zeros(Di, dtype=int) for Di in D], fill="random", rng=rng)
    for i in range(L):
        psi.a[i][8:, :, :] = 0
        psi.a[i][:, :, 8:] = 0
    psi.orthonormalize(mode="left")
    out = {}
    save_mpo(out, "h", h)
    save_mps(out, "psi0", psi)
    dt = 0.01 - 0.05j
    nsteps = 20
    p1 = copy.deepcopy(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        nrm = ptn.tdvp_singlesite(h, p1, dt, nsteps, numiter_lanczos=5)
        o1 = to_chain(psi)
        nrm_o = osw.tdvp_singlesite(h.a, h.qbonds, o1, dt, nsteps, numiter_lanczos=5)
    v_ref = p1.to_vector()
    assert abs(nrm - nrm_o) < 1e-13 and rel(o1.to_vector(), v_ref) < 1e-11, rel(o1.to_vector(), v_ref)
    out["dt"] = np.array(dt); out["nsteps"] = np.array(nsteps); out["k"] = np.array(5)
    out["single/vec"] = v_ref; out["single/nrm"] = np.array(nrm)
    out["single/energy"] = np.array(ptn.mpo_average(p1, h))
    p2 = copy.deepcopy(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ptn.tdvp_twosite(h, p2, dt, 4, numiter_lanczos=10, tol_split=1e-10)
        o2 = to_chain(psi)
        osw.tdvp_twosite(h.a, h.qbonds, o2, dt, 4, numiter_lanczos=10, tol_split=1e-10)
    assert rel(o2.to_vector(), p2.to_vector()) < 1e-10 and o2.bond_dims == p2.bond_dims
    out["two/vec"] = p2.to_vector(); out["two/nsteps"] = np.array(4); out["two/k"] = np.array(10)
    out["two/tol"] = np.array(1e-10); out["two/bond_dims"] = np.array(p2.bond_dims)
    np.savez_compressed(os.path.join(HERE, "tdvp_xxz_L10.npz"), **out)
    print("tdvp_xxz_L10.npz: oracle sweeps == reference; bond dims", p2.bond_dims)


def tdvp_qnumber_case():
    """Reference test_tdvp.py:7-75 restated with a seed: XXZ L=8 with quantum
    numbers (total Sz sector), exercising block-sparse QR/SVD inside the sweeps."""
    L = 8
    h = ptn.heisenberg_xxz_1d_mpo(L, 4.0 / 3, 5.0 / 13, -2.0 / 7)
    spin_tot = 1
    qbonds = [np.array([0])]
    for _ in range(L - 1):
        qbonds.append(np.sort(np.array([q + h.qsite for q in qbonds[-1]]).reshape(-1)))
    qbonds.append(np.array([2 * spin_tot]))
    rng = np.random.default_rng(5)
    psi = ptn.MPS(h.qsite, qbonds, fill="random", rng=rng)
    psi.orthonormalize(mode="left")
    psi.orthonormalize(mode="right")
    for i in range(L):
        psi.a[i][6:, :, :] = 0
        psi.a[i][:, :, 6:] = 0
    psi.orthonormalize(mode="left")
    out = {}
    save_mpo(out, "h", h)
    save_mps(out, "psi0", psi)
    dt = 0.02 - 0.05j
    p1 = copy.deepcopy(psi); p2 = copy.deepcopy(psi)
    o1 = to_chain(psi); o2 = to_chain(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ptn.tdvp_singlesite(h, p1, dt, 6, numiter_lanczos=5)
        ptn.tdvp_twosite(h, p2, dt, 6, numiter_lanczos=10)
        osw.tdvp_singlesite(h.a, h.qbonds, o1, dt, 6, numiter_lanczos=5)
        osw.tdvp_twosite(h.a, h.qbonds, o2, dt, 6, numiter_lanczos=10)
    assert rel(o1.to_vector(), p1.to_vector()) < 1e-11
    assert rel(o2.to_vector(), p2.to_vector()) < 1e-10
    assert all(np.array_equal(x, y) for x, y in zip(o2.qbonds, p2.qbonds)), "sector layout must be bit-exact"
    out["dt"] = np.array(dt); out["nsteps"] = np.array(6)
    out["single/vec"] = p1.to_vector(); out["two/vec"] = p2.to_vector()
    for i, q in enumerate(p1.qbonds):
        out[f"single/qb{i}"] = np.asarray(q)
    for i, q in enumerate(p2.qbonds):
        out[f"two/qb{i}"] = np.asarray(q)
    np.savez_compressed(os.path.join(HERE, "tdvp_xxz_qnum_L8.npz"), **out)
    print("tdvp_xxz_qnum_L8.npz: oracle == reference; two-site bonds", p2.bond_dims)


def dmrg_notebook_case():
    """Seeded known answer of doc/dmrg.ipynb:130,140 (cells 4-10)."""
    L = 6
    h = ptn.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 1.5)
    sector = ptn.encode_quantum_number_pair(7, 1)
    rng = np.random.default_rng(42)
    psi = ptn.MPS.construct_random(L, h.qsite, sector, max_vdim=18, dtype="real", rng=rng)
    out = {}
    save_mpo(out, "h", h)
    save_mps(out, "psi0", psi)
    p = copy.deepcopy(psi); o = to_chain(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        en = ptn.dmrg_twosite(h, p, 4, tol_split=1e-8)
        en_o = osw.dmrg_twosite(h.a, h.qbonds, o, 4, tol_split=1e-8)
    assert abs(en[-1] - (-18.48435890403327)) < 1e-12, en[-1]
    assert p.bond_dims == [1, 4, 16, 30, 16, 4, 1]
    assert np.max(np.abs(en - en_o)) < 1e-12 and o.bond_dims == p.bond_dims
    assert all(np.array_equal(x, y) for x, y in zip(o.qbonds, p.qbonds))
    out["two/en"] = en; out["two/bond_dims"] = np.array(p.bond_dims)
    for i, q in enumerate(p.qbonds):
        out[f"two/qb{i}"] = np.asarray(q)
    out["notebook_e0"] = np.array(-18.48435890403327)
    out["ed_e0"] = np.array(-18.484358962762272)
    # single-site DMRG on the same model, real dtype (test_dmrg.py:5-49 pattern)
    rng = np.random.default_rng(43)
    psi1 = ptn.MPS.construct_random(L, h.qsite, sector, max_vdim=24, dtype="real", rng=rng)
    save_mps(out, "psi1", psi1)
    p = copy.deepcopy(psi1); o = to_chain(psi1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        en1 = ptn.dmrg_singlesite(h, p, 3)
        en1_o = osw.dmrg_singlesite(h.a, h.qbonds, o, 3)
    assert np.max(np.abs(en1 - en1_o)) < 1e-12
    out["single/en"] = en1
    np.savez_compressed(os.path.join(HERE, "dmrg_fermi_hubbard_L6.npz"), **out)
    print("dmrg_fermi_hubbard_L6.npz: e0 =", en[-1], "single-site", en1[-1])


def basics_notebook_case():
    """Seeded known answers of doc/basics.ipynb:32,256,530."""
    rng = np.random.default_rng(42)
    d = 3
    b = [1, 4, 15, 13, 7, 1]
    mps = ptn.MPS(np.zeros(d, dtype=int), [np.zeros(bi, dtype=int) for bi in b], fill="random", rng=rng)
    out = {}
    save_mps(out, "psi0", mps)
    nrm = mps.orthonormalize(mode="left")
    assert nrm == 0.008359386283800499, nrm
    h = ptn.bose_hubbard_1d_mpo(5, d, 1.0, 4.0, -0.5)
    save_mpo(out, "h", h)
    save_mps(out, "psi_left", mps)
    avg = ptn.mpo_average(mps, h)
    assert avg.real == 9.188269028617416, avg
    o = to_chain(type("X", (), {"a": [out[f"psi0/a{i}"] for i in range(5)], "qsite": mps.qsite,
                                "qbonds": [np.zeros(bi, dtype=int) for bi in b]})())
    assert abs(osw.orthonormalize_left(o) - nrm) < 1e-16
    out["norm"] = np.array(nrm); out["average"] = np.array(avg)
    np.savez_compressed(os.path.join(HERE, "basics_notebook.npz"), **out)
    print("basics_notebook.npz: norm", nrm, "average", avg)


def molecular_dmrg_case():
    """BASELINE config 4 at CPU scale: molecular_hamiltonian_mpo from random symmetric integrals
    (8 spin-less orbitals, optimize=False), dmrg_singlesite in the 4-particle sector."""
    rng = np.random.default_rng(2026)
    n = 8
    tkin = rng.normal(size=(n, n)); tkin = 0.5 * (tkin + tkin.T)
    vint = rng.normal(size=(n, n, n, n))
    vint = 0.5 * (vint + vint.transpose(1, 0, 3, 2))
    vint = 0.5 * (vint + vint.transpose(2, 3, 0, 1))
    h = ptn.molecular_hamiltonian_mpo(tkin, vint, optimize=False)
    assert np.allclose(h.to_matrix(), h.to_matrix().conj().T)
    out = {}
    save_mpo(out, "h", h)
    psi = ptn.MPS.construct_random(n, h.qsite, 4, max_vdim=20, dtype="complex", rng=rng)
    save_mps(out, "psi0", psi)
    p = copy.deepcopy(psi); o = to_chain(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        en = ptn.dmrg_singlesite(h, p, 3, numiter_lanczos=12)
        en_o = osw.dmrg_singlesite(h.a, h.qbonds, o, 3, numiter_lanczos=12)
    assert np.max(np.abs(en - en_o)) < 1e-10
    hm = h.to_matrix()
    nocc = np.array([bin(i).count("1") for i in range(2 ** n)])
    sec = np.where(nocc == 4)[0]
    e_ed = np.linalg.eigvalsh(hm[np.ix_(sec, sec)])[0]
    out["single/en"] = en; out["ed_e0_sector"] = np.array(e_ed); out["k"] = np.array(12)
    out["mpo_bond_dims"] = np.array(h.bond_dims)
    np.savez_compressed(os.path.join(HERE, "dmrg_molecular_N8.npz"), **out)
    print("dmrg_molecular_N8.npz: MPO bonds", h.bond_dims, "energies", en, "ED (sector)", e_ed)


def mps_ops_case():
    """apply_mpo, MPS.compress (svd both directions, density), mps_add and
    MPS.from_vector on a Heisenberg chain with quantum numbers (reference test_mps.py / test_chain_ops.py
    style, seeded); the oracle restatement (oracle/mps_ops.py) is checked against the reference here."""
    import oracle.mps_ops as om
    L = 6
    h = ptn.heisenberg_xxz_1d_mpo(L, 0.7, 1.1, 0.3)
    rng = np.random.default_rng(77)
    qbonds = [np.array([0])]
    for i in range(L - 1):
        cand = np.array([q + h.qsite for q in qbonds[-1]]).reshape(-1)
        qbonds.append(np.sort(cand)[rng.permutation(len(cand))[:min(len(cand), 7)]])
    qbonds.append(np.array([2]))
    psi = ptn.MPS(h.qsite, qbonds, fill="random", rng=rng)
    chi = ptn.MPS(h.qsite, qbonds, fill="random", rng=rng)
    out = {}
    save_mpo(out, "h", h); save_mps(out, "psi", psi); save_mps(out, "chi", chi)
    # apply_mpo
    hp = ptn.apply_mpo(h, psi)
    ohp = om.apply_mpo(h.a, h.qbonds, to_chain(psi))
    assert all(np.array_equal(x, y) for x, y in zip(hp.qbonds, ohp.qbonds))
    assert all(rel(x, y) < 1e-14 for x, y in zip(ohp.a, hp.a))
    save_mps(out, "hpsi", hp)
    out["hpsi/vec"] = hp.to_vector()
    # mps_add
    alpha = 0.3 - 0.8j
    sm = ptn.mps.mps_add(psi, chi, alpha)
    osm = om.mps_add(to_chain(psi), to_chain(chi), alpha)
    assert all(np.array_equal(x, y) for x, y in zip(sm.qbonds, osm.qbonds))
    assert all(np.array_equal(x, y) for x, y in zip(osm.a, sm.a))
    out["add/alpha"] = np.array(alpha); out["add/vec"] = sm.to_vector()
    save_mps(out, "add", sm)
    # compress: svd in both directions and density mode, with and without truncation
    for tag, tol, mode, direction in [("svd_l0", 0.0, "svd", "left"), ("svd_r0", 0.0, "svd", "right"),
                                      ("svd_l", 1e-3, "svd", "left"), ("svd_r", 1e-3, "svd", "right"),
                                      ("den", 1e-3, "density", "left"), ("den0", 0.0, "density", "left")]:
        p = copy.deepcopy(hp)
        nrm, scale = p.compress(tol, mode=mode, direction=direction)
        o = to_chain(copy.deepcopy(hp))
        onrm, oscale = (om.compress_svd(o, tol, direction) if mode == "svd" else om.compress_density(o, tol))
        assert abs(nrm - onrm) < 1e-12 * nrm and abs(scale - oscale) < 1e-12, (tag, nrm, onrm, scale, oscale)
        assert all(np.array_equal(x, y) for x, y in zip(p.qbonds, o.qbonds)), tag
        assert rel(o.to_vector(), p.to_vector()) < 1e-11, (tag, rel(o.to_vector(), p.to_vector()))
        out[f"cmp/{tag}/tol"] = np.array(tol); out[f"cmp/{tag}/nrm"] = np.array(nrm)
        out[f"cmp/{tag}/scale"] = np.array(scale); out[f"cmp/{tag}/vec"] = p.to_vector()
        out[f"cmp/{tag}/bond_dims"] = np.array(p.bond_dims)
        for i, q in enumerate(p.qbonds):
            out[f"cmp/{tag}/qb{i}"] = np.asarray(q)
    # from_vector (TT-SVD)
    v = ptn.crandn(3 ** 5, rng)
    for tag, tol in (("fv0", 0.0), ("fv", 1e-2)):
        m = ptn.MPS.from_vector(3, 5, v, tol=tol)
        o = om.from_vector(3, 5, v, tol=tol)
        assert m.bond_dims == o.bond_dims and rel(o.to_vector(), m.to_vector()) < 1e-12
        out[f"{tag}/tol"] = np.array(tol); out[f"{tag}/vec"] = m.to_vector(); out[f"{tag}/bond_dims"] = np.array(m.bond_dims)
    out["fv/input"] = v
    np.savez_compressed(os.path.join(HERE, "mps_ops.npz"), **out)
    print("mps_ops.npz: oracle == reference; H psi bond dims", hp.bond_dims)


if __name__ == "__main__":
    chain_ops_cases()
    mpo_inner_case()
    krylov_cases()
    readme_tdvp_case()
    tdvp_qnumber_case()
    dmrg_notebook_case()
    basics_notebook_case()
    molecular_dmrg_case()
    mps_ops_case()
    print("all golden fixtures written to", HERE)
