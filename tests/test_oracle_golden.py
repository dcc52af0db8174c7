"""CPU: the oracle restatement against the committed golden fixtures that were
generated from the real reference (tests/golden/make_golden.py)."""
import os
import warnings

import numpy as np
import pytest

import oracle
import oracle.blocksparse as ob
import oracle.sweeps as osw


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def load_chain(z, tag, n):
    a = [z[f"{tag}/a{i}"] for i in range(n)]
    qb = [z[f"{tag}/qb{i}"] for i in range(n + 1)]
    return osw.Chain(a, z[f"{tag}/qsite"], qb)


def load_op(z, tag="h"):
    n = int(z[f"{tag}/nsites"])
    return [z[f"{tag}/w{i}"] for i in range(n)], [z[f"{tag}/qb{i}"] for i in range(n + 1)], n


def test_contractions_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "chain_ops.npz"))
    for name in z["names"]:
        g = lambda k: z[f"{name}/{k}"]      # noqa: E731
        assert rel(oracle.apply_local_hamiltonian(g("a"), g("w"), g("l"), g("r")), g("ref_hv")) < 1e-14
        assert rel(oracle.apply_local_bond_contraction(g("c"), g("l"), g("rb")), g("ref_bond")) < 1e-14
        assert rel(oracle.contraction_operator_step_right(g("a"), g("b"), g("w"), g("r")), g("ref_sr")) < 1e-14
        assert rel(oracle.contraction_operator_step_left(g("a"), g("b"), g("w"), g("l")), g("ref_sl")) < 1e-14


def test_mpo_inner_product_block_sparse(golden_dir):
    z = np.load(os.path.join(golden_dir, "mpo_inner.npz"))
    n = int(z["nsites"])
    D = z[f"psi{n-1}"].shape[2]
    t = np.identity(D, dtype=complex).reshape(D, 1, D)
    for i in reversed(range(n)):
        t = oracle.contraction_operator_step_right(z[f"psi{i}"], z[f"chi{i}"], z[f"op{i}"], t)
    assert t.shape == (1, 1, 1)
    assert abs(t[0, 0, 0] - z["value"]) / abs(z["value"]) < 1e-12


def test_lanczos_bit_identical(golden_dir):
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    m = z["m"]
    al, be, V = oracle.lanczos_iteration(lambda x: m @ x, z["v0"], len(z["alpha"]))
    assert np.array_equal(al, z["alpha"]) and np.array_equal(be, z["beta"]) and np.array_equal(V, z["V"])
    ex = oracle.expm_krylov(lambda x: m @ x, z["v0"], complex(z["dt"]), 12)
    assert rel(ex, z["expm"]) < 1e-14
    ew, eu = oracle.eigh_krylov(lambda x: m @ x, z["v0"], 30, 2)
    assert rel(ew, z["eig_w"]) < 1e-14 and rel(eu, z["eig_u"]) < 1e-13


def test_lanczos_breakdown(golden_dir):
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    low = z["low"]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        al, be, V = oracle.lanczos_iteration(lambda x: low @ x, z["vb"], 8)
    assert any("beta[2]" in str(w.message) for w in rec)
    assert np.array_equal(al, z["alpha_b"]) and np.array_equal(be, z["beta_b"]) and V.shape == z["V_b"].shape


def test_tdvp_readme_config(golden_dir):
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_L10.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    nrm = osw.tdvp_singlesite(w, wq, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=int(z["k"]))
    assert abs(nrm - float(z["single/nrm"])) < 1e-13
    assert rel(psi.to_vector(), z["single/vec"]) < 1e-10
    psi = load_chain(z, "psi0", n)
    osw.tdvp_twosite(w, wq, psi, complex(z["dt"]), int(z["two/nsteps"]), numiter_lanczos=int(z["two/k"]),
                     tol_split=float(z["two/tol"]))
    assert psi.bond_dims == list(z["two/bond_dims"])
    assert rel(psi.to_vector(), z["two/vec"]) < 1e-10


def test_tdvp_quantum_numbers(golden_dir):
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_qnum_L8.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    osw.tdvp_twosite(w, wq, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=10)
    assert rel(psi.to_vector(), z["two/vec"]) < 1e-10
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"two/qb{i}"])       # sector layout bit-exact
    psi = load_chain(z, "psi0", n)
    osw.tdvp_singlesite(w, wq, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=5)
    assert rel(psi.to_vector(), z["single/vec"]) < 1e-10


def test_dmrg_notebook_known_answer(golden_dir):
    """doc/dmrg.ipynb:130,140: e0 = -18.48435890403327, bonds [1,4,16,30,16,4,1]."""
    z = np.load(os.path.join(golden_dir, "dmrg_fermi_hubbard_L6.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    en = osw.dmrg_twosite(w, wq, psi, 4, tol_split=1e-8)
    assert abs(en[-1] - (-18.48435890403327)) < 1e-12
    assert psi.bond_dims == [1, 4, 16, 30, 16, 4, 1]
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"two/qb{i}"])
    assert np.max(np.abs(en - z["two/en"])) < 1e-12
    psi = load_chain(z, "psi1", n)
    en1 = osw.dmrg_singlesite(w, wq, psi, 3)
    assert np.max(np.abs(en1 - z["single/en"])) < 1e-12


def test_basics_notebook_known_answer(golden_dir):
    """doc/basics.ipynb:256,530: norm 0.008359386283800499, <H> = 9.188269028617416."""
    z = np.load(os.path.join(golden_dir, "basics_notebook.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    nrm = osw.orthonormalize_left(psi)
    assert abs(nrm - 0.008359386283800499) < 1e-17
    D = psi.a[-1].shape[2]
    t = np.identity(D, dtype=complex).reshape(D, 1, D)
    for i in reversed(range(n)):
        t = oracle.contraction_operator_step_right(psi.a[i], psi.a[i], w[i], t)
    assert abs(t[0, 0, 0].real - 9.188269028617416) < 1e-13


def test_retained_indices_and_sectors():
    rng = np.random.default_rng(3)
    s = np.array([0.5, 0.1, 0.7, 1e-9, 0.2])
    assert list(ob.retained_bond_indices(s, 0.0)) == [0, 1, 2, 3, 4]
    assert list(ob.retained_bond_indices(s, 1e-12)) == [0, 1, 2, 4]
    assert list(ob.retained_bond_indices(np.zeros(3), 0.1)) == []
    q0 = rng.integers(-1, 2, size=12); q1 = rng.integers(-1, 2, size=9)
    a = rng.normal(size=(12, 9)); ob.enforce_qsparsity(a, [q0, -q1])
    u, sv, v, qb = ob.block_sparse_svd(a, q0, q1)
    assert rel((u * sv) @ v, a) < 1e-13
    assert ob.is_qsparse(u, [q0, -qb]) and ob.is_qsparse(v, [qb, -q1])
    assert np.all(np.diff(qb) >= 0)              # sector-ascending layout
    q, r, qi = ob.block_sparse_qr(a, q0, q1)
    assert rel(q @ r, a) < 1e-13 and rel(q.T @ q, np.identity(q.shape[1])) < 1e-13


def test_mps_ops_match_reference(golden_dir):
    """apply_mpo, mps_add, compress (svd / density) and from_vector of oracle/mps_ops.py against the fixture
    generated from the reference."""
    import copy
    import oracle.mps_ops as om
    z = np.load(os.path.join(golden_dir, "mps_ops.npz"))
    hw, hq, n = load_op(z)
    psi, chi = load_chain(z, "psi", n), load_chain(z, "chi", n)
    hp = om.apply_mpo(hw, hq, psi)
    assert all(np.array_equal(hp.qbonds[i], z[f"hpsi/qb{i}"]) for i in range(n + 1))
    assert all(rel(hp.a[i], z[f"hpsi/a{i}"]) < 1e-14 for i in range(n))
    sm = om.mps_add(psi, chi, complex(z["add/alpha"]))
    assert all(np.array_equal(sm.a[i], z[f"add/a{i}"]) for i in range(n))
    for tag, mode, direction in [("svd_l0", "svd", "left"), ("svd_r0", "svd", "right"), ("svd_l", "svd", "left"),
                                 ("svd_r", "svd", "right"), ("den", "density", "left"), ("den0", "density", "left")]:
        p = copy.deepcopy(hp)
        tol = float(z[f"cmp/{tag}/tol"])
        nrm, scale = om.compress_svd(p, tol, direction) if mode == "svd" else om.compress_density(p, tol)
        assert abs(nrm - float(z[f"cmp/{tag}/nrm"])) < 1e-12 * nrm and abs(scale - float(z[f"cmp/{tag}/scale"])) < 1e-12
        assert p.bond_dims == list(z[f"cmp/{tag}/bond_dims"])
        assert all(np.array_equal(p.qbonds[i], z[f"cmp/{tag}/qb{i}"]) for i in range(n + 1))
        assert rel(p.to_vector(), z[f"cmp/{tag}/vec"]) < 1e-11
    for tag in ("fv0", "fv"):
        m = om.from_vector(3, 5, z["fv/input"], tol=float(z[f"{tag}/tol"]))
        assert m.bond_dims == list(z[f"{tag}/bond_dims"]) and rel(m.to_vector(), z[f"{tag}/vec"]) < 1e-12


def test_round2_twosite_fixtures(golden_dir):
    """Round-2 fixtures (tests/golden/make_golden_r2.py): a truncating quantum-number two-site TDVP run and two-site
    sweeps with tol_split = 0 from a product state (rank-deficient splits) -- oracle == reference in bond dimensions,
    sector layouts and state."""
    z = np.load(os.path.join(golden_dir, "tdvp_fh_qnum_trunc_L8.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        osw.tdvp_twosite(w, wq, psi, complex(z["tdvp/dt"]), int(z["tdvp/nsteps"]), numiter_lanczos=int(z["tdvp/k"]),
                         tol_split=float(z["tdvp/tol"]))
    assert psi.bond_dims == list(z["tdvp/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(psi.qbonds[i], z[f"tdvp/qb{i}"])
    assert rel(psi.to_vector(), z["tdvp/vec"]) < 1e-10
    z = np.load(os.path.join(golden_dir, "twosite_rank_deficient_L6.npz"))
    w, wq, n = load_op(z)
    psi = load_chain(z, "psi0", n)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        osw.tdvp_twosite(w, wq, psi, complex(z["tdvp/dt"]), int(z["tdvp/nsteps"]), numiter_lanczos=int(z["tdvp/k"]),
                         tol_split=0)
        psi2 = load_chain(z, "psi0", n)
        en = osw.dmrg_twosite(w, wq, psi2, len(z["dmrg/en"]), numiter_lanczos=int(z["dmrg/k"]), tol_split=0)
    assert psi.bond_dims == list(z["tdvp/bond_dims"]) and psi2.bond_dims == list(z["dmrg/bond_dims"])
    assert rel(psi.to_vector(), z["tdvp/vec"]) < 1e-10
    assert np.max(np.abs(en - z["dmrg/en"])) < 1e-11
