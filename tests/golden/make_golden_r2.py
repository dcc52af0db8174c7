"""
Round-2 golden fixtures from the REAL reference (cmendl/pytenet v1.3.0, /root/reference); same conventions as
make_golden.py.  Run in the dev container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_r2.py

* tdvp_fh_qnum_trunc_L8.npz -- Fermi-Hubbard chain L = 8 with (N, Sz) quantum numbers, `MPS.construct_random` start,
  `tdvp_twosite` WITH truncation (tol_split = 1e-6) and `dmrg_twosite` (tol_split = 1e-8): bond dimensions, sector
  layouts (bit-exact gate) and state / energies.  Pins a truncating quantum-number run against the reference.
* twosite_rank_deficient_L6.npz -- two-site sweeps with tol_split = 0 from a PRODUCT state (every first split is rank
  deficient: LAPACK returns rounding-noise singular values and the reference keeps those bond indices):
  `tdvp_twosite` and `dmrg_twosite` on the XXZ chain with quantum numbers; bond dimensions / sector layouts / state.
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
import oracle.sweeps as osw    # noqa: E402
from make_golden import save_mps, save_mpo, to_chain, rel   # noqa: E402


def truncating_qnumber_case():
    L = 8
    h = ptn.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.5)
    sector = ptn.encode_quantum_number_pair(L, 0)
    rng = np.random.default_rng(2)
    psi = ptn.MPS.construct_random(L, h.qsite, sector, max_vdim=48, dtype="complex", rng=rng)
    psi.orthonormalize(mode="left")
    out = {}
    save_mpo(out, "h", h); save_mps(out, "psi0", psi)
    dt, nsteps, k, tol = 0.05j, 3, 10, 1e-6
    p = copy.deepcopy(psi); o = to_chain(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ptn.tdvp_twosite(h, p, dt, nsteps, numiter_lanczos=k, tol_split=tol)
        osw.tdvp_twosite(h.a, h.qbonds, o, dt, nsteps, numiter_lanczos=k, tol_split=tol)
    assert o.bond_dims == p.bond_dims and rel(o.to_vector(), p.to_vector()) < 1e-10
    out["tdvp/dt"] = np.array(dt); out["tdvp/nsteps"] = np.array(nsteps); out["tdvp/k"] = np.array(k)
    out["tdvp/tol"] = np.array(tol); out["tdvp/bond_dims"] = np.array(p.bond_dims); out["tdvp/vec"] = p.to_vector()
    for i, q in enumerate(p.qbonds):
        out[f"tdvp/qb{i}"] = np.asarray(q)
    p2 = copy.deepcopy(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        en = ptn.dmrg_twosite(h, p2, 3, numiter_lanczos=20, tol_split=1e-8)
    out["dmrg/en"] = en; out["dmrg/bond_dims"] = np.array(p2.bond_dims); out["dmrg/k"] = np.array(20)
    out["dmrg/tol"] = np.array(1e-8)
    for i, q in enumerate(p2.qbonds):
        out[f"dmrg/qb{i}"] = np.asarray(q)
    np.savez_compressed(os.path.join(HERE, "tdvp_fh_qnum_trunc_L8.npz"), **out)
    print("tdvp_fh_qnum_trunc_L8.npz: start bonds", psi.bond_dims, "-> tdvp", p.bond_dims, "dmrg", p2.bond_dims, en)


def rank_deficient_case():
    L = 6
    h = ptn.heisenberg_xxz_1d_mpo(L, 1.0, 0.7, 0.2)
    # Neel product state |up down up down ...> as an MPS with bond dimension 1 and the matching quantum numbers
    qsite = np.asarray(h.qsite)
    occ = [0, 1] * (L // 2)
    qbonds = [np.array([0])]
    for s in occ:
        qbonds.append(np.array([qbonds[-1][0] + qsite[s]]))
    psi = ptn.MPS(h.qsite, qbonds, fill=0.0)
    for i, s in enumerate(occ):
        psi.a[i] = np.zeros((1, 2, 1), dtype=complex)
        psi.a[i][0, s, 0] = 1.0
    out = {}
    save_mpo(out, "h", h); save_mps(out, "psi0", psi)
    dt, nsteps, k = 0.05 - 0.1j, 2, 8
    p = copy.deepcopy(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ptn.tdvp_twosite(h, p, dt, nsteps, numiter_lanczos=k, tol_split=0)
    out["tdvp/dt"] = np.array(dt); out["tdvp/nsteps"] = np.array(nsteps); out["tdvp/k"] = np.array(k)
    out["tdvp/bond_dims"] = np.array(p.bond_dims); out["tdvp/vec"] = p.to_vector()
    for i, q in enumerate(p.qbonds):
        out[f"tdvp/qb{i}"] = np.asarray(q)
    p2 = copy.deepcopy(psi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        en = ptn.dmrg_twosite(h, p2, 2, numiter_lanczos=10, tol_split=0)
    out["dmrg/en"] = en; out["dmrg/bond_dims"] = np.array(p2.bond_dims); out["dmrg/k"] = np.array(10)
    for i, q in enumerate(p2.qbonds):
        out[f"dmrg/qb{i}"] = np.asarray(q)
    np.savez_compressed(os.path.join(HERE, "twosite_rank_deficient_L6.npz"), **out)
    print("twosite_rank_deficient_L6.npz: tdvp bonds", p.bond_dims, "dmrg bonds", p2.bond_dims, "energies", en)


if __name__ == "__main__":
    truncating_qnumber_case()
    rank_deficient_case()
