"""Generator of the cached molecular MPO of BASELINE config 4 (runs in the dev container only).

Imports the REFERENCE (`/root/reference`, `molecular_hamiltonian_mpo(tkin, vint, optimize=False)`,
`pytenet/hamiltonian/molecular.py:612`) on seeded random symmetric integrals and stores the MPO tensors in
sparse coordinate form (`flat index`, `value` per site; the centre tensor `(562,2,2,501)` is 16.8 % dense), plus
`qsite`, `qbonds`, the seed and the integrals, as one compressed `.npz`.  The GPU box has no reference checkout:
`pytenet_b200.hamiltonian.load_cached_mpo` rebuilds the MPO from this file.

    python tests/golden/make_molecular_mpo.py --norb 32      # ~11 min of CPU, writes molecular_mpo_N32.npz
    python tests/golden/make_molecular_mpo.py --norb 10      # seconds; also stores the reference's DMRG energies
"""
import argparse
import copy
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import pytenet as ptn  # noqa: E402  (the reference)

SEED = 2032


def integrals(n, seed=SEED):
    """Random symmetric integrals: t = tᵀ, v[i,j,k,l] = v[j,i,l,k] = v[k,l,i,j] (as tests/golden/make_golden.py)."""
    rng = np.random.default_rng(seed)
    tkin = rng.normal(size=(n, n))
    tkin = 0.5 * (tkin + tkin.T)
    vint = rng.normal(size=(n, n, n, n))
    vint = 0.5 * (vint + vint.transpose(1, 0, 3, 2))
    vint = 0.5 * (vint + vint.transpose(2, 3, 0, 1))
    return tkin, vint, rng


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--norb", type=int, default=32)
    ap.add_argument("--dmrg", action="store_true", help="also store reference dmrg_singlesite energies (small norb)")
    ap.add_argument("--maxd", type=int, default=24)
    args = ap.parse_args()
    n = args.norb
    tkin, vint, rng = integrals(n)
    t0 = time.time()
    h = ptn.molecular_hamiltonian_mpo(tkin, vint, optimize=False)
    build_s = time.time() - t0
    out = {"seed": np.array(SEED), "norb": np.array(n), "tkin": tkin,
           "qsite": np.asarray(h.qsite), "bond_dims": np.array(h.bond_dims), "build_seconds": np.array(build_s)}
    nnz = 0
    for i, w in enumerate(h.a):
        assert w.dtype == np.float64
        flat = w.reshape(-1)
        idx = np.flatnonzero(flat)
        out[f"w{i}_shape"] = np.array(w.shape)
        out[f"w{i}_idx"] = idx.astype(np.int32)
        out[f"w{i}_val"] = flat[idx]
        nnz += idx.size
    for i, q in enumerate(h.qbonds):
        out[f"qb{i}"] = np.asarray(q)
    if args.dmrg:
        psi = ptn.MPS.construct_random(n, h.qsite, n // 2, max_vdim=args.maxd, dtype="complex", rng=rng)
        for i, a in enumerate(psi.a):
            out[f"psi0_a{i}"] = a
        for i, q in enumerate(psi.qbonds):
            out[f"psi0_qb{i}"] = np.asarray(q)
        p = copy.deepcopy(psi)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            en = ptn.dmrg_singlesite(h, p, 2, numiter_lanczos=10)
        out["dmrg_single_en"] = en
        out["dmrg_k"] = np.array(10)
        for i, a in enumerate(p.a):
            out[f"psi1_a{i}"] = a
        print("reference dmrg_singlesite energies", en)
    path = os.path.join(HERE, f"molecular_mpo_N{n}.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: bonds {h.bond_dims} nnz {nnz} build {build_s:.1f}s size {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
