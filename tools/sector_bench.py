"""Block-sparse (quantum-number) effective-H matvec: dense device path vs the sector-banded path at a
config-3-like shape (two-site Fermi-Hubbard: d = 16, chi = 6, (N, Sz) sectors).

    python tools/sector_bench.py [--D 2048] [--profile physical|fragmented]

`physical`: ~40 (N, Sz) sectors with a Gaussian size profile (a few sectors of O(100), what converged
states look like); `fragmented`: the reference's random generator statistics (many sectors of ~10)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import hamiltonian as ham
from pytenet_b200.sectors import HeffSectorPlan

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--profile", default="physical")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
D = args.D
rng = np.random.default_rng(3)
qsite, qb, wbulk, _, _ = ham._fermi_hubbard_bulk(1.0, 4.0, 0.0)
qsite = np.array(qsite); qb = np.array(qb)
d1 = len(qsite)
qs2 = np.add.outer(qsite, qsite).reshape(-1)
w2 = np.einsum("kpqm,mrsn->kprqsn", wbulk, wbulk).reshape(6, d1 * d1, d1 * d1, 6)


def bond_qnumbers(n, nsec_n, nsec_s, width):
    """n indices distributed over (N, Sz) sectors with a Gaussian profile, sorted by encoded value."""
    cand = [(dn, ds) for dn in range(-nsec_n, nsec_n + 1) for ds in range(-nsec_s, nsec_s + 1) if (dn + ds) % 2 == 0]
    wts = np.array([np.exp(-(dn ** 2 + ds ** 2) / (2 * width ** 2)) for dn, ds in cand])
    sizes = np.floor(wts / wts.sum() * n).astype(int)
    sizes[np.argmax(sizes)] += n - sizes.sum()
    q = np.concatenate([np.full(sz, ptb.encode_quantum_number_pair(32 + dn, ds)) for (dn, ds), sz in zip(cand, sizes)])
    return np.sort(q), sizes[sizes > 0]


if args.profile == "physical":
    ql, sizes = bond_qnumbers(D, 4, 4, 1.6)
    qr, _ = bond_qnumbers(D, 4, 4, 1.6)
else:
    ql, sizes = bond_qnumbers(D, 8, 8, 6.0)
    qr, _ = bond_qnumbers(D, 8, 8, 6.0)

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
def crand(*s):
    return torch.randn(*s, dtype=torch.complex128, device=dev, generator=g)
a = crand(D, d1 * d1, D); ptb.enforce_qsparsity(a, [ql, qs2, -qr])
l = crand(D, 6, D); ptb.enforce_qsparsity(l, [ql, qb, -ql])
r = crand(D, 6, D); ptb.enforce_qsparsity(r, [qr, qb, -qr])
w = torch.from_numpy(w2).to(dev)
fill = (a != 0).double().mean().item()
plan = HeffSectorPlan(ql, qs2, qr, qb, qb, cplx=True)

def timeit(fn):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.reps

out_d = ptb.apply_local_hamiltonian(a, w, l, r)
out_s = plan.apply(a, w, l, r)
err = (torch.linalg.norm(out_s - out_d) / torch.linalg.norm(out_d)).item()
ms_dense = timeit(lambda: ptb.apply_local_hamiltonian(a, w, l, r))
ms_band = timeit(lambda: plan.apply(a, w, l, r))
chi, d = 6, d1 * d1
f_alg = 8.0 * (D * d * D * chi * D + chi * d * d * chi * D * D + D * D * chi * d * D)
print(json.dumps({"sector_matvec": {
    "profile": args.profile, "D": D, "d": d, "chi": chi, "sectors": int(len(sizes)), "max_sector": int(sizes.max()),
    "median_sector": float(np.median(sizes)), "tensor_fill": fill, "rel_err_vs_dense": err,
    "ms_dense": ms_dense, "ms_banded": ms_band, "speedup": ms_dense / ms_band,
    "gflops_alg_dense": f_alg / ms_dense / 1e6, "gflops_alg_banded": f_alg / ms_band / 1e6,
    "visit_fraction_step1": plan.visit_fraction[0], "visit_fraction_step3": plan.visit_fraction[1]}}))
