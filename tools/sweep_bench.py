"""Sweep seconds on the GPU (BASELINE metric part 2: "TDVP sweep s at D=2048"; config 2: dmrg_twosite D=1024).

The reference API has no maximum-bond argument, and a full L=100 sweep is hundreds of seconds, so this
tool times complete sweeps of shorter chains whose bulk bonds reach the target D (bond profile
min(2^i, 2^(L-i), D)) and reports the per-site cost of the bulk sites with a breakdown by phase
(matvec / Lanczos vector ops / QR-SVD (cuSOLVER) / environment updates), from which the L=100 sweep
time is extrapolated (labelled as such).

    python tools/sweep_bench.py --algo tdvp1 --D 2048 --L 24 --k 25
    python tools/sweep_bench.py --algo dmrg2 --D 1024 --L 22 --k 25 --tol 1e-8
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import chain_ops, block_sparse_util, krylov, _sweep, mps as pmps, tdvp as ptdvp, dmrg as pdmrg

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="tdvp1", choices=["tdvp1", "dmrg2", "tdvp2", "dmrg1"])
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--L", type=int, default=24)
ap.add_argument("--k", type=int, default=25)
ap.add_argument("--tol", type=float, default=0.0)
args = ap.parse_args()
L, D = args.L, args.D
h = ptb.heisenberg_xxz_1d_mpo(L, 1.0, 0.8, -0.1).zero_qnumbers()
bonds = [min(2 ** i, 2 ** (L - i), D) for i in range(L + 1)]
rng = np.random.default_rng(42)
t0 = time.time()
psi = ptb.MPS(h.qsite, [np.zeros(b, dtype=int) for b in bonds], fill="random", rng=rng)
gen_s = time.time() - t0

# ---- phase timers (synchronising; only used by this tool) ----
acc = {}
def timed(name, fn):
    def wrap(*a, **kw):
        torch.cuda.synchronize(); t = time.perf_counter()
        out = fn(*a, **kw)
        torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        acc[name + "_calls"] = acc.get(name + "_calls", 0) + 1
        return out
    return wrap
_sweep.apply_local_hamiltonian = timed("matvec", chain_ops.apply_local_hamiltonian)
_sweep.apply_local_bond_contraction = timed("bond_matvec", chain_ops.apply_local_bond_contraction)
for mod in (ptdvp, pdmrg):
    mod.env_step_left = timed("env_update", _sweep.env_step_left)
    mod.env_step_right = timed("env_update", _sweep.env_step_right)
qr_t = timed("qr_cusolver", block_sparse_util.block_sparse_qr)
svd_t = timed("svd_cusolver", block_sparse_util.block_sparse_svd)
pmps.block_sparse_qr = qr_t; ptdvp.block_sparse_qr = qr_t
import pytenet_b200.bond_ops as pbo
pbo.block_sparse_svd = svd_t

torch.cuda.synchronize(); t0 = time.perf_counter()
nrm, lb, rb = _sweep.prepare_environments(h, psi)
torch.cuda.synchronize(); prologue_s = time.perf_counter() - t0
acc.clear()
# run one full step / sweep through the public driver (it repeats the prologue on the already
# orthonormal state, which is timed separately above)
torch.cuda.synchronize(); t0 = time.perf_counter()
if args.algo == "tdvp1":
    ptb.tdvp_singlesite(h, psi, 0.01 - 0.05j, 1, numiter_lanczos=args.k)
elif args.algo == "tdvp2":
    ptb.tdvp_twosite(h, psi, 0.01 - 0.05j, 1, numiter_lanczos=args.k, tol_split=args.tol)
elif args.algo == "dmrg1":
    ptb.dmrg_singlesite(h, psi, 1, numiter_lanczos=args.k)
else:
    ptb.dmrg_twosite(h, psi, 1, numiter_lanczos=args.k, tol_split=args.tol)
torch.cuda.synchronize(); total_s = time.perf_counter() - t0
phases = {k: v for k, v in acc.items()}
other = total_s - sum(v for k, v in acc.items() if not k.endswith("_calls"))
bulk = sum(1 for i in range(L) if bonds[i] == D and bonds[i + 1] == D)
print(json.dumps({"sweep_bench": {"algo": args.algo, "L": L, "D": D, "k": args.k, "tol_split": args.tol,
      "bond_dims_after": psi.bond_dims, "sites_with_full_D": bulk, "state_generation_s": gen_s,
      "prologue_s": prologue_s, "sweep_total_s": total_s, "phases_s": phases,
      "lanczos_vector_ops_and_host_s": other}}))
