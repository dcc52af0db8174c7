"""Single-site TDVP step with quantum numbers at D = 2048 (BASELINE config-3 physics: Fermi-Hubbard, (N, Sz)
sectors): the sector path (work lists for matvec, bond contraction and environment updates) against the dense
device path on the same state.

    python tools/qn_sweep_bench.py [--L 12] [--D 2048] [--k 10]

The state comes from the reference's own generator (`MPS.construct_random`, sector N = L, Sz = 0), is brought to
canonical form once (so every bond is grouped by sector) and then evolved by one symmetric TDVP step through the
public driver, once per path.  Reported: seconds per step, agreement of the two final states (overlap), sector
statistics of the largest bond."""
import argparse, json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import _sweep, _prof

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=12)
ap.add_argument("--D", type=int, default=2048)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--dense", type=int, default=1)
ap.add_argument("--algo", default="one", choices=["one", "two"])
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--prof", type=int, default=0, help="device-time split by phase (pytenet_b200/_prof.py)")
args = ap.parse_args()
L, D = args.L, args.D
h = ptb.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.0)
rng = np.random.default_rng(11)
psi0 = ptb.MPS.construct_random(L, h.qsite, ptb.encode_quantum_number_pair(L, 0), max_vdim=D, rng=rng)
psi0.orthonormalize(mode="left"); psi0.orthonormalize(mode="right")
big = int(np.argmax(psi0.bond_dims))
vals, cnt = np.unique(psi0.qbonds[big], return_counts=True)
dt = 0.02j
res = {"model": f"Fermi-Hubbard L={L}, sector (N={L}, Sz=0)", "algo": "tdvp_" + args.algo + "site", "tol_split": args.tol,
       "bond_dims": psi0.bond_dims, "k": args.k,
       "largest_bond_sectors": int(len(vals)), "largest_bond_sector_sizes_median_max": [float(np.median(cnt)), int(cnt.max())]}
finals = {}
for mode in (["auto", "0"] if args.dense else ["auto"]):
    _sweep._SECTOR_MODE = mode
    psi = psi0.copy()
    tag = "sector_path" if mode == "auto" else "dense_path"
    for rep in ("first_step", "later_step"):        # later steps reuse the cached sector plans
        if args.prof:
            _prof.enable(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if args.algo == "one":
            ptb.tdvp_singlesite(h, psi, dt, 1, numiter_lanczos=args.k)
        else:
            ptb.tdvp_twosite(h, psi, dt, 1, numiter_lanczos=args.k, tol_split=args.tol)
        torch.cuda.synchronize()
        res[f"bond_dims_{tag}_{rep}"] = psi.bond_dims
        res[f"seconds_{tag}_{rep}"] = time.perf_counter() - t0
        if args.prof:
            res[f"device_seconds_by_phase_{tag}_{rep}"] = {k: v / 1e3 for k, v in _prof.report().items()}
            _prof.enable(False)
    finals[mode] = psi
    if mode == "auto":
        # the driver's prologue (right-orthonormalisation + all right environment blocks, tdvp.py:44-63) is part of
        # every call: three more steps in ONE call give the cost of a time step without it
        _prof.enable(False) if args.prof else None
        psi3 = psi.copy()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if args.algo == "one":
            ptb.tdvp_singlesite(h, psi3, dt, 3, numiter_lanczos=args.k)
        else:
            ptb.tdvp_twosite(h, psi3, dt, 3, numiter_lanczos=args.k, tol_split=args.tol)
        torch.cuda.synchronize()
        t3 = time.perf_counter() - t0
        res["seconds_three_steps_one_call"] = t3
        res["seconds_per_step_without_prologue"] = (t3 - res[f"seconds_{tag}_later_step"]) / 2
        del psi3
if args.dense:
    ov = ptb.mps_vdot(finals["auto"], finals["0"])
    res["overlap_sector_vs_dense"] = [float(np.real(ov)), float(np.imag(ov))]
    res["one_minus_abs_overlap"] = float(abs(1 - abs(ov)))
    res["speedup_later_step"] = res["seconds_dense_path_later_step"] / res["seconds_sector_path_later_step"]
print(json.dumps({"qn_sweep_bench": res}))
