"""BASELINE config 1 (README.rst:18-48): XXZ L=10, D<=28 (clamped 8), tdvp_singlesite, k=5 -- the
launch-latency-bound end of the path.  Times the device driver and the CPU oracle on the same
fixture state and reports seconds per TDVP step and microseconds per local problem.

    python tools/readme_bench.py [--steps 100]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--cpu", type=int, default=1)
args = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "tdvp_xxz_L10.npz"))
n = int(z["h/nsites"])
qsite = z["h/qsite"]
hq = [z[f"h/qb{i}"] for i in range(n + 1)]
hw = [z[f"h/w{i}"] for i in range(n)]
pq = [z[f"psi0/qb{i}"] for i in range(n + 1)]
pa = [z[f"psi0/a{i}"] for i in range(n)]
dt = complex(z["dt"]); k = int(z["k"])

h = ptb.MPO.from_tensors(qsite, hq, hw)
psi = ptb.MPS.from_tensors(z["psi0/qsite"], pq, pa)
ptb.tdvp_singlesite(h, psi, dt, 2, numiter_lanczos=k)      # warm-up
psi = ptb.MPS.from_tensors(z["psi0/qsite"], pq, pa)
torch.cuda.synchronize(); t0 = time.perf_counter()
ptb.tdvp_singlesite(h, psi, dt, args.steps, numiter_lanczos=k)
torch.cuda.synchronize(); gpu_s = time.perf_counter() - t0
vec_gpu = psi.to_vector()
# local problems per step: 2 sweeps x (L site steps + (L-1) bond steps), minus the shared turning points
local = args.steps * (2 * n + 2 * (n - 1))
res = {"config": "README XXZ L=10 tdvp_singlesite", "steps": args.steps, "k": k, "bond_dims": psi.bond_dims,
       "gpu_s": gpu_s, "gpu_s_per_step": gpu_s / args.steps, "gpu_us_per_local_problem": 1e6 * gpu_s / local,
       "graphs": os.environ.get("PYTENET_B200_GRAPHS", "auto"), "graph_capture_error": ptb.tdvp._StepGraph.last_error}
if args.cpu:
    from oracle import sweeps as osw
    ch = osw.Chain([a.copy() for a in pa], z["psi0/qsite"], [q.copy() for q in pq])
    t0 = time.perf_counter()
    osw.tdvp_singlesite(hw, hq, ch, dt, args.steps, numiter_lanczos=k)
    cpu_s = time.perf_counter() - t0
    vec_cpu = ch.to_vector()
    res.update({"cpu_oracle_s": cpu_s, "cpu_s_per_step": cpu_s / args.steps,
                "rel_diff_state": float(np.linalg.norm(vec_gpu - vec_cpu) / np.linalg.norm(vec_cpu))})
print(json.dumps({"readme_bench": res}))
