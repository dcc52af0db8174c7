"""Multi-GPU check of the MPO-bond-sharded single-site DMRG (run under torchrun, one rank per GPU):
molecular Hamiltonian fixture (reference-built MPO) -> energies must equal the reference's on every rank,
and psi must be identical across ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
        tools/sharded_dmrg_check.py
"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pytenet_b200 as ptb
from pytenet_b200.sharded_dmrg import dmrg_singlesite_sharded

warnings.simplefilter("ignore")
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
z = np.load(os.path.join(ROOT, "tests", "golden", "dmrg_molecular_N8.npz"))
n = int(z["h/nsites"])
h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
psi = ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)], [z[f"psi0/a{i}"] for i in range(n)])
en = dmrg_singlesite_sharded(h, psi, 3, numiter_lanczos=int(z["k"]))
err = float(np.max(np.abs(en - z["single/en"])))
vec = torch.from_numpy(psi.to_vector()).to(device)
if world > 1:
    ref = vec.clone()
    dist.broadcast(torch.view_as_real(ref), src=0)
    drift = (torch.linalg.norm(vec - ref)).item()
    t = torch.tensor([err, drift], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err, drift = t.tolist()
else:
    drift = 0.0
if rank == 0:
    print(f"sharded dmrg_singlesite on {world} rank(s): energies {en}, max |E - E_reference| = {err:.2e}, "
          f"max psi drift across ranks = {drift:.2e}, MPO bonds {h.bond_dims}", flush=True)
assert err < 1e-10 and drift < 1e-12
# ---- sharded single-site TDVP on the README fixture (XXZ, chi = 5 split over the ranks) ----
from pytenet_b200.sharded_dmrg import tdvp_singlesite_sharded
z = np.load(os.path.join(ROOT, "tests", "golden", "tdvp_xxz_L10.npz"))
n = int(z["h/nsites"])
h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
psi = ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)], [z[f"psi0/a{i}"] for i in range(n)])
tdvp_singlesite_sharded(h, psi, complex(z["dt"]), int(z["nsteps"]), numiter_lanczos=int(z["k"]))
v = psi.to_vector()
terr = float(np.linalg.norm(v - z["single/vec"]) / np.linalg.norm(z["single/vec"]))
if world > 1:
    t = torch.tensor([terr], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    terr = t.item()
if rank == 0:
    print(f"sharded tdvp_singlesite on {world} rank(s): rel. state error vs reference = {terr:.2e}", flush=True)
assert terr < 1e-9
if world > 1:
    dist.destroy_process_group()
