"""BASELINE config 4 for real: `dmrg_singlesite` on the 32-orbital molecular Hamiltonian (the reference's own
`molecular_hamiltonian_mpo(tkin, vint, optimize=False)` tensors, cached sparse in tests/golden/molecular_mpo_N32.npz)
with the MPO virtual bond sharded over the GPUs of one box.

    python tools/config4_sweep.py --norb 32 --D 1024 --sweeps 1 --k 25            # one GPU (memory permitting)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
        tools/config4_sweep.py --norb 32 --D 1024 --sweeps 1 --k 25

Prints one JSON line: seconds per sweep (device-synchronised wall time, max over ranks), peak GB per GPU, the energy
after every sweep, the drift of psi across ranks.  `--norb 10` runs the small fixture and checks the energies
against the reference's `dmrg_singlesite` (stored next to the MPO by tests/golden/make_molecular_mpo.py)."""
import argparse, json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pytenet_b200 as ptb
from pytenet_b200.hamiltonian import load_cached_mpo
from pytenet_b200.sharded_dmrg import dmrg_singlesite_sharded

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--norb", type=int, default=32)
ap.add_argument("--D", type=int, default=1024)
ap.add_argument("--sweeps", type=int, default=1)
ap.add_argument("--k", type=int, default=25)
ap.add_argument("--seed", type=int, default=11)
ap.add_argument("--out", default="")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)

path = os.path.join(ROOT, "tests", "golden", f"molecular_mpo_N{args.norb}.npz")
h = load_cached_mpo(path)
z = np.load(path)
n = h.nsites
check = None
if "dmrg_single_en" in z.files and args.norb <= 12:
    psi = ptb.MPS.from_tensors(h.qsite, [z[f"psi0_qb{i}"] for i in range(n + 1)], [z[f"psi0_a{i}"] for i in range(n)])
    check = z["dmrg_single_en"]
    args.k = int(z["dmrg_k"]); args.sweeps = len(check)
else:
    rng = np.random.default_rng(args.seed)                 # same state on every rank
    psi = ptb.MPS.construct_random(n, h.qsite, n // 2, max_vdim=args.D, dtype="complex", rng=rng)
bonds0 = list(psi.bond_dims)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.reset_peak_memory_stats()
t0 = time.perf_counter()
en = dmrg_singlesite_sharded(h, psi, args.sweeps, numiter_lanczos=args.k)
torch.cuda.synchronize()
secs = time.perf_counter() - t0
# psi must be identical on all ranks: compare a cheap fingerprint (all tensors' sums of squares and first entries)
fp = torch.stack([torch.stack([torch.linalg.norm(t).real.double(), t.reshape(-1)[0].real.double(),
                               t.reshape(-1)[-1].real.double()]) for t in psi.a]).reshape(-1)
stat = torch.tensor([secs, torch.cuda.max_memory_allocated() / 1e9], dtype=torch.float64, device=device)
drift = 0.0
if world > 1:
    ref = fp.clone()
    dist.broadcast(ref, src=0)
    d = (fp - ref).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX)
    drift = d.item()
    dist.all_reduce(stat, op=dist.ReduceOp.MAX)
if rank == 0:
    rec = {"config4_dmrg_singlesite_sharded": {
        "n_gpus": world, "norb": args.norb, "D": args.D, "k": args.k, "sweeps": args.sweeps,
        "mpo_bond_max": int(max(h.bond_dims)), "mps_bonds": bonds0, "mps_bond_max": int(max(bonds0)),
        "seconds_total": stat[0].item(), "seconds_per_sweep": stat[0].item() / args.sweeps,
        "peak_mem_gb_per_gpu": stat[1].item(), "energies": [float(e) for e in en],
        "psi_drift_across_ranks": drift}}
    if check is not None:
        rec["config4_dmrg_singlesite_sharded"]["max_abs_err_vs_reference"] = float(np.max(np.abs(en - check)))
    line = json.dumps(rec)
    print(line, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "a") as f:
            f.write(line + "\n")
if check is not None:
    assert np.max(np.abs(en - check)) < 1e-10, (en, check)
assert drift == 0.0
if world > 1:
    dist.destroy_process_group()
