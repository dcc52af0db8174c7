"""TDVP single-site step with MPO-bond-sharded environments at a molecular-like MPO bond dimension
(synthetic sparse real MPO tensors, chi up to 562, 16.8 % dense), N ranks = N GPUs.

    python tools/sharded_sweep_bench.py --D 512 --L 20 --k 10                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 \
        tools/sharded_sweep_bench.py --D 512 --L 20 --k 10

Reports the wall time of one full symmetric TDVP step (left + right sweep) through
`tdvp_singlesite_sharded`, max over ranks.  The MPO is random (not a physical Hamiltonian): the run times
the sweep machinery (gathers, precontractions, matvecs, all-reduces, QR), not physics."""
import argparse, json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pytenet_b200 as ptb
from pytenet_b200.sharded_dmrg import tdvp_singlesite_sharded

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=512)
ap.add_argument("--L", type=int, default=20)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--chi", type=int, default=562)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)
L, D = args.L, args.D
chi = [min(args.chi, 5 * 14 ** min(i - 1, L - 1 - i)) if 0 < i < L else 1 for i in range(L + 1)]
rng = np.random.default_rng(7)                      # same MPO and state on every rank
ws = []
for i in range(L):
    w = rng.normal(size=(chi[i], 2, 2, chi[i + 1])) * (rng.random((chi[i], 2, 2, chi[i + 1])) < 0.168)
    w = w + w.transpose(0, 2, 1, 3)                 # symmetric in the physical indices
    ws.append(w / max(1.0, np.sqrt(chi[i])))
h = ptb.MPO.from_tensors(np.zeros(2, int), [np.zeros(c, int) for c in chi], ws)
bonds = [min(2 ** i, 2 ** (L - i), D) for i in range(L + 1)]
psi = ptb.MPS(np.zeros(2, int), [np.zeros(b, int) for b in bonds], fill="random", rng=rng)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
tdvp_singlesite_sharded(h, psi, 0.01 - 0.02j, 1, numiter_lanczos=args.k)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device=device)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"sharded_tdvp_step": {"n_gpus": world, "L": L, "D": D, "k": args.k, "mpo_bond_max": max(chi),
                                            "seconds": t.item(),
                                            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}}), flush=True)
if world > 1:
    dist.destroy_process_group()
