"""Multi-GPU check + timing of the MPO-bond-sharded matvec (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_check.py [--D 1024]

1. small problem: sharded result == unsharded device matvec (apply_local_hamiltonian) to 1e-12;
2. molecular-like shape (chi_l=562, chi_r=501, d=2): time per matvec, split into compute and exchange.
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pytenet_b200 as ptb
from pytenet_b200.sharded import ShardedEffectiveHamiltonian, PrecontractedShardedHamiltonian

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=1024)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--exchange", default="auto")
ap.add_argument("--mode", default="precontract", choices=["precontract", "exchange"])
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
device = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=device)

# ---- 1. correctness on a small problem (same inputs on every rank) -------------------------
g = torch.Generator(device=device).manual_seed(5)
Dl, d, Dr, cl, cr = 96, 2, 80, 37, 29
a = torch.randn(Dl, d, Dr, dtype=torch.complex128, device=device, generator=g)
l = torch.randn(Dl, cl, Dl, dtype=torch.complex128, device=device, generator=g)
r = torch.randn(Dr, cr, Dr, dtype=torch.complex128, device=device, generator=g)
w = torch.randn(cl, d, d, cr, dtype=torch.float64, device=device, generator=g)
w = w * (torch.rand(cl, d, d, cr, device=device, generator=g) < 0.2)
if args.mode == "precontract":
    make = lambda w_, l_, r_: PrecontractedShardedHamiltonian.from_full(w_, l_, r_)
else:
    make = lambda w_, l_, r_: ShardedEffectiveHamiltonian.from_full(w_, l_, r_, exchange=args.exchange)
heff = make(w, l, r)
out = heff.matvec(a)
ref = ptb.apply_local_hamiltonian(a, w, l, r)
err = (torch.linalg.norm(out - ref) / torch.linalg.norm(ref)).item()
assert err < 1e-12, err
# Lanczos on the sharded operator gives the same Ritz value on every rank
lh = l + l.conj().permute(2, 1, 0); rh = r + r.conj().permute(2, 1, 0); wh = w + w.permute(0, 2, 1, 3)
hs = make(wh, lh, rh)
ev, _ = ptb.eigh_krylov(lambda x: hs.matvec(x.reshape(Dl, d, Dr)).reshape(-1), a.reshape(-1), 12, 1)
ev_ref, _ = ptb.eigh_krylov(lambda x: ptb.apply_local_hamiltonian(x.reshape(Dl, d, Dr), wh, lh, rh).reshape(-1),
                            a.reshape(-1), 12, 1)
assert abs(ev[0] - ev_ref[0]) < 1e-9 * abs(ev_ref[0]), (ev, ev_ref)
if rank == 0:
    print(f"[{heff.exchange}] sharded matvec == unsharded device matvec: rel err {err:.2e}; Ritz value {ev[0]:.12f} vs {ev_ref[0]:.12f}",
          flush=True)
exch_small = heff.exchange
del heff, hs, l, r, w, lh, rh, wh

# ---- 2. timing at the molecular-like shape --------------------------------------------------
D = args.D
cl, cr, d = 562, 501, 2
setup_ms = 0.0
if args.mode == "precontract":
    heff, setup_ms = PrecontractedShardedHamiltonian.synthetic(D, d, D, cl, cr, density=0.168, seed=1, device=device)
else:
    heff = ShardedEffectiveHamiltonian.synthetic(D, d, D, cl, cr, density=0.168, seed=1, device=device,
                                                 exchange=args.exchange)
a = torch.randn(D, d, D, dtype=torch.complex128, device=device) / np.sqrt(D * d * D)
for _ in range(2):
    heff.matvec(a)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    heff.matvec(a)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
t = torch.tensor([ms], dtype=torch.float64, device=device)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
f_alg = 8.0 * (D * d * D * cr * D + cl * d * d * cr * D * D + D * D * cl * d * D)
gather, red = heff.exchange_bytes_per_rank()
if rank == 0:
    print(json.dumps({"sharded_matvec": {"n_gpus": world, "exchange": heff.exchange,
                                         "note": getattr(heff, "_exchange_note", ""), "mode": args.mode,
                                         "precontract_ms_per_site": setup_ms, "D": D, "chi_l": cl, "chi_r": cr, "d": d,
                                         "ms_per_matvec": t.item(), "gflops_alg": f_alg / t.item() / 1e6,
                                         "allgather_bytes_per_rank": gather, "allreduce_bytes_per_rank": red,
                                         "flops_exec_per_rank": heff.flops_per_rank()}}), flush=True)
if world > 1:
    dist.destroy_process_group()
