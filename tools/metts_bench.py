"""METTS sample streams, one per GPU (BASELINE config 5: Ising 1D, independent samples, no communication).

    python tools/metts_bench.py [--L 32] [--samples 4] [--beta 1.0]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
        tools/metts_bench.py --L 64 --samples 8

    # several independent sample streams PER GPU (one process each; the kernels of this regime occupy 1-8 of the 148
    # SMs and a stream is host-bound, so streams interleave on the device):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
        tools/metts_bench.py --L 64 --samples 8 --streams 4

Each rank runs its own chain of samples with its own seed; rank 0 gathers the per-sample energies at the end.
Reported: samples/s per GPU and aggregate, thermal-energy estimate with its standard error, realised bond dims."""
import argparse, json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import pytenet_b200 as ptb

warnings.simplefilter("ignore")
ap = argparse.ArgumentParser()
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--samples", type=int, default=4)
ap.add_argument("--beta", type=float, default=1.0)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--tol", type=float, default=1e-10)
ap.add_argument("--streams", type=int, default=1, help="sample streams (processes) per GPU; WORLD_SIZE = GPUs x streams")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
assert world % args.streams == 0
ngpus = world // args.streams
local = local % ngpus                           # streams of one GPU: ranks g, g + ngpus, g + 2 ngpus, ...
torch.cuda.set_device(local)
if world > 1:
    # scalars only (barrier + final gather): gloo when several ranks share a GPU (NCCL refuses duplicate devices)
    if args.streams > 1:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = ptb.ising_1d_mpo(args.L, 1.0, 0.8, -0.375)
rng = np.random.default_rng(1000 + rank)
ptb.metts_energy_samples(h, args.beta, 1, rng, numsteps=args.steps, numiter_lanczos=args.k, tol_split=args.tol)  # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
stats = {}
vals = ptb.metts_energy_samples(h, args.beta, args.samples, rng, numsteps=args.steps, numiter_lanczos=args.k,
                                tol_split=args.tol, stats=stats)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
allv = [None] * world
if world > 1:
    dist.all_gather_object(allv, (vals.real.tolist(), dt, stats))
else:
    allv = [(vals.real.tolist(), dt, stats)]
if rank == 0:
    e = np.concatenate([np.array(v) for v, _, _ in allv])
    tmax = max(t for _, t, _ in allv)
    mb = np.concatenate([np.array(st["max_bond"]) for _, _, st in allv])
    print(json.dumps({"metts": {"L": args.L, "beta": args.beta, "n_gpus": ngpus, "streams_per_gpu": args.streams,
                                "samples_per_gpu": args.samples * args.streams,
                                "seconds": tmax, "samples_per_s_total": len(e) / tmax,
                                "energy_per_site_mean": float(e.mean() / args.L),
                                "energy_per_site_stderr": float(e.std() / np.sqrt(len(e)) / args.L),
                                "samples_per_s_per_gpu": len(e) / tmax / ngpus,
                                "realised_max_bond_dim": {"max": int(mb.max()), "median": float(np.median(mb)),
                                                          "min": int(mb.min())},
                                "tdvp_steps": args.steps, "k": args.k, "tol_split": args.tol}}))
if world > 1:
    dist.destroy_process_group()
