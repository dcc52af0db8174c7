"""Effective-H matvec over a range of bond dimensions (two-site XXZ, d=4, chi=5, complex128): device path vs
the CPU oracle on the host cores, to show where the GPU path stops being tensor-pipe bound."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import oracle
import pytenet_b200 as ptb
from bench import host_inputs, f_alg

rows = []
for D in [16, 32, 64, 128, 256, 512, 1024, 2048]:
    a, w, l, r = host_inputs(D, 4, 5, seed=3)
    ad, wd, ld, rd = (torch.from_numpy(x).cuda() for x in (a, w, l, r))
    for _ in range(3):
        ptb.apply_local_hamiltonian(ad, wd, ld, rd)
    torch.cuda.synchronize()
    reps = 200 if D <= 256 else (20 if D <= 1024 else 5)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ptb.apply_local_hamiltonian(ad, wd, ld, rd)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # wall clock including Python dispatch (what a Lanczos loop sees)
    t0 = time.perf_counter()
    for _ in range(reps):
        ptb.apply_local_hamiltonian(ad, wd, ld, rd)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / reps * 1e3
    if D <= 1024:
        oracle.apply_local_hamiltonian(a, w, l, r)
        creps = 20 if D <= 128 else 3
        t0 = time.perf_counter()
        for _ in range(creps):
            oracle.apply_local_hamiltonian(a, w, l, r)
        cpu = (time.perf_counter() - t0) / creps * 1e3
    else:
        cpu = None
    F = f_alg(D, 4, 5)
    rows.append({"D": D, "gpu_ms": ms, "gpu_wall_ms": wall, "gpu_gflops": F / ms / 1e6, "cpu_ms": cpu,
                 "cpu_gflops": (F / cpu / 1e6) if cpu else None})
    print(json.dumps(rows[-1]), flush=True)
