"""Per-kernel device time (torch.profiler / CUPTI) of one 25-iteration Lanczos run in the packed space at the
BASELINE config-3 bond dimension (D = 2048, reference generator's fragmented (N, Sz) sectors): the single-site
problem (d = 4), the zero-site (bond) problem (d = 1, identity MPO tensor) and one environment update.

    python tools/packed_lanczos_probe.py [D] [k]
"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import pytenet_b200 as ptb
from pytenet_b200 import _sweep
warnings.simplefilter("ignore")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
k = int(sys.argv[2]) if len(sys.argv) > 2 else 25
L = 16
h = ptb.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.0)
psi = ptb.MPS.construct_random(L, h.qsite, ptb.encode_quantum_number_pair(L, 0), max_vdim=D, rng=np.random.default_rng(11))
psi.orthonormalize(mode="left")
_, lblocks, rblocks = _sweep.prepare_environments(h, psi)
i = L // 2
for j in range(i):
    lblocks[j + 1] = _sweep.env_step_left(psi, h, j, lblocks[j])
qh = h.qbonds
site_plan = _sweep.sector_plan(psi.qbonds[i], psi.qsite, psi.qbonds[i + 1], qh[i], qh[i + 1], psi.a[i], lblocks[i], rblocks[i], h.a[i])
c = torch.randn(psi.a[i].shape[0], psi.a[i].shape[0], dtype=torch.complex128, device="cuda")
ptb.enforce_qsparsity(c, [psi.qbonds[i], -psi.qbonds[i]])
lb = lblocks[i]
bplan = _sweep.bond_plan(psi.qbonds[i], psi.qbonds[i], qh[i], c, lb, lb)


def site():
    return _sweep.local_hamiltonian_step(lblocks[i], rblocks[i], h.a[i], psi.a[i], 0.01j, k, site_plan)


def bond():
    return _sweep.local_bond_step(lb, lb, c, 0.01j, k, bplan)


def env():
    return _sweep.env_step_left(psi, h, i, lblocks[i])


for name, fn in (("site problem (d=4)", site), ("bond problem", bond), ("environment update", env)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"== {name}: a {tuple(psi.a[i].shape)}, wall {wall * 1e3:.2f} ms, kernel busy time {tot / 1e3:.2f} ms in {sum(e.count for e in rows)} launches")
    for e in rows[:10]:
        print(f"{e.device_time_total / 1e3:9.3f} ms {e.count:5d}x {e.device_time_total / e.count:8.1f} us  {e.key[:90]}")
