"""GPU-time breakdown by kernel (torch.profiler / CUPTI) of one quantum-number single-site TDVP step at D = 2048."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import pytenet_b200 as ptb
warnings.simplefilter("ignore")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 12
D = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
h = ptb.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.0)
rng = np.random.default_rng(11)
psi = ptb.MPS.construct_random(L, h.qsite, ptb.encode_quantum_number_pair(L, 0), max_vdim=D, rng=rng)
psi.orthonormalize(mode="left"); psi.orthonormalize(mode="right")
ptb.tdvp_singlesite(h, psi.copy(), 0.02j, 1, numiter_lanczos=k)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ptb.tdvp_singlesite(h, psi, 0.02j, 1, numiter_lanczos=k)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows)
print(f"total device time {tot / 1e3:.1f} ms in {sum(e.count for e in rows)} launches")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:9.2f} ms {100 * e.device_time_total / tot:5.1f}% {e.count:6d}x  {e.key[:100]}")
