"""Per-kernel device time (torch.profiler / CUPTI) of ONE eager time step of the README config (BASELINE config 1)
and of one METTS sample (config 5): what the launch-latency regime consists of."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PYTENET_B200_GRAPHS"] = "0"
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import pytenet_b200 as ptb
warnings.simplefilter("ignore")
z = np.load(os.path.join(ROOT, "tests", "golden", "tdvp_xxz_L10.npz"))
n = int(z["h/nsites"])
h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
psi = ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)], [z[f"psi0/a{i}"] for i in range(n)])
dt = complex(z["dt"]); k = int(z["k"])
ptb.tdvp_singlesite(h, psi, dt, 3, numiter_lanczos=k)
torch.cuda.synchronize()


def show(title, fn):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"== {title}: kernel busy time {tot / 1e3:.3f} ms in {sum(e.count for e in rows)} launches")
    for e in rows[:16]:
        print(f"{e.device_time_total / 1e3:8.3f} ms {e.count:5d}x {e.device_time_total / e.count:7.1f} us  {e.key[:100]}")


show("README config, one tdvp_singlesite step (eager)", lambda: ptb.tdvp_singlesite(h, psi, dt, 1, numiter_lanczos=k))
hm = ptb.ising_1d_mpo(64, 1.0, 0.8, -0.375)
rng = np.random.default_rng(1000)
kw = dict(numsteps=10, numiter_lanczos=8, tol_split=1e-10)
ptb.metts_energy_samples(hm, 1.0, 1, rng, **kw)
show("METTS Ising L=64, one sample", lambda: ptb.metts_energy_samples(hm, 1.0, 1, rng, **kw))
