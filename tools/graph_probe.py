import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pytenet_b200 as ptb
from pytenet_b200 import tdvp
z = np.load("/root/repo/tests/golden/tdvp_xxz_L10.npz")
n = int(z["h/nsites"])
h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
mk = lambda: ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)], [z[f"psi0/a{i}"] for i in range(n)])
dt = complex(z["dt"]); k = int(z["k"])
ptb.tdvp_singlesite(h, mk(), dt, 2, numiter_lanczos=k)
real_init = tdvp._StepGraph.__init__
def spy(self, *a, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    real_init(self, *a, **kw)
    torch.cuda.synchronize(); print("capture s", time.perf_counter() - t0, "ok", self.ok)
tdvp._StepGraph.__init__ = spy
real_replay = tdvp._StepGraph.replay
times = []
def rspy(self):
    t0 = time.perf_counter(); real_replay(self); times.append(time.perf_counter() - t0)
tdvp._StepGraph.replay = rspy
for steps in (10, 100):
    times.clear()
    psi = mk()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ptb.tdvp_singlesite(h, psi, dt, steps, numiter_lanczos=k)
    torch.cuda.synchronize(); print(steps, "steps total", time.perf_counter() - t0, "replays", len(times), "mean replay ms", 1e3 * np.mean(times), "min", 1e3 * np.min(times))
