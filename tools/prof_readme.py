"""cProfile of the steady-state README-config TDVP run (host-side cost of the launch-latency-bound path);
warm-up (CUDA / cuSOLVER initialisation, module loading) is excluded."""
import cProfile, io, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb

z = np.load(os.path.join(ROOT, "tests", "golden", "tdvp_xxz_L10.npz"))
n = int(z["h/nsites"])
h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
mk = lambda: ptb.MPS.from_tensors(z["psi0/qsite"], [z[f"psi0/qb{i}"] for i in range(n + 1)],
                                  [z[f"psi0/a{i}"] for i in range(n)])
dt = complex(z["dt"]); k = int(z["k"])
import warnings; warnings.simplefilter("ignore")
ptb.tdvp_singlesite(h, mk(), dt, 3, numiter_lanczos=k)
psi = mk()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
ptb.tdvp_singlesite(h, psi, dt, 30, numiter_lanczos=k)
torch.cuda.synchronize()
pr.disable()
for key, cnt in (("cumulative", 45), ("tottime", 30)):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(cnt)
    print(s.getvalue()[:9000])
