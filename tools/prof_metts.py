"""cProfile of steady-state METTS sampling (config 5: Ising L=64, two-site TDVP in imaginary time)."""
import cProfile, io, os, pstats, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
warnings.simplefilter("ignore")
h = ptb.ising_1d_mpo(64, 1.0, 0.8, -0.375)
rng = np.random.default_rng(1000)
kw = dict(numsteps=10, numiter_lanczos=8, tol_split=1e-10)
ptb.metts_energy_samples(h, 1.0, 1, rng, **kw)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
ptb.metts_energy_samples(h, 1.0, 1, rng, **kw)
torch.cuda.synchronize()
pr.disable()
for key, cnt in (("cumulative", 40), ("tottime", 22)):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(cnt)
    print(s.getvalue()[:8000])
