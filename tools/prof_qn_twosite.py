"""cProfile + device-phase split of a two-site TDVP step with quantum numbers (BASELINE config-3 algorithm,
Fermi-Hubbard, (N, Sz) sectors) through the public driver.

    python tools/prof_qn_twosite.py [L] [D] [k] [tol]
"""
import cProfile, io, json, os, pstats, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import _prof
warnings.simplefilter("ignore")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
D = int(sys.argv[2]) if len(sys.argv) > 2 else 256
k = int(sys.argv[3]) if len(sys.argv) > 3 else 25
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-6
h = ptb.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.0)
rng = np.random.default_rng(11)
psi = ptb.MPS.construct_random(L, h.qsite, ptb.encode_quantum_number_pair(L, 0), max_vdim=D, rng=rng)
psi.orthonormalize(mode="left"); psi.orthonormalize(mode="right")
ptb.tdvp_twosite(h, psi, 0.02j, 1, numiter_lanczos=k, tol_split=tol)
torch.cuda.synchronize()
res = {"L": L, "D": D, "k": k, "tol_split": tol}
_prof.enable(True)
t0 = time.perf_counter()
ptb.tdvp_twosite(h, psi, 0.02j, 1, numiter_lanczos=k, tol_split=tol)
torch.cuda.synchronize()
res["seconds_per_step"] = time.perf_counter() - t0
res["device_seconds_by_phase"] = {key: v / 1e3 for key, v in _prof.report().items()}
res["bond_dims"] = psi.bond_dims
_prof.enable(False)
print(json.dumps({"qn_twosite": res}))
pr = cProfile.Profile(); pr.enable()
ptb.tdvp_twosite(h, psi, 0.02j, 1, numiter_lanczos=k, tol_split=tol)
torch.cuda.synchronize()
pr.disable()
for key, cnt in (("cumulative", 60), ("tottime", 30)):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(cnt)
    print(s.getvalue()[:11000])
