"""cProfile of the quantum-number single-site TDVP step (sector path) at D = 2048."""
import cProfile, io, os, pstats, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
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
pr = cProfile.Profile(); pr.enable()
ptb.tdvp_singlesite(h, psi, 0.02j, 1, numiter_lanczos=k)
torch.cuda.synchronize()
pr.disable()
for key, cnt in (("cumulative", 45), ("tottime", 25)):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(cnt)
    print(s.getvalue()[:8500])
