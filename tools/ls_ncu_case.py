"""A few launches of the one-kernel local step for ncu (METTS shape and README bulk shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pytenet_b200 import _sweep
rng = np.random.default_rng(3)
cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
crand = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)
def herm(D, chi):
    e = crand(D, chi, D)
    return e + e.conj().transpose(2, 1, 0)
for (Dl, d, Dr, cl, cr, k) in [(4, 4, 4, 3, 3, 8), (16, 2, 28, 5, 5, 5)]:
    l, r = cu(herm(Dl, cl)), cu(herm(Dr, cr))
    w = rng.normal(size=(cl, d, d, cr)); w = w + w.transpose(0, 2, 1, 3); w[np.abs(w) < 0.9] = 0
    x = cu(crand(Dl * d * Dr))
    for _ in range(3):
        _sweep._small_local_step(x, cu(w), l, r, (Dl, d, Dr, cl, cr), k, 0.05j)
    torch.cuda.synchronize()
