import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pytenet_b200 as ptb
from pytenet_b200.hamiltonian import load_cached_mpo
from pytenet_b200 import sharded_dmrg
from pytenet_b200.block_sparse_util import dense_svd
import pytenet_b200.block_sparse_util as bsu
h = load_cached_mpo("/root/repo/tests/golden/molecular_mpo_N32.npz")
for cabi in (True, False, True, False):
    sharded_dmrg._CABI = cabi
    psi = ptb.MPS.construct_random(32, h.qsite, 16, max_vdim=128, dtype="complex", rng=np.random.default_rng(11))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    en = sharded_dmrg.dmrg_singlesite_sharded(h, psi, 1, numiter_lanczos=10)
    torch.cuda.synchronize(); print("cabi", cabi, "sweep s", time.perf_counter() - t0, en[-1])
for n in (256, 384, 512, 768):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    row = {}
    for name, thr in (("gesvd", 10 ** 9), ("polar", 1)):
        bsu._POLAR_MIN = thr
        dense_svd(a); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3): dense_svd(a)
        torch.cuda.synchronize(); row[name] = (time.perf_counter() - t0) / 3 * 1e3
    print(n, row)
