"""A few sector-packed matvecs at the BASELINE config-3 shape (a (2048,16,2048), h2 (6,16,16,6), 37 (N,Sz)
sectors) -- the command the ncu captures of the grouped GEMM kernel are taken from:

    ncu --set full --clock-control none --import-source on -k regex:gemm_grouped -s 4 -c 2 -o gpurun_out/prof \
        python tools/packed_probe.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import hamiltonian as ham
from pytenet_b200.sector_packed import PackedHeffPlan

D = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
qsite, qb, wbulk, _, _ = ham._fermi_hubbard_bulk(1.0, 4.0, 0.0)
qsite = np.array(qsite); qb = np.array(qb)
qs2 = np.add.outer(qsite, qsite).reshape(-1)
w2 = np.einsum("kpqm,mrsn->kprqsn", wbulk, wbulk).reshape(6, 16, 16, 6)
cand = [(dn, ds) for dn in range(-4, 5) for ds in range(-4, 5) if (dn + ds) % 2 == 0]
wts = np.array([np.exp(-(dn ** 2 + ds ** 2) / (2 * 1.6 ** 2)) for dn, ds in cand])
sizes = np.floor(wts / wts.sum() * D).astype(int)
sizes[np.argmax(sizes)] += D - sizes.sum()
q = np.sort(np.concatenate([np.full(sz, ptb.encode_quantum_number_pair(32 + dn, ds)) for (dn, ds), sz in zip(cand, sizes)]))
g = torch.Generator(device="cuda").manual_seed(1)


def tensor(shape, qn):
    t = torch.randn(*shape, dtype=torch.complex128, device="cuda", generator=g)
    ptb.enforce_qsparsity(t, qn)
    return t


a = tensor((D, 16, D), [q, qs2, -q]); l = tensor((D, 6, D), [q, qb, -q]); r = tensor((D, 6, D), [q, qb, -q])
plan = PackedHeffPlan(q, qs2, q, qb, qb, cplx=True)
op = plan.bind(torch.from_numpy(w2).cuda(), l, r)
x = op.pack(a)
for _ in range(reps):
    y = op(x)
torch.cuda.synchronize()
print("packed matvecs done", plan.flop_counts())
