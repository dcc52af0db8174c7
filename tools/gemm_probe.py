"""Steps 1 and 3 of the headline matvec (a (2048,4,2048), chi = 5, complex128) as bare GEMM launches -- the command
behind the ncu traffic / duration captures of the dominant kernel:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:gemm_ws -s 2 -c 2 python tools/gemm_probe.py
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytenet_b200 import _device as dev

D, d, chi = 2048, 4, 5
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a = torch.randn(D * d, D, dtype=torch.complex128, device="cuda")
r = torch.randn(D, chi * D, dtype=torch.complex128, device="cuda")
t1 = torch.empty(D * d, chi * D, dtype=torch.complex128, device="cuda")
l = torch.randn(D * chi, D, dtype=torch.complex128, device="cuda")
t2 = torch.randn(D * chi, d * D, dtype=torch.complex128, device="cuda")
o = torch.empty(D, d * D, dtype=torch.complex128, device="cuda")
for _ in range(reps):
    dev.gemm(a, r, out=t1)
    dev.gemm(l, t2, trans_a=True, out=o)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
e0.record(); dev.gemm(a, r, out=t1); e1.record(); dev.gemm(l, t2, trans_a=True, out=o); e2.record()
torch.cuda.synchronize()
print("group", os.environ.get("PTB_GEMM_GROUP", "default"), "step1 ms", e0.elapsed_time(e1), "step3 ms", e1.elapsed_time(e2))
