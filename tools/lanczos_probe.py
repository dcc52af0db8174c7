"""Lanczos vector kernels at the headline vector length (n = 2048*4*2048 complex128): CUDA-event time per
three-term step and achieved HBM GB/s by algorithmic bytes (dot 2n, axpy+norm 4n, scale 2n elements)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytenet_b200 import _lib, _device as dev
lib = _lib.load()
n = 2048 * 4 * 2048
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
vj = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g); vj /= torch.linalg.norm(vj)
vm = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g); vm /= torch.linalg.norm(vm)
vn = torch.empty_like(w)
scal = torch.zeros(4, dtype=torch.float64, device="cuda"); scal[2] = 0.5
scratch = dev.lanczos_scratch(w.device)
st = torch.cuda.current_stream().cuda_stream
def step():
    rc = lib.ptb_lanczos_ortho_step_z(n, w.data_ptr(), vj.data_ptr(), vm.data_ptr(), scal.data_ptr() + 16,
                                      scal.data_ptr(), scal.data_ptr() + 8, vn.data_ptr(), scratch.data_ptr(), st)
    assert rc == 0
for _ in range(3):
    step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nbytes = 8 * n * 16
print(json.dumps({"lanczos_ortho_step": {"n": n, "ms": ms, "algorithmic_bytes": nbytes, "gb_per_s": nbytes / ms / 1e6}}))
