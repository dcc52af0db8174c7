"""Where the end-to-end (host-buffer) matvec spends its time: PCIe rates (contiguous / 2-D, both directions),
cost of the pinned result buffer, the C call alone, and the sliced GEMMs against the unsliced ones.

    python tools/e2e_probe.py [--D 2048]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200 as ptb
from pytenet_b200 import _lib, _device as dev
from bench import host_inputs

ap = argparse.ArgumentParser()
ap.add_argument("--D", type=int, default=2048)
args = ap.parse_args()
D, d, chi = args.D, 4, 5
lib = _lib.load()
device = dev.default_device()
a, w, l, r = host_inputs(D, d, chi, seed=1, pinned=True)
res = {}


def wall(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


# PCIe rates
ta = torch.from_numpy(a); tl = torch.from_numpy(l)
da = torch.empty_like(ta, device=device); dl = torch.empty_like(tl, device=device)
res["pinned_inputs"] = bool(ta.is_pinned() and tl.is_pinned())
s = wall(lambda: dl.copy_(tl, non_blocking=True)); res["h2d_contig_GBs"] = l.nbytes / s / 1e9
hp = torch.empty(tl.shape, dtype=tl.dtype, pin_memory=True)
s = wall(lambda: hp.copy_(dl, non_blocking=True)); res["d2h_contig_GBs"] = l.nbytes / s / 1e9
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1):
        dl.copy_(tl, non_blocking=True)
    with torch.cuda.stream(s2):
        hp.copy_(dl, non_blocking=True)
s = wall(both); res["bidirectional_GBs_each"] = l.nbytes / s / 1e9
s = wall(lambda: torch.empty((D, d, D), dtype=torch.complex128, pin_memory=True)); res["pinned_alloc_ms"] = s * 1e3

# the public call and the bare C call with a preallocated result buffer
s = wall(lambda: ptb.apply_local_hamiltonian(a, w, l, r), n=4); res["public_call_ms"] = s * 1e3
dims = (D, d, D, chi, chi, d, D, D)
nbytes = lib.ptb_apply_local_hamiltonian_host_workspace_bytes(1, 0, *dims)
ws = dev.workspace(nbytes, device, tag="host")
host = torch.empty((D, d, D), dtype=torch.complex128, pin_memory=True)
def bare():
    st = lib.ptb_apply_local_hamiltonian_host(1, 0, a.ctypes.data, w.ctypes.data, l.ctypes.data, r.ctypes.data,
                                              host.data_ptr(), *dims, ws.data_ptr(), nbytes, dev.stream_ptr(device))
    assert st == 0
s = wall(bare, n=4); res["c_call_ms"] = s * 1e3

# device-resident matvec for reference, and the sliced GEMMs of the host form
ad, wd, ld, rd = (torch.from_numpy(x).to(device) for x in (a, w, l, r))
def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res["device_matvec_ms"] = ev(lambda: ptb.apply_local_hamiltonian(ad, wd, ld, rd))
t1 = torch.empty((D * d, chi * D), dtype=torch.complex128, device=device)
part = torch.empty(8 * D * d * D * 16, dtype=torch.uint8, device=device)
def step1(ns):
    ks = D // ns
    for c in range(ns):
        k0 = c * ks
        st = lib.ptb_gemm_splitk(1, 0, 0, 0, D * d, chi * D, ks, ad.data_ptr() + k0 * 16, D,
                                 rd.data_ptr() + k0 * chi * D * 16, chi * D, t1.data_ptr(), chi * D, 1, 0, 0, 0,
                                 1 if c else 0, 0, part.data_ptr(), part.numel(), dev.stream_ptr(device))
        assert st == 0
for ns in (1, 2, 4, 8):
    res[f"step1_{ns}_slices_ms"] = ev(lambda: step1(ns))
t2 = torch.randn((D, chi * d, D), dtype=torch.complex128, device=device)
out = torch.empty((D, d, D), dtype=torch.complex128, device=device)
def step3(nb):
    ms = D // nb
    for b in range(nb):
        m0 = b * ms
        st = lib.ptb_gemm_splitk(1, 1, 0, 0, ms, d * D, D * chi, ld.data_ptr() + m0 * 16, D, t2.data_ptr(), d * D,
                                 out.data_ptr() + m0 * d * D * 16, d * D, 1, 0, 0, 0, 0, 0, part.data_ptr(),
                                 part.numel(), dev.stream_ptr(device))
        assert st == 0
for nb in (1, 2, 4):
    res[f"step3_{nb}_blocks_ms"] = ev(lambda: step3(nb))
print(json.dumps({"e2e_probe": res}, indent=1))
