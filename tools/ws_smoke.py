"""Quick standalone check of the warp-specialised TMA GEMM (engine 2) against torch fp64.
Run under a short `timeout`; prints one line per case so a hang is easy to localise."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pytenet_b200 import _lib, _device as dev
lib = _lib.load()
eng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
assert lib.ptb_set_gemm_engine(eng) == 0
torch.manual_seed(0)
cases = [(torch.complex128, 128, 64, 16), (torch.complex128, 128, 64, 64), (torch.complex128, 256, 128, 128),
         (torch.complex128, 130, 70, 37), (torch.complex128, 1000, 900, 300), (torch.float64, 128, 128, 16),
         (torch.float64, 256, 384, 128), (torch.float64, 130, 70, 38), (torch.complex128, 4096, 5120, 1024)]
for dt, M, N, K in cases:
    for ta in (False, True):
        for tb in (False, True):
            for cj in ((False, True) if dt.is_complex else (False,)):
                a = torch.randn((K, M) if ta else (M, K), dtype=dt, device="cuda")
                b = torch.randn((N, K) if tb else (K, N), dtype=dt, device="cuda")
                t0 = time.time()
                c = dev.gemm(a, b, trans_a=ta, trans_b=tb, conj_b=cj)
                torch.cuda.synchronize()
                oa = a.T if ta else a
                ob = b.T if tb else b
                if cj:
                    ob = ob.conj()
                ref = oa @ ob
                err = (torch.linalg.norm(c - ref) / torch.linalg.norm(ref)).item()
                print(f"{str(dt):18s} M={M} N={N} K={K} ta={int(ta)} tb={int(tb)} cj={int(cj)} err={err:.2e} "
                      f"{(time.time()-t0)*1e3:.1f} ms", flush=True)
                assert err < 1e-13
print("ws smoke ok")
