"""Where the time of the one-kernel local step (csrc/lanczos_small.cu) and of the batched sector QR / SVD kernels goes:
device time per launch (CUPTI through torch.profiler, 40 back-to-back launches per configuration) for the METTS and
README-config shapes, with the Lanczos run alone (apply_expm = 0) and with the k x k solve + combination, for several
numbers of iterations; the stand-alone k x k solve (ptb_krylov_expm_apply); QR / SVD of single blocks."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import pytenet_b200 as ptb
from pytenet_b200 import _sweep, _lib, _device as dev, block_sparse_util as bsu
warnings.simplefilter("ignore")
rng = np.random.default_rng(3)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def crand(*s):
    return rng.normal(size=s) + 1j * rng.normal(size=s)


def herm(D, chi):
    e = crand(D, chi, D)
    return e + e.conj().transpose(2, 1, 0)


def device_us(fn, reps=40):
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
    rows = [e for e in prof.key_averages() if e.device_time_total > 0]
    return {e.key[:60]: e.device_time_total / e.count for e in rows if e.count >= reps // 2}


def phases():
    """Clock cycles per phase of the one-kernel local step (thread 0 of CTA 0; tools/build_lsprobe.sh)."""
    import ctypes
    so = os.path.join(ROOT, "tools", "_probe", "liblsprobe.so")
    if not os.path.exists(so):
        print("no", so); return
    L = ctypes.CDLL(so)
    i64, vp, dbl, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_double, ctypes.c_int
    L.ptb_local_step_small.argtypes = [ci, vp, vp, ci, vp, vp] + [i64] * 5 + [ci, vp, vp, ci, dbl, dbl, ci, vp, vp, ctypes.c_size_t, vp]
    names = ["setup", "step1", "step2", "step3", "alpha_sum", "axpy+beta_sum", "push+sync", "kxk solve", "combine"]
    for (Dl, d, Dr, cl, cr, k) in [(4, 4, 4, 3, 3, 8), (4, 2, 8, 5, 5, 5), (8, 2, 16, 5, 5, 5), (16, 2, 28, 5, 5, 5), (16, 1, 28, 5, 5, 5)]:
        l, r = cu(herm(Dl, cl)), cu(herm(Dr, cr))
        w = rng.normal(size=(cl, d, d, cr)); w = w + w.transpose(0, 2, 1, 3); w[np.abs(w) < 0.9] = 0
        wd = cu(w) if d > 1 else None
        n = Dl * d * Dr
        x = cu(crand(n)); V = torch.empty((k, n), dtype=x.dtype, device="cuda")
        scal = torch.empty(2 * k, dtype=torch.float64, device="cuda"); out = torch.empty(n, dtype=x.dtype, device="cuda")
        ws = torch.empty(64, dtype=torch.float64, device="cuda")
        buf = (ctypes.c_longlong * 16)()
        reps = 20
        for it in range(2):
            L.lsprobe_read(buf, 1)
            for _ in range(reps):
                st = L.ptb_local_step_small(1, x.data_ptr(), wd.data_ptr() if wd is not None else None, 0, l.data_ptr(), r.data_ptr(),
                                            Dl, d, Dr, cl, cr, k, V.data_ptr(), scal.data_ptr(), 1, 0.0, -0.05, 1, out.data_ptr(),
                                            ws.data_ptr(), 512, None)
                assert st == 0, st
            L.lsprobe_read(buf, 0)
        tot = sum(buf[i] for i in range(9)) / reps
        print(f"dims {(Dl, d, Dr, cl, cr)} k={k}: total {tot:.0f} cycles per launch; " +
              ", ".join(f"{names[i]} {buf[i] / reps:.0f}" for i in range(9)))


print("== phase clocks (cycles per launch, summed over the iterations)")
phases()
print("== one-kernel local step: us per launch")
for (Dl, d, Dr, cl, cr) in [(4, 4, 4, 3, 3), (4, 2, 4, 3, 3), (4, 2, 8, 5, 5), (8, 2, 16, 5, 5), (16, 2, 28, 5, 5), (28, 2, 16, 5, 5),
                            (16, 1, 28, 5, 5)]:
    l, r = cu(herm(Dl, cl)), cu(herm(Dr, cr))
    w = rng.normal(size=(cl, d, d, cr)); w = w + w.transpose(0, 2, 1, 3); w[np.abs(w) < 0.9] = 0
    wd = cu(w) if d > 1 else None
    x = cu(crand(Dl * d * Dr))
    for k in (1, 2, 5, 8):
        for dt in (None, 0.05j):
            res = device_us(lambda: _sweep._small_local_step(x, wd, l, r, (Dl, d, Dr, cl, cr), k, dt,
                                                            *((torch.empty((k, x.numel()), dtype=x.dtype, device="cuda"),
                                                               torch.empty(2 * k, dtype=torch.float64, device="cuda")) if dt is None else ())))
            t = [v for kk, v in res.items() if "lanczos_small" in kk]
            print(f"dims {(Dl, d, Dr, cl, cr)} k={k} expm={dt is not None}: {t[0] if t else float('nan'):7.1f} us")

print("== stand-alone k x k solve (ptb_krylov_expm_apply), n = 64")
lib = _lib.load()
for k in (2, 5, 8, 25):
    n = 64
    alpha = rng.normal(size=k) * 3; beta = np.abs(rng.normal(size=max(k - 1, 0))) + 0.1
    scal = cu(np.concatenate([[1.7], alpha, beta, np.zeros(1)]))
    V = cu(crand(k, n)); out = torch.empty(n, dtype=torch.complex128, device="cuda")
    cws = torch.zeros(lib.ptb_krylov_expm_workspace_bytes() // 8, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    res = device_us(lambda: lib.ptb_krylov_expm_apply(1, n, k, V.data_ptr(), n, scal.data_ptr(), 0.0, -0.05, 1, cws.data_ptr(),
                                                      out.data_ptr(), st))
    print(f"k={k}: " + ", ".join(f"{kk[:40]} {v:.1f} us" for kk, v in res.items()))

print("== batched QR / SVD kernels on single blocks (zero quantum numbers)")
for (m, n) in [(8, 8), (16, 16), (32, 28), (56, 16), (8, 2), (64, 64)]:
    a = cu(crand(m, n))
    q0, q1 = np.zeros(m, dtype=np.int64), np.zeros(n, dtype=np.int64)
    res = device_us(lambda: bsu.block_sparse_qr(a, q0, q1))
    print(f"qr {m}x{n}: " + ", ".join(f"{kk[:40]} {v:.1f} us" for kk, v in res.items()))
    res = device_us(lambda: bsu.block_sparse_svd(a, q0, q1))
    print(f"svd {m}x{n}: " + ", ".join(f"{kk[:40]} {v:.1f} us" for kk, v in res.items()))
