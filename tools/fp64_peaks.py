"""
Measure the FP64 denominators on the B200 this runs on (MEASURED_PEAKS.json has
HBM and bf16 only): cuBLAS DGEMM / ZGEMM through torch.matmul, the raw
DMMA.8x8x4 and DFMA issue peaks (register-resident probes), and the DMMA GEMM
engine of this repository on the matvec's own GEMM shapes.

    python tools/fp64_peaks.py [--out gpurun_out/fp64_peaks.json]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytenet_b200 import _lib, _device as dev  # noqa: E402


def time_ms(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fp64_peaks.json"))
    args = ap.parse_args()
    lib = _lib.load()
    res = {"gpu": torch.cuda.get_device_name(0), "sm_count": torch.cuda.get_device_properties(0).multi_processor_count}

    # cuBLAS through torch (library reference numbers, not the product path)
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    ms = time_ms(lambda: torch.matmul(a, b))
    res["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    del a, b
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda"); b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    ms = time_ms(lambda: torch.matmul(a, b))
    res["cublas_zgemm_4096_tflops"] = 8 * n ** 3 / ms / 1e9
    n = 8192
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda"); b = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    ms = time_ms(lambda: torch.matmul(a, b), warm=1, reps=3)
    res["cublas_zgemm_8192_tflops"] = 8 * n ** 3 / ms / 1e9
    del a, b

    # raw pipe probes
    sms = res["sm_count"]
    out = torch.empty(sms * 8 * 256, dtype=torch.float64, device="cuda")
    fl = ctypes.c_double(0)
    st = torch.cuda.current_stream().cuda_stream
    for name, use, blocks in [("dmma_probe_tflops", 1, sms * 2), ("dmma_probe_occ8_tflops", 1, sms * 8),
                              ("dfma_probe_tflops", 0, sms * 8)]:
        ms = time_ms(lambda: lib.ptb_probe_fp64_pipe(use, blocks, 4000, out.data_ptr(), ctypes.byref(fl), st))
        res[name] = fl.value / ms / 1e9

    # this repository's engines on the matvec shapes (D, d, chi): 1 = cp.async kernel, 2 = TMA kernel
    shapes = []
    for eng, (D, d, chi) in [(e, s) for e in (1, 2) for s in [(1024, 4, 5), (2048, 2, 5), (2048, 4, 5)]]:
        a = torch.randn(D * d, D, dtype=torch.complex128, device="cuda")
        r = torch.randn(D, chi * D, dtype=torch.complex128, device="cuda")
        t1 = torch.empty(D * d, chi * D, dtype=torch.complex128, device="cuda")
        ms1 = time_ms(lambda: dev.gemm(a, r, out=t1, engine=eng), warm=1, reps=3)
        l = torch.randn(D * chi, D, dtype=torch.complex128, device="cuda")
        t2 = torch.randn(D * chi, d * D, dtype=torch.complex128, device="cuda")
        o = torch.empty(D, d * D, dtype=torch.complex128, device="cuda")
        ms3 = time_ms(lambda: dev.gemm(l, t2, trans_a=True, out=o, engine=eng), warm=1, reps=3)
        f = 8.0 * D * d * D * chi * D
        cub1 = time_ms(lambda: torch.matmul(a, r, out=t1), warm=1, reps=3)
        cub3 = time_ms(lambda: torch.matmul(l.T, t2, out=o), warm=1, reps=3)
        shapes.append({"engine": eng, "D": D, "d": d, "chi": chi, "step1_NN_tflops": f / ms1 / 1e9, "step3_TN_tflops": f / ms3 / 1e9,
                       "cublas_step1_tflops": f / cub1 / 1e9, "cublas_step3_tflops": f / cub3 / 1e9,
                       "step1_ms": ms1, "step3_ms": ms3})
        del a, r, t1, l, t2, o
    res["engine"] = shapes
    for eng in (1, 2):
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
        c = torch.empty(n, n, dtype=torch.float64, device="cuda")
        ms = time_ms(lambda: dev.gemm(a, b, out=c, engine=eng), warm=1, reps=3)
        res[f"engine{eng}_dgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
        del a, b, c
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty(n, n, dtype=torch.float64, device="cuda")
    ms = time_ms(lambda: dev.gemm(a, b, out=c), warm=1, reps=3)
    res["engine_dgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
