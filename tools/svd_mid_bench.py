"""Mid-size SVD blocks (64 <= k <= 320, complex128): cuSOLVER gesvd (torch.linalg), the polar driver behind
ptb_svd_polar, and the batched Jacobi kernel (ptb_block_svd, through block_sparse_svd on a single-sector matrix),
time per block and singular-value agreement with LAPACK.  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pytenet_b200.block_sparse_util as bsu

res = {}
for n in (64, 80, 96, 128, 160, 192, 224, 256, 320):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    s_ref = np.linalg.svd(a.cpu().numpy(), compute_uv=False)
    row = {}
    bsu._POLAR_MIN = 1
    fns = [("gesvd", lambda: torch.linalg.svd(a, full_matrices=False, driver="gesvd")),
           ("polar", lambda: bsu.dense_svd(a))]
    q = np.zeros(n, dtype=np.int64)
    if n * n * 2 * 16 <= 200 * 1024:
        fns.append(("jacobi", lambda: bsu.block_sparse_svd(a, q, q)))
    for name, fn in fns:
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        row[name + "_ms"] = (time.perf_counter() - t0) * 1e3 / 3
        s = out[1]
        s = s.cpu().numpy() if isinstance(s, torch.Tensor) else s
        row[name + "_sigma_err"] = float(np.max(np.abs(np.sort(s)[::-1] - s_ref)) / s_ref[0])
    res[str(n)] = row
print(json.dumps({"svd_mid_bench_complex128": res}))
