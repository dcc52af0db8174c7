"""Time of the large dense SVD of a two-site split (2048 x 2048 complex128 at the config-2 shape) through the
available drivers: cuSOLVER gesvd / gesvdj / gesvda (torch.linalg) and the polar-decomposition driver behind
ptb_svd_polar (block_sparse_util.dense_svd).  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pytenet_b200.block_sparse_util import dense_svd

res = {}
for n in (1024, 2048, 4096):
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda")
    row = {}
    for name, fn in [("gesvd", lambda: torch.linalg.svd(a, full_matrices=False, driver="gesvd")),
                     ("gesvdj", lambda: torch.linalg.svd(a, full_matrices=False, driver="gesvdj")),
                     ("polar_gesvdp", lambda: dense_svd(a))]:
        if n == 4096 and name == "gesvdj":
            continue
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        row[name + "_ms"] = (time.perf_counter() - t0) * 1e3
        u, s, vh = out
        row[name + "_recon_err"] = (torch.linalg.norm((u * s) @ vh - a) / torch.linalg.norm(a)).item()
    res[str(n)] = row
print(json.dumps({"svd_bench_complex128": res}))
