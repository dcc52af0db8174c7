"""dense_svd on the rectangular split matrices of a config-2 two-site sweep (2 Dl x 2 Dr, bonds powers of two):
polar driver vs gesvd per shape (complex128), milliseconds."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pytenet_b200.block_sparse_util as bsu

res = {}
for m, n in [(64, 256), (256, 64), (128, 512), (512, 128), (256, 1024), (1024, 256), (512, 2048), (2048, 512),
             (1024, 2048), (2048, 1024), (2048, 2048), (2048, 4096), (4096, 2048)]:
    a = torch.randn(m, n, dtype=torch.complex128, device="cuda")
    row = {}
    for name, pm in (("polar", 1), ("gesvd", 10 ** 9)):
        bsu._POLAR_MIN = pm
        bsu.dense_svd(a); torch.cuda.synchronize()
        t0 = time.perf_counter(); bsu.dense_svd(a); torch.cuda.synchronize()
        row[name + "_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    res[f"{m}x{n}"] = row
    print(f"{m}x{n}", row, flush=True)
print(json.dumps({"svd_rect_bench_complex128": res}))
