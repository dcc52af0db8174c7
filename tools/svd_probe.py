"""cuSOLVER SVD drivers on the two-site split shapes (complex128): time and accuracy of gesvd vs gesvdj."""
import json, sys, time
import torch
for n in (512, 1024, 2048):
    g = torch.Generator(device="cuda").manual_seed(n)
    a = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=g)
    # realistic spectrum: exponentially decaying singular values
    u, _ = torch.linalg.qr(a); v, _ = torch.linalg.qr(a.mH)
    s = torch.exp(-torch.arange(n, device="cuda", dtype=torch.float64) * (30.0 / n))
    m = (u * s) @ v.mH
    res = {"n": n}
    for drv in ("gesvd", "gesvdj"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        uu, ss, vv = torch.linalg.svd(m, full_matrices=False, driver=drv)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        rec = (torch.linalg.norm((uu * ss) @ vv - m) / torch.linalg.norm(m)).item()
        serr = torch.max(torch.abs(ss - s) / s).item()
        orth = torch.linalg.norm(uu.mH @ uu - torch.eye(n, device="cuda", dtype=uu.dtype)).item()
        res[drv] = {"s": dt, "recon": rec, "max_rel_sigma_err": serr, "orth": orth}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    q, r = torch.linalg.qr(m)
    torch.cuda.synchronize(); res["qr_s"] = time.perf_counter() - t0
    print(json.dumps(res), flush=True)
