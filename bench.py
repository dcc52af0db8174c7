#!/usr/bin/env python
"""
Benchmark of the effective-Hamiltonian hot path (BASELINE.json metric:
"effective-H matvec GFLOP/s ... at D=2048").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one effective-H matvec out = L.W.A.R on the headline shape
(two-site XXZ Heisenberg: a (2048,4,2048) complex128, h2 (5,4,4,5) float64,
l / r (2048,5,2048) complex128; SURVEY.md section 8d "Headline shape").
GFLOP/s uses F_alg = 8 (Dl d Dr chi Dr' + chi^2 d^2 Dl Dr' + Dl' Dl chi d Dr'), the
dense all-complex flop count of the reference's own contraction order.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the
reference (oracle/, NumPy + OpenBLAS on all host cores) on a bounded sample.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "xxz_twosite_heff_matvec_D2048"
D_HEAD, d_HEAD, CHI = 2048, 4, 5
METRIC = "effective-H matvec GFLOP/s at D=2048 (two-site XXZ, complex128)"


def f_alg(D, d, chi):
    return 8.0 * (D * d * D * chi * D + chi * d * d * chi * D * D + D * D * chi * d * D)


def xxz_two_site_w():
    """h2 = merge of two bulk XXZ MPO tensors (J=1, Delta=0.8, h=-0.1): (5,4,4,5) float64."""
    from pytenet_b200 import hamiltonian as ham
    _, _, w, _, _ = ham._xxz_bulk(1.0, 0.8, -0.1)
    c0, p0, q0, c1 = w.shape
    t = w.reshape(c0 * p0 * q0, c1) @ w.reshape(c1, -1)
    t = t.reshape(c0, p0, q0, p0, q0, c1).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(t).reshape(c0, p0 * p0, q0 * q0, c1)


def host_inputs(D, d, chi, seed, pinned=False):
    """Synthetic inputs as in pytenet/mps.py:55: crandn / sqrt(Dl d Dr), seeded."""
    rng = np.random.default_rng(seed)

    def crandn(shape):
        return (rng.normal(size=shape) + 1j * rng.normal(size=shape)) / np.sqrt(2)

    a = crandn((D, d, D)) / np.sqrt(D * d * D)
    l = crandn((D, chi, D)) / np.sqrt(D)
    r = crandn((D, chi, D)) / np.sqrt(D)
    w = xxz_two_site_w()
    assert w.shape == (chi, d, d, chi)
    if pinned:
        import torch
        out = []
        for x in (a, w, l, r):
            t = torch.empty(x.shape, dtype=torch.from_numpy(x).dtype, pin_memory=True)
            t.numpy()[...] = x
            out.append(t.numpy())
        return out
    return a, w, l, r


def host_inputs_onesite(D, chi, seed):
    """One-site XXZ inputs: a (D,2,D), bulk MPO tensor (5,2,2,5)."""
    from pytenet_b200 import hamiltonian as ham
    rng = np.random.default_rng(seed)

    def crandn(shape):
        return (rng.normal(size=shape) + 1j * rng.normal(size=shape)) / np.sqrt(2)

    _, _, w, _, _ = ham._xxz_bulk(1.0, 0.8, -0.1)
    return (crandn((D, 2, D)) / np.sqrt(2 * D * D), np.ascontiguousarray(w), crandn((D, chi, D)) / np.sqrt(D),
            crandn((D, chi, D)) / np.sqrt(D))


def block_sparse_block(torch, ptb, device, time_ms, peak_tflops, D=2048):
    """BASELINE config 3 shape (two-site Fermi-Hubbard, (N,Sz) sectors, bonds grouped by sector as the sweeps
    leave them): the dense device matvec, the banded work lists over dense-layout tensors (round 1,
    sectors.HeffSectorPlan) and the sector-packed grouped GEMM (sector_packed.PackedHeffPlan, what the sweeps use)."""
    from pytenet_b200 import hamiltonian as ham
    from pytenet_b200.sectors import HeffSectorPlan
    from pytenet_b200.sector_packed import PackedHeffPlan
    qsite, qb, wbulk, _, _ = ham._fermi_hubbard_bulk(1.0, 4.0, 0.0)
    qsite = np.array(qsite); qb = np.array(qb)
    d1 = len(qsite)
    qs2 = np.add.outer(qsite, qsite).reshape(-1)
    w2 = np.einsum("kpqm,mrsn->kprqsn", wbulk, wbulk).reshape(6, d1 * d1, d1 * d1, 6)
    cand = [(dn, ds) for dn in range(-4, 5) for ds in range(-4, 5) if (dn + ds) % 2 == 0]
    wts = np.array([np.exp(-(dn ** 2 + ds ** 2) / (2 * 1.6 ** 2)) for dn, ds in cand])
    sizes = np.floor(wts / wts.sum() * D).astype(int)
    sizes[np.argmax(sizes)] += D - sizes.sum()
    q = np.sort(np.concatenate([np.full(sz, ptb.encode_quantum_number_pair(32 + dn, ds))
                                for (dn, ds), sz in zip(cand, sizes)]))
    g = torch.Generator(device=device).manual_seed(1)

    def crand(*shape):
        return torch.randn(*shape, dtype=torch.complex128, device=device, generator=g)

    a = crand(D, d1 * d1, D); ptb.enforce_qsparsity(a, [q, qs2, -q])
    l = crand(D, 6, D); ptb.enforce_qsparsity(l, [q, qb, -q])
    r = crand(D, 6, D); ptb.enforce_qsparsity(r, [q, qb, -q])
    w = torch.from_numpy(w2).to(device)
    dense = ptb.apply_local_hamiltonian(a, w, l, r)
    ms_d = time_ms(lambda: ptb.apply_local_hamiltonian(a, w, l, r), reps=2)
    fa = f_alg(D, d1 * d1, 6)
    # ---- round-1 path: banded / segmented GEMMs over the dense layout ----
    plan = HeffSectorPlan(q, qs2, q, qb, qb, cplx=True)
    banded = plan.apply(a, w, l, r)
    err_b = (torch.linalg.norm(banded - dense) / torch.linalg.norm(dense)).item()
    ms_b = time_ms(lambda: plan.apply(a, w, l, r), reps=5)
    fc = plan.flop_counts(nnz_w=int(np.count_nonzero(w2)))
    del banded
    # ---- sector-packed grouped GEMM ----
    pplan = PackedHeffPlan(q, qs2, q, qb, qb, cplx=True)
    op = pplan.bind(w, l, r)
    x = op.pack(a)
    y = op(x)
    err_p = (torch.linalg.norm(op.unpack(y) - dense) / torch.linalg.norm(dense)).item()
    del dense
    ms_p = time_ms(lambda: op(x), reps=10)
    pfc = pplan.flop_counts()
    lib = op.lib
    st = torch.cuda.current_stream().cuda_stream
    n1, n3 = len(pplan.tiles1_host), len(pplan.tiles3_host)
    ms_g1 = time_ms(lambda: lib.ptb_gemm_grouped_v(op.dt, pplan.var1, x.data_ptr(), op.rb.data_ptr(), op.t1.data_ptr(),
                                                   op.tabs["tiles1"].data_ptr(), n1, st), reps=10)
    ms_w = time_ms(lambda: op.wtab.run(lib, op.dt, op.t1, op.t2, st), reps=10)
    ms_g3 = time_ms(lambda: lib.ptb_gemm_grouped_v(op.dt, pplan.var3, op.t2.data_ptr(), op.lp.data_ptr(), op.o.data_ptr(),
                                                   op.tabs["tiles3"].data_ptr(), n3, st), reps=10)
    ms_rp = time_ms(lambda: op.tabs["repack"].run(lib, op.dt, op.o, y, st), reps=10)
    ms_pack = time_ms(lambda: op.pack(a), reps=3)
    ms_unpack = time_ms(lambda: op.unpack(y), reps=3)
    # one Lanczos iteration's vector work on the packed vs the dense vector (HBM-bound, 5 passes: krylov.cu)
    es = 16
    packed = {
        "ms_per_matvec": ms_p, "rel_diff_vs_dense_device_matvec": err_p,
        "flops_exact_sector_blocks": pfc["exact"], "flops_visited": pfc["visited"],
        "visited_over_exact": pfc["visited"] / pfc["exact"],
        "tflops_exact": pfc["exact"] / ms_p / 1e9, "frac_of_fp64_peak_on_exact_flops": pfc["exact"] / ms_p / 1e9 / peak_tflops,
        "tflops_visited_gemm_kernels": pfc["visited"] / (ms_g1 + ms_g3) / 1e9,
        "gemm_kernels_frac_of_fp64_peak": pfc["visited"] / (ms_g1 + ms_g3) / 1e9 / peak_tflops,
        "kernel_ms": {"grouped_gemm_step1": ms_g1, "w_gather": ms_w, "grouped_gemm_step3": ms_g3, "repack": ms_rp},
        "tiles": {"step1": n1, "step3": n3, "tile_variant_step1": pplan.var1, "tile_variant_step3": pplan.var3},
        "once_per_lanczos_run_ms": {"pack": ms_pack, "unpack": ms_unpack},
        "packed_vector_elements": pplan.nX, "dense_vector_elements": D * d1 * d1 * D,
        "intermediate_bytes": {"t1": pplan.nT1 * es, "t2": pplan.nT2 * es, "dense_layout_t1": 16.0 * D * d1 * d1 * 6 * D},
        "speedup_vs_dense": ms_d / ms_p, "speedup_vs_banded_round1": ms_b / ms_p,
    }
    # ---- environment update (contraction_operator_step_right) at the same sector profile, one-site tensor ----
    from pytenet_b200.sectors import EnvSectorPlan
    from pytenet_b200.sector_packed import PackedEnvPlan
    fill = float((a != 0).double().mean().item())
    del a, x, y, op
    w1 = torch.from_numpy(np.ascontiguousarray(wbulk)).to(device)
    a1 = crand(D, d1, D); ptb.enforce_qsparsity(a1, [q, qsite, -q])
    env_dense = ptb.contraction_operator_step_right(a1, a1, w1, r)
    ms_ed = time_ms(lambda: ptb.contraction_operator_step_right(a1, a1, w1, r), reps=2)
    eplan = PackedEnvPlan(q, qsite, q, qb, qb, cplx=True, side="right")
    env_p = eplan.apply(a1, w1, r)
    err_e = (torch.linalg.norm(env_p - env_dense) / torch.linalg.norm(env_dense)).item()
    ms_ep = time_ms(lambda: eplan.apply(a1, w1, r), reps=10)
    bplan = EnvSectorPlan(q, qsite, q, qb, qb, cplx=True)
    ms_eb = time_ms(lambda: bplan.step_right(a1, w1, r), reps=5)
    env = {"op": f"contraction_operator_step_right, a ({D},{d1},{D}), w (6,{d1},{d1},6), same sectors",
           "ms_dense": ms_ed, "ms_banded_round1": ms_eb, "ms_sector_packed": ms_ep,
           "rel_diff_vs_dense": err_e, "speedup_vs_dense": ms_ed / ms_ep, "speedup_vs_banded_round1": ms_eb / ms_ep}
    del env_dense, env_p, a1
    return {"workload": f"two-site Fermi-Hubbard heff matvec a ({D},16,{D}), h2 (6,16,16,6), {int((sizes > 0).sum())} "
                        f"(N,Sz) sectors (Gaussian size profile, max {int(sizes.max())})",
            "tensor_fill": fill, "ms_dense": ms_d, "gflops_alg_dense_path": fa / ms_d / 1e6,
            "sector_packed": packed, "environment_update": env,
            "banded_round1": {"ms_per_matvec": ms_b, "rel_diff_vs_dense": err_b, "flops_visited": fc["visited"],
                              "flops_exact_sector_blocks": fc["exact"], "visited_over_exact": fc["visited"] / fc["exact"],
                              "tflops_exact": fc["exact"] / ms_b / 1e9, "speedup_vs_dense": ms_d / ms_b},
            "gflops_alg_packed_path": fa / ms_p / 1e6}


# ----------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference on the host cores
# ----------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else (os.cpu_count() or 1)
    except Exception:
        return os.cpu_count() or 1


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1, which OpenBLAS honours at load time; the CPU arm is meant to use every
    host core, so the BLAS pool is reset explicitly."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1, user_api="blas")
    except Exception:
        pass


def reference_chain_ops():
    """(module, kind): the reference's own `chain_ops` when an install of it travels with the repo
    (`baseline/_ref`, written by `pip install --no-deps --target baseline/_ref /root/reference`, or a checkout on
    PYTHONPATH), else the oracle restatement.  pytenet 1.3.0's packaging lists `packages = ["pytenet"]` only, so
    the installed tree lacks the `pytenet.hamiltonian` sub-package and `import pytenet` fails in its __init__;
    the hot-path module is therefore loaded from the unmodified installed file under a stub parent package."""
    import importlib
    import types
    for base in (os.path.join(ROOT, "baseline", "_ref"),):
        pkg = os.path.join(base, "pytenet")
        if os.path.exists(os.path.join(pkg, "chain_ops.py")):
            try:
                if "pytenet" not in sys.modules:
                    stub = types.ModuleType("pytenet")
                    stub.__path__ = [pkg]
                    sys.modules["pytenet"] = stub
                return importlib.import_module("pytenet.chain_ops"), "reference"
            except Exception:
                sys.modules.pop("pytenet", None)
    try:
        import pytenet.chain_ops as co           # a full checkout on PYTHONPATH
        return co, "reference"
    except Exception:
        import oracle
        return oracle, "port"


def cpu_time_matvec(D, d, chi, reps=1, impl=None):
    impl = impl or reference_chain_ops()[0]
    use_all_host_threads()
    a, w, l, r = host_inputs(D, d, chi, seed=1)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        impl.apply_local_hamiltonian(a, w, l, r)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_baseline_block():
    """The reference's CPU path on the SAME workload as the GPU arm (D = 2048, always): one matvec, 2.5-6 s."""
    impl, kind = reference_chain_ops()
    cpu_time_matvec(256, d_HEAD, CHI, impl=impl)          # BLAS pool warm-up
    t = cpu_time_matvec(D_HEAD, d_HEAD, CHI, reps=1, impl=impl)
    src = ("pytenet 1.3.0 chain_ops.apply_local_hamiltonian from baseline/_ref" if kind == "reference"
           else "oracle/ NumPy restatement of pytenet/chain_ops.py:237-279")
    return {
        "value": f_alg(D_HEAD, d_HEAD, CHI) / t / 1e9, "unit": "GFLOP/s", "cores": cpu_threads(), "kind": kind,
        "sample": f"1 matvec of the bench workload itself: D={D_HEAD}, d={d_HEAD}, chi={CHI} ({src}, NumPy + "
                  f"OpenBLAS on all host threads, {t:.2f} s)",
        "seconds": t,
    }


def run_reference(args):
    """--impl reference: K timed steps of the reference's CPU path on the bench workload (D = 2048; rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    impl, kind = reference_chain_ops()
    use_all_host_threads()
    steps, warm = args.steps, args.warmup
    D = D_HEAD
    a, w, l, r = host_inputs(D, d_HEAD, CHI, seed=1)
    t0 = time.perf_counter()
    impl.apply_local_hamiltonian(a, w, l, r)
    t_first = time.perf_counter() - t0
    # keep the whole run within a few minutes on a slow host: the untimed warm-up shrinks first, never the workload
    warm_done = 1
    while warm_done < warm and t_first * (steps + warm_done + 1) < 240.0:
        impl.apply_local_hamiltonian(a, w, l, r)
        warm_done += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        impl.apply_local_hamiltonian(a, w, l, r)
    dt = (time.perf_counter() - t0) / steps
    val = f_alg(D, d_HEAD, CHI) / dt / 1e9
    sample = (f"each step = 1 matvec of the bench workload itself (D={D}, d={d_HEAD}, chi={CHI}); "
              f"{warm_done} of {warm} warm-up steps run")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "a": [D, d_HEAD, D], "w": [CHI, d_HEAD, d_HEAD, CHI], "l": [D, CHI, D],
                   "r": [D, CHI, D], "flops_per_step": f_alg(D, d_HEAD, CHI),
                   "note": "reference is pure NumPy (no GPU path); timed on the host cores, all BLAS threads"},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cpu_threads(), "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sharded_parity(torch, dist, device, world):
    """Correctness of the sharded operator on THIS run's ranks (what the timed kernels compute is checked where it
    is timed): a small same-seeded problem on every rank -- sharded matvec vs the unsharded device matvec, and the
    lowest Ritz value of a 12-step Lanczos run on the sharded vs the unsharded operator.  Max over ranks."""
    import pytenet_b200 as ptb
    from pytenet_b200.sharded import PrecontractedShardedHamiltonian, ShardedEffectiveHamiltonian
    g = torch.Generator(device=device).manual_seed(5)
    Dl, d, Dr, cl, cr = 96, 2, 80, 37, 29
    a = torch.randn(Dl, d, Dr, dtype=torch.complex128, device=device, generator=g)
    l = torch.randn(Dl, cl, Dl, dtype=torch.complex128, device=device, generator=g)
    r = torch.randn(Dr, cr, Dr, dtype=torch.complex128, device=device, generator=g)
    w = torch.randn(cl, d, d, cr, dtype=torch.float64, device=device, generator=g)
    w = w * (torch.rand(cl, d, d, cr, device=device, generator=g) < 0.2)
    cls = PrecontractedShardedHamiltonian if world > 1 else ShardedEffectiveHamiltonian
    ref = ptb.apply_local_hamiltonian(a, w, l, r)
    err = (torch.linalg.norm(cls.from_full(w, l, r).matvec(a) - ref) / torch.linalg.norm(ref)).item()
    lh = l + l.conj().permute(2, 1, 0); rh = r + r.conj().permute(2, 1, 0); wh = w + w.permute(0, 2, 1, 3)
    hs = cls.from_full(wh, lh, rh)
    ev, _ = ptb.eigh_krylov(lambda x: hs.matvec(x.reshape(Dl, d, Dr)).reshape(-1), a.reshape(-1), 12, 1)
    ev0, _ = ptb.eigh_krylov(lambda x: ptb.apply_local_hamiltonian(x.reshape(Dl, d, Dr), wh, lh, rh).reshape(-1),
                             a.reshape(-1), 12, 1)
    t = torch.tensor([err, abs(ev[0] - ev0[0]) / abs(ev0[0])], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"parity_rel_err": t[0].item(), "ritz_value_rel_diff": t[1].item(), "ritz_value": float(ev[0]),
            "parity_problem": f"a ({Dl},{d},{Dr}), w ({cl},{d},{d},{cr}) 20% dense, vs unsharded device matvec",
            "parity_ok": bool(t[0].item() < 1e-12 and t[1].item() < 1e-9)}


def molecular_centre_w():
    """Centre tensor w (562,2,2,501) of the reference-built 32-orbital molecular MPO (tests/golden cache)."""
    from pytenet_b200.hamiltonian import cached_mpo_tensors
    path = os.path.join(ROOT, "tests", "golden", "molecular_mpo_N32.npz")
    _, _, ws = cached_mpo_tensors(path)
    return ws[16]


def sharded_block(torch, dist, device, world, rank, D=1024, d=2, reps=3):
    """One effective-H matvec at the centre site of the 32-orbital molecular MPO (BASELINE config 4): the MPO
    tensor is the reference's own w (562,2,2,501) (real, 16.8 % dense, from the committed cache), a (1024,2,1024),
    random environments of the right shape; MPO bond split over all ranks.  Fixed total work, so ms_per_matvec
    across N = 1, 2, 4, 8 is the strong-scaling curve.  Carries its own multi-rank parity check."""
    from pytenet_b200.sharded import ShardedEffectiveHamiltonian, PrecontractedShardedHamiltonian, _CudaOps
    parity = sharded_parity(torch, dist, device, world)
    w = molecular_centre_w()
    cl, _, _, cr = w.shape
    gen = torch.Generator(device=device).manual_seed(1 + 7919 * rank)
    scale = 1.0 / np.sqrt(D)
    setup_ms = 0.0
    if world > 1:
        # all-reduce-only variant: LW_g precontracted once per site from the full left block, two GEMMs per matvec
        P = -(-cr // world)
        q0, q1 = rank * P, min((rank + 1) * P, cr)
        l_full = torch.randn((D, cl, D), dtype=torch.complex128, device=device,
                             generator=torch.Generator(device=device).manual_seed(1)) * scale
        r_shard = torch.zeros((D, P, D), dtype=torch.complex128, device=device)
        r_shard[:, :q1 - q0, :] = torch.randn((D, q1 - q0, D), dtype=torch.complex128, device=device,
                                              generator=gen) * scale
        w3 = np.zeros((d, P, d, cl))
        w3[:, :q1 - q0] = w[:, :, :, q0:q1].transpose(2, 3, 1, 0)          # [(s, kappa_loc, s'), k]
        w3 = torch.from_numpy(w3.reshape(d * P * d, cl)).to(device)
        ops = _CudaOps()
        lw = torch.empty((D, d * P * d, D), dtype=torch.complex128, device=device)
        torch.cuda.synchronize(device)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.precontract(w3, l_full, lw)
        e1.record()
        torch.cuda.synchronize(device)
        setup_ms = e0.elapsed_time(e1)
        del l_full
        heff = PrecontractedShardedHamiltonian(lw.reshape(D * d * P, d, D), r_shard, (D, d, D, d, D, D, P), ops=ops)
    else:
        # one GPU: the three-step contraction (fewer flops) is the baseline the N > 1 runs scale against
        l = torch.randn((D, cl, D), dtype=torch.complex128, device=device, generator=gen) * scale
        r = torch.randn((D, cr, D), dtype=torch.complex128, device=device, generator=gen) * scale
        heff = ShardedEffectiveHamiltonian.from_full(torch.from_numpy(w).to(device), l, r)
        del l, r
    x = torch.randn(D, d, D, dtype=torch.complex128, device=device) / np.sqrt(D * d * D)
    for _ in range(2):
        heff.matvec(x)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        heff.matvec(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        t = torch.tensor([ms, setup_ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, setup_ms = t.tolist()
    fa = 8.0 * (D * d * D * cr * D + cl * d * d * cr * D * D + D * D * cl * d * D)
    gather, red = heff.exchange_bytes_per_rank()
    blk = {"workload": f"config-4 centre-site heff matvec: a ({D},{d},{D}), w ({cl},{d},{d},{cr}) = site 16 of the "
                       f"reference-built 32-orbital molecular MPO (real, {100.0 * np.count_nonzero(w) / w.size:.1f}% "
                       f"dense), random l ({D},{cl},{D}), r ({D},{cr},{D}); MPO bond split over {world} rank(s)",
           "scaling": "strong", "n_gpus": world, "ms_per_matvec": ms, "gflops_alg": fa / ms / 1e6,
           "flops_alg": fa, "flops_exec_per_rank": heff.flops_per_rank(), "exchange": heff.exchange,
           "allgather_bytes_per_rank": gather, "allreduce_bytes_per_rank": red,
           "algorithm": type(heff).__name__, "precontract_ms_per_site": setup_ms,
           "ms_per_matvec_incl_precontract_k25": ms + setup_ms / 25.0, **parity}
    del heff, x
    torch.cuda.empty_cache()
    return blk


def sharded_sweep_block(torch, dist, device, world, rank, D=128, k=10):
    """BASELINE config 4 through its public driver at a bench-sized bond dimension: one sweep of
    `dmrg_singlesite_sharded` on the reference-built 32-orbital molecular MPO (MPO bonds up to 562), random start
    state in the half-filling sector with bonds <= D.  Same start state for every N, so `energy` must agree across
    the N = 1, 2, 4, 8 lines of a scaling run (the all-reduce order differs: agreement to ~1e-10); at N > 1 rank 0
    also reports psi's drift across ranks.  The full-size run (D = 1024, 8 GPUs) is tools/config4_sweep.py."""
    import pytenet_b200 as ptb
    from pytenet_b200.hamiltonian import load_cached_mpo
    from pytenet_b200.sharded_dmrg import dmrg_singlesite_sharded
    # untimed warm-up on the 10-orbital sibling (library handles, plan caches, workspaces)
    h10 = load_cached_mpo(os.path.join(ROOT, "tests", "golden", "molecular_mpo_N10.npz"), device=device)
    dmrg_singlesite_sharded(h10, ptb.MPS.construct_random(10, h10.qsite, 5, max_vdim=16, dtype="complex",
                                                          rng=np.random.default_rng(3), device=device), 1,
                            numiter_lanczos=4)
    h = load_cached_mpo(os.path.join(ROOT, "tests", "golden", "molecular_mpo_N32.npz"), device=device)
    n = h.nsites
    psi = ptb.MPS.construct_random(n, h.qsite, n // 2, max_vdim=D, dtype="complex", rng=np.random.default_rng(11),
                                   device=device)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    t0 = time.perf_counter()
    en = dmrg_singlesite_sharded(h, psi, 1, numiter_lanczos=k)
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    fp = torch.stack([torch.linalg.norm(t).double() for t in psi.a])
    drift = 0.0
    stat = torch.tensor([secs], dtype=torch.float64, device=device)
    if world > 1:
        ref = fp.clone()
        dist.broadcast(ref, src=0)
        dd = (fp - ref).abs().max().reshape(1)
        dist.all_reduce(dd, op=dist.ReduceOp.MAX)
        drift = dd.item()
        dist.all_reduce(stat, op=dist.ReduceOp.MAX)
    blk = {"workload": f"dmrg_singlesite_sharded, 32-orbital molecular MPO (reference-built, bonds <= "
                       f"{max(h.bond_dims)}), psi bonds <= {D}, k = {k}, 1 sweep, {world} rank(s)",
           "n_gpus": world, "seconds_per_sweep": stat.item(), "energy": float(en[-1]),
           "psi_drift_across_ranks": drift, "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
    del h, psi
    torch.cuda.empty_cache()
    return blk


def _xxz_chain(ptb, L, D, device, seed=42):
    h = ptb.heisenberg_xxz_1d_mpo(L, 1.0, 0.8, -0.1).zero_qnumbers()
    bonds = [min(2 ** i, 2 ** (L - i), D) for i in range(L + 1)]
    psi = ptb.MPS(h.qsite, [np.zeros(b, dtype=int) for b in bonds], fill="random",
                  rng=np.random.default_rng(seed), device=device)
    return h, psi, bonds


def _cpu_local_problem(kind, D, k_cpu, k):
    """Seconds the reference's CPU path (oracle restatement: NumPy + OpenBLAS, all host threads) needs for the
    Lanczos runs of ONE bulk local problem at full bond dimension, measured with `k_cpu` iterations and scaled
    linearly to `k` (every iteration is one matvec + the same vector updates).  Returns (seconds at k, raw)."""
    import oracle
    import oracle.lanczos as ol
    use_all_host_threads()
    rng = np.random.default_rng(3)

    def herm(D, chi):
        e = (rng.normal(size=(D, chi, D)) + 1j * rng.normal(size=(D, chi, D))) / np.sqrt(2 * D)
        return e + e.conj().transpose(2, 1, 0)
    t0 = time.perf_counter()
    if kind == "tdvp1":
        a, w, _, _ = host_inputs_onesite(D, CHI, seed=5)
        l, r = herm(D, CHI), herm(D, CHI)
        w = w + w.transpose(0, 2, 1, 3)
        t0 = time.perf_counter()
        ol.expm_krylov(lambda x: oracle.apply_local_hamiltonian(x.reshape(a.shape), w, l, r).reshape(-1),
                       a.reshape(-1), -0.005j, k_cpu)
        c = np.ascontiguousarray(a[:, 0, :])
        ol.expm_krylov(lambda x: oracle.apply_local_bond_contraction(x.reshape(c.shape), l, r).reshape(-1),
                       c.reshape(-1), 0.005j, k_cpu)
        raw = time.perf_counter() - t0
        return raw * k / k_cpu, raw
    a, w, _, _ = host_inputs(D, d_HEAD, CHI, seed=5)
    l, r = herm(D, CHI), herm(D, CHI)
    w = w + w.transpose(0, 2, 1, 3)
    t0 = time.perf_counter()
    ol.eigh_krylov(lambda x: oracle.apply_local_hamiltonian(x.reshape(a.shape), w, l, r).reshape(-1),
                   a.reshape(-1), k_cpu, 1)
    raw = time.perf_counter() - t0
    t1 = time.perf_counter()
    np.linalg.svd(a.reshape(D * 2, 2 * D), full_matrices=False)
    svd = time.perf_counter() - t1
    return raw * k / k_cpu + svd, raw + svd


def sweeps_block(torch, ptb, device):
    """The "sweep seconds" half of BASELINE.json's metric, through the PUBLIC drivers, with the device time split
    by category (CUDA events around the regions of pytenet_b200/_prof.py; "host+other" is what is left of the
    synchronised wall time) and the reference's CPU path beside it.

    Phases: "lanczos" = Krylov runs of the site problems (matvecs + vector kernels + k x k solve), "lanczos_bond" =
    those of the zero-site problems, "env" = environment updates, "qr" / "svd" = factorisations, "glue" = merges
    and gauge absorption.

    * `tdvp_singlesite` (pytenet/tdvp.py:26-118): XXZ chain L = 24 whose bulk bonds reach D = 2048 (bond profile
      min(2^i, 2^(L-i), 2048): the public signature has no maximum-bond argument), complex128, k = 25, ONE full
      symmetric time step (left + right sweep).
    * `dmrg_twosite` (pytenet/dmrg.py:96-178): XXZ chain L = 22 capped at 1024 (config-2 two-site shape
      (1024,4,1024) at the centre), k = 25, tol_split = 0, ONE sweep.
    CPU side: a complete sweep is hours on the host, so the bulk local problem is timed at FULL bond
    dimension with 3 of the 25 Lanczos iterations (scaled linearly in k) and the sweep is extrapolated over the
    sites by their flop counts -- labelled "extrapolated"."""
    from pytenet_b200 import _prof
    out = {}
    k = 25

    def run(name, fn, h, psi):
        torch.cuda.synchronize()
        _prof.enable(True)
        t0 = time.perf_counter()
        fn(h, psi)
        torch.cuda.synchronize()
        secs = time.perf_counter() - t0
        ph = {kk: v / 1e3 for kk, v in _prof.report().items()}
        _prof.enable(False)
        ph["host+other"] = max(0.0, secs - sum(ph.values()))
        return secs, ph

    def site_flops(bonds, d, two_site):
        tot = []
        n = len(bonds) - 1
        for i in range(n - 1 if two_site else n):
            Dl, Dr = bonds[i], bonds[i + 2] if two_site else bonds[i + 1]
            cl = 1 if i == 0 else CHI
            cr = 1 if (i + (2 if two_site else 1)) == n else CHI
            tot.append(8.0 * (Dl * d * Dr * cr * Dr + cl * d * d * cr * Dl * Dr + Dl * Dl * cl * d * Dr))
        return tot

    # ---- single-site TDVP, D = 2048 ----
    L, D = 24, D_HEAD
    h, psi, bonds = _xxz_chain(ptb, L, D, device)
    secs, ph = run("tdvp1", lambda h_, p_: ptb.tdvp_singlesite(h_, p_, 0.01 - 0.05j, 1, numiter_lanczos=k), h, psi)
    fl = site_flops(bonds, 2, False)
    bulk = 8.0 * (2 * 2 * CHI * D ** 3 + 4 * CHI * CHI * D * D)
    # per time step every site problem runs twice (except the last) and every bond once per direction; in units of
    # the bulk site problem (site matvec + bond matvec) the chain is sum(F_site) / F_bulk
    units = 2.0 * sum(fl) / bulk
    cpu_k, cpu_raw = _cpu_local_problem("tdvp1", D, 3, k)
    out["tdvp_singlesite"] = {
        "driver": "pytenet_b200.tdvp_singlesite(H, psi, dt=0.01-0.05j, numsteps=1, numiter_lanczos=25)",
        "workload": f"XXZ L={L}, zero quantum numbers, complex128, bonds min(2^i, 2^(L-i), {D}) "
                    f"({sum(1 for i in range(L) if bonds[i] == D and bonds[i + 1] == D)} sites with both bonds at {D})",
        "seconds_per_step": secs, "device_seconds_by_phase": ph,
        "cpu": {"kind": "port", "cores": cpu_threads(), "bulk_local_problem_seconds_k25": cpu_k,
                "measured_seconds": cpu_raw,
                "sample": f"site step + bond step of one bulk site at D={D}: 3 of 25 Lanczos iterations each, "
                          f"scaled linearly in k",
                "seconds_per_step_extrapolated": cpu_k * units,
                "extrapolation": f"bulk local problem x {units:.2f} bulk-equivalents (sum of per-site flops / bulk "
                                 f"flops, both directions); QR and environment updates not included (favours the CPU)"},
    }
    out["tdvp_singlesite"]["speedup_vs_cpu_extrapolated"] = cpu_k * units / secs
    del h, psi
    torch.cuda.empty_cache()

    # ---- two-site DMRG at the config-2 shape ----
    L2, D2 = 22, 1024
    h, psi, bonds = _xxz_chain(ptb, L2, D2, device, seed=43)
    secs, ph = run("dmrg2", lambda h_, p_: ptb.dmrg_twosite(h_, p_, 1, numiter_lanczos=k, tol_split=0), h, psi)
    after = list(psi.bond_dims)
    fl = site_flops(after, 2 * 2, True)
    bulk2 = f_alg(D2, d_HEAD, CHI)
    units2 = 2.0 * sum(fl) / bulk2
    cpu_k2, cpu_raw2 = _cpu_local_problem("dmrg2", D2, 3, k)
    out["dmrg_twosite"] = {
        "driver": "pytenet_b200.dmrg_twosite(H, psi, numsweeps=1, numiter_lanczos=25, tol_split=0)",
        "workload": f"XXZ L={L2} (config-2 Hamiltonian), complex128, start bonds min(2^i, 2^(L-i), {D2}); "
                    f"tol_split=0 lets the centre bonds grow: bond dims after the sweep {after}",
        "seconds_per_sweep": secs, "device_seconds_by_phase": ph,
        "svd_share": ph.get("svd", 0.0) / secs,
        "cpu": {"kind": "port", "cores": cpu_threads(), "bulk_local_problem_seconds_k25": cpu_k2,
                "measured_seconds": cpu_raw2,
                "sample": f"one (1024,4,1024) two-site problem: 3 of 25 Lanczos iterations scaled linearly in k + "
                          f"one 2048x2048 complex SVD (LAPACK gesdd)",
                "seconds_per_sweep_extrapolated": cpu_k2 * units2,
                "extrapolation": f"bulk pair problem x {units2:.2f} bulk-equivalents (per-pair matvec flops with the "
                                 f"realised bond dimensions / flops of the (1024,4,1024) pair, both directions)"},
    }
    out["dmrg_twosite"]["speedup_vs_cpu_extrapolated"] = cpu_k2 * units2 / secs
    del h, psi
    torch.cuda.empty_cache()

    # ---- single-site TDVP with quantum numbers at D = 2048 (BASELINE config-3 physics, block-sparse path) ----
    # Fermi-Hubbard chain in the (N = L, Sz = 0) sector, random state from the reference generator's sector
    # profile (MPS.construct_random), brought to canonical form once so that every bond is grouped by sector.
    # The first call builds and caches the sector plans (host tables); the second is the steady state of a run.
    L3 = 32
    h = ptb.fermi_hubbard_1d_mpo(L3, 1.0, 4.0, 0.0)
    psi = ptb.MPS.construct_random(L3, h.qsite, ptb.encode_quantum_number_pair(L3, 0), max_vdim=D_HEAD,
                                   rng=np.random.default_rng(11))
    psi.orthonormalize(mode="left")
    fn3 = lambda h_, p_: ptb.tdvp_singlesite(h_, p_, 0.02j, 1, numiter_lanczos=k)      # noqa: E731
    first, _ = run("tdvp1_qn_first", fn3, h, psi)
    secs, ph = run("tdvp1_qn", fn3, h, psi)
    big = int(np.argmax(psi.bond_dims))
    _, cnt = np.unique(psi.qbonds[big], return_counts=True)
    out["tdvp_singlesite_quantum_numbers"] = {
        "driver": "pytenet_b200.tdvp_singlesite(H, psi, dt=0.02j, numsteps=1, numiter_lanczos=25)",
        "workload": f"Fermi-Hubbard L={L3} (t=1, U=4), sector (N={L3}, Sz=0), complex128, MPS.construct_random("
                    f"max_vdim={D_HEAD}): {sum(1 for b in psi.bond_dims if b >= D_HEAD - 64)} bonds at ~{D_HEAD}, "
                    f"{len(cnt)} sectors on the largest bond (median {float(np.median(cnt)):.0f}, max {int(cnt.max())} rows)",
        "seconds_per_step": secs, "seconds_first_step_incl_plan_construction": first,
        "device_seconds_by_phase": ph,
        "path": "sector-packed grouped GEMM for site / bond problems and environment updates, batched sector QR",
    }
    del h, psi
    torch.cuda.empty_cache()
    return out


def latency_block(torch, ptb, device):
    """The launch-latency end of the path, through the public drivers (N = 1 only):
    * BASELINE config 1 -- README XXZ chain L = 10, D <= 28, `tdvp_singlesite`, 100 steps, k = 5 (fixture state of
      tests/golden/tdvp_xxz_L10.npz, generated from the reference) -- with the oracle's CPU run of the same steps
      beside it and the agreement of the two final states;
    * BASELINE config 5 at one sample stream -- METTS for the Ising chain L = 64 (experiments/metts_ising.py:17-46):
      samples per second and the realised bond dimensions."""
    import warnings
    from oracle import sweeps as osw
    z = np.load(os.path.join(ROOT, "tests", "golden", "tdvp_xxz_L10.npz"))
    n = int(z["h/nsites"])
    hq = [z[f"h/qb{i}"] for i in range(n + 1)]
    hw = [z[f"h/w{i}"] for i in range(n)]
    pq = [z[f"psi0/qb{i}"] for i in range(n + 1)]
    pa = [z[f"psi0/a{i}"] for i in range(n)]
    dt = complex(z["dt"]); k = int(z["k"]); steps = 100
    h = ptb.MPO.from_tensors(z["h/qsite"], hq, hw)
    mk = lambda: ptb.MPS.from_tensors(z["psi0/qsite"], pq, pa)          # noqa: E731
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ptb.tdvp_singlesite(h, mk(), dt, 6, numiter_lanczos=k)          # untimed: module loading, first graph capture
        runs = []
        for _ in range(5):               # (a run right after the CPU baseline's BLAS threads can take 1.5x longer)
            psi = mk()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ptb.tdvp_singlesite(h, psi, dt, steps, numiter_lanczos=k)
            torch.cuda.synchronize(); runs.append(time.perf_counter() - t0)
        best = min(runs)
        vec_gpu = psi.to_vector()
        ch = osw.Chain([a.copy() for a in pa], z["psi0/qsite"], [q.copy() for q in pq])
        t0 = time.perf_counter()
        osw.tdvp_singlesite(hw, hq, ch, dt, steps, numiter_lanczos=k)
        cpu_s = time.perf_counter() - t0
        vec_cpu = ch.to_vector()
        out["readme_tdvp_singlesite"] = {
            "workload": f"README XXZ L={n}, bonds {psi.bond_dims}, tdvp_singlesite, {steps} steps, k={k}",
            "seconds": best, "seconds_all_runs": runs, "ms_per_step": 1e3 * best / steps,
            "path": "one CUDA graph launch per time step; one kernel (1-8 CTA cluster) per small local problem "
                    "(csrc/lanczos_small.cu)",
            "cpu": {"kind": "port", "seconds": cpu_s, "cores": cpu_threads()},
            "speedup_vs_cpu": cpu_s / best,
            "rel_diff_final_state_vs_cpu": float(np.linalg.norm(vec_gpu - vec_cpu) / np.linalg.norm(vec_cpu)),
        }
        # METTS, one stream
        hm = ptb.ising_1d_mpo(64, 1.0, 0.8, -0.375)
        rng = np.random.default_rng(1000)
        kw = dict(numsteps=10, numiter_lanczos=8, tol_split=1e-10)
        ptb.metts_energy_samples(hm, 1.0, 1, rng, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        stats = {}
        nsamp = 6
        vals = ptb.metts_energy_samples(hm, 1.0, nsamp, rng, stats=stats, **kw)
        torch.cuda.synchronize(); secs = time.perf_counter() - t0
        out["metts_ising_L64"] = {
            "workload": "Ising L=64 METTS, beta=1, two-site TDVP in imaginary time (10 steps, k=8, tol_split=1e-10)",
            "samples": nsamp, "seconds": secs, "samples_per_s": nsamp / secs,
            "energy_per_site_mean": float(np.mean(vals.real) / 64),
            "realised_max_bond_dim": int(max(stats["max_bond"])) if stats.get("max_bond") else None,
        }
        # the same workload as TWO independent sample streams on this GPU (one process each, no communication): a
        # stream is host-bound and its kernels occupy 1-8 of the SMs, so a second stream fills the gaps of the first
        if os.environ.get("PTB_BENCH_SKIP_METTS_STREAMS") != "1":
            out["metts_ising_L64_two_streams"] = metts_streams(2, device.index or 0)
    return out


def metts_streams(nstreams, gpu_index):
    """tools/metts_bench.py under torch.distributed.run with `nstreams` ranks on ONE GPU (gloo for the barrier and
    the final gather of the scalars); returns its JSON record or the reason it could not run."""
    import subprocess
    env = dict(os.environ)
    env["CUDA_VISIBLE_DEVICES"] = env.get("CUDA_VISIBLE_DEVICES", "").split(",")[gpu_index] if env.get(
        "CUDA_VISIBLE_DEVICES") else str(gpu_index)
    for key in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "GROUP_RANK", "ROLE_RANK",
                "LOCAL_WORLD_SIZE", "TORCHELASTIC_RUN_ID"):
        env.pop(key, None)
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nstreams}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "metts_bench.py"),
           "--L", "64", "--samples", "6", "--streams", str(nstreams)]
    try:
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith('{"metts"')]
        if not lines:
            return {"unavailable": (res.stderr or res.stdout)[-300:]}
        rec = json.loads(lines[-1])["metts"]
        return {"workload": "Ising L=64 METTS, beta=1 (10 two-site TDVP steps, k=8, tol_split=1e-10), "
                            f"{nstreams} sample streams on one GPU, 6 samples each",
                "streams_per_gpu": rec["streams_per_gpu"], "samples": rec["samples_per_gpu"],
                "seconds": rec["seconds"], "samples_per_s_per_gpu": rec["samples_per_s_per_gpu"],
                "energy_per_site_mean": rec["energy_per_site_mean"],
                "realised_max_bond_dim": rec["realised_max_bond_dim"]["max"]}
    except Exception as exc:                                   # the headline line must not depend on this block
        return {"unavailable": repr(exc)[:300]}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import pytenet_b200 as ptb
    from pytenet_b200 import _device as dev, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    D, d, chi = D_HEAD, d_HEAD, CHI
    steps, warm = args.steps, max(args.warmup, 3)
    # page-locked staging memory on the GPU-local NUMA node (before anything is pinned)
    # (N > 1 only: at N = 1 there is no contention, and the CPU baseline of the same process keeps every core)
    numa = dev.bind_host_to_gpu_numa_node(local) if world > 1 else None

    # inputs: pinned host buffers (for e2e) and their device-resident copies (for `value`)
    a_h, w_h, l_h, r_h = host_inputs(D, d, chi, seed=1234 + rank, pinned=True)
    a = torch.from_numpy(a_h).to(device); w = torch.from_numpy(w_h).to(device)
    l = torch.from_numpy(l_h).to(device); r = torch.from_numpy(r_h).to(device)
    out = torch.empty((D, d, D), dtype=torch.complex128, device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ptb.apply_local_hamiltonian(a, w, l, r, out=out)

    # ---- device-resident timing: `value` ----
    for _ in range(warm):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()

    # ---- end to end through the public API with HOST buffers: `e2e` ----
    e2e_steps = max(2, min(steps, 5))
    # warm-up: two results alive at once, as in the timed loop (`res` is rebound while the previous result still
    # exists), so the page-locked result pool holds both buffers before the clock starts -- a cold cudaHostAlloc of
    # 268 MB costs ~90 ms and is not part of the steady-state call
    res = ptb.apply_local_hamiltonian(a_h, w_h, l_h, r_h)
    for _ in range(2):
        res = ptb.apply_local_hamiltonian(a_h, w_h, l_h, r_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = ptb.apply_local_hamiltonian(a_h, w_h, l_h, r_h)   # H2D a,w,l,r + 3 kernels + D2H out
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    h2d = a_h.nbytes + w_h.nbytes + l_h.nbytes + r_h.nbytes
    d2h = res.nbytes
    # the same call as a drop-in NumPy caller makes it: PAGEABLE arrays (the driver stages them through its own
    # bounce buffer); reported next to the pinned number, never instead of it
    pg = [np.array(x) for x in (a_h, w_h, l_h, r_h)]
    res = ptb.apply_local_hamiltonian(*pg)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        res = ptb.apply_local_hamiltonian(*pg)
    barrier()
    e2e_pageable_s = (time.perf_counter() - t0) / 2
    del pg
    if world > 1:
        t = torch.tensor([e2e_pageable_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_pageable_s = t.item()
    F = f_alg(D, d, chi)

    # ---- MPO-bond-sharded matvec (BASELINE config 4 shape), STRONG scaling over the same N ranks ----
    sharded = sharded_sweep = None
    if os.environ.get("PTB_BENCH_SKIP_SHARDED") != "1":
        del a, l, r, out, res
        torch.cuda.empty_cache()
        sharded = sharded_block(torch, dist, device, world, rank)
        sharded_sweep = None
        if os.environ.get("PTB_BENCH_SKIP_SHARDED_SWEEP") != "1":
            sharded_sweep = sharded_sweep_block(torch, dist, device, world, rank)
        a = torch.from_numpy(a_h).to(device); l = torch.from_numpy(l_h).to(device); r = torch.from_numpy(r_h).to(device)
        out = torch.empty((D, d, D), dtype=torch.complex128, device=device)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel roofline (dominant kernel: complex128 DMMA GEMM), CUDA events ----
    def time_ms(fn, reps=3):
        fn(); torch.cuda.synchronize()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    t1 = torch.empty((D * d, chi * D), dtype=torch.complex128, device=device)
    t2 = torch.randn((D * chi, d * D), dtype=torch.float64, device=device).to(torch.complex128)
    ms1 = time_ms(lambda: dev.gemm(a.reshape(D * d, D), r.reshape(D, chi * D), out=t1))
    ms3 = time_ms(lambda: dev.gemm(l.reshape(D * chi, D), t2, trans_a=True, out=out.reshape(D, d * D)))
    f_gemm = 8.0 * D * d * D * chi * D
    # FP64 denominator measured live: cuBLAS ZGEMM (library reference) and the raw DMMA issue peak
    import ctypes
    n = 4096
    x = torch.randn(n, n, dtype=torch.complex128, device=device); y = torch.randn(n, n, dtype=torch.complex128, device=device)
    zg = 8.0 * n ** 3 / time_ms(lambda: torch.matmul(x, y)) / 1e9
    del x, y
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    pout = torch.empty(sms * 2 * 256, dtype=torch.float64, device=device)
    fl = ctypes.c_double(0)
    st = torch.cuda.current_stream().cuda_stream
    pm = time_ms(lambda: lib.ptb_probe_fp64_pipe(1, sms * 2, 4000, pout.data_ptr(), ctypes.byref(fl), st))
    dmma_peak = fl.value / pm / 1e9
    peak = max(zg, dmma_peak)
    ms_dom = max(ms1, ms3)
    achieved = f_gemm / ms_dom / 1e9
    prof_traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    if os.path.exists(tpath):
        try:
            prof_traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            prof_traffic = None
    roofline = {
        "bound": "tensor", "kernel": "gemm_ws_kernel<complex128> (TMA/mbarrier warp-specialised DMMA GEMM; step 1 a.r / step 3 l^T.t2)",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": prof_traffic,
        "peak_source": "measured live in this run: max(cuBLAS ZGEMM 4096^3 via torch.matmul, register-resident "
                       "DMMA.8x8x4 probe); MEASURED_PEAKS.json has no FP64 entry",
        "cublas_zgemm_tflops": zg, "dmma_probe_tflops": dmma_peak,
        "step1_ms": ms1, "step3_ms": ms3, "flops_per_launch": f_gemm,
        "matvec_frac_of_peak": F / (ms / 1e3) / 1e12 / peak,
    }

    # PTB_BENCH_SKIP_CPU=1 only for runs under a profiler (numbers taken there are never bench values)
    cpu = None if (os.environ.get("PTB_BENCH_SKIP_CPU") == "1" or world > 1) else cpu_baseline_block()

    # ---- other shapes of the same path (device-resident, CUDA events) ----
    del t1, t2
    torch.cuda.empty_cache()
    other = []
    for (Do, do, tag) in [(1024, 4, "config 2 two-site (1024,4,1024)"), (2048, 2, "one-site (2048,2,2048)")]:
        ao, wo, lo, ro = host_inputs(Do, do, chi, seed=7) if do == d_HEAD else host_inputs_onesite(Do, chi, seed=7)
        ad, wd, ld, rd = (torch.from_numpy(x).to(device) for x in (ao, wo, lo, ro))
        mso = time_ms(lambda: ptb.apply_local_hamiltonian(ad, wd, ld, rd), reps=5)
        other.append({"shape": tag, "ms_per_matvec": mso, "gflops": f_alg(Do, do, chi) / mso / 1e6})
        del ad, wd, ld, rd
    # the other three chain contractions at the one-site headline shape (2048, 2, 2048), chi = 5
    ao, wo, lo, ro = host_inputs_onesite(D_HEAD, chi, seed=9)
    ad, wd, ld, rd = (torch.from_numpy(x).to(device) for x in (ao, wo, lo, ro))
    cd = ad[:, 0, :].contiguous()
    f_env = f_alg(D_HEAD, 2, chi)
    f_bond = 8.0 * 2 * chi * D_HEAD ** 3
    other_ops = []
    for name, fn, fl in [("contraction_operator_step_left", lambda: ptb.contraction_operator_step_left(ad, ad, wd, ld), f_env),
                         ("contraction_operator_step_right", lambda: ptb.contraction_operator_step_right(ad, ad, wd, rd), f_env),
                         ("apply_local_bond_contraction", lambda: ptb.apply_local_bond_contraction(cd, ld, rd), f_bond)]:
        mso = time_ms(fn, reps=5)
        other_ops.append({"op": name, "shape": "a (2048,2,2048), chi 5", "ms": mso, "gflops": fl / mso / 1e6})
    del ad, wd, ld, rd, cd
    block_sparse = None
    if os.environ.get("PTB_BENCH_SKIP_SECTORS") != "1":
        block_sparse = block_sparse_block(torch, ptb, device, time_ms, peak)
    sweeps = None
    if world == 1 and os.environ.get("PTB_BENCH_SKIP_SWEEPS") != "1":
        del a, l, r, out
        torch.cuda.empty_cache()
        sweeps = sweeps_block(torch, ptb, device)
    latency = None
    if world == 1 and os.environ.get("PTB_BENCH_SKIP_LATENCY") != "1":
        latency = latency_block(torch, ptb, device)

    line = {
        "metric": METRIC, "value": world * F / (ms / 1e3) / 1e9, "unit": "GFLOP/s", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex128", "data": "synthetic",
        "config": {"workload": WORKLOAD, "a": [D, d, D], "w": [chi, d, d, chi], "l": [D, chi, D], "r": [D, chi, D],
                   "flops_per_step": F, "multi_gpu": "independent replicas, one per GPU (no collective)",
                   "l2": "inputs+intermediates (3.6 GB per step) exceed the 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": world * F / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                "call": "pytenet_b200.apply_local_hamiltonian(a, w, l, r) with pinned NumPy host buffers",
                "pageable_inputs": {"ms_per_step": e2e_pageable_s * 1e3,
                                    "value": world * F / e2e_pageable_s / 1e9,
                                    "note": "same call with ordinary (pageable) NumPy arrays"},
                "host_numa_binding": numa},
        # per matvec: step-1 GEMM, the reduction of its split-K tail tiles, the sparse W step, step-3 GEMM
        # (profiles/r02_launches_bench.md: gemm_ws_kernel, tail_reduce_kernel, wapply_csr_kernel, gemm_ws_kernel)
        "gpu_launches": 4 * steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "sharded": sharded,
        "sharded_sweep": sharded_sweep,
        "sweeps": sweeps,
        "latency_regime": latency,
        "other_shapes": other,
        "other_ops": other_ops,
        "block_sparse": block_sparse,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    # exactly ONE line on stdout (the JSON): libraries that write to fd 1 on their own (NCCL prints its version
    # banner there) are diverted to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    for ln in out.getvalue().splitlines():
        if not ln.startswith("{"):
            print(ln, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
