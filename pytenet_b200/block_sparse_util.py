"""
Quantum-number (block-sparse) helpers around the hot path, device-resident
(pytenet/block_sparse_util.py).  Quantum numbers are small host-side integer
arrays; tensors are CUDA torch tensors.  Sector-wise SVD / eigh run on the device
through cuSOLVER (torch.linalg), which north_star names as "not the
optimisation target"; sector-wise QR uses one launch of the batched Householder
kernel (csrc/block_qr.cu) for all sectors that fit in shared memory and cuSOLVER
for the rest; sector order, the stable grouping of indices and the
number of bond indices each sector contributes follow the reference exactly
(block_sparse_util.py:33-37, 82-83, 151-169, 288-303).
"""
import ctypes
import os

import numpy as np
import torch

from . import _device as dev
from . import _lib

__all__ = ["qnumber_outer_sum", "common_qnumbers", "qnumber_flatten", "is_qsparse", "enforce_qsparsity",
           "block_sparse_qr", "block_sparse_eigh", "block_sparse_svd"]

# SVD driver for cuSOLVER: "gesvd" (QR iteration, closest to LAPACK) unless overridden
_SVD_DRIVER = os.environ.get("PYTENET_B200_SVD_DRIVER", "gesvd")
# batched per-sector QR kernel (csrc/block_qr.cu) for sector blocks that fit in shared memory; "0" = cuSOLVER only
_BATCHED_QR = os.environ.get("PYTENET_B200_BATCHED_QR", "1") != "0"
# batched per-sector one-sided Jacobi SVD kernel (csrc/block_svd.cu) for small sector blocks; "0" = cuSOLVER only
_RANK_EPS = 4 * np.finfo(float).eps       # sigma_min <= _RANK_EPS * max(m, n) * sigma_max: numerically rank deficient
_BATCHED_SVD = os.environ.get("PYTENET_B200_BATCHED_SVD", "1") != "0"


def qnumber_outer_sum(qnums):
    """All sums q0[i0] + q1[i1] + ... as a tensor (host integers; :13-30)."""
    if len(qnums) == 0:
        return np.array(0)
    acc = np.asarray(qnums[0])
    for q in qnums[1:]:
        acc = np.add.outer(acc, np.asarray(q))
    return acc


def common_qnumbers(qnums0, qnums1):
    """Sorted quantum numbers present in both lists (:33-37)."""
    return np.intersect1d(qnums0, qnums1)


def qnumber_flatten(qnums):
    """Quantum numbers of a fused index (:40-44)."""
    return qnumber_outer_sum(qnums).reshape(-1)


_QVEC_CACHE = {}


def _qvec(q, device):
    """Quantum numbers of one axis as an int64 device vector, cached by content: sweeps test the same bonds again
    every step, and a host->device copy from pageable memory is a synchronisation point."""
    q = np.ascontiguousarray(q, dtype=np.int64)
    key = (q.tobytes(), device.index)
    t = _QVEC_CACHE.get(key)
    if t is None:
        if len(_QVEC_CACHE) >= 4096:
            _QVEC_CACHE.clear()
        t = _QVEC_CACHE[key] = torch.as_tensor(q, device=device)
    return t


def _forbidden_mask(qnums, device):
    """Boolean device tensor: True where the quantum numbers do not sum to zero.
    Built by broadcasting the per-axis vectors on the device (no dense host mask)."""
    nd = len(qnums)
    total = None
    for ax, q in enumerate(qnums):
        shape = [1] * nd
        shape[ax] = len(q)
        t = _qvec(q, device).reshape(shape)
        total = t if total is None else total + t
    return total != 0


_CAPTURING = False      # set while a sweep step is captured into a CUDA graph (tdvp._StepGraph)


def _all_zero(qnums):
    for q in qnums:
        if (q.any() if isinstance(q, np.ndarray) else np.any(q)):
            return False
    return True


def _violations(a, qnums):
    """0-dim bool device tensor: some forbidden entry of `a` is non-zero (one fused pass, no boolean gather)."""
    return torch.any((a != 0) & _forbidden_mask(qnums, a.device))


def is_qsparse(a, qnums):
    """True iff `a` vanishes wherever the quantum numbers do not sum to zero (:47-53)."""
    if isinstance(a, torch.Tensor) and a.is_cuda:
        if _all_zero(qnums):
            return True                      # all quantum numbers zero: nothing is forbidden
        if _CAPTURING:
            return True                      # a device->host read cannot be captured; the eager steps checked it
        return not bool(_violations(a, qnums).item())
    mask = qnumber_outer_sum(qnums) != 0
    return not np.any(np.asarray(a)[mask])


def _assert_qsparse(a, qnums, message="sparsity pattern must match quantum numbers"):
    """The reference's `assert is_qsparse(...)` (:114, :254).  Inside the sweep drivers (krylov.deferred_checks) the
    test runs on the device and its result is examined when the driver returns, so a factorisation is no
    synchronisation point; elsewhere it is checked immediately."""
    if isinstance(a, torch.Tensor) and a.is_cuda:
        if _all_zero(qnums):
            return                           # all quantum numbers zero: nothing is forbidden
        from . import krylov
        if krylov.deferring() and not _CAPTURING:
            krylov.defer_flag(_violations(a, qnums), message)
            return
    assert is_qsparse(a, qnums), message


def enforce_qsparsity(a, qnums):
    """Zero the forbidden entries of `a` in place (vectorised form of :56-64)."""
    if isinstance(a, torch.Tensor) and a.is_cuda:
        a[_forbidden_mask(qnums, a.device)] = 0
    else:
        a[qnumber_outer_sum(qnums) != 0] = 0


def _sector_plan(q0, q1):
    """(sectors, rows, cols): ascending common quantum numbers and, per sector, the original
    row / column indices in stable order (mergesort of the reference, :82-83)."""
    q0 = np.asarray(q0); q1 = np.asarray(q1)
    sectors = common_qnumbers(q0, q1)
    o0 = np.argsort(q0, kind="stable")
    o1 = np.argsort(q1, kind="stable")
    rows = [o0[q0[o0] == q] for q in sectors]
    cols = [o1[q1[o1] == q] for q in sectors]
    return sectors, rows, cols


def _device_indices(index_lists, device):
    """One host->device copy for all index arrays of a sector loop; returns per-array device views."""
    lens = [len(ix) for ix in index_lists]
    if sum(lens) == 0:
        return [torch.empty(0, dtype=torch.int64, device=device) for _ in index_lists]
    flat = torch.from_numpy(np.concatenate([np.asarray(ix, dtype=np.int64) for ix in index_lists])).to(device)
    return list(torch.split(flat, lens))


def _is_identity_range(idx, n):
    return len(idx) == n and (n == 0 or (idx[0] == 0 and idx[-1] == n - 1 and np.all(np.diff(idx) == 1)))


def _block(a, ri, ci):
    """a[ri][:, ci] on the device; no copy when the sector is the whole matrix."""
    if _is_identity_range(ri, a.shape[0]) and _is_identity_range(ci, a.shape[1]):
        return a
    rt = torch.as_tensor(ri, device=a.device)
    ct = torch.as_tensor(ci, device=a.device)
    return a.index_select(0, rt).index_select(1, ct)


class _QRPlan:
    """Everything block_sparse_qr derives from the quantum numbers alone: sectors, index sets, positions on the
    intermediate bond, which sectors the batched kernel takes, and the kernel's tables on the device.  Cached per
    (q0, q1): sweeps factorise the same bonds again every time step, and a cached plan means no host->device
    transfer (a synchronisation point) on the way."""

    def __init__(self, q0, q1, shape, es, device):
        self.sectors, self.rows, self.cols = _sector_plan(q0, q1)
        rows, cols = self.rows, self.cols
        self.sizes = [min(len(ri), len(ci)) for ri, ci in zip(rows, cols)]
        self.nb = int(sum(self.sizes))
        self.starts = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.qinterm = np.zeros(self.nb, dtype=q0.dtype)
        for qn, p0, sz in zip(self.sectors, self.starts[:-1], self.sizes):
            self.qinterm[p0:p0 + sz] = qn
        limit = _lib.load().ptb_block_qr_max_block_bytes() // es if _BATCHED_QR else 0
        self.one_dense = (len(self.sectors) == 1 and _is_identity_range(rows[0], shape[0])
                          and _is_identity_range(cols[0], shape[1]))
        n = len(self.sectors)
        self.small = [i for i in range(n) if len(rows[i]) * len(cols[i]) <= limit]
        self.large = [i for i in range(n) if len(rows[i]) * len(cols[i]) > limit]
        self.tab = None
        if self.small:
            meta = np.zeros((len(self.small), 8), dtype=np.int32)
            ro = co = 0
            for j, i in enumerate(self.small):
                meta[j, :5] = (len(rows[i]), len(cols[i]), ro, co, self.starts[i])
                ro += len(rows[i]); co += len(cols[i])
            tables = np.concatenate([meta.reshape(-1)] + [np.asarray(rows[i], dtype=np.int32) for i in self.small]
                                    + [np.asarray(cols[i], dtype=np.int32) for i in self.small])
            self.tab = torch.from_numpy(tables).to(device)
            self.row_off = 4 * meta.size
            self.col_off = 4 * (meta.size + ro)
            self.max_elems = int(max(len(rows[i]) * len(cols[i]) for i in self.small))
        self.large_idx = None
        if self.large and not self.one_dense:
            self.large_idx = _device_indices([rows[i] for i in self.large] + [cols[i] for i in self.large], device)
        self._all_idx = None
        self._svd = None
        self.device = device

    def svd_tables(self, es):
        """Sectors the batched Jacobi kernel takes (block + right vectors fit in shared memory) and its tables."""
        if self._svd is None:
            limit = _lib.load().ptb_block_svd_max_block_bytes() // es if _BATCHED_SVD else 0
            rows, cols = self.rows, self.cols

            def work(i):
                m, n = len(rows[i]), len(cols[i])
                return max(m, n) * min(m, n) + min(m, n) ** 2
            n = len(self.sectors)
            small = [i for i in range(n) if work(i) <= limit]
            large = [i for i in range(n) if work(i) > limit]
            tab = None
            row_off = col_off = max_work = 0
            if small:
                meta = np.zeros((len(small), 8), dtype=np.int32)
                ro = co = 0
                for j, i in enumerate(small):
                    meta[j, :5] = (len(rows[i]), len(cols[i]), ro, co, self.starts[i])
                    ro += len(rows[i]); co += len(cols[i])
                tables = np.concatenate([meta.reshape(-1)] + [np.asarray(rows[i], dtype=np.int32) for i in small]
                                        + [np.asarray(cols[i], dtype=np.int32) for i in small])
                tab = torch.from_numpy(tables).to(self.device)
                row_off, col_off = 4 * meta.size, 4 * (meta.size + ro)
                max_work = int(max(work(i) for i in small))
            self._svd = (small, large, tab, row_off, col_off, max_work)
        return self._svd

    def all_indices(self):
        """Device index tensors of every sector (rows then columns), uploaded once (used by the per-sector SVD)."""
        if self._all_idx is None:
            self._all_idx = _device_indices(list(self.rows) + list(self.cols), self.device)
        return self._all_idx


_qr_plans = {}


def _qr_plan(q0, q1, shape, es, device):
    key = (q0.tobytes(), q1.tobytes(), q0.dtype.str, q1.dtype.str, es, device.index)
    plan = _qr_plans.get(key)
    if plan is None:
        if len(_qr_plans) >= 1024:
            _qr_plans.clear()
        plan = _QRPlan(q0, q1, shape, es, device)
        _qr_plans[key] = plan
    return plan


_dense_svd_tables = {}


def single_block_svd(a):
    """
    Thin SVD of a matrix whose quantum numbers are all zero -- ONE sector covering the whole matrix -- through the
    batched Jacobi kernel, without the sector bookkeeping of `block_sparse_svd` (no quantum-number arrays, sparsity
    assertion or plan lookup: in the launch-latency regime, e.g. METTS at bond dimension 4, that host work costs
    more than the factorisation).  Returns `(u, s_host, v, s_dev)` exactly as `block_sparse_svd` would (including
    the cuSOLVER refactorisation of a numerically rank-deficient block), or None when the matrix is too large for
    the kernel (the caller then takes the general path).
    """
    m, n = a.shape
    if m == 0 or n == 0 or not _BATCHED_SVD or a.dtype not in (dev.F64, dev.C128):
        return None
    es = a.element_size()
    key = (m, n, es, a.device.index)
    ent = _dense_svd_tables.get(key)
    if ent is None:
        k = min(m, n)
        work = max(m, n) * k + k * k
        if work > _lib.load().ptb_block_svd_max_block_bytes() // es:
            ent = False
        else:
            meta = np.zeros(8, dtype=np.int32)
            meta[:5] = (m, n, 0, 0, 0)
            tables = np.concatenate([meta, np.arange(m, dtype=np.int32), np.arange(n, dtype=np.int32)])
            ent = (torch.from_numpy(tables).to(a.device), 4 * 8, 4 * (8 + m), int(work))
        if len(_dense_svd_tables) >= 4096:
            _dense_svd_tables.clear()
        _dense_svd_tables[key] = ent
    if ent is False:
        return None
    tab, row_off, col_off, max_work = ent
    a = dev.dense(a)
    nb = min(m, n)
    u = torch.empty((m, nb), dtype=a.dtype, device=a.device)
    v = torch.empty((nb, n), dtype=a.dtype, device=a.device)
    s_dev = torch.empty(nb, dtype=dev.F64, device=a.device)
    tp = tab.data_ptr()
    st = _lib.load().ptb_block_svd(_lib.PTB_COMPLEX128 if a.dtype.is_complex else _lib.PTB_REAL64, a.data_ptr(), n, 1,
                                   tp, max_work, tp + row_off, tp + col_off, u.data_ptr(), nb, s_dev.data_ptr(),
                                   v.data_ptr(), n, dev.stream_ptr(a.device))
    if st != 0:
        _lib.check(st, "ptb_block_svd")
    s_host = s_dev.cpu().numpy()
    if s_host[0] > 0 and s_host[-1] <= _RANK_EPS * max(m, n) * s_host[0]:
        # numerically rank deficient: same treatment as block_sparse_svd (LAPACK's noise-level values and completed
        # basis, so that the retained indices equal the reference's)
        u, s_dev, v = torch.linalg.svd(a, full_matrices=False, driver=_SVD_DRIVER)
        s_host = s_dev.cpu().numpy()
    return u, s_host, v, s_dev


_QR_STREAMS = {}
_QR_MAX_STREAMS = 6


def _qr_streams(device, n):
    pool = _QR_STREAMS.setdefault(device.index, [])
    while len(pool) < min(n, _QR_MAX_STREAMS):
        pool.append(torch.cuda.Stream(device=device))
    return pool[:min(n, _QR_MAX_STREAMS)]


def block_sparse_qr(a, q0, q1):
    """
    Sector-wise reduced QR of a block-sparse matrix (`a[i, j] != 0` only if
    `q0[i] == q1[j]`) -> `(q, r, qinterm)`; `r` is upper triangular only inside
    each sector (:106-180, incl. the no-common-sector case :124-134).
    """
    assert a.ndim == 2
    q0 = np.ascontiguousarray(q0); q1 = np.ascontiguousarray(q1)
    assert len(q0) == a.shape[0] and len(q1) == a.shape[1]
    _assert_qsparse(a, [q0, -q1])
    plan = _qr_plan(q0, q1, tuple(a.shape), a.element_size(), a.device)
    if len(plan.sectors) == 0:
        assert float(torch.linalg.norm(a)) == 0
        q = torch.zeros((a.shape[0], 1), dtype=a.dtype, device=a.device)
        r = torch.zeros((1, a.shape[1]), dtype=a.dtype, device=a.device)
        q[0, 0] = 1
        return q, r, q0[:1]
    if plan.one_dense and plan.large:
        qs, rs = torch.linalg.qr(a, mode="reduced")          # dense fast path: one large sector, no gather
        return qs, rs, plan.qinterm.copy()
    a = dev.dense(a)
    nb = plan.nb
    q = torch.zeros((a.shape[0], nb), dtype=a.dtype, device=a.device)
    r = torch.zeros((nb, a.shape[1]), dtype=a.dtype, device=a.device)
    main = torch.cuda.current_stream(a.device)
    ready = main.record_event() if (len(plan.large) > 1 and not _CAPTURING) else None     # inputs and outputs exist
    if plan.small:
        # all sectors that fit in shared memory: ONE launch of the batched Householder kernel (csrc/block_qr.cu),
        # which gathers each block, factorises it with LAPACK's conventions and scatters Q and R into place
        tab = plan.tab
        st = _lib.load().ptb_block_qr(_lib.PTB_COMPLEX128 if a.dtype.is_complex else _lib.PTB_REAL64, a.data_ptr(),
                                      a.shape[1], len(plan.small), tab.data_ptr(), plan.max_elems,
                                      tab.data_ptr() + plan.row_off, tab.data_ptr() + plan.col_off, q.data_ptr(), nb,
                                      r.data_ptr(), a.shape[1], dev.stream_ptr(a.device))
        _lib.check(st, "ptb_block_qr")
    if plan.large:
        # sectors beyond the shared-memory kernel: cuSOLVER per block.  A panel factorisation of a few hundred
        # columns occupies a handful of SMs for ~0.5 ms, so the independent blocks are spread round-robin over side
        # streams (gather, geqrf, orgqr and the scatter into disjoint parts of q / r all stay on the block's stream)
        nl = len(plan.large)
        streams = _qr_streams(a.device, nl) if ready is not None else []
        for st in streams:
            st.wait_event(ready)            # the batched kernel of the small sectors overlaps too
        for j, i in enumerate(plan.large):
            rt, ct = plan.large_idx[j], plan.large_idx[nl + j]
            p0, sz = plan.starts[i], plan.sizes[i]
            with torch.cuda.stream(streams[j % len(streams)] if streams else main):
                qs, rs = torch.linalg.qr(a.index_select(0, rt).index_select(1, ct), mode="reduced")
                q[rt, p0:p0 + sz] = qs
                r[p0:p0 + sz, ct] = rs
                del qs, rs
        for st in streams:
            main.wait_stream(st)
    return q, r, plan.qinterm.copy()


def block_sparse_eigh(a, q0):
    """
    Sector-wise diagonalisation of a Hermitian block-sparse matrix (`a[i, j] != 0` only if
    `q0[i] == q0[j]`) -> `(u, evals, q)` with `a = u diag(evals) u^H`; `evals` a host float64 array
    (:183-241).  The sectors follow the iteration order of `set(q0)`, as in the reference (:197,212).
    """
    assert a.ndim == 2 and a.shape[0] == a.shape[1]
    q0 = np.asarray(q0)
    assert len(q0) == a.shape[0]
    assert is_qsparse(a, [q0, -q0])
    n = a.shape[0]
    order = np.argsort(q0, kind="stable")
    u = torch.zeros((n, n), dtype=a.dtype, device=a.device)
    ev_dev = torch.zeros(n, dtype=dev.F64, device=a.device)
    q = np.zeros(n, dtype=q0.dtype)
    pos = 0
    for qn in set(q0):
        idx = order[q0[order] == qn]
        ev, us = torch.linalg.eigh(_block(a, idx, idx))
        it = torch.as_tensor(idx, device=a.device)
        u[it, pos:pos + len(idx)] = us
        ev_dev[pos:pos + len(idx)] = ev
        q[pos:pos + len(idx)] = qn
        pos += len(idx)
    assert pos == n
    return u, ev_dev.cpu().numpy(), q


# min(m, n) from which gesvdp is used: measured 2-3x faster than gesvd from 64 up (complex128, ms: 64 3.1 vs 5.9,
# 128 6.5 vs 12.6, 256 8 vs 25, 512 18 vs 62, 1024 44 vs 163, 2048 152 vs 579; profiles/r02_svd_bench.json,
# profiles/r02_svd_mid_bench.json)
_POLAR_MIN = int(os.environ.get("PYTENET_B200_POLAR_SVD_MIN", "64"))


_POLAR_WS = {}


def _polar_workspace(lib, dt, rows, cols):
    """(device bytes, host bytes) of the polar driver for one shape, cached: the query is a cuSOLVER call."""
    key = (dt, rows, cols)
    hit = _POLAR_WS.get(key)
    if hit is None:
        nd, nh = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _lib.check(lib.ptb_svd_polar_workspace_bytes(dt, rows, cols, ctypes.byref(nd), ctypes.byref(nh)),
                   "ptb_svd_polar_workspace_bytes")
        if len(_POLAR_WS) > 4096:
            _POLAR_WS.clear()
        hit = _POLAR_WS[key] = (nd.value, nh.value)
    return hit


_POLAR_SKIP = {}
_POLAR_BACKOFF = 8


def dense_svd(a):
    """
    Thin SVD `a = u diag(s) vh` of one dense device matrix; `s` stays on the device.  Large float64 / complex128
    matrices go to cuSOLVER's polar-decomposition driver through the C ABI (`ptb_svd_polar`: QDWH + Hermitian
    eigensolve, GEMM-shaped work; 2-3x faster than `gesvd` at the 2048 x 2048 complex split of a two-site step),
    everything else -- and any matrix for which the polar driver reports a loss of accuracy -- to `gesvd`.
    """
    m, n = a.shape
    k = min(m, n)
    if k < _POLAR_MIN or a.dtype not in (dev.F64, dev.C128) or not a.is_cuda:
        return torch.linalg.svd(a, full_matrices=False, driver=_SVD_DRIVER)
    # The polar driver perturbs a numerically singular matrix (spectrum graded beyond ~1e-15 of the largest value,
    # as the two-site tensors of converged physical states are) and then reports err_sigma up to 1e-4: such
    # results are discarded below.  A shape whose last attempt was discarded goes straight to gesvd for the next
    # _POLAR_BACKOFF calls instead of paying for both factorisations again.
    skey = (m, n, a.dtype)
    if _POLAR_SKIP.get(skey, 0) > 0:
        _POLAR_SKIP[skey] -= 1
        return torch.linalg.svd(a, full_matrices=False, driver=_SVD_DRIVER)
    lib = _lib.load()
    cplx = a.dtype.is_complex
    dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    # cuSOLVER is column-major: a row-major matrix is handed over as its transpose; the driver gets the tall
    # orientation (rows >= cols), so a tall `a` goes in as its conjugate transpose
    tall = m > n
    work = dev.dense(a.mH) if tall else a.clone()           # row-major (cols_cm x rows_cm), destroyed by the driver
    rows, cols = work.shape[1], work.shape[0]               # column-major view: rows x cols, lda = rows
    ubuf = torch.empty((k, rows), dtype=a.dtype, device=a.device)       # column-major rows x k
    vbuf = torch.empty((k, cols), dtype=a.dtype, device=a.device)       # column-major cols x k
    sdev = torch.empty(k, dtype=dev.F64, device=a.device)
    info = torch.zeros(1, dtype=torch.int32, device=a.device)
    nd, nh = _polar_workspace(lib, dt, rows, cols)
    dws = dev.workspace(max(nd, 16), a.device, tag="svd")
    hws = np.empty(max(nh, 16), dtype=np.uint8)
    err = ctypes.c_double(0.0)
    st = lib.ptb_svd_polar(dt, rows, cols, work.data_ptr(), rows, sdev.data_ptr(), ubuf.data_ptr(), rows,
                           vbuf.data_ptr(), cols, dws.data_ptr(), nd, hws.ctypes.data, nh, info.data_ptr(),
                           ctypes.byref(err), dev.stream_ptr(a.device))
    _lib.check(st, "ptb_svd_polar")
    if int(info.item()) != 0 or not (err.value <= 1e-11):
        _POLAR_SKIP[skey] = _POLAR_BACKOFF
        return torch.linalg.svd(a, full_matrices=False, driver=_SVD_DRIVER)
    # work = U_cm diag(s) V_cm^H with ubuf = U_cm^T and vbuf = V_cm^T as row-major arrays, and work (column-major) is
    # a^T (wide `a`) or conj(a) (tall `a`):   wide: a = conj(V_cm) s U_cm^T;   tall: a = conj(U_cm) s V_cm^T
    if tall:
        return ubuf.mH, sdev, vbuf
    return vbuf.mH, sdev, ubuf


_SIDE_STREAMS = {}


def _side_stream(device):
    st = _SIDE_STREAMS.get(device.index)
    if st is None:
        st = _SIDE_STREAMS[device.index] = torch.cuda.Stream(device=device)
    return st


def dense_svd_batch(mats):
    """
    Thin SVDs of several independent dense device matrices (the sector blocks of one split that are too large for
    the batched Jacobi kernel) -> list of `(u, s, vh)`.  Blocks the polar driver takes (`min(m, n) >= _POLAR_MIN`)
    are factorised CONCURRENTLY by `ptb_svd_polar_batch` (worker streams with their own cuSOLVER handles: a
    mid-size factorisation is a latency-bound chain of small kernels and host synchronisations, 4-8 ms each, so
    the four to eight large blocks of a split overlap almost perfectly); the others, and any block for which the
    driver reports a loss of accuracy, go through `dense_svd` one by one.
    """
    out = [None] * len(mats)
    jobs_idx = [i for i, a in enumerate(mats)
                if min(a.shape) >= _POLAR_MIN and a.dtype in (dev.F64, dev.C128) and a.is_cuda]
    if len(jobs_idx) >= 2 and len({mats[i].dtype for i in jobs_idx}) == 1:
        lib = _lib.load()
        a0 = mats[jobs_idx[0]]
        cplx = a0.dtype.is_complex
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        device = a0.device
        jobs = (_lib.SvdJob * len(jobs_idx))()
        keep = []
        need = []
        for i in jobs_idx:
            a = mats[i]
            m, n = a.shape
            tall = m > n
            work = dev.dense(a.mH) if tall else a.clone()
            rows, cols = work.shape[1], work.shape[0]
            need.append((max(_polar_workspace(lib, dt, rows, cols)[0], 16) + 255) // 256 * 256)
            keep.append((tall, work, rows, cols))
        ws = dev.workspace(sum(need), device, tag="svd_batch")
        infos = torch.zeros(len(jobs_idx), dtype=torch.int32, device=device)
        bufs = []
        pos = 0
        for j, (i, (tall, work, rows, cols)) in enumerate(zip(jobs_idx, keep)):
            k = min(rows, cols)
            ubuf = torch.empty((k, rows), dtype=work.dtype, device=device)
            vbuf = torch.empty((k, cols), dtype=work.dtype, device=device)
            sdev = torch.empty(k, dtype=dev.F64, device=device)
            bufs.append((ubuf, vbuf, sdev))
            jb = jobs[j]
            jb.rows, jb.cols, jb.a, jb.lda = rows, cols, work.data_ptr(), rows
            jb.s, jb.u, jb.ldu, jb.v, jb.ldv = sdev.data_ptr(), ubuf.data_ptr(), rows, vbuf.data_ptr(), cols
            jb.device_ws, jb.device_bytes = ws.data_ptr() + pos, need[j]
            jb.info = infos.data_ptr() + 4 * j
            pos += need[j]
        st = lib.ptb_svd_polar_batch(dt, len(jobs_idx), jobs, 0, dev.stream_ptr(device))
        info_h = infos.cpu().numpy()
        for j, i in enumerate(jobs_idx):
            ok = st == 0 or jobs[j].status == 0
            if ok and jobs[j].status == 0 and info_h[j] == 0 and jobs[j].err_sigma <= 1e-11:
                ubuf, vbuf, sdev = bufs[j]
                out[i] = (ubuf.mH, sdev, vbuf) if keep[j][0] else (vbuf.mH, sdev, ubuf)
    for i, a in enumerate(mats):
        if out[i] is None:
            out[i] = dense_svd(a)
    return out


def block_sparse_svd(a, q0, q1, with_device_sigma=False):
    """
    Sector-wise thin SVD of a block-sparse matrix -> `(u, s, v, q)` with `s` a host
    float64 array ordered sector-ascending, sigma-descending inside a sector (:244-319).
    `with_device_sigma`: additionally return the same singular values as a device vector (the callers scale the
    factors with it without a host->device copy).
    """
    assert a.ndim == 2
    q0 = np.ascontiguousarray(q0); q1 = np.ascontiguousarray(q1)
    assert len(q0) == a.shape[0] and len(q1) == a.shape[1]
    _assert_qsparse(a, [q0, -q1])
    plan = _qr_plan(q0, q1, tuple(a.shape), a.element_size(), a.device)       # same sector structure as the QR
    if len(plan.sectors) == 0:
        assert float(torch.linalg.norm(a)) == 0
        u = torch.zeros((a.shape[0], 1), dtype=a.dtype, device=a.device)
        v = torch.zeros((1, a.shape[1]), dtype=a.dtype, device=a.device)
        if a.shape[0] > 0:
            u[0, 0] = 1
        if with_device_sigma:
            return u, np.zeros(1), v, q0[:1], torch.zeros(1, dtype=dev.F64, device=a.device)
        return u, np.zeros(1), v, q0[:1]
    nb = plan.nb
    small, large, tab, row_off, col_off, max_work = plan.svd_tables(a.element_size())
    if plan.one_dense and large:
        us, ss, vs = dense_svd(a)
        if with_device_sigma:
            return us, ss.cpu().numpy(), vs, plan.qinterm.copy(), ss
        return us, ss.cpu().numpy(), vs, plan.qinterm.copy()
    a = dev.dense(a)
    # one sector covering the whole matrix, taken by the batched kernel: every entry of u, v, s is written
    alloc = torch.empty if (plan.one_dense and small and not large) else torch.zeros
    u = alloc((a.shape[0], nb), dtype=a.dtype, device=a.device)
    v = alloc((nb, a.shape[1]), dtype=a.dtype, device=a.device)
    s_dev = alloc(nb, dtype=dev.F64, device=a.device)
    side = None
    if small:
        # all sectors whose block and right vectors fit in shared memory: ONE launch of the batched one-sided
        # Jacobi kernel (csrc/block_svd.cu), singular values descending per sector as LAPACK returns them.  When
        # larger sectors follow, the kernel runs on a side stream so that it overlaps their factorisations (both
        # are latency bound and use a fraction of the SMs).
        stream = dev.stream_ptr(a.device)
        if large and not _CAPTURING:
            side = _side_stream(a.device)
            side.wait_stream(torch.cuda.current_stream(a.device))
            stream = side.cuda_stream
        st = _lib.load().ptb_block_svd(_lib.PTB_COMPLEX128 if a.dtype.is_complex else _lib.PTB_REAL64, a.data_ptr(),
                                       a.shape[1], len(small), tab.data_ptr(), max_work, tab.data_ptr() + row_off,
                                       tab.data_ptr() + col_off, u.data_ptr(), nb, s_dev.data_ptr(), v.data_ptr(),
                                       a.shape[1], stream)
        _lib.check(st, "ptb_block_svd")
    if large:
        dix = plan.all_indices()
        nsec = len(plan.sectors)
        blocks = dense_svd_batch([a.index_select(0, dix[i]).index_select(1, dix[nsec + i]) for i in large])
        if side is not None:
            torch.cuda.current_stream(a.device).wait_stream(side)       # u, v, s_dev are shared with the kernel
            side = None
        for i, (us, ss, vs) in zip(large, blocks):
            rt, ct = dix[i], dix[nsec + i]
            p0, sz = plan.starts[i], plan.sizes[i]
            u[rt, p0:p0 + sz] = us
            v[p0:p0 + sz, ct] = vs
            s_dev[p0:p0 + sz] = ss
    s_host = s_dev.cpu().numpy()
    if small:
        # A numerically rank-deficient block leaves the Jacobi kernel with column norms that are exactly zero (a
        # structurally zero column: zero singular vector) or pure rounding noise (vectors not reliably
        # orthogonal), where LAPACK -- the reference, block_sparse_util.py:294 -- returns noise of order
        # eps * sigma_max together with a completed orthonormal basis; `retained_bond_indices` (cumsum > tol)
        # keeps such an index at tol = 0.  Those (rare) blocks are refactorised by cuSOLVER so that the retained
        # indices and the isometry of u / v equal the reference's.  An all-zero block has sigma == 0 in LAPACK
        # too and stays.
        dix = None
        for i in small:
            p0, sz = plan.starts[i], plan.sizes[i]
            blk = s_host[p0:p0 + sz]
            if sz > 0 and blk[0] > 0 and blk[-1] <= _RANK_EPS * max(len(plan.rows[i]), len(plan.cols[i])) * blk[0]:
                if dix is None:
                    dix = plan.all_indices()
                rt, ct = dix[i], dix[len(plan.sectors) + i]
                us, ss, vs = torch.linalg.svd(a.index_select(0, rt).index_select(1, ct), full_matrices=False,
                                              driver=_SVD_DRIVER)
                u[rt, p0:p0 + sz] = us
                v[p0:p0 + sz, ct] = vs
                s_dev[p0:p0 + sz] = ss
                s_host[p0:p0 + sz] = ss.cpu().numpy()
    if with_device_sigma:
        return u, s_host, v, plan.qinterm.copy(), s_dev
    return u, s_host, v, plan.qinterm.copy()
