"""
Krylov subspace drivers of the hot path -- same names and signatures as
pytenet/krylov.py (Hermitian/Lanczos branch: tdvp.py:229,238 and
dmrg.py:185-189 are its only callers on the path).

The Lanczos vectors stay resident on the device as one (numiter, n) buffer;
inner products, the three-term update and the normalisation are fused
HBM-bound kernels whose scalars stay on the device, so a complete Lanczos run
is enqueued without host synchronisation.  Only the k alphas / betas come back
(one copy) for the k x k tridiagonal problem, which stays on the host as in the
reference (krylov.py:142-150).
"""
import warnings

import numpy as np
import torch

from . import _lib
from . import _device as dev

__all__ = ["lanczos_iteration", "arnoldi_iteration", "eigh_krylov", "expm_krylov", "eigh_tridiag",
           "deferred_checks"]


def _lanczos_steps(lib, afunc, x, numiter, V, scal, sfx, stream, scratch):
    """The recursion driven from Python, one matvec callback per iteration (any `afunc`)."""
    start = getattr(lib, "ptb_lanczos_start_" + sfx)
    ortho = getattr(lib, "ptb_lanczos_ortho_step_" + sfx)
    closing = getattr(lib, "ptb_lanczos_alpha_" + sfx)
    n = x.shape[0]
    p_nrm = scal.data_ptr()
    p_alpha = p_nrm + 8
    p_beta = p_alpha + 8 * numiter
    _lib.check(start(n, x.data_ptr(), V.data_ptr(), p_nrm, scratch, stream), "lanczos_start")
    row = n * x.element_size()
    for j in range(numiter):
        vj = V.data_ptr() + j * row
        w = afunc(V[j]).reshape(-1)
        if w.dtype != V.dtype:
            w = w.to(V.dtype)
        if not w.is_contiguous() or w.is_conj():
            w = dev.dense(w)
        if w.data_ptr() == vj:
            w = w.clone()
        assert w.shape[0] == n
        if j == numiter - 1:
            # closing matvec only contributes alpha (krylov.py:53-56)
            _lib.check(closing(n, w.data_ptr(), vj, p_alpha + 8 * j, scratch, stream), "lanczos_alpha")
            break
        vjm1 = vj - row if j > 0 else None
        bprev = p_beta + 8 * (j - 1) if j > 0 else None
        # alpha_j, w -= alpha v_j + beta_{j-1} v_{j-1}, beta_j, v_{j+1} = w / beta_j  (krylov.py:41-43,51)
        _lib.check(ortho(n, w.data_ptr(), vj, vjm1, bprev, p_alpha + 8 * j, p_beta + 8 * j,
                         vj + row, scratch, stream), "lanczos_ortho_step")


def _lanczos_device(afunc, vstart, numiter, transient=False):
    """Enqueue the Lanczos recursion on the device; returns (n, V, scal) with V the resident (numiter, n)
    Lanczos-vector buffer and scal = [|vstart|, alpha[0:k], beta[0:k-1]] device doubles.
    All `numiter` steps are enqueued without host synchronisation; the breakdown
    test of krylov.py:44-50 is applied afterwards to the betas (entries up to the
    breakdown index do not depend on later steps, so the truncated results are
    identical to the reference's early exit)."""
    lib = _lib.load()
    assert numiter >= 1
    x = vstart.reshape(-1)
    cplx = x.dtype.is_complex
    x = dev.as_dtype(x, cplx)
    n = x.shape[0]
    device = x.device
    sfx = "z" if cplx else "d"
    stream = dev.stream_ptr(device)
    scratch = dev.lanczos_scratch(device).data_ptr()
    if transient:
        # the caller consumes the Lanczos vectors before the next run on this stream (expm_krylov): a view of the
        # grow-only workspace instead of a fresh allocation per run (sizes change with every split of a two-site
        # sweep, and a caching-allocator miss is a cudaMalloc)
        es = x.element_size()
        V = dev.workspace(numiter * n * es, device, tag="krylov_V")[:numiter * n * es].view(x.dtype).reshape(numiter, n)
    else:
        V = torch.empty((numiter, n), dtype=x.dtype, device=device)
    # device scalars: [nrm, alpha[0:k], beta[0:k-1]]
    scal = torch.zeros(2 * numiter, dtype=dev.F64, device=device)
    # operators that know the fused C entry (one call enqueues the whole run: _sweep.HeffOperator /
    # BondOperator -> ptb_heff_lanczos / ptb_bond_lanczos) take it; any other callable is driven step by step
    fused = getattr(afunc, "ptb_lanczos_run", None)
    if fused is None or not fused(x, numiter, V, scal):
        _lanczos_steps(lib, afunc, x, numiter, V, scal, sfx, stream, scratch)
    return n, V, scal


def _check_scalars(host, n, numiter):
    """Host-side reading of a run's scalars: norm assertion and the breakdown rule of krylov.py:44-50.
    Returns (nrm, alpha, beta) truncated at the breakdown index."""
    nrm = host[0]
    assert nrm > 0
    alpha = host[1:1 + numiter].copy()
    beta = host[1 + numiter:2 * numiter].copy()
    thresh = 100 * n * np.finfo(float).eps
    keep = numiter
    for j in range(numiter - 1):
        if not beta[j] >= thresh:       # also catches NaN of a speculative step past the breakdown
            warnings.warn(f"beta[{j}] ~= 0 encountered during Lanczos iteration.", RuntimeWarning)
            keep = j + 1
            break
    return nrm, alpha[:keep], beta[:keep - 1]


def _lanczos_core(afunc, vstart, numiter):
    """Lanczos run + the one device->host transfer of its scalars.
    Returns (nrm, alpha, beta, Vk) with Vk the first k_eff rows of the resident Lanczos-vector buffer."""
    n, V, scal = _lanczos_device(afunc, vstart, numiter)
    nrm, alpha, beta = _check_scalars(scal.cpu().numpy(), _threshold_length(afunc, n), numiter)
    return nrm, alpha, beta, V[:len(alpha)]


def _threshold_length(afunc, n):
    """Vector length entering the breakdown threshold 100 n eps (krylov.py:44): operators that iterate in a packed
    space (sector_packed.PackedHeffOperator) name the length of the dense vector they stand for."""
    return int(getattr(afunc, "ptb_n_threshold", n))


# ---- deferred scalar checks: keeps a TDVP sweep free of device->host round trips -----------------------------
# expm_krylov solves its tridiagonal problem on the device (ptb_krylov_expm_apply), so nothing forces a
# synchronisation per local step -- except the reference's breakdown warning and norm assertion.  Inside
# `deferred_checks()` (used by the sweep drivers) the scalars of every run are copied asynchronously into a
# page-locked ring and examined when the block ends (or the ring is full): same warnings, issued later.
_DEFER_SLOTS = 2048
_DEFER_WIDTH = 128


_FLAG_SLOTS = 4096


class _Deferred:
    depth = 0
    ring = None
    meta = []
    flag_ring = None       # page-locked bools: device-side assertion results (True = violated)
    flag_meta = []         # message per used slot


def _flush_deferred():
    if not _Deferred.meta and not _Deferred.flag_meta:
        return
    torch.cuda.synchronize()
    if _Deferred.flag_meta:
        flags = _Deferred.flag_ring.numpy()
        msgs, _Deferred.flag_meta = _Deferred.flag_meta, []
        for slot, msg in enumerate(msgs):
            assert not flags[slot], msg
    if not _Deferred.meta:
        return
    host = _Deferred.ring.numpy()
    meta, _Deferred.meta = _Deferred.meta, []
    for slot, (n, numiter) in enumerate(meta):
        _check_scalars(host[slot], n, numiter)


def deferring():
    """True inside a `deferred_checks()` block (the sweep drivers)."""
    return _Deferred.depth > 0


def defer_flag(violated, message):
    """Record a device-side assertion -- `violated` is a 0-dim bool device tensor that must be False -- without
    synchronising: the flag is copied asynchronously into a page-locked ring and asserted when the enclosing
    `deferred_checks()` block ends (the reference's assertions, e.g. block_sparse_util.py:114, issued later)."""
    if _Deferred.flag_ring is None:
        _Deferred.flag_ring = torch.zeros(_FLAG_SLOTS, dtype=torch.bool, pin_memory=True)
    if len(_Deferred.flag_meta) == _FLAG_SLOTS:
        _flush_deferred()
    slot = len(_Deferred.flag_meta)
    _Deferred.flag_ring[slot:slot + 1].copy_(violated.reshape(1), non_blocking=True)
    _Deferred.flag_meta.append(message)


class deferred_checks:
    """Context manager: Lanczos breakdown warnings / norm assertions of expm_krylov calls on device tensors are
    collected and issued when the outermost block exits instead of synchronising after every call."""

    def __enter__(self):
        _Deferred.depth += 1
        return self

    def __exit__(self, *exc):
        _Deferred.depth -= 1
        if _Deferred.depth == 0:
            _flush_deferred()
        return False


def defer_checks(fn):
    """Decorator form of `deferred_checks()` for the sweep drivers."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        with deferred_checks():
            return fn(*args, **kwargs)
    return wrapped


def _defer(scal, n, numiter):
    if _Deferred.ring is None:
        _Deferred.ring = torch.empty((_DEFER_SLOTS, _DEFER_WIDTH), dtype=dev.F64, pin_memory=True)
    if len(_Deferred.meta) == _DEFER_SLOTS:
        _flush_deferred()
    slot = len(_Deferred.meta)
    _Deferred.ring[slot, :2 * numiter].copy_(scal, non_blocking=True)
    _Deferred.meta.append((n, numiter))


def _defer_slot(n, numiter):
    """Reserve the next slot of the page-locked ring for a kernel that writes its scalars there itself (mapped host
    memory: the one-kernel local step) -> the slot as a host tensor of 2 numiter doubles.  Examined like the slots
    filled by `_defer`."""
    if _Deferred.ring is None:
        _Deferred.ring = torch.empty((_DEFER_SLOTS, _DEFER_WIDTH), dtype=dev.F64, pin_memory=True)
    if len(_Deferred.meta) == _DEFER_SLOTS:
        _flush_deferred()
    slot = len(_Deferred.meta)
    _Deferred.meta.append((n, numiter))
    return _Deferred.ring[slot, :2 * numiter]


def _host_afunc(afunc):
    """Adapter for the host-buffer entry: afunc sees / returns NumPy vectors."""
    return lambda t: dev.to_device(np.asarray(afunc(dev.to_host(t))), t.device)


def lanczos_iteration(afunc, vstart, numiter):
    """
    "Matrix free" Lanczos iteration (pytenet/krylov.py:12-57).

    `afunc` maps a flat vector to a flat vector (device tensors when `vstart` is a
    CUDA tensor; the returned vector is modified in place, like the reference
    does at :42).  Returns `(alpha, beta, v)`: `alpha`, `beta` NumPy float64 and
    `v` the `n x k_eff` matrix of Lanczos vectors (a transposed view of the
    resident `(k, n)` buffer, as in the reference).  On breakdown
    (`beta[j] < 100 n eps`, :44-50) the same RuntimeWarning is issued and the
    truncated results are returned.
    """
    host_mode = dev.is_host(vstart)
    x = dev.to_device(vstart)
    _, alpha, beta, Vk = _lanczos_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
    return (alpha, beta, dev.to_host(Vk).T if host_mode else Vk.T)


def _arnoldi_core(afunc, vstart, numiter):
    """Arnoldi recursion on the device: returns (|vstart|, hess (k_eff x k_eff, NumPy), Vk (k_eff x n)).
    Projections and updates are GEMV-shaped calls of the engine (h = conj(V_j) w, w -= V_j^T h), applied
    twice per step (classical Gram-Schmidt with re-orthogonalisation, numerically equivalent to the
    reference's modified Gram-Schmidt loop krylov.py:86-88); norms by the Lanczos start kernel.  All steps
    are enqueued without host synchronisation, the breakdown test (krylov.py:90-96) is applied afterwards."""
    lib = _lib.load()
    assert numiter >= 1
    x = vstart.reshape(-1)
    cplx = x.dtype.is_complex
    x = dev.as_dtype(x, cplx)
    n = x.shape[0]
    device = x.device
    start = lib.ptb_lanczos_start_z if cplx else lib.ptb_lanczos_start_d
    scratch = dev.lanczos_scratch(device).data_ptr()
    V = torch.empty((numiter, n), dtype=x.dtype, device=device)
    hess = torch.zeros((numiter, numiter), dtype=x.dtype, device=device)
    sub = torch.zeros(numiter + 1, dtype=dev.F64, device=device)          # [|vstart|, h[1,0], h[2,1], ...]
    _lib.check(start(n, x.data_ptr(), V.data_ptr(), sub.data_ptr(), scratch, dev.stream_ptr(device)), "lanczos_start")
    for j in range(numiter):
        w = afunc(V[j]).reshape(-1)
        if w.dtype != V.dtype:
            w = w.to(V.dtype)
        w = dev.dense(w)
        if w.data_ptr() == V[j].data_ptr():
            w = w.clone()
        Vj = V[:j + 1]
        for _ in range(2):
            h = dev.gemm(w.reshape(1, n), Vj, trans_b=True, conj_b=True)     # h[0, k] = sum_i w_i conj(V[k, i])
            hess[:j + 1, j] += h.reshape(-1)
            dev.gemm_strided(cplx, 0, 0, 0, 1, n, j + 1, -h, j + 1, Vj, n, w, n, accumulate=True)
        if j < numiter - 1:
            _lib.check(start(n, w.data_ptr(), V[j + 1].data_ptr(), sub.data_ptr() + 8 * (j + 1), scratch,
                             dev.stream_ptr(device)), "lanczos_start")
    hs = hess.cpu().numpy()
    subh = sub.cpu().numpy()
    nrm = subh[0]
    assert nrm > 0
    keep = numiter
    for j in range(numiter - 1):
        if not subh[j + 1] >= 100 * n * np.finfo(float).eps:
            warnings.warn(f"H[{j+1}, {j}] ~= 0 encountered during Arnoldi iteration.", RuntimeWarning)
            keep = j + 1
            break
        hs[j + 1, j] = subh[j + 1]
    return nrm, np.ascontiguousarray(hs[:keep, :keep]), V[:keep]


def arnoldi_iteration(afunc, vstart, numiter):
    """
    "Matrix free" Arnoldi iteration (pytenet/krylov.py:60-107): returns `(hess, v)` with `hess` the
    `k x k` upper Hessenberg matrix (NumPy) and `v` the `n x k` matrix of orthonormal Arnoldi vectors.
    """
    host_mode = dev.is_host(vstart)
    x = dev.to_device(vstart)
    _, hess, Vk = _arnoldi_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
    return hess, (dev.to_host(Vk).T if host_mode else Vk.T)


def eigh_tridiag(d, e):
    """
    Eigen-decomposition of a real symmetric tridiagonal matrix on the host
    (pytenet/krylov.py:142-150: dense assembly + `numpy.linalg.eigh`; k <= 25).
    """
    d = np.asarray(d, dtype=float)
    e = np.asarray(e, dtype=float)
    k = len(d)
    t = np.zeros((k, k))
    i = np.arange(k)
    t[i, i] = d
    if k > 1:
        t[i[:-1], i[1:]] = e
        t[i[1:], i[:-1]] = e
    return np.linalg.eigh(t)


def _combine(Vk, coeff):
    """sum_j coeff[j] Vk[j] on the device; Vk is (k, n) row-major, coeff a host vector."""
    lib = _lib.load()
    assert Vk.ndim == 2 and Vk.stride(1) == 1
    k, n = Vk.shape
    coeff = np.asarray(coeff)
    assert coeff.shape == (k,)
    vc = Vk.dtype.is_complex
    cc = np.iscomplexobj(coeff)
    chost = np.ascontiguousarray(coeff, dtype=np.complex128 if cc else np.float64)
    cdev = torch.from_numpy(chost).to(Vk.device)
    out = torch.empty(n, dtype=dev.C128 if (vc or cc) else dev.F64, device=Vk.device)
    st = lib.ptb_krylov_combine(_lib.PTB_COMPLEX128 if vc else _lib.PTB_REAL64,
                                _lib.PTB_COMPLEX128 if cc else _lib.PTB_REAL64,
                                n, k, Vk.data_ptr(), Vk.stride(0), cdev.data_ptr(), out.data_ptr(),
                                dev.stream_ptr(Vk.device))
    _lib.check(st, "krylov_combine")
    return out


def eigh_krylov(afunc, vstart, numiter, numeig):
    """
    Krylov subspace approximation of eigenvalues and vectors (pytenet/krylov.py:110-119).
    Returns `(w[:numeig], u_ritz)` with `u_ritz` of shape `(n, numeig)`.
    """
    host_mode = dev.is_host(vstart)
    x = dev.to_device(vstart)
    _, alpha, beta, Vk = _lanczos_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
    w_hess, u_hess = eigh_tridiag(alpha, beta)
    cols = [_combine(Vk, u_hess[:, i]) for i in range(min(numeig, u_hess.shape[1]))]
    u_ritz = torch.stack(cols, dim=1)
    return (w_hess[0:numeig], dev.to_host(u_ritz) if host_mode else u_ritz)


def expm_krylov(afunc, vec, dt, numiter, hermitian=False):
    """
    Krylov subspace approximation of `expm(dt*A) @ vec` (pytenet/krylov.py:122-139).
    The Hermitian branch is the hot path (tdvp.py:229,238 pass hermitian=True).
    """
    host_mode = dev.is_host(vec)
    x = dev.to_device(vec)
    if not hermitian:
        # general branch (krylov.py:137-139): Arnoldi on the device, expm of the k x k Hessenberg matrix on the host
        from scipy.linalg import expm
        nrm, hess, Vk = _arnoldi_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
        out = _combine(Vk, nrm * expm(dt * hess)[:, 0])
        return dev.to_host(out) if host_mode else out
    if not host_mode and numiter <= _EXPM_DEVICE_MAX_ITER:
        return _expm_device(afunc, x, dt, numiter)
    nrm, alpha, beta, Vk = _lanczos_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
    w_hess, u_hess = eigh_tridiag(alpha, beta)
    # np.linalg.norm(vec) of the reference (:136) is the norm computed by the start kernel
    coeff = u_hess @ (nrm * np.exp(dt * w_hess) * u_hess[0])
    out = _combine(Vk, coeff)
    return dev.to_host(out) if host_mode else out


_EXPM_DEVICE_MAX_ITER = 64


def _CAPTURE_SAFE_ONLY():
    """True while a sweep step is being captured into a CUDA graph: the workspace must not be (re)allocated then."""
    from . import block_sparse_util as bsu
    return bsu._CAPTURING


def _expm_device(afunc, x, dt, numiter):
    """expm_krylov with the k x k problem solved on the device (ptb_krylov_expm_apply): no device->host
    transfer on the way; the scalar checks follow immediately, or at the end of a `deferred_checks()` block."""
    lib = _lib.load()
    fused = getattr(afunc, "ptb_expm_run", None)
    if fused is not None and numiter >= 1:
        # small local problems: Lanczos run, k x k problem and combination in ONE kernel (csrc/lanczos_small.cu)
        xf = x.reshape(-1)
        xf = dev.as_dtype(xf, xf.dtype.is_complex)
        n = xf.shape[0]
        if _Deferred.depth > 0:
            # inside a sweep driver: the kernel writes alpha / beta / |x| straight into a page-locked slot
            slot = _defer_slot(_threshold_length(afunc, n), numiter)
            nmeta = len(_Deferred.meta) - 1      # (after the reservation: a full ring is flushed inside _defer_slot)
            res = fused(xf, dt, numiter, scal=slot)
            if res is not None:
                return res[0]
            del _Deferred.meta[nmeta:]           # the problem did not qualify: give the slot back
        else:
            res = fused(xf, dt, numiter)
            if res is not None:
                out, scal = res
                _check_scalars(scal.cpu().numpy(), _threshold_length(afunc, n), numiter)
                return out
    n, V, scal = _lanczos_device(afunc, x, numiter, transient=not _CAPTURE_SAFE_ONLY())
    vc = V.dtype.is_complex
    dtc = complex(dt)
    out_cplx = vc or isinstance(dt, (complex, np.complexfloating))
    out = torch.empty(n, dtype=dev.C128 if out_cplx else dev.F64, device=V.device)
    cws = dev.workspace(lib.ptb_krylov_expm_workspace_bytes(), V.device, tag="expm")
    st = lib.ptb_krylov_expm_apply(_lib.PTB_COMPLEX128 if vc else _lib.PTB_REAL64, n, numiter, V.data_ptr(),
                                   V.stride(0), scal.data_ptr(), dtc.real, dtc.imag, int(out_cplx), cws.data_ptr(),
                                   out.data_ptr(), dev.stream_ptr(V.device))
    _lib.check(st, "krylov_expm_apply")
    if _Deferred.depth > 0:
        _defer(scal, _threshold_length(afunc, n), numiter)
    else:
        _check_scalars(scal.cpu().numpy(), _threshold_length(afunc, n), numiter)
    return out
