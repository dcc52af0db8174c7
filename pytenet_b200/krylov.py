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

__all__ = ["lanczos_iteration", "eigh_krylov", "expm_krylov", "eigh_tridiag"]


def _lanczos_core(afunc, vstart, numiter):
    """Run the Lanczos recursion on the device.

    Returns (nrm, alpha, beta, Vk): |vstart|, NumPy alpha (k_eff,), beta (k_eff-1,)
    and the first k_eff rows of the resident (numiter, n) Lanczos-vector buffer.
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
    start = getattr(lib, "ptb_lanczos_start_" + sfx)
    ortho = getattr(lib, "ptb_lanczos_ortho_step_" + sfx)
    closing = getattr(lib, "ptb_lanczos_alpha_" + sfx)
    stream = dev.stream_ptr(device)
    scratch = dev.lanczos_scratch(device).data_ptr()
    V = torch.empty((numiter, n), dtype=x.dtype, device=device)
    # device scalars: [nrm, alpha[0:k], beta[0:k-1]]
    scal = torch.zeros(2 * numiter, dtype=dev.F64, device=device)
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
    host = scal.cpu().numpy()          # the only device->host transfer of the run
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
    return nrm, alpha[:keep], beta[:keep - 1], V[:keep]


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
    Only the Hermitian branch is on the hot path (tdvp.py:229,238 pass hermitian=True).
    """
    if not hermitian:
        raise NotImplementedError(
            "the Arnoldi branch (krylov.py:137-139) is outside the effective-Hamiltonian path")
    host_mode = dev.is_host(vec)
    x = dev.to_device(vec)
    nrm, alpha, beta, Vk = _lanczos_core(_host_afunc(afunc) if host_mode else afunc, x, numiter)
    w_hess, u_hess = eigh_tridiag(alpha, beta)
    # np.linalg.norm(vec) of the reference (:136) is the norm computed by the start kernel
    coeff = u_hess @ (nrm * np.exp(dt * w_hess) * u_hess[0])
    out = _combine(Vk, coeff)
    return dev.to_host(out) if host_mode else out
