"""
NumPy restatement of the Krylov drivers on the hot path (TEST INFRASTRUCTURE,
see oracle/__init__.py).  Only the Hermitian (Lanczos) branch is restated: both
tdvp.py:229,238 and dmrg.py:185-189 pass hermitian=True / use eigh_krylov.
"""
import warnings
import numpy as np


def lanczos_iteration(afunc, vstart, numiter):
    """Three-term Lanczos without re-orthogonalisation.

    Restates pytenet/krylov.py:12-57: normalise vstart (:26-29); for each step
    alpha_j = Re<w, v_j> (:41), w -= alpha_j v_j + beta_{j-1} v_{j-1} (:42),
    beta_j = |w| (:43), breakdown if beta_j < 100 n eps (:44-50) -> truncated
    return; closing matvec for the last alpha (:53-57).
    Returns (alpha, beta, V) with V of shape (n, k_eff).
    """
    x = np.asarray(vstart)
    n = x.shape[0]
    nrm = np.linalg.norm(x)
    assert nrm > 0
    basis = np.zeros((numiter, n), dtype=x.dtype)
    basis[0] = x / nrm
    alpha = np.zeros(numiter)
    beta = np.zeros(numiter - 1)
    threshold = 100 * n * np.finfo(float).eps
    for j in range(numiter):
        w = afunc(basis[j])
        alpha[j] = np.vdot(w, basis[j]).real
        if j == numiter - 1:
            break
        if j > 0:
            w -= alpha[j] * basis[j] + beta[j - 1] * basis[j - 1]
        else:
            w -= alpha[j] * basis[j] + 0
        beta[j] = np.linalg.norm(w)
        if beta[j] < threshold:
            warnings.warn(f"beta[{j}] ~= 0 encountered during Lanczos iteration.",
                          RuntimeWarning)
            keep = j + 1
            return alpha[:keep], beta[:keep - 1], basis[:keep].T
        basis[j + 1] = w / beta[j]
    return alpha, beta, basis.T


def eigh_tridiag(d, e):
    """Dense symmetric-tridiagonal eigenproblem.  Restates pytenet/krylov.py:142-150."""
    k = len(d)
    t = np.zeros((k, k))
    idx = np.arange(k)
    t[idx, idx] = d
    if k > 1:
        t[idx[:-1], idx[1:]] = e
        t[idx[1:], idx[:-1]] = e
    return np.linalg.eigh(t)


def eigh_krylov(afunc, vstart, numiter, numeig):
    """Ritz pairs from a Lanczos run.  Restates pytenet/krylov.py:110-119."""
    alpha, beta, basis = lanczos_iteration(afunc, vstart, numiter)
    evals, evecs = eigh_tridiag(alpha, beta)
    return evals[:numeig], basis @ evecs[:, :numeig]


def expm_krylov(afunc, vec, dt, numiter, hermitian=True):
    """exp(dt A) vec in the Krylov space.  Restates pytenet/krylov.py:122-136
    (Hermitian branch only; U[0] is the first *row* of the eigenvector matrix)."""
    if not hermitian:
        raise NotImplementedError("only the Hermitian branch is on the hot path")
    alpha, beta, basis = lanczos_iteration(afunc, vec, numiter)
    evals, evecs = eigh_tridiag(alpha, beta)
    coeff = evecs @ (np.linalg.norm(vec) * np.exp(dt * evals) * evecs[0])
    return basis @ coeff
