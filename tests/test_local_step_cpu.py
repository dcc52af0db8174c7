"""CPU: the algorithms of the one-kernel local step (pytenet_b200/csrc/lanczos_small.cu, tridiag.cuh) restated in
NumPy and checked against the oracle (test infrastructure only):

* `taylor_expm_coeff` -- the k x k problem of expm_krylov for small Krylov spaces: shift by the mean diagonal, scale
  by 2^-s, sixteen Taylor terms of the tridiagonal matrix, s squarings (tridiag.cuh: tridiag_expm_taylor) -- must
  equal the eigen-decomposition formula of pytenet/krylov.py:122-136 (oracle.eigh_tridiag);
* `sliced_local_step` -- the Lanczos run with the output's right bond index split over C "CTAs": every slice
  computes t1 = v r[:, :, slice], the W step and l^T t2 on its own columns, the scalars are sums of per-slice
  partials in slice order, the new Lanczos vector is assembled from the slices -- must equal
  oracle.lanczos_iteration / expm_krylov on apply_local_hamiltonian, for uneven slices and the zero-site problem.
"""
import numpy as np
import pytest

import oracle
import oracle.lanczos as ol

TAYLOR_MAX_K, TAYLOR_TERMS, TAYLOR_MAX_SQUARINGS, TAYLOR_MAX_REAL_NORM = 16, 16, 10, 1.5


def taylor_expm_coeff(nrm, alpha, beta, dt, thresh):
    """-> (coeff, k_eff) or None when the kernel would fall back to the QL iteration."""
    numiter = len(alpha)
    n = numiter
    for j in range(numiter - 1):
        if not (beta[j] >= thresh):
            n = j + 1
            break
    if n > TAYLOR_MAX_K:
        return None
    mu = float(np.sum(alpha[:n])) / n
    bd = alpha[:n] - mu
    bl = np.concatenate([[0.0], beta[:n - 1]])
    bu = np.concatenate([beta[:n - 1], [0.0]])
    nb = np.max(np.abs(bd) + np.abs(bl) + np.abs(bu))
    na = nb * (abs(dt.real) + abs(dt.imag))
    if not (na < 0.5 * 2 ** TAYLOR_MAX_SQUARINGS):
        return None
    if not (nb * abs(dt.real) <= TAYLOR_MAX_REAL_NORM):
        return None
    s = 0
    while na > 0.5:
        na *= 0.5
        s += 1
    sc = dt * 2.0 ** (-s)
    x = np.eye(n, dtype=complex)
    term = np.eye(n, dtype=complex)
    for m in range(1, TAYLOR_TERMS + 1):
        lo = np.vstack([np.zeros((1, n)), term[:-1]])
        hi = np.vstack([term[1:], np.zeros((1, n))])
        term = (sc / m) * (bl[:, None] * lo + bd[:, None] * term + bu[:, None] * hi)
        x = x + term
    for _ in range(s):
        x = x @ x
    coeff = np.zeros(numiter, dtype=complex)
    coeff[:n] = nrm * np.exp(dt * mu) * x[:, 0]
    return coeff, n


@pytest.mark.parametrize("k", [1, 2, 3, 5, 8, 13, 16])
@pytest.mark.parametrize("dt", [0.05j, -0.05, 0.03 - 0.7j, 2.5j, -0.1, 40j])
def test_taylor_solve_equals_eigen_decomposition(k, dt):
    rng = np.random.default_rng(17 * k + int(abs(dt) * 10))
    alpha = rng.normal(size=k) * 3 - 20.0          # a spectrum far from zero: the shift matters
    beta = np.abs(rng.normal(size=max(k - 1, 0))) + 0.1
    nrm = 1.3
    w, u = oracle.eigh_tridiag(alpha, beta)
    want = u @ (nrm * np.exp(dt * w) * u[0])
    got = taylor_expm_coeff(nrm, alpha, beta, complex(dt), 1e-13)
    if got is None:
        # only the documented exclusions send a small space to the QL path
        nb = np.max(np.abs(alpha - alpha.mean())) + 2 * np.max(beta, initial=0.0)
        assert nb * abs(complex(dt).real) > 1.0
        return
    coeff, n = got
    assert n == k
    assert np.linalg.norm(coeff - want) <= 1e-12 * max(1.0, 3 * abs(dt)) * np.linalg.norm(want)


def test_taylor_solve_refuses_strongly_non_unitary_steps():
    """Imaginary-time steps with |Re dt| |T - mu| > 1.5: the matrix exponential is dominated by one eigen-component
    that e_0 may barely see; the squarings would lose what the eigenvector formula keeps (measured: 1e-8 relative at
    dt = -3 with a 13 x 13 random tridiagonal matrix) -> QL path."""
    rng = np.random.default_rng(5)
    alpha = rng.normal(size=13) * 3 - 20.0
    beta = np.abs(rng.normal(size=12)) + 0.1
    assert taylor_expm_coeff(1.0, alpha, beta, complex(-3.0), 1e-13) is None
    assert taylor_expm_coeff(1.0, alpha, beta, complex(-0.01), 1e-13) is not None


def test_taylor_solve_breakdown_and_fallback_rules():
    alpha = np.array([0.4, -1.0, 2.0, 0.3, 0.1]); beta = np.array([0.7, 1e-15, 0.5, 0.2])
    coeff, n = taylor_expm_coeff(0.9, alpha, beta, -0.1 + 0.4j, 100 * 300 * 2.2e-16)
    w, u = oracle.eigh_tridiag(alpha[:2], beta[:1])
    assert n == 2 and np.allclose(coeff[:2], u @ (0.9 * np.exp((-0.1 + 0.4j) * w) * u[0]), atol=1e-14)
    assert np.all(coeff[2:] == 0)
    # larger than 16, or a norm that needs more than ten squarings: QL path
    assert taylor_expm_coeff(1.0, np.ones(17), np.ones(16), 0.1j, 1e-13) is None
    assert taylor_expm_coeff(1.0, np.array([0.0, 900.0]), np.array([1.0]), 2.0j, 1e-13) is None
    # NaN scalars of speculative steps beyond a breakdown are never touched; NaN inside the space falls back
    assert taylor_expm_coeff(1.0, np.array([1.0, np.nan]), np.array([0.5]), 0.1j, 1e-13) is None


def sliced_local_step(a, w, l, r, numiter, C):
    """Lanczos run of H_eff = (l, w, r) on `a` with the right bond index of the output split over C slices; returns
    (|a|, alpha, beta, V) like the kernel's scal / V outputs (all numiter steps executed)."""
    Dl, d, Dr = a.shape
    cl, cr = l.shape[1], r.shape[1]
    n = a.size
    slices = [(c * Dr // C, (c + 1) * Dr // C) for c in range(C)]
    nrm = np.sqrt(np.sum(np.abs(a.reshape(-1)) ** 2))
    v = a / nrm
    V = np.zeros((numiter, n), dtype=complex)
    V[0] = v.reshape(-1)
    alpha = np.zeros(numiter); beta = np.zeros(max(numiter - 1, 0))
    vprev = np.zeros_like(v)
    for j in range(numiter):
        y = np.zeros_like(v)
        part = []
        for (j0, j1) in slices:
            t1 = np.einsum("isj,jKp->isKp", v, r[:, :, j0:j1])                     # step 1, own columns
            if w is not None:
                t2 = np.einsum("ktsK,isKp->iktp", w, t1)                           # W step, local to the slice
            else:
                t2 = t1.transpose(0, 2, 1, 3)                                      # zero-site: (i, K, s=0, p)
            y[:, :, j0:j1] = np.einsum("ikq,iktp->qtp", l, t2)                     # step 3, own columns
            part.append(np.sum((np.conj(v[:, :, j0:j1]) * y[:, :, j0:j1]).real))   # per-slice partial of alpha
        alpha[j] = sum(part)                                                        # summed in slice order
        if j == numiter - 1:
            break
        y = y - (alpha[j] * v + (beta[j - 1] * vprev if j > 0 else 0))
        beta[j] = np.sqrt(sum(np.sum(np.abs(y[:, :, j0:j1]) ** 2) for (j0, j1) in slices))
        vprev, v = v, y / beta[j]                                                   # every slice receives all slices
        V[j + 1] = v.reshape(-1)
    return nrm, alpha, beta, V


@pytest.mark.parametrize("dims,C", [((4, 4, 4, 3, 3), 1), ((16, 2, 20, 5, 5), 8), ((12, 2, 9, 4, 5), 2),
                                    ((5, 3, 7, 2, 3), 4), ((6, 1, 11, 4, 4), 8)])
def test_sliced_local_step_equals_oracle(dims, C):
    Dl, d, Dr, cl, cr = dims
    rng = np.random.default_rng(Dl * 100 + Dr)
    crand = lambda *s: rng.normal(size=s) + 1j * rng.normal(size=s)      # noqa: E731

    def herm(D, chi):
        e = crand(D, chi, D)
        return e + e.conj().transpose(2, 1, 0)

    l, r = herm(Dl, cl), herm(Dr, cr)
    a = crand(Dl, d, Dr)
    k = 6
    if d > 1:
        w = rng.normal(size=(cl, d, d, cr)); w = w + w.transpose(0, 2, 1, 3); w[np.abs(w) < 0.6] = 0
        hfun = lambda x: oracle.apply_local_hamiltonian(x.reshape(Dl, d, Dr), w, l, r).reshape(-1)   # noqa: E731
    else:
        w = None
        assert cl == cr
        hfun = lambda x: oracle.apply_local_bond_contraction(x.reshape(Dl, Dr), l, r).reshape(-1)   # noqa: E731
    oal, obe, oV = ol.lanczos_iteration(hfun, a.reshape(-1), k)
    nrm, al, be, V = sliced_local_step(a, w, l, r, k, C)
    assert abs(nrm - np.linalg.norm(a)) < 1e-13 * nrm
    assert np.allclose(al, oal, rtol=1e-11, atol=1e-11 * np.abs(oal).max())
    assert np.allclose(be, obe, rtol=1e-10)
    assert np.linalg.norm(V - oV.T) < 1e-9 * np.linalg.norm(oV)
    # the whole step: the kernel's coefficients applied to its Lanczos vectors equal expm_krylov of the oracle
    dt = -0.03j
    coeff, n = taylor_expm_coeff(nrm, al, be, dt, 100 * a.size * 2.220446049250313e-16)
    want = ol.expm_krylov(hfun, a.reshape(-1), dt, k, hermitian=True)
    assert np.linalg.norm(coeff[:n] @ V[:n] - want) < 1e-10 * np.linalg.norm(want)
