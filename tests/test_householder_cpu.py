"""CPU: the algorithm of the batched sector QR kernel (pytenet_b200/csrc/block_qr.cu) restated in NumPy -- unblocked
Householder QR with LAPACK's zgeqr2 / zlarfg / zung2r formulas -- must reproduce numpy.linalg.qr (same reflector
signs, real diagonal of R) on tall, wide, square, 1 x 1 blocks and blocks with zero columns."""
import numpy as np
import pytest


def householder_qr(a):
    a = np.array(a, dtype=complex if np.iscomplexobj(a) else float)
    m, n = a.shape
    k = min(m, n)
    tau = np.zeros(k, dtype=a.dtype)
    for j in range(k):
        x = a[j + 1:, j]
        ss = float(np.vdot(x, x).real)
        alpha = a[j, j]
        if not (ss == 0.0 and np.imag(alpha) == 0.0):
            beta = -np.copysign(np.sqrt(np.real(alpha) ** 2 + np.imag(alpha) ** 2 + ss), np.real(alpha))
            t = (beta - np.real(alpha)) / beta - 1j * np.imag(alpha) / beta
            tau[j] = t if np.iscomplexobj(a) else np.real(t)
            a[j + 1:, j] = x / (alpha - beta)
            a[j, j] = beta
            v = np.concatenate([[1.0], a[j + 1:, j]])
            w = v.conj() @ a[j:, j + 1:]                        # apply H^H = I - conj(tau) v v^H
            a[j:, j + 1:] -= np.conj(tau[j]) * np.outer(v, w)
    r = np.triu(a[:k, :])
    q = a[:, :k].copy()
    for j in range(k - 1, -1, -1):                              # zung2r, in place
        v = np.concatenate([[1.0], q[j + 1:, j]])
        if tau[j] != 0:
            w = v.conj() @ q[j:, j + 1:]
            q[j:, j + 1:] -= tau[j] * np.outer(v, w)
        q[j + 1:, j] = -tau[j] * q[j + 1:, j]
        q[j, j] = 1 - tau[j]
        q[:j, j] = 0
    return q, r


@pytest.mark.parametrize("shape", [(40, 7), (5, 19), (33, 33), (1, 1), (1, 6), (9, 1)])
@pytest.mark.parametrize("cplx", [True, False])
def test_householder_matches_numpy_qr(shape, cplx):
    rng = np.random.default_rng(shape[0] * 100 + shape[1] + int(cplx))
    a = rng.normal(size=shape)
    if cplx:
        a = a + 1j * rng.normal(size=shape)
    if shape[1] > 4:
        a[:, 3] = 0                                             # zero column: tau = 0 at that step once reached
    q, r = householder_qr(a)
    wq, wr = np.linalg.qr(a, mode="reduced")
    assert np.allclose(q @ r, a, atol=1e-13)
    assert np.allclose(q.conj().T @ q, np.eye(q.shape[1]), atol=1e-13)
    assert np.allclose(r, wr, atol=1e-12) and np.allclose(q, wq, atol=1e-12)
    assert np.all(np.abs(np.imag(np.diag(r))) < 1e-15)
