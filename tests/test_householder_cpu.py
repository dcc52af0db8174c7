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


def householder_qr_unscaled(a):
    """The round-2 form of the kernel: column j keeps its UNSCALED sub-diagonal part x; the reflector is
    v = (1, scl_j x), its scaling scl_j and the diagonal entry beta_j live in side arrays, so that the dot products
    x^H a_c of a column step do not wait for the reflector scalars and nothing of column j is rewritten
    (csrc/block_qr.cu: apply_reflector, one CTA barrier per column)."""
    real_input = not np.iscomplexobj(a)
    a = np.array(a, dtype=complex)                # (the real kernel instance is the same code with zero imaginary parts)
    m, n = a.shape
    k = min(m, n)
    tau = np.zeros(k, dtype=complex); scl = np.zeros(k, dtype=complex); dg = np.zeros(k)

    def apply(mat, j, cols, sc, f0):
        x = mat[j + 1:, j]
        for c in cols:
            w = np.conj(sc) * np.vdot(x, mat[j + 1:, c]) + mat[j, c]       # v^H a_c
            f = f0 * w
            mat[j, c] -= f
            mat[j + 1:, c] -= (f * sc) * x

    for j in range(k):
        x = a[j + 1:, j]
        ss = float(np.vdot(x, x).real)
        alpha = a[j, j]
        beta = np.real(alpha)
        if not (ss == 0.0 and np.imag(alpha) == 0.0):
            beta = -np.copysign(np.sqrt(np.real(alpha) ** 2 + np.imag(alpha) ** 2 + ss), np.real(alpha))
            tau[j] = (beta - np.real(alpha)) / beta - 1j * np.imag(alpha) / beta
            scl[j] = 1.0 / (alpha - beta)
        dg[j] = beta
        if tau[j] != 0:
            apply(a, j, range(j + 1, n), scl[j], np.conj(tau[j]))
    r = np.triu(a[:k, :])
    r[np.arange(k), np.arange(k)] = dg
    q = a[:, :k].astype(complex)

    def to_q_form(c):
        q[c + 1:, c] = -(tau[c] * scl[c]) * q[c + 1:, c]
        q[c, c] = 1 - tau[c]
        q[:c, c] = 0

    for j in range(k - 1, -1, -1):
        if j + 1 < k:
            to_q_form(j + 1)                      # by the warp that owns column j+1, right before H_j reaches it
        if tau[j] != 0:
            apply(q, j, range(j + 1, k), scl[j], tau[j])
    if k > 0:
        to_q_form(0)
    if real_input:
        assert np.all(q.imag == 0) and np.all(r.imag == 0)
        q, r = q.real, r.real
    return q, r


@pytest.mark.parametrize("shape", [(40, 7), (5, 19), (33, 33), (1, 1), (1, 6), (9, 1), (32, 28), (56, 16)])
@pytest.mark.parametrize("cplx", [True, False])
def test_unscaled_reflector_variant_matches_numpy_qr(shape, cplx):
    rng = np.random.default_rng(shape[0] * 100 + shape[1] + int(cplx) + 7)
    a = rng.normal(size=shape)
    if cplx:
        a = a + 1j * rng.normal(size=shape)
    if shape[1] > 4:
        a[:, 3] = 0
    q, r = householder_qr_unscaled(a)
    wq, wr = np.linalg.qr(a, mode="reduced")
    assert np.allclose(q @ r, a, atol=1e-13)
    assert np.allclose(q.conj().T @ q, np.eye(q.shape[1]), atol=1e-13)
    assert np.allclose(r, wr, atol=1e-12) and np.allclose(q, wq, atol=1e-12)
    assert np.all(np.abs(np.imag(np.diag(r))) < 1e-15)
    q0, r0 = householder_qr(a)
    assert np.allclose(q, q0, atol=1e-13) and np.allclose(r, r0, atol=1e-13)
