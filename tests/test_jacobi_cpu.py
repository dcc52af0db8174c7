"""CPU: the algorithm of the batched Jacobi SVD kernel (pytenet_b200/csrc/block_svd.cu) restated in NumPy --
the round-robin pairing formula the kernel uses must visit every column pair exactly once per sweep with
disjoint pairs inside a round (rounds are executed by independent warps), and the complex plane rotation with the
kernel's stopping rule must converge to the LAPACK singular values on hard cases (graded, clustered, rank
deficient, wide)."""
import numpy as np
import pytest


def rounds(k):
    """Pairs per round exactly as sector_svd_kernel enumerates them."""
    kk = (k + 1) & ~1
    out = []
    for r in range(kk - 1):
        pairs = []
        for pi in range(kk // 2):
            if pi == 0:
                p, q = r, kk - 1
            else:
                p, q = (r + pi) % (kk - 1), (r - pi + kk - 1) % (kk - 1)
            if p > q:
                p, q = q, p
            if q >= k:
                continue
            pairs.append((p, q))
        out.append(pairs)
    return out


@pytest.mark.parametrize("k", list(range(2, 20)) + [31, 32, 33, 64, 67])
def test_round_robin_covers_every_pair_once(k):
    seen = set()
    for pairs in rounds(k):
        cols = [c for pq in pairs for c in pq]
        assert len(cols) == len(set(cols))                 # disjoint inside a round: no two warps share a column
        for pq in pairs:
            assert pq not in seen and pq[0] < pq[1]
            seen.add(pq)
    assert len(seen) == k * (k - 1) // 2


def jacobi_svd(a, max_sweeps=60):
    """One-sided Jacobi as in the kernel (tall orientation; wide inputs through the conjugate transpose)."""
    m, n = a.shape
    tall = m >= n
    g = np.array(a if tall else a.conj().T, dtype=complex)
    rows, k = g.shape
    v = np.eye(k, dtype=complex)
    eps = np.finfo(float).eps * np.sqrt(rows)
    sweeps = 0
    for sweeps in range(1, max_sweeps + 1):
        rotated = False
        for pairs in rounds(k):
            for p, q in pairs:
                al = np.vdot(g[:, p], g[:, p]).real; be = np.vdot(g[:, q], g[:, q]).real
                c = np.vdot(g[:, p], g[:, q])
                absc = abs(c)
                if absc > eps * np.sqrt(al * be) and absc > 0:
                    ph = c / absc
                    zeta = (be - al) / (2 * absc)
                    t = np.copysign(1.0, zeta) / (abs(zeta) + np.sqrt(1 + zeta * zeta))
                    cs = 1 / np.sqrt(1 + t * t); sn = cs * t
                    for mat in (g, v):
                        xp, xq = mat[:, p].copy(), mat[:, q].copy()
                        mat[:, p] = cs * xp - sn * np.conj(ph) * xq
                        mat[:, q] = sn * ph * xp + cs * xq
                    rotated = True
        if not rotated:
            break
    sig = np.linalg.norm(g, axis=0)
    order = np.argsort(-sig, kind="stable")
    sig = sig[order]
    with np.errstate(divide="ignore", invalid="ignore"):
        un = np.where(sig > 0, g[:, order] / sig, 0)
    vv = v[:, order]
    if tall:
        return un, sig, vv.conj().T, sweeps
    return vv, sig, un.conj().T, sweeps


@pytest.mark.parametrize("case", ["random", "graded", "clustered", "rank_deficient", "wide", "real"])
def test_jacobi_converges_to_lapack_singular_values(case):
    rng = np.random.default_rng(sum(ord(ch) for ch in case))
    m, n = (48, 30)
    a = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
    if case == "graded":
        a = a * np.logspace(0, -12, n)[None, :]
    elif case == "clustered":
        u, _ = np.linalg.qr(a); w, _ = np.linalg.qr(rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)))
        a = (u * np.array([1.0] * 10 + [0.5] * 10 + [1e-3] * 10)) @ w.conj().T
    elif case == "rank_deficient":
        a[:, 7] = 0; a[:, 11] = a[:, 3]
    elif case == "wide":
        a = a.conj().T.copy()
    elif case == "real":
        a = a.real.copy()
    u, s, vh, sweeps = jacobi_svd(a)
    ref = np.linalg.svd(a, compute_uv=False)
    assert sweeps < 30
    assert np.max(np.abs(s - ref)) < 1e-13 * ref[0]
    big = ref > 1e-10 * ref[0]
    assert np.max(np.abs(s[big] - ref[big]) / ref[big]) < 1e-10
    assert np.linalg.norm((u * s) @ vh - a) < 1e-13 * np.linalg.norm(a)
    nz = s > 1e-13 * s[0]
    assert np.linalg.norm(u[:, nz].conj().T @ u[:, nz] - np.eye(nz.sum())) < 1e-10
    assert np.linalg.norm(vh[nz] @ vh[nz].conj().T - np.eye(nz.sum())) < 1e-12
