"""NumPy prototype of the GEMM-based refinement planned for the two-site splits the polar SVD driver cannot finish
(not part of the product path).

Situation: `cusolverDnXgesvdp` returns unitary factors U, V to working precision, but for a numerically singular
matrix it factorises A + E with |E| ~ 3e-10 |A| (section 3.10), and the only fallback today is `gesvd` (580 ms at
2048^2 against 131 ms).  Everything below is matrix products plus an elementwise kernel -- work for the DMMA engine:

    T = U^H A V                                                  two GEMMs
    sigma_i = Re t_ii
    well separated pairs (|sigma_i - sigma_j| > gap): first-order corrections (Ogita & Aishima, "Iterative refinement for
        singular value decomposition based on matrix multiplication", J. Comput. Appl. Math. 369 (2020)) with the
        orthogonality defects R = I - U^H U, S = I - V^H V (two GEMMs; ~1e-14 for the polar start, O(|F|^2) after
        a step):
            alpha_ij = t_ij + sigma_j r_ij,   beta_ij = conj(t_ji) + sigma_j s_ij
            f_ij = (alpha_ij sigma_j + beta_ij sigma_i) / (sigma_j^2 - sigma_i^2)
            g_ij = (alpha_ij sigma_i + beta_ij sigma_j) / (sigma_j^2 - sigma_i^2)
            f_ii = r_ii / 2, g_ii = s_ii / 2 (also inside clusters)
        U <- U (I + F), V <- V (I + G)                           two GEMMs
    clusters of close singular values: the diagonal block of T is diagonalised by a small SVD (the batched Jacobi
        kernel's job) and the cluster's columns of U, V are rotated; a cluster at the noise floor (the numerical null
        space) is left alone -- its block of T is O(|E|^2).
    The second pass removes the O(|F|^2) loss of orthogonality of the first (graded spectra: |F| up to 1e-5).
    Note for the device version: for a graded spectrum the values below ~1e-5 sigma_max form ONE cluster whose block of
    T has entries of order |E| -- as large as or larger than its singular values -- so it needs a real SVD; that block
    has a small norm, hence the polar driver's perturbation of it (3e-10 of ITS norm) is harmless: recurse.

`refine_svd` runs `steps` such passes; `python tools/svd_refine_prototype.py` prints the errors for the test matrices
of tests/test_svd_refine_cpu.py."""
import numpy as np


def clusters_of(sigma, gap):
    """Index ranges [i0, i1) of maximal runs of the descending `sigma` whose neighbours differ by <= gap."""
    out, i0 = [], 0
    for i in range(1, len(sigma) + 1):
        if i == len(sigma) or sigma[i - 1] - sigma[i] > gap:
            if i - i0 > 1:
                out.append((i0, i))
            i0 = i
    return out


def refine_svd(a, u, v, steps=2, rel_gap=1e-5, floor=1e-13):
    """(u, sigma, v) with a ~= u diag(sigma) v^H from approximate unitary factors `u` (m x n), `v` (n x n), m == n or
    thin m > n with range(u) containing range(a) to first order.  `rel_gap`: pairs closer than rel_gap * sigma_max are
    treated as a cluster; values below floor * sigma_max form the null-space cluster."""
    u = np.array(u, dtype=complex); v = np.array(v, dtype=complex)
    n = v.shape[0]
    for _ in range(steps):
        t = u.conj().T @ a @ v
        sig = np.real(np.diag(t)).copy()
        order = np.argsort(-sig)
        if not np.array_equal(order, np.arange(n)):
            u, v, t, sig = u[:, order], v[:, order], t[np.ix_(order, order)], sig[order]
        smax = max(sig[0], np.finfo(float).tiny)
        gap = rel_gap * smax
        # clusters: small SVD of the diagonal block (skipping the null-space cluster), rotate the columns
        cl = clusters_of(sig, gap)
        in_cluster = np.zeros((n, n), dtype=bool)
        for (i0, i1) in cl:
            in_cluster[i0:i1, i0:i1] = True
            if sig[i0] <= floor * smax:
                continue
            p, sc, qh = np.linalg.svd(t[i0:i1, i0:i1])
            u[:, i0:i1] = u[:, i0:i1] @ p
            v[:, i0:i1] = v[:, i0:i1] @ qh.conj().T
        if cl:
            t = u.conj().T @ a @ v
            sig = np.real(np.diag(t)).copy()
        # orthogonality defects of the current factors (zero to 1e-14 for the polar start, O(|F|^2) after a step)
        r = np.eye(n) - u.conj().T @ u
        sm = np.eye(n) - v.conj().T @ v
        sig = np.real(np.diag(t)) / (1.0 - 0.5 * np.real(np.diag(r) + np.diag(sm)))
        si, sj = sig[:, None], sig[None, :]
        den = sj * sj - si * si
        ok = ~in_cluster & ~np.eye(n, dtype=bool)
        den = np.where(ok, den, 1.0)
        alpha = t + sj * r
        beta = t.conj().T + sj * sm
        f = np.where(ok, (alpha * sj + beta * si) / den, 0.5 * r)     # inside clusters / on the diagonal: r / 2
        g = np.where(ok, (alpha * si + beta * sj) / den, 0.5 * sm)
        u = u + u @ f
        v = v + v @ g
    t = u.conj().T @ a @ v
    d = np.diag(t)
    # absorb the phases of the diagonal into u so that sigma is real and non-negative
    ph = np.where(np.abs(d) > 0, d / np.maximum(np.abs(d), np.finfo(float).tiny), 1.0)
    u = u * ph[None, :]
    return u, np.abs(d), v


def polar_like_start(a, eps, rng):
    """What the polar driver hands over for a singular matrix: the exact SVD of a + E, |E| = eps |a|_2."""
    e = rng.normal(size=a.shape) + 1j * rng.normal(size=a.shape)
    e *= eps * np.linalg.norm(a, 2) / np.linalg.norm(e, 2)
    u, s, vh = np.linalg.svd(a + e, full_matrices=False)
    return u, s, vh.conj().T


def test_matrices(n, rng):
    def haar(k):
        q, r = np.linalg.qr(rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k)))
        return q * (np.diag(r) / np.abs(np.diag(r)))
    out = {}
    u0, v0 = haar(n), haar(n)
    out["random"] = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    s = np.concatenate([np.linspace(1.0, 0.1, n // 2), np.zeros(n - n // 2)])
    out["rank deficient (half rank)"] = (u0 * s) @ v0.conj().T
    out["graded over 18 decades"] = (u0 * np.logspace(0, -18, n)) @ v0.conj().T
    s = np.linspace(1.0, 0.2, n); s[3] = s[2]; s[10] = s[9] = s[8]; s[n // 2:] = 1e-17
    out["degenerate pairs + null space"] = (u0 * s) @ v0.conj().T
    return out


def errors(a, u, s, v):
    ref = np.linalg.svd(a, compute_uv=False)
    n = v.shape[0]
    return {"sigma": float(np.max(np.abs(np.sort(s)[::-1] - ref)) / ref[0]),
            "reconstruction": float(np.linalg.norm(a - (u * s) @ v.conj().T, 2) / ref[0]),
            "u isometry": float(np.linalg.norm(u.conj().T @ u - np.eye(n), 2)),
            "v isometry": float(np.linalg.norm(v.conj().T @ v - np.eye(n), 2))}


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for name, a in test_matrices(96, rng).items():
        u, s, v = polar_like_start(a, 3e-10, rng)
        before = errors(a, u, s, v)
        u, s, v = refine_svd(a, u, v)
        after = errors(a, u, s, v)
        print(f"{name}: " + ", ".join(f"{k} {before[k]:.1e} -> {after[k]:.1e}" for k in before))
