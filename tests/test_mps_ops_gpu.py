"""GPU: MPS-level operations next to the hot path -- apply_mpo, mps_add,
MPS.compress (svd / density), MPS.from_vector -- against the fixture generated from the reference
(tests/golden/mps_ops.npz) and the oracle.  Singular / eigen vectors carry a gauge freedom, so the compressed
states are compared through gauge-invariant quantities: the full state vector (1e-9), norms and scale factors
(1e-10), bond dimensions and bond quantum numbers (bit-exact)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def load_mps(ptb, z, tag, n):
    return ptb.MPS.from_tensors(z[f"{tag}/qsite"], [z[f"{tag}/qb{i}"] for i in range(n + 1)],
                                [z[f"{tag}/a{i}"] for i in range(n)])


@pytest.fixture()
def case(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "mps_ops.npz"))
    n = int(z["h/nsites"])
    h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
    return ptb, z, n, h


def test_apply_mpo(case):
    ptb, z, n, h = case
    psi = load_mps(ptb, z, "psi", n)
    hp = ptb.apply_mpo(h, psi)
    assert all(t.is_cuda for t in hp.a)
    for i in range(n):
        assert np.array_equal(hp.qbonds[i], z[f"hpsi/qb{i}"])
        assert rel(hp.a[i].cpu().numpy(), z[f"hpsi/a{i}"]) < 1e-14
    assert rel(hp.to_vector(), z["hpsi/vec"]) < 1e-13
    # <psi| (H psi)> equals the operator average computed on the hot path
    assert abs(ptb.mps_vdot(psi, hp) - ptb.mpo_average(psi, h)) < 1e-12


def test_mps_add(case):
    ptb, z, n, h = case
    psi, chi = load_mps(ptb, z, "psi", n), load_mps(ptb, z, "chi", n)
    alpha = complex(z["add/alpha"])
    sm = ptb.mps_add(psi, chi, alpha)
    for i in range(n):
        assert np.array_equal(sm.qbonds[i], z[f"add/qb{i}"])
        # alpha * a is one complex product per entry: the device fuses multiply-adds, NumPy does not (last-bit)
        got = sm.a[i].cpu().numpy()
        assert got.shape == z[f"add/a{i}"].shape and np.allclose(got, z[f"add/a{i}"], rtol=1e-14, atol=0)
        assert np.array_equal(got == 0, z[f"add/a{i}"] == 0)               # block structure exact
    assert rel((psi - chi).to_vector(), psi.to_vector() - chi.to_vector()) < 1e-13
    assert rel((psi + chi).to_vector(), psi.to_vector() + chi.to_vector()) < 1e-13


@pytest.mark.parametrize("tag,mode,direction", [("svd_l0", "svd", "left"), ("svd_r0", "svd", "right"),
                                                ("svd_l", "svd", "left"), ("svd_r", "svd", "right"),
                                                ("den", "density", "left"), ("den0", "density", "left")])
def test_compress(case, tag, mode, direction):
    ptb, z, n, h = case
    p = load_mps(ptb, z, "hpsi", n)
    nrm, scale = p.compress(float(z[f"cmp/{tag}/tol"]), mode=mode, direction=direction)
    assert abs(nrm - float(z[f"cmp/{tag}/nrm"])) < 1e-10 * nrm
    assert abs(scale - float(z[f"cmp/{tag}/scale"])) < 1e-10
    assert p.bond_dims == list(z[f"cmp/{tag}/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(p.qbonds[i], z[f"cmp/{tag}/qb{i}"])            # sector layout bit-exact
    assert rel(p.to_vector(), z[f"cmp/{tag}/vec"]) < 1e-9
    assert abs(ptb.mps_norm(p) - 1) < 1e-12


def test_compress_rejects_bad_arguments(case):
    ptb, z, n, h = case
    p = load_mps(ptb, z, "psi", n)
    with pytest.raises(ValueError):
        p.compress(0.0, mode="qr")
    with pytest.raises(ValueError):
        p.compress(0.0, mode="svd", direction="up")


@pytest.mark.parametrize("tag", ["fv0", "fv"])
def test_from_vector(case, tag):
    ptb, z, n, h = case
    m = ptb.MPS.from_vector(3, 5, z["fv/input"], tol=float(z[f"{tag}/tol"]))
    assert m.bond_dims == list(z[f"{tag}/bond_dims"])
    assert rel(m.to_vector(), z[f"{tag}/vec"]) < 1e-12


@pytest.mark.parametrize("cplx", [True, False])
def test_batched_sector_qr_matches_lapack(cuda_lib, cplx):
    """block_sparse_qr with the batched Householder kernel (csrc/block_qr.cu): same conventions as LAPACK's
    geqr2 / ung2r, so q, r and the intermediate quantum numbers equal the oracle's (NumPy / LAPACK per sector) to
    rounding -- tall, wide, square and 1 x 1 sector blocks, a zero column, unsorted quantum numbers, and a block too
    large for shared memory that takes the cuSOLVER route in the same call."""
    import oracle.blocksparse as ob
    import pytenet_b200 as ptb
    rng = np.random.default_rng(31 + int(cplx))
    # sector sizes (rows, cols): tall, wide, square, 1x1, large (cuSOLVER route), rows-only, cols-only
    spec = {-2: (40, 7), -1: (5, 19), 0: (33, 33), 1: (1, 1), 2: (700, 300), 3: (4, 0), 5: (0, 6)}
    q0 = np.concatenate([np.full(m, s) for s, (m, n) in spec.items()])
    q1 = np.concatenate([np.full(n, s) for s, (m, n) in spec.items()])
    q0 = q0[rng.permutation(len(q0))]; q1 = q1[rng.permutation(len(q1))]
    a = rng.normal(size=(len(q0), len(q1)))
    if cplx:
        a = a + 1j * rng.normal(size=a.shape)
    a[:, np.nonzero(q1 == -2)[0][3]] = 0                       # a zero column inside a sector
    ob.enforce_qsparsity(a, [q0, -q1])
    wq, wr, wqi = ob.block_sparse_qr(a, q0, q1)
    gq, gr, gqi = ptb.block_sparse_qr(torch.from_numpy(a).cuda(), q0, q1)
    assert np.array_equal(gqi, wqi)
    gq, gr = gq.cpu().numpy(), gr.cpu().numpy()
    assert gq.shape == wq.shape and gr.shape == wr.shape
    assert rel(gq @ gr, a) < 1e-13
    assert rel(gq.conj().T @ gq, np.eye(gq.shape[1])) < 1e-13
    assert rel(gr, wr) < 1e-11 and rel(gq, wq) < 1e-11
    # diagonal of R real (LAPACK convention; the sweeps read the norm from it, mps.py:157)
    for s in np.unique(wqi):
        rows = np.nonzero(wqi == s)[0]; cols = np.nonzero(q1 == s)[0]
        d = np.diag(gr[np.ix_(rows, cols[:len(rows)])])
        assert np.all(np.abs(d.imag) < 1e-14 * (1 + np.abs(d)))


@pytest.mark.parametrize("cplx", [True, False])
def test_batched_sector_svd(cuda_lib, cplx):
    """block_sparse_svd with the batched one-sided Jacobi kernel (csrc/block_svd.cu): singular values equal the
    oracle's (NumPy / LAPACK per sector) to 1e-13 of the largest one, in the same sector-ascending /
    sigma-descending order with identical bond quantum numbers; u s v reconstructs the matrix; left / right vectors
    of non-zero singular values are orthonormal -- tall, wide, square and 1 x 1 blocks, a zero column (exactly zero
    singular value), unsorted quantum numbers, and a block too large for shared memory (cuSOLVER route)."""
    import oracle.blocksparse as ob
    import pytenet_b200 as ptb
    rng = np.random.default_rng(41 + int(cplx))
    spec = {-2: (40, 7), -1: (5, 19), 0: (33, 33), 1: (1, 1), 2: (300, 260), 3: (4, 0), 5: (0, 6), 7: (64, 50),
            9: (20, 12)}
    q0 = np.concatenate([np.full(m, s) for s, (m, n) in spec.items()])
    q1 = np.concatenate([np.full(n, s) for s, (m, n) in spec.items()])
    q0 = q0[rng.permutation(len(q0))]; q1 = q1[rng.permutation(len(q1))]
    a = rng.normal(size=(len(q0), len(q1)))
    if cplx:
        a = a + 1j * rng.normal(size=a.shape)
    a[:, np.nonzero(q1 == -2)[0][3]] = 0
    # a rank-3 sector block (20 x 12): nine singular values are rounding noise in LAPACK and here
    r9, c9 = np.nonzero(q0 == 9)[0], np.nonzero(q1 == 9)[0]
    fl = rng.normal(size=(len(r9), 3)) + (1j * rng.normal(size=(len(r9), 3)) if cplx else 0)
    fr = rng.normal(size=(3, len(c9))) + (1j * rng.normal(size=(3, len(c9))) if cplx else 0)
    a[np.ix_(r9, c9)] = fl @ fr
    # graded singular values in one sector (exercise the relative accuracy of the Jacobi sweeps)
    r7, c7 = np.nonzero(q0 == 7)[0], np.nonzero(q1 == 7)[0]
    a[np.ix_(r7, c7)] *= np.logspace(0, -9, len(c7))[None, :]
    ob.enforce_qsparsity(a, [q0, -q1])
    wu, ws, wv, wq = ob.block_sparse_svd(a, q0, q1)
    gu, gs, gv, gq = ptb.block_sparse_svd(torch.from_numpy(a).cuda(), q0, q1)
    assert np.array_equal(gq, wq) and gs.shape == ws.shape
    assert np.max(np.abs(gs - ws)) < 1e-13 * ws.max()
    keep = ws > 1e-12 * ws.max()
    assert np.max(np.abs(gs[keep] - ws[keep]) / ws[keep]) < 1e-9        # small singular values relatively accurate
    gu, gv = gu.cpu().numpy(), gv.cpu().numpy()
    assert rel((gu * gs) @ gv, a) < 1e-13
    # u and v are isometries on EVERY retained index, the null vectors of rank-deficient blocks included (the
    # reference's LAPACK completes the basis; blocks where the Jacobi kernel meets an exactly zero column norm are
    # refactorised by cuSOLVER, block_sparse_util.block_sparse_svd)
    assert rel(gu.conj().T @ gu, np.eye(gu.shape[1])) < 1e-12
    assert rel(gv @ gv.conj().T, np.eye(gv.shape[0])) < 1e-12
    for s in np.unique(wq):                                             # descending inside every sector
        ss = gs[wq == s]
        assert np.all(np.diff(ss) <= 0)
    # the truncated split used by the sweeps keeps the same indices as the oracle -- also at tol = 0, where the
    # reference keeps the rounding-noise singular values (~1e-16) of the zero column and of the rank-3 sector
    noise = ws < 1e-13 * ws.max()
    assert noise.sum() == 1 + (min(spec[9]) - 3)
    for tol in (0.0, 1e-20, 1e-10, 1e-3):
        assert np.array_equal(ptb.retained_bond_indices(gs, tol), ob.retained_bond_indices(ws, tol)), tol
    assert len(ptb.retained_bond_indices(gs, 0.0)) == len(ws)


@pytest.mark.parametrize("shape,cplx", [((1024, 1024), True), ((900, 1300), False), ((1500, 800), True),
                                        ((256, 300), True), ((400, 260), False), ((64, 64), True), ((100, 70), True),
                                        ((128, 200), False), ((97, 97), False)])
def test_dense_svd_polar_driver(cuda_lib, shape, cplx):
    """block_sparse_util.dense_svd on large blocks (cuSOLVER's polar-decomposition driver through ptb_svd_polar):
    singular values equal LAPACK's to 1e-13 of the largest -- also for a spectrum graded over ten decades --, the
    factors are isometries and reconstruct the matrix; square, wide and tall; complex128 and float64."""
    from pytenet_b200.block_sparse_util import dense_svd, _POLAR_MIN
    assert min(shape) >= _POLAR_MIN
    rng = np.random.default_rng(sum(shape) + int(cplx))
    m, n = shape
    k = min(m, n)
    a = rng.normal(size=shape) + (1j * rng.normal(size=shape) if cplx else 0)
    for graded in (False, True):
        if graded:
            # a = U diag(sigma) V^H with sigma from 1 down to 1e-10
            qa, _ = np.linalg.qr(a if m >= n else a.conj().T)
            qb, _ = np.linalg.qr(rng.normal(size=(k, k)) + (1j * rng.normal(size=(k, k)) if cplx else 0))
            sig = np.logspace(0, -10, k)
            a = (qa * sig) @ qb.conj().T
            if m < n:
                a = a.conj().T
        u, s, vh = dense_svd(torch.from_numpy(np.ascontiguousarray(a)).cuda())
        u, s, vh = u.resolve_conj().cpu().numpy(), s.cpu().numpy(), vh.resolve_conj().cpu().numpy()
        ws = np.linalg.svd(a, compute_uv=False)
        assert u.shape == (m, k) and vh.shape == (k, n)
        assert np.max(np.abs(s - ws)) < 1e-13 * ws[0], graded
        assert np.all(np.diff(s) <= 1e-15 * ws[0])
        assert rel((u * s) @ vh, a) < 1e-13
        assert rel(u.conj().T @ u, np.eye(k)) < 1e-12 and rel(vh @ vh.conj().T, np.eye(k)) < 1e-12


@pytest.mark.parametrize("cplx", [True, False])
def test_dense_svd_batch_concurrent_blocks(cuda_lib, cplx):
    """block_sparse_util.dense_svd_batch (ptb_svd_polar_batch: the independent sector blocks of one split on
    concurrent worker streams): every block equals LAPACK's singular values to 1e-13 of its largest, isometric
    factors, reconstruction -- mid-size, wide, tall, graded, and blocks below the polar threshold in the same batch."""
    from pytenet_b200.block_sparse_util import dense_svd_batch
    rng = np.random.default_rng(77 + int(cplx))
    shapes = [(70, 90), (130, 130), (200, 150), (64, 64), (30, 41), (96, 300), (257, 129), (180, 180), (75, 64),
              (64, 201), (222, 222)]
    mats = []
    for m, n in shapes:
        a = rng.normal(size=(m, n)) + (1j * rng.normal(size=(m, n)) if cplx else 0)
        if (m + n) % 3 == 0:                                 # graded spectrum over eight decades
            k = min(m, n)
            u, _, vh = np.linalg.svd(a, full_matrices=False)
            a = (u * np.logspace(0, -8, k)) @ vh
        mats.append(np.ascontiguousarray(a))
    for rep in range(2):                                     # second call reuses the worker slots
        outs = dense_svd_batch([torch.from_numpy(a).cuda() for a in mats])
        assert len(outs) == len(mats)
        for a, (u, s, vh) in zip(mats, outs):
            m, n = a.shape
            k = min(m, n)
            u, s, vh = u.resolve_conj().cpu().numpy(), s.cpu().numpy(), vh.resolve_conj().cpu().numpy()
            ws = np.linalg.svd(a, compute_uv=False)
            assert u.shape == (m, k) and vh.shape == (k, n)
            assert np.max(np.abs(s - ws)) < 1e-13 * ws[0]
            assert rel((u * s) @ vh, a) < 1e-13
            assert rel(u.conj().T @ u, np.eye(k)) < 1e-12 and rel(vh @ vh.conj().T, np.eye(k)) < 1e-12


def test_block_sparse_svd_with_many_mid_size_sectors_matches_oracle(cuda_lib):
    """A two-site split at config-3-like sector sizes (sector blocks of 100-250 rows: too large for the batched
    Jacobi kernel, factorised concurrently by the polar driver): singular values, sector order and retained
    indices equal the oracle's block_sparse_svd / retained_bond_indices."""
    import pytenet_b200 as ptb
    import oracle.blocksparse as ob
    rng = np.random.default_rng(5)
    sizes0 = [110, 240, 17, 180, 150, 64]
    sizes1 = [130, 200, 25, 190, 90, 70]
    q0 = np.concatenate([np.full(n, q) for q, n in zip([-2, -1, 0, 1, 2, 3], sizes0)])
    q1 = np.concatenate([np.full(n, q) for q, n in zip([-2, -1, 0, 1, 2, 3], sizes1)])
    p0, p1 = rng.permutation(len(q0)), rng.permutation(len(q1))
    q0, q1 = q0[p0], q1[p1]
    a = rng.normal(size=(len(q0), len(q1))) + 1j * rng.normal(size=(len(q0), len(q1)))
    ob.enforce_qsparsity(a, [q0, -q1])
    u, s, v, q = ptb.block_sparse_svd(torch.from_numpy(a).cuda(), q0, q1)
    uo, so, vo, qo = ob.block_sparse_svd(a, q0, q1)
    assert np.array_equal(q, qo)
    assert np.max(np.abs(s - so)) < 1e-13 * so.max()
    u, v = u.resolve_conj().cpu().numpy(), v.resolve_conj().cpu().numpy()
    assert rel((u * s) @ v, a) < 1e-13
    for tol in (0.0, 1e-8, 1e-2):
        assert np.array_equal(ptb.retained_bond_indices(s, tol), ob.retained_bond_indices(so, tol))
