"""GPU: the sector-packed block-sparse path (csrc/gemm_grouped.cuh, csrc/sector_packed.cu,
pytenet_b200/sector_packed.py) -- kernels against NumPy, the packed matvec against the ORACLE on block-sparse inputs
(1e-12, north_star), BASELINE config 3 at reduced D with `MPS.construct_random`'s own sector profile against the
oracle, and at full size (a (2048,16,2048)) against the dense device matvec."""
import numpy as np
import pytest
import torch

import oracle
import oracle.blocksparse as ob

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(x - y) / (n if n > 0 else 1.0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_grouped_gemm_kernel_against_numpy(cuda_lib):
    """ptb_gemm_grouped on a random tile list: ragged m <= 128, n <= 64, contraction lengths 1 ... 300 (not multiples of
    4 or 16), arbitrary leading dimensions and offsets, untouched elements of C outside the tiles."""
    from pytenet_b200.sector_packed import _tile_table
    rng = np.random.default_rng(12)
    nA, nB, nC = 400000, 300000, 200000
    A = rng.normal(size=nA) + 1j * rng.normal(size=nA)
    B = rng.normal(size=nB) + 1j * rng.normal(size=nB)
    C0 = rng.normal(size=nC) + 1j * rng.normal(size=nC)
    tiles, want = [], C0.copy()
    c_pos = 0
    for (m, n, k) in [(128, 64, 16), (128, 64, 300), (1, 1, 1), (5, 3, 2), (37, 64, 50), (128, 9, 7), (33, 17, 129),
                      (128, 64, 4), (64, 64, 33), (100, 50, 1), (2, 64, 18), (128, 1, 5), (77, 31, 255)] * 3:
        lda = m + int(rng.integers(0, 9)); ldb = n + int(rng.integers(0, 9)); ldc = n + int(rng.integers(0, 5))
        a_off = int(rng.integers(0, nA - k * lda - m)); b_off = int(rng.integers(0, nB - k * ldb - n))
        c_off = c_pos
        c_pos += m * ldc + 7
        assert c_pos < nC
        tiles.append((a_off, b_off, c_off, lda, ldb, ldc, m, n, k))
        a = A[a_off + np.arange(k)[:, None] * lda + np.arange(m)[None, :]]
        b = B[b_off + np.arange(k)[:, None] * ldb + np.arange(n)[None, :]]
        want[c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]] = a.T @ b
    tab = _tile_table(tiles)
    dA, dB, dC = cu(A), cu(B), cu(C0)
    dT = torch.from_numpy(tab.view(np.uint8).reshape(-1)).cuda()
    st = cuda_lib.ptb_gemm_grouped(1, dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dT.data_ptr(), len(tab),
                                   torch.cuda.current_stream().cuda_stream)
    assert st == 0
    got = dC.cpu().numpy()
    assert rel(got, want) < 1e-13
    untouched = np.ones(nC, dtype=bool)
    for (a_off, b_off, c_off, lda, ldb, ldc, m, n, k) in tiles:
        untouched[(c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]).reshape(-1)] = False
    assert np.array_equal(got[untouched], C0[untouched])


@pytest.mark.parametrize("cplx", [True, False])
def test_grouped_gemm_small_tile_variant_against_numpy(cuda_lib, cplx):
    """ptb_gemm_grouped_v, variant 1 (64 x 32 complex / 64 x 64 real tiles): ragged extents, contraction lengths 1 ... 300,
    arbitrary (complex) / even (real) leading dimensions and offsets, untouched elements of C outside the tiles."""
    import ctypes
    from pytenet_b200.sector_packed import _tile_table
    bm, bn = ctypes.c_int(), ctypes.c_int()
    assert cuda_lib.ptb_gemm_grouped_tile_shape(1 if cplx else 0, 1, ctypes.byref(bm), ctypes.byref(bn)) == 0
    BM, BN = bm.value, bn.value
    assert (BM, BN) == ((64, 32) if cplx else (64, 64))
    rng = np.random.default_rng(14)
    nA, nB, nC = 300000, 200000, 200000
    mk = (lambda n: rng.normal(size=n) + 1j * rng.normal(size=n)) if cplx else (lambda n: rng.normal(size=n))
    A, B, C0 = mk(nA), mk(nB), mk(nC)
    ev = 1 if cplx else 2
    tiles, want = [], C0.copy()
    c_pos = 0
    for (m, n, k) in [(BM, BN, 16), (BM, BN, 300), (2, 2, 1), (6, 4, 2), (38, BN, 50), (BM, 10, 7), (34, 18, 129),
                      (BM, BN // 2, 4), (BM // 2, BN, 33), (50, 20, 1), (2, BN, 18), (BM, 2, 5), (62, 30, 255)] * 3:
        lda = m + ev * int(rng.integers(0, 5)); ldb = n + ev * int(rng.integers(0, 5)); ldc = n + int(rng.integers(0, 5))
        a_off = ev * int(rng.integers(0, (nA - k * lda - m) // ev)); b_off = ev * int(rng.integers(0, (nB - k * ldb - n) // ev))
        c_off = c_pos
        c_pos += m * ldc + 7
        assert c_pos < nC
        tiles.append((a_off, b_off, c_off, lda, ldb, ldc, m, n, k))
        a = A[a_off + np.arange(k)[:, None] * lda + np.arange(m)[None, :]]
        b = B[b_off + np.arange(k)[:, None] * ldb + np.arange(n)[None, :]]
        want[c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]] = a.T @ b
    tab = _tile_table(tiles)
    dA, dB, dC = cu(A), cu(B), cu(C0)
    dT = torch.from_numpy(tab.view(np.uint8).reshape(-1)).cuda()
    st = cuda_lib.ptb_gemm_grouped_v(1 if cplx else 0, 1, dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dT.data_ptr(),
                                     len(tab), torch.cuda.current_stream().cuda_stream)
    assert st == 0
    got = dC.cpu().numpy()
    assert rel(got, want) < 1e-13
    untouched = np.ones(nC, dtype=bool)
    for (a_off, b_off, c_off, lda, ldb, ldc, m, n, k) in tiles:
        untouched[(c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]).reshape(-1)] = False
    assert np.array_equal(got[untouched], C0[untouched])


@pytest.mark.parametrize("variant", ["0", "1"])
def test_packed_plans_with_forced_tile_variant(cuda_lib, monkeypatch, variant):
    """The packed matvec and both environment updates with every grouped GEMM forced to one tile variant: equal to
    the oracle whichever tile shape the tables were built for."""
    import pytenet_b200.sector_packed as sp
    monkeypatch.setattr(sp, "_FORCE_VARIANT", variant)
    rng = np.random.default_rng(31)
    qn, a, w, l, r = _block_sparse_problem(rng, 230, 3, 310, 4, 5)
    plan = sp.PackedHeffPlan(*qn, cplx=True)
    assert (plan.var1, plan.var3) == (int(variant), int(variant))
    got = plan.bind(cu(w), cu(l), cu(r)).apply_dense(cu(a)).cpu().numpy()
    assert rel(got, oracle.apply_local_hamiltonian(a, w, l, r)) < TOL
    er = sp.PackedEnvPlan(*qn, cplx=True, side="right")
    assert rel(er.apply(cu(a), cu(w), cu(r)).cpu().numpy(), oracle.contraction_operator_step_right(a, a, w, r)) < TOL
    el = sp.PackedEnvPlan(*qn, cplx=True, side="left")
    assert rel(el.apply(cu(a), cu(w), cu(l)).cpu().numpy(), oracle.contraction_operator_step_left(a, a, w, l)) < TOL


def test_grouped_gemm_kernel_float64_against_numpy(cuda_lib):
    """float64 tiles (128 x 128): the bulk copies move 16-byte granules, so offsets, leading dimensions and the m / n
    extents are even (what `PackedHeffPlan(cplx=False)` emits); contraction lengths arbitrary."""
    from pytenet_b200.sector_packed import _tile_table
    rng = np.random.default_rng(13)
    nA, nB, nC = 400000, 300000, 400000
    A, B, C0 = rng.normal(size=nA), rng.normal(size=nB), rng.normal(size=nC)
    tiles, want = [], C0.copy()
    c_pos = 0
    for (m, n, k) in [(128, 128, 16), (128, 128, 300), (2, 2, 1), (6, 4, 2), (38, 64, 50), (128, 10, 7), (34, 18, 129),
                      (128, 64, 4), (64, 128, 33), (100, 50, 1), (2, 128, 18), (128, 2, 5), (78, 32, 255)] * 3:
        lda = m + 2 * int(rng.integers(0, 5)); ldb = n + 2 * int(rng.integers(0, 5)); ldc = n + int(rng.integers(0, 5))
        a_off = 2 * int(rng.integers(0, (nA - k * lda - m) // 2)); b_off = 2 * int(rng.integers(0, (nB - k * ldb - n) // 2))
        c_off = c_pos
        c_pos += m * ldc + 7
        assert c_pos < nC
        tiles.append((a_off, b_off, c_off, lda, ldb, ldc, m, n, k))
        a = A[a_off + np.arange(k)[:, None] * lda + np.arange(m)[None, :]]
        b = B[b_off + np.arange(k)[:, None] * ldb + np.arange(n)[None, :]]
        want[c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]] = a.T @ b
    tab = _tile_table(tiles)
    dA, dB, dC = cu(A), cu(B), cu(C0)
    dT = torch.from_numpy(tab.view(np.uint8).reshape(-1)).cuda()
    st = cuda_lib.ptb_gemm_grouped(0, dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), dT.data_ptr(), len(tab),
                                   torch.cuda.current_stream().cuda_stream)
    assert st == 0
    got = dC.cpu().numpy()
    assert rel(got, want) < 1e-13
    untouched = np.ones(nC, dtype=bool)
    for (a_off, b_off, c_off, lda, ldb, ldc, m, n, k) in tiles:
        untouched[(c_off + np.arange(m)[:, None] * ldc + np.arange(n)[None, :]).reshape(-1)] = False
    assert np.array_equal(got[untouched], C0[untouched])


def _block_sparse_problem(rng, Dl, d, Dr, cl, cr, lo=-2, hi=2, cplx=True):
    ql = np.sort(rng.integers(lo, hi + 1, size=Dl)); qr = np.sort(rng.integers(lo - 1, hi + 1, size=Dr))
    qs = rng.integers(-1, 2, size=d)
    qwl, qwr = rng.integers(-1, 2, size=cl), rng.integers(-1, 2, size=cr)

    def tensor(shape, qn):
        t = rng.normal(size=shape) + (1j * rng.normal(size=shape) if cplx else 0)
        ob.enforce_qsparsity(t, qn)
        return t
    a = tensor((Dl, d, Dr), [ql, qs, -qr])
    l = tensor((Dl, cl, Dl), [ql, qwl, -ql])
    r = tensor((Dr, cr, Dr), [qr, qwr, -qr])
    w = rng.normal(size=(cl, d, d, cr)); ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr])
    return (ql, qs, qr, qwl, qwr), a, w, l, r


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("seed,dims", [(1, (300, 2, 260, 5, 5)), (2, (150, 4, 330, 6, 6)), (3, (420, 3, 129, 4, 5)),
                                       (4, (64, 16, 64, 6, 6)), (5, (201, 3, 177, 3, 4))])
def test_packed_matvec_matches_oracle(cuda_lib, seed, dims, cplx):
    """PackedHeffOperator: pack -> grouped GEMM -> W gather -> grouped GEMM -> repack -> unpack equals the oracle's
    dense contraction; forbidden entries of the result are exactly zero; pack / unpack are inverse on allowed
    entries; the packed matvec is linear and reusable across vectors."""
    import pytenet_b200 as ptb
    from pytenet_b200.sector_packed import PackedHeffPlan
    rng = np.random.default_rng(seed)
    qn, a, w, l, r = _block_sparse_problem(rng, *dims, cplx=cplx)
    plan = PackedHeffPlan(*qn, cplx=cplx)
    assert plan.supported
    op = plan.bind(cu(w), cu(l), cu(r))
    x = op.pack(cu(a))
    assert x.dtype == (torch.complex128 if cplx else torch.float64)
    assert rel(op.unpack(x).cpu().numpy(), a) == 0.0
    got = op.unpack(op(x)).cpu().numpy()
    ref = oracle.apply_local_hamiltonian(a, w, l, r)
    assert rel(got, ref) < TOL
    mask = ob.qnumber_outer_sum([qn[0], qn[1], -qn[2]]) != 0
    assert np.all(got[mask] == 0)
    # second vector on the same operator, linearity in the packed space
    a2 = a[::-1].copy() * ((0.3 - 0.7j) if cplx else 0.3); ob.enforce_qsparsity(a2, [qn[0], qn[1], -qn[2]])
    x2 = op.pack(cu(a2))
    y12 = op(x + 2 * x2)
    assert (torch.linalg.norm(y12 - (op(x) + 2 * op(x2))) / torch.linalg.norm(y12)).item() < 1e-13
    assert rel(op.apply_dense(cu(a2)).cpu().numpy(), oracle.apply_local_hamiltonian(a2, w, l, r)) < TOL
    # inner products in the packed space are those of the dense tensors
    assert abs(torch.vdot(x, x2).item() - np.vdot(a, a2)) < 1e-10 * np.linalg.norm(a) * np.linalg.norm(a2)
    dense = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
    assert rel(got, dense) < 1e-13


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("seed,dims", [(1, (300, 2, 260, 5, 5)), (2, (150, 4, 330, 6, 6)), (3, (421, 3, 129, 4, 5)),
                                       (4, (1, 4, 4, 1, 6)), (5, (700, 4, 700, 6, 6))])
def test_packed_environment_update_matches_oracle(cuda_lib, seed, dims, side, cplx):
    """PackedEnvPlan.apply == the oracle's contraction_operator_step_right / _left on block-sparse inputs (1e-12),
    forbidden entries of the new block exactly zero, and equal to the dense device contraction."""
    import pytenet_b200 as ptb
    from pytenet_b200.sector_packed import PackedEnvPlan
    rng = np.random.default_rng(seed)
    qn, a, w, l, r = _block_sparse_problem(rng, *dims, cplx=cplx)
    plan = PackedEnvPlan(*qn, cplx=cplx, side=side)
    assert plan.supported
    if side == "right":
        got = plan.apply(cu(a), cu(w), cu(r)).cpu().numpy()
        ref = oracle.contraction_operator_step_right(a, a, w, r)
        dense = ptb.contraction_operator_step_right(cu(a), cu(a), cu(w), cu(r)).cpu().numpy()
        mask = ob.qnumber_outer_sum([qn[0], qn[3], -qn[0]]) != 0
    else:
        got = plan.apply(cu(a), cu(w), cu(l)).cpu().numpy()
        ref = oracle.contraction_operator_step_left(a, a, w, l)
        dense = ptb.contraction_operator_step_left(cu(a), cu(a), cu(w), cu(l)).cpu().numpy()
        mask = ob.qnumber_outer_sum([qn[2], qn[4], -qn[2]]) != 0
    assert got.shape == ref.shape
    assert rel(got, ref) < TOL
    assert rel(got, dense) < 1e-13
    assert np.all(got[mask] == 0)


def test_config3_reduced_D_reference_generator_profile_vs_oracle(cuda_lib):
    """BASELINE config 3 at reduced D: Fermi-Hubbard chain with (N, Sz) quantum numbers, random MPS from
    `MPS.construct_random` (the reference generator's fragmented sector profile, bit-equal draws), actual
    right-orthonormalised environments, two-site problem at the centre -- packed matvec and banded matvec against
    the ORACLE on the same host tensors (1e-12), and a Lanczos run in the packed space against the oracle's."""
    import pytenet_b200 as ptb
    import oracle.lanczos as ol
    from pytenet_b200 import _sweep
    from pytenet_b200.sector_packed import PackedHeffPlan
    from pytenet_b200.sectors import HeffSectorPlan
    L, D = 10, 160
    h = ptb.fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.3)
    sector = ptb.encode_quantum_number_pair(L, 0)
    psi = ptb.MPS.construct_random(L, h.qsite, sector, max_vdim=D, dtype="complex", rng=np.random.default_rng(7))
    psi.orthonormalize(mode="left")            # sweeps leave the bonds grouped by sector
    _, lblocks, rblocks = _sweep.prepare_environments(h, psi)
    i = L // 2 - 1
    for j in range(i):
        lblocks[j + 1] = ptb.contraction_operator_step_left(psi.a[j], psi.a[j], h.a[j], lblocks[j])
    merged = ptb.mps_merge_tensor_pair(psi.a[i], psi.a[i + 1])
    w2 = ptb.mpo_merge_tensor_pair(h.a[i], h.a[i + 1])
    qs2 = ptb.qnumber_flatten([psi.qsite, psi.qsite])
    qn = (psi.qbonds[i], qs2, psi.qbonds[i + 2], h.qbonds[i], h.qbonds[i + 2])
    l, r = lblocks[i], rblocks[i + 1]
    assert min(merged.shape[0], merged.shape[2]) >= 100
    ah, wh, lh, rh = (t.cpu().numpy() for t in (merged, w2, l, r))
    ref = oracle.apply_local_hamiltonian(ah, wh, lh, rh)
    plan = PackedHeffPlan(*qn, cplx=True)
    assert plan.supported
    op = plan.bind(w2, l, r)
    assert rel(op.apply_dense(merged).cpu().numpy(), ref) < TOL
    band = HeffSectorPlan(*qn, cplx=True)
    assert rel(band.apply(merged, w2, l, r).cpu().numpy(), ref) < TOL
    # Lanczos in the packed space == the oracle's Lanczos on the dense vectors
    al_o, be_o, _ = ol.lanczos_iteration(lambda v: oracle.apply_local_hamiltonian(v.reshape(ah.shape), wh, lh, rh).reshape(-1),
                                         ah.reshape(-1), 8)
    al, be, _ = ptb.lanczos_iteration(op, op.pack(merged), 8)
    assert np.max(np.abs(al - al_o)) < 1e-10 * np.max(np.abs(al_o))
    assert np.max(np.abs(be - be_o)) < 1e-10 * np.max(np.abs(be_o))
    fc = plan.flop_counts()
    assert fc["visited"] < 40 * fc["exact"]            # fragmented sectors (median ~ 10): tiles mostly padding


def _config3_inputs(ptb, D):
    from pytenet_b200 import hamiltonian as ham
    qsite, qb, wbulk, _, _ = ham._fermi_hubbard_bulk(1.0, 4.0, 0.0)
    qsite = np.array(qsite); qb = np.array(qb)
    qs2 = np.add.outer(qsite, qsite).reshape(-1)
    w2 = np.einsum("kpqm,mrsn->kprqsn", wbulk, wbulk).reshape(6, 16, 16, 6)
    cand = [(dn, ds) for dn in range(-4, 5) for ds in range(-4, 5) if (dn + ds) % 2 == 0]
    wts = np.array([np.exp(-(dn ** 2 + ds ** 2) / (2 * 1.6 ** 2)) for dn, ds in cand])
    sizes = np.floor(wts / wts.sum() * D).astype(int)
    sizes[np.argmax(sizes)] += D - sizes.sum()
    q = np.sort(np.concatenate([np.full(sz, ptb.encode_quantum_number_pair(32 + dn, ds))
                                for (dn, ds), sz in zip(cand, sizes)]))
    return q, qs2, qb, w2


def test_packed_path_at_config3_shape(cuda_lib):
    """Full BASELINE config-3 shape (a (2048,16,2048), h2 (6,16,16,6), 37 (N,Sz) sectors): the sector-packed matvec
    equals the dense device matvec to 1e-13, forbidden entries stay exactly zero, executed / exact flops <= 1.3
    (VERDICT r1 target), and a fixed-size oracle comparison on one output sector block (the full oracle product
    would be 1.4e13 flop)."""
    import pytenet_b200 as ptb
    from pytenet_b200.sector_packed import PackedHeffPlan
    D = 2048
    q, qs2, qb, w2 = _config3_inputs(ptb, D)
    g = torch.Generator(device="cuda").manual_seed(3)

    def tensor(shape, qn):
        t = torch.randn(*shape, dtype=torch.complex128, device="cuda", generator=g)
        ptb.enforce_qsparsity(t, qn)
        return t
    l = tensor((D, 6, D), [q, qb, -q]); r = tensor((D, 6, D), [q, qb, -q])
    w = torch.from_numpy(w2).cuda()
    plan = PackedHeffPlan(q, qs2, q, qb, qb, cplx=True)
    fc = plan.flop_counts()
    assert fc["visited"] / fc["exact"] <= 1.3
    op = plan.bind(w, l, r)
    for _ in range(2):
        a = tensor((D, 16, D), [q, qs2, -q])
        got = op.apply_dense(a)
        want = ptb.apply_local_hamiltonian(a, w, l, r)
        assert (torch.linalg.norm(got - want) / torch.linalg.norm(want)).item() < 1e-13
        mask = (torch.from_numpy(q).cuda()[:, None, None] + torch.from_numpy(qs2).cuda()[None, :, None]
                - torch.from_numpy(q).cuda()[None, None, :]) != 0
        assert not bool(torch.any((got != 0) & mask).item())
        del want, mask
    # oracle on the rows of one left sector: out[i' in alpha', :, :] needs l[:, :, alpha'] only
    vals, starts = np.unique(q, return_index=True)
    s0 = int(starts[len(vals) // 2]); s1 = int(starts[len(vals) // 2 + 1])
    sub = oracle.apply_local_hamiltonian(a.cpu().numpy(), w2, l[:, :, s0:s1].cpu().numpy(), r.cpu().numpy())
    assert rel(got[s0:s1].cpu().numpy(), sub) < TOL


def test_packed_zero_site_problem_and_banded_absorption(cuda_lib):
    """The zero-site (bond) contraction through the sector-packed plan -- the site contraction with a one-dimensional
    physical index and the identity as MPO tensor (`_sweep.bond_plan`) -- and the gauge absorption through the banded
    GEMM (`sectors.AbsorbSectorPlan`), both against the oracle / NumPy on block-sparse inputs."""
    import pytenet_b200 as ptb
    from pytenet_b200 import _sweep
    from pytenet_b200.sector_packed import PackedHeffPlan
    from pytenet_b200.sectors import AbsorbSectorPlan
    rng = np.random.default_rng(19)
    Dl, Dr, chi, d = 300, 270, 5, 3
    qbl = np.sort(rng.integers(-2, 3, size=Dl)); qbr = np.sort(rng.integers(-2, 3, size=Dr))
    qw = rng.integers(-1, 2, size=chi)

    def tensor(shape, qn):
        t = rng.normal(size=shape) + 1j * rng.normal(size=shape)
        ob.enforce_qsparsity(t, qn)
        return t
    c = tensor((Dl, Dr), [qbl, -qbr])
    l = tensor((Dl, chi, Dl), [qbl, qw, -qbl]); r = tensor((Dr, chi, Dr), [qbr, qw, -qbr])
    plan = PackedHeffPlan(qbl, np.zeros(1, dtype=np.int64), qbr, qw, qw)
    assert plan.supported
    op = plan.bind(_sweep._identity_w(chi, torch.device("cuda", 0)), cu(l), cu(r))
    got = op.apply_dense(cu(c).reshape(Dl, 1, Dr)).reshape(Dl, Dr).cpu().numpy()
    assert rel(got, oracle.apply_local_bond_contraction(c, l, r)) < TOL
    # the driver-level entry: exp(-dt K_eff) c through the packed Lanczos run == through the dense operator
    # (same block-sparse l, r: the Lanczos recursion is the same arithmetic on the allowed entries whether or not
    # the operator is Hermitian, so the two runs must agree to rounding)
    import pytenet_b200._sweep as sw
    y_packed = sw.local_bond_step(cu(l), cu(r), cu(c), 0.01j, 6, PackedHeffPlan(qbl, np.zeros(1, dtype=np.int64), qbr, qw, qw))
    y_dense = sw.local_bond_step(cu(l), cu(r), cu(c), 0.01j, 6, None)
    assert (torch.linalg.norm(y_packed - y_dense) / torch.linalg.norm(y_dense)).item() < 1e-10
    # gauge absorption, both sides
    qs = rng.integers(-1, 2, size=d)
    qn = np.sort(rng.integers(-2, 3, size=280))            # the other bond of the site tensor
    a = tensor((Dr, d, 280), [qbr, qs, -qn])                # left bond = columns of c
    got = AbsorbSectorPlan(qbl, qbr, qs, qn, True).apply(cu(c), cu(a)).cpu().numpy()
    assert rel(got, np.tensordot(c, a, 1)) < 1e-13
    b = tensor((280, d, Dl), [qn, qs, -qbl])                # right bond = rows of c
    got = AbsorbSectorPlan(qbl, qbr, qs, qn, False).apply(cu(c), cu(b)).cpu().numpy()
    assert rel(got, np.tensordot(b, c, 1)) < 1e-13
