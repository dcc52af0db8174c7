"""CPU: the host-side tables of the sector-packed matvec (pytenet_b200/sector_packed.py) -- packed layouts,
grouped-GEMM tile lists, block-gather chunk / term lists -- executed by a NumPy emulation of the two device kernels
(`ptb_gemm_grouped`, `ptb_block_gather`, same table semantics as include/pytenet_b200.h) and compared with the
oracle's dense contraction on block-sparse inputs.  No GPU: the kernels themselves are covered in
tests/test_sector_packed_gpu.py."""
import numpy as np
import pytest

import oracle
import oracle.blocksparse as ob


def emu_gather(tables, src, dst):
    ch, tm, wk = tables
    for (ci, r0, nr, _) in wk:
        c = ch[ci]
        for r in range(r0, r0 + nr):
            row = np.zeros(c["cols"], dtype=dst.dtype)
            for t in range(c["t0"], c["t1"]):
                term = tm[t]
                idx = term["src_off"] + r * term["rs"] + np.arange(c["cols"]) * term["cs"]
                coef = term["re"] + 1j * term["im"] if np.iscomplexobj(dst) else term["re"]
                row += coef * (np.conj(src[idx]) if c["flags"] & 1 else src[idx])
            dst[c["dst_off"] + r * c["dst_ld"]: c["dst_off"] + r * c["dst_ld"] + c["cols"]] = row


def emu_grouped(tab, A, B, C):
    for t in tab:
        k, m, n = t["k"], t["m"], t["n"]
        a = A[t["a"] + np.arange(k)[:, None] * t["lda"] + np.arange(m)[None, :]]
        b = B[t["b"] + np.arange(k)[:, None] * t["ldb"] + np.arange(n)[None, :]]
        idx = t["c"] + np.arange(m)[:, None] * t["ldc"] + np.arange(n)[None, :]
        C[idx] = a.T @ b


def sorted_bond(rng, n, lo, hi):
    return np.sort(rng.integers(lo, hi + 1, size=n))


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("seed,Dl,d,Dr,cl,cr", [(1, 70, 2, 61, 4, 5), (2, 150, 4, 90, 3, 3), (3, 33, 3, 200, 5, 2)])
def test_packed_tables_reproduce_the_dense_contraction(cuda_lib, seed, Dl, d, Dr, cl, cr, cplx):
    from pytenet_b200.sector_packed import PackedHeffPlan
    rng = np.random.default_rng(seed)
    ql, qr = sorted_bond(rng, Dl, -2, 2), sorted_bond(rng, Dr, -3, 2)
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
    plan = PackedHeffPlan(ql, qs, qr, qwl, qwr, cplx=cplx)
    assert plan.supported
    z = lambda n: np.zeros(max(n, 1), dtype=complex if cplx else float)          # noqa: E731
    x, rb, lp, t1, t2, o, y = z(plan.nX), z(plan.nRB), z(plan.nLP), z(plan.nT1), z(plan.nT2), z(plan.nO), z(plan.nX)
    t1[:] = np.nan; t2[:] = np.nan; o[:] = np.nan; y[:] = np.nan        # every entry must be written by the tables
    emu_gather(plan.g_host["pack_a"], a.reshape(-1), x)
    emu_gather(plan.g_host["pack_r"], r.reshape(-1), rb)
    emu_gather(plan.g_host["pack_l"], l.reshape(-1), lp)
    # the packed vector holds exactly the allowed entries of a
    if cplx:
        assert plan.nX == int(np.count_nonzero(ob.qnumber_outer_sum([ql, qs, -qr]) == 0))
    else:           # float64: leading dimensions padded to even, offsets of every GEMM operand row 16-byte granular
        for tab in (plan.tiles1_host, plan.tiles3_host):
            for key in ("a", "b", "lda", "ldb", "m", "n"):
                assert np.all(tab[key] % 2 == 0), key
    assert abs(np.linalg.norm(x[:plan.nX]) - np.linalg.norm(a)) < 1e-12
    emu_grouped(plan.tiles1_host, x, rb, t1)
    emu_gather(plan.w_tables_host(w), t1, t2)
    emu_grouped(plan.tiles3_host, t2, lp, o)
    emu_gather(plan.g_host["repack"], o, y)
    written = np.zeros(plan.nX, dtype=bool)            # every non-padding entry of the packed result is written
    ch, _, _ = plan.g_host["repack"]
    for c in ch:
        for r_ in range(c["rows"]):
            written[c["dst_off"] + r_ * c["dst_ld"]: c["dst_off"] + r_ * c["dst_ld"] + c["cols"]] = True
    assert not np.any(np.isnan(y[:plan.nX][written]))
    out = np.zeros(Dl * d * Dr, dtype=complex if cplx else float)
    emu_gather(plan.g_host["unpack"], y, out)
    ref = oracle.apply_local_hamiltonian(a, w, l, r)
    assert np.linalg.norm(out.reshape(Dl, d, Dr) - ref) / np.linalg.norm(ref) < 1e-13
    # tiles are sorted by decreasing contraction length and never exceed the engine's tile
    for tab, (BM, BN) in ((plan.tiles1_host, plan.tile1), (plan.tiles3_host, plan.tile3)):
        assert np.all(np.diff(tab["k"]) <= 0) and tab["m"].max() <= BM and tab["n"].max() <= BN
    fc = plan.flop_counts()
    assert fc["visited"] >= fc["exact"] > 0


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("side", ["right", "left"])
@pytest.mark.parametrize("seed,Dl,d,Dr,cl,cr", [(1, 70, 2, 61, 4, 5), (2, 150, 4, 90, 3, 3), (3, 33, 3, 200, 5, 2)])
def test_packed_environment_tables_reproduce_the_dense_update(cuda_lib, seed, Dl, d, Dr, cl, cr, side, cplx):
    """PackedEnvPlan: the tables of contraction_operator_step_right / _left (pack, T1^T, W gather, conjugated bra
    pack, K-split GEMM, unpack) emulated in NumPy against the oracle's dense update on block-sparse inputs."""
    from pytenet_b200.sector_packed import PackedEnvPlan
    rng = np.random.default_rng(seed)
    ql, qr = sorted_bond(rng, Dl, -2, 2), sorted_bond(rng, Dr, -3, 2)
    qs = rng.integers(-1, 2, size=d)
    qwl, qwr = rng.integers(-1, 2, size=cl), rng.integers(-1, 2, size=cr)

    def tensor(shape, qn):
        t = rng.normal(size=shape) + (1j * rng.normal(size=shape) if cplx else 0)
        ob.enforce_qsparsity(t, qn)
        return t
    a = tensor((Dl, d, Dr), [ql, qs, -qr])
    w = rng.normal(size=(cl, d, d, cr)); ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr])
    if side == "right":
        env = tensor((Dr, cr, Dr), [qr, qwr, -qr])
        ref = oracle.contraction_operator_step_right(a, a, w, env)
    else:
        env = tensor((Dl, cl, Dl), [ql, qwl, -ql])
        ref = oracle.contraction_operator_step_left(a, a, w, env)
    plan = PackedEnvPlan(ql, qs, qr, qwl, qwr, cplx=cplx, side=side)
    assert plan.supported
    base = plan.base
    z = lambda n: np.zeros(max(n, 1), dtype=complex if cplx else float)          # noqa: E731
    x, rb, t1, t2, bt, o = z(base.nX), z(base.nRB), z(base.nT1), z(base.nT2T), z(plan.nBT), z(plan.nO)
    for buf in (rb, t1, t2, bt, o):
        buf[:] = np.nan                                    # every entry that is read must have been written
    emu_gather(plan.g_host["pack_a"], a.reshape(-1), x)
    emu_gather(plan.g_host["pack_r"], env.reshape(-1), rb)
    assert abs(np.linalg.norm(x) - np.linalg.norm(a)) < 1e-12
    emu_grouped(plan.tiles1, rb, x, t1)
    emu_gather(plan.w_tables_host(w), t1, t2)
    emu_gather(plan.g_host["pack_bt"], x, bt)
    emu_grouped(plan.tiles3, t2, bt, o)
    out = np.zeros(ref.size, dtype=complex if cplx else float)
    emu_gather(plan.g_host["unpack"], o, out)
    assert not np.any(np.isnan(out))
    assert np.linalg.norm(out.reshape(ref.shape) - ref) / np.linalg.norm(ref) < 1e-13
    if not cplx:
        for tab in (plan.tiles1, plan.tiles3):
            for key in ("a", "b", "lda", "ldb", "m", "n"):
                assert np.all(tab[key] % 2 == 0), key


def test_unsorted_bonds_are_left_to_the_banded_path(cuda_lib):
    from pytenet_b200.sector_packed import PackedHeffPlan
    plan = PackedHeffPlan([0, 1, 0, 1], [0, 1], [0, 1, 2], [0], [0], cplx=True)
    assert not plan.supported
    assert PackedHeffPlan([0, 0, 1], [0, 1], [0, 1, 2], [0], [0], cplx=False).supported
