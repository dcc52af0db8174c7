"""CPU: the quantum-number sector plan (host logic of the block-sparse path).  The k-tile ranges
must contain every non-zero contribution: a NumPy emulation of the banded, batched GEMMs that only
visits the planned ranges reproduces the oracle's dense result exactly."""
import numpy as np
import pytest

import oracle
import oracle.blocksparse as ob
from pytenet_b200.sectors import HeffSectorPlan, tile_k_ranges


def crand(rng, shape):
    return rng.normal(size=shape) + 1j * rng.normal(size=shape)


def make_case(rng, Dl, d, Dr, cl, cr, sort=True, spread=2):
    qs = rng.integers(-1, 2, size=d)
    ql = rng.integers(-spread, spread + 1, size=Dl)
    qr = rng.integers(-spread, spread + 1, size=Dr)
    qwl = rng.integers(-1, 2, size=cl)
    qwr = rng.integers(-1, 2, size=cr)
    if sort:
        ql = np.sort(ql); qr = np.sort(qr)
    a = crand(rng, (Dl, d, Dr)); ob.enforce_qsparsity(a, [ql, qs, -qr])
    l = crand(rng, (Dl, cl, Dl)); ob.enforce_qsparsity(l, [ql, qwl, -ql])
    r = crand(rng, (Dr, cr, Dr)); ob.enforce_qsparsity(r, [qr, qwr, -qr])
    w = rng.normal(size=(cl, d, d, cr)); ob.enforce_qsparsity(w, [qwl, qs, -qs, -qwr])
    return (a, w, l, r), (ql, qs, qr, qwl, qwr)


def emulate(plan, a, w, l, r):
    """NumPy emulation of HeffSectorPlan.apply visiting only the planned k-tile ranges."""
    Dl, d, Dr, cl, cr, dout, Dlp, Drp = plan.dims
    bm, bn, bk = plan.tile
    t1 = np.zeros((Dl, d, cr * Drp), dtype=complex)
    rm = r.reshape(Dr, cr * Drp)
    for s in range(d):
        tab = plan.tab1_host[s]
        for tm in range(tab.shape[0]):
            for tn in range(tab.shape[1]):
                lo, hi = tab[tm, tn] * bk
                ms = slice(tm * bm, min((tm + 1) * bm, Dl)); ns = slice(tn * bn, min((tn + 1) * bn, cr * Drp))
                t1[ms, s, ns] = a[ms, s, lo:hi] @ rm[lo:hi, ns]
    t2 = np.einsum("kpsc,iscj->ikpj", w, t1.reshape(Dl, d, cr, Drp))
    out = np.zeros((Dlp, dout, Drp), dtype=complex)
    for k in range(cl):
        if not plan.k_active[k]:
            continue
        tabs = plan.tab3_host[k]
        for sp in range(dout):
            for tm in range(tabs.shape[1]):
                for tn in range(tabs.shape[2]):
                    lo, hi = tabs[sp, tm, tn] * bk
                    ms = slice(tm * bm, min((tm + 1) * bm, Dlp)); ns = slice(tn * bn, min((tn + 1) * bn, Drp))
                    out[ms, sp, ns] += l[lo:hi, k, ms].T @ t2[lo:hi, k, sp, ns]
    return out


@pytest.mark.parametrize("sort", [True, False])
def test_plan_contains_all_nonzero_contributions(cuda_lib, sort):
    rng = np.random.default_rng(5 + sort)
    for (Dl, d, Dr, cl, cr) in [(150, 2, 170, 5, 5), (260, 3, 140, 4, 6), (64, 4, 300, 6, 3)]:
        (a, w, l, r), (ql, qs, qr, qwl, qwr) = make_case(rng, Dl, d, Dr, cl, cr, sort=sort)
        plan = HeffSectorPlan(ql, qs, qr, qwl, qwr, cplx=True)
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        got = emulate(plan, a, w, l, r)
        assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)
        if sort:
            f1, f3 = plan.visit_fraction
            assert f1 < 0.75 and f3 < 0.75, (f1, f3)      # sorted sectors must actually skip work


def test_zero_quantum_numbers_give_full_ranges(cuda_lib):
    plan = HeffSectorPlan(np.zeros(130, int), np.zeros(2, int), np.zeros(200, int), np.zeros(5, int), np.zeros(5, int))
    assert plan.trivial()
    bm, bn, bk = plan.tile
    assert np.all(plan.tab1_host[..., 0] == 0) and np.all(plan.tab1_host[..., 1] == -(-200 // bk))
    assert plan.visit_fraction == (1.0, 1.0)


def test_tile_k_ranges_bounds():
    qk = np.array([0, 0, 0, 1, 1, 2, 2, 2, 2, 5])
    tab = tile_k_ranges(np.array([0, 0, 1, 7]), np.array([0, 1, 1, 2]), qk, bm=2, bn=2, bk=2)
    # rows tile 0 -> value 0 -> k in [0,3); cols tile 0 -> values {0,1} -> [0,5): intersection [0,3) -> k-tiles [0,2)
    assert list(tab[0, 0]) == [0, 2]
    # rows tile 1 -> values {1,7}: 7 absent -> [3,5); cols tile 1 -> {1,2} -> [3,9): -> [3,5) -> tiles [1,3)
    assert list(tab[1, 1]) == [1, 3]
    # rows tile 0 ([0,3)) with cols tile 1 ([3,9)): empty
    assert list(tab[0, 1]) == [0, 0]


def emulate_lean(plan, a, w, l, r, t1, t2):
    """NumPy emulation of the lean sector path as the device executes it: step 1 stores only tiles with a
    non-empty k range (in the scheduled order), the W step obeys the row-activity flags, step 3 sums the
    per-tile segment lists.  `t1`, `t2` are persistent buffers whose structural zeros are never rewritten."""
    Dl, d, Dr, cl, cr, dout, Dlp, Drp = plan.dims
    bm, bn, bk = plan.tile
    rm = r.reshape(Dr, cr * Drp)
    tab = plan.tab1_host
    tiles_m, tiles_n = tab.shape[1], tab.shape[2]
    assert sorted(plan.order1_host.tolist()) == list(range(tab[..., 0].size))
    for t in plan.order1_host:                                      # (batch s, tile row, tile column) in table order
        s, rem = divmod(int(t), tiles_m * tiles_n)
        tm, tn = divmod(rem, tiles_n)
        lo, hi = tab[s, tm, tn] * bk
        if hi <= lo:
            continue                                                # accumulate == 2: empty tiles are left untouched
        ms = slice(tm * bm, min((tm + 1) * bm, Dl)); ns = slice(tn * bn, min((tn + 1) * bn, cr * Drp))
        t1[ms, s, ns] = a[ms, s, lo:hi] @ rm[lo:hi, ns]
    # W step with flags per (bm-row block of i, 128-column block of j')
    wm = w.reshape(cl * dout, d * cr)
    ina = plan.in_active_host
    t1v = t1.reshape(Dl, d * cr, Drp)
    for ib in range(ina.shape[0]):
        for jb in range(ina.shape[1]):
            isl = slice(ib * bm, min((ib + 1) * bm, Dl)); jsl = slice(jb * 128, min((jb + 1) * 128, Drp))
            fin = ina[ib, jb]
            fout = np.any((wm != 0) & fin[None, :], axis=1)         # what _w_flags derives from W's pattern
            for m in np.nonzero(fout)[0]:
                cols = np.nonzero((wm[m] != 0) & fin)[0]
                t2[isl, m, jsl] = np.einsum("c,icj->ij", wm[m, cols], t1v[isl][:, cols][:, :, jsl])
    # step 3: segments
    out = np.zeros((Dlp, dout, Drp), dtype=complex)
    t2v = t2.reshape(Dl, cl, dout, Drp)
    sp_, segs = plan.seg_ptr_host, plan.segs_host
    t3m, t3n = -(-Dlp // bm), -(-Drp // bn)
    assert sorted(plan.order3_host.tolist()) == list(range(dout * t3m * t3n))
    for t in plan.order3_host:
        b, rem = divmod(int(t), t3m * t3n)
        tm, tn = divmod(rem, t3n)
        ms = slice(tm * bm, min((tm + 1) * bm, Dlp)); ns = slice(tn * bn, min((tn + 1) * bn, Drp))
        for lo, hi, k, _ in segs[sp_[t]:sp_[t + 1]]:
            assert plan.sel_off_host[k, 0] == k * Dlp and plan.sel_off_host[k, 1] == k * dout * Drp
            out[ms, b, ns] += l[lo * bk:hi * bk, k, ms].T @ t2v[lo * bk:hi * bk, k, b, ns]
    return out


def test_lean_segmented_path_with_persistent_intermediates(cuda_lib):
    """Structural zeros written once: the emulated lean path stays exact over several vectors and MPO tensors
    on the same persistent t1 / t2 buffers (stale entries of earlier applications must never leak)."""
    rng = np.random.default_rng(11)
    (a, w, l, r), (ql, qs, qr, qwl, qwr) = make_case(rng, 300, 3, 280, 5, 4, sort=True, spread=4)
    plan = HeffSectorPlan(ql, qs, qr, qwl, qwr, cplx=True)
    Dl, d, Dr, cl, cr, dout, Dlp, Drp = plan.dims
    t1 = np.zeros((Dl, d, cr * Drp), dtype=complex)
    t2 = np.zeros((Dl, cl * dout, Drp), dtype=complex)
    assert plan.in_active_host.mean() < 0.9                         # the flags actually skip rows
    for trial in range(3):
        a2 = crand(rng, a.shape); ob.enforce_qsparsity(a2, [ql, qs, -qr])
        got = emulate_lean(plan, a2, w, l, r, t1, t2)
        ref = oracle.apply_local_hamiltonian(a2, w, l, r)
        assert np.linalg.norm(got - ref) <= 1e-13 * np.linalg.norm(ref)
    fc = plan.flop_counts(nnz_w=int(np.count_nonzero(w)))
    assert 0 < fc["exact"] <= fc["visited"] <= 8.0 * (Dl * d * Dr * cr * Drp + Dlp * Dl * cl * dout * Drp) * 4


def test_plan_cache_returns_same_plan(cuda_lib):
    from pytenet_b200 import _sweep
    q = np.sort(np.random.default_rng(0).integers(-2, 3, size=300)); qs = np.array([0, 1]); qw = np.array([0, 1, -1])
    p1 = _sweep._cached_plan(HeffSectorPlan, True, q, qs, q, qw, qw)
    p2 = _sweep._cached_plan(HeffSectorPlan, True, q.copy(), qs, q, qw, qw)
    p3 = _sweep._cached_plan(HeffSectorPlan, False, q, qs, q, qw, qw)
    assert p1 is p2 and p3 is not p1 and p3.cplx is False


def test_vectorised_tables_equal_their_definitions(cuda_lib):
    """The plans build their work lists by shift-deduplicated, batched lookups; every table must equal the
    one-at-a-time definition through tile_k_ranges."""
    from pytenet_b200.sectors import EnvSectorPlan, BondSectorPlan
    rng = np.random.default_rng(1)
    ql = np.sort(rng.integers(-3, 4, size=300)); qr = np.sort(rng.integers(-3, 4, size=260))
    qs = np.array([0, 1, -1, 2]); qwl = np.array([0, 1, -1, 0, 2]); qwr = np.array([0, -1, 1])
    d, cl, cr = len(qs), len(qwl), len(qwr)
    hp = HeffSectorPlan(ql, qs, qr, qwl, qwr)
    bm, bn, bk = hp.tile
    cols = (qr[None, :] - qwr[:, None]).reshape(-1)
    for s in range(d):
        assert np.array_equal(hp.tab1_host[s], tile_k_ranges(ql + qs[s], cols, qr, bm, bn, bk))
    for k in range(cl):
        for sp in range(d):
            assert np.array_equal(hp.tab3_host[k][sp], tile_k_ranges(ql - qwl[k], qr - qs[sp] - qwl[k], ql, bm, bn, bk))
    ep = EnvSectorPlan(ql, qs, qr, qwl, qwr)
    for k in range(cl):
        for sp in range(d):
            assert np.array_equal(ep.L1[k][sp], tile_k_ranges(ql + qwl[k], qr - qs[sp], ql, bm, bn, bk))
    for s in range(d):
        assert np.array_equal(ep.L3[s][0], tile_k_ranges(qr - qs[s], cols - qs[s], ql, bm, bn, bk))
        assert np.array_equal(ep.R1[s], tile_k_ranges(ql + qs[s], cols, qr, bm, bn, bk))
        for k in range(cl):
            assert np.array_equal(ep.R3[s][k], tile_k_ranges(ql + qwl[k] + qs[s], ql + qs[s], qr, bm, bn, bk))
    bp = BondSectorPlan(ql, ql, qwl)
    assert bp.seg_ptr_host[-1] == sum(int(np.sum(t[..., 1] > t[..., 0])) for t in bp.B2)
