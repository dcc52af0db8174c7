"""
Sector-packed effective-Hamiltonian matvec: the quantum-number block structure as a device-side grouped GEMM
(BASELINE config 3; north_star: "block_sparse_util/qnumber block structure becomes a device-side grouped, batched
GEMM over the quantum-number sectors").

The reference contracts dense tensors that hold explicit zeros (SURVEY.md headline 3).  With additive quantum
numbers (pytenet/block_sparse_util.py:47-53)

    a[i,s,j]    != 0  only if  ql[i]  + qs[s]  == qr[j]
    r[j,K,j']   != 0  only if  qr[j]  + qwr[K] == qr[j']
    l[i,k,i']   != 0  only if  ql[i]  + qwl[k] == ql[i']
    w[k,s',s,K] != 0  only if  qwl[k] + qs[s'] == qs[s] + qwr[K]

and with the bond indices grouped by sector (block_sparse_qr / block_sparse_svd emit the sectors in ascending order,
block_sparse_util.py:151-169, so every sweep leaves them grouped) all four tensors are unions of dense blocks
indexed by (left sector alpha, right sector beta).  This module stores NO structural zero and multiplies none:

  packed vector space X   for every right sector beta the matrix  AT_beta[j, (alpha,s,i)] = a[i,s,j]  (n_beta x M_beta;
                          the columns stack all blocks (alpha, s) with ql[alpha] + qs[s] == qr[beta])
  step 1 (grouped GEMM)   T1_beta (M_beta x N_beta) = AT_beta^T . RB_beta,   RB_beta[j, (K,j')] = r[j,K,j']  with the columns
                          stacking the sectors gamma(beta, K)
  step 2 (block gather)   T2_alpha'[(alpha,k,i), (s',j')] = sum_{s,K} w[k,s',s,K] T1[(alpha,s,i), (K,j')]   -- one block copy-add per
                          non-zero MPO entry and block
  step 3 (grouped GEMM)   O_alpha' (N'_alpha' x n_alpha') = T2_alpha'^T . LP_alpha',   LP_alpha'[(alpha,k,i), i'] = l[i,k,i']
  repack (block gather)   O -> X

so each group is a GEMM with two large, stacked extents and ONE sector-sized extent (k = n_beta in step 1, n = n_alpha' in
step 3).  With the 128 x 64 complex tile of the DMMA engine the executed flops are 1.2 x the exact sector-block
flops at the config-3 shape (the work lists over dense-layout tensors of sectors.py visit 3.8 x).  A whole
Lanczos run stays in X (`pack` once, `unpack` once): the Lanczos vector kernels then also touch only the
allowed entries (4 % of the dense vector at config 3).  Inner products and norms are sums over the same
non-zero entries, so alphas / betas equal the dense run's up to summation order.

complex128 and float64 (real states: every leading dimension padded to an even number of elements).
Sector layouts are untouched (qbonds / retained indices stay bit-exact).  Bonds that are not grouped by sector
(e.g. `MPS.construct_random` output before the first orthonormalisation) are not supported here -- `supported`
is False and the callers use the banded path of sectors.py.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from . import _device as dev

__all__ = ["PackedHeffPlan", "PackedHeffOperator", "PackedEnvPlan"]

_KSPLIT = 256          # step-3 groups are split along their stacked contraction index into pieces of about this size
_GATHER_ELEMS = 4096   # elements per CTA of the block-gather kernel


def _runs(q):
    """Sectors of a bond grouped by quantum number: (values, offsets, sizes), or None when some value occurs in
    more than one run (bond not grouped by sector)."""
    q = np.asarray(q, dtype=np.int64)
    if len(q) == 0:
        return None
    starts = np.flatnonzero(np.concatenate([[True], q[1:] != q[:-1]]))
    vals = q[starts]
    if len(np.unique(vals)) != len(vals):
        return None
    sizes = np.diff(np.concatenate([starts, [len(q)]]))
    return vals, starts.astype(np.int64), sizes.astype(np.int64)


_CHUNK_DT = np.dtype([("dst_off", "<i8"), ("dst_ld", "<i4"), ("rows", "<i4"), ("cols", "<i4"), ("t0", "<i4"),
                      ("t1", "<i4"), ("flags", "<i4")])
_TERM_DT = np.dtype([("src_off", "<i8"), ("rs", "<i4"), ("cs", "<i4"), ("re", "<f8"), ("im", "<f8")])
_TILE_DT = np.dtype([("a", "<i8"), ("b", "<i8"), ("c", "<i8"), ("lda", "<i4"), ("ldb", "<i4"), ("ldc", "<i4"),
                     ("m", "<i4"), ("n", "<i4"), ("k", "<i4"), ("acc", "<i4"), ("res", "<i4", (3,))])
assert _CHUNK_DT.itemsize == 32 and _TERM_DT.itemsize == 32 and _TILE_DT.itemsize == 64


def _gather_tables(dst_off, dst_ld, rows, cols, term_count, src_off, src_rs, src_cs, coef_re=None, coef_im=None,
                   conj=False):
    """Host tables of one block-gather launch (ptb_block_gather) from per-chunk arrays and per-term arrays; chunk c
    owns the next `term_count[c]` terms.  Empty chunks are dropped.  `conj`: the chunks read the complex conjugate
    of the source.  Returns (chunks, terms, work items)."""
    dst_off, dst_ld, rows, cols, term_count = (np.asarray(v, dtype=np.int64).reshape(-1)
                                               for v in (dst_off, dst_ld, rows, cols, term_count))
    t1 = np.cumsum(term_count)
    t0 = t1 - term_count
    keep = (rows > 0) & (cols > 0)
    ch = np.zeros(int(np.count_nonzero(keep)), dtype=_CHUNK_DT)
    ch["dst_off"], ch["dst_ld"], ch["rows"], ch["cols"] = dst_off[keep], dst_ld[keep], rows[keep], cols[keep]
    ch["t0"], ch["t1"] = t0[keep], t1[keep]
    ch["flags"] = 1 if conj else 0
    nt = len(np.asarray(src_off).reshape(-1))
    tm = np.zeros(max(nt, 1), dtype=_TERM_DT)
    if nt:
        tm["src_off"][:nt], tm["rs"][:nt], tm["cs"][:nt] = src_off, src_rs, src_cs
        tm["re"][:nt] = 1.0 if coef_re is None else coef_re
        tm["im"][:nt] = 0.0 if coef_im is None else coef_im
    # work items: row ranges of about _GATHER_ELEMS elements
    rows_k, cols_k = rows[keep], cols[keep]
    per = np.maximum(1, _GATHER_ELEMS // np.maximum(cols_k, 1))
    nwork = -(-rows_k // per)
    ci = np.repeat(np.arange(len(rows_k)), nwork)
    first = np.cumsum(nwork) - nwork
    r0 = (np.arange(int(nwork.sum())) - first[ci]) * per[ci]
    wk = np.zeros((len(ci), 4), dtype=np.int32)
    wk[:, 0], wk[:, 1], wk[:, 2] = ci, r0, np.minimum(per[ci], rows_k[ci] - r0)
    return ch, tm, wk


def _pack_segments(arrays):
    """One byte buffer holding the given arrays at 64-byte aligned offsets: (buffer, offsets)."""
    offs, pos = [], 0
    for a in arrays:
        offs.append(pos)
        pos += (a.nbytes + 63) // 64 * 64
    buf = np.zeros(max(pos, 64), dtype=np.uint8)
    for a, o in zip(arrays, offs):
        if a.nbytes:
            buf[o:o + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return buf, offs


class _DeviceGather:
    def __init__(self, tables, device, views=None):
        ch, tm, wk = tables
        self.nwork = len(wk)
        if views is None:
            buf, offs = _pack_segments([ch, tm, wk])
            dbuf = torch.from_numpy(buf).to(device)
            views = [dbuf[o:] for o in offs]
        self.ch, self.tm, self.wk = views

    def run(self, lib, dt, src, dst, stream):
        if self.nwork == 0:
            return
        st = lib.ptb_block_gather(dt, src.data_ptr(), dst.data_ptr(), self.ch.data_ptr(), self.tm.data_ptr(),
                                  self.wk.data_ptr(), self.nwork, stream)
        _lib.check(st, "ptb_block_gather")


def _tile_table(tiles):
    """(n, 64-byte) device-ready table from (a_off, b_off, c_off, lda, ldb, ldc, m, n, k) tuples, sorted by
    decreasing k (longest tiles first over the persistent CTAs)."""
    cols = np.asarray(tiles, dtype=np.int64).reshape(-1, 9)
    return _tile_table_cols(*[cols[:, i] for i in range(9)])


def _tile_table_cols(a, b, c, lda, ldb, ldc, m, n, k):
    order = np.argsort(-np.asarray(k, dtype=np.int64), kind="stable")
    tab = np.zeros(len(order), dtype=_TILE_DT)
    for name, v in zip(("a", "b", "c", "lda", "ldb", "ldc", "m", "n", "k"), (a, b, c, lda, ldb, ldc, m, n, k)):
        tab[name] = np.asarray(v, dtype=np.int64)[order]
    return tab


def _seg_offsets(group, size, ngroups):
    """Running (exclusive) sum of `size` inside each group, in the given order of the entries, and the group totals."""
    group = np.asarray(group, dtype=np.int64); size = np.asarray(size, dtype=np.int64)
    tot = np.bincount(group, weights=size, minlength=ngroups).astype(np.int64) if len(group) else np.zeros(ngroups, np.int64)
    if len(group) == 0:
        return np.zeros(0, dtype=np.int64), tot
    order = np.argsort(group, kind="stable")
    g, sz = group[order], size[order]
    cs = np.cumsum(sz) - sz
    first = np.concatenate([[True], g[1:] != g[:-1]])
    base = cs[first][np.cumsum(first) - 1]
    off = np.empty_like(cs)
    off[order] = cs - base
    return off, tot


def _k_pieces(klen, live):
    """Split the contraction range [0, klen[g]) of every live group g into pieces of about _KSPLIT (multiples of 16):
    (group of each piece, k0, k1, pieces per group, index of each group's first piece)."""
    npc = np.where(live, np.maximum(1, -(-klen // _KSPLIT)), 0)
    stepk = -(-klen // np.maximum(npc, 1))
    stepk = -(-stepk // 16) * 16
    npc = np.where(live, -(-klen // np.maximum(stepk, 1)), 0)
    pc_g = np.repeat(np.arange(len(klen)), npc)
    pc_first = np.cumsum(npc) - npc
    pc_p = np.arange(int(npc.sum())) - pc_first[pc_g]
    pc_k0 = pc_p * stepk[pc_g]
    pc_k1 = np.minimum(pc_k0 + stepk[pc_g], klen[pc_g])
    return pc_g, pc_k0, pc_k1, npc, pc_first


# Tile variants of the grouped GEMM (ptb_gemm_grouped_v): 0 = the engine's large tile, 1 = the small tile.  A table is
# built for both and the cheaper one is taken: cost = executed flops / relative efficiency of the variant (the small
# tile issues fewer DMMA per fragment load but measured within 5 % of the large tile's rate on full tiles, and its finer
# granularity balances the persistent CTAs better: it wins on every sector profile measured so far).
_SMALL_TILE_EFFICIENCY = 0.95
_FORCE_VARIANT = os.environ.get("PYTENET_B200_GROUPED_TILE", "auto")
_TILE_SHAPES = {}


def _tile_shapes(cplx):
    hit = _TILE_SHAPES.get(bool(cplx))
    if hit is None:
        lib = _lib.load()
        hit = []
        for variant in (0, 1):
            bm, bn = ctypes.c_int(), ctypes.c_int()
            _lib.check(lib.ptb_gemm_grouped_tile_shape(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, variant,
                                                       ctypes.byref(bm), ctypes.byref(bn)), "ptb_gemm_grouped_tile_shape")
            hit.append((bm.value, bn.value))
        _TILE_SHAPES[bool(cplx)] = hit
    return hit


def _choose_tiles(cplx, build):
    """`build(BM, BN)` -> tile table for that tile shape; returns (variant, table) of the cheaper variant."""
    best = None
    for variant, (BM, BN) in enumerate(_tile_shapes(cplx)):
        if _FORCE_VARIANT in ("0", "1") and int(_FORCE_VARIANT) != variant:
            continue
        tab = build(BM, BN)
        cost = float(BM * BN) * float(np.sum((tab["k"].astype(np.int64) + 3) // 4 * 4))
        cost /= (1.0 if variant == 0 else _SMALL_TILE_EFFICIENCY)
        if best is None or cost < best[0]:
            best = (cost, variant, tab, (BM, BN))
    return best[1], best[2], best[3]


def _tiles_of(groups_m, groups_n, BM, BN):
    """Tile enumeration of a list of (m x n) matrices: per tile (matrix index, row start, column start)."""
    ntm, ntn = -(-groups_m // BM), -(-groups_n // BN)
    cnt = ntm * ntn
    gi = np.repeat(np.arange(len(cnt)), cnt)
    local = np.arange(int(cnt.sum())) - (np.cumsum(cnt) - cnt)[gi]
    ntn_g = np.maximum(ntn[gi], 1)
    return gi, (local // ntn_g) * BM, (local % ntn_g) * BN


class PackedHeffPlan:
    """
    Everything about the sector-packed matvec that depends on the quantum numbers only: sector lists, packed
    layouts, GEMM tile tables and the gather tables that pack `a`, `l`, `r` and unpack the result.
    Arguments as `sectors.HeffSectorPlan` (bra bonds = ket bonds, as in the sweeps).  Built with array operations
    over the (sector, physical index, MPO index) combinations: a two-site TDVP step with truncation changes the
    sector layout of every bond at every split, so plans are rebuilt per local problem there and their construction
    is on the critical path of the sweep.
    """

    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True):
        self.cplx = bool(cplx)
        self.ql = np.asarray(ql, dtype=np.int64); self.qr = np.asarray(qr, dtype=np.int64)
        self.qs = np.asarray(qs, dtype=np.int64)
        self.qwl = np.asarray(qwl, dtype=np.int64); self.qwr = np.asarray(qwr, dtype=np.int64)
        Dl, d, Dr, cl, cr = len(self.ql), len(self.qs), len(self.qr), len(self.qwl), len(self.qwr)
        self.dims = (Dl, d, Dr, cl, cr)
        L, R = _runs(self.ql), _runs(self.qr)
        self.supported = L is not None and R is not None
        self._dev = {}
        self._wcache = {}
        if not self.supported:
            return
        qL, oL, nL = L
        qR, oR, nR = R
        self.L, self.R = L, R
        nLs, nRs = len(qL), len(qR)
        sortL, sortR = np.argsort(qL), np.argsort(qR)

        def lookup(vals, perm, q):
            """Index of the sector with quantum number q (array), -1 where there is none."""
            sv = vals[perm]
            pos = np.minimum(np.searchsorted(sv, q), len(sv) - 1)
            return np.where(sv[pos] == q, perm[pos], -1)
        in_L = lambda q: lookup(qL, sortL, q)          # noqa: E731
        in_R = lambda q: lookup(qR, sortR, q)          # noqa: E731
        # float64: the bulk copies of the grouped GEMM move 16-byte granules, so every leading dimension is rounded
        # up to an even number of elements (complex128 needs no padding).  Padding entries of the packed VECTORS are
        # kept at zero (they take part in the Lanczos inner products); padding of the intermediates is never read.
        ev = (lambda v: v) if self.cplx else (lambda v: v + (v & 1))
        excl = lambda v: np.cumsum(v) - v              # noqa: E731

        # ---- packed vector space X: group b (right sector), columns = stacked (alpha, s) blocks, s-major ----
        xs, xal = np.repeat(np.arange(d), nLs), np.tile(np.arange(nLs), d)
        xb = in_R(qL[xal] + self.qs[xs])
        ok = xb >= 0
        xs, xal, xb = xs[ok], xal[ok], xb[ok]
        xm, M = _seg_offsets(xb, nL[xal], nRs)
        self.XB = np.full((nLs, d), -1, dtype=np.int64); self.XM = np.zeros((nLs, d), dtype=np.int64)
        self.XB[xal, xs], self.XM[xal, xs] = xb, xm
        self.xc = (xs, xal, xb, xm)
        self.M = M
        Mp = self.Mp = ev(M)
        self.offX = excl(nR * Mp)
        self.nX = int(np.sum(nR * Mp))

        # ---- step 1 operand RB: group b, columns = stacked (K, gamma(b, K)) ----
        rb_b, rb_K = np.repeat(np.arange(nRs), cr), np.tile(np.arange(cr), nRs)
        rb_g = in_R(qR[rb_b] + self.qwr[rb_K])
        ok = rb_g >= 0
        rb_b, rb_K, rb_g = rb_b[ok], rb_K[ok], rb_g[ok]
        rb_n, N1 = _seg_offsets(rb_b, nR[rb_g], nRs)
        self.RBG = np.full((nRs, cr), -1, dtype=np.int64); self.RBN = np.zeros((nRs, cr), dtype=np.int64)
        self.RBG[rb_b, rb_K], self.RBN[rb_b, rb_K] = rb_g, rb_n
        self.rbc = (rb_b, rb_K, rb_g, rb_n)
        self.N1 = N1
        N1p = self.N1p = ev(N1)
        self.offRB = excl(nR * N1p)
        self.nRB = int(np.sum(nR * N1p))
        self.offT1 = excl(Mp * N1p)
        self.nT1 = int(np.sum(Mp * N1p))

        # ---- step 3: group a' (left sector): rows of T2 / LP = stacked (k, alpha), columns of T2 = stacked (s', gamma') ----
        tr_ap, tr_k = np.repeat(np.arange(nLs), cl), np.tile(np.arange(cl), nLs)
        tr_al = in_L(qL[tr_ap] - self.qwl[tr_k])
        ok = tr_al >= 0
        tr_ap, tr_k, tr_al = tr_ap[ok], tr_k[ok], tr_al[ok]
        tr_kk, K3 = _seg_offsets(tr_ap, nL[tr_al], nLs)
        self.tr = (tr_ap, tr_k, tr_al, tr_kk)
        tc_ap, tc_sp = np.repeat(np.arange(nLs), d), np.tile(np.arange(d), nLs)
        tc_g = in_R(qL[tc_ap] + self.qs[tc_sp])
        ok = tc_g >= 0
        tc_ap, tc_sp, tc_g = tc_ap[ok], tc_sp[ok], tc_g[ok]
        tc_n, N3 = _seg_offsets(tc_ap, nR[tc_g], nLs)
        self.TCG = np.full((nLs, d), -1, dtype=np.int64); self.TCN = np.zeros((nLs, d), dtype=np.int64)
        self.TCG[tc_ap, tc_sp], self.TCN[tc_ap, tc_sp] = tc_g, tc_n
        self.tc = (tc_ap, tc_sp, tc_g, tc_n)
        self.K3, self.N3 = K3, N3
        N3p, nLp = ev(N3), ev(nL)
        self.N3p, self.nLp = N3p, nLp
        self.offT2 = excl(K3 * N3p)
        self.nT2 = int(np.sum(K3 * N3p))
        self.offLP = excl(K3 * nLp)
        self.nLP = int(np.sum(K3 * nLp))
        self.K3p = ev(K3)                          # transposed intermediates of the environment update
        self.offT2T = excl(N3 * self.K3p)
        self.nT2T = int(np.sum(N3 * self.K3p))
        # K-split of the step-3 groups: piece p of group a' covers stacked rows [k0, k1) and writes its own
        # partial product O_(a',p); the repack gather sums the pieces (fixed order, deterministic)
        pc_ap, pc_k0, pc_k1, npc, pc_first = _k_pieces(K3, (K3 > 0) & (N3 > 0))
        pc_size = N3p[pc_ap] * nLp[pc_ap]
        pc_off = excl(pc_size)
        self.nO = int(pc_size.sum())
        self.pieces = (pc_ap, pc_k0, pc_k1, pc_off, npc, pc_first)

        # ---- GEMM tile tables ----
        live1 = np.flatnonzero((M > 0) & (N1 > 0))

        def build1(BM, BN):
            gi, tm, tn = _tiles_of(Mp[live1], N1p[live1], BM, BN)
            b = live1[gi]
            return _tile_table_cols(self.offX[b] + tm, self.offRB[b] + tn, self.offT1[b] + tm * N1p[b] + tn,
                                    Mp[b], N1p[b], N1p[b], np.minimum(BM, Mp[b] - tm), np.minimum(BN, N1p[b] - tn),
                                    nR[b])

        def build3(BM, BN):
            gi, tm, tn = _tiles_of(N3p[pc_ap], nLp[pc_ap], BM, BN)
            ap, k0 = pc_ap[gi], pc_k0[gi]
            return _tile_table_cols(self.offT2[ap] + k0 * N3p[ap] + tm, self.offLP[ap] + k0 * nLp[ap] + tn,
                                    pc_off[gi] + tm * nLp[ap] + tn, N3p[ap], nLp[ap], nLp[ap],
                                    np.minimum(BM, N3p[ap] - tm), np.minimum(BN, nLp[ap] - tn), pc_k1[gi] - k0)
        self.var1, self.tiles1_host, self.tile1 = _choose_tiles(self.cplx, build1)
        self.var3, self.tiles3_host, self.tile3 = _choose_tiles(self.cplx, build3)

        # ---- gather tables that depend on the quantum numbers only ----
        one = np.ones(len(xb), dtype=np.int64)
        dense_off = (oL[xal] * d + xs) * Dr + oR[xb]
        x_off = self.offX[xb] + xm
        # X[b][j, m_off + i] = a[oL+i, s, oR+j]
        pack_a = _gather_tables(x_off, Mp[xb], nR[xb], nL[xal], one, dense_off, one, one * (d * Dr))
        # out[oL+i, s, oR+j] = X[b][j, m_off + i]
        unpack = _gather_tables(dense_off, one * (d * Dr), nL[xal], nR[xb], one, x_off, one, Mp[xb])
        one = np.ones(len(rb_b), dtype=np.int64)
        pack_r = _gather_tables(self.offRB[rb_b] + rb_n, N1p[rb_b], nR[rb_b], nR[rb_g], one,
                                (oR[rb_b] * cr + rb_K) * Dr + oR[rb_g], one * (cr * Dr), one)
        one = np.ones(len(tr_ap), dtype=np.int64)
        pack_l = _gather_tables(self.offLP[tr_ap] + tr_kk * nLp[tr_ap], nLp[tr_ap], nL[tr_al], nL[tr_ap], one,
                                (oL[tr_al] * cl + tr_k) * Dl + oL[tr_ap], one * (cl * Dl), one)
        # X chunk (a', s') of group g:  X[g][j', m_off + i'] = sum_p O_(a',p)[n_off + j', i']
        cnt = npc[tc_ap]
        ci = np.repeat(np.arange(len(tc_ap)), cnt)
        pidx = pc_first[tc_ap[ci]] + (np.arange(int(cnt.sum())) - excl(cnt)[ci])
        t_ap = tc_ap[ci]
        repack = _gather_tables(self.offX[tc_g] + self.XM[tc_ap, tc_sp], Mp[tc_g], nR[tc_g], nL[tc_ap], cnt,
                                pc_off[pidx] + tc_n[ci] * nLp[t_ap], nLp[t_ap], np.ones(len(ci), dtype=np.int64))
        self.g_host = {"pack_a": pack_a, "unpack": unpack, "pack_r": pack_r, "pack_l": pack_l, "repack": repack}

    # -------------------------------------------------------------------------------------------------
    def flop_counts(self):
        """`exact`: flops of the non-zero sector blocks (two GEMM steps); `visited`: what the grouped GEMMs execute
        (whole 128 x 64 tiles, contraction length rounded up to the DMMA k = 4)."""
        per = 8.0 if self.cplx else 2.0
        _, _, nL = self.L
        _, _, nR = self.R
        exact = per * (float(np.sum(self.M * self.N1 * nR)) + float(np.sum(self.N3 * nL * self.K3)))
        vis = 0.0
        for tab, (BM, BN) in ((self.tiles1_host, self.tile1), (self.tiles3_host, self.tile3)):
            vis += per * BM * BN * float(np.sum((tab["k"].astype(np.int64) + 3) // 4 * 4))
        return {"exact": exact, "visited": vis, "tile_variants": (self.var1, self.var3)}

    def _w_tables(self, w):
        """Block-gather tables of the W step for the MPO tensor `w` (device tensor; cached per tensor version)."""
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype)
        hit = self._wcache.get(key)
        if hit is not None:
            return hit[0]
        tables = _DeviceGather(self.w_tables_host(w.detach().cpu().numpy()), w.device)
        if len(self._wcache) > 8:
            self._wcache.clear()
        self._wcache[key] = (tables, w)
        return tables

    def w_tables_host(self, wh, transposed=False):
        """Host tables (chunks, terms, work) of the W step for the MPO tensor values `wh` (NumPy): one chunk per
        (a', k, s') block of T2, one term per non-zero MPO entry w[k, s', s, K] that connects it to a block of T1.
        `transposed`: source and destination are the transposed intermediates T1^T (N1 x M per group) and T2^T
        (N3 x K3 per group) of the environment update (`PackedEnvPlan`)."""
        Dl, d, Dr, cl, cr = self.dims
        assert wh.shape == (cl, d, d, cr)
        _, _, nL = self.L
        _, _, nR = self.R
        tr_ap, tr_k, tr_al, tr_kk = self.tr
        nrow = len(tr_ap)
        # chunks: (row entry (a', k, alpha), s') with a column block in group a'
        c_row, c_sp = np.repeat(np.arange(nrow), d), np.tile(np.arange(d), nrow)
        c_g = self.TCG[tr_ap[c_row], c_sp]
        ok = c_g >= 0
        c_row, c_sp, c_g = c_row[ok], c_sp[ok], c_g[ok]
        chunk_id = np.full((nrow, d), -1, dtype=np.int64)
        chunk_id[c_row, c_sp] = np.arange(len(c_row))
        # terms: non-zero entries (k, s', s, K) x row entries with the same k
        nk, nsp, ns, nK = np.nonzero(wh)
        vals = wh[nk, nsp, ns, nK]
        rows_of_k = [np.flatnonzero(tr_k == k) for k in range(cl)]
        cnt = np.array([len(rows_of_k[k]) for k in nk], dtype=np.int64)
        e = np.repeat(np.arange(len(nk)), cnt)                       # entry index of every candidate term
        rw = np.concatenate([rows_of_k[k] for k in nk]) if len(nk) else np.zeros(0, dtype=np.int64)
        cid = chunk_id[rw, nsp[e]]
        b = self.XB[tr_al[rw], ns[e]]
        ok = (cid >= 0) & (b >= 0)
        # an MPO entry that violates charge conservation would meet structural zeros: it has no term
        ok &= self.RBG[np.maximum(b, 0), nK[e]] == np.where(cid >= 0, c_g[np.maximum(cid, 0)], -2)
        e, rw, cid, b = e[ok], rw[ok], cid[ok], b[ok]
        order = np.argsort(cid, kind="stable")
        e, rw, cid, b = e[order], rw[order], cid[order], b[order]
        v = vals[e]
        ap = tr_ap[c_row]
        nterms = np.bincount(cid, minlength=len(c_row))
        one = np.ones(len(b), dtype=np.int64)
        if transposed:
            # T2T[a'][(s', j'), (k, alpha, i)] = sum w T1T[b][(K, j'), (alpha, s, i)]
            src = self.offT1[b] + self.RBN[b, nK[e]] * self.Mp[b] + self.XM[tr_al[rw], ns[e]]
            return _gather_tables(self.offT2T[ap] + self.TCN[ap, c_sp] * self.K3p[ap] + tr_kk[c_row], self.K3p[ap],
                                  nR[c_g], nL[tr_al[c_row]], nterms, src, self.Mp[b], one, np.real(v), np.imag(v))
        src = self.offT1[b] + self.XM[tr_al[rw], ns[e]] * self.N1p[b] + self.RBN[b, nK[e]]
        return _gather_tables(self.offT2[ap] + tr_kk[c_row] * self.N3p[ap] + self.TCN[ap, c_sp], self.N3p[ap],
                              nL[tr_al[c_row]], nR[c_g], nterms, src, self.N1p[b], one, np.real(v), np.imag(v))

    def device_tables(self, device):
        key = device.index
        hit = self._dev.get(key)
        if hit is None:
            names = list(self.g_host)
            arrays = [self.tiles1_host, self.tiles3_host] + [t for n in names for t in self.g_host[n]]
            buf, offs = _pack_segments(arrays)
            dbuf = torch.from_numpy(buf).to(device)            # ONE upload for all tables of the plan
            hit = {"tiles1": dbuf[offs[0]:], "tiles3": dbuf[offs[1]:]}
            for i, name in enumerate(names):
                hit[name] = _DeviceGather(self.g_host[name], device, [dbuf[o:] for o in offs[2 + 3 * i: 5 + 3 * i]])
            self._dev[key] = hit
        return hit

    def bind(self, w, l, r):
        """Operator for one local problem: packs `l` and `r` (once; they are fixed during a Lanczos run)."""
        return PackedHeffOperator(self, w, l, r)


class PackedHeffOperator:
    """`H_eff` of one local problem acting on packed vectors (see module docstring).  `pack` / `unpack` convert
    between the dense site tensor and the packed space; `__call__` is the matvec X -> X that the Krylov drivers
    iterate."""

    def __init__(self, plan, w, l, r):
        assert plan.supported
        self.plan = plan
        Dl, d, Dr, cl, cr = plan.dims
        assert tuple(w.shape) == (cl, d, d, cr) and tuple(l.shape) == (Dl, cl, Dl) and tuple(r.shape) == (Dr, cr, Dr)
        self.lib = _lib.load()
        self.dt = _lib.PTB_COMPLEX128 if plan.cplx else _lib.PTB_REAL64
        dtype = dev.C128 if plan.cplx else dev.F64
        device = l.device
        self.device = device
        self.tabs = plan.device_tables(device)
        self.wtab = plan._w_tables(dev.dense(w))
        stream = dev.stream_ptr(device)
        l = dev.as_dtype(l, plan.cplx); r = dev.as_dtype(r, plan.cplx)
        self.rb = torch.empty(max(plan.nRB, 1), dtype=dtype, device=device)
        self.lp = torch.empty(max(plan.nLP, 1), dtype=dtype, device=device)
        self.tabs["pack_r"].run(self.lib, self.dt, r, self.rb, stream)
        self.tabs["pack_l"].run(self.lib, self.dt, l, self.lp, stream)
        es = 16 if plan.cplx else 8
        ws = dev.workspace((plan.nT1 + plan.nT2 + plan.nO + 3) * es + 64, device, tag="packed")
        o1 = (plan.nT1 * es + 15) // 16 * 16
        o2 = o1 + (plan.nT2 * es + 15) // 16 * 16
        self.t1 = ws[:max(plan.nT1, 1) * es].view(dtype)
        self.t2 = ws[o1:o1 + max(plan.nT2, 1) * es].view(dtype)
        self.o = ws[o2:o2 + max(plan.nO, 1) * es].view(dtype)
        self.n = plan.nX
        # the breakdown threshold of krylov.py:44 is 100 n eps with n the length of the DENSE vector
        self.ptb_n_threshold = Dl * d * Dr

    def pack(self, a):
        plan = self.plan
        a = dev.as_dtype(a, plan.cplx)
        alloc = torch.empty if plan.cplx else torch.zeros          # float64: padding entries must be zero
        x = alloc(max(plan.nX, 1), dtype=a.dtype, device=a.device)
        self.tabs["pack_a"].run(self.lib, self.dt, a, x, dev.stream_ptr(a.device))
        return x[:plan.nX]

    def unpack(self, x):
        Dl, d, Dr, _, _ = self.plan.dims
        x = dev.dense(x)
        out = torch.zeros((Dl, d, Dr), dtype=x.dtype, device=x.device)
        self.tabs["unpack"].run(self.lib, self.dt, x, out, dev.stream_ptr(x.device))
        return out

    def __call__(self, x):
        plan, lib, dt = self.plan, self.lib, self.dt
        assert x.shape[0] == plan.nX
        x = dev.as_dtype(x, plan.cplx)
        stream = dev.stream_ptr(x.device)
        y = (torch.empty if plan.cplx else torch.zeros)(max(plan.nX, 1), dtype=x.dtype, device=x.device)
        n1, n3 = len(plan.tiles1_host), len(plan.tiles3_host)
        if n1:
            _lib.check(lib.ptb_gemm_grouped_v(dt, plan.var1, x.data_ptr(), self.rb.data_ptr(), self.t1.data_ptr(),
                                              self.tabs["tiles1"].data_ptr(), n1, stream), "ptb_gemm_grouped(1)")
        self.wtab.run(lib, dt, self.t1, self.t2, stream)
        if n3:
            _lib.check(lib.ptb_gemm_grouped_v(dt, plan.var3, self.t2.data_ptr(), self.lp.data_ptr(), self.o.data_ptr(),
                                              self.tabs["tiles3"].data_ptr(), n3, stream), "ptb_gemm_grouped(3)")
        self.tabs["repack"].run(lib, dt, self.o, y, stream)
        return y[:plan.nX]

    def apply_dense(self, a):
        """Dense in, dense out (tests / one-off calls): pack, one matvec, unpack."""
        return self.unpack(self(self.pack(a)))


class PackedEnvPlan:
    """
    `contraction_operator_step_right / _left` (pytenet/chain_ops.py:16-57, 60-99; bra = ket, as in the sweeps) on
    the sector-packed layouts -- the same two kernels as the matvec, no structural zero stored or multiplied:

        step_right(a, w, r)[i,k,i'] = sum a[i,s,j] r[j,K,j'] w[k,s',s,K] conj(a[i',s',j'])

      1  X = pack(a), RB = pack(r)                            (block gathers, as for the matvec)
      2  T1^T_b (N1 x M) = RB_b^T X_b                         grouped GEMM, contraction over the right sector b
      3  T2^T_a'[(s',j'), (k,alpha,i)] = sum w T1^T[..]       block gather, coefficients = MPO entries
      4  BT_a'[(s',j'), i'] = conj(a[i',s',j'])               block gather from X with the conjugate flag
      5  O_a' (K3 x n_a') = T2^T_a'^T BT_a'                   grouped GEMM, contraction over (s',j') split into pieces
      6  r_next[i,k,i'] = sum of the pieces                   block gather into the dense (zero-initialised) block

    `side="left"` is the same computation on the mirrored tensors a~[j,s,i] = a[i,s,j], w~[K,s',s,k] = w[k,s',s,K],
    r~ = l with quantum numbers (-qr, qs, -ql, -qwr, -qwl): mirroring only changes the strides in the pack table of
    `a` and the (host) index order of the MPO entries, nothing is transposed in memory.
    """

    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True, side="right"):
        assert side in ("left", "right")
        self.side = side
        self.cplx = bool(cplx)
        ql, qs, qr, qwl, qwr = (np.asarray(q, dtype=np.int64) for q in (ql, qs, qr, qwl, qwr))
        self.dims = (len(ql), len(qs), len(qr), len(qwl), len(qwr))
        if side == "left":
            ql, qr, qwl, qwr = -qr, -ql, -qwr, -qwl
        base = self.base = PackedHeffPlan(ql, qs, qr, qwl, qwr, cplx=cplx)
        self.supported = base.supported
        self._dev = {}
        self._wcache = {}
        if not self.supported:
            return
        Dl, d, Dr, cl, cr = base.dims              # of the (possibly mirrored) problem
        _, oL, nL = base.L
        _, oR, nR = base.R
        excl = lambda v: np.cumsum(v) - v          # noqa: E731
        Mp, N1p, N3, K3, K3p, nLp = base.Mp, base.N1p, base.N3, base.K3, base.K3p, base.nLp
        # pack of `a`: X[b][j, m_off + i] = a~[oL+i, s, oR+j]; strides (sx, ss, sy) of a~ in the memory of `a`
        sx, ss, sy = (d * Dr, Dr, 1) if side == "right" else (1, Dl, d * Dl)
        xs, xal, xb, xm = base.xc
        one = np.ones(len(xb), dtype=np.int64)
        pack_a = _gather_tables(base.offX[xb] + xm, Mp[xb], nR[xb], nL[xal], one,
                                oL[xal] * sx + xs * ss + oR[xb] * sy, one * sy, one * sx)
        # T1^T_b = RB_b^T X_b
        live1 = np.flatnonzero((base.M > 0) & (base.N1 > 0))

        def build1(BM, BN):
            gi, tm, tn = _tiles_of(N1p[live1], Mp[live1], BM, BN)
            b = live1[gi]
            return _tile_table_cols(base.offRB[b] + tm, base.offX[b] + tn, base.offT1[b] + tm * Mp[b] + tn,
                                    N1p[b], Mp[b], Mp[b], np.minimum(BM, N1p[b] - tm), np.minimum(BN, Mp[b] - tn),
                                    nR[b])
        self.var1, self.tiles1, _ = _choose_tiles(self.cplx, build1)
        # BT_a'[(n_off(s') + j'), i'] = conj(X[g][j', m_off(a', s') + i'])
        tc_ap, tc_sp, tc_g, tc_n = base.tc
        self.offBT = excl(N3 * nLp)
        self.nBT = int(np.sum(N3 * nLp))
        one = np.ones(len(tc_ap), dtype=np.int64)
        pack_bt = _gather_tables(self.offBT[tc_ap] + tc_n * nLp[tc_ap], nLp[tc_ap], nR[tc_g], nL[tc_ap], one,
                                 base.offX[tc_g] + base.XM[tc_ap, tc_sp], Mp[tc_g], one, conj=True)
        # O_(a',p) (K3 x n_a') = T2^T_a'[rows k0:k1]^T BT_a'[rows k0:k1]
        pc_ap, pc_k0, pc_k1, npc, pc_first = _k_pieces(N3, (K3 > 0) & (N3 > 0))
        pc_size = K3p[pc_ap] * nLp[pc_ap]
        pc_off = excl(pc_size)
        self.nO = int(pc_size.sum())

        def build3(BM, BN):
            gi, tm, tn = _tiles_of(K3p[pc_ap], nLp[pc_ap], BM, BN)
            ap, k0 = pc_ap[gi], pc_k0[gi]
            return _tile_table_cols(base.offT2T[ap] + k0 * K3p[ap] + tm, self.offBT[ap] + k0 * nLp[ap] + tn,
                                    pc_off[gi] + tm * nLp[ap] + tn, K3p[ap], nLp[ap], nLp[ap],
                                    np.minimum(BM, K3p[ap] - tm), np.minimum(BN, nLp[ap] - tn), pc_k1[gi] - k0)
        self.var3, self.tiles3, _ = _choose_tiles(self.cplx, build3)
        # out[oL[alpha]+i, k, oL[a']+i'] = sum_p O_(a',p)[kk + i, i']
        tr_ap, tr_k, tr_al, tr_kk = base.tr
        cnt = npc[tr_ap]
        ci = np.repeat(np.arange(len(tr_ap)), cnt)
        pidx = pc_first[tr_ap[ci]] + (np.arange(int(cnt.sum())) - excl(cnt)[ci])
        t_ap = tr_ap[ci]
        one = np.ones(len(tr_ap), dtype=np.int64)
        unpack = _gather_tables((oL[tr_al] * cl + tr_k) * Dl + oL[tr_ap], one * (cl * Dl), nL[tr_al], nL[tr_ap], cnt,
                                pc_off[pidx] + tr_kk[ci] * nLp[t_ap], nLp[t_ap], np.ones(len(ci), dtype=np.int64))
        self.g_host = {"pack_a": pack_a, "pack_r": base.g_host["pack_r"], "pack_bt": pack_bt, "unpack": unpack}

    def w_tables_host(self, wh):
        """Tables of the W step (T1^T -> T2^T) for the MPO tensor values `wh` (as passed to the contraction)."""
        if self.side == "left":
            wh = np.ascontiguousarray(np.transpose(wh, (3, 1, 2, 0)))
        return self.base.w_tables_host(wh, transposed=True)

    def _w_tables(self, w):
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype)
        hit = self._wcache.get(key)
        if hit is not None:
            return hit[0]
        tables = _DeviceGather(self.w_tables_host(w.detach().cpu().numpy()), w.device)
        if len(self._wcache) > 8:
            self._wcache.clear()
        self._wcache[key] = (tables, w)
        return tables

    def device_tables(self, device):
        key = device.index
        hit = self._dev.get(key)
        if hit is None:
            names = list(self.g_host)
            arrays = [self.tiles1, self.tiles3] + [t for n in names for t in self.g_host[n]]
            buf, offs = _pack_segments(arrays)
            dbuf = torch.from_numpy(buf).to(device)
            hit = {"tiles1": dbuf[offs[0]:], "tiles3": dbuf[offs[1]:]}
            for i, name in enumerate(names):
                hit[name] = _DeviceGather(self.g_host[name], device, [dbuf[o:] for o in offs[2 + 3 * i: 5 + 3 * i]])
            self._dev[key] = hit
        return hit

    def apply(self, a, w, env):
        """The next environment block: `step_right(a, a, w, env)` resp. `step_left(a, a, w, env)` (dense in / out)."""
        base = self.base
        Dl, d, Dr, cl, cr = base.dims
        da, dd, db, dcl, dcr = self.dims
        assert tuple(a.shape) == (da, dd, db) and tuple(w.shape) == (dcl, dd, dd, dcr)
        assert tuple(env.shape) == (Dr, cr, Dr)
        assert dev.any_complex(a, w, env) == self.cplx
        lib = _lib.load()
        dt = _lib.PTB_COMPLEX128 if self.cplx else _lib.PTB_REAL64
        dtype = dev.C128 if self.cplx else dev.F64
        device = a.device
        tabs = self.device_tables(device)
        wtab = self._w_tables(dev.dense(w))
        stream = dev.stream_ptr(device)
        a = dev.as_dtype(a, self.cplx); env = dev.as_dtype(env, self.cplx)
        es = 16 if self.cplx else 8
        sizes = [base.nX, base.nRB, base.nT1, base.nT2T, self.nBT, self.nO]
        offs, pos = [], 0
        for n in sizes:
            offs.append(pos)
            pos += (max(n, 1) * es + 63) // 64 * 64
        ws = dev.workspace(pos, device, tag="packed_env")
        x, rb, t1, t2, bt, o = (ws[o_:o_ + max(n, 1) * es].view(dtype) for o_, n in zip(offs, sizes))
        if not self.cplx:
            x.zero_()                                            # padding columns take part in step 2
        tabs["pack_a"].run(lib, dt, a, x, stream)
        tabs["pack_r"].run(lib, dt, env, rb, stream)
        n1, n3 = len(self.tiles1), len(self.tiles3)
        if n1:
            _lib.check(lib.ptb_gemm_grouped_v(dt, self.var1, rb.data_ptr(), x.data_ptr(), t1.data_ptr(),
                                              tabs["tiles1"].data_ptr(), n1, stream), "ptb_gemm_grouped(env 1)")
        wtab.run(lib, dt, t1, t2, stream)
        tabs["pack_bt"].run(lib, dt, x, bt, stream)
        if n3:
            _lib.check(lib.ptb_gemm_grouped_v(dt, self.var3, t2.data_ptr(), bt.data_ptr(), o.data_ptr(),
                                              tabs["tiles3"].data_ptr(), n3, stream), "ptb_gemm_grouped(env 3)")
        out = torch.zeros((Dl, cl, Dl), dtype=dtype, device=device)
        tabs["unpack"].run(lib, dt, o, out, stream)
        return out
