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

Sector layouts are untouched (qbonds / retained indices stay bit-exact).  Bonds that are not grouped by sector
(e.g. `MPS.construct_random` output before the first orthonormalisation) are not supported here -- `supported`
is False and the callers use the banded path of sectors.py.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import _device as dev

__all__ = ["PackedHeffPlan", "PackedHeffOperator"]

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


class _Gather:
    """Host tables of one block-gather launch (ptb_block_gather): chunks, terms, work items."""

    def __init__(self):
        self.chunks = []      # (dst_off, dst_ld, rows, cols, term_begin, term_end)
        self.terms = []       # (src_off, src_rs, src_cs, coef_re, coef_im)

    def chunk(self, dst_off, dst_ld, rows, cols, terms):
        if rows <= 0 or cols <= 0:
            return
        t0 = len(self.terms)
        self.terms.extend(terms)
        self.chunks.append((dst_off, dst_ld, rows, cols, t0, len(self.terms)))

    def finish(self):
        ch = np.zeros(len(self.chunks), dtype=[("dst_off", "<i8"), ("dst_ld", "<i4"), ("rows", "<i4"), ("cols", "<i4"),
                                               ("t0", "<i4"), ("t1", "<i4"), ("res", "<i4")])
        for i, c in enumerate(self.chunks):
            ch[i] = c + (0,)
        tm = np.zeros(max(len(self.terms), 1), dtype=[("src_off", "<i8"), ("rs", "<i4"), ("cs", "<i4"),
                                                       ("re", "<f8"), ("im", "<f8")])
        for i, t in enumerate(self.terms):
            tm[i] = t
        work = []
        for i, c in enumerate(self.chunks):
            rows, cols = c[2], c[3]
            per = max(1, _GATHER_ELEMS // cols)
            for r0 in range(0, rows, per):
                work.append((i, r0, min(per, rows - r0), 0))
        wk = np.asarray(work, dtype=np.int32).reshape(-1, 4)
        assert ch.dtype.itemsize == 32 and tm.dtype.itemsize == 32
        return ch, tm, wk


class _DeviceGather:
    def __init__(self, tables, device):
        ch, tm, wk = tables
        self.nwork = len(wk)
        self.ch = torch.from_numpy(ch.view(np.uint8).reshape(-1)).to(device) if len(ch) else None
        self.tm = torch.from_numpy(tm.view(np.uint8).reshape(-1)).to(device)
        self.wk = torch.from_numpy(np.ascontiguousarray(wk)).to(device) if len(wk) else None

    def run(self, lib, dt, src, dst, stream):
        if self.nwork == 0:
            return
        st = lib.ptb_block_gather(dt, src.data_ptr(), dst.data_ptr(), self.ch.data_ptr(), self.tm.data_ptr(),
                                  self.wk.data_ptr(), self.nwork, stream)
        _lib.check(st, "ptb_block_gather")


def _tile_table(tiles):
    """(n, 64-byte) device-ready table from (a_off, b_off, c_off, lda, ldb, ldc, m, n, k) tuples, sorted by
    decreasing k (longest tiles first over the persistent CTAs)."""
    tiles = sorted(tiles, key=lambda t: -t[8])
    tab = np.zeros(len(tiles), dtype=[("a", "<i8"), ("b", "<i8"), ("c", "<i8"), ("lda", "<i4"), ("ldb", "<i4"),
                                      ("ldc", "<i4"), ("m", "<i4"), ("n", "<i4"), ("k", "<i4"), ("acc", "<i4"),
                                      ("res", "<i4", (3,))])
    for i, t in enumerate(tiles):
        tab[i] = t + (0, (0, 0, 0))
    assert tab.dtype.itemsize == 64
    return tab


class PackedHeffPlan:
    """
    Everything about the sector-packed matvec that depends on the quantum numbers only: sector lists, packed
    layouts, GEMM tile tables and the gather tables that pack `a`, `l`, `r` and unpack the result.
    Arguments as `sectors.HeffSectorPlan` (bra bonds = ket bonds, as in the sweeps).
    """

    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True):
        self.cplx = bool(cplx)
        self.ql = np.asarray(ql, dtype=np.int64); self.qr = np.asarray(qr, dtype=np.int64)
        self.qs = np.asarray(qs, dtype=np.int64)
        self.qwl = np.asarray(qwl, dtype=np.int64); self.qwr = np.asarray(qwr, dtype=np.int64)
        Dl, d, Dr, cl, cr = len(self.ql), len(self.qs), len(self.qr), len(self.qwl), len(self.qwr)
        self.dims = (Dl, d, Dr, cl, cr)
        L, R = _runs(self.ql), _runs(self.qr)
        # complex128 only: every row of a packed operand is then 16-byte granular for the bulk copies
        self.supported = self.cplx and L is not None and R is not None
        self._dev = {}
        self._wcache = {}
        if not self.supported:
            return
        lib = _lib.load()
        bm, bn, bk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(lib.ptb_gemm_tile_shape(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, ctypes.byref(bm),
                                           ctypes.byref(bn), ctypes.byref(bk)), "ptb_gemm_tile_shape")
        BM, BN = bm.value, bn.value
        self.tile = (BM, BN, bk.value)
        qL, oL, nL = L
        qR, oR, nR = R
        self.L, self.R = L, R
        idxL = {int(v): i for i, v in enumerate(qL)}
        idxR = {int(v): i for i, v in enumerate(qR)}
        nLs, nRs = len(qL), len(qR)

        # ---- packed vector space X: group b (right sector), columns = stacked (alpha, s) blocks ----
        self.x_chunk = {}                       # (alpha, s) -> (b, m_off)
        M = np.zeros(nRs, dtype=np.int64)
        for s in range(d):
            for al in range(nLs):
                b = idxR.get(int(qL[al] + self.qs[s]))
                if b is not None:
                    self.x_chunk[(al, s)] = (b, int(M[b]))
                    M[b] += nL[al]
        self.M = M
        self.offX = np.concatenate([[0], np.cumsum(nR * M)])[:-1]
        self.nX = int(np.sum(nR * M))

        # ---- step 1 operand RB: group b, columns = stacked (K, gamma(b, K)) ----
        self.rb_chunk = {}                      # (b, K) -> (gamma, n_off)
        N1 = np.zeros(nRs, dtype=np.int64)
        for b in range(nRs):
            for K in range(cr):
                g = idxR.get(int(qR[b] + self.qwr[K]))
                if g is not None:
                    self.rb_chunk[(b, K)] = (g, int(N1[b]))
                    N1[b] += nR[g]
        self.N1 = N1
        self.offRB = np.concatenate([[0], np.cumsum(nR * N1)])[:-1]
        self.nRB = int(np.sum(nR * N1))
        self.offT1 = np.concatenate([[0], np.cumsum(M * N1)])[:-1]
        self.nT1 = int(np.sum(M * N1))

        # ---- step 3: group a' (left sector): rows of T2 / LP = stacked (alpha, k), columns of T2 = stacked (s', gamma') ----
        self.t2_row = {}                        # (a', alpha, k) -> kk_off
        self.t2_col = {}                        # (a', s') -> (gamma', n_off)
        K3 = np.zeros(nLs, dtype=np.int64)
        N3 = np.zeros(nLs, dtype=np.int64)
        for ap in range(nLs):
            for k in range(cl):
                al = idxL.get(int(qL[ap] - self.qwl[k]))
                if al is not None:
                    self.t2_row[(ap, al, k)] = int(K3[ap])
                    K3[ap] += nL[al]
            for sp in range(d):
                g = idxR.get(int(qL[ap] + self.qs[sp]))
                if g is not None:
                    self.t2_col[(ap, sp)] = (g, int(N3[ap]))
                    N3[ap] += nR[g]
        self.K3, self.N3 = K3, N3
        self.offT2 = np.concatenate([[0], np.cumsum(K3 * N3)])[:-1]
        self.nT2 = int(np.sum(K3 * N3))
        self.offLP = np.concatenate([[0], np.cumsum(K3 * nL)])[:-1]
        self.nLP = int(np.sum(K3 * nL))
        # K-split of the step-3 groups: piece p of group a' covers stacked rows [kb[p], kb[p+1]) and writes its own
        # partial product O_(a',p); the repack gather sums the pieces (fixed order, deterministic)
        self.k_pieces = []
        o_sizes = []
        for ap in range(nLs):
            kk = int(K3[ap])
            if kk == 0 or N3[ap] == 0:
                self.k_pieces.append([])
                continue
            npc = max(1, -(-kk // _KSPLIT))
            step = -(-kk // npc)
            step = -(-step // 16) * 16
            bounds = list(range(0, kk, step)) + [kk]
            self.k_pieces.append([(bounds[p], bounds[p + 1]) for p in range(len(bounds) - 1)])
            o_sizes.extend([int(N3[ap] * nL[ap])] * (len(bounds) - 1))
        self.offO = []
        pos = 0
        for ap in range(nLs):
            offs = []
            for _ in self.k_pieces[ap]:
                offs.append(pos)
                pos += int(N3[ap] * nL[ap])
            self.offO.append(offs)
        self.nO = pos

        # ---- GEMM tile tables ----
        t1, t3 = [], []
        for b in range(nRs):
            if M[b] == 0 or N1[b] == 0:
                continue
            for tm in range(0, int(M[b]), BM):
                for tn in range(0, int(N1[b]), BN):
                    t1.append((int(self.offX[b] + tm), int(self.offRB[b] + tn), int(self.offT1[b] + tm * N1[b] + tn),
                               int(M[b]), int(N1[b]), int(N1[b]), min(BM, int(M[b]) - tm), min(BN, int(N1[b]) - tn),
                               int(nR[b])))
        for ap in range(nLs):
            for p, (k0, k1) in enumerate(self.k_pieces[ap]):
                for tm in range(0, int(N3[ap]), BM):
                    for tn in range(0, int(nL[ap]), BN):
                        t3.append((int(self.offT2[ap] + k0 * N3[ap] + tm), int(self.offLP[ap] + k0 * nL[ap] + tn),
                                   int(self.offO[ap][p] + tm * nL[ap] + tn), int(N3[ap]), int(nL[ap]), int(nL[ap]),
                                   min(BM, int(N3[ap]) - tm), min(BN, int(nL[ap]) - tn), k1 - k0))
        self.tiles1_host = _tile_table(t1)
        self.tiles3_host = _tile_table(t3)

        # ---- gather tables that depend on the quantum numbers only ----
        one = (1.0, 0.0)
        g_pack_a, g_unpack, g_pack_r, g_pack_l, g_repack = _Gather(), _Gather(), _Gather(), _Gather(), _Gather()
        for (al, s), (b, m_off) in self.x_chunk.items():
            dense_off = int((oL[al] * d + s) * Dr + oR[b])
            x_off = int(self.offX[b] + m_off)
            # X[b][j, m_off + i] = a[oL+i, s, oR+j]
            g_pack_a.chunk(x_off, int(M[b]), int(nR[b]), int(nL[al]), [(dense_off, 1, d * Dr) + one])
            # out[oL+i, s, oR+j] = X[b][j, m_off + i]
            g_unpack.chunk(dense_off, d * Dr, int(nL[al]), int(nR[b]), [(x_off, 1, int(M[b])) + one])
        for (b, K), (g, n_off) in self.rb_chunk.items():
            g_pack_r.chunk(int(self.offRB[b] + n_off), int(N1[b]), int(nR[b]), int(nR[g]),
                           [(int((oR[b] * cr + K) * Dr + oR[g]), cr * Dr, 1) + one])
        for (ap, al, k), kk in self.t2_row.items():
            g_pack_l.chunk(int(self.offLP[ap] + kk * nL[ap]), int(nL[ap]), int(nL[al]), int(nL[ap]),
                           [(int((oL[al] * cl + k) * Dl + oL[ap]), cl * Dl, 1) + one])
        for (ap, sp), (g, n_off) in self.t2_col.items():
            # X chunk (a', s') of group g:  X[g][j', m_off + i'] = sum_p O_(a',p)[n_off + j', i']
            _, m_off = self.x_chunk[(ap, sp)]
            terms = [(int(off + n_off * nL[ap]), int(nL[ap]), 1) + one for off in self.offO[ap]]
            g_repack.chunk(int(self.offX[g] + m_off), int(M[g]), int(nR[g]), int(nL[ap]), terms)
        self.g_host = {"pack_a": g_pack_a.finish(), "unpack": g_unpack.finish(), "pack_r": g_pack_r.finish(),
                       "pack_l": g_pack_l.finish(), "repack": g_repack.finish()}

    # -------------------------------------------------------------------------------------------------
    def flop_counts(self):
        """`exact`: flops of the non-zero sector blocks (two GEMM steps); `visited`: what the grouped GEMMs execute
        (whole 128 x 64 tiles, contraction length rounded up to the DMMA k = 4)."""
        BM, BN, _ = self.tile
        per = 8.0 if self.cplx else 2.0
        _, _, nL = self.L
        _, _, nR = self.R
        exact = per * (float(np.sum(self.M * self.N1 * nR)) + float(np.sum(self.N3 * nL * self.K3)))
        vis = 0.0
        for tab in (self.tiles1_host, self.tiles3_host):
            vis += per * BM * BN * float(np.sum((tab["k"].astype(np.int64) + 3) // 4 * 4))
        return {"exact": exact, "visited": vis}

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

    def w_tables_host(self, wh):
        """Host tables (chunks, terms, work) of the W step for the MPO tensor values `wh` (NumPy)."""
        Dl, d, Dr, cl, cr = self.dims
        assert wh.shape == (cl, d, d, cr)
        _, _, nL = self.L
        _, _, nR = self.R
        nz = {}
        for k, sp, s, K in zip(*np.nonzero(wh)):
            nz.setdefault((int(k), int(sp)), []).append((int(s), int(K), complex(wh[k, sp, s, K])))
        gw = _Gather()
        for (ap, al, k), kk in self.t2_row.items():
            for sp in range(d):
                col = self.t2_col.get((ap, sp))
                if col is None:
                    continue
                g, n_off = col
                terms = []
                for s, K, val in nz.get((k, sp), []):
                    src = self.x_chunk.get((al, s))
                    if src is None:
                        continue
                    b, m_off = src
                    rb = self.rb_chunk.get((b, K))
                    if rb is None or rb[0] != g:
                        continue            # an MPO entry that violates charge conservation would meet structural zeros
                    terms.append((int(self.offT1[b] + m_off * self.N1[b] + rb[1]), int(self.N1[b]), 1,
                                  val.real, val.imag))
                gw.chunk(int(self.offT2[ap] + kk * self.N3[ap] + n_off), int(self.N3[ap]), int(nL[al]), int(nR[g]),
                         terms)
        return gw.finish()

    def device_tables(self, device):
        key = device.index
        hit = self._dev.get(key)
        if hit is None:
            up = lambda t: torch.from_numpy(t.view(np.uint8).reshape(-1)).to(device)      # noqa: E731
            hit = {"tiles1": up(self.tiles1_host), "tiles3": up(self.tiles3_host)}
            for name, tabs in self.g_host.items():
                hit[name] = _DeviceGather(tabs, device)
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
        x = torch.empty(max(plan.nX, 1), dtype=a.dtype, device=a.device)
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
        y = torch.empty(max(plan.nX, 1), dtype=x.dtype, device=x.device)
        n1, n3 = len(plan.tiles1_host), len(plan.tiles3_host)
        if n1:
            _lib.check(lib.ptb_gemm_grouped(dt, x.data_ptr(), self.rb.data_ptr(), self.t1.data_ptr(),
                                            self.tabs["tiles1"].data_ptr(), n1, stream), "ptb_gemm_grouped(1)")
        self.wtab.run(lib, dt, self.t1, self.t2, stream)
        if n3:
            _lib.check(lib.ptb_gemm_grouped(dt, self.t2.data_ptr(), self.lp.data_ptr(), self.o.data_ptr(),
                                            self.tabs["tiles3"].data_ptr(), n3, stream), "ptb_gemm_grouped(3)")
        self.tabs["repack"].run(lib, dt, self.o, y, stream)
        return y[:plan.nX]

    def apply_dense(self, a):
        """Dense in, dense out (tests / one-off calls): pack, one matvec, unpack."""
        return self.unpack(self(self.pack(a)))
