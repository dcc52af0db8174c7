"""
Quantum-number sector plans for the effective-Hamiltonian matvec (BASELINE config 3: block-sparse
path; SURVEY.md headline 3 -- the reference contracts dense tensors that hold explicit zeros).

With additive quantum numbers every tensor on the path is block sparse:

    a[i,s,j]   != 0  only if  ql[i] + qs[s]  - qr[j]   == 0        (pytenet/mps.py:14-26)
    r[j,K,j']  != 0  only if  qr[j] + qwr[K] - qr'[j'] == 0        (pytenet/tdvp.py:61-63)
    l[i,k,i']  != 0  only if  ql[i] + qwl[k] - ql'[i'] == 0
    w[k,s',s,K]!= 0  only if  qwl[k] + qs[s'] - qs[s] - qwr[K] == 0 (pytenet/mpo.py:15-28)

After the first orthonormalisation the bond indices are grouped by sector (block_sparse_qr / svd emit
the sectors in ascending order, block_sparse_util.py:151-169), so the non-zero entries of every GEMM
operand form contiguous blocks.  The plan turns this into *device-side work lists*: the matvec is
issued as GEMMs batched over the physical index (so that one batch sees one sector shift) and every
output tile gets the range of k-tiles that can be non-zero; the kernel (`ptb_gemm_banded`) visits only
those and skips tiles whose range is empty.  Skipped terms are exact zeros in the dense contraction,
so results equal the dense path (and the reference) -- sector layouts are never changed.

Nothing here depends on the indices being sorted: unsorted bonds only make the ranges wider.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import _device as dev

__all__ = ["HeffSectorPlan", "EnvSectorPlan", "BondSectorPlan", "AbsorbSectorPlan", "tile_k_ranges"]

import os
import itertools
_SEGMENTED = os.environ.get("PYTENET_B200_SEGMENTED", "1") != "0"
_SKIP_EMPTY = os.environ.get("PYTENET_B200_SKIP_EMPTY", "1") != "0"
# tile schedules sorted by decreasing work (PYTENET_B200_TILE_ORDER=0 disables them, for A/B measurements)
_ORDERED = os.environ.get("PYTENET_B200_TILE_ORDER", "1") != "0"
_ORDERED_STEP3 = _ORDERED
_PLAN_IDS = itertools.count(1)

_EMPTY_LO = np.iinfo(np.int64).max
_EMPTY_HI = -1


def _tile_shape(cplx):
    lib = _lib.load()
    bm, bn, bk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.ptb_gemm_tile_shape(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64,
                                       ctypes.byref(bm), ctypes.byref(bn), ctypes.byref(bk)), "ptb_gemm_tile_shape")
    return bm.value, bn.value, bk.value


class KIndex:
    """First / last position of every value of the contraction index's quantum numbers `qk`, computed once per
    plan (every work list of a plan looks up the same `qk` dozens of times)."""

    def __init__(self, qk):
        qk = np.asarray(qk, dtype=np.int64)
        self.vals, inv = np.unique(qk, return_inverse=True)
        pos = np.arange(len(qk))
        self.first = np.full(len(self.vals), _EMPTY_LO, dtype=np.int64)
        self.last = np.full(len(self.vals), _EMPTY_HI, dtype=np.int64)
        np.minimum.at(self.first, inv, pos)
        np.maximum.at(self.last, inv, pos + 1)


def _support(need, qk):
    """For every entry of `need` the bounding interval [lo, hi) of positions k with qk[k] == need
    (empty -> (_EMPTY_LO, _EMPTY_HI)).  `qk` is an array or a prepared KIndex."""
    kx = qk if isinstance(qk, KIndex) else KIndex(qk)
    need = np.asarray(need, dtype=np.int64)
    if len(kx.vals) == 0:
        return np.full(len(need), _EMPTY_LO, dtype=np.int64), np.full(len(need), _EMPTY_HI, dtype=np.int64)
    idx_c = np.clip(np.searchsorted(kx.vals, need), 0, len(kx.vals) - 1)
    found = kx.vals[idx_c] == need
    lo = np.where(found, kx.first[idx_c], _EMPTY_LO)
    hi = np.where(found, kx.last[idx_c], _EMPTY_HI)
    return lo, hi


def _tile_reduce(lo, hi, tile):
    """Bounding interval per block of `tile` consecutive entries."""
    n = len(lo)
    starts = np.arange(0, n, tile)
    return np.minimum.reduceat(lo, starts), np.maximum.reduceat(hi, starts)


def shifted_tile_support(base, shifts, kx, tile):
    """Per batch entry b the bounding k interval of every block of `tile` consecutive indices whose required
    quantum number is base[i] + shifts[b]:  (lo, hi) arrays of shape (len(shifts), n_tiles).  Equal shifts are
    looked up once."""
    base = np.asarray(base, dtype=np.int64)
    shifts = np.asarray(shifts, dtype=np.int64).reshape(-1)
    ush, inv = np.unique(shifts, return_inverse=True)
    need = base[None, :] + ush[:, None]
    lo, hi = _support(need.reshape(-1), kx)
    lo = lo.reshape(need.shape); hi = hi.reshape(need.shape)
    starts = np.arange(0, need.shape[1], tile)
    return np.minimum.reduceat(lo, starts, axis=1)[inv], np.maximum.reduceat(hi, starts, axis=1)[inv]


def combine_tile_ranges(rlo, rhi, clo, chi, bk):
    """k-TILE ranges [lo, hi) from row-tile and column-tile supports (broadcast over leading axes):
    (..., tiles_m) x (..., tiles_n) -> (..., tiles_m, tiles_n, 2) int32."""
    lo = np.maximum(rlo[..., :, None], clo[..., None, :])
    hi = np.minimum(rhi[..., :, None], chi[..., None, :])
    empty = hi <= lo
    lo_t = np.where(empty, 0, lo // bk)
    hi_t = np.where(empty, 0, -(-hi // bk))
    return np.stack([lo_t, hi_t], axis=-1).astype(np.int32)


def tile_k_ranges_batched(need_rows, need_cols, qk, bm, bn, bk):
    """Batched form of tile_k_ranges: `need_rows` (B or 1, M) and `need_cols` (B or 1, N) -> (B, tiles_m, tiles_n, 2)
    int32 k-TILE ranges [lo, hi), one table per batch entry, all looked up in one vectorised pass."""
    kx = qk if isinstance(qk, KIndex) else KIndex(qk)
    need_rows = np.atleast_2d(np.asarray(need_rows, dtype=np.int64))
    need_cols = np.atleast_2d(np.asarray(need_cols, dtype=np.int64))

    def reduced(need, tile):
        lo, hi = _support(need.reshape(-1), kx)
        lo = lo.reshape(need.shape); hi = hi.reshape(need.shape)
        starts = np.arange(0, need.shape[1], tile)
        return np.minimum.reduceat(lo, starts, axis=1), np.maximum.reduceat(hi, starts, axis=1)

    rlo, rhi = reduced(need_rows, bm)
    clo, chi = reduced(need_cols, bn)
    return combine_tile_ranges(rlo, rhi, clo, chi, bk)


def tile_k_ranges(need_rows, need_cols, qk, bm, bn, bk):
    """(tiles_m, tiles_n, 2) int32 array of k-TILE ranges [lo, hi): row m of the output only receives
    contributions from k with qk[k] == need_rows[m], column n from k with qk[k] == need_cols[n]."""
    return tile_k_ranges_batched(need_rows, need_cols, qk, bm, bn, bk)[0]


def segment_tables(tabs):
    """Tables of a segmented GEMM from per-selector k-range tables: `tabs` is a list (one entry per selector) of
    (batch, tiles_m, tiles_n, 2) int arrays.  Returns host arrays (seg_ptr, segs, order): per output tile the
    non-empty (lo, hi, selector, 0) segments in selector order, and the tiles sorted by decreasing work."""
    nsel = len(tabs)
    allk = np.stack(tabs, axis=-2)                                # (batch, tiles_m, tiles_n, nsel, 2)
    flat = (allk[..., 1] > allk[..., 0]).reshape(-1, nsel)
    seg_ptr = np.concatenate([[0], np.cumsum(flat.sum(axis=1))]).astype(np.int32)
    tile_idx, sel_idx = np.nonzero(flat)
    lohi = allk.reshape(-1, nsel, 2)[tile_idx, sel_idx]
    segs = np.zeros((max(len(sel_idx), 1), 4), dtype=np.int32)
    segs[:len(sel_idx), 0:2] = lohi
    segs[:len(sel_idx), 2] = sel_idx
    work = np.zeros(flat.shape[0], dtype=np.int64)
    if len(sel_idx):
        np.add.at(work, tile_idx, (lohi[:, 1] - lohi[:, 0]).astype(np.int64))
    order = np.argsort(-work, kind="stable").astype(np.int32)
    return seg_ptr, segs, order


def banded_order(tab):
    """Tiles of a banded GEMM sorted by decreasing k-range length (stable); `tab` is (..., 2)."""
    work = (tab[..., 1] - tab[..., 0]).reshape(-1)
    return np.argsort(-np.maximum(work, 0), kind="stable").astype(np.int32)


class HeffSectorPlan:
    """
    Work lists for `out = L.W.A.R` on one site (or merged site pair) with quantum numbers.

    Args (host integer arrays): `ql`, `qr` ket bond quantum numbers of `a`; `qs_in`, `qs_out` physical
    quantum numbers of the MPO tensor's input / output leg; `qwl`, `qwr` MPO bond quantum numbers;
    `qlp`, `qrp` bra bond quantum numbers (equal to `ql`, `qr` in the sweeps).
    """

    def __init__(self, ql, qs_in, qr, qwl, qwr, qs_out=None, qlp=None, qrp=None, device=None, cplx=True):
        self.ql = np.asarray(ql, dtype=np.int64)
        self.qr = np.asarray(qr, dtype=np.int64)
        self.qs_in = np.asarray(qs_in, dtype=np.int64)
        self.qs_out = self.qs_in if qs_out is None else np.asarray(qs_out, dtype=np.int64)
        self.qwl = np.asarray(qwl, dtype=np.int64)
        self.qwr = np.asarray(qwr, dtype=np.int64)
        self.qlp = self.ql if qlp is None else np.asarray(qlp, dtype=np.int64)
        self.qrp = self.qr if qrp is None else np.asarray(qrp, dtype=np.int64)
        self.device = device
        self.cplx = cplx
        bm, bn, bk = _tile_shape(cplx)
        self.tile = (bm, bn, bk)
        Dl, d, Dr = len(self.ql), len(self.qs_in), len(self.qr)
        cl, cr, dout = len(self.qwl), len(self.qwr), len(self.qs_out)
        Dlp, Drp = len(self.qlp), len(self.qrp)
        self.dims = (Dl, d, Dr, cl, cr, dout, Dlp, Drp)
        # the banded kernel is the TMA engine: float64 operands need even extents (16-byte granularity);
        # complex128 always qualifies.  Otherwise apply() uses the dense device path.
        self.supported = cplx or (Dr % 2 == 0 and Drp % 2 == 0 and Dlp % 2 == 0)

        # step 1, batched over s:  t1[i, s, (K, j')] = sum_j a[i, s, j] r[j, (K, j')]
        #   row i (batch s) needs  qr[j] = ql[i] + qs[s];  column (K, j') needs  qr[j] = qr'[j'] - qwr[K]
        cols1 = (self.qrp[None, :] - self.qwr[:, None]).reshape(-1)
        kx_r, kx_l = KIndex(self.qr), KIndex(self.ql)
        r1 = shifted_tile_support(self.ql, self.qs_in, kx_r, bm)                          # (d, tiles_m)
        c1 = shifted_tile_support(cols1, [0], kx_r, bn)                                   # (1, tiles_n)
        self.tab1_host = np.ascontiguousarray(combine_tile_ranges(r1[0], r1[1], c1[0], c1[1], bk))

        # step 3, one launch per left MPO index k, batched over s':
        #   out[i', s', j'] += sum_i l[i, k, i'] t2[i, k, s', j']
        #   row i' needs  ql[i] = ql'[i'] - qwl[k];  column j' (batch s') needs  ql[i] = qr'[j'] - qs[s'] - qwl[k]
        r3 = shifted_tile_support(self.qlp, -self.qwl, kx_l, bm)                          # (cl, tiles_m)
        c3 = shifted_tile_support(self.qrp, -(self.qs_out[None, :] + self.qwl[:, None]), kx_l, bn)   # (cl*dout, tiles_n)
        all3 = combine_tile_ranges(r3[0][:, None, :], r3[1][:, None, :], c3[0].reshape(cl, dout, -1),
                                   c3[1].reshape(cl, dout, -1), bk)                         # (cl, dout, tm, tn, 2)
        self.tab3_host = [np.ascontiguousarray(all3[k]) for k in range(cl)]
        self.k_active = [bool(np.any(t[..., 1] > t[..., 0])) for t in self.tab3_host]
        self.tab1 = None
        self.tab3 = None
        # step 3 as ONE segmented launch (ptb_gemm_segmented): per output tile the list of (k-tile range, left MPO
        # index) pairs that can contribute; selector k adds k*Dlp / k*dout*Drp elements to the l / t2 base pointers
        self.seg_ptr_host, self.segs_host, self.order3_host = segment_tables(self.tab3_host)
        self.sel_off_host = np.stack([np.arange(cl) * Dlp, np.arange(cl) * dout * Drp], axis=1).astype(np.int64)
        self.seg3 = None
        # tile schedules: tiles sorted by decreasing number of k-tiles (stable), so that the round-robin assignment
        # of work units to the persistent CTAs is balanced although the per-tile work differs widely
        self.order1_host = banded_order(self.tab1_host)
        self.order = None
        # row-activity flags of the W step: for every (BM-row block of i) x (128-column block of j') which input rows
        # (s, K) of t1 can hold entries, i.e. overlap a t1 tile that step 1 writes.  Inactive parts of t1 / t2 are
        # never written -- they keep the zeros of the one-time initialisation of the workspace (see apply) and are
        # neither read nor written by the W kernel, so the dense-layout intermediates cost traffic only where the
        # quantum numbers allow entries.  (The output-row flags follow from these and W's pattern: _w_flags.)
        nonempty1 = self.tab1_host[..., 1] > self.tab1_host[..., 0]                        # (d, tiles_m, tiles_n1)
        ncb = -(-Drp // 128)
        act = np.zeros((nonempty1.shape[1], ncb, d, cr), dtype=bool)
        for K in range(cr):
            for jb in range(ncb):
                c0 = K * Drp + jb * 128
                c1 = K * Drp + min((jb + 1) * 128, Drp)
                act[:, jb, :, K] = np.any(nonempty1[:, :, c0 // bn:(c1 - 1) // bn + 1], axis=2).T
        self.in_active_host = act.reshape(nonempty1.shape[1], ncb, d * cr)
        self._flag_cache = {}
        self._uid = next(_PLAN_IDS)

        # bookkeeping for benchmarks: fraction of the dense k-tile visits that remain
        kt1 = -(-Dr // bk)
        kt3 = -(-Dl // bk)
        vis1 = float(np.sum(self.tab1_host[..., 1] - self.tab1_host[..., 0]))
        den1 = float(self.tab1_host[..., 0].size * kt1)
        vis3 = sum(float(np.sum(t[..., 1] - t[..., 0])) for t in self.tab3_host)
        den3 = float(cl * dout * (-(-Dlp // bm)) * (-(-Drp // bn)) * kt3)
        self.visit_fraction = (vis1 / max(den1, 1.0), vis3 / max(den3, 1.0))

    def flop_counts(self, nnz_w=None):
        """Floating-point operations of one apply(): `visited` = what the banded / segmented GEMMs execute
        (whole k-tiles of whole output tiles), `exact` = the non-zero sector blocks only (what a sector-packed
        contraction would execute), both for the two large GEMM steps; `w_step` = 2 * nnz(W) * 4 (complex x real)
        per (i, j') column of the W step when `nnz_w` is given."""
        Dl, d, Dr, cl, cr, dout, Dlp, Drp = self.dims
        bm, bn, bk = self.tile
        per_ktile = 8.0 * bm * bn * bk if self.cplx else 2.0 * bm * bn * bk
        vis1 = float(np.sum(self.tab1_host[..., 1] - self.tab1_host[..., 0])) * per_ktile
        vis3 = float(np.sum(self.segs_host[:, 1] - self.segs_host[:, 0])) * per_ktile
        per_mac = 8.0 if self.cplx else 2.0

        def count(q):
            vals, cnt = np.unique(q, return_counts=True)
            return dict(zip(vals.tolist(), cnt.tolist()))
        nl, nr, nlp, nrp = count(self.ql), count(self.qr), count(self.qlp), count(self.qrp)
        ex1 = 0.0
        for qi, ni in nl.items():                       # t1[i,s,K,j'] = sum_j a[i,s,j] r[j,K,j']
            for s in range(d):
                nj = nr.get(qi + int(self.qs_in[s]), 0)
                if nj:
                    for K in range(cr):
                        ex1 += ni * nj * nrp.get(qi + int(self.qs_in[s]) + int(self.qwr[K]), 0)
        ex3 = 0.0
        for qip, nip in nlp.items():                    # out[i',s',j'] = sum_{i,k} l[i,k,i'] t2[i,k,s',j']
            for k in range(cl):
                ni = nl.get(qip - int(self.qwl[k]), 0)
                if ni:
                    for sp in range(dout):
                        ex3 += nip * ni * nrp.get(qip + int(self.qs_out[sp]), 0)
        out = {"visited": vis1 + vis3, "exact": per_mac * (ex1 + ex3)}
        if nnz_w is not None:
            out["w_step"] = (4.0 if self.cplx else 2.0) * nnz_w * Dl * Drp
        return out

    def _w_flags(self, csr, key, device):
        """Device array of W-step row flags for the MPO tensor whose CSR form is `csr` (cached per tensor):
        input flags from the plan, output row m active iff some non-zero W[m, c] meets an active input row c."""
        hit = self._flag_cache.get(key)
        if hit is not None:
            return hit
        r_out = self.dims[3] * self.dims[5]
        pat = np.frombuffer(csr[3].pattern, dtype=np.int32)
        rowptr = pat[:r_out + 1].astype(np.int64)
        nnz = int(rowptr[-1])
        col = pat[r_out + 1:].astype(np.int64)[:nnz]
        ina = self.in_active_host                                                 # (nib, ncb, r_in)
        rows = np.repeat(np.arange(r_out), np.diff(rowptr))
        outa = np.zeros(ina.shape[:2] + (r_out,), dtype=bool)
        if nnz:
            np.logical_or.at(outa, (slice(None), slice(None), rows), ina[:, :, col])
        flags = np.ascontiguousarray(np.concatenate([ina, outa], axis=2).astype(np.uint8))
        dev_flags = torch.from_numpy(flags).to(device)
        if len(self._flag_cache) > 8:
            self._flag_cache.clear()
        self._flag_cache[key] = dev_flags
        return dev_flags

    def _upload(self, device):
        if self.tab1 is None or self.tab1.device != device:
            self.tab1 = torch.from_numpy(self.tab1_host).to(device)
            self.tab3 = [torch.from_numpy(t).to(device) if act else None
                         for t, act in zip(self.tab3_host, self.k_active)]
            self.seg3 = (torch.from_numpy(self.seg_ptr_host).to(device), torch.from_numpy(self.segs_host).to(device),
                         torch.from_numpy(self.sel_off_host).to(device))
            self._flag_cache = {}
            self.order = (torch.from_numpy(self.order1_host).to(device), torch.from_numpy(self.order3_host).to(device))

    @classmethod
    def for_site(cls, psi_qbond_l, qsite_in, psi_qbond_r, mpo_qbond_l, mpo_qbond_r, device=None, cplx=True):
        return cls(psi_qbond_l, qsite_in, psi_qbond_r, mpo_qbond_l, mpo_qbond_r, device=device, cplx=cplx)

    def trivial(self):
        """True when all quantum numbers vanish (nothing to skip: use the dense path)."""
        return not any(np.any(q) for q in (self.ql, self.qr, self.qs_in, self.qs_out, self.qwl, self.qwr,
                                           self.qlp, self.qrp))

    def apply(self, a, w, l, r, out=None):
        """Sector-banded `apply_local_hamiltonian(a, w, l, r)` on complex128 (or float64) device tensors."""
        lib = _lib.load()
        Dl, d, Dr, cl, cr, dout, Dlp, Drp = self.dims
        cplx = dev.any_complex(a, l, r, w)
        assert cplx == self.cplx, "plan was built for a different dtype (tile shape differs)"
        if not self.supported:
            from .chain_ops import apply_local_hamiltonian
            return apply_local_hamiltonian(a, w, l, r, out=out)
        a = dev.as_dtype(a, cplx); l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
        w = dev.dense(w)
        assert tuple(a.shape) == (Dl, d, Dr) and tuple(w.shape) == (cl, dout, d, cr)
        assert tuple(l.shape) == (Dl, cl, Dlp) and tuple(r.shape) == (Dr, cr, Drp)
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        es = 16 if cplx else 8
        device = a.device
        self._upload(device)
        stream = dev.stream_ptr(device)
        n1 = Dl * d * cr * Drp
        n2 = Dl * cl * dout * Drp
        nbytes = (n1 + n2) * es + 32
        ws = dev.workspace(nbytes, device, tag="sectors")
        t1 = ws[:n1 * es].view(a.dtype).reshape(Dl, d, cr * Drp)
        off2 = (n1 * es + 15) // 16 * 16
        t2 = ws[off2:off2 + n2 * es].view(a.dtype).reshape(Dl, cl * dout, Drp)
        if out is None:
            out = torch.empty((Dlp, dout, Drp), dtype=a.dtype, device=device)
        csr = dev.w_csr(w) if (cplx or not w.dtype.is_complex) else None
        # Structural zeros are written ONCE per plan: the workspace is zeroed when this plan takes it over; after that
        # step 1 stores only tiles with a non-empty k range and the W kernel touches only active blocks, so every
        # other entry of t1 / t2 keeps its zero across the matvecs of a Lanczos run (k = 25 per plan in a sweep).
        lean = _SKIP_EMPTY and csr is not None
        if lean:
            key = (device.index, stream)
            owner = (self._uid, ws.data_ptr(), nbytes, csr[3].pattern)
            if dev.ws_owner.get(key) != owner:
                ws[:nbytes].zero_()
                dev.ws_owner[key] = owner
        else:
            dev.ws_owner.pop((device.index, stream), None)
        # (1) batched over s, banded in j
        tabs = _lib.SectorTables(self.tab1.data_ptr(), None, None, None, self.order[0].data_ptr() if _ORDERED else None)
        st = lib.ptb_gemm_sector(dt, 0, 0, 0, Dl, cr * Drp, Dr, a.data_ptr(), d * Dr, r.data_ptr(), cr * Drp,
                                 t1.data_ptr(), d * cr * Drp, d, Dr, 0, cr * Drp, 2 if lean else 0,
                                 ctypes.byref(tabs), stream)
        _lib.check(st, "ptb_gemm_sector(step 1)")
        # (2) W step batched over i (t1[i] is (d*cr) x Drp, t2[i] is (cl*dout) x Drp): sparse CSR kernel for
        #     the usual sparse MPO tensors, dense small GEMM otherwise
        if lean:
            rowptr, col, val, _ = csr
            flags = self._w_flags(csr, csr[3].pattern, device)
            st = lib.ptb_wapply_csr_masked(dt, int(w.dtype.is_complex), cl * dout, d * cr, Drp, rowptr.data_ptr(),
                                           col.data_ptr(), val.data_ptr(), t1.data_ptr(), t2.data_ptr(), Dl,
                                           flags.data_ptr(), self.tile[0], stream)
            _lib.check(st, "ptb_wapply_csr_masked")
        elif csr is not None:
            rowptr, col, val, _ = csr
            st = lib.ptb_wapply_csr(dt, int(w.dtype.is_complex), cl * dout, d * cr, Drp, rowptr.data_ptr(),
                                    col.data_ptr(), val.data_ptr(), t1.data_ptr(), t2.data_ptr(), Dl, stream)
            _lib.check(st, "ptb_wapply_csr")
        elif cplx and not w.dtype.is_complex:
            dev.gemm_strided(False, 0, 0, 0, cl * dout, 2 * Drp, d * cr, w, d * cr, torch.view_as_real(t1), 2 * Drp,
                             torch.view_as_real(t2), 2 * Drp, Dl, 0, 2 * d * cr * Drp, 2 * cl * dout * Drp)
        else:
            dev.gemm_strided(cplx, 0, 0, 0, cl * dout, Drp, d * cr, dev.as_dtype(w, cplx), d * cr, t1, Drp, t2, Drp,
                             Dl, 0, d * cr * Drp, cl * dout * Drp)
        # (3) one segmented launch, batched over s': every tile sums its (k-range, left MPO index) segments
        if _SEGMENTED:
            seg_ptr, segs, sel_off = self.seg3
            tabs = _lib.SectorTables(None, seg_ptr.data_ptr(), segs.data_ptr(), sel_off.data_ptr(),
                                     self.order[1].data_ptr() if _ORDERED_STEP3 else None)
            st = lib.ptb_gemm_sector(dt, 1, 0, 0, Dlp, Drp, Dl, l.data_ptr(), cl * Dlp, t2.data_ptr(), cl * dout * Drp,
                                     out.data_ptr(), dout * Drp, dout, 0, Drp, Drp, 0, ctypes.byref(tabs), stream)
            _lib.check(st, "ptb_gemm_sector(step 3)")
            return out
        # alternative (PYTENET_B200_SEGMENTED=0): one banded launch per left MPO index, accumulating into out
        first = True
        for k in range(cl):
            if not self.k_active[k]:
                continue
            a_ptr = l.data_ptr() + k * Dlp * es
            b_ptr = t2.data_ptr() + k * dout * Drp * es
            st = lib.ptb_gemm_banded(dt, 1, 0, 0, Dlp, Drp, Dl, a_ptr, cl * Dlp, b_ptr, cl * dout * Drp,
                                     out.data_ptr(), dout * Drp, dout, 0, Drp, Drp, 0 if first else 1,
                                     self.tab3[k].data_ptr(), stream)
            _lib.check(st, "ptb_gemm_banded(step 3)")
            first = False
        if first:
            out.zero_()
        return out


# ---------------------------------------------------------------------------------------------
# Environment updates and the zero-site (bond) contraction with the same work-list machinery
# ---------------------------------------------------------------------------------------------

def _banded(lib, dt, ta, tb, cj, m, n, k, a_ptr, lda, b_ptr, ldb, c_ptr, ldc, batch, sa, sb, sc, acc, tab, stream,
            what, order=None):
    tabs = _lib.SectorTables(tab.data_ptr(), None, None, None,
                             order.data_ptr() if (order is not None and _ORDERED) else None)
    st = lib.ptb_gemm_sector(dt, ta, tb, cj, m, n, k, a_ptr, lda, b_ptr, ldb, c_ptr, ldc, batch, sa, sb, sc,
                             int(acc), ctypes.byref(tabs), stream)
    _lib.check(st, what)


def _active(tabs):
    return bool(np.any(tabs[..., 1] > tabs[..., 0]))


_WT_CACHE = {}        # transposed MPO matrices, one per (tensor, version); values keep the source alive


class EnvSectorPlan:
    """
    Sector work lists for `contraction_operator_step_left / _right` (pytenet/chain_ops.py:60-99, 16-57) with
    a == b layouts (ket and bra bonds share quantum numbers, as in the sweeps).  Same conventions as
    `HeffSectorPlan`; `ql`, `qr` are the bond quantum numbers of the site tensor, `qwl`, `qwr` of the MPO tensor.
    """

    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True):
        self.ql = np.asarray(ql, dtype=np.int64); self.qr = np.asarray(qr, dtype=np.int64)
        self.qs = np.asarray(qs, dtype=np.int64)
        self.qwl = np.asarray(qwl, dtype=np.int64); self.qwr = np.asarray(qwr, dtype=np.int64)
        self.cplx = cplx
        bm, bn, bk = _tile_shape(cplx)
        Dl, d, Dr, cl, cr = len(self.ql), len(self.qs), len(self.qr), len(self.qwl), len(self.qwr)
        self.dims = (Dl, d, Dr, cl, cr)
        self.supported = cplx or (Dl % 2 == 0 and Dr % 2 == 0)
        ql_, qr_, qs_, qwl_, qwr_ = self.ql, self.qr, self.qs, self.qwl, self.qwr
        kxl, kxr = KIndex(ql_), KIndex(qr_)
        # ---- step_left ----
        # (L1) per k, batch s':  t[i,k,s',j'] = sum_i' l[i,k,i'] conj(b[i',s',j'])      K index i' (left bond)
        rl = shifted_tile_support(ql_, qwl_, kxl, bm)                       # rows need ql[i'] = ql[i] + qwl[k]
        cl_ = shifted_tile_support(qr_, -qs_, kxl, bn)                      # cols need ql[i'] = qr[j'] - qs[s']
        L1 = combine_tile_ranges(rl[0][:, None, :], rl[1][:, None, :], cl_[0][None, :, :], cl_[1][None, :, :], bk)
        self.L1 = [np.ascontiguousarray(L1[k]) for k in range(cl)]
        # (L3) per s:  l_next[j,(K,j')] += sum_i a[i,s,j] t2[i,s,K,j']                   K index i
        cols = (qr_[None, :] - qwr_[:, None]).reshape(-1)
        r3 = shifted_tile_support(qr_, -qs_, kxl, bm)
        c3 = shifted_tile_support(cols, -qs_, kxl, bn)
        L3 = combine_tile_ranges(r3[0], r3[1], c3[0], c3[1], bk)            # (d, tiles_m, tiles_n, 2)
        self.L3 = [np.ascontiguousarray(L3[s][None]) for s in range(d)]
        # ---- step_right ----
        # (R1) batch s:  t1[i,s,(K,j')] = sum_j a[i,s,j] r[j,(K,j')]                      K index j
        r1 = shifted_tile_support(ql_, qs_, kxr, bm)
        c1 = shifted_tile_support(cols, [0], kxr, bn)
        self.R1 = np.ascontiguousarray(combine_tile_ranges(r1[0], r1[1], c1[0], c1[1], bk))
        # (R3) per s', batch k:  r_next[i,k,i'] += sum_j' t2[i,k,s',j'] conj(b[i',s',j'])   K index j'
        rr = shifted_tile_support(ql_, (qs_[:, None] + qwl_[None, :]), kxr, bm)          # (d*cl, tiles_m)
        cr_ = shifted_tile_support(ql_, qs_, kxr, bn)                                    # (d, tiles_n)
        R3 = combine_tile_ranges(rr[0].reshape(d, cl, -1), rr[1].reshape(d, cl, -1), cr_[0][:, None, :],
                                 cr_[1][:, None, :], bk)                                 # (d, cl, tm, tn, 2)
        self.R3 = [np.ascontiguousarray(R3[sp]) for sp in range(d)]
        # (L3) as one segmented launch over s (selector s: offsets s*Dr into a, s*cr*Dr into t2); work-sorted
        # schedules for the banded launches
        self.L3_seg = segment_tables(self.L3)
        self.L3_off = np.stack([np.arange(d) * Dr, np.arange(d) * cr * Dr], axis=1).astype(np.int64)
        self._dev = None
        self._aux = None

    def _upload(self, device):
        if self._dev is None or self._dev[0] != device:
            up = lambda t: torch.from_numpy(np.ascontiguousarray(t)).to(device)      # noqa: E731
            self._dev = (device, [up(t) for t in self.L1], [up(t) for t in self.L3], up(self.R1),
                         [up(t) for t in self.R3])
            self._aux = {"L1": [up(banded_order(t)) for t in self.L1], "R1": up(banded_order(self.R1)),
                         "R3": [up(banded_order(t)) for t in self.R3],
                         "L3": tuple(up(t) for t in self.L3_seg) + (up(self.L3_off),)}
        return self._dev[1:]

    def _prep(self, a, w, env):
        cplx = dev.any_complex(a, env, w)
        assert cplx == self.cplx
        a = dev.as_dtype(a, cplx); env = dev.as_dtype(env, cplx)
        return cplx, a, dev.dense(w), env

    def _w_step(self, lib, dt, cplx, w, transposed, tin, tout, nb, rows_out, rows_in, drp, stream):
        """tout[i] = op(W) tin[i]; op = transpose for step_left (chain_ops.py:96)."""
        wmat = w.reshape(w.shape[0] * w.shape[1], w.shape[2] * w.shape[3])
        if transposed:
            # one transposed copy per MPO tensor (keyed like dev.w_csr): a fresh copy on every call would defeat the
            # CSR cache -- a device->host copy of W per environment update inside the otherwise sync-free sweeps
            key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype)
            hit = _WT_CACHE.get(key)
            if hit is None:
                if len(_WT_CACHE) > 256:
                    _WT_CACHE.clear()
                hit = (dev.dense(wmat.T), w)
                _WT_CACHE[key] = hit
            wmat = hit[0]
        w4 = wmat.reshape(1, rows_out, rows_in, 1)
        csr = dev.w_csr(w4) if (cplx or not w.dtype.is_complex) else None
        if csr is not None:
            rowptr, col, val, _ = csr
            _lib.check(lib.ptb_wapply_csr(dt, int(w.dtype.is_complex), rows_out, rows_in, drp, rowptr.data_ptr(),
                                          col.data_ptr(), val.data_ptr(), tin.data_ptr(), tout.data_ptr(), nb,
                                          stream), "ptb_wapply_csr")
        elif cplx and not w.dtype.is_complex:
            dev.gemm_strided(False, 0, 0, 0, rows_out, 2 * drp, rows_in, wmat, rows_in, torch.view_as_real(tin),
                             2 * drp, torch.view_as_real(tout), 2 * drp, nb, 0, 2 * rows_in * drp, 2 * rows_out * drp)
        else:
            dev.gemm_strided(cplx, 0, 0, 0, rows_out, drp, rows_in, dev.as_dtype(wmat, cplx), rows_in, tin, drp,
                             tout, drp, nb, 0, rows_in * drp, rows_out * drp)

    def step_left(self, a, w, l):
        """l_next[j,K,j'] = sum l[i,k,i'] conj(a[i',s',j']) w[k,s',s,K] a[i,s,j]."""
        lib = _lib.load()
        Dl, d, Dr, cl, cr = self.dims
        cplx, a, w, l = self._prep(a, w, l)
        if not self.supported:
            from .chain_ops import contraction_operator_step_left
            return contraction_operator_step_left(a, a, w, l)
        L1, L3, _, _ = self._upload(a.device)
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        es = 16 if cplx else 8
        stream = dev.stream_ptr(a.device)
        t = torch.zeros((Dl, cl * d, Dr), dtype=a.dtype, device=a.device)
        t2 = torch.empty((Dl, d * cr, Dr), dtype=a.dtype, device=a.device)
        for k in range(cl):                                                            # (L1)
            if not _active(self.L1[k]):
                continue
            _banded(lib, dt, 0, 0, 1, Dl, Dr, Dl, l.data_ptr() + k * Dl * es, cl * Dl, a.data_ptr(), d * Dr,
                    t.data_ptr() + k * d * Dr * es, cl * d * Dr, d, 0, Dr, Dr, False, L1[k], stream, "step_left(1)",
                    self._aux["L1"][k])
        self._w_step(lib, dt, cplx, w, True, t, t2, Dl, d * cr, cl * d, Dr, stream)     # (L2)
        out = torch.empty((Dr, cr, Dr), dtype=a.dtype, device=a.device)
        if _SEGMENTED:                                                                 # (L3), one launch
            seg_ptr, segs, order, off = self._aux["L3"]
            tabs = _lib.SectorTables(None, seg_ptr.data_ptr(), segs.data_ptr(), off.data_ptr(),
                                     order.data_ptr() if _ORDERED else None)
            st = lib.ptb_gemm_sector(dt, 1, 0, 0, Dr, cr * Dr, Dl, a.data_ptr(), d * Dr, t2.data_ptr(), d * cr * Dr,
                                     out.data_ptr(), cr * Dr, 1, 0, 0, 0, 0, ctypes.byref(tabs), stream)
            _lib.check(st, "step_left(3)")
            return out
        first = True
        for s in range(d):                                                             # (L3), one launch per s
            if not _active(self.L3[s]):
                continue
            _banded(lib, dt, 1, 0, 0, Dr, cr * Dr, Dl, a.data_ptr() + s * Dr * es, d * Dr,
                    t2.data_ptr() + s * cr * Dr * es, d * cr * Dr, out.data_ptr(), cr * Dr, 1, 0, 0, 0, not first,
                    L3[s], stream, "step_left(3)")
            first = False
        if first:
            out.zero_()
        return out

    def step_right(self, a, w, r):
        """r_next[i,k,i'] = sum a[i,s,j] r[j,K,j'] w[k,s',s,K] conj(a[i',s',j'])."""
        lib = _lib.load()
        Dl, d, Dr, cl, cr = self.dims
        cplx, a, w, r = self._prep(a, w, r)
        if not self.supported:
            from .chain_ops import contraction_operator_step_right
            return contraction_operator_step_right(a, a, w, r)
        _, _, R1, R3 = self._upload(a.device)
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        es = 16 if cplx else 8
        stream = dev.stream_ptr(a.device)
        t1 = torch.empty((Dl, d * cr, Dr), dtype=a.dtype, device=a.device)
        t2 = torch.empty((Dl, cl * d, Dr), dtype=a.dtype, device=a.device)
        _banded(lib, dt, 0, 0, 0, Dl, cr * Dr, Dr, a.data_ptr(), d * Dr, r.data_ptr(), cr * Dr, t1.data_ptr(),
                d * cr * Dr, d, Dr, 0, cr * Dr, False, R1, stream, "step_right(1)", self._aux["R1"])      # (R1)
        self._w_step(lib, dt, cplx, w, False, t1, t2, Dl, cl * d, d * cr, Dr, stream)                    # (R2)
        out = torch.empty((Dl, cl, Dl), dtype=a.dtype, device=a.device)
        first = True
        for sp in range(d):                                                                              # (R3)
            if not _active(self.R3[sp]):
                continue
            _banded(lib, dt, 0, 1, 1, Dl, Dl, Dr, t2.data_ptr() + sp * Dr * es, cl * d * Dr,
                    a.data_ptr() + sp * Dr * es, d * Dr, out.data_ptr(), cl * Dl, cl, d * Dr, 0, Dl, not first,
                    R3[sp], stream, "step_right(3)", self._aux["R3"][sp])
            first = False
        if first:
            out.zero_()
        return out


class BondSectorPlan:
    """Sector work lists for the zero-site contraction out[i',j'] = sum l[i,k,i'] c[i,j] r[j,k,j']
    (pytenet/chain_ops.py:282-317).  `qbl` are the quantum numbers of the rows of `c` (and both bond legs of
    `l`), `qbr` those of its columns (and of `r`); `c[i,j] != 0` only if `qbl[i] == qbr[j]`."""

    def __init__(self, qbl, qbr, qw, cplx=True):
        self.qbl = np.asarray(qbl, dtype=np.int64)
        self.qbr = np.asarray(qbr, dtype=np.int64)
        self.qw = np.asarray(qw, dtype=np.int64)
        self.cplx = cplx
        bm, bn, bk = _tile_shape(cplx)
        Dl, Dr, chi = len(self.qbl), len(self.qbr), len(self.qw)
        self.dims = (Dl, Dr, chi)
        self.supported = cplx or (Dl % 2 == 0 and Dr % 2 == 0)
        cols = (self.qbr[None, :] - self.qw[:, None]).reshape(-1)
        # (1) t[i,(k,j')] = sum_j c[i,j] r[j,(k,j')]: row i needs qbr[j] = qbl[i]; column (k,j') needs qbr[j] = qbr[j'] - qw[k]
        kxl, kxr = KIndex(self.qbl), KIndex(self.qbr)
        self.B1 = tile_k_ranges(self.qbl, cols, kxr, bm, bn, bk)[None]
        # (2) per k: out[i',j'] += sum_i l[i,k,i'] t[i,k,j']: row i' needs qbl[i] = qbl[i'] - qw[k];
        #     column j' needs qbl[i] = qbr[j'] - qw[k]
        self.B2 = [tile_k_ranges(self.qbl - self.qw[k], self.qbr - self.qw[k], kxl, bm, bn, bk)[None]
                   for k in range(chi)]
        # step 2 as one segmented launch (selector k: offsets k*Dl into l, k*Dr into t), work-sorted schedules
        self.seg_ptr_host, self.segs_host, self.order2_host = segment_tables(self.B2)
        self.sel_off_host = np.stack([np.arange(chi) * Dl, np.arange(chi) * Dr], axis=1).astype(np.int64)
        self.order1_host = banded_order(self.B1)
        self._dev = None

    def apply(self, c, l, r):
        lib = _lib.load()
        Dl, Dr, chi = self.dims
        cplx = dev.any_complex(c, l, r)
        assert cplx == self.cplx
        c = dev.as_dtype(c, cplx); l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
        assert tuple(c.shape) == (Dl, Dr) and tuple(l.shape) == (Dl, chi, Dl) and tuple(r.shape) == (Dr, chi, Dr)
        if not self.supported:
            from .chain_ops import apply_local_bond_contraction
            return apply_local_bond_contraction(c, l, r)
        if self._dev is None or self._dev[0] != c.device:
            up = lambda t: torch.from_numpy(np.ascontiguousarray(t)).to(c.device)     # noqa: E731
            self._dev = (c.device, up(self.B1), [up(t) for t in self.B2], up(self.seg_ptr_host), up(self.segs_host),
                         up(self.sel_off_host), up(self.order1_host), up(self.order2_host))
        _, B1, B2, seg_ptr, segs, sel_off, order1, order2 = self._dev
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        es = 16 if cplx else 8
        stream = dev.stream_ptr(c.device)
        t = torch.empty((Dl, chi * Dr), dtype=c.dtype, device=c.device)
        tabs = _lib.SectorTables(B1.data_ptr(), None, None, None, order1.data_ptr() if _ORDERED else None)
        st = lib.ptb_gemm_sector(dt, 0, 0, 0, Dl, chi * Dr, Dr, c.data_ptr(), Dr, r.data_ptr(), chi * Dr, t.data_ptr(),
                                 chi * Dr, 1, 0, 0, 0, 0, ctypes.byref(tabs), stream)
        _lib.check(st, "bond(1)")
        out = torch.empty((Dl, Dr), dtype=c.dtype, device=c.device)
        if _SEGMENTED:
            tabs = _lib.SectorTables(None, seg_ptr.data_ptr(), segs.data_ptr(), sel_off.data_ptr(),
                                     order2.data_ptr() if _ORDERED else None)
            st = lib.ptb_gemm_sector(dt, 1, 0, 0, Dl, Dr, Dl, l.data_ptr(), chi * Dl, t.data_ptr(), chi * Dr,
                                     out.data_ptr(), Dr, 1, 0, 0, 0, 0, ctypes.byref(tabs), stream)
            _lib.check(st, "bond(2)")
            return out
        first = True
        for k in range(chi):
            if not _active(self.B2[k]):
                continue
            _banded(lib, dt, 1, 0, 0, Dl, Dr, Dl, l.data_ptr() + k * Dl * es, chi * Dl, t.data_ptr() + k * Dr * es,
                    chi * Dr, out.data_ptr(), Dr, 1, 0, 0, 0, not first, B2[k], stream, "bond(2)")
            first = False
        if first:
            out.zero_()
        return out


class AbsorbSectorPlan:
    """
    Gauge absorption of the single-site sweeps (pytenet/tdvp.py:84,112) with the block structure of its operands:
    `c` is block diagonal in the quantum numbers (`c[x, y] != 0` only if `qc_rows[x] == qc_cols[y]`), the site tensor
    `a[i,s,j]` obeys `ql[i] + qs[s] == qr[j]`.

      left = True :  out[x, s, j]  = sum_i c[x, i]  a[i, s, j]      (qc_cols are the quantum numbers of a's left bond)
      left = False:  out[i, s, y]  = sum_j a[i, s, j] c[j, y]       (qc_rows are those of a's right bond)

    One banded GEMM (`ptb_gemm_sector`): every output tile visits only the k-tiles that can be non-zero.
    """

    def __init__(self, qc_rows, qc_cols, qs, q_other, left, cplx=True):
        self.cplx = cplx
        self.left = bool(left)
        qc_rows = np.asarray(qc_rows, dtype=np.int64); qc_cols = np.asarray(qc_cols, dtype=np.int64)
        qs = np.asarray(qs, dtype=np.int64); q_other = np.asarray(q_other, dtype=np.int64)
        bm, bn, bk = _tile_shape(cplx)
        d = len(qs)
        if self.left:
            # rows x of c; columns (s, j) of a viewed as (Dk, d*Dr) with q_other = qr: k = i needs ql[i] == qc_rows[x]
            # and ql[i] == qr[j] - qs[s]
            self.dims = (len(qc_rows), d * len(q_other), len(qc_cols))
            need_rows = qc_rows
            need_cols = (-qs[:, None] + q_other[None, :]).reshape(-1)
            qk = qc_cols
        else:
            # rows (i, s) of a viewed as (Dl*d, Dk) with q_other = ql: k = j needs qr[j] == ql[i] + qs[s] and
            # qr[j] == qc_cols[y]
            self.dims = (len(q_other) * d, len(qc_cols), len(qc_rows))
            need_rows = (q_other[:, None] + qs[None, :]).reshape(-1)
            need_cols = qc_cols
            qk = qc_rows
        m, n, k = self.dims
        self.supported = cplx or (m % 2 == 0 and n % 2 == 0 and k % 2 == 0)
        self.tab_host = np.ascontiguousarray(tile_k_ranges(need_rows, need_cols, KIndex(qk), bm, bn, bk)[None])
        self.order_host = banded_order(self.tab_host)
        self._dev = None

    def apply(self, c, a):
        lib = _lib.load()
        cplx = dev.any_complex(c, a)
        m, n, k = self.dims
        if cplx != self.cplx or not self.supported:
            if self.left:
                return dev.gemm(c, a.reshape(a.shape[0], -1)).reshape((c.shape[0],) + tuple(a.shape[1:]))
            return dev.gemm(a.reshape(-1, a.shape[2]), c).reshape(tuple(a.shape[:2]) + (c.shape[1],))
        c = dev.as_dtype(c, cplx); a = dev.as_dtype(a, cplx)
        if self._dev is None or self._dev[0] != c.device:
            self._dev = (c.device, torch.from_numpy(self.tab_host).to(c.device),
                         torch.from_numpy(self.order_host).to(c.device))
        _, tab, order = self._dev
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        if self.left:
            out = torch.empty((c.shape[0],) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
            assert tuple(c.shape) == (m, k) and a.shape[0] == k and a.shape[1] * a.shape[2] == n
            _banded(lib, dt, 0, 0, 0, m, n, k, c.data_ptr(), k, a.data_ptr(), n, out.data_ptr(), n, 1, 0, 0, 0, False,
                    tab, dev.stream_ptr(c.device), "absorb(left)", order)
        else:
            out = torch.empty(tuple(a.shape[:2]) + (c.shape[1],), dtype=a.dtype, device=a.device)
            assert tuple(c.shape) == (k, n) and a.shape[2] == k and a.shape[0] * a.shape[1] == m
            _banded(lib, dt, 0, 0, 0, m, n, k, a.data_ptr(), k, c.data_ptr(), n, out.data_ptr(), n, 1, 0, 0, 0, False,
                    tab, dev.stream_ptr(c.device), "absorb(right)", order)
        return out
