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

__all__ = ["HeffSectorPlan", "tile_k_ranges"]

_EMPTY_LO = np.iinfo(np.int64).max
_EMPTY_HI = -1


def _tile_shape(cplx):
    lib = _lib.load()
    bm, bn, bk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.check(lib.ptb_gemm_tile_shape(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64,
                                       ctypes.byref(bm), ctypes.byref(bn), ctypes.byref(bk)), "ptb_gemm_tile_shape")
    return bm.value, bn.value, bk.value


def _support(need, qk):
    """For every entry of `need` the bounding interval [lo, hi) of positions k with qk[k] == need
    (empty -> (_EMPTY_LO, _EMPTY_HI))."""
    qk = np.asarray(qk, dtype=np.int64)
    need = np.asarray(need, dtype=np.int64)
    vals, inv = np.unique(qk, return_inverse=True)
    pos = np.arange(len(qk))
    first = np.full(len(vals), _EMPTY_LO, dtype=np.int64)
    last = np.full(len(vals), _EMPTY_HI, dtype=np.int64)
    np.minimum.at(first, inv, pos)
    np.maximum.at(last, inv, pos + 1)
    if len(vals) == 0:
        return np.full(len(need), _EMPTY_LO, dtype=np.int64), np.full(len(need), _EMPTY_HI, dtype=np.int64)
    idx_c = np.clip(np.searchsorted(vals, need), 0, len(vals) - 1)
    found = vals[idx_c] == need
    lo = np.where(found, first[idx_c], _EMPTY_LO)
    hi = np.where(found, last[idx_c], _EMPTY_HI)
    return lo, hi


def _tile_reduce(lo, hi, tile):
    """Bounding interval per block of `tile` consecutive entries."""
    n = len(lo)
    starts = np.arange(0, n, tile)
    return np.minimum.reduceat(lo, starts), np.maximum.reduceat(hi, starts)


def tile_k_ranges(need_rows, need_cols, qk, bm, bn, bk):
    """(tiles_m, tiles_n, 2) int32 array of k-TILE ranges [lo, hi): row m of the output only receives
    contributions from k with qk[k] == need_rows[m], column n from k with qk[k] == need_cols[n]."""
    rlo, rhi = _tile_reduce(*_support(need_rows, qk), bm)
    clo, chi = _tile_reduce(*_support(need_cols, qk), bn)
    lo = np.maximum(rlo[:, None], clo[None, :])
    hi = np.minimum(rhi[:, None], chi[None, :])
    empty = hi <= lo
    lo_t = np.where(empty, 0, lo // bk)
    hi_t = np.where(empty, 0, -(-hi // bk))
    return np.stack([lo_t, hi_t], axis=-1).astype(np.int32)


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
        self.tab1_host = np.ascontiguousarray(np.stack(
            [tile_k_ranges(self.ql + self.qs_in[s], cols1, self.qr, bm, bn, bk) for s in range(d)]))

        # step 3, one launch per left MPO index k, batched over s':
        #   out[i', s', j'] += sum_i l[i, k, i'] t2[i, k, s', j']
        #   row i' needs  ql[i] = ql'[i'] - qwl[k];  column j' (batch s') needs  ql[i] = qr'[j'] - qs[s'] - qwl[k]
        self.tab3_host = []
        self.k_active = []
        for k in range(cl):
            tabs = np.ascontiguousarray(np.stack(
                [tile_k_ranges(self.qlp - self.qwl[k], self.qrp - self.qs_out[sp] - self.qwl[k], self.ql,
                               bm, bn, bk) for sp in range(dout)]))
            self.k_active.append(bool(np.any(tabs[..., 1] > tabs[..., 0])))
            self.tab3_host.append(tabs)
        self.tab1 = None
        self.tab3 = None

        # bookkeeping for benchmarks: fraction of the dense k-tile visits that remain
        kt1 = -(-Dr // bk)
        kt3 = -(-Dl // bk)
        vis1 = float(np.sum(self.tab1_host[..., 1] - self.tab1_host[..., 0]))
        den1 = float(self.tab1_host[..., 0].size * kt1)
        vis3 = sum(float(np.sum(t[..., 1] - t[..., 0])) for t in self.tab3_host)
        den3 = float(cl * dout * (-(-Dlp // bm)) * (-(-Drp // bn)) * kt3)
        self.visit_fraction = (vis1 / max(den1, 1.0), vis3 / max(den3, 1.0))

    def _upload(self, device):
        if self.tab1 is None or self.tab1.device != device:
            self.tab1 = torch.from_numpy(self.tab1_host).to(device)
            self.tab3 = [torch.from_numpy(t).to(device) if act else None
                         for t, act in zip(self.tab3_host, self.k_active)]

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
        ws = dev.workspace((n1 + n2) * es + 32, device, tag="sectors")
        t1 = ws[:n1 * es].view(a.dtype).reshape(Dl, d, cr * Drp)
        off2 = (n1 * es + 15) // 16 * 16
        t2 = ws[off2:off2 + n2 * es].view(a.dtype).reshape(Dl, cl * dout, Drp)
        if out is None:
            out = torch.empty((Dlp, dout, Drp), dtype=a.dtype, device=device)
        # (1) batched over s, banded in j
        st = lib.ptb_gemm_banded(dt, 0, 0, 0, Dl, cr * Drp, Dr, a.data_ptr(), d * Dr, r.data_ptr(), cr * Drp,
                                 t1.data_ptr(), d * cr * Drp, d, Dr, 0, cr * Drp, 0, self.tab1.data_ptr(), stream)
        _lib.check(st, "ptb_gemm_banded(step 1)")
        # (2) W step batched over i (t1[i] is (d*cr) x Drp, t2[i] is (cl*dout) x Drp): sparse CSR kernel for
        #     the usual sparse MPO tensors, dense small GEMM otherwise
        csr = dev.w_csr(w) if (cplx or not w.dtype.is_complex) else None
        if csr is not None:
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
        # (3) one banded launch per left MPO index, batched over s', accumulating into out
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
