"""
Shared machinery of the four sweep drivers (tdvp.py, dmrg.py): environment
bookkeeping and the local Krylov problems, all device-resident.

The closures handed to the Krylov drivers are the integration point named in
SURVEY.md section 8(a10) (pytenet/tdvp.py:223-238, dmrg.py:181-189): they bind
(w, l, r), view the flat Lanczos vector as the local tensor (no copy) and call
the fused contraction chain.
"""
import os

import numpy as np
import torch

from . import _device as dev
from . import _lib
from ._prof import region
from .sectors import HeffSectorPlan, EnvSectorPlan, BondSectorPlan, AbsorbSectorPlan
from .sector_packed import PackedHeffPlan, PackedEnvPlan
from .block_sparse_util import is_qsparse
from .chain_ops import (apply_local_hamiltonian, apply_local_bond_contraction,
                        compute_right_operator_blocks, contraction_operator_step_left,
                        contraction_operator_step_right)
from .krylov import eigh_krylov, expm_krylov


def prepare_environments(hamiltonian, psi):
    """Right-orthonormalise `psi`, build all right blocks and the dummy left block, and run the
    reference's sparsity consistency check (tdvp.py:51-63, dmrg.py:44-56).  Returns (nrm, lblocks, rblocks)."""
    nsites = hamiltonian.nsites
    assert nsites == psi.nsites
    nrm = psi.orthonormalize(mode="right")
    if _use_sectors(max(psi.bond_dims), psi.qsite, *psi.qbonds, *hamiltonian.qbonds):
        # compute_right_operator_blocks (chain_ops.py:102-113) through the sector work lists: at D = 2048 with
        # quantum numbers the dense contraction of every block costs 30x the sector path
        rblocks = [None for _ in range(nsites)]
        rblocks[nsites - 1] = torch.ones((1, 1, 1), dtype=dev.F64, device=psi.a[-1].device)
        for i in reversed(range(nsites - 1)):
            rblocks[i] = env_step_right(psi, hamiltonian, i + 1, rblocks[i + 1])
    else:
        rblocks = compute_right_operator_blocks(psi, hamiltonian)
    lblocks = [None for _ in range(nsites)]
    lblocks[0] = torch.ones((1, 1, 1), dtype=rblocks[0].dtype, device=rblocks[0].device)
    for i, rb in enumerate(rblocks):
        assert is_qsparse(rb, [psi.qbonds[i + 1], hamiltonian.qbonds[i + 1], -psi.qbonds[i + 1]]), \
            "sparsity pattern of operator blocks must match quantum numbers"
    return nrm, lblocks, rblocks


# Sector-banded matvec (pytenet_b200/sectors.py): "auto" uses it when the quantum numbers are
# non-trivial and the bonds are large enough for whole tiles to be skipped; "1" forces it, "0" disables.
_SECTOR_MODE = os.environ.get("PYTENET_B200_SECTORS", "auto")
_SECTOR_MIN_BOND = 256
_PACKED = os.environ.get("PYTENET_B200_PACKED", "1") != "0"


# Plans depend on the quantum numbers only.  Sweeps revisit the same bonds (every time step of TDVP, every sweep of
# DMRG once the sector layouts have settled), so finished plans -- host work lists and their device copies -- are
# kept in a small cache keyed by the quantum-number arrays.
_PLAN_CACHE = {}
_PLAN_CACHE_MAX = 512


def _cached_plan(cls, cplx, *qnums):
    key = (cls.__name__, bool(cplx)) + tuple(np.asarray(q, dtype=np.int64).tobytes() for q in qnums)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        if len(_PLAN_CACHE) >= _PLAN_CACHE_MAX:
            _PLAN_CACHE.clear()
        plan = cls(*qnums, cplx=cplx)
        _PLAN_CACHE[key] = plan
    return plan


def sector_plan(ql, qs, qr, qwl, qwr, like, *others):
    """HeffSectorPlan for a local problem, or None when the dense path is the right choice.  The plan's dtype is
    the promoted dtype of the state tensor `like` and of `others` (environments, MPO tensor), as NumPy promotes."""
    if _SECTOR_MODE == "0":
        return None
    if _SECTOR_MODE != "1" and max(len(ql), len(qr)) < _SECTOR_MIN_BOND:
        return None
    if not (np.any(ql) or np.any(qr) or np.any(qs) or np.any(qwl) or np.any(qwr)):
        return None
    cplx = True if like is None else dev.any_complex(like, *[t for t in others if isinstance(t, torch.Tensor)])
    if _PACKED:
        # sector-packed grouped GEMM (sector_packed.py) whenever the bonds are grouped by sector -- every bond a
        # sweep has orthonormalised is; else the banded work lists over the dense layout (sectors.py)
        plan = _cached_plan(PackedHeffPlan, cplx, ql, qs, qr, qwl, qwr)
        if plan.supported:
            return plan
    return _cached_plan(HeffSectorPlan, cplx, ql, qs, qr, qwl, qwr)


def _use_sectors(nbond, *qnums):
    if _SECTOR_MODE == "0":
        return False
    if _SECTOR_MODE != "1" and nbond < _SECTOR_MIN_BOND:
        return False
    return any(np.any(q) for q in qnums)


class _PackedEnvLeft(PackedEnvPlan):
    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True):
        super().__init__(ql, qs, qr, qwl, qwr, cplx=cplx, side="left")


class _PackedEnvRight(PackedEnvPlan):
    def __init__(self, ql, qs, qr, qwl, qwr, cplx=True):
        super().__init__(ql, qs, qr, qwl, qwr, cplx=cplx, side="right")


def _packed_env(cls, a, w, env, qn):
    """Sector-packed environment update (sector_packed.PackedEnvPlan), or None when the bonds are not grouped by
    sector or the operands are not plain device tensors of the expected shapes."""
    if not _PACKED or w.dtype not in (dev.F64, dev.C128) or not (a.is_cuda and w.is_cuda and env.is_cuda):
        return None
    plan = _cached_plan(cls, dev.any_complex(a, env, w), *qn)
    return plan.apply(a, w, env) if plan.supported else None


def env_step_left(psi, hamiltonian, i, l):
    """lblocks[i+1] from lblocks[i] and site i (tdvp.py:79,179; dmrg.py:72,153), through the sector work lists
    when the quantum numbers are non-trivial and the bonds large, else the dense contraction."""
    a, w = psi.a[i], hamiltonian.a[i]
    qn = (psi.qbonds[i], psi.qsite, psi.qbonds[i + 1], hamiltonian.qbonds[i], hamiltonian.qbonds[i + 1])
    with region("env"):
        if _use_sectors(max(a.shape[0], a.shape[2]), *qn) and isinstance(l, torch.Tensor) and l.shape[0] == a.shape[0]:
            out = _packed_env(_PackedEnvLeft, a, w, l, qn)
            if out is not None:
                return out
            plan = _cached_plan(EnvSectorPlan, dev.any_complex(a, l, w), *qn)
            return plan.step_left(a, w, l)
        return contraction_operator_step_left(a, a, w, l)


def env_step_right(psi, hamiltonian, i, r):
    """rblocks[i-1] from rblocks[i] and site i (tdvp.py:106,197,216; dmrg.py:83,168); see env_step_left."""
    a, w = psi.a[i], hamiltonian.a[i]
    qn = (psi.qbonds[i], psi.qsite, psi.qbonds[i + 1], hamiltonian.qbonds[i], hamiltonian.qbonds[i + 1])
    with region("env"):
        if _use_sectors(max(a.shape[0], a.shape[2]), *qn) and isinstance(r, torch.Tensor) and r.shape[0] == a.shape[2]:
            out = _packed_env(_PackedEnvRight, a, w, r, qn)
            if out is not None:
                return out
            plan = _cached_plan(EnvSectorPlan, dev.any_complex(a, r, w), *qn)
            return plan.step_right(a, w, r)
        return contraction_operator_step_right(a, a, w, r)


def bond_plan(qbl, qbr, qw, c, l, r):
    """Sector plan for the zero-site problem on a bond (rows of `c`: qbl, columns: qbr), or None.  The zero-site
    contraction out[i',j'] = sum l[i,k,i'] c[i,j] r[j,k,j'] is the site contraction with a one-dimensional physical
    index of quantum number 0 and the identity on the MPO bond as MPO tensor, so the sector-packed plan of the site
    problem serves it unchanged (grouped GEMMs over the sector blocks, Lanczos run in the packed space); float64
    problems and bonds that are not grouped by sector use the banded work lists."""
    if not _use_sectors(max(c.shape), qbl, qbr, qw):
        return None
    cplx = dev.any_complex(c, l, r)
    if _PACKED:
        plan = _cached_plan(PackedHeffPlan, cplx, qbl, np.zeros(1, dtype=np.int64), qbr, qw, qw)
        if plan.supported:
            return plan
    return _cached_plan(BondSectorPlan, cplx, qbl, qbr, qw)


_IDENTITY_W = {}


def _identity_w(chi, device):
    """w[k, 0, 0, kappa] = delta(k, kappa): the MPO tensor that turns the site contraction into the zero-site one."""
    key = (chi, device.index)
    w = _IDENTITY_W.get(key)
    if w is None:
        w = torch.eye(chi, dtype=dev.F64, device=device).reshape(chi, 1, 1, chi).contiguous()
        _IDENTITY_W[key] = w
    return w


def absorb_left(c, a, qc_rows, qc_cols, qs, qr):
    """out[x,s,j] = sum_i c[x,i] a[i,s,j] (tdvp.py:84): banded over the sector blocks when that pays, else dense."""
    if _use_sectors(max(c.shape + tuple(a.shape[::2])), qc_rows, qc_cols, qs, qr):
        plan = _cached_plan(_AbsorbLeft, dev.any_complex(c, a), qc_rows, qc_cols, qs, qr)
        return plan.apply(c, a)
    return dev.gemm(c, a.reshape(a.shape[0], -1)).reshape((c.shape[0],) + tuple(a.shape[1:]))


def absorb_right(a, c, qc_rows, qc_cols, qs, ql):
    """out[i,s,y] = sum_j a[i,s,j] c[j,y] (tdvp.py:112)."""
    if _use_sectors(max(c.shape + tuple(a.shape[::2])), qc_rows, qc_cols, qs, ql):
        plan = _cached_plan(_AbsorbRight, dev.any_complex(c, a), qc_rows, qc_cols, qs, ql)
        return plan.apply(c, a)
    return dev.gemm(a.reshape(-1, a.shape[2]), c).reshape(tuple(a.shape[:2]) + (c.shape[1],))


class _AbsorbLeft(AbsorbSectorPlan):
    def __init__(self, qc_rows, qc_cols, qs, q_other, cplx=True):
        super().__init__(qc_rows, qc_cols, qs, q_other, True, cplx=cplx)


class _AbsorbRight(AbsorbSectorPlan):
    def __init__(self, qc_rows, qc_cols, qs, q_other, cplx=True):
        super().__init__(qc_rows, qc_cols, qs, q_other, False, cplx=cplx)


_SMALL_FITS = {}        # (Dl, d, Dr, cl, cr, numiter) -> bool: ptb_local_step_small_fits, asked once per shape


def _small_local_step(x, w, l, r, dims, numiter, dt, V=None, scal=None):
    """The whole local problem in ONE kernel launch (csrc/lanczos_small.cu: ptb_local_step_small) when it is small
    enough -- the launch-latency regime (README config, METTS, chain edges).  `dt is None`: the Lanczos run only
    (fills `V`, `scal`; returns True).  Otherwise returns (exp(-dt H) x ... as `out`, scal); the Lanczos vectors
    then live in a per-stream scratch buffer, and `scal` may be a page-locked HOST tensor (krylov._defer_slot):
    the kernel writes its 2 numiter scalars straight into it (mapped memory), no copy is enqueued.
    Returns None when the problem does not qualify."""
    Dl, d, Dr, cl, cr = dims
    lib = _lib.load()
    key = (Dl, d, Dr, cl, cr, numiter)
    fits = _SMALL_FITS.get(key)
    if fits is None:
        fits = numiter <= 64 and bool(lib.ptb_local_step_small_fits(Dl, d, Dr, cl, cr, numiter))
        _SMALL_FITS[key] = fits
    if not fits:
        return None
    if not (isinstance(x, torch.Tensor) and isinstance(l, torch.Tensor) and isinstance(r, torch.Tensor)
            and x.is_cuda and l.is_cuda and r.is_cuda):
        return None
    cplx = x.dtype.is_complex
    if x.dtype not in (dev.F64, dev.C128):
        return None
    if (l.dtype.is_complex or r.dtype.is_complex) and not cplx:
        return None                      # mixed dtypes: the step-by-step path applies NumPy's promotion rules
    w_cplx = False
    if w is not None:
        if not (isinstance(w, torch.Tensor) and w.is_cuda) or w.dtype not in (dev.F64, dev.C128):
            return None
        w_cplx = w.dtype.is_complex
        if w_cplx and not cplx:
            return None
        w = dev.dense(w)
    device = x.device
    l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
    n = Dl * d * Dr
    dt_code = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    if V is None:
        # internal to the step: (numiter x n) elements of a grow-only, stream-ordered scratch buffer
        V = dev.workspace(numiter * n * (16 if cplx else 8), device, tag="small_step_V")
    if scal is None:
        scal = torch.empty(2 * numiter, dtype=dev.F64, device=device)      # the kernel writes every entry
    out = None
    apply_expm, dre, dim_, out_cplx = 0, 0.0, 0.0, 0
    if dt is not None:
        dtc = complex(dt)
        out_cplx = int(cplx or isinstance(dt, (complex, np.complexfloating)))
        out = torch.empty(n, dtype=dev.C128 if out_cplx else dev.F64, device=device)
        apply_expm, dre, dim_ = 1, dtc.real, dtc.imag
    st = lib.ptb_local_step_small(dt_code, x.data_ptr(), w.data_ptr() if w is not None else None, int(w_cplx),
                                  l.data_ptr(), r.data_ptr(), Dl, d, Dr, cl, cr, numiter, V.data_ptr(),
                                  scal.data_ptr(), apply_expm, dre, dim_, out_cplx,
                                  out.data_ptr() if out is not None else None, None, 0,
                                  dev.stream_ptr(device))
    if st != 0:
        _lib.check(st, "ptb_local_step_small")
    return True if dt is None else (out, scal)


class HeffOperator:
    """The closure of tdvp.py:223-229 / dmrg.py:181-189 as an object: calling it applies the local effective
    Hamiltonian to a flat vector; `ptb_lanczos_run` hands a whole Lanczos run to the fused C entry
    (ptb_heff_lanczos: every iteration enqueued by one call), which krylov._lanczos_core prefers."""

    def __init__(self, w, l, r, shape):
        self.w, self.l, self.r, self.shape = w, l, r, tuple(shape)

    def __call__(self, x):
        return apply_local_hamiltonian(x.reshape(self.shape), self.w, self.l, self.r).reshape(-1)

    def _dims(self, x):
        w, l, r = self.w, self.l, self.r
        Dl, d, Dr = self.shape
        if not all(isinstance(t, torch.Tensor) for t in (w, l, r)) or w.ndim != 4:
            return None
        cl, dout, din, cr = w.shape
        if (dout != d or din != d or tuple(l.shape) != (Dl, cl, Dl) or tuple(r.shape) != (Dr, cr, Dr)
                or x.numel() != Dl * d * Dr):
            return None
        return Dl, d, Dr, cl, cr

    def ptb_expm_run(self, x, dt, numiter, scal=None):
        """exp(dt H_eff) x in one kernel launch when the problem is small (krylov._expm_device prefers it)."""
        dims = self._dims(x)
        return None if dims is None else _small_local_step(x, self.w, self.l, self.r, dims, numiter, dt, scal=scal)

    def ptb_lanczos_run(self, x, numiter, V, scal):
        w, l, r = self.w, self.l, self.r
        Dl, d, Dr = self.shape
        if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in (w, l, r)) or w.ndim != 4:
            return False
        dims = self._dims(x)
        if dims is not None and _small_local_step(x, w, l, r, dims, numiter, None, V, scal):
            return True
        cplx = x.dtype.is_complex
        if (l.dtype.is_complex or r.dtype.is_complex or w.dtype.is_complex) and not cplx:
            return False                 # mixed dtypes: the step-by-step path applies NumPy's promotion rules
        cl, dout, din, cr = w.shape
        if (dout != d or din != d or tuple(l.shape) != (Dl, cl, Dl) or tuple(r.shape) != (Dr, cr, Dr)
                or x.numel() != Dl * d * Dr):
            return False
        lib = _lib.load()
        device = x.device
        l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
        w_cplx = w.dtype.is_complex
        if w.dtype not in (dev.F64, dev.C128):
            return False
        w = dev.dense(w)
        csr = dev.w_csr(w)
        rowptr, col, val = (csr[0].data_ptr(), csr[1].data_ptr(), csr[2].data_ptr()) if csr is not None else (None,) * 3
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        nbytes = lib.ptb_heff_lanczos_workspace_bytes(dt, Dl, d, Dr, cl, cr)
        ws = dev.workspace(nbytes, device, tag="lanczos")
        st = lib.ptb_heff_lanczos(dt, x.data_ptr(), w.data_ptr(), int(w_cplx), rowptr, col, val, l.data_ptr(),
                                  r.data_ptr(), Dl, d, Dr, cl, cr, numiter, V.data_ptr(), scal.data_ptr(),
                                  dev.lanczos_scratch(device).data_ptr(), ws.data_ptr(), nbytes,
                                  dev.stream_ptr(device))
        _lib.check(st, "heff_lanczos")
        return True


class BondOperator:
    """The zero-site closure of tdvp.py:232-238; fused form: ptb_bond_lanczos."""

    def __init__(self, l, r, shape):
        self.l, self.r, self.shape = l, r, tuple(shape)

    def __call__(self, x):
        return apply_local_bond_contraction(x.reshape(self.shape), self.l, self.r).reshape(-1)

    def _dims(self, x):
        l, r = self.l, self.r
        Dl, Dr = self.shape
        if not all(isinstance(t, torch.Tensor) for t in (l, r)):
            return None
        chi = l.shape[1]
        if tuple(l.shape) != (Dl, chi, Dl) or tuple(r.shape) != (Dr, chi, Dr) or x.numel() != Dl * Dr:
            return None
        return Dl, 1, Dr, chi, chi

    def ptb_expm_run(self, x, dt, numiter, scal=None):
        dims = self._dims(x)
        return None if dims is None else _small_local_step(x, None, self.l, self.r, dims, numiter, dt, scal=scal)

    def ptb_lanczos_run(self, x, numiter, V, scal):
        l, r = self.l, self.r
        Dl, Dr = self.shape
        if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in (l, r)):
            return False
        dims = self._dims(x)
        if dims is not None and _small_local_step(x, None, l, r, dims, numiter, None, V, scal):
            return True
        cplx = x.dtype.is_complex
        if (l.dtype.is_complex or r.dtype.is_complex) and not cplx:
            return False
        chi = l.shape[1]
        if tuple(l.shape) != (Dl, chi, Dl) or tuple(r.shape) != (Dr, chi, Dr) or x.numel() != Dl * Dr:
            return False
        lib = _lib.load()
        device = x.device
        l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
        dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
        nbytes = lib.ptb_bond_lanczos_workspace_bytes(dt, Dl, Dr, chi)
        ws = dev.workspace(nbytes, device, tag="lanczos")
        st = lib.ptb_bond_lanczos(dt, x.data_ptr(), l.data_ptr(), r.data_ptr(), Dl, Dr, chi, numiter, V.data_ptr(),
                                  scal.data_ptr(), dev.lanczos_scratch(device).data_ptr(), ws.data_ptr(), nbytes,
                                  dev.stream_ptr(device))
        _lib.check(st, "bond_lanczos")
        return True


def _packed(plan, w, l, r, a):
    """(operator, packed start vector) when `plan` is a sector-packed plan that fits the operands, else None."""
    if not isinstance(plan, PackedHeffPlan):
        return None
    if not all(isinstance(t, torch.Tensor) and t.is_cuda for t in (w, l, r, a)):
        return None
    Dl, d, Dr, cl, cr = plan.dims
    if (tuple(a.shape) != (Dl, d, Dr) or tuple(w.shape) != (cl, d, d, cr) or tuple(l.shape) != (Dl, cl, Dl)
            or tuple(r.shape) != (Dr, cr, Dr) or w.dtype not in (dev.F64, dev.C128)):
        return None
    if dev.any_complex(w, l, r, a) != plan.cplx:
        return None
    op = plan.bind(w, l, r)
    return op, op.pack(a)


def _heff(w, l, r, shape, plan):
    if isinstance(plan, PackedHeffPlan):
        op = plan.bind(w, l, r)
        return lambda x: op.apply_dense(x.reshape(shape)).reshape(-1)        # dense-vector form (fallback only)
    if plan is not None:
        def matvec(x):
            if x.dtype.is_complex != plan.cplx:          # a real state turned complex (or vice versa)
                return apply_local_hamiltonian(x.reshape(shape), w, l, r).reshape(-1)
            return plan.apply(x.reshape(shape), w, l, r).reshape(-1)
        return matvec
    return HeffOperator(w, l, r, shape)


def local_hamiltonian_step(l, r, w, a, dt, numiter: int, plan=None):
    """exp(-dt H_eff) a for the one- or two-site effective Hamiltonian (tdvp.py:223-229)."""
    shape = tuple(a.shape)
    with region("lanczos"):
        pk = _packed(plan, w, l, r, a)
        if pk is not None:
            # the whole Krylov run lives in the packed space: pack once, unpack the result
            op, xp = pk
            return op.unpack(expm_krylov(op, xp, -dt, numiter, hermitian=True))
        return expm_krylov(_heff(w, l, r, shape, plan), a.reshape(-1), -dt, numiter, hermitian=True).reshape(shape)


def local_bond_step(l, r, c, dt, numiter: int, plan=None):
    """exp(-dt K_eff) c for the zero-site (bond) effective Hamiltonian (tdvp.py:232-238)."""
    shape = tuple(c.shape)
    with region("lanczos_bond"):
        if isinstance(plan, PackedHeffPlan):
            pk = _packed(plan, _identity_w(l.shape[1], c.device), l, r, c.reshape(shape[0], 1, shape[1]))
            if pk is not None:
                op, xp = pk
                return op.unpack(expm_krylov(op, xp, -dt, numiter, hermitian=True)).reshape(shape)
            plan = None
        if plan is not None and plan.cplx == (c.dtype.is_complex or l.dtype.is_complex or r.dtype.is_complex):
            def matvec(x):
                if x.dtype.is_complex != plan.cplx:
                    return apply_local_bond_contraction(x.reshape(shape), l, r).reshape(-1)
                return plan.apply(x.reshape(shape), l, r).reshape(-1)
            return expm_krylov(matvec, c.reshape(-1), -dt, numiter, hermitian=True).reshape(shape)
        return expm_krylov(BondOperator(l, r, shape), c.reshape(-1), -dt, numiter, hermitian=True).reshape(shape)


def minimize_local_energy(w, l, r, a_start, numiter: int, plan=None):
    """Lowest Ritz pair of the local effective Hamiltonian (dmrg.py:181-189)."""
    shape = tuple(a_start.shape)
    with region("lanczos"):
        pk = _packed(plan, w, l, r, a_start)
        if pk is not None:
            op, xp = pk
            ev, u_ritz = eigh_krylov(op, xp, numiter, 1)
            return ev[0], op.unpack(dev.dense(u_ritz[:, 0]))
        ev, u_ritz = eigh_krylov(_heff(w, l, r, shape, plan), a_start.reshape(-1), numiter, 1)
        return ev[0], dev.dense(u_ritz[:, 0]).reshape(shape)
