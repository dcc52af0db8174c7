"""
MPO-bond-sharded effective Hamiltonian (BASELINE config 4: molecular
Hamiltonians with MPO bond dimension chi ~ O(N^2)).

    out[i',s',j'] = sum_{k,kappa}  l[i,k,i']  w[k,s',s,kappa]  a[i,s,j]  r[j,kappa,j']

is a sum over MPO-bond index pairs.  With G ranks (one process per GPU):

  * the RIGHT MPO bond kappa is split into G ranges: rank g holds r[:, kappa_g, :] and computes its
    slice of step 1,  t1_g[i,s,kappa_g,j'] = a . r_g                       (1/G of the step-1 flops)
  * one all-gather over NVLink makes t1 visible to every rank (the only bulk exchange)
  * the LEFT MPO bond k is split into G ranges: rank g holds l[:, k_g, :] and w[k_g, ...] and computes
    t2_g[i,k_g,s',j'] = sum_{g'} w[k_g, :, :, kappa_g'] t1_g'              (1/G of step 2)
    out_g = l_g^T . t2_g                                                    (1/G of step 3)
  * the partial results out_g are summed with one all-reduce (the matvec result, D*d*D elements),
    after which every rank holds the full Lanczos vector -- alpha / beta are then computed
    redundantly and identically on every rank, so the Krylov drivers run unchanged.

All three GEMM steps scale with 1/G; communication per matvec is one all-gather of t1 and one
all-reduce of `out`.  The collectives go through torch.distributed (NCCL over NVLink on GPUs; the
same code runs under gloo on CPU tensors in the tests, with the arithmetic supplied by an `ops`
object -- the product default is the CUDA engine, there is no CPU arithmetic in this package).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _device as dev

__all__ = ["ShardedEffectiveHamiltonian", "PrecontractedShardedHamiltonian", "bond_partition"]


def bond_partition(chi, nparts):
    """Contiguous, nearly equal ranges [(start, stop)] of an MPO bond of dimension `chi`."""
    base, extra = divmod(chi, nparts)
    ranges = []
    pos = 0
    for g in range(nparts):
        size = base + (1 if g < extra else 0)
        ranges.append((pos, pos + size))
        pos += size
    return ranges


def _flat_real(t):
    """1-D float64 view of a dense tensor (complex128 as interleaved pairs): what goes on the wire."""
    return (torch.view_as_real(t) if t.dtype.is_complex else t).reshape(-1)


class _CudaOps:
    """Arithmetic of the three steps on the DMMA engine (C ABI: ptb_gemm)."""

    @staticmethod
    def step1(a2d, r2d, out):
        return dev.gemm(a2d, r2d, out=out)

    @staticmethod
    def wapply(wblk, tin, tout, accumulate):
        """tout[i] (+)= wblk @ tin[i]  for all i.  wblk (R_out, R_in) real or complex;
        tin (B, R_in, Drp), tout (B, R_out, Drp) complex128 or float64, dense."""
        nb, rin, drp = tin.shape
        rout = tout.shape[1]
        if tin.dtype.is_complex and not wblk.dtype.is_complex:
            # real W on complex t: real GEMM over (re, im)-interleaved columns
            dev.gemm_strided(False, 0, 0, 0, rout, 2 * drp, rin, wblk, rin, torch.view_as_real(tin), 2 * drp,
                             torch.view_as_real(tout), 2 * drp, nb, 0, 2 * rin * drp, 2 * rout * drp, accumulate)
        else:
            cplx = tin.dtype.is_complex
            dev.gemm_strided(cplx, 0, 0, 0, rout, drp, rin, dev.as_dtype(wblk, cplx), rin, tin, drp, tout, drp,
                             nb, 0, rin * drp, rout * drp, accumulate)
        return tout

    @staticmethod
    def step3(l2d, t2d, out):
        return dev.gemm(l2d, t2d, trans_a=True, out=out)

    @staticmethod
    def precontract(w3, l, out):
        """out[i] = w3 @ l[i]:  w3 (R, chi_l) real or complex, l (Dl, chi_l, Dlp), out (Dl, R, Dlp)."""
        nb, cl, dlp = l.shape
        rows = w3.shape[0]
        if l.dtype.is_complex and not w3.dtype.is_complex:
            dev.gemm_strided(False, 0, 0, 0, rows, 2 * dlp, cl, w3, cl, torch.view_as_real(l), 2 * dlp,
                             torch.view_as_real(out), 2 * dlp, nb, 0, 2 * cl * dlp, 2 * rows * dlp)
        else:
            cplx = l.dtype.is_complex
            dev.gemm_strided(cplx, 0, 0, 0, rows, dlp, cl, dev.as_dtype(w3, cplx), cl, l, dlp, out, dlp,
                             nb, 0, cl * dlp, rows * dlp)
        return out

    @staticmethod
    def contract_lw(lw, t1, out):
        """out[i',s',j'] = sum_K lw[K, s', i'] t1[K, j']  (K = (i, s, kappa_loc)): one GEMM batched over s'
        with the contraction index split over work units (split-K) because the output has few tiles."""
        from . import _lib
        lib = _lib.load()
        kk, dout, dlp = lw.shape
        drp = t1.shape[1]
        cplx = lw.dtype.is_complex
        es = 16 if cplx else 8
        ws = dev.workspace(8 * dout * dlp * drp * es, lw.device, tag="splitk")
        st = lib.ptb_gemm_splitk(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, 1, 0, 0, dlp, drp, kk,
                                 lw.data_ptr(), dout * dlp, t1.data_ptr(), drp, out.data_ptr(), dout * drp,
                                 dout, dlp, 0, drp, 0, 0, ws.data_ptr(), ws.numel(), dev.stream_ptr(lw.device))
        _lib.check(st, "ptb_gemm_splitk")
        return out


class ShardedEffectiveHamiltonian:
    """
    One- or two-site effective Hamiltonian with both MPO bonds split across the ranks of a process
    group.  Build it once per site (environments and MPO tensor are fixed during a Lanczos run) and
    call :meth:`matvec` inside the Krylov drivers:

        heff = ShardedEffectiveHamiltonian.from_full(w, l, r)          # every rank passes the full tensors
        ptn.eigh_krylov(lambda x: heff.matvec(x.reshape(shape)).reshape(-1), a.reshape(-1), k, 1)
    """

    def __init__(self, w_blocks, l_shard, r_shard, dims, group=None, ops=None, exchange="auto"):
        # exchange of t1: step-1 GEMM followed by an NCCL all-gather.  (Round 1 also had a GEMM whose epilogue
        # stored its tiles into every rank's gathered buffer over NVLink; at 8 GPUs it lost to the plain NCCL
        # all-gather -- 148.2 vs 137.9 ms, the peer stores stalled the DMMA warps -- and both lose to the
        # all-reduce-only form below (91.4 ms), so it was removed.)
        self.exchange = exchange
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = ops if ops is not None else _CudaOps()
        self.w_blocks = w_blocks          # list over source rank g' of (k_g*dout, din*P) matrices
        self.l_shard = l_shard            # (Dl, k_g, Dlp)
        self.r_shard = r_shard            # (Dr, P, Drp), zero padded to the common shard size P
        (self.Dl, self.d_in, self.Dr, self.d_out, self.Dlp, self.Drp, self.kg, self.P) = dims
        self._t1 = self._gathered = self._t2 = None

    @classmethod
    def from_full(cls, w, l, r, group=None, ops=None, device=None, exchange="auto"):
        """Slice the full tensors (same on every rank) into this rank's shards."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        cl, dout, din, cr = w.shape
        Dl, cl2, Dlp = l.shape
        Dr, cr2, Drp = r.shape
        assert cl2 == cl and cr2 == cr
        kparts = bond_partition(cl, world)
        P = -(-cr // world)                                   # common (padded) kappa shard size
        k0, k1 = kparts[rank]
        kg = k1 - k0
        # NumPy-style promotion over all operands (the state the operator acts on is promoted in matvec)
        cplx = bool(l.dtype.is_complex or r.dtype.is_complex
                    or (w.dtype.is_complex if isinstance(w, torch.Tensor) else np.iscomplexobj(w)))

        def dense(x, want_cplx):
            x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            if device is not None:
                x = x.to(device)
            if want_cplx and not x.dtype.is_complex:
                x = x.to(torch.complex128)
            return x

        l_shard = dense(l[:, k0:k1, :], cplx).contiguous()
        r_pad = torch.zeros((Dr, P, Drp), dtype=l_shard.dtype, device=l_shard.device)
        q0, q1 = rank * P, min((rank + 1) * P, cr)
        if q1 > q0:
            r_pad[:, :q1 - q0, :] = dense(r[:, q0:q1, :], cplx)
        w_t = dense(w, False)
        w_blocks = []
        for gp in range(world):
            p0, p1 = gp * P, min((gp + 1) * P, cr)
            blk = torch.zeros((kg, dout, din, P), dtype=w_t.dtype, device=l_shard.device)
            if p1 > p0:
                blk[:, :, :, :p1 - p0] = w_t[k0:k1, :, :, p0:p1]
            w_blocks.append(blk.reshape(kg * dout, din * P).contiguous())
        return cls(w_blocks, l_shard, r_pad, (Dl, din, Dr, dout, Dlp, Drp, kg, P), group=group, ops=ops,
                   exchange=exchange)

    @classmethod
    def synthetic(cls, Dl, d, Dr, cl, cr, density=0.168, seed=0, device=None, group=None, dtype=torch.complex128,
                  exchange="auto"):
        """Random shards of the given global shape generated directly on this rank (benchmark input:
        no rank ever materialises the full environments).  `density` is the fraction of non-zero
        MPO-tensor entries (16.8 % at the centre of the 32-orbital molecular MPO)."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        device = device if device is not None else dev.default_device()
        gen = torch.Generator(device=device).manual_seed(seed + 7919 * rank)
        k0, k1 = bond_partition(cl, world)[rank]
        kg = k1 - k0
        P = -(-cr // world)
        scale = 1.0 / np.sqrt(Dl)
        l_shard = torch.randn((Dl, kg, Dl), dtype=dtype, device=device, generator=gen) * scale
        r_shard = torch.randn((Dr, P, Dr), dtype=dtype, device=device, generator=gen) * scale
        q1 = min((rank + 1) * P, cr) - rank * P
        if q1 < P:
            r_shard[:, max(q1, 0):, :] = 0
        w_blocks = []
        for gp in range(world):
            blk = torch.randn((kg, d, d, P), dtype=torch.float64, device=device, generator=gen)
            mask = torch.rand((kg, d, d, P), device=device, generator=gen) < density
            blk = blk * mask
            p1 = min((gp + 1) * P, cr) - gp * P
            if p1 < P:
                blk[:, :, :, max(p1, 0):] = 0
            w_blocks.append(blk.reshape(kg * d, d * P).contiguous())
        return cls(w_blocks, l_shard, r_shard, (Dl, d, Dr, d, Dl, Dr, kg, P), group=group, exchange=exchange)

    def _buffers(self, like):
        n1 = (self.Dl, self.d_in * self.P, self.Drp)
        if self._t1 is None or self._t1.dtype != like.dtype:
            self._t2 = torch.empty((self.Dl, self.kg * self.d_out, self.Drp), dtype=like.dtype, device=like.device)
            self.exchange = "nccl" if self.world > 1 else "local"
            self._t1 = torch.empty(n1, dtype=like.dtype, device=like.device)
            self._gathered = torch.empty((self.world,) + n1, dtype=like.dtype, device=like.device)
        return self._t1, self._gathered, self._t2

    def matvec(self, a):
        """Apply the sharded effective Hamiltonian to `a` (Dl, d, Dr); every rank gets the full result."""
        assert tuple(a.shape) == (self.Dl, self.d_in, self.Dr)
        if a.dtype.is_complex and not self.l_shard.dtype.is_complex:
            # a complex state on real shards: promote the operator once (never demote the state)
            self.l_shard = self.l_shard.to(torch.complex128)
            self.r_shard = self.r_shard.to(torch.complex128)
            self._t1 = self._gathered = self._t2 = None
        a = a.to(self.l_shard.dtype) if a.dtype != self.l_shard.dtype else a
        a = a.contiguous()
        t1, gathered, t2 = self._buffers(a)
        a2d = a.reshape(self.Dl * self.d_in, self.Dr)
        r2d = self.r_shard.reshape(self.Dr, self.P * self.Drp)
        # step 1 on this rank's kappa range:  t1[(i,s),(kappa_loc,j')] = a r_g
        self.ops.step1(a2d, r2d, t1.reshape(self.Dl * self.d_in, self.P * self.Drp))
        # exchange: every rank receives all kappa ranges of t1
        if self.world > 1:
            dist.all_gather_into_tensor(_flat_real(gathered), _flat_real(t1), group=self.group)
        else:
            gathered = t1.reshape((1,) + tuple(t1.shape))
        # step 2 on this rank's k range, accumulating over the source ranges
        for gp in range(self.world):
            # gathered[gp] is [i, (s, kappa_loc), j'] because t1 rows are (i, s) and columns (kappa_loc, j')
            self.ops.wapply(self.w_blocks[gp], gathered[gp], t2, accumulate=(gp > 0))
        # step 3 on this rank's k range, then sum the partial results
        out = torch.empty((self.Dlp, self.d_out, self.Drp), dtype=a.dtype, device=a.device)
        if self.kg > 0:
            self.ops.step3(self.l_shard.reshape(self.Dl * self.kg, self.Dlp),
                           t2.reshape(self.Dl * self.kg, self.d_out * self.Drp),
                           out.reshape(self.Dlp, self.d_out * self.Drp))
        else:
            out.zero_()
        if self.world > 1:
            dist.all_reduce(_flat_real(out), op=dist.ReduceOp.SUM, group=self.group)
        return out

    # ---- accounting used by bench.py / DESIGN.md -------------------------------------------
    def flops_per_rank(self):
        """Executed complex-MAC-equivalent flops of one matvec on this rank (dense count, real W at half)."""
        s1 = self.Dl * self.d_in * self.Dr * self.P * self.Drp
        s2 = self.kg * self.d_out * self.d_in * self.P * self.world * self.Dl * self.Drp
        s3 = self.Dlp * self.Dl * self.kg * self.d_out * self.Drp
        return 8.0 * (s1 + s3) + 4.0 * s2

    def exchange_bytes_per_rank(self):
        es = self.l_shard.element_size()
        gather = (self.world - 1) * self.Dl * self.d_in * self.P * self.Drp * es
        reduce_ = 2.0 * (self.world - 1) / max(self.world, 1) * self.Dlp * self.d_out * self.Drp * es
        return gather, reduce_


class PrecontractedShardedHamiltonian:
    r"""
    MPO-bond-sharded effective Hamiltonian whose only communication is the all-reduce of the result.

    The right MPO bond is split into G ranges.  Once per site (environments and MPO tensor are fixed
    during a Lanczos run) rank g contracts the left environment with its slice of the MPO tensor,

        LW_g[i, s, kappa, s', i'] = sum_k  w[k, s', s, kappa] l[i, k, i']        (kappa in range g),

    and every matvec is then two GEMMs on that rank followed by one all-reduce:

        t1_g[(i,s,kappa), j'] = a[(i,s), j] r[j, (kappa, j')]                    (kappa in range g)
        out_g[i', s', j']     = sum_{(i,s,kappa)} LW_g[(i,s,kappa), s', i'] t1_g[(i,s,kappa), j']
        out                   = all-reduce(sum_g out_g)                           (D*d*D elements)

    Per-rank flops are 8 (Dl d Dr P Dr' + Dl' d' Dr' Dl d P) with P = chi_r / G: both GEMMs scale with
    1/G and no intermediate ever crosses NVLink.  Compared with the three-step contraction this trades
    the W step for a d-times larger final GEMM (about 12 % more flops at the 32-orbital molecular
    shape), which is why the one-GPU path keeps the three-step form and the sharded path uses this one.
    """

    def __init__(self, lw, r_shard, dims, group=None, ops=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.ops = ops if ops is not None else _CudaOps()
        self.lw = lw                      # (Dl * d_in * P, d_out, Dlp)
        self.r_shard = r_shard            # (Dr, P, Drp), zero padded
        (self.Dl, self.d_in, self.Dr, self.d_out, self.Dlp, self.Drp, self.P) = dims
        self._t1 = None
        self.exchange = "allreduce-only"

    @staticmethod
    def _w3(w_t, rank, P, cr):
        """W3[(s, kappa_loc, s'), k] = w[k, s', s, kappa] for this rank's zero-padded kappa range."""
        cl, dout, din, _ = w_t.shape
        q0, q1 = rank * P, min((rank + 1) * P, cr)
        blk = torch.zeros((din, P, dout, cl), dtype=w_t.dtype, device=w_t.device)
        if q1 > q0:
            blk[:, :q1 - q0] = w_t[:, :, :, q0:q1].permute(2, 3, 1, 0)
        return blk.reshape(din * P * dout, cl).contiguous()

    @classmethod
    def from_full(cls, w, l, r, group=None, ops=None, device=None):
        """Every rank passes the full (w, l, r); the rank keeps r[:, kappa_g, :] and builds LW_g."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        ops = ops if ops is not None else _CudaOps()

        def tens(x):
            x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            return x.to(device) if device is not None else x

        w_t, l_t, r_t = tens(w), tens(l), tens(r)
        w_t = w_t.to(l_t.device)
        cl, dout, din, cr = w_t.shape
        Dl, _, Dlp = l_t.shape
        Dr, _, Drp = r_t.shape
        cplx = l_t.dtype.is_complex or r_t.dtype.is_complex or w_t.dtype.is_complex
        dt = torch.complex128 if cplx else torch.float64
        l_t = l_t.to(dt).contiguous()
        P = -(-cr // world)
        r_pad = torch.zeros((Dr, P, Drp), dtype=dt, device=l_t.device)
        q0, q1 = rank * P, min((rank + 1) * P, cr)
        if q1 > q0:
            r_pad[:, :q1 - q0, :] = r_t[:, q0:q1, :].to(dt)
        w3 = cls._w3(w_t, rank, P, cr)
        lw = torch.empty((Dl, din * P * dout, Dlp), dtype=dt, device=l_t.device)
        ops.precontract(w3, l_t, lw)
        return cls(lw.reshape(Dl * din * P, dout, Dlp), r_pad, (Dl, din, Dr, dout, Dlp, Drp, P), group=group, ops=ops)

    @classmethod
    def synthetic(cls, Dl, d, Dr, cl, cr, density=0.168, seed=0, device=None, group=None, dtype=torch.complex128):
        """Random shards generated on this rank (benchmark input): the left environment is drawn in
        full (it is what a site update would hold after its all-gather), LW_g is built from it by the
        same precontraction GEMM a real run uses.  Returns (operator, precontraction milliseconds)."""
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        device = device if device is not None else dev.default_device()
        gen = torch.Generator(device=device).manual_seed(seed + 7919 * rank)
        P = -(-cr // world)
        scale = 1.0 / np.sqrt(Dl)
        l_full = torch.randn((Dl, cl, Dl), dtype=dtype, device=device, generator=gen) * scale
        r_shard = torch.randn((Dr, P, Dr), dtype=dtype, device=device, generator=gen) * scale
        q1 = min((rank + 1) * P, cr) - rank * P
        if q1 < P:
            r_shard[:, max(q1, 0):, :] = 0
        w3 = torch.randn((d * P * d, cl), dtype=torch.float64, device=device, generator=gen)
        w3 = w3 * (torch.rand((d * P * d, cl), device=device, generator=gen) < density)
        if q1 < P:
            w3.reshape(d, P, d, cl)[:, max(q1, 0):] = 0
        ops = _CudaOps()
        lw = torch.empty((Dl, d * P * d, Dl), dtype=dtype, device=device)
        torch.cuda.synchronize(device)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.precontract(w3.contiguous(), l_full, lw)
        e1.record()
        torch.cuda.synchronize(device)
        setup_ms = e0.elapsed_time(e1)
        del l_full
        return cls(lw.reshape(Dl * d * P, d, Dl), r_shard, (Dl, d, Dr, d, Dl, Dr, P), group=group, ops=ops), setup_ms

    def matvec(self, a):
        """Apply the sharded effective Hamiltonian to `a` (Dl, d, Dr); every rank gets the full result."""
        assert tuple(a.shape) == (self.Dl, self.d_in, self.Dr)
        if a.dtype.is_complex and not self.lw.dtype.is_complex:
            # a complex state on a real operator: promote the operator once (never demote the state)
            self.lw = self.lw.to(torch.complex128)
            self.r_shard = self.r_shard.to(torch.complex128)
            self._t1 = None
        a = a.to(self.lw.dtype) if a.dtype != self.lw.dtype else a
        a = a.contiguous()
        rows = self.Dl * self.d_in
        if self._t1 is None or self._t1.dtype != a.dtype:
            self._t1 = torch.empty((rows, self.P * self.Drp), dtype=a.dtype, device=a.device)
        # step 1 on this rank's kappa range; its row-major memory is also [(i, s, kappa_loc), j']
        self.ops.step1(a.reshape(rows, self.Dr), self.r_shard.reshape(self.Dr, self.P * self.Drp), self._t1)
        out = torch.empty((self.Dlp, self.d_out, self.Drp), dtype=a.dtype, device=a.device)
        self.ops.contract_lw(self.lw, self._t1.reshape(rows * self.P, self.Drp), out)
        if self.world > 1:
            dist.all_reduce(_flat_real(out), op=dist.ReduceOp.SUM, group=self.group)
        return out

    def flops_per_rank(self):
        s1 = self.Dl * self.d_in * self.Dr * self.P * self.Drp
        s2 = self.Dlp * self.d_out * self.Drp * self.Dl * self.d_in * self.P
        return 8.0 * (s1 + s2)

    def exchange_bytes_per_rank(self):
        es = self.lw.element_size()
        return 0, 2.0 * (self.world - 1) / max(self.world, 1) * self.Dlp * self.d_out * self.Drp * es
