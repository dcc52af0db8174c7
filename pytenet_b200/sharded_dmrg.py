"""
Single-site DMRG with the MPO virtual bond split across GPUs (BASELINE config 4: molecular
Hamiltonians, MPO bond dimension O(N^2)).

Every environment block is stored sharded over its MPO-bond index: rank g of G holds the contiguous
range g of bond b (zero padded to P_b = ceil(chi_b / G)), so the environment lists cost 1/G of the
reference's memory per GPU (the reference keeps both full lists, dmrg.py:49-51).  The state `psi` and
the MPO are replicated.  One site of a left-to-right sweep is

  1. all-gather of the left block l_i  (once per site; the only bulk transfer, ~1 % of the site's time)
  2. LW_g = sum_k w[k, :, :, kappa_g] l_i[:, k, :]                       precontraction, 1/G of the W flops
  3. Lanczos on  x -> all-reduce( LW_g^T (x r_{i,g}) )                   two GEMMs + one all-reduce per matvec
  4. QR of the optimised tensor (replicated, cuSOLVER)
  5. l_{i+1,g} = a^T (LW_g conj(a))                                      two GEMMs, no communication: the new block
                                                                          is born sharded over the right MPO bond

and the right-to-left sweep is the mirror image, obtained by transposing the site / MPO tensors
(`_mirror_*`) so that the same left-form kernels serve both directions.  All ranks execute the same
Lanczos recursion on identical data (the all-reduce returns identical results everywhere), so psi stays
bit-identical across ranks without further synchronisation.

The arithmetic goes through an `ops` object (default: the CUDA engine, pytenet_b200/sharded.py); the
multi-rank logic of the environment update is covered on CPU/gloo with a test-only ops object.
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _device as dev
from . import _lib
from .mps import MPS, mps_local_orthonormalize_left_qr, mps_local_orthonormalize_right_qr
from .mpo import MPO
from .krylov import eigh_krylov
from .sharded import _CudaOps, _flat_real

__all__ = ["dmrg_singlesite_sharded", "tdvp_singlesite_sharded", "ShardedSite", "shard_env", "gather_env",
           "ptb_communicator", "destroy_communicators"]


_CABI = os.environ.get("PYTENET_B200_CABI_SHARDED", "1") != "0"
_comms = {}          # process group -> opaque ptb_comm handle (ctypes.c_void_p) of the C ABI


def ptb_communicator(group=None):
    """The C ABI's communicator (include/pytenet_b200.h: ptb_comm_*) for the ranks of `group`, created once:
    rank 0 draws the NCCL id through `ptb_comm_unique_id`, torch.distributed only carries the 128 bytes to the other
    ranks, every rank calls `ptb_comm_init` on its current device.  None for a single rank.  With it the sharded
    matvec is ONE C call (two GEMMs + the all-reduce on the same stream) that a host without torch could issue
    just as well."""
    world, rank = _world(group)
    if world == 1:
        return None
    key = id(group) if group is not None else 0
    if key not in _comms:
        lib = _lib.load()
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            _lib.check(lib.ptb_comm_unique_id(buf), "ptb_comm_unique_id")
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        handle = ctypes.c_void_p()
        torch.cuda.synchronize()
        idbuf = ctypes.create_string_buffer(box[0], 128)
        _lib.check(lib.ptb_comm_init(ctypes.byref(handle), world, rank, idbuf), "ptb_comm_init")
        _comms[key] = handle
    return _comms[key]


def destroy_communicators():
    """Destroy the C ABI communicators created by `ptb_communicator` (call before the process group goes away)."""
    lib = _lib.load()
    for handle in _comms.values():
        lib.ptb_comm_destroy(handle)
    _comms.clear()


def _world(group):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_env(block, group=None):
    """This rank's zero-padded MPO-bond range of a full environment block (D, chi, D')."""
    world, rank = _world(group)
    D, chi, Dp = block.shape
    P = -(-chi // world)
    out = torch.zeros((D, P, Dp), dtype=block.dtype, device=block.device)
    q0, q1 = rank * P, min((rank + 1) * P, chi)
    if q1 > q0:
        out[:, :q1 - q0, :] = block[:, q0:q1, :]
    return out


def gather_env(shard, chi, group=None):
    """Full environment block (D, chi, D') from the per-rank shards (D, P, D')."""
    world, _ = _world(group)
    D, P, Dp = shard.shape
    if world == 1:
        return shard[:, :chi, :].contiguous()
    buf = torch.empty((world, D, P, Dp), dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(_flat_real(buf), _flat_real(shard.contiguous()), group=group)
    return buf.permute(1, 0, 2, 3).reshape(D, world * P, Dp)[:, :chi, :].contiguous()


def _mirror_site(a):
    return dev.dense(a.permute(2, 1, 0))


def _mirror_w(w):
    return dev.dense(w.permute(3, 1, 2, 0))


class ShardedSite:
    """
    Left-form sharded local problem: the full "left-like" block `lfull` (D, chi_l, D') is contracted with this
    rank's range of the MPO tensor's right bond; `r_shard` (Dr, P, Dr') is this rank's range of the
    "right-like" block.  Serves the matvec and the environment update of one site.
    """

    def __init__(self, w, lfull, r_shard, group=None, ops=None, cplx=None):
        self.group = group
        self.world, self.rank = _world(group)
        self.ops = ops if ops is not None else _CudaOps()
        cl, dout, din, cr = w.shape
        Dl, cl2, Dlp = lfull.shape
        assert cl2 == cl
        P = -(-cr // self.world)
        assert r_shard.shape[1] == P
        self.dims = (Dl, din, r_shard.shape[0], dout, Dlp, r_shard.shape[2], P)
        # `cplx` = dtype of the state the site will act on (environments start as the real dummy block)
        cplx = bool(cplx) or lfull.dtype.is_complex or r_shard.dtype.is_complex or w.dtype.is_complex
        self.dtype = torch.complex128 if cplx else torch.float64
        lfull = lfull.to(self.dtype).contiguous()
        self.r_shard = r_shard.to(self.dtype).contiguous()
        # W3[(s, kappa_loc, s'), k] = w[k, s', s, kappa] on this rank's zero-padded kappa range
        q0, q1 = self.rank * P, min((self.rank + 1) * P, cr)
        w3 = torch.zeros((din, P, dout, cl), dtype=w.dtype, device=lfull.device)
        if q1 > q0:
            w3[:, :q1 - q0] = w[:, :, :, q0:q1].permute(2, 3, 1, 0)
        lw = torch.empty((Dl, din * P * dout, Dlp), dtype=self.dtype, device=lfull.device)
        # the CUDA engine is driven through the sharded entries of the C ABI (csrc/sharded.cu), communicator
        # included; a test-only `ops` object (CPU / gloo) keeps the step-by-step form
        self.cabi = _CABI and isinstance(self.ops, _CudaOps) and lfull.is_cuda
        w3 = w3.reshape(din * P * dout, cl).contiguous()
        if self.cabi:
            self.lib = _lib.load()
            self.dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
            self.comm = ptb_communicator(group)
            if w3.dtype.is_complex and not cplx:
                raise TypeError("complex MPO tensor on a real state: pass cplx=True")
            w3 = w3 if w3.dtype in (dev.F64, dev.C128) else w3.to(dev.F64)
            _lib.check(self.lib.ptb_sharded_precontract(self.dt, int(w3.dtype.is_complex), w3.data_ptr(),
                                                        lfull.data_ptr(), lw.data_ptr(), Dl, cl, Dlp, din * P * dout,
                                                        dev.stream_ptr(lfull.device)), "ptb_sharded_precontract")
        else:
            self.ops.precontract(w3, lfull, lw)
        self.lw = lw.reshape(Dl * din * P, dout, Dlp)          # [(i, s, kappa_loc), s', i']
        self._t1 = None

    def matvec(self, a):
        """out[i',s',j'] (full, identical on every rank) = L.W.A.R applied to a (Dl, d, Dr)."""
        Dl, d, Dr, dout, Dlp, Drp, P = self.dims
        a = a.to(self.dtype).contiguous()
        if self.cabi:
            # ONE C call: a.r_g, LW_g^T.t1 (split-K) and the all-reduce, all on the current stream
            out = torch.empty((Dlp, dout, Drp), dtype=self.dtype, device=a.device)
            nb = self.lib.ptb_apply_local_hamiltonian_sharded_workspace_bytes(self.dt, Dl, d, Dr, P, dout, Dlp, Drp)
            ws = dev.workspace(nb, a.device, tag="sharded")
            _lib.check(self.lib.ptb_apply_local_hamiltonian_sharded(
                self.comm, self.dt, a.data_ptr(), self.lw.data_ptr(), self.r_shard.data_ptr(), out.data_ptr(),
                Dl, d, Dr, P, dout, Dlp, Drp, ws.data_ptr(), nb, dev.stream_ptr(a.device)),
                "ptb_apply_local_hamiltonian_sharded")
            return out
        if self._t1 is None:
            self._t1 = torch.empty((Dl * d, P * Drp), dtype=self.dtype, device=a.device)
        self.ops.step1(a.reshape(Dl * d, Dr), self.r_shard.reshape(Dr, P * Drp), self._t1)
        out = torch.empty((Dlp, dout, Drp), dtype=self.dtype, device=a.device)
        self.ops.contract_lw(self.lw, self._t1.reshape(Dl * d * P, Drp), out)
        if self.world > 1:
            dist.all_reduce(_flat_real(out), op=dist.ReduceOp.SUM, group=self.group)
        return out

    def next_env_shard(self, a, b=None):
        """This rank's range of the next block:  l_next[j, kappa, j'] = sum l[i,k,i'] conj(b[i',s',j'])
        w[k,s',s,kappa] a[i,s,j]  (pytenet/chain_ops.py:60-99) for kappa in the rank's range; no communication."""
        Dl, d, Dr, dout, Dlp, Drp, P = self.dims
        b = a if b is None else b
        a = a.to(self.dtype).contiguous()
        # the orthonormalised tensor may have a smaller right bond than the one the local problem was solved on
        # (a sector-wise QR keeps min(rows, cols) indices per sector)
        assert a.shape[0] == Dl and a.shape[1] == d and b.shape[0] == Dlp and b.shape[1] == dout
        Dr = a.shape[2]
        if self.cabi:
            b = dev.dense(b.to(self.dtype))
            Drb = b.shape[2]
            nxt = torch.empty((Dr, P, Drb), dtype=self.dtype, device=a.device)
            nb = self.lib.ptb_env_step_left_sharded_workspace_bytes(self.dt, Dl, d, P, Drb)
            ws = dev.workspace(nb, a.device, tag="sharded")
            _lib.check(self.lib.ptb_env_step_left_sharded(
                self.dt, a.data_ptr(), b.data_ptr(), self.lw.data_ptr(), nxt.data_ptr(), Dl, d, Dr, P, dout, Dlp, Drb,
                ws.data_ptr(), nb, dev.stream_ptr(a.device)), "ptb_env_step_left_sharded")
            return nxt
        # conj(b) with rows ordered (s', i') to match the column order of LW
        b2 = dev.dense(b.to(self.dtype).permute(1, 0, 2)).reshape(dout * Dlp, b.shape[2])
        x = torch.empty((Dl * d * P, b.shape[2]), dtype=self.dtype, device=a.device)
        self.ops.env_x(self.lw.reshape(Dl * d * P, dout * Dlp), b2, x)           # X = LW conj(B)
        nxt = torch.empty((Dr, P * b.shape[2]), dtype=self.dtype, device=a.device)
        self.ops.step3(a.reshape(Dl * d, Dr), x.reshape(Dl * d, P * b.shape[2]), nxt)   # a^T X
        return nxt.reshape(Dr, P, b.shape[2])


def _ensure_env_ops(ops):
    """The CUDA ops gain the conj-B GEMM used by the environment update."""
    if not hasattr(ops, "env_x"):
        def env_x(lw2d, b2, out):
            return dev.gemm(lw2d, b2, conj_b=True, out=out)
        ops.env_x = env_x
    return ops


def dmrg_singlesite_sharded(hamiltonian: MPO, psi: MPS, numsweeps: int, numiter_lanczos: int = 25, group=None,
                            ops=None):
    """
    Single-site DMRG (same semantics and return value as `dmrg_singlesite`, pytenet/dmrg.py:22-93) with every
    environment block sharded over its MPO-bond index across the ranks of `group`.  `psi` and `hamiltonian`
    must be identical on all ranks; `psi` is updated in place (identically on every rank).
    """
    ops = _ensure_env_ops(ops if ops is not None else _CudaOps())
    nsites = hamiltonian.nsites
    assert nsites == psi.nsites
    ham = hamiltonian.a
    chi = hamiltonian.bond_dims
    k = numiter_lanczos
    cplx = any(t.dtype.is_complex for t in psi.a) or any(t.dtype.is_complex for t in ham)
    psi.orthonormalize(mode="right")
    device = psi.a[0].device
    one = torch.ones((1, 1, 1), dtype=dev.F64, device=device)

    # right blocks, right to left (chain_ops.py:102-113), each born sharded over bond i + 1
    rshards = [None] * nsites
    rshards[nsites - 1] = shard_env(one, group)
    for i in reversed(range(nsites - 1)):
        rfull = gather_env(rshards[i + 1], chi[i + 2], group)
        # mirrored left form: "left-like" block = r_{i+1}, "right-like" = dummy; only the update is needed
        site = ShardedSite(_mirror_w(ham[i + 1]), rfull, _dummy_right(psi.a[i + 1].shape[0], chi[i + 1], group,
                                                                     rfull.dtype, device), group, ops, cplx)
        rshards[i] = site.next_env_shard(_mirror_site(psi.a[i + 1]))
    lshards = [None] * nsites
    lshards[0] = shard_env(one, group)

    def minimize(site, a_start, mirrored):
        shape = tuple(a_start.shape)

        def afunc(x):
            t = x.reshape(shape)
            if mirrored:
                return _mirror_site(site.matvec(_mirror_site(t))).reshape(-1)
            return site.matvec(t).reshape(-1)

        ev, u = eigh_krylov(afunc, a_start.reshape(-1), k, 1)
        return ev[0], dev.dense(u[:, 0]).reshape(shape)

    en_min = np.zeros(numsweeps)
    for n in range(numsweeps):
        en = 0
        for i in range(nsites - 1):                                            # dmrg.py:65-73
            lfull = gather_env(lshards[i], chi[i], group)
            site = ShardedSite(ham[i], lfull, rshards[i], group, ops, cplx)
            en, psi.a[i] = minimize(site, psi.a[i], mirrored=False)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = mps_local_orthonormalize_left_qr(
                psi.a[i], psi.a[i + 1], psi.qsite, psi.qbonds[i:i + 2])
            lshards[i + 1] = site.next_env_shard(psi.a[i])
        for i in reversed(range(1, nsites)):                                   # dmrg.py:76-84
            rfull = gather_env(rshards[i], chi[i + 1], group)
            site = ShardedSite(_mirror_w(ham[i]), rfull, lshards[i], group, ops, cplx)
            en, psi.a[i] = minimize(site, psi.a[i], mirrored=True)
            psi.a[i], psi.a[i - 1], psi.qbonds[i] = mps_local_orthonormalize_right_qr(
                psi.a[i], psi.a[i - 1], psi.qsite, psi.qbonds[i:i + 2])
            rshards[i - 1] = site.next_env_shard(_mirror_site(psi.a[i]))
        psi.a[0], _, psi.qbonds[0] = mps_local_orthonormalize_right_qr(psi.a[0], one, psi.qsite, psi.qbonds[:2])
        en_min[n] = en
    return en_min


def _dummy_right(D, chi, group, dtype, device):
    """Placeholder right-like shard of the correct shape for sites where only the environment update is used."""
    world, _ = _world(group)
    P = -(-chi // world)
    return torch.zeros((D, P, D), dtype=dtype, device=device)


# ---------------------------------------------------------------------------------------------
# Single-site TDVP on the same sharded environments
# ---------------------------------------------------------------------------------------------

def _sharded_bond_matvec(c, l_shard, r_shard, ops, group):
    """Zero-site matvec out[i',j'] = sum_k l[i,k,i'] c[i,j] r[j,k,j'] (pytenet/chain_ops.py:282-317) with both
    blocks sharded over the SAME MPO bond: every rank contracts its range of k, one all-reduce sums them."""
    Dl, P, Dlp = l_shard.shape
    Dr, P2, Drp = r_shard.shape
    assert P == P2
    dt = torch.complex128 if (c.dtype.is_complex or l_shard.dtype.is_complex or r_shard.dtype.is_complex) \
        else torch.float64
    c = c.to(dt).contiguous()
    if _CABI and isinstance(ops, _CudaOps) and c.is_cuda:
        lib = _lib.load()
        code = _lib.PTB_COMPLEX128 if dt.is_complex else _lib.PTB_REAL64
        ls, rs = dev.dense(l_shard.to(dt)), dev.dense(r_shard.to(dt))
        out = torch.empty((Dlp, Drp), dtype=dt, device=c.device)
        nb = lib.ptb_apply_local_bond_contraction_sharded_workspace_bytes(code, Dl, P, Drp)
        ws = dev.workspace(nb, c.device, tag="sharded")
        _lib.check(lib.ptb_apply_local_bond_contraction_sharded(
            ptb_communicator(group), code, c.data_ptr(), ls.data_ptr(), rs.data_ptr(), out.data_ptr(), Dl, Dr, P, Dlp,
            Drp, ws.data_ptr(), nb, dev.stream_ptr(c.device)), "ptb_apply_local_bond_contraction_sharded")
        return out
    t = torch.empty((Dl, P * Drp), dtype=dt, device=c.device)
    ops.step1(c, r_shard.to(dt).reshape(Dr, P * Drp), t)                                  # t[i,(k,j')]
    out = torch.empty((Dlp, Drp), dtype=dt, device=c.device)
    ops.step3(l_shard.to(dt).reshape(Dl * P, Dlp), t.reshape(Dl * P, Drp), out)             # l^T t
    world, _ = _world(group)
    if world > 1:
        dist.all_reduce(_flat_real(out), op=dist.ReduceOp.SUM, group=group)
    return out


def tdvp_singlesite_sharded(hamiltonian: MPO, psi: MPS, dt, numsteps: int, numiter_lanczos: int = 25, group=None,
                            ops=None):
    """
    Symmetric single-site TDVP (same semantics and return value as `tdvp_singlesite`, pytenet/tdvp.py:26-118)
    with every environment block sharded over its MPO-bond index across the ranks of `group`.  Site steps use
    the precontracted left form (`ShardedSite`), bond steps contract the two shards of the shared bond and
    all-reduce; the right-to-left half of the sweep runs on mirrored tensors.
    """
    from .krylov import expm_krylov
    from .block_sparse_util import qnumber_flatten, block_sparse_qr
    ops = _ensure_env_ops(ops if ops is not None else _CudaOps())
    nsites = hamiltonian.nsites
    assert nsites == psi.nsites
    ham = hamiltonian.a
    chi = hamiltonian.bond_dims
    k = numiter_lanczos
    nrm = psi.orthonormalize(mode="right")
    device = psi.a[0].device
    one = torch.ones((1, 1, 1), dtype=dev.F64, device=device)
    cplx = True if isinstance(dt, complex) else (any(t.dtype.is_complex for t in psi.a)
                                                 or any(t.dtype.is_complex for t in ham))

    rshards = [None] * nsites
    rshards[nsites - 1] = shard_env(one, group)
    for i in reversed(range(nsites - 1)):
        rfull = gather_env(rshards[i + 1], chi[i + 2], group)
        site = ShardedSite(_mirror_w(ham[i + 1]), rfull, _dummy_right(psi.a[i + 1].shape[0], chi[i + 1], group,
                                                                     rfull.dtype, device), group, ops, cplx)
        rshards[i] = site.next_env_shard(_mirror_site(psi.a[i + 1]))
    lshards = [None] * nsites
    lshards[0] = shard_env(one, group)

    def evolve_site(site, a, tau, mirrored):
        shape = tuple(a.shape)

        def afunc(x):
            t = x.reshape(shape)
            if mirrored:
                return _mirror_site(site.matvec(_mirror_site(t))).reshape(-1)
            return site.matvec(t).reshape(-1)

        return expm_krylov(afunc, a.reshape(-1), -tau, k, hermitian=True).reshape(shape)

    def evolve_bond(l_shard, r_shard, c, tau):
        shape = tuple(c.shape)
        return expm_krylov(lambda x: _sharded_bond_matvec(x.reshape(shape), l_shard, r_shard, ops, group).reshape(-1),
                           c.reshape(-1), -tau, k, hermitian=True).reshape(shape)

    for _ in range(numsteps):
        for i in range(nsites - 1):                                            # tdvp.py:68-84
            site = ShardedSite(ham[i], gather_env(lshards[i], chi[i], group), rshards[i], group, ops, cplx)
            psi.a[i] = evolve_site(site, psi.a[i], 0.5 * dt, False)
            b0, d, b1 = psi.a[i].shape
            q, c, psi.qbonds[i + 1] = block_sparse_qr(
                psi.a[i].reshape(b0 * d, b1), qnumber_flatten((psi.qbonds[i], psi.qsite)), psi.qbonds[i + 1])
            psi.a[i] = dev.dense(q.reshape(b0, d, q.shape[1]))
            lshards[i + 1] = site.next_env_shard(psi.a[i])
            c = evolve_bond(lshards[i + 1], rshards[i], dev.dense(c), -0.5 * dt)
            nxt = psi.a[i + 1]
            psi.a[i + 1] = dev.gemm(c, nxt.reshape(nxt.shape[0], -1)).reshape((c.shape[0],) + tuple(nxt.shape[1:]))
        i = nsites - 1                                                         # tdvp.py:87-89
        site = ShardedSite(ham[i], gather_env(lshards[i], chi[i], group), rshards[i], group, ops, cplx)
        psi.a[i] = evolve_site(site, psi.a[i], dt, False)
        for i in reversed(range(1, nsites)):                                   # tdvp.py:92-115
            at = dev.dense(psi.a[i].permute(2, 1, 0))
            b1, d, b0 = at.shape
            q, c, qbond = block_sparse_qr(
                at.reshape(b1 * d, b0), qnumber_flatten((-psi.qbonds[i + 1], psi.qsite)), -psi.qbonds[i])
            psi.qbonds[i] = -qbond
            psi.a[i] = dev.dense(q.reshape(b1, d, q.shape[1]).permute(2, 1, 0))
            site = ShardedSite(_mirror_w(ham[i]), gather_env(rshards[i], chi[i + 1], group), lshards[i], group, ops,
                               cplx)
            rshards[i - 1] = site.next_env_shard(_mirror_site(psi.a[i]))
            c = evolve_bond(lshards[i], rshards[i - 1], dev.dense(c.T), -0.5 * dt)
            prv = psi.a[i - 1]
            psi.a[i - 1] = dev.gemm(prv.reshape(-1, prv.shape[2]), c).reshape(tuple(prv.shape[:2]) + (c.shape[1],))
            site = ShardedSite(_mirror_w(ham[i - 1]), gather_env(rshards[i - 1], chi[i], group), lshards[i - 1],
                               group, ops, cplx)
            psi.a[i - 1] = evolve_site(site, psi.a[i - 1], 0.5 * dt, True)
    return nrm
