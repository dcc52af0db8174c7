"""CPU, world_size 2 and 3 over gloo: the host-side logic of the MPO-bond-sharded effective
Hamiltonian (bond partition, zero-padded shards, packed W blocks, all-gather + all-reduce) with the
arithmetic supplied by a test-only torch-CPU `ops` object, checked against the oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


class CpuOps:
    """Test infrastructure: the three steps with torch CPU matmul (the product default is the CUDA engine)."""

    @staticmethod
    def step1(a2d, r2d, out):
        out.copy_(a2d @ r2d)
        return out

    @staticmethod
    def wapply(wblk, tin, tout, accumulate):
        res = torch.matmul(wblk.to(tin.dtype), tin)
        if accumulate:
            tout += res
        else:
            tout.copy_(res)
        return tout

    @staticmethod
    def step3(l2d, t2d, out):
        out.copy_(l2d.T @ t2d)
        return out

    @staticmethod
    def precontract(w3, l, out):
        out.copy_(torch.matmul(w3.to(l.dtype), l))
        return out

    @staticmethod
    def contract_lw(lw, t1, out):
        out.copy_(torch.einsum("ksm,kn->msn", lw, t1))
        return out

    @staticmethod
    def env_x(lw2d, b2, out):
        out.copy_(lw2d @ b2.conj())
        return out


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, shapes, seed, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pytenet_b200.sharded import ShardedEffectiveHamiltonian
        rng = np.random.default_rng(seed)           # same inputs on every rank
        Dl, d, Dr, cl, cr, Dlp, Drp = shapes

        def crand(*s):
            return rng.normal(size=s) + 1j * rng.normal(size=s)

        a = crand(Dl, d, Dr); l = crand(Dl, cl, Dlp); r = crand(Dr, cr, Drp)
        w = rng.normal(size=(cl, d, d, cr)); w[rng.random(w.shape) < 0.8] = 0
        heff = ShardedEffectiveHamiltonian.from_full(w, torch.from_numpy(l), torch.from_numpy(r), ops=CpuOps())
        out = heff.matvec(torch.from_numpy(a)).numpy()
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        # second call reuses the buffers
        out2 = heff.matvec(torch.from_numpy(2 * a)).numpy()
        err2 = np.linalg.norm(out2 - 2 * ref) / np.linalg.norm(ref)
        # the all-reduce-only (precontracted) variant
        from pytenet_b200.sharded import PrecontractedShardedHamiltonian
        hpre = PrecontractedShardedHamiltonian.from_full(w, torch.from_numpy(l), torch.from_numpy(r), ops=CpuOps())
        out3 = hpre.matvec(torch.from_numpy(a)).numpy()
        err3 = np.linalg.norm(out3 - ref) / np.linalg.norm(ref)
        results[rank] = max(err, err2, err3)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shapes", [
    (2, (6, 2, 5, 7, 9, 6, 5)),          # odd MPO bonds: uneven k ranges, zero-padded kappa shards
    (3, (4, 2, 4, 2, 5, 3, 4)),          # more ranks than left-bond indices on one rank (k_g can be 0 or 1)
    (2, (8, 4, 8, 5, 5, 8, 8)),
])
def test_sharded_matvec_matches_oracle(world, shapes):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shapes, 123, results), nprocs=world, join=True)
    assert len(results) == world
    for rank, err in results.items():
        assert err < 1e-13, (rank, err)


def test_bond_partition():
    from pytenet_b200.sharded import bond_partition
    assert bond_partition(562, 8) == [(0, 71), (71, 142), (142, 212), (212, 282), (282, 352), (352, 422),
                                      (422, 492), (492, 562)]
    assert bond_partition(2, 3) == [(0, 1), (1, 2), (2, 2)]
    assert sum(b - a for a, b in bond_partition(501, 8)) == 501


def _env_worker(rank, world, port, seed, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pytenet_b200.sharded_dmrg import ShardedSite, shard_env, gather_env
        rng = np.random.default_rng(seed)
        Dl, d, Dr, cl, cr = 5, 2, 6, 7, 9

        def crand(*s):
            return rng.normal(size=s) + 1j * rng.normal(size=s)

        a = crand(Dl, d, Dr); l = crand(Dl, cl, Dl); r = crand(Dr, cr, Dr)
        w = rng.normal(size=(cl, d, d, cr)); w[rng.random(w.shape) < 0.7] = 0
        T = torch.from_numpy
        errs = []
        # shard / gather round trip
        rs = shard_env(T(r)); ls = shard_env(T(l))
        errs.append(float(torch.linalg.norm(gather_env(rs, cr) - T(r))))
        # left form: matvec and the next left block (born sharded over the right MPO bond)
        site = ShardedSite(T(w), T(l), rs, ops=CpuOps())
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        errs.append(np.linalg.norm(site.matvec(T(a)).numpy() - ref) / np.linalg.norm(ref))
        lnext = gather_env(site.next_env_shard(T(a)), cr).numpy()
        ref = oracle.contraction_operator_step_left(a, a, w, l)
        errs.append(np.linalg.norm(lnext - ref) / np.linalg.norm(ref))
        # mirrored form: the next right block (sharded over the left MPO bond) from the full right block
        site = ShardedSite(T(np.ascontiguousarray(w.transpose(3, 1, 2, 0))), T(r), ls, ops=CpuOps())
        am = T(np.ascontiguousarray(a.transpose(2, 1, 0)))
        rnext = gather_env(site.next_env_shard(am), cl).numpy()
        ref = oracle.contraction_operator_step_right(a, a, w, r)
        errs.append(np.linalg.norm(rnext - ref) / np.linalg.norm(ref))
        out = site.matvec(am).numpy().transpose(2, 1, 0)
        ref = oracle.apply_local_hamiltonian(a, w, l, r)
        errs.append(np.linalg.norm(out - ref) / np.linalg.norm(ref))
        results[rank] = max(errs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_site_matvec_and_environment_updates(world):
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_env_worker, args=(world, _free_port(), 99, results), nprocs=world, join=True)
    assert len(results) == world
    for rank, err in results.items():
        assert err < 1e-13, (rank, err)
