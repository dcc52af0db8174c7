"""GPU: MPS-level operations next to the hot path -- apply_mpo, mps_add,
MPS.compress (svd / density), MPS.from_vector -- against the fixture generated from the reference
(tests/golden/mps_ops.npz) and the oracle.  Singular / eigen vectors carry a gauge freedom, so the compressed
states are compared through gauge-invariant quantities: the full state vector (1e-9), norms and scale factors
(1e-10), bond dimensions and bond quantum numbers (bit-exact)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def load_mps(ptb, z, tag, n):
    return ptb.MPS.from_tensors(z[f"{tag}/qsite"], [z[f"{tag}/qb{i}"] for i in range(n + 1)],
                                [z[f"{tag}/a{i}"] for i in range(n)])


@pytest.fixture()
def case(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "mps_ops.npz"))
    n = int(z["h/nsites"])
    h = ptb.MPO.from_tensors(z["h/qsite"], [z[f"h/qb{i}"] for i in range(n + 1)], [z[f"h/w{i}"] for i in range(n)])
    return ptb, z, n, h


def test_apply_mpo(case):
    ptb, z, n, h = case
    psi = load_mps(ptb, z, "psi", n)
    hp = ptb.apply_mpo(h, psi)
    assert all(t.is_cuda for t in hp.a)
    for i in range(n):
        assert np.array_equal(hp.qbonds[i], z[f"hpsi/qb{i}"])
        assert rel(hp.a[i].cpu().numpy(), z[f"hpsi/a{i}"]) < 1e-14
    assert rel(hp.to_vector(), z["hpsi/vec"]) < 1e-13
    # <psi| (H psi)> equals the operator average computed on the hot path
    assert abs(ptb.mps_vdot(psi, hp) - ptb.mpo_average(psi, h)) < 1e-12


def test_mps_add(case):
    ptb, z, n, h = case
    psi, chi = load_mps(ptb, z, "psi", n), load_mps(ptb, z, "chi", n)
    alpha = complex(z["add/alpha"])
    sm = ptb.mps_add(psi, chi, alpha)
    for i in range(n):
        assert np.array_equal(sm.qbonds[i], z[f"add/qb{i}"])
        # alpha * a is one complex product per entry: the device fuses multiply-adds, NumPy does not (last-bit)
        got = sm.a[i].cpu().numpy()
        assert got.shape == z[f"add/a{i}"].shape and np.allclose(got, z[f"add/a{i}"], rtol=1e-14, atol=0)
        assert np.array_equal(got == 0, z[f"add/a{i}"] == 0)               # block structure exact
    assert rel((psi - chi).to_vector(), psi.to_vector() - chi.to_vector()) < 1e-13
    assert rel((psi + chi).to_vector(), psi.to_vector() + chi.to_vector()) < 1e-13


@pytest.mark.parametrize("tag,mode,direction", [("svd_l0", "svd", "left"), ("svd_r0", "svd", "right"),
                                                ("svd_l", "svd", "left"), ("svd_r", "svd", "right"),
                                                ("den", "density", "left"), ("den0", "density", "left")])
def test_compress(case, tag, mode, direction):
    ptb, z, n, h = case
    p = load_mps(ptb, z, "hpsi", n)
    nrm, scale = p.compress(float(z[f"cmp/{tag}/tol"]), mode=mode, direction=direction)
    assert abs(nrm - float(z[f"cmp/{tag}/nrm"])) < 1e-10 * nrm
    assert abs(scale - float(z[f"cmp/{tag}/scale"])) < 1e-10
    assert p.bond_dims == list(z[f"cmp/{tag}/bond_dims"])
    for i in range(n + 1):
        assert np.array_equal(p.qbonds[i], z[f"cmp/{tag}/qb{i}"])            # sector layout bit-exact
    assert rel(p.to_vector(), z[f"cmp/{tag}/vec"]) < 1e-9
    assert abs(ptb.mps_norm(p) - 1) < 1e-12


def test_compress_rejects_bad_arguments(case):
    ptb, z, n, h = case
    p = load_mps(ptb, z, "psi", n)
    with pytest.raises(ValueError):
        p.compress(0.0, mode="qr")
    with pytest.raises(ValueError):
        p.compress(0.0, mode="svd", direction="up")


@pytest.mark.parametrize("tag", ["fv0", "fv"])
def test_from_vector(case, tag):
    ptb, z, n, h = case
    m = ptb.MPS.from_vector(3, 5, z["fv/input"], tol=float(z[f"{tag}/tol"]))
    assert m.bond_dims == list(z[f"{tag}/bond_dims"])
    assert rel(m.to_vector(), z[f"{tag}/vec"]) < 1e-12
