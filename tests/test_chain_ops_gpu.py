"""GPU: the four chain contractions through the C ABI against the oracle and the
golden fixtures of the reference.  Tolerance 1e-12 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

TOL = 1e-12


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def rnd(rng, shape, cplx):
    x = rng.normal(size=shape)
    if cplx:
        x = (x + 1j * rng.normal(size=shape)) / np.sqrt(2)
    return x


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_golden_fixtures_device_and_host_entry(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "chain_ops.npz"))
    for name in z["names"]:
        g = lambda k: z[f"{name}/{k}"]      # noqa: E731
        a, b, w, l, r, c, rb = (g(k) for k in "a b w l r c rb".split())
        # device-resident entry
        out = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r))
        assert isinstance(out, torch.Tensor) and out.is_cuda
        assert rel(out.cpu().numpy(), g("ref_hv")) < TOL, name
        assert rel(ptb.apply_local_bond_contraction(cu(c), cu(l), cu(rb)).cpu().numpy(), g("ref_bond")) < TOL, name
        assert rel(ptb.contraction_operator_step_right(cu(a), cu(b), cu(w), cu(r)).cpu().numpy(), g("ref_sr")) < TOL
        assert rel(ptb.contraction_operator_step_left(cu(a), cu(b), cu(w), cu(l)).cpu().numpy(), g("ref_sl")) < TOL
        # host-buffer entry (NumPy in, NumPy out): the e2e call
        out = ptb.apply_local_hamiltonian(a, w, l, r)
        assert isinstance(out, np.ndarray) and out.dtype == g("ref_hv").dtype and out.shape == g("ref_hv").shape
        assert rel(out, g("ref_hv")) < TOL


SHAPES = [
    # Dl, d, Dr, chil, chir, Dl', Dr'
    (1, 2, 1, 1, 1, 1, 1),
    (1, 2, 2, 1, 4, 1, 2),
    (2, 2, 4, 4, 5, 2, 4),
    (16, 2, 28, 5, 5, 16, 28),          # config 1 largest site
    (28, 2, 16, 5, 5, 28, 16),
    (31, 3, 17, 4, 6, 29, 19),
    (64, 4, 64, 5, 5, 64, 64),          # two-site Heisenberg
    (48, 16, 40, 6, 6, 48, 40),         # two-site Fermi-Hubbard
    (96, 2, 96, 37, 41, 96, 96),        # large MPO bond (molecular-like)
    (130, 2, 257, 5, 5, 131, 255),
]


@pytest.mark.parametrize("cplx_state", [True, False])
@pytest.mark.parametrize("cplx_w", [False, True])
def test_seeded_shapes_against_oracle(cuda_lib, cplx_state, cplx_w):
    import pytenet_b200 as ptb
    rng = np.random.default_rng(11 + 2 * cplx_state + cplx_w)
    for (Dl, d, Dr, cl, cr, Dlp, Drp) in SHAPES:
        a = rnd(rng, (Dl, d, Dr), cplx_state)
        b = rnd(rng, (Dlp, d, Drp), cplx_state)
        w = rnd(rng, (cl, d, d, cr), cplx_w)
        w[rng.random(w.shape) < 0.7] = 0
        l = rnd(rng, (Dl, cl, Dlp), cplx_state)
        r = rnd(rng, (Dr, cr, Drp), cplx_state)
        c = rnd(rng, (Dl, Dr), cplx_state)
        rb = rnd(rng, (Dr, cl, Drp), cplx_state)
        tag = (Dl, d, Dr, cl, cr, Dlp, Drp)
        got = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
        want = oracle.apply_local_hamiltonian(a, w, l, r)
        assert got.dtype == want.dtype and got.shape == want.shape
        assert rel(got, want) < TOL, ("hv", tag)
        got = ptb.apply_local_bond_contraction(cu(c), cu(l), cu(rb)).cpu().numpy()
        assert rel(got, oracle.apply_local_bond_contraction(c, l, rb)) < TOL, ("bond", tag)
        got = ptb.contraction_operator_step_right(cu(a), cu(b), cu(w), cu(r)).cpu().numpy()
        assert rel(got, oracle.contraction_operator_step_right(a, b, w, r)) < TOL, ("sr", tag)
        got = ptb.contraction_operator_step_left(cu(a), cu(b), cu(w), cu(l)).cpu().numpy()
        assert rel(got, oracle.contraction_operator_step_left(a, b, w, l)) < TOL, ("sl", tag)


def test_integer_dummy_edge_block(cuda_lib):
    """The reference seeds environments with the int64 block [[[1]]] (chain_ops.py:110)."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(3)
    a = rnd(rng, (4, 2, 1), True); w = rnd(rng, (5, 2, 2, 1), False)
    r = np.array([[[1]]])
    got = ptb.contraction_operator_step_right(a, a, w, r)
    want = oracle.contraction_operator_step_right(a, a, w, r)
    assert got.shape == want.shape == (4, 5, 4) and rel(got, want) < TOL


def test_block_sparse_mpo_inner_product(cuda_lib, golden_dir):
    """reference test_chain_ops.py:32-65 pattern (a != b, quantum numbers), seeded fixture."""
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "mpo_inner.npz"))
    n = int(z["nsites"])
    D = z[f"psi{n-1}"].shape[2]
    t = torch.eye(D, dtype=torch.complex128, device="cuda").reshape(D, 1, D)
    for i in reversed(range(n)):
        t = ptb.contraction_operator_step_right(cu(z[f"psi{i}"]), cu(z[f"chi{i}"]), cu(z[f"op{i}"]), t)
    assert tuple(t.shape) == (1, 1, 1)
    val = t.reshape(-1)[0].item()
    assert abs(val - complex(z["value"])) / abs(complex(z["value"])) < TOL


def test_shape_errors_raise_like_the_reference(cuda_lib):
    import pytenet_b200 as ptb
    a = np.zeros((2, 2, 2)); w = np.zeros((3, 2, 2, 3)); l = np.zeros((2, 3, 2)); r = np.zeros((2, 3, 2))
    with pytest.raises(AssertionError):
        ptb.apply_local_hamiltonian(a[0], w, l, r)          # rank assert, chain_ops.py:268
    with pytest.raises(AssertionError):
        ptb.apply_local_hamiltonian(a, w, l, np.zeros((3, 3, 2)))


def test_hermiticity_and_linearity_at_scale(cuda_lib):
    """Size-independent properties at a bench-like size (D=512, d=4, chi=5): with
    Hermitian environments / MPO the effective Hamiltonian is Hermitian, <x|H y> = <H x|y>,
    and H(x + 2y) = Hx + 2Hy."""
    import pytenet_b200 as ptb
    D, d, chi = 512, 4, 5
    g = torch.Generator(device="cuda").manual_seed(5)
    def rc(*s):
        return torch.randn(*s, dtype=torch.complex128, device="cuda", generator=g)
    l = rc(D, chi, D); r = rc(D, chi, D)
    l = l + l.conj().permute(2, 1, 0); r = r + r.conj().permute(2, 1, 0)
    w = torch.randn(chi, d, d, chi, dtype=torch.float64, device="cuda", generator=g)
    w = w + w.permute(0, 2, 1, 3)
    x = rc(D, d, D); y = rc(D, d, D)
    hx = ptb.apply_local_hamiltonian(x, w, l, r)
    hy = ptb.apply_local_hamiltonian(y, w, l, r)
    lhs = torch.vdot(x.reshape(-1), hy.reshape(-1))
    rhs = torch.vdot(hx.reshape(-1), y.reshape(-1))
    assert abs((lhs - rhs).item()) / abs(lhs.item()) < 1e-11
    hxy = ptb.apply_local_hamiltonian(x + 2 * y, w, l, r)
    assert (torch.linalg.norm(hxy - (hx + 2 * hy)) / torch.linalg.norm(hxy)).item() < 1e-13


def test_sparse_and_dense_w_step_agree(cuda_lib):
    """The CSR W kernel (sparse MPO tensors) and the GEMM W step (dense / large MPO tensors) give the same
    matvec; both against the oracle.  The dense path is forced by lowering the CSR threshold."""
    import pytenet_b200 as ptb
    from pytenet_b200 import _device as dev
    rng = np.random.default_rng(41)
    Dl, d, Dr, cl, cr = 70, 3, 66, 9, 8
    a = rnd(rng, (Dl, d, Dr), True); l = rnd(rng, (Dl, cl, Dl), True); r = rnd(rng, (Dr, cr, Dr), True)
    for cplx_w in (False, True):
        w = rnd(rng, (cl, d, d, cr), cplx_w)
        w[rng.random(w.shape) < 0.85] = 0
        want = oracle.apply_local_hamiltonian(a, w, l, r)
        assert dev.w_csr(cu(w)) is not None
        got_sparse = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
        old = dev._CSR_MAX_NNZ
        try:
            dev._CSR_MAX_NNZ = 0
            dev._csr_cache.clear()
            got_dense = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
        finally:
            dev._CSR_MAX_NNZ = old
            dev._csr_cache.clear()
        assert rel(got_sparse, want) < TOL and rel(got_dense, want) < TOL
    # real state with a real sparse W
    ar, lr, rr = rnd(rng, (Dl, d, Dr), False), rnd(rng, (Dl, cl, Dl), False), rnd(rng, (Dr, cr, Dr), False)
    w = rnd(rng, (cl, d, d, cr), False); w[rng.random(w.shape) < 0.85] = 0
    got = ptb.apply_local_hamiltonian(cu(ar), cu(w), cu(lr), cu(rr)).cpu().numpy()
    assert got.dtype == np.float64 and rel(got, oracle.apply_local_hamiltonian(ar, w, lr, rr)) < TOL


def _host_call(ptb, a, w, l, r):
    from pytenet_b200 import chain_ops
    old = chain_ops._HOST_OVERLAP_MIN_BYTES
    try:
        chain_ops._HOST_OVERLAP_MIN_BYTES = 0
        return ptb.apply_local_hamiltonian(a, w, l, r)
    finally:
        chain_ops._HOST_OVERLAP_MIN_BYTES = old


@pytest.mark.parametrize("cplx,w_cplx,dense_w", [(True, False, False), (False, False, False), (True, True, False),
                                                 (True, False, True)])
def test_host_entry_small(cuda_lib, cplx, w_cplx, dense_w):
    """Host-buffer form of the C ABI (ptb_apply_local_hamiltonian_host) on ragged shapes: one slice, CSR and
    dense W step, real / complex state and MPO tensor, pageable NumPy inputs."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(17)
    Dl, d, Dr, Dlp, Drp = 160, 2, 150, 141, 133
    cl, cr = (48, 45) if dense_w else (5, 4)
    a = rnd(rng, (Dl, d, Dr), cplx); l = rnd(rng, (Dl, cl, Dlp), cplx); r = rnd(rng, (Dr, cr, Drp), cplx)
    w = rnd(rng, (cl, d, d, cr), w_cplx)
    if not dense_w:
        w[rng.random(w.shape) < 0.6] = 0
    got = _host_call(ptb, a, w, l, r)
    want = oracle.apply_local_hamiltonian(a, w, l, r)
    assert isinstance(got, np.ndarray) and got.dtype == want.dtype and got.shape == want.shape
    assert rel(got, want) < TOL


@pytest.mark.parametrize("pinned", [False, True])
def test_host_entry_sliced_pipeline(cuda_lib, pinned):
    """Sizes at which the host form slices step 1 along the contraction index (2-D copies of `a`, accumulated
    GEMMs) and step 3 along the rows of `out` (copied back block by block); ragged extents; page-locked and
    pageable inputs; the dummy int64 edge block of chain_ops.py:110 is not involved here."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(23)
    Dl, d, Dr, cl, cr, Dlp, Drp = 700, 2, 777, 5, 5, 1100, 650
    a = rnd(rng, (Dl, d, Dr), True); l = rnd(rng, (Dl, cl, Dlp), True); r = rnd(rng, (Dr, cr, Drp), True)
    w = rnd(rng, (cl, d, d, cr), False); w[rng.random(w.shape) < 0.8] = 0
    if pinned:
        keep = []
        def pin(x):
            t = torch.empty(x.shape, dtype=torch.from_numpy(x).dtype, pin_memory=True)
            t.numpy()[...] = x
            keep.append(t)
            return t.numpy()
        a, l, r = pin(a), pin(l), pin(r)
    got = _host_call(ptb, a, w, l, r)
    assert rel(got, oracle.apply_local_hamiltonian(a, w, l, r)) < TOL
    # twice in a row on the same workspace (stream ordering of the internal copy streams)
    got2 = _host_call(ptb, a, w, l, r)
    assert np.array_equal(got, got2)


def test_config2_shape_direct_parity_with_oracle(cuda_lib):
    """Direct parity at BASELINE config-2 size (two-site XXZ, a (1024,4,1024), l/r (1024,5,1024)): the device
    matvec against the CPU oracle on the same seeded inputs (the oracle needs < 1 s here), 1e-12 relative."""
    import pytenet_b200 as ptb
    from bench import host_inputs
    a, w, l, r = host_inputs(1024, 4, 5, seed=2026)
    want = oracle.apply_local_hamiltonian(a, w, l, r)
    got = ptb.apply_local_hamiltonian(cu(a), cu(w), cu(l), cu(r)).cpu().numpy()
    assert rel(got, want) < TOL
    # environment updates at the same size (single-site tensors)
    a1 = a[:, :2, :].copy(); w1 = w[:, :2, :2, :].copy()
    got = ptb.contraction_operator_step_left(cu(a1), cu(a1), cu(w1), cu(l)).cpu().numpy()
    assert rel(got, oracle.contraction_operator_step_left(a1, a1, w1, l)) < TOL
    got = ptb.contraction_operator_step_right(cu(a1), cu(a1), cu(w1), cu(r)).cpu().numpy()
    assert rel(got, oracle.contraction_operator_step_right(a1, a1, w1, r)) < TOL
    c = a[:, 0, :].copy()
    got = ptb.apply_local_bond_contraction(cu(c), cu(l), cu(r)).cpu().numpy()
    assert rel(got, oracle.apply_local_bond_contraction(c, l, r)) < TOL


def test_headline_shape_direct_parity_with_oracle(cuda_lib):
    """Direct parity at the HEADLINE shape the bench number is quoted on (two-site XXZ, a (2048,4,2048),
    l/r (2048,5,2048), complex128): all four contractions against the CPU oracle on the same seeded inputs, 1e-12
    relative.  The oracle needs ~2.5 s per two-site matvec on 16 host cores (BENCH_r01 cpu_baseline), about a
    minute for everything here on 8."""
    import pytenet_b200 as ptb
    from bench import host_inputs
    a, w, l, r = host_inputs(2048, 4, 5, seed=2048)
    ld, rd, wd = cu(l), cu(r), cu(w)
    got = ptb.apply_local_hamiltonian(cu(a), wd, ld, rd).cpu().numpy()
    want = oracle.apply_local_hamiltonian(a, w, l, r)
    assert rel(got, want) < TOL
    # the same call through the host-buffer C entry (sliced copy / compute pipeline)
    assert rel(ptb.apply_local_hamiltonian(a, w, l, r), want) < TOL
    del got, want
    # single-site tensors of the same bond dimension: both environment updates and the zero-site contraction
    a1 = np.ascontiguousarray(a[:, :2, :]); w1 = np.ascontiguousarray(w[:, :2, :2, :])
    got = ptb.contraction_operator_step_left(cu(a1), cu(a1), cu(w1), ld).cpu().numpy()
    assert rel(got, oracle.contraction_operator_step_left(a1, a1, w1, l)) < TOL
    got = ptb.contraction_operator_step_right(cu(a1), cu(a1), cu(w1), rd).cpu().numpy()
    assert rel(got, oracle.contraction_operator_step_right(a1, a1, w1, r)) < TOL
    c = np.ascontiguousarray(a[:, 0, :])
    got = ptb.apply_local_bond_contraction(cu(c), ld, rd).cpu().numpy()
    assert rel(got, oracle.apply_local_bond_contraction(c, l, r)) < TOL
    got = ptb.apply_local_hamiltonian(cu(a1), cu(w1), ld, rd).cpu().numpy()
    assert rel(got, oracle.apply_local_hamiltonian(a1, w1, l, r)) < TOL


def test_headline_shape_properties(cuda_lib):
    """BASELINE headline shape (two-site XXZ, a (2048,4,2048), l/r (2048,5,2048), complex128): size-independent
    properties of the device matvec on top of the direct oracle comparison above -- Hermiticity <x|H y> = <H x|y> for
    Hermitian environments, linearity, and agreement of the device-resident path with the host-buffer C entry
    (sliced copy/compute pipeline, accumulated step-1 slices)."""
    import pytenet_b200 as ptb
    from bench import xxz_two_site_w
    D, d, chi = 2048, 4, 5
    g = torch.Generator(device="cuda").manual_seed(7)

    def crand(*shape):
        return torch.randn(*shape, dtype=torch.complex128, device="cuda", generator=g)

    def herm_env():
        e = crand(D, chi, D)
        return (e + e.conj().permute(2, 1, 0)).contiguous()       # e[:, k, :] Hermitian for every k

    l, r = herm_env(), herm_env()
    w = xxz_two_site_w()
    w = np.ascontiguousarray(w + w.transpose(0, 2, 1, 3))           # real, symmetric in (s', s) for every (k, kappa):
    #                                                                 with Hermitian l[:, k, :], r[:, kappa, :] H_eff is Hermitian
    wd = torch.from_numpy(w).cuda()
    x, y = crand(D, d, D), crand(D, d, D)
    hx = ptb.apply_local_hamiltonian(x, wd, l, r)
    hy = ptb.apply_local_hamiltonian(y, wd, l, r)
    lhs = torch.vdot(x.reshape(-1), hy.reshape(-1)).item()
    rhs = torch.vdot(hx.reshape(-1), y.reshape(-1)).item()
    scale = (torch.linalg.norm(x) * torch.linalg.norm(hy)).item()
    assert abs(lhs - rhs) < 1e-12 * scale
    al, be = 0.3 - 1.1j, -0.7 + 0.2j
    hz = ptb.apply_local_hamiltonian(al * x + be * y, wd, l, r)
    assert (torch.linalg.norm(hz - (al * hx + be * hy)) / torch.linalg.norm(hz)).item() < 1e-13
    del hz, hy, y
    got = ptb.apply_local_hamiltonian(x.cpu().numpy(), w, l.cpu().numpy(), r.cpu().numpy())
    assert isinstance(got, np.ndarray)
    assert rel(got, hx.cpu().numpy()) < 1e-13


@pytest.mark.parametrize("cplx,w_cplx", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("dims", [(16, 2, 28, 5, 5, 2, 16, 28), (1, 2, 2, 1, 4, 2, 1, 2), (33, 3, 17, 4, 6, 3, 29, 21),
                                  (64, 2, 64, 5, 5, 2, 64, 64), (7, 4, 9, 3, 3, 4, 250, 8), (20, 16, 12, 2, 3, 16, 20, 12)])
def test_fused_small_matvec_kernel(cuda_lib, cplx, w_cplx, dims):
    """The one-kernel matvec of the launch-latency regime (csrc/heff_small.cu), reached through the dense-w C entry
    and through the zero-site entry, against the oracle on ragged small shapes (non-square H, d_out up to 16)."""
    from pytenet_b200 import _lib, _device as dev
    lib = cuda_lib
    Dl, d, Dr, cl, cr, dout, Dlp, Drp = dims
    rng = np.random.default_rng(sum(dims) + int(cplx) + 2 * int(w_cplx))
    a = rnd(rng, (Dl, d, Dr), cplx); l = rnd(rng, (Dl, cl, Dlp), cplx); r = rnd(rng, (Dr, cr, Drp), cplx)
    w = rnd(rng, (cl, dout, d, cr), w_cplx)
    ad, wd, ld, rd = cu(a), cu(w), cu(l), cu(r)
    out = torch.full((Dlp, dout, Drp), 3.0, dtype=ad.dtype, device="cuda")
    dt = 1 if cplx else 0
    nbytes = lib.ptb_apply_local_hamiltonian_workspace_bytes(dt, Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    if cplx:
        st = lib.ptb_apply_local_hamiltonian_z(ad.data_ptr(), wd.data_ptr(), int(w_cplx), ld.data_ptr(), rd.data_ptr(),
                                               out.data_ptr(), Dl, d, Dr, cl, cr, dout, Dlp, Drp, ws.data_ptr(), nbytes,
                                               stream)
    else:
        st = lib.ptb_apply_local_hamiltonian_d(ad.data_ptr(), wd.data_ptr(), ld.data_ptr(), rd.data_ptr(),
                                               out.data_ptr(), Dl, d, Dr, cl, cr, dout, Dlp, Drp, ws.data_ptr(), nbytes,
                                               stream)
    assert st == 0
    assert rel(out.cpu().numpy(), oracle.apply_local_hamiltonian(a, w, l, r)) < TOL
    # zero-site form: c (Dl, Dr), l (Dl, chi, Dlp), r (Dr, chi, Drp)
    import pytenet_b200 as ptb
    c = rnd(rng, (Dl, Dr), cplx); r2 = rnd(rng, (Dr, cl, Drp), cplx)
    got = ptb.apply_local_bond_contraction(cu(c), ld, cu(r2)).cpu().numpy()
    assert rel(got, oracle.apply_local_bond_contraction(c, l, r2)) < TOL
