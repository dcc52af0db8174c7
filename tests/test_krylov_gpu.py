"""GPU: device-resident Lanczos / Krylov drivers against the golden fixtures of the
reference (tests/golden/krylov.npz) and the reference's own test properties
(test/test_krylov.py:6-122, restated with seeds)."""
import os
import warnings

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(np.asarray(x) - np.asarray(y)) / (n if n > 0 else 1.0)


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def dev_matvec(m):
    from pytenet_b200 import _device as dev
    md = cu(m)
    return lambda x: dev.gemm(md, x.reshape(-1, 1)).reshape(-1)


def test_lanczos_against_golden(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    m = z["m"]; k = len(z["alpha"])
    al, be, V = ptb.lanczos_iteration(dev_matvec(m), cu(z["v0"]), k)
    assert isinstance(V, torch.Tensor) and tuple(V.shape) == z["V"].shape
    Vh = V.cpu().numpy()
    # first vectors agree to rounding; later ones drift with the loss of orthogonality, as in
    # any Lanczos run, so compare the first few directly and pin the rest by the defining relations
    assert rel(al[:6], z["alpha"][:6]) < 1e-12 and rel(be[:6], z["beta"][:6]) < 1e-12
    assert rel(Vh[:, :6], z["V"][:, :6]) < 1e-11
    # reference test_krylov.py:19-30: V^H V = I and V^H A V = T  (rtol 1e-12 there)
    assert np.allclose(Vh.conj().T @ Vh, np.identity(k), rtol=1e-11, atol=1e-11)
    T = np.diag(al) + np.diag(be, 1) + np.diag(be, -1)
    assert np.allclose(Vh.conj().T @ m @ Vh, T, rtol=1e-10, atol=1e-10)


def test_expm_and_eigh_krylov_against_golden(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    m = z["m"]
    ex = ptb.expm_krylov(dev_matvec(m), cu(z["v0"]), complex(z["dt"]), 12, hermitian=True)
    assert rel(ex.cpu().numpy(), z["expm"]) < 1e-11
    ew, eu = ptb.eigh_krylov(dev_matvec(m), cu(z["v0"]), 30, 2)
    assert rel(ew, z["eig_w"]) < 1e-11
    euh = eu.cpu().numpy()
    assert euh.shape == z["eig_u"].shape
    # Ritz vectors are defined up to the sign chosen by eigh of the tridiagonal matrix
    for i in range(2):
        ph = np.vdot(z["eig_u"][:, i], euh[:, i])
        assert abs(abs(ph) - 1) < 1e-9
        assert rel(euh[:, i] * np.conj(ph) / abs(ph), z["eig_u"][:, i]) < 1e-8


def test_host_buffer_entry_matches_oracle(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    m = z["m"]
    ex = ptb.expm_krylov(lambda x: m @ x, z["v0"], complex(z["dt"]), 12, hermitian=True)
    assert isinstance(ex, np.ndarray) and rel(ex, z["expm"]) < 1e-11


def test_breakdown_truncates_and_warns(cuda_lib, golden_dir):
    import pytenet_b200 as ptb
    z = np.load(os.path.join(golden_dir, "krylov.npz"))
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        al, be, V = ptb.lanczos_iteration(dev_matvec(z["low"]), cu(z["vb"]), 8)
    assert any("beta[2] ~= 0 encountered during Lanczos iteration." in str(w.message) for w in rec)
    assert al.shape == z["alpha_b"].shape and be.shape == z["beta_b"].shape and tuple(V.shape) == z["V_b"].shape
    assert rel(al, z["alpha_b"]) < 1e-11 and rel(be, z["beta_b"]) < 1e-11


def test_real_vectors_and_complex_time_step(cuda_lib):
    """float64 state with a complex dt (a real MPS turns complex after the first TDVP step)."""
    import pytenet_b200 as ptb
    rng = np.random.default_rng(9)
    n = 200
    m = rng.normal(size=(n, n)); m = 0.5 * (m + m.T)
    v = rng.normal(size=n)
    dt = 0.1 - 0.3j
    got = ptb.expm_krylov(dev_matvec(m), cu(v), dt, 20, hermitian=True)
    want = oracle.expm_krylov(lambda x: m @ x, v, dt, 20)
    assert got.dtype == torch.complex128 and rel(got.cpu().numpy(), want) < 1e-10
    ew, eu = ptb.eigh_krylov(dev_matvec(m), cu(v), 25, 1)
    ow, ou = oracle.eigh_krylov(lambda x: m @ x, v, 25, 1)
    assert eu.dtype == torch.float64 and abs(ew[0] - ow[0]) < 1e-10


def test_vector_kernels_at_scale(cuda_lib):
    """Size-independent property at a bench-like length (n = 2048*4*2048/8): after one
    ortho step w is orthogonal to v_j, |v_next| = 1 and alpha, beta match torch fp64."""
    from pytenet_b200 import _lib, _device as dev
    lib = cuda_lib
    n = 2048 * 4 * 256
    g = torch.Generator(device="cuda").manual_seed(2)
    w = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
    vj = torch.randn(n, dtype=torch.complex128, device="cuda", generator=g)
    vj = vj / torch.linalg.norm(vj)
    w0 = w.clone()
    scal = torch.zeros(2, dtype=torch.float64, device="cuda")
    vnext = torch.empty_like(w)
    scratch = dev.lanczos_scratch(w.device)
    st = lib.ptb_lanczos_ortho_step_z(n, w.data_ptr(), vj.data_ptr(), None, None, scal.data_ptr(),
                                      scal.data_ptr() + 8, vnext.data_ptr(), scratch.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    assert st == 0
    alpha, beta = scal.cpu().numpy()
    a_ref = torch.vdot(vj, w0).real.item()
    assert abs(alpha - a_ref) < 1e-12 * max(1, abs(a_ref))
    resid = w0 - a_ref * vj
    assert abs(beta - torch.linalg.norm(resid).item()) / beta < 1e-13
    assert abs(torch.linalg.norm(vnext).item() - 1) < 1e-13
    assert abs(torch.vdot(vj, vnext).real.item()) < 1e-12


@pytest.mark.parametrize("n", [1, 7, 896, 8192, 8193, 40000])
def test_ortho_step_small_and_large_kernels(cuda_lib, n):
    """The single-CTA form (n*2 <= 16384 doubles) and the grid form of the three-term step against NumPy,
    including the previous-vector term and aliasing v_next == w."""
    from pytenet_b200 import _device as dev
    lib = cuda_lib
    rng = np.random.default_rng(n)
    w = rng.normal(size=n) + 1j * rng.normal(size=n)
    vj = rng.normal(size=n) + 1j * rng.normal(size=n); vj /= np.linalg.norm(vj)
    vm = rng.normal(size=n) + 1j * rng.normal(size=n); vm /= np.linalg.norm(vm)
    bprev = 0.37
    wd, vjd, vmd = cu(w), cu(vj), cu(vm)
    scal = torch.tensor([0.0, 0.0, bprev], dtype=torch.float64, device="cuda")
    scratch = dev.lanczos_scratch(wd.device)
    st = lib.ptb_lanczos_ortho_step_z(n, wd.data_ptr(), vjd.data_ptr(), vmd.data_ptr(), scal.data_ptr() + 16,
                                      scal.data_ptr(), scal.data_ptr() + 8, wd.data_ptr(), scratch.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    assert st == 0
    alpha, beta, _ = scal.cpu().numpy()
    a_ref = np.vdot(vj, w).real
    res = w - (a_ref * vj + bprev * vm)
    assert abs(alpha - a_ref) < 1e-13 * max(1.0, abs(a_ref))
    assert abs(beta - np.linalg.norm(res)) < 1e-13 * np.linalg.norm(res)
    assert rel(wd.cpu().numpy(), res / np.linalg.norm(res)) < 1e-13
    # start kernel: v0 = x / |x|
    x = cu(w); v0 = torch.empty_like(x); nrm = torch.zeros(1, dtype=torch.float64, device="cuda")
    assert lib.ptb_lanczos_start_z(n, x.data_ptr(), v0.data_ptr(), nrm.data_ptr(), scratch.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream) == 0
    assert abs(nrm.item() - np.linalg.norm(w)) < 1e-13 * np.linalg.norm(w)
    assert rel(v0.cpu().numpy(), w / np.linalg.norm(w)) < 1e-14


@pytest.mark.parametrize("cplx,dims", [(True, (16, 2, 28, 5, 5)), (False, (12, 2, 9, 4, 5)), (True, (130, 4, 121, 5, 5)),
                                       (True, (2, 2, 3, 2, 4))])
def test_fused_lanczos_run_equals_stepwise(cuda_lib, cplx, dims):
    """ptb_heff_lanczos / ptb_bond_lanczos (one call enqueues the whole run) give exactly the alphas, betas and
    Lanczos vectors of the step-by-step path (same kernels), and match the oracle's Lanczos on the same operator."""
    import pytenet_b200 as ptb
    from pytenet_b200 import krylov, _sweep
    Dl, d, Dr, cl, cr = dims
    rng = np.random.default_rng(Dl * 100 + Dr)

    def rnd(shape, c):
        x = rng.normal(size=shape)
        return x + 1j * rng.normal(size=shape) if c else x

    def herm_env(D, chi):
        # environments of a Hermitian operator: e[:, k, :] Hermitian for every k (then H_eff is Hermitian for
        # w[k, s', s, kappa] symmetric under s <-> s')
        e = rnd((D, chi, D), cplx)
        return np.ascontiguousarray(e + e.conj().transpose(2, 1, 0))

    l, r = herm_env(Dl, cl), herm_env(Dr, cr)
    w = rnd((cl, d, d, cr), False); w = w + w.transpose(0, 2, 1, 3); w[rng.random(w.shape) < 0.5] = 0
    w = np.ascontiguousarray(np.minimum(w, w.transpose(0, 2, 1, 3)))      # keep it symmetric after masking
    a = rnd((Dl, d, Dr), cplx)
    k = 6
    op = _sweep.HeffOperator(cu(w), cu(l), cu(r), (Dl, d, Dr))
    n1, al1, be1, V1 = krylov._lanczos_core(op, cu(a).reshape(-1), k)
    n2, al2, be2, V2 = krylov._lanczos_core(lambda x: op(x), cu(a).reshape(-1), k)
    if Dl * d * Dr > 20000:
        # same kernels on both paths: bit-identical
        assert n1 == n2 and np.array_equal(al1, al2) and np.array_equal(be1, be2) and torch.equal(V1, V2)
    else:
        # small bond dimensions: the fused run uses the one-kernel matvec (csrc/heff_small.cu), the step-by-step
        # path the three-kernel chain -- same numbers up to summation order
        assert n1 == n2 and np.allclose(al1, al2, rtol=1e-12, atol=1e-12 * np.abs(al2).max())
        assert np.allclose(be1, be2, rtol=1e-11) and rel(V1.cpu().numpy(), V2.cpu().numpy()) < 1e-10
    oal, obe, oV = oracle.lanczos_iteration(
        lambda x: oracle.apply_local_hamiltonian(x.reshape(Dl, d, Dr), w, l, r).reshape(-1), a.reshape(-1), k)
    assert np.allclose(al1, oal, rtol=1e-9, atol=1e-9 * np.abs(oal).max()) and np.allclose(be1, obe, rtol=1e-8)
    # zero-site (bond) operator: needs a common MPO bond
    r2 = herm_env(Dr, cl)
    c = rnd((Dl, Dr), cplx)
    bop = _sweep.BondOperator(cu(l), cu(r2), (Dl, Dr))
    n1, al1, be1, V1 = krylov._lanczos_core(bop, cu(c).reshape(-1), k)
    n2, al2, be2, V2 = krylov._lanczos_core(lambda x: bop(x), cu(c).reshape(-1), k)
    if cuda_lib.ptb_local_step_small_fits(Dl, 1, Dr, cl, cl, k):
        # small problems: the fused run is the one-kernel local step (csrc/lanczos_small.cu) -- same numbers up to
        # summation order
        assert n1 == n2 and np.allclose(al1, al2, rtol=1e-12, atol=1e-12 * np.abs(al2).max())
        assert np.allclose(be1, be2, rtol=1e-11) and rel(V1.cpu().numpy(), V2.cpu().numpy()) < 1e-10
    else:
        assert n1 == n2 and np.array_equal(al1, al2) and np.array_equal(be1, be2) and torch.equal(V1, V2)   # same entry point
    oal, obe, oV = oracle.lanczos_iteration(
        lambda x: oracle.apply_local_bond_contraction(x.reshape(Dl, Dr), l, r2).reshape(-1), c.reshape(-1), k)
    assert np.allclose(al1, oal, rtol=1e-9, atol=1e-9 * np.abs(oal).max()) and np.allclose(be1, obe, rtol=1e-8)


@pytest.mark.parametrize("k", [1, 2, 5, 8, 16, 17, 25, 64])
@pytest.mark.parametrize("vc,dt", [(True, 0.3 - 0.7j), (False, -0.45), (False, 0.2j), (True, 40.0j), (False, -3.0), (True, 200.0j)])
def test_expm_tridiagonal_on_device(cuda_lib, k, vc, dt):
    """ptb_krylov_expm_apply (k x k problem on the device + combination) against the reference formula
    v @ (U (|vec| exp(dt w) U[0])) evaluated with numpy.linalg.eigh (krylov.py:122-136, :142-150): Krylov spaces
    up to 16 through the shifted / scaled Taylor series with squaring (tridiag.cuh: tridiag_expm_taylor), larger
    ones -- and time steps whose norm would need more than ten squarings -- through the implicit QL iteration."""
    from pytenet_b200 import _lib, _device as dev
    lib = cuda_lib
    rng = np.random.default_rng(1000 * k + int(vc))
    n = 777
    alpha = rng.normal(size=k) * 3
    beta = np.abs(rng.normal(size=max(k - 1, 0))) + 0.1
    nrm = 1.7
    V = rng.normal(size=(k, n)) + (1j * rng.normal(size=(k, n)) if vc else 0)
    scal = np.concatenate([[nrm], alpha, beta, np.zeros(2 * k - 1 - len(alpha) - len(beta))])
    w_h, u_h = oracle.eigh_tridiag(alpha, beta)
    want = V.T @ (u_h @ (nrm * np.exp(dt * w_h) * u_h[0]))
    out_cplx = vc or isinstance(dt, complex)
    Vd, sd = cu(V), cu(scal)
    out = torch.empty(n, dtype=torch.complex128 if out_cplx else torch.float64, device="cuda")
    cws = torch.zeros(lib.ptb_krylov_expm_workspace_bytes() // 8, dtype=torch.float64, device="cuda")
    dtc = complex(dt)
    st = lib.ptb_krylov_expm_apply(1 if vc else 0, n, k, Vd.data_ptr(), n, sd.data_ptr(), dtc.real, dtc.imag,
                                   int(out_cplx), cws.data_ptr(), out.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    assert st == 0
    # the phases exp(dt w) carry the eigenvalue errors (eps |T|) times |dt|, whatever the algorithm
    assert rel(out.cpu().numpy(), want) < 1e-12 * max(1.0, 3 * abs(dt))
    assert int(cws[128:129].view(torch.int32)[0].item()) == k


def test_expm_on_device_truncates_at_breakdown(cuda_lib):
    """A beta below 100 n eps ends the Krylov space on the device exactly as krylov.py:44-50 does on the host
    (later rows of V may hold NaN from the division by ~0 and must not be touched)."""
    lib = cuda_lib
    n, k = 300, 6
    rng = np.random.default_rng(5)
    alpha = rng.normal(size=k); beta = np.array([0.8, 0.5, 1e-14, 0.3, 0.2])
    V = rng.normal(size=(k, n)) + 1j * rng.normal(size=(k, n)); V[3:] = np.nan
    nrm, dt = 0.9, -0.1 + 0.4j
    w_h, u_h = oracle.eigh_tridiag(alpha[:3], beta[:2])
    want = V[:3].T @ (u_h @ (nrm * np.exp(dt * w_h) * u_h[0]))
    scal = np.concatenate([[nrm], alpha, beta])
    Vd, sd = cu(V), cu(scal)
    out = torch.empty(n, dtype=torch.complex128, device="cuda")
    cws = torch.zeros(lib.ptb_krylov_expm_workspace_bytes() // 8, dtype=torch.float64, device="cuda")
    assert lib.ptb_krylov_expm_apply(1, n, k, Vd.data_ptr(), n, sd.data_ptr(), dt.real, dt.imag, 1, cws.data_ptr(),
                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    assert rel(out.cpu().numpy(), want) < 1e-12


def test_deferred_breakdown_warning(cuda_lib):
    """Inside deferred_checks() the RuntimeWarning of a Lanczos breakdown is issued when the block exits."""
    import pytenet_b200 as ptb
    from pytenet_b200 import _device as dev
    m = np.diag([1.0, 2.0, 3.0]); v = np.array([1.0, 1.0, 1.0])
    md = cu(m)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        with ptb.deferred_checks():
            got = ptb.expm_krylov(lambda x: dev.gemm(md, x.reshape(-1, 1)).reshape(-1), cu(v), -0.5, 6, hermitian=True)
            assert not any(issubclass(r.category, RuntimeWarning) for r in rec)
        assert any(issubclass(r.category, RuntimeWarning) and "Lanczos" in str(r.message) for r in rec)
    assert got.dtype == torch.float64 and rel(got.cpu().numpy(), np.exp(-0.5 * np.array([1.0, 2.0, 3.0]))) < 1e-12


@pytest.mark.parametrize("cplx", [True, False])
def test_arnoldi_and_general_expm(cuda_lib, cplx):
    """Arnoldi branch (krylov.py:60-107, :137-139) for a non-Hermitian operator: Hessenberg relation
    V^H A V = H, orthonormal V, and expm_krylov(hermitian=False) against scipy's dense expm."""
    import pytenet_b200 as ptb
    from scipy.linalg import expm
    rng = np.random.default_rng(9)
    n, k = 60, 60
    m = rng.normal(size=(n, n)) / np.sqrt(n)
    v = rng.normal(size=n)
    if cplx:
        m = m + 1j * rng.normal(size=(n, n)) / np.sqrt(n); v = v + 1j * rng.normal(size=n)
    hess, V = ptb.arnoldi_iteration(dev_matvec(m), cu(v), 12)
    Vh = V.cpu().numpy()
    assert rel(Vh.conj().T @ Vh, np.eye(12)) < 1e-13
    assert rel(Vh.conj().T @ m @ Vh, hess) < 1e-12
    assert np.allclose(np.tril(hess, -2), 0)
    dt = 0.4 - 0.3j if cplx else 0.4
    got = ptb.expm_krylov(dev_matvec(m), cu(v), dt, k)            # default: hermitian=False, as the reference
    assert rel(got.cpu().numpy(), expm(dt * m) @ v) < 1e-10
    # host-buffer entry
    got_h = ptb.expm_krylov(lambda x: m @ x, v, dt, 25)
    assert isinstance(got_h, np.ndarray) and rel(got_h, expm(dt * m) @ v) < 1e-10


@pytest.mark.parametrize("cplx,wc,dims", [(True, False, (16, 2, 20, 5, 5)), (False, False, (12, 2, 9, 4, 5)),
                                          (True, True, (7, 3, 5, 3, 2)), (True, False, (4, 4, 4, 3, 3)),
                                          (True, False, (2, 2, 3, 1, 4)), (False, False, (16, 2, 16, 5, 5)),
                                          # README-config bulk sites (clusters of 8 CTAs, unequal slices of the right bond)
                                          (True, False, (16, 2, 28, 5, 5)), (True, False, (28, 2, 16, 5, 5)),
                                          # l too large for shared memory (read through L1/L2), one column per CTA
                                          (True, False, (48, 2, 4, 5, 5)), (False, False, (3, 2, 2, 1, 4)),
                                          (True, True, (9, 2, 13, 4, 4))])
@pytest.mark.parametrize("k", [1, 5, 8])
def test_one_kernel_local_step_matches_oracle(cuda_lib, cplx, wc, dims, k):
    """ptb_local_step_small (csrc/lanczos_small.cu: start, all Lanczos iterations, k x k problem and combination in ONE
    kernel) against the oracle: alphas / betas / Lanczos vectors of krylov.lanczos_iteration and the result of
    expm_krylov, for the site problem (real and complex MPO tensor, real and complex state, real-time and
    imaginary-time steps) and the zero-site problem."""
    import oracle.lanczos as ol
    from pytenet_b200 import krylov, _sweep
    Dl, d, Dr, cl, cr = dims
    assert cuda_lib.ptb_local_step_small_fits(Dl, d, Dr, cl, cr, k)
    rng = np.random.default_rng(Dl * 1000 + Dr * 10 + k)

    def rnd(shape, c):
        x = rng.normal(size=shape)
        return x + 1j * rng.normal(size=shape) if c else x

    def herm_env(D, chi):
        e = rnd((D, chi, D), cplx)
        return np.ascontiguousarray(e + e.conj().transpose(2, 1, 0))

    l, r = herm_env(Dl, cl), herm_env(Dr, cr)
    w = rnd((cl, d, d, cr), wc)
    w = np.ascontiguousarray(w + w.conj().transpose(0, 2, 1, 3))     # Hermitian in (s', s) for every (k, K)
    w[np.abs(w) < 0.6] = 0
    a = rnd((Dl, d, Dr), cplx)
    hfun = lambda x: oracle.apply_local_hamiltonian(x.reshape(Dl, d, Dr), w, l, r).reshape(-1)      # noqa: E731
    op = _sweep.HeffOperator(cu(w), cu(l), cu(r), (Dl, d, Dr))
    for dt in ((0.05j, -0.1) if not wc else (0.05j,)):
        oal, obe, oV = ol.lanczos_iteration(hfun, a.reshape(-1), k)
        res = op.ptb_expm_run(cu(a).reshape(-1), dt, k)
        assert res is not None
        out, scal = res
        sc = scal.cpu().numpy()
        assert abs(sc[0] - np.linalg.norm(a)) < 1e-13 * np.linalg.norm(a)
        assert np.allclose(sc[1:1 + len(oal)], oal, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(oal).max()))
        assert np.allclose(sc[1 + k:1 + k + len(obe)], obe, rtol=1e-9)
        want = ol.expm_krylov(hfun, a.reshape(-1), dt, k, hermitian=True)
        assert rel(out.cpu().numpy(), want) < 1e-10
        # the public driver takes the same path
        got = krylov.expm_krylov(op, cu(a).reshape(-1), dt, k, hermitian=True)
        assert torch.equal(got, out)
    # Lanczos run only (eigh_krylov's use)
    n1, al1, be1, V1 = krylov._lanczos_core(op, cu(a).reshape(-1), k)
    assert np.allclose(al1, oal, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(oal).max()))
    assert rel(V1.cpu().numpy()[:oV.shape[1]], oV.T) < 1e-8
    # zero-site problem
    r2 = herm_env(Dr, cl)
    c = rnd((Dl, Dr), cplx)
    kfun = lambda x: oracle.apply_local_bond_contraction(x.reshape(Dl, Dr), l, r2).reshape(-1)      # noqa: E731
    bop = _sweep.BondOperator(cu(l), cu(r2), (Dl, Dr))
    res = bop.ptb_expm_run(cu(c).reshape(-1), -0.03j, k)
    assert res is not None
    want = ol.expm_krylov(kfun, c.reshape(-1), -0.03j, k, hermitian=True)
    assert rel(res[0].cpu().numpy(), want) < 1e-10
