"""GPU: the DMMA GEMM engine (ptb_gemm) against NumPy on seeded inputs.
Tolerance: 1e-13 relative (Frobenius) -- north_star asks 1e-12 for the contractions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-13


def rel(x, y):
    n = np.linalg.norm(y)
    return np.linalg.norm(x - y) / (n if n > 0 else 1.0)


def rnd(rng, shape, cplx):
    x = rng.normal(size=shape)
    if cplx:
        x = x + 1j * rng.normal(size=shape)
    return x


def run_gemm(lib, cplx, ta, tb, cj, M, N, K, rng, batch=1, accumulate=False, pad=(0, 0, 0), engine=0):
    from pytenet_b200 import _lib
    lda = (M if ta else K) + pad[0]
    ldb = (K if tb else N) + pad[1]
    ldc = N + pad[2]
    A = rnd(rng, (batch, K if ta else M, lda), cplx)
    B = rnd(rng, (batch, N if tb else K, ldb), cplx)
    C0 = rnd(rng, (batch, M, ldc), cplx)
    dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, C0))
    st = lib.ptb_gemm_engine(engine, _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, ta, tb, cj, M, N, K,
                             dA.data_ptr(), lda, dB.data_ptr(), ldb, dC.data_ptr(), ldc, batch,
                             A.shape[1] * lda, B.shape[1] * ldb, M * ldc, int(accumulate),
                             torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.ptb_status_string(st)
    torch.cuda.synchronize()
    got = dC.cpu().numpy()
    opA = A[:, :, :M].transpose(0, 2, 1) if ta else A[:, :, :K]
    opB = B[:, :, :K].transpose(0, 2, 1) if tb else B[:, :, :N]
    if cj:
        opB = opB.conj()
    want = C0.copy()
    prod = opA @ opB
    want[:, :, :N] = prod + (C0[:, :, :N] if accumulate else 0)
    # padding columns of C must be untouched
    assert np.array_equal(got[:, :, N:], C0[:, :, N:])
    return rel(got[:, :, :N], want[:, :, :N])


@pytest.fixture(params=[0, 1, 2], ids=["engine-auto", "engine-cpasync", "engine-tma"])
def engine(request, cuda_lib):
    """Run every GEMM case on both kernel generations (2 = warp-specialised TMA kernel required); the engine is an
    argument of ptb_gemm_engine -- the library holds no process-wide mode."""
    return request.param


SHAPES = [(1, 1, 1), (3, 5, 7), (8, 8, 4), (17, 9, 33), (128, 64, 8), (129, 65, 9), (130, 200, 77),
          (64, 257, 19), (300, 70, 130), (1, 300, 5), (257, 1, 64), (256, 256, 256)]


@pytest.mark.parametrize("cplx", [True, False])
@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_gemm_layouts_and_ragged_shapes(cuda_lib, engine, cplx, ta, tb):
    if engine == 2 and not cplx:
        pytest.skip("odd float64 extents are served by the cp.async kernel")
    rng = np.random.default_rng(100 + 4 * cplx + 2 * ta + tb)
    for (M, N, K) in SHAPES:
        for cj in ([0, 1] if cplx else [0]):
            e = run_gemm(cuda_lib, cplx, ta, tb, cj, M, N, K, rng, engine=engine)
            assert e < TOL, (cplx, ta, tb, cj, M, N, K, e)


@pytest.mark.parametrize("cplx", [True, False])
def test_gemm_batched_accumulate_and_leading_dims(cuda_lib, engine, cplx):
    rng = np.random.default_rng(7 + cplx)
    # odd leading dimensions force the 8-byte copy path for float64
    pads = [(0, 0, 0), (1, 3, 5), (2, 2, 2)] if (cplx or engine != 2) else []
    for pad in pads:
        for (ta, tb) in [(0, 0), (1, 0), (0, 1), (1, 1)]:
            e = run_gemm(cuda_lib, cplx, ta, tb, 0, 37, 45, 29, rng, batch=5, accumulate=True, pad=pad, engine=engine)
            assert e < TOL, (cplx, pad, ta, tb, e)
    e = run_gemm(cuda_lib, cplx, 0, 0, 0, 10, 130, 10, rng, batch=300, engine=engine)
    assert e < TOL
    # float64 with even extents is eligible for the TMA kernel in every layout
    for (ta, tb) in [(0, 0), (1, 0), (0, 1), (1, 1)]:
        e = run_gemm(cuda_lib, cplx, ta, tb, 0, 70, 36, 50, rng, batch=3, accumulate=True, pad=(2, 4, 6), engine=engine)
        assert e < TOL, (cplx, ta, tb, e)
        e = run_gemm(cuda_lib, cplx, ta, tb, 0, 200, 300, 40, rng, engine=engine)
        assert e < TOL, (cplx, ta, tb, e)


def test_gemm_exact_zeros_preserved(cuda_lib):
    """Structural zeros of block-sparse operands must give exactly 0.0 outputs (SURVEY.md headline 3)."""
    from pytenet_b200 import _device as dev
    rng = np.random.default_rng(5)
    a = rnd(rng, (40, 30), True); b = rnd(rng, (30, 50), True)
    a[:20, 15:] = 0; a[20:, :15] = 0
    b[:15, 25:] = 0; b[15:, :25] = 0
    c = dev.gemm(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy()
    assert np.all(c[:20, 25:] == 0) and np.all(c[20:, :25] == 0)
    assert rel(c, a @ b) < TOL


def test_gemm_large_against_torch(cuda_lib, engine):
    """Full-size sanity (size-independent check): 2048 x 1280 x 1024 complex, T/N layout, vs torch fp64."""
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(1024, 2048, dtype=torch.complex128, device="cuda", generator=g)
    b = torch.randn(1024, 1280, dtype=torch.complex128, device="cuda", generator=g)
    c = torch.empty(2048, 1280, dtype=torch.complex128, device="cuda")
    st = cuda_lib.ptb_gemm_engine(engine, 1, 1, 0, 0, 2048, 1280, 1024, a.data_ptr(), 2048, b.data_ptr(), 1280,
                                  c.data_ptr(), 1280, 1, 0, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    assert st == 0
    ref = a.T @ b
    assert (torch.linalg.norm(c - ref) / torch.linalg.norm(ref)).item() < TOL


@pytest.mark.parametrize("cplx", [True, False])
def test_gemm_split_k(cuda_lib, cplx):
    """Split-K (partial tiles + ordered reduction) against NumPy: forced factors, automatic choice,
    ragged K, batches and accumulate."""
    from pytenet_b200 import _lib
    rng = np.random.default_rng(21 + cplx)
    lib = cuda_lib
    for (M, N, K, batch, split, acc, ta) in [(130, 70, 1000, 1, 2, False, 0), (128, 64, 777, 1, 3, True, 1),
                                             (256, 128, 4096, 2, 4, False, 1), (100, 60, 2048, 1, 0, False, 1),
                                             (64, 64, 50, 1, 8, False, 0),
                                             # > 148 tiles with a small last wave: tail splitting (auto mode)
                                             (1664, 768, 512, 1, 0, False, 0), (1600, 700, 300, 1, 0, True, 1),
                                             (640, 520, 200, 3, 0, False, 1)]:
        A = rnd(rng, (batch, K, M) if ta else (batch, M, K), cplx)
        B = rnd(rng, (batch, K, N), cplx)
        C0 = rnd(rng, (batch, M, N), cplx)
        dA, dB, dC = (torch.from_numpy(x).cuda() for x in (A, B, C0))
        ws = torch.empty(batch * 8 * M * N * (2 if cplx else 1), dtype=torch.float64, device="cuda")
        st = lib.ptb_gemm_splitk(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, ta, 0, 0, M, N, K,
                                 dA.data_ptr(), M if ta else K, dB.data_ptr(), N, dC.data_ptr(), N, batch,
                                 M * K, K * N, M * N, int(acc), split, ws.data_ptr(), ws.numel() * 8,
                                 torch.cuda.current_stream().cuda_stream)
        assert st == 0, lib.ptb_status_string(st)
        want = (A.transpose(0, 2, 1) if ta else A) @ B + (C0 if acc else 0)
        assert rel(dC.cpu().numpy(), want) < TOL, (M, N, K, batch, split, acc, ta)


@pytest.mark.parametrize("cplx,conj", [(True, False), (True, True), (False, False)])
def test_gemm_segmented(cuda_lib, cplx, conj):
    """ptb_gemm_segmented: every output tile sums its own list of (k-tile range, selector) segments, the selector
    adding element offsets to the A / B base pointers; tiles without segments are zero.  Checked tile by tile
    against NumPy on ragged extents with two batches."""
    import ctypes
    lib = cuda_lib
    rng = np.random.default_rng(42 + int(cplx) + 2 * int(conj))
    bm, bn, bk = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.ptb_gemm_tile_shape(1 if cplx else 0, ctypes.byref(bm), ctypes.byref(bn), ctypes.byref(bk)) == 0
    bm, bn, bk = bm.value, bn.value, bk.value
    M, N, K, batch, nsel = 2 * bm + 38, 3 * bn + 10, 5 * bk + 6, 2, 3
    KT = -(-K // bk)
    tm, tn = -(-M // bm), -(-N // bn)

    def rnd(*shape):
        x = rng.normal(size=shape)
        return x + 1j * rng.normal(size=shape) if cplx else x

    A = rnd(batch, nsel, K, M)                 # selector s of batch b: element offset s*K*M, batch stride nsel*K*M
    B = rnd(batch, nsel, K, N)
    want = np.zeros((batch, M, N), dtype=A.dtype)
    seg_ptr, segs = [0], []
    for b in range(batch):
        for i in range(tm):
            for j in range(tn):
                nseg = int(rng.integers(0, 4))
                for _ in range(nseg):
                    lo = int(rng.integers(0, KT)); hi = int(rng.integers(lo, KT + 1)); sel = int(rng.integers(0, nsel))
                    segs.append([lo, hi, sel, 0])
                    rows = slice(i * bm, min((i + 1) * bm, M)); cols = slice(j * bn, min((j + 1) * bn, N))
                    ks = slice(lo * bk, min(hi * bk, K))
                    bb = B[b, sel, ks, cols]
                    want[b, rows, cols] += A[b, sel, ks, rows].T @ (bb.conj() if conj else bb)
                seg_ptr.append(len(segs))
    if not segs:
        segs.append([0, 0, 0, 0])
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.full((batch, M, N), 7.0, dtype=dA.dtype, device="cuda")
    dptr = torch.tensor(seg_ptr, dtype=torch.int32, device="cuda")
    dseg = torch.tensor(segs, dtype=torch.int32, device="cuda")
    doff = torch.tensor([[s * K * M, s * K * N] for s in range(nsel)], dtype=torch.int64, device="cuda")
    st = lib.ptb_gemm_segmented(1 if cplx else 0, int(conj), M, N, K, dA.data_ptr(), M, dB.data_ptr(), N, dC.data_ptr(),
                                N, batch, nsel * K * M, nsel * K * N, M * N, 0, dptr.data_ptr(), dseg.data_ptr(),
                                doff.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0
    got = dC.cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < TOL
    assert np.array_equal(got == 0, want == 0)              # tiles without segments are exactly zero
    # accumulate on top of an existing C
    st = lib.ptb_gemm_segmented(1 if cplx else 0, int(conj), M, N, K, dA.data_ptr(), M, dB.data_ptr(), N, dC.data_ptr(),
                                N, batch, nsel * K * M, nsel * K * N, M * N, 1, dptr.data_ptr(), dseg.data_ptr(),
                                doff.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0
    assert np.linalg.norm(dC.cpu().numpy() - 2 * want) / np.linalg.norm(want) < TOL
