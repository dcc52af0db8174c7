"""
Device-buffer plumbing between the Python API and the C ABI.

torch tensors are used only as device memory + stream handles; every
arithmetic kernel on the path is in libpytenet_b200.so.
"""
import numpy as np
import torch

from . import _lib

C128 = torch.complex128
F64 = torch.float64

_workspaces = {}     # (device index, tag) -> torch.uint8 buffer
_scratch = {}        # device index -> zero-initialised Lanczos reduction scratch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pytenet_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def default_device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def is_host(x):
    """True for NumPy arrays / Python scalars / lists (the host-buffer, end-to-end entry)."""
    return not isinstance(x, torch.Tensor)


def to_device(x, device=None):
    """NumPy array or torch tensor -> torch tensor on the CUDA device (no dtype change)."""
    if isinstance(x, torch.Tensor):
        if x.is_cuda:
            return x
        return x.to(device or default_device())
    arr = np.asarray(x)
    if arr.dtype not in (np.float64, np.complex128):
        # the reference's dummy edge blocks are int64 [[[1]]] (chain_ops.py:110)
        arr = arr.astype(np.complex128 if np.iscomplexobj(arr) else np.float64)
    t = torch.from_numpy(np.ascontiguousarray(arr))
    # page-locked source (e.g. a NumPy view of a pinned buffer): stream-ordered DMA that overlaps with
    # kernels of other streams; pageable memory is staged synchronously by the driver either way
    return t.to(device or default_device(), non_blocking=t.is_pinned())


def to_host(x):
    """Device tensor -> fresh NumPy array.  The copy lands in pinned memory from torch's
    caching host allocator (full-speed DMA); every call returns its own buffer, so results
    never alias each other."""
    x = x.detach()
    if not x.is_cuda:
        return x.numpy()
    if x.numel() * x.element_size() < (1 << 20):
        return x.cpu().numpy()
    buf = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    buf.copy_(x)
    return buf.numpy()


def dense(x):
    """Materialised, C-contiguous tensor whose memory is exactly its values: torch marks
    conjugation / negation lazily (bit flags on a view, e.g. on the factors returned by
    torch.linalg.svd), which a raw data_ptr() consumer would silently ignore."""
    return x.resolve_conj().resolve_neg().contiguous()


def as_dtype(x, cplx):
    """Contiguous float64 / complex128 view-or-copy of a device tensor (safe to pass by pointer)."""
    want = C128 if cplx else F64
    if x.dtype != want:
        if x.dtype.is_complex and not cplx:
            raise TypeError("cannot demote a complex tensor to float64")
        x = x.to(want)
    return dense(x)


def any_complex(*tensors):
    return any(t.dtype.is_complex for t in tensors)


def stream_ptr(device):
    """Raw cudaStream_t of torch's current stream on `device` (the fast path of torch.cuda.current_stream)."""
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device() if idx is None else idx)


def workspace(nbytes, device, tag="main"):
    """Grow-only per-device workspace; safe because every user is stream-ordered."""
    key = (device.index, tag, stream_ptr(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            del _workspaces[key]
            buf = None
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def lanczos_scratch(device):
    key = (device.index, stream_ptr(device))
    buf = _scratch.get(key)
    if buf is None:
        nbytes = _lib.load().ptb_lanczos_scratch_bytes()
        buf = torch.zeros(nbytes // 8, dtype=F64, device=device)   # ticket must start at zero
        _scratch[key] = buf
    return buf


_side_streams = {}


def side_stream(device):
    """A second stream per device for copies that overlap compute (host-buffer entry points)."""
    st = _side_streams.get(device.index)
    if st is None:
        st = torch.cuda.Stream(device)
        _side_streams[device.index] = st
    return st


ws_owner = {}       # (device index, stream) -> token of the sector plan whose structural zeros are in the workspace


def release_workspaces():
    _workspaces.clear()
    _scratch.clear()
    ws_owner.clear()


_GEMM_SCRATCH_BYTES = 32 << 20


def gemm(a, b, trans_a=False, trans_b=False, conj_b=False, out=None, engine=0):
    """2-D GEMM through the DMMA engine on contiguous device matrices.

    op(a) (M x K) @ op(b) (K x N) -> (M x N).  Used for the small GEMM-shaped steps
    that sit between the hot contractions (merge, gauge absorption).  `engine` != 0 pins the
    kernel generation (ptb_gemm_engine; measurements only).
    """
    lib = _lib.load()
    cplx = any_complex(a, b)
    a = as_dtype(a, cplx)
    b = as_dtype(b, cplx)
    assert a.ndim == 2 and b.ndim == 2
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    k2, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    assert k == k2, "inner dimensions must agree"
    if out is None:
        out = torch.empty((m, n), dtype=a.dtype, device=a.device)
    if m == 0 or n == 0:
        return out
    if k == 0:
        return out.zero_()
    if engine:
        st = lib.ptb_gemm_engine(int(engine), _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, int(trans_a),
                                 int(trans_b), int(conj_b), m, n, k, a.data_ptr(), a.shape[1], b.data_ptr(),
                                 b.shape[1], out.data_ptr(), n, 1, 0, 0, 0, 0, stream_ptr(a.device))
        _lib.check(st, "ptb_gemm_engine")
        return out
    # a fixed 32 MB scratch lets the engine split K when the output has few tiles, and split the tiles of
    # the last partial wave of large outputs (both deterministic; see csrc/gemm_ws.cuh)
    ws = workspace(_GEMM_SCRATCH_BYTES, a.device, tag="gemm")
    st = lib.ptb_gemm_splitk(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, int(trans_a), int(trans_b),
                             int(conj_b), m, n, k, a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1],
                             out.data_ptr(), n, 1, 0, 0, 0, 0, 0, ws.data_ptr(), ws.numel(), stream_ptr(a.device))
    _lib.check(st, "ptb_gemm_splitk")
    return out


def gemm_strided(cplx, trans_a, trans_b, conj_b, m, n, k, a, lda, b, ldb, c, ldc,
                 batch=1, stride_a=0, stride_b=0, stride_c=0, accumulate=False):
    """Thin call-through to ptb_gemm on tensors that are already dense device buffers of the
    right dtype (float64 when `cplx` is False -- a complex tensor may be passed as its float64
    view for the real-W trick -- or complex128).  Leading dimensions / strides in elements."""
    lib = _lib.load()
    st = lib.ptb_gemm(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, int(trans_a), int(trans_b), int(conj_b),
                      int(m), int(n), int(k), a.data_ptr(), int(lda), b.data_ptr(), int(ldb), c.data_ptr(), int(ldc),
                      int(batch), int(stride_a), int(stride_b), int(stride_c), int(bool(accumulate)),
                      stream_ptr(c.device))
    _lib.check(st, "ptb_gemm")
    return c


# ---- CSR form of small sparse MPO tensors (cached per tensor version) ---------------------------
_csr_cache = {}
_CSR_MAX_NNZ = 4096          # beyond this the dense GEMM W step is the better kernel


def w_csr_from_host(w_host, device):
    """CSR arrays on `device` built from a host (NumPy) MPO tensor; None when too dense for the sparse kernel."""
    cl, dout, din, cr = w_host.shape
    mat = np.ascontiguousarray(w_host).reshape(cl * dout, din * cr)
    if mat.dtype not in (np.float64, np.complex128):
        mat = mat.astype(np.complex128 if np.iscomplexobj(mat) else np.float64)
    return _csr_arrays(mat, device)


class _Nnz(int):
    """Number of non-zeros that also carries the sparsity pattern (row pointers + column indices as bytes): a
    content-based key for everything derived from W's pattern (pointer / version keys can be recycled)."""
    pattern = b""


def _csr_arrays(mat, device):
    rows, cols = np.nonzero(mat)
    if len(rows) > _CSR_MAX_NNZ:
        return None
    rowptr = np.zeros(mat.shape[0] + 1, dtype=np.int32)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr, dtype=np.int64).astype(np.int32)
    vals = np.ascontiguousarray(mat[rows, cols])
    if len(rows) == 0:                           # all-zero MPO tensor: keep the arrays non-empty (non-null pointers)
        cols = np.zeros(1, dtype=np.int64)
        vals = np.zeros(1, dtype=mat.dtype)
    nnz = _Nnz(len(rows))
    nnz.pattern = rowptr.tobytes() + cols.astype(np.int32).tobytes()
    return (torch.from_numpy(rowptr).to(device), torch.from_numpy(cols.astype(np.int32)).to(device),
            torch.from_numpy(vals).to(device), nnz)


def csr_arrays_any(mat):
    """CSR device arrays of a 2-D device matrix of any density (no non-zero limit; one small device->host copy)."""
    host = mat.detach().cpu().numpy()
    rows, cols = np.nonzero(host)
    rowptr = np.zeros(host.shape[0] + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    vals = np.ascontiguousarray(host[rows, cols])
    if len(rows) == 0:
        cols = np.zeros(1, dtype=np.int64)
        vals = np.zeros(1, dtype=host.dtype)
    return (torch.from_numpy(rowptr).to(mat.device), torch.from_numpy(cols.astype(np.int32)).to(mat.device),
            torch.from_numpy(vals).to(mat.device), len(rows))


def w_csr(w):
    """(rowptr, col, val, nnz) device arrays of w reshaped to (chi_l*d_out) x (d_in*chi_r), or None when w
    is too dense / large for the sparse W kernel.  Built once per tensor (one small device->host copy of w)."""
    key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype)
    hit = _csr_cache.get(key)
    if hit is not None:
        return hit[0]
    cl, dout, din, cr = w.shape
    mat = w.detach().reshape(cl * dout, din * cr).cpu().numpy()
    result = _csr_arrays(mat, w.device)
    if len(_csr_cache) > 256:
        _csr_cache.clear()
    _csr_cache[key] = (result, w)           # keep `w` alive so the data_ptr key cannot be recycled
    return result


def bind_host_to_gpu_numa_node(device_index=None):
    """Restrict this process to the CPUs of the NUMA node the GPU hangs off (sysfs: the PCI device's `numa_node`
    and the node's `cpulist`), so that page-locked staging buffers allocated afterwards are first-touched on the
    GPU-local node and host<->device copies do not cross the socket interconnect.  With one process per GPU on a
    two-socket box, eight ranks otherwise contend for one socket's memory controllers (measured: end-to-end time
    +15 % at N = 8).  Returns a short description, or None when the topology is not exposed (then nothing changes)."""
    import os
    try:
        idx = torch.cuda.current_device() if device_index is None else int(device_index)
        props = torch.cuda.get_device_properties(idx)
        # torch exposes the PCI address as three integers (domain, bus, device); sysfs wants dddd:bb:dd.f
        bus = "%04x:%02x:%02x.0" % (int(getattr(props, "pci_domain_id", 0)), int(props.pci_bus_id),
                                    int(props.pci_device_id))
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return f"gpu {idx} ({bus}) -> numa node {node}, {len(allowed)} cpus"
    except Exception:
        return None
