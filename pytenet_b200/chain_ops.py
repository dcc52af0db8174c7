"""
Chain contractions of the effective-Hamiltonian path -- same names, argument
order and error behaviour as pytenet/chain_ops.py, executed by hand-written
sm_100a kernels through the C ABI (include/pytenet_b200.h).

Inputs may be CUDA torch tensors (device-resident path used by the sweeps; a
CUDA tensor is returned) or NumPy arrays (host-buffer entry: inputs are copied
to the device, the result is copied back and returned as ndarray -- this is the
end-to-end call measured as `e2e` by bench.py).
"""
import numpy as np
import torch

from . import _lib
from . import _device as dev

__all__ = ["contraction_operator_step_right", "contraction_operator_step_left",
           "compute_right_operator_blocks", "mpo_average", "mpo_inner_product",
           "apply_local_hamiltonian", "apply_local_bond_contraction", "apply_mpo"]


def _prep(tensors, ranks, names):
    """Common entry: rank asserts as in the reference, host->device, dtype promotion.

    Returns (host_mode, device, cplx, tensors-on-device)."""
    host_mode = dev.is_host(tensors[0])
    device = None
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            device = t.device
            break
    if device is None:
        device = dev.default_device()
    out = []
    for t, rk, nm in zip(tensors, ranks, names):
        td = dev.to_device(t, device)
        assert td.ndim == rk, f"`{nm}` must be a rank-{rk} tensor"      # chain_ops.py:45-48 etc.
        if td.dtype not in (dev.F64, dev.C128):
            td = td.to(dev.C128 if td.dtype.is_complex else dev.F64)
        out.append(td)
    return host_mode, device, dev.any_complex(*out), out


def _finish(out, host_mode):
    return dev.to_host(out) if host_mode else out


def _apply_local_hamiltonian_host(a, w, l, r):
    """Host-buffer entry for large tensors: one call of the C ABI's host form (csrc/host_entry.cu), which moves
    the operands in slices so that the PCIe copies overlap the three contraction steps.  The result lands in a
    page-locked buffer from torch's caching host allocator and is returned as a fresh ndarray."""
    lib = _lib.load()
    device = dev.default_device()
    for t, rk, nm in zip((a, w, l, r), (3, 4, 3, 3), "awlr"):
        assert np.ndim(t) == rk, f"`{nm}` must be a rank-{rk} tensor"
    cplx = any(np.iscomplexobj(t) for t in (a, l, r, w))
    w_cplx = bool(np.iscomplexobj(w))
    sdt = np.complex128 if cplx else np.float64
    # no-ops for C-ordered arrays of the common dtype (the int64 dummy blocks of chain_ops.py:110 are promoted)
    a, l, r = (np.ascontiguousarray(t, dtype=sdt) for t in (a, l, r))
    w = np.ascontiguousarray(w, dtype=np.complex128 if w_cplx else np.float64)
    Dl, d, Dr = a.shape
    cl, dout, din, cr = w.shape
    assert din == d and l.shape[0] == Dl and l.shape[1] == cl, "shape mismatch between a, w and l"
    assert r.shape[0] == Dr and r.shape[1] == cr, "shape mismatch between a, w and r"
    Dlp, Drp = l.shape[2], r.shape[2]
    dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    dims = (Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    nbytes = lib.ptb_apply_local_hamiltonian_host_workspace_bytes(dt, int(w_cplx), *dims)
    ws = dev.workspace(nbytes, device, tag="host")
    host = torch.empty((Dlp, dout, Drp), dtype=dev.C128 if cplx else dev.F64, pin_memory=True)
    st = lib.ptb_apply_local_hamiltonian_host(dt, int(w_cplx), a.ctypes.data, w.ctypes.data, l.ctypes.data,
                                              r.ctypes.data, host.data_ptr(), *dims, ws.data_ptr(), nbytes,
                                              dev.stream_ptr(device))
    _lib.check(st, "apply_local_hamiltonian (host buffers)")
    return host.numpy()


_HOST_OVERLAP_MIN_BYTES = 32 << 20


def apply_local_hamiltonian(a, w, l, r, out=None):
    r"""
    Apply a local Hamiltonian operator (pytenet/chain_ops.py:237-279)::

        out[i',s',j'] = sum l[i,k,i'] w[k,s',s,kappa] a[i,s,j] r[j,kappa,j']

    `a` (Dl,d,Dr), `w` (chi_l,d_out,d_in,chi_r), `l` (Dl,chi_l,Dl'), `r` (Dr,chi_r,Dr').
    """
    lib = _lib.load()
    if (out is None and all(isinstance(t, np.ndarray) for t in (a, w, l, r)) and np.ndim(l) == 3
            and l.nbytes >= _HOST_OVERLAP_MIN_BYTES):
        return _apply_local_hamiltonian_host(a, w, l, r)
    host_mode, device, cplx, (a, w, l, r) = _prep((a, w, l, r), (3, 4, 3, 3), "awlr")
    Dl, d, Dr = a.shape
    cl, dout, din, cr = w.shape
    assert din == d and l.shape[0] == Dl and l.shape[1] == cl, "shape mismatch between a, w and l"
    assert r.shape[0] == Dr and r.shape[1] == cr, "shape mismatch between a, w and r"
    Dlp, Drp = l.shape[2], r.shape[2]
    a = dev.as_dtype(a, cplx); l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
    w_cplx = w.dtype.is_complex
    w = dev.dense(w)
    dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    if out is None:
        out = torch.empty((Dlp, dout, Drp), dtype=a.dtype, device=device)
    else:
        assert out.shape == (Dlp, dout, Drp) and out.dtype == a.dtype and out.is_contiguous()
    nbytes = lib.ptb_apply_local_hamiltonian_workspace_bytes(dt, Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    ws = dev.workspace(nbytes, device)
    dims = (Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    tail = (ws.data_ptr(), nbytes, dev.stream_ptr(device))
    csr = dev.w_csr(w) if (cplx or not w_cplx) else None     # sparse W kernel for 5-17 % dense MPO tensors
    if csr is not None:
        rowptr, col, val, _ = csr
        if cplx:
            st = lib.ptb_apply_local_hamiltonian_csr_z(a.data_ptr(), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                                       int(w_cplx), l.data_ptr(), r.data_ptr(), out.data_ptr(),
                                                       *dims, *tail)
        else:
            st = lib.ptb_apply_local_hamiltonian_csr_d(a.data_ptr(), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                                       l.data_ptr(), r.data_ptr(), out.data_ptr(), *dims, *tail)
    elif cplx:
        st = lib.ptb_apply_local_hamiltonian_z(a.data_ptr(), w.data_ptr(), int(w_cplx), l.data_ptr(), r.data_ptr(),
                                               out.data_ptr(), *dims, *tail)
    else:
        st = lib.ptb_apply_local_hamiltonian_d(a.data_ptr(), w.data_ptr(), l.data_ptr(), r.data_ptr(),
                                               out.data_ptr(), *dims, *tail)
    _lib.check(st, "apply_local_hamiltonian")
    return _finish(out, host_mode)


def apply_local_bond_contraction(c, l, r, out=None):
    r"""
    Apply a "zero-site" bond contraction (pytenet/chain_ops.py:282-317)::

        out[i',j'] = sum l[i,k,i'] c[i,j] r[j,k,j']
    """
    lib = _lib.load()
    host_mode, device, cplx, (c, l, r) = _prep((c, l, r), (2, 3, 3), "clr")
    Dl, Dr = c.shape
    chi = l.shape[1]
    assert l.shape[0] == Dl and r.shape[0] == Dr and r.shape[1] == chi, "shape mismatch between c, l and r"
    Dlp, Drp = l.shape[2], r.shape[2]
    c = dev.as_dtype(c, cplx); l = dev.as_dtype(l, cplx); r = dev.as_dtype(r, cplx)
    dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    if out is None:
        out = torch.empty((Dlp, Drp), dtype=c.dtype, device=device)
    else:
        assert out.shape == (Dlp, Drp) and out.dtype == c.dtype and out.is_contiguous()
    nbytes = lib.ptb_apply_local_bond_contraction_workspace_bytes(dt, Dl, Dr, chi, Dlp, Drp)
    ws = dev.workspace(nbytes, device)
    fn = lib.ptb_apply_local_bond_contraction_z if cplx else lib.ptb_apply_local_bond_contraction_d
    st = fn(c.data_ptr(), l.data_ptr(), r.data_ptr(), out.data_ptr(), Dl, Dr, chi, Dlp, Drp,
            ws.data_ptr(), nbytes, dev.stream_ptr(device))
    _lib.check(st, "apply_local_bond_contraction")
    return _finish(out, host_mode)


def _env_step(which, a, b, w, env):
    lib = _lib.load()
    host_mode, device, cplx, (a, b, w, env) = _prep((a, b, w, env), (3, 3, 4, 3), ("a", "b", "w", "lr"[which]))
    Dl, d, Dr = a.shape
    Dlp, dout, Drp = b.shape
    cl, dw_out, dw_in, cr = w.shape
    assert dw_in == d and dw_out == dout, "physical dimensions of a, b and w must agree"
    if which == 0:
        assert tuple(env.shape) == (Dl, cl, Dlp), "shape mismatch between l and a, w, b"
        oshape = (Dr, cr, Drp)
    else:
        assert tuple(env.shape) == (Dr, cr, Drp), "shape mismatch between r and a, w, b"
        oshape = (Dl, cl, Dlp)
    a = dev.as_dtype(a, cplx); b = dev.as_dtype(b, cplx); env = dev.as_dtype(env, cplx)
    w_cplx = w.dtype.is_complex
    w = dev.dense(w)
    dt = _lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64
    out = torch.empty(oshape, dtype=a.dtype, device=device)
    nbytes = lib.ptb_env_step_workspace_bytes(dt, Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    ws = dev.workspace(nbytes, device)
    dims = (Dl, d, Dr, cl, cr, dout, Dlp, Drp)
    tail = (ws.data_ptr(), nbytes, dev.stream_ptr(device))
    if cplx:
        fn = lib.ptb_env_step_left_z if which == 0 else lib.ptb_env_step_right_z
        st = fn(a.data_ptr(), b.data_ptr(), w.data_ptr(), int(w_cplx), env.data_ptr(), out.data_ptr(), *dims, *tail)
    else:
        fn = lib.ptb_env_step_left_d if which == 0 else lib.ptb_env_step_right_d
        st = fn(a.data_ptr(), b.data_ptr(), w.data_ptr(), env.data_ptr(), out.data_ptr(), *dims, *tail)
    _lib.check(st, "contraction_operator_step_" + ("left" if which == 0 else "right"))
    return _finish(out, host_mode)


def contraction_operator_step_left(a, b, w, l):
    r"""
    Contraction step from left to right with an MPO tensor sandwiched in between
    (pytenet/chain_ops.py:60-99)::

        l_next[j,kappa,j'] = sum l[i,k,i'] conj(b[i',s',j']) w[k,s',s,kappa] a[i,s,j]
    """
    return _env_step(0, a, b, w, l)


def contraction_operator_step_right(a, b, w, r):
    r"""
    Contraction step from right to left with an MPO tensor sandwiched in between
    (pytenet/chain_ops.py:16-57)::

        r_next[i,k,i'] = sum a[i,s,j] r[j,kappa,j'] w[k,s',s,kappa] conj(b[i',s',j'])
    """
    return _env_step(1, a, b, w, r)


def compute_right_operator_blocks(psi, op):
    """
    Compute all partial contractions from the right (pytenet/chain_ops.py:102-113).
    `blocks[nsites-1]` is the 1 x 1 x 1 dummy block.
    """
    nsites = psi.nsites
    assert nsites == op.nsites
    blocks = [None for _ in range(nsites)]
    device = psi.a[-1].device if isinstance(psi.a[-1], torch.Tensor) else dev.default_device()
    blocks[nsites - 1] = torch.ones((1, 1, 1), dtype=dev.F64, device=device)
    for i in reversed(range(nsites - 1)):
        blocks[i] = contraction_operator_step_right(psi.a[i + 1], psi.a[i + 1], op.a[i + 1], blocks[i + 1])
    return blocks


def mpo_inner_product(chi, op, psi):
    """
    Compute `<chi | op | psi>` (pytenet/chain_ops.py:166-189) by repeated
    `contraction_operator_step_right`, on the device; returns a Python scalar.
    """
    assert chi.nsites == op.nsites
    assert psi.nsites == op.nsites
    if psi.nsites == 0:
        return 0
    last = dev.to_device(psi.a[-1])
    assert chi.a[-1].shape[2] == psi.a[-1].shape[2]
    n = last.shape[2]
    t = torch.eye(n, dtype=last.dtype, device=last.device).reshape(n, 1, n)
    for i in reversed(range(psi.nsites)):
        t = contraction_operator_step_right(dev.to_device(psi.a[i]), dev.to_device(chi.a[i]),
                                            dev.to_device(op.a[i]), t)
    assert tuple(t.shape) == (1, 1, 1)
    return t.reshape(-1)[0].item()


def mpo_average(psi, op):
    """Expectation value `<psi | op | psi>` (pytenet/chain_ops.py:151-163)."""
    return mpo_inner_product(psi, op, psi)


def apply_mpo(op, psi):
    """
    Apply an operator in MPO form to a state in MPS form (pytenet/chain_ops.py:215-234): per site
    `t[(k,i), s', (kappa,j)] = sum_s w[k,s',s,kappa] a[i,s,j]`, virtual bonds grouped MPO-major.

    The contraction index has length d, so the step is pure data movement: it runs on the HBM-bound sparse
    W kernel (`ptb_wapply_csr`, batched over the left bond index i, one pass over the output) followed by
    the regrouping of the bond indices.
    """
    from .mps import MPS
    from .block_sparse_util import qnumber_flatten, is_qsparse
    lib = _lib.load()
    assert np.array_equal(psi.qsite, op.qsite)
    assert psi.nsites == op.nsites
    qbonds = [qnumber_flatten((op.qbonds[i], psi.qbonds[i])) for i in range(psi.nsites + 1)]
    out = MPS(psi.qsite, qbonds, fill="postpone", device=psi.device)
    for i in range(psi.nsites):
        a = dev.to_device(psi.a[i], psi.device)
        w = dev.dense(dev.to_device(op.a[i], psi.device))
        cplx = dev.any_complex(a, w)
        a = dev.as_dtype(a, cplx)
        Dl, d, Dr = a.shape
        cl, dout, din, cr = w.shape
        assert din == d
        # W' = w viewed as ((k, s', kappa), s); t[i, (k, s', kappa), j] = sum_s W'[(k,s',kappa), s] a[i, s, j]
        wp = dev.dense(w.permute(0, 1, 3, 2)).reshape(cl * dout * cr, d)
        rowptr, col, val, _ = dev.csr_arrays_any(wp)
        t = torch.empty((Dl, cl, dout, cr, Dr), dtype=a.dtype, device=a.device)
        st = lib.ptb_wapply_csr(_lib.PTB_COMPLEX128 if cplx else _lib.PTB_REAL64, int(w.dtype.is_complex),
                                cl * dout * cr, d, Dr, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                a.data_ptr(), t.data_ptr(), Dl, dev.stream_ptr(a.device))
        _lib.check(st, "ptb_wapply_csr(apply_mpo)")
        out.a[i] = dev.dense(t.permute(1, 0, 2, 3, 4)).reshape(cl * Dl, dout, cr * Dr)
        assert is_qsparse(out.a[i], (out.qbonds[i], out.qsite, -out.qbonds[i + 1])), \
            "sparsity pattern of MPS tensor does not match quantum numbers"
    return out
