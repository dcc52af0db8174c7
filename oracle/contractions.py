"""
NumPy restatement of the chain contractions (TEST INFRASTRUCTURE, see
oracle/__init__.py).

Every contraction is written as explicit reshape + matmul in the same
contraction *order* as the reference's ``np.tensordot`` calls, so the GEMM
shapes (and therefore the flop count F_alg used by bench.py) are the
reference's own.

Index names: i, j ket bonds; i', j' bra bonds; k, kappa MPO bonds; s in,
s' out physical index.
"""
import numpy as np


def _as3(t):
    t = np.asarray(t)
    if t.ndim != 3:
        raise AssertionError("expected a rank-3 tensor")
    return t


def apply_local_hamiltonian(a, w, l, r):
    """out[i',s',j'] = sum l[i,k,i'] w[k,s',s,kappa] a[i,s,j] r[j,kappa,j'].

    Restates pytenet/chain_ops.py:237-279 (three tensordots at :273, :276, :278).
    """
    a = _as3(a); l = _as3(l); r = _as3(r)
    w = np.asarray(w)
    assert w.ndim == 4
    Dl, d, Dr = a.shape
    cl, dout, din, cr = w.shape
    assert din == d and l.shape[0] == Dl and l.shape[1] == cl
    assert r.shape[0] == Dr and r.shape[1] == cr
    Dlp, Drp = l.shape[2], r.shape[2]
    # (chain_ops.py:273)  t1[(i,s),(kappa,j')] = a[(i,s),j] r[j,(kappa,j')]
    t1 = a.reshape(Dl * d, Dr) @ r.reshape(Dr, cr * Drp)
    # (chain_ops.py:276)  t2[(k,s'),(i,j')] = w[(k,s'),(s,kappa)] t1[(s,kappa),(i,j')]
    t1 = t1.reshape(Dl, d * cr, Drp).transpose(1, 0, 2).reshape(d * cr, Dl * Drp)
    t2 = w.reshape(cl * dout, d * cr) @ t1
    # (chain_ops.py:278)  out[i',(s',j')] = l[(i,k),i']^T t2[(i,k),(s',j')]
    t2 = t2.reshape(cl, dout, Dl, Drp).transpose(2, 0, 1, 3).reshape(Dl * cl, dout * Drp)
    out = l.reshape(Dl * cl, Dlp).T @ t2
    return out.reshape(Dlp, dout, Drp)


def apply_local_bond_contraction(c, l, r):
    """out[i',j'] = sum l[i,k,i'] c[i,j] r[j,k,j'].

    Restates pytenet/chain_ops.py:282-317 (tensordots at :314, :316).
    """
    c = np.asarray(c); l = _as3(l); r = _as3(r)
    assert c.ndim == 2
    Dl, Dr = c.shape
    chi = l.shape[1]
    assert l.shape[0] == Dl and r.shape[0] == Dr and r.shape[1] == chi
    Dlp, Drp = l.shape[2], r.shape[2]
    t = c @ r.reshape(Dr, chi * Drp)                      # t[i,(k,j')]
    out = l.reshape(Dl * chi, Dlp).T @ t.reshape(Dl * chi, Drp)
    return out


def contraction_operator_step_right(a, b, w, r):
    """r_next[i,k,i'] = sum a[i,s,j] r[j,kappa,j'] w[k,s',s,kappa] conj(b[i',s',j']).

    Restates pytenet/chain_ops.py:16-57 (tensordots at :50, :52, transpose :54, :56).
    """
    a = _as3(a); b = _as3(b); r = _as3(r)
    w = np.asarray(w)
    assert w.ndim == 4
    Dl, d, Dr = a.shape
    Dlp, dout, Drp = b.shape
    cl, dw_out, dw_in, cr = w.shape
    assert dw_in == d and dw_out == dout
    assert r.shape == (Dr, cr, Drp)
    t1 = a.reshape(Dl * d, Dr) @ r.reshape(Dr, cr * Drp)              # :50
    t1 = t1.reshape(Dl, d * cr, Drp).transpose(1, 0, 2).reshape(d * cr, Dl * Drp)
    t2 = w.reshape(cl * dout, d * cr) @ t1                            # :52
    t2 = t2.reshape(cl, dout, Dl, Drp).transpose(2, 0, 1, 3)          # :54
    t2 = t2.reshape(Dl * cl, dout * Drp)
    r_next = t2 @ b.conj().reshape(Dlp, dout * Drp).T                 # :56
    return r_next.reshape(Dl, cl, Dlp)


def contraction_operator_step_left(a, b, w, l):
    """l_next[j,kappa,j'] = sum l[i,k,i'] conj(b[i',s',j']) w[k,s',s,kappa] a[i,s,j].

    Restates pytenet/chain_ops.py:60-99 (tensordots at :94, :96, :98).
    """
    a = _as3(a); b = _as3(b); l = _as3(l)
    w = np.asarray(w)
    assert w.ndim == 4
    Dl, d, Dr = a.shape
    Dlp, dout, Drp = b.shape
    cl, dw_out, dw_in, cr = w.shape
    assert dw_in == d and dw_out == dout
    assert l.shape == (Dl, cl, Dlp)
    # :94  t[(i,k),(s',j')] = l[(i,k),i'] conj(b)[i',(s',j')]
    t = l.reshape(Dl * cl, Dlp) @ b.conj().reshape(Dlp, dout * Drp)
    # :96  t2[(s,kappa),(i,j')] = w[(k,s'),(s,kappa)]^T t[(k,s'),(i,j')]
    t = t.reshape(Dl, cl * dout, Drp).transpose(1, 0, 2).reshape(cl * dout, Dl * Drp)
    t2 = w.reshape(cl * dout, d * cr).T @ t
    # :98  l_next[j,(kappa,j')] = a[(i,s),j]^T t2[(i,s),(kappa,j')]
    t2 = t2.reshape(d, cr, Dl, Drp).transpose(2, 0, 1, 3).reshape(Dl * d, cr * Drp)
    l_next = a.reshape(Dl * d, Dr).T @ t2
    return l_next.reshape(Dr, cr, Drp)


def compute_right_operator_blocks(a_list, w_list):
    """All right environments, right to left, seeded with the integer [[[1]]].

    Restates pytenet/chain_ops.py:102-113.  Takes the tensor lists directly
    (``psi.a``, ``op.a``).
    """
    n = len(a_list)
    assert n == len(w_list)
    blocks = [None] * n
    blocks[n - 1] = np.array([[[1]]])
    for i in range(n - 2, -1, -1):
        blocks[i] = contraction_operator_step_right(
            a_list[i + 1], a_list[i + 1], w_list[i + 1], blocks[i + 1])
    return blocks


def merge_mps_pair(a0, a1):
    """Two-site tensor (Dl, d0*d1, Dr).  Restates pytenet/mps.py:528-535."""
    Dl, d0, D = a0.shape
    D2, d1, Dr = a1.shape
    assert D == D2
    t = a0.reshape(Dl * d0, D) @ a1.reshape(D, d1 * Dr)
    return t.reshape(Dl, d0 * d1, Dr)


def merge_mpo_pair(w0, w1):
    """Two-site MPO tensor (chi0, d*d, d*d, chi2).  Restates pytenet/mpo.py:314-322."""
    c0, p0, q0, c1 = w0.shape
    c1b, p1, q1, c2 = w1.shape
    assert c1 == c1b
    t = w0.reshape(c0 * p0 * q0, c1) @ w1.reshape(c1, p1 * q1 * c2)
    t = t.reshape(c0, p0, q0, p1, q1, c2).transpose(0, 1, 3, 2, 4, 5)
    return np.ascontiguousarray(t).reshape(c0, p0 * p1, q0 * q1, c2)


def flops_matvec(Dl, d, Dr, cl, cr, Dlp=None, Drp=None, dout=None):
    """F_alg of one effective-H matvec: dense, all-complex count of the
    reference's contraction order."""
    Dlp = Dl if Dlp is None else Dlp
    Drp = Dr if Drp is None else Drp
    dout = d if dout is None else dout
    return 8 * (Dl * d * Dr * cr * Drp + cl * dout * d * cr * Dl * Drp + Dlp * Dl * cl * dout * Drp)


def flops_bond(Dl, Dr, chi, Dlp=None, Drp=None):
    """F_alg of one zero-site bond matvec (8 * 2 chi D^3 for uniform D)."""
    Dlp = Dl if Dlp is None else Dlp
    Drp = Dr if Drp is None else Drp
    return 8 * (Dl * Dr * chi * Drp + Dlp * Dl * chi * Drp)
