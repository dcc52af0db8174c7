"""
NumPy restatement of the sweep drivers that call the hot path (TEST
INFRASTRUCTURE, see oracle/__init__.py): right-orthonormalisation, one- and
two-site TDVP, one- and two-site DMRG, on plain lists of arrays.

State container: ``Chain`` with ``a`` (list of (Dl,d,Dr) arrays), ``qsite``
and ``qbonds`` -- the three attributes of pytenet.MPS the sweeps touch.
Operators are passed as (list of (chi_l,d,d,chi_r) arrays, list of bond
quantum-number arrays).
"""
import numpy as np
from . import contractions as oc
from . import lanczos as ol
from . import blocksparse as ob


class Chain:
    """Minimal stand-in for pytenet.MPS (mps.py:14-139): tensors + quantum numbers."""

    def __init__(self, a, qsite, qbonds):
        self.a = [np.array(t) for t in a]
        self.qsite = np.array(qsite)
        self.qbonds = [np.array(q) for q in qbonds]

    @property
    def nsites(self):
        return len(self.a)

    @property
    def bond_dims(self):
        return [t.shape[0] for t in self.a] + [self.a[-1].shape[2]]

    def to_vector(self):
        """Restates mps.py:292-301."""
        t = self.a[0]
        for nxt in self.a[1:]:
            t = oc.merge_mps_pair(t, nxt)
        return t.reshape(-1)


def local_orthonormalize_left_qr(a, a_next, qsite, qbonds):
    """Restates mps.py:460-473."""
    Dl, d, Dr = a.shape
    q, r, qb = ob.block_sparse_qr(a.reshape(Dl * d, Dr), ob.qnumber_flatten((qbonds[0], qsite)), qbonds[1])
    a = q.reshape(Dl, d, q.shape[1])
    a_next = (r @ a_next.reshape(a_next.shape[0], -1)).reshape((r.shape[0],) + a_next.shape[1:])
    return a, a_next, qb


def local_orthonormalize_right_qr(a, a_prev, qsite, qbonds):
    """Restates mps.py:476-491."""
    at = a.transpose(2, 1, 0)
    Dr, d, Dl = at.shape
    q, r, qb = ob.block_sparse_qr(at.reshape(Dr * d, Dl), ob.qnumber_flatten((-qbonds[1], qsite)), -qbonds[0])
    a = q.reshape(Dr, d, q.shape[1]).transpose(2, 1, 0)
    a_prev = np.tensordot(a_prev, r, (2, 1))
    return a, a_prev, -qb


def orthonormalize_right(psi):
    """Right-orthonormalise in place, return the norm.  Restates mps.py:163-177."""
    for i in range(psi.nsites - 1, 0, -1):
        psi.a[i], psi.a[i - 1], psi.qbonds[i] = local_orthonormalize_right_qr(
            psi.a[i], psi.a[i - 1], psi.qsite, psi.qbonds[i:i + 2])
    psi.a[0], t, psi.qbonds[0] = local_orthonormalize_right_qr(
        psi.a[0], np.array([[[1]]]), psi.qsite, psi.qbonds[:2])
    nrm = t[0, 0, 0].real
    if nrm < 0:
        psi.a[0] = -psi.a[0]
        nrm = -nrm
    return nrm


def orthonormalize_left(psi):
    """Left-orthonormalise in place, return the norm.  Restates mps.py:150-162."""
    n = psi.nsites
    for i in range(n - 1):
        psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = local_orthonormalize_left_qr(
            psi.a[i], psi.a[i + 1], psi.qsite, psi.qbonds[i:i + 2])
    psi.a[-1], t, psi.qbonds[-1] = local_orthonormalize_left_qr(
        psi.a[-1], np.array([[[1]]]), psi.qsite, psi.qbonds[-2:])
    nrm = t[0, 0, 0].real
    if nrm < 0:
        psi.a[-1] = -psi.a[-1]
        nrm = -nrm
    return nrm


def split_tensor_svd(a, qsite0, qsite1, qbonds_outer, svd_distr, tol=0):
    """Restates mps.py:538-568."""
    qsite0 = np.asarray(qsite0); qsite1 = np.asarray(qsite1)
    d0, d1 = len(qsite0), len(qsite1)
    Dl, dd, Dr = a.shape
    assert dd == d0 * d1
    q0 = ob.qnumber_flatten([qbonds_outer[0], qsite0])
    q1 = ob.qnumber_flatten([-qsite1, qbonds_outer[1]])
    u, sigma, v, qb = ob.split_block_sparse_matrix_svd(a.reshape(Dl * d0, d1 * Dr), q0, q1, tol)
    a0 = u.reshape(Dl, d0, len(sigma))
    a1 = v.reshape(len(sigma), d1, Dr)
    if svd_distr == "left":
        a0 = a0 * sigma
    elif svd_distr == "right":
        a1 = a1 * sigma[:, None, None]
    elif svd_distr == "sqrt":
        rt = np.sqrt(sigma)
        a0 = a0 * rt
        a1 = a1 * rt[:, None, None]
    else:
        raise ValueError('`svd_distr` parameter must be "left", "right" or "sqrt".')
    return a0, a1, qb


def _evolve_site(l, r, w, a, dt, k):
    """Restates tdvp.py:223-229."""
    shp = a.shape
    return ol.expm_krylov(lambda x: oc.apply_local_hamiltonian(x.reshape(shp), w, l, r).reshape(-1),
                          a.reshape(-1), -dt, k, hermitian=True).reshape(shp)


def _evolve_bond(l, r, c, dt, k):
    """Restates tdvp.py:232-238."""
    shp = c.shape
    return ol.expm_krylov(lambda x: oc.apply_local_bond_contraction(x.reshape(shp), l, r).reshape(-1),
                          c.reshape(-1), -dt, k, hermitian=True).reshape(shp)


def _ground_site(w, l, r, a, k):
    """Restates dmrg.py:181-189."""
    shp = a.shape
    ev, ritz = ol.eigh_krylov(lambda x: oc.apply_local_hamiltonian(x.reshape(shp), w, l, r).reshape(-1),
                              a.reshape(-1), k, 1)
    return ev[0], ritz[:, 0].reshape(shp)


def _prologue(w_list, w_qbonds, psi):
    """Shared start of all four drivers (tdvp.py:47-63, dmrg.py:41-56)."""
    n = len(w_list)
    assert n == psi.nsites
    nrm = orthonormalize_right(psi)
    rblocks = oc.compute_right_operator_blocks(psi.a, w_list)
    lblocks = [None] * n
    lblocks[0] = np.array([[[1]]], dtype=rblocks[0].dtype)
    for i, rb in enumerate(rblocks):
        assert ob.is_qsparse(rb, [psi.qbonds[i + 1], w_qbonds[i + 1], -psi.qbonds[i + 1]])
    return nrm, lblocks, rblocks


def tdvp_singlesite(w_list, w_qbonds, psi, dt, numsteps, numiter_lanczos=25):
    """Restates tdvp.py:26-118."""
    n = len(w_list)
    nrm, lb, rb = _prologue(w_list, w_qbonds, psi)
    k = numiter_lanczos
    for _ in range(numsteps):
        for i in range(n - 1):
            psi.a[i] = _evolve_site(lb[i], rb[i], w_list[i], psi.a[i], 0.5 * dt, k)
            Dl, d, Dr = psi.a[i].shape
            q, c, psi.qbonds[i + 1] = ob.block_sparse_qr(
                psi.a[i].reshape(Dl * d, Dr), ob.qnumber_flatten((psi.qbonds[i], psi.qsite)), psi.qbonds[i + 1])
            psi.a[i] = q.reshape(Dl, d, q.shape[1])
            lb[i + 1] = oc.contraction_operator_step_left(psi.a[i], psi.a[i], w_list[i], lb[i])
            c = _evolve_bond(lb[i + 1], rb[i], c, -0.5 * dt, k)
            psi.a[i + 1] = np.tensordot(c, psi.a[i + 1], (1, 0))
        i = n - 1
        psi.a[i] = _evolve_site(lb[i], rb[i], w_list[i], psi.a[i], dt, k)
        for i in range(n - 1, 0, -1):
            at = psi.a[i].transpose(2, 1, 0)
            Dr, d, Dl = at.shape
            q, c, qb = ob.block_sparse_qr(
                at.reshape(Dr * d, Dl), ob.qnumber_flatten((-psi.qbonds[i + 1], psi.qsite)), -psi.qbonds[i])
            psi.qbonds[i] = -qb
            psi.a[i] = q.reshape(Dr, d, q.shape[1]).transpose(2, 1, 0)
            rb[i - 1] = oc.contraction_operator_step_right(psi.a[i], psi.a[i], w_list[i], rb[i])
            c = _evolve_bond(lb[i], rb[i - 1], c.T, -0.5 * dt, k)
            psi.a[i - 1] = np.tensordot(psi.a[i - 1], c, (2, 0))
            psi.a[i - 1] = _evolve_site(lb[i - 1], rb[i - 1], w_list[i - 1], psi.a[i - 1], 0.5 * dt, k)
    return nrm


def tdvp_twosite(w_list, w_qbonds, psi, dt, numsteps, numiter_lanczos=25, tol_split=0):
    """Restates tdvp.py:121-220."""
    n = len(w_list)
    assert n >= 2
    nrm, lb, rb = _prologue(w_list, w_qbonds, psi)
    k = numiter_lanczos
    h2 = [oc.merge_mpo_pair(w_list[i], w_list[i + 1]) for i in range(n - 1)]
    qs = psi.qsite
    for _ in range(numsteps):
        for i in range(n - 2):
            m = oc.merge_mps_pair(psi.a[i], psi.a[i + 1])
            m = _evolve_site(lb[i], rb[i + 1], h2[i], m, 0.5 * dt, k)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = split_tensor_svd(
                m, qs, qs, (psi.qbonds[i], psi.qbonds[i + 2]), "right", tol=tol_split)
            lb[i + 1] = oc.contraction_operator_step_left(psi.a[i], psi.a[i], w_list[i], lb[i])
            psi.a[i + 1] = _evolve_site(lb[i + 1], rb[i + 1], w_list[i + 1], psi.a[i + 1], -0.5 * dt, k)
        i = n - 2
        m = oc.merge_mps_pair(psi.a[i], psi.a[i + 1])
        m = _evolve_site(lb[i], rb[i + 1], h2[i], m, dt, k)
        psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = split_tensor_svd(
            m, qs, qs, (psi.qbonds[i], psi.qbonds[i + 2]), "left", tol=tol_split)
        rb[i] = oc.contraction_operator_step_right(psi.a[i + 1], psi.a[i + 1], w_list[i + 1], rb[i + 1])
        for i in range(n - 3, -1, -1):
            psi.a[i + 1] = _evolve_site(lb[i + 1], rb[i + 1], w_list[i + 1], psi.a[i + 1], -0.5 * dt, k)
            m = oc.merge_mps_pair(psi.a[i], psi.a[i + 1])
            m = _evolve_site(lb[i], rb[i + 1], h2[i], m, 0.5 * dt, k)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = split_tensor_svd(
                m, qs, qs, (psi.qbonds[i], psi.qbonds[i + 2]), "left", tol=tol_split)
            rb[i] = oc.contraction_operator_step_right(psi.a[i + 1], psi.a[i + 1], w_list[i + 1], rb[i + 1])
    return nrm


def dmrg_singlesite(w_list, w_qbonds, psi, numsweeps, numiter_lanczos=25):
    """Restates dmrg.py:22-93."""
    n = len(w_list)
    _, lb, rb = _prologue(w_list, w_qbonds, psi)
    k = numiter_lanczos
    en_min = np.zeros(numsweeps)
    for sweep in range(numsweeps):
        en = 0
        for i in range(n - 1):
            en, psi.a[i] = _ground_site(w_list[i], lb[i], rb[i], psi.a[i], k)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = local_orthonormalize_left_qr(
                psi.a[i], psi.a[i + 1], psi.qsite, psi.qbonds[i:i + 2])
            lb[i + 1] = oc.contraction_operator_step_left(psi.a[i], psi.a[i], w_list[i], lb[i])
        for i in range(n - 1, 0, -1):
            en, psi.a[i] = _ground_site(w_list[i], lb[i], rb[i], psi.a[i], k)
            psi.a[i], psi.a[i - 1], psi.qbonds[i] = local_orthonormalize_right_qr(
                psi.a[i], psi.a[i - 1], psi.qsite, psi.qbonds[i:i + 2])
            rb[i - 1] = oc.contraction_operator_step_right(psi.a[i], psi.a[i], w_list[i], rb[i])
        psi.a[0], _, psi.qbonds[0] = local_orthonormalize_right_qr(
            psi.a[0], np.array([[[1]]]), psi.qsite, psi.qbonds[:2])
        en_min[sweep] = en
    return en_min


def dmrg_twosite(w_list, w_qbonds, psi, numsweeps, numiter_lanczos=25, tol_split=0):
    """Restates dmrg.py:96-178."""
    n = len(w_list)
    _, lb, rb = _prologue(w_list, w_qbonds, psi)
    k = numiter_lanczos
    en_min = np.zeros(numsweeps)
    h2 = [oc.merge_mpo_pair(w_list[i], w_list[i + 1]) for i in range(n - 1)]
    qs = psi.qsite
    for sweep in range(numsweeps):
        en = 0
        for i in range(n - 2):
            m = oc.merge_mps_pair(psi.a[i], psi.a[i + 1])
            en, m = _ground_site(h2[i], lb[i], rb[i + 1], m, k)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = split_tensor_svd(
                m, qs, qs, [psi.qbonds[i], psi.qbonds[i + 2]], "right", tol=tol_split)
            lb[i + 1] = oc.contraction_operator_step_left(psi.a[i], psi.a[i], w_list[i], lb[i])
        for i in range(n - 2, -1, -1):
            m = oc.merge_mps_pair(psi.a[i], psi.a[i + 1])
            en, m = _ground_site(h2[i], lb[i], rb[i + 1], m, k)
            psi.a[i], psi.a[i + 1], psi.qbonds[i + 1] = split_tensor_svd(
                m, qs, qs, [psi.qbonds[i], psi.qbonds[i + 2]], "left", tol=tol_split)
            rb[i] = oc.contraction_operator_step_right(psi.a[i + 1], psi.a[i + 1], w_list[i + 1], rb[i + 1])
        psi.a[0], _, psi.qbonds[0] = local_orthonormalize_right_qr(
            psi.a[0], np.array([[[1]]]), psi.qsite, psi.qbonds[:2])
        en_min[sweep] = en
    return en_min
