"""CPU: the directly written MPO input generators represent the same operators, with
consistent quantum numbers, as the tensors the reference's builders produced (golden fixtures)."""
import os

import numpy as np

import oracle
import oracle.blocksparse as ob
from pytenet_b200 import hamiltonian as ham


def dense(tensors):
    t = tensors[0]
    for nxt in tensors[1:]:
        t = oracle.merge_mpo_pair(t, nxt)
    assert t.shape[0] == 1 and t.shape[3] == 1
    return t[0, :, :, 0]


def check_qnumbers(qsite, qbonds, tensors):
    for i, w in enumerate(tensors):
        assert ob.is_qsparse(w, [qbonds[i], qsite, -np.asarray(qsite), -qbonds[i + 1]])


def test_xxz_matches_reference_operator(golden_dir):
    z = np.load(os.path.join(golden_dir, "tdvp_xxz_L10.npz"))
    n = int(z["h/nsites"])
    ref = dense([z[f"h/w{i}"] for i in range(n)])
    qsite, qb, w, first, last = ham._xxz_bulk(1.0, 0.8, -0.1)
    tensors, qbonds = ham._chain_tensors(qb, w, n, first, last)
    assert np.max(np.abs(dense(tensors) - ref)) < 1e-13
    check_qnumbers(qsite, qbonds, tensors)
    assert [t.shape[0] for t in tensors] + [1] == [1] + [5] * (n - 1) + [1]


def test_fermi_hubbard_matches_reference_operator(golden_dir):
    z = np.load(os.path.join(golden_dir, "dmrg_fermi_hubbard_L6.npz"))
    n = int(z["h/nsites"])
    ref = dense([z[f"h/w{i}"] for i in range(n)])
    qsite, qb, w, first, last = ham._fermi_hubbard_bulk(1.0, 4.0, 1.5)
    tensors, qbonds = ham._chain_tensors(qb, w, n, first, last)
    assert np.max(np.abs(dense(tensors) - ref)) < 1e-13
    assert list(qsite) == list(z["h/qsite"])
    check_qnumbers(qsite, qbonds, tensors)


def test_ising_is_hermitian_and_local():
    qsite, qb, w, first, last = ham._ising_bulk(1.0, 0.3, -0.7)
    tensors, qbonds = ham._chain_tensors(qb, w, 5, first, last)
    h = dense(tensors)
    assert np.allclose(h, h.T)
    sx = np.array([[0., 1.], [1., 0.]]); sz = np.diag([1., -1.]); id2 = np.identity(2)
    def op(o, i):
        m = np.identity(1)
        for j in range(5):
            m = np.kron(m, o if j == i else id2)
        return m
    ref = sum(op(sz, i) @ op(sz, i + 1) for i in range(4)) + sum(0.3 * op(sz, i) - 0.7 * op(sx, i) for i in range(5))
    assert np.max(np.abs(h - ref)) < 1e-13
