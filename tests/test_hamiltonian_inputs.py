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


def test_cached_molecular_mpo_config4(golden_dir):
    """BASELINE config 4 input: the 32-orbital `molecular_hamiltonian_mpo(optimize=False)` tensors cached sparse by
    tests/golden/make_molecular_mpo.py (which imports the reference).  Shapes / bond dimensions as SURVEY 8(d)
    measured them, every non-zero entry allowed by the quantum numbers, centre tensor 16.8 % dense."""
    import os
    from pytenet_b200.hamiltonian import cached_mpo_tensors
    qsite, qbonds, ws = cached_mpo_tensors(os.path.join(golden_dir, "molecular_mpo_N32.npz"))
    assert list(qsite) == [0, 1] and len(ws) == 32
    bonds = [w.shape[0] for w in ws] + [ws[-1].shape[3]]
    assert bonds == [1, 5, 72, 81, 94, 111, 132, 157, 186, 219, 256, 297, 342, 391, 444, 501, 562,
                     501, 444, 391, 342, 297, 256, 219, 186, 157, 132, 111, 94, 81, 72, 5, 1]
    assert ws[16].shape == (562, 2, 2, 501) and ws[16].dtype == np.float64
    assert abs(np.count_nonzero(ws[16]) / ws[16].size - 0.168) < 0.002
    vals, counts = np.unique(qbonds[16], return_counts=True)
    assert dict(zip(vals.tolist(), counts.tolist())) == {-2: 120, -1: 32, 0: 258, 1: 32, 2: 120}
    for i, w in enumerate(ws):
        assert [len(qbonds[i]), len(qbonds[i + 1])] == [w.shape[0], w.shape[3]]
        qsum = (qbonds[i][:, None, None, None] + qsite[None, :, None, None] - qsite[None, None, :, None]
                - qbonds[i + 1][None, None, None, :])
        assert np.all(w[qsum != 0] == 0), i


def test_cached_molecular_mpo_small_is_hermitian(golden_dir):
    """The 10-orbital sibling of the config-4 cache (same generator): the full operator is Hermitian and the
    reference's `dmrg_singlesite` energies stored with it lie above its exact ground state in the sector."""
    import os
    from pytenet_b200.hamiltonian import cached_mpo_tensors
    path = os.path.join(golden_dir, "molecular_mpo_N10.npz")
    qsite, qbonds, ws = cached_mpo_tensors(path)
    t = ws[0]
    for w in ws[1:]:
        t = np.einsum("kpqm,mrsn->kprqsn", t, w).reshape(t.shape[0], t.shape[1] * 2, t.shape[2] * 2, w.shape[3])
    hm = t[0, :, :, 0]
    assert np.allclose(hm, hm.T)
    nocc = np.array([bin(i).count("1") for i in range(2 ** 10)])
    sec = np.where(nocc == 5)[0]
    e0 = np.linalg.eigvalsh(hm[np.ix_(sec, sec)])[0]
    en = np.load(path)["dmrg_single_en"]
    assert np.all(en >= e0 - 1e-10) and en[-1] - e0 < 1.0
