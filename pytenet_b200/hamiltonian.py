"""
Input generators: nearest-neighbour Hamiltonians as MPOs, written down directly
in the textbook "operator-valued matrix" form.

The reference builds these through its symbolic operator-graph toolchain
(pytenet/hamiltonian/*.py, opgraph.py) -- a one-off host-side construction that
is out of scope for the hot path.  The MPOs here
represent the same operators (checked against the reference's tensors in
tests/test_hamiltonian_inputs.py) with the same quantum-number conventions,
so they can feed the sweeps and the benchmark on a box that has no reference
checkout.  MPO tensors produced by the reference itself can be wrapped with
`MPO.from_tensors`.
"""
import numpy as np

from .mpo import MPO
from .scalars import encode_quantum_number_pair

__all__ = ["heisenberg_xxz_1d_mpo", "ising_1d_mpo", "fermi_hubbard_1d_mpo", "load_cached_mpo"]


def _chain_tensors(qbond, wbulk, nsites, first_row, last_col):
    """Host tensors and bond quantum numbers of an open chain built from one bulk tensor:
    the first site keeps row `first_row`, the last site column `last_col` of the
    operator-valued matrix."""
    tensors = []
    qbonds = []
    for i in range(nsites):
        w = wbulk
        ql = qbond
        if i == 0:
            w = w[first_row:first_row + 1]
            ql = qbond[first_row:first_row + 1]
        if i == nsites - 1:
            w = w[:, :, :, last_col:last_col + 1]
        tensors.append(np.ascontiguousarray(w))
        qbonds.append(np.asarray(ql))
    qbonds.append(np.asarray(qbond[last_col:last_col + 1]))
    return tensors, qbonds


def _chain_mpo(qsite, qbond, wbulk, nsites, first_row, last_col, device=None):
    tensors, qbonds = _chain_tensors(qbond, wbulk, nsites, first_row, last_col)
    return MPO.from_tensors(qsite, qbonds, tensors, device=device)


def _xxz_bulk(J, D, h):
    sup = np.array([[0., 1.], [0., 0.]])
    sdn = np.array([[0., 0.], [1., 0.]])
    sz = np.array([[0.5, 0.], [0., -0.5]])
    id2 = np.identity(2)
    w = np.zeros((5, 2, 2, 5))
    # bond states: 0 = nothing yet, 1 = S+ placed, 2 = S- placed, 3 = Sz placed, 4 = done
    w[0, :, :, 0] = id2
    w[0, :, :, 1] = 0.5 * J * sup
    w[0, :, :, 2] = 0.5 * J * sdn
    w[0, :, :, 3] = D * sz
    w[0, :, :, 4] = -h * sz
    w[1, :, :, 4] = sdn
    w[2, :, :, 4] = sup
    w[3, :, :, 4] = sz
    w[4, :, :, 4] = id2
    return [1, -1], np.array([0, 2, -2, 0, 0]), w, 0, 4


def heisenberg_xxz_1d_mpo(nsites: int, J: float, D: float, h: float, device=None) -> MPO:
    """
    XXZ Heisenberg chain `sum J (X X + Y Y) + D Z Z - h Z` (spin 1/2), MPO bond
    dimension 5; physical quantum numbers are 2 Sz = (1, -1) as in
    pytenet/hamiltonian/heisenberg.py:14-55.
    """
    qsite, qb, w, first, last = _xxz_bulk(J, D, h)
    return _chain_mpo(qsite, qb, w, nsites, first, last, device)


def _ising_bulk(J, h, g):
    sx = np.array([[0., 1.], [1., 0.]])
    sz = np.array([[1., 0.], [0., -1.]])
    id2 = np.identity(2)
    w = np.zeros((3, 2, 2, 3))
    w[0, :, :, 0] = id2
    w[0, :, :, 1] = J * sz
    w[0, :, :, 2] = h * sz + g * sx
    w[1, :, :, 2] = sz
    w[2, :, :, 2] = id2
    return [0, 0], np.zeros(3, dtype=int), w, 0, 2


def ising_1d_mpo(nsites: int, J: float, h: float, g: float, device=None) -> MPO:
    """
    Ising chain `sum J Z Z + h Z + g X` (Pauli matrices), MPO bond dimension 3, all
    quantum numbers zero (pytenet/hamiltonian/ising.py:14-69).
    """
    qsite, qb, w, first, last = _ising_bulk(J, h, g)
    return _chain_mpo(qsite, qb, w, nsites, first, last, device)


def _fermi_hubbard_bulk(t, u, mu):
    qsite = [encode_quantum_number_pair(n, s) for n, s in zip([0, 1, 1, 2], [0, -1, 1, 0])]
    id2 = np.identity(2)
    cr = np.array([[0., 0.], [1., 0.]])      # creation
    an = np.array([[0., 1.], [0., 0.]])      # annihilation
    nu = np.array([[0., 0.], [0., 1.]])
    z = np.array([[1., 0.], [0., -1.]])
    id4 = np.identity(4)
    onsite = -mu * (np.kron(nu, id2) + np.kron(id2, nu)) + u * np.diag([0.25, -0.25, -0.25, 0.25])
    w = np.zeros((6, 4, 4, 6))
    # bond states: 0 start; 1..4 pending hopping partner; 5 done
    w[0, :, :, 0] = id4
    w[0, :, :, 1] = -t * np.kron(cr, z)      # c+_up (string over own down mode) ... a_up on the next site
    w[0, :, :, 2] = -t * np.kron(an, z)      # a_up ... c+_up
    w[0, :, :, 3] = -t * np.kron(id2, cr)    # c+_dn ... (string over next up mode) a_dn
    w[0, :, :, 4] = -t * np.kron(id2, an)    # a_dn ... c+_dn
    w[0, :, :, 5] = onsite
    w[1, :, :, 5] = np.kron(an, id2)
    w[2, :, :, 5] = np.kron(cr, id2)
    w[3, :, :, 5] = np.kron(z, an)
    w[4, :, :, 5] = np.kron(z, cr)
    w[5, :, :, 5] = id4
    qb = np.array([0,
                   encode_quantum_number_pair(1, 1), encode_quantum_number_pair(-1, -1),
                   encode_quantum_number_pair(1, -1), encode_quantum_number_pair(-1, 1),
                   0])
    return qsite, qb, w, 0, 5


def fermi_hubbard_1d_mpo(nsites: int, t: float, u: float, mu: float, device=None) -> MPO:
    """
    Fermi-Hubbard chain with nearest-neighbour hopping `t`, interaction
    `u (n_up - 1/2)(n_dn - 1/2)` and chemical potential `mu`, Jordan-Wigner ordered
    (up, down) per site; MPO bond dimension 6; physical quantum numbers are
    (particle number, spin) pairs (pytenet/hamiltonian/fermi_hubbard.py:14-85).
    """
    qsite, qb, w, first, last = _fermi_hubbard_bulk(t, u, mu)
    return _chain_mpo(qsite, qb, w, nsites, first, last, device)


def cached_mpo_tensors(path):
    """Host side of `load_cached_mpo`: (qsite, qbonds, dense NumPy tensors) rebuilt from the sparse cache file."""
    z = np.load(path)
    nsites = len([k for k in z.files if k.endswith("_shape")])
    tensors = []
    for i in range(nsites):
        shape = tuple(int(x) for x in z[f"w{i}_shape"])
        flat = np.zeros(int(np.prod(shape)), dtype=z[f"w{i}_val"].dtype)
        flat[z[f"w{i}_idx"].astype(np.int64)] = z[f"w{i}_val"]
        tensors.append(flat.reshape(shape))
    qbonds = [z[f"qb{i}"] for i in range(nsites + 1)]
    return z["qsite"], qbonds, tensors


def load_cached_mpo(path, device=None):
    """
    MPO from a sparse cache file written by `tests/golden/make_molecular_mpo.py` (per site: tensor shape, flat
    indices and values of the non-zero entries; plus `qsite` and the bond quantum numbers).  This is how
    BASELINE config 4 -- `molecular_hamiltonian_mpo(tkin, vint, optimize=False)` on 32 orbitals
    (pytenet/hamiltonian/molecular.py:612), a nine-minute symbolic host construction in the reference -- reaches
    a GPU box without a reference checkout: the tensors are the reference's own, bit for bit.
    """
    qsite, qbonds, tensors = cached_mpo_tensors(path)
    return MPO.from_tensors(qsite, qbonds, tensors, device=device)
