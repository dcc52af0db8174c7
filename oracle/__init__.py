"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU (NumPy) restatement of the PyTeNet effective-Hamiltonian hot path
(reference: cmendl/pytenet v1.3.0).  It exists to *check* the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``pytenet_b200/`` imports, links or executes anything from here; the
product path fails loudly when the CUDA library is missing.

Parity pin: every function in this package is checked against the real
reference (imported from /root/reference in the dev container) by
``tests/golden/make_golden.py``, which also writes the committed fixtures in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the oracle
against those fixtures on any machine (no reference needed), including the two
seeded known-answer values the reference records (doc/dmrg.ipynb:130,140;
doc/basics.ipynb:32,256,530).

Each function cites the reference file:line it restates.
"""
from .contractions import (            # noqa: F401
    apply_local_hamiltonian, apply_local_bond_contraction,
    contraction_operator_step_left, contraction_operator_step_right,
    compute_right_operator_blocks, merge_mps_pair, merge_mpo_pair,
    flops_matvec, flops_bond)
from .lanczos import (                 # noqa: F401
    lanczos_iteration, eigh_krylov, expm_krylov, eigh_tridiag)
