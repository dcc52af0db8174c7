"""
pytenet_b200 -- B200-native drop-in for the effective-Hamiltonian hot path of
cmendl/pytenet (L.W.A.R contraction inside every Lanczos/Krylov step of TDVP and
DMRG, plus the left/right environment updates).

The public names mirror pytenet's flat namespace (pytenet/__init__.py:10-36) for
the modules on that path.  All arithmetic runs in hand-written sm_100a CUDA
kernels behind the C ABI of include/pytenet_b200.h; there is no CPU fallback.
"""
from .scalars import *            # noqa: F401,F403
from .block_sparse_util import *  # noqa: F401,F403
from .bond_ops import *           # noqa: F401,F403
from .mps import *                # noqa: F401,F403
from .mpo import *                # noqa: F401,F403
from .chain_ops import *          # noqa: F401,F403
from .krylov import *             # noqa: F401,F403
from .tdvp import *               # noqa: F401,F403
from .dmrg import *               # noqa: F401,F403
from .hamiltonian import *        # noqa: F401,F403
from .metts import *              # noqa: F401,F403
from .sectors import *            # noqa: F401,F403
from .sharded import *            # noqa: F401,F403
from .sharded_dmrg import *       # noqa: F401,F403

__version__ = "0.1.0"
