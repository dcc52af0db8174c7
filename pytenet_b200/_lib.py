"""
ctypes binding of libpytenet_b200.so (the C ABI declared in include/pytenet_b200.h).

There is no CPU fallback: if the library is missing or fails to load, importing
any compute entry point raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpytenet_b200.so")

PTB_REAL64 = 0
PTB_COMPLEX128 = 1

_i64 = ctypes.c_int64
_int = ctypes.c_int
_ptr = ctypes.c_void_p
_sz = ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/pytenet_b200.h line by line
_DIMS8 = [_i64] * 8
SIGNATURES = {
    "ptb_version": (_int, []),
    "ptb_status_string": (ctypes.c_char_p, [_int]),
    "ptb_gemm_engine": (_int, [_int, _int, _int, _int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                               _i64, _i64, _i64, _i64, _int, _ptr]),
    "ptb_gemm_splitk": (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                               _i64, _i64, _i64, _i64, _int, _int, _ptr, _sz, _ptr]),
    "ptb_gemm_banded": (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                               _i64, _i64, _i64, _i64, _int, _ptr, _ptr]),
    "ptb_gemm_segmented": (_int, [_int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64, _i64, _i64, _i64, _i64,
                                  _int, _ptr, _ptr, _ptr, _ptr]),
    "ptb_gemm_tile_shape": (_int, [_int, ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "ptb_gemm": (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                        _i64, _i64, _i64, _i64, _int, _ptr]),
    "ptb_apply_local_hamiltonian_workspace_bytes": (_sz, [_int] + _DIMS8),
    "ptb_apply_local_hamiltonian_z": (_int, [_ptr, _ptr, _int, _ptr, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_hamiltonian_d": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_hamiltonian_csr_z": (_int, [_ptr, _ptr, _ptr, _ptr, _int, _ptr, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_hamiltonian_csr_d": (_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_hamiltonian_host_workspace_bytes": (_sz, [_int, _int] + _DIMS8),
    "ptb_apply_local_hamiltonian_host": (_int, [_int, _int, _ptr, _ptr, _ptr, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_wapply_csr": (_int, [_int, _int, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "ptb_wapply_csr_masked": (_int, [_int, _int, _i64, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr]),
    "ptb_apply_local_bond_contraction_workspace_bytes": (_sz, [_int] + [_i64] * 5),
    "ptb_apply_local_bond_contraction_z": (_int, [_ptr] * 4 + [_i64] * 5 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_bond_contraction_d": (_int, [_ptr] * 4 + [_i64] * 5 + [_ptr, _sz, _ptr]),
    "ptb_env_step_workspace_bytes": (_sz, [_int] + _DIMS8),
    "ptb_env_step_left_z": (_int, [_ptr, _ptr, _ptr, _int, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_env_step_left_d": (_int, [_ptr] * 5 + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_env_step_right_z": (_int, [_ptr, _ptr, _ptr, _int, _ptr, _ptr] + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_env_step_right_d": (_int, [_ptr] * 5 + _DIMS8 + [_ptr, _sz, _ptr]),
    "ptb_lanczos_scratch_bytes": (_sz, []),
    "ptb_lanczos_start_d": (_int, [_i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ptb_lanczos_start_z": (_int, [_i64, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "ptb_lanczos_ortho_step_d": (_int, [_i64] + [_ptr] * 9),
    "ptb_lanczos_ortho_step_z": (_int, [_i64] + [_ptr] * 9),
    "ptb_lanczos_alpha_d": (_int, [_i64] + [_ptr] * 5),
    "ptb_lanczos_alpha_z": (_int, [_i64] + [_ptr] * 5),
    "ptb_krylov_combine": (_int, [_int, _int, _i64, _i64, _ptr, _i64, _ptr, _ptr, _ptr]),
    "ptb_heff_lanczos_workspace_bytes": (_sz, [_int] + [_i64] * 5),
    "ptb_heff_lanczos": (_int, [_int, _ptr, _ptr, _int, _ptr, _ptr, _ptr, _ptr, _ptr] + [_i64] * 5
                         + [_int, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "ptb_bond_lanczos_workspace_bytes": (_sz, [_int] + [_i64] * 3),
    "ptb_bond_lanczos": (_int, [_int, _ptr, _ptr, _ptr] + [_i64] * 3 + [_int, _ptr, _ptr, _ptr, _ptr, _sz, _ptr]),
    "ptb_local_step_small_fits": (_int, [_i64] * 5 + [_int]),
    "ptb_local_step_small_workspace_bytes": (_sz, [_int] + [_i64] * 5),
    "ptb_local_step_small": (_int, [_int, _ptr, _ptr, _int, _ptr, _ptr] + [_i64] * 5 + [_int, _ptr, _ptr, _int,
                                    ctypes.c_double, ctypes.c_double, _int, _ptr, _ptr, _sz, _ptr]),
    "ptb_krylov_expm_workspace_bytes": (_sz, []),
    "ptb_krylov_expm_apply": (_int, [_int, _i64, _int, _ptr, _i64, _ptr, ctypes.c_double, ctypes.c_double, _int, _ptr,
                                     _ptr, _ptr]),
    "ptb_block_qr_max_block_bytes": (_sz, []),
    "ptb_block_qr": (_int, [_int, _ptr, _i64, _int, _ptr, _int, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr]),
    "ptb_block_svd_max_block_bytes": (_sz, []),
    "ptb_block_svd": (_int, [_int, _ptr, _i64, _int, _ptr, _int, _ptr, _ptr, _ptr, _i64, _ptr, _ptr, _i64, _ptr]),
    "ptb_gemm_grouped": (_int, [_int, _ptr, _ptr, _ptr, _ptr, _int, _ptr]),
    "ptb_gemm_grouped_tile_shape": (_int, [_int, _int, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "ptb_gemm_grouped_v": (_int, [_int, _int, _ptr, _ptr, _ptr, _ptr, _int, _ptr]),
    "ptb_block_gather": (_int, [_int, _ptr, _ptr, _ptr, _ptr, _ptr, _int, _ptr]),
    "ptb_svd_polar_workspace_bytes": (_int, [_int, _i64, _i64, ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "ptb_svd_polar": (_int, [_int, _i64, _i64, _ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _sz, _ptr, _sz, _ptr,
                             ctypes.POINTER(ctypes.c_double), _ptr]),
    "ptb_comm_unique_id": (_int, [_ptr]),
    "ptb_comm_init": (_int, [ctypes.POINTER(_ptr), _int, _int, _ptr]),
    "ptb_comm_destroy": (_int, [_ptr]),
    "ptb_comm_info": (_int, [_ptr, ctypes.POINTER(_int), ctypes.POINTER(_int)]),
    "ptb_allreduce_sum": (_int, [_ptr, _int, _ptr, _i64, _ptr]),
    "ptb_sharded_precontract": (_int, [_int, _int, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr]),
    "ptb_apply_local_hamiltonian_sharded_workspace_bytes": (_sz, [_int] + [_i64] * 7),
    "ptb_apply_local_hamiltonian_sharded": (_int, [_ptr, _int, _ptr, _ptr, _ptr, _ptr] + [_i64] * 7 + [_ptr, _sz, _ptr]),
    "ptb_env_step_left_sharded_workspace_bytes": (_sz, [_int] + [_i64] * 4),
    "ptb_env_step_left_sharded": (_int, [_int, _ptr, _ptr, _ptr, _ptr] + [_i64] * 7 + [_ptr, _sz, _ptr]),
    "ptb_apply_local_bond_contraction_sharded_workspace_bytes": (_sz, [_int] + [_i64] * 3),
    "ptb_apply_local_bond_contraction_sharded": (_int, [_ptr, _int, _ptr, _ptr, _ptr, _ptr] + [_i64] * 5
                                                 + [_ptr, _sz, _ptr]),
    "ptb_probe_fp64_pipe": (_int, [_int, _int, _int, _ptr, ctypes.POINTER(ctypes.c_double), _ptr]),
}



class SectorTables(ctypes.Structure):
    """ptb_sector_tables of include/pytenet_b200.h (device pointers; None = NULL)."""
    _fields_ = [("ktab", _ptr), ("seg_ptr", _ptr), ("segs", _ptr), ("sel_off", _ptr), ("order", _ptr)]


class SvdJob(ctypes.Structure):
    """ptb_svd_job of include/pytenet_b200.h."""
    _fields_ = [("rows", _i64), ("cols", _i64), ("a", _ptr), ("lda", _i64), ("s", _ptr), ("u", _ptr), ("ldu", _i64),
                ("v", _ptr), ("ldv", _i64), ("device_ws", _ptr), ("device_bytes", _sz), ("info", _ptr),
                ("err_sigma", ctypes.c_double), ("status", _int), ("reserved", _int)]


SIGNATURES["ptb_svd_polar_batch"] = (_int, [_int, _int, ctypes.POINTER(SvdJob), _int, _ptr])
SIGNATURES["ptb_gemm_sector"] = (_int, [_int, _int, _int, _int, _i64, _i64, _i64, _ptr, _i64, _ptr, _i64, _ptr, _i64,
                                        _i64, _i64, _i64, _i64, _int, ctypes.POINTER(SectorTables), _ptr])

_lib = None


class PtbError(RuntimeError):
    """Non-zero status from the C ABI."""


def load():
    """Load the shared library (once).  Raises ImportError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m pytenet_b200._build` "
            "(there is no CPU fallback for the pytenet_b200 compute path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status, what=""):
    """Convert a C-ABI status into the reference's error behaviour."""
    if status == 0:
        return
    msg = load().ptb_status_string(int(status)).decode()
    if status == -1:
        # the reference asserts on ranks / shapes (chain_ops.py:45-48, 89-92, 268-271, 310-312)
        raise AssertionError(f"{what}: {msg}")
    raise PtbError(f"{what}: status {status}: {msg}")
