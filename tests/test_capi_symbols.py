"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares
(no compute calls -- there is no GPU on the CPU test box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pytenet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_survey_minimum():
    names = declared_symbols()
    for stem in ["apply_local_hamiltonian", "apply_local_bond_contraction", "env_step_left", "env_step_right",
                 "lanczos_ortho_step"]:
        for sfx in ("_d", "_z"):
            assert f"ptb_{stem}{sfx}" in names
    assert "ptb_krylov_combine" in names
    # communicator handle + sharded variants ("*_comm_init/_destroy and sharded variants of the first four";
    # the sharded right update is the left one on mirrored tensors, csrc/sharded.cu)
    for name in ["ptb_comm_init", "ptb_comm_destroy", "ptb_comm_unique_id", "ptb_apply_local_hamiltonian_sharded",
                 "ptb_apply_local_bond_contraction_sharded", "ptb_env_step_left_sharded", "ptb_sharded_precontract"]:
        assert name in names
    assert "ptb_apply_local_hamiltonian_workspace_bytes" in names


def test_library_exports_every_declared_symbol(cuda_lib):
    from pytenet_b200 import _lib
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in include/pytenet_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in pytenet_b200/_lib.py"
    assert cuda_lib.ptb_version() >= 100
    assert cuda_lib.ptb_status_string(0) == b"ok"
    assert b"workspace" in cuda_lib.ptb_status_string(-3)


def test_host_side_argument_checks(cuda_lib):
    """Bad arguments are rejected on the host before any launch (no GPU needed)."""
    assert cuda_lib.ptb_gemm(0, 0, 0, 0, 4, 4, 4, None, 4, None, 4, None, 4, 1, 0, 0, 0, 0, None) == -1
    assert cuda_lib.ptb_gemm(7, 0, 0, 0, 4, 4, 4, 16, 4, 16, 4, 16, 4, 1, 0, 0, 0, 0, None) == -2
    nb = cuda_lib.ptb_apply_local_hamiltonian_workspace_bytes(1, 8, 2, 8, 5, 5, 2, 8, 8)
    # t1 + t2 + split-K partials (8 copies of the output or 4 of t1 when the first GEMM is small)
    assert nb == 2 * 8 * 2 * 5 * 8 * 16 + max(8 * (8 * 2 * 8), 4 * (8 * 2 * 5 * 8)) * 16
    # workspace too small
    assert cuda_lib.ptb_apply_local_hamiltonian_z(16, 16, 0, 16, 16, 16, 8, 2, 8, 5, 5, 2, 8, 8, 16, 10, None) == -3
    # non-positive extent
    assert cuda_lib.ptb_apply_local_hamiltonian_z(16, 16, 0, 16, 16, 16, 0, 2, 8, 5, 5, 2, 8, 8, 16, nb, None) == -1


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (or /root/reference)."""
    pkg = os.path.join(ROOT, "pytenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_local_step_plan_rules_on_the_host(cuda_lib):
    """ptb_local_step_small_fits is host arithmetic (cluster size, shared-memory layout, slots per thread): the
    shapes of the launch-latency regime qualify, larger ones and bad arguments do not; argument checks of the entry
    itself reject before any launch."""
    fits = cuda_lib.ptb_local_step_small_fits
    for dims in [(4, 4, 4, 3, 3), (1, 2, 2, 1, 4), (4, 2, 8, 5, 5), (8, 2, 16, 5, 5), (16, 2, 28, 5, 5),
                 (28, 2, 16, 5, 5), (16, 1, 28, 5, 5), (28, 1, 28, 5, 5), (48, 2, 4, 5, 5)]:
        assert fits(*dims, 5) == 1, dims                       # README config (D <= 28), METTS (D = 4), chain edges
        assert fits(*dims, 64) == 1 and fits(*dims, 65) == 0   # numiter <= 64 (tridiagonal solve on the device)
    # more than 400 000 multiply-adds per matvec: the multi-kernel path
    assert fits(48, 2, 40, 5, 5, 6) == 0 and fits(64, 2, 64, 5, 5, 5) == 0 and fits(2048, 2, 2048, 5, 5, 25) == 0
    assert fits(0, 2, 4, 3, 3, 5) == 0 and fits(4, 2, 4, 3, 3, 0) == 0 and fits(4, -1, 4, 3, 3, 5) == 0
    assert cuda_lib.ptb_local_step_small_workspace_bytes(1, 16, 2, 28, 5, 5) >= 16
    args = [16, 2, 28, 5, 5, 5]
    # null pointers / unknown dtype / zero-site problem with a physical index: rejected on the host
    assert cuda_lib.ptb_local_step_small(1, None, None, 0, 16, 16, *args, 16, 16, 0, 0.0, 0.0, 1, None, None, 0, None) == -1
    assert cuda_lib.ptb_local_step_small(7, 16, 16, 0, 16, 16, *args, 16, 16, 0, 0.0, 0.0, 1, None, None, 0, None) == -2
    assert cuda_lib.ptb_local_step_small(1, 16, None, 0, 16, 16, *args, 16, 16, 0, 0.0, 0.0, 1, None, None, 0, None) == -1
    # real state with a complex time step needs a complex output
    assert cuda_lib.ptb_local_step_small(0, 16, 16, 0, 16, 16, *args, 16, 16, 1, 0.0, 0.5, 0, 16, None, 0, None) == -1
