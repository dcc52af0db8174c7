"""
In-tree build of the CUDA library (nvcc, sm_100a only).

    python -m pytenet_b200._build          # or __graft_entry__.build()

Produces pytenet_b200/libpytenet_b200.so next to this file.  The .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpytenet_b200.so")
SOURCES = ["block_qr.cu", "block_svd.cu", "chain_ops.cu", "dense_svd.cu", "heff_small.cu", "host_entry.cu", "krylov.cu", "lanczos_heff.cu", "lanczos_small.cu", "sector_packed.cu", "sharded.cu", "wapply.cu", "probe.cu"]


def _dependencies():
    """Every file the library is compiled from (all of csrc/ plus the public header)."""
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(HERE, "..", "include", "pytenet_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-warn-spills",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libpytenet_b200.so")
    return exe


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _dependencies())


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
