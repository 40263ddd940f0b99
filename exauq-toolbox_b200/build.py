"""Build libexauq_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libexauq_b200.so")
SOURCES = [os.path.join(HERE, "csrc", "exq_api.cu")]
HEADERS = [
    os.path.join(HERE, "csrc", n)
    for n in ("exq_common.cuh", "gemm_dmma.cuh", "leaf.cuh", "kernels.cuh", "ozaki_i8.cuh", "ozaki_gemm.cuh", "fused_node.cuh", "pei_fused.cuh", "pivot.cuh")
] + [os.path.join(os.path.dirname(HERE), "include", "exauq_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("EXQ_NVCC_EXTRA", "").split()      # experiments only (e.g. -DEXQ_GRAD_SKIP_MATH)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
