"""Build libgpcsd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgpcsd_b200.so")
SOURCES = ["gpcsd_gemm.cu", "gpcsd_tma.cu", "gpcsd_kernels.cu", "gpcsd_eig.cu", "gpcsd_aux.cu", "gpcsd_plan.cu"]
HEADERS = ["common.h", "dmma_gemm.cuh", "dc_core.h", os.path.join("..", "..", "include", "gpcsd_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA sources into gpcsd_b200/libgpcsd_b200.so; returns the library path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + os.environ.get("GPCSD_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES + \
          ["-lcusolver", "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
