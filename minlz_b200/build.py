"""Builds minlz_b200/libminlz_cuda.so for sm_100a with nvcc (in-tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libminlz_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    hdr = os.path.join(os.path.dirname(HERE), "include", "minlz_cuda.h")
    return any(os.path.getmtime(f) > t for f in sources() + [hdr])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, os.path.join(CSRC, "mz_api.cu")]
    subprocess.check_call(cmd)
    return SO


def build_variant(name, defines):
    """Profiling / experiment builds: minlz_b200/libminlz_cuda_<name>.so with extra -D flags.
    Loaded instead of the product library when MINLZ_CUDA_SO points at it (see _lib.py)."""
    out = os.path.join(HERE, "libminlz_cuda_%s.so" % name)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-o", out, os.path.join(CSRC, "mz_api.cu")])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
