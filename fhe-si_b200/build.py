"""Build libfhesi_b200.so (sm_100a) in-tree with nvcc.  No JIT cache: the .so travels with
the repo snapshot to the GPU box."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfhesi_b200.so")
SOURCES = ["fhesi_lib.cu"]
HEADERS = ["modarith.cuh", "kernels_generic.cuh", "kernels_fused.cuh", "kernels_fused2k.cuh", "../../include/fhesi.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-extended-lambda", "-diag-suppress", "550",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build(force=False, verbose=False):
    from buildlock import build_lock, publish
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "fhesi.h")]
    if not force and not _stale(LIB, deps):
        return LIB
    with build_lock(LIB):  # one builder per tree; the other ranks wait, then find it fresh
        if not force and not _stale(LIB, deps):
            return LIB
        nvcc = nvcc_path()
        if not os.path.exists(nvcc):
            if os.path.exists(LIB):
                return LIB  # GPU box without a toolchain: use the .so that travelled with the snapshot
            raise RuntimeError("nvcc not found and no prebuilt libfhesi_b200.so")
        tmp = LIB + ".tmp.%d" % os.getpid()
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
        subprocess.check_call(cmd, cwd=CSRC)
        publish(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
