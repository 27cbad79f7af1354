"""TEST INFRASTRUCTURE ONLY: compile the kernel sources under fhe-si_b200/csrc with g++
against tests/emu/cuda_emu.h (CUDA threads = OS threads) so that `pytest -m "not gpu"` can
check kernel logic against the oracle on a GPU-less machine.  The product never loads this."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "fhe-si_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libfhesi_emu.so")


def build_emu(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "cuda_emu.h"),
                                                               os.path.join(ROOT, "include", "fhesi.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    import sys
    sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200"))
    from buildlock import build_lock, publish
    fresh = lambda: os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps)
    with build_lock(OUT):  # torchrun ranks and xdist workers may build concurrently
        if not force and fresh():
            return OUT
        tmp = OUT + ".tmp.%d" % os.getpid()
        cmd = ["g++", "-std=c++20", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-DFHESI_EMU=1",
               "-include", os.path.join(HERE, "cuda_emu.h"), "-x", "c++", os.path.join(CSRC, "fhesi_lib.cu"),
               "-o", tmp]
        subprocess.check_call(cmd)
        publish(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build_emu(force=True))
