"""Compile the reference's own client programs, UNCHANGED, against the host layer.

The sources (Test_AddMul.cpp, Test_General.cpp, Test_Regression.cpp, Test_Statistics.cpp,
Regression.h, Statistics.h, Matrix.*) are copied from the mounted reference tree to a scratch
directory for the duration of the compile and never enter the repository; only the binaries
are kept (tests/cpp/_ref_build/, git-ignored; they travel to the GPU box like the built .so
files, where the reference tree does not exist).  Test infrastructure: nothing in the product
depends on these programs."""
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200", "host"))
from build_host import compile_client  # noqa: E402

REF = "/root/reference"
CLIENTS = ("Test_AddMul", "Test_General", "Test_Regression", "Test_Statistics")
SHARED = ("Regression.h", "Statistics.h", "Matrix.h", "Matrix.cpp")
GPU_DIR = os.path.join(ROOT, "tests", "cpp", "_ref_build")


def build_ref_clients(backend, out_dir, ref=REF):
    """-> {name: path of <name>_x}; empty when the reference tree is not mounted."""
    if not os.path.isdir(ref):
        return {}
    os.makedirs(out_dir, exist_ok=True)
    hdir = os.path.join(ROOT, "fhe-si_b200", "host")
    newest = max(os.path.getmtime(os.path.join(hdir, f)) for f in os.listdir(hdir) if f.endswith((".h", ".cpp")))
    exes = {c: os.path.join(out_dir, c + "_x") for c in CLIENTS}
    if all(os.path.exists(e) and os.path.getmtime(e) >= newest for e in exes.values()):
        return exes
    with tempfile.TemporaryDirectory() as tmp:
        for f in SHARED + tuple(c + ".cpp" for c in CLIENTS):
            shutil.copy(os.path.join(ref, f), os.path.join(tmp, f))
        for c in CLIENTS:
            exes[c] = compile_client([os.path.join(tmp, c + ".cpp")], backend, os.path.join(out_dir, c + "_x"))
    return exes


def prebuilt(out_dir=GPU_DIR):
    return {c: os.path.join(out_dir, c + "_x") for c in CLIENTS if os.path.exists(os.path.join(out_dir, c + "_x"))}


if __name__ == "__main__":
    backend = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fhe-si_b200", "libfhesi_b200.so")
    print(build_ref_clients(backend, GPU_DIR))
