"""Build oracle/_ref: the reference's OWN library sources, unmodified and compiled where they
lie under /root/reference, against the NTL stand-in in oracle/ntl_compat (NTL and GMP are not in
this image).  TEST INFRASTRUCTURE: the binaries pin the oracle (tests/golden/make_ref_golden.py)
and give an extra CPU data point (bench.py); nothing in the product links or runs them.

Outputs (git-ignored, they travel to the GPU box with the snapshot):
  oracle/_ref/golden_client_ref   tests/cpp/host_client.cpp compiled against the REFERENCE's headers
                                  and objects: writes the byte files the golden vectors are made of
  oracle/_ref/Test_AddMul_ref     the reference's own test program
  oracle/_ref/ref_bench           oracle/ref_bench.cpp: times c *= b; ApplyKeySwitch(c)
  oracle/_ref/Test_Regression_ref the reference's own Test_Regression.cpp
  oracle/_ref/ref_regression      oracle/ref_regression.cpp: Regression.h's phases with explicit (logQ, xi) and a
                                  block limit -- the CPU baseline of the regression wall-time metric
  oracle/_ref/api_probe_ref       tests/cpp/api_probe.cpp (every client-facing symbol of SURVEY.md §8b)

What is and is not the reference here: DoubleCRT, Cmodulus/Bluestein, Ciphertext, FHE-SI (keys,
Encrypt/Decrypt, key switching), Util, NumbTh samplers, PlaintextSpace, Serialization -- all the
reference's code.  NTL's big integers, zz_p/fftRep arithmetic and both random streams (NTL's and
libc rand(), which NumbTh.h maps lrand48 to) are ours (ntl_compat.h, libc_rand.cpp; one SplitMix64
stream), written against NTL's documented semantics.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
COMPAT = os.path.join(HERE, "ntl_compat")
# SRC of the reference's Makefile:13-17
SRC = ["PlaintextSpace", "CModulus", "FHEContext", "PAlgebra", "SingleCRT", "DoubleCRT", "NumbTh", "bluestein",
       "IndexSet", "Plaintext", "Util", "FHE-SI", "Ciphertext", "Serialization", "Matrix"]
CXX = ["g++", "-std=c++17", "-O2", "-w", "-I", COMPAT]
PROGRAMS = {
    "golden_client_ref": os.path.join(ROOT, "tests", "cpp", "host_client.cpp"),
    "ref_bench": os.path.join(HERE, "ref_bench.cpp"),
    "Test_AddMul_ref": os.path.join(REF, "Test_AddMul.cpp"),
    # the reference's own regression driver, unchanged, and our parameterised driver over its Regression.h
    "Test_Regression_ref": os.path.join(REF, "Test_Regression.cpp"),
    "ref_regression": os.path.join(HERE, "ref_regression.cpp"),
    # tests/cpp/api_probe.cpp against the reference itself: proves the probe only uses real reference API
    "api_probe_ref": os.path.join(ROOT, "tests", "cpp", "api_probe.cpp"),
}


def available():
    return {k: os.path.join(OUT, k) for k in PROGRAMS if os.path.exists(os.path.join(OUT, k))}


def build_ref(force=False):
    """-> {name: path}; {} when the reference tree is not mounted and nothing was prebuilt."""
    if not os.path.isdir(REF):
        return available()
    obj = os.path.join(OUT, "obj")
    os.makedirs(obj, exist_ok=True)
    deps = [os.path.join(COMPAT, "ntl_compat.h"), os.path.join(COMPAT, "libc_rand.cpp"),
            os.path.join(ROOT, "fhe-si_b200", "host", "ntl_shim.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    objs, jobs = [], []
    for name in SRC + ["libc_rand"]:
        src = os.path.join(COMPAT, "libc_rand.cpp") if name == "libc_rand" else os.path.join(REF, name + ".cpp")
        o = os.path.join(obj, name + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(newest, os.path.getmtime(src)):
            jobs.append(subprocess.Popen(CXX + ["-c", src, "-o", o]))
    for j in jobs:
        if j.wait():
            raise RuntimeError("oracle/_ref: compiling the reference sources failed")
    exes = {}
    for name, main in PROGRAMS.items():
        exe = os.path.join(OUT, name)
        exes[name] = exe
        if force or jobs or not os.path.exists(exe) or os.path.getmtime(exe) < max(newest, os.path.getmtime(main)):
            subprocess.check_call(CXX + ["-I", REF, main] + objs + ["-o", exe])
    return exes


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
