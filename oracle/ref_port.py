"""ctypes binding of oracle/ref_restate.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The C file restates the reference's *algorithm* (per-prime Bluestein transforms, DoubleCRT
rows on the reference chain, incremental big-integer CRT).  It is used (a) to pin
fhesi_oracle.py with an independent implementation and (b) as the `cpu_baseline` /
`--impl reference` arm of bench.py, labelled kind="port" (the NTL build itself cannot be
compiled in this image)."""
import ctypes as C
import os
import subprocess

import numpy as np

import fhesi_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libref_restate.so")


def build(force=False):
    import fcntl
    src = os.path.join(HERE, "ref_restate.c")
    stale = lambda: not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB)
    if force or stale():
        with open(LIB + ".lock", "w") as lk:  # the all-core reference arm forks many workers
            fcntl.flock(lk, fcntl.LOCK_EX)
            if force or stale():
                subprocess.check_call(["make", "-C", HERE, "-s", "-B", "libref_restate.so"])
    return LIB


def _lib():
    lib = C.CDLL(build())
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p,
                               C.c_void_p, C.c_int]
    lib.ref_mult_relin.argtypes = [C.c_void_p] * 5
    lib.ref_rows.argtypes = [C.c_void_p] * 3
    lib.ref_to_poly.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
    lib.ref_transform_count.restype = C.c_uint64
    lib.ref_transform_count.argtypes = [C.c_void_p]
    lib.ref_set_faithful.argtypes = [C.c_void_p, C.c_int]
    return lib


class RefPort:
    """The reference algorithm on one oracle Context (its 60-bit chain and roots)."""

    def __init__(self, octx: O.Context, faithful: bool = False):
        self.lib = _lib()
        self.octx = octx
        self.L = len(octx.primes)
        pr = np.asarray(octx.primes, dtype=np.uint64)
        rt = np.asarray(octx.roots, dtype=np.uint64)
        self.h = self.lib.ref_create(octx.m, octx.logQ, octx.p, octx.decompSize, self.L, pr.ctypes.data,
                                     rt.ctypes.data, 1 if faithful else 0)
        self.n, self.W = octx.phim, O.words_per_coeff(octx.logQ)
        self.ksw = None

    def rows(self, poly) -> np.ndarray:
        """DoubleCRT(const ZZX&) -> rows [L][n]."""
        w = O.pack_poly_words(poly, self.octx.logQ)
        out = np.empty((self.L, self.n), dtype=np.uint64)
        self.lib.ref_rows(self.h, w.ctypes.data, out.ctypes.data)
        return out

    def to_poly(self, rows: np.ndarray, wout: int = 40):
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        out = np.empty((self.n, wout), dtype=np.uint32)
        self.lib.ref_to_poly(self.h, rows.ctypes.data, out.ctypes.data, wout)
        return O.unpack_poly_words(out)

    def set_key_switch(self, ks: O.KeySwitch):
        """keySwitchMatrix as DoubleCRT rows [2][3D][L][n] (FHE-SI.cpp:205-208)."""
        mats = []
        for row in (ks.b, ks.A):
            mats.append(np.stack([self.rows(O.reduce_poly(k, self.octx.logQ)) for k in row]))
        self.ksw = np.ascontiguousarray(np.stack(mats))

    def mult_relin(self, a_words: np.ndarray, b_words: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a_words, dtype=np.uint32)
        b = np.ascontiguousarray(b_words, dtype=np.uint32)
        out = np.empty((2, self.n, self.W), dtype=np.uint32)
        self.lib.ref_mult_relin(self.h, a.ctypes.data, b.ctypes.data, self.ksw.ctypes.data, out.ctypes.data)
        return out

    def transforms(self) -> int:
        return self.lib.ref_transform_count(self.h)
