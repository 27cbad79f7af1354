"""Key generation for Python callers of the C ABI, through the C++ host layer
(fhe-si_b200/host: FHESISecKey / FHESIPubKey / KeySwitchSI -- FHE-SI.cpp:42-62,86-91,153-239).
Returns `poly`-format uint32 arrays ready for Context.key_create / Context.ksw_create."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)


_LIBS = {}


def _host_lib(backend_path: str) -> C.CDLL:
    if backend_path not in _LIBS:
        sys.path.insert(0, os.path.join(_PKG, "host"))
        from build_host import build_host
        C.CDLL(backend_path, mode=C.RTLD_GLOBAL)  # the host library resolves fhesi_* from the backend
        _LIBS[backend_path] = C.CDLL(build_host(backend_path))
    return _LIBS[backend_path]


def prepare(lib_path: str | None = None) -> None:
    """Build (if stale) and load the host library now, so that a later keygen() call is key
    generation only -- callers that time their set-up phase call this during start-up."""
    from . import DEFAULT_LIB
    _host_lib(lib_path or DEFAULT_LIB)


def keygen(ctx, seed: int, g: int, rot_k=(), lib_path: str | None = None):
    """-> dict(sk [2][n][W], pk [2][n][W], ks_b / ks_A [3D][n][W], rot_b / rot_A [len(rot_k)][2D][n][W])."""
    from . import DEFAULT_LIB
    lib = _host_lib(lib_path or DEFAULT_LIB)
    i = ctx.info
    n, W, D = ctx.n, ctx.W, ctx.D
    rot = np.asarray(list(rot_k), dtype=np.uint32)
    out = {"sk": np.zeros((2, n, W), np.uint32), "pk": np.zeros((2, n, W), np.uint32),
           "ks_b": np.zeros((3 * D, n, W), np.uint32), "ks_A": np.zeros((3 * D, n, W), np.uint32),
           "rot_b": np.zeros((max(len(rot), 1), 2 * D, n, W), np.uint32),
           "rot_A": np.zeros((max(len(rot), 1), 2 * D, n, W), np.uint32)}
    fn = lib.fhesih_keygen
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32] + \
        [C.c_void_p] * 7
    rc = fn(i.m, i.logQ, i.p, g, i.decompSize, i.xi, seed, len(rot), rot.ctypes.data if len(rot) else None,
            out["sk"].ctypes.data, out["pk"].ctypes.data, out["ks_b"].ctypes.data, out["ks_A"].ctypes.data,
            out["rot_b"].ctypes.data, out["rot_A"].ctypes.data)
    if rc:
        raise RuntimeError(f"fhesih_keygen failed ({rc})")
    return out


def keydraws(ctx, seed: int, g: int, rot_k=(), lib_path: str | None = None):
    """The random draws of keygen() (same stream, same order) without the key-switch arithmetic, for
    Context.ksw_generate: dict(sk int32 [n], pk [2][n][W], s2_src [3][n], s2_A [3D][n][W], s2_e [3D][n],
    rot_src [R][2][n], rot_A [R][2D][n][W], rot_e [R][2D][n])."""
    from . import DEFAULT_LIB
    lib = _host_lib(lib_path or DEFAULT_LIB)
    i = ctx.info
    n, W, D = ctx.n, ctx.W, ctx.D
    rot = np.asarray(list(rot_k), dtype=np.uint32)
    R = max(len(rot), 1)
    out = {"sk": np.zeros(n, np.int32), "pk": np.zeros((2, n, W), np.uint32),
           "s2_src": np.zeros((3, n), np.int32), "s2_A": np.zeros((3 * D, n, W), np.uint32),
           "s2_e": np.zeros((3 * D, n), np.int32), "rot_src": np.zeros((R, 2, n), np.int32),
           "rot_A": np.zeros((R, 2 * D, n, W), np.uint32), "rot_e": np.zeros((R, 2 * D, n), np.int32)}
    fn = lib.fhesih_keydraws
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32] + \
        [C.c_void_p] * 9
    rc = fn(i.m, i.logQ, i.p, g, i.decompSize, i.xi, seed, len(rot), rot.ctypes.data if len(rot) else None,
            out["sk"].ctypes.data, out["pk"].ctypes.data, out["s2_src"].ctypes.data, out["s2_A"].ctypes.data,
            out["s2_e"].ctypes.data, out["rot_src"].ctypes.data, out["rot_A"].ctypes.data, out["rot_e"].ctypes.data)
    if rc:
        raise RuntimeError(f"fhesih_keydraws failed ({rc})")
    return out


def keydraws_flat(ctx, seed: int, g: int, rot_k=(), lib_path: str | None = None):
    """The draws of keygen() (same stream, same order) written straight into the arrays
    Context.keygen_batch consumes: dict(parts [M], sk int32 [n], src int32 [3 + 2R][n],
    A uint32 [(3 + 2R) D + 1][n][W], e int32 [(3 + 2R) D + 1][n]); the public key's c1 / e are the last entry."""
    from . import DEFAULT_LIB
    lib = _host_lib(lib_path or DEFAULT_LIB)
    i = ctx.info
    n, W, D = ctx.n, ctx.W, ctx.D
    rot = np.asarray(list(rot_k), dtype=np.uint32)
    R = len(rot)
    Kt = (3 + 2 * R) * D + 1
    out = {"parts": np.array([3] + [2] * R, np.uint32), "sk": np.empty(n, np.int32),
           "src": np.empty((3 + 2 * R, n), np.int32), "A": np.empty((Kt, n, W), np.uint32),
           "e": np.empty((Kt, n), np.int32)}
    fn = lib.fhesih_keydraws_flat
    fn.restype = C.c_int
    fn.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32] + \
        [C.c_void_p] * 5
    rc = fn(i.m, i.logQ, i.p, g, i.decompSize, i.xi, seed, R, rot.ctypes.data if R else None,
            out["sk"].ctypes.data, out["src"].ctypes.data, out["A"].ctypes.data, out["e"].ctypes.data)
    if rc:
        raise RuntimeError(f"fhesih_keydraws_flat failed ({rc})")
    return out


def sk_words(ctx, sk: np.ndarray) -> np.ndarray:
    """The secret key (1, s) in `poly` format for Context.key_create."""
    skw = np.zeros((2, ctx.n, ctx.W), np.uint32)
    skw[0, 0, 0] = 1                                             # sKeys[0] = 1
    skw[1] = (sk.astype(np.int64)[:, None] >> (32 * np.arange(ctx.W))[None, :]).astype(np.uint32)  # sign-extended
    return skw


def device_keys(ctx, seed: int, g: int, rot_k=(), lib_path: str | None = None):
    """Keys for a device context generated ON the device in one pass (fhesi_keygen_batch) from host
    draws: -> (ksw, [rotation ksw ...], pk handle, sk handle).  Same keys as keygen() + ksw_create /
    key_create for the same seed."""
    d = keydraws_flat(ctx, seed, g, rot_k, lib_path)
    ksws, pk = ctx.keygen_batch(d["parts"], d["src"], d["sk"], d["A"], d["e"], with_pk=True)
    return ksws[0], ksws[1:], pk, ctx.key_create(sk_words(ctx, d["sk"]))
