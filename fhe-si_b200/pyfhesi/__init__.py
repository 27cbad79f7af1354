"""pyfhesi -- thin ctypes binding of include/fhesi.h (libfhesi_b200.so).

Plumbing only: the product is the CUDA library.  The binding never computes anything on
the CPU; if the CUDA library is missing or no GPU is usable it raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(os.path.dirname(_HERE), "libfhesi_b200.so")
FHESI_MAX_PRIMES = 40


class FhesiError(RuntimeError):
    pass


class Info(C.Structure):
    _fields_ = [("m", C.c_uint32), ("n", C.c_uint32), ("logQ", C.c_uint32), ("W", C.c_uint32),
                ("decompSize", C.c_uint32), ("D", C.c_uint32), ("N", C.c_uint32), ("Lt", C.c_uint32),
                ("Lk", C.c_uint32), ("Le", C.c_uint32), ("p", C.c_uint64), ("xi", C.c_uint64),
                ("primes", C.c_uint32 * FHESI_MAX_PRIMES), ("device", C.c_int), ("Ls", C.c_uint32),
                ("split_words", C.c_uint32)]


# every symbol include/fhesi.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_SZ = C.c_size_t
_U32 = C.c_uint32
SYMBOLS = {
    "fhesi_last_error": (C.c_char_p, []),
    "fhesi_version": (C.c_char_p, []),
    "fhesi_ctx_create": (C.c_int, [_U32, _U32, C.c_uint64, _U32, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "fhesi_ctx_destroy": (None, [_P]),
    "fhesi_ctx_info": (C.c_int, [_P, C.POINTER(Info)]),
    "fhesi_ctx_set_stream": (C.c_int, [_P, _P]),
    "fhesi_sync": (C.c_int, [_P]),
    "fhesi_malloc": (C.c_int, [_P, _SZ, C.POINTER(_P)]),
    "fhesi_free": (C.c_int, [_P, _P]),
    "fhesi_h2d": (C.c_int, [_P, _P, _P, _SZ]),
    "fhesi_h2d_async": (C.c_int, [_P, _P, _P, _SZ]),
    "fhesi_d2h": (C.c_int, [_P, _P, _P, _SZ]),
    "fhesi_crt_fallbacks": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "fhesi_host_alloc": (C.c_int, [_SZ, C.c_int, C.POINTER(_P)]),
    "fhesi_host_free": (C.c_int, [_P]),
    "fhesi_d2d": (C.c_int, [_P, _P, _P, _SZ]),
    "fhesi_ct_mul_plain_dev": (C.c_int, [_P, _P, _P, _U32, _SZ]),
    "fhesi_ct_bytes": (_SZ, [_P, _U32]),
    "fhesi_tprod_bytes": (_SZ, [_P, _U32]),
    "fhesi_ksw_create": (C.c_int, [_P, _P, _P, _U32, C.POINTER(_P)]),
    "fhesi_ksw_generate": (C.c_int, [_P, _P, _P, _P, _P, _U32, C.POINTER(_P), _P, _P]),
    "fhesi_keygen_batch": (C.c_int, [_P, _U32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fhesi_ksw_destroy": (None, [_P]),
    "fhesi_key_create": (C.c_int, [_P, _P, _U32, C.POINTER(_P)]),
    "fhesi_key_destroy": (None, [_P]),
    "fhesi_mult_relin_dev": (C.c_int, [_P, _P, _P, _P, _P, _SZ]),
    "fhesi_mult_relin_host": (C.c_int, [_P, _P, _P, _P, _P, _SZ]),
    "fhesi_mult_relin_host_async": (C.c_int, [_P, _P, _P, _P, _P, _SZ]),
    "fhesi_sync_all": (C.c_int, [_P]),
    "fhesi_ct_add_dev": (C.c_int, [_P, _P, _P, _U32, _SZ]),
    "fhesi_ct_sum_dev": (C.c_int, [_P, _P, _P, _U32, _SZ]),
    "fhesi_ct_mul_scalar_dev": (C.c_int, [_P, _P, C.c_int64, _U32, _SZ]),
    "fhesi_ct_tensor_dev": (C.c_int, [_P, _P, _U32, _P, _U32, _P, _SZ, C.c_int]),
    "fhesi_tprod_add_dev": (C.c_int, [_P, _P, _P, _U32, _SZ]),
    "fhesi_tprod_mul_scalar_dev": (C.c_int, [_P, _P, C.c_int64, _U32, _SZ]),
    "fhesi_scaledown_dev": (C.c_int, [_P, _P, _U32, _P, _SZ]),
    "fhesi_keyswitch_dev": (C.c_int, [_P, _P, _P, _P, _SZ]),
    "fhesi_encrypt_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, _SZ]),
    "fhesi_decrypt_dev": (C.c_int, [_P, _P, _P, _U32, _P, _SZ]),
    "fhesi_rotate_keyswitch_dev": (C.c_int, [_P, _P, _P, _U32, _P, _SZ]),
    "fhesi_tprod_add_poly_dev": (C.c_int, [_P, _P, _U32, _P, _U32, _SZ]),
    "fhesi_tprod_mul_poly_dev": (C.c_int, [_P, _P, _U32, _P, _U32, _SZ]),
    "fhesi_tprod_automorph_dev": (C.c_int, [_P, _P, _U32, _U32, _P, _SZ]),
    "fhesi_trim_cache": (C.c_int, [C.c_int]),
    "fhesi_embed_slots_dev": (C.c_int, [_P, _P, _U32, _P, _P, _SZ]),
    "fhesi_decode_slots_dev": (C.c_int, [_P, _P, _U32, _P, _P, _SZ]),
    "fhesi_ct_automorph_dev": (C.c_int, [_P, _P, _U32, _U32, _P, _SZ]),
    "fhesi_reduce_wide_dev": (C.c_int, [_P, _P, _U32, _P, _U32, _SZ]),
    "fhesi_ref_rows_host": (C.c_int, [_P, _P, _U32, _P, _P, _U32, _P]),
    "fhesi_tprod_reduce_gathered_dev": (C.c_int, [_P, _P, _U32, _U32, _P]),
    "fhesi_modmul_peak": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "fhesi_pipe_peak": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "fhesi_profile_enable": (C.c_int, [_P, C.c_int]),
    "fhesi_profile_launches": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "fhesi_profile_report": (C.c_int, [_P, C.c_char_p, _SZ]),
}


def load_library(path: Optional[str] = None) -> C.CDLL:
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise FhesiError(f"{path} not found: build it with `python fhe-si_b200/build.py` "
                         "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def _ptr(a) -> int:
    """numpy array / torch tensor / int -> raw address."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


class DeviceBuffer:
    """A cudaMalloc'd buffer owned through the C ABI (fhesi_malloc / fhesi_free)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx, self.nbytes = ctx, nbytes
        p = _P()
        ctx._ck(ctx.lib.fhesi_malloc(ctx.h, nbytes, C.byref(p)))
        self.ptr = p.value

    def upload(self, arr: np.ndarray) -> "DeviceBuffer":
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._ck(self.ctx.lib.fhesi_h2d(self.ctx.h, self.ptr, arr.ctypes.data, arr.nbytes))
        return self

    def download(self, shape, dtype=np.uint32) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        self.ctx._ck(self.ctx.lib.fhesi_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.fhesi_free(self.ctx.h, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """FHEcontext image on one GPU: fhesi_ctx_create (FHEContext.h:105-118 +
    FHEContext.cpp:83-85)."""

    def __init__(self, m: int, logQ: int, p: int, decompSize: int = 3, xi: int = 1, device: int = 0,
                 lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        h = _P()
        rc = self.lib.fhesi_ctx_create(m, logQ, p, decompSize, xi, device, C.byref(h))
        if rc:
            raise FhesiError(f"fhesi_ctx_create: {self.lib.fhesi_last_error().decode()} (rc={rc})")
        self.h = h.value
        self.info = Info()
        self._ck(self.lib.fhesi_ctx_info(self.h, C.byref(self.info)))
        i = self.info
        self.n, self.N, self.W, self.D = i.n, i.N, i.W, i.D
        self.Lt, self.Lk, self.Le, self.Ls = i.Lt, i.Lk, i.Le, i.Ls
        self.primes = [i.primes[k] for k in range(i.Lt)]

    def _ck(self, rc: int):
        if rc:
            raise FhesiError(f"{self.lib.fhesi_last_error().decode()} (rc={rc})")

    def close(self):
        if getattr(self, "h", None):
            self.lib.fhesi_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def set_stream(self, stream_ptr: int):
        self._ck(self.lib.fhesi_ctx_set_stream(self.h, stream_ptr))

    def sync(self):
        self._ck(self.lib.fhesi_sync(self.h))

    def alloc(self, nbytes: int) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, arr: np.ndarray) -> DeviceBuffer:
        arr = np.ascontiguousarray(arr)
        return DeviceBuffer(self, max(arr.nbytes, 16)).upload(arr)

    def crt_fallbacks(self) -> int:
        """Coefficients whose ScaleDown took k_crt_direct's exact fallback since the context was created."""
        v = C.c_uint64()
        self._ck(self.lib.fhesi_crt_fallbacks(self.h, C.byref(v)))
        return v.value

    def host_alloc(self, shape, dtype=np.uint32, write_combined: bool = False) -> np.ndarray:
        """Page-locked host array for the *_host entry points (fhesi_host_alloc); write_combined=True for operands the
        CPU only writes.  The memory lives until host_free(array)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = _P()
        self._ck(self.lib.fhesi_host_alloc(nbytes, 1 if write_combined else 0, C.byref(p)))
        buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        self._host_blocks = getattr(self, "_host_blocks", {})
        self._host_blocks[arr.ctypes.data] = p.value
        return arr

    def host_free(self, arr: np.ndarray):
        p = getattr(self, "_host_blocks", {}).pop(arr.ctypes.data, None)
        if p is not None:
            self._ck(self.lib.fhesi_host_free(p))

    def ct_words(self, parts: int) -> int:
        return parts * self.n * self.W

    def tprod_words(self, parts: int) -> int:
        return parts * self.Lt * self.N

    # ---- keys
    def ksw_create(self, b: np.ndarray, A: np.ndarray, src_parts: int) -> int:
        b = np.ascontiguousarray(b, dtype=np.uint32)
        A = np.ascontiguousarray(A, dtype=np.uint32)
        assert b.shape == A.shape == (src_parts * self.D, self.n, self.W)
        k = _P()
        self._ck(self.lib.fhesi_ksw_create(self.h, b.ctypes.data, A.ctypes.data, src_parts, C.byref(k)))
        return k.value

    def ksw_generate(self, src: np.ndarray, t: np.ndarray, A: np.ndarray, e: np.ndarray, want_host=False):
        """KeySwitchSI::Init on the device from explicit draws -> handle (and b, A' words if want_host)."""
        src = np.ascontiguousarray(src, dtype=np.int32)
        t = np.ascontiguousarray(t, dtype=np.int32)
        A = np.ascontiguousarray(A, dtype=np.uint32)
        e = np.ascontiguousarray(e, dtype=np.int32)
        parts = src.shape[0]
        assert src.shape == (parts, self.n) and t.shape == (self.n,)
        assert A.shape == (parts * self.D, self.n, self.W) and e.shape == (parts * self.D, self.n)
        k = _P()
        b_out = np.empty_like(A) if want_host else None
        a_out = np.empty_like(A) if want_host else None
        self._ck(self.lib.fhesi_ksw_generate(self.h, src.ctypes.data, t.ctypes.data, A.ctypes.data, e.ctypes.data,
                                             parts, C.byref(k), b_out.ctypes.data if want_host else None,
                                             a_out.ctypes.data if want_host else None))
        return (k.value, b_out, a_out) if want_host else k.value

    def keygen_batch(self, parts, src, t, A, e, with_pk=True, want_host=False):
        """All key-switch matrices of a set-up and (with_pk) the public key in one pass of kernels
        (fhesi_keygen_batch).  parts: list of source-part counts, one per matrix; src int32 [sum parts][n];
        t int32 [n]; A uint32 [sum parts*D (+1)][n][W]; e int32 [sum parts*D (+1)][n] -- the public key's
        c1 / e last.  -> (list of ksw handles, pk handle or None[, b, A', pk words])."""
        parts = np.ascontiguousarray(parts, dtype=np.uint32)
        src = np.ascontiguousarray(src, dtype=np.int32).reshape(-1, self.n)
        t = np.ascontiguousarray(t, dtype=np.int32)
        A = np.ascontiguousarray(A, dtype=np.uint32)
        e = np.ascontiguousarray(e, dtype=np.int32)
        M, rows = len(parts), int(parts.sum())
        Kt = rows * self.D + (1 if with_pk else 0)
        assert src.shape[0] == rows and t.shape == (self.n,)
        assert A.shape == (Kt, self.n, self.W) and e.shape == (Kt, self.n), (A.shape, e.shape, Kt)
        outs = (_P * max(M, 1))()
        pk = _P()
        b_out = np.empty((rows * self.D, self.n, self.W), np.uint32) if want_host else None
        a_out = np.empty_like(b_out) if want_host else None
        pk_out = np.empty((2, self.n, self.W), np.uint32) if (want_host and with_pk) else None
        self._ck(self.lib.fhesi_keygen_batch(
            self.h, M, parts.ctypes.data if M else None, src.ctypes.data if rows else None, t.ctypes.data,
            A.ctypes.data, e.ctypes.data, C.cast(outs, _P) if M else None, _ptr(b_out), _ptr(a_out),
            C.cast(C.pointer(pk), _P) if with_pk else None, _ptr(pk_out)))
        ksws = [outs[i] for i in range(M)]
        if want_host:
            return ksws, (pk.value if with_pk else None), b_out, a_out, pk_out
        return ksws, (pk.value if with_pk else None)

    def key_create(self, polys: np.ndarray) -> int:
        polys = np.ascontiguousarray(polys, dtype=np.uint32)
        assert polys.shape[1:] == (self.n, self.W)
        k = _P()
        self._ck(self.lib.fhesi_key_create(self.h, polys.ctypes.data, polys.shape[0], C.byref(k)))
        return k.value

    # ---- ops on device pointers (numpy arrays are NOT accepted here: use *_host or to_device)
    def mult_relin_dev(self, ksw, a, b, out, count):
        self._ck(self.lib.fhesi_mult_relin_dev(self.h, ksw, _ptr(a), _ptr(b), _ptr(out), count))

    def mult_relin_host(self, ksw, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        count = a.shape[0]
        assert a.shape == b.shape == (count, 2, self.n, self.W)
        out = np.empty_like(a)
        self._ck(self.lib.fhesi_mult_relin_host(self.h, ksw, a.ctypes.data, b.ctypes.data,
                                                out.ctypes.data, count))
        return out

    def mult_relin_host_async(self, ksw, a, b, out, count):
        """fhesi_mult_relin_host_async: host pointers / page-locked arrays; returns when enqueued (sync_all waits)."""
        self._ck(self.lib.fhesi_mult_relin_host_async(self.h, ksw, _ptr(a), _ptr(b), _ptr(out), count))

    def sync_all(self):
        self._ck(self.lib.fhesi_sync_all(self.h))

    def ct_add_dev(self, io, other, parts, count):
        self._ck(self.lib.fhesi_ct_add_dev(self.h, _ptr(io), _ptr(other), parts, count))

    def ct_sum_dev(self, inp, out, parts, count):
        self._ck(self.lib.fhesi_ct_sum_dev(self.h, _ptr(inp), _ptr(out), parts, count))

    def ct_mul_scalar_dev(self, io, l, parts, count):
        self._ck(self.lib.fhesi_ct_mul_scalar_dev(self.h, _ptr(io), l, parts, count))

    def ct_mul_plain_dev(self, io, plain, parts, count):
        self._ck(self.lib.fhesi_ct_mul_plain_dev(self.h, _ptr(io), _ptr(plain), parts, count))

    def ct_tensor_dev(self, a, pa, b, pb, tprod, count, accumulate=False):
        self._ck(self.lib.fhesi_ct_tensor_dev(self.h, _ptr(a), pa, _ptr(b), pb, _ptr(tprod), count,
                                              1 if accumulate else 0))

    def tprod_add_dev(self, io, other, parts, count):
        self._ck(self.lib.fhesi_tprod_add_dev(self.h, _ptr(io), _ptr(other), parts, count))

    def tprod_mul_scalar_dev(self, io, l, parts, count):
        self._ck(self.lib.fhesi_tprod_mul_scalar_dev(self.h, _ptr(io), l, parts, count))

    def scaledown_dev(self, tprod, parts, out, count):
        self._ck(self.lib.fhesi_scaledown_dev(self.h, _ptr(tprod), parts, _ptr(out), count))

    def keyswitch_dev(self, ksw, inp, out, count):
        self._ck(self.lib.fhesi_keyswitch_dev(self.h, ksw, _ptr(inp), _ptr(out), count))

    def encrypt_dev(self, pk, msg, r, e, out, count):
        self._ck(self.lib.fhesi_encrypt_dev(self.h, pk, _ptr(msg), _ptr(r), _ptr(e), _ptr(out), count))

    def decrypt_dev(self, sk, inp, parts, msg, count):
        self._ck(self.lib.fhesi_decrypt_dev(self.h, sk, _ptr(inp), parts, _ptr(msg), count))

    def rotate_keyswitch_dev(self, ksw, inp, k, out, count):
        self._ck(self.lib.fhesi_rotate_keyswitch_dev(self.h, ksw, _ptr(inp), k, _ptr(out), count))

    def tprod_add_poly_dev(self, tprod, parts, poly, win, count):
        self._ck(self.lib.fhesi_tprod_add_poly_dev(self.h, _ptr(tprod), parts, _ptr(poly), win, count))

    def tprod_mul_poly_dev(self, tprod, parts, poly, win, count):
        self._ck(self.lib.fhesi_tprod_mul_poly_dev(self.h, _ptr(tprod), parts, _ptr(poly), win, count))

    def tprod_automorph_dev(self, inp, parts, k, out, count):
        self._ck(self.lib.fhesi_tprod_automorph_dev(self.h, _ptr(inp), parts, k, _ptr(out), count))

    def embed_slots_dev(self, basis, nslots, vals, msg, count):
        self._ck(self.lib.fhesi_embed_slots_dev(self.h, _ptr(basis), nslots, _ptr(vals), _ptr(msg), count))

    def decode_slots_dev(self, vander, nslots: int, msg, vals, count: int):
        """PlaintextSpace::DecodeSlots for a batch: vals[c][k] = msg[c](root_k) mod p (vander[j][k] = root_k^j)."""
        self._ck(self.lib.fhesi_decode_slots_dev(self.h, _ptr(vander), nslots, _ptr(msg), _ptr(vals), count))

    def ct_automorph_dev(self, inp, parts, k, out_wide, count):
        self._ck(self.lib.fhesi_ct_automorph_dev(self.h, _ptr(inp), parts, k, _ptr(out_wide), count))

    def reduce_wide_dev(self, inp, Win, out, parts, count):
        self._ck(self.lib.fhesi_reduce_wide_dev(self.h, _ptr(inp), Win, _ptr(out), parts, count))

    def tprod_reduce_gathered_dev(self, gathered, world, parts, out):
        self._ck(self.lib.fhesi_tprod_reduce_gathered_dev(self.h, _ptr(gathered), world, parts, _ptr(out)))

    def ref_rows_host(self, poly: np.ndarray, primes, roots) -> np.ndarray:
        poly = np.ascontiguousarray(poly, dtype=np.uint32)
        L = len(primes)
        pr = np.asarray(primes, dtype=np.uint64)
        rt = np.asarray(roots, dtype=np.uint64)
        rows = np.empty((L, self.n), dtype=np.int64)
        self._ck(self.lib.fhesi_ref_rows_host(self.h, poly.ctypes.data, poly.shape[1], pr.ctypes.data,
                                              rt.ctypes.data, L, rows.ctypes.data))
        return rows

    def profile_enable(self, on: bool = True):
        self._ck(self.lib.fhesi_profile_enable(self.h, 1 if on else 0))

    def launches(self) -> int:
        v = C.c_uint64()
        self._ck(self.lib.fhesi_profile_launches(self.h, C.byref(v)))
        return v.value

    def profile_report(self) -> dict:
        """kernel name -> (launches, total_ms) since profile_enable(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.fhesi_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.rsplit(" ", 2)
            out[name] = (int(cnt), float(ms))
        return out

    def pipe_peak(self, kind: int) -> float:
        v = C.c_double()
        self._ck(self.lib.fhesi_pipe_peak(self.h, kind, C.byref(v)))
        return v.value

    def modmul_peak(self, word_bits: int) -> float:
        v = C.c_double()
        self._ck(self.lib.fhesi_modmul_peak(self.h, word_bits, C.byref(v)))
        return v.value
