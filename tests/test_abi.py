"""The C-ABI library loads and exports every symbol include/fhesi.h declares (no compute
calls without a GPU), and fails loudly -- no CPU fallback -- when no device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "fhesi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhesi_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    import pyfhesi
    assert declared_symbols() == sorted(pyfhesi.SYMBOLS)


def test_library_exports_every_declared_symbol(cuda_lib):
    lib = ctypes.CDLL(cuda_lib)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_built_for_sm100a(cuda_lib):
    import pyfhesi
    lib = pyfhesi.load_library(cuda_lib)
    assert b"sm_100a" in lib.fhesi_version()


def test_no_cpu_fallback_without_gpu(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pyfhesi
    with pytest.raises(pyfhesi.FhesiError):
        pyfhesi.Context(22, 80, 23, lib_path=cuda_lib)


def test_bad_parameters_rejected(emu_lib):
    import pyfhesi
    # m out of range, phi(m) < 2, phi(m) > 1024, logQ too small, p < 2 (any other m is served: test_emu_general_m)
    for args in ((2, 80, 23), (9000, 80, 23), (6151, 80, 23), (22, 4, 23), (22, 80, 1)):
        with pytest.raises(pyfhesi.FhesiError):
            pyfhesi.Context(*args, lib_path=emu_lib)


def test_host_alloc_roundtrip(emu_lib):
    """fhesi_host_alloc / fhesi_host_free (page-locked staging; write-combined for operands): the array is writable,
    readable and released; the fallback counter of the windowed CRT starts at zero."""
    import numpy as np
    import pyfhesi
    ctx = pyfhesi.Context(22, 80, 23, lib_path=emu_lib)
    for wc in (False, True):
        a = ctx.host_alloc((3, ctx.n, ctx.W), np.uint32, write_combined=wc)
        a[...] = np.arange(a.size, dtype=np.uint32).reshape(a.shape)
        assert int(a.sum()) == a.size * (a.size - 1) // 2
        ctx.host_free(a)
    assert ctx.crt_fallbacks() == 0
    ctx.close()
