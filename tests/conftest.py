import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fhe-si_b200"), os.path.join(ROOT, "tests"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


# The golden vectors and the oracle draw from the documented deterministic TEST stream (SplitMix64); the host
# layer's production generator is ChaCha20 keyed from the OS or from the full seed (ntl_shim.h).  Every test
# process and every client program the tests start inherits this.
os.environ.setdefault("FHESI_TEST_RNG", "splitmix64")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def emu_lib():
    """Kernel-logic emulator (tests/emu): the product sources compiled for the host with CUDA
    threads as OS threads.  Test infrastructure only -- never loaded by the product."""
    from emu.build_emu import build_emu
    return build_emu()


@pytest.fixture(scope="session")
def cuda_lib():
    import build as fhesi_build  # fhe-si_b200/build.py
    return fhesi_build.build()
