"""fhesi_ksw_generate (KeySwitchSI::Init on the device from explicit draws) on the
B200: b and A' against the oracle on the same draws, and a mult+relin through the generated matrix.
(Collected last on purpose: the newest entry point of the round.)"""
import pytest

import parity_checks as P
from common import CONFIGS, Scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg2"])
def test_ksw_generate(name, cuda_lib):
    P.check_ksw_generate(Scenario(*CONFIGS[name], seed=9, lib_path=cuda_lib), CONFIGS[name][2])


def test_device_keys_equal_host_keys(cuda_lib):
    """pyfhesi.hostkeys.keydraws + Context.ksw_generate reproduce keygen()'s matrices for the same seed."""
    import numpy as np
    import pyfhesi
    from pyfhesi.hostkeys import keydraws, keygen
    logq, p, g = CONFIGS["cfg4"]
    dev = pyfhesi.Context(p - 1, logq, p, 3, 391, 0, lib_path=cuda_lib)
    k = keygen(dev, 5, g, rot_k=[3, 9], lib_path=cuda_lib)
    d = keydraws(dev, 5, g, rot_k=[3, 9], lib_path=cuda_lib)
    assert np.array_equal(k["pk"], d["pk"])
    h, b, a = dev.ksw_generate(d["s2_src"], d["sk"], d["s2_A"], d["s2_e"], want_host=True)
    assert np.array_equal(b, k["ks_b"]) and np.array_equal(a, k["ks_A"])
    for r in range(2):
        h, b, a = dev.ksw_generate(d["rot_src"][r], d["sk"], d["rot_A"][r], d["rot_e"][r], want_host=True)
        assert np.array_equal(b, k["rot_b"][r]) and np.array_equal(a, k["rot_A"][r])


@pytest.mark.gpu
def test_keygen_batch_equals_host_keys_gpu(cuda_lib):
    P.check_keygen_batch(cuda_lib, CONFIGS["cfg4"], 391, [3, 9, 81])
    P.check_keygen_batch(cuda_lib, CONFIGS["cfg2"], 1, [3])
