"""CPU CI: the kernel sources compiled against tests/emu/cuda_emu.h (CUDA threads as OS
threads) must match the oracle bit for bit.  This checks kernel LOGIC only; the GPU parity
tests proper are in test_gpu_parity.py and run on the B200."""
import pytest

import parity_checks as P
from common import CONFIGS, Scenario


@pytest.fixture(scope="module")
def sc1(emu_lib):
    return Scenario(*CONFIGS["cfg1"], seed=11, xi=4, lib_path=emu_lib)


def test_emu_mult_relin_cfg1(sc1):
    P.check_mult_relin(sc1, count=3)
    P.check_mult_relin(sc1, count=2, host=True)
    P.check_mult_relin(sc1, count=2, random_inputs=True)


def test_emu_mult_relin_host_async_cfg1(sc1):
    P.check_mult_relin_host_async(sc1, count=3)


def test_emu_mult_relin_wide_cfg1(emu_lib):
    P.check_mult_relin_wide(*CONFIGS["cfg1"], emu_lib, pairs=16, seeds=(101, 202))


def test_emu_pieces_cfg1(sc1):
    P.check_pieces(sc1, count=2)
    P.check_tensor_accumulate(sc1, count=3)
    P.check_tprod_scalar(sc1)
    P.check_gathered_reduce(sc1)


def test_emu_encrypt_decrypt_cfg1(sc1):
    P.check_encrypt_decrypt(sc1, count=3)


def test_emu_coeff_ops_cfg1(sc1):
    P.check_coeff_ops(sc1, count=3)
    P.check_mul_plain(sc1, count=2)


def test_emu_ref_rows_cfg1(sc1):
    P.check_ref_rows(sc1)


def test_emu_ksw_generate_cfg1(sc1):
    P.check_ksw_generate(sc1, 7)


def test_emu_keygen_batch_cfg1(emu_lib):
    P.check_keygen_batch(emu_lib, CONFIGS["cfg1"], 4, [7, 5])


def test_emu_embed_slots_cfg1(sc1):
    P.check_embed_slots(sc1, 7)


def test_emu_edge_cases_cfg1(sc1):
    P.check_edge_cases(sc1)


def test_emu_golden_cfg1(emu_lib):
    P.check_golden("cfg1", emu_lib)


def test_emu_other_small_rings(emu_lib):
    # m = 2*17 (n = 16: N = 2n exactly, the fold guard) and m = 2*13, odd logQ
    for m_half, logq in ((17, 61), (13, 100)):
        p = 2 * m_half + 1
        sc = Scenario(logq, p, 3, seed=3, lib_path=emu_lib)
        P.check_mult_relin(sc, count=2)
        P.check_encrypt_decrypt(sc, count=1)
        P.check_coeff_ops(sc, count=2)


def test_emu_mult_relin_cfg2(emu_lib):
    sc = Scenario(*CONFIGS["cfg2"], seed=2, lib_path=emu_lib)
    P.check_mult_relin(sc, count=1)
    P.check_rotate_keyswitch(sc, CONFIGS["cfg2"][2], count=1, compare_steps=False)  # fused N=1024 digit kernel


def test_emu_fused_2048_p2027(emu_lib):
    """phi(m) = 1012 (p = 2027): the fused N = 2048 kernels (256 threads per transform, kernels_fused2k.cuh)."""
    sc = Scenario(*CONFIGS["p2027_176"], seed=5, lib_path=emu_lib)
    assert sc.dev.N == 2048 and sc.dev.Ls and sc.dev.Ls < sc.dev.Lk
    P.check_mult_relin(sc, count=1)
    P.check_pieces(sc, count=1)


@pytest.mark.parametrize("name", ["m16", "m17", "m36", "m45", "m105", "m128"])
def test_emu_general_m(name, emu_lib):
    """Any m, not only 2 * (odd prime): the remainder by Phi_m and the automorphisms as sparse integer matrices
    (DevCtx::red, k_automorph_csr) against the oracle's NTL-style rem (bluestein.cpp:93-144 / CModulus.cpp:128-129
    serve every m in the reference)."""
    from common import GENERAL_M
    logq, p, g, m = GENERAL_M[name]
    sc = Scenario(logq, p, g, seed=17, xi=3, lib_path=emu_lib, m=m)
    assert sc.dev.n == sc.octx.phim and sc.dev.info.m == m
    P.check_mult_relin(sc, count=2)
    P.check_mult_relin(sc, count=1, random_inputs=True)
    P.check_pieces(sc, count=1)
    P.check_tensor_accumulate(sc, count=3)
    P.check_encrypt_decrypt(sc, count=2)
    P.check_coeff_ops(sc, count=2)
    P.check_mul_plain(sc, count=1)
    P.check_rotate_keyswitch(sc, g, count=1)
    P.check_edge_cases(sc, counts=(0, 1, 3))
    P.check_ref_rows(sc)


def test_emu_general_m_fused(emu_lib):
    """256 < phi(m) <= 512 takes the fused N = 1024 kernels' general-m instances (GEN = true)."""
    from common import GENERAL_M
    logq, p, g, m = GENERAL_M["m1320"]
    sc = Scenario(logq, p, g, seed=23, lib_path=emu_lib, m=m)
    assert sc.dev.N == 1024 and sc.dev.n == 320
    sc.dev.profile_enable(True)
    P.check_mult_relin(sc, count=1)
    assert "k_fused_keyswitch_split" in sc.dev.profile_report() or "k_fused_keyswitch<true>" in sc.dev.profile_report()
    sc.dev.profile_enable(False)
    P.check_rotate_keyswitch(sc, g, count=1)


def test_emu_general_m_fused_2048(emu_lib):
    """phi(m) = 1024 (m = 5 * 257): the N = 2048 fused kernels' general-m instances."""
    from common import GENERAL_M
    logq, p, g, m = GENERAL_M["m1285"]
    sc = Scenario(logq, p, g, seed=29, lib_path=emu_lib, m=m)
    assert sc.dev.N == 2048 and sc.dev.n == 1024
    sc.dev.profile_enable(True)
    P.check_mult_relin(sc, count=1)
    assert "k_fused_keyswitch_split_2k" in sc.dev.profile_report()


@pytest.mark.parametrize("name", ["cfg1", "m36"])
def test_emu_crt_direct_paths(name, emu_lib):
    """cfg1 (logQ = 80: the limb window starts at limb 0, nothing is cut) and m = 36 at logQ = 100 (the window starts
    at limb 1: guard limb and cut-off carry in play).  The BASELINE sizes run on the B200 (test_crt_direct_paths)."""
    from common import GENERAL_M
    if name in CONFIGS:
        P.check_crt_direct_paths(CONFIGS[name], emu_lib, count=1)
    else:
        logq, p, g, m = GENERAL_M[name]
        P.check_crt_direct_paths((logq, p, g), emu_lib, count=1, m=m)
