"""GPU parity tests (run on the B200 with `pytest -m gpu`): every C-ABI entry point of
libfhesi_b200.so against the oracle, bit for bit, on all five BASELINE.json configs."""
import numpy as np
import pytest

import parity_checks as P
from common import CONFIGS, Scenario

pytestmark = pytest.mark.gpu

_cache = {}


def scenario(name, cuda_lib, xi=1):
    key = (name, xi)
    if key not in _cache:
        _cache[key] = Scenario(*CONFIGS[name], seed=20240611, xi=xi, lib_path=cuda_lib)
    return _cache[key]


def test_library_is_cuda(cuda_lib):
    import pyfhesi
    lib = pyfhesi.load_library(cuda_lib)
    assert b"sm_100a" in lib.fhesi_version()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5_128", "cfg5_512"])
def test_mult_relin(name, cuda_lib):
    sc = scenario(name, cuda_lib)
    P.check_mult_relin(sc, count=2 if name == "cfg5_512" else 3)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5_128", "cfg5_512"])
def test_mult_relin_wide(name, cuda_lib):
    """64 pairs x 3 key sets per parameter set (24 x 3 at logQ = 512, where the CPU checker is slowest)."""
    P.check_mult_relin_wide(*CONFIGS[name], cuda_lib, pairs=24 if name == "cfg5_512" else 64)


def test_fused_2048_p2027(cuda_lib):
    """The README's second family, p = 2027 (phi(m) = 1012, N = 2048): fused 256-thread kernels against the
    oracle and the C restatement, and against the generic kernels (FHESI_NO_FUSED) byte for byte."""
    import os
    sc = scenario("p2027", cuda_lib)
    assert sc.dev.N == 2048 and sc.dev.Ls
    P.check_mult_relin(sc, count=3)
    P.check_mult_relin(sc, count=2, host=True)
    P.check_pieces(sc, count=2)
    P.check_tensor_accumulate(sc, count=5)
    P.check_encrypt_decrypt(sc, count=2)
    P.check_edge_cases(sc, counts=(0, 1, 5))
    P.check_rotate_keyswitch(sc, CONFIGS["p2027"][2], count=2)
    P.check_mult_relin_wide(*CONFIGS["p2027_176"], cuda_lib, pairs=16, seeds=(7,))
    _, cts = sc.fresh(8)
    fused = sc.dev_mult_relin(cts[:4], cts[4:])
    os.environ["FHESI_NO_FUSED"] = "1"
    try:
        gen = Scenario(*CONFIGS["p2027"], seed=20240611, lib_path=cuda_lib)
        gen.ks = sc.ks  # same key-switch matrix, same ciphertexts, generic kernels
        assert np.array_equal(gen.dev_mult_relin(cts[:4], cts[4:]), fused)
    finally:
        del os.environ["FHESI_NO_FUSED"]


@pytest.mark.parametrize("name", ["m16", "m17", "m36", "m45", "m105", "m128", "m1320", "m771", "m1285"])
def test_general_m(name, cuda_lib):
    """Any m (bluestein.cpp:93-144 and CModulus.cpp:110-132 serve every m in the reference): the remainder by
    Phi_m and the automorphisms as sparse integer matrices.  m1320 and m771 (phi = 320, 512) run the fused
    N = 1024 kernels' general-m instances; the others the generic kernels."""
    from common import GENERAL_M
    logq, p, g, m = GENERAL_M[name]
    sc = Scenario(logq, p, g, seed=20240611, xi=3, lib_path=cuda_lib, m=m)
    assert sc.dev.n == sc.octx.phim and sc.dev.info.m == m
    big = sc.dev.N >= 1024
    sc.dev.profile_enable(True)
    P.check_mult_relin(sc, count=2 if big else 4)
    prof = sc.dev.profile_report()
    assert any(k.startswith("k_fused_keyswitch") for k in prof) == big, prof
    sc.dev.profile_enable(False)
    P.check_mult_relin(sc, count=2, random_inputs=True)
    P.check_mult_relin(sc, count=2, host=True)
    P.check_pieces(sc, count=2)
    P.check_tensor_accumulate(sc, count=3)
    P.check_encrypt_decrypt(sc, count=2)
    P.check_coeff_ops(sc, count=2)
    P.check_mul_plain(sc, count=1)
    P.check_rotate_keyswitch(sc, g, count=2)
    P.check_edge_cases(sc, counts=(0, 1, 5))
    if not big:
        P.check_ref_rows(sc)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5_128", "cfg5_512", "p2027_176"])
def test_crt_direct_paths(name, cuda_lib):
    """ScaleDown through the windowed explicit CRT, through its exact fallback on every thread, and through the
    mixed-radix kernel: three implementations, one set of bytes (the oracle's)."""
    P.check_crt_direct_paths(CONFIGS[name], cuda_lib, count=2 if name == "cfg5_512" else 3)


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_mult_relin_host_async(name, cuda_lib):
    """Batches in flight across calls (the next call's uploads overlap this call's last kernels and downloads)."""
    P.check_mult_relin_host_async(scenario(name, cuda_lib), count=5 if name == "cfg1" else 3)
    import pyfhesi  # a long batch too: many pipeline chunks per call, both staging halves, against the device path
    sc = scenario("cfg2", cuda_lib)
    if name == "cfg2":
        rng = np.random.default_rng(3)
        shape = (700, 2, sc.dev.n, sc.dev.W)
        ins = [(rng.integers(0, 2**32, size=shape, dtype=np.uint32), rng.integers(0, 2**32, size=shape, dtype=np.uint32))
               for _ in range(3)]
        outs = [np.zeros(shape, np.uint32) for _ in ins]
        for (a, b), o in zip(ins, outs):
            sc.dev.mult_relin_host_async(sc.ksw, a, b, o, shape[0])
        sc.dev.sync_all()
        for (a, b), o in zip(ins, outs):
            da, db, do = sc.dev.to_device(a), sc.dev.to_device(b), sc.dev.alloc(a.nbytes)
            sc.dev.mult_relin_dev(sc.ksw, da.ptr, db.ptr, do.ptr, shape[0])
            sc.dev.sync()
            assert np.array_equal(o, do.download(shape))


def test_mult_relin_host_and_random_cfg2(cuda_lib):
    sc = scenario("cfg2", cuda_lib)
    P.check_mult_relin(sc, count=2, host=True)
    P.check_mult_relin(sc, count=2, random_inputs=True)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4"])
def test_pieces(name, cuda_lib):
    sc = scenario(name, cuda_lib, xi=8)
    P.check_pieces(sc, count=2)
    P.check_tensor_accumulate(sc, count=5)
    P.check_tprod_scalar(sc)
    P.check_gathered_reduce(sc)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5_128", "cfg5_512"])
def test_encrypt_decrypt(name, cuda_lib):
    P.check_encrypt_decrypt(scenario(name, cuda_lib), count=3)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4"])
def test_coeff_ops(name, cuda_lib):
    P.check_coeff_ops(scenario(name, cuda_lib), count=3)
    P.check_mul_plain(scenario(name, cuda_lib), count=2)


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_ref_rows(name, cuda_lib):
    P.check_ref_rows(scenario(name, cuda_lib))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg5_128"])
def test_edge_cases(name, cuda_lib):
    P.check_edge_cases(scenario(name, cuda_lib), counts=(0, 1, 7, 13))


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4"])
def test_golden_vectors(name, cuda_lib):
    P.check_golden(name, cuda_lib)


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg4", "cfg5_512"])
def test_rotate_keyswitch(name, cuda_lib):
    P.check_rotate_keyswitch(scenario(name, cuda_lib), CONFIGS[name][2], count=3)


@pytest.mark.parametrize("name", ["cfg1", "cfg2"])
def test_embed_slots(name, cuda_lib):
    P.check_embed_slots(scenario(name, cuda_lib), CONFIGS[name][2], count=6)


def test_generic_and_fused_paths_agree(cuda_lib, monkeypatch):
    """The fused N=1024 kernels and the generic kernels are two CUDA implementations of the
    same arithmetic; they must agree bit for bit (and both with the oracle, above)."""
    sc = scenario("cfg2", cuda_lib)
    A, B = sc.random_cts(4), sc.random_cts(4)
    fused = sc.dev_mult_relin(A, B)
    monkeypatch.setenv("FHESI_NO_FUSED", "1")
    sc2 = Scenario(*CONFIGS["cfg2"], seed=20240611, lib_path=cuda_lib)
    sc2.ks = sc.ks
    generic = sc2.dev_mult_relin(A, B)
    assert np.array_equal(fused, generic)


def test_full_batch_properties_cfg2(cuda_lib):
    """BASELINE-size batch: size-independent properties instead of the (slow) oracle.
    (1) a batch equals the concatenation of its halves (no cross-talk, chunking exact);
    (2) mult_relin(a, b) decrypts to the plaintext product for every element
        (Test_AddMul.cpp:84-86 identity), checked on device;
    (3) commutativity: mult_relin(a,b) and mult_relin(b,a) decrypt identically."""
    import fhesi_oracle as O
    sc = scenario("cfg2", cuda_lib)
    d = sc.dev
    count, n = 512, d.n
    rng = np.random.default_rng(7)
    msgs = rng.integers(0, sc.p, size=(2 * count, n), dtype=np.uint32)
    rs = rng.integers(0, 2, size=(2 * count, n), dtype=np.uint8)
    es = np.rint(rng.normal(0, 3.2, size=(2 * count, 2, n))).astype(np.int32)
    dct = d.alloc(2 * count * d.ct_words(2) * 4)
    dmsg, drs, des = d.to_device(msgs), d.to_device(rs), d.to_device(es)  # keep alive until sync
    d.encrypt_dev(sc.dpk, dmsg.ptr, drs.ptr, des.ptr, dct.ptr, 2 * count)
    half = count * d.ct_words(2) * 4
    dout = d.alloc(half)
    dout2 = d.alloc(half)
    d.mult_relin_dev(sc.ksw, dct.ptr, dct.ptr + half, dout.ptr, count)
    d.mult_relin_dev(sc.ksw, dct.ptr + half, dct.ptr, dout2.ptr, count)
    dm, dm2 = d.alloc(count * n * 4), d.alloc(count * n * 4)
    d.decrypt_dev(sc.dsk, dout.ptr, 2, dm.ptr, count)
    d.decrypt_dev(sc.dsk, dout2.ptr, 2, dm2.ptr, count)
    d.sync()
    full = dout.download((count, 2, n, d.W))
    dec, dec2 = dm.download((count, n)), dm2.download((count, n))
    assert np.array_equal(dec, dec2)
    ring = sc.octx.ring
    for i in (0, 1, count // 2, count - 1):
        want = [c % sc.p for c in ring.mul(msgs[i].tolist(), msgs[count + i].tolist())]
        assert dec[i].tolist() == want
    # all elements: plaintext product through numpy (exact int64, mod p)
    k = 200  # halves
    d.mult_relin_dev(sc.ksw, dct.ptr, dct.ptr + half, dout2.ptr, k)
    d.sync()
    assert np.array_equal(dout2.download((k, 2, n, d.W)), full[:k])
    # one oracle spot check inside the big batch
    cts = dct.download((2 * count, 2, n, d.W))
    a = O.Ciphertext(sc.octx, sc.unpack_ct(cts[5]))
    b = O.Ciphertext(sc.octx, sc.unpack_ct(cts[count + 5]))
    from common import assert_ct_equal
    assert_ct_equal(sc, full[5], O.mult_relin(sc.ks, a, b), "big batch element 5")


def test_host_pipeline_full_size_cfg2(cuda_lib):
    """BASELINE.json's batch (8192 pairs per GPU) through fhesi_mult_relin_host -- two compute lanes,
    tapered chunk schedule -- must equal the device-resident path element for element; ragged counts
    exercise every shape of the schedule (no taper, taper with a short middle chunk, single chunk)."""
    sc = scenario("cfg2", cuda_lib)
    d = sc.dev
    count, n, W = 8192, d.n, d.W
    rng = np.random.default_rng(11)
    msgs = rng.integers(0, sc.p, size=(2 * count, n), dtype=np.uint32)
    rs = rng.integers(0, 2, size=(2 * count, n), dtype=np.uint8)
    es = np.rint(rng.normal(0, 3.2, size=(2 * count, 2, n))).astype(np.int32)
    dct = d.alloc(2 * count * d.ct_words(2) * 4)
    dmsg, drs, des = d.to_device(msgs), d.to_device(rs), d.to_device(es)
    d.encrypt_dev(sc.dpk, dmsg.ptr, drs.ptr, des.ptr, dct.ptr, 2 * count)
    half = count * d.ct_words(2) * 4
    dout = d.alloc(half)
    d.mult_relin_dev(sc.ksw, dct.ptr, dct.ptr + half, dout.ptr, count)
    dm = d.alloc(count * n * 4)
    d.decrypt_dev(sc.dsk, dout.ptr, 2, dm.ptr, count)
    d.sync()
    want = dout.download((count, 2, n, W))
    dec = dm.download((count, n))
    cts = dct.download((2 * count, 2, n, W))
    ha, hb = np.ascontiguousarray(cts[:count]), np.ascontiguousarray(cts[count:])
    for cnt in (count, 3457, 600, 1):
        got = d.mult_relin_host(sc.ksw, ha[:cnt], hb[:cnt])
        assert np.array_equal(got, want[:cnt]), f"host pipeline differs at count {cnt}"
    # the Test_AddMul.cpp:84-86 identity on a sample of the batch (plaintext product mod Phi_m, mod p)
    ring = sc.octx.ring
    for i in (0, 1, 4095, 4096, count - 1):
        assert dec[i].tolist() == [c % sc.p for c in ring.mul(msgs[i].tolist(), msgs[count + i].tolist())]
