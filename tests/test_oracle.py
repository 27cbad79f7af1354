"""The oracle against (i) the reference's own self-consistency identities
(Test_AddMul.cpp:84-86), (ii) chain independence, (iii) the C restatement of the reference's
algorithm (oracle/ref_restate.c), (iv) serialization round trips, (v) the committed golden
vectors.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import fhesi_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def addmul_identities(logq, p, g, seed):
    """runTest of Test_AddMul.cpp:11-113."""
    ctx = O.Context(p - 1, logq, p, g).setup_si()
    rng = O.Rng(seed)
    sk = O.SecKey.generate(ctx, rng)
    pk = O.PubKey.generate(sk, rng)
    n, R = ctx.phim, ctx.ring
    m1 = [rng.random_bnd(p) for _ in range(n)]
    m2 = [rng.random_bnd(p) for _ in range(n)]
    modp = lambda a: [c % p for c in a]
    prod = modp(R.mul(m1, m2))
    prod2 = modp(R.mul(prod, prod))
    c1, c2 = O.encrypt_rng(pk, m1, rng), O.encrypt_rng(pk, m2, rng)
    csum = c1.copy().add(c2)
    csm = c2.copy()
    for _ in range(6):
        csm.add(c2)
    cprod = c1.copy().mul(c2)
    assert O.decrypt(sk, csum) == modp([a + b for a, b in zip(m1, m2)])
    assert O.decrypt(sk, csm) == modp([7 * b for b in m2])
    ks = O.KeySwitch.init_s2(sk, rng)
    O.apply_key_switch(ks, cprod)
    assert O.decrypt(sk, cprod) == prod
    cprod.mul(cprod)
    tmp, cq = cprod.copy(), cprod.copy()
    O.apply_key_switch(ks, cprod)
    assert O.decrypt(sk, cprod) == prod2
    for _ in range(8):
        cq.add(tmp)
    O.apply_key_switch(ks, cq)
    cq.mul(cprod)
    O.apply_key_switch(ks, cq)
    assert O.decrypt(sk, cq) == modp([9 * c for c in R.mul(prod2, prod2)])


@pytest.mark.parametrize("seed", range(12))
def test_addmul_identities_cfg1(seed):
    addmul_identities(80, 23, 7, seed)


def test_addmul_identities_cfg2():
    addmul_identities(256, 1019, 3, 1)


def test_addmul_identities_p2027():
    addmul_identities(128, 2027, 3, 2)  # README's other example parameter set (m = 2*1013)


def test_chain_matches_survey():
    # first primes of the 60-bit and 50-bit chains, and L60/L50 per config
    c = O.Context(1018, 256, 1019, 3).setup_si()
    assert c.primes[0] == 1152921504606820681 and len(c.primes) == 10 and c.primes[-1] == 4073
    c50 = O.Context(1018, 256, 1019, 3).setup_si(start_bits=50)
    assert c50.primes[0] == 1125899906824669 and len(c50.primes) == 11
    assert len(O.Context(22, 80, 23, 7).setup_si().primes) == 3
    assert len(O.Context(1018, 176, 1019, 3).setup_si(xi=391).primes) == 7
    assert len(O.Context(1018, 100, 1019, 3).setup_si(xi=40).primes) == 4
    for q in c.primes + c50.primes:
        assert O.is_prime(q) and q % (2 * 1018) == 1


def test_g2_is_not_a_unit_mod_1018():
    # BASELINE's "g=2" cannot be used with m=1018; automorph must refuse it
    ring = O.Ring(1018)
    with pytest.raises(ValueError):
        ring.automorph([1] * ring.phim, 2)
    ring.automorph([1] * ring.phim, 3)


def test_rows_roundtrip_and_chain_independence_cfg1():
    """DoubleCRT rows by direct evaluation and the literal toPoly (inverse transform +
    incremental CRT) invert each other under both chains -- coefficient results do not
    depend on the chain."""
    rng = O.Rng(9)
    for sb in (60, 50):
        ctx = O.Context(22, 80, 23, 7).setup_si(start_bits=sb)
        a = O.sample_random(rng, ctx.q << 40, ctx.phim)  # wider than q, still < P/2
        assert O.dcrt_to_poly(ctx, O.dcrt_rows(ctx, a)) == a


@pytest.mark.parametrize("cfg,start_bits", [((80, 23, 7), 60), ((80, 23, 7), 50), ((256, 1019, 3), 60),
                                            ((256, 1019, 3), 50), ((100, 1019, 3), 60)])
def test_reference_algorithm_port_agrees(cfg, start_bits):
    """oracle/ref_restate.c (Bluestein per prime on the reference chain + incremental CRT)
    gives the same mult+relin ciphertext as the exact-integer oracle, under both chains."""
    import ref_port
    logq, p, g = cfg
    ctx = O.Context(p - 1, logq, p, g).setup_si(start_bits=start_bits)
    rng = O.Rng(77)
    sk = O.SecKey.generate(ctx, rng)
    pk = O.PubKey.generate(sk, rng)
    ks = O.KeySwitch.init_s2(sk, rng)
    port = ref_port.RefPort(ctx)
    port.set_key_switch(ks)
    c1 = O.encrypt_rng(pk, [rng.random_bnd(p) for _ in range(ctx.phim)], rng)
    c2 = O.encrypt_rng(pk, [rng.random_bnd(p) for _ in range(ctx.phim)], rng)
    pack = lambda ct: np.stack([O.pack_poly_words(x, logq) for x in ct.parts])
    t0 = port.transforms()
    out = port.mult_relin(pack(c1), pack(c2))
    L, D = len(ctx.primes), ctx.ndigits
    assert port.transforms() - t0 == (4 + 3 * D) * L + 5 * L  # SURVEY.md §3.2/§3.3 transform counts
    want = O.mult_relin(ks, c1, c2)
    assert [O.unpack_poly_words(out[i]) for i in range(2)] == want.parts
    if p == 23:
        a = O.sample_random(rng, ctx.q, ctx.phim)
        assert port.rows(a).tolist() == O.dcrt_rows(ctx, a)


def test_reduce_matches_reference_semantics():
    # Util.cpp:3-26 on a few hand-checked values
    assert O.reduce_q(5, 4) == 5 and O.reduce_q(8, 4) == -8 and O.reduce_q(-9, 4) == 7
    assert O.reduce_q(-9, 4, True) == 7 and O.reduce_q(24, 4, True) == 8 and O.reduce_q(-16, 4) == 0
    for v in range(-70, 70):
        r = O.reduce_q(v, 5)
        assert -16 <= r < 16 and (r - v) % 32 == 0


def test_serialization_formats_and_roundtrip():
    # Serialization.cpp:3-13: u32 nBytes, bool neg, little-endian magnitude; zero -> nBytes = 0
    assert O.export_zz(0) == bytes.fromhex("0000000000")
    assert O.export_zz(-255) == bytes.fromhex("0100000001ff")
    assert O.export_zz(1 << 16) == bytes.fromhex("0300000000000001")
    assert O.export_zzx([0, 0, 0]) == bytes.fromhex("ffffffff")  # deg(0) = -1
    assert O.export_zzx([3, 0, -1, 0])[:4] == (2).to_bytes(4, "little")
    ctx = O.Context(22, 80, 23, 7).setup_si()
    rng = O.Rng(4)
    sk = O.SecKey.generate(ctx, rng)
    pk = O.PubKey.generate(sk, rng)
    ct = O.encrypt_rng(pk, [1, 2, 3], rng)
    blob = O.export_ciphertext(ct)
    back, off = O.import_ciphertext(ctx, blob)
    assert off == len(blob) and back.parts == ct.parts and O.export_ciphertext(back) == blob
    cblob = O.export_context(ctx)
    ctx2 = O.import_context(cblob)
    assert (ctx2.m, ctx2.logQ, ctx2.p, ctx2.g, ctx2.primes, ctx2.roots) == (22, 80, 23, 7, ctx.primes, ctx.roots)
    assert O.export_context(ctx2) == cblob
    rows = O.dcrt_rows(ctx, pk.pk[0])
    rblob = O.export_dcrt(rows)
    assert O.import_dcrt(rblob, 0) == (rows, len(rblob))
    # a scaledUp ciphertext exports its ScaleDown'd parts (Serialization.cpp:109-114)
    t = ct.copy().mul(ct)
    assert O.export_ciphertext(t) == O.export_ciphertext(t.copy().scale_down())


def _golden_scenario(name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    p = GOLD["configs"][name]["params"]
    return mg, mg.scenario(p["logQ"], p["p"], p["g"], GOLD["seed"])


def test_golden_cfg1_full_vectors():
    mg, (ctx, sk, pk, ks, msgs, rand, cts) = _golden_scenario("cfg1")
    g = GOLD["configs"]["cfg1"]
    assert O.export_context(ctx).hex() == g["context"]
    assert [O.export_zzx(x).hex() for x in pk.pk] == g["pk"]
    assert [O.export_zzx(x).hex() for x in ks.b] == g["ksw_b"]
    assert [O.export_ciphertext(c).hex() for c in cts] == g["cts"]
    out = mg.outputs(ctx, sk, ks, cts)
    assert {k: v.hex() for k, v in out.items()} == g["out"]
    # and from the stored inputs alone (no RNG): import keys + randomness, recompute
    imp = lambda h: O.import_zzx(bytes.fromhex(h), 0, ctx.phim)[0]
    pk2 = O.PubKey(ctx, [imp(h) for h in g["pk"]])
    cts2 = [O.encrypt(pk2, m, r, e) for m, r, e in zip(g["msgs"], g["r"], g["e"])]
    assert [O.export_ciphertext(c).hex() for c in cts2] == g["cts"]
    ks2 = O.KeySwitch(ctx, [imp(h) for h in g["ksw_b"]], [imp(h) for h in g["ksw_A"]])
    assert O.export_ciphertext(O.mult_relin(ks2, cts2[0], cts2[1])).hex() == g["out"]["mult_relin"]


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_golden_digests(name):
    mg, (ctx, sk, pk, ks, msgs, rand, cts) = _golden_scenario(name)
    g = GOLD["configs"][name]
    assert ctx.primes == g["chain"]
    assert [hashlib.sha256(O.export_ciphertext(c)).hexdigest() for c in cts] == g["cts_sha256"]
    out = mg.outputs(ctx, sk, ks, cts)
    assert {k: hashlib.sha256(v).hexdigest() for k, v in out.items()} == g["out_sha256"]


# ---- the reference's own code (oracle/_ref: reference sources + NTL stand-in) pins the oracle
REF_GOLD = json.load(open(os.path.join(HERE, "golden", "ref_golden.json")))


def _oracle_files(mg, logq, p, g):
    """Everything tests/cpp/host_client.cpp writes, computed by the oracle."""
    ctx, sk, pk, ks, msgs, rand, cts = mg.scenario(logq, p, g, REF_GOLD["seed"])
    files = {"context": O.export_context(ctx), "ct0": O.export_ciphertext(cts[0]), "ct1": O.export_ciphertext(cts[1])}
    files.update(mg.outputs(ctx, sk, ks, cts))
    # vector<DoubleCRT>: u32 count, then each DoubleCRT's rows on the reference chain and roots
    files["pk"] = (2).to_bytes(4, "little") + b"".join(O.export_dcrt(O.dcrt_rows(ctx, x)) for x in pk.pk)
    files["mult_relin_roundtrip"] = files["mult_relin"]
    files["pk_roundtrip"] = files["pk"]
    vec = lambda polys: len(polys).to_bytes(4, "little") + b"".join(O.export_dcrt(O.dcrt_rows(ctx, x)) for x in polys)
    files["sk"] = vec(sk.s)
    if ctx.phim <= 64:  # rows of the 6D key-switch polynomials by direct evaluation: small rings only
        # KeySwitchSI::Export: vector<vector<DoubleCRT>> = {b, A}; A is stored unreduced (FHE-SI.cpp:178-180)
        files["ksw"] = (2).to_bytes(4, "little") + vec(ks.b) + vec(ks.A)
    return files


@pytest.mark.parametrize("name", sorted(REF_GOLD["configs"]))
def test_oracle_matches_reference_golden(name):
    """Byte equality with what the reference's own DoubleCRT / Bluestein / Ciphertext / FHE-SI code
    wrote for the same seed (tests/golden/make_ref_golden.py), file by file."""
    mg, _ = _golden_scenario("cfg1")
    ref = REF_GOLD["configs"][name]
    P = ref["params"]
    if name == "cfg5_512" and os.environ.get("FHESI_SKIP_SLOW"):
        pytest.skip("slow")
    files = _oracle_files(mg, P["logQ"], P["p"], P["g"])
    # ksw: rows of 6D polynomials, small rings only; embed_slots: the oracle has no plaintext space (checked
    # against the host layer and the CUDA kernel in test_host_cpp.py / test_gpu_parity.py)
    assert set(files) | {"ksw", "embed_slots"} == set(ref["sha256"])
    for f, blob in files.items():
        assert hashlib.sha256(blob).hexdigest() == ref["sha256"][f], (name, f)
        if "hex" in ref:
            assert blob.hex() == ref["hex"][f], (name, f)


@pytest.mark.skipif(not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "golden_client_ref")),
                    reason="oracle/_ref not built (needs the reference tree at build time)")
def test_reference_golden_is_reproducible(tmp_path):
    """Re-run the reference build on cfg1 and cfg3: the committed JSON is what it writes."""
    import subprocess
    exe = os.path.join(HERE, "..", "oracle", "_ref", "golden_client_ref")
    for name in ("cfg1", "cfg3"):
        P = REF_GOLD["configs"][name]["params"]
        d = tmp_path / name
        d.mkdir()
        subprocess.check_call([exe, str(P["logQ"]), str(P["p"]), str(P["g"]), str(REF_GOLD["seed"]), str(d)],
                              stdout=subprocess.DEVNULL)
        for f, h in REF_GOLD["configs"][name]["sha256"].items():
            assert hashlib.sha256(open(d / (f + ".bin"), "rb").read()).hexdigest() == h, (name, f)
