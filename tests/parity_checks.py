"""Parity checks shared by the CPU (emulator) and GPU test modules.  Each check runs one
C-ABI entry point on explicit inputs and compares with the oracle bit for bit."""
import numpy as np

import fhesi_oracle as O
from common import Scenario, assert_ct_equal


def check_mult_relin(sc: Scenario, count=2, host=False, random_inputs=False):
    if random_inputs:
        A, B = sc.random_cts(count), sc.random_cts(count)
    else:
        _, cts = sc.fresh(2 * count)
        A, B = cts[:count], cts[count:]
    out = sc.dev_mult_relin(A, B, host=host)
    for i in range(count):
        want = O.mult_relin(sc.ks, A[i], B[i])
        assert_ct_equal(sc, out[i], want, f"mult_relin[{i}]")


def check_mult_relin_wide(logq, p, g, lib_path, pairs=64, seeds=(101, 202, 303), device=0, oracle_pairs=2):
    """Breadth: `pairs` fresh ciphertext pairs x len(seeds) independent key sets, every output word compared.
    The checker for all pairs is oracle/ref_restate.c -- the reference's ALGORITHM restated in C (m-point
    Bluestein per 60-bit chain prime, incremental big-integer CRT), itself pinned to the exact-integer oracle by
    tests/test_oracle.py; the first `oracle_pairs` of every seed are also checked against the Python oracle
    directly.  Half of the pairs are fresh encryptions, half uniformly random reduced parts (full range)."""
    import ref_port
    for seed in seeds:
        sc = Scenario(logq, p, g, seed=seed, lib_path=lib_path, device=device)
        port = ref_port.RefPort(sc.octx)
        port.set_key_switch(sc.ks)
        nf = pairs // 2
        _, cts = sc.fresh(2 * nf)
        A = cts[:nf] + sc.random_cts(pairs - nf)
        B = cts[nf:] + sc.random_cts(pairs - nf)
        out = sc.dev_mult_relin(A, B)
        pa, pb = sc.pack_cts(A), sc.pack_cts(B)
        for i in range(pairs):
            want = port.mult_relin(pa[i], pb[i])
            if not np.array_equal(out[i], want):
                bad = np.argwhere(out[i] != want)
                raise AssertionError(f"seed {seed} pair {i}: {len(bad)} words differ from the reference algorithm, "
                                     f"first at (part, coeff, word) = {bad[0].tolist()}")
        for i in list(range(oracle_pairs)) + [pairs - 1]:
            assert_ct_equal(sc, out[i], O.mult_relin(sc.ks, A[i], B[i]), f"seed {seed} mult_relin[{i}] vs oracle")
        sc.dev.close()


def check_ksw_generate(sc: Scenario, g):
    """fhesi_ksw_generate = KeySwitchSI::Init on the device from explicit draws: b and A' against the
    oracle's KeySwitch.init on the same draws (s^2 -> s matrix and a rotation matrix), and a mult+relin
    through the generated matrix against the oracle's."""
    import copy
    d, octx, logq = sc.dev, sc.octx, sc.logq
    n, D = octx.phim, octx.ndigits
    pack = lambda polys: np.stack([O.pack_poly_words(a, logq) for a in polys])
    s1 = sc.sk.s[1]
    cases = {"s2": [sc.sk.s[0], s1, octx.ring.mul(s1, s1)], "rot": [octx.ring.automorph(x, g) for x in sc.sk.s]}
    handles = {}
    for name, src in cases.items():
        rng = O.Rng(1234 + len(src))
        rng2 = copy.deepcopy(rng)
        A, E = [], []
        for _ in range(len(src) * D):           # FHE-SI.cpp:176,188: SampleRandom then sampleGaussian per entry
            A.append(O.sample_random(rng, octx.q, n))
            E.append(O.sample_gaussian(rng, n, octx.stdev))
        want = O.KeySwitch.init(octx, src, s1, rng2)
        h, b_out, a_out = d.ksw_generate(np.array(src, dtype=np.int64), np.array(s1, dtype=np.int64), pack(A),
                                         np.array(E, dtype=np.int64), want_host=True)
        for k in range(len(src) * D):
            assert O.unpack_poly_words(b_out[k]) == list(want.b[k]), f"{name}: b[{k}]"
            assert O.unpack_poly_words(a_out[k]) == O.reduce_poly(want.A[k], logq), f"{name}: A[{k}]"
        handles[name] = (h, want)
    # the generated s^2 matrix relinearises like the oracle's
    _, cts = sc.fresh(2)
    da, db = d.to_device(sc.pack_cts(cts[:1])), d.to_device(sc.pack_cts(cts[1:]))
    do = d.alloc(d.ct_words(2) * 4)
    d.mult_relin_dev(handles["s2"][0], da.ptr, db.ptr, do.ptr, 1)
    d.sync()
    assert_ct_equal(sc, do.download((1, 2, d.n, d.W))[0], O.mult_relin(handles["s2"][1], cts[0], cts[1]),
                    "mult_relin with a device-generated matrix")
    for h, _ in handles.values():
        d.lib.fhesi_ksw_destroy(h)


def check_rotate_keyswitch(sc: Scenario, g, count=2, compare_steps=True):
    """One SumBatchedData step -- tmp >>= k; KeySwitchSI(sk, k).ApplyKeySwitch(tmp) (Regression.h:166-178)
    -- through fhesi_rotate_keyswitch_dev (rotation folded into the digit extraction) against the
    oracle, and against the three separate calls."""
    d, logq = sc.dev, sc.logq
    rot = O.KeySwitch.init_automorph(sc.sk, g, sc.rng)
    pack = lambda polys: np.stack([O.pack_poly_words(a, logq) for a in polys])
    rksw = d.ksw_create(pack(rot.b), pack([O.reduce_poly(a, logq) for a in rot.A]), 2)
    _, cts = sc.fresh(count)
    din = d.to_device(sc.pack_cts(cts))
    dout = d.alloc(count * d.ct_words(2) * 4)
    d.rotate_keyswitch_dev(rksw, din.ptr, g, dout.ptr, count)
    if compare_steps:
        dw = d.alloc(count * 2 * d.n * (d.W + 1) * 4)
        dr, do2 = d.alloc(count * d.ct_words(2) * 4), d.alloc(count * d.ct_words(2) * 4)
        d.ct_automorph_dev(din.ptr, 2, g, dw.ptr, count)
        d.reduce_wide_dev(dw.ptr, d.W + 1, dr.ptr, 2, count)
        d.keyswitch_dev(rksw, dr.ptr, do2.ptr, count)
    d.sync()
    got = dout.download((count, 2, d.n, d.W))
    if compare_steps:
        assert np.array_equal(got, do2.download((count, 2, d.n, d.W))), \
            "fused rotation differs from automorph + reduce + key switch"
    for i in range(count):
        assert_ct_equal(sc, got[i], O.apply_key_switch(rot, cts[i].copy().automorph(g)), f"rotate_keyswitch[{i}]")
    d.lib.fhesi_ksw_destroy(rksw)


def check_pieces(sc: Scenario, count=2):
    """tensor -> (tprod add) -> ScaleDown -> key switch as separate calls, plus decrypt."""
    d = sc.dev
    msgs, cts = sc.fresh(2 * count)
    A, B = cts[:count], cts[count:]
    da, db = d.to_device(sc.pack_cts(A)), d.to_device(sc.pack_cts(B))
    dt = d.alloc(count * d.tprod_words(3) * 4)
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, count)
    # c*c + c*c in tensor form (Test_AddMul.cpp:73-75 adds tProd-form ciphertexts)
    dt2 = d.alloc(count * d.tprod_words(3) * 4)
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt2.ptr, count)
    d.tprod_add_dev(dt2.ptr, dt.ptr, 3, count)
    dc = d.alloc(count * d.ct_words(3) * 4)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, count)
    d.sync()
    c3 = dc.download((count, 3, d.n, d.W))
    wants = []
    for i in range(count):
        w = A[i].copy().mul(B[i])
        w2 = w.copy().add(w)
        w.scale_down()
        assert_ct_equal(sc, c3[i], w, f"scaledown[{i}]")
        wants.append((w, w2))
    d.scaledown_dev(dt2.ptr, 3, dc.ptr, count)
    d.sync()
    c3b = dc.download((count, 3, d.n, d.W))
    for i in range(count):
        assert_ct_equal(sc, c3b[i], wants[i][1].copy().scale_down(), f"tprod_add+scaledown[{i}]")
    # key switch of the scaled-down 3-part ciphertexts, then decrypt
    dc.upload(c3)
    do = d.alloc(count * d.ct_words(2) * 4)
    d.keyswitch_dev(sc.ksw, dc.ptr, do.ptr, count)
    dm = d.alloc(count * d.n * 4)
    d.decrypt_dev(sc.dsk, do.ptr, 2, dm.ptr, count)
    d.sync()
    o2 = do.download((count, 2, d.n, d.W))
    dec = dm.download((count, d.n))
    for i in range(count):
        w = O.apply_key_switch(sc.ks, wants[i][0])
        assert_ct_equal(sc, o2[i], w, f"keyswitch[{i}]")
        m = O.decrypt(sc.sk, w)
        assert dec[i].tolist() == m, f"decrypt[{i}]"
        prod = [c % sc.p for c in sc.octx.ring.mul(msgs[i], msgs[count + i])]
        assert m == prod, "homomorphic product"  # Test_AddMul.cpp:84-86 identity


def check_tensor_accumulate(sc: Scenario, count=3):
    """sum_i a_i * b_i in tensor form (Matrix.cpp:80-97), then ScaleDown."""
    d = sc.dev
    _, cts = sc.fresh(2 * count)
    A, B = cts[:count], cts[count:]
    da, db = d.to_device(sc.pack_cts(A)), d.to_device(sc.pack_cts(B))
    dt = d.alloc(d.tprod_words(3) * 4)
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, count, accumulate=True)
    dc = d.alloc(d.ct_words(3) * 4)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    got = dc.download((3, d.n, d.W))
    acc = A[0].copy().mul(B[0])
    for i in range(1, count):
        acc.add(A[i].copy().mul(B[i]))
    assert_ct_equal(sc, got, acc.scale_down(), "tensor accumulate")


def check_encrypt_decrypt(sc: Scenario, count=2):
    d = sc.dev
    n = d.n
    rng = sc.rng
    msgs = [[rng.random_bnd(sc.p) for _ in range(n)] for _ in range(count)]
    rs = [[rng.random_bnd(2) for _ in range(n)] for _ in range(count)]
    es = [[O.sample_gaussian(rng, n, sc.octx.stdev) for _ in range(2)] for _ in range(count)]
    dmsg = d.to_device(np.array(msgs, dtype=np.uint32))
    dr = d.to_device(np.array(rs, dtype=np.uint8))
    de = d.to_device(np.array(es, dtype=np.int32))
    dout = d.alloc(count * d.ct_words(2) * 4)
    d.encrypt_dev(sc.dpk, dmsg.ptr, dr.ptr, de.ptr, dout.ptr, count)
    dm = d.alloc(count * n * 4)
    d.decrypt_dev(sc.dsk, dout.ptr, 2, dm.ptr, count)
    d.sync()
    got = dout.download((count, 2, n, d.W))
    dec = dm.download((count, n))
    for i in range(count):
        want = O.encrypt(sc.pk, msgs[i], rs[i], es[i])
        assert_ct_equal(sc, got[i], want, f"encrypt[{i}]")
        assert dec[i].tolist() == msgs[i], f"decrypt(encrypt)[{i}]"


def check_coeff_ops(sc: Scenario, count=3):
    """+=, batch sum, *= long, >>= and Reduce in coefficient form."""
    d = sc.dev
    A, B = sc.random_cts(count), sc.random_cts(count)
    pa = sc.pack_cts(A)
    da, db = d.to_device(pa), d.to_device(sc.pack_cts(B))
    d.ct_add_dev(da.ptr, db.ptr, 2, count)
    d.sync()
    got = da.download(pa.shape)
    for i in range(count):
        assert_ct_equal(sc, got[i], A[i].copy().add(B[i]), f"ct_add[{i}]")
    # batch sum
    da.upload(pa)
    ds = d.alloc(d.ct_words(2) * 4)
    d.ct_sum_dev(da.ptr, ds.ptr, 2, count)
    d.sync()
    acc = A[0].copy()
    for i in range(1, count):
        acc.add(A[i])
    assert_ct_equal(sc, ds.download((2, d.n, d.W)), acc, "ct_sum")
    # scalar multiply, positive and negative
    for l in (7, -3, 123456789012345):
        da.upload(pa)
        d.ct_mul_scalar_dev(da.ptr, l, 2, count)
        d.sync()
        got = da.download(pa.shape)
        for i in range(count):
            assert_ct_equal(sc, got[i], A[i].copy().mul_scalar(l), f"ct_mul_scalar {l} [{i}]")
    # automorphism (not reduced mod q) and Reduce
    units = sc.octx.ring.units
    for k in (units[1], units[len(units) // 2], units[-1]):
        da.upload(pa)
        dw = d.alloc(count * 2 * d.n * (d.W + 1) * 4)
        d.ct_automorph_dev(da.ptr, 2, k, dw.ptr, count)
        dr = d.alloc(count * d.ct_words(2) * 4)
        d.reduce_wide_dev(dw.ptr, d.W + 1, dr.ptr, 2, count)
        d.sync()
        wide = dw.download((count, 2, d.n, d.W + 1))
        red = dr.download(pa.shape)
        for i in range(count):
            want = A[i].copy().automorph(k)
            for part in range(2):
                assert O.unpack_poly_words(wide[i, part]) == want.parts[part], f"automorph k={k} [{i}]"
                assert O.unpack_poly_words(red[i, part]) == O.reduce_poly(want.parts[part], sc.logq)


def check_tprod_scalar(sc: Scenario):
    d = sc.dev
    _, cts = sc.fresh(2)
    da, db = d.to_device(sc.pack_cts(cts[:1])), d.to_device(sc.pack_cts(cts[1:]))
    dt = d.alloc(d.tprod_words(3) * 4)
    dc = d.alloc(d.ct_words(3) * 4)
    for l in (9, -2):
        d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
        d.tprod_mul_scalar_dev(dt.ptr, l, 3, 1)
        d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
        d.sync()
        want = cts[0].copy().mul(cts[1]).mul_scalar(l).scale_down()
        assert_ct_equal(sc, dc.download((3, d.n, d.W)), want, f"tprod *= {l}")


def check_ref_rows(sc: Scenario):
    """DoubleCRT rows on the reference chain (key export parity)."""
    a = O.sample_random(sc.rng, sc.octx.q, sc.octx.phim)
    rows = sc.dev.ref_rows_host(O.pack_poly_words(a, sc.logq), sc.octx.primes, sc.octx.roots)
    want = O.dcrt_rows(sc.octx, a)
    assert rows.tolist() == want


def check_gathered_reduce(sc: Scenario, world=3):
    d = sc.dev
    _, cts = sc.fresh(2 * world)
    bufs = []
    acc = None
    for w in range(world):
        da, db = d.to_device(sc.pack_cts([cts[2 * w]])), d.to_device(sc.pack_cts([cts[2 * w + 1]]))
        dt = d.alloc(d.tprod_words(3) * 4)
        d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
        d.sync()
        bufs.append(dt.download((3, d.Lt, d.N)))
        t = cts[2 * w].copy().mul(cts[2 * w + 1])
        acc = t if acc is None else acc.add(t)
    dg = d.to_device(np.stack(bufs))
    do = d.alloc(d.tprod_words(3) * 4)
    d.tprod_reduce_gathered_dev(dg.ptr, world, 3, do.ptr)
    dc = d.alloc(d.ct_words(3) * 4)
    d.scaledown_dev(do.ptr, 3, dc.ptr, 1)
    d.sync()
    assert_ct_equal(sc, dc.download((3, d.n, d.W)), acc.scale_down(), "gathered reduce")


def check_golden(name, lib_path, device=0):
    """Run the committed golden inputs (tests/golden/golden.json, made by make_golden.py from
    the oracle) through the C ABI and compare the Export bytes of every output."""
    import hashlib
    import importlib.util
    import json
    import os
    import pyfhesi
    here = os.path.dirname(os.path.abspath(__file__))
    gold = json.load(open(os.path.join(here, "golden", "golden.json")))
    g = gold["configs"][name]
    P = g["params"]
    logq, p = P["logQ"], P["p"]
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    ctx, sk, pk, ks, msgs, rand, cts = mg.scenario(logq, p, P["g"], gold["seed"])
    rot, rot_k, plain = ks.rot_g, ks.rot_k, ks.plain
    if "cts" in g:  # full vectors: take every input from the fixture, not from the RNG
        imp = lambda h: O.import_zzx(bytes.fromhex(h), 0, ctx.phim)[0]
        sk = O.SecKey(ctx, [imp(h) for h in g["sk"]])
        pk = O.PubKey(ctx, [imp(h) for h in g["pk"]])
        ks = O.KeySwitch(ctx, [imp(h) for h in g["ksw_b"]], [imp(h) for h in g["ksw_A"]])
        cts = [O.import_ciphertext(ctx, bytes.fromhex(h))[0] for h in g["cts"]]
    d = pyfhesi.Context(p - 1, logq, p, 3, 1, device, lib_path=lib_path)
    pack = lambda polys: np.stack([O.pack_poly_words(a, logq) for a in polys])
    n, W = d.n, d.W
    ksw = d.ksw_create(pack(ks.b), pack([O.reduce_poly(a, logq) for a in ks.A]), 3)
    dsk = d.key_create(pack(sk.s))
    da, db = d.to_device(pack(cts[0].parts)[None]), d.to_device(pack(cts[1].parts)[None])
    unpack = lambda arr: O.Ciphertext(ctx, [O.unpack_poly_words(arr[i]) for i in range(arr.shape[0])])
    got = {}
    # add
    dx = d.to_device(pack(cts[0].parts)[None])
    d.ct_add_dev(dx.ptr, db.ptr, 2, 1)
    d.sync()
    got["add"] = O.export_ciphertext(unpack(dx.download((2, n, W))))
    # tensor, ScaleDown
    dt = d.alloc(d.tprod_words(3) * 4)
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
    dc = d.alloc(d.ct_words(3) * 4)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    got["tensor_scaledown"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    # mult+relin, decrypt, square
    do = d.alloc(d.ct_words(2) * 4)
    d.mult_relin_dev(ksw, da.ptr, db.ptr, do.ptr, 1)
    dm = d.alloc(n * 4)
    d.decrypt_dev(dsk, do.ptr, 2, dm.ptr, 1)
    do2 = d.alloc(d.ct_words(2) * 4)
    d.mult_relin_dev(ksw, do.ptr, do.ptr, do2.ptr, 1)
    d.sync()
    got["mult_relin"] = O.export_ciphertext(unpack(do.download((2, n, W))))
    got["decrypt_mult_relin"] = O.export_zzx(dm.download((n,)).tolist())
    got["square_relin"] = O.export_ciphertext(unpack(do2.download((2, n, W))))
    # scalar, automorphism (not reduced mod q)
    dx.upload(pack(cts[0].parts)[None])
    d.ct_mul_scalar_dev(dx.ptr, -7, 2, 1)
    dw = d.alloc(2 * n * (W + 1) * 4)
    d.ct_automorph_dev(da.ptr, 2, 3, dw.ptr, 1)
    d.sync()
    got["mul_scalar_m7"] = O.export_ciphertext(unpack(dx.download((2, n, W))))
    got["automorph_3"] = O.export_ciphertext(unpack(dw.download((2, n, W + 1))))
    # ---- second group: tensor-form accumulation a*b + b*b, scalar multiple in tensor form, key switch
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
    dt2 = d.alloc(d.tprod_words(3) * 4)
    d.ct_tensor_dev(db.ptr, 2, db.ptr, 2, dt2.ptr, 1)
    d.tprod_add_dev(dt.ptr, dt2.ptr, 3, 1)
    # the same sum through the batch-accumulating form: pairs (a, b), (b, b) summed into one tprod
    dab = d.to_device(np.stack([pack(cts[0].parts), pack(cts[1].parts)]))
    dbb = d.to_device(np.stack([pack(cts[1].parts), pack(cts[1].parts)]))
    d.ct_tensor_dev(dab.ptr, 2, dbb.ptr, 2, dt2.ptr, 2, accumulate=True)
    dc2 = d.alloc(d.ct_words(3) * 4)
    d.scaledown_dev(dt2.ptr, 3, dc2.ptr, 1)
    d.sync()
    got["tensor_accumulate_batched"] = O.export_ciphertext(unpack(dc2.download((3, n, W))))
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.keyswitch_dev(ksw, dc.ptr, do.ptr, 1)
    d.sync()
    got["tensor_accumulate"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    got["accumulate_relin"] = O.export_ciphertext(unpack(do.download((2, n, W))))
    d.tprod_mul_scalar_dev(dt.ptr, 5, 3, 1)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    got["tensor_mul_scalar"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    # plaintext operands: *= ZZX, and += ZZX as the host layer issues it (floor(c 2^logQ / p) added to part 0)
    dx.upload(pack(cts[0].parts)[None])
    dp = d.to_device(np.array(plain, dtype=np.uint32))
    d.ct_mul_plain_dev(dx.ptr, dp.ptr, 2, 1)
    d.sync()
    got["mul_plain"] = O.export_ciphertext(unpack(dx.download((2, n, W))))
    dx.upload(pack(cts[0].parts)[None])
    scaled = O.reduce_poly([(c << logq) // p for c in plain], logq)
    dsc = d.to_device(pack([scaled])[None])
    d.ct_add_dev(dx.ptr, dsc.ptr, 1, 1)
    d.sync()
    got["add_plain"] = O.export_ciphertext(unpack(dx.download((2, n, W))))
    # 3-part + 2-part: ScaleDown(a*b), then the first two parts += a
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.ct_add_dev(dc.ptr, da.ptr, 2, 1)
    d.sync()
    got["add_3part"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    # rotation: signed permutation, reduction mod q, key switch with the (1, s(X^k)) -> s matrix
    rksw = d.ksw_create(pack(rot.b), pack([O.reduce_poly(a, logq) for a in rot.A]), 2)
    d.ct_automorph_dev(da.ptr, 2, rot_k, dw.ptr, 1)
    dr = d.alloc(d.ct_words(2) * 4)
    d.reduce_wide_dev(dw.ptr, W + 1, dr.ptr, 2, 1)
    d.keyswitch_dev(rksw, dr.ptr, do.ptr, 1)
    d.decrypt_dev(dsk, do.ptr, 2, dm.ptr, 1)
    d.sync()
    got["rotate_keyswitch"] = O.export_ciphertext(unpack(do.download((2, n, W))))
    got["decrypt_rotate"] = O.export_zzx(dm.download((n,)).tolist())
    # the same through the fused entry point (rotation folded into the digit extraction), batch of 3
    d3 = d.to_device(np.stack([pack(cts[0].parts)] * 3))
    do3 = d.alloc(3 * d.ct_words(2) * 4)
    d.rotate_keyswitch_dev(rksw, d3.ptr, rot_k, do3.ptr, 3)
    d.sync()
    fused = do3.download((3, 2, n, W))
    for b in range(3):
        assert O.export_ciphertext(unpack(fused[b])) == got["rotate_keyswitch"], "fhesi_rotate_keyswitch_dev"
    d.lib.fhesi_ksw_destroy(rksw)
    assert got.pop("tensor_accumulate_batched") == got["tensor_accumulate"]
    # ---- third group: tensor-form += ZZX, >>=, *= ZZX (Ciphertext.cpp:157-159, 269-273, 252-256)
    pack1 = lambda poly: np.stack([O.pack_poly_words(poly, 32 * (W + 1))])  # W + 1 words per coefficient
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
    dpl = d.to_device(pack1([(c << logq) // p for c in plain]))
    d.tprod_add_poly_dev(dt.ptr, 3, dpl.ptr, W + 1, 1)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    got["tensor_add_plain"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    d.ct_tensor_dev(da.ptr, 2, db.ptr, 2, dt.ptr, 1)
    d.tprod_automorph_dev(dt.ptr, 3, rot_k, dt2.ptr, 1)
    d.scaledown_dev(dt2.ptr, 3, dc.ptr, 1)
    d.sync()
    got["tensor_automorph"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    dmul = d.to_device(pack1([1, 1] + [0] * (n - 2)))
    d.tprod_mul_poly_dev(dt.ptr, 3, dmul.ptr, W + 1, 1)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    got["tensor_mul_plain"] = O.export_ciphertext(unpack(dc.download((3, n, W))))
    # ---- fourth group: the documented deviation, pinned.  x = a; x >>= k; x *= b.  The reference multiplies the
    # UNREDUCED a(X^k) (coefficients in (-q, q)); the C ABI's tensor product takes W-word operands, so a caller
    # reduces first (fhesi_reduce_wide_dev) -- what the host layer's Ciphertext *= does.  The two results differ
    # exactly by the carried multiple of q: with a(X^k) = r + q K (r reduced), the reference's tensor is
    # p (r + q K) * b, ours p r * b, and ScaleDown turns the difference into Reduce(p K * b) per part.
    d.ct_automorph_dev(da.ptr, 2, rot_k, dw.ptr, 1)
    d.reduce_wide_dev(dw.ptr, W + 1, dr.ptr, 2, 1)
    d.ct_tensor_dev(dr.ptr, 2, db.ptr, 2, dt.ptr, 1)
    d.scaledown_dev(dt.ptr, 3, dc.ptr, 1)
    d.sync()
    ours = unpack(dc.download((3, n, W)))
    wide = cts[0].copy().automorph(rot_k)                                  # the reference's operand
    red = O.Ciphertext(ctx, [O.reduce_poly(x, logq) for x in wide.parts])  # ours
    assert O.export_ciphertext(ours) == O.export_ciphertext(red.copy().mul(cts[1])), "reduce-first product"
    K = [[(w - r) >> logq for w, r in zip(wp, rp)] for wp, rp in zip(wide.parts, red.parts)]
    assert all(set(k) <= {-1, 0, 1} for k in K)
    ref_ct = wide.copy().mul(cts[1]).scale_down()
    corr = [[0] * ctx.phim for _ in range(3)]
    for i in range(2):
        for j in range(2):
            pr = ctx.ring.mul([p * c for c in K[i]], cts[1].parts[j])
            corr[i + j] = [x + y for x, y in zip(corr[i + j], pr)]
    want = [O.reduce_poly([x + y for x, y in zip(op, cp)], logq) for op, cp in zip(ours.parts, corr)]
    assert [list(x) for x in ref_ct.parts] == want, "reference = ours + Reduce(p K * b)"
    got["unreduced_rot_mul"] = O.export_ciphertext(ref_ct) if any(any(k) for k in K) else O.export_ciphertext(ours)
    assert any(any(k) for k in K) == (O.export_ciphertext(ref_ct) != O.export_ciphertext(ours))
    if "out" in g:
        assert {k: v.hex() for k, v in got.items()} == g["out"]
    else:
        assert {k: hashlib.sha256(v).hexdigest() for k, v in got.items()} == g["out_sha256"]


def check_embed_slots(sc: Scenario, g, count=5):
    """fhesi_embed_slots_dev = PlaintextSpace::EmbedInSlots for a batch (PlaintextSpace.cpp:112-134):
    exact against integer arithmetic, and the defining property -- the embedded polynomial evaluates
    to the k-th value at the k-th slot root (DecodeSlots, :136-146); ragged (short) value rows."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "apps"))
    from fhesi_app import Slots
    d, p, m = sc.dev, sc.p, sc.octx.m
    n = d.n
    phi = [int(c) for c in sc.octx.ring.PhimX]
    sl = Slots(m, p, g, phi)
    nslots = sl.total
    vals = np.zeros((count, nslots), dtype=np.uint32)
    for c in range(count):
        width = nslots if c == 0 else (1 if c == 1 else sc.rng.random_bnd(nslots) + 1)
        vals[c, :width] = [sc.rng.random_bnd(p) for _ in range(width)]
    if count > 2:
        vals[2] = 0                                   # the zero plaintext
    dbasis = d.to_device(sl.basis.astype(np.uint32))
    dvals = d.to_device(vals)
    dmsg = d.alloc(count * n * 4)
    d.embed_slots_dev(dbasis.ptr, nslots, dvals.ptr, dmsg.ptr, count)
    d.sync()
    got = dmsg.download((count, n))
    B = [[int(x) for x in row] for row in sl.basis]
    for c in range(count):
        want = [sum(int(vals[c, k]) * B[k][j] for k in range(nslots)) % p for j in range(n)]
        assert got[c].tolist() == want, f"embed_slots row {c}"
        for k in range(nslots):                       # evaluate at the slot's root
            acc = 0
            for coef in reversed(want):
                acc = (acc * sl.roots[k] + coef) % p
            assert acc == int(vals[c, k]), f"slot {k} of row {c} does not decode"
    # DecodeSlots on the device (fhesi_decode_slots_dev) undoes the embedding: exact against Horner evaluation
    V = np.array([[pow(r, j, p) for r in sl.roots] for j in range(n)], dtype=np.uint32)
    dV = d.to_device(V)
    dback = d.alloc(count * nslots * 4)
    d.decode_slots_dev(dV.ptr, nslots, dmsg.ptr, dback.ptr, count)
    d.sync()
    assert np.array_equal(dback.download((count, nslots)), vals), "decode_slots(embed_slots(v)) != v"
    # Against the REFERENCE's own PlaintextSpace::EmbedInSlots (oracle/_ref via tests/golden/ref_golden.json:
    # slot k holds (7k + 3) mod p).  Which root is slot 0 follows from the factoring order of Phi_m mod p, which
    # the reference does not pin, so the two agree up to a cyclic shift of the slot vector: the reference's
    # polynomial must be the kernel's embedding of ONE of the n rotations -- all n computed in one launch.
    import hashlib
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.json")))
    for cfg in ref["configs"].values():
        P = cfg["params"]
        if (P["logQ"], P["p"], P["g"]) != (sc.logq, p, g) or "embed_slots" not in cfg["sha256"]:
            continue
        base = [(7 * k + 3) % p for k in range(nslots)]
        rot = np.array([base[k:] + base[:k] for k in range(nslots)], dtype=np.uint32)
        drot = d.to_device(rot)
        dout = d.alloc(nslots * n * 4)
        d.embed_slots_dev(dbasis.ptr, nslots, drot.ptr, dout.ptr, nslots)
        d.sync()
        polys = dout.download((nslots, n))
        digests = {hashlib.sha256(O.export_zzx(row.tolist())).hexdigest() for row in polys}
        assert cfg["sha256"]["embed_slots"] in digests, "reference EmbedInSlots is not a slot rotation of the kernel's"
        break
    else:
        raise AssertionError("no reference golden for this parameter set")


def check_mul_plain(sc: Scenario, count=2):
    """Ciphertext *= ZZX in coefficient form (Ciphertext.cpp:246-250)."""
    d = sc.dev
    A = sc.random_cts(count)
    pt = [sc.rng.random_bnd(sc.p) for _ in range(d.n)]
    da = d.to_device(sc.pack_cts(A))
    dp = d.to_device(np.array(pt, dtype=np.uint32))
    d.ct_mul_plain_dev(da.ptr, dp.ptr, 2, count)
    d.sync()
    got = da.download((count, 2, d.n, d.W))
    for i in range(count):
        assert_ct_equal(sc, got[i], A[i].copy().mul_plain(pt), f"ct *= plain [{i}]")


def check_edge_cases(sc: Scenario, counts=(0, 1, 7)):
    """Extreme operands and ragged batch sizes: every coefficient at -q/2 or q/2-1, the zero
    ciphertext, unit polynomials; batch sizes 0, 1 and one that is not a multiple of the kernels'
    group size."""
    octx, n, logq = sc.octx, sc.octx.phim, sc.logq
    lo, hi = -(1 << (logq - 1)), (1 << (logq - 1)) - 1
    mk = lambda parts: O.Ciphertext(octx, parts)
    specials = [
        mk([[lo] * n, [lo] * n]), mk([[hi] * n, [hi] * n]), mk([[0] * n, [0] * n]),
        mk([[1] + [0] * (n - 1), [0] * (n - 1) + [-1]]), mk([[hi if i % 2 else lo for i in range(n)], [lo] + [hi] * (n - 1)]),
    ]
    pairs = [(a, b) for a in specials for b in specials[:3]] + [(specials[3], specials[4])]
    A, B = [p[0] for p in pairs], [p[1] for p in pairs]
    out = sc.dev_mult_relin(A, B)
    for i in range(len(pairs)):
        assert_ct_equal(sc, out[i], O.mult_relin(sc.ks, A[i], B[i]), f"edge mult_relin[{i}]")
    host = sc.dev_mult_relin(A[:5], B[:5], host=True)
    assert np.array_equal(host, out[:5])
    for cnt in counts:
        if cnt == 0:
            d = sc.dev
            buf = d.alloc(16)
            d.mult_relin_dev(sc.ksw, buf.ptr, buf.ptr, buf.ptr, 0)   # empty batch: a no-op, not an error
            d.ct_add_dev(buf.ptr, buf.ptr, 2, 0)
            d.sync()
            continue
        X, Y = sc.random_cts(cnt), sc.random_cts(cnt)
        got = sc.dev_mult_relin(X, Y)
        for i in (0, cnt - 1):
            assert_ct_equal(sc, got[i], O.mult_relin(sc.ks, X[i], Y[i]), f"count={cnt} [{i}]")


def check_keygen_batch(lib, cfg, xi, rot, seed=5):
    """keydraws_flat + Context.keygen_batch (one pass of kernels for every matrix and the public key) give
    byte for byte what the C++ classes compute on the host (keygen(): FHESIPubKey::Init, KeySwitchSI::Init)."""
    import numpy as np
    import pyfhesi
    from pyfhesi.hostkeys import keydraws, keydraws_flat, keygen
    logq, p, g = cfg
    dev = pyfhesi.Context(p - 1, logq, p, 3, xi, 0, lib_path=lib)
    k = keygen(dev, seed, g, rot_k=rot, lib_path=lib)
    f = keydraws_flat(dev, seed, g, rot_k=rot, lib_path=lib)
    d = keydraws(dev, seed, g, rot_k=rot, lib_path=lib)
    D = dev.D
    # the flat draws are the ZZX draws
    assert np.array_equal(f["sk"], d["sk"]) and np.array_equal(f["src"][:3], d["s2_src"])
    assert np.array_equal(f["A"][:3 * D], d["s2_A"]) and np.array_equal(f["e"][:3 * D], d["s2_e"])
    for r in range(len(rot)):
        o = (3 + 2 * r) * D
        assert np.array_equal(f["src"][3 + 2 * r:5 + 2 * r], d["rot_src"][r])
        assert np.array_equal(f["A"][o:o + 2 * D], d["rot_A"][r]) and np.array_equal(f["e"][o:o + 2 * D], d["rot_e"][r])
    ksws, pk, b, a, pkw = dev.keygen_batch(f["parts"], f["src"], f["sk"], f["A"], f["e"], with_pk=True, want_host=True)
    assert len(ksws) == 1 + len(rot) and pk
    assert np.array_equal(pkw, k["pk"]), "public key generated on the device differs from FHESIPubKey::Init"
    assert np.array_equal(b[:3 * D], k["ks_b"]) and np.array_equal(a[:3 * D], k["ks_A"])
    for r in range(len(rot)):
        o = (3 + 2 * r) * D
        assert np.array_equal(b[o:o + 2 * D], k["rot_b"][r]) and np.array_equal(a[o:o + 2 * D], k["rot_A"][r])
    # public key alone
    _, pk2, _, _, pkw2 = dev.keygen_batch([], np.zeros((0, dev.n), np.int32), f["sk"], f["A"][-1:], f["e"][-1:],
                                          with_pk=True, want_host=True)
    assert pk2 and np.array_equal(pkw2, k["pk"])




def check_crt_direct_paths(cfg, lib_path, count=3, m=None, seed=5):
    """ScaleDown three ways -- k_crt_direct (windowed explicit CRT), the same kernel with every thread forced down its
    exact mixed-radix routine (FHESI_CRT_FORCE_EXACT=1), and k_crt (FHESI_NO_CRT_DIRECT=1) -- must each give the
    oracle's bytes: mult+relin (ScaleDown fused with the digits) on fresh and on full-range operands, tensor +
    ScaleDown, and the extreme operands of check_edge_cases (zero and unit polynomials make x tiny, the case in which
    the quotient k cannot be read off the fixed-point sum and the kernel must fall back on its own)."""
    import os
    for env in ({}, {"FHESI_CRT_FORCE_EXACT": "1"}, {"FHESI_NO_CRT_DIRECT": "1"}):
        os.environ.update(env)
        try:
            sc = Scenario(*cfg, seed=seed, lib_path=lib_path, m=m)
        finally:
            for k in env:
                del os.environ[k]
        sc.dev.profile_enable(True)
        check_mult_relin(sc, count=count)
        check_mult_relin(sc, count=count, random_inputs=True)
        check_pieces(sc, count=2)
        check_edge_cases(sc, counts=(1, 5))
        prof = sc.dev.profile_report()
        assert ("k_crt_direct<ML>" in prof) == ("FHESI_NO_CRT_DIRECT" not in env), prof
        if not env:
            # the fallback is for the rare coefficient: tiny x (the zero and unit operands above: counted) aside, fresh
            # encryptions must stay on the windowed path -- a wide "unsure" band is a 2 x slowdown, not a wrong bit
            before = sc.dev.crt_fallbacks()
            assert before > 0, "the extreme operands must have taken the exact routine"
            _, cts = sc.fresh(8)
            sc.dev_mult_relin(cts[:4], cts[4:])
            assert sc.dev.crt_fallbacks() - before <= max(2, 4 * 3 * sc.dev.n // 1000), "k_crt_direct falls back too often"
        sc.dev.close()


def check_mult_relin_host_async(sc: Scenario, count=3):
    """fhesi_mult_relin_host_async: three batches in flight (the staging halves are reused by the third), one
    fhesi_sync_all; every batch must equal the oracle and the blocking call."""
    batches = []
    for _ in range(3):
        A, B = sc.random_cts(count), sc.random_cts(count)
        a = np.ascontiguousarray(sc.pack_cts(A), dtype=np.uint32)
        b = np.ascontiguousarray(sc.pack_cts(B), dtype=np.uint32)
        batches.append((A, B, a, b, np.zeros_like(a)))
    for A, B, a, b, out in batches:
        sc.dev.mult_relin_host_async(sc.ksw, a, b, out, count)
    sc.dev.sync_all()
    for A, B, a, b, out in batches:
        for i in range(count):
            assert_ct_equal(sc, out[i], O.mult_relin(sc.ks, A[i], B[i]), f"async batch [{i}]")
        assert np.array_equal(out, sc.dev.mult_relin_host(sc.ksw, a, b))
