"""Generate the committed golden vectors from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no known-answer vectors
(SURVEY.md §4, §8c); these are the ORACLE's outputs, and tests/golden/ref_golden.json
(make_ref_golden.py) holds the same scenario run through the reference's own sources
(oracle/_ref) -- tests/test_oracle.py requires the two to agree byte for byte.  Any later
change to oracle/ or to the CUDA path must still reproduce them.

cfg1 (logQ=80, p=23, g=7): full inputs and outputs in the reference's Export byte format
(Serialization.cpp:3-119), hex-encoded.  cfg2/cfg3/cfg4 (p=1019, g=3): inputs are derived from
the seed, outputs are pinned by SHA-256 of their Export bytes (keeps the fixture small)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import fhesi_oracle as O  # noqa: E402

SEED = 20240611


def scenario(logq, p, g, seed=SEED, nct=2):
    # the client builds its context before it seeds the stream: the chain's roots of unity are
    # drawn from a fresh stream (CModulus.cpp:66-76), everything else from the seeded one
    ctx = O.Context(p - 1, logq, p, g).setup_si(rng=O.Rng(0))
    rng = O.Rng(seed)
    sk = O.SecKey.generate(ctx, rng)
    pk = O.PubKey.generate(sk, rng)
    ks = O.KeySwitch.init_s2(sk, rng)
    msgs = [[rng.random_bnd(p) for _ in range(ctx.phim)] for _ in range(nct)]
    rand = []
    for _ in range(nct):
        r = [rng.random_bnd(2) for _ in range(ctx.phim)]
        e = [O.sample_gaussian(rng, ctx.phim, ctx.stdev) for _ in range(2)]
        rand.append((r, e))
    cts = [O.encrypt(pk, m, r, e) for m, (r, e) in zip(msgs, rand)]
    # later draws of tests/cpp/host_client.cpp: one more encryption (its import check), then the
    # rotation key of its second group
    O.encrypt_rng(pk, msgs[0], rng)
    ks.rot_g = O.KeySwitch.init_automorph(sk, g, rng)
    ks.rot_k = g
    ks.plain = msgs[1]
    return ctx, sk, pk, ks, msgs, rand, cts


def outputs(ctx, sk, ks, cts):
    a, b = cts
    out = {}
    out["add"] = O.export_ciphertext(a.copy().add(b))
    t = a.copy().mul(b)
    out["tensor_scaledown"] = O.export_ciphertext(t)          # Export applies ScaleDown
    mr = O.mult_relin(ks, a, b)
    out["mult_relin"] = O.export_ciphertext(mr)
    out["decrypt_mult_relin"] = O.export_zzx(O.decrypt(sk, mr))
    sq = O.mult_relin(ks, mr, mr)
    out["square_relin"] = O.export_ciphertext(sq)
    out["mul_scalar_m7"] = O.export_ciphertext(a.copy().mul_scalar(-7))
    out["automorph_3"] = O.export_ciphertext(a.copy().automorph(3 if ctx.m % 3 else 5))  # as host_client.cpp
    # second group
    acc = a.copy().mul(b).add(b.copy().mul(b))
    out["tensor_accumulate"] = O.export_ciphertext(acc)
    out["tensor_mul_scalar"] = O.export_ciphertext(acc.copy().mul_scalar(5))
    out["accumulate_relin"] = O.export_ciphertext(O.apply_key_switch(ks, acc.copy()))
    out["mul_plain"] = O.export_ciphertext(a.copy().mul_plain(ks.plain))
    out["add_plain"] = O.export_ciphertext(a.copy().add_plain(ks.plain))
    out["add_3part"] = O.export_ciphertext(a.copy().mul(b).scale_down().add(a))
    rot = O.apply_key_switch(ks.rot_g, a.copy().automorph(ks.rot_k))
    out["rotate_keyswitch"] = O.export_ciphertext(rot)
    out["decrypt_rotate"] = O.export_zzx(O.decrypt(sk, rot))
    # third group: tensor-form branches of += ZZX, >>=, *= ZZX
    out["tensor_add_plain"] = O.export_ciphertext(a.copy().mul(b).add_plain(ks.plain))
    out["tensor_automorph"] = O.export_ciphertext(a.copy().mul(b).automorph(ks.rot_k))
    out["tensor_mul_plain"] = O.export_ciphertext(a.copy().mul(b).mul_plain([1, 1] + [0] * (ctx.phim - 2)))
    # fourth group: a product whose left operand is the unreduced output of >>= (the reference's semantics:
    # the extra multiple of q is carried into the tensor product)
    out["unreduced_rot_mul"] = O.export_ciphertext(a.copy().automorph(ks.rot_k).mul(b))
    return out


def main():
    gold = {"seed": SEED, "configs": {}}
    ctx, sk, pk, ks, msgs, rand, cts = scenario(80, 23, 7)
    full = {
        "params": {"logQ": 80, "p": 23, "g": 7},
        "context": O.export_context(ctx).hex(),
        "sk": [O.export_zzx(s).hex() for s in sk.s],
        "pk": [O.export_zzx(x).hex() for x in pk.pk],
        "ksw_b": [O.export_zzx(x).hex() for x in ks.b],
        "ksw_A": [O.export_zzx(x).hex() for x in ks.A],
        "msgs": msgs,
        "r": [r for r, _ in rand],
        "e": [e for _, e in rand],
        "cts": [O.export_ciphertext(c).hex() for c in cts],
        "out": {k: v.hex() for k, v in outputs(ctx, sk, ks, cts).items()},
    }
    gold["configs"]["cfg1"] = full
    for name, (logq, p, g) in {"cfg2": (256, 1019, 3), "cfg3": (100, 1019, 3), "cfg4": (176, 1019, 3)}.items():
        ctx, sk, pk, ks, msgs, rand, cts = scenario(logq, p, g)
        gold["configs"][name] = {
            "params": {"logQ": logq, "p": p, "g": g},
            "chain": ctx.primes,
            "cts_sha256": [hashlib.sha256(O.export_ciphertext(c)).hexdigest() for c in cts],
            "out_sha256": {k: hashlib.sha256(v).hexdigest() for k, v in outputs(ctx, sk, ks, cts).items()},
        }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "golden.json"))


if __name__ == "__main__":
    main()
