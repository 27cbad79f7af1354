"""Golden vectors made by the REFERENCE's own code: runs oracle/_ref/golden_client_ref (the
reference's library sources compiled against the NTL stand-in, oracle/build_ref.py; the client is
tests/cpp/host_client.cpp built against the reference's headers) on every configuration and
records the bytes it writes -- context, two fresh ciphertexts, add, tensor + ScaleDown, mult +
relinearise, its decryption, a second-level square, scalar multiple, automorphism, the public key
as DoubleCRT rows, tensor-form accumulation and its key switch, tensor-form scalar multiple,
plaintext multiply and add, a 3-part + 2-part sum, a rotation with its key switch -- in tests/golden/ref_golden.json (hex for cfg1, SHA-256 otherwise).

Run in the build container (needs /root/reference):  python tests/golden/make_ref_golden.py
The JSON is committed; tests compare the oracle, the host layer and the CUDA path against it."""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from build_ref import build_ref  # noqa: E402

SEED = 20240611
CONFIGS = {  # name: (logQ, p, g) -- BASELINE.json configs 1-5 (g = 3 for p = 1019)
    "cfg1": (80, 23, 7), "cfg2": (256, 1019, 3), "cfg3": (100, 1019, 3), "cfg4": (176, 1019, 3),
    "cfg5_128": (128, 1019, 3), "cfg5_512": (512, 1019, 3),
    # m = p - 1 = 2 q^k, k >= 2: not 2 * prime (the remainder by Phi_m is no longer the alternating fold), yet Z_m^* is
    # still cyclic, which the reference's PlaintextSpace::FindSlots needs (PlaintextSpace.cpp:80-96).  m = 18, 162,
    # 250, 1458 (phi = 6, 54, 100, 486: transform lengths 16, 128, 256 and the fused 1024).
    "gm18": (80, 19, 5), "gm162": (100, 163, 5), "gm250": (100, 251, 3), "gm1458": (128, 1459, 5),
}
FILES = ["context", "ct0", "ct1", "add", "tensor_scaledown", "mult_relin", "decrypt_mult_relin", "square_relin",
         "mul_scalar_m7", "automorph_3", "pk", "mult_relin_roundtrip", "pk_roundtrip", "tensor_accumulate",
         "tensor_mul_scalar", "accumulate_relin", "mul_plain", "add_plain", "add_3part", "rotate_keyswitch",
         "decrypt_rotate", "tensor_add_plain", "tensor_automorph", "tensor_mul_plain", "sk", "ksw",
         "unreduced_rot_mul", "embed_slots"]


def run(exe, logq, p, g, seed=SEED):
    with tempfile.TemporaryDirectory() as d:
        # the golden files are made with the deterministic TEST stream (SplitMix64, shared with the oracle)
        subprocess.check_call([exe, str(logq), str(p), str(g), str(seed), d], stdout=subprocess.DEVNULL,
                              env=dict(os.environ, FHESI_TEST_RNG="splitmix64"))
        return {f: open(os.path.join(d, f + ".bin"), "rb").read() for f in FILES}


def main():
    exes = build_ref()
    if "golden_client_ref" not in exes:
        raise SystemExit("oracle/_ref is not built and /root/reference is not mounted")
    gold = {"seed": SEED, "generator": "oracle/_ref/golden_client_ref: reference sources + oracle/ntl_compat",
            "configs": {}}
    only = sys.argv[1:]  # names: regenerate these and keep the other entries of the committed file
    if only:
        gold = json.load(open(os.path.join(HERE, "ref_golden.json")))
    for name, (logq, p, g) in CONFIGS.items():
        if only and name not in only:
            continue
        out = run(exes["golden_client_ref"], logq, p, g)
        entry = {"params": {"logQ": logq, "p": p, "g": g},
                 "sha256": {f: hashlib.sha256(b).hexdigest() for f, b in out.items()}}
        if name == "cfg1":
            entry["hex"] = {f: b.hex() for f, b in out.items()}
        gold["configs"][name] = entry
        print(name, "done", flush=True)
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote", os.path.join(HERE, "ref_golden.json"))


if __name__ == "__main__":
    main()
