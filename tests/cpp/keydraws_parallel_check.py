"""Run in a subprocess WITHOUT FHESI_TEST_RNG (the ChaCha20 stream): fhesih_keydraws_flat with the long draws filled on
several cores from their stream positions must return exactly what the in-order walk (FHESIH_SEQ_DRAWS=1) returns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "fhe-si_b200"))
import pyfhesi  # noqa: E402
from pyfhesi.hostkeys import keydraws_flat  # noqa: E402

assert "FHESI_TEST_RNG" not in os.environ
lib = sys.argv[1]
for (m, logq, p, g, rot) in ((22, 80, 23, 7, [7, 5]), (1018, 176, 1019, 3, [3, 9, 81])):
    ctx = pyfhesi.Context(m, logq, p, 3, 4, 0, lib_path=lib)
    for seed in (1, 20240611):
        os.environ.pop("FHESIH_SEQ_DRAWS", None)
        par = keydraws_flat(ctx, seed, g, rot_k=rot, lib_path=lib)
        os.environ["FHESIH_SEQ_DRAWS"] = "1"
        seq = keydraws_flat(ctx, seed, g, rot_k=rot, lib_path=lib)
        os.environ.pop("FHESIH_SEQ_DRAWS")
        os.environ["FHESIH_TEST_REJECT"] = "1"  # as if a Gaussian word had been rejected: rewind, walk in order
        redo = keydraws_flat(ctx, seed, g, rot_k=rot, lib_path=lib)
        os.environ.pop("FHESIH_TEST_REJECT")
        for k in ("sk", "src", "A", "e"):
            assert np.array_equal(par[k], seq[k]), (m, seed, k)
            assert np.array_equal(redo[k], seq[k]), (m, seed, k, "after a rejection")
        other = keydraws_flat(ctx, seed + 1, g, rot_k=rot, lib_path=lib)
        assert not np.array_equal(other["A"], seq["A"])
    ctx.close()
print("keydraws parallel == sequential")
