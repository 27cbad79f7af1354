"""Shared parity harness: build a scenario with the oracle, run the same explicit inputs
through the C ABI (CUDA library on the GPU box, or the kernel-logic emulator in CPU CI) and
compare bit for bit."""
import numpy as np

import fhesi_oracle as O
import pyfhesi

CONFIGS = {
    # name: (logQ, p, g)  -- BASELINE.json configs; p=1019 uses g=3
    "cfg1": (80, 23, 7),
    "cfg2": (256, 1019, 3),
    "cfg3": (100, 1019, 3),
    "cfg4": (176, 1019, 3),
    "cfg5_128": (128, 1019, 3),
    "cfg5_512": (512, 1019, 3),
    # the reference README's second parameter family (README:35-37): p = 2027, m = 2026, phi(m) = 1012, N = 2048
    "p2027": (256, 2027, 3),
    "p2027_176": (176, 2027, 3),
}

# General m (not 2 * odd prime): Phi_m is taken as a sparse remainder table (DevCtx::red).  (logQ, p, g, m):
# a power of two, an odd prime, prime powers, three odd prime factors (Phi_105 has a coefficient -2; X^j mod Phi_105
# reaches +-2), phi(m) > m / 2 and < m / 2, and sizes that take the fused kernels (256 < phi(m) <= 1024).
GENERAL_M = {
    "m16": (80, 17, 3, 16),
    "m17": (80, 2, 3, 17),
    "m36": (100, 37, 5, 36),
    "m45": (80, 181, 2, 45),
    "m105": (80, 211, 2, 105),
    "m128": (128, 257, 3, 128),
    "m1320": (100, 1321, 13, 1320),
    "m771": (128, 2, 5, 771),
    "m1285": (100, 257, 2, 1285),  # phi = 1024: the fused N = 2048 kernels' general-m instances, n = N / 2 exactly
}


class Scenario:
    def __init__(self, logq, p, g, seed=1, xi=1, lib_path=None, device=0, m=None):
        m = p - 1 if m is None else m  # the reference's clients take m = p - 1; the library takes any m
        self.octx = O.Context(m, logq, p, g).setup_si(xi)
        self.rng = O.Rng(seed)
        self.sk = O.SecKey.generate(self.octx, self.rng)
        self.pk = O.PubKey.generate(self.sk, self.rng)
        self.ks = O.KeySwitch.init_s2(self.sk, self.rng)
        self.dev = pyfhesi.Context(m, logq, p, 3, xi, device, lib_path=lib_path)
        self.logq, self.p = logq, p
        self._ksw = None
        self._pk = None
        self._sk = None

    # ---- packing
    def pack(self, polys):
        return np.stack([O.pack_poly_words(a, self.logq) for a in polys])

    def pack_cts(self, cts):
        return np.stack([self.pack(ct.parts) for ct in cts])

    def unpack_ct(self, arr):
        return [O.unpack_poly_words(arr[i]) for i in range(arr.shape[0])]

    # ---- device keys
    @property
    def ksw(self):
        if self._ksw is None:
            A_mod = [O.reduce_poly(a, self.logq) for a in self.ks.A]  # value mod q only matters
            self._ksw = self.dev.ksw_create(self.pack(self.ks.b), self.pack(A_mod), 3)
        return self._ksw

    @property
    def dpk(self):
        if self._pk is None:
            self._pk = self.dev.key_create(self.pack(self.pk.pk))
        return self._pk

    @property
    def dsk(self):
        if self._sk is None:
            self._sk = self.dev.key_create(self.pack(self.sk.s))
        return self._sk

    # ---- fresh ciphertexts through the oracle
    def fresh(self, count):
        msgs, cts = [], []
        for _ in range(count):
            m = [self.rng.random_bnd(self.p) for _ in range(self.octx.phim)]
            msgs.append(m)
            cts.append(O.encrypt_rng(self.pk, m, self.rng))
        return msgs, cts

    def random_cts(self, count, parts=2):
        """Uniformly random reduced parts (not valid encryptions): exercises full range."""
        q = self.octx.q
        return [O.Ciphertext(self.octx, [O.sample_random(self.rng, q, self.octx.phim) for _ in range(parts)])
                for _ in range(count)]

    # ---- device runs
    def dev_mult_relin(self, cts_a, cts_b, host=False):
        a, b = self.pack_cts(cts_a), self.pack_cts(cts_b)
        if host:
            return self.dev.mult_relin_host(self.ksw, a, b)
        da, db = self.dev.to_device(a), self.dev.to_device(b)
        dout = self.dev.alloc(a.nbytes)
        self.dev.mult_relin_dev(self.ksw, da.ptr, db.ptr, dout.ptr, len(cts_a))
        self.dev.sync()
        return dout.download(a.shape)


def assert_ct_equal(sc, arr, ct, what=""):
    got = sc.unpack_ct(arr)
    assert len(got) == len(ct.parts), what
    for i, (g, e) in enumerate(zip(got, ct.parts)):
        if g != list(e):
            bad = [k for k in range(len(g)) if g[k] != e[k]]
            raise AssertionError(f"{what}: part {i} differs at {len(bad)} coefficients, first {bad[:4]}: "
                                 f"got {g[bad[0]]} want {e[bad[0]]}")
