"""CPU oracle for the FHE-SI ciphertext-arithmetic hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker.

What this is
------------
An exact-integer restatement of the reference's arithmetic (dwu4/fhe-si), function by
function, with Python ``int`` standing in for NTL ``ZZ`` and dense coefficient lists for
``ZZX``.  Each function cites the reference ``file:line`` it follows.  By SURVEY.md §0.3 the
coefficient-domain results of the reference do not depend on its prime chain: DoubleCRT is
only an exact multiplication engine for Z[X]/Phi_m, sized so nothing wraps
(FHEContext.cpp:83-85).  The oracle therefore multiplies polynomials exactly (Kronecker
substitution on big integers) and asserts that every value the reference would hold in
DoubleCRT form is below P/2 (P = product of the reference chain), i.e. that the reference
itself would not have wrapped.

Parity status: pinned against the reference's own sources, with NTL substituted.  The
reference ships no golden vectors or known-answer tests and NTL/GMP are
absent here, so oracle/build_ref.py compiles the reference's library sources -- unmodified,
where they lie -- against oracle/ntl_compat (our stand-in for the NTL interface they use)
into oracle/_ref.  tests/golden/ref_golden.json holds what that build writes for a seeded
scenario at every BASELINE configuration, and tests/test_oracle.py requires this oracle to
reproduce every byte (context, ciphertexts, add, tensor + ScaleDown, mult + relin, decrypt,
square, scalar, automorphism, DoubleCRT key rows).  Not pinned: NTL's own big-integer code
and both random streams (NTL's, libc rand()), which the stand-in replaces.  Also: (i) the
reference's self-consistency identities from Test_AddMul.cpp:84-86 over many seeds,
(ii) chain independence, (iii) an independent second implementation of the reference's
*algorithm* (oracle/ref_restate.c), (iv) serialization round trips.

Randomness: the reference draws from NTL's PRNG / lrand48 (= rand(), NumbTh.h:32-35) / libm
Box-Muller; the streams themselves are not pinned.  ``Rng`` below is a
SplitMix64 counter stream that the oracle, the tests, the C++ host layer and oracle/_ref
share, and every sampler follows the reference's order and number of draws.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

Poly = List[int]  # dense, low-to-high, length == phi(m) unless stated otherwise

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# RNG shared by oracle, tests and the C++ host layer (fhe-si_b200/host/ntl_shim.h)
# --------------------------------------------------------------------------------------
class Rng:
    """SplitMix64 stream.  Replaces NTL SetSeed/RandomBnd and srand48/lrand48
    (Test_AddMul.cpp:15-16); the stream itself is ours."""

    def __init__(self, seed: int):
        self.state = seed & MASK64

    def next64(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & MASK64
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)

    def random_bits(self, nbits: int) -> int:
        if nbits <= 0:
            return 0
        words = (nbits + 63) // 64
        v = 0
        for i in range(words):
            v |= self.next64() << (64 * i)
        return v & ((1 << nbits) - 1)

    def rand31(self) -> int:
        """libc rand() stand-in: 31 bits of the same stream.  The reference's sampleHWt draws
        from lrand48(), which NumbTh.h:32-35 maps to rand()."""
        return self.next64() >> 33

    def random_bnd(self, n: int) -> int:
        """Uniform in [0, n) -- NTL RandomBnd semantics, our stream."""
        if n <= 1:
            return 0
        nbits = (n - 1).bit_length()
        while True:
            v = self.random_bits(nbits)
            if v < n:
                return v


# --------------------------------------------------------------------------------------
# Number theory (NumbTh.cpp, PAlgebra.cpp)
# --------------------------------------------------------------------------------------
_MR_BASES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n: int) -> bool:
    """Deterministic Miller-Rabin for n < 3.3e24; stands in for NTL ProbPrime
    (FHEContext.cpp:34,108)."""
    if n < 2:
        return False
    for p in _MR_BASES:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _MR_BASES:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def factorize(n: int) -> List[int]:
    fs, d = [], 2
    while d * d <= n:
        if n % d == 0:
            fs.append(d)
            while n % d == 0:
                n //= d
        d += 1
    if n > 1:
        fs.append(n)
    return fs


def units_of(m: int) -> List[int]:
    """Z_m^* in increasing order; PAlgebra.cpp:52-53 (zmsIdx)."""
    return [i for i in range(m) if math.gcd(i, m) == 1]


def cyclotomic(m: int) -> Poly:
    """Phi_m(X) as a dense coefficient list of length phi(m)+1; NumbTh.cpp:142-159."""
    def polydiv_exact(a: Poly, b: Poly) -> Poly:
        a = a[:]
        out = [0] * (len(a) - len(b) + 1)
        for i in range(len(out) - 1, -1, -1):
            c = a[i + len(b) - 1] // b[-1]
            out[i] = c
            for j, bj in enumerate(b):
                a[i + j] -= c * bj
        assert not any(a)
        return out

    phi = {1: [-1, 1]}
    for d in range(2, m + 1):
        if m % d:
            continue
        num = [-1] + [0] * (d - 1) + [1]
        for e in range(1, d):
            if d % e == 0:
                num = polydiv_exact(num, phi[e])
        phi[d] = num
    return phi[m]


def root_of_unity_2m(p: int, m: int) -> int:
    """A primitive 2m-th root of unity mod p (needs p == 1 mod 2m).  The reference picks a
    random one (CModulus.cpp:66-76, NumbTh FindPrimitiveRoot); ours is deterministic --
    x^((p-1)/2m) for the smallest x = 2, 3, ... that gives exact order 2m -- and recorded
    in the exported context (FHEContext.cpp:52-59)."""
    e = 2 * m
    assert (p - 1) % e == 0
    fs = factorize(e)
    x = 2
    while True:
        r = pow(x, (p - 1) // e, p)
        if all(pow(r, e // f, p) != 1 for f in fs):
            return r
        x += 1


def find_primitive_root(rng: "Rng", q: int, e: int) -> int:
    """FindPrimRootT, NumbTh.cpp:84-118, as Cmod::privateInit calls it for e = 2m
    (CModulus.cpp:66-76): random s, root = s^(phi(q)/e), accepted when its order is exactly e.
    The draws come from the caller's stream, as the reference takes them from NTL's."""
    assert (q - 1) % e == 0
    fs = factorize(e)
    exp = (q - 1) // e
    for _ in range(1000):
        s = rng.random_bnd(q)                       # :103 random(s)
        root = pow(s, exp, q)                       # :104
        if pow(root, e, q) != 1:                    # :105-106 (s = 0)
            continue
        if all(pow(root, e // f, q) != 1 for f in fs):  # :108-113
            return root
    raise RuntimeError("FindPrimitiveRoot(): gave up after 1000 trials")


def add_primes_by_size(m: int, total_size: float, start_bits: int = 60) -> List[int]:
    """Prime chain builder; literal restatement of AddPrimesBySize, FHEContext.cpp:88-115.
    ``start_bits`` is NTL_SP_NBITS (FHEContext.cpp:92; 60 on modern NTL, 50 on 2013-era)."""
    chain: List[int] = []
    p = (1 << start_bits) - 1
    two_m = 2 * m
    p -= p % two_m
    p += two_m + 1
    last = False
    left = total_size
    while left > 0.0:
        if left < math.log(float(p)) and not last:
            last = True
            p = int(math.ceil(math.exp(left)))
            p -= (p % two_m) - 1
            two_m = -two_m
        while True:
            p -= two_m
            if is_prime(p):
                break
        if p not in chain:
            assert p % (2 * m) == 1  # AddPrime invariant, FHEContext.cpp:34
            chain.append(p)
            left -= math.log(float(p))
    return chain


# --------------------------------------------------------------------------------------
# Reduce (Util.cpp:3-33)
# --------------------------------------------------------------------------------------
def reduce_q(v: int, logq: int, positive: bool = False) -> int:
    """v mod 2^logq into [-q/2, q/2) (or [0, q) if positive); Util.cpp:3-26.
    NTL '>>' on a negative ZZ shifts the magnitude and keeps the sign (Util.cpp:16); the
    net effect of lines 15-19 is the non-negative residue, which Python's & gives directly."""
    q = 1 << logq
    v &= q - 1
    if not positive:
        s = 1 << (logq - 1)
        v = (v ^ s) - s
    return v


def reduce_poly(a: Poly, logq: int, positive: bool = False) -> Poly:
    """ReduceCoefficients, Util.cpp:28-33."""
    return [reduce_q(c, logq, positive) for c in a]


# --------------------------------------------------------------------------------------
# Exact arithmetic in Z[X]/Phi_m(X)
# --------------------------------------------------------------------------------------
def _kron_pack(a: Sequence[int], bits: int) -> int:
    """sum a_i 2^(bits*i) for signed a_i."""
    nb = bits // 8
    half = 1 << (bits - 1)
    buf = bytearray(nb * len(a))
    for i, c in enumerate(a):
        buf[i * nb:(i + 1) * nb] = (c + half).to_bytes(nb, "little")
    bias = int.from_bytes(bytes([0] * (nb - 1) + [0x80]) * len(a), "little")
    return int.from_bytes(buf, "little") - bias


def _kron_unpack(v: int, bits: int, count: int) -> List[int]:
    nb = bits // 8
    half = 1 << (bits - 1)
    bias = int.from_bytes(bytes([0] * (nb - 1) + [0x80]) * count, "little")
    v += bias
    assert v >= 0
    raw = v.to_bytes(nb * count + 1, "little")
    assert raw[nb * count] == 0
    return [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") - half for i in range(count)]


def poly_mul_full(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Plain integer polynomial product (NTL ZZX '*'), by Kronecker substitution."""
    if not a or not b:
        return []
    ma = max((abs(c) for c in a), default=0)
    mb = max((abs(c) for c in b), default=0)
    if ma == 0 or mb == 0:
        return [0] * (len(a) + len(b) - 1)
    bits = ma.bit_length() + mb.bit_length() + min(len(a), len(b)).bit_length() + 2
    bits = (bits + 7) // 8 * 8
    prod = _kron_pack(a, bits) * _kron_pack(b, bits)
    return _kron_unpack(prod, bits, len(a) + len(b) - 1)


@dataclass
class Ring:
    """Z[X]/Phi_m(X); PAlgebra.cpp:40-56 holds m, phi(m), Phi_m(X), the unit index table."""
    m: int
    phim: int = 0
    PhimX: Poly = field(default_factory=list)
    units: List[int] = field(default_factory=list)

    def __post_init__(self):
        self.units = units_of(self.m)
        self.phim = len(self.units)
        self.PhimX = cyclotomic(self.m)
        assert len(self.PhimX) == self.phim + 1 and self.PhimX[-1] == 1
        h = self.m // 2
        # m = 2p', p' odd prime: Phi_m(X) = Phi_p'(-X) = sum (-1)^i X^i  
        self.two_pprime = (self.m % 2 == 0 and h % 2 == 1 and is_prime(h))

    def rem(self, a: Sequence[int]) -> Poly:
        """a mod Phi_m(X), dense length phi(m); NTL rem (Ciphertext.cpp:30)."""
        n, m = self.phim, self.m
        a = list(a)
        if len(a) <= n:
            return a + [0] * (n - len(a))
        if self.two_pprime:
            h = m // 2  # X^h = -1 mod Phi_m
            v = [0] * h
            for i, c in enumerate(a):
                k, r = divmod(i, h)
                v[r] += -c if (k & 1) else c
            top = v[n]  # n == h-1
            return [v[i] - (top if i % 2 == 0 else -top) for i in range(n)]
        # general m: fold modulo X^m - 1, then schoolbook remainder by monic Phi_m
        v = [0] * m
        for i, c in enumerate(a):
            v[i % m] += c
        phi = self.PhimX
        for i in range(m - 1, n - 1, -1):
            c = v[i]
            if c:
                for j in range(n + 1):
                    v[i - n + j] -= c * phi[j]
        return v[:n]

    def mul(self, a: Sequence[int], b: Sequence[int]) -> Poly:
        return self.rem(poly_mul_full(a, b))

    def automorph(self, a: Sequence[int], k: int) -> Poly:
        """a(X) -> a(X^k) mod Phi_m; DoubleCRT.cpp:439-465 (row permutation
        new[j] = old[j*k mod m]) restated in coefficient form."""
        if math.gcd(k, self.m) != 1:
            raise ValueError("DoubleCRT::automorph: k not in Zm*")  # DoubleCRT.cpp:442-443
        v = [0] * self.m
        for i, c in enumerate(a):
            v[(i * k) % self.m] += c
        return self.rem(v)


# --------------------------------------------------------------------------------------
# Context (FHEContext.h:105-118, FHEContext.cpp:83-115)
# --------------------------------------------------------------------------------------
class Context:
    def __init__(self, m: int, logq: int, p: int, g: int, decomp_size: int = 3):
        """FHEcontext::Init, FHEContext.h:105-118."""
        self.m, self.logQ, self.p, self.g = m, logq, p, g
        self.stdev = 3.2
        self.ring = Ring(m)
        self.phim = self.ring.phim
        self.q = 1 << logq
        self.decompSize = decomp_size
        self.ndigits = (logq + 8 * decomp_size - 1) // (8 * decomp_size)
        self.primes: List[int] = []
        self.roots: List[int] = []
        self.xi = 1

    def setup_si(self, xi: int = 1, start_bits: int = 60, roots: Optional[List[int]] = None,
                 rng: Optional["Rng"] = None):
        """SetUpSIContext, FHEContext.cpp:83-85.  Each AddPrime constructs a Cmodulus, which
        draws its 2m-th root of unity from the random stream (CModulus.cpp:66-76): pass ``rng``
        to follow those draws (a client that has not called SetSeed yet is ``Rng(0)``); without
        it the roots are the deterministic ones of root_of_unity_2m.  Roots only matter for
        DoubleCRT rows and the exported context, never for coefficient-domain results."""
        total = (math.log(self.q) * 2 + math.log(self.p) + math.log(self.phim) * 2
                 + math.log(2) + math.log(xi))
        self.xi = xi
        self.primes = add_primes_by_size(self.m, total, start_bits)
        if roots:
            self.roots = list(roots)
        elif rng is not None:
            self.roots = [find_primitive_root(rng, q, 2 * self.m) for q in self.primes]
        else:
            self.roots = [root_of_unity_2m(q, self.m) for q in self.primes]
        return self

    def set_chain(self, primes: List[int], roots: List[int]):
        """ImportSIContext's AddPrime loop, FHEContext.cpp:74-80."""
        for q in primes:
            assert is_prime(q) and q % (2 * self.m) == 1
        self.primes, self.roots = list(primes), list(roots)
        return self

    @property
    def P(self) -> int:
        out = 1
        for q in self.primes:
            out *= q
        return out

    def check_no_wrap(self, a: Sequence[int], what: str = ""):
        """The reference holds this value mod P centred (DoubleCRT::toPoly,
        DoubleCRT.cpp:349-398); the restatement is exact only if it does not wrap."""
        if self.primes:
            half = self.P // 2
            for c in a:
                if not (-half <= c <= half):
                    raise OverflowError(f"reference chain would wrap in {what}")


# --------------------------------------------------------------------------------------
# DoubleCRT rows: reference-chain evaluation form (CModulus.cpp:90-132)
# --------------------------------------------------------------------------------------
def dcrt_rows(ctx: Context, a: Sequence[int]) -> List[List[int]]:
    """DoubleCRT(const ZZX&), DoubleCRT.cpp:244-257: row i, column j = a(zeta_i^{u_j}) mod
    p_i with zeta_i = root_i^2 a primitive m-th root and u_j the j-th unit
    (CModulus.cpp:90-107).  Direct evaluation (Horner); Bluestein is an implementation
    detail of the reference with the same result (bluestein.cpp:58-60)."""
    rows = []
    for q, root in zip(ctx.primes, ctx.roots):
        zeta = root * root % q
        ar = [c % q for c in a]
        row = []
        for u in ctx.ring.units:
            x = pow(zeta, u, q)
            acc = 0
            for c in reversed(ar):
                acc = (acc * x + c) % q
            row.append(acc)
        rows.append(row)
    return rows


def dcrt_to_poly(ctx: Context, rows: Sequence[Sequence[int]]) -> Poly:
    """DoubleCRT::toPoly, DoubleCRT.cpp:349-398 with Cmodulus::iFFT (CModulus.cpp:110-132)
    and intVecCRT (NumbTh.cpp:307-335): per-prime inverse transform, then incremental CRT,
    centred at every step."""
    m, n = ctx.m, ctx.phim
    units = ctx.ring.units
    phi = ctx.ring.PhimX
    acc: Optional[List[int]] = None
    prod = 1
    for q, root, row in zip(ctx.primes, ctx.roots, rows):
        zinv = pow(root * root % q, q - 2, q)
        minv = pow(m, q - 2, q)
        # m-point inverse DFT of the scattered row (zeros off the units)
        c = [0] * m
        for t in range(m):
            x = pow(zinv, t, q)
            s = 0
            for u, r in zip(units, row):
                s = (s + r * pow(x, u, q)) % q
            c[t] = s * minv % q
        # rem by Phi_m mod q (CModulus.cpp:127-129)
        for i in range(m - 1, n - 1, -1):
            ci = c[i]
            if ci:
                for j in range(n + 1):
                    c[i - n + j] = (c[i - n + j] - ci * phi[j]) % q
        c = c[:n]
        if acc is None:
            half = q // 2
            acc = [v - q if v > half else v for v in c]  # DoubleCRT.cpp:375-376
            prod = q
        else:
            pinv = pow(prod % q, q - 2, q)
            qhalf = q // 2
            for i in range(n):
                d = (c[i] - acc[i]) % q * pinv % q
                if d > qhalf:
                    d -= q  # NumbTh.cpp:318
                acc[i] += d * prod
            prod *= q
    return acc or [0] * n


# --------------------------------------------------------------------------------------
# Samplers (NumbTh.cpp:340-404, Util.cpp:49-56) on our stream
# --------------------------------------------------------------------------------------
def sample_hwt(rng: Rng, hwt: int, n: int) -> Poly:
    """sampleHWt, NumbTh.cpp:340-359."""
    a = [0] * n
    hwt = min(hwt, n)
    i = 0
    while i < hwt:
        u = rng.rand31() % n           # NumbTh.cpp:349
        if a[u] == 0:
            a[u] = (rng.rand31() & 2) - 1  # :351-352
            i += 1
    return a


def sample_gaussian(rng: Rng, n: int, stdev: float) -> Poly:
    """sampleGaussian (Box-Muller, rounded), NumbTh.cpp:377-404."""
    bignum = 0xFFFFFFF
    a = [0] * n
    for i in range(0, n, 2):
        r1 = (1 + rng.random_bnd(bignum)) / (float(bignum) + 1)
        r2 = (1 + rng.random_bnd(bignum)) / (float(bignum) + 1)
        theta = 2 * (4.0 * math.atan(1.0)) * r1
        rr = math.sqrt(-2.0 * math.log(r2)) * stdev
        assert rr < 8 * stdev
        a[i] = int(math.floor(rr * math.cos(theta) + 0.5))
        if i + 1 < n:
            a[i + 1] = int(math.floor(rr * math.sin(theta) + 0.5))
    return a


def sample_random(rng: Rng, modulus: int, n: int) -> Poly:
    """SampleRandom, Util.cpp:49-56: RandomBnd(modulus) - modulus/2."""
    off = modulus // 2
    return [rng.random_bnd(modulus) - off for _ in range(n)]


# --------------------------------------------------------------------------------------
# Keys (FHE-SI.cpp:42-62, 86-91, 153-239)
# --------------------------------------------------------------------------------------
@dataclass
class SecKey:
    """FHESISecKey: sKeys = (1, s); FHE-SI.cpp:86-91.  Held as exact integer polys."""
    ctx: Context
    s: List[Poly]

    @staticmethod
    def generate(ctx: Context, rng: Rng) -> "SecKey":
        one = [1] + [0] * (ctx.phim - 1)
        return SecKey(ctx, [one, sample_hwt(rng, 64, ctx.phim)])


@dataclass
class PubKey:
    """FHESIPubKey::Init, FHE-SI.cpp:42-62.  pk = (c0, c1) as coefficient polys
    (the reference stores DoubleCRT(c0), DoubleCRT(c1))."""
    ctx: Context
    pk: List[Poly]

    @staticmethod
    def generate(sk: SecKey, rng: Rng) -> "PubKey":
        ctx = sk.ctx
        c0 = sample_gaussian(rng, ctx.phim, ctx.stdev)            # :45
        c1 = sample_random(rng, ctx.q, ctx.phim)                   # :46
        tmp = ctx.ring.mul(sk.s[1], c1)                            # :48-53 (rem applied to the sum)
        c0 = [a + b for a, b in zip(c0, tmp)]
        c1 = [-c for c in c1]                                      # :54
        return PubKey(ctx, [reduce_poly(c0, ctx.logQ), reduce_poly(c1, ctx.logQ)])  # :56-57


@dataclass
class KeySwitch:
    """KeySwitchSI: keySwitchMatrix[0]=b, [1]=A, each of length ndigits * src.size();
    FHE-SI.cpp:153-209.  Entries are exact integer polys (the reference's DoubleCRT content)."""
    ctx: Context
    b: List[Poly]
    A: List[Poly]

    @staticmethod
    def init(ctx: Context, src: List[Poly], dst_t: Poly, rng: Rng) -> "KeySwitch":
        """KeySwitchSI::Init, FHE-SI.cpp:153-209."""
        D, n = ctx.ndigits, ctx.phim
        s_coeff = [list(x) for x in src]
        A, b = [], []
        for i in range(len(src)):
            for _ in range(D):
                poly = sample_random(rng, ctx.q, n)                 # :176
                A.append([-c for c in poly])                        # :178-180
                bc = ctx.ring.mul(poly, dst_t)                      # :182-185
                ctx.check_no_wrap(bc, "KeySwitchSI::Init b*t")
                err = sample_gaussian(rng, n, ctx.stdev)            # :187-188
                bc = [x + e + s for x, e, s in zip(bc, err, s_coeff[i])]  # :190-192
                s_coeff[i] = [c << (8 * ctx.decompSize) for c in s_coeff[i]]  # :194-196
                b.append(reduce_poly(bc, ctx.logQ))                 # :198-199
        return KeySwitch(ctx, b, A)

    @staticmethod
    def init_s2(sk: SecKey, rng: Rng) -> "KeySwitch":
        """InitS2, FHE-SI.cpp:211-227: src = (1, s, s^2), dst = s."""
        ctx = sk.ctx
        t = [sk.s[0], sk.s[1], ctx.ring.mul(sk.s[1], sk.s[1])]
        SecKey.generate(ctx, rng)  # :222 constructs a fresh FHESISecKey before overwriting it: draws
        return KeySwitch.init(ctx, t, sk.s[1], rng)

    @staticmethod
    def init_automorph(sk: SecKey, k: int, rng: Rng) -> "KeySwitch":
        """InitAutomorph, FHE-SI.cpp:229-239: src = (1, s(X^k)), dst = s."""
        ctx = sk.ctx
        src = [ctx.ring.automorph(x, k) for x in sk.s]
        SecKey.generate(ctx, rng)  # :233, as in InitS2: a throw-away key is sampled first
        return KeySwitch.init(ctx, src, sk.s[1], rng)


# --------------------------------------------------------------------------------------
# Ciphertext (Ciphertext.cpp)
# --------------------------------------------------------------------------------------
class Ciphertext:
    """Either ``parts`` (coefficient domain) or ``tprod`` (the reference's DoubleCRT
    tensor form, flag scaledUp; here exact integer polys); Ciphertext.h:46-97."""

    def __init__(self, ctx: Context, parts: Optional[List[Poly]] = None):
        self.ctx = ctx
        self.parts: List[Poly] = [list(p) for p in (parts or [])]
        self.tprod: List[Poly] = []
        self.scaled_up = False

    def copy(self) -> "Ciphertext":
        c = Ciphertext(self.ctx, self.parts)
        c.tprod = [list(t) for t in self.tprod]
        c.scaled_up = self.scaled_up
        return c

    def size(self) -> int:
        return len(self.tprod) if self.scaled_up else len(self.parts)

    # -- operator+=(const Ciphertext&), Ciphertext.cpp:123-145
    def add(self, other: "Ciphertext") -> "Ciphertext":
        assert self.scaled_up == other.scaled_up
        ctx = self.ctx
        if not self.scaled_up:
            k = min(len(self.parts), len(other.parts))
            for i in range(k):
                self.parts[i] = reduce_poly([a + b for a, b in zip(self.parts[i], other.parts[i])], ctx.logQ)
            self.parts += [list(p) for p in other.parts[k:]]
        else:
            k = min(len(self.tprod), len(other.tprod))
            for i in range(k):
                self.tprod[i] = [a + b for a, b in zip(self.tprod[i], other.tprod[i])]
                ctx.check_no_wrap(self.tprod[i], "tProd +=")
            self.tprod += [list(p) for p in other.tprod[k:]]
        return self

    # -- operator+=(const ZZX&), Ciphertext.cpp:147-161
    def add_plain(self, other: Sequence[int]) -> "Ciphertext":
        ctx = self.ctx
        sc = [(c << ctx.logQ) // ctx.p for c in other]
        sc = sc + [0] * (ctx.phim - len(sc))
        if not self.scaled_up:
            self.parts[0] = reduce_poly([a + b for a, b in zip(self.parts[0], sc)], ctx.logQ)
        else:
            self.tprod[0] = [a + b for a, b in zip(self.tprod[0], sc)]
        return self

    # -- operator*=(const Ciphertext&), Ciphertext.cpp:167-192
    def mul(self, other: "Ciphertext") -> "Ciphertext":
        ctx = self.ctx
        c1 = [[c * ctx.p for c in part] for part in self.parts]       # :170-172
        c2 = [list(part) for part in other.parts]                      # :174-176
        t = [[0] * ctx.phim for _ in range(len(c1) + len(c2) - 1)]
        for i, a in enumerate(c1):
            for j, b in enumerate(c2):
                pr = ctx.ring.mul(a, b)
                t[i + j] = [x + y for x, y in zip(t[i + j], pr)]       # :179-186
        for x in t:
            ctx.check_no_wrap(x, "Ciphertext *=")
        self.tprod, self.parts, self.scaled_up = t, [], True          # :188-189
        return self

    # -- ScaleDown, Ciphertext.cpp:194-218
    def scale_down(self) -> "Ciphertext":
        if not self.scaled_up:
            return self
        ctx = self.ctx
        q, q2 = ctx.q, 2 * ctx.q
        self.parts = [reduce_poly([(2 * c + q) // q2 for c in t], ctx.logQ) for t in self.tprod]
        self.tprod, self.scaled_up = [], False
        return self

    # -- ByteDecomp, Ciphertext.cpp:82-121: part-major, digit-minor, little-endian digits
    def byte_decomp(self) -> List[Poly]:
        ctx = self.ctx
        w = 8 * ctx.decompSize
        mask = (1 << w) - 1
        out: List[Poly] = []
        for part in self.parts:
            pos = reduce_poly(part, ctx.logQ, True)                    # :93
            for d in range(ctx.ndigits):
                out.append([(c >> (w * d)) & mask for c in pos])
        return out

    # -- operator*=(long), Ciphertext.cpp:233-244, :21-27
    def mul_scalar(self, l: int) -> "Ciphertext":
        ctx = self.ctx
        if not self.scaled_up:
            self.parts = [reduce_poly([c * l for c in part], ctx.logQ) for part in self.parts]
        else:
            self.tprod = [[c * l for c in t] for t in self.tprod]
            for t in self.tprod:
                ctx.check_no_wrap(t, "tProd *= long")
        return self

    # -- operator*=(const ZZX&), Ciphertext.cpp:246-258, :29-36
    def mul_plain(self, other: Sequence[int]) -> "Ciphertext":
        ctx = self.ctx
        if not self.scaled_up:
            self.parts = [reduce_poly(ctx.ring.mul(part, other), ctx.logQ) for part in self.parts]
        else:
            self.tprod = [ctx.ring.mul(t, other) for t in self.tprod]
            for t in self.tprod:
                ctx.check_no_wrap(t, "tProd *= ZZX")
        return self

    # -- operator>>=(long), Ciphertext.cpp:264-275, :54-59.  Output NOT reduced mod q.
    def automorph(self, k: int) -> "Ciphertext":
        ring = self.ctx.ring
        if not self.scaled_up:
            self.parts = [ring.automorph(p, k) for p in self.parts]
        else:
            self.tprod = [ring.automorph(t, k) for t in self.tprod]
        return self


def encrypt(pk: PubKey, msg: Sequence[int], r: Sequence[int], e: Sequence[Sequence[int]]) -> Ciphertext:
    """FHESIPubKey::Encrypt, FHE-SI.cpp:10-36, with the randomness made explicit:
    r = phi(m) bits (:14-17), e[i] = Gaussian poly per part (:24)."""
    ctx = pk.ctx
    parts = []
    for i in range(2):
        c = ctx.ring.mul(pk.pk[i], r)                                  # :27
        c = [x + ctx.p * y for x, y in zip(c, e[i])]                   # :24-25,28
        ctx.check_no_wrap(c, "Encrypt")
        parts.append(c)                                                # :29
    scale = ctx.q // ctx.p                                             # :31
    m = list(msg) + [0] * (ctx.phim - len(msg))
    parts[0] = [x + scale * y for x, y in zip(parts[0], m)]
    return Ciphertext(ctx, [reduce_poly(p, ctx.logQ) for p in parts])  # :33-35


def encrypt_rng(pk: PubKey, msg: Sequence[int], rng: Rng) -> Ciphertext:
    """Same draw order as FHE-SI.cpp:14-25: r first, then e[0], e[1]."""
    ctx = pk.ctx
    r = [rng.random_bnd(2) for _ in range(ctx.phim)]
    e = [sample_gaussian(rng, ctx.phim, ctx.stdev) for _ in range(2)]
    return encrypt(pk, msg, r, e)


def decrypt(sk: SecKey, ct: Ciphertext) -> Poly:
    """FHESISecKey::Decrypt, FHE-SI.cpp:93-119.  Uses parts 0..sKeys.size()-1 only."""
    ctx = sk.ctx
    z = [0] * ctx.phim
    for i in range(len(sk.s)):
        pr = ctx.ring.mul(ct.parts[i], sk.s[i])                        # :96-103
        z = [a + b for a, b in zip(z, pr)]
    ctx.check_no_wrap(z, "Decrypt")
    q, q2, p = ctx.q, 2 * ctx.q, ctx.p
    return [((2 * p * c + q) // q2) % p for c in z]                    # :111-118


def apply_key_switch(ks: KeySwitch, ct: Ciphertext) -> Ciphertext:
    """KeySwitchSI::ApplyKeySwitch, FHE-SI.cpp:241-260."""
    ctx = ks.ctx
    ct.scale_down()                                                    # :243
    digits = ct.byte_decomp()                                          # :244
    assert len(digits) == len(ks.b)
    new_parts = []
    for row in (ks.b, ks.A):                                           # :251-257
        acc = [0] * ctx.phim
        for k, d in zip(row, digits):
            pr = ctx.ring.mul(k, d)
            acc = [a + b for a, b in zip(acc, pr)]
        ctx.check_no_wrap(acc, "ApplyKeySwitch")
        new_parts.append(reduce_poly(acc, ctx.logQ))
    ct.parts = new_parts                                               # :259
    return ct


def mult_relin(ks: KeySwitch, a: Ciphertext, b: Ciphertext) -> Ciphertext:
    """The metric op: c = a; c *= b; ks.ApplyKeySwitch(c)  (Test_AddMul.cpp:59-66)."""
    c = a.copy()
    c.mul(b)
    return apply_key_switch(ks, c)


# --------------------------------------------------------------------------------------
# Serialization (Serialization.cpp:3-119, FHEContext.cpp:45-81); host-endian LP64
# --------------------------------------------------------------------------------------
def export_zz(v: int) -> bytes:
    """Serialization.cpp:3-13: u32 nBytes, 1-byte neg, little-endian magnitude."""
    mag = abs(v)
    nb = (mag.bit_length() + 7) // 8
    return struct.pack("<I?", nb, v < 0) + mag.to_bytes(nb, "little")


def import_zz(buf: bytes, off: int) -> Tuple[int, int]:
    nb, neg = struct.unpack_from("<I?", buf, off)
    off += 5
    v = int.from_bytes(buf[off:off + nb], "little")
    return (-v if neg else v), off + nb


def export_zzx(a: Sequence[int]) -> bytes:
    """Serialization.cpp:29-36: i32 degree (-1 for zero), then degree+1 ZZ."""
    deg = len(a) - 1
    while deg >= 0 and a[deg] == 0:
        deg -= 1
    return struct.pack("<i", deg) + b"".join(export_zz(a[i]) for i in range(deg + 1))


def import_zzx(buf: bytes, off: int, n: int) -> Tuple[Poly, int]:
    (deg,) = struct.unpack_from("<i", buf, off)
    off += 4
    a = [0] * max(n, deg + 1)
    for i in range(deg + 1):
        a[i], off = import_zz(buf, off)
    return a, off


def export_vec_long(v: Sequence[int]) -> bytes:
    """Serialization.cpp:83-89."""
    return struct.pack("<I", len(v)) + b"".join(struct.pack("<q", x) for x in v)


def export_dcrt(rows: Sequence[Sequence[int]]) -> bytes:
    """Serialization.cpp:56-65: u32 card, then per row (i64 index, vec_long)."""
    out = struct.pack("<I", len(rows))
    for i, r in enumerate(rows):
        out += struct.pack("<q", i) + export_vec_long(r)
    return out


def import_dcrt(buf: bytes, off: int) -> Tuple[List[List[int]], int]:
    (card,) = struct.unpack_from("<I", buf, off)
    off += 4
    rows: List[List[int]] = []
    for _ in range(card):
        (_idx,) = struct.unpack_from("<q", buf, off)
        off += 8
        (ln,) = struct.unpack_from("<I", buf, off)
        off += 4
        rows.append(list(struct.unpack_from(f"<{ln}q", buf, off)))
        off += 8 * ln
    return rows, off


def export_ciphertext(ct: Ciphertext) -> bytes:
    """Serialization.cpp:109-114: ScaleDown on a copy, then vector<CiphertextPart>."""
    c = ct.copy().scale_down()
    return struct.pack("<I", len(c.parts)) + b"".join(export_zzx(p) for p in c.parts)


def import_ciphertext(ctx: Context, buf: bytes, off: int = 0) -> Tuple[Ciphertext, int]:
    (cnt,) = struct.unpack_from("<I", buf, off)
    off += 4
    parts = []
    for _ in range(cnt):
        a, off = import_zzx(buf, off, ctx.phim)
        parts.append(a)
    return Ciphertext(ctx, parts), off


def export_context(ctx: Context) -> bytes:
    """ExportSIContext, FHEContext.cpp:45-60."""
    out = struct.pack("<II", ctx.m, ctx.logQ) + export_zz(ctx.p) + struct.pack("<II", ctx.g, ctx.decompSize)
    out += struct.pack("<I", len(ctx.primes))
    for q, r in zip(ctx.primes, ctx.roots):
        out += struct.pack("<qq", q, r)
    return out


def import_context(buf: bytes) -> Context:
    """ImportSIContext, FHEContext.cpp:62-81."""
    m, logq = struct.unpack_from("<II", buf, 0)
    p, off = import_zz(buf, 8)
    g, ds, cnt = struct.unpack_from("<III", buf, off)
    off += 12
    ctx = Context(m, logq, p, g, ds)
    primes, roots = [], []
    for _ in range(cnt):
        q, r = struct.unpack_from("<qq", buf, off)
        off += 16
        primes.append(q)
        roots.append(r)
    return ctx.set_chain(primes, roots)


# --------------------------------------------------------------------------------------
# Fixed-width two's-complement coefficient packing used at the C-ABI (include/fhesi.h)
# --------------------------------------------------------------------------------------
def words_per_coeff(logq: int) -> int:
    return (logq + 31) // 32


def pack_poly_words(a: Sequence[int], logq: int):
    """-> numpy uint32 [len(a)][W], little-endian words, two's complement of the centred
    value sign-extended to 32*W bits."""
    import numpy as np
    W = words_per_coeff(logq)
    mod = 1 << (32 * W)
    raw = b"".join((c % mod).to_bytes(4 * W, "little") for c in a)
    return np.frombuffer(raw, dtype="<u4").reshape(len(a), W).copy()


def unpack_poly_words(arr) -> Poly:
    import numpy as np
    arr = np.ascontiguousarray(arr, dtype="<u4")
    n, W = arr.shape
    raw = arr.tobytes()
    half, mod = 1 << (32 * W - 1), 1 << (32 * W)
    out = []
    for i in range(n):
        v = int.from_bytes(raw[i * 4 * W:(i + 1) * 4 * W], "little")
        out.append(v - mod if v >= half else v)
    return out
