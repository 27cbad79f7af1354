// fhesi_host.cpp -- bodies of the reference-named classes (fhesi_host.h) over the C ABI.
// Compiled with g++ into libfhesi_host.so and linked against libfhesi_b200.so (or, in GPU-less
// CI, against the kernel-logic emulator build of the same sources, tests/emu).
#include "fhesi_host.h"

#include <sched.h>

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

#include <map>

FHEcontext *activeContext = NULL;

static void Check(int rc, const char *what) {
  if (rc) {
    std::cerr << what << ": " << fhesi_last_error() << std::endl;
    std::abort();
  }
}

// ------------------------------------------------------------------------------- number theory
ZZX Cyclotomic(long N) {  // NumbTh.cpp:142-159
  std::map<long, std::vector<long long>> phi;
  phi[1] = {-1, 1};
  for (long d = 2; d <= N; ++d) {
    if (N % d) continue;
    std::vector<long long> num(d + 1, 0);
    num[0] = -1, num[d] = 1;
    for (long e = 1; e < d; ++e) {
      if (d % e) continue;
      const std::vector<long long> &b = phi[e];
      std::vector<long long> q(num.size() - b.size() + 1, 0);
      for (long i = (long)q.size() - 1; i >= 0; --i) {
        long long c = num[i + b.size() - 1] / b.back();
        q[i] = c;
        for (size_t j = 0; j < b.size(); ++j) num[i + j] -= c * b[j];
      }
      num = q;
    }
    phi[d] = num;
  }
  ZZX r;
  const std::vector<long long> &c = phi[N];
  r.rep.v.resize(c.size());
  for (size_t i = 0; i < c.size(); ++i) r.rep.v[i] = ZZ((long)c[i]);
  r.normalize();
  return r;
}

void PAlgebra::init(unsigned mm, unsigned gen) {
  if (m == mm) return;
  m = mm;
  g = gen;
  zmsIdx.assign(m, -1);
  long idx = 0;
  for (unsigned i = 0; i < m; i++)
    if (GCD(i, m) == 1) zmsIdx[i] = idx++;
  phim = idx;
  Phi_mX = Cyclotomic(m);
}

// a mod Phi_m for m = 2h (h odd prime): X^h = -1, Phi_m = sum (-1)^i X^i; generic otherwise
static void RemPhim(ZZX &a, const PAlgebra &zms) {
  const unsigned m = zms.M(), n = zms.phiM(), h = m / 2;
  if (deg(a) < (long)n) return;
  if (m % 2 == 0 && n == h - 1) {
    std::vector<ZZ> v(h);
    for (long i = 0; i <= deg(a); ++i) {
      if (a.rep.v[i].is_zero()) continue;
      if ((i / h) & 1) v[i % h] -= a.rep.v[i];
      else v[i % h] += a.rep.v[i];
    }
    ZZ top = v[n];
    a.rep.v.assign(n, ZZ());
    for (unsigned i = 0; i < n; ++i) a.rep.v[i] = (i & 1) ? v[i] + top : v[i] - top;
    a.normalize();
  } else {
    rem(a, a, zms.PhimX());
  }
}
// a * b mod Phi_m.  Secret keys are ternary and sparse (Hamming weight 64), so the products key
// generation needs are shift-and-add; they run on fixed-width two's-complement word arrays
// (no allocation per coefficient), which keeps KeySwitchSI::Init at set-up-time cost.
// Run fn(0..total-1) over the host's cores (at most 16 workers); fn must not touch shared state.
template <class F>
static void ParallelFor(size_t total, F fn) {
  static const unsigned cap = [] {
    const char *e = getenv("FHESIH_THREADS");  // 1 = serial
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) hw = std::min<unsigned>(hw, std::max(1, CPU_COUNT(&set)));
    return e && atoi(e) > 0 ? (unsigned)atoi(e) : std::min(hw, 16u);
  }();
  unsigned workers = std::min<unsigned>((unsigned)total, cap);
  if (workers <= 1) {
    for (size_t i = 0; i < total; ++i) fn(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (unsigned w = 0; w < workers; ++w)
    pool.emplace_back([&] {
      for (size_t i; (i = next.fetch_add(1)) < total;) fn(i);
    });
  for (auto &th : pool) th.join();
}

// Inner loops of the host-side samplers/packers, cloned for AVX2 (resolved at load time: the
// library is built in one container and runs on another host).
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define FHESI_SIMD_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define FHESI_SIMD_CLONES
#endif
FHESI_SIMD_CLONES static void MacRow32(uint64_t *acc, const uint32_t *b, uint32_t c, long n) {
  for (long k = 0; k < n; ++k) acc[k] += (uint64_t)b[k] * c;
}
FHESI_SIMD_CLONES static void AddRow64(int64_t *d, const int64_t *w, size_t n) {
  for (size_t t = 0; t < n; ++t) d[t] += w[t];
}
FHESI_SIMD_CLONES static void SubRow64(int64_t *d, const int64_t *w, size_t n) {
  for (size_t t = 0; t < n; ++t) d[t] -= w[t];
}

static ZZX MulModPhim(const ZZX &a, const ZZX &b, const PAlgebra &zms) {
  auto tiny = [](const ZZX &x) {
    for (auto &c : x.rep.v)
      if (c.mag.size() > 1 || (c.mag.size() == 1 && c.mag[0] > 1)) return false;
    return true;
  };
  const ZZX *big = &a, *sm = &b;
  if (!tiny(b) && tiny(a)) big = &b, sm = &a;
  ZZX r;
  if (tiny(*sm) && !big->rep.v.empty() && !sm->rep.v.empty()) {
    // one operand has coefficients in {-1,0,1} (secret keys, encryption randomness): the product is
    // a few hundred shifted signed adds.  Limbs are accumulated carry-save in int64 lanes (a sum of
    // < 2^31 signed 32-bit limbs cannot overflow), so the inner loop is a plain vector add; carries
    // are propagated once per output coefficient.
    size_t Wd = 1;
    for (auto &c : big->rep.v) Wd = std::max(Wd, c.mag.size());
    const size_t nb = big->rep.v.size(), len = nb + sm->rep.v.size() - 1;
    std::vector<int64_t> src(nb * Wd, 0), acc(len * Wd, 0);
    for (size_t i = 0; i < nb; ++i) {
      const ZZ &c = big->rep.v[i];
      for (size_t k = 0; k < c.mag.size(); ++k) src[i * Wd + k] = c.neg ? -(int64_t)c.mag[k] : (int64_t)c.mag[k];
    }
    const size_t span = nb * Wd;
    for (size_t j = 0; j < sm->rep.v.size(); ++j) {
      const ZZ &c = sm->rep.v[j];
      if (c.is_zero()) continue;
      int64_t *d = &acc[j * Wd];
      const int64_t *w = src.data();
      if (c.neg) SubRow64(d, w, span);
      else AddRow64(d, w, span);
    }
    r.rep.v.resize(len);
    std::vector<uint32_t> limbs(Wd + 2);
    for (size_t i = 0; i < len; ++i) {  // signed carry propagation -> two's complement -> sign + magnitude
      const int64_t *d = &acc[i * Wd];
      int64_t cy = 0;
      for (size_t k = 0; k < Wd; ++k) {
        int64_t v = d[k] + cy;
        limbs[k] = (uint32_t)v;
        cy = v >> 32;
      }
      limbs[Wd] = (uint32_t)cy;
      limbs[Wd + 1] = (uint32_t)(cy >> 32);
      const bool neg = cy < 0;
      if (neg) {
        uint64_t c2 = 1;
        for (size_t k = 0; k < Wd + 2; ++k) {
          c2 += (uint32_t)~limbs[k];
          limbs[k] = (uint32_t)c2;
          c2 >>= 32;
        }
      }
      ZZ &c = r.rep.v[i];
      c.mag.assign(limbs.begin(), limbs.end());
      c.neg = neg;
      c.trim();
    }
    r.normalize();
  } else {
    r = a * b;
  }
  RemPhim(r, zms);
  return r;
}

// ------------------------------------------------------------------------------- Util / samplers
void Reduce(ZZ &val, unsigned logQ, bool positive) {  // Util.cpp:3-26, q = 2^logQ: masks, no division
  const size_t words = (logQ + 31) / 32, tb = logQ & 31;
  // |val| mod q
  if (val.mag.size() > words) val.mag.resize(words);
  if (tb && val.mag.size() == words) val.mag[words - 1] &= (1u << tb) - 1;
  bool neg = val.neg;
  val.neg = false;
  val.trim();
  if (neg && !val.mag.empty()) {  // q - r
    val.mag.resize(words, 0);
    uint64_t cy = 1;
    for (size_t k = 0; k < words; ++k) {
      cy += (uint32_t)~val.mag[k];
      val.mag[k] = (uint32_t)cy;
      cy >>= 32;
    }
    if (tb) val.mag[words - 1] &= (1u << tb) - 1;
    val.trim();
  }
  if (!positive && val.bits() == logQ) {  // >= q/2: subtract q
    val.mag.resize(words, 0);
    uint64_t cy = 1;
    for (size_t k = 0; k < words; ++k) {
      cy += (uint32_t)~val.mag[k];
      val.mag[k] = (uint32_t)cy;
      cy >>= 32;
    }
    if (tb) val.mag[words - 1] &= (1u << tb) - 1;
    val.neg = true;
    val.trim();
  }
}
void ReduceCoefficients(ZZX &poly, unsigned logQ, bool positive) {
  for (long i = 0; i <= deg(poly); i++) Reduce(poly.rep[i], logQ, positive);
  poly.normalize();
}
void SampleRandom(ZZX &poly, const ZZ &modulus, unsigned d) {  // Util.cpp:49-56
  ZZ offset = modulus / 2;
  poly.rep.v.assign(d, ZZ());
  const size_t k = modulus.bits() - 1;  // modulus == 2^k ?
  bool pow2 = !modulus.is_zero() && k >= 2 && (modulus.mag.back() & (modulus.mag.back() - 1)) == 0;
  for (size_t i = 0; pow2 && i + 1 < modulus.mag.size(); ++i) pow2 = modulus.mag[i] == 0;
  if (pow2) {
    // RandomBnd(2^k) is RandomBits(k): ceil(k/64) words of the stream, low k bits kept, never
    // rejected.  v - 2^(k-1) is formed on the limbs directly: same draws, same values, no ZZ temporaries.
    const size_t words64 = (k + 63) / 64, limbs = (k + 31) / 32, top = (k - 1) / 32;
    const uint32_t topbit = 1u << ((k - 1) % 32);
    uint32_t w[2 * 64];
    if (words64 <= 64) {
      for (unsigned i = 0; i < d; i++) {
        for (size_t t = 0; t < words64; ++t) {
          uint64_t x = GlobalRandomStream().next64();
          w[2 * t] = (uint32_t)x, w[2 * t + 1] = (uint32_t)(x >> 32);
        }
        if (k % 32) w[limbs - 1] &= (1u << (k % 32)) - 1;
        ZZ &c = poly.rep.v[i];
        if (w[top] & topbit) {  // v >= q/2: v - q/2 clears the top bit
          w[top] &= ~topbit;
          c.mag.assign(w, w + limbs);
          c.neg = false;
        } else {  // -(q/2 - v)
          uint64_t borrow = 0;
          for (size_t t = 0; t < limbs; ++t) {
            uint64_t sub = (uint64_t)w[t] + borrow, from = t == top ? topbit : 0u;
            borrow = from < sub;
            w[t] = (uint32_t)(from - sub);
          }
          c.mag.assign(w, w + limbs);
          c.neg = true;
        }
        c.trim();
      }
      poly.normalize();
      return;
    }
  }
  for (unsigned i = 0; i < d; i++) poly.rep.v[i] = RandomBnd(modulus) - offset;
  poly.normalize();
}
void sampleHWt(ZZX &poly, long Hwt, long n) {  // NumbTh.cpp:340-359
  if (n <= 0) n = deg(poly) + 1;
  if (n <= 0) return;
  poly.rep.v.assign(n, ZZ());
  if (Hwt > n) Hwt = n;
  long i = 0;
  while (i < Hwt) {
    long u = RandomLibc31() % n;  // :349 -- lrand48(), i.e. rand() (NumbTh.h:32-35)
    if (poly.rep.v[u].is_zero()) {
      long b = (RandomLibc31() & 2) - 1;  // :351-352
      poly.rep.v[u] = ZZ(b);
      i++;
    }
  }
  poly.normalize();
}
void sampleSmall(ZZX &poly, long n) {  // NumbTh.cpp:361-375
  if (n <= 0) n = deg(poly) + 1;
  if (n <= 0) return;
  poly.rep.v.assign(n, ZZ());
  for (long i = 0; i < n; i++) {
    long u = RandomLibc31();  // :367
    if (u & 1) poly.rep.v[i] = ZZ((long)(u & 2) - 1);
  }
  poly.normalize();
}
// Box-Muller as NumbTh.cpp:377-404 draws it: two uniforms per pair of outputs, rounded to nearest.  Split in
// two so that a caller with many polynomials can take the draws from the (sequential) stream first and do
// the floating-point part on all cores afterwards: raw holds 2 * ceil(n / 2) uniforms in draw order.
static void SampleGaussianRaw(uint32_t *raw, long n) {
  static long const bignum = 0xfffffff;
  for (long i = 0; i < n; i += 2) {
    raw[i] = (uint32_t)RandomBnd(bignum);
    raw[i + 1] = (uint32_t)RandomBnd(bignum);
  }
}
// The same draws from an explicit stream, one 64-bit word each and NO retry: returns false if a word would have been
// rejected by RandomBnd (probability 2^-28 per draw) -- the caller then repeats the whole set-up sequentially
static bool SampleGaussianRawNoRetry(RandomStream &rs, uint32_t *raw, long n) {
  static const uint64_t bignum = 0xfffffff;  // RandomBnd(2^28 - 1): 28 bits, accepted when < bignum
  bool ok = true;
  for (long i = 0; i < n; i += 2) {
    const uint64_t a = rs.next64() & bignum, b = rs.next64() & bignum;
    ok &= a < bignum && b < bignum;
    raw[i] = (uint32_t)a, raw[i + 1] = (uint32_t)b;
  }
  return ok;
}
static void GaussianFromRaw(int32_t *out, const uint32_t *raw, long n, double stdev) {
  static double const Pi = 4.0 * atan(1.0);
  static long const bignum = 0xfffffff;
  for (long i = 0; i < n; i += 2) {
    double r1 = (1 + (long)raw[i]) / ((double)bignum + 1);
    double r2 = (1 + (long)raw[i + 1]) / ((double)bignum + 1);
    double theta = 2 * Pi * r1;
    double rr = sqrt(-2.0 * std::log(r2)) * stdev;
    assert(rr < 8 * stdev);
    out[i] = (int32_t)floor(rr * cos(theta) + 0.5);
    if (i + 1 < n) out[i + 1] = (int32_t)floor(rr * sin(theta) + 0.5);
  }
}
static void SampleGaussianInts(int32_t *out, long n, double stdev) {
  std::vector<uint32_t> raw((size_t)n + 1);
  SampleGaussianRaw(raw.data(), n);
  GaussianFromRaw(out, raw.data(), n, stdev);
}
void sampleGaussian(ZZX &poly, long n, double stdev) {  // NumbTh.cpp:377-404
  if (n <= 0) n = deg(poly) + 1;
  if (n <= 0) return;
  std::vector<int32_t> g(n);
  SampleGaussianInts(g.data(), n, stdev);
  poly.rep.v.assign(n, ZZ());
  for (long i = 0; i < n; ++i) poly.rep.v[i] = ZZ((long)g[i]);
  poly.normalize();
}

// ------------------------------------------------------------------------------- PlaintextSpace
void PlaintextSpace::Init(const ZZX &PhiX, const ZZ &pp, unsigned gen) {
  generator = gen;
  Init(PhiX, pp);
}
void PlaintextSpace::Init(const ZZX &PhiX, const ZZ &pp) {
  ZZ_p::init(pp);
  p = pp;
  const long P = to_long(pp), n = deg(PhiX);
  totalSlots = usableSlots = 0;
  roots.clear();
  basis.clear();
  if (m == 0 || !ProbPrime(P) || (P - 1) % m != 0) return;  // no linear slots: packing unavailable
  // primitive m-th root of unity mod p
  std::vector<long> fs;
  for (long t = m, d = 2; t > 1; ++d)
    if (t % d == 0) {
      fs.push_back(d);
      while (t % d == 0) t /= d;
    }
  long rho = 0;
  for (long x = 2; !rho; ++x) {
    long r = PowerMod(x, (P - 1) / m, P);
    bool ok = true;
    for (long f : fs)
      if (PowerMod(r, m / f, P) == 1) ok = false;
    if (ok) rho = r;
  }
  // slot j <-> rho^(g^j): the order PlaintextSpace::ReorderSlots (PlaintextSpace.cpp:87-110)
  // produces, up to its arbitrary starting factor
  long e = 1;
  for (long j = 0; j < n; ++j) {
    roots.push_back(PowerMod(rho, e, P));
    e = MulMod(e, generator % m, m);
    if (e == 1 && j + 1 < n) break;
  }
  if ((long)roots.size() != n) {
    roots.clear();
    return;  // generator does not generate Z_m^* (the reference asserts here)
  }
  totalSlots = n;
  usableSlots = 1;
  for (unsigned t = totalSlots; t > 1; t >>= 1) usableSlots <<= 1;
  // the CRT idempotents are built on first use (EnsureBasis): key generation and pure ciphertext
  // arithmetic never embed a plaintext
  phiModP.assign(n + 1, 0);
  for (long i = 0; i <= n; ++i) phiModP[i] = to_long(coeff(PhiX, i) % pp);
  basis.clear();
  basis32.clear();
}
void PlaintextSpace::EnsureBasis() const {  // basis_j = Phi/(X - r_j) / Phi'(r_j)
  if (!basis.empty() || !totalSlots) return;
  const long P = to_long(p), n = totalSlots;
  const std::vector<long> &phi = phiModP;
  basis.assign(n, std::vector<long>(n));
  for (long j = 0; j < n; ++j) {
    std::vector<long> &b = basis[j];
    long r = roots[j], carry = phi[n];
    for (long i = n - 1; i >= 0; --i) {  // synthetic division by (X - r)
      b[i] = carry;
      carry = AddMod(phi[i], MulMod(carry, r, P), P);
    }
    long d = 0;
    for (long i = n - 1; i >= 0; --i) d = AddMod(MulMod(d, r, P), b[i], P);
    long di = InvMod(d, P);
    for (long i = 0; i < n; ++i) b[i] = MulMod(b[i], di, P);
  }
  if (P < (1L << 26)) {
    basis32.resize((size_t)n * n);
    for (long j = 0; j < n; ++j)
      for (long i = 0; i < n; ++i) basis32[(size_t)j * n + i] = (uint32_t)basis[j][i];
  }
}
void PlaintextSpace::EmbedInSlots(ZZ_pX &embedded, const vector<ZZ_pX> &msgs, bool onlyUsable) const {
  if (!totalSlots) Error("PlaintextSpace: slots need a prime p = 1 mod m and a generator of Z_m^*");
  EnsureBasis();
  const long P = to_long(p), n = totalSlots;
  embedded.rep.v.assign(n, ZZ_p());
  auto slotValue = [&](const ZZ_pX &mi, unsigned i) -> long {
    // a slot only sees the message modulo its factor X - r_i
    return deg(mi) <= 0 ? (deg(mi) < 0 ? 0 : mi.rep.v[0].v) : eval(mi, ZZ_p(roots[i])).v;
  };
  unsigned msgInd = 0;
  if (P < (1L << 26)) {  // products < 2^52, at most 2^12 of them per sum: plain 64-bit lanes
    std::vector<uint64_t> acc(n, 0);
    for (unsigned i = 0; i < totalSlots && msgInd < msgs.size(); i++) {
      if (onlyUsable && i >= usableSlots) break;
      const uint32_t c = (uint32_t)slotValue(msgs[msgInd++], i);
      if (!c) continue;
      MacRow32(acc.data(), &basis32[(size_t)i * n], c, n);
    }
    for (long k = 0; k < n; ++k) embedded.rep.v[k].v = (long)(acc[k] % (uint64_t)P);
  } else {
    std::vector<unsigned __int128> acc(n, 0);
    for (unsigned i = 0; i < totalSlots && msgInd < msgs.size(); i++) {
      if (onlyUsable && i >= usableSlots) break;
      long c = slotValue(msgs[msgInd++], i);
      if (!c) continue;
      for (long k = 0; k < n; ++k) acc[k] += (unsigned __int128)basis[i][k] * c;
    }
    for (long k = 0; k < n; ++k) embedded.rep.v[k].v = (long)(acc[k] % (unsigned long)P);
  }
  embedded.normalize();
}
void PlaintextSpace::DecodeSlot(ZZ_pX &val, const ZZ_pX &msg, unsigned ind) const {
  if (!totalSlots) Error("PlaintextSpace: slots need a prime p = 1 mod m and a generator of Z_m^*");
  val = to_ZZ_pX(eval(msg, ZZ_p(roots[ind])));
}
void PlaintextSpace::DecodeSlots(vector<ZZ_pX> &msgBatch, const ZZ_pX &msg, bool onlyUsable) const {
  msgBatch.resize(totalSlots);
  for (unsigned i = 0; i < totalSlots; i++) {
    if (onlyUsable && i >= usableSlots) break;
    DecodeSlot(msgBatch[i], msg, i);
  }
}

// ------------------------------------------------------------------------------- FHEcontext
void FHEcontext::Init(unsigned m, unsigned lq, const ZZ &p, unsigned generator, unsigned ds) {
  stdev = 3.2;
  zMstar.init(m, generator);
  logQ = lq;
  modulusQ = ZZ(1L) << (long)lq;
  decompSize = ds;
  ndigits = (lq + 8 * ds - 1) / (8 * ds);
  ptxtSpace.m = m;
  ptxtSpace.Init(zMstar.PhimX(), p, generator);
}
FHEcontext::~FHEcontext() {
  if (dev) fhesi_ctx_destroy(dev);
  if (activeContext == this) activeContext = NULL;
}
// FindPrimRootT (NumbTh.cpp:84-118) as Cmod::privateInit calls it for e = 2m (CModulus.cpp:66-76):
// random s from the global stream, root = s^(phi(q)/e), accepted when its order is exactly e
static long FindPrimitiveRoot2m(long p, long m) {
  long e = 2 * m;
  std::vector<long> fs;
  for (long t = e, d = 2; t > 1; ++d)
    if (t % d == 0) {
      fs.push_back(d);
      while (t % d == 0) t /= d;
    }
  for (int it = 0; it < 1000; ++it) {
    long s = RandomBnd(p);
    long r = PowerMod(s, (p - 1) / e, p);
    if (PowerMod(r, e, p) != 1) continue;
    bool ok = true;
    for (long f : fs)
      if (PowerMod(r, e / f, p) == 1) ok = false;
    if (ok) return r;
  }
  Error("FindPrimitiveRoot(): gave up after 1000 trials");
}
void FHEcontext::AddPrime(long p, bool special, long root) {  // FHEContext.cpp:30-43
  long twoM = 2 * zMstar.M();
  assert(ProbPrime(p) && p % twoM == 1 && !inChain(p));
  CmodulusInfo mo;
  mo.q = p;
  mo.root = root ? root : FindPrimitiveRoot2m(p, zMstar.M());
  long i = moduli.size();
  moduli.push_back(mo);
  if (special) specialPrimes.insert(i);
  else ctxtPrimes.insert(i);
}
// The chain rule of FHEContext.cpp:88-115, stated as two searches over the values = 1 (mod 2m): word-size
// primes downwards from just above 2^NTL_SP_NBITS - 1 while a whole one still fits into the size that is
// left, then one closing prime searched UPWARDS from exp(what is left) -- so the chain covers totalSize
// without overshooting by more than one stride.  (tests: test_chain_matches_survey pins the result.)
double AddPrimesBySize(FHEcontext &context, double totalSize, bool special) {
  const long m = context.zMstar.M();
  if (m <= 0 || m > (1L << 20)) Error("AddModuli1: m undefined or larger than 2^20");
  const long stride = 2 * m;
  const auto nextPrime = [](long from, long step) {
    do from += step;
    while (!ProbPrime(from));
    return from;
  };
  long cand = (((1L << NTL_SP_NBITS) - 1) / stride + 1) * stride + 1;
  double remaining = totalSize;
  bool closing = false;
  while (remaining > 0.0) {
    if (!closing && remaining < std::log((double)cand)) {
      closing = true;
      const long target = (long)ceil(exp(remaining));
      cand = target - target % stride + 1;
    }
    cand = nextPrime(cand, closing ? stride : -stride);
    if (context.inChain(cand)) continue;
    context.AddPrime(cand, special);
    remaining -= std::log((double)cand);
  }
  return totalSize - remaining;
}
// Natural log of the size SetUpSIContext(1) asks of the chain (FHEContext.cpp:83-85)
static double SIChainBaseSize(const FHEcontext &c) {
  return NTL::log(c.modulusQ) * 2 + NTL::log(c.ModulusP()) + std::log((double)c.zMstar.phiM()) * 2 + std::log(2.0);
}
// Host-only users of the class surface (fhesih_keygen) switch this off: they never touch a device.
static bool g_eagerDevice = true;
void FHEcontext::SetUpSIContext(long xi) {  // FHEContext.cpp:83-85
  xiHint = xi < 1 ? 1 : xi;
  AddPrimesBySize(*this, SIChainBaseSize(*this) + std::log((double)xiHint), false);
  // The reference builds every Cmodulus (roots, Bluestein tables) here; the device counterpart --
  // CUDA context, twiddle / CRT tables, kernel images -- is created here too, not at the first
  // operator, so that a client's own timers see operators and not start-up.
  if (g_eagerDevice) {
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    Dev();
  }
}
fhesi_ctx *FHEcontext::Dev() const {
  if (dev && devXi != xiHint) {
    // the head-room changed after the device image was made (SetUpSIContext / ImportSIContext called again):
    // objects of the old image would silently keep the old tensor chain
    Error("FHEcontext: xi changed after the device context was created");
  }
  if (!dev) {
    Check(fhesi_ctx_create(zMstar.M(), logQ, (uint64_t)to_long(ModulusP()), decompSize, (uint64_t)xiHint, device, &dev),
          "fhesi_ctx_create");
    devXi = xiHint;
  }
  return dev;
}
void FHEcontext::ExportSIContext(ofstream &out) {  // FHEContext.cpp:45-60
  unsigned m = zMstar.M(), g = Generator();
  Export(out, m);
  Export(out, logQ);
  Export(out, ModulusP());
  Export(out, g);
  Export(out, decompSize);
  uint32_t size = moduli.size();
  Export(out, size);
  for (auto &mo : moduli) {
    long q = mo.q, root = mo.root;
    Export(out, q);
    Export(out, root);
  }
}
void FHEcontext::ImportSIContext(ifstream &in) {  // FHEContext.cpp:62-81
  unsigned m, lq, generator, ds;
  ZZ p;
  Import(in, m);
  Import(in, lq);
  Import(in, p);
  Import(in, generator);
  Import(in, ds);
  Init(m, lq, p, generator, ds);
  uint32_t size;
  Import(in, size);
  long q, root;
  double chain = 0;
  for (unsigned i = 0; i < size; i++) {
    Import(in, q);
    Import(in, root);
    if (!in.good() || q < 3) Error("ImportSIContext: truncated or corrupt context file");
    AddPrime(q, false, root);
    chain += std::log((double)q);
  }
  // The file does not store xi; the chain it stores does: SetUpSIContext(xi) sized it to cover
  // base + log(xi).  The device's own tensor chain gets the same head-room for sums of tensor products
  // (Matrix sums in Regression.h / Statistics.h), instead of the default xi = 1.
  const double room = chain - SIChainBaseSize(*this);
  xiHint = room > 0 ? std::max(1L, (long)std::floor(std::exp(std::min(room, 40.0)))) : 1;
}
ostream &operator<<(ostream &os, const FHEcontext &context) {
  os << "logQ: " << context.logQ << endl << "p: " << context.ModulusP() << endl
     << "g: " << context.Generator() << endl << "primes: [";
  for (auto &mo : context.moduli) os << mo.q << ", ";
  return os << "]" << endl;
}

// ------------------------------------------------------------------------------- word packing
static void PackZZ(uint32_t *w, const ZZ &c, unsigned W) {  // two's complement, W words
  for (unsigned k = 0; k < W; ++k) w[k] = k < c.mag.size() ? c.mag[k] : 0u;
  if (c.neg) {
    uint64_t carry = 1;
    for (unsigned k = 0; k < W; ++k) {
      carry += (uint32_t)~w[k];
      w[k] = (uint32_t)carry;
      carry >>= 32;
    }
  }
}
static ZZ UnpackZZ(const uint32_t *w, unsigned W) {
  ZZ c;
  bool neg = w[W - 1] >> 31;
  c.mag.assign(w, w + W);
  if (neg) {
    uint64_t carry = 1;
    for (unsigned k = 0; k < W; ++k) {
      carry += (uint32_t)~c.mag[k];
      c.mag[k] = (uint32_t)carry;
      carry >>= 32;
    }
  }
  c.neg = neg;
  c.trim();
  return c;
}
static std::vector<uint32_t> PackPoly(const ZZX &a, unsigned n, unsigned W) {
  std::vector<uint32_t> w((size_t)n * W, 0);
  for (long i = 0; i <= deg(a) && i < (long)n; ++i) PackZZ(&w[(size_t)i * W], a.rep.v[i], W);
  return w;
}
static ZZX UnpackPoly(const uint32_t *w, unsigned n, unsigned W) {
  ZZX a;
  a.rep.v.resize(n);
  for (unsigned i = 0; i < n; ++i) a.rep.v[i] = UnpackZZ(w + (size_t)i * W, W);
  a.normalize();
  return a;
}

// ------------------------------------------------------------------------------- DoubleCRT
DoubleCRT::DoubleCRT() : context(*activeContext) {}
DoubleCRT::DoubleCRT(const FHEcontext &c) : context(c) {}
DoubleCRT::DoubleCRT(const ZZX &p) : context(*activeContext), poly(p) { wrap(); }
DoubleCRT::DoubleCRT(const ZZX &p, const FHEcontext &c) : context(c), poly(p) { wrap(); }
void DoubleCRT::wrap() {  // the matrix represents the polynomial modulo the chain product, centred
  RemPhim(poly, context.zMstar);
  ZZ P = context.productOfPrimes(), half = P / 2;
  for (auto &c : poly.rep.v) {
    if (c > half || c < -half) {
      c %= P;
      if (c > half) c -= P;
    }
  }
  poly.normalize();
}
DoubleCRT &DoubleCRT::operator=(const DoubleCRT &o) {
  if (&context != &o.context) Error("DoubleCRT assigment: incompatible contexts");
  poly = o.poly;
  return *this;
}
DoubleCRT &DoubleCRT::operator=(const ZZX &p) { poly = p; wrap(); return *this; }
DoubleCRT &DoubleCRT::operator=(const ZZ &num) { poly = to_ZZX(num); wrap(); return *this; }
DoubleCRT &DoubleCRT::operator+=(const DoubleCRT &o) {
  if (&context != &o.context) Error("DoubleCRT::Op: incompatible objects");
  poly += o.poly;
  wrap();
  return *this;
}
DoubleCRT &DoubleCRT::operator-=(const DoubleCRT &o) {
  if (&context != &o.context) Error("DoubleCRT::Op: incompatible objects");
  poly -= o.poly;
  wrap();
  return *this;
}
DoubleCRT &DoubleCRT::operator*=(const DoubleCRT &o) {
  if (&context != &o.context) Error("DoubleCRT::Op: incompatible objects");
  poly = MulModPhim(poly, o.poly, context.zMstar);
  wrap();
  return *this;
}
DoubleCRT &DoubleCRT::operator+=(const ZZ &c) {  // adds c to every evaluation = to the constant term
  SetCoeff(poly, 0, coeff(poly, 0) + c);
  wrap();
  return *this;
}
DoubleCRT &DoubleCRT::operator*=(const ZZ &c) { poly *= c; wrap(); return *this; }
void DoubleCRT::toPoly(ZZX &p, bool positive) const {
  p = poly;
  if (positive) {
    ZZ P = context.productOfPrimes();
    for (auto &c : p.rep.v)
      if (c < 0L) c += P;
  }
  p.normalize();
}
IndexMap<vec_long> DoubleCRT::getMap() const {
  IndexMap<vec_long> m;
  vector<vector<long>> rows = getRows();
  for (size_t i = 0; i < rows.size(); ++i) {
    m.insert((long)i);
    m[(long)i].v = rows[i];
  }
  return m;
}
IndexSet DoubleCRT::getIndexSet() const { return context.ctxtPrimes; }
DoubleCRT &DoubleCRT::operator/=(const ZZ &num) {  // every row times num^-1 mod its prime == poly * num^-1 mod P
  const ZZ P = context.productOfPrimes();
  ZZ r = num % P;
  poly *= InvMod(r, P);
  wrap();
  return *this;
}
void DoubleCRT::Exp(long e) {
  vector<vector<long>> rows = getRows();
  for (size_t i = 0; i < rows.size(); ++i) {
    const long pi = context.ithPrime(i);
    for (auto &v : rows[i]) v = PowerMod(v, e, pi);
  }
  setRows(rows);
}
void DoubleCRT::randomize(const ZZ *seed) {
  if (seed != NULL) SetSeed(*seed);
  const unsigned n = context.zMstar.phiM(), L = context.numPrimes();
  vector<vector<long>> rows(L, vector<long>(n));
  for (unsigned i = 0; i < L; ++i) {
    const long pi = context.ithPrime(i);
    for (unsigned j = 0; j < n; ++j) rows[i][j] = RandomBnd(pi);
  }
  setRows(rows);
}
void DoubleCRT::automorph(long k) {  // DoubleCRT.cpp:439-465 in coefficient form
  const PAlgebra &z = context.zMstar;
  if (!z.inZmStar(k)) Error("DoubleCRT::automorph: k not in Zm*");
  const long m = z.M();
  ZZX r;
  r.rep.v.assign(m, ZZ());
  for (long i = 0; i <= deg(poly); ++i) r.rep.v[MulMod(i, k, m)] += poly.rep.v[i];
  r.normalize();
  poly = r;
  // fold modulo X^m - 1 happened above; now modulo Phi_m
  if (m % 2 == 0 && z.phiM() == (unsigned)(m / 2 - 1)) {
    ZZX t;  // X^(m/2) = -1
    const long h = m / 2;
    t.rep.v.assign(h, ZZ());
    for (long i = 0; i <= deg(poly); ++i) {
      if (i >= h) t.rep.v[i - h] -= poly.rep.v[i];
      else t.rep.v[i] += poly.rep.v[i];
    }
    t.normalize();
    poly = t;
  }
  wrap();
}
vector<vector<long>> DoubleCRT::getRows() const {
  const unsigned n = context.zMstar.phiM(), L = context.numPrimes();
  unsigned Win = (unsigned)((NumBits(context.productOfPrimes()) + 32) / 32) + 1;
  std::vector<uint32_t> words = PackPoly(poly, n, Win);
  std::vector<uint64_t> primes(L), roots(L);
  for (unsigned i = 0; i < L; ++i) primes[i] = context.ithPrime(i), roots[i] = context.ithModulus(i).root;
  std::vector<int64_t> flat((size_t)L * n);
  Check(fhesi_ref_rows_host(context.Dev(), words.data(), Win, primes.data(), roots.data(), L, flat.data()),
        "fhesi_ref_rows_host");
  vector<vector<long>> rows(L, vector<long>(n));
  for (unsigned i = 0; i < L; ++i)
    for (unsigned j = 0; j < n; ++j) rows[i][j] = flat[(size_t)i * n + j];
  return rows;
}
void DoubleCRT::setRows(const vector<vector<long>> &rows) {
  // Cmodulus::iFFT per prime (CModulus.cpp:110-132), then DoubleCRT::toPoly's incremental CRT
  // (DoubleCRT.cpp:349-398, NumbTh.cpp:307-335); import-time only
  const PAlgebra &z = context.zMstar;
  const long m = z.M(), n = z.phiM();
  std::vector<long> units;
  for (long i = 0; i < m; ++i)
    if (z.inZmStar(i)) units.push_back(i);
  std::vector<long> phi(n + 1);
  std::vector<ZZ> acc(n);
  ZZ prod(1L);
  for (size_t l = 0; l < rows.size(); ++l) {
    const long q = context.ithPrime(l), root = context.ithModulus(l).root;
    const long zinv = InvMod(MulMod(root, root, q), q), minv = InvMod(m % q, q);
    for (long i = 0; i <= n; ++i) phi[i] = to_long(coeff(z.PhimX(), i) % ZZ(q));
    std::vector<long> c(m), zpow(m);
    zpow[0] = 1;
    for (long t = 1; t < m; ++t) zpow[t] = MulMod(zpow[t - 1], zinv, q);
    for (long t = 0; t < m; ++t) {
      unsigned __int128 s = 0;
      for (long j = 0; j < n; ++j) s += (unsigned long)MulMod(rows[l][j], zpow[(units[j] * t) % m], q);
      c[t] = MulMod((long)(s % (unsigned long)q), minv, q);
    }
    for (long i = m - 1; i >= n; --i) {
      long ci = c[i];
      if (!ci) continue;
      for (long j = 0; j <= n; ++j) c[i - n + j] = SubMod(c[i - n + j], MulMod(ci, phi[j], q), q);
    }
    if (l == 0) {
      for (long j = 0; j < n; ++j) acc[j] = ZZ(c[j] > q / 2 ? c[j] - q : c[j]);
      prod = ZZ(q);
    } else {
      const long pinv = InvMod(rem(prod, q), q), qh = q / 2;
      for (long j = 0; j < n; ++j) {
        long d = MulMod(SubMod(c[j], rem(acc[j], q), q), pinv, q);
        if (d > qh) d -= q;
        acc[j] += prod * ZZ(d);
      }
      prod *= ZZ(q);
    }
  }
  poly.rep.v = acc;
  poly.normalize();
}

// ------------------------------------------------------------------------------- Ciphertext
DevBuf::DevBuf(fhesi_ctx *c, size_t b) : ctx(c), bytes(b) {
  void *p = nullptr;
  Check(fhesi_malloc(c, b, &p), "fhesi_malloc");
  ptr = (uint32_t *)p;
}
DevBuf::~DevBuf() {
  if (ptr) fhesi_free(ctx, ptr);
}
Ciphertext::Ciphertext(const FHESIPubKey &pk) : context(&pk.GetContext()) {}
Ciphertext::Ciphertext(const Ciphertext &o)
    : context(o.context), nparts(o.nparts), wordsPer(o.wordsPer), scaledUp(o.scaledUp), hostStale(o.hostStale),
      hostDirty(o.hostDirty), parts(o.parts) {
  if (o.buf) {
    buf = make_shared<DevBuf>(o.buf->ctx, o.buf->bytes);
    Check(fhesi_d2d(buf->ctx, buf->ptr, o.buf->ptr, buf->bytes), "fhesi_d2d");
  }
}
Ciphertext &Ciphertext::operator=(const Ciphertext &o) {
  if (this == &o) return *this;
  context = o.context;
  nparts = o.nparts, wordsPer = o.wordsPer, scaledUp = o.scaledUp, hostStale = o.hostStale;
  hostDirty = o.hostDirty;
  parts = o.parts;
  buf.reset();
  if (o.buf) {
    buf = make_shared<DevBuf>(o.buf->ctx, o.buf->bytes);
    Check(fhesi_d2d(buf->ctx, buf->ptr, o.buf->ptr, buf->bytes), "fhesi_d2d");
  }
  return *this;
}
void Ciphertext::Alloc(unsigned np, unsigned words) {
  fhesi_ctx *d = context->Dev();
  nparts = np, wordsPer = words;
  buf = make_shared<DevBuf>(d, (size_t)np * context->zMstar.phiM() * words * 4);
}
void Ciphertext::Initialize(unsigned n, const FHEcontext &c) {
  context = &c;
  Clear();
  if (n) {
    Alloc(n, c.Words());
    std::vector<uint32_t> z(buf->bytes / 4, 0);
    Check(fhesi_h2d(buf->ctx, buf->ptr, z.data(), buf->bytes), "fhesi_h2d");
  }
  hostStale = true;
}
void Ciphertext::Clear() {
  buf.reset();
  nparts = 0, wordsPer = 0, scaledUp = false, hostStale = false, hostDirty = false;
  parts.clear();
}
void Ciphertext::SyncHost() const {
  Ciphertext *self = const_cast<Ciphertext *>(this);
  if (!hostStale) return;
  self->parts.clear();
  if (!scaledUp && buf) {
    const unsigned n = context->zMstar.phiM();
    std::vector<uint32_t> w(buf->bytes / 4);
    Check(fhesi_d2h(buf->ctx, w.data(), buf->ptr, buf->bytes), "fhesi_d2h");
    for (unsigned i = 0; i < nparts; ++i) {
      CiphertextPart part(*context);
      part.poly = UnpackPoly(&w[(size_t)i * n * wordsPer], n, wordsPer);
      self->parts.push_back(part);
    }
  }
  hostStale = false;
}
void Ciphertext::UploadHost() {  // after Import / a write through operator[]
  const unsigned n = context->zMstar.phiM(), W = context->Words();
  Alloc(parts.size(), W + 1);  // head-room word: imported parts need not be reduced
  std::vector<uint32_t> w;
  for (auto &p : parts) {
    std::vector<uint32_t> pw = PackPoly(p.poly, n, W + 1);
    w.insert(w.end(), pw.begin(), pw.end());
  }
  if (!w.empty()) Check(fhesi_h2d(buf->ctx, buf->ptr, w.data(), w.size() * 4), "fhesi_h2d");
  scaledUp = false;
  hostStale = false;
  hostDirty = false;
}
// A write through the non-const operator[] reaches the device here, before any operator reads the buffer.
void Ciphertext::Flush() const {
  if (hostDirty) const_cast<Ciphertext *>(this)->UploadHost();
}
void Ciphertext::EnsureReduced() {
  Flush();
  if (scaledUp || !buf || wordsPer == context->Words()) return;
  const unsigned W = context->Words();
  auto nb = make_shared<DevBuf>(buf->ctx, (size_t)nparts * context->zMstar.phiM() * W * 4);
  Check(fhesi_reduce_wide_dev(buf->ctx, buf->ptr, wordsPer, nb->ptr, nparts, 1), "fhesi_reduce_wide_dev");
  buf = nb;
  wordsPer = W;
  hostStale = true;
}
CiphertextPart Ciphertext::GetPart(unsigned ind) const {
  SyncHost();
  return parts[ind];
}
CiphertextPart &Ciphertext::operator[](unsigned ind) {
  SyncHost();
  if (!scaledUp) hostDirty = true;  // the caller holds a writable reference into the mirror
  return parts[ind];
}
Ciphertext &Ciphertext::operator+=(const Ciphertext &o) {  // Ciphertext.cpp:123-145
  Flush();
  o.Flush();
  assert(scaledUp == o.scaledUp);
  if (!o.buf || o.nparts == 0) return *this;
  if (!buf || nparts == 0) return *this = o;
  fhesi_ctx *d = context->Dev();
  const unsigned n = context->zMstar.phiM();
  Ciphertext tmp(*context);
  const Ciphertext *rhs = &o;
  if (!scaledUp) {
    EnsureReduced();
    if (o.wordsPer != context->Words()) {
      tmp = o;
      tmp.EnsureReduced();
      rhs = &tmp;
    }
  }
  const size_t partBytes = scaledUp ? fhesi_tprod_bytes(d, 1) : (size_t)n * wordsPer * 4;
  if (rhs->nparts > nparts) {  // "parts.push_back(other.parts[i])" for the missing ones
    auto nb = make_shared<DevBuf>(d, partBytes * rhs->nparts);
    Check(fhesi_d2d(d, nb->ptr, buf->ptr, partBytes * nparts), "fhesi_d2d");
    Check(fhesi_d2d(d, (char *)nb->ptr + partBytes * nparts, (char *)rhs->buf->ptr + partBytes * nparts,
                    partBytes * (rhs->nparts - nparts)), "fhesi_d2d");
    unsigned common = nparts;
    buf = nb;
    nparts = rhs->nparts;
    if (scaledUp) Check(fhesi_tprod_add_dev(d, buf->ptr, rhs->buf->ptr, common, 1), "fhesi_tprod_add_dev");
    else Check(fhesi_ct_add_dev(d, buf->ptr, rhs->buf->ptr, common, 1), "fhesi_ct_add_dev");
  } else {
    if (scaledUp) Check(fhesi_tprod_add_dev(d, buf->ptr, rhs->buf->ptr, rhs->nparts, 1), "fhesi_tprod_add_dev");
    else Check(fhesi_ct_add_dev(d, buf->ptr, rhs->buf->ptr, rhs->nparts, 1), "fhesi_ct_add_dev");
  }
  hostStale = true;
  return *this;
}
Ciphertext &Ciphertext::operator+=(const ZZX &other) {  // Ciphertext.cpp:147-161
  Flush();
  fhesi_ctx *d = context->Dev();
  const unsigned n = context->zMstar.phiM(), W = context->Words();
  if (scaledUp) {  // :157-159  tProd[0] += scaledConstant, nothing reduced
    if (!buf) return *this;
    ZZX sc = other;
    for (long i = 0; i <= deg(sc); i++) {
      sc.rep[i] <<= (long)context->logQ;
      sc.rep[i] /= context->ModulusP();
    }
    std::vector<uint32_t> w = PackPoly(sc, n, W + 1);
    DevBuf t(d, w.size() * 4);
    Check(fhesi_h2d_async(d, t.ptr, w.data(), w.size() * 4), "fhesi_h2d_async");
    Check(fhesi_tprod_add_poly_dev(d, buf->ptr, nparts, t.ptr, W + 1, 1), "fhesi_tprod_add_poly_dev");
    return *this;
  }
  EnsureReduced();
  ZZX sc = other;
  for (long i = 0; i <= deg(sc); i++) {
    sc.rep[i] <<= (long)context->logQ;
    sc.rep[i] /= context->ModulusP();
    Reduce(sc.rep[i], context->logQ);
  }
  std::vector<uint32_t> w = PackPoly(sc, n, W);
  DevBuf t(d, w.size() * 4);
  Check(fhesi_h2d_async(d, t.ptr, w.data(), w.size() * 4), "fhesi_h2d_async");
  Check(fhesi_ct_add_dev(d, buf->ptr, t.ptr, 1, 1), "fhesi_ct_add_dev");
  hostStale = true;
  return *this;
}
Ciphertext &Ciphertext::operator*=(const Ciphertext &o) {  // Ciphertext.cpp:167-192
  if (scaledUp || o.scaledUp) Error("Ciphertext *= : operands must not be in tensor form (Ciphertext.cpp:169-176)");
  Flush();
  o.Flush();
  fhesi_ctx *d = context->Dev();
  Ciphertext rhs(o);  // also covers self-multiplication
  EnsureReduced();
  rhs.EnsureReduced();
  const unsigned po = nparts + rhs.nparts - 1;
  auto nb = make_shared<DevBuf>(d, fhesi_tprod_bytes(d, po));
  Check(fhesi_ct_tensor_dev(d, buf->ptr, nparts, rhs.buf->ptr, rhs.nparts, nb->ptr, 1, 0), "fhesi_ct_tensor_dev");
  buf = nb;
  nparts = po;
  wordsPer = 0;
  scaledUp = true;
  parts.clear();
  hostStale = true;
  return *this;
}
Ciphertext &Ciphertext::operator*=(long l) {  // Ciphertext.cpp:233-244
  Flush();
  if (!buf) return *this;
  fhesi_ctx *d = context->Dev();
  if (!scaledUp) {
    EnsureReduced();
    Check(fhesi_ct_mul_scalar_dev(d, buf->ptr, l, nparts, 1), "fhesi_ct_mul_scalar_dev");
  } else {
    Check(fhesi_tprod_mul_scalar_dev(d, buf->ptr, l, nparts, 1), "fhesi_tprod_mul_scalar_dev");
  }
  hostStale = true;
  return *this;
}
Ciphertext &Ciphertext::operator*=(const ZZX &other) {  // Ciphertext.cpp:246-258
  Flush();
  if (!buf) return *this;
  fhesi_ctx *d = context->Dev();
  const unsigned n = context->zMstar.phiM();
  if (scaledUp) {  // :252-256  tProd[i] *= DoubleCRT(other), `other` as an integer polynomial
    const unsigned W = context->Words();
    ZZX o = other;
    RemPhim(o, context->zMstar);
    std::vector<uint32_t> w = PackPoly(o, n, W + 1);
    DevBuf t(d, w.size() * 4);
    Check(fhesi_h2d_async(d, t.ptr, w.data(), w.size() * 4), "fhesi_h2d_async");
    Check(fhesi_tprod_mul_poly_dev(d, buf->ptr, nparts, t.ptr, W + 1, 1), "fhesi_tprod_mul_poly_dev");
    return *this;
  }
  EnsureReduced();
  std::vector<uint32_t> pt(n, 0);
  const ZZ &P = context->ModulusP();
  for (long i = 0; i <= deg(other) && i < (long)n; ++i) pt[i] = (uint32_t)to_long(other.rep.v[i] % P);
  // NOTE: the reference multiplies by `other` as an integer polynomial; callers pass to_ZZX of a
  // ZZ_pX (coefficients in [0,p)), for which reducing mod p is the identity.
  DevBuf t(d, n * 4);
  Check(fhesi_h2d_async(d, t.ptr, pt.data(), n * 4), "fhesi_h2d_async");
  Check(fhesi_ct_mul_plain_dev(d, buf->ptr, t.ptr, nparts, 1), "fhesi_ct_mul_plain_dev");
  hostStale = true;
  return *this;
}
Ciphertext &Ciphertext::operator>>=(long k) {  // Ciphertext.cpp:264-275
  Flush();
  if (!buf) return *this;
  fhesi_ctx *d = context->Dev();
  if (scaledUp) {  // :269-273  tProd[i] >>= k
    auto nb = make_shared<DevBuf>(d, fhesi_tprod_bytes(d, nparts));
    Check(fhesi_tprod_automorph_dev(d, buf->ptr, nparts, (uint32_t)k, nb->ptr, 1), "fhesi_tprod_automorph_dev");
    buf = nb;
    return *this;
  }
  EnsureReduced();
  const unsigned n = context->zMstar.phiM(), W = context->Words();
  auto nb = make_shared<DevBuf>(d, (size_t)nparts * n * (W + 1) * 4);
  Check(fhesi_ct_automorph_dev(d, buf->ptr, nparts, (uint32_t)k, nb->ptr, 1), "fhesi_ct_automorph_dev");
  buf = nb;
  wordsPer = W + 1;
  hostStale = true;
  return *this;
}
void Ciphertext::ScaleDown() {  // Ciphertext.cpp:194-218
  if (!scaledUp) return;
  fhesi_ctx *d = context->Dev();
  auto nb = make_shared<DevBuf>(d, fhesi_ct_bytes(d, nparts));
  Check(fhesi_scaledown_dev(d, buf->ptr, nparts, nb->ptr, 1), "fhesi_scaledown_dev");
  buf = nb;
  wordsPer = context->Words();
  scaledUp = false;
  hostStale = true;
}
ostream &operator<<(ostream &os, const Ciphertext &c) {
  Ciphertext t(c);
  t.ScaleDown();
  t.SyncHost();
  for (unsigned i = 0; i < t.size(); i++) os << t.parts[i] << ", ";
  return os;
}

// ------------------------------------------------------------------------------- keys
void FHESISecKey::Init(const FHEcontext &c) {  // FHE-SI.cpp:86-91
  sKeys.assign(2, DoubleCRT(c));
  sKeys[0] = 1;
  sKeys[1].sampleHWt(64);
  devKey.reset();
}
static shared_ptr<fhesi_key> UploadKey(const FHEcontext &c, const vector<DoubleCRT> &rep) {
  const unsigned n = c.zMstar.phiM(), W = c.Words();
  std::vector<uint32_t> w;
  for (auto &d : rep) {
    ZZX p;
    d.toPoly(p);
    ReduceCoefficients(p, c.logQ);
    std::vector<uint32_t> pw = PackPoly(p, n, W);
    w.insert(w.end(), pw.begin(), pw.end());
  }
  fhesi_key *k = nullptr;
  Check(fhesi_key_create(c.Dev(), w.data(), rep.size(), &k), "fhesi_key_create");
  return shared_ptr<fhesi_key>(k, [](fhesi_key *p) { fhesi_key_destroy(p); });
}
void FHESISecKey::Decrypt(Plaintext &ptxt, const Ciphertext &ct) const {  // FHE-SI.cpp:93-119
  if (!devKey) devKey = UploadKey(context, sKeys);
  Ciphertext c(ct);
  c.ScaleDown();
  c.EnsureReduced();
  if (c.size() < sKeys.size()) Error("Decrypt: ciphertext has fewer parts than the secret key");
  fhesi_ctx *d = context.Dev();
  const unsigned n = context.zMstar.phiM();
  DevBuf m(d, n * 4);
  Check(fhesi_decrypt_dev(d, devKey.get(), c.buf->ptr, c.size(), m.ptr, 1), "fhesi_decrypt_dev");
  std::vector<uint32_t> h(n);
  Check(fhesi_d2h(d, h.data(), m.ptr, n * 4), "fhesi_d2h");
  ptxt.message.rep.v.assign(n, ZZ_p());
  for (unsigned i = 0; i < n; ++i) ptxt.message.rep.v[i].v = h[i];
  ptxt.message.normalize();
}
void FHESISecKey::Export(ofstream &out) const { ::Export(out, sKeys); }
void FHESISecKey::Import(ifstream &in) { ::Import(in, sKeys); devKey.reset(); }

// When set, KeySwitchSI::Init only performs its random draws (in the reference's order) and hands them
// over instead of doing the arithmetic: fhesih_keydraws feeds them to fhesi_ksw_generate on the device.
struct KeyDrawSink {
  struct Matrix {
    vector<ZZX> src, polys, errs;
    ZZX t;
  };
  vector<Matrix> matrices;
};
static KeyDrawSink *g_drawSink = nullptr;
static void SampleRandomWords(uint32_t *dst, unsigned n, unsigned W, unsigned k, RandomStream &rs);
static void SampleRandomWords(uint32_t *dst, unsigned n, unsigned W, unsigned k) {
  SampleRandomWords(dst, n, W, k, GlobalRandomStream());
}
// every coefficient fits an int32 (s, s^2, s(X^k): yes; an arbitrary imported key: maybe not)
static bool SmallPoly(std::vector<int32_t> &dst, const ZZX &a, unsigned n) {
  dst.assign(n, 0);
  if (deg(a) >= (long)n) return false;
  for (long i = 0; i <= deg(a); ++i) {
    const ZZ &c = a.rep.v[i];
    if (c.mag.size() > 1 || (c.mag.size() == 1 && c.mag[0] > 0x7fffffffu)) return false;
    dst[i] = (int32_t)to_long(c);
  }
  return true;
}
void FHESIPubKey::Init(const FHESISecKey &secKey) {  // FHE-SI.cpp:42-62
  const unsigned n = context.zMstar.phiM(), W = context.Words();
  ZZX s;
  secKey.GetRepresentation()[1].toPoly(s);
  std::vector<int32_t> sv;
  publicKey.clear();
  hostWords.clear();
  devKey.reset();
  if (g_eagerDevice && !g_drawSink && SmallPoly(sv, s, n)) {
    // on the device: c0 = e + s * c1 reduced, c1' = Reduce(-c1) -- one key-generation entry with no source
    // term.  Draw order as the reference: the Gaussian first, then the uniform polynomial (:44-47).
    std::vector<int32_t> e(n);
    std::vector<uint32_t> c1((size_t)n * W);
    SampleGaussianInts(e.data(), n, context.stdev);
    SampleRandomWords(c1.data(), n, W, context.logQ);
    hostWords.resize((size_t)2 * n * W);
    fhesi_key *k = nullptr;
    Check(fhesi_keygen_batch(context.Dev(), 0, nullptr, nullptr, sv.data(), c1.data(), e.data(), nullptr, nullptr, nullptr,
                             &k, hostWords.data()),
          "fhesi_keygen_batch (public key)");
    devKey = shared_ptr<fhesi_key>(k, [](fhesi_key *p) { fhesi_key_destroy(p); });
    return;
  }
  ZZX c0, c1;
  sampleGaussian(c0, context.zMstar.phiM(), context.stdev);
  SampleRandom(c1, context.modulusQ, context.zMstar.phiM());
  c0 += MulModPhim(s, c1, context.zMstar);
  c1 *= -1;
  ReduceCoefficients(c0, context.logQ);
  ReduceCoefficients(c1, context.logQ);
  publicKey.push_back(DoubleCRT(c0, context));
  publicKey.push_back(DoubleCRT(c1, context));
}
void FHESIPubKey::Materialize() const {
  if (hostWords.empty()) return;
  const unsigned n = context.zMstar.phiM(), W = context.Words();
  publicKey.clear();
  for (int i = 0; i < 2; ++i) publicKey.push_back(DoubleCRT(UnpackPoly(hostWords.data() + (size_t)i * n * W, n, W), context));
  hostWords.clear();
}
void FHESIPubKey::Encrypt(Ciphertext &ctxt, const Plaintext &ptxt) const {  // FHE-SI.cpp:10-36
  if (!devKey) devKey = UploadKey(context, GetRepresentation());
  fhesi_ctx *d = context.Dev();
  const unsigned n = context.zMstar.phiM();
  // one staging block: message | e0 | e1 | r, one stream-ordered copy, no synchronisation -- the
  // next call's sampling overlaps this call's device work
  std::vector<uint32_t> stage(3 * n + (n + 3) / 4, 0);
  uint8_t *r = (uint8_t *)&stage[3 * n];
  for (unsigned i = 0; i < n; i++) r[i] = (uint8_t)RandomBnd(2L);        // :14-17
  for (unsigned i = 0; i < 2; ++i)                                         // :24, one draw per part
    SampleGaussianInts((int32_t *)&stage[(1 + i) * n], n, context.stdev);
  for (long j = 0; j <= deg(ptxt.message) && j < (long)n; ++j) stage[j] = (uint32_t)ptxt.message.rep.v[j].v;
  auto ds = make_shared<DevBuf>(d, stage.size() * 4);
  Check(fhesi_h2d_async(d, ds->ptr, stage.data(), stage.size() * 4), "fhesi_h2d_async");
  const uint32_t *dw = (const uint32_t *)ds->ptr;
  ctxt.context = &context;
  ctxt.Clear();
  ctxt.Alloc(2, context.Words());
  Check(fhesi_encrypt_dev(d, devKey.get(), dw, (const uint8_t *)(dw + 3 * n), (const int32_t *)(dw + n),
                          ctxt.buf->ptr, 1), "fhesi_encrypt_dev");
  ctxt.hostStale = true;
}
void FHESIPubKey::Export(ofstream &out) const { ::Export(out, GetRepresentation()); }
void FHESIPubKey::Import(ifstream &in) { hostWords.clear(); ::Import(in, publicKey); devKey.reset(); }

void KeySwitchSI::Init(const FHESISecKey &src, const FHESISecKey &dst) {  // FHE-SI.cpp:153-209
  const vector<DoubleCRT> &s = src.GetRepresentation();
  vector<ZZX> sCoeff(s.size());
  for (size_t i = 0; i < s.size(); i++) s[i].toPoly(sCoeff[i]);
  ZZX t;
  dst.GetRepresentation()[1].toPoly(t);
  const size_t n = src.GetSize(), D = context.ndigits, total = n * D;
  static const bool timing = getenv("FHESIH_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tt0 = now();
  keySwitchMatrix.clear();
  hostB.clear(), drawA.clear();
  entries = total;
  if (g_eagerDevice && !g_drawSink && InitOnDevice(sCoeff, t)) return;
  // 1. the random draws, in the reference's order (:176 SampleRandom, :188 sampleGaussian per entry)
  vector<ZZX> polys(total), errs(total);
  for (size_t ind = 0; ind < total; ++ind) {
    SampleRandom(polys[ind], context.modulusQ, context.zMstar.phiM());
    sampleGaussian(errs[ind], context.zMstar.phiM(), context.stdev);
  }
  if (g_drawSink) {
    g_drawSink->matrices.push_back(KeyDrawSink::Matrix{sCoeff, polys, errs, t});
    keySwitchMatrix.assign(2, vector<DoubleCRT>());
    devKsw.reset();
    return;
  }
  const double tt1 = now();
  // 2. the arithmetic of each entry is independent of the others: b = poly * t + err + s_i * 2^(24 j)
  // reduced mod q (:182-199), A = -poly unreduced (:178-180); spread over the host's cores
  vector<DoubleCRT> A(total, DoubleCRT(context)), b(total, DoubleCRT(context));
  auto entry = [&](size_t ind) {
    const size_t i = ind / D, j = ind % D;
    ZZX bCoeff = MulModPhim(polys[ind], t, context.zMstar);
    bCoeff += errs[ind];
    ZZX sh = sCoeff[i];
    for (long k = 0; k <= deg(sh); k++) sh.rep[k] <<= (long)(8 * context.decompSize * j);
    bCoeff += sh;
    ReduceCoefficients(bCoeff, context.logQ);
    polys[ind] *= -1;
    A[ind] = DoubleCRT(polys[ind], context);
    b[ind] = DoubleCRT(bCoeff, context);
  };
  ParallelFor(total, entry);
  if (timing) fprintf(stderr, "KeySwitchSI::Init: draws %.4f  arithmetic %.4f s (%zu entries)\n", tt1 - tt0, now() - tt1, total);
  keySwitchMatrix.resize(2);
  keySwitchMatrix[0] = std::move(b);
  keySwitchMatrix[1] = std::move(A);
  devKsw.reset();
}
// The same matrix generated on the device: this class makes the draws (same stream positions and values as
// the loop above, flat arrays), fhesi_keygen_batch does b = A t + e + s_i 2^(24 j) and the key images.
bool KeySwitchSI::InitOnDevice(const vector<ZZX> &sCoeff, const ZZX &t) {
  const unsigned n = context.zMstar.phiM(), W = context.Words(), D = context.ndigits;
  const size_t parts = sCoeff.size(), K = parts * D;
  if (parts < 1 || parts > 3) return false;
  std::vector<int32_t> src(parts * n), tv, one;
  if (!SmallPoly(tv, t, n)) return false;
  for (size_t i = 0; i < parts; ++i) {
    if (!SmallPoly(one, sCoeff[i], n)) return false;
    memcpy(&src[i * n], one.data(), n * 4);
  }
  drawA.resize(K * n * W);
  hostB.resize(K * n * W);
  std::vector<int32_t> e(K * n);
  for (size_t k = 0; k < K; ++k) {  // :176 SampleRandom, :188 sampleGaussian per entry
    SampleRandomWords(&drawA[k * n * W], n, W, context.logQ);
    SampleGaussianInts(&e[k * n], n, context.stdev);
  }
  fhesi_ksw *k = nullptr;
  const uint32_t p32 = (uint32_t)parts;
  Check(fhesi_keygen_batch(context.Dev(), 1, &p32, src.data(), tv.data(), drawA.data(), e.data(), &k, hostB.data(), nullptr,
                           nullptr, nullptr),
        "fhesi_keygen_batch");
  devKsw = shared_ptr<fhesi_ksw>(k, [](fhesi_ksw *p) { fhesi_ksw_destroy(p); });
  return true;
}
// The reference's image of the matrix, when somebody asks: row 0 = b (reduced, as :198), row 1 = A' = -A
// (NOT reduced: -(-q/2) stays +q/2, :178-180), both as DoubleCRT over the reference chain.
void KeySwitchSI::Materialize() const {
  if (hostB.empty()) return;
  const unsigned n = context.zMstar.phiM(), W = context.Words();
  vector<DoubleCRT> A(entries, DoubleCRT(context)), b(entries, DoubleCRT(context));
  ParallelFor(entries, [&](size_t k) {
    ZZX a = UnpackPoly(drawA.data() + k * n * W, n, W);
    a *= -1;
    A[k] = DoubleCRT(a, context);
    b[k] = DoubleCRT(UnpackPoly(hostB.data() + k * n * W, n, W), context);
  });
  keySwitchMatrix.resize(2);
  keySwitchMatrix[0] = std::move(b);
  keySwitchMatrix[1] = std::move(A);
  hostB.clear(), drawA.clear();
}
// s^2 -> s (FHE-SI.cpp:211-227).  A tensor product of two k-part ciphertexts is linear in the 2k-1
// monomials 1, s, ..., s^(2k-2); they are the source key, s is the target.
void KeySwitchSI::InitS2(const FHESISecKey &s) {
  const vector<DoubleCRT> &base = s.GetRepresentation();  // (1, s)
  vector<DoubleCRT> powers;
  powers.reserve(2 * base.size() - 1);
  powers.push_back(base[0]);
  while (powers.size() < 2 * base.size() - 1) {
    DoubleCRT next = base[1];
    if (powers.size() > 1) next *= powers.back();
    powers.push_back(next);
  }
  FHESISecKey source(s.GetContext());  // the reference constructs -- hence samples -- a key here (:222): same draws
  source.UpdateRepresentation(powers);
  Init(source, s);
}
// s(X^k) -> s (FHE-SI.cpp:229-239): the key under which a ciphertext decrypts after `>>= k`
void KeySwitchSI::InitAutomorph(const FHESISecKey &s, unsigned k) {
  vector<DoubleCRT> rotated = s.GetRepresentation();
  FHESISecKey source(s.GetContext());  // sampled, then overwritten, as at :233
  for (DoubleCRT &part : rotated) part.automorph(k);
  source.UpdateRepresentation(rotated);
  Init(source, s);
}
const fhesi_ksw *KeySwitchSI::Dev() const {
  if (!devKsw) {
    Materialize();
    const unsigned n = context.zMstar.phiM(), W = context.Words();
    std::vector<uint32_t> wb, wA;
    for (int r = 0; r < 2; ++r) {
      for (auto &d : keySwitchMatrix[r]) {
        ZZX p;
        d.toPoly(p);
        ReduceCoefficients(p, context.logQ);  // only the value mod q matters
        std::vector<uint32_t> pw = PackPoly(p, n, W);
        (r ? wA : wb).insert((r ? wA : wb).end(), pw.begin(), pw.end());
      }
    }
    fhesi_ksw *k = nullptr;
    Check(fhesi_ksw_create(context.Dev(), wb.data(), wA.data(), keySwitchMatrix[0].size() / context.ndigits, &k),
          "fhesi_ksw_create");
    devKsw = shared_ptr<fhesi_ksw>(k, [](fhesi_ksw *p) { fhesi_ksw_destroy(p); });
  }
  return devKsw.get();
}
void KeySwitchSI::ApplyKeySwitch(Ciphertext &ctxt) const {  // FHE-SI.cpp:241-260
  ctxt.ScaleDown();
  ctxt.EnsureReduced();
  if (!entries || ctxt.size() * context.ndigits != entries)
    Error("ApplyKeySwitch: ciphertext size does not match the key-switch matrix");
  fhesi_ctx *d = context.Dev();
  auto nb = make_shared<DevBuf>(d, fhesi_ct_bytes(d, 2));
  Check(fhesi_keyswitch_dev(d, Dev(), ctxt.buf->ptr, nb->ptr, 1), "fhesi_keyswitch_dev");
  ctxt.buf = nb;
  ctxt.nparts = 2;
  ctxt.hostStale = true;
}
void KeySwitchSI::Export(ofstream &out) const { ::Export(out, GetRepresentation()); }
void KeySwitchSI::Import(ifstream &in) {
  hostB.clear(), drawA.clear();
  ::Import(in, keySwitchMatrix);
  entries = keySwitchMatrix.empty() ? 0 : keySwitchMatrix[0].size();
  devKsw.reset();
}

// ------------------------------------------------------------------------------- Serialization
void Export(ofstream &out, const ZZ &val) {  // Serialization.cpp:3-13
  uint32_t nBytes = NumBytes(val);
  out.write((char *)&nBytes, sizeof(uint32_t));
  bool neg = (val < 0L);
  out.write((char *)&neg, sizeof(bool));
  std::vector<unsigned char> data(nBytes + 1);
  BytesFromZZ(data.data(), val, nBytes);
  out.write((char *)data.data(), nBytes);
}
// Files come from another party: lengths are bounded before anything is allocated, and a short read is an
// error, not a silently truncated value.  No integer of this scheme is wider than the chain product
// (a few thousand bits); 1 MiB per integer, 2^24 coefficients / rows per vector are far above any real file.
static const size_t kMaxImportBytes = (size_t)1 << 20, kMaxImportItems = (size_t)1 << 24;
static void ImportCheck(ifstream &in, const char *what) {
  if (!in.good()) Error((std::string("Import: truncated or unreadable ") + what).c_str());
}
void Import(ifstream &in, ZZ &val) {
  uint32_t nBytes = 0;
  in.read((char *)&nBytes, sizeof(uint32_t));
  bool neg = false;
  in.read((char *)&neg, sizeof(bool));
  ImportCheck(in, "integer header");
  if ((size_t)nBytes > kMaxImportBytes) Error("Import: integer length out of range");
  std::vector<unsigned char> data((size_t)nBytes + 1);
  in.read((char *)data.data(), nBytes);
  ImportCheck(in, "integer");
  ZZFromBytes(val, data.data(), nBytes);
  if (neg) val *= -1;
}
void Export(ofstream &out, const ZZX &poly) {  // Serialization.cpp:29-36
  int32_t degree = deg(poly);
  out.write((char *)&degree, sizeof(int32_t));
  for (int i = 0; i <= degree; i++) Export(out, poly.rep[i]);
}
void Import(ifstream &in, ZZX &poly) {
  poly = ZZX::zero();
  int32_t degree = -1;
  in.read((char *)&degree, sizeof(int32_t));
  ImportCheck(in, "polynomial header");
  if (degree == -1) return;
  if (degree < -1 || (size_t)degree >= kMaxImportItems) Error("Import: polynomial degree out of range");
  poly.rep.v.resize((size_t)degree + 1);
  for (int i = 0; i <= degree; i++) Import(in, poly.rep.v[i]);
  poly.normalize();
}
void Export(ofstream &out, const vec_long &vec) {  // Serialization.cpp:83-89
  uint32_t len = vec.length();
  Export(out, len);
  for (long i = 0; i < vec.length(); i++) Export(out, vec[i]);
}
void Import(ifstream &in, vec_long &vec) {
  uint32_t size = 0;
  Import(in, size);
  ImportCheck(in, "vector header");
  if ((size_t)size > kMaxImportItems) Error("Import: vector length out of range");
  vec.SetLength(size);
  for (uint32_t i = 0; i < size; i++) Import(in, vec[i]);
  ImportCheck(in, "vector");
}
void Export(ofstream &out, const DoubleCRT &poly) {  // Serialization.cpp:56-65
  vector<vector<long>> rows = poly.getRows();
  uint32_t size = rows.size();
  Export(out, size);
  for (long i = 0; i < (long)rows.size(); ++i) {
    Export(out, i);
    vec_long v;
    v.v = rows[i];
    Export(out, v);
  }
}
void Import(ifstream &in, DoubleCRT &poly) {
  uint32_t size;
  Import(in, size);
  vector<vector<long>> rows(size);
  for (unsigned i = 0; i < size; i++) {
    long key;
    Import(in, key);
    vec_long v;
    Import(in, v);
    if (key < 0 || key >= (long)size) Error("Import(DoubleCRT): bad row index");
    rows[key] = v.v;
  }
  poly.setRows(rows);
}
void Export(ofstream &out, const CiphertextPart &part) { Export(out, part.poly); }
void Import(ifstream &in, CiphertextPart &part) { Import(in, part.poly); }
void Export(ofstream &out, const Ciphertext &ctxt) {  // Serialization.cpp:109-114
  Ciphertext copy = ctxt;
  copy.ScaleDown();
  copy.SyncHost();
  Export(out, copy.parts);
}
void Import(ifstream &in, Ciphertext &ctxt) {
  ctxt.Clear();
  Import(in, ctxt.parts);
  ctxt.UploadHost();
}

// ------------------------------------------------------------------------------- C entry: key generation
// Host-side key generation for callers that drive the C ABI directly (bench.py,
// apps/regression_sharded.py through ctypes): the same FHESISecKey / FHESIPubKey / KeySwitchSI code
// paths as the C++ clients, results as `poly`-format words ready for fhesi_key_create /
// fhesi_ksw_create.  Pure host work; no device is touched.
extern "C" int fhesih_keygen(uint32_t m, uint32_t logQ, uint64_t p, uint32_t g, uint32_t decompSize, uint64_t xi,
                             uint64_t seed, uint32_t n_rot, const uint32_t *rot_k, uint32_t *sk_words,
                             uint32_t *pk_words, uint32_t *ks_b, uint32_t *ks_A, uint32_t *rot_b, uint32_t *rot_A) {
  FHEcontext *saved = activeContext;
  struct HostOnly {  // no device context for this FHEcontext
    bool was = g_eagerDevice;
    HostOnly() { g_eagerDevice = false; }
    ~HostOnly() { g_eagerDevice = was; }
  } hostOnly;
  const bool timing = getenv("FHESIH_TIMING") != nullptr;  // phase times on stderr, for tuning
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = now(), t1, t2, t3, t4;
  {
    FHEcontext context(m, logQ, to_ZZ((unsigned long)p), g, decompSize);
    activeContext = &context;
    context.SetUpSIContext((long)xi);
    SetSeed(ZZ((unsigned long)seed));
    const unsigned n = context.zMstar.phiM(), W = context.Words(), D = context.ndigits;
    const size_t pw = (size_t)n * W;
    auto put = [&](uint32_t *dst, const DoubleCRT &d) {
      ZZX poly;
      d.toPoly(poly);
      ReduceCoefficients(poly, logQ);
      std::vector<uint32_t> w = PackPoly(poly, n, W);
      memcpy(dst, w.data(), pw * 4);
    };
    t1 = now();
    FHESISecKey sk(context);
    FHESIPubKey pk(sk);
    KeySwitchSI ks(sk);
    t2 = now();
    std::vector<std::pair<uint32_t *, const DoubleCRT *>> jobs;
    for (int i = 0; i < 2; ++i) {
      jobs.emplace_back(sk_words + i * pw, &sk.GetRepresentation()[i]);
      jobs.emplace_back(pk_words + i * pw, &pk.GetRepresentation()[i]);
    }
    for (unsigned k = 0; k < 3 * D; ++k) {
      jobs.emplace_back(ks_b + k * pw, &ks.GetRepresentation()[0][k]);
      jobs.emplace_back(ks_A + k * pw, &ks.GetRepresentation()[1][k]);
    }
    std::vector<std::unique_ptr<KeySwitchSI>> rks;
    for (uint32_t r = 0; r < n_rot; ++r) {
      rks.emplace_back(new KeySwitchSI(sk, rot_k[r]));
      for (unsigned k = 0; k < 2 * D; ++k) {
        jobs.emplace_back(rot_b + ((size_t)r * 2 * D + k) * pw, &rks.back()->GetRepresentation()[0][k]);
        jobs.emplace_back(rot_A + ((size_t)r * 2 * D + k) * pw, &rks.back()->GetRepresentation()[1][k]);
      }
    }
    t3 = now();
    ParallelFor(jobs.size(), [&](size_t j) { put(jobs[j].first, *jobs[j].second); });
    t4 = now();
  }
  if (timing)
    fprintf(stderr, "fhesih_keygen: context %.4f  sk+pk+s2 matrix %.4f  rotation matrices %.4f  pack %.4f  s\n", t1 - t0,
            t2 - t1, t3 - t2, t4 - t3);
  activeContext = saved;
  return 0;
}

// ---- flat draws: the same stream positions and values as SampleRandom / sampleGaussian, written straight
// into the arrays the device consumes (no ZZ / ZZX temporaries).
// SampleRandom(poly, 2^k, n) (Util.cpp:49-56): per coefficient RandomBits(k) - 2^(k-1), as W-word two's
// complement: subtracting 2^(k-1) mod 2^k flips bit k-1, and the flipped bit is the sign.
static void SampleRandomWords(uint32_t *dst, unsigned n, unsigned W, unsigned k, RandomStream &rs) {
  const size_t words64 = (k + 63) / 64, limbs = (k + 31) / 32, top = (k - 1) / 32;
  const uint32_t topbit = 1u << ((k - 1) % 32);
  if (limbs != W) Error("SampleRandomWords: word count does not match logQ");
  std::vector<uint32_t> w(2 * words64);
  for (unsigned i = 0; i < n; ++i) {
    for (size_t t = 0; t < words64; ++t) {
      const uint64_t x = rs.next64();
      w[2 * t] = (uint32_t)x, w[2 * t + 1] = (uint32_t)(x >> 32);
    }
    if (k % 32) w[limbs - 1] &= (1u << (k % 32)) - 1;
    w[top] ^= topbit;
    if (w[top] & topbit) w[top] |= ~(topbit | (topbit - 1));  // sign-extend inside the top word
    memcpy(dst + (size_t)i * W, w.data(), W * 4);
  }
}
// One KeySwitchSI::Init worth of draws (FHE-SI.cpp:174-189: per entry SampleRandom, then sampleGaussian)
// raw: [K][n + 1] uniforms of the Gaussians, turned into e by GaussianFromRaw afterwards
static void DrawMatrixFlat(uint32_t *A, uint32_t *raw, size_t K, unsigned n, unsigned W, unsigned logQ) {
  for (size_t k = 0; k < K; ++k) {
    SampleRandomWords(A + k * (size_t)n * W, n, W, logQ);
    SampleGaussianRaw(raw + k * ((size_t)n + 1), n);
  }
}
// sampleHWt (NumbTh.cpp:340-359) on a plain array: the same draws, the same polynomial
static void SampleHWtSmall(int32_t *out, long Hwt, long n) {
  std::fill(out, out + n, 0);
  if (Hwt > n) Hwt = n;
  for (long i = 0; i < Hwt;) {
    const long u = RandomLibc31() % n;
    if (out[u] == 0) {
      out[u] = (int32_t)((RandomLibc31() & 2) - 1);
      i++;
    }
  }
}
// a * b mod Phi_m and a(X^k) mod Phi_m for small-coefficient polynomials when m = 2h, h odd:
// X^h = -1 and Phi_m = sum_{i<h} (-X)^i, so the X^(h-1) term folds as w_j = v_j - (-1)^j v_(h-1)
static void FoldTopSmall(std::vector<int64_t> &v, int32_t *out, unsigned n) {
  const int64_t top = v[n];
  for (unsigned j = 0; j < n; ++j) out[j] = (int32_t)((j & 1) ? v[j] + top : v[j] - top);
}
static void MulSmallPhim(const int32_t *a, const int32_t *b, int32_t *out, unsigned n) {
  const unsigned h = n + 1;
  std::vector<int64_t> v(h, 0);
  for (unsigned i = 0; i < n; ++i) {
    if (!a[i]) continue;
    for (unsigned j = 0; j < n; ++j) {
      if (!b[j]) continue;
      const unsigned e = i + j;
      if (e >= h) v[e - h] -= (int64_t)a[i] * b[j];
      else v[e] += (int64_t)a[i] * b[j];
    }
  }
  FoldTopSmall(v, out, n);
}
static void AutomorphSmallPhim(const int32_t *a, int32_t *out, unsigned n, unsigned m, unsigned k) {
  const unsigned h = n + 1;
  std::vector<int64_t> v(h, 0);
  for (unsigned i = 0; i < n; ++i) {
    if (!a[i]) continue;
    unsigned e = (unsigned)(((uint64_t)i * k) % m);
    if (e >= h) v[e - h] -= a[i];
    else v[e] += a[i];
  }
  FoldTopSmall(v, out, n);
}
// Any other m: the linear product / the index map, then the remainder by Phi_m (RemPhim's general branch)
static void RemSmallGeneral(const std::vector<int64_t> &v, int32_t *out, const PAlgebra &zms) {
  ZZX a;
  a.rep.v.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) a.rep.v[i] = ZZ((long)v[i]);
  a.normalize();
  RemPhim(a, zms);
  const unsigned n = zms.phiM();
  for (unsigned j = 0; j < n; ++j) out[j] = (long)j <= deg(a) ? (int32_t)to_long(a.rep.v[j]) : 0;
}
static void MulSmallGeneral(const int32_t *a, const int32_t *b, int32_t *out, const PAlgebra &zms) {
  const unsigned n = zms.phiM();
  std::vector<int64_t> v(2 * (size_t)n - 1, 0);
  for (unsigned i = 0; i < n; ++i)
    if (a[i])
      for (unsigned j = 0; j < n; ++j) v[i + j] += (int64_t)a[i] * b[j];
  RemSmallGeneral(v, out, zms);
}
static void AutomorphSmallGeneral(const int32_t *a, int32_t *out, const PAlgebra &zms, unsigned k) {
  const unsigned n = zms.phiM(), m = zms.M();
  std::vector<int64_t> v(m, 0);
  for (unsigned i = 0; i < n; ++i) v[(unsigned)(((uint64_t)i * k) % m)] += a[i];
  RemSmallGeneral(v, out, zms);
}
// The whole set-up's draws in the layout fhesi_keygen_batch consumes: secret key; public key (its Gaussian,
// then its uniform polynomial: FHE-SI.cpp:44-47) stored as the LAST entry; the s^2 -> s matrix (3 source rows
// 1, s, s^2) preceded by the throw-away key FHE-SI.cpp:222 samples; one rotation matrix per rot_k (2 source
// rows 1, s(X^k)), each likewise preceded by its throw-away key (:233).  Same seed => the same keys as the
// C++ classes (fhesih_keygen) generate.  sk_out int32 [n]; src int32 [3 + 2 n_rot][n];
// A uint32 [(3 + 2 n_rot) D + 1][n][W]; e int32 [(3 + 2 n_rot) D + 1][n].
extern "C" int fhesih_keydraws_flat(uint32_t m, uint32_t logQ, uint64_t p, uint32_t g, uint32_t decompSize, uint64_t xi,
                                    uint64_t seed, uint32_t n_rot, const uint32_t *rot_k, int32_t *sk_out, int32_t *src,
                                    uint32_t *A, int32_t *e) {
  FHEcontext *saved = activeContext;
  struct HostOnly {
    bool was = g_eagerDevice;
    HostOnly() { g_eagerDevice = false; }
    ~HostOnly() { g_eagerDevice = was; }
  } hostOnly;
  {
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool timing = getenv("FHESIH_TIMING") != nullptr;
    const double T0 = now();
    FHEcontext context(m, logQ, to_ZZ((unsigned long)p), g, decompSize);
    activeContext = &context;
    context.SetUpSIContext((long)xi);
    const double T1 = now();
    SetSeed(ZZ((unsigned long)seed));
    const unsigned n = context.zMstar.phiM(), W = context.Words(), D = context.ndigits;
    const size_t pw = (size_t)n * W, Km = (size_t)(3 + 2 * n_rot) * D;
    std::vector<uint32_t> raw((Km + 1) * ((size_t)n + 1));
    const bool twoh = m % 2 == 0 && (m / 2) % 2 == 1 && n + 1 == m / 2;  // m = 2h, h an odd prime (every device context)
    std::vector<int32_t> tmp(n);
    // The long draws -- a uniform polynomial and a Gaussian's uniforms per matrix entry -- have a fixed length in
    // stream words, and the ChaCha20 stream is seekable: the walk below only runs the short, data-dependent draws
    // (the Hamming-weight keys) and notes where every long one starts; they are then filled on all cores, each from
    // its own position of the same stream.  Same words as the sequential walk (tests/test_host_cpp.py); the
    // SplitMix test stream, FHESIH_SEQ_DRAWS=1 and the one-in-2^28 rejected Gaussian word take the sequential walk.
    RandomStream &gs = GlobalRandomStream();
    const uint64_t pos_start = gs.test ? 0 : gs.position();
    struct Seg { uint32_t *A, *raw; uint64_t pos; bool gauss_first; };
    std::vector<Seg> segs;
    const uint64_t lenA = (uint64_t)n * ((logQ + 63) / 64), lenG = 2 * (((uint64_t)n + 1) / 2);
    const char *seq_env = getenv("FHESIH_SEQ_DRAWS");
    bool parallel = !gs.test && !(seq_env && *seq_env);
    for (int pass = 0; pass < 2; ++pass) {
      uint64_t pos = 0;
      auto long_draws = [&](uint32_t *Ad, uint32_t *rawd, size_t K) {  // K entries: uniform polynomial, then Gaussian
        if (!parallel) return DrawMatrixFlat(Ad, rawd, K, n, W, logQ);
        for (size_t k = 0; k < K; ++k, pos += lenA + lenG) segs.push_back(Seg{Ad + k * pw, rawd + k * ((size_t)n + 1), pos, false});
      };
      auto short_draw = [&](int32_t *out) {
        if (parallel) gs.seek(pos);
        SampleHWtSmall(out, 64, n);
        if (parallel) pos = gs.position();
      };
      segs.clear();
      if (parallel) pos = gs.position();
      short_draw(sk_out);                                       // FHESISecKey::Init
      if (parallel) {                                           // FHESIPubKey::Init: c0's Gaussian, then c1
        segs.push_back(Seg{A + Km * pw, raw.data() + Km * ((size_t)n + 1), pos, true});
        pos += lenA + lenG;
      } else {
        SampleGaussianRaw(raw.data() + Km * ((size_t)n + 1), n);
        SampleRandomWords(A + Km * pw, n, W, logQ);
      }
      {                                                         // KeySwitchSI(sk): InitS2
        short_draw(tmp.data());                                 // the throw-away key of FHE-SI.cpp:222
        std::fill(src, src + n, 0);
        src[0] = 1;
        memcpy(src + n, sk_out, n * 4);
        if (twoh) MulSmallPhim(sk_out, sk_out, src + 2 * (size_t)n, n);
        else MulSmallGeneral(sk_out, sk_out, src + 2 * (size_t)n, context.zMstar);
        long_draws(A, raw.data(), 3 * (size_t)D);
      }
      for (uint32_t r = 0; r < n_rot; ++r) {                    // KeySwitchSI(sk, k): InitAutomorph
        short_draw(tmp.data());                                 // :233
        int32_t *rs = src + (3 + 2 * (size_t)r) * n;
        std::fill(rs, rs + n, 0);
        rs[0] = 1;                                              // 1(X^k) = 1
        if (twoh) AutomorphSmallPhim(sk_out, rs + n, n, m, rot_k[r]);
        else AutomorphSmallGeneral(sk_out, rs + n, context.zMstar, rot_k[r]);
        const size_t off = (3 + 2 * (size_t)r) * D;
        long_draws(A + off * pw, raw.data() + off * ((size_t)n + 1), 2 * (size_t)D);
      }
      if (!parallel) break;
      gs.seek(pos);  // the stream continues after the last long draw
      std::atomic<bool> ok{true};
      const RandomStream keyed = gs;  // same key; every worker seeks its own copy
      ParallelFor(segs.size(), [&](size_t i) {
        RandomStream t = keyed;
        t.seek(segs[i].pos);
        bool good = true;
        if (segs[i].gauss_first) {
          good = SampleGaussianRawNoRetry(t, segs[i].raw, n);
          SampleRandomWords(segs[i].A, n, W, logQ, t);
        } else {
          SampleRandomWords(segs[i].A, n, W, logQ, t);
          good = SampleGaussianRawNoRetry(t, segs[i].raw, n);
        }
        if (!good) ok = false;
      });
      if (ok && !getenv("FHESIH_TEST_REJECT")) break;  // (the variable: a test's way to take the next two lines)
      parallel = false;  // a rejected word shifts everything after it: walk the stream in order instead
      gs.seek(pos_start);
    }
    const double T2 = now();
    // the floating-point half of every Gaussian polynomial, on all cores
    const double stdev = context.stdev;
    ParallelFor(Km + 1, [&](size_t k) { GaussianFromRaw(e + k * n, raw.data() + k * ((size_t)n + 1), n, stdev); });
    if (timing) fprintf(stderr, "keydraws_flat: context %.4f  draws %.4f  gaussian math %.4f\n", T1 - T0, T2 - T1, now() - T2);
  }
  activeContext = saved;
  return 0;
}

// The random draws of fhesih_keygen without the key-switch arithmetic: same objects, same order and
// number of draws (secret key, public key, s^2 matrix incl. its throw-away key, one rotation matrix per
// rot_k), with KeySwitchSI::Init diverted into a sink.  The caller passes the draws to
// fhesi_ksw_generate, which does b = A t + e + src 2^(24 j) on the device.  sk: int32 [n] (the
// polynomial s); pk_words: [2][n][W]; per matrix: src int32 [parts][n], A uint32 [parts*D][n][W],
// e int32 [parts*D][n] (parts = 3 for s^2, 2 for rotations).
extern "C" int fhesih_keydraws(uint32_t m, uint32_t logQ, uint64_t p, uint32_t g, uint32_t decompSize, uint64_t xi,
                               uint64_t seed, uint32_t n_rot, const uint32_t *rot_k, int32_t *sk_out,
                               uint32_t *pk_words, int32_t *s2_src, uint32_t *s2_A, int32_t *s2_e, int32_t *rot_src,
                               uint32_t *rot_A, int32_t *rot_e) {
  FHEcontext *saved = activeContext;
  struct HostOnly {
    bool was = g_eagerDevice;
    HostOnly() { g_eagerDevice = false; }
    ~HostOnly() { g_eagerDevice = was; }
  } hostOnly;
  KeyDrawSink sink;
  {
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool timing = getenv("FHESIH_TIMING") != nullptr;
    const double T0 = now();
    FHEcontext context(m, logQ, to_ZZ((unsigned long)p), g, decompSize);
    activeContext = &context;
    const double T1 = now();
    context.SetUpSIContext((long)xi);
    const double T2 = now();
    SetSeed(ZZ((unsigned long)seed));
    const unsigned n = context.zMstar.phiM(), W = context.Words(), D = context.ndigits;
    const size_t pw = (size_t)n * W;
    FHESISecKey sk(context);
    FHESIPubKey pk(sk);
    const double T3 = now();
    g_drawSink = &sink;
    KeySwitchSI ks(sk);
    for (uint32_t r = 0; r < n_rot; ++r) KeySwitchSI rk(sk, rot_k[r]);
    g_drawSink = nullptr;
    if (timing) fprintf(stderr, "keydraws: context %.4f setup %.4f sk+pk %.4f matrices %.4f\n", T1 - T0, T2 - T1, T3 - T2, now() - T3);
    auto small = [&](int32_t *dst, const ZZX &a) {
      for (long i = 0; i < (long)n; ++i) dst[i] = i <= deg(a) ? (int32_t)to_long(a.rep.v[i]) : 0;
    };
    ZZX s;
    sk.GetRepresentation()[1].toPoly(s);
    small(sk_out, s);
    for (int i = 0; i < 2; ++i) {
      ZZX poly;
      pk.GetRepresentation()[i].toPoly(poly);
      ReduceCoefficients(poly, logQ);
      std::vector<uint32_t> w = PackPoly(poly, n, W);
      memcpy(pk_words + i * pw, w.data(), pw * 4);
    }
    for (size_t mi = 0; mi < sink.matrices.size(); ++mi) {
      const KeyDrawSink::Matrix &M = sink.matrices[mi];
      const size_t parts = M.src.size(), K = parts * D;
      int32_t *src = mi == 0 ? s2_src : rot_src + (mi - 1) * parts * n;
      uint32_t *A = mi == 0 ? s2_A : rot_A + (mi - 1) * K * pw;
      int32_t *e = mi == 0 ? s2_e : rot_e + (mi - 1) * K * n;
      for (size_t i = 0; i < parts; ++i) small(src + i * n, M.src[i]);
      ParallelFor(K, [&](size_t k) {
        std::vector<uint32_t> w = PackPoly(M.polys[k], n, W);
        memcpy(A + k * pw, w.data(), pw * 4);
        small(e + k * n, M.errs[k]);
      });
    }
  }
  activeContext = saved;
  return 0;
}
