// ntl_compat.h -- TEST INFRASTRUCTURE.  The rest of the NTL interface that the reference's own
// library sources use (PlaintextSpace.cpp, CModulus.cpp, FHEContext.cpp, PAlgebra.cpp,
// SingleCRT.cpp, DoubleCRT.cpp, NumbTh.cpp, bluestein.cpp, IndexSet.cpp, Plaintext.cpp, Util.cpp,
// FHE-SI.cpp, Ciphertext.cpp, Serialization.cpp, Matrix.cpp), on top of the client-facing slice in
// fhe-si_b200/host/ntl_shim.h.  With it oracle/Makefile compiles those sources, unmodified and
// where they lie under /root/reference, into oracle/_ref -- the reference's own DoubleCRT /
// Bluestein / Ciphertext / key-switch code then runs here, and its outputs pin the oracle
// (tests/golden/make_ref_golden.py).  NTL proper and GMP are absent from this image; this header is
// written from scratch against NTL's documented semantics and contains no NTL code.
#pragma once
#include "../../fhe-si_b200/host/ntl_shim.h"
#include <array>
#include <map>

#ifndef NTL_CLIENT
#define NTL_CLIENT using namespace std; using namespace NTL;
#endif

namespace NTL {

enum INIT_VAL_TYPE { INIT_VAL };

// ------------------------------------------------------------------------------- ZZ extras
inline ZZ abs(const ZZ &a) { ZZ r = a; r.neg = false; return r; }
inline long bit(const ZZ &a, long i) {
  return (i >= 0 && (size_t)(i / 32) < a.mag.size()) ? (a.mag[i / 32] >> (i % 32)) & 1 : 0;
}
// NTL: bitwise operators act on magnitudes, the result is non-negative
inline ZZ &operator&=(ZZ &a, const ZZ &b) {
  size_t n = std::min(a.mag.size(), b.mag.size());
  a.mag.resize(n);
  for (size_t i = 0; i < n; ++i) a.mag[i] &= b.mag[i];
  a.neg = false;
  a.trim();
  return a;
}
inline ZZ &operator|=(ZZ &a, const ZZ &b) {
  if (a.mag.size() < b.mag.size()) a.mag.resize(b.mag.size(), 0);
  for (size_t i = 0; i < b.mag.size(); ++i) a.mag[i] |= b.mag[i];
  a.neg = false;
  a.trim();
  return a;
}
inline ZZ &operator^=(ZZ &a, const ZZ &b) {
  if (a.mag.size() < b.mag.size()) a.mag.resize(b.mag.size(), 0);
  for (size_t i = 0; i < b.mag.size(); ++i) a.mag[i] ^= b.mag[i];
  a.neg = false;
  a.trim();
  return a;
}
inline ZZ operator&(const ZZ &a, const ZZ &b) { ZZ r = a; r &= b; return r; }
inline ZZ operator|(const ZZ &a, const ZZ &b) { ZZ r = a; r |= b; return r; }
inline ZZ operator^(const ZZ &a, const ZZ &b) { ZZ r = a; r ^= b; return r; }
inline void RandomBits(ZZ &x, long l) { x = RandomBits_ZZ(l); }
inline long RandomBits_long(long l) { return to_long(RandomBits_ZZ(l)); }
inline unsigned long RandomBits_ulong(long l) { return (unsigned long)RandomBits_long(l); }
inline long power_long(long a, long e) {
  long r = 1;
  while (e-- > 0) r *= a;
  return r;
}
inline ZZ SqrRoot(const ZZ &a) {  // floor(sqrt(a)), Newton
  if (a <= 0L) return ZZ();
  ZZ x = ZZ(1L) << (long)((a.bits() + 1) / 2), y;
  for (;;) {
    y = (x + a / x) >> 1;
    if (y >= x) return x;
    x = y;
  }
}
inline long SqrRoot(long a) { return to_long(SqrRoot(ZZ(a))); }
inline ZZ GCD(const ZZ &a, const ZZ &b) {
  ZZ x = abs(a), y = abs(b);
  while (!y.is_zero()) {
    ZZ t = x % y;
    x = y;
    y = t;
  }
  return x;
}
inline void conv(ZZ &x, int a) { x = ZZ((long)a); }
inline void conv(ZZ &x, unsigned a) { x = ZZ((unsigned long)a); }
inline void conv(ZZ &x, unsigned long a) { x = ZZ(a); }
inline void conv(double &x, const ZZ &a) { x = to_double(a); }
inline void conv(long &x, long a) { x = a; }
inline void conv(int &x, const ZZ &a) { x = (int)to_long(a); }
inline long to_int(const ZZ &a) { return to_long(a); }
inline void set(ZZ &x) { x = ZZ(1L); }
inline void negate(ZZ &x, const ZZ &a) { x = -a; }
inline void DivRem(ZZ &q, ZZ &r, const ZZ &a, const ZZ &b) { ZZ::DivRem(q, r, a, b); }
inline void div(ZZ &q, const ZZ &a, const ZZ &b) { q = a / b; }
inline void div(ZZ &q, const ZZ &a, long b) { q = a / ZZ(b); }
inline void rem(ZZ &r, const ZZ &a, long b) { r = a % ZZ(b); }
inline ZZ MulMod(const ZZ &a, const ZZ &b, const ZZ &n) { return (a * b) % n; }
inline ZZ AddMod(const ZZ &a, const ZZ &b, const ZZ &n) { return (a + b) % n; }
inline ZZ SubMod(const ZZ &a, const ZZ &b, const ZZ &n) { return (a - b) % n; }
inline void MulMod(ZZ &x, const ZZ &a, const ZZ &b, const ZZ &n) { x = (a * b) % n; }
inline void AddMod(ZZ &x, const ZZ &a, const ZZ &b, const ZZ &n) { x = (a + b) % n; }
inline void SubMod(ZZ &x, const ZZ &a, const ZZ &b, const ZZ &n) { x = (a - b) % n; }
inline void InvMod(ZZ &x, const ZZ &a, const ZZ &n) { x = InvMod(a, n); }
inline void PowerMod(ZZ &x, const ZZ &a, const ZZ &e, const ZZ &n) { x = PowerMod(a, e, n); }
inline ZZ PowerMod(const ZZ &a, long e, const ZZ &n) { return PowerMod(a, ZZ(e), n); }

// single-precision multiplication with a "preconditioned" operand: plain MulMod here
typedef long mulmod_precon_t;
inline mulmod_precon_t PrepMulModPrecon(long, long) { return 0; }
inline long MulModPrecon(long a, long b, long n, mulmod_precon_t) { return MulMod(a, b, n); }

// the small primes in increasing order
class PrimeSeq {
  long cur = 1;

 public:
  void reset(long b) { cur = b > 1 ? b - 1 : 1; }
  long next() {
    for (++cur;; ++cur) {
      bool ok = cur >= 2;
      for (long d = 2; ok && d * d <= cur; ++d)
        if (cur % d == 0) ok = false;
      if (ok) return cur;
    }
  }
};

// ------------------------------------------------------------------------------- ZZX extras
inline void trunc(ZZX &x, const ZZX &a, long m) {
  ZZX r;
  for (long i = 0; i < m && i <= deg(a); ++i) r.rep.v.push_back(a.rep.v[i]);
  r.normalize();
  x = r;
}
inline ZZX trunc(const ZZX &a, long m) { ZZX r; trunc(r, a, m); return r; }
inline void SetX(ZZX &x) { x = ZZX(); SetCoeff(x, 1, 1L); }
inline void set(ZZX &x) { x = ZZX(1L); }
inline void conv(ZZX &x, long a) { x = ZZX(a); }
inline void conv(ZZX &x, const ZZ &a) { x = to_ZZX(a); }
inline void conv(ZZX &x, const ZZX &a) { x = a; }
inline ZZX operator+(const ZZX &a, const ZZ &c) { ZZX r = a; SetCoeff(r, 0, coeff(r, 0) + c); return r; }
inline ZZX operator+(const ZZX &a, long c) { return a + ZZ(c); }
inline ZZX operator-(const ZZX &a, const ZZ &c) { return a + (-c); }
inline ZZX operator-(const ZZX &a, long c) { return a + ZZ(-c); }
inline ZZX &operator+=(ZZX &a, const ZZ &c) { SetCoeff(a, 0, coeff(a, 0) + c); return a; }
inline ZZX &operator+=(ZZX &a, long c) { return a += ZZ(c); }
inline ZZX &operator-=(ZZX &a, const ZZ &c) { SetCoeff(a, 0, coeff(a, 0) - c); return a; }
inline ZZX &operator-=(ZZX &a, long c) { return a -= ZZ(c); }
inline void mul(ZZX &x, const ZZX &a, const ZZX &b) { x = a * b; }
inline void mul(ZZX &x, const ZZX &a, const ZZ &b) { x = a * b; }
inline void mul(ZZX &x, const ZZX &a, long b) { x = a * b; }
inline void add(ZZX &x, const ZZX &a, const ZZX &b) { x = a + b; }
inline void sub(ZZX &x, const ZZX &a, const ZZX &b) { x = a - b; }
inline void negate(ZZX &x, const ZZX &a) { x = -a; }
inline void add(ZZX &x, const ZZX &a, const ZZ &b) { x = a + b; }
inline void sub(ZZX &x, const ZZX &a, const ZZ &b) { x = a - b; }
inline ZZX LeftShift(const ZZX &a, long n) {
  ZZX r;
  if (IsZero(a)) return r;
  r.rep.v.assign(n, ZZ());
  r.rep.v.insert(r.rep.v.end(), a.rep.v.begin(), a.rep.v.end());
  return r;
}
// exact or pseudo division is not needed: divisors here are monic (cyclotomics, X^k - 1)
inline void DivRem(ZZX &q, ZZX &r, const ZZX &a, const ZZX &b) {
  const long db = deg(b);
  if (db < 0) Error("ZZX DivRem: division by zero");
  ZZX t = a;
  q = ZZX();
  const ZZ &lc = LeadCoeff(b);
  for (long i = deg(t); i >= db; --i) {
    ZZ c = coeff(t, i);
    if (c.is_zero()) continue;
    if (!IsOne(lc)) {
      ZZ qq, rr;
      ZZ::DivRem(qq, rr, c, lc);
      if (!rr.is_zero()) Error("ZZX DivRem: non-exact leading coefficient");
      c = qq;
    }
    SetCoeff(q, i - db, c);
    for (long j = 0; j <= db; ++j) t.rep.v[i - db + j] -= c * b.rep.v[j];
  }
  t.normalize();
  r = t;
}
inline void div(ZZX &q, const ZZX &a, const ZZX &b) { ZZX r; DivRem(q, r, a, b); }
inline ZZX operator/(const ZZX &a, const ZZX &b) { ZZX q, r; DivRem(q, r, a, b); return q; }
inline ZZX &operator/=(ZZX &a, const ZZX &b) { a = a / b; return a; }
inline ZZX &operator%=(ZZX &a, const ZZX &b) { a = a % b; return a; }
inline long divide(ZZX &q, const ZZX &a, const ZZX &b) {
  ZZX r;
  DivRem(q, r, a, b);
  return IsZero(r);
}
inline void MulMod(ZZX &x, const ZZX &a, const ZZX &b, const ZZX &f) { x = (a * b) % f; }
inline ZZX MulMod(const ZZX &a, const ZZX &b, const ZZX &f) { return (a * b) % f; }
inline ZZX diff(const ZZX &a) {
  ZZX r;
  for (long i = 1; i <= deg(a); ++i) SetCoeff(r, i - 1, coeff(a, i) * ZZ(i));
  return r;
}
inline const ZZ &ConstTerm(const ZZX &a) { return coeff(a, 0); }
inline std::istream &operator>>(std::istream &is, ZZX &a) {
  a = ZZX();
  char ch;
  if (!(is >> ch) || ch != '[') return is;
  for (long i = 0;; ++i) {
    is >> std::ws;
    if (is.peek() == ']') {
      is.get();
      break;
    }
    ZZ c;
    if (!(is >> c)) break;
    SetCoeff(a, i, c);
  }
  a.normalize();
  return is;
}

// ------------------------------------------------------------------------------- ZZ_p / ZZ_pX extras
class ZZ_pContext {
  long p = 0;
  ZZ pz;

 public:
  ZZ_pContext() {}
  explicit ZZ_pContext(const ZZ &q) : p(to_long(q)), pz(q) {}
  void save() { p = ZZ_p::mod(), pz = ZZ_p::modZZ(); }
  void restore() const {
    if (p) ZZ_p::init(pz);
  }
};
typedef ZZ_pContext ZZ_pBak_c;
inline bool operator!=(const ZZ_p &a, long b) { return !(a == b); }
inline void conv(ZZ_p &x, long a) { x = ZZ_p(a); }
inline void conv(ZZ_p &x, int a) { x = ZZ_p((long)a); }
inline void conv(ZZ_p &x, const ZZ &a) { x = to_ZZ_p(a); }
inline void conv(ZZ &x, const ZZ_p &a) { x = rep(a); }
inline void set(ZZ_p &x) { x = ZZ_p(1); }
inline void clear(ZZ_p &x) { x = ZZ_p(0); }
inline bool IsOne(const ZZ_p &a) { return a.v == 1 % ZZ_p::mod(); }
inline void power(ZZ_p &x, const ZZ_p &a, long e) { x = e >= 0 ? power(a, e) : power(inv(a), -e); }
inline void power(ZZ_p &x, const ZZ_p &a, const ZZ &e) { x = ZZ_p(to_long(PowerMod(ZZ(a.v), e, ZZ_p::modulus()))); }
inline ZZ_p power(const ZZ_p &a, const ZZ &e) { ZZ_p r; power(r, a, e); return r; }
inline void inv(ZZ_p &x, const ZZ_p &a) { x = inv(a); }
inline ZZ_p operator/(const ZZ_p &a, const ZZ_p &b) { return a * inv(b); }
inline ZZ_p &operator/=(ZZ_p &a, const ZZ_p &b) { a = a * inv(b); return a; }
inline ZZ_p operator+(const ZZ_p &a, long b) { return a + ZZ_p(b); }
inline ZZ_p operator-(const ZZ_p &a, long b) { return a - ZZ_p(b); }
inline ZZ_p operator*(const ZZ_p &a, long b) { return a * ZZ_p(b); }
inline ZZ_p operator*(long a, const ZZ_p &b) { return ZZ_p(a) * b; }
inline void mul(ZZ_p &x, const ZZ_p &a, const ZZ_p &b) { x = a * b; }
inline void add(ZZ_p &x, const ZZ_p &a, const ZZ_p &b) { x = a + b; }
inline void sub(ZZ_p &x, const ZZ_p &a, const ZZ_p &b) { x = a - b; }
inline void negate(ZZ_p &x, const ZZ_p &a) { x = -a; }

inline void conv(ZZ_pX &x, long a) { x = to_ZZ_pX(a); }
inline void conv(ZZ_pX &x, const ZZ_p &a) { x = to_ZZ_pX(a); }
inline void conv(ZZ_pX &x, const ZZ &a) { x = to_ZZ_pX(a); }
inline void set(ZZ_pX &x) { x = to_ZZ_pX(1L); }
inline void SetX(ZZ_pX &x) { x = ZZ_pX(); SetCoeff(x, 1); }
inline bool IsOne(const ZZ_pX &a) { return deg(a) == 0 && IsOne(a.rep.v[0]); }
inline void mul(ZZ_pX &x, const ZZ_pX &a, const ZZ_pX &b) { x = a * b; }
inline void mul(ZZ_pX &x, const ZZ_pX &a, const ZZ_p &b) { x = a * b; }
inline void add(ZZ_pX &x, const ZZ_pX &a, const ZZ_pX &b) { x = a + b; }
inline void sub(ZZ_pX &x, const ZZ_pX &a, const ZZ_pX &b) { x = a - b; }
inline void div(ZZ_pX &q, const ZZ_pX &a, const ZZ_pX &b) { q = a / b; }
inline void negate(ZZ_pX &x, const ZZ_pX &a) { x = ZZ_pX() - a; }
inline ZZ_pX operator-(const ZZ_pX &a) { return ZZ_pX() - a; }
inline ZZ_pX operator+(const ZZ_pX &a, const ZZ_p &c) { ZZ_pX r = a; SetCoeff(r, 0, coeff(r, 0) + c); return r; }
inline ZZ_pX operator+(const ZZ_pX &a, long c) { return a + ZZ_p(c); }
inline ZZ_pX operator-(const ZZ_pX &a, const ZZ_p &c) { return a + (-c); }
inline ZZ_pX operator-(const ZZ_pX &a, long c) { return a + ZZ_p(-c); }
inline ZZ_pX &operator+=(ZZ_pX &a, const ZZ_p &c) { a = a + c; return a; }
inline ZZ_pX &operator-=(ZZ_pX &a, const ZZ_p &c) { a = a - c; return a; }
inline ZZ_pX &operator+=(ZZ_pX &a, long c) { a = a + c; return a; }
inline ZZ_pX &operator-=(ZZ_pX &a, long c) { a = a - c; return a; }
inline ZZ_pX &operator/=(ZZ_pX &a, const ZZ_p &c) { a *= inv(c); return a; }
inline ZZ_pX &operator/=(ZZ_pX &a, long c) { a *= inv(ZZ_p(c)); return a; }
inline void MakeMonic(ZZ_pX &a) {
  if (!IsZero(a)) a *= inv(LeadCoeff(a));
}
inline ZZ_pX GCD(const ZZ_pX &a, const ZZ_pX &b) {
  ZZ_pX x = a, y = b;
  while (!IsZero(y)) {
    ZZ_pX t = x % y;
    x = y;
    y = t;
  }
  MakeMonic(x);
  return x;
}
inline void GCD(ZZ_pX &d, const ZZ_pX &a, const ZZ_pX &b) { d = GCD(a, b); }
inline ZZ_pX diff(const ZZ_pX &a) {
  ZZ_pX r;
  for (long i = 1; i <= deg(a); ++i) SetCoeff(r, i - 1, coeff(a, i) * ZZ_p(i));
  r.normalize();
  return r;
}
inline void trunc(ZZ_pX &x, const ZZ_pX &a, long m) {
  ZZ_pX r;
  for (long i = 0; i < m && i <= deg(a); ++i) r.rep.v.push_back(a.rep.v[i]);
  r.normalize();
  x = r;
}
inline ZZ_pX PowerMod(const ZZ_pX &a, const ZZ &e, const ZZ_pX &f) {
  ZZ_pX r = to_ZZ_pX(1L), b = a % f;
  for (size_t i = 0; i < e.bits(); ++i) {
    if (bit(e, (long)i)) r = MulMod(r, b, f);
    b = MulMod(b, b, f);
  }
  return r;
}
inline ZZ_pX PowerMod(const ZZ_pX &a, long e, const ZZ_pX &f) { return PowerMod(a, ZZ(e), f); }
inline ZZ_pX PowerXMod(const ZZ &e, const ZZ_pX &f) { ZZ_pX x; SetX(x); return PowerMod(x, e, f); }
inline ZZ_pX PowerXMod(long e, const ZZ_pX &f) { return PowerXMod(ZZ(e), f); }
inline void eval(ZZ_p &y, const ZZ_pX &a, const ZZ_p &x) { y = eval(a, x); }
// a stand-in for NTL's precomputed-modulus type: plain remainders
class ZZ_pXModulus {
 public:
  ZZ_pX f;
  ZZ_pXModulus() {}
  ZZ_pXModulus(const ZZ_pX &ff) : f(ff) {}
  operator const ZZ_pX &() const { return f; }
  const ZZ_pX &val() const { return f; }
};
inline void build(ZZ_pXModulus &F, const ZZ_pX &f) { F.f = f; }
inline long deg(const ZZ_pXModulus &F) { return deg(F.f); }

// Factorisation of a monic square-free polynomial over Z_p into monic irreducibles
// (distinct-degree, then Cantor-Zassenhaus equal-degree splitting with the shared random stream).
// p is odd in every configuration of this repository.
inline void EDF(vec_ZZ_pX &factors, const ZZ_pX &f, long d) {
  factors.v.clear();
  std::vector<ZZ_pX> work{f};
  const ZZ p = ZZ_p::modulus();
  ZZ e = (power(p, d) - 1L) / 2L;
  while (!work.empty()) {
    ZZ_pX g = work.back();
    work.pop_back();
    if (deg(g) == d) {
      factors.v.push_back(g);
      continue;
    }
    for (;;) {
      // splitting polynomials come from a private stream, so that factoring Phi_m in the
      // FHEcontext constructor leaves the shared stream where the client seeded it
      static RandomStream local = [] {  // a fixed private SplitMix64 stream: factoring is deterministic
        RandomStream r;
        r.test = true;
        r.state = 0x5DEECE66Dull;
        return r;
      }();
      ZZ_pX r;
      r.rep.v.resize(deg(g));
      for (auto &c : r.rep.v) c = ZZ_p((long)(local.next64() % (uint64_t)ZZ_p::mod()));
      r.normalize();
      ZZ_pX h = GCD(g, PowerMod(r, e, g) - 1L);
      if (deg(h) > 0 && deg(h) < deg(g)) {
        work.push_back(h);
        work.push_back(g / h);
        break;
      }
    }
  }
}
inline void EDF(vec_ZZ_pX &factors, const ZZ_pXModulus &F, const ZZ_pX & /*X^p mod f*/, long d, long = 0) {
  EDF(factors, F.f, d);
}
inline void SFCanZass(vec_ZZ_pX &factors, const ZZ_pX &ff, long = 0) {
  factors.v.clear();
  ZZ_pX f = ff;
  MakeMonic(f);
  const ZZ p = ZZ_p::modulus();
  ZZ_pX x, h;
  SetX(x);
  h = x;
  for (long d = 1; 2 * d <= deg(f); ++d) {
    h = PowerMod(h, p, f);  // X^(p^d) mod f
    ZZ_pX g = GCD(f, h - x);
    if (deg(g) > 0) {
      vec_ZZ_pX part;
      EDF(part, g, d);
      for (auto &q : part.v) factors.v.push_back(q);
      f = f / g;
      h = h % f;
    }
  }
  if (deg(f) > 0) factors.v.push_back(f);
}
inline vec_ZZ_pX SFCanZass(const ZZ_pX &f, long verbose = 0) { vec_ZZ_pX r; SFCanZass(r, f, verbose); return r; }

// ------------------------------------------------------------------------------- zz_p / zz_pX
class zz_p {
 public:
  long v = 0;
  static long &mod() {
    static long m = 0;
    return m;
  }
  static void init(long q) { mod() = q; }
  static long modulus() { return mod(); }
  zz_p() {}
  explicit zz_p(long x) { v = ((x % mod()) + mod()) % mod(); }
  zz_p &operator=(long x) { return *this = zz_p(x); }
  zz_p &operator+=(const zz_p &o) { v = AddMod(v, o.v, mod()); return *this; }
  zz_p &operator-=(const zz_p &o) { v = SubMod(v, o.v, mod()); return *this; }
  zz_p &operator*=(const zz_p &o) { v = MulMod(v, o.v, mod()); return *this; }
  bool operator==(const zz_p &o) const { return v == o.v; }
  bool operator!=(const zz_p &o) const { return v != o.v; }
  bool operator==(long o) const { return v == zz_p(o).v; }
  bool operator!=(long o) const { return v != zz_p(o).v; }
};
inline zz_p operator+(zz_p a, const zz_p &b) { return a += b; }
inline zz_p operator-(zz_p a, const zz_p &b) { return a -= b; }
inline zz_p operator*(zz_p a, const zz_p &b) { return a *= b; }
inline zz_p operator-(const zz_p &a) { return zz_p(-a.v); }
inline zz_p inv(const zz_p &a) { return zz_p(InvMod(a.v, zz_p::mod())); }
inline zz_p operator/(const zz_p &a, const zz_p &b) { return a * inv(b); }
inline zz_p power(const zz_p &a, long e) {
  return e >= 0 ? zz_p(PowerMod(a.v, e, zz_p::mod())) : zz_p(PowerMod(InvMod(a.v, zz_p::mod()), -e, zz_p::mod()));
}
inline void power(zz_p &x, const zz_p &a, long e) { x = power(a, e); }
inline long rep(const zz_p &a) { return a.v; }
inline zz_p to_zz_p(long x) { return zz_p(x); }
inline zz_p to_zz_p(const ZZ &x) { return zz_p(to_long(x % ZZ(zz_p::mod()))); }
inline void conv(zz_p &x, long a) { x = zz_p(a); }
inline void conv(zz_p &x, int a) { x = zz_p((long)a); }
inline void conv(zz_p &x, const ZZ &a) { x = to_zz_p(a); }
inline void conv(long &x, const zz_p &a) { x = a.v; }
inline void conv(ZZ &x, const zz_p &a) { x = ZZ(a.v); }
inline bool IsZero(const zz_p &a) { return a.v == 0; }
inline bool IsOne(const zz_p &a) { return a.v == 1 % zz_p::mod(); }
inline void clear(zz_p &a) { a.v = 0; }
inline void set(zz_p &a) { a = zz_p(1); }
inline void random(zz_p &x) { x = zz_p(RandomBnd(zz_p::mod())); }
inline zz_p random_zz_p() { return zz_p(RandomBnd(zz_p::mod())); }
inline std::ostream &operator<<(std::ostream &os, const zz_p &a) { return os << a.v; }
typedef Vec<zz_p> vec_zz_p;

class zz_pContext {
  long p = 0;

 public:
  zz_pContext() {}
  explicit zz_pContext(long q, long /*maxroot*/ = 0) : p(q) {}
  void save() { p = zz_p::mod(); }
  void restore() const {
    if (p) zz_p::init(p);
  }
};
class zz_pBak {
  long p = 0;
  bool armed = false;

 public:
  void save() { p = zz_p::mod(), armed = true; }
  void restore() {
    if (armed) zz_p::init(p);
    armed = false;
  }
  ~zz_pBak() { restore(); }
};
class ZZ_pBak {
  ZZ p;
  bool armed = false;

 public:
  void save() { p = ZZ_p::modZZ(), armed = true; }
  void restore() {
    if (armed && !p.is_zero()) ZZ_p::init(p);
    armed = false;
  }
  ~ZZ_pBak() { restore(); }
};

class zz_pX {
 public:
  vec_zz_p rep;
  zz_pX() {}
  zz_pX(INIT_SIZE_TYPE, long n) { rep.v.reserve(n); }
  static const zz_pX &zero() {
    static const zz_pX z;
    return z;
  }
  void normalize() {
    while (!rep.v.empty() && rep.v.back().v == 0) rep.v.pop_back();
  }
  void SetMaxLength(long n) { rep.SetMaxLength(n); }
  void SetLength(long n) { rep.SetLength(n); }
  bool operator==(const zz_pX &o) const { return rep == o.rep; }
  bool operator!=(const zz_pX &o) const { return !(rep == o.rep); }
};
inline long deg(const zz_pX &a) { return (long)a.rep.v.size() - 1; }
inline zz_p coeff(const zz_pX &a, long i) { return (i >= 0 && i < (long)a.rep.v.size()) ? a.rep.v[i] : zz_p(); }
inline void SetCoeff(zz_pX &a, long i, const zz_p &c) {
  if ((long)a.rep.v.size() <= i) a.rep.v.resize(i + 1);
  a.rep.v[i] = c;
  a.normalize();
}
inline void SetCoeff(zz_pX &a, long i, long c) { SetCoeff(a, i, zz_p(c)); }
inline void SetCoeff(zz_pX &a, long i) { SetCoeff(a, i, zz_p(1)); }
inline void clear(zz_pX &a) { a.rep.v.clear(); }
inline bool IsZero(const zz_pX &a) { return a.rep.v.empty(); }
inline zz_p LeadCoeff(const zz_pX &a) { return a.rep.v.empty() ? zz_p() : a.rep.v.back(); }
inline zz_pX &operator+=(zz_pX &a, const zz_pX &o) {
  if (a.rep.v.size() < o.rep.v.size()) a.rep.v.resize(o.rep.v.size());
  for (size_t i = 0; i < o.rep.v.size(); ++i) a.rep.v[i] += o.rep.v[i];
  a.normalize();
  return a;
}
inline zz_pX &operator-=(zz_pX &a, const zz_pX &o) {
  if (a.rep.v.size() < o.rep.v.size()) a.rep.v.resize(o.rep.v.size());
  for (size_t i = 0; i < o.rep.v.size(); ++i) a.rep.v[i] -= o.rep.v[i];
  a.normalize();
  return a;
}
inline zz_pX operator+(zz_pX a, const zz_pX &b) { return a += b; }
inline zz_pX operator-(zz_pX a, const zz_pX &b) { return a -= b; }
inline zz_pX operator*(const zz_pX &a, const zz_pX &b) {
  zz_pX r;
  if (IsZero(a) || IsZero(b)) return r;
  const long m = zz_p::mod();
  std::vector<unsigned __int128> acc(a.rep.v.size() + b.rep.v.size() - 1, 0);
  for (size_t i = 0; i < a.rep.v.size(); ++i) {
    if (!a.rep.v[i].v) continue;
    for (size_t j = 0; j < b.rep.v.size(); ++j)
      acc[i + j] = (acc[i + j] + (unsigned __int128)a.rep.v[i].v * b.rep.v[j].v) % (unsigned long)m;
  }
  r.rep.v.resize(acc.size());
  for (size_t i = 0; i < acc.size(); ++i) r.rep.v[i].v = (long)acc[i];
  r.normalize();
  return r;
}
inline zz_pX operator*(const zz_pX &a, const zz_p &c) {
  zz_pX r = a;
  for (auto &x : r.rep.v) x *= c;
  r.normalize();
  return r;
}
inline zz_pX &operator*=(zz_pX &a, const zz_pX &b) { a = a * b; return a; }
inline zz_pX &operator*=(zz_pX &a, const zz_p &c) { a = a * c; return a; }
inline void mul(zz_pX &x, const zz_pX &a, const zz_pX &b) { x = a * b; }
inline zz_pX &operator/=(zz_pX &a, const zz_p &c) { a = a * inv(c); return a; }
inline zz_pX &operator/=(zz_pX &a, long c) { a = a * inv(zz_p(c)); return a; }
inline void DivRem(zz_pX &q, zz_pX &r, const zz_pX &a, const zz_pX &b) {
  const long db = deg(b);
  if (db < 0) Error("zz_pX DivRem: division by zero");
  zz_pX t = a;
  q = zz_pX();
  const zz_p li = inv(LeadCoeff(b));
  for (long i = deg(t); i >= db; --i) {
    zz_p c = coeff(t, i) * li;
    if (IsZero(c)) continue;
    if ((long)q.rep.v.size() <= i - db) q.rep.v.resize(i - db + 1);
    q.rep.v[i - db] = c;
    for (long j = 0; j <= db; ++j) t.rep.v[i - db + j] -= c * b.rep.v[j];
  }
  t.normalize();
  q.normalize();
  r = t;
}
inline void rem(zz_pX &r, const zz_pX &a, const zz_pX &b) { zz_pX q; DivRem(q, r, a, b); }
inline zz_pX operator%(const zz_pX &a, const zz_pX &b) { zz_pX q, r; DivRem(q, r, a, b); return r; }
inline zz_pX operator/(const zz_pX &a, const zz_pX &b) { zz_pX q, r; DivRem(q, r, a, b); return q; }
inline zz_pX &operator%=(zz_pX &a, const zz_pX &b) { a = a % b; return a; }
inline void conv(zz_pX &x, const ZZX &a) {  // each coefficient reduced mod the current modulus
  x.rep.v.resize(a.rep.v.size());
  for (size_t i = 0; i < a.rep.v.size(); ++i) x.rep.v[i] = to_zz_p(a.rep.v[i]);
  x.normalize();
}
inline void conv(ZZX &x, const zz_pX &a) {
  x.rep.v.resize(a.rep.v.size());
  for (size_t i = 0; i < a.rep.v.size(); ++i) x.rep.v[i] = ZZ(a.rep.v[i].v);
  x.normalize();
}
inline void conv(zz_pX &x, const zz_pX &a) { x = a; }
inline void conv(zz_pX &x, long a) { x = zz_pX(); SetCoeff(x, 0, a); }
inline zz_pX to_zz_pX(const ZZX &a) { zz_pX r; conv(r, a); return r; }
inline ZZX to_ZZX(const zz_pX &a) { ZZX r; conv(r, a); return r; }
inline std::ostream &operator<<(std::ostream &os, const zz_pX &a) {
  os << "[";
  for (size_t i = 0; i < a.rep.v.size(); ++i) os << (i ? " " : "") << a.rep.v[i];
  return os << "]";
}

// ------------------------------------------------------------------------------- FFT representation
// A length-2^k cyclic "FFT representation" of a polynomial over Z_q: the images under three
// 62-bit NTT primes, enough to hold any cyclic product of two vectors with entries < 2^63 and
// length <= 2^20 exactly; FromRep recombines (Garner) and reduces mod q.
namespace fftdetail {
typedef unsigned long u64;
typedef unsigned __int128 u128;
struct NttPrime {
  u64 p, g;  // prime = c * 2^32 + 1, generator of the 2^32-torsion source
};
inline u64 mulmod(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
inline u64 powmod(u64 a, u64 e, u64 p) {
  u64 r = 1;
  for (; e; e >>= 1, a = mulmod(a, a, p))
    if (e & 1) r = mulmod(r, a, p);
  return r;
}
inline const std::array<NttPrime, 3> &primes() {
  static const std::array<NttPrime, 3> P = [] {
    std::array<NttPrime, 3> r{};
    int found = 0;
    for (u64 c = (1ull << 30) - 1; found < 3; --c) {  // p = c * 2^32 + 1 < 2^62
      u64 p = (c << 32) + 1;
      if (!ProbPrime((long)p)) continue;
      // an element of order exactly 2^32: x^((p-1)/2^32) with x a non-residue
      for (u64 x = 2;; ++x)
        if (powmod(x, (p - 1) / 2, p) == p - 1) {
          r[found++] = NttPrime{p, powmod(x, (p - 1) >> 32, p)};
          break;
        }
    }
    return r;
  }();
  return P;
}
inline void ntt(std::vector<u64> &a, int k, const NttPrime &P, bool inverse) {
  const size_t n = (size_t)1 << k;
  for (size_t i = 1, j = 0; i < n; ++i) {
    size_t b = n >> 1;
    for (; j & b; b >>= 1) j ^= b;
    j ^= b;
    if (i < j) std::swap(a[i], a[j]);
  }
  u64 w_n = powmod(P.g, (u64)1 << (32 - k), P.p);  // order 2^k
  if (inverse) w_n = powmod(w_n, P.p - 2, P.p);
  for (size_t len = 2; len <= n; len <<= 1) {
    const u64 wl = powmod(w_n, n / len, P.p);
    for (size_t s = 0; s < n; s += len) {
      u64 w = 1;
      for (size_t j = 0; j < len / 2; ++j) {
        u64 u = a[s + j], v = mulmod(a[s + j + len / 2], w, P.p);
        a[s + j] = u + v >= P.p ? u + v - P.p : u + v;
        a[s + j + len / 2] = u >= v ? u - v : u + P.p - v;
        w = mulmod(w, wl, P.p);
      }
    }
  }
  if (inverse) {
    const u64 ni = powmod((u64)n % P.p, P.p - 2, P.p);
    for (auto &x : a) x = mulmod(x, ni, P.p);
  }
}
}  // namespace fftdetail

class fftRep {
 public:
  long k = -1;
  std::array<std::vector<unsigned long>, 3> img;
  fftRep() {}
  fftRep(INIT_SIZE_TYPE, long kk) { SetSize(kk); }
  void SetSize(long kk) {
    k = kk;
    for (auto &v : img) v.assign(kk >= 0 ? (size_t)1 << kk : 0, 0);
  }
};
typedef fftRep FFTRep;

// y = representation of x[lo..hi] (coefficient lo becomes position 0), folded modulo X^(2^k) - 1
template <class PX>
inline void ToRepGeneric(fftRep &y, const PX &x, long k, long lo, long hi) {
  using namespace fftdetail;
  y.SetSize(k);
  const size_t n = (size_t)1 << k;
  hi = std::min(hi, deg(x));
  for (int t = 0; t < 3; ++t) {
    const NttPrime &P = primes()[t];
    std::vector<u64> &a = y.img[t];
    for (long i = lo; i <= hi; ++i) {
      size_t pos = (size_t)(i - lo) & (n - 1);
      u64 v = (u64)x.rep.v[i].v % P.p;
      a[pos] = a[pos] + v >= P.p ? a[pos] + v - P.p : a[pos] + v;
    }
    ntt(a, (int)k, P, false);
  }
}
inline void mul(fftRep &z, const fftRep &x, const fftRep &y) {
  using namespace fftdetail;
  if (x.k != y.k) Error("fftRep mul: size mismatch");
  fftRep r;
  r.k = x.k;
  for (int t = 0; t < 3; ++t) {
    const u64 p = primes()[t].p;
    r.img[t].resize(x.img[t].size());
    for (size_t i = 0; i < x.img[t].size(); ++i) r.img[t][i] = mulmod(x.img[t][i], y.img[t][i], p);
  }
  z = r;
}
// coefficients lo..hi of the cyclic product, reduced mod q, into x (coefficient lo -> constant term)
inline void FromRepGeneric(std::vector<unsigned long> &out, const fftRep &y, long lo, long hi, unsigned long q) {
  using namespace fftdetail;
  const size_t n = (size_t)1 << y.k;
  std::array<std::vector<u64>, 3> c = y.img;
  for (int t = 0; t < 3; ++t) ntt(c[t], (int)y.k, primes()[t], true);
  const u64 p0 = primes()[0].p, p1 = primes()[1].p, p2 = primes()[2].p;
  const u64 i01 = powmod(p0 % p1, p1 - 2, p1);                       // p0^-1 mod p1
  const u64 i012 = powmod(mulmod(p0 % p2, p1 % p2, p2), p2 - 2, p2);  // (p0 p1)^-1 mod p2
  const u64 p0q = p0 % q, p01q = (u64)((u128)p0q * (p1 % q) % q);
  out.clear();
  for (long i = lo; i <= hi; ++i) {
    if (i < 0 || (size_t)i >= n) {
      out.push_back(0);
      continue;
    }
    // Garner: v = a0 + p0 (a1 + p1 a2)
    const u64 r0 = c[0][i], r1 = c[1][i], r2 = c[2][i];
    const u64 a0 = r0;
    const u64 a1 = mulmod((r1 + p1 - a0 % p1) % p1, i01, p1);
    const u64 t2 = (u64)(((u128)a0 % p2 + (u128)(p0 % p2) * (a1 % p2)) % p2);
    const u64 a2 = mulmod((r2 + p2 - t2) % p2, i012, p2);
    const u64 v = (u64)(((u128)(a0 % q) + (u128)p0q * (a1 % q) % q + (u128)p01q * (a2 % q) % q) % q);
    out.push_back(v);
  }
}
inline void TofftRep(fftRep &y, const zz_pX &x, long k, long lo, long hi) { ToRepGeneric(y, x, k, lo, hi); }
inline void TofftRep(fftRep &y, const zz_pX &x, long k) { ToRepGeneric(y, x, k, 0, deg(x)); }
inline void FromfftRep(zz_pX &x, fftRep &y, long lo, long hi) {
  std::vector<unsigned long> out;
  FromRepGeneric(out, y, lo, hi, (unsigned long)zz_p::mod());
  x.rep.v.resize(out.size());
  for (size_t i = 0; i < out.size(); ++i) x.rep.v[i].v = (long)out[i];
  x.normalize();
}
inline void ToFFTRep(FFTRep &y, const ZZ_pX &x, long k, long lo, long hi) { ToRepGeneric(y, x, k, lo, hi); }
inline void ToFFTRep(FFTRep &y, const ZZ_pX &x, long k) { ToRepGeneric(y, x, k, 0, deg(x)); }
inline void FromFFTRep(ZZ_pX &x, FFTRep &y, long lo, long hi) {
  std::vector<unsigned long> out;
  FromRepGeneric(out, y, lo, hi, (unsigned long)ZZ_p::mod());
  x.rep.v.resize(out.size());
  for (size_t i = 0; i < out.size(); ++i) x.rep.v[i].v = (long)out[i];
  x.normalize();
}

// ------------------------------------------------------------------------------- GF(2) stand-ins
// PAlgebra.h / NumbTh.h declare helpers over GF2X that the Brakerski path never calls; the types
// only have to exist.
class GF2 {
 public:
  long v = 0;
};
inline long rep(const GF2 &a) { return a.v; }
class GF2X {
 public:
  Vec<long> rep;
  void SetMaxLength(long n) { rep.SetMaxLength(n); }
};
inline long deg(const GF2X &a) { return (long)a.rep.v.size() - 1; }
inline bool IsZero(const GF2X &a) { return a.rep.v.empty(); }
inline void SetCoeff(GF2X &a, long i, long c = 1) {
  if ((long)a.rep.v.size() <= i) a.rep.v.resize(i + 1, 0);
  a.rep.v[i] = c & 1;
  while (!a.rep.v.empty() && !a.rep.v.back()) a.rep.v.pop_back();
}
inline GF2 coeff(const GF2X &a, long i) {
  GF2 r;
  r.v = (i >= 0 && i < (long)a.rep.v.size()) ? a.rep.v[i] : 0;
  return r;
}
inline bool IsOne(const GF2 &a) { return a.v == 1; }
typedef Vec<GF2X> vec_GF2X;
typedef Vec<GF2> vec_GF2;
class GF2E {};
class GF2EX {};
class GF2XModulus {};
class zz_pE {};
class zz_pEX {};
class ZZ_pE {};
class ZZ_pEX {};

// ------------------------------------------------------------------------------- matrices
template <class T>
class Mat {
  std::vector<Vec<T>> rows;
  long nc = 0;

 public:
  void SetDims(long r, long c) {
    rows.resize(r);
    nc = c;
    for (auto &v : rows) v.SetLength(c);
  }
  long NumRows() const { return (long)rows.size(); }
  long NumCols() const { return nc; }
  Vec<T> &operator[](long i) { return rows[i]; }
  const Vec<T> &operator[](long i) const { return rows[i]; }
  void kill() { rows.clear(), nc = 0; }
};
typedef Mat<long> mat_long;
typedef Mat<ZZ> mat_ZZ;
typedef Vec<vec_long> vec_vec_long;

}  // namespace NTL
