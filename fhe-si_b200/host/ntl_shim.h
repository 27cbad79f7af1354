// ntl_shim.h -- the slice of NTL's interface that the reference's CLIENTS use
// (Test_AddMul.cpp, Test_General.cpp, Test_Regression.cpp, Test_Statistics.cpp, Regression.h,
// Statistics.h, Matrix.*; SURVEY.md §8b last row), written from scratch.
//
// This is host-side glue, not the hot path: big integers here only carry plaintext-side
// values, key generation and (de)serialisation.  Ciphertext arithmetic never goes through ZZ --
// it runs in libfhesi_b200.so on the GPU.  Semantics honoured: ZZ '/' is floor
// division, '%' is non-negative for a positive modulus, '>>' shifts the magnitude and keeps the
// sign, BytesFromZZ is the little-endian magnitude, NumBytes(0) = 0, deg(0) = -1.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

namespace NTL {

enum INIT_SIZE_TYPE { INIT_SIZE };

[[noreturn]] inline void Error(const char *msg) {
  std::cerr << msg << std::endl;
  std::abort();
}

// Limb storage for ZZ: up to 20 limbs (640 bits -- every coefficient, chain product and key value
// of the five configurations) live inline, so ZZ temporaries do not touch the heap.
class Limbs {
  static constexpr uint32_t kInline = 20;
  uint32_t n_ = 0, cap_ = kInline;
  uint32_t *p_ = buf_;
  uint32_t buf_[kInline];
  void grow(size_t want) {
    size_t cap = std::max<size_t>(want, 2 * (size_t)cap_);
    uint32_t *q = (uint32_t *)std::malloc(cap * 4);
    if (!q) std::abort();
    std::memcpy(q, p_, (size_t)n_ * 4);
    if (p_ != buf_) std::free(p_);
    p_ = q;
    cap_ = (uint32_t)cap;
  }

 public:
  Limbs() {}
  Limbs(size_t n, uint32_t v) { assign(n, v); }
  Limbs(const Limbs &o) { assign(o.p_, o.p_ + o.n_); }
  Limbs(Limbs &&o) noexcept { steal(o); }
  ~Limbs() {
    if (p_ != buf_) std::free(p_);
  }
  Limbs &operator=(const Limbs &o) {
    if (this != &o) assign(o.p_, o.p_ + o.n_);
    return *this;
  }
  Limbs &operator=(Limbs &&o) noexcept {
    if (this != &o) {
      if (p_ != buf_) std::free(p_);
      p_ = buf_, cap_ = kInline, n_ = 0;
      steal(o);
    }
    return *this;
  }
  void steal(Limbs &o) {
    if (o.p_ == o.buf_) {
      std::memcpy(buf_, o.buf_, (size_t)o.n_ * 4);
      n_ = o.n_;
    } else {
      p_ = o.p_, cap_ = o.cap_, n_ = o.n_;
      o.p_ = o.buf_, o.cap_ = kInline;
    }
    o.n_ = 0;
  }
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  uint32_t &operator[](size_t i) { return p_[i]; }
  const uint32_t &operator[](size_t i) const { return p_[i]; }
  uint32_t &back() { return p_[n_ - 1]; }
  const uint32_t &back() const { return p_[n_ - 1]; }
  uint32_t *data() { return p_; }
  const uint32_t *data() const { return p_; }
  uint32_t *begin() { return p_; }
  uint32_t *end() { return p_ + n_; }
  const uint32_t *begin() const { return p_; }
  const uint32_t *end() const { return p_ + n_; }
  void clear() { n_ = 0; }
  void pop_back() { --n_; }
  void push_back(uint32_t v) {
    if (n_ == cap_) grow(n_ + 1);
    p_[n_++] = v;
  }
  void resize(size_t n, uint32_t v = 0) {
    if (n > cap_) grow(n);
    for (size_t i = n_; i < n; ++i) p_[i] = v;
    n_ = (uint32_t)n;
  }
  void assign(size_t n, uint32_t v) {
    n_ = 0;
    resize(n, v);
  }
  template <class It>
  void assign(It a, It b) {
    size_t n = (size_t)(b - a);
    if (n > cap_) {
      n_ = 0;
      grow(n);
    }
    for (size_t i = 0; i < n; ++i) p_[i] = (uint32_t)a[i];
    n_ = (uint32_t)n;
  }
  void swap(Limbs &o) {
    Limbs t(std::move(o));
    o = std::move(*this);
    *this = std::move(t);
  }
};

// ------------------------------------------------------------------------------- ZZ
class ZZ {
 public:
  Limbs mag;  // little-endian, no leading zero limbs; empty = 0
  bool neg = false;

  ZZ() {}
  ZZ(long v) { set_long(v); }
  ZZ(int v) { set_long(v); }
  ZZ(unsigned v) { set_ulong(v); }
  ZZ(unsigned long v) { set_ulong(v); }
  static const ZZ &zero() {
    static const ZZ z;
    return z;
  }
  void set_ulong(unsigned long long v) {
    mag.clear();
    neg = false;
    while (v) {
      mag.push_back((uint32_t)v);
      v >>= 32;
    }
  }
  void set_long(long long v) {
    bool n = v < 0;
    set_ulong(n ? 0ull - (unsigned long long)v : (unsigned long long)v);
    neg = n && !mag.empty();
  }
  ZZ &operator=(long v) {
    set_long(v);
    return *this;
  }
  bool is_zero() const { return mag.empty(); }
  void trim() {
    while (!mag.empty() && mag.back() == 0) mag.pop_back();
    if (mag.empty()) neg = false;
  }
  size_t bits() const {
    if (mag.empty()) return 0;
    return 32 * (mag.size() - 1) + (32 - __builtin_clz(mag.back()));
  }

  // ---- magnitude helpers
  static int cmp_mag(const Limbs &a, const Limbs &b) {
    if (a.size() != b.size()) return a.size() < b.size() ? -1 : 1;
    for (size_t i = a.size(); i-- > 0;)
      if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
    return 0;
  }
  static void add_mag(Limbs &a, const Limbs &b) {
    if (a.size() < b.size()) a.resize(b.size(), 0);
    uint64_t c = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      c += (uint64_t)a[i] + (i < b.size() ? b[i] : 0);
      a[i] = (uint32_t)c;
      c >>= 32;
    }
    if (c) a.push_back((uint32_t)c);
  }
  static void sub_mag(Limbs &a, const Limbs &b) {  // a >= b
    int64_t br = 0;
    for (size_t i = 0; i < a.size(); ++i) {
      int64_t t = (int64_t)a[i] - (i < b.size() ? b[i] : 0) - br;
      br = t < 0;
      a[i] = (uint32_t)(t + (br ? (1ll << 32) : 0));
    }
  }
  static int cmp(const ZZ &a, const ZZ &b) {
    if (a.neg != b.neg) return a.neg ? -1 : 1;
    int c = cmp_mag(a.mag, b.mag);
    return a.neg ? -c : c;
  }

  ZZ &operator+=(const ZZ &o) {
    if (neg == o.neg) {
      add_mag(mag, o.mag);
    } else if (cmp_mag(mag, o.mag) >= 0) {
      sub_mag(mag, o.mag);
    } else {
      Limbs t = o.mag;
      sub_mag(t, mag);
      mag.swap(t);
      neg = o.neg;
    }
    trim();
    return *this;
  }
  ZZ operator-() const {
    ZZ r = *this;
    if (!r.mag.empty()) r.neg = !r.neg;
    return r;
  }
  ZZ &operator-=(const ZZ &o) { return *this += -o; }
  ZZ &operator*=(const ZZ &o) {
    if (mag.empty() || o.mag.empty()) {
      mag.clear();
      neg = false;
      return *this;
    }
    Limbs r(mag.size() + o.mag.size(), 0);
    for (size_t i = 0; i < mag.size(); ++i) {
      uint64_t c = 0;
      for (size_t j = 0; j < o.mag.size(); ++j) {
        c += (uint64_t)mag[i] * o.mag[j] + r[i + j];
        r[i + j] = (uint32_t)c;
        c >>= 32;
      }
      r[i + o.mag.size()] += (uint32_t)c;
    }
    mag.swap(r);
    neg = neg != o.neg;
    trim();
    return *this;
  }
  // truncated magnitude division: q = |a| / |b|, r = |a| % |b|
  static void divmod_mag(const Limbs &a, const Limbs &b,
                         Limbs &q, Limbs &r) {
    if (b.empty()) Error("ZZ: division by zero");
    q.assign(a.size(), 0);
    r.clear();
    if (b.size() == 1) {
      uint64_t rem = 0;
      for (size_t i = a.size(); i-- > 0;) {
        uint64_t cur = (rem << 32) | a[i];
        q[i] = (uint32_t)(cur / b[0]);
        rem = cur % b[0];
      }
      if (rem) r.push_back((uint32_t)rem);
    } else {  // bitwise long division: sizes here are tiny (plaintext-side only)
      for (size_t i = a.size() * 32; i-- > 0;) {
        // r = 2r + bit
        uint32_t carry = (a[i / 32] >> (i % 32)) & 1;
        for (size_t k = 0; k < r.size(); ++k) {
          uint32_t nc = r[k] >> 31;
          r[k] = (r[k] << 1) | carry;
          carry = nc;
        }
        if (carry) r.push_back(carry);
        while (!r.empty() && r.back() == 0) r.pop_back();
        if (cmp_mag(r, b) >= 0) {
          sub_mag(r, b);
          while (!r.empty() && r.back() == 0) r.pop_back();
          q[i / 32] |= 1u << (i % 32);
        }
      }
    }
    while (!q.empty() && q.back() == 0) q.pop_back();
  }
  // floor division and non-negative-for-positive-modulus remainder (NTL semantics)
  static void DivRem(ZZ &q, ZZ &r, const ZZ &a, const ZZ &b) {
    Limbs qm, rm;
    divmod_mag(a.mag, b.mag, qm, rm);
    q.mag = qm;
    q.neg = (a.neg != b.neg) && !qm.empty();
    r.mag = rm;
    r.neg = a.neg && !rm.empty();
    if (!r.mag.empty() && (a.neg != b.neg)) {  // adjust truncation to floor
      q -= ZZ(1L);
      r += b;
    }
    q.trim();
    r.trim();
  }
  ZZ &operator/=(const ZZ &o) {
    ZZ q, r;
    DivRem(q, r, *this, o);
    return *this = q;
  }
  ZZ &operator%=(const ZZ &o) {
    ZZ q, r;
    DivRem(q, r, *this, o);
    return *this = r;
  }
  ZZ &operator<<=(long k) {
    if (k < 0) return *this >>= -k;
    if (mag.empty() || k == 0) return *this;
    size_t ws = k / 32, bs = k % 32;
    Limbs r(mag.size() + ws + 1, 0);
    for (size_t i = 0; i < mag.size(); ++i) {
      uint64_t v = (uint64_t)mag[i] << bs;
      r[i + ws] |= (uint32_t)v;
      r[i + ws + 1] |= (uint32_t)(v >> 32);
    }
    mag.swap(r);
    trim();
    return *this;
  }
  ZZ &operator>>=(long k) {  // magnitude shift, sign kept (Util.cpp:16 relies on this)
    if (k < 0) return *this <<= -k;
    size_t ws = k / 32, bs = k % 32;
    if (ws >= mag.size()) {
      mag.clear();
      neg = false;
      return *this;
    }
    Limbs r(mag.size() - ws, 0);
    for (size_t i = 0; i < r.size(); ++i) {
      uint64_t v = mag[i + ws];
      if (i + ws + 1 < mag.size()) v |= (uint64_t)mag[i + ws + 1] << 32;
      r[i] = (uint32_t)(v >> bs);
    }
    mag.swap(r);
    trim();
    return *this;
  }
  ZZ &operator+=(long v) { return *this += ZZ(v); }
  ZZ &operator-=(long v) { return *this -= ZZ(v); }
  ZZ &operator*=(long v) { return *this *= ZZ(v); }
  ZZ &operator/=(long v) { return *this /= ZZ(v); }
  ZZ &operator%=(long v) { return *this %= ZZ(v); }
  ZZ &operator++() { return *this += 1L; }
  ZZ &operator--() { return *this -= 1L; }
};

#define FHESI_ZZ_BINOP(op)                                                        \
  inline ZZ operator op(const ZZ &a, const ZZ &b) { ZZ r = a; r op## = b; return r; } \
  inline ZZ operator op(const ZZ &a, long b) { ZZ r = a; r op## = ZZ(b); return r; }  \
  inline ZZ operator op(long a, const ZZ &b) { ZZ r(a); r op## = b; return r; }
FHESI_ZZ_BINOP(+)
FHESI_ZZ_BINOP(-)
FHESI_ZZ_BINOP(*)
FHESI_ZZ_BINOP(/)
#undef FHESI_ZZ_BINOP
inline ZZ operator%(const ZZ &a, const ZZ &b) { ZZ r = a; r %= b; return r; }
inline long to_long(const ZZ &a);
inline long operator%(const ZZ &a, long b) { ZZ r = a; r %= ZZ(b); return to_long(r); }  // NTL: long
inline ZZ operator<<(const ZZ &a, long k) { ZZ r = a; r <<= k; return r; }
inline ZZ operator>>(const ZZ &a, long k) { ZZ r = a; r >>= k; return r; }
#define FHESI_ZZ_CMP(op)                                                          \
  inline bool operator op(const ZZ &a, const ZZ &b) { return ZZ::cmp(a, b) op 0; } \
  inline bool operator op(const ZZ &a, long b) { return ZZ::cmp(a, ZZ(b)) op 0; }  \
  inline bool operator op(long a, const ZZ &b) { return ZZ::cmp(ZZ(a), b) op 0; }
FHESI_ZZ_CMP(==)
FHESI_ZZ_CMP(!=)
FHESI_ZZ_CMP(<)
FHESI_ZZ_CMP(<=)
FHESI_ZZ_CMP(>)
FHESI_ZZ_CMP(>=)
#undef FHESI_ZZ_CMP

inline ZZ to_ZZ(long v) { return ZZ(v); }
inline ZZ to_ZZ(int v) { return ZZ((long)v); }
inline ZZ to_ZZ(unsigned v) { return ZZ((unsigned long)v); }
inline ZZ to_ZZ(unsigned long v) { return ZZ(v); }
inline const ZZ &to_ZZ(const ZZ &v) { return v; }
inline long sign(const ZZ &a) { return a.is_zero() ? 0 : (a.neg ? -1 : 1); }
inline bool IsZero(const ZZ &a) { return a.is_zero(); }
inline bool IsOne(const ZZ &a) { return !a.neg && a.mag.size() == 1 && a.mag[0] == 1; }
inline void clear(ZZ &a) { a = ZZ(); }
inline long NumBits(const ZZ &a) { return (long)a.bits(); }
inline long NumBytes(const ZZ &a) { return (long)((a.bits() + 7) / 8); }
inline long to_long(const ZZ &a) {
  unsigned long long v = 0;
  for (size_t i = 0; i < a.mag.size() && i < 2; ++i) v |= (unsigned long long)a.mag[i] << (32 * i);
  return a.neg ? -(long)v : (long)v;
}
inline unsigned long to_ulong(const ZZ &a) { return (unsigned long)to_long(a); }
inline void conv(long &x, const ZZ &a) { x = to_long(a); }
inline void conv(ZZ &x, long a) { x = ZZ(a); }
inline void conv(ZZ &x, const ZZ &a) { x = a; }
inline double to_double(double d) { return d; }
inline double to_double(long d) { return (double)d; }
inline double to_double(const ZZ &a) {
  double r = 0;
  for (size_t i = a.mag.size(); i-- > 0;) r = r * 4294967296.0 + a.mag[i];
  return a.neg ? -r : r;
}
inline double log(const ZZ &a) {  // natural log of |a|, fine for parameter sizing
  size_t b = a.bits();
  if (b <= 1000) return std::log(to_double(a));
  ZZ t = a >> (long)(b - 960);
  return std::log(std::fabs(to_double(t))) + (double)(b - 960) * 0.6931471805599453;
}
inline void BytesFromZZ(unsigned char *p, const ZZ &a, long n) {
  for (long i = 0; i < n; ++i) {
    size_t w = i / 4;
    p[i] = w < a.mag.size() ? (unsigned char)(a.mag[w] >> (8 * (i % 4))) : 0;
  }
}
inline void ZZFromBytes(ZZ &x, const unsigned char *p, long n) {
  x.mag.assign((n + 3) / 4, 0);
  x.neg = false;
  for (long i = 0; i < n; ++i) x.mag[i / 4] |= (uint32_t)p[i] << (8 * (i % 4));
  x.trim();
}
inline ZZ ZZFromBytes(const unsigned char *p, long n) {
  ZZ x;
  ZZFromBytes(x, p, n);
  return x;
}
inline long rem(const ZZ &a, long m) { return to_long(a % ZZ(m)); }
inline void rem(ZZ &r, const ZZ &a, const ZZ &m) { r = a % m; }
inline long divide(const ZZ &a, long d) { return d != 0 && rem(a, d) == 0; }
inline void mul(ZZ &x, const ZZ &a, const ZZ &b) { x = a * b; }
inline void mul(ZZ &x, long a, const ZZ &b) { x = ZZ(a) * b; }
inline void add(ZZ &x, const ZZ &a, const ZZ &b) { x = a + b; }
inline void sub(ZZ &x, const ZZ &a, const ZZ &b) { x = a - b; }
inline void power(ZZ &x, const ZZ &a, long e) {
  ZZ r(1L), b = a;
  while (e > 0) {
    if (e & 1) r *= b;
    b *= b;
    e >>= 1;
  }
  x = r;
}
inline ZZ power(const ZZ &a, long e) {
  ZZ x;
  power(x, a, e);
  return x;
}
inline ZZ power2_ZZ(long e) { return ZZ(1L) << e; }
inline ZZ PowerMod(const ZZ &a, const ZZ &e, const ZZ &n) {
  ZZ r(1L), b = a % n;
  for (size_t i = 0; i < e.bits(); ++i) {
    if ((e.mag[i / 32] >> (i % 32)) & 1) r = (r * b) % n;
    b = (b * b) % n;
  }
  return r;
}
inline ZZ InvMod(const ZZ &a, const ZZ &n) {  // extended Euclid
  ZZ r0 = n, r1 = a % n, s0(0L), s1(1L);
  while (!r1.is_zero()) {
    ZZ q = r0 / r1, t = r0 - q * r1;
    r0 = r1, r1 = t;
    t = s0 - q * s1;
    s0 = s1, s1 = t;
  }
  if (!IsOne(r0)) Error("InvMod: inverse undefined");
  return s0 % n;
}
inline long InvMod(long a, long n) { return to_long(InvMod(ZZ(a), ZZ(n))); }
inline long MulMod(long a, long b, long n) { return (long)((unsigned __int128)a * b % n); }
inline long AddMod(long a, long b, long n) { long r = a + b; return r >= n ? r - n : r; }
inline long SubMod(long a, long b, long n) { long r = a - b; return r < 0 ? r + n : r; }
inline long NegateMod(long a, long n) { return a ? n - a : 0; }
inline long PowerMod(long a, long e, long n) {
  long r = 1 % n;
  a %= n;
  while (e > 0) {
    if (e & 1) r = MulMod(r, a, n);
    a = MulMod(a, a, n);
    e >>= 1;
  }
  return r;
}
inline long GCD(long a, long b) {
  while (b) {
    long t = a % b;
    a = b, b = t;
  }
  return a < 0 ? -a : a;
}
inline long ProbPrime(long n) {
  if (n < 2) return 0;
  static const long B[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (long b : B)
    if (n % b == 0) return n == b;
  long d = n - 1, s = 0;
  while (!(d & 1)) d >>= 1, ++s;
  for (long a : B) {
    long x = PowerMod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (long i = 1; i < s && comp; ++i) {
      x = MulMod(x, x, n);
      if (x == n - 1) comp = false;
    }
    if (comp) return 0;
  }
  return 1;
}
inline long ProbPrime(const ZZ &n) { return ProbPrime(to_long(n)); }
inline long NextPowerOfTwo(long m) {
  long k = 0;
  while ((1L << k) < m) ++k;
  return k;
}

inline std::ostream &operator<<(std::ostream &os, const ZZ &a) {
  if (a.is_zero()) return os << "0";
  std::string s;
  Limbs m = a.mag;
  while (!m.empty()) {
    uint64_t rem = 0;
    for (size_t i = m.size(); i-- > 0;) {
      uint64_t cur = (rem << 32) | m[i];
      m[i] = (uint32_t)(cur / 1000000000u);
      rem = cur % 1000000000u;
    }
    while (!m.empty() && m.back() == 0) m.pop_back();
    char buf[16];
    snprintf(buf, sizeof buf, m.empty() ? "%u" : "%09u", (unsigned)rem);
    s = std::string(buf) + s;
  }
  return os << (a.neg ? "-" : "") << s;
}
inline std::istream &operator>>(std::istream &is, ZZ &a) {
  std::string tok;
  if (!(is >> tok)) return is;
  a = ZZ();
  size_t i = 0;
  bool neg = false;
  if (tok[0] == '-' || tok[0] == '+') neg = tok[0] == '-', i = 1;
  for (; i < tok.size(); ++i) {
    if (tok[i] < '0' || tok[i] > '9') break;
    a *= ZZ(10L);
    a += ZZ((long)(tok[i] - '0'));
  }
  if (neg && !a.is_zero()) a.neg = true;
  return is;
}

// -------------------------------------------------------------------- PRNG
// NTL's stream is a cryptographic PRG and is not pinned by the reference.  Here:
//  * ChaCha20 (the RFC 8439 block function; 256-bit key, 64-bit block counter) is the generator.  Unseeded,
//    its key comes from the operating system (getrandom / /dev/urandom); SetSeed(seed) derives the key from
//    EVERY byte of the seed, so equal seeds give equal streams and the key space is the seed's.
//  * SplitMix64 is the documented deterministic TEST stream: the one oracle/fhesi_oracle.py::Rng draws from and
//    the golden vectors were made with.  It has 64 bits of state and is NOT for keys that protect anything; it
//    is selected only explicitly, by the environment variable FHESI_TEST_RNG=splitmix64 (read once per thread;
//    tests/conftest.py and the golden generators set it) or by UseTestRandomStream(true).
// The stream is per thread (as in NTL's thread-safe build): no shared mutable state, no locking.
struct RandomStream {
  bool test = false;
  uint64_t state = 0;                    // SplitMix64
  static const unsigned LANES = 4;       // ChaCha20 blocks per refill, computed side by side (SSE2 lanes)
  uint32_t key[8] = {0}, block[16 * LANES] = {0};
  uint64_t counter = 0;
  unsigned have = 0;                     // unread 64-bit words left in `block`

  static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
  static void quarter(uint32_t *x, int a, int b, int c, int d) {
    x[a] += x[b], x[d] = rotl(x[d] ^ x[a], 16);
    x[c] += x[d], x[b] = rotl(x[b] ^ x[c], 12);
    x[a] += x[b], x[d] = rotl(x[d] ^ x[a], 8);
    x[c] += x[d], x[b] = rotl(x[b] ^ x[c], 7);
  }
  static void chacha_block(uint32_t out[16], const uint32_t k[8], uint64_t ctr, uint32_t n0, uint32_t n1) {
    const uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, k[0], k[1], k[2], k[3],
                             k[4], k[5], k[6], k[7], (uint32_t)ctr, (uint32_t)(ctr >> 32), n0, n1};
    uint32_t x[16];
    for (int i = 0; i < 16; ++i) x[i] = in[i];
    for (int r = 0; r < 10; ++r) {
      quarter(x, 0, 4, 8, 12), quarter(x, 1, 5, 9, 13), quarter(x, 2, 6, 10, 14), quarter(x, 3, 7, 11, 15);
      quarter(x, 0, 5, 10, 15), quarter(x, 1, 6, 11, 12), quarter(x, 2, 7, 8, 13), quarter(x, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + in[i];
  }
  // LANES consecutive blocks (counters ctr .. ctr + LANES - 1) of the output stream, word-sliced so that every
  // step is the same operation on LANES independent values: the compiler turns the inner loops into SIMD.
  // Same bytes as LANES calls of chacha_block (checked by tests/test_host_cpp.py::test_random_stream_is_keyed_emu).
  void chacha_blocks(uint64_t ctr) {
    typedef uint32_t v4 __attribute__((vector_size(16)));  // GCC / clang vector extension: SSE2 on x86-64
    static_assert(LANES == 4, "one 128-bit vector of block lanes");
    static const uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    v4 x[16], in[16];
    for (int i = 0; i < 4; ++i) in[i] = v4{sigma[i], sigma[i], sigma[i], sigma[i]};
    for (int i = 0; i < 8; ++i) in[4 + i] = v4{key[i], key[i], key[i], key[i]};
    in[12] = v4{(uint32_t)ctr, (uint32_t)(ctr + 1), (uint32_t)(ctr + 2), (uint32_t)(ctr + 3)};
    in[13] = v4{(uint32_t)(ctr >> 32), (uint32_t)((ctr + 1) >> 32), (uint32_t)((ctr + 2) >> 32), (uint32_t)((ctr + 3) >> 32)};
    in[14] = in[15] = v4{0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) x[i] = in[i];
#define FHESI_QR(a, b, c, d)                                            \
  x[a] += x[b], x[d] ^= x[a], x[d] = (x[d] << 16) | (x[d] >> 16);       \
  x[c] += x[d], x[b] ^= x[c], x[b] = (x[b] << 12) | (x[b] >> 20);       \
  x[a] += x[b], x[d] ^= x[a], x[d] = (x[d] << 8) | (x[d] >> 24);        \
  x[c] += x[d], x[b] ^= x[c], x[b] = (x[b] << 7) | (x[b] >> 25);
    for (int r = 0; r < 10; ++r) {
      FHESI_QR(0, 4, 8, 12) FHESI_QR(1, 5, 9, 13) FHESI_QR(2, 6, 10, 14) FHESI_QR(3, 7, 11, 15)
      FHESI_QR(0, 5, 10, 15) FHESI_QR(1, 6, 11, 12) FHESI_QR(2, 7, 8, 13) FHESI_QR(3, 4, 9, 14)
    }
#undef FHESI_QR
    for (int i = 0; i < 16; ++i) {
      const v4 o = x[i] + in[i];
      for (unsigned l = 0; l < LANES; ++l) block[16 * l + i] = o[l];
    }
  }
  void rekey(const uint32_t k[8]) {
    for (int i = 0; i < 8; ++i) key[i] = k[i];
    counter = 0, have = 0;
  }
  // key <- absorb(words): 8 words at a time are XORed into the key, which is then replaced by a block of
  // its own keystream (domain-separated from the output stream by the nonce)
  void seed_words(const uint32_t *w, size_t n, uint32_t sign) {
    uint32_t k[8] = {0x66686573u, 0x692d7369u, (uint32_t)n, sign, 0, 0, 0, 0}, out[16];
    for (size_t i = 0; i < n || i == 0; i += 8) {
      for (size_t j = 0; j < 8 && i + j < n; ++j) k[j] ^= w[i + j];
      chacha_block(out, k, i, 0x73656564u, 0x6b657921u);
      for (int j = 0; j < 8; ++j) k[j] = out[j];
    }
    rekey(k);
  }
  void seed_from_os() {
    uint32_t k[8];
    size_t got = 0;
#if defined(__linux__)
    FILE *f = fopen("/dev/urandom", "rb");
    if (f) {
      got = fread(k, 1, sizeof k, f);
      fclose(f);
    }
#endif
    if (got != sizeof k) Error("RandomStream: the operating system gave no entropy");
    rekey(k);
  }
  RandomStream() {
    const char *e = getenv("FHESI_TEST_RNG");
    if (e && std::string(e) == "splitmix64") test = true;
    else seed_from_os();
  }
  // The ChaCha20 stream is block-counter mode: the number of 64-bit words drawn so far, and random access to any
  // word of the stream (fhesih_keydraws_flat fills the long, fixed-length draws of a key set-up on several cores)
  uint64_t position() const { return counter * 8 - have; }
  void seek(uint64_t pos) {
    const uint64_t c0 = pos / (8 * LANES) * LANES;
    chacha_blocks(c0);
    counter = c0 + LANES;
    have = (unsigned)(8 * LANES - (pos - c0 * 8));
  }
  uint64_t next64() {
    if (test) {
      state += 0x9E3779B97F4A7C15ull;
      uint64_t z = state;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      return z ^ (z >> 31);
    }
    if (!have) {
      chacha_blocks(counter);
      counter += LANES;
      have = 8 * LANES;
    }
    const unsigned i = 8 * LANES - have--;
    return (uint64_t)block[2 * i] | ((uint64_t)block[2 * i + 1] << 32);
  }
};
inline RandomStream &GlobalRandomStream() {
  static thread_local RandomStream s;
  return s;
}
// Test infrastructure switch (see above); returns the previous setting.
inline bool UseTestRandomStream(bool on) {
  RandomStream &s = GlobalRandomStream();
  const bool was = s.test;
  s.test = on;
  return was;
}
inline void SetSeed(const ZZ &seed) {
  RandomStream &s = GlobalRandomStream();
  if (s.test) {  // the oracle's stream: state = the seed's low 64 bits
    uint64_t v = 0;
    for (size_t i = 0; i < seed.mag.size() && i < 2; ++i) v |= (uint64_t)seed.mag[i] << (32 * i);
    s.state = v;
    return;
  }
  s.seed_words(seed.mag.data(), seed.mag.size(), seed.neg ? 1u : 0u);
}
inline ZZ RandomBits_ZZ(long nbits) {
  ZZ r;
  if (nbits <= 0) return r;
  size_t words = (nbits + 63) / 64;
  for (size_t i = 0; i < words; ++i) {
    uint64_t w = GlobalRandomStream().next64();
    r.mag.push_back((uint32_t)w);
    r.mag.push_back((uint32_t)(w >> 32));
  }
  size_t keep = (nbits + 31) / 32;
  r.mag.resize(keep);
  if (nbits % 32) r.mag.back() &= (1u << (nbits % 32)) - 1;
  r.trim();
  return r;
}
inline ZZ RandomBnd(const ZZ &n) {
  if (n <= 1L) return ZZ();
  long nbits = NumBits(n - 1L);
  for (;;) {
    ZZ v = RandomBits_ZZ(nbits);
    if (v < n) return v;
  }
}
inline long RandomBnd(long n) {  // same draws as the ZZ version: one 64-bit word per attempt
  if (n <= 1) return 0;
  const int nbits = 64 - __builtin_clzl((unsigned long)(n - 1));
  const uint64_t mask = nbits >= 64 ? ~0ull : (1ull << nbits) - 1;
  for (;;) {
    uint64_t v = GlobalRandomStream().next64() & mask;
    if (v < (uint64_t)n) return (long)v;
  }
}
inline void RandomBnd(ZZ &x, const ZZ &n) { x = RandomBnd(n); }
// libc rand() stand-in (the reference's sampleHWt draws from lrand48(), which NumbTh.h:32-35 maps
// to rand()): 31 bits of the same stream
inline long RandomLibc31() { return (long)(GlobalRandomStream().next64() >> 33); }

// ------------------------------------------------------------------------------- vectors
template <class T>
class Vec {
 public:
  std::vector<T> v;
  Vec() {}
  long length() const { return (long)v.size(); }
  void SetLength(long n) { v.resize(n); }
  void FixLength(long n) { v.resize(n); }
  void SetMaxLength(long n) { v.reserve(n); }
  void kill() { v.clear(); }
  T &operator[](long i) { return v[i]; }
  const T &operator[](long i) const { return v[i]; }
  T &operator()(long i) { return v[i - 1]; }
  bool operator==(const Vec &o) const { return v == o.v; }
  bool operator!=(const Vec &o) const { return !(v == o.v); }
};
typedef Vec<ZZ> vec_ZZ;
typedef Vec<long> vec_long;

// ------------------------------------------------------------------------------- ZZX
class ZZX {
 public:
  vec_ZZ rep;
  ZZX() {}
  explicit ZZX(long c) {
    if (c) rep.v.push_back(ZZ(c));
  }
  static const ZZX &zero() {
    static const ZZX z;
    return z;
  }
  void normalize() {
    while (!rep.v.empty() && rep.v.back().is_zero()) rep.v.pop_back();
  }
  void SetMaxLength(long n) { rep.SetMaxLength(n); }
  void SetLength(long n) { rep.SetLength(n); }
  bool operator==(const ZZX &o) const { return rep == o.rep; }
  bool operator!=(const ZZX &o) const { return !(rep == o.rep); }
  ZZX &operator+=(const ZZX &o) {
    if (rep.v.size() < o.rep.v.size()) rep.v.resize(o.rep.v.size());
    for (size_t i = 0; i < o.rep.v.size(); ++i) rep.v[i] += o.rep.v[i];
    normalize();
    return *this;
  }
  ZZX &operator-=(const ZZX &o) {
    if (rep.v.size() < o.rep.v.size()) rep.v.resize(o.rep.v.size());
    for (size_t i = 0; i < o.rep.v.size(); ++i) rep.v[i] -= o.rep.v[i];
    normalize();
    return *this;
  }
  ZZX &operator*=(const ZZ &c) {
    for (auto &x : rep.v) x *= c;
    normalize();
    return *this;
  }
  ZZX &operator*=(long c) { return *this *= ZZ(c); }
  ZZX &operator*=(const ZZX &o) {
    if (rep.v.empty() || o.rep.v.empty()) {
      rep.v.clear();
      return *this;
    }
    std::vector<ZZ> r(rep.v.size() + o.rep.v.size() - 1);
    for (size_t i = 0; i < rep.v.size(); ++i) {
      if (rep.v[i].is_zero()) continue;
      for (size_t j = 0; j < o.rep.v.size(); ++j)
        if (!o.rep.v[j].is_zero()) r[i + j] += rep.v[i] * o.rep.v[j];
    }
    rep.v.swap(r);
    normalize();
    return *this;
  }
};
inline long deg(const ZZX &a) { return (long)a.rep.v.size() - 1; }
inline const ZZ &coeff(const ZZX &a, long i) {
  return (i >= 0 && i < (long)a.rep.v.size()) ? a.rep.v[i] : ZZ::zero();
}
inline void SetCoeff(ZZX &a, long i, const ZZ &c) {
  if (i >= (long)a.rep.v.size()) a.rep.v.resize(i + 1);
  a.rep.v[i] = c;
  a.normalize();
}
inline void SetCoeff(ZZX &a, long i, long c) { SetCoeff(a, i, ZZ(c)); }
inline const ZZ &LeadCoeff(const ZZX &a) { return a.rep.v.empty() ? ZZ::zero() : a.rep.v.back(); }
inline void clear(ZZX &a) { a.rep.v.clear(); }
inline bool IsZero(const ZZX &a) { return a.rep.v.empty(); }
inline ZZX operator+(const ZZX &a, const ZZX &b) { ZZX r = a; r += b; return r; }
inline ZZX operator-(const ZZX &a, const ZZX &b) { ZZX r = a; r -= b; return r; }
inline ZZX operator*(const ZZX &a, const ZZX &b) { ZZX r = a; r *= b; return r; }
inline ZZX operator*(const ZZX &a, const ZZ &c) { ZZX r = a; r *= c; return r; }
inline ZZX operator*(const ZZ &c, const ZZX &a) { ZZX r = a; r *= c; return r; }
inline ZZX operator*(const ZZX &a, long c) { ZZX r = a; r *= ZZ(c); return r; }
inline ZZX operator*(long c, const ZZX &a) { ZZX r = a; r *= ZZ(c); return r; }
inline ZZX operator-(const ZZX &a) { ZZX r = a; r *= ZZ(-1L); return r; }
inline ZZX to_ZZX(long c) { return ZZX(c); }
inline ZZX to_ZZX(const ZZ &c) {
  ZZX r;
  if (!c.is_zero()) r.rep.v.push_back(c);
  return r;
}
// remainder by a monic polynomial
inline void rem(ZZX &r, const ZZX &a, const ZZX &f) {
  ZZX t = a;
  long df = deg(f);
  if (df < 0 || !IsOne(f.rep.v.back())) Error("rem: modulus must be monic");
  for (long i = deg(t); i >= df; --i) {
    ZZ c = coeff(t, i);
    if (c.is_zero()) continue;
    if (i >= (long)t.rep.v.size()) continue;
    for (long j = 0; j <= df; ++j)
      if (!f.rep.v[j].is_zero()) t.rep.v[i - df + j] -= c * f.rep.v[j];
  }
  if ((long)t.rep.v.size() > df) t.rep.v.resize(df);
  t.normalize();
  r = t;
}
inline ZZX operator%(const ZZX &a, const ZZX &f) {
  ZZX r;
  rem(r, a, f);
  return r;
}
inline std::ostream &operator<<(std::ostream &os, const ZZX &a) {
  os << "[";
  for (size_t i = 0; i < a.rep.v.size(); ++i) os << (i ? " " : "") << a.rep.v[i];
  return os << "]";
}

// ------------------------------------------------------------------------------- ZZ_p
// Plaintext moduli are word sized in every supported parameter set (p < 2^31).
class ZZ_p {
 public:
  long v = 0;
  static long &mod() {
    static long m = 2;
    return m;
  }
  static ZZ &modZZ() {
    static ZZ m(2L);
    return m;
  }
  static void init(const ZZ &p) {
    if (p.bits() > 31) Error("ZZ_p::init: plaintext modulus must be below 2^31");
    mod() = to_long(p);
    modZZ() = p;
  }
  static const ZZ &modulus() { return modZZ(); }
  ZZ_p() {}
  explicit ZZ_p(long x) { v = ((x % mod()) + mod()) % mod(); }
  ZZ_p &operator+=(const ZZ_p &o) { v = AddMod(v, o.v, mod()); return *this; }
  ZZ_p &operator-=(const ZZ_p &o) { v = SubMod(v, o.v, mod()); return *this; }
  ZZ_p &operator*=(const ZZ_p &o) { v = MulMod(v, o.v, mod()); return *this; }
  ZZ_p &operator*=(long o) { return *this *= ZZ_p(o); }
  ZZ_p &operator=(long x) { return *this = ZZ_p(x); }
  bool operator==(const ZZ_p &o) const { return v == o.v; }
  bool operator!=(const ZZ_p &o) const { return v != o.v; }
  bool operator==(long o) const { return v == ZZ_p(o).v; }
};
inline ZZ_p operator+(ZZ_p a, const ZZ_p &b) { return a += b; }
inline ZZ_p operator-(ZZ_p a, const ZZ_p &b) { return a -= b; }
inline ZZ_p operator*(ZZ_p a, const ZZ_p &b) { return a *= b; }
inline ZZ_p operator-(const ZZ_p &a) { return ZZ_p(-a.v); }
inline ZZ_p to_ZZ_p(long x) { return ZZ_p(x); }
inline ZZ_p to_ZZ_p(const ZZ &x) { return ZZ_p(to_long(x % ZZ_p::modulus())); }
inline ZZ rep(const ZZ_p &a) { return ZZ(a.v); }
inline bool IsZero(const ZZ_p &a) { return a.v == 0; }
inline ZZ_p inv(const ZZ_p &a) { return ZZ_p(InvMod(a.v, ZZ_p::mod())); }
inline ZZ_p power(const ZZ_p &a, long e) { return ZZ_p(PowerMod(a.v, e, ZZ_p::mod())); }
inline ZZ_p random_ZZ_p() { return ZZ_p(RandomBnd(ZZ_p::mod())); }
inline void random(ZZ_p &x) { x = random_ZZ_p(); }
inline std::ostream &operator<<(std::ostream &os, const ZZ_p &a) { return os << a.v; }
typedef Vec<ZZ_p> vec_ZZ_p;

// ------------------------------------------------------------------------------- ZZ_pX
class ZZ_pX {
 public:
  vec_ZZ_p rep;
  ZZ_pX() {}
  ZZ_pX(INIT_SIZE_TYPE, long n) { rep.v.reserve(n); }
  static const ZZ_pX &zero() {
    static const ZZ_pX z;
    return z;
  }
  void normalize() {
    while (!rep.v.empty() && rep.v.back().v == 0) rep.v.pop_back();
  }
  void SetMaxLength(long n) { rep.SetMaxLength(n); }
  void SetLength(long n) { rep.SetLength(n); }
  bool operator==(const ZZ_pX &o) const { return rep == o.rep; }
  bool operator!=(const ZZ_pX &o) const { return !(rep == o.rep); }
  ZZ_pX &operator+=(const ZZ_pX &o) {
    if (rep.v.size() < o.rep.v.size()) rep.v.resize(o.rep.v.size());
    for (size_t i = 0; i < o.rep.v.size(); ++i) rep.v[i] += o.rep.v[i];
    normalize();
    return *this;
  }
  ZZ_pX &operator-=(const ZZ_pX &o) {
    if (rep.v.size() < o.rep.v.size()) rep.v.resize(o.rep.v.size());
    for (size_t i = 0; i < o.rep.v.size(); ++i) rep.v[i] -= o.rep.v[i];
    normalize();
    return *this;
  }
  ZZ_pX &operator*=(const ZZ_p &c) {
    for (auto &x : rep.v) x *= c;
    normalize();
    return *this;
  }
  ZZ_pX &operator*=(long c) { return *this *= ZZ_p(c); }
  ZZ_pX &operator*=(const ZZ_pX &o) {
    if (rep.v.empty() || o.rep.v.empty()) {
      rep.v.clear();
      return *this;
    }
    const long m = ZZ_p::mod();
    std::vector<unsigned __int128> acc(rep.v.size() + o.rep.v.size() - 1, 0);
    for (size_t i = 0; i < rep.v.size(); ++i) {
      if (!rep.v[i].v) continue;
      for (size_t j = 0; j < o.rep.v.size(); ++j) acc[i + j] += (unsigned __int128)rep.v[i].v * o.rep.v[j].v;
    }
    rep.v.resize(acc.size());
    for (size_t i = 0; i < acc.size(); ++i) rep.v[i].v = (long)(acc[i] % (unsigned long)m);
    normalize();
    return *this;
  }
};
inline long deg(const ZZ_pX &a) { return (long)a.rep.v.size() - 1; }
inline ZZ_p coeff(const ZZ_pX &a, long i) { return (i >= 0 && i < (long)a.rep.v.size()) ? a.rep.v[i] : ZZ_p(); }
inline void SetCoeff(ZZ_pX &a, long i, const ZZ_p &c) {
  if (i >= (long)a.rep.v.size()) a.rep.v.resize(i + 1);
  a.rep.v[i] = c;
  a.normalize();
}
inline void SetCoeff(ZZ_pX &a, long i, long c) { SetCoeff(a, i, ZZ_p(c)); }
inline void SetCoeff(ZZ_pX &a, long i) { SetCoeff(a, i, ZZ_p(1)); }
inline void clear(ZZ_pX &a) { a.rep.v.clear(); }
inline bool IsZero(const ZZ_pX &a) { return a.rep.v.empty(); }
inline ZZ_p LeadCoeff(const ZZ_pX &a) { return a.rep.v.empty() ? ZZ_p() : a.rep.v.back(); }
inline ZZ_p ConstTerm(const ZZ_pX &a) { return a.rep.v.empty() ? ZZ_p() : a.rep.v[0]; }
inline ZZ_pX operator+(const ZZ_pX &a, const ZZ_pX &b) { ZZ_pX r = a; r += b; return r; }
inline ZZ_pX operator-(const ZZ_pX &a, const ZZ_pX &b) { ZZ_pX r = a; r -= b; return r; }
inline ZZ_pX operator*(const ZZ_pX &a, const ZZ_pX &b) { ZZ_pX r = a; r *= b; return r; }
inline ZZ_pX operator*(const ZZ_pX &a, long c) { ZZ_pX r = a; r *= ZZ_p(c); return r; }
inline ZZ_pX operator*(long c, const ZZ_pX &a) { ZZ_pX r = a; r *= ZZ_p(c); return r; }
inline ZZ_pX operator*(const ZZ_pX &a, const ZZ_p &c) { ZZ_pX r = a; r *= c; return r; }
inline ZZ_pX operator*(const ZZ_p &c, const ZZ_pX &a) { ZZ_pX r = a; r *= c; return r; }
inline void DivRem(ZZ_pX &q, ZZ_pX &r, const ZZ_pX &a, const ZZ_pX &b) {
  if (b.rep.v.empty()) Error("ZZ_pX: division by zero");
  ZZ_pX t = a;
  long db = deg(b);
  ZZ_p li = inv(b.rep.v.back());
  q.rep.v.assign(deg(a) >= db ? deg(a) - db + 1 : 0, ZZ_p());
  for (long i = deg(t); i >= db; --i) {
    ZZ_p c = t.rep.v[i] * li;
    if (!c.v) continue;
    q.rep.v[i - db] = c;
    for (long j = 0; j <= db; ++j) t.rep.v[i - db + j] -= c * b.rep.v[j];
  }
  if ((long)t.rep.v.size() > db) t.rep.v.resize(db);
  t.normalize();
  q.normalize();
  r = t;
}
inline void rem(ZZ_pX &r, const ZZ_pX &a, const ZZ_pX &b) {
  ZZ_pX q;
  DivRem(q, r, a, b);
}
inline ZZ_pX operator%(const ZZ_pX &a, const ZZ_pX &b) { ZZ_pX q, r; DivRem(q, r, a, b); return r; }
inline ZZ_pX operator/(const ZZ_pX &a, const ZZ_pX &b) { ZZ_pX q, r; DivRem(q, r, a, b); return q; }
inline ZZ_pX &operator%=(ZZ_pX &a, const ZZ_pX &b) { a = a % b; return a; }
inline ZZ_pX &operator/=(ZZ_pX &a, const ZZ_pX &b) { a = a / b; return a; }
inline void MulMod(ZZ_pX &x, const ZZ_pX &a, const ZZ_pX &b, const ZZ_pX &f) { x = (a * b) % f; }
inline ZZ_pX MulMod(const ZZ_pX &a, const ZZ_pX &b, const ZZ_pX &f) { return (a * b) % f; }
inline void InvMod(ZZ_pX &x, const ZZ_pX &a, const ZZ_pX &f) {  // extended Euclid over Z_p[X]
  ZZ_pX r0 = f, r1 = a % f, s0, s1;
  SetCoeff(s1, 0, 1);
  while (!IsZero(r1)) {
    ZZ_pX q, r;
    DivRem(q, r, r0, r1);
    r0 = r1, r1 = r;
    ZZ_pX t = s0 - q * s1;
    s0 = s1, s1 = t;
  }
  if (deg(r0) != 0) Error("InvMod(ZZ_pX): not invertible");
  x = (s0 * inv(r0.rep.v[0])) % f;
}
inline void random(ZZ_pX &x, long n) {
  x.rep.v.resize(n);
  for (long i = 0; i < n; ++i) x.rep.v[i] = random_ZZ_p();
  x.normalize();
}
inline ZZ_pX to_ZZ_pX(long c) { ZZ_pX r; SetCoeff(r, 0, ZZ_p(c)); return r; }
inline ZZ_pX to_ZZ_pX(int c) { return to_ZZ_pX((long)c); }
inline ZZ_pX to_ZZ_pX(unsigned c) { return to_ZZ_pX((long)c); }
inline ZZ_pX to_ZZ_pX(unsigned long c) { return to_ZZ_pX((long)(c % (unsigned long)ZZ_p::mod())); }
inline ZZ_pX to_ZZ_pX(const ZZ_p &c) { ZZ_pX r; SetCoeff(r, 0, c); return r; }
inline ZZ_pX to_ZZ_pX(const ZZ &c) { return to_ZZ_pX(to_ZZ_p(c)); }
inline const ZZ_pX &to_ZZ_pX(const ZZ_pX &a) { return a; }
inline ZZ_pX to_ZZ_pX(const ZZX &a) {
  ZZ_pX r;
  r.rep.v.resize(a.rep.v.size());
  for (size_t i = 0; i < a.rep.v.size(); ++i) r.rep.v[i] = to_ZZ_p(a.rep.v[i]);
  r.normalize();
  return r;
}
inline ZZX to_ZZX(const ZZ_pX &a) {
  ZZX r;
  r.rep.v.resize(a.rep.v.size());
  for (size_t i = 0; i < a.rep.v.size(); ++i) r.rep.v[i] = ZZ(a.rep.v[i].v);
  r.normalize();
  return r;
}
inline void conv(ZZX &x, const ZZ_pX &a) { x = to_ZZX(a); }
inline void conv(ZZ_pX &x, const ZZX &a) { x = to_ZZ_pX(a); }
inline ZZ_p eval(const ZZ_pX &a, const ZZ_p &x) {
  ZZ_p acc;
  for (size_t i = a.rep.v.size(); i-- > 0;) acc = acc * x + a.rep.v[i];
  return acc;
}
inline std::ostream &operator<<(std::ostream &os, const ZZ_pX &a) {
  os << "[";
  for (size_t i = 0; i < a.rep.v.size(); ++i) os << (i ? " " : "") << a.rep.v[i];
  return os << "]";
}
typedef Vec<ZZ_pX> vec_ZZ_pX;

// extended-exponent double: long double has range to spare for the parameter sizing it is used for
class xdouble {
 public:
  long double v = 0;
  xdouble() {}
  xdouble(double d) : v(d) {}
  xdouble(long double d, int) : v(d) {}
};
inline xdouble to_xdouble(const ZZ &a) {
  long double r = 0;
  for (size_t i = a.mag.size(); i-- > 0;) r = r * 4294967296.0L + a.mag[i];
  return xdouble(a.neg ? -r : r, 0);
}
inline xdouble to_xdouble(double d) { return xdouble(d); }
inline xdouble to_xdouble(long d) { return xdouble((long double)d, 0); }
inline void conv(xdouble &x, const ZZ &a) { x = to_xdouble(a); }
inline void conv(xdouble &x, double a) { x = xdouble(a); }
inline double to_double(const xdouble &x) { return (double)x.v; }
inline double log(const xdouble &x) { return (double)std::log(x.v); }
inline xdouble operator*(const xdouble &a, const xdouble &b) { return xdouble(a.v * b.v, 0); }
inline xdouble operator/(const xdouble &a, const xdouble &b) { return xdouble(a.v / b.v, 0); }
inline xdouble operator+(const xdouble &a, const xdouble &b) { return xdouble(a.v + b.v, 0); }
inline xdouble operator-(const xdouble &a, const xdouble &b) { return xdouble(a.v - b.v, 0); }
inline xdouble &operator*=(xdouble &a, const xdouble &b) { a.v *= b.v; return a; }
inline xdouble &operator/=(xdouble &a, const xdouble &b) { a.v /= b.v; return a; }
inline xdouble &operator+=(xdouble &a, const xdouble &b) { a.v += b.v; return a; }
inline bool operator<(const xdouble &a, const xdouble &b) { return a.v < b.v; }
inline bool operator>(const xdouble &a, const xdouble &b) { return a.v > b.v; }
inline std::ostream &operator<<(std::ostream &os, const xdouble &x) { return os << (double)x.v; }

}  // namespace NTL

#define NTL_CLIENT      \
  using namespace std;  \
  using namespace NTL;
#define NTL_SP_NBITS 60
#define NTL_SP_BOUND (1L << NTL_SP_NBITS)
