// fhesi_host.h -- the reference's C++ class surface re-implemented from scratch
// on top of the C ABI in include/fhesi.h.  Client sources written against dwu4/fhe-si
// (Test_AddMul.cpp, Test_General.cpp, Regression.h, Statistics.h, Matrix.*) compile unchanged
// with `-I fhe-si_b200/host`: the headers they name (FHEContext.h, FHE-SI.h, Ciphertext.h,
// DoubleCRT.h, Plaintext.h, PlaintextSpace.h, Util.h, Serialization.h, NTL/*.h, ...) all forward
// here.
//
// Where things live
//   Ciphertext      parts / tProd are HBM buffers owned through the C ABI; every operator is one
//                   or more kernel launches; `parts` (public in the reference, Ciphertext.h:68)
//                   is a host mirror refreshed by SyncHost() -- no client touches it directly.
//   DoubleCRT       host object holding the exact coefficient polynomial the reference's matrix
//                   represents (centred mod the chain product); rows over the reference chain are
//                   produced on demand by fhesi_ref_rows_host for Export.  Only keys are DoubleCRT
//                   in the client-visible API (FHE-SI.h:31-32,64-65,105-106).
//   keys            generated on the host (set-up time), uploaded once (fhesi_key_create /
//                   fhesi_ksw_create) on first use.
// Errors: like the reference (NTL::Error / assert) the layer prints and aborts.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "fhesi.h"
#include "ntl_shim.h"

using namespace std;
using namespace NTL;

// ------------------------------------------------------------------------------- PAlgebra
// Structure of Z_m^* (PAlgebra.h:30-88, PAlgebra.cpp:40-56)
class PAlgebra {
  unsigned m = 0, g = 0, phim = 0;
  vector<long> zmsIdx;
  ZZX Phi_mX;

 public:
  void init(unsigned mm, unsigned gen);
  unsigned M() const { return m; }
  unsigned phiM() const { return phim; }
  const ZZX &PhimX() const { return Phi_mX; }
  bool inZmStar(unsigned t) const { return t < m && zmsIdx[t] >= 0; }
  long indexInZmstar(unsigned t) const { return t < m ? zmsIdx[t] : -1; }
  bool operator==(const PAlgebra &o) const { return m == o.m; }
};

// a dense index set {0..n-1}: all the SI path ever uses (SURVEY.md §2 "Index containers")
class IndexSet {
  long n = 0;

 public:
  IndexSet() {}
  IndexSet(long lo, long hi) : n(hi + 1) { assert(lo == 0); }
  void insert(long i) { if (i + 1 > n) n = i + 1; }
  long card() const { return n; }
  long first() const { return 0; }
  long last() const { return n - 1; }
  long next(long i) const { return i + 1; }
  bool contains(long i) const { return i >= 0 && i < n; }
  bool operator==(const IndexSet &o) const { return n == o.n; }
};
inline long card(const IndexSet &s) { return s.card(); }

// The reference's IndexMap<T> (IndexMap.h): a map over a dynamic index set.  Index sets on the
// Brakerski path are always the whole chain {0..L-1}, so a vector is enough.
template <class T>
class IndexMap {
  std::vector<T> items;
  IndexSet set;

 public:
  IndexMap() {}
  const IndexSet &getIndexSet() const { return set; }
  T &operator[](long j) { assert(set.contains(j)); return items[j]; }
  const T &operator[](long j) const { assert(set.contains(j)); return items[j]; }
  void insert(long j) {
    set.insert(j);
    if ((long)items.size() < set.card()) items.resize(set.card());
  }
  void insert(const IndexSet &s) {
    for (long i = s.first(); i <= s.last(); i = s.next(i)) insert(i);
  }
  void clear() { items.clear(), set = IndexSet(); }
};

// ------------------------------------------------------------------------------- PlaintextSpace
// Slot structure of Z_p[X]/Phi_m (PlaintextSpace.h, PlaintextSpace.cpp:22-134) for p = 1 mod m,
// where Phi_m splits into linear factors: slot j <-> root rho^(g^j).
class PlaintextSpace {
 public:
  PlaintextSpace() {}
  void Init(const ZZX &PhiX, const ZZ &p);
  void Init(const ZZX &PhiX, const ZZ &p, unsigned generator);
  unsigned GetUsableSlots() const { return usableSlots; }
  unsigned GetTotalSlots() const { return totalSlots; }
  void EmbedInSlots(ZZ_pX &embedded, const vector<ZZ_pX> &msgs, bool onlyUsable = true) const;
  void DecodeSlots(vector<ZZ_pX> &msgBatch, const ZZ_pX &msg, bool onlyUsable = true) const;
  void DecodeSlot(ZZ_pX &val, const ZZ_pX &msg, unsigned ind) const;

  ZZ p;
  unsigned generator = 0;

 private:
  unsigned m = 0, totalSlots = 0, usableSlots = 0;
  vector<long> roots;            // roots[j] = rho^(g^j) mod p
  void EnsureBasis() const;
  vector<long> phiModP;                  // Phi_m mod p
  mutable vector<vector<long>> basis;    // basis[j] = CRT idempotent of slot j, phi(m) coefficients (lazy)
  mutable vector<uint32_t> basis32;      // the same, flat, when p < 2^26 (vectorisable embedding)
  friend class FHEcontext;
};

// ------------------------------------------------------------------------------- FHEcontext
class FHEcontext;
extern FHEcontext *activeContext;  // FHEContext.h:43, FHEContext.cpp:21

struct CmodulusInfo {  // the client-visible part of Cmodulus (CModulus.h:42-170)
  long q = 0, root = 0;
  long getQ() const { return q; }
  long getRoot() const { return root; }
};

class FHEcontext {
  vector<CmodulusInfo> moduli;
  PlaintextSpace ptxtSpace;
  mutable fhesi_ctx *dev = nullptr;
  mutable long devXi = 0;
  long xiHint = 1;

 public:
  PAlgebra zMstar;
  IndexSet ctxtPrimes, specialPrimes;
  double stdev = 3.2;
  ZZ modulusQ;
  unsigned logQ = 0, decompSize = 3, ndigits = 0, primesNeeded = 0;
  int device = 0;  // CUDA device this context's image lives on

  FHEcontext(unsigned m, unsigned logQ, unsigned p, unsigned generator, unsigned decompSize = 3) {
    Init(m, logQ, to_ZZ(p), generator, decompSize);
  }
  FHEcontext(unsigned m, unsigned logQ, const ZZ &p, unsigned generator, unsigned decompSize = 3) {
    Init(m, logQ, p, generator, decompSize);
  }
  FHEcontext(ifstream &in) { ImportSIContext(in); }
  FHEcontext(const FHEcontext &) = delete;
  ~FHEcontext();

  void Init(unsigned m, unsigned logQ, const ZZ &p, unsigned generator, unsigned decompSize = 3);
  void ExportSIContext(ofstream &out);
  void ImportSIContext(ifstream &in);
  void SetUpSIContext(long xi = 1);

  unsigned Generator() const { return ptxtSpace.generator; }
  const ZZ &ModulusP() const { return ptxtSpace.p; }
  const PlaintextSpace &GetPlaintextSpace() const { return ptxtSpace; }
  long ithPrime(unsigned i) const { return i < moduli.size() ? moduli[i].q : 0; }
  const CmodulusInfo &ithModulus(unsigned i) const { return moduli[i]; }
  long numPrimes() const { return (long)moduli.size(); }
  bool inChain(long p) const {
    for (auto &mo : moduli)
      if (mo.q == p) return true;
    return false;
  }
  void productOfPrimes(ZZ &p, const IndexSet &s) const {
    p = 1;
    for (long i = s.first(); i <= s.last(); i = s.next(i)) p *= ZZ(ithPrime(i));
  }
  ZZ productOfPrimes(const IndexSet &s) const { ZZ p; productOfPrimes(p, s); return p; }
  ZZ productOfPrimes() const { return productOfPrimes(ctxtPrimes); }
  void AddPrime(long p, bool special, long root = 0);

  // the device image (created on first use; xi as given to SetUpSIContext)
  fhesi_ctx *Dev() const;
  unsigned Words() const { return (logQ + 31) / 32; }

  friend ostream &operator<<(ostream &os, const FHEcontext &context);
};
double AddPrimesBySize(FHEcontext &context, double totalSize, bool special = false);

// ------------------------------------------------------------------------------- Util.h
void Reduce(ZZ &val, unsigned logQ, bool positive = false);
void ReduceCoefficients(ZZX &poly, unsigned logQ, bool positive = false);
void SampleRandom(ZZX &poly, const ZZ &modulus, unsigned deg);
template <typename T>
void PrintVector(const vector<T> &vec, ostream &out = std::cout) {
  for (unsigned i = 0; i < vec.size(); i++) out << vec[i] << " ";
}
template <typename T>
void PrintVector(const vector<vector<T>> &vec, ostream &out = std::cout) {
  for (unsigned i = 0; i < vec.size(); i++) {
    PrintVector(vec[i], out);
    out << endl;
  }
}
template <typename T>
static void DotProduct(T &res, const vector<T> &v1, const vector<T> &v2) {
  if (v1.empty()) return;
  res = v1[0];
  res *= v2[0];
  for (unsigned i = 1; i < v1.size(); i++) {
    T val = v1[i];
    val *= v2[i];
    res += val;
  }
}
// NumbTh.h samplers (NumbTh.cpp:340-404) on the shared SplitMix64 stream
void sampleHWt(ZZX &poly, long Hwt, long n = 0);
void sampleSmall(ZZX &poly, long n = 0);
void sampleGaussian(ZZX &poly, long n = 0, double stdev = 1.0);
ZZX Cyclotomic(long N);

// ------------------------------------------------------------------------------- DoubleCRT
class DoubleCRT {
  const FHEcontext &context;
  ZZX poly;  // centred mod productOfPrimes(): exactly what toPoly() yields in the reference
  void wrap();

 public:
  DoubleCRT();
  DoubleCRT(const FHEcontext &context);
  DoubleCRT(const ZZX &poly);
  DoubleCRT(const ZZX &poly, const FHEcontext &context);
  DoubleCRT(const DoubleCRT &o) : context(o.context), poly(o.poly) {}
  DoubleCRT &operator=(const DoubleCRT &other);
  DoubleCRT &operator=(const ZZX &p);
  DoubleCRT &operator=(const ZZ &num);
  DoubleCRT &operator=(long num) { return *this = to_ZZ(num); }

  DoubleCRT &operator+=(const DoubleCRT &o);
  DoubleCRT &operator-=(const DoubleCRT &o);
  DoubleCRT &operator*=(const DoubleCRT &o);
  DoubleCRT &operator+=(const ZZX &p) { return *this += DoubleCRT(p, context); }
  DoubleCRT &operator-=(const ZZX &p) { return *this -= DoubleCRT(p, context); }
  DoubleCRT &operator*=(const ZZX &p) { return *this *= DoubleCRT(p, context); }
  DoubleCRT &operator+=(const ZZ &c);
  DoubleCRT &operator-=(const ZZ &c) { return *this += -c; }
  DoubleCRT &operator*=(const ZZ &c);
  DoubleCRT &operator+=(long c) { return *this += to_ZZ(c); }
  DoubleCRT &operator-=(long c) { return *this -= to_ZZ(c); }
  DoubleCRT &operator*=(long c) { return *this *= to_ZZ(c); }

  void toPoly(ZZX &p, bool positive = false) const;
  void automorph(long k);
  DoubleCRT &operator>>=(long k) { automorph(k); return *this; }
  DoubleCRT &operator/=(const ZZ &num);                    // DoubleCRT.cpp:406-420
  DoubleCRT &operator/=(long num) { return *this /= to_ZZ(num); }
  void Exp(long e);                                        // :422-435, pointwise power of the rows
  void randomize(const ZZ *seed = NULL);                   // :466-480, uniformly random rows
  const FHEcontext &getContext() const { return context; }
  // rows over the reference chain as the reference exposes them (DoubleCRT.h:296-297); computed on demand
  IndexMap<vec_long> getMap() const;
  IndexSet getIndexSet() const;

  // rows over the context's (reference) chain: DoubleCRT(const ZZX&), DoubleCRT.cpp:244-257
  vector<vector<long>> getRows() const;
  void setRows(const vector<vector<long>> &rows);  // inverse transform + incremental CRT

  void sampleSmall() { ZZX p; ::sampleSmall(p, context.zMstar.phiM()); *this = p; }
  void sampleHWt(long Hwt) { ZZX p; ::sampleHWt(p, Hwt, context.zMstar.phiM()); *this = p; }
  void sampleGaussian(double sd = 0.0) {
    ZZX p;
    ::sampleGaussian(p, context.zMstar.phiM(), sd == 0.0 ? context.stdev : sd);
    *this = p;
  }
  friend ostream &operator<<(ostream &os, const DoubleCRT &d) { return os << d.poly; }
};
inline ZZX to_ZZX(const DoubleCRT &d) { ZZX p; d.toPoly(p); return p; }

// ------------------------------------------------------------------------------- Plaintext
class Plaintext {
 public:
  Plaintext() : context(*activeContext) {}
  Plaintext(const FHEcontext &context) : context(context) {}
  Plaintext(const ZZ_pX &msg) : context(*activeContext) { Init(msg); }
  Plaintext(const FHEcontext &context, const ZZ_pX &msg) : context(context) { Init(msg); }
  template <typename T>
  Plaintext(const T &msg) : context(*activeContext) { Init(to_ZZ_pX(msg)); }
  template <typename T>
  Plaintext(const FHEcontext &context, const T &msg) : context(context) { Init(to_ZZ_pX(msg)); }
  Plaintext(const vector<ZZ_pX> &msgs) : context(*activeContext) { Init(msgs); }
  Plaintext(const FHEcontext &context, const vector<ZZ_pX> &msgs) : context(context) { Init(msgs); }
  template <typename T>
  Plaintext(const vector<T> &msgs) : context(*activeContext) { Init(msgs); }
  template <typename T>
  Plaintext(const FHEcontext &context, const vector<T> &msgs) : context(context) { Init(msgs); }
  Plaintext(const Plaintext &o) : message(o.message), context(o.context) {}

  void Init() {}
  void Init(const ZZ_pX &msg) { message = msg; }
  void Init(const vector<ZZ_pX> &msgs) { EmbedInSlots(msgs); }
  template <typename T>
  void Init(const vector<T> &msgs) {
    vector<ZZ_pX> m(msgs.size());
    for (unsigned i = 0; i < msgs.size(); i++) m[i] = to_ZZ_pX(msgs[i]);
    EmbedInSlots(m);
  }
  void EmbedInSlots(const vector<ZZ_pX> &msgs, bool onlyUsable = true) {
    context.GetPlaintextSpace().EmbedInSlots(message, msgs, onlyUsable);
  }
  void DecodeSlots(vector<ZZ_pX> &msgBatch, bool onlyUsable = true) {
    context.GetPlaintextSpace().DecodeSlots(msgBatch, message, onlyUsable);
  }
  void DecodeSlot(ZZ_pX &val, unsigned slot) { context.GetPlaintextSpace().DecodeSlot(val, message, slot); }
  Plaintext &operator=(const Plaintext &other) {
    assert(&context == &other.context);
    message = other.message;
    return *this;
  }
  bool operator==(const Plaintext &other) const { return &context == &other.context && message == other.message; }
  bool operator==(const ZZ_pX &other) const { return message == other; }
  friend ostream &operator<<(ostream &os, const Plaintext &ptxt) { return os << ptxt.message; }

  ZZ_pX message;

  void Randomize() { random(message, context.zMstar.phiM()); }
  static Plaintext Random(const FHEcontext &context) { Plaintext r(context); r.Randomize(); return r; }
  Plaintext &operator+=(const Plaintext &o) { message += o.message; return *this; }
  Plaintext &operator-=(const Plaintext &o) { message -= o.message; return *this; }
  Plaintext &operator*=(const Plaintext &o) {
    MulMod(message, message, o.message, to_ZZ_pX(context.zMstar.PhimX()));
    return *this;
  }
  Plaintext &operator>>=(long k) {
    vector<ZZ_pX> a;
    DecodeSlots(a, false);
    vector<ZZ_pX> r = a;
    for (unsigned i = 0; i < a.size(); i++) r[(i + a.size() - k) % a.size()] = a[i];
    EmbedInSlots(r, false);
    return *this;
  }
  Plaintext operator+(const Plaintext &o) { return Plaintext(context, message + o.message); }
  Plaintext operator*(const Plaintext &o) {
    return Plaintext(context, MulMod(message, o.message, to_ZZ_pX(context.zMstar.PhimX())));
  }

 private:
  const FHEcontext &context;
};

// ------------------------------------------------------------------------------- Ciphertext
class FHESIPubKey;
class FHESISecKey;
class KeySwitchSI;

class CiphertextPart {
  const FHEcontext &context;

 public:
  ZZX poly;
  CiphertextPart() : context(*activeContext) {}
  CiphertextPart(const FHEcontext &context) : context(context) {}
  CiphertextPart(const long val) : context(*activeContext) { poly = to_ZZX(val); }
  explicit CiphertextPart(const ZZX &poly) : context(*activeContext) { this->poly = poly; }
  CiphertextPart(const CiphertextPart &o) : context(o.context), poly(o.poly) {}
  bool operator==(const CiphertextPart &o) const { return poly == o.poly; }
  CiphertextPart &operator=(const CiphertextPart &o) {
    if (&context != &o.context) Error("Incompatible contexts.");
    poly = o.poly;
    return *this;
  }
  friend ostream &operator<<(ostream &os, const CiphertextPart &c) { return os << c.poly; }
};

// a device allocation owned through the C ABI
struct DevBuf {
  fhesi_ctx *ctx = nullptr;
  uint32_t *ptr = nullptr;
  size_t bytes = 0;
  DevBuf(fhesi_ctx *c, size_t b);
  ~DevBuf();
  DevBuf(const DevBuf &) = delete;
};

class Ciphertext {
  const FHEcontext *context;
  shared_ptr<DevBuf> buf;  // !scaledUp: [nparts][n][wordsPer]; scaledUp: tprod [nparts][Lt][N]
  unsigned nparts = 0;
  unsigned wordsPer = 0;   // W, or W+1 right after >>= (not reduced mod q, Ciphertext.cpp:54-59)
  bool scaledUp = false;
  mutable bool hostStale = false;
  // set by the non-const operator[]: the caller may have written through the reference (the reference's
  // `ctxt[i].poly = ...`, FHE-SI.cpp:29), so the host mirror is the truth until the next operator flushes it
  mutable bool hostDirty = false;
  void Flush() const;

  void Alloc(unsigned parts, unsigned words);
  void EnsureReduced();     // Reduce a wide (W+1 words) image to W words
  void UploadHost();        // parts (host) -> device
  friend class FHESIPubKey;
  friend class FHESISecKey;
  friend class KeySwitchSI;
  friend void Export(ofstream &out, const Ciphertext &ctxt);
  friend void Import(ifstream &in, Ciphertext &ctxt);

 public:
  Ciphertext() : context(activeContext) {}
  Ciphertext(const FHEcontext &context) : context(&context) {}
  Ciphertext(const FHESIPubKey &pk);
  Ciphertext(const Ciphertext &other);

  vector<CiphertextPart> parts;  // host mirror; call SyncHost() before reading
  void SyncHost() const;

  void Initialize(unsigned n, const FHEcontext &context);
  unsigned size() const { return nparts; }
  bool IsScaledUp() const { return scaledUp; }

  Ciphertext &operator+=(const Ciphertext &other);
  Ciphertext &operator+=(const ZZX &other);
  Ciphertext &operator+=(const ZZ_pX &other) { return operator+=(to_ZZX(other)); }
  Ciphertext &operator*=(const Ciphertext &other);
  Ciphertext &operator*=(long l);
  Ciphertext &operator*=(int l) { return operator*=((long)l); }
  Ciphertext &operator*=(const ZZX &other);
  Ciphertext &operator*=(const ZZ_pX &other) { return operator*=(to_ZZX(other)); }
  Ciphertext &operator>>=(long k);

  void Clear();
  void ScaleDown();
  CiphertextPart GetPart(unsigned ind) const;
  CiphertextPart &operator[](unsigned ind);
  Ciphertext &operator=(const Ciphertext &other);
  friend ostream &operator<<(ostream &os, const Ciphertext &ctxt);

  // raw access for batch-aware callers
  const uint32_t *DevWords() const { Flush(); return buf ? buf->ptr : nullptr; }
};

// ------------------------------------------------------------------------------- keys
class FHESISecKey {
 public:
  FHESISecKey() : context(*activeContext) { Init(*activeContext); }
  FHESISecKey(const FHEcontext &context) : context(context) { Init(context); }
  void Init(const FHEcontext &context);
  void Decrypt(Plaintext &plaintext, const Ciphertext &ciphertext) const;
  const FHEcontext &GetContext() const { return context; }
  size_t GetSize() const { return sKeys.size(); }
  const vector<DoubleCRT> &GetRepresentation() const { return sKeys; }
  void UpdateRepresentation(vector<DoubleCRT> &rep) { sKeys = rep; devKey.reset(); }
  void Export(ofstream &out) const;
  void Import(ifstream &in);
  friend ostream &operator<<(ostream &os, const FHESISecKey &k) {
    for (auto &d : k.sKeys) os << d;
    return os;
  }

 private:
  // an empty key to be filled by UpdateRepresentation (KeySwitchSI::InitS2 / InitAutomorph build
  // their source keys this way; the reference constructs-and-discards a random key there, which
  // only burns randomness)
  struct Empty {};
  FHESISecKey(const FHEcontext &context, Empty) : context(context) {}
  friend class KeySwitchSI;
  vector<DoubleCRT> sKeys;
  const FHEcontext &context;
  mutable shared_ptr<fhesi_key> devKey;
};

class FHESIPubKey {
  friend class Ciphertext;
  const FHEcontext &context;
  // Generated on the device (Init); the host image is made from the downloaded words only when a
  // caller asks for it (GetRepresentation, Export, operator<<)
  mutable vector<DoubleCRT> publicKey;
  mutable std::vector<uint32_t> hostWords;  // [2][n][W], pending materialisation when non-empty
  mutable shared_ptr<fhesi_key> devKey;
  void Materialize() const;

 public:
  FHESIPubKey(const FHEcontext &context) : context(context) {}
  FHESIPubKey(const FHESISecKey &secKey) : context(secKey.GetContext()) { Init(secKey); }
  FHESIPubKey(const FHESISecKey &secKey, const FHEcontext &context) : context(context) { Init(secKey); }
  void Encrypt(Ciphertext &ctxt, const Plaintext &ptxt) const;
  void Init(const FHESISecKey &secKey);
  const FHEcontext &GetContext() const { return context; }
  const vector<DoubleCRT> &GetRepresentation() const { Materialize(); return publicKey; }
  void UpdateRepresentation(const vector<DoubleCRT> &rep) { publicKey = rep; hostWords.clear(); devKey.reset(); }
  void Export(ofstream &out) const;
  void Import(ifstream &in);
  friend ostream &operator<<(ostream &os, const FHESIPubKey &k) {
    k.Materialize();
    return os << k.publicKey[0] << ", " << k.publicKey[1];
  }
};

class KeySwitchSI {
 public:
  KeySwitchSI() : context(*activeContext) {}
  KeySwitchSI(FHEcontext &context) : context(context) {}
  KeySwitchSI(const FHESISecKey &src, const FHESISecKey &dst) : context(*activeContext) { Init(src, dst); }
  KeySwitchSI(const FHESISecKey &src, const FHESISecKey &dst, const FHEcontext &context) : context(context) {
    Init(src, dst);
  }
  KeySwitchSI(const FHESISecKey &s) : context(*activeContext) { InitS2(s); }
  KeySwitchSI(const FHESISecKey &s, const FHEcontext &context) : context(context) { InitS2(s); }
  KeySwitchSI(const FHESISecKey &s, unsigned k) : context(*activeContext) { InitAutomorph(s, k); }
  KeySwitchSI(const FHESISecKey &s, const FHEcontext &context, unsigned k) : context(context) {
    InitAutomorph(s, k);
  }
  KeySwitchSI(const KeySwitchSI &o)
      : context(o.context), keySwitchMatrix(o.keySwitchMatrix), hostB(o.hostB), drawA(o.drawA), entries(o.entries),
        devKsw(o.devKsw) {}

  void Init(const FHESISecKey &src, const FHESISecKey &dst);
  void InitS2(const FHESISecKey &s);
  void InitAutomorph(const FHESISecKey &s, unsigned k);
  void ApplyKeySwitch(Ciphertext &ctxt) const;
  const vector<vector<DoubleCRT>> &GetRepresentation() const { Materialize(); return keySwitchMatrix; }
  void UpdateRepresentation(const vector<vector<DoubleCRT>> &rep) {
    keySwitchMatrix = rep;
    hostB.clear(), drawA.clear();
    entries = rep.empty() ? 0 : rep[0].size();
    devKsw.reset();
  }
  void Export(ofstream &out) const;
  void Import(ifstream &in);
  KeySwitchSI &operator=(const KeySwitchSI &other) {
    if (&context != &other.context) Error("KeySwitchSI assignment: context mismatch");
    keySwitchMatrix = other.keySwitchMatrix;
    hostB = other.hostB, drawA = other.drawA, entries = other.entries;
    devKsw = other.devKsw;
    return *this;
  }
  friend ostream &operator<<(ostream &os, const KeySwitchSI &k) {
    k.Materialize();
    PrintVector(k.keySwitchMatrix, os);
    return os;
  }
  // the uploaded matrix, for batch-aware callers (fhesi_mult_relin_dev)
  const fhesi_ksw *Dev() const;

 private:
  const FHEcontext &context;
  // Init generates the matrix on the device (fhesi_keygen_batch) from this class's draws; b comes back as
  // words, A' = -A is known from the draws.  The vector<DoubleCRT> image the reference exposes is built from
  // them only on demand (GetRepresentation, Export, operator<<) -- ApplyKeySwitch never needs it.
  mutable vector<vector<DoubleCRT>> keySwitchMatrix;
  mutable std::vector<uint32_t> hostB, drawA;  // [entries][n][W] each, pending materialisation when non-empty
  size_t entries = 0;                           // source parts * ndigits
  mutable shared_ptr<fhesi_ksw> devKsw;
  void Materialize() const;
  bool InitOnDevice(const vector<ZZX> &sCoeff, const ZZX &t);
};

// ------------------------------------------------------------------------------- Serialization.h
void Export(ofstream &out, const ZZ &val);
void Export(ofstream &out, const ZZX &val);
void Export(ofstream &out, const DoubleCRT &val);
void Export(ofstream &out, const vec_long &vec);
void Import(ifstream &in, ZZ &val);
void Import(ifstream &in, ZZX &val);
void Import(ifstream &in, DoubleCRT &val);
void Import(ifstream &in, vec_long &vec);
void Export(ofstream &out, const CiphertextPart &part);
void Export(ofstream &out, const Ciphertext &ctxt);
void Import(ifstream &in, CiphertextPart &part);
void Import(ifstream &in, Ciphertext &ctxt);
template <typename T>
void Export(ofstream &out, const T &val) { out.write((char *)&val, sizeof(T)); }
template <typename T>
void Import(ifstream &in, T &val) { in.read((char *)&val, sizeof(T)); }
template <typename T>
void Export(ofstream &out, const vector<T> &vec) {
  uint32_t size = vec.size();
  Export(out, size);
  for (unsigned i = 0; i < vec.size(); i++) Export(out, vec[i]);
}
template <typename T>
void Import(ifstream &in, vector<T> &vec) {
  uint32_t size;
  Import(in, size);
  vec.resize(size);
  for (unsigned i = 0; i < size; i++) Import(in, vec[i]);
}
// Matrix<T> (a client type, Matrix.h) serialises as u32 rows, u32 cols, row-major elements
template <template <class> class M, typename T>
auto Export(ofstream &out, const M<T> &mat) -> decltype(mat.NumRows(), void()) {
  Export(out, mat.NumRows());
  Export(out, mat.NumCols());
  for (unsigned i = 0; i < mat.NumRows(); i++)
    for (unsigned j = 0; j < mat.NumCols(); j++) Export(out, mat(i, j));
}
template <template <class> class M, typename T>
auto Import(ifstream &in, M<T> &mat) -> decltype(mat.NumRows(), void()) {
  uint32_t nRows, nCols;
  Import(in, nRows);
  Import(in, nCols);
  mat.Resize(nRows, nCols);
  for (unsigned i = 0; i < nRows; i++)
    for (unsigned j = 0; j < nCols; j++) Import(in, mat(i, j));
}
