// api_probe.cpp -- uses every symbol of the reference's client-facing C++ surface that SURVEY.md §8b
// lists, once.  Our own code.  It must compile and run both against the reference's headers/objects
// (oracle/build_ref.py builds that as a validity check of the probe itself) and against
// fhe-si_b200/host (tests/test_host_cpp.py), and print the same self-checks.
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "Ciphertext.h"
#include "DoubleCRT.h"
#include "FHE-SI.h"
#include "FHEContext.h"
#include "Plaintext.h"
#include "PlaintextSpace.h"
#include "Serialization.h"

#define REQUIRE(x)                                          \
  do {                                                      \
    if (!(x)) {                                             \
      std::cout << "api_probe FAILED: " #x << std::endl;    \
      return 1;                                             \
    }                                                       \
  } while (0)

int main(int argc, char **argv) {
  std::string dir = argc > 1 ? argv[1] : ".";
  // FHEcontext: both constructors, set-up, public members, export / import
  FHEcontext context(22, 80, 23, 7);
  FHEcontext context2(22, 80, to_ZZ(23L), 7, 3);
  activeContext = &context;
  context.SetUpSIContext();
  context2.SetUpSIContext(4);
  REQUIRE(context.zMstar.M() == 22 && context.zMstar.phiM() == 10 && deg(context.zMstar.PhimX()) == 10);
  REQUIRE(context.logQ == 80 && context.ndigits == 4 && context.decompSize == 3 && context.stdev > 3.0);
  REQUIRE(context.modulusQ == (to_ZZ(1L) << 80) && context.ModulusP() == 23 && context.Generator() == 7);
  REQUIRE(context.numPrimes() == 3 && context.ctxtPrimes.card() == 3);
  ZZ P = context.productOfPrimes(context.ctxtPrimes);
  REQUIRE(P % context.ithPrime(0) == 0 && context.ithModulus(1).getQ() == context.ithPrime(1));
  {
    std::ofstream out(dir + "/probe_context.bin", std::ios::binary);
    context.ExportSIContext(out);
  }
  {
    std::ifstream in(dir + "/probe_context.bin", std::ios::binary);
    FHEcontext back(in);  // the constructor imports (FHEContext.h:101-103)
    REQUIRE(back.numPrimes() == 3 && back.ithPrime(2) == context.ithPrime(2) && back.logQ == 80);
  }
  SetSeed(to_ZZ(77L));
  srand48(1);
  (void)lrand48();
  // keys
  FHESISecKey secretKey(context);
  const FHESIPubKey &publicKey(secretKey);  // implicit conversion bound to a temporary
  KeySwitchSI keySwitch(secretKey), rotKey(secretKey, 7u);
  std::vector<KeySwitchSI> keys;
  keys.push_back(rotKey);
  REQUIRE(secretKey.GetSize() == 2 && &secretKey.GetContext() == &context && &publicKey.GetContext() == &context);
  std::vector<DoubleCRT> skRep = secretKey.GetRepresentation();
  secretKey.UpdateRepresentation(skRep);
  REQUIRE(publicKey.GetRepresentation().size() == 2 && keySwitch.GetRepresentation().size() == 2);
  {
    std::ofstream out(dir + "/probe_sk.bin", std::ios::binary);
    secretKey.Export(out);
  }
  {
    std::ifstream in(dir + "/probe_sk.bin", std::ios::binary);
    FHESISecKey sk2(context);
    sk2.Import(in);
    ZZX a, b;
    sk2.GetRepresentation()[1].toPoly(a);
    secretKey.GetRepresentation()[1].toPoly(b);
    REQUIRE(a == b);
  }
  // DoubleCRT
  DoubleCRT d1(context), d2(to_ZZX(5L)), d3(to_ZZX(3L), context);
  d1 = d2;
  d1 += d3;
  d1 -= to_ZZ(1L);
  d1 *= 4L;
  d1 /= to_ZZ(4L);
  d1 *= d3;
  ZZX dp;
  d1.toPoly(dp);
  REQUIRE(dp == to_ZZX(21L));
  d1.automorph(7);
  d1 >>= 3;
  d1.toPoly(dp);
  REQUIRE(dp == to_ZZX(21L));
  REQUIRE(d1.getMap().getIndexSet().card() == 3 && &d1.getContext() == &context);
  DoubleCRT ds(context);
  ds.sampleHWt(4);
  ds.sampleSmall();
  ds.sampleGaussian();
  ds.randomize();
  // Plaintext: constructors, slots, arithmetic
  const PlaintextSpace &ps = context.GetPlaintextSpace();
  unsigned slots = ps.GetUsableSlots();
  REQUIRE(slots == 8 && ps.GetTotalSlots() == 10);
  std::vector<ZZ> vals(slots);
  for (unsigned i = 0; i < slots; ++i) vals[i] = to_ZZ((long)(i + 2));
  Plaintext p0(context, vals), p1(context, to_ZZ(3L)), pz(context), pm(context, to_ZZ_pX(4L));
  Plaintext pr = Plaintext::Random(context);
  (void)pr;
  std::vector<ZZ_pX> dec;
  p0.DecodeSlots(dec);
  REQUIRE(dec.size() >= slots && dec[0] == to_ZZ_pX(2L) && dec[7] == to_ZZ_pX(9L));
  Plaintext prod = p0;
  prod *= p1;
  Plaintext sum = p0;
  sum += p1;
  sum -= p1;
  REQUIRE(sum == p0);
  Plaintext rot = p0;
  rot >>= 1;
  // Ciphertext: constructors, every operator, accessors, stream output
  Ciphertext c0, c1(context), a(publicKey), b(publicKey);
  publicKey.Encrypt(a, p0);
  publicKey.Encrypt(b, p1);
  c0 = a;
  c0 += b;
  c0 += to_ZZX(p1.message);
  c0 += p1.message;
  c0 *= 2L;
  c0 *= to_ZZX(p1.message);
  c0 *= p1.message;
  Ciphertext cm = a;
  cm *= b;
  REQUIRE(cm.size() == 3);
  cm.ScaleDown();
  keySwitch.ApplyKeySwitch(cm);
  REQUIRE(cm.size() == 2 && deg(cm.GetPart(0).poly) <= 9 && deg(cm[1].poly) <= 9 && cm.parts.size() == 2);
  Plaintext got(context);
  secretKey.Decrypt(got, cm);
  REQUIRE(got == prod);
  // a write through the non-const operator[] is a write to the ciphertext (FHE-SI.cpp:29 does exactly this)
  Ciphertext cw = a;
  cw[0].poly = b[0].poly;
  cw[1].poly = b[1].poly;
  secretKey.Decrypt(got, cw);
  REQUIRE(got == p1);
  cw += a;  // ... and the next operator sees it
  secretKey.Decrypt(got, cw);
  Plaintext p01 = p0;
  p01 += p1;
  REQUIRE(got == p01);
  Ciphertext cr = a;
  cr >>= 7;
  keys[0].ApplyKeySwitch(cr);
  secretKey.Decrypt(got, cr);
  std::vector<ZZ_pX> d2v, d0v;
  got.DecodeSlots(d2v, false);
  p0.DecodeSlots(d0v, false);
  bool rotated = false;  // some cyclic shift of the slot vector
  for (unsigned s = 0; s < d0v.size() && !rotated; ++s) {
    bool ok = true;
    for (unsigned i = 0; i < d0v.size() && ok; ++i) ok = d2v[(i + s) % d0v.size()] == d0v[i];
    rotated = ok;
  }
  REQUIRE(rotated);
  std::ostringstream os;
  os << cm;
  REQUIRE(!os.str().empty());
  c1.Clear();
  REQUIRE(c1.size() == 0);
  // Serialization overload set
  {
    std::ofstream out(dir + "/probe_misc.bin", std::ios::binary);
    Export(out, to_ZZ(-12345L));
    Export(out, to_ZZX(p0.message));
    Export(out, cm);
    std::vector<Ciphertext> vc(2, cm);
    Export(out, vc);
    Export(out, 7u);
  }
  {
    std::ifstream in(dir + "/probe_misc.bin", std::ios::binary);
    ZZ z;
    ZZX zx;
    Ciphertext cb;
    std::vector<Ciphertext> vb;
    unsigned u = 0;
    Import(in, z);
    Import(in, zx);
    Import(in, cb);
    Import(in, vb);
    Import(in, u);
    REQUIRE(z == to_ZZ(-12345L) && zx == to_ZZX(p0.message) && vb.size() == 2 && u == 7);
    secretKey.Decrypt(got, cb);
    REQUIRE(got == prod);
  }
  ZZ_p rz = random_ZZ_p();
  (void)power(rz, 3);
  (void)RandomBnd(10L);
  std::cout << "api_probe ok" << std::endl;
  return 0;
}
