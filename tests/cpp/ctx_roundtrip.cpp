// ctx_roundtrip.cpp -- ExportSIContext -> ImportSIContext keeps the head-room for sums of tensor products
// (FHEContext.cpp:45-81; the file stores the chain, not xi): a server that imports a context sized for xi
// products, sums xi of them in tensor form (Matrix sums in Regression.h) and key-switches must decrypt to the
// plaintext sum.  Also: Export -> Import -> Export is byte-identical.  Run by tests/test_host_cpp.py.
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "Ciphertext.h"
#include "FHE-SI.h"
#include "FHEContext.h"
#include "Plaintext.h"

static std::string slurp(const char *path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  const std::string dir = argv[1];
  const long xi = 6;
  const std::string f1 = dir + "/ctx1.bin", f2 = dir + "/ctx2.bin";
  size_t nprimes = 0;
  {
    FHEcontext context(22, 80, to_ZZ(23L), 7, 3);
    activeContext = &context;
    context.SetUpSIContext(xi);
    nprimes = context.numPrimes();
    std::ofstream out(f1, std::ios::binary);
    context.ExportSIContext(out);
  }
  std::ifstream in(f1, std::ios::binary);
  FHEcontext context(in);
  activeContext = &context;
  if (context.numPrimes() != nprimes) { std::cout << "chain length FAIL\n"; return 1; }
  {
    std::ofstream out(f2, std::ios::binary);
    context.ExportSIContext(out);
  }
  if (slurp(f1.c_str()) != slurp(f2.c_str())) { std::cout << "re-export FAIL\n"; return 1; }
  SetSeed(to_ZZ(11L));
  FHESISecKey secretKey(context);
  const FHESIPubKey &publicKey(secretKey);
  KeySwitchSI keySwitch(secretKey);
  Plaintext want(context);
  Ciphertext sum(publicKey);
  for (long i = 0; i < xi; ++i) {
    Plaintext a = Plaintext::Random(context), b = Plaintext::Random(context);
    Ciphertext ca(publicKey), cb(publicKey);
    publicKey.Encrypt(ca, a);
    publicKey.Encrypt(cb, b);
    ca *= cb;  // tensor form
    if (i == 0) sum = ca;
    else sum += ca;
    a *= b;
    if (i == 0) want = a;
    else want += a;
  }
  keySwitch.ApplyKeySwitch(sum);
  Plaintext got(context);
  secretKey.Decrypt(got, sum);
  if (!(got == want)) { std::cout << "sum of xi products FAIL\n"; return 1; }
  std::cout << "ctx roundtrip ok\n";
  return 0;
}
