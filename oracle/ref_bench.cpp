// ref_bench.cpp -- TEST INFRASTRUCTURE.  Times the reference's own `c = a; c *= b;
// keySwitch.ApplyKeySwitch(c)` (Test_AddMul.cpp:60-66) -- the reference's sources, compiled by
// oracle/build_ref.py against the NTL stand-in -- and checks each result by decryption.  Our own
// driver code, written against the reference's public API.
//
//   ref_bench <logQ> <p> <g> <ops> [seed]   -> one JSON line
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "Ciphertext.h"
#include "FHE-SI.h"
#include "FHEContext.h"
#include "Plaintext.h"

int main(int argc, char **argv) {
  if (argc < 5) return 2;
  unsigned logQ = atoi(argv[1]), p = atoi(argv[2]), g = atoi(argv[3]);
  int ops = atoi(argv[4]);
  long seed = argc > 5 ? atol(argv[5]) : 20240611;
  FHEcontext context(p - 1, logQ, p, g, 3);
  activeContext = &context;
  context.SetUpSIContext();
  SetSeed(to_ZZ(seed));
  FHESISecKey secretKey(context);
  const FHESIPubKey &publicKey(secretKey);
  KeySwitchSI keySwitch(secretKey);
  Plaintext p0 = Plaintext::Random(context), p1 = Plaintext::Random(context);
  Ciphertext a(publicKey), b(publicKey);
  publicKey.Encrypt(a, p0);
  publicKey.Encrypt(b, p1);
  Plaintext want = p0;
  want *= p1;
  {
    Ciphertext c = a;  // warm-up: the cached Bluestein tables exist afterwards, as in a long run
    c *= b;
    keySwitch.ApplyKeySwitch(c);
  }
  bool ok = true;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < ops; ++i) {
    Ciphertext c = a;
    c *= b;
    keySwitch.ApplyKeySwitch(c);
    if (i == ops - 1) {
      Plaintext got(context);
      secretKey.Decrypt(got, c);
      ok = got == want;
    }
  }
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("{\"logQ\": %u, \"p\": %u, \"ops\": %d, \"seconds\": %.6f, \"ops_per_s\": %.6f, \"decrypt_ok\": %s}\n", logQ, p,
         ops, s, ops / s, ok ? "true" : "false");
  return ok ? 0 : 1;
}
