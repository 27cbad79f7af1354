// host_bench.cpp -- per-operator latency of the C++ host layer (one synchronous device call per
// Ciphertext operator, as the reference's clients issue them).  Our own code, written against the
// class surface of SURVEY.md §8b; not part of the product.
//
//   host_bench <logQ> <p> <g> [iters]
#include <chrono>
#include <cstdio>
#include <vector>

#include "Ciphertext.h"
#include "FHE-SI.h"
#include "FHEContext.h"
#include "Plaintext.h"

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  unsigned logQ = atoi(argv[1]), p = atoi(argv[2]), g = atoi(argv[3]);
  int iters = argc > 4 ? atoi(argv[4]) : 200;
  double t0 = now();
  FHEcontext context(p - 1, logQ, p, g, 3);
  activeContext = &context;
  context.SetUpSIContext();
  double t1 = now();
  SetSeed(to_ZZ(1L));
  FHESISecKey secretKey(context);
  const FHESIPubKey &publicKey(secretKey);
  KeySwitchSI keySwitch(secretKey);
  KeySwitchSI rotKey(secretKey, g);
  double t2 = now();
  printf("{\"logQ\": %u, \"p\": %u, \"iters\": %d, \"context_s\": %.6f, \"keygen_s\": %.6f", logQ, p, iters, t1 - t0,
         t2 - t1);

  unsigned slots = context.GetPlaintextSpace().GetUsableSlots();
  std::vector<ZZ> vals(slots);
  for (unsigned i = 0; i < slots; ++i) vals[i] = to_ZZ((long)(i * 7 + 3) % p);
  double t = now();
  for (int i = 0; i < iters; ++i) Plaintext pt(context, vals);
  printf(", \"embed_us\": %.2f", (now() - t) / iters * 1e6);
  Plaintext pt0(context, vals), pt1(context, to_ZZ(5L));

  Ciphertext a(publicKey), b(publicKey);
  publicKey.Encrypt(a, pt0);  // warm: key upload
  publicKey.Encrypt(b, pt1);
  t = now();
  for (int i = 0; i < iters; ++i) publicKey.Encrypt(a, pt0);
  printf(", \"encrypt_us\": %.2f", (now() - t) / iters * 1e6);

  Ciphertext c = a;
  c *= b;
  keySwitch.ApplyKeySwitch(c);  // warm: matrix upload
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
  }
  printf(", \"copy_us\": %.2f", (now() - t) / iters * 1e6);
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
    x += b;
  }
  printf(", \"copy_add_us\": %.2f", (now() - t) / iters * 1e6);
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
    x *= b;
  }
  printf(", \"copy_mul_us\": %.2f", (now() - t) / iters * 1e6);
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
    x *= b;
    keySwitch.ApplyKeySwitch(x);
  }
  printf(", \"copy_mul_relin_us\": %.2f", (now() - t) / iters * 1e6);
  {
    Ciphertext acc = a;
    acc *= b;
    t = now();
    for (int i = 0; i < iters; ++i) {
      Ciphertext x = a;
      x *= b;
      acc += x;  // tensor-form accumulation (Matrix sums)
    }
    printf(", \"copy_mul_accumulate_us\": %.2f", (now() - t) / iters * 1e6);
  }
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
    x >>= g;
    rotKey.ApplyKeySwitch(x);
  }
  printf(", \"rotate_keyswitch_us\": %.2f", (now() - t) / iters * 1e6);
  t = now();
  for (int i = 0; i < iters; ++i) {
    Ciphertext x = a;
    x *= 7L;
  }
  printf(", \"copy_mul_scalar_us\": %.2f", (now() - t) / iters * 1e6);
  Plaintext out(context);
  t = now();
  for (int i = 0; i < iters; ++i) secretKey.Decrypt(out, c);
  printf(", \"decrypt_us\": %.2f", (now() - t) / iters * 1e6);
  std::vector<ZZ_pX> msgs;
  t = now();
  for (int i = 0; i < iters; ++i) out.DecodeSlots(msgs);
  printf(", \"decode_us\": %.2f}\n", (now() - t) / iters * 1e6);
  return 0;
}
