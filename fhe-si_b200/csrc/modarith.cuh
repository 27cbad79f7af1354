// modarith.cuh -- 32-bit Montgomery arithmetic (R = 2^32) for primes 2^29 < p < 2^30.
//
// Replaces NTL's AddMod/SubMod/MulMod on `long` residues (DoubleCRT.h:206-251,
// DoubleCRT.cpp:79-113).  Word size is a design choice evidenced in DESIGN.md: one 64-bit
// Montgomery product costs ~4x the IMADs of a 32-bit one, while covering the same dynamic
// range needs only 2x as many 30-bit limbs.
//
// Lazy ranges: with p < 2^30, mont_mul(a, b) < 2p whenever a*b < 2^32 * p, in particular
// for a < 4p, b < p and for a, b < 2p.  Values in [0, 4p) always fit a uint32_t.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#ifndef FHESI_EMU  // tests/emu/cuda_emu.h (CPU kernel-logic tests) pre-defines these
#include <cuda_runtime.h>
#define FHESI_LAUNCH(kern, grid, block, smem, stream, ...) \
  kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FHESI_SMEM(name) extern __shared__ uint32_t name[]
#endif

typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define FHESI_HD __host__ __device__ __forceinline__

// (a*b + m*p) / 2^32 with m = a*b*(-p^-1) mod 2^32.  Result < 2p (see header).
FHESI_HD u32 mont_mul(u32 a, u32 b, u32 p, u32 pinv) {
  u64 t = (u64)a * b;
  u32 m = (u32)t * pinv;
  u64 u = t + (u64)m * p;
  return (u32)(u >> 32);
}
// Montgomery reduction of a 64-bit lazy sum t < 2^32 * p * k: result < (k+1) p.
FHESI_HD u32 mont_red64(u64 t, u32 p, u32 pinv) {
  u32 m = (u32)t * pinv;
  // t + m*p may carry out of 64 bits when t is close to 2^64; callers keep t < 2^63.
  u64 u = t + (u64)m * p;
  return (u32)(u >> 32);
}
// a + b forced onto the ALU pipe (VIADDMNMX).  ptxas otherwise turns many plain adds into
// IMAD.IADD, which occupies the integer-multiply pipe -- the pipe that bounds these kernels
// (profiles/r01_summary_v4.md).
FHESI_HD u32 add_alu(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  return __viaddmax_u32(a, b, 0u);
#else
  return a + b;
#endif
}
// a - b (mod 2^32), likewise kept off the integer-multiply pipe (ptxas would emit IMAD.IADD)
FHESI_HD u32 sub_alu(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  return (u32)__viaddmax_s32((int)a, -(int)b, (int)0x80000000);
#else
  return a - b;
#endif
}
// x in [0, 2*c) -> [0, c) by one conditional subtract (unsigned-min trick).
FHESI_HD u32 csub(u32 x, u32 c) {
  u32 y = x - c;
  return y < x ? y : x;
}
FHESI_HD u32 full_reduce(u32 x, u32 p) { return csub(csub(x, 2 * p), p); }  // [0,4p) -> [0,p)

// Shoup-style product by a fixed w with the quotient estimated on the FP64 pipe instead of by IMAD.HI
// (which costs two slots of the integer-multiply pipe that bounds these kernels; the FP64 pipe is idle).
//   c = k 2^-50 with k = floor(w 2^50 / p)  (a double with at most 50 significant bits),  K = 2^52 - 4k.
// The register pair (x, 0x43300000) IS the double 2^52 + x, so fma(2^52 + x, c, K) = x c + 2^52 exactly
// before its single rounding: the low word of the result is q = round(x c), |q - x w / p| < 1/2 + 2^-18.
// Then x w + p - q p lies in (p/2 - p 2^-18, 3p/2 + p 2^-18): inside [0, 2p), the contract of mulw().
// Valid for every 32-bit x; checked exhaustively on the B200 (scripts/micro/shoup_dfma.cu).
// hic, zop: registers holding 0x43300000 and 0 that the compiler cannot see through.  The pair's high
// word is produced per use by one ALU-pipe instruction, hic | (x & zop): ptxas materialises a known or
// loop-invariant value into the pair with IMAD.MOV -- on the very pipe this function is meant to relieve.
FHESI_HD u32 mulw_dfma(u32 x, u32 w, double c, double K, u32 p, u32 negp, u32 hic, u32 zop) {
#if defined(__CUDA_ARCH__)
  const double Q = __fma_rn(__hiloint2double((int)(hic | (x & zop)), (int)x), c, K);  // zop == 0: one LOP3 per use
  const u32 q = (u32)__double2loint(Q);
#else
  const u64 xb = 0x4330000000000000ull | x;
  double xd;
  memcpy(&xd, &xb, 8);
  const double Q = fma(xd, c, K);
  u64 qb;
  memcpy(&qb, &Q, 8);
  const u32 q = (u32)qb;
#endif
  return q * negp + (x * w + p);
}

// ---- host-side helpers (setup time) -------------------------------------------------
static inline u64 h_mulmod(u64 a, u64 b, u64 m) { return (u64)((unsigned __int128)a * b % m); }
static inline u64 h_powmod(u64 a, u64 e, u64 m) {
  u64 r = 1 % m;
  a %= m;
  while (e) {
    if (e & 1) r = h_mulmod(r, a, m);
    a = h_mulmod(a, a, m);
    e >>= 1;
  }
  return r;
}
static inline bool h_is_prime(u64 n) {
  if (n < 2) return false;
  static const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  for (u64 b : bases) {
    if (n % b == 0) return n == b;
  }
  u64 d = n - 1;
  int s = 0;
  while ((d & 1) == 0) {
    d >>= 1;
    ++s;
  }
  for (u64 a : bases) {
    u64 x = h_powmod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < s; ++i) {
      x = h_mulmod(x, x, n);
      if (x == n - 1) {
        comp = false;
        break;
      }
    }
    if (comp) return false;
  }
  return true;
}
static inline u64 h_invmod(u64 a, u64 p) { return h_powmod(a, p - 2, p); }  // p prime
