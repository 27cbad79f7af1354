// kernels_fused.cuh -- the hot path for N = 1024 (phi(m) = 508, every p = 1019 config).
//
// One 128-thread group per transform, 8 coefficients per thread held in registers:
//   pass 1  position bits 9,8,7   (thread holds i = j*128 + t)
//   pass 2  position bits 6,5,4   (thread holds i = hi*128 + j*16 + lo,  t = hi*16 + lo)
//   pass 3  position bits 3,2,1   (thread holds i = u*16 + j*2 + b0,     t = u*2 + b0)
//   last    position bit 0        one warp shuffle (lane ^ 1), twiddle 1
// Two padded shared-memory exchanges sit between the passes (PAD(i) = i + 2*(i>>4) makes
// every exchange bank-conflict free); the group synchronises with its own named barrier,
// never the whole CTA.  All 21 twiddles a thread ever needs depend only on its thread index,
// so they are loaded once per CTA and stay in registers across all transforms of the CTA.
// The transform-domain storage order store_index() is this kernel's natural output order
// (thread-major, 8 consecutive words per thread), so key tiles are read with 128-bit loads.
//
// k_fused_tensor     a8 + a3      Ciphertext::operator*= and the per-prime half of ScaleDown
//                                 (Ciphertext.cpp:167-192, CModulus.cpp:110-132)
// k_fused_keyswitch  a10/a11      ByteDecomp digits -> 3D forward transforms -> two inner
//                                 products with keySwitchMatrix -> inverse transform
//                                 (FHE-SI.cpp:244-257, Util.h:80-98)
#pragma once
#include "kernels_generic.cuh"

// [P][L][N] -> [L][P][N]
__global__ void k_transpose_key(const u32 *in, u32 *out, u32 P, u32 L, u32 N) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)P * L * N;
  if (idx >= total) return;
  u32 e = (u32)(idx % N);
  u32 l = (u32)((idx / N) % L);
  u32 q = (u32)(idx / ((size_t)N * L));
  out[((size_t)l * P + q) * N + e] = in[idx];
}

#ifndef FHESI_EMU
__device__ __forceinline__ void fhesi_group_sync(unsigned g) {
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
}
#endif

#define FN 1024u
#define FPADN 1152u  // FN + 2*(FN>>4)
__device__ __forceinline__ u32 fpad(u32 i) { return i + ((i >> 4) << 1); }

struct Tw21 {
  u32 a[7], b[7], c[7];
};
// tw: the prime's [N] table (forward or inverse), index h + j (kernels_generic.cuh)
__device__ __forceinline__ void load_tw21(Tw21 &w, const u32 *__restrict__ tw, u32 tg) {
  const u32 lo = tg & 15, b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    w.a[j] = __ldg(tw + 512 + j * 128 + tg);
    w.b[j] = __ldg(tw + 64 + j * 16 + lo);
    w.c[j] = __ldg(tw + 8 + j * 2 + b0);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    w.a[4 + j] = __ldg(tw + 256 + j * 128 + tg);
    w.b[4 + j] = __ldg(tw + 32 + j * 16 + lo);
    w.c[4 + j] = __ldg(tw + 4 + j * 2 + b0);
  }
  w.a[6] = __ldg(tw + 128 + tg);
  w.b[6] = __ldg(tw + 16 + lo);
  w.c[6] = __ldg(tw + 2 + b0);
}

// Gentleman-Sande butterfly: values in [0,2p) in and out
#define GS(X, Y, W)                              \
  do {                                           \
    u32 s_ = (X) + (Y), d_ = (X) + p2 - (Y);     \
    (X) = csub(s_, p2);                          \
    (Y) = mont_mul(d_, (W), p, pinv);            \
  } while (0)
// Cooley-Tukey butterfly: values in [0,2p) in and out
#define CT(X, Y, W)                              \
  do {                                           \
    u32 t_ = mont_mul((Y), (W), p, pinv);        \
    u32 s_ = (X) + t_, d_ = (X) + p2 - t_;       \
    (X) = csub(s_, p2);                          \
    (Y) = csub(d_, p2);                          \
  } while (0)

// three DIF stages on the 8 registers; w[0..3] first stage, w[4..5] second, w[6] third
__device__ __forceinline__ void dif8(u32 *x, const u32 *w, u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
#pragma unroll
  for (int j = 0; j < 4; ++j) GS(x[j], x[j + 4], w[j]);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    GS(x[j], x[j + 2], w[4 + j]);
    GS(x[j + 4], x[j + 6], w[4 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; j += 2) GS(x[j], x[j + 1], w[6]);
}
// the mirror image: three DIT stages, innermost first
__device__ __forceinline__ void dit8(u32 *x, const u32 *w, u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
#pragma unroll
  for (int j = 0; j < 8; j += 2) CT(x[j], x[j + 1], w[6]);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    CT(x[j], x[j + 2], w[4 + j]);
    CT(x[j + 4], x[j + 6], w[4 + j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) CT(x[j], x[j + 4], w[j]);
}

// Forward transform of a polynomial whose upper half is zero.  In: x[0..3] = coefficients
// j*128 + tg (values < 2p; x[4..7] ignored).  Out: x[r] = transform value at storage index
// tg*8 + r, fully reduced to [0,p).  bufA/bufB: two FPADN-word exchange buffers of the group.
__device__ __forceinline__ void fwd1024(u32 *x, const Tw21 &w, u32 *bufA, u32 *bufB, u32 g, u32 tg,
                                        u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
  // pass 1; its first stage sees (X, 0): X stays, the partner becomes X * w
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j + 4] = mont_mul(x[j], w.a[j], p, pinv);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    GS(x[j], x[j + 2], w.a[4 + j]);
    GS(x[j + 4], x[j + 6], w.a[4 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; j += 2) GS(x[j], x[j + 1], w.a[6]);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[fpad(j * 128 + tg)] = x[j];
  fhesi_group_sync(g);
  const u32 hi = tg >> 4, lo = tg & 15;
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[fpad(hi * 128 + j * 16 + lo)];
  dif8(x, w.b, p, pinv);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[fpad(hi * 128 + j * 16 + lo)] = x[j];
  fhesi_group_sync(g);
  const u32 u = tg >> 1, b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[fpad(u * 16 + j * 2 + b0)];
  dif8(x, w.c, p, pinv);
  // last stage (position bit 0) across lane pairs, twiddle 1
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    u32 v = b0 ? o + p2 - x[j] : x[j] + o;
    x[j] = csub(csub(v, p2), p);
  }
}
// Inverse transform.  In: x[r] = value at storage index tg*8 + r (< 2p).  Out: the natural
// order result (unscaled) is written to nat[0..1024) (nat may alias bufA) and the group is
// synchronised, so every thread may read any coefficient afterwards.
__device__ __forceinline__ void inv1024(u32 *x, const Tw21 &w, u32 *bufA, u32 *bufB, u32 *nat, u32 g,
                                        u32 tg, u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
  const u32 u = tg >> 1, b0 = tg & 1, hi = tg >> 4, lo = tg & 15;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    u32 v = b0 ? o + p2 - x[j] : x[j] + o;
    x[j] = csub(v, p2);
  }
  dit8(x, w.c, p, pinv);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[fpad(u * 16 + j * 2 + b0)] = x[j];
  fhesi_group_sync(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[fpad(hi * 128 + j * 16 + lo)];
  dit8(x, w.b, p, pinv);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[fpad(hi * 128 + j * 16 + lo)] = x[j];
  fhesi_group_sync(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[fpad(j * 128 + tg)];
  dit8(x, w.a, p, pinv);
  // every thread is past its bufA reads (it passed the bufB barrier), so nat may alias bufA
#pragma unroll
  for (int j = 0; j < 8; ++j) nat[j * 128 + tg] = x[j];
  fhesi_group_sync(g);
}
// Phi_m fold for m = 2h (X^h = -1, Phi_m = sum (-1)^i X^i) from the natural-order product in
// nat[] (values < 2p), storing n = h-1 fully reduced residues to dst[0..n).
__device__ __forceinline__ void phim_store_1024(const u32 *nat, u32 *__restrict__ dst, u32 h, u32 tg,
                                                u32 p) {
  const u32 p2 = 2 * p, n = h - 1;
  u32 top = nat[n];
  if (n + h < FN) top = csub(top + p2 - nat[n + h], p2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = j * 128 + tg;
    if (i < n) {
      u32 v = nat[i];
      if (i + h < FN) v = csub(v + p2 - nat[i + h], p2);
      v = (i & 1) ? v + top : v + p2 - top;
      dst[i] = full_reduce(v, p);
    }
  }
}

static bool fused_supported(const DevCtx &dc) { return dc.N == FN && dc.n <= 512; }

// ---------------------------------------------------------------------------------------
// tensor product + inverse transform + Phi_m fold, one CTA per (prime, ciphertext pair)
// ---------------------------------------------------------------------------------------
struct FusedTensorArgs {
  const u32 *a, *b;  // [count][2][n][W]
  u32 *res;          // [count][3][Lt][n]
  u32 Lt;
};
#define FT_SMEM_WORDS (4 * 2 * FPADN + 4 * FN)
__global__ void __launch_bounds__(512, 1) k_fused_tensor(DevCtx c, FusedTensorArgs a) {
  FHESI_SMEM(sm);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  const u32 l = blockIdx.x;
  const size_t op = blockIdx.y;
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u32 *bufA = sm + g * 2 * FPADN, *bufB = bufA + FPADN;
  u32 *F = sm + 4 * 2 * FPADN;  // [4][FN] transform images, storage order
  Tw21 w;
  load_tw21(w, c.tw_fwd + (size_t)l * FN, tg);
  // group g transforms a0, a1 (scaled by p_pt / N, Montgomery form) or b0, b1 (plain)
  const u32 *src = (g < 2 ? a.a + (op * 2 + g) * (size_t)c.n * c.W : a.b + (op * 2 + (g - 2)) * (size_t)c.n * c.W);
  u32 x[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = j * 128 + tg;
    u32 r = 0;
    if (i < c.n) {
      r = residue_from_words(src + (size_t)i * c.W, c.W, c.cword + (size_t)l * c.CW, p, pinv);
      if (g < 2) r = mont_mul(r, pc.tensor_c, p, pinv);
    }
    x[j] = r;
  }
  fwd1024(x, w, bufA, bufB, g, tg, p, pinv);
#pragma unroll
  for (int j = 0; j < 8; ++j) F[g * FN + tg * 8 + j] = x[j];
  __syncthreads();
  if (g == 3) return;
  // tProd[g] = sum_{i+j=g} a_i * b_j   (Ciphertext.cpp:179-186)
  const u32 *A0 = F, *A1 = F + FN, *B0 = F + 2 * FN, *B1 = F + 3 * FN;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const u32 s = tg * 8 + j;
    u32 v;
    if (g == 0) v = mont_mul(A0[s], B0[s], p, pinv);
    else if (g == 2) v = mont_mul(A1[s], B1[s], p, pinv);
    else v = csub(mont_mul(A0[s], B1[s], p, pinv) + mont_mul(A1[s], B0[s], p, pinv), p2);
    x[j] = v;
  }
  load_tw21(w, c.tw_inv + (size_t)l * FN, tg);
  inv1024(x, w, bufA, bufB, bufA, g, tg, p, pinv);
  phim_store_1024(bufA, a.res + ((op * 3 + g) * a.Lt + l) * (size_t)c.n, c.h, tg, p);
}

// Same forward half, but leaving the tensor in transform-domain (tprod) form:
// Ciphertext::operator*= without the ScaleDown (Ciphertext.cpp:167-192).
struct FusedTprodArgs {
  const u32 *a, *b;  // [count][2][n][W]
  u32 *tprod;        // [count][3][Lt][N]
  u32 Lt;
};
__global__ void __launch_bounds__(512, 1) k_fused_tprod(DevCtx c, FusedTprodArgs a) {
  FHESI_SMEM(sm);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  const u32 l = blockIdx.x;
  const size_t op = blockIdx.y;
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u32 *bufA = sm + g * 2 * FPADN, *bufB = bufA + FPADN;
  u32 *F = sm + 4 * 2 * FPADN;
  Tw21 w;
  load_tw21(w, c.tw_fwd + (size_t)l * FN, tg);
  const u32 *src = (g < 2 ? a.a + (op * 2 + g) * (size_t)c.n * c.W : a.b + (op * 2 + (g - 2)) * (size_t)c.n * c.W);
  u32 x[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = j * 128 + tg;
    u32 r = 0;
    if (i < c.n) {
      r = residue_from_words(src + (size_t)i * c.W, c.W, c.cword + (size_t)l * c.CW, p, pinv);
      if (g < 2) r = mont_mul(r, pc.tensor_c, p, pinv);
    }
    x[j] = r;
  }
  fwd1024(x, w, bufA, bufB, g, tg, p, pinv);
#pragma unroll
  for (int j = 0; j < 8; ++j) F[g * FN + tg * 8 + j] = x[j];
  __syncthreads();
  if (g == 3) return;
  const u32 *A0 = F, *A1 = F + FN, *B0 = F + 2 * FN, *B1 = F + 3 * FN;
  u32 *dst = a.tprod + ((op * 3 + g) * a.Lt + l) * (size_t)FN + tg * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const u32 s = tg * 8 + j;
    u32 v;
    if (g == 0) v = mont_mul(A0[s], B0[s], p, pinv);
    else if (g == 2) v = mont_mul(A1[s], B1[s], p, pinv);
    else v = csub(mont_mul(A0[s], B1[s], p, pinv) + mont_mul(A1[s], B0[s], p, pinv), p2);
    dst[j] = csub(v, p);
  }
}

// ---------------------------------------------------------------------------------------
// key switch: one 128-thread group per (prime, ciphertext); KG ciphertexts per CTA share the
// prime's key tiles through L1
// ---------------------------------------------------------------------------------------
struct FusedKsArgs {
  const u32 *digits;  // [count][K][n]   dbits-wide digits, part-major digit-minor
  const u32 *key;     // [Lk][K][2][N]   key form, storage order
  u32 *res;           // [count][2][Lk][n]
  u32 K, Lk, count;
};
#define KG 4
#define FK_SMEM_WORDS (KG * 2 * FPADN)
__global__ void __launch_bounds__(KG * 128, 1) k_fused_keyswitch(DevCtx c, FusedKsArgs a) {
  FHESI_SMEM(sm);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  const u32 l = blockIdx.x;
  const size_t op = (size_t)blockIdx.y * KG + g;
  if (op >= a.count) return;  // whole group leaves together; only group barriers are used
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u32 *bufA = sm + g * 2 * FPADN, *bufB = bufA + FPADN;
  Tw21 w;
  load_tw21(w, c.tw_fwd + (size_t)l * FN, tg);
  u64 acc0[8], acc1[8];
  u32 t0[8], t1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0, t0[j] = t1[j] = 0;
  const u32 *dig = a.digits + op * a.K * (size_t)c.n;
  const u32 *key = a.key + (size_t)l * a.K * 2 * FN + tg * 8;
  for (u32 k = 0; k < a.K; ++k) {
    u32 x[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const u32 i = j * 128 + tg;
      x[j] = i < c.n ? __ldg(dig + (size_t)k * c.n + i) : 0u;
    }
    const uint4 *kp = (const uint4 *)(key + (size_t)k * 2 * FN);
    const uint4 ka0 = __ldg(kp), ka1 = __ldg(kp + 1);
    const uint4 kb0 = __ldg(kp + FN / 4), kb1 = __ldg(kp + FN / 4 + 1);
    fwd1024(x, w, bufA, bufB, g, tg, p, pinv);
    const u32 kb[8] = {ka0.x, ka0.y, ka0.z, ka0.w, ka1.x, ka1.y, ka1.z, ka1.w};
    const u32 kA[8] = {kb0.x, kb0.y, kb0.z, kb0.w, kb1.x, kb1.y, kb1.z, kb1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc0[j] += (u64)x[j] * kb[j];
      acc1[j] += (u64)x[j] * kA[j];
    }
    if ((k & 7) == 7 || k + 1 == a.K) {  // <= 8 products of < p^2 each: the sum stays < 2^63
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t0[j] = csub(t0[j] + csub(mont_red64(acc0[j], p, pinv), p2), p2);
        t1[j] = csub(t1[j] + csub(mont_red64(acc1[j], p, pinv), p2), p2);
        acc0[j] = acc1[j] = 0;
      }
    }
  }
  load_tw21(w, c.tw_inv + (size_t)l * FN, tg);
  inv1024(t0, w, bufA, bufB, bufA, g, tg, p, pinv);
  phim_store_1024(bufA, a.res + ((op * 2 + 0) * a.Lk + l) * (size_t)c.n, c.h, tg, p);
  fhesi_group_sync(g);  // bufA is rewritten by the next inverse transform
  inv1024(t1, w, bufA, bufB, bufA, g, tg, p, pinv);
  phim_store_1024(bufA, a.res + ((op * 2 + 1) * a.Lk + l) * (size_t)c.n, c.h, tg, p);
}

// ByteDecomp (Ciphertext.cpp:82-121) of coefficient-form parts into the digit layout above:
// in [npolys][n][W] -> out [npolys][D][n]
__global__ void k_digits(DevCtx c, const u32 *in, u32 *out, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * c.n) return;
  size_t poly = idx / c.n;
  u32 i = (u32)(idx % c.n);
  const u32 *w = in + idx * c.W;
  for (u32 d = 0; d < c.D; ++d) out[(poly * c.D + d) * c.n + i] = digit_from_words(w, c.W, c.logQ, c.dbits, d);
}

static int fused_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_fused_tensor, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(FT_SMEM_WORDS * 4));
  if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(k_fused_tprod, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(FT_SMEM_WORDS * 4));
  if (e != cudaSuccess) return -1;
  e = cudaFuncSetAttribute(k_fused_keyswitch, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)(FK_SMEM_WORDS * 4));
  return e == cudaSuccess ? 0 : -1;
}
