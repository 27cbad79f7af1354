// kernels_generic.cuh -- the general (any N = 2^k, any part count) CUDA path.
//
// These kernels are the building blocks behind every C-ABI entry point; the fused N=1024
// kernels in kernels_fused.cuh replace the hot sequence (tensor -> ScaleDown -> key switch)
// but produce bit-identical results and share the data layouts defined here.
//
// Reference map:
//   k_fwd        a2/a4  Cmodulus::FFT / DoubleCRT(const ZZX&)      CModulus.cpp:90-107, DoubleCRT.cpp:244-257
//   k_inv        a3     Cmodulus::iFFT (incl. rem by Phi_m)        CModulus.cpp:110-132
//   k_crt        a7/a9/a12/a15  toPoly + intVecCRT + ScaleDown / Reduce / Decrypt rounding
//                                DoubleCRT.cpp:349-398, NumbTh.cpp:307-335, Ciphertext.cpp:194-218,
//                                Util.cpp:3-26, FHE-SI.cpp:111-118
//   k_tensor_pw  a5/a8  DoubleCRT::Op(MulMod/AddMod) in the tensor  Ciphertext.cpp:179-186
//   k_dot        a11    DotProduct(keySwitchMatrix[i], byteDecomp)  Util.h:80-98, FHE-SI.cpp:251-257
//   k_fwd(DIGIT) a10    ByteDecomp fused into the transform's load  Ciphertext.cpp:82-121
#pragma once
#include "modarith.cuh"

struct PrimeConst {
  u32 p, pinv;      // prime, -p^-1 mod 2^32
  u32 r1, r2;       // R mod p, R^2 mod p
  u32 ninv_r;       // N^-1 * R      : mont(x, .) = x / N            (plain)
  u32 ninv_r2;      // N^-1 * R^2    : mont(x, .) = x / N * R        (key form)
  u32 tensor_c;     // p_pt * N^-1 * R^2 : mont(x, .) = x * p_pt / N * R
  u32 ptxt_r;       // p_pt * R
  u32 scale_r;      // floor(q / p_pt) * R
  // operands of mulw_dfma that must reach the kernel as loaded values (neither nvcc nor ptxas can fold
  // what comes from memory): -p mod 2^32, the high word of the double 2^52, and zero
  u32 negp, hic, zop;
  u32 pad[4];
};

struct DevCtx {
  u32 n, N, logN, W, logQ, D, dbits, h;  // h = m/2 (X^h = -1 mod Phi_m), n = h-1
  u32 Lmax, CW;                           // CW = row length of cword
  u32 ptxt;                               // plaintext modulus p
  const PrimeConst *pc;                   // [Lmax]
  const u32 *tw_fwd, *tw_inv;             // [Lmax][N]  Montgomery form
  const u32 *cword;                       // [Lmax][CW] 2^(32k) * R mod p
  const u32 *garner;                      // [Lmax][Lmax] garner[j][i] = p_i^-1 * R mod p_j (i<j)
  const u32 *Pfull, *Phalf;               // [Lmax+1][Lmax]  words of prod_{i<l} p_i and its half
  // Shoup tables for the fused kernels: (w, floor(w * 2^32 / p)), plain (non-Montgomery) w
  const uint2 *tws_fwd, *tws_inv;         // [Lmax][N] same index h + j as tw_fwd / tw_inv
  // FP64-quotient companions of tws_fwd (mulw_dfma): (c, K) per twiddle, [Lmax][N]
  const double2 *twd_fwd;
  // tws_fwd / tws_inv re-laid-out per thread for the fused N = 1024 kernels (kernels_fused.cuh, FTW layout):
  // [Lmax][FTW_ENTRIES]; NULL when N != 1024
  const uint2 *ftw_fwd, *ftw_inv;
  // multiword -> residue constants: cwr[l][v][k] = 2^(32k) * R * s_v mod p for k < W, and
  // cwr[l][v][W] = p - (2^(32W) * s_v mod p); s_0 = 1 (plain result after Montgomery
  // reduction), s_1 = p_pt / N * R (Montgomery form of the tensor's left operand)
  const u32 *cwr;                         // [Lmax][2][CW]
  u32 sshift;                             // store_index() block: 1 (16 positions) or 2 (32, N = 2048)
  // General m (h == 0): the remainder by Phi_m (CModulus.cpp:128-129, NTL rem) as a sparse gather.  Row i lists the
  // non-zero coefficients of X^i in X^(n+j) mod Phi_m, j < n - 1: [n + 1 row starts][(j, coefficient) pairs].
  // NULL for m = 2h (h an odd prime), whose fold is written out in phim_reduce_store / phim_store_*.
  const u32 *red;
  const u32 *red_wide;                    // the same with j < N - n: every position a transform-domain vector can hold
};

// one row of the general-m remainder: x[0..2n-1) holds a product (values < 2p), returns coefficient i in [0,p)
__device__ __forceinline__ u32 phim_row_csr(const u32 *x, const u32 *__restrict__ red, u32 n, u32 i, u32 p) {
  const u32 p2 = 2 * p;
  u32 s = x[i];
  const u32 *ent = red + n + 1;
  for (u32 t = red[i], e = red[i + 1]; t < e; ++t) {
    const u32 v = x[n + ent[2 * t]];
    int cf = (int)ent[2 * t + 1];  // small: +-1 for every m with at most two odd prime factors
    if (cf > 0) {
      do s = csub(s + v, p2); while (--cf);
    } else {
      do s = csub(s + p2 - v, p2); while (++cf);
    }
  }
  return full_reduce(s, p);
}

// Storage order of transform-domain vectors.  Position i of the in-place DIF output lives at
// store_index(i): inside every block of 16 positions the 8 even ones come first, then the 8
// odd ones.  This is the natural register order of the fused N=1024 kernels
// (kernels_fused.cuh: thread t, register r <-> storage index 8t + r), which lets them read
// key tiles with 128-bit loads; pointwise kernels do not care.  Needs N >= 16.
// With sshift = 2 (N = 2048: 256 threads per transform, kernels_fused2k.cuh) the block is 32 positions: the low
// five bits (j:3, b:2) of a position are stored as (b:2, j:3).
__device__ __forceinline__ u32 store_index(u32 i, u32 sshift = 1) {
  const u32 blk = (8u << sshift) - 1u;
  return (i & ~blk) | ((i & ((1u << sshift) - 1u)) << 3) | ((i >> sshift) & 7u);
}

// ---------------------------------------------------------------------------------------
// shared-memory radix-2 NTT, one butterfly per thread per stage
// ---------------------------------------------------------------------------------------
// forward: decimation in frequency, natural order in, bit-reversed ("position") order out.
__device__ __forceinline__ void ntt_fwd_smem(u32 *x, const u32 *__restrict__ tw, u32 N, u32 p,
                                             u32 pinv) {
  const u32 p2 = 2 * p;
  for (u32 h = N >> 1; h >= 1; h >>= 1) {
    for (u32 b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
      u32 j = b & (h - 1);
      u32 i = ((b - j) << 1) | j;
      u32 X = x[i], Y = x[i + h];
      u32 w = __ldg(tw + h + j);
      x[i] = csub(X + Y, p2);
      x[i + h] = mont_mul(X + p2 - Y, w, p, pinv);
    }
    __syncthreads();
  }
}
// inverse: decimation in time with the inverse root; position order in, natural order out,
// unscaled (the factor 1/N is folded into the constants that produced the input).
__device__ __forceinline__ void ntt_inv_smem(u32 *x, const u32 *__restrict__ itw, u32 N, u32 p,
                                             u32 pinv) {
  const u32 p2 = 2 * p;
  for (u32 h = 1; h < N; h <<= 1) {
    for (u32 b = threadIdx.x; b < (N >> 1); b += blockDim.x) {
      u32 j = b & (h - 1);
      u32 i = ((b - j) << 1) | j;
      u32 X = x[i];
      u32 T = mont_mul(x[i + h], __ldg(itw + h + j), p, pinv);
      x[i] = csub(X + T, p2);
      x[i + h] = csub(X + p2 - T, p2);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// residue of one coefficient
// ---------------------------------------------------------------------------------------
// two's-complement multiword (Win words) -> [0, 2p)
__device__ __forceinline__ u32 residue_from_words(const u32 *__restrict__ w, u32 Win,
                                                  const u32 *__restrict__ cw, u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
  u32 r = 0, top = 0;
  for (u32 k = 0; k < Win; ++k) {
    top = w[k];
    r = csub(r + mont_mul(top, __ldg(cw + k), p, pinv), p2);
  }
  if (top >> 31) r = csub(r + p2 - mont_mul(1u, __ldg(cw + Win), p, pinv), p2);
  return r;
}
// digit d (dbits wide, little-endian) of the non-negative residue mod 2^logQ
// (Ciphertext::ByteDecompPart, Ciphertext.cpp:92-103)
__device__ __forceinline__ u32 digit_from_words(const u32 *__restrict__ w, u32 W, u32 logQ,
                                                u32 dbits, u32 d) {
  u32 o = dbits * d;
  u32 wi = o >> 5, sh = o & 31;
  u32 lo = w[wi];
  u32 hi = (wi + 1 < W) ? w[wi + 1] : 0u;
  u32 v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
  u32 valid = min(dbits, logQ - o);
  return v & ((1u << valid) - 1u);
}

// SRC_RES: per-prime residues [npolys][L][n] (e.g. the output of k_inv)
enum { SRC_POLY = 0, SRC_DIGIT = 1, SRC_U8 = 2, SRC_I32 = 3, SRC_U32 = 4, SRC_RES = 5 };
// SC_MONT: to Montgomery form (x R), no 1/N
enum { SC_NONE = 0, SC_TENSOR = 1, SC_KEYFORM = 2, SC_NINV = 3, SC_MONT = 4 };

struct FwdArgs {
  const void *src;
  u32 Win;       // SRC_POLY: words per coefficient
  u32 L;         // number of primes (prefix of the chain)
  u32 src_mode, scale_mode;
  u32 *dst;      // [npolys][L][N]
};

// grid (npolys, L), block N/2.  dynamic smem: N words.
__global__ void k_fwd(DevCtx c, FwdArgs a) {
  FHESI_SMEM(sm);
  u32 *x = sm;
  const u32 q = blockIdx.x, l = blockIdx.y;
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv;
  u32 sc = 0;
  if (a.scale_mode == SC_TENSOR) sc = pc.tensor_c;
  else if (a.scale_mode == SC_KEYFORM) sc = pc.ninv_r2;
  else if (a.scale_mode == SC_NINV) sc = pc.ninv_r;
  else if (a.scale_mode == SC_MONT) sc = pc.r2;
  for (u32 i = threadIdx.x; i < c.N; i += blockDim.x) {
    u32 r = 0;
    if (i < c.n) {
      switch (a.src_mode) {
        case SRC_POLY:
          r = residue_from_words((const u32 *)a.src + ((size_t)q * c.n + i) * a.Win, a.Win,
                                 c.cword + (size_t)l * c.CW, p, pinv);
          break;
        case SRC_DIGIT: {
          u32 sp = q / c.D, d = q - sp * c.D;
          r = digit_from_words((const u32 *)a.src + ((size_t)sp * c.n + i) * c.W, c.W, c.logQ,
                               c.dbits, d);
          break;
        }
        case SRC_U8: r = ((const uint8_t *)a.src)[(size_t)q * c.n + i]; break;
        case SRC_I32: {
          int v = ((const int *)a.src)[(size_t)q * c.n + i];
          r = v < 0 ? p - ((u32)(-v)) % p : ((u32)v) % p;
          break;
        }
        case SRC_RES: r = ((const u32 *)a.src)[((size_t)q * a.L + l) * c.n + i]; break;
        default: r = ((const u32 *)a.src)[(size_t)q * c.n + i] % p; break;
      }
      if (a.scale_mode != SC_NONE) r = mont_mul(r, sc, p, pinv);
    }
    x[i] = r;
  }
  __syncthreads();
  ntt_fwd_smem(x, c.tw_fwd + (size_t)l * c.N, c.N, p, pinv);
  u32 *dst = a.dst + ((size_t)q * a.L + l) * c.N;
  for (u32 i = threadIdx.x; i < c.N; i += blockDim.x) dst[store_index(i, c.sshift)] = full_reduce(x[i], p);
}

// Phi_m reduction for m = 2h, h odd prime: X^h = -1 and Phi_m = sum_{i<h} (-1)^i X^i.
// x: N unreduced-product coefficients in [0,2p) (zero beyond 2n-2); writes n residues in [0,p).
// y: scratch of h words.
template <class F>
__device__ __forceinline__ void phim_reduce_store(const u32 *x, u32 *y, u32 N, u32 h, u32 p,
                                                  F store) {
  const u32 p2 = 2 * p, n = h - 1;
  for (u32 r = threadIdx.x; r < h; r += blockDim.x) {
    u32 v = x[r];
    if (r + h < N) v = csub(v + p2 - x[r + h], p2);
    y[r] = v;
  }
  __syncthreads();
  const u32 top = y[n];
  for (u32 i = threadIdx.x; i < n; i += blockDim.x) {
    u32 v = (i & 1) ? y[i] + top : y[i] + p2 - top;
    store(i, full_reduce(v, p));
  }
}

struct InvArgs {
  const u32 *src;  // [npolys][L][N]
  u32 L;
  u32 *dst;        // [npolys][L][n]
  // Encrypt addend (FHE-SI.cpp:24-31): + p_pt*e[q][i] + (q even ? floor(q/p_pt)*msg[q/2][i] : 0)
  const int *e;
  const u32 *msg;
  // key-generation addend (FHE-SI.cpp:187-196): + add1[q][i] + sh_src[q / sh_D][i] * 2^(dbits (q % sh_D));
  // sh_pow[j][l] = 2^(dbits j) mod p_l in Montgomery form
  const int *add1 = nullptr;
  const int *sh_src = nullptr;
  const u32 *sh_pow = nullptr;
  u32 sh_D = 1;
};

// grid (npolys, L), block N/2.  dynamic smem: N + h words.
__global__ void k_inv(DevCtx c, InvArgs a) {
  FHESI_SMEM(sm);
  u32 *x = sm, *y = sm + c.N;
  const u32 q = blockIdx.x, l = blockIdx.y;
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv;
  const u32 *src = a.src + ((size_t)q * a.L + l) * c.N;
  for (u32 i = threadIdx.x; i < c.N; i += blockDim.x) x[i] = src[store_index(i, c.sshift)];
  __syncthreads();
  ntt_inv_smem(x, c.tw_inv + (size_t)l * c.N, c.N, p, pinv);
  u32 *dst = a.dst + ((size_t)q * a.L + l) * c.n;
  const int *e = a.e ? a.e + (size_t)q * c.n : nullptr;
  const u32 *msg = (a.msg && !(q & 1)) ? a.msg + (size_t)(q >> 1) * c.n : nullptr;
  auto store = [&](u32 i, u32 v) {
    if (e) {
      int ev = e[i];
      u32 er = ev < 0 ? p - ((u32)(-ev)) % p : ((u32)ev) % p;
      v = csub(v + csub(mont_mul(er, pc.ptxt_r, p, pinv), p), p);
      if (msg) v = csub(v + csub(mont_mul(msg[i] % p, pc.scale_r, p, pinv), p), p);
    }
    if (a.add1) {
      const int ev = a.add1[(size_t)q * c.n + i];
      v = csub(v + (ev < 0 ? p - ((u32)(-ev)) % p : ((u32)ev) % p), p);
    }
    if (a.sh_src) {
      const int sv = a.sh_src[(size_t)(q / a.sh_D) * c.n + i];
      if (sv) {
        const u32 sr = sv < 0 ? p - ((u32)(-sv)) % p : ((u32)sv) % p;
        v = csub(v + csub(mont_mul(sr, a.sh_pow[(size_t)(q % a.sh_D) * a.L + l], p, pinv), p), p);
      }
    }
    dst[i] = v;
  };
  if (c.h) {
    phim_reduce_store(x, y, c.N, c.h, p, store);
  } else {  // general m
    for (u32 i = threadIdx.x; i < c.n; i += blockDim.x) store(i, phim_row_csr(x, c.red_wide, c.n, i, p));
  }
}

// ---------------------------------------------------------------------------------------
// pointwise kernels in the transform domain
// ---------------------------------------------------------------------------------------
struct TensArgs {
  const u32 *A;  // [count][pa][L][N]  scaled by tensor_c (Montgomery form of p_pt*a/N)
  const u32 *B;  // [count][pb][L][N]  plain
  u32 pa, pb, L;
  u32 *out;      // [count or 1][pa+pb-1][L][N]
  u32 count;
  int accumulate;
};
// one thread per (b, l, e); with accumulate, one thread per (l, e) looping over b.
__global__ void k_tensor_pw(DevCtx c, TensArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)a.L * c.N;
  size_t total = a.accumulate ? per : per * a.count;
  if (idx >= total) return;
  u32 b0 = a.accumulate ? 0 : (u32)(idx / per);
  u32 b1 = a.accumulate ? a.count : b0 + 1;
  size_t le = idx % per;
  u32 l = (u32)(le / c.N);
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  const u32 po = a.pa + a.pb - 1;
  u32 acc[5] = {0, 0, 0, 0, 0};
  for (u32 b = b0; b < b1; ++b) {
    u32 av[3], bv[3];
    for (u32 i = 0; i < a.pa; ++i) av[i] = a.A[((size_t)b * a.pa + i) * per + le];
    for (u32 j = 0; j < a.pb; ++j) bv[j] = a.B[((size_t)b * a.pb + j) * per + le];
    for (u32 i = 0; i < a.pa; ++i)
      for (u32 j = 0; j < a.pb; ++j)
        acc[i + j] = csub(acc[i + j] + mont_mul(av[i], bv[j], p, pinv), p2);
  }
  for (u32 k = 0; k < po; ++k) a.out[((size_t)b0 * po + k) * per + le] = csub(acc[k], p);
}

struct DotArgs {
  const u32 *in;   // [count][K][L][N] plain
  const u32 *key;  // [L][K][J][N]     key form (value * R / N)
  u32 K, J, L;
  u32 *out;        // [count][J][L][N]
  u32 count;
};
__global__ void k_dot(DevCtx c, DotArgs a) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t per = (size_t)a.L * c.N;
  if (idx >= per * a.count) return;
  u32 b = (u32)(idx / per);
  size_t le = idx % per;
  u32 l = (u32)(le / c.N), e = (u32)(le % c.N);
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u64 acc[2] = {0, 0};
  u32 ts[2] = {0, 0};
  for (u32 k = 0; k < a.K; ++k) {
    u32 v = a.in[((size_t)b * a.K + k) * per + le];
    const u32 *kp = a.key + (((size_t)l * a.K + k) * a.J) * c.N + e;
    for (u32 j = 0; j < a.J; ++j) acc[j] += (u64)v * kp[(size_t)j * c.N];
    if ((k & 3) == 3 || k + 1 == a.K) {
      for (u32 j = 0; j < a.J; ++j) {
        ts[j] = csub(ts[j] + mont_red64(acc[j], p, pinv), p2);
        acc[j] = 0;
      }
    }
  }
  for (u32 j = 0; j < a.J; ++j) a.out[((size_t)b * a.J + j) * per + le] = csub(ts[j], p);
}

// io += other (mod p_l) / io *= scalar, over [count][parts][L][N]
__global__ void k_tprod_add(DevCtx c, u32 *io, const u32 *other, u32 L, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  u32 l = (u32)((idx / c.N) % L);
  u32 p = c.pc[l].p;
  io[idx] = csub(io[idx] + other[idx], p);
}
// scal[l] = (scalar mod p_l) * R mod p_l
__global__ void k_tprod_mul_scalar(DevCtx c, u32 *io, const u32 *scal, u32 L, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  u32 l = (u32)((idx / c.N) % L);
  const PrimeConst pc = c.pc[l];
  io[idx] = csub(mont_mul(io[idx], scal[l], pc.p, pc.pinv), pc.p);
}
// Key upload staging: interleave b and A to [K][2][n][W] (one transform launch then covers both) and,
// when split is non-null, write the halves of (K mod q) >= 0 = lo + 2^(32 ws) hi as [K][4][n][W]
// (b_lo, b_hi, A_lo, A_hi), each a non-negative W-word polynomial.  tb = logQ mod 32.
__global__ void k_key_stage(const u32 *b, const u32 *A, u32 *inter, u32 *split, u32 K, u32 n, u32 W, u32 ws,
                            u32 tb) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)K * 2 * n) return;
  const u32 i = (u32)(idx % n), r = (u32)((idx / n) & 1), k = (u32)(idx / ((size_t)2 * n));
  const u32 *src = (r ? A : b) + ((size_t)k * n + i) * W;
  u32 *dst = inter + (((size_t)k * 2 + r) * n + i) * W;
  u32 *lo = split ? split + (((size_t)k * 4 + 2 * r) * n + i) * W : nullptr;
  u32 *hi = split ? lo + (size_t)n * W : nullptr;
  for (u32 w = 0; w < W; ++w) {
    u32 v = src[w];
    dst[w] = v;
    if (split) {
      if (w == W - 1 && tb) v &= (1u << tb) - 1;  // two's complement -> residue in [0, q)
      lo[w] = w < ws ? v : 0u;
      if (w >= ws) hi[w - ws] = v;
    }
  }
  if (split)
    for (u32 w = W - ws; w < W; ++w) hi[w] = 0u;
}
// io[b][part][l][e] *= img[l][e] (img in Montgomery form): DoubleCRT *= DoubleCRT with one shared
// right-hand side, the tensor-form branch of Ciphertext::operator*=(const ZZX&) (Ciphertext.cpp:252-256)
__global__ void k_tprod_mul_img(DevCtx c, u32 *io, const u32 *img, u32 L, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const size_t per = (size_t)L * c.N;
  u32 l = (u32)((idx / c.N) % L);
  const PrimeConst pc = c.pc[l];
  io[idx] = csub(mont_mul(io[idx], img[idx % per], pc.p, pc.pinv), pc.p);
}
// DoubleCRT::automorph (DoubleCRT.cpp:439-465) on coefficient residues: in, out [npolys][L][n];
// tab[e] = (source index << 1) | negate for the h = n + 1 positions of the permuted polynomial,
// 0xFFFFFFFF where no source lands; the X^n position is folded with Phi_m: w_j = v_j - (-1)^j v_n
__global__ void k_automorph_res(DevCtx c, const u32 *in, const u32 *tab, u32 *out, u32 L, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * L * c.n) return;
  const u32 j = (u32)(idx % c.n);
  const size_t row = idx / c.n;  // poly * L + l
  const u32 p = c.pc[row % L].p;
  const u32 *base = in + row * c.n;
  const u32 tj = tab[j], tt = tab[c.n];
  u32 a = 0, t = 0;
  if (tj != 0xFFFFFFFFu) {
    a = base[tj >> 1];
    if ((tj & 1) && a) a = p - a;
  }
  if (tt != 0xFFFFFFFFu) {
    t = base[tt >> 1];
    if ((tt & 1) && t) t = p - t;
  }
  out[idx] = (j & 1) ? csub(a + t, p) : csub(a + p - t, p);
}
// io[e] += sum_b in[b][e] over a batch of tprods (the data-phase sums of Matrix.cpp:80-97,149-173)
__global__ void k_tprod_batch_sum(DevCtx c, const u32 *in, u32 count, u32 L, size_t per, u32 *io) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  u32 l = (u32)((idx / c.N) % L);
  u32 p = c.pc[l].p;
  u32 s = io[idx];
  for (u32 b = 0; b < count; ++b) s = csub(s + in[(size_t)b * per + idx], p);
  io[idx] = s;
}
// out[part][l][e] = sum_w in[w][part][l][e]  (multi-GPU combine after all-gather)
__global__ void k_tprod_reduce_world(DevCtx c, const u32 *in, u32 world, u32 L, size_t per,
                                     u32 *out) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  u32 l = (u32)((idx / c.N) % L);
  u32 p = c.pc[l].p;
  u32 s = 0;
  for (u32 w = 0; w < world; ++w) s = csub(s + in[(size_t)w * per + idx], p);
  out[idx] = s;
}

// General m: a(X^k) mod Phi_m is a sparse integer matrix (the index map i -> i k mod m followed by the
// remainder by Phi_m); tab = [n + 1 row starts][(source index, coefficient) pairs], one row per output coefficient
__global__ void k_automorph_res_csr(DevCtx c, const u32 *in, const u32 *tab, u32 *out, u32 L, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * L * c.n) return;
  const u32 j = (u32)(idx % c.n);
  const size_t row = idx / c.n;  // poly * L + l
  const u32 p = c.pc[row % L].p;
  const u32 *base = in + row * c.n, *ent = tab + c.n + 1;
  u32 s = 0;
  for (u32 t = tab[j], e = tab[j + 1]; t < e; ++t) {
    const u32 v = base[ent[2 * t]];
    int cf = (int)ent[2 * t + 1];
    if (cf > 0) {
      do s = csub(s + v, p); while (--cf);
    } else {
      do s = csub(s + p - v, p); while (++cf);
    }
  }
  out[idx] = s;
}
// ---------------------------------------------------------------------------------------
// CRT: Garner mixed radix -> multiword -> centre -> mode-specific rounding
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u32 sext_top(u32 w, u32 logQ) {
  const u32 tb = (logQ - 1) & 31;
  return tb == 31 ? w : (u32)((int)(w << (31 - tb)) >> (31 - tb));
}
enum { CRT_REDUCE_Q = 0, CRT_SCALEDOWN = 1, CRT_DECRYPT = 2, CRT_WIDE = 3, CRT_SCALEDOWN_DIGITS = 4 };

struct CrtArgs {
  const u32 *res;  // [npolys][L][n]
  u32 L, mode;
  u32 *out;        // REDUCE_Q/SCALEDOWN: [npolys][n][W]; DECRYPT: [npolys][n]; WIDE: [npolys][n][Wout]
                   // SCALEDOWN_DIGITS: [npolys][D][n] -- ScaleDown fused with ByteDecomp
  u32 Wout;
  size_t total;    // npolys * n
};

// Per-launch constants passed BY VALUE: they sit in the kernel-parameter constant bank, and since
// every loop below is fully unrolled with compile-time indices they become immediate constant
// operands -- no loads at all (the first version fetched them with dependent __ldg's and stalled
// on long_scoreboard for 7 of every 8 issue slots, profiles/r01_summary_v6.md).
template <int ML>
struct CrtTables {
  u32 p[ML], pinv[ML];
  u32 garner[ML][ML];   // [j][i] = p_i^-1 mod p_j for i < j (plain), and its Shoup quotient
  u32 garnerq[ML][ML];  // floor(garner * 2^32 / p_j)
  u32 Pfull[ML], Phalf[ML];  // words of prod_{i<L} p_i and of its half
};
// dynamic smem: ML * blockDim.x words (only used for the runtime-offset shifts)
template <int ML>
__global__ void __launch_bounds__(128) k_crt(DevCtx c, CrtArgs a, const __grid_constant__ CrtTables<ML> T) {
  FHESI_SMEM(sm);
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < a.total;
  const size_t poly = active ? idx / c.n : 0;
  const u32 coef = active ? (u32)(idx % c.n) : 0;
  const int L = (int)a.L;
  u32 v[ML];
#pragma unroll
  for (int j = 0; j < ML; ++j)
    v[j] = (active && j < L) ? a.res[((size_t)poly * L + j) * c.n + coef] : 0u;
    // mixed-radix digits: v_j = (..((r_j - v_0)/p_0 - v_1)/p_1 ..)/p_{j-1} mod p_j
#pragma unroll
  for (int j = 1; j < ML; ++j) {
    if (j < L) {
      const u32 p = T.p[j], p2 = 2 * p;
      u32 t = v[j];
#pragma unroll
      for (int i = 0; i < j; ++i) {  // Shoup product: any 32-bit input -> [0, 2p)
        const u32 d = t + p2 - v[i];
        t = d * T.garner[j][i] - __umulhi(d, T.garnerq[j][i]) * p;
      }
      v[j] = csub(t, p);
    }
  }
  // x = v_0 + p_0 (v_1 + p_1 (v_2 + ...)), Horner from the top, ML words
  u32 acc[ML];
#pragma unroll
  for (int k = 0; k < ML; ++k) acc[k] = 0;
#pragma unroll
  for (int i = ML - 1; i >= 0; --i) {
    if (i < L) {
      const u32 p = T.p[i];
      // before this step acc < prod_{i<j<L} p_j < 2^(30 (ML-1-i)): only ML-i words can change.
      // The ML-i word products are independent (no carry threaded through the multiplier); the
      // carries ripple through a separate add chain on the ALU pipe.
      u32 plo[ML], phi[ML];
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        if (k <= ML - 1 - i) {
          const u64 t = (u64)acc[k] * p;
          plo[k] = (u32)t;
          phi[k] = (u32)(t >> 32);
        }
      }
      u64 carry = v[i];
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        if (k <= ML - 1 - i) {
          carry += (u64)plo[k] + (k > 0 ? phi[k - 1] : 0u);
          acc[k] = (u32)carry;
          carry >>= 32;
        }
      }
    }
  }
  // centre: if x > P/2 then x -= P   (DoubleCRT.cpp:375-376, NumbTh.cpp:316-318)
  {
    bool gt = false, decided = false;
#pragma unroll
    for (int k = ML - 1; k >= 0; --k) {
      u32 ph = T.Phalf[k];
      if (!decided && acc[k] != ph) {
        gt = acc[k] > ph;
        decided = true;
      }
    }
    if (gt) {
      u32 borrow = 0;
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        u32 pf = T.Pfull[k];
        u64 t = (u64)acc[k] - pf - borrow;
        acc[k] = (u32)t;
        borrow = (u32)(t >> 63);
      }
    }
  }
  const u32 W = c.W, logQ = c.logQ;
  if (a.mode == CRT_WIDE) {
    if (active) {
      u32 *o = a.out + idx * a.Wout;
      u32 sign = (acc[ML - 1] >> 31) ? 0xFFFFFFFFu : 0u;
#pragma unroll
      for (int k = 0; k < ML; ++k)
        if (k < (int)a.Wout) o[k] = acc[k];
      for (u32 k = ML; k < a.Wout; ++k) o[k] = sign;
    }
    return;
  }
  if (a.mode == CRT_REDUCE_Q) {
    if (active) {
      u32 *o = a.out + idx * W;
      const u32 tb = (logQ - 1) & 31;  // sign bit position inside the top word
#pragma unroll
      for (int k = 0; k < ML; ++k) {
        if (k < (int)W) {
          u32 w = acc[k];
          if (k == (int)W - 1 && tb != 31) w = (u32)((int)(w << (31 - tb)) >> (31 - tb));
          o[k] = w;
        }
      }
    }
    return;
  }
  // SCALEDOWN: y = (x + q/2) >> logQ, keep logQ bits centred   (Ciphertext.cpp:206-213)
  // DECRYPT:   y = (p_pt*x + q/2) >> logQ, then mod p_pt       (FHE-SI.cpp:111-118)
  if (a.mode == CRT_DECRYPT) {
    u64 carry = 0;
    const u32 pt = c.ptxt;
    // two's complement multiply by a small positive constant: multiply magnitude-agnostic
    // works mod 2^(32 ML) because acc is already sign-extended over ML words.
#pragma unroll
    for (int k = 0; k < ML; ++k) {
      u64 t = (u64)acc[k] * pt + carry;
      acc[k] = (u32)t;
      carry = t >> 32;
    }
  }
  {  // += 2^(logQ-1)
    const u32 hw = (logQ - 1) >> 5, hb = (logQ - 1) & 31;
    u32 carry = 0;
#pragma unroll
    for (int k = 0; k < ML; ++k) {
      u32 add = (k == (int)hw) ? (1u << hb) : 0u;
      u64 t = (u64)acc[k] + add + carry;
      acc[k] = (u32)t;
      carry = (u32)(t >> 32);
    }
  }
  // runtime-offset funnel shift through shared memory (column per thread)
  u32 *col = sm + threadIdx.x;
  const u32 stride = blockDim.x;
#pragma unroll
  for (int k = 0; k < ML; ++k) col[k * stride] = acc[k];
  const u32 sign = (acc[ML - 1] >> 31) ? 0xFFFFFFFFu : 0u;
  const u32 ws = logQ >> 5, bs = logQ & 31;
  auto word_at = [&](u32 k) -> u32 { return k < (u32)ML ? col[k * stride] : sign; };
  if (!active) return;
  if (a.mode == CRT_SCALEDOWN_DIGITS) {
    // digit d of the non-negative residue mod q of y = (x + q/2) >> logQ: bits
    // [logQ + d*dbits, +dbits) of the shifted sum, clipped at 2*logQ  (Ciphertext.cpp:92-103)
    for (u32 d = 0; d < c.D; ++d) {
      const u32 o = logQ + c.dbits * d;
      const u32 wi = o >> 5, sh = o & 31;
      u32 lo = word_at(wi), hi = word_at(wi + 1);
      u32 v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
      const u32 valid = min(c.dbits, logQ - c.dbits * d);
      a.out[((size_t)poly * c.D + d) * c.n + coef] = v & ((1u << valid) - 1u);
    }
    return;
  }
  if (a.mode == CRT_SCALEDOWN) {
    u32 *o = a.out + idx * W;
    const u32 tb = (logQ - 1) & 31;
    for (u32 k = 0; k < W; ++k) {
      u32 lo = word_at(ws + k), hi = word_at(ws + k + 1);
      u32 w = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
      if (k == W - 1 && tb != 31) w = (u32)((int)(w << (31 - tb)) >> (31 - tb));
      o[k] = w;
    }
  } else {  // DECRYPT: |y| < 2^62 by construction (x < 2^(logQ+20), p_pt < 2^31)
    u32 w0, w1;
    {
      u32 a0 = word_at(ws), a1 = word_at(ws + 1), a2 = word_at(ws + 2);
      w0 = bs ? ((a0 >> bs) | (a1 << (32 - bs))) : a0;
      w1 = bs ? ((a1 >> bs) | (a2 << (32 - bs))) : a1;
    }
    i64 y = (i64)(((u64)w1 << 32) | w0);
    i64 r = y % (i64)c.ptxt;
    if (r < 0) r += c.ptxt;
    a.out[idx] = (u32)r;
  }
}

// ---------------------------------------------------------------------------------------
// ScaleDown without the mixed-radix triangle.  Only the bits [logQ - 1, 2 logQ) of the centred integer x are
// needed (the rounding bit and the logQ bits that survive the shift), so x is taken from the explicit CRT sum
//     x = sum_i y_i (P / p_i)  -  (k + c) P,      y_i = r_i (P / p_i)^-1 mod p_i,
//     k = floor(sum_i y_i / p_i),   c = [x mod P > P / 2],
// evaluated on a WINDOW of 28-bit limbs j0 .. j1 only (j0 = two limbs below the rounding bit's limb, j1 = the limb of
// bit 2 logQ - 1).  In radix 2^28 a column sum_i y_i C_ij of up to 40 products (30 x 28 bits) fits 64 bits, so the inner
// loop is ONE multiply-accumulate per term with no carry handling: L Shoup products + (L + 1) NL multiply-accumulates
// instead of L (L - 1) / 2 Shoup products + L (L + 1) / 2 multiply-accumulates with carry chains.  What the window
// cannot see is bounded, and a thread that cannot PROVE its result exact recomputes it with the full mixed-radix
// routine (crt_exact_limbs):
//   * k and c come from the fixed-point sum F = sum_i y_i floor(2^84 / p_i), kept to 56 fraction bits:
//     sum y_i / p_i lies in [T, T + 256) 2^-56 with T = floor(F / 2^28), so they are certain unless the fraction of T
//     is within 256 of 1 (k) or of 1/2 from below (c).  (x of a tensor product is roughly normal around 0 with
//     |x| ~ 2^-17 P: with 22 fraction bits, the first version, one coefficient in fifty was "unsure".)
//   * the limbs below j0 contribute a carry of less than L 2^30 + 1 into limb j0, i.e. at most +161 / -1 into limb
//     j0 + 1; if that limb is not within 256 of wrapping, limbs j0 + 2 and up are exact.  (j0 = 0: nothing is cut.)
// The first event needs |x| < 2^-48 P, the second has probability 2e-6 per coefficient.  FHESI_CRT_FORCE_EXACT=1 sends every
// thread down the exact routine, FHESI_NO_CRT_DIRECT=1 uses k_crt; the parity tests run all three and require
// equal bytes.
// ---------------------------------------------------------------------------------------
#define CRT_LB 28u
#define CRT_LMASK ((1u << CRT_LB) - 1u)
template <int ML, int NL>
struct CrtDirectTables {
  u32 p[ML], yinv[ML], yinvq[ML];  // (P / p_i)^-1 mod p_i and its Shoup quotient
  u32 rfix_hi[ML], rfix_lo[ML];    // floor(2^84 / p_i) = rfix_hi 2^28 + rfix_lo
  u32 C[ML][NL];                   // limbs j0 .. j0 + NL - 1 of P / p_i
  u32 NP[NL];                      // the same limbs of 2^(28 (j0 + nl)) - P: adding (k + c) NP subtracts (k + c) P
  u32 j0, nl, force_exact, pad_;
  unsigned long long *fallbacks;   // threads that took the exact routine (fhesi_crt_fallbacks), device counter
};
// the exact centred x over the context's first L primes, limbs j0 .. j0 + nl - 1 (two's complement): the rare path
static __device__ __noinline__ void crt_exact_limbs(const DevCtx &c, const u32 *res, size_t stride, int L, u32 j0, u32 nl,
                                                    u32 *out) {
  u32 v[FHESI_MAX_PRIMES], acc[FHESI_MAX_PRIMES + 1];
  for (int j = 0; j < L; ++j) v[j] = res[(size_t)j * stride];
  for (int j = 1; j < L; ++j) {  // mixed-radix digits (NumbTh.cpp:307-335 computes the same integer incrementally)
    const u32 p = c.pc[j].p, pinv = c.pc[j].pinv;
    u32 t = v[j];
    for (int i = 0; i < j; ++i) {
      const u32 vi = csub(v[i], p);  // v_i < p_i < 2 p_j
      // garner[j][i] = p_i^-1 R mod p_j: the Montgomery product is (t - v_i) / p_i mod p_j
      t = csub(mont_mul(t + p - vi, c.garner[(size_t)j * c.Lmax + i], p, pinv), p);
    }
    v[j] = t;
  }
  for (int k = 0; k <= L; ++k) acc[k] = 0;
  for (int i = L - 1; i >= 0; --i) {
    u64 carry = v[i];
    const u32 p = c.pc[i].p;
    for (int k = 0; k < L; ++k) {
      carry += (u64)acc[k] * p;
      acc[k] = (u32)carry;
      carry >>= 32;
    }
  }
  const u32 *Ph = c.Phalf + (size_t)L * c.Lmax, *Pf = c.Pfull + (size_t)L * c.Lmax;
  bool gt = false;
  for (int k = L - 1; k >= 0; --k)
    if (acc[k] != Ph[k]) {
      gt = acc[k] > Ph[k];
      break;
    }
  if (gt) {
    u32 borrow = 0;
    for (int k = 0; k < L; ++k) {
      const u64 t = (u64)acc[k] - Pf[k] - borrow;
      acc[k] = (u32)t;
      borrow = (u32)(t >> 63);
    }
  }
  const u32 sign = (acc[L - 1] >> 31) ? 0xFFFFFFFFu : 0u;
  auto word = [&](u32 k) -> u64 { return k < (u32)L ? acc[k] : sign; };
  for (u32 j = 0; j < nl; ++j) {
    const u32 o = CRT_LB * (j0 + j), w = o >> 5, sh = o & 31;
    out[j] = (u32)(((word(w) | (word(w + 1) << 32)) >> sh)) & CRT_LMASK;
  }
}
// dynamic smem: NL * blockDim.x words.  Modes CRT_SCALEDOWN and CRT_SCALEDOWN_DIGITS only.
// FULLW: the window is exactly NL limbs (the bound becomes a compile-time constant: no predicates in the rows)
template <int ML, int NL, bool FULLW>
__global__ void __launch_bounds__(128) k_crt_direct(DevCtx c, CrtArgs a, const __grid_constant__ CrtDirectTables<ML, NL> T) {
  FHESI_SMEM(sm);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < a.total;
  const size_t poly = active ? idx / c.n : 0;
  const u32 coef = active ? (u32)(idx % c.n) : 0;
  const int L = (int)a.L;
  const u32 j0 = T.j0, nl = FULLW ? (u32)NL : T.nl;
  const u32 *res = a.res + (size_t)poly * L * c.n + coef;
  u32 rr[ML];  // every residue is requested before the first is used: one round trip to HBM, not L
#pragma unroll
  for (int i = 0; i < ML; ++i) rr[i] = (active && i < L) ? __ldg(res + (size_t)i * c.n) : 0u;
  u64 col[NL];
#pragma unroll
  for (int j = 0; j < NL; ++j) col[j] = 0;
  u64 Fh = 0, Fl = 0;
#pragma unroll
  for (int i = 0; i < ML; ++i) {
    if (i < L) {
      const u32 r = rr[i];
      const u32 p = T.p[i];
      const u32 y = csub(r * T.yinv[i] - __umulhi(r, T.yinvq[i]) * p, p);
      Fh += (u64)y * T.rfix_hi[i];  // < 2^57 each
      Fl += (u64)y * T.rfix_lo[i];  // < 2^58 each
#pragma unroll
      for (int j = 0; j < NL; ++j)
        if ((u32)j < nl) col[j] += (u64)y * T.C[i][j];  // < 2^58 each, at most 40 of them
    }
  }
  const u64 Tq = Fh + (Fl >> 28);  // floor(F / 2^28): sum y_i / p_i in units of 2^-56, low by less than 256
  const u64 frac = Tq & ((1ull << 56) - 1), D = 256, HALF = 1ull << 55;
  const u32 kc = (u32)(Tq >> 56) + (frac > HALF ? 1u : 0u);
  bool unsure = T.force_exact || frac >= (1ull << 56) - D || (frac <= HALF && frac + D > HALF);
  u32 limb[NL];
  {
    u64 carry = 0;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      limb[j] = 0;
      if ((u32)j < nl) {
        carry += col[j] + (u64)kc * T.NP[j];
        limb[j] = (u32)carry & CRT_LMASK;
        carry >>= CRT_LB;
      }
    }
  }
  if (j0 && (limb[1] < 256u || limb[1] >= (1u << CRT_LB) - 256u)) unsure = true;
  if (unsure && active) {  // (through a buffer of its own: handing out limb[] would move it from registers to local memory)
    u32 ex[NL];
    if (!T.force_exact) atomicAdd(T.fallbacks, 1ull);
    crt_exact_limbs(c, res, c.n, L, j0, nl, ex);
#pragma unroll
    for (int j = 0; j < NL; ++j)
      if ((u32)j < nl) limb[j] = ex[j];
  }
  const u32 W = c.W, logQ = c.logQ, base = CRT_LB * j0;
  {  // += 2^(logQ-1)
    const u32 hb = logQ - 1 - base, hl = hb / CRT_LB, hs = hb - hl * CRT_LB;
    u32 carry = 0;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const u32 t = limb[j] + ((j == (int)hl) ? (1u << hs) : 0u) + carry;
      limb[j] = t & CRT_LMASK;
      carry = t >> CRT_LB;
    }
  }
  // runtime-offset bit fields through shared memory (column per thread); limbs outside the window are never part of
  // a kept bit, so they read as zero
  u32 *colm = sm + threadIdx.x;
  const u32 stride = blockDim.x;
#pragma unroll
  for (int j = 0; j < NL; ++j) colm[j * stride] = limb[j];
  if (!active) return;
  auto limb_at = [&](u32 q) -> u64 { return q < nl ? colm[q * stride] : 0u; };
  auto bits_at = [&](u32 o, u32 nb) -> u32 {  // nb <= 32 bits from absolute bit position o >= base
    const u32 q = (o - base) / CRT_LB, r = (o - base) - q * CRT_LB;
    const u64 v = limb_at(q) | (limb_at(q + 1) << CRT_LB) | (limb_at(q + 2) << (2 * CRT_LB));
    const u32 w = (u32)(v >> r);
    return nb >= 32 ? w : (w & ((1u << nb) - 1u));
  };
  if (a.mode == CRT_SCALEDOWN_DIGITS) {
    // digit d of the non-negative residue mod q of y = (x + q/2) >> logQ  (Ciphertext.cpp:92-103)
    for (u32 d = 0; d < c.D; ++d)
      a.out[((size_t)poly * c.D + d) * c.n + coef] = bits_at(logQ + c.dbits * d, min(c.dbits, logQ - c.dbits * d));
    return;
  }
  u32 *o = a.out + idx * W;  // CRT_SCALEDOWN: logQ bits, the top word sign-extended
  const u32 tb = (logQ - 1) & 31;
  for (u32 k = 0; k < W; ++k) {
    u32 w = bits_at(logQ + 32 * k, (k == W - 1) ? tb + 1 : 32);
    if (k == W - 1 && tb != 31) w = (u32)((int)(w << (31 - tb)) >> (31 - tb));
    o[k] = w;
  }
}

// CRT for the split-key key switch: two non-negative values lo, hi < P/2 (Ls primes each) ->
// Reduce(lo + 2^(32 ws) hi).  res: [npolys][2][L][n];  out: [npolys][n][W].
struct CrtSplitArgs {
  const u32 *res;
  u32 L, ws;
  u32 *out;
  size_t total;  // npolys * n
};
template <int ML>
__global__ void __launch_bounds__(128) k_crt_split(DevCtx c, CrtSplitArgs a, const __grid_constant__ CrtTables<ML> T) {
  FHESI_SMEM(sm);  // 2 * ML * blockDim.x words
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < a.total;
  const size_t poly = active ? idx / c.n : 0;
  const u32 coef = active ? (u32)(idx % c.n) : 0;
  const int L = (int)a.L;
  const u32 stride = blockDim.x;
  u32 vv[2][ML];  // both halves' residues are requested before the first is used (one round trip to HBM)
#pragma unroll
  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
    for (int j = 0; j < ML; ++j)
      vv[hf][j] = (active && j < L) ? __ldg(a.res + (((size_t)poly * 2 + hf) * L + j) * c.n + coef) : 0u;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    u32 v[ML];
#pragma unroll
    for (int j = 0; j < ML; ++j) v[j] = vv[hf][j];
#pragma unroll
    for (int j = 1; j < ML; ++j) {
      if (j < L) {
        const u32 p = T.p[j], p2 = 2 * p;
        u32 t = v[j];
#pragma unroll
        for (int i = 0; i < j; ++i) {
          const u32 d = t + p2 - v[i];
          t = d * T.garner[j][i] - __umulhi(d, T.garnerq[j][i]) * p;
        }
        v[j] = csub(t, p);
      }
    }
    u32 acc[ML];
#pragma unroll
    for (int k = 0; k < ML; ++k) acc[k] = 0;
#pragma unroll
    for (int i = ML - 1; i >= 0; --i) {
      if (i < L) {
        const u32 p = T.p[i];
        u32 plo[ML], phi[ML];
#pragma unroll
        for (int k = 0; k < ML; ++k) {
          if (k <= ML - 1 - i) {
            const u64 t = (u64)acc[k] * p;
            plo[k] = (u32)t;
            phi[k] = (u32)(t >> 32);
          }
        }
        u64 carry = v[i];
#pragma unroll
        for (int k = 0; k < ML; ++k) {
          if (k <= ML - 1 - i) {
            carry += (u64)plo[k] + (k > 0 ? phi[k - 1] : 0u);
            acc[k] = (u32)carry;
            carry >>= 32;
          }
        }
      }
    }
    // the Phi_m fold makes the sums signed: centre at P/2 like every other toPoly
    {
      bool gt = false, decided = false;
#pragma unroll
      for (int k = ML - 1; k >= 0; --k) {
        const u32 ph = T.Phalf[k];
        if (!decided && acc[k] != ph) {
          gt = acc[k] > ph;
          decided = true;
        }
      }
      if (gt) {
        u32 borrow = 0;
#pragma unroll
        for (int k = 0; k < ML; ++k) {
          u64 t = (u64)acc[k] - T.Pfull[k] - borrow;
          acc[k] = (u32)t;
          borrow = (u32)(t >> 63);
        }
      }
    }
    u32 *col = sm + (size_t)hf * ML * stride + threadIdx.x;
#pragma unroll
    for (int k = 0; k < ML; ++k) col[k * stride] = acc[k];
  }
  if (!active) return;
  const u32 *lo = sm + threadIdx.x, *hi = sm + (size_t)ML * stride + threadIdx.x;
  const u32 slo = (lo[(ML - 1) * stride] >> 31) ? 0xFFFFFFFFu : 0u;
  const u32 shi = (hi[(ML - 1) * stride] >> 31) ? 0xFFFFFFFFu : 0u;
  u32 *o = a.out + idx * c.W;
  u32 carry = 0;
  for (u32 k = 0; k < c.W; ++k) {
    const u32 a0 = k < (u32)ML ? lo[k * stride] : slo;
    const u32 a1 = k >= a.ws ? ((k - a.ws < (u32)ML) ? hi[(k - a.ws) * stride] : shi) : 0u;
    const u64 t = (u64)a0 + a1 + carry;
    carry = (u32)(t >> 32);
    o[k] = (k == c.W - 1) ? sext_top((u32)t, c.logQ) : (u32)t;
  }
}

// PlaintextSpace::EmbedInSlots over a batch (PlaintextSpace.cpp:112-134; BatchData, Regression.h:43-66):
// msg[c][j] = sum_k vals[c][k] * basis[k][j] mod p_pt, with basis[k] the CRT idempotent of slot k.
// One thread per output coefficient; a block shares its row of slot values through shared memory.
// Products are < p_pt^2 <= 2^52 (p_pt < 2^26), so a 64-bit sum of up to 2^12 of them cannot overflow.
__global__ void __launch_bounds__(128) k_embed_slots(const u32 *basis, const u32 *vals, u32 *msg, u32 nslots, u32 n,
                                                     u32 p_pt) {
  FHESI_SMEM(sv);
  const size_t row = blockIdx.y;
  for (u32 k = threadIdx.x; k < nslots; k += blockDim.x) sv[k] = vals[row * nslots + k];
  __syncthreads();
  const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  u64 acc = 0;
  for (u32 k = 0; k < nslots; ++k) {
    const u32 v = sv[k];
    if (v) acc += (u64)v * basis[(size_t)k * n + j];
  }
  msg[row * n + j] = (u32)(acc % p_pt);
}

// ---------------------------------------------------------------------------------------
// coefficient-domain multiword kernels (one thread per coefficient)
// ---------------------------------------------------------------------------------------
// io = Reduce(io + other)   (Ciphertext.cpp:128-131)
__global__ void k_ct_add(DevCtx c, u32 *io, const u32 *other, size_t ncoef) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncoef) return;
  u32 *a = io + idx * c.W;
  const u32 *b = other + idx * c.W;
  u32 carry = 0;
  for (u32 k = 0; k < c.W; ++k) {
    u64 t = (u64)a[k] + b[k] + carry;
    carry = (u32)(t >> 32);
    a[k] = (k == c.W - 1) ? sext_top((u32)t, c.logQ) : (u32)t;
  }
}
// out = Reduce(sum_b in[b])  over a batch; per = parts*n coefficients per ciphertext
__global__ void k_ct_sum(DevCtx c, const u32 *in, u32 *out, size_t per, u32 count) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  u32 acc[17];
  for (u32 k = 0; k < c.W; ++k) acc[k] = 0;
  for (u32 b = 0; b < count; ++b) {
    const u32 *s = in + ((size_t)b * per + idx) * c.W;
    u32 carry = 0;
    for (u32 k = 0; k < c.W; ++k) {
      u64 t = (u64)acc[k] + s[k] + carry;
      acc[k] = (u32)t;
      carry = (u32)(t >> 32);
    }
  }
  u32 *o = out + idx * c.W;
  for (u32 k = 0; k < c.W; ++k) o[k] = (k == c.W - 1) ? sext_top(acc[k], c.logQ) : acc[k];
}
// io = Reduce(io * l)   (Ciphertext.cpp:21-27); l given as sign + magnitude
__global__ void k_ct_mul_scalar(DevCtx c, u32 *io, u64 mag, int neg, size_t ncoef) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncoef) return;
  u32 *a = io + idx * c.W;
  u32 x[17], r[17];
  const u32 W = c.W;
  u32 borrow = neg ? 1u : 0u;  // two's complement negate: ~x + 1
  for (u32 k = 0; k < W; ++k) {
    u32 w = neg ? ~a[k] : a[k];
    u64 t = (u64)w + borrow;
    x[k] = (u32)t;
    borrow = (u32)(t >> 32);
    r[k] = 0;
  }
  const u32 m0 = (u32)mag, m1 = (u32)(mag >> 32);
  u64 carry = 0;
  for (u32 k = 0; k < W; ++k) {
    u64 t = (u64)x[k] * m0 + carry;
    r[k] = (u32)t;
    carry = t >> 32;
  }
  carry = 0;
  for (u32 k = 0; k + 1 < W; ++k) {
    u64 t = (u64)x[k] * m1 + r[k + 1] + carry;
    r[k + 1] = (u32)t;
    carry = t >> 32;
  }
  for (u32 k = 0; k < W; ++k) a[k] = (k == W - 1) ? sext_top(r[k], c.logQ) : r[k];
}
// Reduce of a wide ([..][Win]) poly into [..][W]
__global__ void k_reduce_wide(DevCtx c, const u32 *in, u32 Win, u32 *out, size_t ncoef) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= ncoef) return;
  const u32 *a = in + idx * Win;
  u32 *o = out + idx * c.W;
  for (u32 k = 0; k < c.W; ++k) o[k] = (k == c.W - 1) ? sext_top(a[k], c.logQ) : a[k];
}
// a(X) -> a(X^k) mod Phi_m in coefficient form (m = 2h): a signed permutation into h slots
// followed by the Phi_m fold.  tab[e] = (source index << 1 | negate) or 0xFFFFFFFF if slot e
// receives nothing.  out is [..][n][W+1] (not reduced mod q; Ciphertext.cpp:54-59).
__global__ void k_automorph(DevCtx c, const u32 *in, const u32 *tab, u32 *out, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * c.n) return;
  size_t poly = idx / c.n;
  u32 j = (u32)(idx % c.n);
  const u32 W = c.W, Wo = W + 1;
  const u32 *base = in + poly * c.n * W;
  u32 tj = tab[j], tt = tab[c.n];
  // w_j = v_j - (-1)^j v_n
  u32 neg_top = (j & 1) ? 0u : 1u;  // subtract top when j even
  u32 borrow_a = 0, borrow_b = 0;
  u32 *o = out + idx * Wo;
  // term A = +-in[tj>>1], term B = -+(...)in[tt>>1]; add as sign-extended (W+1)-word values
  u32 ca = (tj != 0xFFFFFFFFu && (tj & 1)) ? 1u : 0u;          // negate A
  u32 cb = (tt != 0xFFFFFFFFu && ((tt & 1) ^ neg_top)) ? 1u : 0u;  // negate B
  borrow_a = ca;
  borrow_b = cb;
  u32 c2 = 0;
  for (u32 k = 0; k < Wo; ++k) {
    u32 wa = 0, wb = 0;
    if (tj != 0xFFFFFFFFu) {
      const u32 *s = base + (size_t)(tj >> 1) * W;
      wa = k < W ? s[k] : ((s[W - 1] >> 31) ? 0xFFFFFFFFu : 0u);
      if (ca) wa = ~wa;
      u64 t = (u64)wa + borrow_a;
      wa = (u32)t;
      borrow_a = (u32)(t >> 32);
    }
    if (tt != 0xFFFFFFFFu) {
      const u32 *s = base + (size_t)(tt >> 1) * W;
      wb = k < W ? s[k] : ((s[W - 1] >> 31) ? 0xFFFFFFFFu : 0u);
      if (cb) wb = ~wb;
      u64 t = (u64)wb + borrow_b;
      wb = (u32)t;
      borrow_b = (u32)(t >> 32);
    }
    u64 t = (u64)wa + wb + c2;
    o[k] = (u32)t;
    c2 = (u32)(t >> 32);
  }
}
// General m: the same through the sparse matrix of k_automorph_res_csr, on two's-complement words.
// out is [..][n][W+1]: a row's coefficients sum to far less than 2^32 in magnitude.
__global__ void k_automorph_csr(DevCtx c, const u32 *in, const u32 *tab, u32 *out, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * c.n) return;
  const size_t poly = idx / c.n;
  const u32 j = (u32)(idx % c.n), W = c.W, Wo = W + 1;
  const u32 *base = in + poly * c.n * W, *ent = tab + c.n + 1;
  u32 acc[17];  // W <= 16
#pragma unroll
  for (int k = 0; k < 17; ++k) acc[k] = 0;
  for (u32 t = tab[j], e = tab[j + 1]; t < e; ++t) {
    const u32 *src = base + (size_t)ent[2 * t] * W;
    const int cf = (int)ent[2 * t + 1];
    const u32 neg = cf < 0 ? 0xFFFFFFFFu : 0u, ext = (src[W - 1] >> 31) ? 0xFFFFFFFFu : 0u;
    for (int rep = cf < 0 ? -cf : cf; rep > 0; --rep) {
      u32 carry = neg & 1u;  // -v = ~v + 1
#pragma unroll
      for (int k = 0; k < 17; ++k) {
        if ((u32)k < Wo) {
          const u64 sum = (u64)acc[k] + (((u32)k < W ? src[k] : ext) ^ neg) + carry;
          acc[k] = (u32)sum;
          carry = (u32)(sum >> 32);
        }
      }
    }
  }
  u32 *o = out + idx * Wo;
#pragma unroll
  for (int k = 0; k < 17; ++k)
    if ((u32)k < Wo) o[k] = acc[k];
}

// ---------------------------------------------------------------------------------------
// reference-chain DoubleCRT rows (64-bit primes, direct evaluation; setup/export time only)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mont_mul64(u64 a, u64 b, u64 p, u64 pinv) {
  u64 lo = a * b, hi = __umul64hi(a, b);
  u64 m = lo * pinv;
  u64 mh = __umul64hi(m, p);
  u64 r = hi + mh + (lo != 0);
  return r >= p ? r - p : r;
}
struct RefRowArgs {
  const u32 *poly;   // [n][Win]
  u32 Win, L;
  const u64 *prime;  // [L] p, and per-prime constants below
  const u64 *pinv;   // -p^-1 mod 2^64
  const u64 *r2;     // 2^128 mod p
  const u64 *zeta;   // m-th root (root^2), plain
  const u32 *units;  // [n]
  i64 *rows;         // [L][n]
};
// grid (ceil(n/128), L): thread = one unit j of one prime
__global__ void k_ref_rows(DevCtx c, RefRowArgs a) {
  u32 j = blockIdx.x * blockDim.x + threadIdx.x;
  u32 l = blockIdx.y;
  if (j >= c.n) return;
  const u64 p = a.prime[l], pinv = a.pinv[l], r2 = a.r2[l];
  const u64 one_m = mont_mul64(1, r2, p, pinv);
  const u64 two32_m = mont_mul64(1ull << 32, r2, p, pinv);
  // x = zeta^{u_j} in Montgomery form
  u64 base = mont_mul64(a.zeta[l], r2, p, pinv), x = one_m;
  for (u32 e = a.units[j]; e; e >>= 1) {
    if (e & 1) x = mont_mul64(x, base, p, pinv);
    base = mont_mul64(base, base, p, pinv);
  }
  u64 acc = 0;  // Horner over coefficients, high to low, Montgomery form
  for (int i = (int)c.n - 1; i >= 0; --i) {
    const u32 *w = a.poly + (size_t)i * a.Win;
    // coefficient mod p: Horner over words (top word signed)
    u64 r = 0;
    for (int k = (int)a.Win - 1; k >= 0; --k) {
      r = mont_mul64(r, two32_m, p, pinv);
      u64 wk = mont_mul64((u64)w[k], r2, p, pinv);
      r += wk;
      if (r >= p) r -= p;
    }
    if (w[a.Win - 1] >> 31) {  // subtract 2^(32 Win)
      u64 t = one_m;
      for (u32 k = 0; k < a.Win; ++k) t = mont_mul64(t, two32_m, p, pinv);
      r = r >= t ? r - t : r + p - t;
    }
    acc = mont_mul64(acc, x, p, pinv) + r;
    if (acc >= p) acc -= p;
  }
  a.rows[(size_t)l * c.n + j] = (i64)mont_mul64(acc, 1, p, pinv);
}

// ---------------------------------------------------------------------------------------
// modmul peak microbenchmarks: register-resident, ILP 8
// ---------------------------------------------------------------------------------------
__global__ void k_peak32(u32 *out, u32 p, u32 pinv, int iters) {
  u32 x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = (threadIdx.x * 8 + k + blockIdx.x) % p;
  u32 w = 123456789u % p;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = mont_mul(x[k], w, p, pinv);
  }
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x[k];
  if (s == 0xFFFFFFFFu) out[0] = s;
}
// single-instruction-class chains: which pipe costs what (DESIGN.md "integer pipe model")
//  kind 0: 32x32->lo32 multiply-add   1: 32x32->64 multiply-add   2: hi32 multiply
//  kind 3: Shoup product x*w - hi(x*w')*p   4: conditional subtract (ALU)   5: fp64 FMA
template <int KIND>
__global__ void k_pipe(u32 *out, u32 p, int iters) {
  u32 x[8];
  u64 y[8];
  double z[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = (threadIdx.x * 8 + k + blockIdx.x) | 1u;
    y[k] = x[k];
    z[k] = (double)x[k];
  }
  const u32 w = 0x2468ace1u % p, wq = (u32)(((u64)w << 32) / p);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (KIND == 0) x[k] = x[k] * w + p;
      else if (KIND == 1) y[k] = (u64)(u32)y[k] * w + y[k];
      else if (KIND == 2) x[k] = __umulhi(x[k], wq) + w;
      else if (KIND == 3) x[k] = x[k] * w - __umulhi(x[k], wq) * p;
      else if (KIND == 4) x[k] = csub(x[k] + w, p);
      else z[k] = z[k] * 1.0000001 + 0.5;
    }
  }
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x[k] ^ (u32)y[k] ^ (u32)(y[k] >> 32) ^ (u32)z[k];
  if (s == 0xFFFFFFFFu) out[0] = s;
}
__global__ void k_peak64(u64 *out, u64 p, u64 pinv, int iters) {
  u64 x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = (threadIdx.x * 8 + k + blockIdx.x) % p;
  u64 w = 1234567890123ull % p;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = mont_mul64(x[k], w, p, pinv);
  }
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s ^= x[k];
  if (s == ~0ull) out[0] = s;
}
