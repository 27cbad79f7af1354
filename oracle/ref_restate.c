/*
 * ref_restate.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into the product).
 *
 * A plain-C restatement of the reference's *algorithm* for the hot path, as opposed to
 * fhesi_oracle.py which restates its *arithmetic* with exact integers:
 *
 *   - per-prime m-point Bluestein transform through a power-of-two cyclic convolution
 *     (bluestein.cpp:93-144), the convolution done multi-modularly over three FFT primes as
 *     NTL's fftRep does for a modulus that is not itself an FFT prime;
 *   - Cmodulus::FFT / iFFT incl. gather/scatter on Z_m^*, division by m and rem by Phi_m
 *     (CModulus.cpp:90-132);
 *   - DoubleCRT from ZZX, pointwise ops, toPoly with incremental CRT on big integers
 *     (DoubleCRT.cpp:79-113, 244-257, 349-398; NumbTh.cpp:307-335);
 *   - Ciphertext *=, ScaleDown, ByteDecomp (Ciphertext.cpp:82-121, 167-218);
 *   - KeySwitchSI::ApplyKeySwitch with DotProduct (FHE-SI.cpp:241-260, Util.h:80-98).
 *
 * The reference itself cannot be built here (NTL/GMP absent), so this is a
 * "port", labelled so wherever its timings are reported.  `faithful` = 1 reproduces the
 * reference's table bug: the Bluestein `powers` table is recomputed on every transform
 * because the cache test compares deg(powers) with n (bluestein.cpp:103-109).
 *
 * Known deviations from the NTL build (timing only, results are exact): NTL's FFT primes,
 * its FFT-based polynomial rem and its MulMod are replaced by straightforward equivalents
 * (three 62-bit FFT primes, schoolbook rem by Phi_m, 128-bit remainder).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

#define MAXL 48
#define BIGK 40 /* capacity: 64-bit limbs of the fixed-width two's-complement big integers */
static int BK = BIGK; /* limbs actually used, sized from the chain product in ref_create */

static inline u64 mulmod(u64 a, u64 b, u64 p) { return (u64)((u128)a * b % p); }
static u64 powmod(u64 a, u64 e, u64 p) {
  u64 r = 1;
  a %= p;
  while (e) {
    if (e & 1) r = mulmod(r, a, p);
    a = mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
static u64 invmod(u64 a, u64 p) { return powmod(a, p - 2, p); }
static int is_prime64(u64 n) {
  static const u64 B[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (int i = 0; i < 12; ++i)
    if (n % B[i] == 0) return n == B[i];
  u64 d = n - 1;
  int s = 0;
  while (!(d & 1)) d >>= 1, ++s;
  for (int i = 0; i < 12; ++i) {
    u64 x = powmod(B[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; ++r) {
      x = mulmod(x, x, n);
      if (x == n - 1) {
        comp = 0;
        break;
      }
    }
    if (comp) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------ FFT primes (NTL fftRep) */
typedef struct {
  u64 q;
  u64 *tw, *twp;   /* forward twiddles and Shoup quotients, per stage packed [h + j] */
  u64 *itw, *itwp; /* inverse */
  u64 ninv, ninvp;
} FFTPrime;

static inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static inline u64 mulshoup(u64 x, u64 w, u64 wp, u64 q) {
  u64 hi = (u64)(((u128)x * wp) >> 64);
  u64 r = x * w - hi * q;
  return r >= q ? r - q : r;
}
static void fftprime_init(FFTPrime *F, u64 q, int k) {
  u64 N = 1ull << k;
  F->q = q;
  F->tw = malloc(N * 8), F->twp = malloc(N * 8), F->itw = malloc(N * 8), F->itwp = malloc(N * 8);
  u64 z = 2;
  while (powmod(z, (q - 1) / 2, q) != q - 1) ++z;
  u64 w = powmod(z, (q - 1) / N, q), wi = invmod(w, q);
  for (u64 h = 1; h < N; h <<= 1) {
    u64 st = powmod(w, N / (2 * h), q), ist = powmod(wi, N / (2 * h), q), a = 1, b = 1;
    for (u64 j = 0; j < h; ++j) {
      F->tw[h + j] = a, F->twp[h + j] = shoup(a, q);
      F->itw[h + j] = b, F->itwp[h + j] = shoup(b, q);
      a = mulmod(a, st, q), b = mulmod(b, ist, q);
    }
  }
  F->ninv = invmod(N % q, q);
  F->ninvp = shoup(F->ninv, q);
}
static void fft_fwd(const FFTPrime *F, u64 *x, int k) { /* DIF, bit-reversed output */
  u64 N = 1ull << k, q = F->q;
  for (u64 h = N >> 1; h >= 1; h >>= 1)
    for (u64 s = 0; s < N; s += 2 * h)
      for (u64 j = 0; j < h; ++j) {
        u64 X = x[s + j], Y = x[s + j + h];
        u64 a = X + Y;
        x[s + j] = a >= q ? a - q : a;
        x[s + j + h] = mulshoup(X + q - Y, F->tw[h + j], F->twp[h + j], q);
      }
}
static void fft_inv(const FFTPrime *F, u64 *x, int k) { /* DIT from bit-reversed, scaled */
  u64 N = 1ull << k, q = F->q;
  for (u64 h = 1; h < N; h <<= 1)
    for (u64 s = 0; s < N; s += 2 * h)
      for (u64 j = 0; j < h; ++j) {
        u64 X = x[s + j], T = mulshoup(x[s + j + h], F->itw[h + j], F->itwp[h + j], q);
        u64 a = X + T, b = X + q - T;
        x[s + j] = a >= q ? a - q : a;
        x[s + j + h] = b >= q ? b - q : b;
      }
  for (u64 i = 0; i < N; ++i) x[i] = mulshoup(x[i], F->ninv, F->ninvp, q);
}

/* ------------------------------------------------------------------ context */
typedef struct {
  u64 q, root, rinv;
  u64 *powers, *ipowers;     /* root^{i^2}, rinv^{i^2}                 bluestein.cpp:103-109 */
  u64 *Rb[3], *iRb[3];       /* FFT images of the chirp filter b       bluestein.cpp:121-136 */
  int have_fwd, have_inv;
  u64 minv;
} Cmod;

typedef struct {
  u32 m, n, logQ, W, D, dbits, L, k;
  u64 p_pt;
  int faithful;
  u32 *units; /* [n] */
  i64 *phi;   /* Phi_m coefficients, [n+1] */
  Cmod mod[MAXL];
  FFTPrime F[3];
  u64 *scratch[3];
  u64 transforms; /* statistics */
} RefCtx;

static void cyclo(RefCtx *c) { /* Phi_m for the shapes the reference's parameter sets use */
  u32 m = c->m, n = c->n, h = m / 2;
  c->phi = calloc(n + 1, sizeof(i64));
  if (m % 2 == 0 && n == h - 1) { /* m = 2p': Phi_m(X) = Phi_p'(-X) */
    for (u32 i = 0; i <= n; ++i) c->phi[i] = (i & 1) ? -1 : 1;
  } else if (n == m - 1) { /* m prime */
    for (u32 i = 0; i <= n; ++i) c->phi[i] = 1;
  } else {
    abort();
  }
}

void *ref_create(u32 m, u32 logQ, u64 p_pt, u32 decompSize, u32 L, const u64 *primes, const u64 *roots,
                 int faithful) {
  RefCtx *c = calloc(1, sizeof(RefCtx));
  c->m = m, c->logQ = logQ, c->p_pt = p_pt, c->L = L, c->faithful = faithful;
  c->W = (logQ + 31) / 32, c->dbits = 8 * decompSize, c->D = (logQ + c->dbits - 1) / c->dbits;
  c->units = malloc(m * 4);
  u32 n = 0;
  for (u32 i = 0; i < m; ++i) {
    u32 a = i, b = m;
    while (b) {
      u32 t = a % b;
      a = b, b = t;
    }
    if (a == 1) c->units[n++] = i;
  }
  c->n = n;
  cyclo(c);
  int k = 0;
  while ((1u << k) < 2 * m - 1) ++k; /* NextPowerOfTwo(2n-1) with n := m, bluestein.cpp:116 */
  c->k = k;
  u64 N = 1ull << k;
  /* three FFT primes just below 2^62, = 1 mod 2^k */
  u64 q = ((1ull << 62) / N) * N + 1;
  for (int j = 0; j < 3;) {
    q -= N;
    if (is_prime64(q)) fftprime_init(&c->F[j++], q, k);
  }
  for (int j = 0; j < 3; ++j) c->scratch[j] = malloc(N * 8);
  double bits = 0;
  for (u32 l = 0; l < L; ++l) {
    u64 t = primes[l];
    while (t) bits += 1, t >>= 1;
  }
  BK = (int)(bits + 64 + 2) / 64 + 1;
  if (BK > BIGK) abort();
  for (u32 l = 0; l < L; ++l) {
    Cmod *M = &c->mod[l];
    M->q = primes[l], M->root = roots[l], M->rinv = invmod(roots[l], primes[l]);
    M->powers = malloc(m * 8), M->ipowers = malloc(m * 8);
    for (int j = 0; j < 3; ++j) M->Rb[j] = malloc(N * 8), M->iRb[j] = malloc(N * 8);
    M->minv = invmod(m % M->q, M->q);
  }
  return c;
}

/* BluesteinFFT, bluestein.cpp:93-144.  x[0..n) <- DFT of a[0..n) at root^2 (n = m here). */
static void bluestein(RefCtx *c, const Cmod *M, u64 *x, const u64 *a, u64 root, u64 *powers,
                      u64 **Rb, int *have) {
  const u32 n = c->m;
  const u64 p = M->q;
  const int k = (int)c->k;
  const u64 N = 1ull << k;
  c->transforms++;
  if (!*have || c->faithful) { /* powers table: rebuilt every call in the reference (bug) */
    powers[0] = 1;
    for (u32 i = 1; i < n; ++i) powers[i] = powmod(root, ((u64)i * i) % (2 * n), p);
  }
  if (!*have) { /* Rb is cached correctly (bluestein.cpp:121) */
    u64 rinv = invmod(root, p);
    u64 *b = calloc(N, 8);
    b[n - 1] = 1;
    for (u32 i = 1; i < n; ++i) {
      u64 bi = powmod(rinv, ((u64)i * i) % (2 * n), p);
      b[n - 1 + i] = b[n - 1 - i] = bi;
    }
    for (int j = 0; j < 3; ++j) {
      for (u64 i = 0; i < N; ++i) Rb[j][i] = b[i] % c->F[j].q;
      fft_fwd(&c->F[j], Rb[j], k);
    }
    free(b);
    *have = 1;
  }
  /* Ra = a .* powers, then cyclic convolution with b through the three FFT primes */
  for (int j = 0; j < 3; ++j) memset(c->scratch[j], 0, N * 8);
  for (u32 i = 0; i < n; ++i) {
    u64 v = mulmod(a[i], powers[i], p);
    for (int j = 0; j < 3; ++j) c->scratch[j][i] = v % c->F[j].q;
  }
  for (int j = 0; j < 3; ++j) {
    const FFTPrime *F = &c->F[j];
    u64 *s = c->scratch[j];
    fft_fwd(F, s, k);
    for (u64 i = 0; i < N; ++i) s[i] = mulmod(s[i], Rb[j][i], F->q);
    fft_inv(F, s, k);
  }
  /* CRT the three images (Garner) and reduce mod p; keep coefficients n-1 .. 2(n-1) */
  const u64 q0 = c->F[0].q, q1 = c->F[1].q, q2 = c->F[2].q;
  const u64 i01 = invmod(q0 % q1, q1), i012 = invmod(mulmod(q0 % q2, q1 % q2, q2), q2);
  const u64 q0p = q0 % p, q01p = mulmod(q0 % p, q1 % p, p);
  for (u32 i = 0; i < n; ++i) {
    u64 r0 = c->scratch[0][n - 1 + i], r1 = c->scratch[1][n - 1 + i], r2 = c->scratch[2][n - 1 + i];
    u64 v1 = mulmod((r1 + q1 - r0 % q1) % q1, i01, q1);
    u64 t = (r0 % q2 + mulmod(q0 % q2, v1 % q2, q2)) % q2;
    u64 v2 = mulmod((r2 + q2 - t) % q2, i012, q2);
    u64 val = (r0 % p + mulmod(q0p, v1 % p, p) + mulmod(q01p, v2 % p, p)) % p;
    x[i] = mulmod(val, powers[i], p);
  }
}

/* ------------------------------------------------------------------ big integers */
typedef struct {
  u64 w[BIGK];
} Big; /* two's complement, little-endian limbs */

static void big_from_words(Big *b, const u32 *w, u32 W) { /* sign-extend W 32-bit words */
  u64 ext = (w[W - 1] >> 31) ? ~0ull : 0ull;
  for (int i = 0; i < BK; ++i) {
    u64 lo = (2u * i < W) ? w[2 * i] : (u32)ext;
    u64 hi = (2u * i + 1 < W) ? w[2 * i + 1] : (u32)ext;
    b->w[i] = lo | (hi << 32);
  }
}
static inline int big_neg(const Big *b) { return (int)(b->w[BK - 1] >> 63); }
static void big_negate(Big *b) {
  u64 c = 1;
  for (int i = 0; i < BK; ++i) {
    u64 t = ~b->w[i] + c;
    c = (c && t == 0);
    b->w[i] = t;
  }
}
static void big_add(Big *a, const Big *b) {
  u64 c = 0;
  for (int i = 0; i < BK; ++i) {
    u128 t = (u128)a->w[i] + b->w[i] + c;
    a->w[i] = (u64)t;
    c = (u64)(t >> 64);
  }
}
static void big_mul_small(Big *a, u64 s) { /* a *= s (s >= 0), two's complement safe */
  int neg = big_neg(a);
  if (neg) big_negate(a);
  u64 c = 0;
  for (int i = 0; i < BK; ++i) {
    u128 t = (u128)a->w[i] * s + c;
    a->w[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  if (neg) big_negate(a);
}
static u64 big_mod_small(const Big *a, u64 p) { /* non-negative residue (NTL rem) */
  Big t = *a;
  int neg = big_neg(&t);
  if (neg) big_negate(&t);
  u64 r = 0;
  for (int i = BK - 1; i >= 0; --i) r = (u64)((((u128)r << 64) | t.w[i]) % p);
  return (neg && r) ? p - r : r;
}
static int big_cmp(const Big *a, const Big *b) { /* signed compare */
  int na = big_neg(a), nb = big_neg(b);
  if (na != nb) return na ? -1 : 1;
  for (int i = BK - 1; i >= 0; --i)
    if (a->w[i] != b->w[i]) return a->w[i] > b->w[i] ? 1 : -1;
  return 0;
}
static void big_sar(Big *a, u32 sh) { /* arithmetic shift right = floor division by 2^sh */
  u64 ext = big_neg(a) ? ~0ull : 0ull;
  u32 ws = sh / 64, bs = sh % 64;
  for (int i = 0; i < BK; ++i) {
    u64 lo = (i + ws < BIGK) ? a->w[i + ws] : ext;
    u64 hi = (i + ws + 1 < BIGK) ? a->w[i + ws + 1] : ext;
    a->w[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
  }
}
static void big_set_pow2(Big *a, u32 e) {
  memset(a, 0, sizeof *a);
  a->w[e / 64] = 1ull << (e % 64);
}
/* Reduce, Util.cpp:3-26: low logQ bits, centred (or positive) */
static void big_reduce(Big *a, u32 logQ, int positive) {
  u32 ws = logQ / 64, bs = logQ % 64;
  for (int i = 0; i < BK; ++i) {
    if ((u32)i > ws || ((u32)i == ws && bs == 0)) a->w[i] = 0;
    else if ((u32)i == ws) a->w[i] &= (1ull << bs) - 1;
  }
  if (!positive) {
    u32 sb = logQ - 1;
    if ((a->w[sb / 64] >> (sb % 64)) & 1) { /* subtract 2^logQ */
      Big q;
      big_set_pow2(&q, logQ);
      big_negate(&q);
      big_add(a, &q);
    }
  }
}
static void big_to_words(const Big *a, u32 *w, u32 W) {
  u32 ext = big_neg(a) ? 0xFFFFFFFFu : 0u;
  for (u32 i = 0; i < W; ++i) w[i] = (i / 2 < (u32)BK) ? (u32)(a->w[i / 2] >> (32 * (i & 1))) : ext;
}

/* ------------------------------------------------------------------ Cmodulus::FFT / iFFT */
/* CModulus.cpp:90-107: row[j] = poly(zeta^{u_j}) mod q */
static void cmod_fft(RefCtx *c, u32 l, u64 *row, const Big *poly, u32 len) {
  Cmod *M = &c->mod[l];
  const u32 m = c->m;
  u64 *in = calloc(m, 8), *out = malloc(m * 8);
  for (u32 i = 0; i < len && i < m; ++i) in[i] = big_mod_small(&poly[i], M->q); /* conv(in, x) */
  bluestein(c, M, out, in, M->root, M->powers, M->Rb, &M->have_fwd);
  for (u32 j = 0; j < c->n; ++j) row[j] = out[c->units[j]];
  free(in), free(out);
}
/* CModulus.cpp:110-132: coefficients mod q from a row */
static void cmod_ifft(RefCtx *c, u32 l, u64 *coef, const u64 *row) {
  Cmod *M = &c->mod[l];
  const u32 m = c->m, n = c->n;
  const u64 p = M->q;
  u64 *in = calloc(m, 8), *out = malloc(m * 8);
  for (u32 j = 0; j < n; ++j) in[c->units[j]] = row[j];
  bluestein(c, M, out, in, M->rinv, M->ipowers, M->iRb, &M->have_inv);
  for (u32 i = 0; i < m; ++i) out[i] = mulmod(out[i], M->minv, p); /* out /= m */
  for (int i = (int)m - 1; i >= (int)n; --i) {                     /* rem(out, out, Phi_m) */
    u64 ci = out[i];
    if (!ci) continue;
    for (u32 j = 0; j <= n; ++j) {
      u64 t = c->phi[j] < 0 ? ci : p - ci; /* out[i-n+j] -= ci * phi[j] */
      if (c->phi[j]) out[i - n + j] = (out[i - n + j] + t) % p;
    }
  }
  memcpy(coef, out, n * 8);
  free(in), free(out);
}

/* DoubleCRT(const ZZX&), DoubleCRT.cpp:244-257 */
static void dcrt_from_poly(RefCtx *c, u64 *rows, const Big *poly) {
  for (u32 l = 0; l < c->L; ++l) cmod_fft(c, l, rows + (size_t)l * c->n, poly, c->n);
}
/* DoubleCRT::toPoly, DoubleCRT.cpp:349-398 with intVecCRT, NumbTh.cpp:307-335 */
static void dcrt_to_poly(RefCtx *c, Big *poly, const u64 *rows) {
  const u32 n = c->n;
  u64 *cur = malloc(n * 8);
  Big P;
  memset(&P, 0, sizeof P);
  for (u32 l = 0; l < c->L; ++l) {
    const u64 q = c->mod[l].q;
    cmod_ifft(c, l, cur, rows + (size_t)l * n);
    if (l == 0) {
      P.w[0] = q;
      for (u32 j = 0; j < n; ++j) {
        memset(&poly[j], 0, sizeof(Big));
        poly[j].w[0] = cur[j];
        if (cur[j] > q / 2) { /* vp[j] -= p */
          Big t;
          memset(&t, 0, sizeof t);
          t.w[0] = q;
          big_negate(&t);
          big_add(&poly[j], &t);
        }
      }
      continue;
    }
    const u64 pinv = invmod(big_mod_small(&P, q), q), qh = q / 2;
    for (u32 j = 0; j < n; ++j) {
      u64 vp = big_mod_small(&poly[j], q);
      u64 d = mulmod((cur[j] + q - vp) % q, pinv, q);
      Big t = P;
      if (d > qh) {
        big_mul_small(&t, q - d);
        big_negate(&t);
      } else {
        big_mul_small(&t, d);
      }
      big_add(&poly[j], &t);
    }
    big_mul_small(&P, q);
  }
  free(cur);
}

/* ------------------------------------------------------------------ the op */
/* c = a; c *= b; ks.ApplyKeySwitch(c).  a, b, out: [2][n][W] words.  ksw rows: [2][3D][L][n]. */
void ref_mult_relin(void *ctx, const u32 *a, const u32 *b, const u64 *ksw_rows, u32 *out) {
  RefCtx *c = ctx;
  const u32 n = c->n, L = c->L, W = c->W, D = c->D;
  const size_t R = (size_t)L * n;
  Big *poly = malloc(sizeof(Big) * n);
  u64 *c1 = malloc(2 * R * 8), *c2 = malloc(2 * R * 8), *tp = calloc(3 * R, 8), *tmp = malloc(R * 8);
  /* Ciphertext::operator*=, Ciphertext.cpp:167-192 */
  for (u32 i = 0; i < 2; ++i) {
    for (u32 j = 0; j < n; ++j) {
      big_from_words(&poly[j], a + ((size_t)i * n + j) * W, W);
      big_mul_small(&poly[j], c->p_pt); /* parts[i].poly * p */
    }
    dcrt_from_poly(c, c1 + i * R, poly);
  }
  for (u32 i = 0; i < 2; ++i) {
    for (u32 j = 0; j < n; ++j) big_from_words(&poly[j], b + ((size_t)i * n + j) * W, W);
    dcrt_from_poly(c, c2 + i * R, poly);
  }
  for (u32 i = 0; i < 2; ++i)
    for (u32 j = 0; j < 2; ++j) {
      for (u32 l = 0; l < L; ++l) {
        const u64 q = c->mod[l].q;
        for (u32 e = 0; e < n; ++e) {
          size_t x = (size_t)l * n + e;
          tmp[x] = mulmod(c1[i * R + x], c2[j * R + x], q);
          tp[(i + j) * R + x] = (tp[(i + j) * R + x] + tmp[x]) % q;
        }
      }
    }
  /* ScaleDown, Ciphertext.cpp:194-218; then ByteDecomp, Ciphertext.cpp:82-121 */
  Big *parts = malloc(sizeof(Big) * 3 * n);
  Big q2h;
  big_set_pow2(&q2h, c->logQ); /* q */
  for (u32 t = 0; t < 3; ++t) {
    dcrt_to_poly(c, poly, tp + t * R);
    for (u32 j = 0; j < n; ++j) {
      Big v = poly[j];
      big_add(&v, &poly[j]); /* 2x */
      big_add(&v, &q2h);     /* + q */
      big_sar(&v, c->logQ + 1); /* / 2q, floor */
      big_reduce(&v, c->logQ, 0);
      parts[(size_t)t * n + j] = v;
    }
  }
  const u32 K = 3 * D;
  u64 *dig = malloc((size_t)K * R * 8);
  for (u32 t = 0; t < 3; ++t)
    for (u32 d = 0; d < D; ++d) {
      for (u32 j = 0; j < n; ++j) {
        Big v = parts[(size_t)t * n + j];
        big_reduce(&v, c->logQ, 1);
        big_sar(&v, c->dbits * d);
        u64 dv = v.w[0] & ((1ull << c->dbits) - 1);
        memset(&poly[j], 0, sizeof(Big));
        poly[j].w[0] = dv;
      }
      dcrt_from_poly(c, dig + (size_t)(t * D + d) * R, poly); /* FHE-SI.cpp:246-249 */
    }
  /* DotProduct + toPoly + Reduce, FHE-SI.cpp:251-257 */
  u64 *acc = malloc(R * 8);
  for (u32 r = 0; r < 2; ++r) {
    memset(acc, 0, R * 8);
    for (u32 k = 0; k < K; ++k)
      for (u32 l = 0; l < L; ++l) {
        const u64 q = c->mod[l].q;
        const u64 *kr = ksw_rows + ((size_t)(r * K + k) * L + l) * n;
        const u64 *dr = dig + (size_t)k * R + (size_t)l * n;
        for (u32 e = 0; e < n; ++e) acc[(size_t)l * n + e] = (acc[(size_t)l * n + e] + mulmod(kr[e], dr[e], q)) % q;
      }
    dcrt_to_poly(c, poly, acc);
    for (u32 j = 0; j < n; ++j) {
      big_reduce(&poly[j], c->logQ, 0);
      big_to_words(&poly[j], out + ((size_t)r * n + j) * W, W);
    }
  }
  free(acc), free(dig), free(parts), free(poly), free(c1), free(c2), free(tp), free(tmp);
}

/* DoubleCRT rows of one poly ([n][W] words) -> rows [L][n]; used to build the key matrix */
void ref_rows(void *ctx, const u32 *poly_words, u64 *rows) {
  RefCtx *c = ctx;
  Big *poly = malloc(sizeof(Big) * c->n);
  for (u32 j = 0; j < c->n; ++j) big_from_words(&poly[j], poly_words + (size_t)j * c->W, c->W);
  dcrt_from_poly(c, rows, poly);
  free(poly);
}
/* toPoly of rows, written as Wout-word two's complement */
void ref_to_poly(void *ctx, const u64 *rows, u32 *out, u32 Wout) {
  RefCtx *c = ctx;
  Big *poly = malloc(sizeof(Big) * c->n);
  dcrt_to_poly(c, poly, rows);
  for (u32 j = 0; j < c->n; ++j) big_to_words(&poly[j], out + (size_t)j * Wout, Wout);
  free(poly);
}
u64 ref_transform_count(void *ctx) { return ((RefCtx *)ctx)->transforms; }
void ref_set_faithful(void *ctx, int f) { ((RefCtx *)ctx)->faithful = f; }
void ref_destroy(void *ctx) { free(ctx); /* tables leak: process-lifetime test helper */ }
