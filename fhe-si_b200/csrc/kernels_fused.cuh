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
#define FPADN 1152u  // FN + 128
// exchange 1 (pass-1 <-> pass-2 ownership): 16 words of padding per 128; exchange 2
// (pass-2 <-> pass-3 ownership): 2 words per 16.  Both are bank-conflict free on both sides.
__device__ __forceinline__ u32 fpad1(u32 i) { return i + ((i >> 7) << 4); }
__device__ __forceinline__ u32 fpad2(u32 i) { return i + ((i >> 4) << 1); }

// Shoup multiplication by a fixed w: (w, floor(w 2^32 / p)); any x < 2^32 -> [0, 2p)
#ifdef MULW_NEGP
// x w - q p as (x w) + q (-p): the low product no longer waits for the quotient (dependent depth 2 instead of 3)
__device__ __forceinline__ u32 mulw(u32 x, uint2 w, u32 p) {
  const u32 t = x * w.x, q = __umulhi(x, w.y);
  u32 r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(q), "r"(0u - p), "r"(t));
  return r;
}
#else
__device__ __forceinline__ u32 mulw(u32 x, uint2 w, u32 p) { return x * w.x - __umulhi(x, w.y) * p; }
#endif

// The 21 twiddles a thread needs live in shared memory as tws[s * 128 + tg], s = 0..20
// (0-6 pass 1, 7-13 pass 2, 14-20 pass 3); one table per direction, shared by the CTA.
#define FTW_P2 (7u * 128u)            // entry offset of the pass-2 block
#define FTW_P3 (7u * 128u + 7u * 16u) // entry offset of the pass-3 block
#define FTW_ENTRIES (7u * 128u + 7u * 16u + 7u * 2u)
#define FTW_WORDS (FTW_ENTRIES * 2u + 4u)  // uint2 entries, padded to a 16-byte multiple
// FP64-quotient companions (mulw_dfma) for the pass-2 and pass-3 twiddles only: 7 * 16 + 7 * 2 entries of
// (c, K).  Pass 1's twiddles differ per thread -- 16 more bytes per lane per butterfly would make shared-
// memory bandwidth the limiter -- so pass 1 keeps the all-integer product.
#define FTD_ENTRIES (7u * 16u + 7u * 2u)
#define FTD_WORDS (FTD_ENTRIES * 4u)
// tws layout: pass 1: [s][tg] (s < 7, 128 per row); pass 2: [s][lo] (16 per row); pass 3: [s][b0]
// index into the [N] Shoup table of entry e of the per-thread layout above
__host__ __device__ __forceinline__ u32 ftw_source_index(u32 e) {
  if (e < FTW_P2) {
    const u32 s = e >> 7, t = e & 127;
    return s < 4 ? 512 + s * 128 + t : (s < 6 ? 256 + (s - 4) * 128 + t : 128 + t);
  }
  if (e < FTW_P3) {
    const u32 s = (e - FTW_P2) >> 4, lo = (e - FTW_P2) & 15;
    return s < 4 ? 64 + s * 16 + lo : (s < 6 ? 32 + (s - 4) * 16 + lo : 16 + lo);
  }
  const u32 s = (e - FTW_P3) >> 1, b0 = (e - FTW_P3) & 1;
  return s < 4 ? 8 + s * 2 + b0 : (s < 6 ? 4 + (s - 4) * 2 + b0 : 2 + b0);
}
// The per-prime tables are stored in HBM already in this layout (DevCtx::ftw_fwd / ftw_inv, built once at
// context creation), so a CTA's fill is a straight coalesced copy -- the index arithmetic above used to cost
// a CTA of the tensor kernel a third of one ciphertext pair's work.
__device__ __forceinline__ void fill_tw_table(uint2 *tws, const uint2 *__restrict__ laid_out) {
  for (u32 e = threadIdx.x; e < FTW_ENTRIES; e += blockDim.x) tws[e] = __ldg(laid_out + e);
}
// twd[e - FTW_P2] for the pass-2 / pass-3 entries e of fill_tw_table, same source index
__device__ __forceinline__ void fill_twd_table(double2 *twd, const double2 *__restrict__ table) {
  for (u32 e = FTW_P2 + threadIdx.x; e < FTW_ENTRIES; e += blockDim.x) twd[e - FTW_P2] = table[ftw_source_index(e)];
}

#define GSW(X, Y, W)                          \
  do {                                        \
    u32 s_ = add_alu((X), (Y)), d_ = (X) + p2 - (Y);  \
    (X) = csub(s_, p2);                       \
    (Y) = mulw(d_, (W), p);                   \
  } while (0)
#define CTW(X, Y, W)                          \
  do {                                        \
    u32 t_ = mulw((Y), (W), p);               \
    u32 s_ = add_alu((X), t_), d_ = (X) + p2 - t_;    \
    (X) = csub(s_, p2);                       \
    (Y) = csub(d_, p2);                       \
  } while (0)

#define GSWD(X, Y, W, C)                      \
  do {                                        \
    u32 s_ = add_alu((X), (Y)), d_ = (X) + p2 - (Y);  \
    (X) = csub(s_, p2);                       \
    (Y) = mulw_dfma(d_, (W).x, (C).x, (C).y, p, negp, hic, zop);  \
  } while (0)
// dif8 with the twiddle products' quotients on the FP64 pipe; td[] is indexed like tw[]
template <int ST>
__device__ __forceinline__ void dif8d(u32 *x, const uint2 *tw, const double2 *td, u32 p, u32 negp, u32 hic, u32 zop) {
  const u32 p2 = 2 * p;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint2 w = tw[j * ST];
    const double2 c = td[j * ST];
    GSWD(x[j], x[j + 4], w, c);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint2 w = tw[(4 + j) * ST];
    const double2 c = td[(4 + j) * ST];
    GSWD(x[j], x[j + 2], w, c);
    GSWD(x[j + 4], x[j + 6], w, c);
  }
  const uint2 w = tw[6 * ST];
  const double2 c = td[6 * ST];
#pragma unroll
  for (int j = 0; j < 8; j += 2) GSWD(x[j], x[j + 1], w, c);
}
// three DIF stages on the 8 registers, twiddles tw[0..3], tw[4..5], tw[6] (stride 128 apart)
template <int ST>
__device__ __forceinline__ void dif8(u32 *x, const uint2 *tw, u32 p) {
  const u32 p2 = 2 * p;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint2 w = tw[j * ST];
    GSW(x[j], x[j + 4], w);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint2 w = tw[(4 + j) * ST];
    GSW(x[j], x[j + 2], w);
    GSW(x[j + 4], x[j + 6], w);
  }
  const uint2 w = tw[6 * ST];
#pragma unroll
  for (int j = 0; j < 8; j += 2) GSW(x[j], x[j + 1], w);
}
template <int ST>
__device__ __forceinline__ void dit8(u32 *x, const uint2 *tw, u32 p) {
  const u32 p2 = 2 * p;
  {
    const uint2 w = tw[6 * ST];
#pragma unroll
    for (int j = 0; j < 8; j += 2) CTW(x[j], x[j + 1], w);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint2 w = tw[(4 + j) * ST];
    CTW(x[j], x[j + 2], w);
    CTW(x[j + 4], x[j + 6], w);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint2 w = tw[j * ST];
    CTW(x[j], x[j + 4], w);
  }
}

// per-thread shared-memory offsets of the three ownerships (loop invariant)
struct XAddr {
  u32 a1, a2a, a2b, a3;  // pass-1 view of bufA, pass-2 view of bufA, of bufB, pass-3 view of bufB
};
__device__ __forceinline__ XAddr make_xaddr(u32 tg) {
  const u32 hi = tg >> 4, lo = tg & 15, u = tg >> 1, b0 = tg & 1;
  XAddr r;
  r.a1 = tg;                    // fpad1(j*128 + tg)          = a1  + 144 j
  r.a2a = hi * 144 + lo;        // fpad1(hi*128 + j*16 + lo)  = a2a + 16 j
  r.a2b = hi * 144 + lo;        // fpad2(hi*128 + j*16 + lo)  = a2b + 18 j
  r.a3 = u * 18 + b0;           // fpad2(u*16 + j*2 + b0)     = a3  + 2 j
  return r;
}

// Forward transform of a polynomial whose upper half is zero.  In: x[0..3] = coefficients
// j*128 + tg (< 2p; x[4..7] ignored).  Out: x[r] = value at storage index tg*8 + r, in [0,p);
// with OFFS the output is the signed value v - (p >> 1), v in [0,p): a balanced residue of the
// value minus the constant (p >> 1) -- one subtract instead of a compare/select/subtract, the
// constant's contribution to an inner product is added back from a per-key table (k_split_corr).
// With DF (twd non-null at compile time) passes 2 and 3 take their quotients from the FP64 pipe.
template <bool OFFS = false, bool DF = false>
__device__ __forceinline__ void fwd1024(u32 *x, const uint2 *twf, const XAddr &A, u32 *bufA, u32 *bufB,
                                        u32 g, u32 tg, u32 p, const double2 *twd = nullptr, u32 negp = 0, u32 hic = 0, u32 zop = 0) {
  const u32 p2 = 2 * p;
  const uint2 *tw = twf + tg;
  // pass 1; its first stage sees (X, 0): X stays, the partner becomes X * w
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j + 4] = mulw(x[j], tw[j * 128], p);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint2 w = tw[(4 + j) * 128];
    GSW(x[j], x[j + 2], w);
    GSW(x[j + 4], x[j + 6], w);
  }
  {
    const uint2 w = tw[6 * 128];
#pragma unroll
    for (int j = 0; j < 8; j += 2) GSW(x[j], x[j + 1], w);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[A.a1 + 144 * j] = x[j];
  fhesi_group_sync(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[A.a2a + 16 * j];
  if (DF) dif8d<16>(x, twf + FTW_P2 + (tg & 15), twd + (tg & 15), p, negp, hic, zop);
  else dif8<16>(x, twf + FTW_P2 + (tg & 15), p);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[A.a2b + 18 * j] = x[j];
  // exchange 2 stays inside a warp: warp w owns positions [256w, 256w+256) in both ownerships
  // (pass 2: hi in {2w, 2w+1}; pass 3: u in [16w, 16w+16)), so a warp barrier is enough
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[A.a3 + 2 * j];
  if (DF) dif8d<2>(x, twf + FTW_P3 + (tg & 1), twd + (FTW_P3 - FTW_P2) + (tg & 1), p, negp, hic, zop);
  else dif8<2>(x, twf + FTW_P3 + (tg & 1), p);
  // last stage (position bit 0) across lane pairs, twiddle 1
  const u32 b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    u32 v = b0 ? o + p2 - x[j] : add_alu(x[j], o);
    x[j] = csub(csub(v, p2), p);
    if (OFFS) x[j] = sub_alu(x[j], p >> 1);
  }
}
// Inverse transform.  In: x[r] = value at storage index tg*8 + r (< 2p).  Out: the natural
// order result (unscaled, < 2p) is written to nat[0..1024) (nat may alias bufA) and the group
// is synchronised, so every thread may read any coefficient afterwards.
__device__ __forceinline__ void inv1024(u32 *x, const uint2 *twi, const XAddr &A, u32 *bufA, u32 *bufB,
                                        u32 *nat, u32 g, u32 tg, u32 p) {
  const u32 p2 = 2 * p;
  const uint2 *tw = twi + tg;
  const u32 b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    u32 v = b0 ? o + p2 - x[j] : add_alu(x[j], o);
    x[j] = csub(v, p2);
  }
  dit8<2>(x, twi + FTW_P3 + (tg & 1), p);
  __syncwarp();  // the warp's previous reads of its bufB region (last forward transform) are done
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[A.a3 + 2 * j] = x[j];
  __syncwarp();  // intra-warp exchange (see fwd1024)
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[A.a2b + 18 * j];
  dit8<16>(x, twi + FTW_P2 + (tg & 15), p);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[A.a2a + 16 * j] = x[j];
  fhesi_group_sync(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[A.a1 + 144 * j];
  dit8<128>(x, tw, p);
  // nat aliases bufA: every thread must be past its bufA reads before anyone writes
  fhesi_group_sync(g);
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

// General m (DevCtx::red, h == 0): the sparse remainder rows instead of the fold; T threads per transform.  The fused
// kernels take the choice as a template parameter GEN, so that the m = 2h instances compile exactly as without it
// (k_fused_keyswitch_split sits at 128 registers: a run-time branch costs it spills).
template <int T>
__device__ __forceinline__ void phim_store_csr(const u32 *nat, u32 *__restrict__ dst, const u32 *__restrict__ red, u32 n,
                                               u32 tg, u32 p) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = j * T + tg;
    if (i < n) dst[i] = phim_row_csr(nat, red, n, i, p);
  }
}
#define PHIM_STORE_1024(NAT, DST)                                   \
  do {                                                              \
    if (!GEN) phim_store_1024((NAT), (DST), c.h, tg, p);            \
    else phim_store_csr<128>((NAT), (DST), c.red, c.n, tg, p);      \
  } while (0)

static bool fused_supported(const DevCtx &dc) { return (dc.N == FN && dc.n <= 512) || (dc.N == 2048 && dc.n <= 1024); }

// multiword two's complement (W words) -> residue in [0,2p), scaled by s_v (DevCtx::cwr).
// Lazy: 4 word products are summed in 64 bits (< 2^64) and Montgomery-reduced once.
__device__ __forceinline__ u32 residue_lazy(const u32 *__restrict__ w, u32 W, const u32 *__restrict__ cw,
                                            u32 p, u32 pinv) {
  const u32 p2 = 2 * p;
  u32 r = 0, top = 0;
  for (u32 k0 = 0; k0 < W; k0 += 4) {
    u64 t = 0;
    if (k0 + 4 <= W && (W & 3) == 0) {
      const uint4 v = *(const uint4 *)(w + k0);
      t = (u64)v.x * __ldg(cw + k0) + (u64)v.y * __ldg(cw + k0 + 1) + (u64)v.z * __ldg(cw + k0 + 2) +
          (u64)v.w * __ldg(cw + k0 + 3);
      top = v.w;
    } else {
      for (u32 k = k0; k < W && k < k0 + 4; ++k) {
        top = w[k];
        t += (u64)top * __ldg(cw + k);
      }
    }
    // t < 4 * 2^32 * p.  t - m*p == 0 mod 2^32 with m = t_lo * p^-1; quotient t_hi - hi(m p)
    const u32 m = (u32)t * (0u - pinv);
    const u32 th = (u32)(t >> 32), hm = __umulhi(m, p);
    u32 q = th - hm;
    if (th < hm) q += p;                   // q in [0, 4p)
    r = csub(r + csub(q, p2), p2);
  }
  if (top >> 31) r = csub(r + __ldg(cw + W), p2);
  return r;
}

// ---------------------------------------------------------------------------------------
// multiword -> residues for all tensor primes, one thread per coefficient (the `conv(in, x)`
// of Cmodulus::FFT, CModulus.cpp:96).  a-parts come out scaled by p_pt/N in Montgomery form.
// in a, b: [count][2][n][W] -> out [count][4][Lt][n]  (poly order a0, a1, b0, b1)
// ---------------------------------------------------------------------------------------
struct ResidueArgs {
  const u32 *a, *b;
  u32 *out;
  u32 Lt;
  size_t count;
};
__global__ void __launch_bounds__(128) k_residues(DevCtx c, ResidueArgs a) {
  FHESI_SMEM(sm);  // cwr table of the first Lt primes: [Lt][2][CW], then (p, p^-1) per prime
  const u32 tabw = a.Lt * 2 * c.CW;
  for (u32 e = threadIdx.x; e < tabw; e += blockDim.x) sm[e] = __ldg(c.cwr + e);
  for (u32 e = threadIdx.x; e < a.Lt; e += blockDim.x) {
    sm[tabw + 2 * e] = c.pc[e].p;
    sm[tabw + 2 * e + 1] = 0u - c.pc[e].pinv;
  }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = a.count * 4 * c.n;
  if (idx >= total) return;
  const u32 i = (u32)(idx % c.n);
  const size_t pq = idx / c.n;  // op * 4 + q
  const u32 q = (u32)(pq & 3);
  const size_t op = pq >> 2;
  const u32 W = c.W;
  const u32 *src = (q < 2 ? a.a : a.b) + ((op * 2 + (q & 1)) * c.n + i) * (size_t)W;
  u32 w[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) w[k] = 0;
  if ((W & 3) == 0) {
#pragma unroll
    for (int k = 0; k < 16; k += 4)
      if ((u32)k < W) {
        const uint4 v = *(const uint4 *)(src + k);
        w[k] = v.x, w[k + 1] = v.y, w[k + 2] = v.z, w[k + 3] = v.w;
      }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if ((u32)k < W) w[k] = src[k];
  }
  const bool neg = (src[W - 1] >> 31) != 0;
  const u32 v = q < 2 ? 1u : 0u;
  u32 *dst = a.out + pq * a.Lt * (size_t)c.n + i;
#pragma unroll 2
  for (u32 l = 0; l < a.Lt; ++l) {
    const u32 p = sm[tabw + 2 * l], p2 = 2 * p, ipinv = sm[tabw + 2 * l + 1];
    const u32 *cw = sm + (l * 2 + v) * c.CW;
    u32 r = 0;
#pragma unroll
    for (int k0 = 0; k0 < 16; k0 += 4) {
      if ((u32)k0 < W) {
        // words beyond W are zero, so a partial last group needs no special case
        const u64 t = (u64)w[k0] * cw[k0] + (u64)w[k0 + 1] * (k0 + 1 < (int)W ? cw[k0 + 1] : 0u) +
                      (u64)w[k0 + 2] * (k0 + 2 < (int)W ? cw[k0 + 2] : 0u) +
                      (u64)w[k0 + 3] * (k0 + 3 < (int)W ? cw[k0 + 3] : 0u);
        const u32 m = (u32)t * ipinv;
        const u32 th = (u32)(t >> 32), hm = __umulhi(m, p);
        u32 qv = th - hm;
        if (th < hm) qv += p;  // [0, 4p)
        r = csub(r + csub(qv, p2), p2);
      }
    }
    if (neg) r = csub(r + cw[W], p2);
    dst[(size_t)l * c.n] = csub(r, p);
  }
}

// The same with the word count W a compile-time constant (the runtime-W kernel above spends most of its
// instructions on `k < W` tests and selects): all word loops unrolled, the per-prime constants read as two or
// three 128-bit shared-memory loads from a table padded to 4-word rows.  TW = 4, 6, 8, 16 cover logQ = 128,
// 176, 256, 512; anything else takes the generic kernel.
template <int TW>
__global__ void __launch_bounds__(128) k_residues_t(DevCtx c, ResidueArgs a) {
  constexpr int ROW = (TW + 1 + 3) & ~3;  // cw[0..TW] padded
  FHESI_SMEM(sm);  // [Lt][2][ROW] constants, then (p, -p^-1) per prime
  const u32 tabw = a.Lt * 2 * ROW;
  for (u32 e = threadIdx.x; e < tabw; e += blockDim.x) {
    const u32 k = e % ROW, lv = e / ROW;
    sm[e] = k <= (u32)TW ? __ldg(c.cwr + (size_t)lv * c.CW + k) : 0u;
  }
  for (u32 e = threadIdx.x; e < a.Lt; e += blockDim.x) {
    sm[tabw + 2 * e] = c.pc[e].p;
    sm[tabw + 2 * e + 1] = 0u - c.pc[e].pinv;
  }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.count * 4 * c.n) return;
  const u32 i = (u32)(idx % c.n);
  const size_t pq = idx / c.n;  // op * 4 + q
  const u32 q = (u32)(pq & 3);
  const size_t op = pq >> 2;
  const u32 *src = (q < 2 ? a.a : a.b) + ((op * 2 + (q & 1)) * c.n + i) * (size_t)TW;
  u32 w[TW];
  if (TW % 4 == 0) {
#pragma unroll
    for (int k = 0; k < TW; k += 4) {
      const uint4 v = *(const uint4 *)(src + k);
      w[k] = v.x, w[k + 1] = v.y, w[k + 2] = v.z, w[k + 3] = v.w;
    }
  } else if (TW % 2 == 0) {
#pragma unroll
    for (int k = 0; k < TW; k += 2) {
      const uint2 v = *(const uint2 *)(src + k);
      w[k] = v.x, w[k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < TW; ++k) w[k] = src[k];
  }
  const bool neg = (w[TW - 1] >> 31) != 0;
  const u32 *cwb = sm + (q < 2 ? ROW : 0);  // a-parts take the scaled constants (v = 1)
  u32 *dst = a.out + pq * a.Lt * (size_t)c.n + i;
  const u32 n = c.n;
#pragma unroll 2
  for (u32 l = 0; l < a.Lt; ++l) {
    const uint2 pp = *(const uint2 *)(sm + tabw + 2 * l);
    const u32 p = pp.x, p2 = 2 * p, ipinv = pp.y;
    u32 cw[ROW];
#pragma unroll
    for (int k = 0; k < ROW; k += 4) {
      const uint4 v = *(const uint4 *)(cwb + l * 2 * ROW + k);
      cw[k] = v.x, cw[k + 1] = v.y, cw[k + 2] = v.z, cw[k + 3] = v.w;
    }
    u32 r = 0;
#pragma unroll
    for (int k0 = 0; k0 < TW; k0 += 4) {
      u64 t = (u64)w[k0] * cw[k0];
#pragma unroll
      for (int k = k0 + 1; k < k0 + 4 && k < TW; ++k) t += (u64)w[k] * cw[k];
      // t < 4 * 2^32 * p.  t - m*p == 0 mod 2^32 with m = t_lo * p^-1; quotient t_hi - hi(m p) in (-p, 4p)
      const u32 m = (u32)t * ipinv;
      const u32 th = (u32)(t >> 32), hm = __umulhi(m, p);
      u32 qv = th - hm;
      if (th < hm) qv += p;  // [0, 4p)
      r = k0 == 0 ? csub(qv, p2) : csub(r + csub(qv, p2), p2);
    }
    if (neg) r = csub(r + cw[TW], p2);
    *dst = csub(r, p);
    dst += n;
  }
}

// ---------------------------------------------------------------------------------------
// tensor product, one 128-thread group per (prime, ciphertext pair); a group is independent
// of the other groups of its CTA (they only share the prime's twiddle tables)
// ---------------------------------------------------------------------------------------
struct FusedTensorArgs {
  const u32 *resin;  // [count][4][Lt][n]   residues of a0, a1 (scaled), b0, b1 from k_residues
  u32 *res;          // to_tprod == 0: [count][3][Lt][n]   coefficient residues after the Phi_m fold
                     // to_tprod == 1: [count][3][Lt][N]   transform-domain tprod
  u32 Lt, count, ops_per_group, to_tprod;
};
#ifndef KG
#define KG 2         // transform groups per CTA: 2 x 3 CTAs per SM measured 3 % faster than 6 x 1
#endif
#ifndef KG_MINB
#define KG_MINB 3    // resident CTAs per SM asked of ptxas
#endif
#define FUSED_SMEM_WORDS (2 * FTW_WORDS + KG * 3 * FPADN)
template <bool GEN>
__global__ void __launch_bounds__(KG * 128, KG_MINB) k_fused_tensor(DevCtx c, FusedTensorArgs a) {
  FHESI_SMEM(sm);
  uint2 *twf = (uint2 *)sm, *twi = (uint2 *)(sm + FTW_WORDS);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  const u32 l = blockIdx.x;
  fill_tw_table(twf, c.ftw_fwd + (size_t)l * FTW_ENTRIES);
  fill_tw_table(twi, c.ftw_inv + (size_t)l * FTW_ENTRIES);
  __syncthreads();
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  // two alternating inter-warp exchange buffers (a warp may run one transform ahead of the
  // others, since only the first exchange of a transform has a group barrier) + one intra-warp
  u32 *bufA0 = sm + 2 * FTW_WORDS + g * 3 * FPADN, *bufA1 = bufA0 + FPADN, *bufB = bufA1 + FPADN;
  u32 flip = 0;
  const XAddr A = make_xaddr(tg);
  for (u32 it = 0; it < a.ops_per_group; ++it) {
    const size_t op = ((size_t)blockIdx.y * a.ops_per_group + it) * KG + g;
    if (op >= a.count) return;  // only group barriers from here on
    // images of a0, a1 (scaled by p_pt/N, Montgomery form) and b0, b1 (plain), in registers
    u32 F[4][8];
    const u32 *rin = a.resin + ((op * 4) * a.Lt + l) * (size_t)c.n;
    const size_t qstride = (size_t)a.Lt * c.n;
    u32 xn[4];  // next polynomial's coefficients, prefetched one transform ahead
#pragma unroll
    for (int j = 0; j < 4; ++j) xn[j] = (j * 128 + tg < c.n) ? __ldg(rin + j * 128 + tg) : 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) F[q][j] = xn[j];
      if (q < 3) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          xn[j] = (j * 128 + tg < c.n) ? __ldg(rin + (q + 1) * qstride + j * 128 + tg) : 0u;
      }
      fwd1024(F[q], twf, A, (flip++ & 1) ? bufA1 : bufA0, bufB, g, tg, p);
    }
    // tProd[k] = sum_{i+j=k} a_i * b_j   (Ciphertext.cpp:179-186)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      u32 y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k == 0) y[j] = mont_mul(F[0][j], F[2][j], p, pinv);
        else if (k == 2) y[j] = mont_mul(F[1][j], F[3][j], p, pinv);
        else y[j] = csub(mont_mul(F[0][j], F[3][j], p, pinv) + mont_mul(F[1][j], F[2][j], p, pinv), p2);
      }
      if (a.to_tprod) {
        u32 *dst = a.res + ((op * 3 + k) * a.Lt + l) * (size_t)FN + tg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = csub(y[j], p);
      } else {
        u32 *bufA = (flip++ & 1) ? bufA1 : bufA0;
        inv1024(y, twi, A, bufA, bufB, bufA, g, tg, p);
        PHIM_STORE_1024(bufA, a.res + ((op * 3 + k) * a.Lt + l) * (size_t)c.n);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// key switch: one 128-thread group per (prime, ciphertext); the KG ciphertexts of a CTA share
// the prime's twiddle tables (shared memory) and key tiles (L1)
// ---------------------------------------------------------------------------------------
struct FusedKsArgs {
  const u32 *digits;  // [count][K][n]   dbits-wide digits, part-major digit-minor
  const u32 *key;     // [Lk][K][2][N]   key form, storage order; TFREE: balanced (int32 in (-p/2, p/2])
  u32 *res;           // [count][2][Lk][n]
  u32 K, Lk, count;
};
// G transform groups (ciphertexts) per CTA.  TFREE: every prime satisfies K * (p/2)^2 < 2^63, so
// the whole inner product of balanced residues fits one signed 64-bit accumulator and no
// intermediate reduction (nor its 16 partial-sum registers) is needed.
#define KSG 6
#define KS_SMEM_WORDS (2 * FTW_WORDS + KSG * 3 * FPADN)
template <bool TFREE, bool GEN>
__global__ void __launch_bounds__(KSG * 128, 1) k_fused_keyswitch(DevCtx c, FusedKsArgs a) {
  FHESI_SMEM(sm);
  uint2 *twf = (uint2 *)sm, *twi = (uint2 *)(sm + FTW_WORDS);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
  const u32 l = blockIdx.x;
  fill_tw_table(twf, c.ftw_fwd + (size_t)l * FTW_ENTRIES);
  fill_tw_table(twi, c.ftw_inv + (size_t)l * FTW_ENTRIES);
  __syncthreads();
  const size_t op = (size_t)blockIdx.y * KSG + g;
  if (op >= a.count) return;  // whole group leaves together; only group barriers from here on
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p, half = p >> 1;
  u32 *bufA0 = sm + 2 * FTW_WORDS + g * 3 * FPADN, *bufA1 = bufA0 + FPADN, *bufB = bufA1 + FPADN;
  const XAddr A = make_xaddr(tg);
  u64 acc0[8], acc1[8];
  u32 t0[8], t1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0, t0[j] = t1[j] = 0;
  const u32 *dig = a.digits + op * a.K * (size_t)c.n;
  const u32 *key = a.key + (size_t)l * a.K * 2 * FN + tg * 8;
  u32 xn[4];  // next digit, prefetched one transform ahead
#pragma unroll
  for (int j = 0; j < 4; ++j) xn[j] = (j * 128 + tg < c.n) ? __ldg(dig + j * 128 + tg) : 0u;
  for (u32 k = 0; k < a.K; ++k) {
    u32 x[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = xn[j];
    if (k + 1 < a.K) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        xn[j] = (j * 128 + tg < c.n) ? __ldg(dig + (size_t)(k + 1) * c.n + j * 128 + tg) : 0u;
    }
    fwd1024(x, twf, A, (k & 1) ? bufA1 : bufA0, bufB, g, tg, p);
    const uint4 *kp = (const uint4 *)(key + (size_t)k * 2 * FN);
    const uint4 ka0 = __ldg(kp), ka1 = __ldg(kp + 1);
    const uint4 kb0 = __ldg(kp + FN / 4), kb1 = __ldg(kp + FN / 4 + 1);
    const u32 kb[8] = {ka0.x, ka0.y, ka0.z, ka0.w, ka1.x, ka1.y, ka1.z, ka1.w};
    const u32 kA[8] = {kb0.x, kb0.y, kb0.z, kb0.w, kb1.x, kb1.y, kb1.z, kb1.w};
    if (TFREE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int xb = (int)x[j] - (x[j] > half ? (int)p : 0);  // balanced residue
        acc0[j] += (u64)((i64)xb * (int)kb[j]);
        acc1[j] += (u64)((i64)xb * (int)kA[j]);
      }
    } else {
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
  }
  if (TFREE) {  // |sum| < 2^63: reduce the magnitude, then restore the sign
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const i64 s0 = (i64)acc0[j], s1 = (i64)acc1[j];
      u32 r0 = csub(mont_red64((u64)(s0 < 0 ? -s0 : s0), p, pinv), p2);
      u32 r1 = csub(mont_red64((u64)(s1 < 0 ? -s1 : s1), p, pinv), p2);
      t0[j] = s0 < 0 ? csub(p2 - r0, p2) : r0;
      t1[j] = s1 < 0 ? csub(p2 - r1, p2) : r1;
    }
  }
  fhesi_group_sync(g);  // every warp is done with the forward transforms' buffers
  inv1024(t0, twi, A, bufA0, bufB, bufA0, g, tg, p);
  PHIM_STORE_1024(bufA0, a.res + ((op * 2 + 0) * a.Lk + l) * (size_t)c.n);
  inv1024(t1, twi, A, bufA1, bufB, bufA1, g, tg, p);
  PHIM_STORE_1024(bufA1, a.res + ((op * 2 + 1) * a.Lk + l) * (size_t)c.n);
}
// ---------------------------------------------------------------------------------------
// split-key key switch.  Every key polynomial K (mod q) is stored as two non-negative halves,
// K = K_lo + 2^(32 ws) K_hi, so each inner product sum_k digit_k * K_half,k is bounded by
// ~2^(24 + 32 ws) * 3D * 2n instead of 2^(24 + logQ): it needs Ls ~ 0.6 Lk primes.  The 3D digit
// transforms -- the dominant cost -- are shared by the four accumulators (b_lo, b_hi, A_lo, A_hi),
// so forward transforms drop from 3D*Lk to 3D*Ls (330 -> 198 at logQ = 256) for 20 % more
// multiply-accumulates.  Requires the single-accumulator (TFREE) bound.
// key: [Ls][K][4][N] balanced;  res: [count][4][Ls][n] in the order b_lo, b_hi, A_lo, A_hi.
// ---------------------------------------------------------------------------------------
#ifndef KSS
#define KSS 1        // transform groups (ciphertexts) per CTA: 1 x 4 CTAs per SM measured 2 % faster than 4 x 1
#endif
#ifndef KSS_MINB
#define KSS_MINB 4   // resident CTAs per SM asked of ptxas (KSS * KSS_MINB * 128 threads at <= 128 registers)
#endif
#ifndef KSS_DFMA
#define KSS_DFMA false
#endif
#define KSS_BUFA 2048u  // word distance of the two alternating exchange buffers (a power of two: XOR toggle)
#define KSS_SMEM_WORDS (2 * FTW_WORDS + FTD_WORDS + KSS * (2 * KSS_BUFA + FPADN))
// key: [Ls][K][4][N] balanced, followed by the offset-correction table [Ls][4][N] (k_split_corr)
template <bool GEN>
__global__ void __launch_bounds__(KSS * 128, KSS_MINB) k_fused_keyswitch_split(DevCtx c, FusedKsArgs a) {
  FHESI_SMEM(sm);
  uint2 *twf = (uint2 *)sm, *twi = (uint2 *)(sm + FTW_WORDS);
  const u32 g = threadIdx.x >> 7, tg = threadIdx.x & 127;
    const u32 l = blockIdx.x;
  double2 *twd = (double2 *)(sm + 2 * FTW_WORDS);
  fill_tw_table(twf, c.ftw_fwd + (size_t)l * FTW_ENTRIES);
  fill_tw_table(twi, c.ftw_inv + (size_t)l * FTW_ENTRIES);
  if (KSS_DFMA) fill_twd_table(twd, c.twd_fwd + (size_t)l * FN);
  __syncthreads();
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  const u32 negp = pc.negp, hic = pc.hic, zop = pc.zop;  // loaded, hence opaque: see mulw_dfma
  u32 *bufA0 = sm + 2 * FTW_WORDS + FTD_WORDS + g * (2 * KSS_BUFA + FPADN), *bufB = bufA0 + 2 * KSS_BUFA;
  const XAddr A = make_xaddr(tg);
  const size_t op = (size_t)blockIdx.y * KSS + g;
  if (op >= a.count) return;  // only group barriers from here on
  u64 acc[4][8];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h][j] = 0;
  // digit rows are read unconditionally (a row is n <= 512 words; positions >= n fall into the next
  // row or the scratch area that follows the digit buffer) and masked: N = 1024 means n > 256, so
  // only the upper two of a thread's four coefficients can lie beyond n
  const u32 m2 = (256 + tg < c.n) ? 0xFFFFFFFFu : 0u, m3 = (384 + tg < c.n) ? 0xFFFFFFFFu : 0u;
  const u32 *dp = a.digits + op * a.K * (size_t)c.n + tg;
  const uint4 *kp = (const uint4 *)(a.key + (size_t)l * a.K * 4 * FN + tg * 8);
  // The next digit's four words per thread are fetched one transform ahead by an asynchronous copy
  // (cp.async, LDGSTS) into the unused tail of the exchange buffer the current transform does NOT use: no
  // registers are held across the transform and the fetch cannot be scheduled late (with plain loads ptxas
  // sank them to the end of the loop body to save registers, exposing their whole latency).  A thread reads
  // back only what it wrote itself, so cp.async.wait_group is all the synchronisation needed.
  u32 *stage0 = bufA0 + FPADN + tg, *stage1 = stage0 + KSS_BUFA;  // 4 words each at +0, +128, +256, +384
  auto fetch = [&](u32 *st, const u32 *src) {
#ifdef FHESI_EMU
    st[0] = src[0], st[128] = src[128], st[256] = src[256], st[384] = src[384];
#else
    const u32 sa = (u32)__cvta_generic_to_shared(st);
    asm volatile(
        "cp.async.ca.shared.global [%0], [%1], 4;\n\t"
        "cp.async.ca.shared.global [%0 + 512], [%1 + 512], 4;\n\t"
        "cp.async.ca.shared.global [%0 + 1024], [%1 + 1024], 4;\n\t"
        "cp.async.ca.shared.global [%0 + 1536], [%1 + 1536], 4;\n\t"
        "cp.async.commit_group;" ::"r"(sa), "l"(src) : "memory");
#endif
  };
  fetch(stage0, dp);
  u32 tgl = 0;
  for (u32 k = 0; k < a.K; ++k) {
    u32 x[8];
#ifndef FHESI_EMU
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    const u32 *st = (k & 1) ? stage1 : stage0;
    x[0] = st[0], x[1] = st[128], x[2] = st[256] & m2, x[3] = st[384] & m3;
    if (k + 1 < a.K) {
      dp += c.n;
      fetch((k & 1) ? stage0 : stage1, dp);
    }
    fwd1024<true, KSS_DFMA>(x, twf, A, bufA0 + tgl, bufB, g, tg, p, twd, negp, hic, zop);
    tgl ^= KSS_BUFA;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const uint4 k0 = __ldg(kp + h * (FN / 4)), k1 = __ldg(kp + h * (FN / 4) + 1);
      const u32 kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[h][j] += (u64)((i64)(int)x[j] * (int)kv[j]);
    }
    kp += FN;  // 4 * FN words
  }
  fhesi_group_sync(g);  // every warp is done with the forward transforms' buffers
  const u32 *corr = a.key + (size_t)a.Lk * a.K * 4 * FN + (size_t)l * 4 * FN + tg * 8;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const uint4 c0 = __ldg((const uint4 *)(corr + h * FN)), c1 = __ldg((const uint4 *)(corr + h * FN) + 1);
    const u32 cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    u32 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const i64 s0 = (i64)acc[h][j];
      const u32 r0 = csub(mont_red64((u64)(s0 < 0 ? -s0 : s0), p, pinv), p2);
      t[j] = csub((s0 < 0 ? csub(p2 - r0, p2) : r0) + cv[j], p2);
    }
    u32 *bufA = bufA0 + ((h & 1) ? KSS_BUFA : 0u);
    inv1024(t, twi, A, bufA, bufB, bufA, g, tg, p);
    PHIM_STORE_1024(bufA, a.res + ((op * 4 + h) * a.Lk + l) * (size_t)c.n);
  }
}
// Offset correction of the split key switch: the kernel accumulates (x - ch) * K with ch = p >> 1,
// so sum_k x_k K_k = acc + ch * sum_k K_k; the second term depends on the key only.  One thread per
// (prime, half, position): corr = ch * sum_k K_k * R^-1 mod p in [0,p) (the accumulator is Montgomery-
// reduced before corr is added).  key: [Ls][K][4][N] balanced; corr: [Ls][4][N]
__global__ void k_split_corr(DevCtx c, const u32 *key, u32 *corr, u32 K, u32 Ls) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Ls * 4 * c.N) return;
  const u32 e = (u32)(idx % c.N), h = (u32)((idx / c.N) & 3), l = (u32)(idx / ((size_t)4 * c.N));
  const PrimeConst pc = c.pc[l];
  i64 s = 0;
  for (u32 k = 0; k < K; ++k) s += (i64)(int)key[(((size_t)l * K + k) * 4 + h) * c.N + e];
  i64 r = s % (i64)pc.p;
  if (r < 0) r += pc.p;
  corr[idx] = csub(mont_mul(pc.p >> 1, (u32)r, pc.p, pc.pinv), pc.p);
}

// key form [0,p) -> balanced (-p/2, p/2], same layout [L][P][N]
__global__ void k_balance_key(DevCtx c, const u32 *in, u32 *out, u32 P, size_t total) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const u32 l = (u32)(idx / ((size_t)P * c.N));
  const u32 p = c.pc[l].p, v = in[idx];
  out[idx] = v > (p >> 1) ? v - p : v;
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

// Rotation fused into the digit stage (SumBatchedData, Regression.h:166-178: tmp >>= k;
// ApplyKeySwitch(tmp)): in coefficient form a(X) -> a(X^k) mod Phi_m is a signed index permutation
// plus the X^n fold, so the digits of Reduce(a(X^k)) are produced directly from the ciphertext words --
// no wide intermediate, no separate reduction, no transform.  in [npolys][n][W] -> out [npolys][D][n];
// tab as in k_automorph.
__global__ void k_digits_automorph(DevCtx c, const u32 *in, const u32 *tab, u32 *out, size_t npolys) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npolys * c.n) return;
  const size_t poly = idx / c.n;
  const u32 j = (u32)(idx % c.n), W = c.W;
  const u32 *base = in + poly * c.n * W;
  const u32 tj = tab[j], tt = tab[c.n];
  const u32 neg_top = (j & 1) ? 0u : 1u;  // w_j = v_j - (-1)^j v_n
  const bool ha = tj != 0xFFFFFFFFu, hb = tt != 0xFFFFFFFFu;
  const u32 ca = (ha && (tj & 1)) ? 1u : 0u, cb = (hb && ((tt & 1) ^ neg_top)) ? 1u : 0u;
  const u32 *sa = base + (size_t)(tj >> 1) * W, *sb = base + (size_t)(tt >> 1) * W;
  u32 w[17];  // W <= 16 words of the sum; bits above logQ are dropped by the digit extraction
  u32 ba = ca, bb = cb, c2 = 0;
  for (u32 k = 0; k < W; ++k) {
    u32 wa = 0, wb = 0;
    if (ha) {
      u64 t = (u64)(ca ? ~sa[k] : sa[k]) + ba;
      wa = (u32)t, ba = (u32)(t >> 32);
    }
    if (hb) {
      u64 t = (u64)(cb ? ~sb[k] : sb[k]) + bb;
      wb = (u32)t, bb = (u32)(t >> 32);
    }
    u64 t = (u64)wa + wb + c2;
    w[k] = (u32)t, c2 = (u32)(t >> 32);
  }
  for (u32 d = 0; d < c.D; ++d) out[(poly * c.D + d) * c.n + j] = digit_from_words(w, W, c.logQ, c.dbits, d);
}

template <class K>
static int fused_set_smem(K kern, size_t words) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(words * 4)) == cudaSuccess ? 0 : -1;
}
static int fused_configure() {
  return fused_set_smem(k_fused_tensor<false>, FUSED_SMEM_WORDS) || fused_set_smem(k_fused_tensor<true>, FUSED_SMEM_WORDS) ||
         fused_set_smem(k_fused_keyswitch<true, false>, KS_SMEM_WORDS) || fused_set_smem(k_fused_keyswitch<false, false>, KS_SMEM_WORDS) ||
         fused_set_smem(k_fused_keyswitch<true, true>, KS_SMEM_WORDS) || fused_set_smem(k_fused_keyswitch<false, true>, KS_SMEM_WORDS) ||
         fused_set_smem(k_fused_keyswitch_split<false>, KSS_SMEM_WORDS) || fused_set_smem(k_fused_keyswitch_split<true>, KSS_SMEM_WORDS);
}
