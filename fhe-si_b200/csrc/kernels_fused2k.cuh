// kernels_fused2k.cuh -- the fused hot path for N = 2048 (phi(m) in (512, 1024]: the reference README's second
// parameter family, p = 2027, m = 2026, phi(m) = 1012; README:35-37).
//
// Same design as kernels_fused.cuh (N = 1024) with 256 threads per transform, written over the group size T so
// that the index arithmetic is stated once:
//   N = 8 T coefficients, 8 per thread in registers; S2 = T/8, S3 = T/64
//   pass 1  thread t holds i = j*T + t                       (position bits 10,9,8 for T = 256)
//   pass 2  thread t = hi*S2 + lo holds i = hi*T + j*S2 + lo  (bits 7,6,5)
//   pass 3  thread t = u*S3 + b  holds i = u*S2 + j*S3 + b    (bits 4,3,2)
//   then log2(S3) stages across lanes by warp shuffles (bit 1 with a twiddle, bit 0 with twiddle 1)
// Exchange 1 (pass 1 <-> 2) needs no padding at T = 256 (a warp's 32 lanes read 32 consecutive words);
// exchange 2 (pass 2 <-> 3) pads S3 words per S2 positions and stays inside a warp: warp w owns positions
// [T w, T w + T) in both ownerships.  Transform-domain vectors are stored thread-major (thread t, register r
// <-> word 8t + r), which for N = 2048 is store_index() with a 32-position block (DevCtx::sshift = 2).
//
// k_fused_tensor_2k           Ciphertext::operator*= + the per-prime half of ScaleDown (Ciphertext.cpp:167-192)
// k_fused_keyswitch_split_2k  digits -> 3D forward transforms -> four inner products with the split key-switch
//                             matrix -> inverse transforms (FHE-SI.cpp:244-257, Util.h:80-98)
#pragma once
#include "kernels_fused.cuh"

template <int T>
struct FT {
  static constexpr u32 N = 8u * T, S2 = T / 8u, S3 = T / 64u;
  static constexpr u32 P2 = 7u * T, P3 = P2 + 7u * S2, PS = P3 + 7u * S3;  // table offsets: pass 2, pass 3, shuffle stage
  static constexpr u32 ENTRIES = PS + (S3 > 2u ? S3 / 2u : 0u);
  static constexpr u32 WORDS = (ENTRIES * 2u + 3u) & ~3u;
  static constexpr u32 PAD1 = (T == 128 ? 16u : 0u);       // words of padding per T positions, exchange 1
  static constexpr u32 BUFA = N + 8u * PAD1;                // exchange-1 buffer
  static constexpr u32 BUFB = N + (N / S2) * S3;            // exchange-2 buffer
};
// index into the [N] Shoup table (entry h + j of stage h) of entry e of the per-thread layout
template <int T>
__host__ __device__ __forceinline__ u32 ftw_source_index_t(u32 e) {
  typedef FT<T> F;
  u32 s, r, unit;
  if (e < F::P2) s = e / T, r = e % T, unit = T;
  else if (e < F::P3) s = (e - F::P2) / F::S2, r = (e - F::P2) % F::S2, unit = F::S2;
  else if (e < F::PS) s = (e - F::P3) / F::S3, r = (e - F::P3) % F::S3, unit = F::S3;
  else return F::S3 / 2u + (e - F::PS);  // the twiddled shuffle stage, h = S3 / 2
  return s < 4 ? 4 * unit + s * unit + r : (s < 6 ? 2 * unit + (s - 4) * unit + r : unit + r);
}
template <int T>
__device__ __forceinline__ void fill_tw_table_t(uint2 *tws, const uint2 *__restrict__ laid_out) {
  for (u32 e = threadIdx.x; e < FT<T>::ENTRIES; e += blockDim.x) tws[e] = __ldg(laid_out + e);
}
#ifndef FHESI_EMU
template <int T>
__device__ __forceinline__ void fhesi_group_sync_t(unsigned g) {
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
}
#endif

template <int T>
struct XAddrT {
  u32 a1, a2a, a2b, a3;
};
template <int T>
__device__ __forceinline__ XAddrT<T> make_xaddr_t(u32 tg) {
  typedef FT<T> F;
  const u32 hi = tg / F::S2, lo = tg % F::S2, u = tg / F::S3, b = tg % F::S3;
  XAddrT<T> r;
  r.a1 = tg;                                // pass-1 view of bufA:  a1  + (T + PAD1) j
  r.a2a = hi * (T + F::PAD1) + lo;          // pass-2 view of bufA:  a2a + S2 j
  r.a2b = hi * (T + 8 * F::S3) + lo;        // pass-2 view of bufB:  a2b + (S2 + S3) j
  r.a3 = u * (F::S2 + F::S3) + b;           // pass-3 view of bufB:  a3  + S3 j
  return r;
}

// Forward transform of a polynomial whose upper half is zero.  In: x[0..3] = coefficients j*T + tg (< 2p).
// Out: x[r] = value at storage index 8 tg + r, in [0,p) (OFFS: minus p >> 1, see fwd1024).
template <int T, bool OFFS>
__device__ __forceinline__ void fwd_t(u32 *x, const uint2 *twf, const XAddrT<T> &A, u32 *bufA, u32 *bufB, u32 g, u32 tg,
                                      u32 p) {
  typedef FT<T> F;
  const u32 p2 = 2 * p;
  const uint2 *tw = twf + tg;
#pragma unroll
  for (int j = 0; j < 4; ++j) x[j + 4] = mulw(x[j], tw[j * T], p);  // first stage sees (X, 0)
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const uint2 w = tw[(4 + j) * T];
    GSW(x[j], x[j + 2], w);
    GSW(x[j + 4], x[j + 6], w);
  }
  {
    const uint2 w = tw[6 * T];
#pragma unroll
    for (int j = 0; j < 8; j += 2) GSW(x[j], x[j + 1], w);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[A.a1 + (T + F::PAD1) * j] = x[j];
  fhesi_group_sync_t<T>(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[A.a2a + F::S2 * j];
  dif8<(int)F::S2>(x, twf + F::P2 + (tg % F::S2), p);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[A.a2b + (F::S2 + F::S3) * j] = x[j];
  __syncwarp();  // exchange 2 stays inside a warp
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[A.a3 + F::S3 * j];
  dif8<(int)F::S3>(x, twf + F::P3 + (tg % F::S3), p);
  // stages across lanes: position bit k pairs lane l with l ^ (1 << k)
  if (F::S3 > 2) {  // h = 2: twiddle tw[2 + (position & 1)]
    const u32 up = tg & 2;
    const uint2 w = twf[F::PS + (tg & 1)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32 o = __shfl_xor_sync(0xffffffffu, x[j], 2);
      const u32 s_ = csub(add_alu(x[j], o), p2);   // lower lane keeps X + Y
      const u32 d_ = mulw(o + p2 - x[j], w, p);     // upper lane: (X - Y) w, X came from the lower lane
      x[j] = up ? d_ : s_;
    }
  }
  const u32 b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    const u32 v = b0 ? o + p2 - x[j] : add_alu(x[j], o);
    x[j] = csub(csub(v, p2), p);
    if (OFFS) x[j] = sub_alu(x[j], p >> 1);
  }
}
// Inverse transform.  In: x[r] = value at storage index 8 tg + r (< 2p).  Out: the natural-order result
// (unscaled, < 2p) in nat[0..N) (nat may alias bufA); the group is synchronised afterwards.
template <int T>
__device__ __forceinline__ void inv_t(u32 *x, const uint2 *twi, const XAddrT<T> &A, u32 *bufA, u32 *bufB, u32 *nat, u32 g,
                                      u32 tg, u32 p) {
  typedef FT<T> F;
  const u32 p2 = 2 * p;
  const u32 b0 = tg & 1;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const u32 o = __shfl_xor_sync(0xffffffffu, x[j], 1);
    const u32 v = b0 ? o + p2 - x[j] : add_alu(x[j], o);
    x[j] = csub(v, p2);
  }
  if (F::S3 > 2) {  // h = 2, decimation in time: the upper lane multiplies before the exchange
    const u32 up = tg & 2;
    const uint2 w = twi[F::PS + (tg & 1)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const u32 mine = up ? mulw(x[j], w, p) : x[j];
      const u32 o = __shfl_xor_sync(0xffffffffu, mine, 2);
      const u32 v = up ? o + p2 - mine : add_alu(mine, o);
      x[j] = csub(v, p2);
    }
  }
  dit8<(int)F::S3>(x, twi + F::P3 + (tg % F::S3), p);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) bufB[A.a3 + F::S3 * j] = x[j];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufB[A.a2b + (F::S2 + F::S3) * j];
  dit8<(int)F::S2>(x, twi + F::P2 + (tg % F::S2), p);
#pragma unroll
  for (int j = 0; j < 8; ++j) bufA[A.a2a + F::S2 * j] = x[j];
  fhesi_group_sync_t<T>(g);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = bufA[A.a1 + (T + F::PAD1) * j];
  dit8<T>(x, twi + tg, p);
  fhesi_group_sync_t<T>(g);  // nat aliases bufA
#pragma unroll
  for (int j = 0; j < 8; ++j) nat[j * T + tg] = x[j];
  fhesi_group_sync_t<T>(g);
}
// Phi_m fold for m = 2h from the natural-order product in nat[] (values < 2p): n = h - 1 residues to dst
template <int T>
__device__ __forceinline__ void phim_store_t(const u32 *nat, u32 *__restrict__ dst, u32 h, u32 tg, u32 p) {
  const u32 p2 = 2 * p, n = h - 1, N = FT<T>::N;
  u32 top = nat[n];
  if (n + h < N) top = csub(top + p2 - nat[n + h], p2);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const u32 i = j * T + tg;
    if (i < n) {
      u32 v = nat[i];
      if (i + h < N) v = csub(v + p2 - nat[i + h], p2);
      v = (i & 1) ? v + top : v + p2 - top;
      dst[i] = full_reduce(v, p);
    }
  }
}

// ---------------------------------------------------------------------------------------
// tensor product: one T-thread group per (prime, ciphertext pair)
// ---------------------------------------------------------------------------------------
#ifndef KG2
#define KG2 1        // transform groups per CTA (1 x 3 CTAs per SM measured 6 % faster than 3 x 1)
#endif
#ifndef KG2_MINB
#define KG2_MINB 3   // resident CTAs per SM asked of ptxas
#endif
#define T2K 256
#define FUSED2K_SMEM_WORDS (2 * FT<T2K>::WORDS + KG2 * (2 * FT<T2K>::BUFA + FT<T2K>::BUFB))
template <bool GEN>
__global__ void __launch_bounds__(KG2 *T2K, KG2_MINB) k_fused_tensor_2k(DevCtx c, FusedTensorArgs a) {
  typedef FT<T2K> F;
  FHESI_SMEM(sm);
  uint2 *twf = (uint2 *)sm, *twi = (uint2 *)(sm + F::WORDS);
  const u32 g = threadIdx.x / T2K, tg = threadIdx.x % T2K;
  const u32 l = blockIdx.x;
  fill_tw_table_t<T2K>(twf, c.ftw_fwd + (size_t)l * F::ENTRIES);
  fill_tw_table_t<T2K>(twi, c.ftw_inv + (size_t)l * F::ENTRIES);
  __syncthreads();
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u32 *bufA0 = sm + 2 * F::WORDS + g * (2 * F::BUFA + F::BUFB), *bufA1 = bufA0 + F::BUFA, *bufB = bufA1 + F::BUFA;
  u32 flip = 0;
  const XAddrT<T2K> A = make_xaddr_t<T2K>(tg);
  for (u32 it = 0; it < a.ops_per_group; ++it) {
    const size_t op = ((size_t)blockIdx.y * a.ops_per_group + it) * KG2 + g;
    if (op >= a.count) return;  // only group barriers from here on
    u32 Fq[4][8];
    const u32 *rin = a.resin + ((op * 4) * a.Lt + l) * (size_t)c.n;
    const size_t qstride = (size_t)a.Lt * c.n;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int j = 0; j < 4; ++j) Fq[q][j] = (j * T2K + tg < c.n) ? __ldg(rin + q * qstride + j * T2K + tg) : 0u;
      fwd_t<T2K, false>(Fq[q], twf, A, (flip++ & 1) ? bufA1 : bufA0, bufB, g, tg, p);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {  // tProd[k] = sum_{i+j=k} a_i * b_j   (Ciphertext.cpp:179-186)
      u32 y[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (k == 0) y[j] = mont_mul(Fq[0][j], Fq[2][j], p, pinv);
        else if (k == 2) y[j] = mont_mul(Fq[1][j], Fq[3][j], p, pinv);
        else y[j] = csub(mont_mul(Fq[0][j], Fq[3][j], p, pinv) + mont_mul(Fq[1][j], Fq[2][j], p, pinv), p2);
      }
      if (a.to_tprod) {
        u32 *dst = a.res + ((op * 3 + k) * a.Lt + l) * (size_t)F::N + tg * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = csub(y[j], p);
      } else {
        u32 *bufA = (flip++ & 1) ? bufA1 : bufA0;
        inv_t<T2K>(y, twi, A, bufA, bufB, bufA, g, tg, p);
        if (!GEN) phim_store_t<T2K>(bufA, a.res + ((op * 3 + k) * a.Lt + l) * (size_t)c.n, c.h, tg, p);
        else phim_store_csr<T2K>(bufA, a.res + ((op * 3 + k) * a.Lt + l) * (size_t)c.n, c.red, c.n, tg, p);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// split-key key switch (see k_fused_keyswitch_split): key [Ls][K][4][N] balanced + correction table [Ls][4][N];
// res [count][4][Ls][n] in the order b_lo, b_hi, A_lo, A_hi
// ---------------------------------------------------------------------------------------
#ifndef KSS2
#define KSS2 1       // transform groups (ciphertexts) per CTA (1 x 2 CTAs per SM measured 10 % faster than 2 x 1)
#endif
#ifndef KSS2_MINB
#define KSS2_MINB 2
#endif
#define KSS2K_SMEM_WORDS (2 * FT<T2K>::WORDS + KSS2 * (2 * FT<T2K>::BUFA + FT<T2K>::BUFB))
template <bool GEN>
__global__ void __launch_bounds__(KSS2 *T2K, KSS2_MINB) k_fused_keyswitch_split_2k(DevCtx c, FusedKsArgs a) {
  typedef FT<T2K> F;
  FHESI_SMEM(sm);
  uint2 *twf = (uint2 *)sm, *twi = (uint2 *)(sm + F::WORDS);
  const u32 g = threadIdx.x / T2K, tg = threadIdx.x % T2K;
  const u32 l = blockIdx.x;
  fill_tw_table_t<T2K>(twf, c.ftw_fwd + (size_t)l * F::ENTRIES);
  fill_tw_table_t<T2K>(twi, c.ftw_inv + (size_t)l * F::ENTRIES);
  __syncthreads();
  const size_t op = (size_t)blockIdx.y * KSS2 + g;
  if (op >= a.count) return;
  const PrimeConst pc = c.pc[l];
  const u32 p = pc.p, pinv = pc.pinv, p2 = 2 * p;
  u32 *bufA0 = sm + 2 * F::WORDS + g * (2 * F::BUFA + F::BUFB), *bufA1 = bufA0 + F::BUFA, *bufB = bufA1 + F::BUFA;
  const XAddrT<T2K> A = make_xaddr_t<T2K>(tg);
  u64 acc[4][8];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[h][j] = 0;
  const u32 *dig = a.digits + op * a.K * (size_t)c.n;
  const uint4 *kp = (const uint4 *)(a.key + (size_t)l * a.K * 4 * F::N + tg * 8);
  u32 xn[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) xn[j] = (j * T2K + tg < c.n) ? __ldg(dig + j * T2K + tg) : 0u;
  for (u32 k = 0; k < a.K; ++k) {
    u32 x[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = xn[j];
    if (k + 1 < a.K) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        xn[j] = (j * T2K + tg < c.n) ? __ldg(dig + (size_t)(k + 1) * c.n + j * T2K + tg) : 0u;
    }
    fwd_t<T2K, true>(x, twf, A, (k & 1) ? bufA1 : bufA0, bufB, g, tg, p);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const uint4 k0 = __ldg(kp + h * (F::N / 4)), k1 = __ldg(kp + h * (F::N / 4) + 1);
      const u32 kv[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[h][j] += (u64)((i64)(int)x[j] * (int)kv[j]);
    }
    kp += F::N;
  }
  fhesi_group_sync_t<T2K>(g);  // every warp is done with the forward transforms' buffers
  const u32 *corr = a.key + (size_t)a.Lk * a.K * 4 * F::N + (size_t)l * 4 * F::N + tg * 8;
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const uint4 c0 = __ldg((const uint4 *)(corr + h * F::N)), c1 = __ldg((const uint4 *)(corr + h * F::N) + 1);
    const u32 cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    u32 t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const i64 s0 = (i64)acc[h][j];
      const u32 r0 = csub(mont_red64((u64)(s0 < 0 ? -s0 : s0), p, pinv), p2);
      t[j] = csub((s0 < 0 ? csub(p2 - r0, p2) : r0) + cv[j], p2);
    }
    u32 *bufA = (h & 1) ? bufA1 : bufA0;
    inv_t<T2K>(t, twi, A, bufA, bufB, bufA, g, tg, p);
    if (!GEN) phim_store_t<T2K>(bufA, a.res + ((op * 4 + h) * a.Lk + l) * (size_t)c.n, c.h, tg, p);
        else phim_store_csr<T2K>(bufA, a.res + ((op * 4 + h) * a.Lk + l) * (size_t)c.n, c.red, c.n, tg, p);
  }
}

static int fused2k_configure() {
  return fused_set_smem(k_fused_tensor_2k<false>, FUSED2K_SMEM_WORDS) || fused_set_smem(k_fused_tensor_2k<true>, FUSED2K_SMEM_WORDS) ||
         fused_set_smem(k_fused_keyswitch_split_2k<false>, KSS2K_SMEM_WORDS) ||
         fused_set_smem(k_fused_keyswitch_split_2k<true>, KSS2K_SMEM_WORDS);
}
