/*
 * fhesi.h -- C ABI of libfhesi_b200.so: the B200-native (sm_100a) implementation of the
 * FHE-SI ciphertext-arithmetic hot path (dwu4/fhe-si).
 *
 * The reference has no FFI: its boundary is the C++ class surface in FHEContext.h,
 * FHE-SI.h, Ciphertext.h, DoubleCRT.h.  This header is the layer that sits
 * *underneath* a re-implementation of those classes (fhe-si_b200/host/).  Every entry
 * point cites the reference function(s) whose work it performs.
 *
 * Conventions
 *   - every call returns 0 on success or a negative FHESI_ERR_* code; nothing throws;
 *     fhesi_last_error() returns a thread-local message for the last failure;
 *   - `*_dev` entry points take DEVICE pointers (cudaMalloc'd / torch tensors' data_ptr)
 *     and enqueue kernels on the context's stream -- they are asynchronous;
 *     `*_host` entry points take HOST pointers, copy in, compute, copy out and block;
 *   - there is no CPU fallback: if no CUDA device is usable, fhesi_ctx_create fails.
 *
 * Data formats (all little-endian uint32 words)
 *   poly   coefficient-domain polynomial: [n][W] words, W = ceil(logQ/32); coefficient i is
 *          the two's complement of the centred value in [-q/2, q/2), sign-extended to 32*W
 *          bits (this is CiphertextPart::poly after ReduceCoefficients, Util.cpp:28-33).
 *   ct     ciphertext with `parts` polys: [parts][n][W]; a batch is [count][parts][n][W].
 *   tprod  tensor-form ("scaledUp") ciphertext, the device image of the reference's
 *          vector<DoubleCRT> tProd (Ciphertext.h:51): [parts][Lt][N] residues in [0,p_i),
 *          in the library's own transform domain: N-point cyclic NTT of the
 *          zero-padded polynomial, scaled by N^-1, over the library's own 30-bit primes.
 *          By SURVEY.md §0.3 coefficient-domain results do not depend on this choice.
 *   rows   reference-chain DoubleCRT rows (DoubleCRT.h:83-365): int64 [L][n], row i column j
 *          = poly(zeta_i^{u_j}) mod q_i -- only used for key import/export parity.
 */
#ifndef FHESI_H_
#define FHESI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHESI_OK 0
#define FHESI_ERR_INVALID (-1)      /* bad argument                                       */
#define FHESI_ERR_UNSUPPORTED (-2)  /* parameter set outside what the kernels implement   */
#define FHESI_ERR_CUDA (-3)         /* CUDA runtime error (message has the cudaError_t)   */
#define FHESI_ERR_NOMEM (-4)

#define FHESI_MAX_PRIMES 40

typedef struct fhesi_ctx fhesi_ctx; /* FHEcontext image on one GPU (FHEContext.h:45-206) */
typedef struct fhesi_ksw fhesi_ksw; /* KeySwitchSI::keySwitchMatrix in HBM (FHE-SI.h:116) */
typedef struct fhesi_key fhesi_key; /* vector<DoubleCRT> key (pk or sk) in HBM            */

typedef struct fhesi_info {
  uint32_t m, n, logQ, W, decompSize, D; /* n = phi(m), D = ndigits (FHEContext.h:115) */
  uint32_t N;                            /* transform length, power of two >= 2n-1      */
  uint32_t Lt, Lk, Le;                   /* primes used by tensor / key-switch / enc-dec */
  uint64_t p, xi;
  uint32_t primes[FHESI_MAX_PRIMES];     /* the library's own chain (30-bit, = 1 mod N)  */
  int device;
  uint32_t Ls, split_words;              /* split-key key switch: primes used, words in the low half
                                            (0 = not used); see DESIGN.md "split keys"          */
} fhesi_info;

const char *fhesi_last_error(void);
const char *fhesi_version(void);

/* ---- context: FHEcontext::Init + SetUpSIContext (FHEContext.h:105-118, FHEContext.cpp:83-85).
 * Any m in [3, 8192] with 2 <= phi(m) <= 1024, as the reference's Bluestein transform serves any m
 * (bluestein.cpp:93-144, CModulus.cpp:110-132).  m = 2*p' with p' an odd prime (every parameter set the
 * reference's clients use, README:35-37) has the remainder by Phi_m written into the kernels
 * as a fold; every other m takes it from a sparse table built here (X^j mod Phi_m must have coefficients
 * within +-64, true of every m in range).  xi is the number of tensor products that may be summed in
 * tprod form before ScaleDown (SetUpSIContext's argument). */
int fhesi_ctx_create(uint32_t m, uint32_t logQ, uint64_t p, uint32_t decompSize, uint64_t xi,
                     int device, fhesi_ctx **out);
void fhesi_ctx_destroy(fhesi_ctx *ctx);
/* Device memory of a destroyed context (tables, scratch, pooled buffers) is kept in a per-device cache for
 * the next context instead of going back to the driver (cudaFree synchronises the whole device; a client that
 * builds one context per request would pay for it on every request).  This returns the cache to the driver. */
int fhesi_trim_cache(int device);
int fhesi_ctx_info(const fhesi_ctx *ctx, fhesi_info *out);
/* Use an externally owned cudaStream_t (e.g. torch's current stream); NULL = own stream. */
int fhesi_ctx_set_stream(fhesi_ctx *ctx, void *cuda_stream);
int fhesi_sync(fhesi_ctx *ctx);

/* ---- device memory helpers (so host code above the ABI needs no CUDA headers) */
int fhesi_malloc(fhesi_ctx *ctx, size_t bytes, void **dptr);
int fhesi_free(fhesi_ctx *ctx, void *dptr);
int fhesi_h2d(fhesi_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
/* Stream-ordered copy without the trailing synchronisation.  For pageable `src_host` the source has
 * been consumed when the call returns; a pinned source must stay untouched until fhesi_sync.  The
 * host layer uses it so that a client's next operator (sampling, packing) overlaps the device
 * work of the previous one. */
int fhesi_h2d_async(fhesi_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int fhesi_d2h(fhesi_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);
/* Page-locked host memory for the buffers handed to the *_host entry points (new: the reference has no device).
 * write_combined != 0 is for buffers the CPU only WRITES (operands on their way up): the device reads them
 * without snooping the CPU caches.  Measured with all eight GPUs of a box copying at once: the slowest ranks move
 * a step's bytes in 47.7 ms instead of 52.0 ms (profiles/r02c_pcie_probe_8gpu.txt); CPU reads of such memory are
 * slow, so results belong in ordinary page-locked memory (write_combined = 0). */
/* ScaleDown (Ciphertext.cpp:194-218) reads the centred integer off a windowed explicit CRT sum and falls back, per
 * coefficient, on the full mixed-radix reconstruction when it cannot prove the window exact (kernels_generic.cuh,
 * k_crt_direct).  This is the number of coefficients that took the fallback since the context was created
 * (expected: about two per million). */
int fhesi_crt_fallbacks(fhesi_ctx *ctx, uint64_t *count);
int fhesi_host_alloc(size_t bytes, int write_combined, void **out);
int fhesi_host_free(void *p);
int fhesi_d2d(fhesi_ctx *ctx, void *dst_dev, const void *src_dev, size_t bytes);
size_t fhesi_ct_bytes(const fhesi_ctx *ctx, uint32_t parts);    /* parts*n*W*4      */
size_t fhesi_tprod_bytes(const fhesi_ctx *ctx, uint32_t parts); /* parts*Lt*N*4     */

/* ---- keys.  Inputs are HOST coefficient polys in `poly` format.
 * fhesi_ksw_create: uploads keySwitchMatrix (FHE-SI.cpp:205-208): b and A are
 *   [src_parts*D][n][W] each, in the reference's order (part-major, digit-minor,
 *   FHE-SI.cpp:171-172).  A may hold +q/2 (FHE-SI.cpp:178-180 negates an unreduced poly);
 *   only its value mod q matters.
 * fhesi_key_create: a vector<DoubleCRT> key of `parts` polys (publicKey FHE-SI.cpp:59-61,
 *   sKeys FHE-SI.cpp:86-91). */
int fhesi_ksw_create(fhesi_ctx *ctx, const uint32_t *h_b, const uint32_t *h_A, uint32_t src_parts,
                     fhesi_ksw **out);
/* KeySwitchSI::Init (FHE-SI.cpp:153-209) on the device from the caller's random draws (the order and
 * number of draws are the caller's business).  For entry (i, j), i < parts, j < D:
 *   b = A * t + e + src_i * 2^(8 decompSize j)  reduced mod q,   A' = Reduce(-A).
 * h_src: HOST int32 [parts][n], the source key polynomials (1, s, s^2 or 1, s(X^k): small integers);
 * h_t: int32 [n], the target key s; h_A: uint32 [parts*D][n][W], the SampleRandom polynomials (centred,
 * two's complement); h_e: int32 [parts*D][n], the Gaussians.  The matrix is left on the device as a
 * fhesi_ksw; h_b_out / h_A_out (nullable, [parts*D][n][W]) receive b and A' for callers that also
 * need them on the host (Export). */
int fhesi_ksw_generate(fhesi_ctx *ctx, const int32_t *h_src, const int32_t *h_t, const uint32_t *h_A,
                       const int32_t *h_e, uint32_t parts, fhesi_ksw **out, uint32_t *h_b_out,
                       uint32_t *h_A_out);
/* The whole set-up of a client in one pass of kernels (Regression.h:68-81: `secretKey, publicKey(secretKey),
 * keySwitch(secretKey)` and one KeySwitchSI(secretKey, k) per rotation): M key-switch matrices, matrix m with
 * parts[m] source polynomials, and -- when pk_out is non-NULL -- FHESIPubKey::Init (FHE-SI.cpp:42-62), which is
 * one more entry of the same shape with no source term: c0 = e + s * c1 reduced, c1' = Reduce(-c1).
 * Draws are concatenated in matrix order: h_src int32 [sum parts][n]; h_A uint32 [sum parts*D (+1)][n][W]; h_e
 * int32 [sum parts*D (+1)][n]; the public key's c1 and e are the LAST entry of h_A / h_e.  h_t: the secret key s.
 * out[M] receives the matrices; h_b_out / h_A_out (nullable, [sum parts*D][n][W]) b and A'; h_pk_out (nullable,
 * [2][n][W]) the public key polynomials.  M = 0 with pk_out generates only the public key. */
int fhesi_keygen_batch(fhesi_ctx *ctx, uint32_t M, const uint32_t *parts, const int32_t *h_src, const int32_t *h_t,
                       const uint32_t *h_A, const int32_t *h_e, fhesi_ksw **out, uint32_t *h_b_out,
                       uint32_t *h_A_out, fhesi_key **pk_out, uint32_t *h_pk_out);
void fhesi_ksw_destroy(fhesi_ksw *ksw);
int fhesi_key_create(fhesi_ctx *ctx, const uint32_t *h_polys, uint32_t parts, fhesi_key **out);
void fhesi_key_destroy(fhesi_key *key);

/* ---- the metric op, batched: for each i, c = a[i]; c *= b[i]; ks.ApplyKeySwitch(c)
 * (Ciphertext.cpp:167-192 + FHE-SI.cpp:241-260; Test_AddMul.cpp:59-66).
 * a, b: [count][2][n][W]; out: [count][2][n][W]. */
int fhesi_mult_relin_dev(fhesi_ctx *ctx, const fhesi_ksw *ksw, const uint32_t *d_a,
                         const uint32_t *d_b, uint32_t *d_out, size_t count);
int fhesi_mult_relin_host(fhesi_ctx *ctx, const fhesi_ksw *ksw, const uint32_t *h_a,
                          const uint32_t *h_b, uint32_t *h_out, size_t count);
/* The same, returning as soon as the batch is enqueued: a server that feeds batch after batch (Matrix.cpp's
 * products over a stream of blocks) overlaps the next batch's uploads and kernels with this batch's last kernels
 * and downloads.  h_a / h_b must stay untouched and h_out unread until fhesi_sync_all returns; calls in flight at
 * the same time need an h_out each.  fhesi_sync_all waits for the context's stream AND the pipeline's copy streams
 * (fhesi_sync waits for the context's stream only). */
int fhesi_mult_relin_host_async(fhesi_ctx *ctx, const fhesi_ksw *ksw, const uint32_t *h_a,
                                const uint32_t *h_b, uint32_t *h_out, size_t count);
int fhesi_sync_all(fhesi_ctx *ctx);

/* ---- the pieces, each batched over `count` independent ciphertexts -------------------- */

/* Ciphertext::operator+=(const Ciphertext&), !scaledUp branch (Ciphertext.cpp:126-134):
 * io[i] = Reduce(io[i] + other[i]); both [count][parts][n][W]. */
int fhesi_ct_add_dev(fhesi_ctx *ctx, uint32_t *d_io, const uint32_t *d_other, uint32_t parts,
                     size_t count);
/* Same, accumulating a batch into one ciphertext: out = Reduce(sum_i in[i]) -- the
 * Matrix<Ciphertext> row sums (Matrix.cpp:80-97) in coefficient form. */
int fhesi_ct_sum_dev(fhesi_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint32_t parts,
                     size_t count);
/* CiphertextPart::operator*=(long) (Ciphertext.cpp:21-27): io = Reduce(io * l). */
int fhesi_ct_mul_scalar_dev(fhesi_ctx *ctx, uint32_t *d_io, int64_t l, uint32_t parts, size_t count);

/* CiphertextPart::operator*=(const ZZX&) (Ciphertext.cpp:29-36, 246-250): every part of io times
 * the same plaintext polynomial mod Phi_m, then Reduce.  plain: DEVICE uint32 [n], coefficients
 * in [0, 2^29) (to_ZZX of a ZZ_pX).  io: [count][parts][n][W]. */
int fhesi_ct_mul_plain_dev(fhesi_ctx *ctx, uint32_t *d_io, const uint32_t *d_plain, uint32_t parts,
                           size_t count);

/* Ciphertext::operator*=(const Ciphertext&) (Ciphertext.cpp:167-192): tensor product into
 * tprod form.  a: [count][pa][n][W], b: [count][pb][n][W], out: [count][pa+pb-1][Lt][N].
 * If accumulate != 0, out[0] (a single tprod) receives the SUM over the batch instead
 * (the data-phase sums of Regression.h:106-108 / Matrix.cpp:80-97,149-173). */
int fhesi_ct_tensor_dev(fhesi_ctx *ctx, const uint32_t *d_a, uint32_t pa, const uint32_t *d_b,
                        uint32_t pb, uint32_t *d_tprod, size_t count, int accumulate);
/* Ciphertext::operator+=, scaledUp branch (Ciphertext.cpp:135-143): io += other, mod p_i. */
int fhesi_tprod_add_dev(fhesi_ctx *ctx, uint32_t *d_io, const uint32_t *d_other, uint32_t parts,
                        size_t count);
/* Ciphertext::operator*=(long), scaledUp branch (Ciphertext.cpp:239-242). */
int fhesi_tprod_mul_scalar_dev(fhesi_ctx *ctx, uint32_t *d_io, int64_t l, uint32_t parts,
                               size_t count);
/* Ciphertext::ScaleDown (Ciphertext.cpp:194-218): toPoly, round(x/q), Reduce.
 * tprod: [count][parts][Lt][N] -> out: [count][parts][n][W]. */
int fhesi_scaledown_dev(fhesi_ctx *ctx, const uint32_t *d_tprod, uint32_t parts, uint32_t *d_out,
                        size_t count);
/* KeySwitchSI::ApplyKeySwitch after ScaleDown (FHE-SI.cpp:244-259): ByteDecomp, 3D forward
 * transforms, two inner products, toPoly, Reduce.  in: [count][parts][n][W] with
 * parts == ksw's src_parts; out: [count][2][n][W]. */
int fhesi_keyswitch_dev(fhesi_ctx *ctx, const fhesi_ksw *ksw, const uint32_t *d_in,
                        uint32_t *d_out, size_t count);

/* FHESIPubKey::Encrypt (FHE-SI.cpp:10-36) with explicit randomness:
 * msg: uint32 [count][n] coefficients in [0,p) (to_ZZX(ptxt.message));
 * r:   uint8  [count][n] bits (FHE-SI.cpp:14-17);
 * e:   int32  [count][2][n] rounded Gaussians (FHE-SI.cpp:24);
 * out: [count][2][n][W]. */
int fhesi_encrypt_dev(fhesi_ctx *ctx, const fhesi_key *pk, const uint32_t *d_msg,
                      const uint8_t *d_r, const int32_t *d_e, uint32_t *d_out, size_t count);
/* FHESISecKey::Decrypt (FHE-SI.cpp:93-119): in [count][parts>=2][n][W] (parts 0,1 used),
 * out: uint32 [count][n] message coefficients in [0,p). */
int fhesi_decrypt_dev(fhesi_ctx *ctx, const fhesi_key *sk, const uint32_t *d_in, uint32_t parts,
                      uint32_t *d_msg, size_t count);

/* SumBatchedData's step `tmp >>= k; autoKeySwitch.ApplyKeySwitch(tmp)` (Regression.h:166-178,
 * Statistics.h:146-158) in one call: in [count][2][n][W] reduced ciphertexts, ksw the (1, s(X^k)) -> s
 * matrix (KeySwitchSI(sk, k)), out [count][2][n][W].  Same result as fhesi_ct_automorph_dev +
 * fhesi_reduce_wide_dev + fhesi_keyswitch_dev; the rotation is folded into the digit extraction. */
int fhesi_rotate_keyswitch_dev(fhesi_ctx *ctx, const fhesi_ksw *ksw, const uint32_t *d_in, uint32_t k,
                               uint32_t *d_out, size_t count);

/* Tensor-form (scaledUp) branches of the plaintext and automorphism operators.  tprod buffers are
 * [count][parts][Lt][N] as fhesi_ct_tensor_dev writes them.  As in the reference, nothing here reduces
 * modulo q, and results are only meaningful while the represented integers stay inside the prime
 * chain's range (the reference wraps modulo ITS chain product beyond that).
 *  - Ciphertext::operator+=(const ZZX&) (Ciphertext.cpp:153-159): tProd[0] += DoubleCRT(poly), poly =
 *    the caller's floor(c * 2^logQ / p) polynomial, DEVICE [count][n][Win] two's-complement words,
 *    Win <= W + 1 (use W + 1 with a zero top word for non-negative values of logQ bits).
 *  - Ciphertext::operator*=(const ZZX&) (Ciphertext.cpp:252-256): every part of every tprod times
 *    DoubleCRT(poly), ONE polynomial [n][Win] shared by the batch.
 *  - Ciphertext::operator>>=(long k) (Ciphertext.cpp:269-273; DoubleCRT::automorph,
 *    DoubleCRT.cpp:439-465): a(X) -> a(X^k) mod Phi_m on every part; out may not alias in. */
int fhesi_tprod_add_poly_dev(fhesi_ctx *ctx, uint32_t *d_tprod, uint32_t parts, const uint32_t *d_poly,
                             uint32_t Win, size_t count);
int fhesi_tprod_mul_poly_dev(fhesi_ctx *ctx, uint32_t *d_tprod, uint32_t parts, const uint32_t *d_poly,
                             uint32_t Win, size_t count);
int fhesi_tprod_automorph_dev(fhesi_ctx *ctx, const uint32_t *d_in, uint32_t parts, uint32_t k,
                              uint32_t *d_out, size_t count);

/* PlaintextSpace::EmbedInSlots (PlaintextSpace.cpp:112-134) for a batch of plaintexts, as BatchData
 * issues it (Regression.h:43-66, Test_Statistics.cpp:35-63): msg[c] = sum_k vals[c][k] * basis[k] mod p.
 * basis: DEVICE uint32 [nslots][n], row k = the CRT idempotent of slot k (values in [0,p)); vals:
 * DEVICE uint32 [count][nslots] slot values in [0,p); msg: [count][n] coefficients in [0,p), the layout
 * fhesi_encrypt_dev takes.  Needs p < 2^26 and nslots <= 4096. */
int fhesi_embed_slots_dev(fhesi_ctx *ctx, const uint32_t *d_basis, uint32_t nslots, const uint32_t *d_vals,
                          uint32_t *d_msg, size_t count);
/* PlaintextSpace::DecodeSlots (PlaintextSpace.cpp:136-146) for a batch: vals[c][k] = msg[c](root_k) mod p, the
 * value of the message polynomial at the root of slot k -- the same integer matrix product with the roles
 * swapped.  vander: DEVICE uint32 [n][nslots], vander[j][k] = root_k^j mod p; msg: DEVICE uint32 [count][n]
 * (what fhesi_decrypt_dev writes); vals: [count][nslots].  Needs p < 2^26 and n <= 4096. */
int fhesi_decode_slots_dev(fhesi_ctx *ctx, const uint32_t *d_vander, uint32_t nslots, const uint32_t *d_msg,
                           uint32_t *d_vals, size_t count);

/* CiphertextPart::operator>>=(long k) (Ciphertext.cpp:54-59; DoubleCRT::automorph,
 * DoubleCRT.cpp:439-465): a(X) -> a(X^k) mod Phi_m.  The reference does NOT reduce the
 * result mod q; out is therefore [count][parts][n][W+1] words (one word of head-room). */
int fhesi_ct_automorph_dev(fhesi_ctx *ctx, const uint32_t *d_in, uint32_t parts, uint32_t k,
                           uint32_t *d_out_wide, size_t count);
/* Reduce (Util.cpp:3-26) of a [count][parts][n][Win] wide poly into `poly` format. */
int fhesi_reduce_wide_dev(fhesi_ctx *ctx, const uint32_t *d_in_wide, uint32_t Win, uint32_t *d_out,
                          uint32_t parts, size_t count);

/* Cmodulus::FFT / DoubleCRT(const ZZX&) on the REFERENCE chain (CModulus.cpp:90-107,
 * DoubleCRT.cpp:244-257), for key export parity: h_poly [n][Win] two's-complement words ->
 * h_rows int64 [L][n].  primes/roots: the reference chain and its 2m-th roots. */
int fhesi_ref_rows_host(fhesi_ctx *ctx, const uint32_t *h_poly, uint32_t Win, const uint64_t *primes,
                        const uint64_t *roots, uint32_t L, int64_t *h_rows);

/* ---- multi-GPU combine: all ranks' partial tprod sums were all-gathered
 * into d_gathered [world][parts][Lt][N]; out = sum over world, mod p_i. */
int fhesi_tprod_reduce_gathered_dev(fhesi_ctx *ctx, const uint32_t *d_gathered, uint32_t world,
                                    uint32_t parts, uint32_t *d_out);

/* ---- measurement helpers (bench.py): register-resident Montgomery-multiply peak, in
 * modmul/s, for word size 32 or 64 (SURVEY.md §8d "modmul_peak"). */
int fhesi_modmul_peak(fhesi_ctx *ctx, int word_bits, double *modmul_per_s);
/* Single-instruction-class throughput, ops/s over the whole GPU: kind 0 = 32x32->lo32
 * multiply-add, 1 = 32x32->64 multiply-add, 2 = hi32 multiply, 3 = Shoup modular product
 * (the butterfly multiply of the fused kernels), 4 = conditional subtract (ALU), 5 = fp64 FMA. */
int fhesi_pipe_peak(fhesi_ctx *ctx, int kind, double *ops_per_s);

/* Launch accounting and a per-kernel CUDA-event profiler: with profiling on, every kernel
 * launch is bracketed by events on the context's stream.  fhesi_profile_report writes one
 * line per kernel: "<name> <launches> <total_ms>\n".  fhesi_profile_enable resets both. */
int fhesi_profile_enable(fhesi_ctx *ctx, int on);
int fhesi_profile_launches(fhesi_ctx *ctx, uint64_t *launches);
int fhesi_profile_report(fhesi_ctx *ctx, char *buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* FHESI_H_ */
