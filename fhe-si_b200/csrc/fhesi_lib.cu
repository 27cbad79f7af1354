// fhesi_lib.cu -- context set-up and the C ABI of libfhesi_b200.so (include/fhesi.h).
// Built for sm_100a only; there is no CPU fallback anywhere in this library.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <string>
#include <vector>

#include "../../include/fhesi.h"
#include "kernels_generic.cuh"
#include "kernels_fused.cuh"
#include "kernels_fused2k.cuh"

// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(FHESI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));  \
  } while (0)
#define CKL()                                                                            \
  do {                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                 \
    if (e_ != cudaSuccess)                                                               \
      return fail(FHESI_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

struct ProfRec {
  int id;
  cudaEvent_t e0, e1;
};
struct Arena {
  void *ptr = nullptr;
  size_t cap = 0;
};
// a temporary device allocation released on every exit path (cudaFree synchronises implicitly)
struct DevTmp {
  void *p = nullptr;
  ~DevTmp() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
  u32 *u() const { return (u32 *)p; }
  void *release() {
    void *r = p;
    p = nullptr;
    return r;
  }
};

// ---------------------------------------------------------------------------------------
// Per-device cache of device memory released by destroyed contexts.  A client that builds a context per
// request (one regression, one statistics job) would otherwise pay cudaMalloc for every table, scratch arena
// and pooled buffer again -- and cudaFree, a device-wide synchronisation, for every block of the context it
// just dropped; on a busy box those calls stall for tens of milliseconds.  Blocks enter the cache only from
// fhesi_ctx_destroy (after its cudaDeviceSynchronize: nothing in flight can still touch them) and leave it
// for a new owner of at least the same size; fhesi_trim_cache gives everything back to the driver.
// ---------------------------------------------------------------------------------------
#include <mutex>
struct DeviceCache {
  std::mutex mu;
  std::multimap<size_t, void *> blocks;
  size_t bytes = 0;
};
static DeviceCache &device_cache(int dev) {
  static std::mutex mu;
  static std::map<int, DeviceCache> caches;
  std::lock_guard<std::mutex> lk(mu);
  return caches[dev];
}
static const size_t kCacheCapBytes = (size_t)24 << 30;
// a cached block of at least `bytes` and at most 2x + 1 MiB of it, or nullptr; *got = its real size
static void *cache_take(int dev, size_t bytes, size_t *got) {
  DeviceCache &dc = device_cache(dev);
  std::lock_guard<std::mutex> lk(dc.mu);
  auto it = dc.blocks.lower_bound(bytes);
  if (it == dc.blocks.end() || it->first > 2 * bytes + ((size_t)1 << 20)) return nullptr;
  void *p = it->second;
  *got = it->first;
  dc.bytes -= it->first;
  dc.blocks.erase(it);
  return p;
}
static void cache_put(int dev, void *p, size_t bytes) {
  if (!p) return;
  DeviceCache &dc = device_cache(dev);
  std::lock_guard<std::mutex> lk(dc.mu);
  if (dc.bytes + bytes > kCacheCapBytes) {
    cudaFree(p);
    return;
  }
  dc.blocks.emplace(bytes, p);
  dc.bytes += bytes;
}
// cudaMalloc through the cache: *cap receives the real size of the block
static cudaError_t cached_malloc(int dev, void **p, size_t bytes, size_t *cap) {
  size_t got = 0;
  if ((*p = cache_take(dev, bytes, &got))) {
    *cap = got;
    return cudaSuccess;
  }
  *cap = bytes;
  return cudaMalloc(p, bytes);
}
int fhesi_trim_cache(int device) {
  int cur = 0;
  cudaGetDevice(&cur);
  if (cudaSetDevice(device) != cudaSuccess) return FHESI_ERR_CUDA;
  DeviceCache &dc = device_cache(device);
  std::lock_guard<std::mutex> lk(dc.mu);
  cudaDeviceSynchronize();
  for (auto &kv : dc.blocks) cudaFree(kv.second);
  dc.blocks.clear();
  dc.bytes = 0;
  cudaSetDevice(cur);
  return 0;
}

struct fhesi_ctx {
  fhesi_info info{};
  DevCtx dc{};
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::vector<std::pair<void *, size_t>> tables;  // device allocations owned by the context (pointer, bytes)
  Arena scratch;
  std::vector<PrimeConst> h_pc;
  std::vector<u32> h_garner, h_Pfull, h_Phalf;  // host copies for the by-value CRT tables
  std::map<std::pair<int, u32>, std::vector<unsigned char>> crt_tables;
  std::map<u32, u32 *> automorph_tabs;  // Galois element -> device permutation table
  // general m (not 2 * odd prime): column j lists the non-zero (i, coefficient) of X^(n+j) mod Phi_m,
  // j < max(n - 1, m - n); empty for m = 2h
  std::vector<std::vector<std::pair<u32, int>>> red_cols;
  // fhesi_malloc / fhesi_free pool
  std::map<size_t, std::vector<void *>> pool_free;
  std::map<void *, size_t> pool_size;
  size_t pool_idle_bytes = 0, pool_cap_bytes = (size_t)8 << 30;
  u32 chunk = 128;        // ciphertexts per pass through the scratch arena (generic path)
  u32 fused_chunk = 8192; // same for the fused path: large, so the grid is many waves deep
  bool use_fused = true;
  bool crt_direct = true;        // ScaleDown through k_crt_direct (FHESI_NO_CRT_DIRECT=1: k_crt)
  bool crt_force_exact = false;  // FHESI_CRT_FORCE_EXACT=1: every thread of k_crt_direct takes the exact routine
  unsigned long long *d_crt_fallbacks = nullptr;  // device counter of k_crt_direct threads that were not certain
  bool tfree = false;  // every prime satisfies 3D (p/2)^2 < 2^63 (single-accumulator key switch)
  // launch accounting / per-kernel CUDA-event profiler (bench.py "roofline", "gpu_launches")
  uint64_t launches = 0;
  bool prof_on = false;
  std::vector<std::string> prof_names;
  std::vector<ProfRec> prof_recs;
  Arena stage;  // device staging for the *_host entry points
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // copy streams of the host pipeline
  std::vector<cudaEvent_t> pipe_events;
  // fhesi_mult_relin_host_async: two staging halves used alternately; half_done[h] = everything of the last call that
  // used half h has finished (recorded on the download stream after its last copy)
  cudaEvent_t half_done[2] = {nullptr, nullptr};
  bool half_used[2] = {false, false};
  unsigned host_calls = 0;
  u32 pipe_chunk = 0;  // 0 = choose from the batch size; FHESI_PIPE_CHUNK overrides
  u32 pipe_taper = 1;  // FHESI_PIPE_TAPER=0: equal chunks
  int sm_count = 148;  // from the device at context creation
  // second compute lane of the host pipeline: alternate chunks run on their own stream with their
  // own scratch, so one chunk's first waves fill the SMs that the other chunk's last wave leaves idle
  u32 pipe_lanes = 2;  // FHESI_PIPE_LANES=1: single compute stream
  cudaStream_t lane_stream = nullptr;
  Arena lane_scratch;
  Arena work;   // tprod / scaled-down intermediates of the generic mult_relin composition
};
// scratch for set-up paths (key upload): blocks come from the context's pool (fhesi_malloc /
// fhesi_free), so repeated key creation costs no cudaMalloc / cudaFree and no implicit device sync
struct PoolTmp {
  fhesi_ctx *c;
  void *p = nullptr;
  explicit PoolTmp(fhesi_ctx *ctx) : c(ctx) {}
  ~PoolTmp() {
    if (p) fhesi_free(c, p);
  }
  int alloc(size_t bytes) { return fhesi_malloc(c, bytes, &p); }
  u32 *u() const { return (u32 *)p; }
  void *release() {
    void *r = p;
    p = nullptr;
    return r;
  }
};
static void prof_clear(fhesi_ctx *c);
static void prof_begin(fhesi_ctx *c, const char *name) {
  c->launches++;
  if (!c->prof_on) return;
  int id = -1;
  for (size_t i = 0; i < c->prof_names.size(); ++i)
    if (c->prof_names[i] == name) id = (int)i;
  if (id < 0) {
    id = (int)c->prof_names.size();
    c->prof_names.push_back(name);
  }
  ProfRec r{id, nullptr, nullptr};
  cudaEventCreate(&r.e0);
  cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, c->stream);
  c->prof_recs.push_back(r);
}
static void prof_end(fhesi_ctx *c) {
  if (c->prof_on) cudaEventRecord(c->prof_recs.back().e1, c->stream);
}
#define KLN(c, name, kern, grid, block, smem, ...)                         \
  do {                                                                     \
    prof_begin((c), name);                                                 \
    FHESI_LAUNCH(kern, grid, block, smem, (c)->stream, __VA_ARGS__);       \
    prof_end((c));                                                         \
  } while (0)
#define KL(c, kern, ...) KLN(c, #kern, kern, __VA_ARGS__)
// One device allocation shared by the key images made in one call (a whole set-up's matrices come out of
// fhesi_keygen_batch together): cudaMalloc costs ~0.1 ms a piece, three per matrix used to dominate key set-up
struct KeyBlock {
  void *ptr;
  int refs;
};
struct fhesi_ksw {
  fhesi_ctx *ctx;
  u32 *d_key;      // [Lk][parts*D][2][N] key form, residues in [0,p)
  u32 *d_key_bal;  // same, balanced residues (fused T-free path); NULL if unused
  u32 *d_key_split;  // [Ls][parts*D][4][N] balanced key form of (b_lo, b_hi, A_lo, A_hi), then the
                     // offset-correction table [Ls][4][N] (k_split_corr); NULL if unused
  u32 parts;
  KeyBlock *blk;
};
struct fhesi_key {
  fhesi_ctx *ctx;
  u32 *d_key;  // [Le][parts][1][N] key form  (and its transpose view [Le][1][parts][N])
  u32 parts;
};

const char *fhesi_last_error(void) { return g_err.c_str(); }
const char *fhesi_version(void) { return "fhesi_b200 0.1 (sm_100a)"; }

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
static u32 ilog2_ceil(u64 x) {
  u32 k = 0;
  while ((1ull << k) < x) ++k;
  return k;
}
// Phi_m(X) = prod_{d | m} (X^d - 1)^mu(m/d), dense, degree phi(m)  (NumbTh.cpp:142-159 computes the same polynomial)
static std::vector<long long> h_cyclotomic(u32 m) {
  auto mobius = [](u32 x) {
    int mu = 1;
    for (u32 q = 2; q * q <= x; ++q)
      if (x % q == 0) {
        x /= q;
        if (x % q == 0) return 0;
        mu = -mu;
      }
    return x > 1 ? -mu : mu;
  };
  std::vector<long long> a(1, 1);
  for (u32 d = 1; d <= m; ++d)  // numerator
    if (m % d == 0 && mobius(m / d) == 1) {
      std::vector<long long> b(a.size() + d, 0);
      for (size_t i = 0; i < a.size(); ++i) b[i + d] += a[i], b[i] -= a[i];
      a.swap(b);
    }
  for (u32 d = 1; d <= m; ++d)  // exact division by X^d - 1: a[i] = q[i - d] - q[i]
    if (m % d == 0 && mobius(m / d) == -1) {
      std::vector<long long> q(a.size() - d, 0);
      for (size_t i = 0; i < q.size(); ++i) q[i] = (i >= d ? q[i - d] : 0) - a[i];
      a.swap(q);
    }
  return a;
}
// columns X^(n+j) mod Phi_m for j < J as sparse (row, coefficient) lists; false if a coefficient leaves [-64, 64]
static bool h_reduction_columns(const std::vector<long long> &phi, u32 J, std::vector<std::vector<std::pair<u32, int>>> &cols) {
  const u32 n = (u32)phi.size() - 1;
  std::vector<long long> r(n);
  for (u32 i = 0; i < n; ++i) r[i] = -phi[i];  // X^n
  cols.assign(J, {});
  for (u32 j = 0; j < J; ++j) {
    for (u32 i = 0; i < n; ++i)
      if (r[i]) {
        if (r[i] > 64 || r[i] < -64) return false;
        cols[j].push_back(std::make_pair(i, (int)r[i]));
      }
    const long long top = r[n - 1];  // times X, minus top * Phi_m
    for (u32 i = n - 1; i > 0; --i) r[i] = r[i - 1] - top * phi[i];
    r[0] = -top * phi[0];
  }
  return true;
}

template <class T>
static int upload(fhesi_ctx *c, const std::vector<T> &h, const T **d) {
  void *p = nullptr;
  size_t cap = 0;
  CK(cached_malloc(c->device, &p, h.size() * sizeof(T) + 16, &cap));
  CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  c->tables.push_back(std::make_pair(p, cap));
  *d = (const T *)p;
  return 0;
}

int fhesi_ctx_create(uint32_t m, uint32_t logQ, uint64_t p_pt, uint32_t decompSize, uint64_t xi,
                     int device, fhesi_ctx **out) {
  if (!out) return fail(FHESI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (m < 3 || m > 8192) return fail(FHESI_ERR_UNSUPPORTED, "m must be in [3, 8192]");
  // m = 2h with h an odd prime (every parameter set of the reference's clients): Phi_m = sum (-X)^i and the
  // remainder is a fold written into the kernels.  Any other m takes Phi_m as a sparse remainder table.
  const bool twoh = m >= 6 && !(m & 1) && ((m / 2) & 1) && h_is_prime(m / 2);
  const u32 h = twoh ? m / 2 : 0;
  std::vector<long long> phi;
  if (!twoh) phi = h_cyclotomic(m);
  if (logQ < 8 || logQ > 512) return fail(FHESI_ERR_UNSUPPORTED, "logQ must be in [8, 512]");
  if (decompSize < 1 || decompSize > 3)
    return fail(FHESI_ERR_UNSUPPORTED, "decompSize must be 1..3 (digits must stay below 2^29)");
  if (p_pt < 2 || p_pt >= (1ull << 29)) return fail(FHESI_ERR_UNSUPPORTED, "p must be in [2, 2^29)");
  if (xi < 1) xi = 1;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(FHESI_ERR_CUDA, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(FHESI_ERR_INVALID, "bad device index");
  CK(cudaSetDevice(device));

  fhesi_ctx *c = new fhesi_ctx();
  c->device = device;
  const u32 n = twoh ? h - 1 : (u32)phi.size() - 1;
  u32 N = 16;  // store_index() needs N >= 16
  while (N < 2 * n - 1) N <<= 1;
  if (N > 2048 || n < 2) {
    delete c;
    return fail(FHESI_ERR_UNSUPPORTED, "phi(m) must be in [2, 1024]");
  }
  // growth of one ring product: |a b mod Phi_m|_inf <= G |a|_inf |b|_inf.  2h: the fold adds three runs of the
  // linear convolution, G = 2n - 2 (kept at 2n).  General m: from the remainder table, row by row.
  double G = 2.0 * n;
  std::vector<u32> red, red_wide;
  if (!twoh) {
    const u32 J = std::max(N - n, m - n);
    if (!h_reduction_columns(phi, J, c->red_cols)) {
      delete c;
      return fail(FHESI_ERR_UNSUPPORTED, "X^j mod Phi_m has coefficients beyond +-64 for this m");
    }
    std::vector<double> g(n);
    for (u32 i = 0; i < n; ++i) g[i] = i + 1;  // terms of the linear convolution at degree i
    for (u32 j = 0; j + 1 < n; ++j)
      for (auto &e : c->red_cols[j]) g[e.first] += std::abs(e.second) * (double)(2 * n - 1 - (n + j));
    G = *std::max_element(g.begin(), g.end());
    // rows of the gather, columns j < jmax: [n + 1 row starts][(j, coefficient) pairs]
    auto csr = [&](u32 jmax) {
      std::vector<u32> cnt(n + 1, 0);
      for (u32 j = 0; j < jmax; ++j)
        for (auto &e : c->red_cols[j]) cnt[e.first + 1]++;
      for (u32 i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
      std::vector<u32> t(n + 1 + 2 * (size_t)cnt[n], 0), fill(cnt.begin(), cnt.end() - 1);
      for (u32 i = 0; i <= n; ++i) t[i] = cnt[i];
      for (u32 j = 0; j < jmax; ++j)
        for (auto &e : c->red_cols[j]) {
          const u32 at = fill[e.first]++;
          t[n + 1 + 2 * at] = j;
          t[n + 1 + 2 * at + 1] = (u32)e.second;
        }
      return t;
    };
    red = csr(n - 1);        // a product of two ring elements: degree <= 2n - 2
    red_wide = csr(N - n);   // anything a transform-domain vector can hold (tensor form times a plaintext)
  }
  const u32 W = (logQ + 31) / 32, dbits = 8 * decompSize;
  const u32 D = (logQ + dbits - 1) / dbits;

  // number of 30-bit primes needed by each stage (DESIGN.md "chain sizing")
  const double lg2n = std::log2(G), lgp = std::log2((double)p_pt);
  double need_t = 2.0 * logQ - 2 + lgp + lg2n + std::log2(3.0) + std::log2((double)xi) + 1 + 0.1;
  // General m is no throughput configuration: give the tensor chain at least the reference's own sizing rule
  // (FHEContext.cpp:83-85: 2 log q + log p + 2 log phi(m) + log 2 + log xi), so that the tensor-form plaintext
  // operators (Ciphertext.cpp:157-159,252-256) have the head-room they have there.
  if (!twoh) need_t = std::max(need_t, 2.0 * logQ + lgp + 2 * std::log2((double)n) + 1 + std::log2((double)xi) + 0.1);
  const double need_k = dbits + (logQ - 1) + lg2n + std::log2(3.0 * D) + 1 + 0.1;
  const double need_e = logQ + lg2n + 10 + 1 + 0.1;
  std::vector<u32> primes;
  double bits = 0;
  u32 Lt = 0, Lk = 0, Le = 0;
  // T-free key switch (kernels_fused.cuh): with balanced residues the 3D-term inner product fits a
  // signed 64-bit accumulator if 3D (p/2)^2 < 2^63.  Start the chain below that cap when it costs
  // no extra prime at logQ <= 256 (cap ~ 2^29.98 for D = 11); otherwise keep 30-bit primes.
  u64 start = (1ull << 30) - 1;
  {
    const double cap = std::floor(std::sqrt(std::ldexp(1.0, 65) / (3.0 * D))) - 1;
    if (cap > 1.02 * std::ldexp(1.0, 29)) {
      if (cap < (double)start) start = (u64)cap;
      c->tfree = true;
    }
  }
  for (u64 q = start / N * N + 1; q > (1ull << 29) && Lt == 0; q -= N) {
    if (q >= (1ull << 30) || !h_is_prime(q)) continue;
    primes.push_back((u32)q);
    bits += std::log2((double)q);
    if (!Lk && bits >= need_k) Lk = (u32)primes.size();
    if (!Le && bits >= need_e) Le = (u32)primes.size();
    if (bits >= need_t) Lt = (u32)primes.size();
    if (primes.size() >= FHESI_MAX_PRIMES) break;
  }
  if (!Lt || !Lk || !Le) {
    delete c;
    return fail(FHESI_ERR_UNSUPPORTED, "parameter set needs more than FHESI_MAX_PRIMES primes");
  }
  if (Lk > Lt) Lk = Lt;
  if (Le > Lt) Le = Lt;
  const u32 L = Lt;
  // split-key key switch: halves of 32*ws bits (kernels_fused.cuh).  Needs W >= 2 and the
  // single-accumulator bound; inner products are non-negative and must stay below P_s / 2.
  u32 Ls = 0, ws = 0;
  if (c->tfree && W >= 2 && (N == 1024 || N == 2048)) {
    ws = (W + 1) / 2;
    const double need_s = dbits + 32.0 * ws + lg2n + std::log2(3.0 * D) + 1 + 0.1;
    double b2 = 0;
    for (u32 i = 0; i < L && !Ls; ++i) {
      b2 += std::log2((double)primes[i]);
      if (b2 >= need_s) Ls = i + 1;
    }
    if (!Ls || Ls >= Lk) Ls = 0, ws = 0;
  }

  fhesi_info &I = c->info;
  I.m = m; I.n = n; I.logQ = logQ; I.W = W; I.decompSize = decompSize; I.D = D; I.N = N;
  I.Lt = Lt; I.Lk = Lk; I.Le = Le; I.p = p_pt; I.xi = xi; I.device = device;
  I.Ls = Ls; I.split_words = ws;
  for (u32 i = 0; i < L; ++i) I.primes[i] = primes[i];

  // per-prime constants and tables
  const u32 CW = W + 2;
  std::vector<PrimeConst> pc(L);
  std::vector<u32> twf((size_t)L * N), twi((size_t)L * N), cw((size_t)L * CW), gar((size_t)L * L, 0);
  std::vector<uint2> twsf((size_t)L * N), twsi((size_t)L * N);
  std::vector<double2> twdf((size_t)L * N);
  std::vector<u32> cwr((size_t)L * 2 * CW);
  // floor(2^logQ / p_pt) mod q needs 2^logQ mod p_pt
  const u64 rem_q = h_powmod(2, logQ, p_pt);
  for (u32 l = 0; l < L; ++l) {
    const u64 q = primes[l];
    PrimeConst &P = pc[l];
    memset(&P, 0, sizeof(P));
    P.p = (u32)q;
    u32 inv = 1;  // Newton: inv = q^-1 mod 2^32
    for (int it = 0; it < 5; ++it) inv *= 2 - (u32)q * inv;
    P.pinv = (u32)(0 - inv);
    P.negp = 0u - (u32)q;
    P.hic = 0x43300000u;
    P.zop = 0u;
    const u64 R = (1ull << 32) % q, R2 = h_mulmod(R, R, q);
    const u64 ninv = h_invmod(N % q, q);
    P.r1 = (u32)R; P.r2 = (u32)R2;
    P.ninv_r = (u32)h_mulmod(ninv, R, q);
    P.ninv_r2 = (u32)h_mulmod(ninv, R2, q);
    P.tensor_c = (u32)h_mulmod(h_mulmod(p_pt % q, ninv, q), R2, q);
    P.ptxt_r = (u32)h_mulmod(p_pt % q, R, q);
    const u64 two_q = h_powmod(2, logQ, q);
    const u64 scale = h_mulmod((two_q + q - rem_q % q) % q, h_invmod(p_pt % q, q), q);
    P.scale_r = (u32)h_mulmod(scale, R, q);
    // primitive N-th root of unity: z^((q-1)/N) for a quadratic non-residue z
    u64 z = 2;
    while (h_powmod(z, (q - 1) / 2, q) != q - 1) ++z;
    const u64 w = h_powmod(z, (q - 1) / N, q), wi = h_invmod(w, q);
    for (u32 hh = 1; hh < N; hh <<= 1) {
      const u64 step = h_powmod(w, N / (2 * hh), q), istep = h_powmod(wi, N / (2 * hh), q);
      u64 a = R, b = R;  // Montgomery form of 1
      for (u32 j = 0; j < hh; ++j) {
        twf[(size_t)l * N + hh + j] = (u32)a;
        twi[(size_t)l * N + hh + j] = (u32)b;
        const u64 ap = h_mulmod(a, h_invmod(R, q), q), bp = h_mulmod(b, h_invmod(R, q), q);  // plain
        twsf[(size_t)l * N + hh + j] = make_uint2((u32)ap, (u32)((ap << 32) / q));
        {
          const u64 kq = (u64)(((unsigned __int128)ap << 50) / q);  // < 2^50
          twdf[(size_t)l * N + hh + j] = make_double2(std::ldexp((double)kq, -50), 4503599627370496.0 - 4.0 * (double)kq);
        }
        twsi[(size_t)l * N + hh + j] = make_uint2((u32)bp, (u32)((bp << 32) / q));
        a = h_mulmod(a, step, q);
        b = h_mulmod(b, istep, q);
      }
    }
    twf[(size_t)l * N] = twi[(size_t)l * N] = (u32)R;
    u64 t = R;  // 2^(32k) * R
    for (u32 k = 0; k < CW; ++k) {
      cw[(size_t)l * CW + k] = (u32)t;
      t = h_mulmod(t, R, q);
    }
    twsf[(size_t)l * N] = twsi[(size_t)l * N] = make_uint2(1u, (u32)((1ull << 32) / q));
    {
      const u64 kq = (u64)(((unsigned __int128)1 << 50) / q);
      twdf[(size_t)l * N] = make_double2(std::ldexp((double)kq, -50), 4503599627370496.0 - 4.0 * (double)kq);
    }
    for (u32 v = 0; v < 2; ++v) {
      const u64 sv = v ? h_mulmod(h_mulmod(p_pt % q, ninv, q), R, q) : 1;
      u64 c2 = h_mulmod(R, sv, q);  // 2^(32k) * R * s_v
      for (u32 k = 0; k < CW; ++k) {
        // entry W holds the correction for a negative top word: p - 2^(32W) * s_v (no extra R:
        // it is added after the Montgomery reduction)
        cwr[((size_t)l * 2 + v) * CW + k] = (u32)c2;
        c2 = h_mulmod(c2, R, q);
      }
      const u64 topc = h_mulmod(h_powmod(2, 32ull * W, q), sv, q);
      cwr[((size_t)l * 2 + v) * CW + W] = (u32)((q - topc) % q);
    }
    for (u32 i = 0; i < l; ++i)
      gar[(size_t)l * L + i] = (u32)h_mulmod(h_invmod(primes[i] % q, q), R, q);
  }
  // prefix products and their halves, L words each
  std::vector<u32> Pf((size_t)(L + 1) * L, 0), Ph((size_t)(L + 1) * L, 0);
  {
    std::vector<u32> cur(L, 0);
    cur[0] = 1;
    for (u32 l = 0; l <= L; ++l) {
      for (u32 k = 0; k < L; ++k) Pf[(size_t)l * L + k] = cur[k];
      u32 carry = 0;
      for (int k = (int)L - 1; k >= 0; --k) {
        Ph[(size_t)l * L + k] = (cur[k] >> 1) | (carry << 31);
        carry = cur[k] & 1;
      }
      if (l < L) {
        u64 cy = 0;
        for (u32 k = 0; k < L; ++k) {
          u64 t = (u64)cur[k] * primes[l] + cy;
          cur[k] = (u32)t;
          cy = t >> 32;
        }
      }
    }
  }
  DevCtx &dc = c->dc;
  dc.n = n; dc.N = N; dc.logN = ilog2_ceil(N); dc.W = W; dc.logQ = logQ; dc.D = D; dc.dbits = dbits;
  dc.h = h; dc.Lmax = L; dc.CW = CW; dc.ptxt = (u32)p_pt;
  dc.sshift = N == 2048 ? 2 : 1;
  c->h_pc = pc;
  c->h_garner = gar;
  c->h_Pfull = Pf;
  c->h_Phalf = Ph;
  int rc = 0;
  if ((rc = upload(c, pc, &dc.pc)) || (rc = upload(c, twf, &dc.tw_fwd)) ||
      (rc = upload(c, twi, &dc.tw_inv)) || (rc = upload(c, cw, &dc.cword)) ||
      (rc = upload(c, gar, &dc.garner)) || (rc = upload(c, Pf, &dc.Pfull)) ||
      (rc = upload(c, Ph, &dc.Phalf)) || (rc = upload(c, twsf, &dc.tws_fwd)) ||
      (rc = upload(c, twsi, &dc.tws_inv)) || (rc = upload(c, cwr, &dc.cwr)) || (rc = upload(c, twdf, &dc.twd_fwd))) {
    fhesi_ctx_destroy(c);
    return rc;
  }
  dc.red = dc.red_wide = nullptr;
  if (!twoh && ((rc = upload(c, red, &dc.red)) || (rc = upload(c, red_wide, &dc.red_wide)))) {
    fhesi_ctx_destroy(c);
    return rc;
  }
  if (N == FN || N == FT<T2K>::N) {  // the fused kernels' per-thread twiddle layout, laid out once
    const u32 entries = N == FN ? FTW_ENTRIES : FT<T2K>::ENTRIES;
    std::vector<uint2> lf((size_t)L * entries), li((size_t)L * entries);
    for (u32 l = 0; l < L; ++l)
      for (u32 e = 0; e < entries; ++e) {
        const u32 src = N == FN ? ftw_source_index(e) : ftw_source_index_t<T2K>(e);
        lf[(size_t)l * entries + e] = twsf[(size_t)l * N + src];
        li[(size_t)l * entries + e] = twsi[(size_t)l * N + src];
      }
    if ((rc = upload(c, lf, &dc.ftw_fwd)) || (rc = upload(c, li, &dc.ftw_inv))) {
      fhesi_ctx_destroy(c);
      return rc;
    }
  }
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  {
    void *pc = nullptr;
    size_t cap = 0;
    CK(cached_malloc(c->device, &pc, 16, &cap));
    CK(cudaMemset(pc, 0, 16));
    c->tables.push_back(std::make_pair(pc, cap));
    c->d_crt_fallbacks = (unsigned long long *)pc;
  }
  {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c->device));
    if (prop.multiProcessorCount > 0) c->sm_count = prop.multiProcessorCount;
  }
  const char *ev = getenv("FHESI_CHUNK");
  if (ev && atoi(ev) > 0) c->chunk = (u32)atoi(ev);
  ev = getenv("FHESI_FUSED_CHUNK");
  if (ev && atoi(ev) > 0) c->fused_chunk = (u32)atoi(ev);
  ev = getenv("FHESI_PIPE_CHUNK");
  if (ev && atoi(ev) > 0) c->pipe_chunk = (u32)atoi(ev);
  ev = getenv("FHESI_PIPE_TAPER");
  if (ev) c->pipe_taper = atoi(ev) > 0;
  ev = getenv("FHESI_PIPE_LANES");
  if (ev && atoi(ev) > 0) c->pipe_lanes = atoi(ev) > 1 ? 2 : 1;
  ev = getenv("FHESI_NO_CRT_DIRECT");
  if (ev && atoi(ev) > 0) c->crt_direct = false;
  ev = getenv("FHESI_CRT_FORCE_EXACT");
  if (ev && atoi(ev) > 0) c->crt_force_exact = true;
  ev = getenv("FHESI_NO_SPLIT");
  if (ev && atoi(ev) > 0) c->info.Ls = 0, c->info.split_words = 0;
  ev = getenv("FHESI_NO_FUSED");
  if (ev && atoi(ev) > 0) c->use_fused = false;
  if (!fused_supported(dc)) c->use_fused = false;
  if (dc.N != FN && !c->info.Ls) c->use_fused = false;  // N = 2048 has the split-key kernel only
  if (c->use_fused) {
    rc = dc.N == FN ? fused_configure() : fused2k_configure();
    if (rc) {
      fhesi_ctx_destroy(c);
      return fail(FHESI_ERR_CUDA, "cudaFuncSetAttribute failed for the fused kernels");
    }
  }
  *out = c;
  return 0;
}

void fhesi_ctx_destroy(fhesi_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  // everything goes to the device's cache (see DeviceCache), not back to the driver: the device is idle now
  for (auto &t : c->tables) cache_put(c->device, t.first, t.second);
  for (auto &kv : c->pool_free)
    for (void *p : kv.second) cache_put(c->device, p, kv.first);
  // blocks still handed out (a caller's buffers, key images whose handle was never destroyed) die with
  // the context they were allocated from
  for (auto &kv : c->pool_size) cache_put(c->device, kv.first, kv.second);
  cache_put(c->device, c->scratch.ptr, c->scratch.cap);
  cache_put(c->device, c->lane_scratch.ptr, c->lane_scratch.cap);
  cache_put(c->device, c->stage.ptr, c->stage.cap);
  cache_put(c->device, c->work.ptr, c->work.cap);
  prof_clear(c);
  for (auto e : c->pipe_events) cudaEventDestroy(e);
  for (auto e : c->half_done)
    if (e) cudaEventDestroy(e);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->lane_stream) cudaStreamDestroy(c->lane_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}
int fhesi_ctx_info(const fhesi_ctx *c, fhesi_info *out) {
  if (!c || !out) return fail(FHESI_ERR_INVALID, "null argument");
  *out = c->info;
  return 0;
}
int fhesi_ctx_set_stream(fhesi_ctx *c, void *s) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return 0;
}
int fhesi_sync(fhesi_ctx *c) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int fhesi_sync_all(fhesi_ctx *c) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  if (c->d2h_stream) CK(cudaStreamSynchronize(c->d2h_stream));
  if (c->lane_stream) CK(cudaStreamSynchronize(c->lane_stream));
  if (c->h2d_stream) CK(cudaStreamSynchronize(c->h2d_stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
// Device allocations made through the ABI are pooled per context: the host layer's value-semantics
// Ciphertext allocates and frees a buffer per operator, and every operator is enqueued on the
// context's one stream, so handing a freed block to the next fhesi_malloc is stream-ordered safe and
// needs neither cudaFree nor a synchronisation.
int fhesi_malloc(fhesi_ctx *c, size_t bytes, void **d) {
  if (!c || !d) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  size_t sz = ((bytes ? bytes : 1) + 255) & ~(size_t)255;
  auto it = c->pool_free.find(sz);
  if (it != c->pool_free.end() && !it->second.empty()) {
    *d = it->second.back();
    it->second.pop_back();
    c->pool_idle_bytes -= sz;
  } else {
    CK(cached_malloc(c->device, d, sz, &sz));  // sz becomes the block's real size: it returns to that bucket
  }
  c->pool_size[*d] = sz;
  return 0;
}
int fhesi_free(fhesi_ctx *c, void *d) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  if (!d) return 0;
  CK(cudaSetDevice(c->device));
  auto it = c->pool_size.find(d);
  if (it == c->pool_size.end()) {  // not ours: release for real
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(d));
    return 0;
  }
  const size_t sz = it->second;
  c->pool_size.erase(it);
  if (c->pool_idle_bytes + sz > c->pool_cap_bytes) {
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(d));
  } else {
    c->pool_free[sz].push_back(d);
    c->pool_idle_bytes += sz;
  }
  return 0;
}
int fhesi_h2d(fhesi_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int fhesi_h2d_async(fhesi_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
int fhesi_crt_fallbacks(fhesi_ctx *c, uint64_t *count) {
  if (!c || !count) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, c->d_crt_fallbacks, sizeof v, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *count = v;
  return 0;
}
int fhesi_host_alloc(size_t bytes, int write_combined, void **out) {
  if (!out) return fail(FHESI_ERR_INVALID, "out is NULL");
  *out = nullptr;
  CK(cudaHostAlloc(out, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
  return 0;
}
int fhesi_host_free(void *p) {
  if (p) CK(cudaFreeHost(p));
  return 0;
}
int fhesi_d2h(fhesi_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
int fhesi_d2d(fhesi_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
size_t fhesi_ct_bytes(const fhesi_ctx *c, uint32_t parts) {
  return c ? (size_t)parts * c->info.n * c->info.W * 4 : 0;
}
size_t fhesi_tprod_bytes(const fhesi_ctx *c, uint32_t parts) {
  return c ? (size_t)parts * c->info.Lt * c->info.N * 4 : 0;
}

static int scratch(fhesi_ctx *c, size_t bytes, u32 **p) {
  if (c->scratch.cap < bytes) {
    CK(cudaStreamSynchronize(c->stream));  // the arena is only ever used by work on this stream
    cache_put(c->device, c->scratch.ptr, c->scratch.cap);
    c->scratch.ptr = nullptr;
    c->scratch.cap = 0;
    CK(cached_malloc(c->device, &c->scratch.ptr, bytes, &c->scratch.cap));
  }
  *p = (u32 *)c->scratch.ptr;
  return 0;
}
static inline size_t al(size_t words) { return (words + 63) & ~(size_t)63; }

// ---------------------------------------------------------------------------------------
// launch helpers for the generic kernels
// ---------------------------------------------------------------------------------------
static int launch_fwd(fhesi_ctx *c, const void *src, u32 src_mode, u32 Win, u32 scale_mode, u32 L,
                      u32 *dst, size_t npolys) {
  if (!npolys) return 0;
  FwdArgs a{src, Win, L, src_mode, scale_mode, dst};
  const DevCtx &dc = c->dc;
  dim3 grid((unsigned)npolys, L), block(dc.N / 2 < 32 ? 32 : dc.N / 2);
  KL(c, k_fwd, grid, block, dc.N * 4, dc, a);
  CKL();
  return 0;
}
static int launch_inv(fhesi_ctx *c, const u32 *src, u32 L, u32 *dst, size_t npolys,
                      const int *e = nullptr, const u32 *msg = nullptr) {
  if (!npolys) return 0;
  InvArgs a{src, L, dst, e, msg};
  const DevCtx &dc = c->dc;
  dim3 grid((unsigned)npolys, L), block(dc.N / 2 < 32 ? 32 : dc.N / 2);
  KL(c, k_inv, grid, block, (dc.N + dc.h) * 4, dc, a);
  CKL();
  return 0;
}
// the by-value table for (ML, L) is built once per context and cached
template <int ML>
static const CrtTables<ML> &crt_tables(fhesi_ctx *c, u32 L) {
  std::vector<unsigned char> &raw = c->crt_tables[std::make_pair(ML, L)];
  if (raw.empty()) {
    raw.assign(sizeof(CrtTables<ML>), 0);
    CrtTables<ML> &T = *reinterpret_cast<CrtTables<ML> *>(raw.data());
    const u32 LM = c->dc.Lmax;
    for (u32 j = 0; j < L && j < (u32)ML; ++j) {
      T.p[j] = c->h_pc[j].p;
      T.pinv[j] = c->h_pc[j].pinv;
      for (u32 i = 0; i < j; ++i) {
        const u64 q = c->h_pc[j].p, gi = h_invmod(c->h_pc[i].p % q, q);
        T.garner[j][i] = (u32)gi;
        T.garnerq[j][i] = (u32)((gi << 32) / q);
      }
    }
    for (u32 k = 0; k < LM && k < (u32)ML; ++k) {
      T.Pfull[k] = c->h_Pfull[(size_t)L * LM + k];
      T.Phalf[k] = c->h_Phalf[(size_t)L * LM + k];
    }
  }
  return *reinterpret_cast<const CrtTables<ML> *>(raw.data());
}
// tables of k_crt_direct for the first L primes and the limb window of this context's logQ; cached like crt_tables
template <int ML, int NL>
static const CrtDirectTables<ML, NL> &crt_direct_tables(fhesi_ctx *c, u32 L, u32 j0, u32 nl) {
  std::vector<unsigned char> &raw = c->crt_tables[std::make_pair(1000 + ML, L)];
  if (raw.empty()) {
    raw.assign(sizeof(CrtDirectTables<ML, NL>), 0);
    CrtDirectTables<ML, NL> &T = *reinterpret_cast<CrtDirectTables<ML, NL> *>(raw.data());
    const u32 LM = c->dc.Lmax;
    const u32 *P = &c->h_Pfull[(size_t)L * LM];  // LM words of prod_{i<L} p_i
    T.j0 = j0, T.nl = nl, T.force_exact = c->crt_force_exact;
    T.fallbacks = c->d_crt_fallbacks;
    // 28-bit limb `j` of a little-endian word array
    auto limb = [&](const u32 *w, u32 j) -> u32 {
      const u32 o = CRT_LB * j, k = o >> 5, sh = o & 31;
      const u64 lo = k < LM ? w[k] : 0u, hi = k + 1 < LM ? w[k + 1] : 0u;
      return (u32)((lo | (hi << 32)) >> sh) & CRT_LMASK;
    };
    {  // limbs of 2^(28 (j0 + nl)) - P: complement every limb of P up to the window's top, plus one
      u32 carry = 1;
      for (u32 j = 0; j < j0 + nl; ++j) {
        const u32 t = (CRT_LMASK ^ limb(P, j)) + carry;
        carry = t >> CRT_LB;
        if (j >= j0) T.NP[j - j0] = t & CRT_LMASK;
      }
    }
    std::vector<u32> q(LM);
    for (u32 i = 0; i < L; ++i) {
      const u64 p = c->h_pc[i].p;
      u64 rem = 0;
      for (int k = (int)LM - 1; k >= 0; --k) {  // C_i = P / p_i, exact
        const u64 cur = (rem << 32) | P[k];
        q[k] = (u32)(cur / p);
        rem = cur % p;
      }
      u64 cm = 0;  // C_i mod p_i
      for (int k = (int)LM - 1; k >= 0; --k) cm = (u64)((((unsigned __int128)cm << 32) | q[k]) % p);
      const u64 yi = h_invmod(cm, p);
      T.p[i] = (u32)p;
      T.yinv[i] = (u32)yi;
      T.yinvq[i] = (u32)((yi << 32) / p);
      {
        const unsigned __int128 rf = ((unsigned __int128)1 << 84) / p;  // < 2^55
        T.rfix_hi[i] = (u32)(rf >> 28);
        T.rfix_lo[i] = (u32)rf & ((1u << 28) - 1u);
      }
      for (u32 j = 0; j < nl; ++j) T.C[i][j] = limb(q.data(), j0 + j);
    }
  }
  return *reinterpret_cast<const CrtDirectTables<ML, NL> *>(raw.data());
}
template <int ML, int NL>
static bool launch_crt_direct_t(fhesi_ctx *c, const CrtArgs &a, u32 j0, u32 nl) {
  if (nl > (u32)NL) return false;
  const int B = 128;
  const unsigned g = (unsigned)((a.total + B - 1) / B);
  const CrtDirectTables<ML, NL> &T = crt_direct_tables<ML, NL>(c, a.L, j0, nl);
  if (nl == (u32)NL) {
    const auto kern = k_crt_direct<ML, NL, true>;
    KLN(c, "k_crt_direct<ML>", kern, g, B, NL * B * 4, c->dc, a, T);
  } else {
    const auto kern = k_crt_direct<ML, NL, false>;
    KLN(c, "k_crt_direct<ML>", kern, g, B, NL * B * 4, c->dc, a, T);
  }
  return true;
}
// ScaleDown modes through k_crt_direct when the window fits an instantiation; false = use k_crt
static bool launch_crt_direct(fhesi_ctx *c, const CrtArgs &a) {
  if (!c->crt_direct || (a.mode != CRT_SCALEDOWN && a.mode != CRT_SCALEDOWN_DIGITS)) return false;
  const u32 logQ = c->dc.logQ, hl = (logQ - 1) / CRT_LB, j0 = hl >= 2 ? hl - 2 : 0, j1 = (2 * logQ - 1) / CRT_LB;
  const u32 nl = j1 - j0 + 1, L = a.L;
  if (29.0 * L < 2.0 * logQ + 2) return false;  // the chain must reach past bit 2 logQ (always true of a tensor chain)
  if (L <= 8) return launch_crt_direct_t<8, 7>(c, a, j0, nl);
  if (L <= 10) return launch_crt_direct_t<10, 8>(c, a, j0, nl);
  if (L <= 13) return launch_crt_direct_t<13, 9>(c, a, j0, nl);
  if (L <= 18) return launch_crt_direct_t<18, 12>(c, a, j0, nl);
  if (L <= 20) return launch_crt_direct_t<20, 13>(c, a, j0, nl);
  if (L <= 28) return launch_crt_direct_t<28, 17>(c, a, j0, nl);
  if (L <= 36) return launch_crt_direct_t<36, 21>(c, a, j0, nl);
  return launch_crt_direct_t<40, 25>(c, a, j0, nl);
}
template <int ML>
static void launch_crt_t(fhesi_ctx *c, const CrtArgs &a) {
  const int B = 128;
  unsigned g = (unsigned)((a.total + B - 1) / B);
  const CrtTables<ML> &T = crt_tables<ML>(c, a.L);
  KL(c, k_crt<ML>, g, B, ML * B * 4, c->dc, a, T);
}
template <int ML>
static void launch_crt_split_t(fhesi_ctx *c, const CrtSplitArgs &a) {
  const int B = 128;
  unsigned g = (unsigned)((a.total + B - 1) / B);
  const CrtTables<ML> &T = crt_tables<ML>(c, a.L);
  KL(c, k_crt_split<ML>, g, B, 2 * ML * B * 4, c->dc, a, T);
}
static int launch_crt_split(fhesi_ctx *c, const u32 *res, u32 L, u32 *out, size_t npolys) {
  if (!npolys) return 0;
  CrtSplitArgs a{res, L, c->info.split_words, out, npolys * c->dc.n};
  if (L <= 6) launch_crt_split_t<6>(c, a);  // exact sizes for the common chains: no dead words in the Horner rows
  else if (L <= 8) launch_crt_split_t<8>(c, a);
  else if (L <= 12) launch_crt_split_t<12>(c, a);
  else if (L <= 20) launch_crt_split_t<20>(c, a);
  else return fail(FHESI_ERR_UNSUPPORTED, "split CRT: too many primes");
  CKL();
  return 0;
}
static int launch_crt(fhesi_ctx *c, const u32 *res, u32 L, u32 mode, u32 *out, u32 Wout,
                      size_t npolys) {
  if (!npolys) return 0;
  CrtArgs a{res, L, mode, out, Wout, npolys * c->dc.n};
  if (launch_crt_direct(c, a)) {
    CKL();
    return 0;
  }
  // DECRYPT multiplies by p_pt before the shift: two words of head-room
  u32 need = L + (mode == CRT_DECRYPT ? 2 : 0);
  if (need <= 8) launch_crt_t<8>(c, a);
  else if (need <= 10) launch_crt_t<10>(c, a);
  else if (need <= 13) launch_crt_t<13>(c, a);
  else if (need <= 18) launch_crt_t<18>(c, a);
  else if (need <= 20) launch_crt_t<20>(c, a);
  else if (need <= 28) launch_crt_t<28>(c, a);
  else if (need <= 36) launch_crt_t<36>(c, a);
  else launch_crt_t<42>(c, a);
  CKL();
  return 0;
}
static inline unsigned nblk(size_t total, int B = 256) { return (unsigned)((total + B - 1) / B); }

// k_residues + k_fused_tensor over one chunk; resid: scratch of cnt*4*Lt*n words
static int launch_fused_tensor(fhesi_ctx *c, const u32 *a, const u32 *b, u32 *resid, u32 *out, size_t cnt,
                               int to_tprod) {
  const fhesi_info &I = c->info;
  ResidueArgs r{a, b, resid, I.Lt, cnt};
  const unsigned rg = nblk(cnt * 4 * I.n, 128);
  const size_t rsm = [&](u32 row) { return (size_t)(I.Lt * 2 * row + 2 * I.Lt) * 4; }((I.W + 1 + 3) & ~3u);
  switch (I.W) {  // word count as a compile-time constant for the common moduli (logQ = 128, 176, 256, 512)
    case 4: KL(c, k_residues_t<4>, rg, 128, rsm, c->dc, r); break;
    case 6: KL(c, k_residues_t<6>, rg, 128, rsm, c->dc, r); break;
    case 8: KL(c, k_residues_t<8>, rg, 128, rsm, c->dc, r); break;
    case 16: KL(c, k_residues_t<16>, rg, 128, rsm, c->dc, r); break;
    default: KL(c, k_residues, rg, 128, (I.Lt * 2 * c->dc.CW + 2 * I.Lt) * 4, c->dc, r);
  }
  CKL();
  // ops per group: more of them amortise the CTA's twiddle-table fill (a straight 16 KB copy, about
  // 0.04 of one op's work), fewer keep the last wave full; pick the best product of the two
  u32 opg = 1;
  {
    double best = 0;
    for (u32 o = 1; o <= 4; ++o) {
      const u32 kg = I.N == FN ? KG : KG2;
      const double ctas = (double)I.Lt * (double)((cnt + kg * o - 1) / (kg * o));
      const double waves = ctas / (c->sm_count * (I.N == FN ? KG_MINB : KG2_MINB));
      const double eff = waves / std::ceil(waves) / (1.0 + 0.04 / o);
      if (eff > best * 1.0001) best = eff, opg = o;
    }
  }
  FusedTensorArgs t{resid, out, I.Lt, (u32)cnt, opg, (u32)to_tprod};
  if (I.N == FN) {
    dim3 grid(I.Lt, (unsigned)((cnt + KG * opg - 1) / (KG * opg)));
    const auto kern = c->dc.h ? k_fused_tensor<false> : k_fused_tensor<true>;
    KLN(c, "k_fused_tensor", kern, grid, KG * 128, FUSED_SMEM_WORDS * 4, c->dc, t);
  } else {
    dim3 grid(I.Lt, (unsigned)((cnt + KG2 * opg - 1) / (KG2 * opg)));
    const auto kern = c->dc.h ? k_fused_tensor_2k<false> : k_fused_tensor_2k<true>;
    KLN(c, "k_fused_tensor_2k", kern, grid, KG2 * T2K, FUSED2K_SMEM_WORDS * 4, c->dc, t);
  }
  CKL();
  return 0;
}

// ---------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------
// Device half of the key upload: b, A as [K][n][W] coefficient words in HBM -> key-form images
// (prime-major, balanced, split halves).  Enqueued on the context's stream; the caller synchronises.
// words of scratch and of key images one matrix needs (ksw_build_from_device)
static void ksw_sizes(const fhesi_ctx *c, uint32_t parts, size_t *scratch_words, size_t *key_words) {
  const fhesi_info &I = c->info;
  const size_t K = (size_t)parts * I.D, polyw = (size_t)I.n * I.W, full = K * 2 * I.Lk * I.N;
  const bool bal = c->use_fused && c->tfree, split = bal && I.Ls;
  const size_t sp = split ? K * 4 * I.Ls * I.N : 0;
  *scratch_words = al(K * 2 * polyw) + al(full) + (split ? al(K * 4 * polyw) + 2 * al(sp) : 0);
  *key_words = al(full) + (bal ? al(full) : 0) + (split ? al(sp + (size_t)I.Ls * 4 * I.N) : 0);
}
// b, A [K][n][W] on the device -> key-form images in `keymem` (ksw_sizes words), using `scr` as scratch.
// Enqueued on the context's stream; the caller synchronises.  blk: the allocation keymem lives in.
static int ksw_build_from_device(fhesi_ctx *c, const u32 *d_b, const u32 *d_A, uint32_t parts, u32 *scr, u32 *keymem,
                                 KeyBlock *blk, fhesi_ksw **out) {
  const fhesi_info &I = c->info;
  const u32 K = parts * I.D, Lk = I.Lk;
  const size_t polyw = (size_t)I.n * I.W, full = (size_t)K * 2 * Lk * I.N;
  const bool bal = c->use_fused && c->tfree, split = bal && I.Ls;
  u32 *d_in = scr, *d_tmp = d_in + al((size_t)K * 2 * polyw), *d_in2 = d_tmp + al(full);
  u32 *d_key = keymem, *d_bal = bal ? d_key + al(full) : nullptr, *d_split = split ? d_bal + al(full) : nullptr;
  int rc = 0;
  // interleave to [K][2] so that one transform launch writes [K*2][Lk][N]; K mod q (non-negative)
  // = lo + 2^(32 ws) hi, each half as a non-negative W-word polynomial
  KL(c, k_key_stage, nblk((size_t)K * 2 * I.n), 256, 0, d_b, d_A, d_in, split ? d_in2 : (u32 *)nullptr, K, I.n, I.W,
     I.split_words, I.logQ & 31);
  CKL();
  if ((rc = launch_fwd(c, d_in, SRC_POLY, I.W, SC_KEYFORM, Lk, d_tmp, (size_t)K * 2))) return rc;
  // [K*2][Lk][N] -> [Lk][K*2][N]
  KL(c, k_transpose_key, nblk(full), 256, 0, d_tmp, d_key, K * 2, Lk, I.N);
  CKL();
  if (bal) {
    KL(c, k_balance_key, nblk(full), 256, 0, c->dc, d_key, d_bal, K * 2, full);
    CKL();
  }
  if (split) {
    const u32 Ls = I.Ls;
    const size_t total = (size_t)K * 4 * Ls * I.N;
    u32 *d_tmp2 = d_in2 + al((size_t)K * 4 * polyw), *d_t2 = d_tmp2 + al(total);
    if ((rc = launch_fwd(c, d_in2, SRC_POLY, I.W, SC_KEYFORM, Ls, d_tmp2, (size_t)K * 4))) return rc;
    KL(c, k_transpose_key, nblk(total), 256, 0, d_tmp2, d_t2, K * 4, Ls, I.N);
    CKL();
    KL(c, k_balance_key, nblk(total), 256, 0, c->dc, d_t2, d_split, K * 4, total);
    CKL();
    KL(c, k_split_corr, nblk((size_t)Ls * 4 * I.N), 256, 0, c->dc, d_split, d_split + total, K, Ls);  // + correction table
    CKL();
  }
  blk->refs++;
  *out = new fhesi_ksw{c, d_key, d_bal, d_split, parts, blk};
  return 0;
}
// M matrices from consecutive entries of d_b / d_A: one scratch block, one key block for all of them
static int ksw_build_many(fhesi_ctx *c, const u32 *d_b, const u32 *d_A, uint32_t M, const uint32_t *parts,
                          fhesi_ksw **out) {
  const size_t polyw = (size_t)c->info.n * c->info.W;
  size_t scr_max = 0, key_total = 0;
  for (u32 m = 0; m < M; ++m) {
    size_t sw, kw;
    ksw_sizes(c, parts[m], &sw, &kw);
    scr_max = std::max(scr_max, sw);
    key_total += kw;
  }
  PoolTmp scr(c), keys(c);
  int rc = scr.alloc(scr_max * 4);
  if (!rc) rc = keys.alloc(key_total * 4);
  if (rc) return rc;
  KeyBlock *blk = new KeyBlock{keys.u(), 0};
  size_t off = 0, koff = 0;
  for (u32 m = 0; m < M; ++m) {
    out[m] = nullptr;
    size_t sw, kw;
    ksw_sizes(c, parts[m], &sw, &kw);
    // stream order makes the shared scratch safe: matrix m+1's kernels run after matrix m's
    if ((rc = ksw_build_from_device(c, d_b + off * polyw, d_A + off * polyw, parts[m], scr.u(), keys.u() + koff, blk,
                                    &out[m]))) {
      for (u32 k = 0; k < m; ++k) delete out[k], out[k] = nullptr;
      delete blk;
      return rc;  // `keys` is still owned by the PoolTmp and goes back to the pool
    }
    off += (size_t)parts[m] * c->info.D;
    koff += kw;
  }
  keys.release();  // now owned by blk
  return 0;
}
int fhesi_ksw_create(fhesi_ctx *c, const uint32_t *h_b, const uint32_t *h_A, uint32_t parts,
                     fhesi_ksw **out) {
  if (!c || !h_b || !h_A || !out || parts < 1 || parts > 3)
    return fail(FHESI_ERR_INVALID, "fhesi_ksw_create: bad argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const size_t bytes = (size_t)parts * I.D * I.n * I.W * 4;
  PoolTmp d_b(c), d_A(c);
  int rc = d_b.alloc(bytes);
  if (!rc) rc = d_A.alloc(bytes);
  if (rc) return rc;
  CK(cudaMemcpyAsync(d_b.u(), h_b, bytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_A.u(), h_A, bytes, cudaMemcpyHostToDevice, c->stream));
  rc = ksw_build_many(c, d_b.u(), d_A.u(), 1, &parts, out);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));  // h_b / h_A may be pinned: the caller gets them back consumed
  return 0;
}
// KeySwitchSI::Init (FHE-SI.cpp:153-209) and FHESIPubKey::Init (FHE-SI.cpp:42-62) on the device, from the
// caller's draws (the reference's order and number of draws stay with the caller).  Every
// entry q of every matrix has the same shape,
//   b_q = A_q * t + e_q + src_{q / D} * 2^(dbits (q % D))  reduced mod q  (:182-199),   A'_q = -A_q  (:178-180),
// and the public key is one more entry with no source term (c0 = e + s * c1, c1' = -c1, :44-58) -- so all
// matrices of a set-up (s^2 -> s and the log2(slots) rotation keys: 9 at p = 1019) and the public key go
// through ONE pass of kernels over the concatenated entries: one upload, one transform launch, one CRT.
// d_A [Kt][n][W], d_e [Kt][n], d_src [rows][n] (rows >= ceil(Kt / D)), d_t [n]  ->  d_b, d_An [Kt][n][W].
static int keygen_entries(fhesi_ctx *c, u32 Kt, const u32 *d_A, const int *d_e, const int *d_src, const int *d_t,
                          u32 *d_b, u32 *d_An) {
  const fhesi_info &I = c->info;
  const u32 Le = I.Le, n = I.n;
  const size_t polyw = (size_t)n * I.W, per = (size_t)Le * I.N;
  PoolTmp d_pow(c), s0(c), s1(c), s2(c), s3(c), s4(c);
  int rc = 0;
  if ((rc = d_pow.alloc((size_t)I.D * Le * 4)) || (rc = s0.alloc(per * 4)) || (rc = s1.alloc(per * 4)) ||
      (rc = s2.alloc(Kt * per * 4)) || (rc = s3.alloc(Kt * per * 4)) || (rc = s4.alloc((size_t)Kt * Le * n * 4)))
    return rc;
  // 2^(dbits j) mod p_l in Montgomery form, [D][Le]
  std::vector<u32> pw((size_t)I.D * Le);
  for (u32 l = 0; l < Le; ++l) {
    const u64 q = I.primes[l];
    u64 v = c->h_pc[l].r1 % q, two = 1;
    for (u32 t = 0; t < 8 * I.decompSize; ++t) two = two * 2 % q;
    for (u32 j = 0; j < I.D; ++j) {
      pw[(size_t)j * Le + l] = (u32)v;
      v = h_mulmod(v, two, q);
    }
  }
  CK(cudaMemcpyAsync(d_pow.u(), pw.data(), pw.size() * 4, cudaMemcpyHostToDevice, c->stream));  // pageable: consumed on return
  // t in key form [Le][1][1][N]; images of every A; pointwise products; back with the addend
  if ((rc = launch_fwd(c, d_t, SRC_I32, 0, SC_KEYFORM, Le, s0.u(), 1))) return rc;
  KL(c, k_transpose_key, nblk(per), 256, 0, s0.u(), s1.u(), 1, Le, I.N);
  CKL();
  if ((rc = launch_fwd(c, d_A, SRC_POLY, I.W, SC_NONE, Le, s2.u(), Kt))) return rc;
  DotArgs d{s2.u(), s1.u(), 1, 1, Le, s3.u(), Kt};
  KL(c, k_dot, nblk(per * Kt), 256, 0, c->dc, d);
  CKL();
  {
    InvArgs a{s3.u(), Le, s4.u(), nullptr, nullptr};
    a.add1 = d_e;
    a.sh_src = d_src;
    a.sh_pow = d_pow.u();
    a.sh_D = I.D;
    const DevCtx &dc = c->dc;
    dim3 grid(Kt, Le), block(dc.N / 2 < 32 ? 32 : dc.N / 2);
    KL(c, k_inv, grid, block, (dc.N + dc.h) * 4, dc, a);
    CKL();
  }
  if ((rc = launch_crt(c, s4.u(), Le, CRT_REDUCE_Q, d_b, I.W, Kt))) return rc;
  // A' = Reduce(-A)
  CK(cudaMemcpyAsync(d_An, d_A, Kt * polyw * 4, cudaMemcpyDeviceToDevice, c->stream));
  return fhesi_ct_mul_scalar_dev(c, d_An, -1, Kt, 1);
}
static int key_build_from_device(fhesi_ctx *c, const u32 *d_polys, uint32_t parts, fhesi_key **out);

int fhesi_keygen_batch(fhesi_ctx *c, uint32_t M, const uint32_t *parts, const int32_t *h_src, const int32_t *h_t,
                       const uint32_t *h_A, const int32_t *h_e, fhesi_ksw **out, uint32_t *h_b_out, uint32_t *h_A_out,
                       fhesi_key **pk_out, uint32_t *h_pk_out) {
  if (!c || !h_t || !h_A || !h_e || (M && (!parts || !h_src || !out)) || (!M && !pk_out))
    return fail(FHESI_ERR_INVALID, "fhesi_keygen_batch: bad argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const u32 n = I.n;
  u32 rows = 0;
  for (u32 m = 0; m < M; ++m) {
    if (parts[m] < 1 || parts[m] > 3) return fail(FHESI_ERR_INVALID, "fhesi_keygen_batch: parts must be 1..3");
    rows += parts[m];
  }
  const u32 Km = rows * I.D, Kt = Km + (pk_out ? 1 : 0);  // the public key is the last entry, with a zero source row
  const size_t polyw = (size_t)n * I.W;
  static const bool timing = getenv("FHESI_TIMING") && atoi(getenv("FHESI_TIMING")) > 0;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double T0 = now();
  PoolTmp d_A(c), d_An(c), d_b(c), d_e(c), d_src(c), d_t(c);
  int rc = 0;
  if ((rc = d_A.alloc(Kt * polyw * 4)) || (rc = d_An.alloc(Kt * polyw * 4)) || (rc = d_b.alloc(Kt * polyw * 4)) ||
      (rc = d_e.alloc((size_t)Kt * n * 4)) || (rc = d_src.alloc((size_t)(rows + 1) * n * 4)) || (rc = d_t.alloc(n * 4)))
    return rc;
  CK(cudaMemcpyAsync(d_A.u(), h_A, Kt * polyw * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d_e.u(), h_e, (size_t)Kt * n * 4, cudaMemcpyHostToDevice, c->stream));
  if (rows) CK(cudaMemcpyAsync(d_src.u(), h_src, (size_t)rows * n * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemsetAsync(d_src.u() + (size_t)rows * n, 0, n * 4, c->stream));
  CK(cudaMemcpyAsync(d_t.u(), h_t, n * 4, cudaMemcpyHostToDevice, c->stream));
  if (timing) cudaStreamSynchronize(c->stream);
  const double T1 = now();
  if ((rc = keygen_entries(c, Kt, d_A.u(), (const int *)d_e.u(), (const int *)d_src.u(), (const int *)d_t.u(),
                           d_b.u(), d_An.u())))
    return rc;
  if (timing) cudaStreamSynchronize(c->stream);
  const double T2 = now();
  if (M && (rc = ksw_build_many(c, d_b.u(), d_An.u(), M, parts, out))) return rc;
  PoolTmp d_pk(c);
  if (pk_out) {  // publicKey = (c0, c1') (FHE-SI.cpp:59-61)
    if ((rc = d_pk.alloc(2 * polyw * 4))) return rc;
    CK(cudaMemcpyAsync(d_pk.u(), d_b.u() + (size_t)Km * polyw, polyw * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(d_pk.u() + polyw, d_An.u() + (size_t)Km * polyw, polyw * 4, cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = key_build_from_device(c, d_pk.u(), 2, pk_out))) return rc;
    if (h_pk_out) CK(cudaMemcpyAsync(h_pk_out, d_pk.u(), 2 * polyw * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if (h_b_out && Km) CK(cudaMemcpyAsync(h_b_out, d_b.u(), Km * polyw * 4, cudaMemcpyDeviceToHost, c->stream));
  if (h_A_out && Km) CK(cudaMemcpyAsync(h_A_out, d_An.u(), Km * polyw * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));  // the caller's host buffers are consumed / filled on return
  if (timing)
    fprintf(stderr, "fhesi_keygen_batch: alloc+upload %.4f  entries %.4f  key images + download %.4f s (%u entries)\n",
            T1 - T0, T2 - T1, now() - T2, Kt);
  return 0;
}
int fhesi_ksw_generate(fhesi_ctx *c, const int32_t *h_src, const int32_t *h_t, const uint32_t *h_A,
                       const int32_t *h_e, uint32_t parts, fhesi_ksw **out, uint32_t *h_b_out,
                       uint32_t *h_A_out) {
  if (!c || !h_src || !h_t || !h_A || !h_e || !out || parts < 1 || parts > 3)
    return fail(FHESI_ERR_INVALID, "fhesi_ksw_generate: bad argument");
  return fhesi_keygen_batch(c, 1, &parts, h_src, h_t, h_A, h_e, out, h_b_out, h_A_out, nullptr, nullptr);
}
void fhesi_ksw_destroy(fhesi_ksw *k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);  // (the host pipeline's lane streams are drained by its own call)
  if (--k->blk->refs == 0) {
    fhesi_free(k->ctx, k->blk->ptr);
    delete k->blk;
  }
  delete k;
}
static int key_build_from_device(fhesi_ctx *c, const u32 *d_polys, uint32_t parts, fhesi_key **out) {
  const fhesi_info &I = c->info;
  DevTmp t_key;
  PoolTmp t_tmp(c);
  int rc = t_tmp.alloc((size_t)parts * I.Le * I.N * 4);
  if (rc) return rc;
  CK(t_key.alloc((size_t)parts * I.Le * I.N * 4));
  rc = launch_fwd(c, d_polys, SRC_POLY, I.W, SC_KEYFORM, I.Le, t_tmp.u(), parts);
  if (rc) return rc;
  KL(c, k_transpose_key, nblk((size_t)parts * I.Le * I.N), 256, 0, t_tmp.u(), t_key.u(), parts, I.Le, I.N);
  CKL();
  *out = new fhesi_key{c, (u32 *)t_key.release(), parts};
  return 0;
}
int fhesi_key_create(fhesi_ctx *c, const uint32_t *h_polys, uint32_t parts, fhesi_key **out) {
  if (!c || !h_polys || !out || parts < 1 || parts > 3)
    return fail(FHESI_ERR_INVALID, "fhesi_key_create: bad argument");
  CK(cudaSetDevice(c->device));
  const size_t polyw = (size_t)c->info.n * c->info.W;
  PoolTmp t_in(c);
  int rc = t_in.alloc(parts * polyw * 4);
  if (rc) return rc;
  CK(cudaMemcpyAsync(t_in.u(), h_polys, parts * polyw * 4, cudaMemcpyHostToDevice, c->stream));
  if ((rc = key_build_from_device(c, t_in.u(), parts, out))) return rc;
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
void fhesi_key_destroy(fhesi_key *k) {
  if (!k) return;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  cudaFree(k->d_key);
  delete k;
}

// ---------------------------------------------------------------------------------------
// coefficient-domain ops
// ---------------------------------------------------------------------------------------
int fhesi_ct_add_dev(fhesi_ctx *c, uint32_t *io, const uint32_t *other, uint32_t parts, size_t count) {
  if (!c || !io || !other) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  size_t ncoef = count * parts * c->info.n;
  if (!ncoef) return 0;
  KL(c, k_ct_add, nblk(ncoef), 256, 0, c->dc, io, other, ncoef);
  CKL();
  return 0;
}
int fhesi_ct_sum_dev(fhesi_ctx *c, const uint32_t *in, uint32_t *out, uint32_t parts, size_t count) {
  if (!c || !in || !out) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  size_t per = (size_t)parts * c->info.n;
  KL(c, k_ct_sum, nblk(per, 64), 64, 0, c->dc, in, out, per, (u32)count);
  CKL();
  return 0;
}
int fhesi_ct_mul_scalar_dev(fhesi_ctx *c, uint32_t *io, int64_t l, uint32_t parts, size_t count) {
  if (!c || !io) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  size_t ncoef = count * parts * c->info.n;
  if (!ncoef) return 0;
  u64 mag = l < 0 ? (u64)0 - (u64)l : (u64)l;
  KL(c, k_ct_mul_scalar, nblk(ncoef, 128), 128, 0, c->dc, io, mag, l < 0, ncoef);
  CKL();
  return 0;
}
int fhesi_reduce_wide_dev(fhesi_ctx *c, const uint32_t *in, uint32_t Win, uint32_t *out,
                          uint32_t parts, size_t count) {
  if (!c || !in || !out || Win < c->info.W) return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  size_t ncoef = count * parts * c->info.n;
  if (!ncoef) return 0;
  KL(c, k_reduce_wide, nblk(ncoef), 256, 0, c->dc, in, Win, out, ncoef);
  CKL();
  return 0;
}
// device table of the signed index permutation of X -> X^k (one per Galois element, kept for reuse)
static int automorph_table(fhesi_ctx *c, uint32_t k, u32 **out) {
  const fhesi_info &I = c->info;
  u32 *&d_tab = c->automorph_tabs[k % I.m];
  if (!d_tab) {
    std::vector<u32> tab;
    if (c->dc.h) {
      const u32 h = c->dc.h;
      tab.assign(h, 0xFFFFFFFFu);
      for (u32 i = 0; i < I.n; ++i) {
        u32 e = (u32)(((u64)i * k) % I.m), neg = 0;
        if (e >= h) { e -= h; neg = 1; }
        tab[e] = (i << 1) | neg;
      }
    } else {
      // general m: X^i -> X^(i k mod m), positions >= n expanded through X^(n+j) mod Phi_m; rows by output index
      std::vector<std::vector<std::pair<u32, int>>> rows(I.n);
      for (u32 i = 0; i < I.n; ++i) {
        const u32 e = (u32)(((u64)i * k) % I.m);
        if (e < I.n) rows[e].push_back(std::make_pair(i, 1));
        else
          for (auto &t : c->red_cols[e - I.n]) rows[t.first].push_back(std::make_pair(i, t.second));
      }
      tab.assign(I.n + 1, 0);
      for (u32 i = 0; i < I.n; ++i) tab[i + 1] = tab[i] + (u32)rows[i].size();
      for (u32 i = 0; i < I.n; ++i)
        for (auto &t : rows[i]) {
          tab.push_back(t.first);
          tab.push_back((u32)t.second);
        }
    }
    void *pt = nullptr;
    size_t cap = 0;
    CK(cached_malloc(c->device, &pt, tab.size() * 4, &cap));
    c->tables.push_back(std::make_pair(pt, cap));
    d_tab = (u32 *)pt;
    CK(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // tab is a stack-lifetime host buffer
  }
  *out = d_tab;
  return 0;
}
// ---- tensor-form branches of the plaintext / automorphism operators (Ciphertext.cpp:153-159,252-256,269-273)
int fhesi_tprod_add_poly_dev(fhesi_ctx *c, uint32_t *tprod, uint32_t parts, const uint32_t *poly, uint32_t Win,
                             size_t count) {
  if (!c || !tprod || !poly || !parts) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  if (!Win || Win > I.W + 1) return fail(FHESI_ERR_INVALID, "fhesi_tprod_add_poly_dev: 1 <= Win <= W + 1");
  CK(cudaSetDevice(c->device));
  if (!count) return 0;
  const size_t per = (size_t)I.Lt * I.N;
  PoolTmp img(c);
  int rc = img.alloc(count * per * 4);
  if (rc) return rc;
  // DoubleCRT(poly) in the tprod convention (images carry 1/N)
  if ((rc = launch_fwd(c, poly, SRC_POLY, Win, SC_NINV, I.Lt, img.u(), count))) return rc;
  for (size_t b = 0; b < count; ++b) {  // tProd[0] += ...
    KL(c, k_tprod_add, nblk(per), 256, 0, c->dc, tprod + b * parts * per, img.u() + b * per, I.Lt, per);
    CKL();
  }
  return 0;
}
int fhesi_tprod_mul_poly_dev(fhesi_ctx *c, uint32_t *tprod, uint32_t parts, const uint32_t *poly, uint32_t Win,
                             size_t count) {
  if (!c || !tprod || !poly || !parts) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  if (!Win || Win > I.W + 1) return fail(FHESI_ERR_INVALID, "fhesi_tprod_mul_poly_dev: 1 <= Win <= W + 1");
  CK(cudaSetDevice(c->device));
  const size_t per = (size_t)I.Lt * I.N, total = count * parts * per;
  if (!total) return 0;
  PoolTmp img(c);
  int rc = img.alloc(per * 4);
  if (rc) return rc;
  if ((rc = launch_fwd(c, poly, SRC_POLY, Win, SC_MONT, I.Lt, img.u(), 1))) return rc;
  KL(c, k_tprod_mul_img, nblk(total), 256, 0, c->dc, tprod, img.u(), I.Lt, total);
  CKL();
  return 0;
}
int fhesi_tprod_automorph_dev(fhesi_ctx *c, const uint32_t *in, uint32_t parts, uint32_t k, uint32_t *out,
                              size_t count) {
  if (!c || !in || !out || !parts) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  u64 a = k % I.m, b = I.m;
  while (b) { u64 t = a % b; a = b; b = t; }
  if (a != 1) return fail(FHESI_ERR_INVALID, "DoubleCRT::automorph: k not in Zm*");
  CK(cudaSetDevice(c->device));
  const size_t npolys = count * parts;
  if (!npolys) return 0;
  u32 *d_tab = nullptr;
  int rc = automorph_table(c, k, &d_tab);
  if (rc) return rc;
  PoolTmp r0(c), r1(c);
  const size_t words = npolys * I.Lt * I.n;
  if ((rc = r0.alloc(words * 4)) || (rc = r1.alloc(words * 4))) return rc;
  // transform domain -> coefficient residues (Phi_m fold included) -> signed permutation -> back
  if ((rc = launch_inv(c, in, I.Lt, r0.u(), npolys))) return rc;
  if (c->dc.h) KL(c, k_automorph_res, nblk(words), 256, 0, c->dc, r0.u(), d_tab, r1.u(), I.Lt, npolys);
  else KL(c, k_automorph_res_csr, nblk(words), 256, 0, c->dc, r0.u(), d_tab, r1.u(), I.Lt, npolys);
  CKL();
  return launch_fwd(c, r1.u(), SRC_RES, 0, SC_NINV, I.Lt, out, npolys);
}
int fhesi_embed_slots_dev(fhesi_ctx *c, const uint32_t *basis, uint32_t nslots, const uint32_t *vals,
                          uint32_t *msg, size_t count) {
  if (!c || !basis || !vals || !msg) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  if (!nslots || nslots > 4096 || I.p >= (1ull << 26))
    return fail(FHESI_ERR_INVALID, "fhesi_embed_slots_dev: needs 1 <= nslots <= 4096 and p < 2^26");
  CK(cudaSetDevice(c->device));
  if (!count) return 0;
  for (size_t off = 0; off < count; off += 65535) {  // gridDim.y limit
    const size_t cnt = count - off < 65535 ? count - off : 65535;
    dim3 grid((I.n + 127) / 128, (unsigned)cnt);
    KL(c, k_embed_slots, grid, 128, nslots * 4, basis, vals + off * nslots, msg + off * I.n, nslots, I.n, (u32)I.p);
    CKL();
  }
  return 0;
}
int fhesi_decode_slots_dev(fhesi_ctx *c, const uint32_t *vander, uint32_t nslots, const uint32_t *msg,
                           uint32_t *vals, size_t count) {
  if (!c || !vander || !msg || !vals) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  if (!nslots || I.n > 4096 || I.p >= (1ull << 26))
    return fail(FHESI_ERR_INVALID, "fhesi_decode_slots_dev: needs nslots >= 1, n <= 4096 and p < 2^26");
  CK(cudaSetDevice(c->device));
  if (!count) return 0;
  for (size_t off = 0; off < count; off += 65535) {  // gridDim.y limit
    const size_t cnt = count - off < 65535 ? count - off : 65535;
    dim3 grid((nslots + 127) / 128, (unsigned)cnt);
    // the embedding kernel with the roles swapped: n "slots" in (the coefficients), nslots values out
    KL(c, k_embed_slots, grid, 128, I.n * 4, vander, msg + off * I.n, vals + off * nslots, I.n, nslots, (u32)I.p);
    CKL();
  }
  return 0;
}
int fhesi_ct_automorph_dev(fhesi_ctx *c, const uint32_t *in, uint32_t parts, uint32_t k,
                           uint32_t *out, size_t count) {
  if (!c || !in || !out) return fail(FHESI_ERR_INVALID, "null argument");
  const fhesi_info &I = c->info;
  u64 a = k % I.m, b = I.m;
  while (b) { u64 t = a % b; a = b; b = t; }
  if (a != 1) return fail(FHESI_ERR_INVALID, "DoubleCRT::automorph: k not in Zm*");
  CK(cudaSetDevice(c->device));
  size_t npolys = count * parts;
  u32 *d_tab = nullptr;
  int rc = automorph_table(c, k, &d_tab);
  if (rc) return rc;
  if (npolys && c->dc.h) KL(c, k_automorph, nblk(npolys * I.n, 128), 128, 0, c->dc, in, d_tab, out, npolys);
  else if (npolys) KL(c, k_automorph_csr, nblk(npolys * I.n, 128), 128, 0, c->dc, in, d_tab, out, npolys);
  CKL();
  return 0;
}

// every part times one plaintext polynomial: key-form transform of the parts, plain transform of
// the plaintext, pointwise product, inverse transform, CRT + Reduce
int fhesi_ct_mul_plain_dev(fhesi_ctx *c, uint32_t *io, const uint32_t *plain, uint32_t parts, size_t count) {
  if (!c || !io || !plain || parts < 1 || parts > 3) return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const u32 Le = I.Le;
  const size_t per = (size_t)Le * I.N, CH = c->chunk;
  size_t n1 = al(CH * parts * per), n2 = al(per), n3 = al(CH * parts * Le * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (n1 + n2 + n1 + n3) * 4, &s);
  if (rc) return rc;
  u32 *sA = s, *sP = s + n1, *sO = sP + n2, *sR = sO + n1;
  if ((rc = launch_fwd(c, plain, SRC_U32, 0, SC_NONE, Le, sP, 1))) return rc;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    u32 *cur = io + off * parts * I.n * I.W;
    if ((rc = launch_fwd(c, cur, SRC_POLY, I.W, SC_KEYFORM, Le, sA, cnt * parts))) return rc;
    // out[q] = A[q] * P : a 1x1 "tensor" per part, the plaintext image broadcast (count 1 each)
    for (size_t q = 0; q < cnt * parts; ++q) {
      TensArgs t{sA + q * per, sP, 1, 1, Le, sO + q * per, 1, 0};
      KL(c, k_tensor_pw, nblk(per), 256, 0, c->dc, t);
      CKL();
    }
    if ((rc = launch_inv(c, sO, Le, sR, cnt * parts))) return rc;
    if ((rc = launch_crt(c, sR, Le, CRT_REDUCE_Q, cur, I.W, cnt * parts))) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// tensor-form ops
// ---------------------------------------------------------------------------------------
int fhesi_ct_tensor_dev(fhesi_ctx *c, const uint32_t *a, uint32_t pa, const uint32_t *b, uint32_t pb,
                        uint32_t *tprod, size_t count, int accumulate) {
  if (!c || !a || !b || !tprod || pa < 1 || pa > 3 || pb < 1 || pb > 3)
    return fail(FHESI_ERR_INVALID, "fhesi_ct_tensor_dev: bad argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const size_t per = (size_t)I.Lt * I.N;
  const u32 po = pa + pb - 1;
  const size_t CH = c->chunk;
  if (c->use_fused && pa == 2 && pb == 2) {
    const size_t ctw = (size_t)I.n * I.W;
    size_t FC = accumulate ? 512 : c->fused_chunk;
    if (count < FC) FC = count ? count : 1;
    u32 *sres = nullptr;
    const size_t nres = al(FC * 4 * I.Lt * I.n), ntp = accumulate ? al(FC * 3 * per) : 0;
    int rc2 = scratch(c, (nres + ntp) * 4, &sres);
    if (rc2) return rc2;
    if (accumulate) CK(cudaMemsetAsync(tprod, 0, 3 * per * 4, c->stream));
    for (size_t off = 0; off < count; off += FC) {
      size_t cnt = count - off < FC ? count - off : FC;
      u32 *dst = accumulate ? sres + nres : tprod + off * 3 * per;
      if ((rc2 = launch_fused_tensor(c, a + off * 2 * ctw, b + off * 2 * ctw, sres, dst, cnt, 1))) return rc2;
      if (accumulate) {  // fold the chunk's tprods into the single running sum
        KL(c, k_tprod_batch_sum, nblk(3 * per), 256, 0, c->dc, dst, (u32)cnt, I.Lt, 3 * per, tprod);
        CKL();
      }
    }
    return 0;
  }
  u32 *s = nullptr;
  size_t na = al(CH * pa * per), nb_ = al(CH * pb * per), nacc = al(po * per);
  int rc = scratch(c, (na + nb_ + nacc) * 4, &s);
  if (rc) return rc;
  u32 *sA = s, *sB = s + na, *sAcc = sB + nb_;
  if (accumulate) CK(cudaMemsetAsync(tprod, 0, po * per * 4, c->stream));
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    if ((rc = launch_fwd(c, a + off * pa * I.n * I.W, SRC_POLY, I.W, SC_TENSOR, I.Lt, sA, cnt * pa))) return rc;
    if ((rc = launch_fwd(c, b + off * pb * I.n * I.W, SRC_POLY, I.W, SC_NONE, I.Lt, sB, cnt * pb))) return rc;
    TensArgs t{sA, sB, pa, pb, I.Lt, accumulate ? sAcc : tprod + off * po * per, (u32)cnt, accumulate};
    size_t total = accumulate ? per : per * cnt;
    KL(c, k_tensor_pw, nblk(total), 256, 0, c->dc, t);
    CKL();
    if (accumulate) {
      KL(c, k_tprod_add, nblk(po * per), 256, 0, c->dc, tprod, sAcc, I.Lt, po * per);
      CKL();
    }
  }
  return 0;
}
int fhesi_tprod_add_dev(fhesi_ctx *c, uint32_t *io, const uint32_t *other, uint32_t parts, size_t count) {
  if (!c || !io || !other) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  size_t total = count * parts * c->info.Lt * c->info.N;
  if (!total) return 0;
  KL(c, k_tprod_add, nblk(total), 256, 0, c->dc, io, other, c->info.Lt, total);
  CKL();
  return 0;
}
int fhesi_tprod_mul_scalar_dev(fhesi_ctx *c, uint32_t *io, int64_t l, uint32_t parts, size_t count) {
  if (!c || !io) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  size_t total = count * parts * I.Lt * I.N;
  if (!total) return 0;
  std::vector<u32> sc(I.Lt);
  for (u32 i = 0; i < I.Lt; ++i) {
    i64 q = I.primes[i];
    i64 r = l % q;
    if (r < 0) r += q;
    sc[i] = (u32)h_mulmod((u64)r, c->h_pc[i].r1, (u64)q);
  }
  PoolTmp t_sc(c);  // pooled, stream-ordered: no cudaMalloc / cudaFree / synchronisation per call
  int rc = t_sc.alloc(I.Lt * 4);
  if (rc) return rc;
  u32 *d_sc = t_sc.u();
  // `sc` is pageable: the copy has consumed it when cudaMemcpyAsync returns
  CK(cudaMemcpyAsync(d_sc, sc.data(), I.Lt * 4, cudaMemcpyHostToDevice, c->stream));
  KL(c, k_tprod_mul_scalar, nblk(total), 256, 0, c->dc, io, d_sc, I.Lt, total);
  CKL();
  return 0;
}
int fhesi_tprod_reduce_gathered_dev(fhesi_ctx *c, const uint32_t *g, uint32_t world, uint32_t parts,
                                    uint32_t *out) {
  if (!c || !g || !out || !world) return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  size_t per = (size_t)parts * c->info.Lt * c->info.N;
  KL(c, k_tprod_reduce_world, nblk(per), 256, 0, c->dc, g, world, c->info.Lt, per, out);
  CKL();
  return 0;
}
int fhesi_scaledown_dev(fhesi_ctx *c, const uint32_t *tprod, uint32_t parts, uint32_t *out, size_t count) {
  if (!c || !tprod || !out) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const size_t CH = c->chunk;
  u32 *s = nullptr;
  int rc = scratch(c, al(CH * parts * I.Lt * I.n) * 4, &s);
  if (rc) return rc;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    if ((rc = launch_inv(c, tprod + off * parts * I.Lt * I.N, I.Lt, s, cnt * parts))) return rc;
    if ((rc = launch_crt(c, s, I.Lt, CRT_SCALEDOWN, out + off * parts * I.n * I.W, I.W, cnt * parts))) return rc;
  }
  return 0;
}

// generic key switch on already scaled-down parts
static int keyswitch_generic(fhesi_ctx *c, const fhesi_ksw *ksw, const u32 *in, u32 *out, size_t count) {
  const fhesi_info &I = c->info;
  const u32 K = ksw->parts * I.D, Lk = I.Lk;
  const size_t per = (size_t)Lk * I.N;
  const size_t CH = c->chunk;
  size_t nd = al(CH * K * per), no = al(CH * 2 * per), nr = al(CH * 2 * Lk * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (nd + no + nr) * 4, &s);
  if (rc) return rc;
  u32 *sD = s, *sO = s + nd, *sR = sO + no;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    if ((rc = launch_fwd(c, in + off * ksw->parts * I.n * I.W, SRC_DIGIT, I.W, SC_NONE, Lk, sD, cnt * K))) return rc;
    DotArgs d{sD, ksw->d_key, K, 2, Lk, sO, (u32)cnt};
    KL(c, k_dot, nblk(per * cnt), 256, 0, c->dc, d);
    CKL();
    if ((rc = launch_inv(c, sO, Lk, sR, cnt * 2))) return rc;
    if ((rc = launch_crt(c, sR, Lk, CRT_REDUCE_Q, out + off * 2 * I.n * I.W, I.W, cnt * 2))) return rc;
  }
  return 0;
}
// ---------------------------------------------------------------------------------------
// fused N=1024 path (kernels_fused.cuh)
// ---------------------------------------------------------------------------------------
static int fused_ks_from_digits(fhesi_ctx *c, const fhesi_ksw *ksw, const u32 *digits, u32 *res, u32 *out,
                                size_t cnt) {
  const fhesi_info &I = c->info;
  if (I.Ls && ksw->d_key_split) {  // split-key path: res holds [cnt][4][Ls][n]
    FusedKsArgs k{digits, ksw->d_key_split, res, ksw->parts * I.D, I.Ls, (u32)cnt};
    if (I.N == FN) {
      dim3 grid(I.Ls, (unsigned)((cnt + KSS - 1) / KSS));
      const auto kern = c->dc.h ? k_fused_keyswitch_split<false> : k_fused_keyswitch_split<true>;
      KLN(c, "k_fused_keyswitch_split", kern, grid, KSS * 128, KSS_SMEM_WORDS * 4, c->dc, k);
    } else {
      dim3 grid(I.Ls, (unsigned)((cnt + KSS2 - 1) / KSS2));
      const auto kern = c->dc.h ? k_fused_keyswitch_split_2k<false> : k_fused_keyswitch_split_2k<true>;
      KLN(c, "k_fused_keyswitch_split_2k", kern, grid, KSS2 * T2K, KSS2K_SMEM_WORDS * 4, c->dc, k);
    }
    CKL();
    return launch_crt_split(c, res, I.Ls, out, cnt * 2);
  }
  const bool tfree = c->tfree && ksw->d_key_bal;
  FusedKsArgs k{digits, tfree ? ksw->d_key_bal : ksw->d_key, res, ksw->parts * I.D, I.Lk, (u32)cnt};
  dim3 grid(I.Lk, (unsigned)((cnt + KSG - 1) / KSG));
  const auto kern = tfree ? (c->dc.h ? k_fused_keyswitch<true, false> : k_fused_keyswitch<true, true>)
                          : (c->dc.h ? k_fused_keyswitch<false, false> : k_fused_keyswitch<false, true>);
  KLN(c, tfree ? "k_fused_keyswitch<true>" : "k_fused_keyswitch<false>", kern, grid, KSG * 128, KS_SMEM_WORDS * 4, c->dc, k);
  CKL();
  return launch_crt(c, res, I.Lk, CRT_REDUCE_Q, out, I.W, cnt * 2);
}
static int fused_keyswitch(fhesi_ctx *c, const fhesi_ksw *ksw, const u32 *in, u32 *out, size_t count) {
  const fhesi_info &I = c->info;
  const u32 K = ksw->parts * I.D;
  const size_t CH = c->fused_chunk;
  const size_t rw = (size_t)(2 * I.Lk > 4 * I.Ls ? 2 * I.Lk : 4 * I.Ls);
  size_t nd = al(CH * K * I.n), nr = al(CH * rw * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (nd + nr) * 4, &s);
  if (rc) return rc;
  u32 *sD = s, *sR = s + nd;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    size_t npolys = cnt * ksw->parts;
    KL(c, k_digits, nblk(npolys * I.n), 256, 0, c->dc, in + off * ksw->parts * I.n * I.W, sD, npolys);
    CKL();
    if ((rc = fused_ks_from_digits(c, ksw, sD, sR, out + off * 2 * I.n * I.W, cnt))) return rc;
  }
  return 0;
}
// tmp >>= k; ApplyKeySwitch(tmp) with the rotation folded into the digit extraction
static int fused_rotate_keyswitch(fhesi_ctx *c, const fhesi_ksw *ksw, const u32 *in, const u32 *d_tab, u32 *out,
                                  size_t count) {
  const fhesi_info &I = c->info;
  const u32 K = ksw->parts * I.D;
  const size_t CH = c->fused_chunk;
  const size_t rw = (size_t)(2 * I.Lk > 4 * I.Ls ? 2 * I.Lk : 4 * I.Ls);
  size_t nd = al(CH * K * I.n), nr = al(CH * rw * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (nd + nr) * 4, &s);
  if (rc) return rc;
  u32 *sD = s, *sR = s + nd;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    size_t npolys = cnt * ksw->parts;
    KL(c, k_digits_automorph, nblk(npolys * I.n), 256, 0, c->dc, in + off * ksw->parts * I.n * I.W, d_tab, sD, npolys);
    CKL();
    if ((rc = fused_ks_from_digits(c, ksw, sD, sR, out + off * 2 * I.n * I.W, cnt))) return rc;
  }
  return 0;
}
static int fused_mult_relin(fhesi_ctx *c, const fhesi_ksw *ksw, const u32 *a, const u32 *b, u32 *out,
                            size_t count) {
  const fhesi_info &I = c->info;
  const u32 K = 3 * I.D;
  size_t CH = c->fused_chunk;
  const size_t ctw = (size_t)I.n * I.W;
  if (count < CH) CH = count ? count : 1;
  const size_t rw = (size_t)(2 * I.Lk > 4 * I.Ls ? 2 * I.Lk : 4 * I.Ls);
  size_t n0 = al(CH * 4 * I.Lt * I.n), n1 = al(CH * 3 * I.Lt * I.n), nd = al(CH * K * I.n),
         n2 = al(CH * rw * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (n0 + n1 + nd + n2) * 4, &s);
  if (rc) return rc;
  u32 *sR0 = s, *sR1 = s + n0, *sD = sR1 + n1, *sR2 = sD + nd;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    if ((rc = launch_fused_tensor(c, a + off * 2 * ctw, b + off * 2 * ctw, sR0, sR1, cnt, 0))) return rc;
    if ((rc = launch_crt(c, sR1, I.Lt, CRT_SCALEDOWN_DIGITS, sD, I.W, cnt * 3))) return rc;
    if ((rc = fused_ks_from_digits(c, ksw, sD, sR2, out + off * 2 * ctw, cnt))) return rc;
  }
  return 0;
}

int fhesi_keyswitch_dev(fhesi_ctx *c, const fhesi_ksw *ksw, const uint32_t *in, uint32_t *out, size_t count) {
  if (!c || !ksw || !in || !out) return fail(FHESI_ERR_INVALID, "null argument");
  if (ksw->ctx != c) return fail(FHESI_ERR_INVALID, "key-switch matrix belongs to another context");
  CK(cudaSetDevice(c->device));
  if (c->use_fused) return fused_keyswitch(c, ksw, in, out, count);
  return keyswitch_generic(c, ksw, in, out, count);
}

int fhesi_rotate_keyswitch_dev(fhesi_ctx *c, const fhesi_ksw *ksw, const uint32_t *in, uint32_t k, uint32_t *out,
                               size_t count) {
  if (!c || !ksw || !in || !out) return fail(FHESI_ERR_INVALID, "null argument");
  if (ksw->ctx != c) return fail(FHESI_ERR_INVALID, "key-switch matrix belongs to another context");
  if (ksw->parts != 2) return fail(FHESI_ERR_INVALID, "rotation needs the (1, s(X^k)) -> s matrix (2 source parts)");
  const fhesi_info &I = c->info;
  u64 a = k % I.m, b = I.m;
  while (b) { u64 t = a % b; a = b; b = t; }
  if (a != 1) return fail(FHESI_ERR_INVALID, "DoubleCRT::automorph: k not in Zm*");
  CK(cudaSetDevice(c->device));
  if (!count) return 0;
  u32 *d_tab = nullptr;
  int rc = automorph_table(c, k, &d_tab);
  if (rc) return rc;
  if (c->use_fused && c->dc.h) return fused_rotate_keyswitch(c, ksw, in, d_tab, out, count);
  // generic kernels, or general m (the rotation is a sparse matrix, not a signed permutation): three steps
  PoolTmp wide(c), red(c);
  const size_t npolys = count * 2;
  if ((rc = wide.alloc(npolys * I.n * (I.W + 1) * 4)) || (rc = red.alloc(npolys * I.n * I.W * 4))) return rc;
  if (c->dc.h) KL(c, k_automorph, nblk(npolys * I.n, 128), 128, 0, c->dc, in, d_tab, wide.u(), npolys);
  else KL(c, k_automorph_csr, nblk(npolys * I.n, 128), 128, 0, c->dc, in, d_tab, wide.u(), npolys);
  CKL();
  if ((rc = fhesi_reduce_wide_dev(c, wide.u(), I.W + 1, red.u(), 2, count))) return rc;
  return fhesi_keyswitch_dev(c, ksw, red.u(), out, count);
}

int fhesi_mult_relin_dev(fhesi_ctx *c, const fhesi_ksw *ksw, const uint32_t *a, const uint32_t *b,
                         uint32_t *out, size_t count) {
  if (!c || !ksw || !a || !b || !out) return fail(FHESI_ERR_INVALID, "null argument");
  if (ksw->ctx != c) return fail(FHESI_ERR_INVALID, "key-switch matrix belongs to another context");
  if (ksw->parts != 3) return fail(FHESI_ERR_INVALID, "mult_relin needs the s^2 -> s matrix (3 source parts)");
  CK(cudaSetDevice(c->device));
  if (c->use_fused) return fused_mult_relin(c, ksw, a, b, out, count);
  const fhesi_info &I = c->info;
  // generic composition: tensor -> ScaleDown -> key switch, chunked through HBM scratch
  const size_t CH = c->chunk;
  const size_t ctw = (size_t)I.n * I.W;
  const size_t nt = al(CH * 3 * I.Lt * I.N), nc = al(CH * 3 * ctw);
  if (c->work.cap < (nt + nc) * 4) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->work.ptr) CK(cudaFree(c->work.ptr));
    c->work.ptr = nullptr;
    c->work.cap = 0;
    CK(cached_malloc(c->device, &c->work.ptr, (nt + nc) * 4, &c->work.cap));
  }
  u32 *d_t = (u32 *)c->work.ptr, *d_c = d_t + nt;
  int rc = 0;
  for (size_t off = 0; off < count && !rc; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    rc = fhesi_ct_tensor_dev(c, a + off * 2 * ctw, 2, b + off * 2 * ctw, 2, d_t, cnt, 0);
    if (!rc) rc = fhesi_scaledown_dev(c, d_t, 3, d_c, cnt);
    if (!rc) rc = keyswitch_generic(c, ksw, d_c, out + off * 2 * ctw, cnt);
  }
  return rc;
}
// Host-buffer entry point.  The batch is cut into pipeline chunks: chunk i+1 is uploaded on a
// copy stream while chunk i computes on the context's stream and chunk i-1 is downloaded on a
// second copy stream, so PCIe in both directions overlaps the kernels.
// The same without the final wait: the call returns when everything is enqueued, and the NEXT call's uploads and
// kernels overlap this call's last kernels and downloads (two staging halves, used alternately).  h_a and h_b must
// stay untouched and h_out unread until fhesi_sync_all (or a later blocking fhesi_mult_relin_host) returns; a caller
// that keeps several calls in flight gives each its own h_out.
int fhesi_mult_relin_host_async(fhesi_ctx *c, const fhesi_ksw *ksw, const uint32_t *h_a, const uint32_t *h_b,
                                uint32_t *h_out, size_t count) {
  if (!c || !ksw || !h_a || !h_b || !h_out) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  if (!count) return 0;
  const size_t ctb = fhesi_ct_bytes(c, 2), bytes = count * ctb;
  if (c->stage.cap < 2 * (3 * bytes + 64)) {  // grow-only staging area (two halves), reused across calls
    int rc = fhesi_sync_all(c);
    if (rc) return rc;
    if (c->stage.ptr) CK(cudaFree(c->stage.ptr));
    c->stage.ptr = nullptr;
    c->stage.cap = 0;
    c->half_used[0] = c->half_used[1] = false;
    CK(cached_malloc(c->device, &c->stage.ptr, 2 * (3 * bytes + 64), &c->stage.cap));
  }
  const unsigned half = c->host_calls++ & 1;
  for (int h = 0; h < 2; ++h)
    if (!c->half_done[h]) CK(cudaEventCreateWithFlags(&c->half_done[h], cudaEventDisableTiming));
  if (!c->h2d_stream) {
    CK(cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->lane_stream, cudaStreamNonBlocking));
  }
  // pipeline chunk: ~21 chunks per call, a multiple of 12, between 252 (below that the grids are too few waves
  // deep) and 384 (above, the compute stream's lag behind the upload -- one chunk -- costs more than the better
  // kernel efficiency returns).  Measured on a B200 behind PCIe gen5 with two compute lanes, batch 8192, with the
  // round-2 kernels (8.6 ms of compute against 10.5 ms of copies; scripts/gpu/e2e_sweep.py): half chunk + 384s + half
  // chunk 730 k ops/s, 384s alone 721 k, 288s 719 k, 576s with quarter and half chunks at the ends (the round-1
  // optimum, when compute took 10 ms) 704 k, 768s 707 k, 192s 697 k
  size_t PC = c->pipe_chunk;
  if (!PC) {
    PC = ((count / 21 + 11) / 12) * 12;
    if (PC < 252) PC = 252;
    if (PC > 384) PC = 384;
  }
  // chunk schedule: full chunks in the middle, a half chunk at either end when the batch is long enough -- the
  // first upload and the last compute + download are the only parts of the call that do not overlap anything
  // (FHESI_PIPE_TAPER=0 turns the taper off)
  std::vector<size_t> sizes;
  if (const char *sched = getenv("FHESI_PIPE_SCHED")) {
    // explicit schedule for tuning runs: "192,384,*1536,384,192" -- head sizes, one repeated
    // middle size (marked *), tail sizes; anything that does not fit the batch is dropped
    std::vector<size_t> head, tail;
    size_t mid = PC;
    bool after = false;
    for (const char *q = sched; *q;) {
      const bool star = *q == '*';
      if (star) ++q;
      const size_t v = strtoul(q, (char **)&q, 10);
      if (*q == ',') ++q;
      if (!v) break;
      if (star) mid = v, after = true;
      else (after ? tail : head).push_back(v);
    }
    size_t rest = count, tsum = 0;
    for (size_t v : tail) tsum += v;
    for (size_t v : head)
      if (rest > tsum + v) sizes.push_back(v), rest -= v;
    if (rest <= tsum) tail.clear(), tsum = 0;
    rest -= tsum;
    while (rest) {
      const size_t t = rest < mid ? rest : mid;
      sizes.push_back(t);
      rest -= t;
    }
    sizes.insert(sizes.end(), tail.begin(), tail.end());
  } else {
    size_t rest = count;
    const size_t h = ((PC / 2 + 11) / 12) * 12;
    const bool taper = c->pipe_taper && count >= 6 * PC;
    std::vector<size_t> tail;
    if (taper) {
      sizes.push_back(h);
      tail.push_back(h);
      rest -= 2 * h;
    }
    while (rest) {
      const size_t t = rest < PC ? rest : PC;
      sizes.push_back(t);
      rest -= t;
    }
    sizes.insert(sizes.end(), tail.begin(), tail.end());
  }
  const size_t nchunks = sizes.size();
  while (c->pipe_events.size() < 2 * nchunks) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->pipe_events.push_back(e);
  }
  char *d = (char *)c->stage.ptr + (size_t)half * (c->stage.cap / 2 / 64 * 64);
  char *da = d, *db = d + bytes, *dout = d + 2 * bytes;
  // the copy stream must not overtake work already queued on the compute stream, nor the last call that used
  // this staging half
  CK(cudaEventRecord(c->pipe_events[0], c->stream));
  CK(cudaStreamWaitEvent(c->h2d_stream, c->pipe_events[0], 0));
  CK(cudaStreamWaitEvent(c->lane_stream, c->pipe_events[0], 0));
  if (c->half_used[half]) CK(cudaStreamWaitEvent(c->h2d_stream, c->half_done[half], 0));
  // FHESI_PIPE_TRACE=1: per-chunk timeline (upload done / compute done / download done, ms from the
  // start of the call) on stderr -- a debugging aid, off by default
  static const bool trace = getenv("FHESI_PIPE_TRACE") && atoi(getenv("FHESI_PIPE_TRACE")) > 0;
  std::vector<cudaEvent_t> tev;
  if (trace) {
    tev.resize(3 * nchunks + 1);
    for (auto &e : tev) CK(cudaEventCreate(&e));
    CK(cudaEventRecord(tev[3 * nchunks], c->stream));
  }
  size_t off = 0;
  for (size_t ci = 0; ci < nchunks; ++ci) {
    const size_t cnt = sizes[ci];
    cudaEvent_t ev_in = trace ? tev[3 * ci] : c->pipe_events[2 * ci];
    cudaEvent_t ev_done = trace ? tev[3 * ci + 1] : c->pipe_events[2 * ci + 1];
    CK(cudaMemcpyAsync(da + off * ctb, (const char *)h_a + off * ctb, cnt * ctb, cudaMemcpyHostToDevice, c->h2d_stream));
    CK(cudaMemcpyAsync(db + off * ctb, (const char *)h_b + off * ctb, cnt * ctb, cudaMemcpyHostToDevice, c->h2d_stream));
    CK(cudaEventRecord(ev_in, c->h2d_stream));
    const bool lane1 = c->pipe_lanes > 1 && (ci & 1);
    if (lane1) std::swap(c->stream, c->lane_stream), std::swap(c->scratch, c->lane_scratch);
    int rc = cudaStreamWaitEvent(c->stream, ev_in, 0) == cudaSuccess ? 0 : fail(FHESI_ERR_CUDA, "cudaStreamWaitEvent");
    if (!rc)
      rc = fhesi_mult_relin_dev(c, ksw, (const u32 *)(da + off * ctb), (const u32 *)(db + off * ctb),
                                (u32 *)(dout + off * ctb), cnt);
    if (!rc && cudaEventRecord(ev_done, c->stream) != cudaSuccess) rc = fail(FHESI_ERR_CUDA, "cudaEventRecord");
    if (lane1) std::swap(c->stream, c->lane_stream), std::swap(c->scratch, c->lane_scratch);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(c->d2h_stream, ev_done, 0));
    CK(cudaMemcpyAsync((char *)h_out + off * ctb, dout + off * ctb, cnt * ctb, cudaMemcpyDeviceToHost, c->d2h_stream));
    if (trace) CK(cudaEventRecord(tev[3 * ci + 2], c->d2h_stream));
    off += cnt;
  }
  if (trace) {
    CK(cudaStreamSynchronize(c->d2h_stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->lane_stream));
    for (size_t ci = 0; ci < nchunks; ++ci) {
      float t[3];
      for (int k = 0; k < 3; ++k) CK(cudaEventElapsedTime(&t[k], tev[3 * nchunks], tev[3 * ci + k]));
      fprintf(stderr, "pipe chunk %2zu size %4zu  h2d %.3f  compute %.3f  d2h %.3f\n", ci, sizes[ci], t[0], t[1], t[2]);
    }
    for (auto &e : tev) cudaEventDestroy(e);
  }
  CK(cudaEventRecord(c->half_done[half], c->d2h_stream));  // every kernel of the call precedes its last download
  c->half_used[half] = true;
  return 0;
}
int fhesi_mult_relin_host(fhesi_ctx *c, const fhesi_ksw *ksw, const uint32_t *h_a, const uint32_t *h_b,
                          uint32_t *h_out, size_t count) {
  int rc = fhesi_mult_relin_host_async(c, ksw, h_a, h_b, h_out, count);
  if (rc || !count) return rc;
  return fhesi_sync_all(c);
}

// ---------------------------------------------------------------------------------------
// Encrypt / Decrypt
// ---------------------------------------------------------------------------------------
int fhesi_encrypt_dev(fhesi_ctx *c, const fhesi_key *pk, const uint32_t *msg, const uint8_t *r,
                      const int32_t *e, uint32_t *out, size_t count) {
  if (!c || !pk || !msg || !r || !e || !out) return fail(FHESI_ERR_INVALID, "null argument");
  if (pk->ctx != c || pk->parts != 2) return fail(FHESI_ERR_INVALID, "bad public key");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const u32 Le = I.Le;
  const size_t per = (size_t)Le * I.N, CH = c->chunk;
  size_t n1 = al(CH * per), n2 = al(CH * 2 * per), n3 = al(CH * 2 * Le * I.n);
  u32 *s = nullptr;
  int rc = scratch(c, (n1 + n2 + n3) * 4, &s);
  if (rc) return rc;
  u32 *sR = s, *sO = s + n1, *sC = sO + n2;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    if ((rc = launch_fwd(c, r + off * I.n, SRC_U8, 0, SC_NONE, Le, sR, cnt))) return rc;
    // out[b][j] = NTT(r_b) . pk_j : K = 1 input, J = 2 outputs; key layout [Le][1][2][N]
    DotArgs d{sR, pk->d_key, 1, 2, Le, sO, (u32)cnt};
    KL(c, k_dot, nblk(per * cnt), 256, 0, c->dc, d);
    CKL();
    if ((rc = launch_inv(c, sO, Le, sC, cnt * 2, e + off * 2 * I.n, msg + off * I.n))) return rc;
    if ((rc = launch_crt(c, sC, Le, CRT_REDUCE_Q, out + off * 2 * I.n * I.W, I.W, cnt * 2))) return rc;
  }
  return 0;
}
int fhesi_decrypt_dev(fhesi_ctx *c, const fhesi_key *sk, const uint32_t *in, uint32_t parts,
                      uint32_t *msg, size_t count) {
  if (!c || !sk || !in || !msg) return fail(FHESI_ERR_INVALID, "null argument");
  if (sk->ctx != c || parts < sk->parts) return fail(FHESI_ERR_INVALID, "bad secret key / part count");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  const u32 Le = I.Le, K = sk->parts;
  const size_t per = (size_t)Le * I.N, CH = c->chunk;
  size_t n1 = al(CH * K * per), n2 = al(CH * per), n3 = al(CH * Le * I.n), n4 = al(CH * K * I.n * I.W);
  u32 *s = nullptr;
  int rc = scratch(c, (n1 + n2 + n3 + n4) * 4, &s);
  if (rc) return rc;
  u32 *sF = s, *sO = s + n1, *sC = sO + n2, *sIn = sC + n3;
  const size_t ctw = (size_t)I.n * I.W;
  for (size_t off = 0; off < count; off += CH) {
    size_t cnt = count - off < CH ? count - off : CH;
    const u32 *src = in + off * parts * ctw;
    if (parts != K) {  // compact the first K parts of each ciphertext
      CK(cudaMemcpy2DAsync(sIn, K * ctw * 4, src, parts * ctw * 4, K * ctw * 4, cnt,
                           cudaMemcpyDeviceToDevice, c->stream));
      src = sIn;
    }
    if ((rc = launch_fwd(c, src, SRC_POLY, I.W, SC_NONE, Le, sF, cnt * K))) return rc;
    DotArgs d{sF, sk->d_key, K, 1, Le, sO, (u32)cnt};  // key layout [Le][K][1][N]
    KL(c, k_dot, nblk(per * cnt), 256, 0, c->dc, d);
    CKL();
    if ((rc = launch_inv(c, sO, Le, sC, cnt))) return rc;
    if ((rc = launch_crt(c, sC, Le, CRT_DECRYPT, msg + off * I.n, 1, cnt))) return rc;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------
// reference-chain rows (key export parity)
// ---------------------------------------------------------------------------------------
int fhesi_ref_rows_host(fhesi_ctx *c, const uint32_t *h_poly, uint32_t Win, const uint64_t *primes,
                        const uint64_t *roots, uint32_t L, int64_t *h_rows) {
  if (!c || !h_poly || !primes || !roots || !h_rows || !L || !Win)
    return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  const fhesi_info &I = c->info;
  std::vector<u64> pinv(L), r2(L), zeta(L);
  for (u32 l = 0; l < L; ++l) {
    u64 q = primes[l];
    if (!(q & 1) || q >> 62) return fail(FHESI_ERR_INVALID, "reference primes must be odd and < 2^62");
    u64 inv = 1;
    for (int it = 0; it < 6; ++it) inv *= 2 - q * inv;
    pinv[l] = 0 - inv;
    u64 R = (u64)(((unsigned __int128)1 << 64) % q);
    r2[l] = h_mulmod(R, R, q);
    zeta[l] = h_mulmod(roots[l] % q, roots[l] % q, q);
  }
  std::vector<u32> units;
  for (u32 i = 0; i < I.m; ++i) {
    u32 a = i, b = I.m;
    while (b) { u32 t = a % b; a = b; b = t; }
    if (a == 1) units.push_back(i);
  }
  char *d = nullptr;
  size_t o_poly = 0, o_p = o_poly + al((size_t)I.n * Win) * 4, o_pi = o_p + L * 8, o_r2 = o_pi + L * 8,
         o_z = o_r2 + L * 8, o_u = o_z + L * 8, o_rows = o_u + al(I.n) * 4, tot = o_rows + (size_t)L * I.n * 8;
  CK(cudaMalloc(&d, tot));
  CK(cudaMemcpyAsync(d + o_poly, h_poly, (size_t)I.n * Win * 4, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d + o_p, primes, L * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d + o_pi, pinv.data(), L * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d + o_r2, r2.data(), L * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d + o_z, zeta.data(), L * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d + o_u, units.data(), I.n * 4, cudaMemcpyHostToDevice, c->stream));
  RefRowArgs a{(const u32 *)(d + o_poly), Win, L, (const u64 *)(d + o_p), (const u64 *)(d + o_pi),
               (const u64 *)(d + o_r2), (const u64 *)(d + o_z), (const u32 *)(d + o_u), (i64 *)(d + o_rows)};
  dim3 grid((I.n + 127) / 128, L);
  KL(c, k_ref_rows, grid, 128, 0, c->dc, a);
  CKL();
  CK(cudaMemcpyAsync(h_rows, d + o_rows, (size_t)L * I.n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaFree(d));
  return 0;
}

// ---------------------------------------------------------------------------------------
// launch accounting and per-kernel event profiler
// ---------------------------------------------------------------------------------------
static void prof_clear(fhesi_ctx *c) {
  for (auto &r : c->prof_recs) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  c->prof_recs.clear();
  c->prof_names.clear();
}
int fhesi_profile_enable(fhesi_ctx *c, int on) {
  if (!c) return fail(FHESI_ERR_INVALID, "null ctx");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  prof_clear(c);
  c->prof_on = on != 0;
  c->launches = 0;
  return 0;
}
int fhesi_profile_launches(fhesi_ctx *c, uint64_t *launches) {
  if (!c || !launches) return fail(FHESI_ERR_INVALID, "null argument");
  *launches = c->launches;
  return 0;
}
int fhesi_profile_report(fhesi_ctx *c, char *buf, size_t cap) {
  if (!c || !buf || !cap) return fail(FHESI_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  std::vector<double> ms(c->prof_names.size(), 0.0);
  std::vector<uint64_t> cnt(c->prof_names.size(), 0);
  for (auto &r : c->prof_recs) {
    float t = 0;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms[r.id] += t;
      cnt[r.id]++;
    }
  }
  std::string out;
  for (size_t i = 0; i < ms.size(); ++i) {
    char line[256];
    snprintf(line, sizeof line, "%s %llu %.6f\n", c->prof_names[i].c_str(), (unsigned long long)cnt[i], ms[i]);
    out += line;
  }
  if (out.size() + 1 > cap) return fail(FHESI_ERR_INVALID, "report buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return 0;
}

// ---------------------------------------------------------------------------------------
// modmul peak
// ---------------------------------------------------------------------------------------
template <int KIND>
static void launch_pipe(fhesi_ctx *c, int blocks, int threads, void *d, int iters) {
  FHESI_LAUNCH(k_pipe<KIND>, blocks, threads, 0, c->stream, (u32 *)d, c->h_pc[0].p, iters);
}
int fhesi_pipe_peak(fhesi_ctx *c, int kind, double *out) {
  if (!c || !out || kind < 0 || kind > 5) return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  void *d = nullptr;
  CK(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, c->stream));
    switch (kind) {
      case 0: launch_pipe<0>(c, blocks, threads, d, iters); break;
      case 1: launch_pipe<1>(c, blocks, threads, d, iters); break;
      case 2: launch_pipe<2>(c, blocks, threads, d, iters); break;
      case 3: launch_pipe<3>(c, blocks, threads, d, iters); break;
      case 4: launch_pipe<4>(c, blocks, threads, d, iters); break;
      default: launch_pipe<5>(c, blocks, threads, d, iters); break;
    }
    CKL();
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *out = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
  return 0;
}
int fhesi_modmul_peak(fhesi_ctx *c, int word_bits, double *out) {
  if (!c || !out || (word_bits != 32 && word_bits != 64)) return fail(FHESI_ERR_INVALID, "bad argument");
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  void *d = nullptr;
  CK(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0, c->stream));
    if (word_bits == 32)
      KL(c, k_peak32, blocks, threads, 0, (u32 *)d, c->h_pc[0].p, c->h_pc[0].pinv, iters);
    else
      KL(c, k_peak64, blocks, threads, 0, (u64 *)d, 1152921504606820681ull, 0x9d1c9a1c8c4a3b47ull | 1, iters);
    CKL();
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *out = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
  return 0;
}
