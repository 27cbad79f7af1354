// kernels_fused.cuh -- fused N=1024 hot-path kernels (placeholder while the generic path is
// brought up; replaced below).
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

static bool fused_supported(const DevCtx &) { return false; }
static int fused_configure() { return 0; }
struct fhesi_ctx;
struct fhesi_ksw;
static int fused_keyswitch(fhesi_ctx *, const fhesi_ksw *, const u32 *, u32 *, size_t) { return -2; }
static int fused_mult_relin(fhesi_ctx *, const fhesi_ksw *, const u32 *, const u32 *, u32 *, size_t) { return -2; }
