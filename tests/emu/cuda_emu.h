// cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal CUDA execution-model emulator so that the *same* kernel sources under
// fhe-si_b200/csrc can be compiled with g++ and their logic checked against the oracle in
// this GPU-less build container (`pytest -m "not gpu"`).  Every CUDA thread of a block is a
// real OS thread; __syncthreads() and warp shuffles are barriers.  It is slow and is never
// loaded by the product: the Python/C++ host layers only ever open libfhesi_b200.so, which
// is compiled by nvcc for sm_100a and has no CPU path.  The emulator library is built by
// tests/emu/build_emu.py into tests/emu/_build/ and opened explicitly by tests/conftest.py.
#pragma once
#ifndef FHESI_EMU
#define FHESI_EMU 1
#endif
#include <algorithm>
#include <barrier>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu {
  unsigned x, y, z;
};
namespace emu {
struct BlockState {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<std::unique_ptr<std::barrier<>>> group_bar;  // 128-thread named barriers
  std::vector<std::unique_ptr<std::barrier<>>> group_bar256;  // 256-thread named barriers
  std::vector<uint64_t> shfl;  // one slot per thread
  std::vector<uint32_t> smem;
  unsigned nthreads = 0;
};
inline thread_local uint3_emu t_threadIdx, t_blockIdx;
inline thread_local dim3 t_blockDim, t_gridDim;
inline thread_local BlockState *t_block = nullptr;
inline uint32_t *smem() { return t_block->smem.data(); }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const unsigned T = block.x * block.y * block.z;
  BlockState st;
  st.nthreads = T;
  st.bar.reset(new std::barrier<>(T));
  for (unsigned w = 0; w < (T + 31) / 32; ++w) {
    unsigned cnt = std::min(32u, T - w * 32);
    st.warp_bar.emplace_back(new std::barrier<>(cnt));
  }
  for (unsigned g = 0; g < (T + 127) / 128; ++g) {
    unsigned cnt = std::min(128u, T - g * 128);
    st.group_bar.emplace_back(new std::barrier<>(cnt));
  }
  for (unsigned g = 0; g < (T + 255) / 256; ++g) {
    unsigned cnt = std::min(256u, T - g * 256);
    st.group_bar256.emplace_back(new std::barrier<>(cnt));
  }
  st.shfl.assign(T, 0);
  st.smem.assign(smem_bytes / 4 + 64, 0);
  auto worker = [&](unsigned tid) {
    t_block = &st;
    t_blockDim = block;
    t_gridDim = grid;
    t_threadIdx = {tid % block.x, (tid / block.x) % block.y, tid / (block.x * block.y)};
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = {bx, by, bz};
          body();
          st.bar->arrive_and_wait();  // block finished: shared memory may be reused
        }
  };
  if (T == 1) {
    worker(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(T);
  for (unsigned t = 0; t < T; ++t) th.emplace_back(worker, t);
  for (auto &x : th) x.join();
}
inline unsigned linear_tid() {
  return t_threadIdx.x + t_blockDim.x * (t_threadIdx.y + t_blockDim.y * t_threadIdx.z);
}
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

inline void __syncthreads() { emu::t_block->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::t_block->warp_bar[emu::linear_tid() / 32]->arrive_and_wait();
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  unsigned tid = emu::linear_tid(), w = tid / 32;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  emu::t_block->shfl[tid] = raw;
  emu::t_block->warp_bar[w]->arrive_and_wait();
  uint64_t got = emu::t_block->shfl[(tid & ~31u) | ((tid ^ (unsigned)lane_mask) & 31u)];
  emu::t_block->warp_bar[w]->arrive_and_wait();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  unsigned tid = emu::linear_tid(), w = tid / 32;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  emu::t_block->shfl[tid] = raw;
  emu::t_block->warp_bar[w]->arrive_and_wait();
  uint64_t got = emu::t_block->shfl[(tid & ~31u) | ((unsigned)src_lane & 31u)];
  emu::t_block->warp_bar[w]->arrive_and_wait();
  T out;
  memcpy(&out, &got, sizeof(T));
  return out;
}
struct uint4 {
  unsigned x, y, z, w;
};
struct uint2 {
  unsigned x, y;
};
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct double2 {
  double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
template <class T>
inline T __ldg(const T *p) { return *p; }
// named barrier over the 128-thread group `g` of the block (bar.sync g+1, 128 on the GPU)
inline void fhesi_group_sync(unsigned g) { emu::t_block->group_bar[g]->arrive_and_wait(); }
template <int T>
inline void fhesi_group_sync_t(unsigned g) {
  static_assert(T == 128 || T == 256, "named barrier sizes the emulator knows");
  (T == 128 ? emu::t_block->group_bar[g] : emu::t_block->group_bar256[g])->arrive_and_wait();
}
inline uint64_t __umul64hi(uint64_t a, uint64_t b) {
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
}
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
using std::max;
using std::min;

#define FHESI_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (smem), [&] { kern(__VA_ARGS__); })
#define FHESI_SMEM(name) uint32_t *name = emu::smem()

// ---- the slice of the CUDA runtime API the library uses, on host memory ----------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  int multiProcessorCount = 1;
};
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
enum { cudaHostAllocDefault = 0, cudaHostAllocWriteCombined = 4 };
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
  return 0;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = nullptr; return 0; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return 0; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
