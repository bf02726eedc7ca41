// Shared helpers for the sm_100a kernels of libsfm_b200.  Internal header (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>

#define SFM_OK 0
#define SFM_ERR_INVALID (-1)
#define SFM_ERR_CUDA (-2)
#define SFM_ERR_WORKSPACE (-3)
#define SFM_ERR_STATE (-4)
#define SFM_ERR_UNSUPPORTED (-5)

namespace sfm {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launches(long long n);

#define SFM_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) return ::sfm::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

// every kernel launch site is followed by SFM_LAUNCH_CHECK(); `n` = launches covered by the check
#define SFM_LAUNCH_CHECK_N(n)            \
  do {                                   \
    ::sfm::count_launches(n);            \
    SFM_CUDA(cudaGetLastError());        \
  } while (0)
#define SFM_LAUNCH_CHECK() SFM_LAUNCH_CHECK_N(1)

#define SFM_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) {                                   \
      ::sfm::set_error(__VA_ARGS__);                 \
      return SFM_ERR_INVALID;                        \
    }                                                \
  } while (0)

#define SFM_TRY(expr)                   \
  do {                                  \
    int rc__ = (expr);                  \
    if (rc__ != SFM_OK) return rc__;    \
  } while (0)

// One-time per-DEVICE setup of a kernel (opt-in to > 48 KB of dynamic shared memory is a per-device function
// attribute).  One instance per kernel instantiation; a bit per device ordinal, set after the setup succeeded,
// so a second GPU in the same process -- or a second thread racing on the first call -- still gets configured
// (the setup is idempotent).
struct PerDeviceOnce {
  std::atomic<unsigned long long> done[2];
  PerDeviceOnce() { done[0] = 0; done[1] = 0; }
  template <typename F>
  int run(F&& setup) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__);
    std::atomic<unsigned long long>& word = done[(dev >> 6) & 1];
    const unsigned long long bit = 1ull << (dev & 63);
    if (word.load(std::memory_order_acquire) & bit) return SFM_OK;
    const int rc = setup();
    if (rc == SFM_OK) word.fetch_or(bit, std::memory_order_release);
    return rc;
  }
};
// opt a kernel in to `smem` bytes of dynamic shared memory on the current device (once per device)
#define SFM_SMEM_OPT_IN(kernel, smem)                                                                         \
  do {                                                                                                        \
    static ::sfm::PerDeviceOnce once__;                                                                       \
    SFM_TRY(once__.run([&]() -> int {                                                                         \
      SFM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));       \
      return SFM_OK;                                                                                          \
    }));                                                                                                      \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (torch owns the memory).
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool overflow;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0), overflow(false) {}
  template <typename T>
  T* get(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (base == nullptr || off + bytes > cap) {
      overflow = true;
      off += bytes;
      return nullptr;
    }
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (max, sum) pair of a running log-sum-exp; merge is associative.
struct Lse {
  float m, s;
};
__device__ __forceinline__ Lse lse_merge(Lse a, Lse b) {
  float m = fmaxf(a.m, b.m);
  if (m == -INFINITY) return Lse{m, 0.f};
  return Lse{m, a.s * __expf(a.m - m) + b.s * __expf(b.m - m)};
}

}  // namespace sfm
