// Error plumbing, workspace carving and small device helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200boot.h"

namespace b200boot {

// ---- thread-local last error -------------------------------------------------
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define B2_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return fail(B200BOOT_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                      \
  } while (0)

// every kernel launch of the library goes through this macro, which also counts it
inline std::atomic<long long>& launch_counter() {
  static std::atomic<long long> n{0};
  return n;
}
#define B2_LAUNCH_CHECK()                                                  \
  do {                                                                     \
    b200boot::launch_counter().fetch_add(1, std::memory_order_relaxed);   \
    B2_CUDA(cudaGetLastError());                                           \
  } while (0)

// One-time work per device (cudaFuncSetAttribute and the like) under std::call_once: exports
// are called from several host threads (ctypes drops the GIL).  run() returns f's status on the
// call that executes it and cudaSuccess afterwards.
constexpr int kMaxDevices = 64;
struct PerDeviceOnce {
  std::once_flag flags[kMaxDevices];
  template <typename F>
  cudaError_t run(F&& f) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kMaxDevices) return f();
    std::call_once(flags[dev], [&] { e = f(); });
    return e;
  }
};

// ---- workspace carving ---------------------------------------------------------
struct Carver {
  unsigned char* base;
  size_t size;
  size_t used;
  bool ok;
  Carver(void* p, size_t n) : base((unsigned char*)p), size(n), used(0), ok(true) {}
  template <typename T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* out = (T*)(base + used);
    used += bytes;
    if (base == nullptr || used > size) ok = false;
    return out;
  }
};

inline size_t pad256(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

inline int sm_count() {
  static std::atomic<int> cached[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  int n = cached[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

#if defined(__CUDACC__)
// ---- warp / block reductions (fixed order => deterministic) -------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// ---- mbarrier + 1-D bulk (TMA) copy, sm_90+/sm_100a ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

}  // namespace b200boot
