// common.cuh — error plumbing, launch accounting and small device helpers shared by all
// translation units of libghf_b200.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace ghf {

constexpr int kErrLen = 512;
char* err_buf();                       // thread-local message buffer (api.cu)
extern std::atomic<int64_t> g_launches;  // kernels launched by this library (api.cu)

inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), kErrLen, fmt, ap);
  va_end(ap);
  return 1;
}

#define GHF_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t e__ = (expr);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return ::ghf::fail("%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                         cudaGetErrorString(e__));                                  \
  } while (0)

// after a <<<>>> launch: count it and surface launch-configuration errors immediately
#define GHF_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ::ghf::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    GHF_CUDA(cudaGetLastError());                                                   \
  } while (0)

#define GHF_REQUIRE(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) return ::ghf::fail(__VA_ARGS__);                                   \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t align_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

int sm_count();  // SMs of the current device (cached; api.cu)

// GHF_PREC_F16 range guard: one device word per device, set by every kernel that writes an fp16 shadow of h when a
// magnitude exceeds the fp16 range (api.cu).  nullptr when it cannot be allocated.
int* f16_overflow_flag();
#ifdef __CUDACC__
__device__ __forceinline__ void flag_f16_overflow(float max_abs, int* flag) {
  if (max_abs > 65504.f) atomicOr(flag, 1);
}
#endif

// stream-ordered scratch that is returned to the pool at scope exit
struct TempBuf {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    return cudaMallocAsync(&p, bytes ? bytes : 1, stream);
  }
  ~TempBuf() {
    if (p) cudaFreeAsync(p, s);
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace ghf
