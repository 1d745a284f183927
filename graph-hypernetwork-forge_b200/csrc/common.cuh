// common.cuh — error plumbing, launch accounting and small device helpers shared by all
// translation units of libghf_b200.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <functional>
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

int sm_count();  // SMs of the current device (cached per device; api.cu)

// ghf_set_presync_hook (api.cu): a host callback the NEXT blocking entry point of this thread (ghf_select_edges,
// ghf_dedup_texts, ghf_graph_build) calls once, after its kernels are enqueued and right before it waits for a size
// to come back from the device.  Work the callback enqueues (on any stream) fills the GPU during the round trip.
// The wait itself is for an EVENT recorded right after the read-back copy, not for the stream: work the hook puts on
// the same stream does not delay the host.
// readback_wait: record the event on `stream`, run `first` (an internal hook, may be NULL) and the caller's hook, wait.
int readback_wait(cudaStream_t stream, const std::function<int()>* first = nullptr);

// cudaFuncSetAttribute is per DEVICE: kernels that opt in to large shared memory configure themselves on first use
// on each device of the process.  -> true the first time it is called with these flags on the current device.
inline bool first_use_on_device(bool (&flags)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;   // unknown device: configure again
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// ---- fp16 shadow of h (GHF_PREC_F16) ------------------------------------------------------------------------
// A shadow is (h16, scale) with scale = float[2] in device memory: h = h16 * scale[0] (scale[0] is an exact power of
// two), scale[1] = max |h| (written by the producer of h, read by the kernel that picks the scale).  Everything is
// decided on the device: no host round trip, no dependence on the data range.
#ifdef __CUDACC__
// power of two s with amax * s in [2^13, 2^14): inside the fp16 range with headroom; 1 for amax = 0 / inf / nan
__device__ __forceinline__ float f16_scale_for(float amax) {
  if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
  int e;
  frexpf(amax, &e);  // amax = f * 2^e, f in [0.5, 1)
  e = 14 - e;
  return ldexpf(1.f, e > 120 ? 120 : (e < -120 ? -120 : e));
}
// max over non-negative floats (their IEEE bit patterns order like integers)
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
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
