// mp_fuse.cuh — small device helpers shared by the message-passing kernels: acquire loads / bounded spinning
// on grid-wide progress words (mp_f16.cu) and vector row accesses (the layer epilogue in mp.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>

namespace ghf {
namespace fuse {

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void spin_until_at_least(const int32_t* p, int target) {
  for (uint32_t spins = 0; ld_acquire(p) < target; ++spins) {
    __nanosleep(64);
    if (spins > (1u << 24)) __trap();
  }
}

// V consecutive floats as one vector access (V = 1, 2, 4)
template <int V> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };
template <int V>
__device__ __forceinline__ void vload_cg(float (&dst)[V], const float* p) {   // L2 (skips L1)
  using T = typename VecT<V>::type;
  const T v = __ldcg(reinterpret_cast<const T*>(p));
  memcpy(dst, &v, sizeof(T));
}
template <int V>
__device__ __forceinline__ void vload_nc(float (&dst)[V], const float* p) {   // read-only path
  using T = typename VecT<V>::type;
  const T v = __ldg(reinterpret_cast<const T*>(p));
  memcpy(dst, &v, sizeof(T));
}
template <int V>
__device__ __forceinline__ void vstore(float* p, const float (&src)[V]) {
  using T = typename VecT<V>::type;
  T v;
  memcpy(&v, src, sizeof(T));
  *reinterpret_cast<T*>(p) = v;
}

}  // namespace fuse
}  // namespace ghf
