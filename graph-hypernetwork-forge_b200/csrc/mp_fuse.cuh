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

// ---- dropout on torch's Philox stream (DropoutArgs, mp.cuh) ----------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// multipliers (0 or scale) of elements 4v .. 4v + 3
template <class DA>
__device__ __forceinline__ void dropout_mult4(const DA& da, uint64_t v, float (&m)[4]) {
  uint64_t t, s;
  if ((v >> 32) == 0) {   // 32-bit division when it fits (any tensor below 2^34 elements)
    t = (uint32_t)v % da.threads;
    s = (uint32_t)v / da.threads;
  } else {
    t = v % da.threads;
    s = v / da.threads;
  }
  const uint64_t ctr = da.ctr0 + s;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)t, (uint32_t)(t >> 32)),
                                make_uint2(da.key0, da.key1));
  // curand_uniform: x * 2^-32 + 2^-33 in one fused multiply-add, in (0, 1]
  m[0] = __fmaf_rn(__uint2float_rn(r.x), 2.3283064e-10f, 1.1641532e-10f) < da.keep ? da.scale : 0.f;
  m[1] = __fmaf_rn(__uint2float_rn(r.y), 2.3283064e-10f, 1.1641532e-10f) < da.keep ? da.scale : 0.f;
  m[2] = __fmaf_rn(__uint2float_rn(r.z), 2.3283064e-10f, 1.1641532e-10f) < da.keep ? da.scale : 0.f;
  m[3] = __fmaf_rn(__uint2float_rn(r.w), 2.3283064e-10f, 1.1641532e-10f) < da.keep ? da.scale : 0.f;
}
// multiplier of the single element e
template <class DA>
__device__ __forceinline__ float dropout_mult1(const DA& da, uint64_t e) {
  float m[4];
  dropout_mult4(da, e >> 2, m);
  const int c = (int)(e & 3);
  return c == 0 ? m[0] : (c == 1 ? m[1] : (c == 2 ? m[2] : m[3]));
}
// multipliers of the V consecutive elements starting at e0 (V = 1, 2, 4; e0 % V == 0): one Philox call
template <int V, class DA>
__device__ __forceinline__ void dropout_multv(const DA& da, uint64_t e0, float (&mv)[V]) {
  float m[4];
  dropout_mult4(da, e0 >> 2, m);
  if constexpr (V == 4) {
    mv[0] = m[0]; mv[1] = m[1]; mv[2] = m[2]; mv[3] = m[3];
  } else if constexpr (V == 2) {
    const bool hi = (e0 & 2) != 0;
    mv[0] = hi ? m[2] : m[0];
    mv[1] = hi ? m[3] : m[1];
  } else {
    const int c = (int)(e0 & 3);
    mv[0] = c == 0 ? m[0] : (c == 1 ? m[1] : (c == 2 ? m[2] : m[3]));
  }
}

}  // namespace fuse
}  // namespace ghf
