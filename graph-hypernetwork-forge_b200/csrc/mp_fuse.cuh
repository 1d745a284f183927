// mp_fuse.cuh — the in-place fused layer epilogue shared by the tcgen05 contraction kernels
// (mp_umma.cu: weights in shared memory; mp_umma_ts.cu: weights in tensor memory).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstring>

namespace ghf {
namespace fuse {

// ---- in-place fusion of the layer epilogue -----------------------------------------------------------------
// The output rows double as the accumulator.  `epi_ctas` CTAs (lowest block ids) do no contraction: they zero
// the rows of super-block ("phase") p just before its units start reducing into them, and as soon as the last
// unit of phase p has signalled they turn those rows into LayerNorm(relu(acc / max(indeg,1) + h)) in place -
// while rows, h[dst] rows and in-degrees are still in L2.  No separate clear pass, no separate epilogue pass,
// no accumulator round trip through HBM.
//   zero_done[p]  : epilogue CTAs that have zeroed their share of phase p      (== epi_ctas -> units may start)
//   units_done[p] : units of phase p whose reductions are complete            (== phase_units[p] -> epilogue)
struct FuseArgs {
  const int32_t* unit_phase;
  const int32_t* phase_units;
  int32_t* zero_done;
  int32_t* units_done;
  const int32_t* indeg;
  const float* ln_w;
  const float* ln_b;
  float* upd;  // optional pre-residual update (parity taps)
  float eps;
  int64_t num_local;
  int32_t sb_nodes, num_phases, epi_ctas;
};

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

template <int D>
__device__ void epilogue_cta(const FuseArgs& fa, float* __restrict__ out, const float* __restrict__ h,
                             int64_t dst_lo) {
  constexpr int V = D / 32;  // consecutive floats per lane: a warp covers one row with one vector access
  constexpr int kRows = 8;   // rows in flight per warp (latency hiding: ~8 KB of loads per warp)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = fa.epi_ctas * (blockDim.x >> 5);
  const int slot = blockIdx.x * (blockDim.x >> 5) + warp;
  float lw[V], lb[V];
  vload_nc<V>(lw, fa.ln_w + lane * V);
  vload_nc<V>(lb, fa.ln_b + lane * V);
  const float zeros[V] = {};

  auto zero_phase = [&](int p) {
    const int64_t lo = (int64_t)p * fa.sb_nodes, hi = min(lo + fa.sb_nodes, fa.num_local);
    for (int64_t r = lo + slot; r < hi; r += nslots) vstore<V>(out + r * D + lane * V, zeros);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&fa.zero_done[p], 1);
    }
  };

  zero_phase(0);
  if (fa.num_phases > 1) zero_phase(1);
  for (int p = 0; p < fa.num_phases; ++p) {
    if (threadIdx.x == 0) spin_until_at_least(&fa.units_done[p], fa.phase_units[p]);
    __syncthreads();
    __threadfence();
    const int64_t lo = (int64_t)p * fa.sb_nodes, hi = min(lo + fa.sb_nodes, fa.num_local);
    for (int64_t r0 = lo + slot; r0 < hi; r0 += (int64_t)kRows * nslots) {
      float a[kRows][V], hv[kRows][V];
      int deg[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int64_t r = r0 + (int64_t)k * nslots;
        if (r < hi) {
          vload_cg<V>(a[k], out + r * D + lane * V);  // reductions live in L2, never in L1
          vload_nc<V>(hv[k], h + (dst_lo + r) * D + lane * V);
          deg[k] = __ldg(fa.indeg + r);
        }
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        const int64_t r = r0 + (int64_t)k * nslots;
        if (r >= hi) break;
        const float inv = 1.f / (float)max(deg[k], 1);
        float x[V], u[V], sum = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          u[j] = a[k][j] * inv;
          x[j] = fmaxf(u[j] + hv[k][j], 0.f);
          sum += x[j];
        }
        if (fa.upd) vstore<V>(fa.upd + r * D + lane * V, u);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
        const float mean = sum / (float)D;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j) var += (x[j] - mean) * (x[j] - mean);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
        const float rstd = rsqrtf(var / (float)D + fa.eps);
        float y[V];
#pragma unroll
        for (int j = 0; j < V; ++j) y[j] = (x[j] - mean) * rstd * lw[j] + lb[j];
        vstore<V>(out + r * D + lane * V, y);
      }
    }
    if (p + 2 < fa.num_phases) zero_phase(p + 2);
  }
}

}  // namespace fuse
}  // namespace ghf
