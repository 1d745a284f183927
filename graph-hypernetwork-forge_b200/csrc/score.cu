// score.cu — batched link-prediction scores fused with the gather of head/tail rows.
// The reference scores pairs with `score_triple(embs[src], embs[dst])` (HG:304-318, demo.py:90-94): two [B, d] gathers
// materialised, then a product and a row sum.  Here one warp reads the two rows in place and reduces them:
//   out[b] = <emb[heads[b]], emb[tails[b]]>            (2 * d * 4 bytes read per pair, 4 written)
// and the backward scatters  g_emb[heads[b]] += g[b] emb[tails[b]],  g_emb[tails[b]] += g[b] emb[heads[b]].
#include "common.cuh"
#include "ghf_b200.h"

namespace ghf {
namespace {

template <bool VEC>
__global__ void __launch_bounds__(256)
score_pairs_kernel(const float* __restrict__ emb, int64_t N, int d, const int64_t* __restrict__ heads,
                   const int64_t* __restrict__ tails, int64_t B, float* __restrict__ out, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    const int64_t hd = heads[b], tl = tails[b];
    if (hd < 0 || hd >= N || tl < 0 || tl >= N) {
      if (lane == 0) { *bad = 1; out[b] = 0.f; }
      continue;
    }
    const float* x = emb + hd * d;
    const float* y = emb + tl * d;
    float s = 0.f;
    if (VEC) {
      for (int c = lane * 4; c < d; c += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x + c)), v = __ldg(reinterpret_cast<const float4*>(y + c));
        s += a.x * v.x + a.y * v.y + a.z * v.z + a.w * v.w;
      }
    } else {
      for (int c = lane; c < d; c += 32) s = fmaf(x[c], y[c], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[b] = s;
  }
}

__global__ void __launch_bounds__(256)
score_pairs_bwd_kernel(const float* __restrict__ emb, int64_t N, int d, const int64_t* __restrict__ heads,
                       const int64_t* __restrict__ tails, int64_t B, const float* __restrict__ g_out,
                       float* __restrict__ g_emb) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < B; b += warps) {
    const int64_t hd = heads[b], tl = tails[b];
    if (hd < 0 || hd >= N || tl < 0 || tl >= N) continue;
    const float g = g_out[b];
    for (int c = lane; c < d; c += 32) {
      const float xh = emb[hd * d + c], xt = emb[tl * d + c];
      atomicAdd(g_emb + hd * d + c, g * xt);
      atomicAdd(g_emb + tl * d + c, g * xh);
    }
  }
}

unsigned pair_grid(int64_t B) {
  const int64_t want = cdiv(B, 8), cap = (int64_t)sm_count() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" int ghf_score_pairs(const float* d_emb, int64_t N, int d, const int64_t* d_heads, const int64_t* d_tails,
                               int64_t B, float* d_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(N >= 0 && d > 0 && B >= 0, "ghf_score_pairs: bad dims N=%lld d=%d B=%lld", (long long)N, d, (long long)B);
  if (B == 0) return 0;
  GHF_REQUIRE(d_emb && d_heads && d_tails && d_out, "ghf_score_pairs: NULL argument");
  TempBuf bad;
  GHF_CUDA(bad.alloc(sizeof(int), stream));
  GHF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream));
  const bool vec = d % 4 == 0 && reinterpret_cast<uintptr_t>(d_emb) % 16 == 0;
  if (vec)
    score_pairs_kernel<true><<<pair_grid(B), 256, 0, stream>>>(d_emb, N, d, d_heads, d_tails, B, d_out, bad.as<int>());
  else
    score_pairs_kernel<false><<<pair_grid(B), 256, 0, stream>>>(d_emb, N, d, d_heads, d_tails, B, d_out, bad.as<int>());
  GHF_LAUNCH_CHECK();
  int h_bad = 0;
  GHF_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
  GHF_CUDA(cudaStreamSynchronize(stream));
  GHF_REQUIRE(h_bad == 0, "ghf_score_pairs: a head or tail id lies outside [0, %lld)", (long long)N);
  return 0;
}

extern "C" int ghf_score_pairs_backward(const float* d_emb, int64_t N, int d, const int64_t* d_heads,
                                        const int64_t* d_tails, int64_t B, const float* d_g_out, float* d_g_emb,
                                        void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(N >= 0 && d > 0 && B >= 0, "ghf_score_pairs_backward: bad dims");
  GHF_REQUIRE(d_g_emb != nullptr || N == 0, "ghf_score_pairs_backward: NULL gradient buffer");
  GHF_CUDA(cudaMemsetAsync(d_g_emb, 0, (size_t)N * d * sizeof(float), stream));
  if (B == 0) return 0;
  score_pairs_bwd_kernel<<<pair_grid(B), 256, 0, stream>>>(d_emb, N, d, d_heads, d_tails, B, d_g_out, d_g_emb);
  GHF_LAUNCH_CHECK();
  return 0;
}
