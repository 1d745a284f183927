// mp_backward.cu — gradient kernels of one message-passing layer (the reference is trainable,
// tests/test_hypergnn.py:183-226 and demo.py:79-101 call loss.backward() through HyperGNN.forward).
//
// Forward (mp.cu):   acc_v = sum_{e:(u->v), r} ( h_u W_msg[r] + h_v W_self[r] + bias[r] )
//                    upd_v = acc_v / max(indeg_v, 1);  x_v = relu(upd_v + h_v);  out_v = LayerNorm(x_v)
// Backward, given g_out:
//   (1) ghf_mp_epilogue_backward:  g_pre_v = dL/d(upd_v + h_v)  (LayerNorm and ReLU undone), g_acc_v = g_pre_v / cnt_v,
//                                  g_ln_w, g_ln_b (column sums over the rows)
//   (2) dL/dh = g_pre + sum_{e:(u->v)} g_acc_v W_msg[r]^T (at u) + sum_{e:(.->v)} g_acc_v W_self[r]^T (at v):
//       the SAME contraction as the forward, run by the host side through ghf_mp_contract on the reversed graph
//       (transposed W_msg) and on the graph itself (transposed W_self) - no new kernel;
//   (3) ghf_mp_weight_grad:  g_W_msg[r] = sum_{e in r} h_u^T g_acc_v,  g_W_self[r] = sum_{e in r} h_v^T g_acc_v,
//                            g_bias[r] = sum_{e in r} g_acc_v      - a contraction over the EDGES of each unit.
#include "common.cuh"
#include "ghf_b200.h"
#include "graph.cuh"
#include "mp.cuh"
#include "mp_fuse.cuh"

namespace ghf {
namespace {

constexpr int kBwdMaxV = 8;  // hidden_dim <= 256 keeps a row in registers (one warp per row)

__global__ void __launch_bounds__(256)
mp_epilogue_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ upd,
                       const int32_t* __restrict__ indeg, const float* __restrict__ h, int64_t dst_lo,
                       int64_t num_local, int d, const float* __restrict__ ln_w, float eps,
                       float* __restrict__ g_pre, float* __restrict__ g_acc, float* __restrict__ g_lnw,
                       float* __restrict__ g_lnb, const DropoutArgs da) {
  extern __shared__ float red[];  // [2][d] block partials of the LayerNorm parameter gradients
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  float pw[kBwdMaxV], pb[kBwdMaxV], lw[kBwdMaxV];
#pragma unroll
  for (int i = 0; i < kBwdMaxV; ++i) {
    pw[i] = pb[i] = 0.f;
    const int c = lane + 32 * i;
    lw[i] = c < d ? ln_w[c] : 0.f;
  }
  const float inv_d = 1.f / (float)d;
  for (int64_t v = w0; v < num_local; v += warps) {
    const float inv_cnt = 1.f / (float)max(indeg[v], 1);
    const float* up = upd + v * d;
    const float* hv = h + (dst_lo + v) * d;
    const float* go = g_out + v * d;
    float x[kBwdMaxV], gy[kBwdMaxV], dm[kBwdMaxV];   // dm: dropout multiplier (0 or 1 / keep; 1 without dropout)
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kBwdMaxV; ++i) {
      const int c = lane + 32 * i;
      x[i] = gy[i] = 0.f;
      dm[i] = 1.f;
      if (c < d) {
        if (da.keep > 0.f) dm[i] = fuse::dropout_mult1(da, (uint64_t)(dst_lo + v) * (uint64_t)d + c);
        x[i] = fmaxf(up[c] + hv[c], 0.f) * dm[i];        // what the LayerNorm saw
        gy[i] = go[c];
        sum += x[i];
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float mean = sum * inv_d;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < kBwdMaxV; ++i)
      if (lane + 32 * i < d) var += (x[i] - mean) * (x[i] - mean);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
    const float rstd = rsqrtf(var * inv_d + eps);
    float s1 = 0.f, s2 = 0.f;
    float xh[kBwdMaxV];
#pragma unroll
    for (int i = 0; i < kBwdMaxV; ++i) {
      xh[i] = (x[i] - mean) * rstd;
      if (lane + 32 * i < d) {
        const float gh = gy[i] * lw[i];
        s1 += gh;
        s2 += gh * xh[i];
        pw[i] += gy[i] * xh[i];
        pb[i] += gy[i];
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, s);
      s2 += __shfl_xor_sync(0xffffffffu, s2, s);
    }
    s1 *= inv_d;
    s2 *= inv_d;
#pragma unroll
    for (int i = 0; i < kBwdMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < d) {
        const float gx = rstd * (gy[i] * lw[i] - s1 - xh[i] * s2) * dm[i];
        const float gp = x[i] > 0.f ? gx : 0.f;  // relu'(0) = 0, as torch (x > 0 iff kept and relu input > 0)
        g_pre[v * d + c] = gp;
        g_acc[v * d + c] = gp * inv_cnt;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kBwdMaxV; ++i) {
    const int c = lane + 32 * i;
    if (c < d) {
      atomicAdd(&red[c], pw[i]);
      atomicAdd(&red[d + c], pb[i]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(&g_lnw[c], red[c]);
    atomicAdd(&g_lnb[c], red[d + c]);
  }
}

// Any hidden_dim beyond 32 * kBwdMaxV: the row lives in shared memory instead of registers (one warp per row, four
// warps per CTA: 8 d floats of row buffers + 2 d of LayerNorm-gradient partials), four passes over it.
__global__ void __launch_bounds__(128)
mp_epilogue_bwd_big_kernel(const float* __restrict__ g_out, const float* __restrict__ upd,
                           const int32_t* __restrict__ indeg, const float* __restrict__ h, int64_t dst_lo,
                           int64_t num_local, int d, const float* __restrict__ ln_w, float eps,
                           float* __restrict__ g_pre, float* __restrict__ g_acc, float* __restrict__ g_lnw,
                           float* __restrict__ g_lnb, const DropoutArgs da) {
  extern __shared__ float sm[];  // red[2][d] | per warp: x[d], gy[d]
  float* red = sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* xs = sm + 2 * d + (size_t)warp * 2 * d;
  float* gys = xs + d;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  const float inv_d = 1.f / (float)d;
  for (int64_t v = w0; v < num_local; v += warps) {
    const float inv_cnt = 1.f / (float)max(indeg[v], 1);
    const float* up = upd + v * d;
    const float* hv = h + (dst_lo + v) * d;
    const float* go = g_out + v * d;
    const uint64_t e_row = (uint64_t)(dst_lo + v) * (uint64_t)d;
    float sum = 0.f;
    for (int c = lane; c < d; c += 32) {
      float x = fmaxf(up[c] + hv[c], 0.f);
      if (da.keep > 0.f) x *= fuse::dropout_mult1(da, e_row + c);   // what the LayerNorm saw
      xs[c] = x;
      gys[c] = go[c];
      sum += x;
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
    const float mean = sum * inv_d;
    float var = 0.f;
    for (int c = lane; c < d; c += 32) var += (xs[c] - mean) * (xs[c] - mean);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
    const float rstd = rsqrtf(var * inv_d + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float xh = (xs[c] - mean) * rstd, gh = gys[c] * ln_w[c];
      s1 += gh;
      s2 += gh * xh;
      atomicAdd(&red[c], gys[c] * xh);
      atomicAdd(&red[d + c], gys[c]);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, s);
      s2 += __shfl_xor_sync(0xffffffffu, s2, s);
    }
    s1 *= inv_d;
    s2 *= inv_d;
    for (int c = lane; c < d; c += 32) {
      const float xh = (xs[c] - mean) * rstd;
      float gx = rstd * (gys[c] * ln_w[c] - s1 - xh * s2);
      if (da.keep > 0.f) gx *= fuse::dropout_mult1(da, e_row + c);
      const float gp = xs[c] > 0.f ? gx : 0.f;
      g_pre[v * d + c] = gp;
      g_acc[v * d + c] = gp * inv_cnt;
    }
    __syncwarp();                                        // the row buffers are reused by this warp's next row
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(&g_lnw[c], red[c]);
    atomicAdd(&g_lnb[c], red[d + c]);
  }
}

// The same for hidden_dim 32 / 64 / 128: a lane owns D/32 CONSECUTIVE columns (one vector access per row and
// operand), four rows in flight per warp; optionally max |g_acc| for the fp16 shadow the next kernels gather from.
template <int D, bool DROP>
__global__ void __launch_bounds__(256)
mp_epilogue_bwd_vec_kernel(const float* __restrict__ g_out, const float* __restrict__ upd,
                           const int32_t* __restrict__ indeg, const float* __restrict__ h, int64_t dst_lo,
                           int64_t num_local, const float* __restrict__ ln_w, float eps, float* __restrict__ g_pre,
                           float* __restrict__ g_acc, float* __restrict__ g_lnw, float* __restrict__ g_lnb,
                           float* __restrict__ g_acc_scale, const DropoutArgs da) {
  using namespace fuse;
  constexpr int V = D / 32;
  constexpr int kRows = 4;
  __shared__ float red[2 * D];
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  float lw[V], pw[V], pb[V];
  vload_nc<V>(lw, ln_w + lane * V);
#pragma unroll
  for (int j = 0; j < V; ++j) pw[j] = pb[j] = 0.f;
  float amax = 0.f;
  const float inv_d = 1.f / (float)D;
  for (int64_t r0 = w0; r0 < num_local; r0 += kRows * warps) {
    float up[kRows][V], hv[kRows][V], gy[kRows][V];
    int deg[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = r0 + k * warps;
      if (r < num_local) {
        vload_nc<V>(up[k], upd + r * D + lane * V);
        vload_nc<V>(hv[k], h + (dst_lo + r) * D + lane * V);
        vload_nc<V>(gy[k], g_out + r * D + lane * V);
        deg[k] = __ldg(indeg + r);
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = r0 + k * warps;
      if (r >= num_local) break;
      const float inv_cnt = 1.f / (float)max(deg[k], 1);
      float x[V], dm[V], sum = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) dm[j] = 1.f;
      if constexpr (DROP) dropout_multv<V>(da, (uint64_t)(dst_lo + r) * D + lane * V, dm);
#pragma unroll
      for (int j = 0; j < V; ++j) {
        x[j] = fmaxf(up[k][j] + hv[k][j], 0.f) * dm[j];  // what the LayerNorm saw
        sum += x[j];
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
      const float mean = sum * inv_d;
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) var += (x[j] - mean) * (x[j] - mean);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
      const float rstd = rsqrtf(var * inv_d + eps);
      float xh[V], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        xh[j] = (x[j] - mean) * rstd;
        const float gh = gy[k][j] * lw[j];
        s1 += gh;
        s2 += gh * xh[j];
        pw[j] += gy[k][j] * xh[j];
        pb[j] += gy[k][j];
      }
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, s);
        s2 += __shfl_xor_sync(0xffffffffu, s2, s);
      }
      s1 *= inv_d;
      s2 *= inv_d;
      float gp[V], ga[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const float gx = rstd * (gy[k][j] * lw[j] - s1 - xh[j] * s2) * dm[j];
        gp[j] = x[j] > 0.f ? gx : 0.f;
        ga[j] = gp[j] * inv_cnt;
        amax = fmaxf(amax, fabsf(ga[j]));
      }
      vstore<V>(g_pre + r * D + lane * V, gp);
      vstore<V>(g_acc + r * D + lane * V, ga);
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    atomicAdd(&red[lane * V + j], pw[j]);
    atomicAdd(&red[D + lane * V + j], pb[j]);
  }
  if (g_acc_scale) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
    if (lane == 0) atomic_max_nonneg(g_acc_scale + 1, amax);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(&g_lnw[c], red[c]);
    atomicAdd(&g_lnb[c], red[D + c]);
  }
}

// ---- weight gradients, CUDA cores (any hidden_dim): one CTA per (unit, 64-row tile of [W_msg ; W_self], 64-column
// tile); the unit's edges are the contraction dimension, staged 32 at a time.
constexpr int kWgTM = 64, kWgTN = 64, kWgTE = 32;

__global__ void __launch_bounds__(256)
mp_wgrad_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
                const int32_t* __restrict__ unit_rel, const int32_t* __restrict__ src_sorted,
                const int32_t* __restrict__ dst_sorted, const float* __restrict__ h, const float* __restrict__ g_acc,
                int64_t dst_lo, int d, float* __restrict__ gW_msg, float* __restrict__ gW_self,
                float* __restrict__ gbias) {
  __shared__ __align__(16) float sA[kWgTE][kWgTM];
  __shared__ __align__(16) float sG[kWgTE][kWgTN];
  __shared__ int64_t s_src[kWgTE], s_dst[kWgTE];
  const int64_t u = blockIdx.x;
  const int m0 = blockIdx.y * kWgTM, n0 = blockIdx.z * kWgTN;
  const int start = unit_start[u], count = unit_count[u];
  const int64_t r = unit_rel[u];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  float bsum[4] = {};
  for (int e0 = 0; e0 < count; e0 += kWgTE) {
    const int rows = min(kWgTE, count - e0);
    __syncthreads();
    if (threadIdx.x < kWgTE) {
      const bool ok = (int)threadIdx.x < rows;
      s_src[threadIdx.x] = ok ? src_sorted[start + e0 + threadIdx.x] : -1;
      s_dst[threadIdx.x] = ok ? dst_sorted[start + e0 + threadIdx.x] : -1;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kWgTE * kWgTM; idx += 256) {
      const int e = idx / kWgTM, m = m0 + idx % kWgTM;
      float a = 0.f;
      if (s_src[e] >= 0 && m < 2 * d)
        a = m < d ? h[s_src[e] * d + m] : h[(dst_lo + s_dst[e]) * d + (m - d)];
      sA[e][idx % kWgTM] = a;
      const int n = n0 + idx % kWgTN;
      sG[e][idx % kWgTN] = (s_src[e] >= 0 && n < d) ? g_acc[s_dst[e] * d + n] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int e = 0; e < kWgTE; ++e) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[e][ty * 4]);
      const float4 g = *reinterpret_cast<const float4*>(&sG[e][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], gv[j], acc[i][j]);
      if (blockIdx.y == 0 && ty == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += gv[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= 2 * d) continue;
    float* row = m < d ? gW_msg + (r * d + m) * d : gW_self + (r * d + (m - d)) * d;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d) atomicAdd(row + n, acc[i][j]);
    }
  }
  if (blockIdx.y == 0 && ty == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < d) atomicAdd(gbias + r * d + n, bsum[j]);
    }
  }
}

}  // namespace
}  // namespace ghf

using namespace ghf;

static int epilogue_backward_impl(const ghf_graph* g, const float* d_g_out, const float* d_upd, const float* d_h,
                                  const float* d_ln_w, float eps, float* d_g_pre, float* d_g_acc, float* d_g_ln_w,
                                  float* d_g_ln_b, float* d_g_acc_scale, const DropoutArgs da, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(g != nullptr, "ghf_mp_epilogue_backward: graph is NULL");
  const int d = g->hidden_dim;
  GHF_REQUIRE(d_g_out && d_upd && d_h && d_ln_w && d_g_pre && d_g_acc && d_g_ln_w && d_g_ln_b,
              "ghf_mp_epilogue_backward: NULL argument");
  GHF_CUDA(cudaMemsetAsync(d_g_ln_w, 0, d * sizeof(float), stream));
  GHF_CUDA(cudaMemsetAsync(d_g_ln_b, 0, d * sizeof(float), stream));
  if (d_g_acc_scale) GHF_CUDA(cudaMemsetAsync(d_g_acc_scale, 0, 2 * sizeof(float), stream));
  if (g->num_local == 0) return 0;
  const int64_t want = cdiv(g->num_local, 8 * 4);
  const int64_t cap = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  const bool aligned = (reinterpret_cast<uintptr_t>(d_g_out) | reinterpret_cast<uintptr_t>(d_upd) |
                        reinterpret_cast<uintptr_t>(d_h) | reinterpret_cast<uintptr_t>(d_ln_w) |
                        reinterpret_cast<uintptr_t>(d_g_pre) | reinterpret_cast<uintptr_t>(d_g_acc)) % 16 == 0;
#define GHF_BWD_VEC(D, DROP)                                                                                        \
  mp_epilogue_bwd_vec_kernel<D, DROP><<<grid, 256, 0, stream>>>(d_g_out, d_upd, g->indeg, d_h, g->dst_lo,           \
                                                                g->num_local, d_ln_w, eps, d_g_pre, d_g_acc,        \
                                                                d_g_ln_w, d_g_ln_b, d_g_acc_scale, da)
  const bool drop = da.keep > 0.f;
  if (aligned && d == 128) { if (drop) GHF_BWD_VEC(128, true); else GHF_BWD_VEC(128, false); }
  else if (aligned && d == 64) { if (drop) GHF_BWD_VEC(64, true); else GHF_BWD_VEC(64, false); }
  else if (aligned && d == 32) { if (drop) GHF_BWD_VEC(32, true); else GHF_BWD_VEC(32, false); }
  else if (d > 32 * kBwdMaxV) {                          // the row does not fit in registers: shared-memory variant
    GHF_REQUIRE(d_g_acc_scale == nullptr, "ghf_mp_epilogue_backward: max|g_acc| needs hidden_dim 32/64/128");
    const size_t smem = (size_t)(2 + 2 * 4) * d * sizeof(float);
    GHF_REQUIRE(smem <= 200 * 1024, "ghf_mp_epilogue_backward: hidden_dim %d needs %zu bytes of shared memory", d, smem);
    static bool configured[64] = {false};
    if (first_use_on_device(configured))
      GHF_CUDA(cudaFuncSetAttribute(mp_epilogue_bwd_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int64_t want_big = cdiv(g->num_local, 4);
    mp_epilogue_bwd_big_kernel<<<(unsigned)(want_big < cap ? want_big : cap), 128, smem, stream>>>(
        d_g_out, d_upd, g->indeg, d_h, g->dst_lo, g->num_local, d, d_ln_w, eps, d_g_pre, d_g_acc, d_g_ln_w, d_g_ln_b,
        da);
  } else {
    GHF_REQUIRE(d_g_acc_scale == nullptr, "ghf_mp_epilogue_backward: max|g_acc| needs hidden_dim 32/64/128");
    mp_epilogue_bwd_kernel<<<grid, 256, 2 * d * sizeof(float), stream>>>(
        d_g_out, d_upd, g->indeg, d_h, g->dst_lo, g->num_local, d, d_ln_w, eps, d_g_pre, d_g_acc, d_g_ln_w, d_g_ln_b,
        da);
  }
#undef GHF_BWD_VEC
  GHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int ghf_mp_epilogue_backward(const ghf_graph* g, const float* d_g_out, const float* d_upd,
                                        const float* d_h, const float* d_ln_w, float eps, float* d_g_pre,
                                        float* d_g_acc, float* d_g_ln_w, float* d_g_ln_b, float* d_g_acc_scale,
                                        void* stream_) {
  return epilogue_backward_impl(g, d_g_out, d_upd, d_h, d_ln_w, eps, d_g_pre, d_g_acc, d_g_ln_w, d_g_ln_b,
                                d_g_acc_scale, DropoutArgs(), stream_);
}

extern "C" int ghf_mp_epilogue_backward_dropout(const ghf_graph* g, const float* d_g_out, const float* d_upd,
                                                const float* d_h, const float* d_ln_w, float eps, float p_drop,
                                                uint64_t seed, uint64_t offset, float* d_g_pre, float* d_g_acc,
                                                float* d_g_ln_w, float* d_g_ln_b, float* d_g_acc_scale,
                                                void* stream_) {
  GHF_REQUIRE(g != nullptr, "ghf_mp_epilogue_backward_dropout: graph is NULL");
  DropoutArgs da;
  if (int rc = make_dropout_args(p_drop, seed, offset, g->num_nodes * (int64_t)g->hidden_dim, &da, nullptr)) return rc;
  return epilogue_backward_impl(g, d_g_out, d_upd, d_h, d_ln_w, eps, d_g_pre, d_g_acc, d_g_ln_w, d_g_ln_b,
                                d_g_acc_scale, da, stream_);
}

extern "C" int ghf_mp_weight_grad(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                                  const float* d_g_acc, const void* d_g16, const float* d_g16_scale, int precision,
                                  float* d_gW_msg, float* d_gW_self, float* d_gbias, void* d_workspace,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(g != nullptr, "ghf_mp_weight_grad: graph is NULL");
  GHF_REQUIRE(d_h && d_g_acc && d_gW_msg && d_gW_self && d_gbias, "ghf_mp_weight_grad: NULL argument");
  const int d = g->hidden_dim;
  const size_t wbytes = (size_t)g->num_rel * d * d * sizeof(float);
  GHF_CUDA(cudaMemsetAsync(d_gW_msg, 0, wbytes, stream));
  GHF_CUDA(cudaMemsetAsync(d_gW_self, 0, wbytes, stream));
  GHF_CUDA(cudaMemsetAsync(d_gbias, 0, (size_t)g->num_rel * d * sizeof(float), stream));
  if (g->num_units == 0) return 0;
  if (precision == GHF_PREC_F16 && mp_f16_supported(d)) {
    // tensor-core path (mp_wgrad_f16.cu).  Workspace regions as in ghf_mp_layer: [sync words][accumulator rows]
    // [weight images][scale words][fp16 h]; the accumulator region holds the fp16 shadow of g_acc here.
    GHF_REQUIRE(d_workspace != nullptr, "ghf_mp_weight_grad: workspace is NULL");
    GHF_REQUIRE(d_h16 == nullptr || d_h16_scale != nullptr, "ghf_mp_weight_grad: d_h16 needs d_h16_scale");
    int* counter = reinterpret_cast<int*>(align_up(reinterpret_cast<int64_t>(d_workspace), 256));
    char* acc_ws = reinterpret_cast<char*>(counter) + mp_f16_sync_bytes(g);
    char* pack = acc_ws + align_up(g->num_local * (int64_t)d * 4, 256);
    float* sc = reinterpret_cast<float*>(pack + mp_f16_pack_bytes(g->num_rel));
    const void* h16 = d_h16;
    const float* h_scale = d_h16_scale;
    if (h16 == nullptr) {
      void* conv = reinterpret_cast<char*>(sc) + 256;
      if (int rc = mp_f16_absmax(d_h, g->num_nodes * (int64_t)d, sc, stream)) return rc;
      if (int rc = mp_f16_convert(d_h, g->num_nodes * (int64_t)d, conv, sc, /*rescue=*/false, stream)) return rc;
      h16 = conv;
      h_scale = sc;
    }
    const void* g16 = d_g16;
    const float* g_scale = d_g16_scale;
    GHF_REQUIRE(g16 == nullptr || g_scale != nullptr, "ghf_mp_weight_grad: d_g16 needs d_g16_scale");
    if (g16 == nullptr) {
      float* gs = sc + 16;
      if (int rc = mp_f16_absmax(d_g_acc, g->num_local * (int64_t)d, gs, stream)) return rc;
      if (int rc = mp_f16_convert(d_g_acc, g->num_local * (int64_t)d, acc_ws, gs, /*rescue=*/false, stream)) return rc;
      g16 = acc_ws;
      g_scale = gs;
    }
    return mp_wgrad_f16_launch(g, h16, h_scale, g16, g_scale, d_gW_msg, d_gW_self, d_gbias, counter, stream);
  }
  const dim3 grid((unsigned)g->num_units, (unsigned)cdiv(2 * d, kWgTM), (unsigned)cdiv(d, kWgTN));
  GHF_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "ghf_mp_weight_grad: hidden_dim %d too large", d);
  mp_wgrad_kernel<<<grid, 256, 0, stream>>>(g->unit_start, g->unit_count, g->unit_rel, g->src_sorted, g->dst_sorted,
                                            d_h, d_g_acc, g->dst_lo, d, d_gW_msg, d_gW_self, d_gbias);
  GHF_LAUNCH_CHECK();
  return 0;
}
