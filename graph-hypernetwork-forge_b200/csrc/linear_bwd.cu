// linear_bwd.cu — ghf_linear_backward: the gradients of Y = alpha * act(X W^T + b) (ghf_linear; HG:261, WG:97-107,
// WG:138-140) in two fp32 kernels, replacing the three library GEMMs + elementwise passes of round 1:
//
//   g_pre[m,n] = gY[m,n] * [Y[m,n] > 0]        (ReLU; alpha = exp(log_scale) > 0, so Y and the pre-activation share
//                                               their sign; without ReLU g_pre = gY)
//   gX[m,k]    = alpha * sum_n g_pre[m,n] W[n,k]        linear_bwd_gx_kernel   (the contraction index n is split over
//                                                        blockIdx.z when there are few row tiles: generator heads)
//   gW[n,k]    = alpha * sum_m g_pre[m,n] X[m,k]        linear_bwd_gw_kernel   (m split over blockIdx.z: 2.5M rows for
//   gb[n]      = alpha * sum_m g_pre[m,n]                the input projection); the k-tile-0 CTAs also sum gb and
//   g_ls       = sum_{m,n} gY[m,n] Y[m,n]                g_ls from the tiles they load anyway
//
// The ReLU mask is applied while the tiles are loaded, so g_pre is never written.  Both operands of gW have the
// contraction index as their slow index (rows of gY and of X), i.e. the tile rows are contiguous in memory exactly as
// the FFMA engine keeps them in shared memory (A[k][row]): float4 loads go straight to float4 stores, no transpose.
#include "ffma_gemm.cuh"
#include "ghf_b200.h"

namespace ghf {
namespace {

__device__ __forceinline__ float relu_mask(float g, float y) { return y > 0.f ? g : 0.f; }
__device__ __forceinline__ float4 relu_mask4(float4 g, float4 y) {
  return make_float4(relu_mask(g.x, y.x), relu_mask(g.y, y.y), relu_mask(g.z, y.z), relu_mask(g.w, y.w));
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(kFfmaThreads)
linear_bwd_gx_kernel(const float* __restrict__ gY, const float* __restrict__ Y, const float* __restrict__ W, int64_t M,
                     int N, int K, int relu, const float* __restrict__ log_scale, int n_per_split,
                     float* __restrict__ gX, int atomic) {
  __shared__ FfmaSmem<BN> sm;
  constexpr int TN = BN / 16;
  const int64_t m0 = (int64_t)blockIdx.x * kFfmaBM;
  const int k0 = blockIdx.y * BN;
  const int nb = blockIdx.z * n_per_split;
  const int ne = min(N, nb + n_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // A(row = m, kk = n): g_pre, contiguous along the contraction index
  auto loadA = [&](int row, int kk) -> float {
    const int64_t m = m0 + row;
    const int n = nb + kk;
    if (m >= M || n >= ne) return 0.f;
    const float g = gY[m * N + n];
    return relu ? relu_mask(g, Y[m * N + n]) : g;
  };
  auto loadA4 = [&](int row, int kk) -> float4 {   // N % 4 == 0 and nb % 16 == 0: n < ne implies n + 3 < ne
    const int64_t m = m0 + row;
    const int n = nb + kk;
    if (m >= M || n >= ne) return make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 g = *reinterpret_cast<const float4*>(gY + m * N + n);
    return relu ? relu_mask4(g, *reinterpret_cast<const float4*>(Y + m * N + n)) : g;
  };
  // B(kk = n, col = k): W[n, k], contiguous along the column
  auto loadB = [&](int kk, int col) -> float {
    const int n = nb + kk, k = k0 + col;
    return (n < ne && k < K) ? W[(int64_t)n * K + k] : 0.f;
  };
  auto loadB4 = [&](int kk, int col) -> float4 {
    const int n = nb + kk, k = k0 + col;
    return (n < ne && k < K) ? *reinterpret_cast<const float4*>(W + (int64_t)n * K + k)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  ffma_mainloop<BN, VEC, /*B_KMAJOR=*/false>(sm, ne - nb, loadA, loadA4, loadB, loadB4, acc);

  const float alpha = log_scale ? expf(*log_scale) : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ffma_row(ty, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int k = k0 + ffma_col<BN>(tx, j);
      if (k >= K) continue;
      if (atomic) atomicAdd(&gX[m * K + k], acc[i][j] * alpha);
      else gX[m * K + k] = acc[i][j] * alpha;
    }
  }
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(kFfmaThreads)
linear_bwd_gw_kernel(const float* __restrict__ gY, const float* __restrict__ Y, const float* __restrict__ X, int64_t M,
                     int N, int K, int relu, const float* __restrict__ log_scale, int64_t m_per_split,
                     float* __restrict__ gW, float* __restrict__ gb, float* __restrict__ gls) {
  __shared__ FfmaSmem<BN> sm;
  __shared__ float side_b[kFfmaBM];
  __shared__ float side_ls;
  constexpr int TN = BN / 16;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * kFfmaBM;
  const int k0 = blockIdx.y * BN;
  const int64_t mb = (int64_t)blockIdx.z * m_per_split;
  const int64_t me = mb + m_per_split < M ? mb + m_per_split : M;
  const bool want_b = gW == nullptr ? gb != nullptr : (blockIdx.y == 0 && gb != nullptr);
  const bool want_ls = gW == nullptr ? gls != nullptr : (blockIdx.y == 0 && gls != nullptr);
  const bool need_y = relu || want_ls;
  if (tid < kFfmaBM) side_b[tid] = 0.f;
  if (tid == 0) side_ls = 0.f;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // per k-tile of 16 contraction rows: A = g_pre[m, n0 .. n0+128), B = X[m, k0 .. k0+BN)
  constexpr int AV = VEC ? kFfmaBM * kFfmaBK / 4 / kFfmaThreads : kFfmaBM * kFfmaBK / kFfmaThreads;          // 2 | 8
  constexpr int BV = VEC ? (BN * kFfmaBK / 4 + kFfmaThreads - 1) / kFfmaThreads : BN * kFfmaBK / kFfmaThreads;
  float4 ra4[VEC ? AV : 1], rb4[VEC ? BV : 1];
  float ras[VEC ? 1 : AV], rbs[VEC ? 1 : BV];
  float ps[4] = {0.f, 0.f, 0.f, 0.f};   // bias-gradient partials of this thread's rows (VEC: 4 rows, else 1)
  float pl = 0.f;                       // log-scale-gradient partial

  auto fetch = [&](int64_t mt) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        const int f = tid + i * kFfmaThreads;
        const int64_t m = mt + (f >> 5);
        const int n = n0 + (f & 31) * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < me && n < N) {                      // N % 4 == 0: n < N implies n + 3 < N
          g = *reinterpret_cast<const float4*>(gY + m * N + n);
          if (need_y) {
            const float4 y = *reinterpret_cast<const float4*>(Y + m * N + n);
            if (want_ls) pl += g.x * y.x + g.y * y.y + g.z * y.z + g.w * y.w;
            if (relu) g = relu_mask4(g, y);
          }
          if (want_b) { ps[0] += g.x; ps[1] += g.y; ps[2] += g.z; ps[3] += g.w; }
        }
        ra4[i] = g;
      }
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK / 4) {
          const int64_t m = mt + f / (BN / 4);
          const int k = k0 + (f % (BN / 4)) * 4;
          rb4[i] = (X != nullptr && m < me && k < K) ? *reinterpret_cast<const float4*>(X + m * K + k)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        const int f = tid + i * kFfmaThreads;
        const int64_t m = mt + (f >> 7);
        const int n = n0 + (f & 127);
        float g = 0.f;
        if (m < me && n < N) {
          g = gY[m * N + n];
          if (need_y) {
            const float y = Y[m * N + n];
            if (want_ls) pl += g * y;
            if (relu) g = relu_mask(g, y);
          }
          if (want_b) ps[0] += g;
        }
        ras[i] = g;
      }
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        const int f = tid + i * kFfmaThreads;
        const int64_t m = mt + f / BN;
        const int k = k0 + f % BN;
        rbs[i] = (X != nullptr && m < me && k < K) ? X[m * K + k] : 0.f;
      }
    }
  };
  auto stash = [&](int buf) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        const int f = tid + i * kFfmaThreads;
        *reinterpret_cast<float4*>(&sm.A[buf][f >> 5][(f & 31) * 4]) = ra4[i];
      }
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK / 4)
          *reinterpret_cast<float4*>(&sm.B[buf][f / (BN / 4)][(f % (BN / 4)) * 4]) = rb4[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < AV; ++i) {
        const int f = tid + i * kFfmaThreads;
        sm.A[buf][f >> 7][f & 127] = ras[i];
      }
#pragma unroll
      for (int i = 0; i < BV; ++i) {
        const int f = tid + i * kFfmaThreads;
        sm.B[buf][f / BN][f % BN] = rbs[i];
      }
    }
  };

  const int64_t nk = me > mb ? (me - mb + kFfmaBK - 1) / kFfmaBK : 0;
  if (nk > 0) {
    fetch(mb);
    stash(0);
  }
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = (int)(kt & 1);
    if (kt + 1 < nk) fetch(mb + (kt + 1) * kFfmaBK);
    if (gW != nullptr) ffma_compute<BN>(sm, buf, tx, ty, acc);
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }

  const float alpha = log_scale ? expf(*log_scale) : 1.f;
  if (gW != nullptr && nk > 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = n0 + ffma_row(ty, i);
      if (n >= N) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int k = k0 + ffma_col<BN>(tx, j);
        if (k < K) atomicAdd(&gW[(int64_t)n * K + k], acc[i][j] * alpha);
      }
    }
  }
  if (want_b) {
    if constexpr (VEC) {
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(&side_b[(tid & 31) * 4 + c], ps[c]);
    } else {
      atomicAdd(&side_b[tid & 127], ps[0]);
    }
  }
  if (want_ls) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pl += __shfl_xor_sync(0xffffffffu, pl, o);
    if ((tid & 31) == 0) atomicAdd(&side_ls, pl);
  }
  if (want_b || want_ls) {
    __syncthreads();
    if (want_b && tid < kFfmaBM && n0 + tid < N) atomicAdd(&gb[n0 + tid], side_b[tid] * alpha);
    if (want_ls && tid == 0) atomicAdd(gls, side_ls);
  }
}

template <int BN>
int launch_gx(const float* gY, const float* Y, const float* W, int64_t M, int N, int K, int relu,
              const float* log_scale, float* gX, cudaStream_t stream) {
  const int64_t tiles = cdiv(M, kFfmaBM) * cdiv(K, BN);
  // few row tiles and a long contraction (the generator heads: 535 x 16384): split n, reduce with atomics
  int64_t splits = tiles >= sm_count() ? 1 : cdiv(2 * (int64_t)sm_count(), tiles);
  if (splits > cdiv(N, 64)) splits = cdiv(N, 64);
  if (splits < 1) splits = 1;
  const int n_per_split = (int)align_up(cdiv(N, splits), 16);
  splits = cdiv(N, n_per_split);
  const bool vec = N % 4 == 0 && K % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(gY) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(W)) % 16 == 0;
  if (splits > 1) GHF_CUDA(cudaMemsetAsync(gX, 0, (size_t)M * K * sizeof(float), stream));
  dim3 grid((unsigned)cdiv(M, kFfmaBM), (unsigned)cdiv(K, BN), (unsigned)splits);
  if (vec)
    linear_bwd_gx_kernel<BN, true><<<grid, kFfmaThreads, 0, stream>>>(gY, Y, W, M, N, K, relu, log_scale, n_per_split,
                                                                      gX, splits > 1);
  else
    linear_bwd_gx_kernel<BN, false><<<grid, kFfmaThreads, 0, stream>>>(gY, Y, W, M, N, K, relu, log_scale, n_per_split,
                                                                       gX, splits > 1);
  GHF_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int launch_gw(const float* gY, const float* Y, const float* X, int64_t M, int N, int K, int relu,
              const float* log_scale, float* gW, float* gb, float* gls, cudaStream_t stream) {
  const int64_t tiles = cdiv(N, kFfmaBM) * (gW ? cdiv(K, BN) : 1);
  int64_t splits = tiles >= 2 * sm_count() ? 1 : cdiv(2 * (int64_t)sm_count(), tiles);
  if (splits > cdiv(M, 64)) splits = cdiv(M, 64);
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  const int64_t m_per_split = align_up(cdiv(M, splits), kFfmaBK);
  splits = cdiv(M, m_per_split);
  const bool vec = N % 4 == 0 && K % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(gY) | reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(X)) % 16 == 0;
  dim3 grid((unsigned)cdiv(N, kFfmaBM), (unsigned)(gW ? cdiv(K, BN) : 1), (unsigned)splits);
  if (vec)
    linear_bwd_gw_kernel<BN, true><<<grid, kFfmaThreads, 0, stream>>>(gY, Y, X, M, N, K, relu, log_scale, m_per_split,
                                                                      gW, gb, gls);
  else
    linear_bwd_gw_kernel<BN, false><<<grid, kFfmaThreads, 0, stream>>>(gY, Y, X, M, N, K, relu, log_scale,
                                                                       m_per_split, gW, gb, gls);
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" int ghf_linear_backward(const float* d_X, int64_t M, int K, const float* d_W, int N, int relu,
                                   const float* d_log_scale, const float* d_Y, const float* d_gY, float* d_gX,
                                   float* d_gW, float* d_gb, float* d_gls, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(M >= 0 && K > 0 && N > 0, "ghf_linear_backward: bad dims M=%lld K=%d N=%d", (long long)M, K, N);
  GHF_REQUIRE(d_gY != nullptr || M == 0, "ghf_linear_backward: d_gY is NULL");
  GHF_REQUIRE(d_Y != nullptr || M == 0 || (!relu && d_gls == nullptr),
              "ghf_linear_backward: d_Y is needed for the ReLU mask and for the log-scale gradient");
  GHF_REQUIRE(d_gX == nullptr || d_W != nullptr, "ghf_linear_backward: d_gX needs d_W");
  GHF_REQUIRE(d_gW == nullptr || d_X != nullptr, "ghf_linear_backward: d_gW needs d_X");
  GHF_REQUIRE(d_gls == nullptr || d_log_scale != nullptr, "ghf_linear_backward: d_gls without d_log_scale");
  GHF_REQUIRE(cdiv(K, 32) <= 65535 && cdiv(N, kFfmaBM) <= 0x7FFFFFFF, "ghf_linear_backward: dims too large");
  // the sums over rows are reduced with atomics: start from zero
  if (d_gW) GHF_CUDA(cudaMemsetAsync(d_gW, 0, (size_t)N * K * sizeof(float), stream));
  if (d_gb) GHF_CUDA(cudaMemsetAsync(d_gb, 0, (size_t)N * sizeof(float), stream));
  if (d_gls) GHF_CUDA(cudaMemsetAsync(d_gls, 0, sizeof(float), stream));
  if (M == 0) return 0;
  const int bn = K <= 32 ? 32 : (K <= 64 ? 64 : 128);
  if (d_gX) {
    int rc = bn == 32   ? launch_gx<32>(d_gY, d_Y, d_W, M, N, K, relu, d_log_scale, d_gX, stream)
             : bn == 64 ? launch_gx<64>(d_gY, d_Y, d_W, M, N, K, relu, d_log_scale, d_gX, stream)
                        : launch_gx<128>(d_gY, d_Y, d_W, M, N, K, relu, d_log_scale, d_gX, stream);
    if (rc) return rc;
  }
  if (d_gW || d_gb || d_gls) {
    int rc = bn == 32   ? launch_gw<32>(d_gY, d_Y, d_X, M, N, K, relu, d_log_scale, d_gW, d_gb, d_gls, stream)
             : bn == 64 ? launch_gw<64>(d_gY, d_Y, d_X, M, N, K, relu, d_log_scale, d_gW, d_gb, d_gls, stream)
                        : launch_gw<128>(d_gY, d_Y, d_X, M, N, K, relu, d_log_scale, d_gW, d_gb, d_gls, stream);
    if (rc) return rc;
  }
  return 0;
}
