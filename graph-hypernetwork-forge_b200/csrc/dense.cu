// dense.cu — ghf_linear: Y = alpha * act(X W^T + b), fp32 FFMA.
// Replaces F.relu(input_proj(x)) (HG:261), the generator MLP Linears (WG:97-107) and the scaled
// final Linear (WG:138-140: flat * exp(log_scale)).
#include "ffma_gemm.cuh"
#include "ghf_b200.h"

namespace ghf {
// tcgen05 3xTF32 path for K = 128, N a multiple of 128 and enough outputs to fill the machine (linear_umma.cu)
bool linear_umma_eligible(int64_t M, int K, int N, int relu, const void* log_scale, const void* X, const void* W,
                          const void* Y);
int linear_umma_launch(const float* X, int64_t M, const float* W, const float* b, int N, int relu,
                       const float* log_scale, float* Y, void* Y16, float* y16_scale, cudaStream_t stream);
int mp_f16_absmax(const float* x, int64_t elems, float* scale, cudaStream_t stream);                      // mp_f16.cu
int mp_f16_convert(const float* h, int64_t elems, void* h16, float* scale, bool rescue, cudaStream_t stream);
namespace {

template <int BN, bool VEC>
__global__ void __launch_bounds__(kFfmaThreads)
linear_kernel(const float* __restrict__ X, int64_t M, int K, const float* __restrict__ W,
              const float* __restrict__ b, int N, int relu, const float* __restrict__ log_scale,
              float* __restrict__ Y) {
  __shared__ FfmaSmem<BN> sm;
  constexpr int TN = BN / 16;
  const int64_t m0 = (int64_t)blockIdx.x * kFfmaBM;
  const int n0 = blockIdx.y * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto loadA = [&](int row, int k) -> float {
    const int64_t m = m0 + row;
    return (m < M && k < K) ? X[m * K + k] : 0.f;
  };
  auto loadA4 = [&](int row, int k) -> float4 {
    const int64_t m = m0 + row;
    return (m < M && k < K) ? *reinterpret_cast<const float4*>(X + m * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  auto loadB = [&](int k, int n) -> float {
    const int nn = n0 + n;
    return (nn < N && k < K) ? W[(int64_t)nn * K + k] : 0.f;
  };
  auto loadB4 = [&](int k, int n) -> float4 {
    const int nn = n0 + n;
    return (nn < N && k < K) ? *reinterpret_cast<const float4*>(W + (int64_t)nn * K + k)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  ffma_mainloop<BN, VEC, /*B_KMAJOR=*/true>(sm, K, loadA, loadA4, loadB, loadB4, acc);

  const float alpha = log_scale ? expf(*log_scale) : 1.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ffma_row(ty, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + ffma_col<BN>(tx, j);
      if (n >= N) continue;
      float v = acc[i][j] + (b ? b[n] : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      Y[m * N + n] = v * alpha;
    }
  }
}

template <int BN>
int launch_linear(const float* X, int64_t M, int K, const float* W, const float* b, int N, int relu,
                  const float* log_scale, float* Y, cudaStream_t stream) {
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) % 16 == 0);
  dim3 grid((unsigned)cdiv(M, kFfmaBM), (unsigned)cdiv(N, BN));
  if (vec)
    linear_kernel<BN, true><<<grid, kFfmaThreads, 0, stream>>>(X, M, K, W, b, N, relu, log_scale, Y);
  else
    linear_kernel<BN, false><<<grid, kFfmaThreads, 0, stream>>>(X, M, K, W, b, N, relu, log_scale, Y);
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" int ghf_linear(const float* d_X, int64_t M, int K, const float* d_W, const float* d_b, int N,
                          int relu, const float* d_log_scale, float* d_Y, void* stream_) {
  return ghf_linear_f16out(d_X, M, K, d_W, d_b, N, relu, d_log_scale, d_Y, nullptr, nullptr, stream_);
}

extern "C" int ghf_linear_f16out(const float* d_X, int64_t M, int K, const float* d_W, const float* d_b, int N,
                                 int relu, const float* d_log_scale, float* d_Y, void* d_Y16, float* d_Y16_scale,
                                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(M >= 0 && K > 0 && N > 0, "ghf_linear: bad dims M=%lld K=%d N=%d", (long long)M, K, N);
  GHF_REQUIRE(cdiv(N, 32) <= 65535, "ghf_linear: N=%d too large", N);
  GHF_REQUIRE(reinterpret_cast<uintptr_t>(d_Y16) % 16 == 0, "ghf_linear_f16out: d_Y16 must be 16-byte aligned");
  GHF_REQUIRE(d_Y16 == nullptr || d_Y16_scale != nullptr, "ghf_linear_f16out: d_Y16 needs d_Y16_scale (float[2])");
  if (M == 0) return 0;
  if (linear_umma_eligible(M, K, N, relu, d_log_scale, d_X, d_W, d_Y)) {
    // fused: the kernel writes fp16(Y) and max|Y|; a rescue pass rewrites the shadow only if the range needs a scale
    if (d_Y16) GHF_CUDA(cudaMemsetAsync(d_Y16_scale + 1, 0, sizeof(float), stream));
    if (int rc = linear_umma_launch(d_X, M, d_W, d_b, N, relu, d_log_scale, d_Y, d_Y16, d_Y16_scale, stream)) return rc;
    return d_Y16 ? mp_f16_convert(d_Y, M * (int64_t)N, d_Y16, d_Y16_scale, /*rescue=*/true, stream) : 0;
  }
  // column-tile width: as wide as N allows, narrowed while the grid would leave most SMs idle (the generator's
  // hidden Linears have a few hundred rows)
  int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  while (bn > 32 && cdiv(M, kFfmaBM) * cdiv(N, bn) < sm_count()) bn >>= 1;
  int rc;
  if (bn == 32) rc = launch_linear<32>(d_X, M, K, d_W, d_b, N, relu, d_log_scale, d_Y, stream);
  else if (bn == 64) rc = launch_linear<64>(d_X, M, K, d_W, d_b, N, relu, d_log_scale, d_Y, stream);
  else rc = launch_linear<128>(d_X, M, K, d_W, d_b, N, relu, d_log_scale, d_Y, stream);
  if (rc == 0 && d_Y16) rc = mp_f16_absmax(d_Y, M * (int64_t)N, d_Y16_scale, stream);   // fused only on the tcgen05 path
  if (rc == 0 && d_Y16) rc = mp_f16_convert(d_Y, M * (int64_t)N, d_Y16, d_Y16_scale, /*rescue=*/false, stream);
  return rc;
}
