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

// A GROUP of equally shaped Linears in one launch (blockIdx.z = problem): the hidden layers of the weight generators -
// three MLPs per HyperGNN layer, all layers at once (they depend on the text embeddings only) - are a few hundred rows
// each; one launch of 3 L problems instead of 3 L launches of a few CTAs.
constexpr int kMaxGroup = 48;
struct LinearGroup {
  const float* X[kMaxGroup];
  const float* W[kMaxGroup];
  const float* b[kMaxGroup];
  const float* log_scale[kMaxGroup];
  float* Y[kMaxGroup];
};

template <int BN, bool VEC>
__global__ void __launch_bounds__(kFfmaThreads)
linear_group_kernel(const __grid_constant__ LinearGroup grp, int64_t M, int K, int N, int relu) {
  __shared__ FfmaSmem<BN> sm;
  constexpr int TN = BN / 16;
  const int z = blockIdx.z;
  const float* __restrict__ X = grp.X[z];
  const float* __restrict__ W = grp.W[z];
  const float* __restrict__ b = grp.b[z];
  float* __restrict__ Y = grp.Y[z];
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
  const float alpha = grp.log_scale[z] ? expf(*grp.log_scale[z]) : 1.f;
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
int launch_linear_group(const LinearGroup& grp, int count, int64_t M, int K, int N, int relu, cudaStream_t stream) {
  bool vec = K % 4 == 0;
  for (int z = 0; z < count; ++z)
    vec = vec && (reinterpret_cast<uintptr_t>(grp.X[z]) | reinterpret_cast<uintptr_t>(grp.W[z])) % 16 == 0;
  dim3 grid((unsigned)cdiv(M, kFfmaBM), (unsigned)cdiv(N, BN), (unsigned)count);
  if (vec)
    linear_group_kernel<BN, true><<<grid, kFfmaThreads, 0, stream>>>(grp, M, K, N, relu);
  else
    linear_group_kernel<BN, false><<<grid, kFfmaThreads, 0, stream>>>(grp, M, K, N, relu);
  GHF_LAUNCH_CHECK();
  return 0;
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

// ---- WG:120-143 for several generators at once ------------------------------------------------------------------
static int linear_group(const LinearGroup& grp, int count, int64_t M, int K, int N, int relu, cudaStream_t stream) {
  if (M == 0 || count == 0) return 0;
  int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  while (bn > 32 && cdiv(M, kFfmaBM) * cdiv(N, bn) * count < sm_count()) bn >>= 1;
  if (bn == 32) return launch_linear_group<32>(grp, count, M, K, N, relu, stream);
  if (bn == 64) return launch_linear_group<64>(grp, count, M, K, N, relu, stream);
  return launch_linear_group<128>(grp, count, M, K, N, relu, stream);
}

extern "C" int64_t ghf_weight_generators_scratch_bytes(int64_t U, int32_t H, int32_t depth, int32_t n_gen) {
  if (U < 0 || H < 0 || depth < 0 || n_gen < 0) return -1;
  return depth > 0 ? (int64_t)2 * n_gen * 3 * (U > 0 ? U : 1) * (H > 0 ? H : 1) * (int64_t)sizeof(float) + 256 : 256;
}

extern "C" int ghf_weight_generators(const float* d_text_emb, int64_t U, int32_t T, int32_t H, int32_t depth,
                                     int32_t n_gen, int32_t d_in, int32_t d_out, const float* const* h_params,
                                     const float* const* h_log_scales, float* const* h_out, void* d_scratch,
                                     int32_t skip_big, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(U >= 0 && T > 0 && depth >= 0 && (H > 0 || depth == 0) && n_gen >= 0 && d_in > 0 && d_out > 0,
              "ghf_weight_generators: bad dimensions");
  GHF_REQUIRE(3 * n_gen <= kMaxGroup, "ghf_weight_generators: at most %d generators per call", kMaxGroup / 3);
  GHF_REQUIRE(h_params && h_log_scales && h_out && (d_scratch || depth == 0), "ghf_weight_generators: NULL argument");
  if (U == 0 || n_gen == 0) return 0;
  const int n_prob = 3 * n_gen;
  const int per_mlp = 2 * (depth + 1);                       // {weight, bias} per Linear
  auto param = [&](int gen, int mlp, int lin, int which) {
    return h_params[((int64_t)gen * 3 + mlp) * per_mlp + 2 * lin + which];
  };
  float* hid[2] = {reinterpret_cast<float*>(d_scratch),
                   reinterpret_cast<float*>(d_scratch) + (int64_t)n_prob * U * (H > 0 ? H : 1)};
  // hidden layers: one launch per depth level for all MLPs of all generators
  for (int i = 0; i < depth; ++i) {
    LinearGroup grp{};
    for (int p = 0; p < n_prob; ++p) {
      grp.X[p] = i == 0 ? d_text_emb : hid[(i - 1) & 1] + (int64_t)p * U * H;
      grp.W[p] = param(p / 3, p % 3, i, 0);
      grp.b[p] = param(p / 3, p % 3, i, 1);
      grp.log_scale[p] = nullptr;
      grp.Y[p] = hid[i & 1] + (int64_t)p * U * H;
    }
    // many relations (zero-shot vocabularies, BASELINE config 4: U = 20k): a hidden Linear is large enough to fill
    // the machine by itself and goes to the tcgen05 Linear (3xTF32) problem by problem; otherwise one grouped launch
    const int in_i = i == 0 ? T : H;
    if (linear_umma_eligible(U, in_i, H, 1, nullptr, grp.X[0], grp.W[0], grp.Y[0])) {
      for (int p = 0; p < n_prob; ++p)
        if (int rc = ghf_linear(grp.X[p], U, in_i, grp.W[p], grp.b[p], H, 1, nullptr, grp.Y[p], stream_)) return rc;
    } else if (int rc = linear_group(grp, n_prob, U, in_i, H, 1, stream)) {
      return rc;
    }
  }
  const int in_dim = depth > 0 ? H : T;
  auto last_in = [&](int p) { return depth > 0 ? hid[(depth - 1) & 1] + (int64_t)p * U * H : d_text_emb; };
  // last Linears: the bias generators of all layers in one launch; the two [U, d_in * d_out] ones per generator go
  // to ghf_linear (tcgen05 3xTF32 when they are large enough) unless the caller writes operand images itself
  {
    LinearGroup grp{};
    for (int gdx = 0; gdx < n_gen; ++gdx) {
      const int p = 3 * gdx + 2;
      grp.X[gdx] = last_in(p);
      grp.W[gdx] = param(gdx, 2, depth, 0);
      grp.b[gdx] = param(gdx, 2, depth, 1);
      grp.log_scale[gdx] = h_log_scales[3 * gdx + 2];
      grp.Y[gdx] = h_out[3 * gdx + 2];
    }
    if (int rc = linear_group(grp, n_gen, U, in_dim, d_out, 0, stream)) return rc;
  }
  if (skip_big) return 0;
  for (int gdx = 0; gdx < n_gen; ++gdx)
    for (int m = 0; m < 2; ++m)
      if (int rc = ghf_linear(last_in(3 * gdx + m), U, in_dim, param(gdx, m, depth, 0), param(gdx, m, depth, 1),
                              d_in * d_out, 0, h_log_scales[3 * gdx + m], h_out[3 * gdx + m], stream_))
        return rc;
  return 0;
}
