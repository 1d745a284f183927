// red_bw.cu — microbenchmark: SM-side throughput of scattering 512-byte fp32 rows with atomic adds into an
// L2-resident window (25 MB), three ways:  red.global.add.v4.f32 per lane | red.global.add.f32 per lane |
// cp.reduce.async.bulk (one 512 B bulk reduction per row from shared memory) | plain st.v4 (no atomics).
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void scatter_kernel(float* __restrict__ out, const int* __restrict__ idx, long n_rows) {
  __shared__ __align__(128) float rows[16][128];   // one 512 B row per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = lane; j < 128; j += 32) rows[warp][j] = 1.0f;
  __syncwarp();
  const long gw = (long)blockIdx.x * nwarps + warp, total = (long)gridDim.x * nwarps;
  const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
  for (long r = gw; r < n_rows; r += total) {
    float* dst = out + (long)idx[r] * 128;
    if (MODE == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + lane * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + j * 32 + lane), "f"(v.x) : "memory");
    } else if (MODE == 2) {
      if (lane == 0) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(dst), "r"(smem_u32(rows[warp])) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
      }
    } else {
      *reinterpret_cast<float4*>(dst + lane * 4) = v;
    }
  }
  if (MODE == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const long WIN = 49152, E = 16000000;   // rows in the window (25 MB), rows scattered per launch
  float* out; int* idx;
  cudaMalloc(&out, WIN * 512);
  cudaMalloc(&idx, E * 4);
  cudaMemset(out, 0, WIN * 512);
  std::vector<int> hi(E);
  unsigned long long s = 88172645463325252ull;
  for (long i = 0; i < E; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; hi[i] = (int)(s % WIN); }
  cudaMemcpy(idx, hi.data(), E * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"red.v4.f32", "red.f32 x4", "cp.reduce.async.bulk 512B", "st.v4 (no atomics)"};
  printf("%-28s warps/SM     ms   Mrows/s  GB/s   cycles/row/SM@1.9GHz\n", "mode");
  auto run = [&](int mode, int warps) {
    auto go = [&](auto kern) {
      for (int i = 0; i < 2; ++i) kern<<<148, warps * 32>>>(out, idx, E);
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) kern<<<148, warps * 32>>>(out, idx, E);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
      printf("%-28s %8d %6.3f %9.0f %5.0f %10.1f  %s\n", names[mode], warps, ms, E / ms / 1e3, E * 512.0 / ms / 1e6,
             ms * 1e-3 * 1.9e9 / (E / 148.0), cudaGetErrorString(cudaGetLastError()));
    };
    if (mode == 0) go(scatter_kernel<0>); else if (mode == 1) go(scatter_kernel<1>);
    else if (mode == 2) go(scatter_kernel<2>); else go(scatter_kernel<3>);
  };
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 8, 16}) run(mode, warps);
  return 0;
}
