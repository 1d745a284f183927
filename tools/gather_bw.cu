// gather_bw.cu — microbenchmark: achievable HBM bandwidth for random 512-byte row gathers on B200 as a
// function of bytes in flight per SM (cp.async into shared memory, the way the contraction kernel gathers).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_bw.cu -o gpurun_out/gather_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// each warp: loop over its rows; STAGES groups of `rows_per_group` rows in flight (cp.async commit groups)
template <int STAGES>
__global__ void gather_kernel(const float* __restrict__ h, const int* __restrict__ idx, long n_rows, int D,
                              int rows_per_group) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long warps_total = (long)gridDim.x * nwarps;
  const long gw = (long)blockIdx.x * nwarps + warp;
  const int row_bytes = D * 4;                       // 512
  unsigned base = (unsigned)__cvta_generic_to_shared(smem) + warp * STAGES * rows_per_group * row_bytes;
  long g = 0;
  const long groups = n_rows / (warps_total * rows_per_group);
  for (; g < groups + STAGES - 1; ++g) {
    if (g < groups) {
      const int st = g % STAGES;
      for (int r = 0; r < rows_per_group; ++r) {
        const long row = (g * warps_total + gw) * rows_per_group + r;
        const long src = idx[row];
        cp16(base + (st * rows_per_group + r) * row_bytes + lane * 16, h + src * D + lane * 4);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
  }
}

int main() {
  const long N = 2500000, E = 16000000;
  const int D = 128;
  float* h; int* idx;
  cudaMalloc(&h, N * D * 4);
  cudaMalloc(&idx, E * 4);
  cudaMemset(h, 0, N * D * 4);
  std::vector<int> hi(E);
  unsigned long long s = 88172645463325252ull;
  for (long i = 0; i < E; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; hi[i] = (int)(s % N); }
  cudaMemcpy(idx, hi.data(), E * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("blocks/SM warps rows/grp stages  KB-in-flight/SM   GB/s\n");
  auto run = [&](int bps, int warps, int rpg, int stages) {
    const int smem = warps * stages * rpg * 512;
    auto launch = [&](auto kern) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int i = 0; i < 2; ++i) kern<<<148 * bps, warps * 32, smem>>>(h, idx, E, D, rpg);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) kern<<<148 * bps, warps * 32, smem>>>(h, idx, E, D, rpg);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      cudaError_t err = cudaGetLastError();
      printf("%9d %5d %8d %6d %16.0f %7.0f %s\n", bps, warps, rpg, stages, bps * smem / 1024.0,
             5.0 * E * 512 / (ms * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
    };
    if (stages == 2) launch(gather_kernel<2>);
    else if (stages == 4) launch(gather_kernel<4>);
    else launch(gather_kernel<8>);
  };
  run(1, 4, 8, 2);    // 32 KB
  run(1, 4, 8, 4);    // 64 KB
  run(1, 5, 8, 4);    // 80 KB  (what the contraction kernel has today)
  run(1, 4, 8, 8);    // 128 KB
  run(1, 8, 8, 4);    // 128 KB, more warps
  run(1, 12, 8, 4);   // 192 KB
  run(1, 13, 8, 4);   // 208 KB
  run(2, 6, 8, 4);    // 2 x 96 KB
  run(4, 4, 8, 2);    // 4 x 32 KB
  run(1, 13, 4, 8);   // 208 KB, smaller groups
  run(1, 16, 2, 4);   // 64 KB, many warps
  return 0;
}
