// l2_paths.cu — microbenchmark behind the message-passing kernel's design (B200):
//   (1) is the scatter of 512-byte fp32 rows into an L2-resident window limited per SM or chip-wide?
//       -> same kernel on 148 and on 74 CTAs, 16/32 warps, destination ids prefetched 32 at a time
//   (2) do the random row gather (HBM -> shared memory, cp.async) and the scatter (SM -> L2 red) share a
//       limit?  -> one kernel where half of the warps gather and half scatter, against each alone
//   (3) gather of 256-byte rows (fp16 features) against 512-byte rows at equal bytes in flight
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_paths.cu -o tools/bin/l2_paths
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// MODE bit 0: gather warps active, bit 1: scatter warps active.  Warps [0, gw) gather, [gw, nw) scatter.
// Gather: each warp keeps STAGES groups of 8 rows in flight.  Scatter: 32 destination ids per lane-load,
// one coalesced red.f32 per 128 B quarter-row (the shape the kernel's epilogue emits).
template <int STAGES>
__global__ void paths_kernel(const unsigned char* __restrict__ h, const int* __restrict__ src_idx, long n_gather,
                             int row_bytes, int gw, float* __restrict__ acc, const int* __restrict__ dst_idx,
                             long n_scatter, int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (warp < gw) {
    if (!(mode & 1)) return;
    const long warps_total = (long)gridDim.x * gw;
    const long me = (long)blockIdx.x * gw + warp;
    const int lanes_per_row = row_bytes / 16;            // 32 (512 B) or 16 (256 B)
    const int rows_per_op = 32 / lanes_per_row;          // 1 or 2
    const int rpg = 8;
    unsigned base = (unsigned)__cvta_generic_to_shared(smem) + warp * STAGES * rpg * row_bytes;
    const long groups = n_gather / (warps_total * rpg);
    int my_idx = 0;
    for (long g = 0; g < groups + STAGES - 1; ++g) {
      if (g < groups) {
        const int st = g % STAGES;
        if ((g & 3) == 0) my_idx = src_idx[((g >> 2) * warps_total + me) * 32 + lane];   // 32 ids = 4 groups
        for (int r = 0; r < rpg; r += rows_per_op) {
          const int rr = r + (lane / lanes_per_row);
          const long src = __shfl_sync(0xffffffffu, my_idx, (int)((g & 3) * 8 + rr));
          cp16(base + (st * rpg + rr) * row_bytes + (lane % lanes_per_row) * 16,
               h + src * row_bytes + (lane % lanes_per_row) * 16);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    }
  } else {
    if (!(mode & 2)) return;
    const int sw = nw - gw;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (mode & 4) {
      // the kernel's shape: four warps share a row block, warp q writes quarter q of each of its 32 rows
      const int q = (warp - gw) & 3, team = (warp - gw) >> 2, teams = sw >> 2;
      const long teams_total = (long)gridDim.x * teams;
      const long me = (long)blockIdx.x * teams + team;
      const long blocks = n_scatter / (teams_total * 32);
      for (long b = 0; b < blocks; ++b) {
        const int my_dst = dst_idx[(b * teams_total + me) * 32 + lane];
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
          const long d = __shfl_sync(0xffffffffu, my_dst, e);
          float* p = acc + d * 128 + 32 * q + lane;
          if (mode & 8) asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(1.0f), "l"(pol) : "memory");
          else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(1.0f) : "memory");
        }
      }
      return;
    }
    const long warps_total = (long)gridDim.x * sw;
    const long me = (long)blockIdx.x * sw + (warp - gw);
    const long blocks = n_scatter / (warps_total * 32);
    for (long b = 0; b < blocks; ++b) {
      const int my_dst = dst_idx[(b * warps_total + me) * 32 + lane];
#pragma unroll 4
      for (int e = 0; e < 32; ++e) {
        const long d = __shfl_sync(0xffffffffu, my_dst, e);
        float* p = acc + d * 128 + lane;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (mode & 8) asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p + 32 * q), "f"(1.0f), "l"(pol) : "memory");
          else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 32 * q), "f"(1.0f) : "memory");
        }
      }
    }
  }
}

int main() {
  const long N = 2500000, E = 16000000, WIN = 49152;
  unsigned char* h; int *src, *dst; float* acc;
  cudaMalloc(&h, N * 512);
  cudaMalloc(&src, E * 4);
  cudaMalloc(&dst, E * 4);
  cudaMalloc(&acc, N * 512);          // 1.28 GB: the moving-window cases walk all of it
  cudaMemset(h, 0, N * 512);
  cudaMemset(acc, 0, N * 512);
  int* dst_mv; cudaMalloc(&dst_mv, E * 4);
  std::vector<int> a(E), b(E);
  unsigned long long s = 88172645463325252ull;
  for (long i = 0; i < E; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; a[i] = (int)(s % N); b[i] = (int)((s >> 32) % WIN); }
  cudaMemcpy(src, a.data(), E * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dst, b.data(), E * 4, cudaMemcpyHostToDevice);
  {  // moving window: edge i scatters into super-block i / (E / 51), a random row of its 49152
    const long per = E / 51 + 1;
    for (long i = 0; i < E; ++i) { long r = (i / per) * WIN + b[i]; b[i] = (int)(r < N ? r : r - WIN); }
    cudaMemcpy(dst_mv, b.data(), E * 4, cudaMemcpyHostToDevice);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%-34s ctas gw sw row_B  ms   gather GB/s  scatter GB/s  cyc/row/SM@1.9\n", "case");
  auto run = [&](const char* name, int ctas, int gw, int sw, int row_bytes, int mode) {
    const int* dsts = (mode & 16) ? dst_mv : dst;
    constexpr int ST = 4;
    const int smem = gw * ST * 8 * row_bytes;
    cudaFuncSetAttribute(paths_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    auto go = [&]() { paths_kernel<ST><<<ctas, (gw + sw) * 32, smem>>>(h, src, E, row_bytes, gw, acc, dsts, E, mode); };
    go(); go();
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) go();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("%-34s %4d %2d %2d %4d %6.3f %10.0f %12.0f %10.1f  %s\n", name, ctas, gw, sw, row_bytes, ms,
           (mode & 1) ? E * (double)row_bytes / ms / 1e6 : 0.0, (mode & 2) ? E * 512.0 / ms / 1e6 : 0.0,
           ms * 1e-3 * 1.9e9 / (E / (double)ctas), cudaGetErrorString(cudaGetLastError()));
  };
  run("scatter, 4 warps share a row", 148, 0, 8, 512, 2 | 4);
  run("scatter, 4 warps share a row", 148, 0, 16, 512, 2 | 4);
  run("scatter, evict_last hint", 148, 0, 8, 512, 2 | 8);
  run("scatter, shared row + hint", 148, 0, 8, 512, 2 | 4 | 8);
  run("scatter, moving window 1.28 GB", 148, 0, 8, 512, 2 | 16);
  run("scatter, moving+shared+hint", 148, 0, 8, 512, 2 | 4 | 8 | 16);
  run("scatter only", 148, 0, 8, 512, 2);
  run("scatter only", 148, 0, 16, 512, 2);
  run("scatter only", 148, 0, 32, 512, 2);
  run("scatter only, half the SMs", 74, 0, 16, 512, 2);
  run("scatter only, half the SMs", 74, 0, 32, 512, 2);
  run("gather only 512 B rows", 148, 12, 0, 512, 1);
  run("gather only 512 B rows", 148, 8, 0, 512, 1);
  run("gather only 256 B rows", 148, 12, 0, 256, 1);
  run("gather only 256 B rows", 148, 16, 0, 256, 1);
  run("gather only 256 B rows", 148, 24, 0, 256, 1);
  run("gather 512 + scatter together", 148, 12, 8, 512, 3);
  run("gather 512 + scatter together", 148, 12, 16, 512, 3);
  run("gather 256 + scatter together", 148, 16, 8, 256, 3);
  run("gather 256 + scatter together", 148, 16, 16, 256, 3);
  return 0;
}
