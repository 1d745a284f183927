// scatter_paths.cu — round-2 microbenchmark: which path carries 16M partial result rows from the SMs to their
// destination rows fastest on B200?  Round 1 (tools/l2_paths.cu) measured 24 cycles per 512 B row per SM for
// warp-coalesced red.global.add.f32 into an L2-resident window.  Candidates measured here, same 16M rows:
//   red32     red.global.add.f32, one 128 B line per warp instruction (the round-1 shape)
//   redv2/v4  red.global.add.v2/v4.f32: 256 / 512 B per warp instruction
//   tma       cp.reduce.async.bulk.global.shared::cta.add.f32: one 512 B row per bulk operation, from shared memory
//   st512     plain st.global.v4 of the 512 B row (no reduction: the two-phase alternative's store side)
//   st256     plain 256 B rows (fp16 partial results)
//   redh2     red.global.add.noftz.f16x2, 256 B rows
//   dsmem     red.shared::cluster.add.f32 into the shared memory of a random CTA of the cluster (8 or 16 CTAs)
//   dsmemst   st.shared::cluster.v4.f32 (no reduction), same addressing
// Each can run next to random 256 B row gathers (cp.async into shared memory), as in the contraction kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/scatter_paths.cu -o tools/bin/scatter_paths
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

enum Path { RED32 = 0, REDV2, REDV4, TMA, ST512, ST256, REDH2, DSMEM, DSMEMST, NONE };

__device__ __forceinline__ void cp16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int kGatherStages = 4, kRowsPerGroup = 8;
constexpr int kTmaRows = 64;          // staging rows per CTA for the bulk reductions
constexpr int kDsmemRows = 256;       // accumulator rows per CTA for the cluster variants (128 KiB)

// warps [0, gw): gather 256 B rows; warps [gw, gw + sw): scatter.
template <int PATH>
__global__ void scatter_kernel(const unsigned char* __restrict__ h, const int* __restrict__ src_idx, long n_gather,
                               int gw, float* __restrict__ acc, const int* __restrict__ dst_idx, long n_scatter,
                               int cluster) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int sw = nw - gw;
  unsigned char* gather_smem = smem;
  unsigned char* stage_smem = smem + gw * kGatherStages * kRowsPerGroup * 256;   // TMA staging / DSMEM accumulators
  if (PATH == DSMEM || PATH == DSMEMST) {
    for (int i = threadIdx.x; i < kDsmemRows * 128; i += blockDim.x) reinterpret_cast<float*>(stage_smem)[i] = 0.f;
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  } else if (PATH == TMA) {
    for (int i = threadIdx.x; i < kTmaRows * 128; i += blockDim.x) reinterpret_cast<float*>(stage_smem)[i] = 1.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (warp < gw) {
    const long warps_total = (long)gridDim.x * gw;
    const long me = (long)blockIdx.x * gw + warp;
    const unsigned base = smem_u32(gather_smem) + warp * kGatherStages * kRowsPerGroup * 256;
    const long groups = n_gather / (warps_total * kRowsPerGroup);
    int my_idx = 0;
    for (long g = 0; g < groups + kGatherStages - 1; ++g) {
      if (g < groups) {
        const int st = g % kGatherStages;
        if ((g & 3) == 0) my_idx = src_idx[((g >> 2) * warps_total + me) * 32 + lane];
        for (int r = 0; r < kRowsPerGroup; r += 2) {
          const int rr = r + (lane >> 4);
          const long src = __shfl_sync(0xffffffffu, my_idx, (int)((g & 3) * 8 + rr));
          cp16(base + (st * kRowsPerGroup + rr) * 256 + (lane & 15) * 16, h + src * 256 + (lane & 15) * 16);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kGatherStages - 1) : "memory");
    }
  } else if (PATH != NONE) {
    const long warps_total = (long)gridDim.x * sw;
    const long me = (long)blockIdx.x * sw + (warp - gw);
    const long blocks = n_scatter / (warps_total * 32);
    if (PATH == TMA) {
      // every lane owns one staging row and issues the bulk reductions of "its" destination: 32 rows per warp step,
      // up to 4 bulk groups in flight per thread
      const unsigned srow = smem_u32(stage_smem) + (((warp - gw) * 32 + lane) % kTmaRows) * 512;
      for (long b = 0; b < blocks; ++b) {
        const long d = dst_idx[(b * warps_total + me) * 32 + lane];
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(acc + d * 128),
                     "r"(srow)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      return;
    }
    unsigned my_rank = 0;
    if (PATH == DSMEM || PATH == DSMEMST) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(my_rank));
    for (long b = 0; b < blocks; ++b) {
      const int my_dst = dst_idx[(b * warps_total + me) * 32 + lane];
#pragma unroll 4
      for (int e = 0; e < 32; ++e) {
        const long d = __shfl_sync(0xffffffffu, my_dst, e);
        if (PATH == RED32) {
          float* p = acc + d * 128 + lane;
#pragma unroll
          for (int q = 0; q < 4; ++q) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p + 32 * q), "f"(1.0f) : "memory");
        } else if (PATH == REDV2) {
          float* p = acc + d * 128 + 2 * lane;
#pragma unroll
          for (int q = 0; q < 2; ++q)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %1};" ::"l"(p + 64 * q), "f"(1.0f) : "memory");
        } else if (PATH == REDV4) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(acc + d * 128 + 4 * lane), "f"(1.0f)
                       : "memory");
        } else if (PATH == ST512) {
          asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(acc + d * 128 + 4 * lane), "f"((float)e)
                       : "memory");
        } else if (PATH == ST256) {
          if (lane < 16)
            asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(acc + d * 64 + 4 * lane), "f"((float)e)
                         : "memory");
        } else if (PATH == REDH2) {
          unsigned* p = reinterpret_cast<unsigned*>(acc) + d * 64 + lane;
#pragma unroll
          for (int q = 0; q < 2; ++q)
            asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(p + 32 * q), "r"(0x3c003c00u) : "memory");
        } else if (PATH == DSMEM || PATH == DSMEMST) {
          // destination d -> CTA (d % cluster) of this cluster, row (d / cluster) % kDsmemRows
          const unsigned cta = (unsigned)(d % cluster);
          const unsigned local = smem_u32(stage_smem) + (unsigned)((d / cluster) % kDsmemRows) * 512u;
          unsigned remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta));
          if (PATH == DSMEM) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              asm volatile("red.shared::cluster.add.f32 [%0], %1;" ::"r"(remote + (unsigned)(lane + 32 * q) * 4u), "f"(1.0f)
                           : "memory");
          } else {
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(remote + (unsigned)lane * 16u),
                         "f"((float)e)
                         : "memory");
          }
        }
      }
    }
  }
  if (PATH == DSMEM || PATH == DSMEMST)
    asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}

int main() {
  const long N = 2500000, E = 16000000;
  unsigned char* h; int *src, *dst, *dst_big; float* acc;
  cudaMalloc(&h, N * 256);
  cudaMalloc(&src, E * 4);
  cudaMalloc(&dst, E * 4);
  cudaMalloc(&dst_big, E * 4);
  cudaMalloc(&acc, N * 512);
  cudaMemset(h, 0, N * 256);
  cudaMemset(acc, 0, N * 512);
  std::vector<int> a(E), b(E), c(E);
  unsigned long long s = 88172645463325252ull;
  const long WIN = 49152;
  for (long i = 0; i < E; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    a[i] = (int)(s % N); b[i] = (int)((s >> 32) % WIN); c[i] = (int)((s >> 20) % N);
  }
  cudaMemcpy(src, a.data(), E * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dst, b.data(), E * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dst_big, c.data(), E * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%-44s ctas gw sw   ms    rows/us  cyc/row/SM@1.9  status\n", "case");
  auto run = [&](const char* name, auto kernel, int path, int ctas, int gw, int sw, int cluster, bool big) {
    int smem = gw * kGatherStages * kRowsPerGroup * 256;
    if (path == TMA) smem += kTmaRows * 512;
    if (path == DSMEM || path == DSMEMST) smem += kDsmemRows * 512;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cluster > 8) cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3((gw + sw) * 32); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster > 1 ? cluster : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const long ng = gw ? E : 0;
    auto go = [&]() { cudaLaunchKernelEx(&cfg, kernel, (const unsigned char*)h, (const int*)src, ng, gw, acc,
                                         (const int*)(big ? dst_big : dst), E, cluster > 1 ? cluster : 1); };
    go(); go();
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-44s FAILED: %s\n", name, cudaGetErrorString(err)); cudaGetLastError(); return; }
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) go();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("%-44s %4d %2d %2d %6.3f %9.0f %10.1f  %s\n", name, ctas, gw, sw, ms, E / (ms * 1e3),
           ms * 1e-3 * 1.9e9 / (E / (double)ctas), cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
  };
#define RUN(name, P, ...) run(name, scatter_kernel<P>, P, __VA_ARGS__)
  RUN("red32  window 25 MB", RED32, 148, 0, 16, 1, false);
  RUN("redv2  window 25 MB", REDV2, 148, 0, 16, 1, false);
  RUN("redv4  window 25 MB", REDV4, 148, 0, 16, 1, false);
  RUN("redv4  window 25 MB, 8 warps", REDV4, 148, 0, 8, 1, false);
  RUN("redv4  window 25 MB, 4 warps", REDV4, 148, 0, 4, 1, false);
  RUN("redv4  window 25 MB, half the SMs", REDV4, 74, 0, 16, 1, false);
  RUN("tma    window 25 MB, 4 warps", TMA, 148, 0, 4, 1, false);
  RUN("tma    window 25 MB, 8 warps", TMA, 148, 0, 8, 1, false);
  RUN("tma    window 25 MB, half the SMs", TMA, 74, 0, 8, 1, false);
  RUN("st512  window 25 MB", ST512, 148, 0, 16, 1, false);
  RUN("st512  window 25 MB, 8 warps", ST512, 148, 0, 8, 1, false);
  RUN("st256  window 12.5 MB", ST256, 148, 0, 16, 1, false);
  RUN("redh2  window 12.5 MB", REDH2, 148, 0, 16, 1, false);
  RUN("red32  all 1.28 GB (HBM)", RED32, 148, 0, 16, 1, true);
  RUN("redv4  all 1.28 GB (HBM)", REDV4, 148, 0, 16, 1, true);
  RUN("tma    all 1.28 GB (HBM)", TMA, 148, 0, 8, 1, true);
  RUN("st512  all 1.28 GB (HBM)", ST512, 148, 0, 16, 1, true);
  RUN("st256  all 0.64 GB (HBM)", ST256, 148, 0, 16, 1, true);
  RUN("dsmem red, cluster 8", DSMEM, 144, 0, 16, 8, false);
  RUN("dsmem red, cluster 16", DSMEM, 128, 0, 16, 16, false);
  RUN("dsmem st,  cluster 8", DSMEMST, 144, 0, 16, 8, false);
  RUN("dsmem st,  cluster 16", DSMEMST, 128, 0, 16, 16, false);
  RUN("gather 256 alone", NONE, 148, 16, 0, 1, false);
  RUN("gather 256 + red32", RED32, 148, 16, 16, 1, false);
  RUN("gather 256 + redv4", REDV4, 148, 16, 16, 1, false);
  RUN("gather 256 + redv4, 8 warps", REDV4, 148, 16, 8, 1, false);
  RUN("gather 256 + tma, 8 warps", TMA, 148, 16, 8, 1, false);
  RUN("gather 256 + st512", ST512, 148, 16, 16, 1, false);
  RUN("gather 256 + st256", ST256, 148, 16, 16, 1, false);
  RUN("gather 256 + st256 (HBM)", ST256, 148, 16, 16, 1, true);
  RUN("gather 256 + dsmem red, cluster 16", DSMEM, 128, 16, 16, 16, false);
  return 0;
}
