// ffma_gemm.cuh — fp32 CUDA-core tile engine shared by ghf_linear (dense.cu) and the exact-fp32
// message-passing path (mp_fp32.cu).  128 x BN x 16 tiles, 256 threads, 8 x (BN/16) outputs per
// thread, register-prefetch double buffering.  This is the exact (rtol 1e-5) path; the tensor-core
// path lives in mp_umma.cu.
#pragma once

#include "common.cuh"

namespace ghf {

constexpr int kFfmaBM = 128;
constexpr int kFfmaBK = 16;
constexpr int kFfmaThreads = 256;

template <int BN>
struct FfmaSmem {
  float A[2][kFfmaBK][kFfmaBM + 4];
  float B[2][kFfmaBK][BN + 4];
};

// acc[i][j]: rows {ty*4+i (i<4), 64+ty*4+(i-4)}, cols: TN==8 -> {tx*4+j (j<4), 64+tx*4+(j-4)},
// TN==4 -> tx*4+j, TN==2 -> tx*2+j.
template <int BN>
__device__ __forceinline__ int ffma_col(int tx, int j) {
  constexpr int TN = BN / 16;
  if constexpr (TN == 8) return j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4);
  return tx * TN + j;
}
__device__ __forceinline__ int ffma_row(int ty, int i) { return i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4); }

template <int BN>
__device__ __forceinline__ void ffma_compute(const FfmaSmem<BN>& sm, int buf, int tx, int ty,
                                             float (&acc)[8][BN / 16]) {
  constexpr int TN = BN / 16;
#pragma unroll
  for (int k = 0; k < kFfmaBK; ++k) {
    float a[8], b[TN];
    const float4 a0 = *reinterpret_cast<const float4*>(&sm.A[buf][k][ty * 4]);
    const float4 a1 = *reinterpret_cast<const float4*>(&sm.A[buf][k][64 + ty * 4]);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    if constexpr (TN == 8) {
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.B[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sm.B[buf][k][64 + tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
      b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    } else if constexpr (TN == 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(&sm.B[buf][k][tx * 4]);
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
    } else {
      const float2 b0 = *reinterpret_cast<const float2*>(&sm.B[buf][k][tx * 2]);
      b[0] = b0.x; b[1] = b0.y;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// Generic mainloop.  LoadA(row, k) / LoadB(k, n) return one element (0 outside bounds); the
// 4-wide variants return 4 consecutive-k (A, and B when B is K-contiguous) or consecutive-n (B when
// N-contiguous) elements and are used when the caller guarantees 4-alignment.
template <int BN, bool VEC, bool B_KMAJOR, class FA, class FA4, class FB, class FB4>
__device__ __forceinline__ void ffma_mainloop(FfmaSmem<BN>& sm, int Ktot, FA loadA, FA4 loadA4, FB loadB,
                                              FB4 loadB4, float (&acc)[8][BN / 16]) {
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int A4 = kFfmaBM * kFfmaBK / 4 / kFfmaThreads;          // float4 per thread for A (=2)
  constexpr int B4 = (BN * kFfmaBK / 4 + kFfmaThreads - 1) / kFfmaThreads;
  constexpr int AS = kFfmaBM * kFfmaBK / kFfmaThreads;              // scalars per thread (=8)
  constexpr int BS = (BN * kFfmaBK + kFfmaThreads - 1) / kFfmaThreads;
  float4 ra4[A4], rb4[B4];
  float ras[VEC ? 1 : AS], rbs[VEC ? 1 : BS];

  auto fetch = [&](int k0) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < A4; ++i) {
        const int f = tid + i * kFfmaThreads;
        ra4[i] = loadA4(f >> 2, k0 + (f & 3) * 4);
      }
#pragma unroll
      for (int i = 0; i < B4; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK / 4) {
          if constexpr (B_KMAJOR) rb4[i] = loadB4(k0 + (f & 3) * 4, f >> 2);           // (k, n): 4 along k
          else          rb4[i] = loadB4(k0 + f / (BN / 4), (f % (BN / 4)) * 4);  // 4 along n
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < AS; ++i) {
        const int f = tid + i * kFfmaThreads;
        ras[i] = loadA(f / kFfmaBK, k0 + f % kFfmaBK);
      }
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK) {
          if constexpr (B_KMAJOR) rbs[i] = loadB(k0 + f % kFfmaBK, f / kFfmaBK);
          else          rbs[i] = loadB(k0 + f / BN, f % BN);
        }
      }
    }
  };
  auto stash = [&](int buf) {
    if constexpr (VEC) {
#pragma unroll
      for (int i = 0; i < A4; ++i) {
        const int f = tid + i * kFfmaThreads;
        const int row = f >> 2, kq = (f & 3) * 4;
        sm.A[buf][kq + 0][row] = ra4[i].x; sm.A[buf][kq + 1][row] = ra4[i].y;
        sm.A[buf][kq + 2][row] = ra4[i].z; sm.A[buf][kq + 3][row] = ra4[i].w;
      }
#pragma unroll
      for (int i = 0; i < B4; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK / 4) {
          if constexpr (B_KMAJOR) {
            const int n = f >> 2, kq = (f & 3) * 4;
            sm.B[buf][kq + 0][n] = rb4[i].x; sm.B[buf][kq + 1][n] = rb4[i].y;
            sm.B[buf][kq + 2][n] = rb4[i].z; sm.B[buf][kq + 3][n] = rb4[i].w;
          } else {
            *reinterpret_cast<float4*>(&sm.B[buf][f / (BN / 4)][(f % (BN / 4)) * 4]) = rb4[i];
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < AS; ++i) {
        const int f = tid + i * kFfmaThreads;
        sm.A[buf][f % kFfmaBK][f / kFfmaBK] = ras[i];
      }
#pragma unroll
      for (int i = 0; i < BS; ++i) {
        const int f = tid + i * kFfmaThreads;
        if (f < BN * kFfmaBK) {
          if constexpr (B_KMAJOR) sm.B[buf][f % kFfmaBK][f / kFfmaBK] = rbs[i];
          else          sm.B[buf][f / BN][f % BN] = rbs[i];
        }
      }
    }
  };

  const int nk = (Ktot + kFfmaBK - 1) / kFfmaBK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * kFfmaBK);
    ffma_compute<BN>(sm, buf, tx, ty, acc);
    if (kt + 1 < nk) stash(buf ^ 1);
    __syncthreads();
  }
}

}  // namespace ghf
