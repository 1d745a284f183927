// linear_umma.cu — Y = alpha * act(X W^T + b) at K = 128 and N a multiple of 128 on tcgen05: the node-feature
// projection (HG:261, N = 128) and the last Linear of the weight generators (WG:138-140, N = d*d, alpha =
// exp(log_scale)); blockIdx.y selects the 128-wide block of output features.  fp32-grade through a 3 x TF32 split:  x = xh + xl, w = wh + wl  (h = top 19 bits, l = the remainder),
//     x.w ~ xh.wh + xl.wh + xh.wl            (the dropped xl.wl term is ~2^-22 relative)
//
// Layout mirrors mp_umma_ts.cu.  The transposed tile  Yt[128 n, 128 rows] = W[128 n, 128 k] * Xt  is computed
// with A = W from TENSOR MEMORY (wh in columns [256,384), wl in [384,512), loaded once per CTA: nn.Linear's
// weight [out, in] is already "lane = n, column = k") and B = X tiles staged in shared memory by the producers,
// which split every value on the fly (hi tile + lo tile per 32-wide K chunk, 128 B swizzle, K-major).
// Epilogue: TMEM -> registers -> transpose through shared memory -> + bias, ReLU -> coalesced 512 B row stores.
//
// Warp roles (544 threads, 1 CTA / SM, persistent over row tiles):
//   0-7 epilogue (2 groups; warps 0-3 also load W into TMEM at start) | 8-15 producers (2 groups, alternating
//   chunks) | 16 MMA issuer + TMEM allocator
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "ghf_b200.h"
#include "umma.cuh"

namespace ghf {
int mp_f16_absmax(const float* x, int64_t elems, float* scale, cudaStream_t stream);   // mp_f16.cu
namespace {

using namespace ptx;

constexpr int kD = 128;                     // K and N
constexpr int kTile = 128;                  // rows per tile
constexpr int kChunks = kD / 32;            // 4 K-chunks of 32
constexpr int kSub = kTile * 128;           // 16 KiB: one chunk of one tile (hi or lo)
constexpr int kStage = 2 * kSub;            // hi + lo
constexpr int kStages = 6;
constexpr int kStaging = 32 * kD * 4;       // per epilogue group
constexpr int kSmem = 1024 + kStages * kStage + 2 * kStaging + 512;
constexpr int kWarpProd = 8, kWarpMma = 16;
constexpr int kThreads = 32 * (kWarpMma + 1);
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kWhCol = 256, kWlCol = 384;

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// kImage: instead of fp32 rows, the epilogue writes the result straight into the per-relation fp16 OPERAND IMAGES of
// the hidden-64/256 contraction (mp_f16_ss.cu): row m = relation, output feature j = k * d + n = entry (k, n) of the
// generated matrix, scaled by the relation's power of two `row_scale[m]` - the fp32 weights never exist (SURVEY 8f
// rank 1).  Image of relation m: K-chunks of 64 (k + which * d counts through [W_msg; W_self]), inside a chunk
// 64-wide n blocks of [64 k][128 B], 16-byte groups swizzled by k % 8 - an MN-major SWIZZLE_128B operand.
struct ImageOut {
  uint8_t* img;             // [M][image_bytes]
  const float* row_scale;   // [M]
  int64_t image_bytes;
  int d, which;             // hidden_dim; 0 = W_msg half, 1 = W_self half
};

template <bool kImage>
__global__ void __launch_bounds__(kThreads, 1)
linear128_umma_kernel(const float* __restrict__ X, int64_t M, const float* __restrict__ W,
                      const float* __restrict__ bias, int relu, const float* __restrict__ log_scale,
                      float* __restrict__ Y, int64_t ldy, __half* __restrict__ Y16,
                      float* __restrict__ y16_scale, ImageOut io) {
  // this CTA's block of 128 output features: rows [128 y, 128 y + 128) of W, the same columns of Y
  W += (int64_t)blockIdx.y * kD * kD;
  if (!kImage) Y += (int64_t)blockIdx.y * kD;
  if (Y16) Y16 += (int64_t)blockIdx.y * kD;   // optional fp16 copy of Y (same leading dimension)
  if (bias) bias += (int64_t)blockIdx.y * kD;
  const float alpha = log_scale ? expf(*log_scale) : 1.f;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sStg = sA + kStages * kStage;
  const uint32_t sBar = sStg + 2 * kStaging;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return sBar + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = sBar + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (M + kTile - 1) / kTile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 128);   // the 128 threads of the producer group that owns the chunk
      mbar_init(empty(s), 1);    // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 256);
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // W -> TMEM once: thread n of warps 0-3 owns row n of W (lane n): 128 k values as hi and lo columns
  if (warp < 4) {
    const float* wrow = W + (int64_t)(warp * 32 + lane) * kD;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < kChunks; ++c) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(wrow + c * 32 + 4 * j);
        const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float h = to_tf32_rna(f[t]);
          hi[4 * j + t] = __float_as_uint(h);
          lo[4 * j + t] = __float_as_uint(to_tf32_rna(f[t] - h));
        }
      }
      tmem_st_32x32(t_row + kWhCol + 32u * c, hi);
      tmem_st_32x32(t_row + kWlCol + 32u * c, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < kWarpProd) {
    // ------------------------------------------------------------------ epilogue
    const int grp = warp >> 2, q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem_raw + (sStg - raw) + grp * kStaging);
    const float4* stg4 = reinterpret_cast<const float4*>(stg);
    const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    float amax = 0.f;   // max |Y| seen by this thread (the fp16 shadow's range check)
    int64_t img_off = 0;   // kImage: where this lane's 4 consecutive n of entry (k, n..n+3) sit inside a relation's image
    if constexpr (kImage) {
      const int j0 = (int)blockIdx.y * kD + 4 * lane;          // flat output feature = k * d + n
      const int kk = j0 / io.d + io.which * io.d, n = j0 % io.d;
      img_off = (int64_t)(kk >> 6) * (io.d * 128) + (int64_t)(n >> 6) * (64 * 128) + (kk & 63) * 128 +
                (((((n & 63) >> 3) ^ (kk & 7))) << 4) + (n & 7) * 2;
    }
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int64_t row0 = tile * kTile;
      mbar_wait(acc_full(a), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int bb = 0; bb < 2; ++bb) {
        const int e0 = 64 * grp + 32 * bb;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTile + e0), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) stg[e * kD + q * 32 + lane] = __uint_as_float(r[e]);
        float sc8[8];                                  // kImage: this thread's 8 relation scales, in flight over the barrier
        if constexpr (kImage) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t row = row0 + e0 + 8 * q + i;
            sc8[i] = row < M ? __ldg(io.row_scale + row) : 0.f;
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = 8 * q + i;
          const int64_t row = row0 + e0 + e;
          float4 v = stg4[e * (kD / 4) + lane];
          if (row < M) {
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
            if constexpr (kImage) {
              const float sc = sc8[i];
              const __half2 p0 = __floats2half2_rn(v.x * sc, v.y * sc), p1 = __floats2half2_rn(v.z * sc, v.w * sc);
              *reinterpret_cast<uint2*>(io.img + row * io.image_bytes + img_off) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
              continue;
            }
            *reinterpret_cast<float4*>(Y + row * ldy + 4 * lane) = v;
            if (Y16) {
              amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
              const __half2 p0 = __floats2half2_rn(v.x, v.y), p1 = __floats2half2_rn(v.z, v.w);
              *reinterpret_cast<uint2*>(Y16 + row * ldy + 4 * lane) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
            }
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      tc_fence_before();
      mbar_arrive(acc_empty(a));
    }
    if (Y16) {   // speculative unscaled shadow: publish the range, the rescue pass decides whether to rescale
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, sft));
      if (lane == 0) atomic_max_nonneg(y16_scale + 1, amax);
    }
  } else if (warp < kWarpMma) {
    // ------------------------------------------------------------------ producers: split X into hi / lo tiles
    // group g (4 warps) owns chunks with (global chunk counter) % 2 == g, so two chunks are in flight per SM
    const int grp = (warp - kWarpProd) >> 2, pw = (warp - kWarpProd) & 3;
    const int cj = lane & 7;
    int64_t cc = grp;  // global chunk counter of this group's next chunk
    // this group's chunks in order: (tile, c) with c = grp, grp + 2.  The loads of the NEXT chunk are issued before
    // the current one is converted and stored, so the L2 / HBM latency of one chunk hides behind the other's work
    // (a generator Linear re-reads a small, L2-resident X for each of its hundreds of feature blocks: latency, not
    // bandwidth, is what its producers wait for).
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * 2;
    auto load = [&](int64_t idx, float4 (&v)[8]) {
      const int64_t row0 = (blockIdx.x + (idx >> 1) * gridDim.x) * kTile;
      const int c = grp + 2 * (int)(idx & 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t row = row0 + pw * 32 + 4 * i + (lane >> 3);
        v[i] = row < M ? __ldg(reinterpret_cast<const float4*>(X + row * kD + c * 32 + 4 * cj))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store = [&](const float4 (&v)[8]) {   // convert + store one chunk into the ring slot of chunk counter cc
      const int stage = (int)(cc % kStages);
      const uint32_t phase = (uint32_t)((cc / kStages) & 1);
      mbar_wait(empty(stage), phase ^ 1u);
      const uint32_t hi_base = sA + stage * kStage, lo_base = hi_base + kSub;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = pw * 32 + 4 * i + (lane >> 3);
        const uint32_t off = row * 128 + ((cj ^ (row & 7)) << 4);
        if constexpr (kImage) {   // one TF32 product: X rounded to nearest, no remainder tile
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "f"(to_tf32_rna(v[i].x)),
                       "f"(to_tf32_rna(v[i].y)), "f"(to_tf32_rna(v[i].z)), "f"(to_tf32_rna(v[i].w))
                       : "memory");
          continue;
        }
        const float4 h = make_float4(tf32_hi(v[i].x), tf32_hi(v[i].y), tf32_hi(v[i].z), tf32_hi(v[i].w));
        const float4 l = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "f"(h.x), "f"(h.y), "f"(h.z),
                     "f"(h.w)
                     : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_base + off), "f"(l.x), "f"(l.y), "f"(l.z),
                     "f"(l.w)
                     : "memory");
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
      mbar_arrive(full(stage));
      cc += 2;
    };
    // two register sets, alternating: a chunk's loads are in flight while the previous chunk is converted and stored
    float4 va[8], vb[8];
    if (total > 0) load(0, va);
    for (int64_t idx = 0; idx < total; idx += 2) {
      if (idx + 1 < total) load(idx + 1, vb);
      store(va);
      if (idx + 2 < total) load(idx + 2, va);
      if (idx + 1 < total) store(vb);
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    constexpr uint32_t idesc = umma_idesc_tf32(kD, kTile);
    int64_t cc = 0;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * kTile);
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c, ++cc) {
        const int stage = (int)(cc % kStages);
        mbar_wait(full(stage), (uint32_t)((cc / kStages) & 1));
        tc_fence_after();
        const uint32_t hi_addr = sA + stage * kStage;
        if (elect_one()) {
          const uint64_t bh = umma_desc_k128(hi_addr), bl = umma_desc_k128(hi_addr + kSub);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t ah = tmem_base + kWhCol + (uint32_t)(32 * c + 8 * j);
            const uint32_t al = tmem_base + kWlCol + (uint32_t)(32 * c + 8 * j);
            if constexpr (kImage) {
              // the result is rounded to fp16 (11-bit significand) anyway: one TF32 product (operands rounded to
              // nearest) instead of the 3-term split - a third of the tensor work
              umma_tf32_ts(d_tmem, ah, bh + 2 * j, idesc, (uint32_t)(c | j));
            } else {
              umma_tf32_ts(d_tmem, ah, bl + 2 * j, idesc, (uint32_t)(c | j));  // small terms first
              umma_tf32_ts(d_tmem, al, bh + 2 * j, idesc, 1u);
              umma_tf32_ts(d_tmem, ah, bh + 2 * j, idesc, 1u);
            }
          }
          umma_commit(empty(stage));
          if (c + 1 == kChunks) umma_commit(acc_full(a));
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc<kTmemCols>(tmem_base);
}

// ================================================================================================================
// Image mode on fp16 operands (round 2).  The tf32 image kernel above re-stages the fp32 activations of every row
// tile for every feature block (64 KiB in per 32 KiB out: ncu showed 18 % tensor pipe, 19 % DRAM - the SM <-> L2
// interface again).  The result is rounded to fp16 anyway, so the operands can be fp16 too (the same 11-bit
// significand TF32 keeps, power-of-two scales): the activations are converted ONCE into the swizzled tile layout
// the tensor core reads (32 KiB per tile, one bulk copy), the weight block sits in 64 TMEM columns, kind::f16.
// ================================================================================================================
constexpr int kXTileBytes = 2 * kSub;          // 128 rows x 128 k fp16: two 128-byte swizzle atoms of 64 k
constexpr int kFStages = 4;
constexpr int kFSmem = 1024 + kFStages * kXTileBytes + 2 * kStaging + 512;
constexpr uint32_t kFWCol = 256;               // weight block: TMEM columns [256, 320)
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kD >> 4) << 24);

// X [M,128] fp32 -> fp16 tiles in operand order: tile t, k-half s, row r, 16-byte chunk c at chunk (c ^ (r & 7));
// rows >= M are zero.  scale[0] = 1 / s with s the power of two chosen from scale[1] = max |X| (mp_f16_absmax).
__global__ void __launch_bounds__(256)
x16_tiles_kernel(const float* __restrict__ X, int64_t M, int64_t tiles, __half* __restrict__ out,
                 float* __restrict__ scale) {
  const float s = f16_scale_for(scale[1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[0] = 1.f / s;
  const int64_t total = tiles * kTile * 16;    // 16-byte chunks
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i >> 4;
    const int c = (int)(i & 15);               // chunk of 8 k
    const int64_t t = row / kTile;
    const int r = (int)(row % kTile);
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (row < M) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(X + row * kD + 8 * c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(X + row * kD + 8 * c + 4));
      const __half2 p0 = __floats2half2_rn(a.x * s, a.y * s), p1 = __floats2half2_rn(a.z * s, a.w * s);
      const __half2 p2 = __floats2half2_rn(b.x * s, b.y * s), p3 = __floats2half2_rn(b.z * s, b.w * s);
      o.x = *reinterpret_cast<const uint32_t*>(&p0); o.y = *reinterpret_cast<const uint32_t*>(&p1);
      o.z = *reinterpret_cast<const uint32_t*>(&p2); o.w = *reinterpret_cast<const uint32_t*>(&p3);
    }
    const int sub = c >> 3, cc = c & 7;
    uint8_t* dst = reinterpret_cast<uint8_t*>(out) + t * kXTileBytes + sub * kSub + r * 128 + ((cc ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// W [N,128] fp32 -> fp16 in TMEM-load order per 128-feature block: block b | lane quarter q | 16-byte group j (8 k) |
// lane l | 8 halfs, so that loader warp q reads 512 contiguous bytes per instruction and lane l's 16 loads are its
// row's 128 k = 64 TMEM columns.  scale as above (scale[1] = max |W| given).
__global__ void __launch_bounds__(256)
w16_blocks_kernel(const float* __restrict__ W, int64_t N, __half* __restrict__ out, float* __restrict__ scale) {
  const float s = f16_scale_for(scale[1]);
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[0] = 1.f / s;
  const int64_t total = N * 16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i >> 4;
    const int j = (int)(i & 15);
    const float4 a = __ldg(reinterpret_cast<const float4*>(W + n * kD + 8 * j));
    const float4 b = __ldg(reinterpret_cast<const float4*>(W + n * kD + 8 * j + 4));
    const __half2 p0 = __floats2half2_rn(a.x * s, a.y * s), p1 = __floats2half2_rn(a.z * s, a.w * s);
    const __half2 p2 = __floats2half2_rn(b.x * s, b.y * s), p3 = __floats2half2_rn(b.z * s, b.w * s);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&p0); o.y = *reinterpret_cast<const uint32_t*>(&p1);
    o.z = *reinterpret_cast<const uint32_t*>(&p2); o.w = *reinterpret_cast<const uint32_t*>(&p3);
    const int64_t blk = n / kD;
    const int nn = (int)(n % kD);
    const int64_t idx = (((blk * 4 + (nn >> 5)) * 16 + j) * 32 + (nn & 31));   // in 16-byte units
    reinterpret_cast<uint4*>(out)[idx] = o;
  }
}

__device__ __forceinline__ void umma_f16_ts_l(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Yt[128 features, 128 rows] = W16_block[128, 128 k] * X16_tile^T, written as fp16 operand images (ImageOut).
// Warp roles (320 threads, 1 CTA / SM): 0-7 epilogue (warps 0-3 load the weight block into TMEM first) | 8 bulk-copy
// producer (one lane) | 9 MMA issuer + TMEM allocator.
constexpr int kFThreads = 32 * 10;
__global__ void __launch_bounds__(kFThreads, 1)
linear128_f16_images_kernel(const __half* __restrict__ X16, int64_t M, const __half* __restrict__ W16,
                            const float* __restrict__ bias, const float* __restrict__ log_scale,
                            const float* __restrict__ x_scale, const float* __restrict__ w_scale, ImageOut io) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sStg = sA + kFStages * kXTileBytes;
  const uint32_t sBar = sStg + 2 * kStaging;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kFStages + s); };
  auto acc_full = [&](int a) { return sBar + 8u * (2 * kFStages + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * kFStages + 2 + a); };
  const uint32_t tmem_slot = sBar + 8u * (2 * kFStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (M + kTile - 1) / kTile;
  if (bias) bias += (int64_t)blockIdx.y * kD;
  // exp(log_scale) and the two operand scales (exact powers of two) in one factor
  const float alpha = (log_scale ? expf(*log_scale) : 1.f) * x_scale[0] * w_scale[0];

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFStages; ++s) {
      mbar_init(full(s), 1);     // the producer's expect_tx arrival; the bulk copy completes the transaction bytes
      mbar_init(empty(s), 1);    // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 256);
    }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {   // the weight block -> 64 TMEM columns (lane = feature, column = k pair)
    const uint8_t* src = reinterpret_cast<const uint8_t*>(W16) + ((int64_t)blockIdx.y * 4 + warp) * (16 * 32 * 16) + lane * 16;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + kFWCol;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 v = ldg_nc_v4(src + (half * 8 + j) * (32 * 16));
        r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
      }
      tmem_st_32x32(t_row + 32u * half, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue (as in the tf32 image kernel)
    const int grp = warp >> 2, q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem_raw + (sStg - raw) + grp * kStaging);
    const float4* stg4 = reinterpret_cast<const float4*>(stg);
    const float4 b4 = bias ? *reinterpret_cast<const float4*>(bias + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float eb = log_scale ? expf(*log_scale) : 1.f;     // the bias is in real units: only exp(log_scale) applies
    const int j0 = (int)blockIdx.y * kD + 4 * lane;          // flat output feature = k * d + n
    const int kk = j0 / io.d + io.which * io.d, n = j0 % io.d;
    const int64_t img_off = (int64_t)(kk >> 6) * (io.d * 128) + (int64_t)(n >> 6) * (64 * 128) + (kk & 63) * 128 +
                            (((((n & 63) >> 3) ^ (kk & 7))) << 4) + (n & 7) * 2;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int a = it & 1;
      const int64_t row0 = tile * kTile;
      mbar_wait(acc_full(a), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int bb = 0; bb < 2; ++bb) {
        const int e0 = 64 * grp + 32 * bb;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTile + e0), r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) stg[e * kD + q * 32 + lane] = __uint_as_float(r[e]);
        float sc8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t row = row0 + e0 + 8 * q + i;
          sc8[i] = row < M ? __ldg(io.row_scale + row) : 0.f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = 8 * q + i;
          const int64_t row = row0 + e0 + e;
          const float4 v = stg4[e * (kD / 4) + lane];
          if (row < M) {
            const float sc = sc8[i];
            const float y0 = (v.x * alpha + b4.x * eb) * sc, y1 = (v.y * alpha + b4.y * eb) * sc;
            const float y2 = (v.z * alpha + b4.z * eb) * sc, y3 = (v.w * alpha + b4.w * eb) * sc;
            const __half2 p0 = __floats2half2_rn(y0, y1), p1 = __floats2half2_rn(y2, y3);
            *reinterpret_cast<uint2*>(io.img + row * io.image_bytes + img_off) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
      }
      tc_fence_before();
      mbar_arrive(acc_empty(a));
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ producer: one bulk copy per row tile
    if (lane == 0) {
      int64_t cc = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++cc) {
        const int stage = (int)(cc % kFStages);
        mbar_wait(empty(stage), (uint32_t)(((cc / kFStages) & 1) ^ 1));
        mbar_arrive_expect_tx(full(stage), kXTileBytes);
        bulk_g2s(sA + stage * kXTileBytes, reinterpret_cast<const uint8_t*>(X16) + tile * kXTileBytes, kXTileBytes,
                 full(stage));
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    int64_t cc = 0;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it, ++cc) {
      const int a = it & 1;
      mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1u);
      const int stage = (int)(cc % kFStages);
      mbar_wait(full(stage), (uint32_t)((cc / kFStages) & 1));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * kTile);
      const uint32_t addr = sA + stage * kXTileBytes;
      if (elect_one()) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const uint64_t bdesc = umma_desc_k128(addr + sub * kSub);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            umma_f16_ts_l(d_tmem, tmem_base + kFWCol + (uint32_t)(32 * sub + 8 * j), bdesc + 2 * j, kIdescF16,
                          (uint32_t)(sub | j));
        }
        umma_commit(empty(stage));
        umma_commit(acc_full(a));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

bool linear_umma_eligible(int64_t M, int K, int N, int relu, const void* log_scale, const void* X, const void* W,
                          const void* Y) {
  const char* env = getenv("GHF_LINEAR_UMMA");
  if (env && env[0] == '0') return false;
  (void)relu; (void)log_scale;
  // enough work to fill the machine: >= 2^21 outputs (node projection: many rows; generator: many columns)
  return K == kD && N % kD == 0 && N / kD <= 65535 && M * (int64_t)N >= (1 << 21) && M >= 64 &&
         ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(Y)) % 16 == 0);
}

// CTAs per feature block: the grid is gx x nblocks persistent CTAs, each walking ceil(tiles / gx) row tiles after
// loading its block of W into tensor memory (about one tile's worth of time).  Pick gx for the fewest tile-times on
// the critical path - with many feature blocks (weight generators) that is not always 1: 512 blocks on 148 SMs are
// 4 waves of 157 tiles with gx = 1, but 7 waves of 79 with gx = 2.
static int64_t pick_gx(int64_t tiles, int nblocks) {
  const int sms = sm_count();
  int64_t best = 1, best_cost = INT64_MAX;
  const int64_t hi = tiles < sms ? tiles : sms;
  for (int64_t gx = 1; gx <= hi; ++gx) {
    const int64_t waves = (gx * nblocks + sms - 1) / sms;
    const int64_t cost = waves * ((tiles + gx - 1) / gx + 1);
    if (cost < best_cost) { best_cost = cost; best = gx; }
  }
  return best;
}

int linear_umma_launch(const float* X, int64_t M, const float* W, const float* b, int N, int relu,
                       const float* log_scale, float* Y, void* Y16, float* y16_scale, cudaStream_t stream) {
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(linear128_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int nblocks = N / kD;
  const int64_t gx = pick_gx(tiles, nblocks);
  linear128_umma_kernel<false><<<dim3((unsigned)gx, (unsigned)nblocks), kThreads, kSmem, stream>>>(
      X, M, W, b, relu, log_scale, Y, (int64_t)N, reinterpret_cast<__half*>(Y16), y16_scale, ImageOut{});
  GHF_LAUNCH_CHECK();
  return 0;
}

// The last Linear of a weight generator written as fp16 operand images: image[m] entry (k + which * d, n) =
// fp16( exp(log_scale) * (X[m] . W[k * d + n] + b[k * d + n]) * row_scale[m] ).   K = 128, N = d * d.
// the same through fp16 operands (linear128_f16_images_kernel); GHF_IMAGES_TF32=1 keeps the tf32 kernel
static int linear_f16_to_images(const float* X, int64_t M, const float* W, const float* b, int d, int which,
                                const float* log_scale, const float* row_scale, void* images, int64_t image_bytes,
                                cudaStream_t stream) {
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(linear128_f16_images_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmem));
  }
  const int N = d * d;
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int nblocks = N / kD;
  TempBuf x16, w16, scales;
  GHF_CUDA(x16.alloc((size_t)tiles * kXTileBytes, stream));
  GHF_CUDA(w16.alloc((size_t)N * kD * sizeof(__half), stream));
  GHF_CUDA(scales.alloc(4 * sizeof(float), stream));
  float* xs = scales.as<float>();
  float* ws = xs + 2;
  if (int rc = mp_f16_absmax(X, M * (int64_t)kD, xs, stream)) return rc;
  if (int rc = mp_f16_absmax(W, (int64_t)N * kD, ws, stream)) return rc;
  const unsigned cap = (unsigned)sm_count() * 8;
  const int64_t xc = tiles * kTile * 16, wc = (int64_t)N * 16;
  x16_tiles_kernel<<<(unsigned)(cdiv(xc, 256) < cap ? cdiv(xc, 256) : cap), 256, 0, stream>>>(X, M, tiles,
                                                                                            x16.as<__half>(), xs);
  GHF_LAUNCH_CHECK();
  w16_blocks_kernel<<<(unsigned)(cdiv(wc, 256) < cap ? cdiv(wc, 256) : cap), 256, 0, stream>>>(W, N, w16.as<__half>(), ws);
  GHF_LAUNCH_CHECK();
  const int64_t gx = pick_gx(tiles, nblocks);
  ImageOut io{reinterpret_cast<uint8_t*>(images), row_scale, image_bytes, d, which};
  linear128_f16_images_kernel<<<dim3((unsigned)gx, (unsigned)nblocks), kFThreads, kFSmem, stream>>>(
      x16.as<__half>(), M, w16.as<__half>(), b, log_scale, xs, ws, io);
  GHF_LAUNCH_CHECK();
  return 0;
}

int linear_umma_to_images(const float* X, int64_t M, const float* W, const float* b, int d, int which,
                          const float* log_scale, const float* row_scale, void* images, int64_t image_bytes,
                          cudaStream_t stream) {
  {
    const char* env = getenv("GHF_IMAGES_TF32");
    const int N = d * d;
    if (!(env && env[0] == '1') && N % kD == 0 && d % 4 == 0 && N / kD <= 65535 && M > 0 &&
        (reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(W)) % 16 == 0)
      return linear_f16_to_images(X, M, W, b, d, which, log_scale, row_scale, images, image_bytes, stream);
  }
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(linear128_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const int N = d * d;
  GHF_REQUIRE(N % kD == 0 && d % 4 == 0 && N / kD <= 65535, "linear_umma_to_images: hidden_dim %d", d);
  const int64_t tiles = (M + kTile - 1) / kTile;
  const int nblocks = N / kD;
  const int64_t gx = pick_gx(tiles, nblocks);
  ImageOut io{reinterpret_cast<uint8_t*>(images), row_scale, image_bytes, d, which};
  linear128_umma_kernel<true><<<dim3((unsigned)gx, (unsigned)nblocks), kThreads, kSmem, stream>>>(
      X, M, W, b, 0, log_scale, nullptr, (int64_t)N, nullptr, nullptr, io);
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace ghf
