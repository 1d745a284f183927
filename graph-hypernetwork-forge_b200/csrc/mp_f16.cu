// mp_f16.cu — message-passing contraction for hidden_dim 128 with HALF-WIDTH TRANSPORT (GHF_PREC_F16).
//
// What bounds the layer on B200 (tools/l2_paths.cu, profiles/r01_l2_paths_microbench.txt):
//   - the scatter of fp32 result rows into the L2-resident accumulator window costs 24 cycles per 512 B row per
//     SM, whatever the warp count (the SM's store path), i.e. 1.35 ms for 16M edges - a floor no layout removes;
//   - the random row gather runs at HBM speed (7+ TB/s) when >= 128 KB are in flight per SM, and it shares the
//     SM <-> L2 path with the scatter: gathering 512 B rows next to the scatter takes 1.95 ms, 256 B rows 1.5 ms.
// So the node features travel as fp16: h16 = fp16(h) (11-bit significand, the same as the TF32 operand the
// tensor core would round an fp32 value to; power-of-two scales keep values inside the fp16 range), the
// generated weights as fp16(W * 2^k_r) with a per-relation power of two, products accumulate in fp32 in TMEM
// (tcgen05 kind::f16) and the epilogue undoes the scales exactly.  Half the gather bytes, half the shared
// memory per tile (3.5 tiles in flight instead of 1.5), half the tensor-pipe time, and the weight image of a
// relation shrinks to 128 TMEM columns, which leaves room to double-buffer it.
//
// Transposed tile, as in mp_umma_ts.cu:   Dt[128 out-cols, 128 edges] = Wt_r[128, 256] * [h16_src | h16_dst]^T
//   A = Wt_r : TMEM columns [256,384) / [384,512)  (two units' weights; lane = output column, column = k pair)
//   B = rows : gathered by cp.async into a 7 x 32 KiB ring (stage = the src halves or the dst halves of a tile)
//   D = Dt   : TMEM columns [0,128) / [128,256)
// The epilogue needs no shared-memory transpose: thread = output column, so one red.global.add.f32 per edge is
// a coalesced 128 B line of acc[dst_e].
//
// Sixteen epilogue warps, not eight: a warp's 32 reds of a block go to 32 different rows, and that shape needs
// 16 warps to saturate the store path (tools/l2_paths.cu: 2.13 ms with 8 warps, 1.42 ms with 16).
//
// Warp roles (832 threads with 4 producer warps, 1 CTA / SM, persistent):
//   0-15  epilogue (group g = warp/4 owns edges [32g, 32g+32) of each tile, warp%4 = TMEM lane quarter)
//   16-19 weight loaders: global -> registers -> tcgen05.st, one unit ahead of the MMA
//   20    MMA issuer + TMEM allocator
//   21    scheduler: draws units from the global counter, publishes TILE descriptors in shared memory and clears
//         this CTA's share of the next super-block's accumulator rows
//   22-   row-gather producers (4 by default, 8 with GHF_F16_PROD=8): warp p owns 128/P rows of each tile; ids are
//         prefetched one tile ahead
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "mp.cuh"
#include "mp_fuse.cuh"
#include "umma.cuh"

namespace ghf {
namespace {

using namespace ptx;

constexpr int kD = 128;
constexpr int kTile = 128;                 // edges per tile (the N of the transposed product)
constexpr int kRowBytes = kD * 2;          // one fp16 feature row
constexpr int kSub = kTile * 128;          // 128 rows x 128 B (64 halfs of K): 16 KiB, one swizzle atom column
constexpr int kStageBytes = 2 * kSub;      // the src halves (K 0..127) or the dst halves (K 128..255) of a tile
constexpr int kStages = 7;                 // 224 KiB ring = 3.5 tiles of gathers in flight
constexpr int kQueue = 16;                 // tile descriptors between the scheduler and the other roles
constexpr int kEpiWarps = 16, kLoadWarps = 4;
// warp layout: [epilogue 16][weight loaders 4][MMA][scheduler][producers P]  (P = 4 or 8, template parameter)
constexpr int kWarpLoad = kEpiWarps, kWarpMma = kWarpLoad + kLoadWarps, kWarpSched = kWarpMma + 1,
              kWarpProd = kWarpSched + 1;
constexpr int threads_for(int prod_warps) { return 32 * (kWarpProd + prod_warps); }
constexpr uint32_t kTmemCols = 512;
// TMEM: two accumulators [0,256) + two weight buffers [256,512): the next unit's weights load while the current
// unit computes.
// What the per-role cycle accounting (GHF_F16_TRACE=1) shows at c3: per 128-edge tile the 512 reds occupy the SM's
// store path for ~3800 cycles (7.4 cycles per 128 B red), the tensor pipe is busy ~1100 cycles, MMA and producers
// wait ~2900 cycles for the epilogue; descriptor + barrier + TMEM load cost the epilogue warps ~800 cycles per tile.
// Tried against that bubble, A/B in one run: two sets of 8 epilogue warps on alternate tiles with three accumulators
// and one weight buffer (2.25 ms vs 2.23 ms) - no gain, the queued reds already cover the bubble.  The kernel runs
// at the store-path rate; only fewer or narrower reds would make it faster.
constexpr int kAccBufs = 2, kWBufs = 2;
constexpr uint32_t kWCol = kAccBufs * kTile;   // first TMEM column of the weight buffers
static_assert(kAccBufs * kTile + kWBufs * kD <= 512, "TMEM columns");
constexpr int kImageBytes = kD * 2 * kD * 2;   // one relation's Wt image: 128 x 256 fp16 = 64 KiB
constexpr int kBarBytes = 1024;
constexpr int kSmem = 1024 + kStages * kStageBytes + kQueue * 16 + kBarBytes;

constexpr uint32_t kFlagSrcEvictFirst = 1u, kFlagDstEvictLast = 2u, kFlagRedEvictLast = 4u, kFlagWEvictLast = 8u,
                   kFlagWEvictFirst = 16u;
constexpr uint32_t kDefaultFlags = kFlagSrcEvictFirst | kFlagDstEvictLast | kFlagRedEvictLast;
// timing experiments only (results are wrong): drop the reductions / the gathers
constexpr uint32_t kDbgNoRed = 32u, kDbgNoGather = 64u;
// accumulate into the rows as they are (ghf_mp_contract with accumulate != 0): no clearing, same synchronisation
constexpr uint32_t kFlagNoClear = 128u;
// one half of K absent (gradient contractions, ghf_mp_contract with a NULL weight tensor): its rows are neither
// gathered nor multiplied.  Bit 256 << s drops stage s (0 = source rows, 1 = destination rows).
constexpr uint32_t kFlagSkipSrc = 256u, kFlagSkipDst = 512u;

constexpr uint32_t kTileFirst = 1u, kTileLast = 2u, kTileWbuf = 4u;   // descriptor flags

// Instruction descriptor, kind::f16: D fp32, A and B fp16, both K-major.
//   [4,6) D fmt (1 = f32) | [7,10) A fmt (0 = f16) | [10,13) B fmt (0 = f16) | [17,23) N>>3 | [24,29) M>>4
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kD >> 4) << 24);

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- weight image ------------------------------------------------------------------------------------------
// Element (n, k) of Wt_r (k < 128 -> W_msg[r][k][n], else W_self[r][k-128][n]) scaled by 2^k_r, as fp16, at
// half index   piece c = k/64 | quarter q = n/32 | 16-byte group j = (k%64)/8 | lane l = n%32 | k%8
// so that load j of piece c by loader warp q is one contiguous 512 B line set and the 8 loads of a lane are its
// 64 consecutive k, i.e. the 32 TMEM columns [32c, 32c+32) of lane n (two halfs per column, even k low).
__device__ __forceinline__ int wt_half_index(int n, int k) {
  return ((((k >> 6) * 4 + (n >> 5)) * 8 + ((k & 63) >> 3)) * 32 + (n & 31)) * 8 + (k & 7);
}

// One CTA per relation: amax over [W_msg[r]; W_self[r]], k_r = 14 - floor(log2(amax)), scaled fp16 image.
// A NULL tensor stands for zeros; `transposed`: the image is built from W[r]^T (element (k, n) read at [n][k]).
__global__ void __launch_bounds__(256)
pack_f16_kernel(const float* __restrict__ W_msg, const float* __restrict__ W_self, __half* __restrict__ pack,
                float* __restrict__ inv_scale, int transposed) {
  __shared__ float s_max[8];
  __shared__ float s_scale;
  const int64_t r = blockIdx.x;
  float m = 0.f;
  for (int which = 0; which < 2; ++which) {
    const float* W = which ? W_self : W_msg;
    if (W == nullptr) continue;
    const float4* w4 = reinterpret_cast<const float4*>(W + r * kD * kD);
    for (int i = threadIdx.x; i < kD * kD / 4; i += 256) {
      const float4 a = w4[i];
      m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = 0.f;
    for (int i = 0; i < 8; ++i) mm = fmaxf(mm, s_max[i]);
    int e = 0;
    float scale = 1.f;
    if (mm > 0.f && isfinite(mm)) {
      frexpf(mm, &e);                           // mm = f * 2^e, f in [0.5, 1)  ->  mm * 2^(15-e) in [2^14, 2^15)
      e = 15 - e;
      e = e > 100 ? 100 : (e < -100 ? -100 : e);
      scale = ldexpf(1.f, e);
    }
    s_scale = scale;
    inv_scale[r] = 1.f / scale;                 // exact: a power of two
  }
  __syncthreads();
  const float scale = s_scale;
  __half* img = pack + r * (int64_t)(2 * kD * kD);
  // thread = (8 consecutive k, one n): lanes walk n, so every read is a coalesced 128 B line of one W row and
  // the eight halfs of a thread are one 16 B store, consecutive across the warp (512 B).  (Transposed: a thread
  // reads its 8 values from one row - 32 B pieces, the 64 KiB of a relation come through L2 either way.)
  for (int i = threadIdx.x; i < (2 * kD / 8) * kD; i += 256) {
    const int n = i % kD, k0 = (i / kD) * 8;
    const float* W = k0 < kD ? W_msg : W_self;
    const int kk = k0 < kD ? k0 : k0 - kD;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (W != nullptr) {
      const float* src = transposed ? W + (r * kD + n) * kD + kk : W + (r * kD + kk) * kD + n;
      const int step = transposed ? 1 : kD;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __half2 p = __floats2half2_rn(src[(2 * j) * step] * scale, src[(2 * j + 1) * step] * scale);
        w[j] = *reinterpret_cast<const uint32_t*>(&p);
      }
    }
    *reinterpret_cast<uint4*>(img + wt_half_index(n, k0)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// scale[1] = max |x| over `n8` groups of 8 floats (scale[1] must be zero at launch)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, int64_t n8, float* __restrict__ scale) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                       fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0) atomic_max_nonneg(scale + 1, m);
}

// h16 = fp16(h * s), s = f16_scale_for(scale[1]); scale[0] = 1 / s.  RESCUE = true: the shadow was already written
// unscaled by the producer of h (speculatively, fused); only if the range demands it is it rewritten with a scale.
template <bool RESCUE>
__global__ void __launch_bounds__(256)
to_f16_kernel(const float* __restrict__ h, int64_t n8, __half* __restrict__ h16, float* __restrict__ scale) {
  const float amax = scale[1];
  float s = f16_scale_for(amax);
  if (RESCUE && ((amax >= 0.0625f && amax < 32768.f) || !(amax > 0.f))) s = 1.f;   // the unscaled shadow is fine
  if (blockIdx.x == 0 && threadIdx.x == 0) scale[0] = 1.f / s;
  if (RESCUE && s == 1.f) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(h) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(h) + 2 * i + 1);
    const __half2 p0 = __floats2half2_rn(a.x * s, a.y * s), p1 = __floats2half2_rn(a.z * s, a.w * s);
    const __half2 p2 = __floats2half2_rn(b.x * s, b.y * s), p3 = __floats2half2_rn(b.z * s, b.w * s);
    uint4 o;
    o.x = *reinterpret_cast<const uint32_t*>(&p0); o.y = *reinterpret_cast<const uint32_t*>(&p1);
    o.z = *reinterpret_cast<const uint32_t*>(&p2); o.w = *reinterpret_cast<const uint32_t*>(&p3);
    reinterpret_cast<uint4*>(h16)[i] = o;
  }
}

// ---- deterministic mode: a bound on |message| -----------------------------------------------------------------
// words[0] = max over relations r and output columns j of sum_k (|W_msg[r][k][j]| + |W_self[r][k][j]|) (float bits),
// words[1] = max |bias| (float bits); non-negative floats order like their bit patterns, so atomicMax on ints works
// and the result does not depend on the order of the atomics.
__global__ void __launch_bounds__(kD)
det_wnorm_kernel(const float* __restrict__ W_msg, const float* __restrict__ W_self, const float* __restrict__ bias,
                 int32_t* __restrict__ words) {
  __shared__ float s_max[kD / 32], s_bmax[kD / 32];
  const int64_t r = blockIdx.x;
  const int j = threadIdx.x;
  float sum = 0.f;
  for (int k = 0; k < kD; ++k)
    sum += fabsf(W_msg[(r * kD + k) * kD + j]) + fabsf(W_self[(r * kD + k) * kD + j]);
  float b = bias ? fabsf(bias[r * kD + j]) : 0.f;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    sum = fmaxf(sum, __shfl_xor_sync(0xffffffffu, sum, s));
    b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, s));
  }
  if ((j & 31) == 0) { s_max[j >> 5] = sum; s_bmax[j >> 5] = b; }
  __syncthreads();
  if (j == 0) {
    for (int i = 1; i < kD / 32; ++i) { sum = fmaxf(sum, s_max[i]); b = fmaxf(b, s_bmax[i]); }
    atomicMax(words, __float_as_int(sum));
    atomicMax(words + 1, __float_as_int(b));
  }
}
// words[2] = eB with |message| < 2^eB: |h_u W_msg + h_v W_self + bias| <= max|h| * column sum + max|bias|, one more
// bit for the rounding of the fp16 operands
__global__ void det_exponent_kernel(const float* __restrict__ h_scale, int32_t* __restrict__ words) {
  const float bound = h_scale[1] * __int_as_float(words[0]) + __int_as_float(words[1]);
  int e = -100;
  if (bound > 0.f && isfinite(bound)) {
    frexpf(bound, &e);        // bound = f * 2^e, f in [0.5, 1): bound < 2^e
    e += 1;
  }
  words[2] = e > 60 ? 60 : (e < -100 ? -100 : e);
}

template <int kProdWarps, bool kTrace, bool kDet>
__global__ void __launch_bounds__(threads_for(kProdWarps), 1)
mp_f16_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
              const int32_t* __restrict__ unit_rel, int64_t num_units, const int32_t* __restrict__ src_sorted,
              const int32_t* __restrict__ dst_sorted, const __half* __restrict__ h16, int64_t dst_lo,
              const float* __restrict__ h_scale, const __half* __restrict__ wpack,
              const float* __restrict__ w_inv_scale,
              const float* __restrict__ bias, float* __restrict__ acc, int* __restrict__ unit_counter,
              const int32_t* __restrict__ unit_phase, int* __restrict__ zero_done, int num_phases, int phase_lo,
              int sb_nodes, int64_t num_local, uint32_t flags, long long* __restrict__ trace,
              const int32_t* __restrict__ indeg, const int32_t* __restrict__ det_words) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sQ = sA + kStages * kStageBytes;          // kQueue x int4 tile descriptors
  const uint32_t sBar = sQ + kQueue * 16;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  const uint32_t bar2 = sBar + 8u * (2 * kStages);
  auto acc_full = [&](int a) { return bar2 + 8u * a; };             // up to 4 accumulators
  auto acc_empty = [&](int a) { return bar2 + 32u + 8u * a; };
  auto w_full = [&](int b) { return bar2 + 64u + 8u * b; };          // up to 2 weight buffers
  auto w_empty = [&](int b) { return bar2 + 80u + 8u * b; };
  const uint32_t q_full0 = bar2 + 96u;
  const uint32_t q_empty0 = q_full0 + 8u * kQueue;
  const uint32_t tmem_slot = q_empty0 + 8u * kQueue;
  volatile int4* q_ptr = reinterpret_cast<volatile int4*>(smem_raw + (sQ - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cycle accounting of CTA 0 (kTrace instantiation only, GHF_F16_TRACE=1): trace[role * 8 + slot] += cycles
  const bool tracing = kTrace && blockIdx.x == 0 && lane == 0;
  long long tr[kTrace ? 6 : 1] = {};
  auto tick = [&]() -> long long {
    if constexpr (kTrace) return tracing ? clock64() : 0;
    return 0;
  };
  auto tadd = [&](int k, long long v) {
    if constexpr (kTrace) tr[k] += v;
  };
  constexpr int kConsumers = kEpiWarps + kProdWarps + 1 + kLoadWarps;   // warps that read every descriptor
  constexpr int kRowsPerWarp = kTile / kProdWarps;                      // 32 or 16

  // tile-descriptor queue (consumer side).  acquire() blocks until descriptor `idx` is published and returns it
  // {x = first sorted edge (-1: no more work), y = rows, z = relation, w = flags}; release() frees the slot once
  // per warp.  Roles may hold several descriptors (the producers look one tile ahead).
  auto q_acquire = [&](uint32_t idx) -> int4 {
    mbar_wait(q_full0 + 8u * (idx % kQueue), (idx / kQueue) & 1u);
    const volatile int4* p = q_ptr + (idx % kQueue);
    int4 t;
    t.x = p->x; t.y = p->y; t.z = p->z; t.w = p->w;
    return t;
  };
  auto q_release = [&](uint32_t idx) {
    __syncwarp();
    if (lane == 0) mbar_arrive(q_empty0 + 8u * (idx % kQueue));
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 32 * kProdWarps);  // one cp.async-completion arrival per producer thread
      mbar_init(empty(s), 1);               // tcgen05.commit
    }
    for (int a = 0; a < kAccBufs; ++a) {
      mbar_init(acc_full(a), 1);                // tcgen05.commit
      mbar_init(acc_empty(a), kEpiWarps);       // one arrival per epilogue warp
    }
    for (int b = 0; b < kWBufs; ++b) {
      mbar_init(w_full(b), 32 * kLoadWarps);
      mbar_init(w_empty(b), 1);                 // tcgen05.commit
    }
    for (int q = 0; q < kQueue; ++q) {
      mbar_init(q_full0 + 8u * q, 1);
      mbar_init(q_empty0 + 8u * q, kConsumers);
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < kEpiWarps) {
    // ------------------------------------------------------------------ epilogue: Dt -> red.f32 rows
    // All 16 warps work on every tile: group g = warp / 4 owns edges [32g, 32g + 32), warp % 4 = TMEM lane quarter.
    const int grp = warp >> 2, q = warp & 3;
    const int col = 32 * q + lane;                       // this thread's output column = its TMEM lane
    float* acc_col = acc + col;
    const float h_inv = h_scale[0];                      // h = h16 * h_inv (exact power of two)
    const int e0 = 32 * grp;
    // Everything a tile's reductions need from global memory (the destination id of edge e0 + lane, the relation's
    // bias entry and scale) is fetched one tile ahead, so that no load latency sits between "accumulator ready" and
    // the first red.  (Loaded values are kept raw - arithmetic on them here would wait for the load right away.)
    struct TileRegs { int dst; float bias_n, w_inv; int deg; };
    auto fetch = [&](const int4& t) -> TileRegs {
      TileRegs x{-1, 0.f, 1.f, 1};
      if (t.x < 0) return x;
      if (e0 + lane < t.y) {
        x.dst = dst_sorted[t.x + e0 + lane];
        if constexpr (kDet) x.deg = indeg[x.dst];
      }
      x.bias_n = bias ? bias[(int64_t)t.z * kD + col] : 0.f;
      x.w_inv = w_inv_scale[t.z];
      return x;
    };
    // Deterministic mode: every contribution is rounded to fixed point BEFORE it is added - integer additions
    // commute, so neither the order of the atomics nor the order of the edges inside a run matters.  The scale of
    // destination v is 2^k_v, k_v = 30 - bits(indeg_v) - eB with |message| < 2^eB (det_words[2], from a bound on
    // |h| and on the column sums of the relation matrices): |sum_v| 2^k_v < 2^30 cannot overflow an int32.
    int det_eB = 0;
    if constexpr (kDet) det_eB = det_words[2];
    int4 cur = q_acquire(0);
    TileRegs cr = fetch(cur);
    for (uint32_t it = 0; cur.x >= 0; ++it) {
      long long t0 = tick();
      const int4 nxt = q_acquire(it + 1);
      const TileRegs nr = fetch(nxt);
      const int a = (int)(it % kAccBufs);
      long long t1 = tick();
      mbar_wait(acc_full(a), (it / kAccBufs) & 1u);
      tc_fence_after();
      long long t2 = tick();
      tadd(0, t1 - t0); tadd(1, t2 - t1); tadd(5, 1);
      if (e0 < cur.y) {
        const float inv = cr.w_inv * h_inv;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTile + e0), r);
        tmem_ld_wait();
        t1 = tick();
        tadd(2, t1 - t2);
        if constexpr (kDet) {
          int run = 0, run_dst = -1;
          int* acc_i = reinterpret_cast<int*>(acc_col);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int dsti = __shfl_sync(0xffffffffu, cr.dst, e);
            const int degi = __shfl_sync(0xffffffffu, cr.deg, e);
            if (dsti != run_dst) {
              if (run_dst >= 0) atomicAdd(acc_i + (int64_t)run_dst * kD, run);
              run = 0;
              run_dst = dsti;
            }
            const int k = 30 - (32 - __clz(max(degi, 1))) - det_eB;
            const float scale = __int_as_float((uint32_t)(127 + max(-126, min(127, k))) << 23);
            run += __float2int_rn(fmaf(__uint_as_float(r[e]), inv, cr.bias_n) * scale);
          }
          if (run_dst >= 0) atomicAdd(acc_i + (int64_t)run_dst * kD, run);
        } else if (!(flags & kDbgNoRed)) {
          // Segmented sum along the sorted order: edges of one (destination, relation) pair are adjacent, so their
          // contributions are added in a register and leave as ONE red (multigraphs, hub destinations); with all
          // destinations distinct this is one red per edge.
          float run = 0.f;
          int run_dst = -1;
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const int dsti = __shfl_sync(0xffffffffu, cr.dst, e);
            if (dsti != run_dst) {
              if (run_dst >= 0) red_add_f32(acc_col + (int64_t)run_dst * kD, run);
              run = 0.f;
              run_dst = dsti;
            }
            run += fmaf(__uint_as_float(r[e]), inv, cr.bias_n);
          }
          if (run_dst >= 0) red_add_f32(acc_col + (int64_t)run_dst * kD, run);
        }
        t2 = tick();
        tadd(3, t2 - t1);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(a));
      q_release(it);
      tadd(4, tick() - t2);
      cur = nxt;
      cr = nr;
    }
    if constexpr (kTrace)
      if (tracing && warp == 0)                          // acquire+fetch | acc_full wait | tmem ld | reds | release | tiles
        for (int k = 0; k < 6; ++k) trace[k] = tr[k];
  } else if (warp >= kWarpProd) {
    // ------------------------------------------------------------------ row-gather producers
    // Warp pw owns kRowsPerWarp consecutive rows of every tile; lane l keeps the source and destination id of
    // its row l % kRowsPerWarp.  One cp.async instruction moves two 256 B rows (16 lanes x 16 B each).
    const int pw = warp - kWarpProd;
    const int l16 = lane & 15, hi = lane >> 4;
    const uint64_t pol_src = (flags & kFlagSrcEvictFirst) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_dst = (flags & kFlagDstEvictLast) ? policy_evict_last() : policy_evict_normal();
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h16) + l16 * 16;
    struct Ids { int src, dst; };                        // raw loads: converted where they are used
    auto ids_of = [&](const int4& t) -> Ids {
      const int row = kRowsPerWarp * pw + lane % kRowsPerWarp;
      if (t.x < 0 || row >= t.y) return Ids{-1, -1};
      return Ids{src_sorted[t.x + row], dst_sorted[t.x + row]};
    };
    int stage = 0;
    uint32_t phase = 0;
    int4 cur = q_acquire(0);
    Ids ids = ids_of(cur);
    for (uint32_t it = 0; cur.x >= 0; ++it) {
      const int4 nxt = q_acquire(it + 1);
      const Ids ids_nxt = ids_of(nxt);                   // in flight while this tile's rows are issued
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {                      // s = 0: source rows, s = 1: destination rows
        if (flags & (kFlagSkipSrc << s)) continue;       // this half of K is absent
        const long long t0 = tick();
        mbar_wait(empty(stage), phase ^ 1u);
        tadd(0, tick() - t0);
        const uint32_t base = sA + stage * kStageBytes + (l16 >> 3) * kSub;
        const uint64_t pol = s ? pol_dst : pol_src;
        const int mine = s ? ids.dst : ids.src;
        const uint8_t* table = s ? hb + dst_lo * kRowBytes : hb;   // destination ids are local to the rank's range
        if (!(flags & kDbgNoGather)) {
#pragma unroll
          for (int i = 0; i < kRowsPerWarp / 2; ++i) {
            const int rl = 2 * i + hi;                   // row within the warp's share
            const int idx = __shfl_sync(0xffffffffu, mine, rl);
            const int row = kRowsPerWarp * pw + rl;
            const uint32_t to = base + row * 128 + (((l16 & 7) ^ (row & 7)) << 4);
            if (idx >= 0) cp_async_16_hint(to, table + (int64_t)idx * kRowBytes, pol);
          }
        }
        cp_async_arrive_noinc(full(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      q_release(it);
      cur = nxt;
      ids = ids_nxt;
      tadd(5, 1);
    }
    if constexpr (kTrace)
      if (tracing && warp == kWarpProd) {
        trace[16] = tr[0];                               // empty(stage) wait
        trace[21] = tr[5];
      }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane)
    int stage = 0;
    uint32_t phase = 0, wph = 0;                         // wph bit b: parity of the next w_full(b) wait
    const long long t_begin = tick();
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      if (t.x < 0) break;
      const uint32_t tf = (uint32_t)t.w;
      q_release(it);
      const int a = (int)(it % kAccBufs);
      const int wb = (tf & kTileWbuf) ? 1 : 0;
      long long t0 = tick();
      mbar_wait(acc_empty(a), ((it / kAccBufs) & 1u) ^ 1u);
      long long t1 = tick();
      if (tf & kTileFirst) {
        mbar_wait(w_full(wb), (wph >> wb) & 1u);
        wph ^= 1u << wb;
      }
      tadd(0, t1 - t0); tadd(1, tick() - t1); tadd(5, 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * kTile);
      const uint32_t w_tmem = tmem_base + kWCol + (uint32_t)(wb * kD);
      const bool leader = elect_one();                   // the one lane that issues (and commits) this tile's MMAs
      bool first = true;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        if (flags & (kFlagSkipSrc << s)) continue;       // this half of K is absent (gradient contractions)
        t0 = tick();
        mbar_wait(full(stage), phase);
        tadd(2, tick() - t0);
        fence_proxy_async();                             // cp.async (generic proxy) writes -> tensor-core reads
        tc_fence_after();
        const uint32_t stage_addr = sA + stage * kStageBytes;
        if (leader) {
#pragma unroll
          for (int cs = 0; cs < 2; ++cs) {
            const uint64_t bdesc = umma_desc_k128(stage_addr + cs * kSub);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16_ts(d_tmem, w_tmem + (uint32_t)(32 * (2 * s + cs) + 8 * j), bdesc + 2 * j, kIdesc,
                          first ? (uint32_t)(cs | j) : 1u);
          }
          umma_commit(empty(stage));
        }
        first = false;
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      if (leader) {
        umma_commit(acc_full(a));
        if (tf & kTileLast) umma_commit(w_empty(wb));    // every MMA that reads this unit's weights is done
      }
      __syncwarp();
    }
    if constexpr (kTrace)
      if (tracing) {
        for (int k = 0; k < 6; ++k) trace[8 + k] = tr[k];  // acc_empty | w_full | full(stage) waits | - | - | tiles
        trace[14] = tick() - t_begin;                      // cycles from the first to the last tile of this CTA
      }
  } else if (warp == kWarpSched) {
    // ------------------------------------------------------------------ scheduler (+ accumulator clearing)
    // The accumulator rows of super-block ("phase") p are zeroed INSIDE this kernel, one phase ahead of their
    // first reduction: every CTA clears its 1/gridDim share of phase p + 1 when it first draws a unit of phase p,
    // and no unit of phase p is published before all shares of p are in (zero_done[p] == gridDim.x).  The zero
    // lines are created in L2 by full-line stores, so the reductions never fetch accumulator lines from HBM and
    // there is no separate 1.28 GB clear pass.  (All CTAs are co-resident: 1 CTA / SM, grid <= SM count.)
    uint32_t qi = 0, wb = 0;
    // this launch covers super-blocks [phase_lo, num_phases) of the graph (the units it was given lie in them)
    int my_zeroed = phase_lo - 1, ready_phase = -1;
    auto publish = [&](int start, int rows, int rel, uint32_t tf) {
      if (lane == 0) {
        mbar_wait(q_empty0 + 8u * (qi % kQueue), ((qi / kQueue) & 1u) ^ 1u);
        volatile int4* p = q_ptr + (qi % kQueue);
        p->x = start; p->y = rows; p->z = rel; p->w = (int)tf;
        mbar_arrive(q_full0 + 8u * (qi % kQueue));       // release: the descriptor is visible to the waiters
      }
      ++qi;
    };
    auto draw = [&]() -> int64_t {
      int v = 0;
      if (lane == 0) v = atomicAdd(unit_counter, 1);
      return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    auto clear_share = [&](int p) {
      const int64_t lo = (int64_t)p * sb_nodes;
      const int64_t hi = min(lo + (int64_t)sb_nodes, num_local);
      const int64_t share = (hi - lo + gridDim.x - 1) / gridDim.x;
      const int64_t r0 = lo + (int64_t)blockIdx.x * share, r1 = min(hi, r0 + share);
      float4* row = reinterpret_cast<float4*>(acc) + lane;
      if (flags & kFlagNoClear) {
      } else if (flags & kFlagRedEvictLast) {                   // keep the zero lines in L2 until their reductions arrive
        const uint64_t pol = policy_evict_last();
        for (int64_t r = r0; r < r1; ++r)
          asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%1,%1,%1}, %2;" ::"l"(row + r * (kD / 4)), "f"(0.f),
                       "l"(pol)
                       : "memory");
      } else {
        for (int64_t r = r0; r < r1; ++r) row[r * (kD / 4)] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        atomicAdd(&zero_done[p], 1);
      }
    };
    int64_t u = draw();
    while (u < num_units) {
      const int start = unit_start[u], count = unit_count[u], rel = unit_rel[u], ph = unit_phase[u];
      const int64_t u_next = draw();                     // its latency hides behind the work below
      const int ahead = min(ph + 1, num_phases - 1);
      while (my_zeroed < ahead) clear_share(++my_zeroed);
      if (ph != ready_phase) {
        if (lane == 0) fuse::spin_until_at_least(&zero_done[ph], (int)gridDim.x);
        __syncwarp();
        ready_phase = ph;
      }
      for (int t0 = 0; t0 < count; t0 += kTile) {
        const uint32_t tf = (t0 == 0 ? kTileFirst : 0u) | (t0 + kTile >= count ? kTileLast : 0u) |
                            (wb ? kTileWbuf : 0u);
        publish(start + t0, min(kTile, count - t0), rel, tf);
      }
      wb = (wb + 1u) % kWBufs;
      u = u_next;
    }
    while (my_zeroed < num_phases - 1) clear_share(++my_zeroed);   // rows without in-edges are cleared too
    publish(-1, 0, 0, 0);
    publish(-1, 0, 0, 0);                                // producers and epilogue look one descriptor ahead
  } else if (warp >= kWarpLoad && warp < kWarpMma) {
    // ------------------------------------------------------------------ weight loaders: Wt_r -> TMEM
    // thread = output column n = 32 * quarter + lane = TMEM lane; 128 columns (256 k) in 4 pieces of 32
    const int quarter = warp & 3;   // a warp reaches TMEM lanes [32 (warp % 4), +32) only
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + kWCol;
    const uint64_t pol_w = (flags & kFlagWEvictLast) ? policy_evict_last()
                           : (flags & kFlagWEvictFirst) ? policy_evict_first() : policy_evict_normal();
    uint32_t eph = 0;                                    // bit b: parity of the next w_empty(b) wait
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      if (t.x < 0) break;
      const uint32_t tf = (uint32_t)t.w;
      q_release(it);
      if (!(tf & kTileFirst)) continue;
      const int wb = (tf & kTileWbuf) ? 1 : 0;
      const uint8_t* img = reinterpret_cast<const uint8_t*>(wpack) + (int64_t)t.z * kImageBytes +
                           quarter * (8 * 32 * 16) + lane * 16;
      uint32_t r[32];
      auto fetch = [&](int piece) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = ldg_v4_hint(img + piece * (4 * 8 * 32 * 16) + j * (32 * 16), pol_w);
          r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
      };
      fetch(0);                                          // in flight while the buffer is still being read
      mbar_wait(w_empty(wb), ((eph >> wb) & 1u) ^ 1u);
      eph ^= 1u << wb;
      tc_fence_after();
      const uint32_t dst = t_row + (uint32_t)(wb * kD);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tmem_st_32x32(dst + 32u * c, r);
        tmem_st_wait();
        if (c < 3) fetch(c + 1);
      }
      tc_fence_before();
      mbar_arrive(w_full(wb));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc<kTmemCols>(tmem_base);
}

uint32_t env_flags() {
  const char* env = getenv("GHF_MP_FLAGS");
  return env ? (uint32_t)atoi(env) : kDefaultFlags;
}

}  // namespace

bool mp_f16_supported(int d) { return d == kD; }

// the three words of the deterministic mode inside the sync words (ints 8..10: clear of the unit counter's line)
constexpr int kDetWordsAt = 8;
const int32_t* mp_f16_det_words(const int* sync_words) { return sync_words + kDetWordsAt; }

// [0, 256): counters | then two int32 per super-block (mp_f16_kernel uses one, mp_f16_fused_kernel both)
int64_t mp_f16_sync_bytes(const ghf_graph* g) { return align_up(256 + g->num_phases * 8, 256); }

int64_t mp_f16_pack_bytes(int num_rel) {
  return align_up((int64_t)num_rel * kImageBytes, 256) + align_up((int64_t)num_rel * 4, 256);
}

int mp_f16_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch,
                cudaStream_t stream, bool transposed) {
  GHF_REQUIRE(g->hidden_dim == kD, "mp_f16: hidden_dim must be %d", kD);
  return mp_f16_pack_rel(g->num_rel, W_msg, W_self, pack_scratch, stream, transposed);
}

int mp_f16_pack_rel(int num_rel, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream,
                    bool transposed) {
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(W_msg) | reinterpret_cast<uintptr_t>(W_self) |
               reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_f16: W_msg / W_self / scratch must be 16-byte aligned");
  __half* img = reinterpret_cast<__half*>(pack_scratch);
  float* inv = reinterpret_cast<float*>(reinterpret_cast<char*>(pack_scratch) +
                                        align_up((int64_t)num_rel * kImageBytes, 256));
  pack_f16_kernel<<<(unsigned)num_rel, 256, 0, stream>>>(W_msg, W_self, img, inv, transposed ? 1 : 0);
  GHF_LAUNCH_CHECK();
  return 0;
}

static unsigned stride_grid(int64_t n8) {
  const int64_t want = cdiv(n8, 256), cap = (int64_t)sm_count() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

int mp_f16_absmax(const float* x, int64_t elems, float* scale, cudaStream_t stream) {
  GHF_REQUIRE(elems % 8 == 0, "mp_f16: element count must be a multiple of 8");
  GHF_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, "mp_f16: x must be 16-byte aligned");
  GHF_CUDA(cudaMemsetAsync(scale + 1, 0, sizeof(float), stream));
  if (elems == 0) return 0;
  absmax_kernel<<<stride_grid(elems / 8), 256, 0, stream>>>(x, elems / 8, scale);
  GHF_LAUNCH_CHECK();
  return 0;
}

int mp_f16_convert(const float* h, int64_t elems, void* h16, float* scale, bool rescue, cudaStream_t stream) {
  GHF_REQUIRE(elems % 8 == 0, "mp_f16: element count must be a multiple of 8");
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(h16)) % 16 == 0,
              "mp_f16: h / h16 must be 16-byte aligned");
  GHF_REQUIRE(scale != nullptr, "mp_f16: the fp16 shadow needs its scale words");
  __half* out = reinterpret_cast<__half*>(h16);
  if (rescue)
    to_f16_kernel<true><<<stride_grid(elems / 8), 256, 0, stream>>>(h, elems / 8, out, scale);
  else
    to_f16_kernel<false><<<stride_grid(elems / 8), 256, 0, stream>>>(h, elems / 8, out, scale);
  GHF_LAUNCH_CHECK();
  return 0;
}

int mp_f16_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias, float* acc,
                  const void* pack_scratch, int* sync_words, cudaStream_t stream, bool keep_acc, int skip_half,
                  int phase_lo, int phase_hi, const float* det_W_msg, const float* det_W_self) {
  const bool det = det_W_msg != nullptr && det_W_self != nullptr;
  GHF_REQUIRE(!det || (!keep_acc && skip_half == 0), "mp_f16: the deterministic mode covers the plain layer only");
  if (phase_hi < 0) phase_hi = (int)g->num_phases;
  GHF_REQUIRE(0 <= phase_lo && phase_lo <= phase_hi && phase_hi <= g->num_phases, "mp_f16: bad super-block range");
  GHF_REQUIRE(h16_scale != nullptr, "mp_f16: the fp16 shadow needs its scale words");
  GHF_REQUIRE(g->hidden_dim == kD, "mp_f16: hidden_dim must be %d", kD);
  GHF_REQUIRE(g->unit_edges % kTile == 0, "mp_f16: unit_edges=%d must be a multiple of %d", g->unit_edges, kTile);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h16) | reinterpret_cast<uintptr_t>(acc) |
               reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_f16: h16 / acc / scratch must be 16-byte aligned");
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_kernel<4, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_kernel<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_kernel<4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  const char* penv = getenv("GHF_F16_PROD");
  const int prod = penv ? atoi(penv) : 4;
  const __half* img = reinterpret_cast<const __half*>(pack_scratch);
  const float* inv = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pack_scratch) +
                                                    align_up((int64_t)g->num_rel * kImageBytes, 256));
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  long long* trace = nullptr;   // GHF_F16_TRACE=1: cycle accounting of CTA 0 printed after the launch (synchronises)
  TempBuf trace_buf;
  if (getenv("GHF_F16_TRACE")) {
    GHF_CUDA(trace_buf.alloc(32 * sizeof(long long), stream));
    GHF_CUDA(cudaMemsetAsync(trace_buf.p, 0, 32 * sizeof(long long), stream));
    trace = trace_buf.as<long long>();
  }
  // sync words (zero at launch): [0] unit counter, [64 + p] cleared shares of phase p
  GHF_CUDA(cudaMemsetAsync(sync_words, 0, mp_f16_sync_bytes(g), stream));
  int* unit_counter = sync_words;
  int* zero_done = sync_words + 64;
  int32_t* det_words = sync_words + kDetWordsAt;        // [0] column-sum max, [1] |bias| max, [2] eB (mp_f16_det_words)
  if (det) {
    det_wnorm_kernel<<<(unsigned)g->num_rel, kD, 0, stream>>>(det_W_msg, det_W_self, bias, det_words);
    GHF_LAUNCH_CHECK();
    det_exponent_kernel<<<1, 1, 0, stream>>>(h16_scale, det_words);
    GHF_LAUNCH_CHECK();
  }
#define GHF_F16_LAUNCH(P, T, D)                                                                                   \
  mp_f16_kernel<P, T, D><<<(unsigned)grid, threads_for(P), kSmem, stream>>>(                                     \
      g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted,                    \
      reinterpret_cast<const __half*>(h16), g->dst_lo, h16_scale, img, inv, bias, acc, unit_counter, g->unit_phase, \
      zero_done, phase_hi, phase_lo, g->sb_nodes, g->num_local, env_flags() | (keep_acc ? kFlagNoClear : 0u) | (skip_half == 1 ? kFlagSkipSrc : skip_half == 2 ? kFlagSkipDst : 0u), trace, g->indeg, det_words)
  if (det) GHF_F16_LAUNCH(4, false, true);
  else if (trace) GHF_F16_LAUNCH(4, true, false);
  else if (prod == 8) GHF_F16_LAUNCH(8, false, false);
  else GHF_F16_LAUNCH(4, false, false);
#undef GHF_F16_LAUNCH
  GHF_LAUNCH_CHECK();
  if (trace) {
    long long t[32];
    GHF_CUDA(cudaMemcpyAsync(t, trace, sizeof(t), cudaMemcpyDeviceToHost, stream));
    GHF_CUDA(cudaStreamSynchronize(stream));
    auto per = [&](int i, int n) { return t[n] ? (double)t[i] / (double)t[n] : 0.0; };
    fprintf(stderr,
            "mp_f16 trace (CTA 0, cycles per tile): epilogue[acquire %.0f | acc_full %.0f | tmem %.0f | reds %.0f | "
            "release %.0f] x %lld   mma[acc_empty %.0f | w_full %.0f | full %.0f] x %lld   producer[empty %.0f] x %lld"
            "   total %lld cycles\n",
            per(0, 5), per(1, 5), per(2, 5), per(3, 5), per(4, 5), t[5], per(8, 13), per(9, 13), per(10, 13), t[13],
            per(16, 21), t[21], t[14]);
  }
  return 0;
}

}  // namespace ghf
