// mp_f16_fused.cu — the whole layer in ONE kernel for hidden_dim 128 on the f16 engine (HG:201-230 + HG:289-296):
// contraction, per-destination mean, self-loop, residual, ReLU, LayerNorm and the fp16 shadow of the result.
//
// What round 1 measured (profiles/r01_ncu_mp_f16_summary.txt, profiles/r02_scatter_paths_microbench.txt):
//   * the per-edge fp32 reductions are the wall: 24 cycles per 512 B row per SM whatever instruction carries them
//     (red.f32, red.v4.f32, the TMA bulk reduction), 6.0 TB/s chip-wide - IF the rows they hit are in L2;
//   * they were not: the accumulator was [N, d] (1.28 GB at c3) with a moving window of "hot" rows, every line of it
//     marked evict_last at some point, and 58 % of the reduction sectors missed L2 (a scatter into a moving window
//     costs 1.87 ms against 1.37 ms into a fixed one, tools/l2_paths.cu);
//   * the separate epilogue kernel re-read that accumulator from HBM (4.2 GB per layer, 17 % of the forward).
// So the accumulator is no longer an [N, d] array.  It is a RING of `slots` windows of `sb_nodes` rows (2-3 x 16-25 MB,
// the same addresses for the whole kernel, hence L2-resident): super-block ("phase") p reduces into slot p % slots,
// and as soon as every CTA has left phase p the CTAs run the row epilogue of p straight out of L2 - mean, residual,
// ReLU, LayerNorm, fp32 + fp16 result rows - and hand the slot back zeroed.  The accumulator never reaches HBM, the
// epilogue kernel and the in-kernel clearing pass of mp_f16.cu are gone.
//
// Grid-wide protocol (all CTAs co-resident: 1 CTA / SM, grid <= SM count; sync words zero at launch):
//   left[p]  += 1 per epilogue warp once all its reductions into phase p are performed (fence, then atomic)
//   epi[p]   += 1 per epilogue warp once its rows of phase p are written and the slot rows are zero again
//   a CTA's scheduler publishes, in queue order:  tiles of phase q  only after  epi[q - slots] is complete,
//                                                 LEAVE(a..b)       when it draws the first unit beyond phase b-1,
//                                                 EPI(a)            once left[a] is complete (checked at every unit
//                                                                   draw; forced before it would block on epi[a]).
//   Progress: the smallest phase any CTA waits for only needs CTAs that are not blocked (they are in a smaller or
//   equal phase, working on tiles), so some CTA can always move.
//
// Tile pipeline (unchanged from mp_f16.cu): transposed product Dt[128 out, 128 edges] = Wt_r[128, 256] * [h16_src |
// h16_dst]^T, A = Wt_r from TMEM (double-buffered), B = rows gathered by cp.async into a 7 x 32 KiB ring, D in TMEM
// (double-buffered); 16 epilogue warps (thread = output column: one coalesced 128 B red per edge and warp),
// 4 weight loaders, 1 MMA warp, 1 scheduler warp, 4 gather producers.
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
constexpr int kTile = 128;
constexpr int kRowBytes = kD * 2;
constexpr int kSub = kTile * 128;
constexpr int kStageBytes = 2 * kSub;
constexpr int kStages = 7;
constexpr int kQueue = 16;
constexpr int kEpiWarps = 16, kLoadWarps = 4;
// gather producers: 2 warps (24 warps in all: 80 registers per thread) or 4 (26 warps: 72) - template parameter
constexpr int kWarpLoad = kEpiWarps, kWarpMma = kWarpLoad + kLoadWarps, kWarpSched = kWarpMma + 1,
              kWarpProd = kWarpSched + 1;
constexpr int threads_for(int prod_warps) { return 32 * (kWarpProd + prod_warps); }
constexpr uint32_t kTmemCols = 512;
constexpr int kAccBufs = 2, kWBufs = 2;
constexpr uint32_t kWCol = kAccBufs * kTile;
constexpr int kImageBytes = kD * 2 * kD * 2;
constexpr int kBarBytes = 1024;
constexpr int kSmem = 1024 + kStages * kStageBytes + kQueue * 16 + kBarBytes;

// behaviour flags (GHF_FUSED_FLAGS)
constexpr uint32_t kFlagSrcEvictFirst = 1u, kFlagDstEvictLast = 2u, kFlagRingEvictLast = 4u, kFlagWEvictLast = 8u;
constexpr uint32_t kDefaultFlags = kFlagSrcEvictFirst | kFlagRingEvictLast;
// timing experiments only (results are wrong): row epilogue without its body / tiles without reductions
constexpr uint32_t kDbgNoRowWork = 32u, kDbgNoRed = 64u;

// descriptor flags (int4.w, low byte; the phase of a tile sits above)
constexpr uint32_t kTileFirst = 1u, kTileLast = 2u, kTileWbuf = 4u, kDescLeave = 8u, kDescEpi = 16u, kDescEnd = 32u;
constexpr uint32_t kDescNotTile = kDescLeave | kDescEpi | kDescEnd;

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

__device__ __forceinline__ float4 ld_cg_v4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_v4_hint(float* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}

struct FusedParams {
  const int32_t *unit_start, *unit_count, *unit_rel, *unit_phase;
  int64_t num_units;
  const int32_t *src_sorted, *dst_sorted;
  const __half* h16;
  int64_t dst_lo;
  const float* h_scale;
  const __half* wpack;
  const float* w_inv_scale;
  const float* bias;
  float* ring;          // slots x sb_nodes rows of kD floats
  int slots, sb_nodes;
  int64_t num_local;
  int num_phases;
  int* sync;            // [0] unit counter | [1] ring rows zeroed (warps) | [64 + p] left[p] | [64 + phases + p] epi[p]
  const int32_t* indeg;
  const float* h;       // fp32 master rows (residual)
  const float *ln_w, *ln_b;
  float eps;
  float* out;           // [num_local, kD]
  float* upd;           // optional tap: acc / max(indeg, 1)
  __half* out16;        // optional fp16 shadow of out (+ its scale words)
  float* out16_scale;
  uint32_t flags;
  long long* trace;     // GHF_FUSED_TRACE: cycle accounting of CTA 0 (kTrace instantiation)
  PeerPush push;        // multi-GPU: result rows also go to the tables of the peers that read them (mp.cuh)
};

// Rows [r0, r1) of phase a that this CTA finishes: out = LN(relu(acc / max(indeg, 1) + h)), slot rows back to zero.
// Everything it needs is derived here (nothing but `p` stays live in the tile loop of the caller).
__device__ __forceinline__ void row_epilogue(const FusedParams& p, int a, int warp, int lane) {
  int* const epi = p.sync + 64 + p.num_phases;
  const uint64_t pol_ring = (p.flags & kFlagRingEvictLast) ? policy_evict_last() : policy_evict_normal();
  __threadfence();                                    // the other CTAs' reductions (ordered before left[a])
  const int64_t lo = (int64_t)a * p.sb_nodes;
  const int64_t hi = min(lo + (int64_t)p.sb_nodes, p.num_local);
  const int64_t share = (hi - lo + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = lo + (int64_t)blockIdx.x * share;
  const int64_t r1 = min(hi, r0 + share);
  const int64_t delta = ((int64_t)(a % p.slots) - a) * p.sb_nodes;
  const float4 lw = __ldg(reinterpret_cast<const float4*>(p.ln_w) + lane);
  const float4 lb = __ldg(reinterpret_cast<const float4*>(p.ln_b) + lane);
  // fp16 shadow of the result: |LayerNorm(x)_c| <= sqrt(D-1) |w_c| + |b_c|; the power of two that fits the largest
  // such bound fits every row, and every warp (and every rank) picks the same one from w and b alone
  float s16 = 1.f;
  if (p.out16) {
    const float k = sqrtf((float)(kD - 1));
    float bound = fmaxf(fmaxf(k * fabsf(lw.x) + fabsf(lb.x), k * fabsf(lw.y) + fabsf(lb.y)),
                        fmaxf(k * fabsf(lw.z) + fabsf(lb.z), k * fabsf(lw.w) + fabsf(lb.w)));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, s));
    s16 = f16_scale_for(bound);
    if (a == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
      p.out16_scale[0] = 1.f / s16;
      p.out16_scale[1] = bound;
    }
  }
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kRows = 2;
  if (p.flags & kDbgNoRowWork) {
    __syncwarp();
    if (lane == 0) atomicAdd(&epi[a], 1);
    return;
  }
  for (int k = 0; k < 2 * kRows; ++k) {                  // warm-up of the prefetch distance
    const int64_t rp = r0 + warp + k * kEpiWarps;
    if (rp < r1 && (lane & 7) == 0) prefetch_l2(p.h + (p.dst_lo + rp) * kD + lane * 4);
  }
  // Two rows in flight per warp (32 KiB per SM) is all the registers allow; the residual rows come from HBM, so the
  // rows of the iteration after next are prefetched into L2 (no registers), and the loads below mostly hit L2.
  constexpr int kAhead = 2;
  for (int64_t rb = r0 + warp; rb < r1; rb += kRows * kEpiWarps) {
    float4 av[kRows], hv[kRows];
    int deg[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t rp = rb + (kAhead * kRows + k) * kEpiWarps;
      if (rp < r1 && (lane & 7) == 0) prefetch_l2(p.h + (p.dst_lo + rp) * kD + lane * 4);
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = rb + k * kEpiWarps;
      if (r < r1) {
        av[k] = ld_cg_v4(p.ring + (r + delta) * kD + lane * 4);
        hv[k] = __ldg(reinterpret_cast<const float4*>(p.h + (p.dst_lo + r) * kD) + lane);
        deg[k] = __ldg(p.indeg + r);
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = rb + k * kEpiWarps;
      if (r >= r1) break;
      st_v4_hint(p.ring + (r + delta) * kD + lane * 4, z, pol_ring);   // the slot row is free again
      const float inv = 1.f / (float)max(deg[k], 1);
      const float4 u = make_float4(av[k].x * inv, av[k].y * inv, av[k].z * inv, av[k].w * inv);
      if (p.upd) *reinterpret_cast<float4*>(p.upd + r * kD + lane * 4) = u;
      const float x0 = fmaxf(u.x + hv[k].x, 0.f), x1 = fmaxf(u.y + hv[k].y, 0.f);
      const float x2 = fmaxf(u.z + hv[k].z, 0.f), x3 = fmaxf(u.w + hv[k].w, 0.f);
      float sum = (x0 + x1) + (x2 + x3);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
      const float mean = sum * (1.f / (float)kD);
      float var = (x0 - mean) * (x0 - mean) + (x1 - mean) * (x1 - mean) + (x2 - mean) * (x2 - mean) +
                  (x3 - mean) * (x3 - mean);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
      const float rstd = rsqrtf(var * (1.f / (float)kD) + p.eps);
      const float4 y = make_float4((x0 - mean) * rstd * lw.x + lb.x, (x1 - mean) * rstd * lw.y + lb.y,
                                   (x2 - mean) * rstd * lw.z + lb.z, (x3 - mean) * rstd * lw.w + lb.w);
      __stcs(reinterpret_cast<float4*>(p.out + r * kD) + lane, y);
      if (p.out16) {
        const __half2 p0 = __floats2half2_rn(y.x * s16, y.y * s16), p1 = __floats2half2_rn(y.z * s16, y.w * s16);
        const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        *reinterpret_cast<uint2*>(p.out16 + r * kD + lane * 4) = packed;
        if (p.push.mask)                                 // NVLink stores, while later super-blocks are contracted
          for (int q = 0; q < p.push.world; ++q)
            if (q != p.push.me && p.push.mask[q * p.push.mask_stride + r])
              *reinterpret_cast<uint2*>(p.push.tables[q] + (p.push.table_row0 + r) * kD + lane * 4) = packed;
      }
    }
  }
  __threadfence();                                    // rows written, slot rows zero: visible before epi[a]
  __syncwarp();
  if (lane == 0) atomicAdd(&epi[a], 1);
}

template <int kProdWarps, bool kTrace>
__global__ void __launch_bounds__(threads_for(kProdWarps), 1)
mp_f16_fused_kernel(const __grid_constant__ FusedParams p) {
  constexpr int kRowsPerWarp = kTile / kProdWarps;
  constexpr int kIdsPerLane = kRowsPerWarp / 32;
  static_assert(kRowsPerWarp % 32 == 0, "a producer lane keeps whole ids");
  constexpr int kConsumers = kEpiWarps + kProdWarps + 1 + kLoadWarps;
  const bool tracing = kTrace && blockIdx.x == 0 && (threadIdx.x & 31) == 0;
  auto tick = [&]() -> long long {
    if constexpr (kTrace) return tracing ? clock64() : 0;
    return 0;
  };
  auto tadd = [&](int k, long long v) {
    if constexpr (kTrace)
      if (tracing) atomicAdd(reinterpret_cast<unsigned long long*>(p.trace + k), (unsigned long long)v);
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sQ = sA + kStages * kStageBytes;
  const uint32_t sBar = sQ + kQueue * 16;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  const uint32_t bar2 = sBar + 8u * (2 * kStages);
  auto acc_full = [&](int a) { return bar2 + 8u * a; };
  auto acc_empty = [&](int a) { return bar2 + 32u + 8u * a; };
  auto w_full = [&](int b) { return bar2 + 64u + 8u * b; };
  auto w_empty = [&](int b) { return bar2 + 80u + 8u * b; };
  const uint32_t q_full0 = bar2 + 96u;
  const uint32_t q_empty0 = q_full0 + 8u * kQueue;
  const uint32_t tmem_slot = q_empty0 + 8u * kQueue;
  volatile int4* q_ptr = reinterpret_cast<volatile int4*>(smem_raw + (sQ - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  int* const left = p.sync + 64;
  int* const epi = p.sync + 64 + p.num_phases;
  const int warps_total = (int)gridDim.x * kEpiWarps;      // the target of every grid-wide count

  auto q_acquire = [&](uint32_t idx) -> int4 {
    mbar_wait(q_full0 + 8u * (idx % kQueue), (idx / kQueue) & 1u);
    const volatile int4* q = q_ptr + (idx % kQueue);
    int4 t;
    t.x = q->x; t.y = q->y; t.z = q->z; t.w = q->w;
    return t;
  };
  auto q_release = [&](uint32_t idx) {
    __syncwarp();
    if (lane == 0) mbar_arrive(q_empty0 + 8u * (idx % kQueue));
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 32 * kProdWarps);
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < kAccBufs; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), kEpiWarps);
    }
    for (int b = 0; b < kWBufs; ++b) {
      mbar_init(w_full(b), 32 * kLoadWarps);
      mbar_init(w_empty(b), 1);
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
    // ------------------------------------------------------------------ epilogue warps
    const int grp = warp >> 2, q = warp & 3;
    const int col = 32 * q + lane;
    const float h_inv = p.h_scale[0];
    const int e0 = 32 * grp;

    // (0) the ring starts zeroed: this launch's share of slots x sbr rows, one row per warp step
    {
      const uint64_t pol_ring = (p.flags & kFlagRingEvictLast) ? policy_evict_last() : policy_evict_normal();
      const int64_t sbr = p.num_phases > 1 ? (int64_t)p.sb_nodes : p.num_local;   // rows per ring slot
      const int64_t total = (int64_t)p.slots * sbr;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t r = (int64_t)blockIdx.x * kEpiWarps + warp; r < total; r += warps_total)
        st_v4_hint(p.ring + r * kD + lane * 4, z, pol_ring);
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(p.sync + 1, 1);
    }
    struct TileRegs { int dst; float bias_n, w_inv; };
    auto fetch = [&](const int4& t) -> TileRegs {
      TileRegs x{-1, 0.f, 1.f};
      if ((uint32_t)t.w & kDescNotTile) return x;
      if (e0 + lane < t.y) x.dst = p.dst_sorted[t.x + e0 + lane];
      x.bias_n = p.bias ? p.bias[(int64_t)t.z * kD + col] : 0.f;
      x.w_inv = p.w_inv_scale[t.z];
      return x;
    };
    int4 cur = q_acquire(0);
    TileRegs cr = fetch(cur);
    uint32_t tile_it = 0, it = 0;
    while (!((uint32_t)cur.w & kDescEnd)) {
      if ((uint32_t)cur.w & (kDescLeave | kDescEpi)) {
        // grid-level work between tiles; nothing of the tile pipeline is live here (no look-ahead)
        const long long t0 = tick();
        if ((uint32_t)cur.w & kDescLeave) {
          __threadfence();                                // every reduction this warp issued so far is performed
          __syncwarp();
          if (lane == 0)
            for (int a = cur.x; a < cur.y; ++a) atomicAdd(&left[a], 1);
          if (warp == 0) tadd(3, tick() - t0);
        } else {
          row_epilogue(p, cur.x, warp, lane);
          if (warp == 0) { tadd(2, tick() - t0); tadd(7, 1); }
        }
        q_release(it);
        ++it;
        cur = q_acquire(it);
        cr = fetch(cur);
        continue;
      }
      const int4 nxt = q_acquire(it + 1);
      const TileRegs nr = fetch(nxt);                     // in flight while this tile's reductions are issued
      const int a = (int)(tile_it % kAccBufs);
      const long long tw = tick();
      mbar_wait(acc_full(a), (tile_it / kAccBufs) & 1u);
      if (warp == 0) { tadd(6, tick() - tw); tadd(5, 1); }
      tc_fence_after();
      if (e0 < cur.y && !(p.flags & kDbgNoRed)) {
        const int phase = (int)((uint32_t)cur.w >> 8);
        float* acc_col = p.ring + ((int64_t)(phase % p.slots) - phase) * p.sb_nodes * kD + col;
        const float inv = cr.w_inv * h_inv;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTile + e0), r);
        tmem_ld_wait();
        // runs of equal destinations (adjacent in the sorted order) leave as ONE reduction
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(a));
      ++tile_it;
      q_release(it);
      ++it;
      cur = nxt;
      cr = nr;
    }
  } else if (warp >= kWarpProd) {
    // ------------------------------------------------------------------ row-gather producers
    const int pw = warp - kWarpProd;
    const int l16 = lane & 15, hi = lane >> 4;
    const uint64_t pol_src = (p.flags & kFlagSrcEvictFirst) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_dst = (p.flags & kFlagDstEvictLast) ? policy_evict_last() : policy_evict_normal();
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(p.h16) + l16 * 16;
    // lane l keeps the ids of rows l, l + 32, ... of the warp's share (kIdsPerLane of them)
    struct Ids { int src[kIdsPerLane], dst[kIdsPerLane]; };
    auto ids_of = [&](const int4& t) -> Ids {
      Ids x;
#pragma unroll
      for (int j = 0; j < kIdsPerLane; ++j) {
        const int row = kRowsPerWarp * pw + 32 * j + lane;
        const bool ok = !((uint32_t)t.w & kDescNotTile) && row < t.y;
        x.src[j] = ok ? p.src_sorted[t.x + row] : -1;
        x.dst[j] = ok ? p.dst_sorted[t.x + row] : -1;
      }
      return x;
    };
    int stage = 0;
    uint32_t phase = 0;
    int4 cur = q_acquire(0);
    Ids ids = ids_of(cur);
    for (uint32_t it = 0; !((uint32_t)cur.w & kDescEnd); ++it) {
      const int4 nxt = q_acquire(it + 1);
      const Ids ids_nxt = ids_of(nxt);
      if (!((uint32_t)cur.w & kDescNotTile)) {
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          mbar_wait(empty(stage), phase ^ 1u);
          const uint32_t base = sA + stage * kStageBytes + (l16 >> 3) * kSub;
          const uint64_t pol = s ? pol_dst : pol_src;
          const uint8_t* table = s ? hb + p.dst_lo * kRowBytes : hb;
#pragma unroll
          for (int i = 0; i < kRowsPerWarp / 2; ++i) {
            const int rl = 2 * i + hi;                     // row within the warp's share; 2 i / 32 is compile-time
            const int mine = s ? ids.dst[(2 * i) / 32] : ids.src[(2 * i) / 32];
            const int idx = __shfl_sync(0xffffffffu, mine, rl & 31);
            const int row = kRowsPerWarp * pw + rl;
            const uint32_t to = base + row * 128 + (((l16 & 7) ^ (row & 7)) << 4);
            if (idx >= 0) cp_async_16_hint(to, table + (int64_t)idx * kRowBytes, pol);
          }
          cp_async_arrive_noinc(full(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      q_release(it);
      cur = nxt;
      ids = ids_nxt;
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0, wph = 0, tile_it = 0;
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      const uint32_t tf = (uint32_t)t.w;
      q_release(it);
      if (tf & kDescEnd) break;
      if (tf & kDescNotTile) continue;
      const int a = (int)(tile_it % kAccBufs);
      const int wb = (tf & kTileWbuf) ? 1 : 0;
      mbar_wait(acc_empty(a), ((tile_it / kAccBufs) & 1u) ^ 1u);
      if (tf & kTileFirst) {
        mbar_wait(w_full(wb), (wph >> wb) & 1u);
        wph ^= 1u << wb;
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(a * kTile);
      const uint32_t w_tmem = tmem_base + kWCol + (uint32_t)(wb * kD);
      const bool leader = elect_one();
      bool first = true;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        mbar_wait(full(stage), phase);
        fence_proxy_async();
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
        if (tf & kTileLast) umma_commit(w_empty(wb));
      }
      __syncwarp();
      ++tile_it;
    }
  } else if (warp == kWarpSched) {
    // ------------------------------------------------------------------ scheduler
    uint32_t qi = 0, wb = 0;
    auto publish = [&](int x, int y, int z, uint32_t w) {
      if (lane == 0) {
        mbar_wait(q_empty0 + 8u * (qi % kQueue), ((qi / kQueue) & 1u) ^ 1u);
        volatile int4* q = q_ptr + (qi % kQueue);
        q->x = x; q->y = y; q->z = z; q->w = (int)w;
        mbar_arrive(q_full0 + 8u * (qi % kQueue));
      }
      ++qi;
    };
    auto draw = [&]() -> int64_t {
      int v = 0;
      if (lane == 0) v = atomicAdd(p.sync, 1);
      return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    auto complete = [&](const int* word) -> bool {         // non-blocking, warp-uniform
      int v = 0;
      if (lane == 0) v = fuse::ld_acquire(word);
      return __shfl_sync(0xffffffffu, v, 0) >= warps_total;
    };
    auto wait_for = [&](const int* word, int slot) {
      const long long t0 = tick();
      if (lane == 0) fuse::spin_until_at_least(word, warps_total);
      __syncwarp();
      tadd(slot, tick() - t0);
    };
    const long long t_begin = tick();
    int left_upto = 0;    // LEAVE published for phases < left_upto
    int epi_upto = 0;     // EPI published for phases < epi_upto
    int cur_phase = -1;
    wait_for(p.sync + 1, 8);                                  // every slot row is zero
    int64_t u = draw();
    while (u < p.num_units) {
      const int start = p.unit_start[u], count = p.unit_count[u], rel = p.unit_rel[u], ph = p.unit_phase[u];
      const int64_t u_next = draw();
      if (ph > cur_phase) {
        if (ph > left_upto) {                              // this CTA issues no more reductions into phases < ph
          publish(left_upto, ph, 0, kDescLeave);
          left_upto = ph;
        }
        // slot ph % slots must be handed back by EVERY CTA (epilogue of phase ph - slots) before it is reduced into;
        // our own share of every phase up to that one goes first
        for (; epi_upto <= ph - p.slots; ++epi_upto) {
          wait_for(&left[epi_upto], 0);
          publish(epi_upto, 0, 0, kDescEpi);
        }
        if (ph >= p.slots) wait_for(&epi[ph - p.slots], 1);
        cur_phase = ph;
      }
      if (epi_upto < left_upto && complete(&left[epi_upto])) {   // an earlier phase is ready for its epilogue
        publish(epi_upto, 0, 0, kDescEpi);
        ++epi_upto;
      }
      for (int t0 = 0; t0 < count; t0 += kTile) {
        const uint32_t tf = (t0 == 0 ? kTileFirst : 0u) | (t0 + kTile >= count ? kTileLast : 0u) |
                            (wb ? kTileWbuf : 0u) | ((uint32_t)ph << 8);
        publish(start + t0, min(kTile, count - t0), rel, tf);
      }
      wb = (wb + 1u) % kWBufs;
      u = u_next;
    }
    if (left_upto < p.num_phases) publish(left_upto, p.num_phases, 0, kDescLeave);
    for (; epi_upto < p.num_phases; ++epi_upto) {          // rows without in-edges get their epilogue too
      wait_for(&left[epi_upto], 9);
      publish(epi_upto, 0, 0, kDescEpi);
    }
    tadd(4, tick() - t_begin);
    publish(0, 0, 0, kDescEnd);
    publish(0, 0, 0, kDescEnd);                            // producers and epilogue look one descriptor ahead
  } else if (warp >= kWarpLoad && warp < kWarpMma) {
    // ------------------------------------------------------------------ weight loaders: Wt_r -> TMEM
    const int quarter = warp & 3;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + kWCol;
    const uint64_t pol_w = (p.flags & kFlagWEvictLast) ? policy_evict_last() : policy_evict_normal();
    uint32_t eph = 0;
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      const uint32_t tf = (uint32_t)t.w;
      q_release(it);
      if (tf & kDescEnd) break;
      if ((tf & kDescNotTile) || !(tf & kTileFirst)) continue;
      const int wb = (tf & kTileWbuf) ? 1 : 0;
      const uint8_t* img = reinterpret_cast<const uint8_t*>(p.wpack) + (int64_t)t.z * kImageBytes +
                           quarter * (8 * 32 * 16) + lane * 16;
      uint32_t r[32];
      auto fetch = [&](int piece) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = ldg_v4_hint(img + piece * (4 * 8 * 32 * 16) + j * (32 * 16), pol_w);
          r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
      };
      fetch(0);
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

}  // namespace

constexpr int kMaxSlots = 4;

int mp_f16_fused_slots(const ghf_graph* g) {
  const char* env = getenv("GHF_FUSED_SLOTS");
  int slots = env ? atoi(env) : 3;
  slots = slots < 2 ? 2 : (slots > kMaxSlots ? kMaxSlots : slots);
  return g->num_phases < slots ? (int)g->num_phases : slots;
}

// rows (hidden_dim floats each) to reserve for the accumulator ring, whatever GHF_FUSED_SLOTS says at launch time
int64_t mp_f16_fused_ring_rows(const ghf_graph* g) {
  if (g->num_phases <= 1) return g->num_local;
  return (g->num_phases < kMaxSlots ? g->num_phases : (int64_t)kMaxSlots) * g->sb_nodes;
}

// Off unless GHF_MP_FUSED=1: measured at c3 (profiles/r02_fused_layer_kernel.txt) the fused kernel takes 3.50 ms per
// layer against 3.13 ms for contraction + separate epilogue.  Fusion removes HBM traffic (the 1.28 GB accumulator
// never leaves L2) but not one byte of SM <-> L2 traffic, and that interface - 32 B/clk/SM for both directions
// together, ~8.3 TB/s chip-wide (tools/scatter_paths.cu) - is what the layer is bound by; inside one kernel the row
// epilogue's stores queue behind the reductions on the same port instead of overlapping with anything.
bool mp_f16_fused_enabled(const ghf_graph* g) {
  const char* env = getenv("GHF_MP_FUSED");
  return g->hidden_dim == kD && g->num_units > 0 && env && env[0] == '1';
}

int mp_f16_fused_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias,
                        const void* pack_scratch, float* ring, int* sync_words, const float* h, const float* ln_w,
                        const float* ln_b, float eps, float* out, float* upd, void* out16, float* out16_scale,
                        cudaStream_t stream, const PeerPush* push) {
  GHF_REQUIRE(h16_scale != nullptr, "mp_f16_fused: the fp16 shadow needs its scale words");
  GHF_REQUIRE(g->hidden_dim == kD && g->num_units > 0, "mp_f16_fused: hidden_dim must be %d and the graph non-empty", kD);
  GHF_REQUIRE(g->unit_edges % kTile == 0, "mp_f16_fused: unit_edges=%d must be a multiple of %d", g->unit_edges, kTile);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h16) | reinterpret_cast<uintptr_t>(ring) | reinterpret_cast<uintptr_t>(h) |
               reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(upd) | reinterpret_cast<uintptr_t>(out16) |
               reinterpret_cast<uintptr_t>(ln_w) | reinterpret_cast<uintptr_t>(ln_b) |
               reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_f16_fused: buffers must be 16-byte aligned");
  GHF_REQUIRE(g->num_phases < (1 << 23), "mp_f16_fused: too many super-blocks");
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_fused_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_fused_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_fused_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  FusedParams p{};
  p.unit_start = g->unit_start; p.unit_count = g->unit_count; p.unit_rel = g->unit_rel; p.unit_phase = g->unit_phase;
  p.num_units = g->num_units; p.src_sorted = g->src_sorted; p.dst_sorted = g->dst_sorted;
  p.h16 = reinterpret_cast<const __half*>(h16); p.dst_lo = g->dst_lo; p.h_scale = h16_scale;
  p.wpack = reinterpret_cast<const __half*>(pack_scratch);
  p.w_inv_scale = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pack_scratch) +
                                                 align_up((int64_t)g->num_rel * kImageBytes, 256));
  p.bias = bias; p.ring = ring; p.slots = mp_f16_fused_slots(g); p.sb_nodes = g->sb_nodes;
  p.num_local = g->num_local; p.num_phases = (int)g->num_phases; p.sync = sync_words; p.indeg = g->indeg; p.h = h;
  p.ln_w = ln_w; p.ln_b = ln_b; p.eps = eps; p.out = out; p.upd = upd; p.out16 = reinterpret_cast<__half*>(out16);
  p.out16_scale = out16_scale;
  if (push) {
    GHF_REQUIRE(out16 != nullptr, "mp_f16_fused: the peer push needs the fp16 output");
    p.push = *push;
  }
  const char* fenv = getenv("GHF_FUSED_FLAGS");
  p.flags = fenv ? (uint32_t)atoi(fenv) : kDefaultFlags;
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  GHF_CUDA(cudaMemsetAsync(sync_words, 0, mp_f16_sync_bytes(g), stream));
  const char* penv = getenv("GHF_FUSED_PROD");
  const int prod = penv ? atoi(penv) : 2;
  TempBuf trace_buf;
  if (getenv("GHF_FUSED_TRACE")) {
    GHF_CUDA(trace_buf.alloc(16 * sizeof(long long), stream));
    GHF_CUDA(cudaMemsetAsync(trace_buf.p, 0, 16 * sizeof(long long), stream));
    p.trace = trace_buf.as<long long>();
    mp_f16_fused_kernel<2, true><<<(unsigned)grid, threads_for(2), kSmem, stream>>>(p);
  } else if (prod == 4) {
    mp_f16_fused_kernel<4, false><<<(unsigned)grid, threads_for(4), kSmem, stream>>>(p);
  } else {
    mp_f16_fused_kernel<2, false><<<(unsigned)grid, threads_for(2), kSmem, stream>>>(p);
  }
  GHF_LAUNCH_CHECK();
  if (p.trace) {   // synchronises: diagnostics only
    long long t[16];
    GHF_CUDA(cudaMemcpyAsync(t, p.trace, sizeof(t), cudaMemcpyDeviceToHost, stream));
    GHF_CUDA(cudaStreamSynchronize(stream));
    fprintf(stderr,
            "mp_f16_fused trace (CTA 0, cycles): total %lld | scheduler blocked: ring-zero %lld, left(forced) %lld, "
            "epi(slot) %lld, tail left %lld | epilogue warp 0: %lld tiles, acc_full wait %lld, row epilogue %lld in %lld "
            "calls, leave fences %lld\n",
            t[4], t[8], t[0], t[1], t[9], t[5], t[6], t[2], t[7], t[3]);
  }
  return 0;
}

}  // namespace ghf
