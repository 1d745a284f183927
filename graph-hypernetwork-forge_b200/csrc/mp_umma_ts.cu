// mp_umma_ts.cu — message-passing contraction for hidden_dim 128 with the generated weights resident in
// TENSOR MEMORY, so that all of shared memory is one deep gather ring.
//
// Why: the gather of h rows is latency-bound, its bandwidth is (bytes in flight per SM) / (loaded HBM latency).
// tools/gather_bw.cu on B200: 80 KB in flight per SM -> 4.2 TB/s, 192 KB -> saturated.  With the 128 KB operand
// image [W_msg; W_self][r] in shared memory only 5 x 16 KB stages fit (mp_umma.cu); here it lives in TMEM and
// 13 stages (208 KB) fit.
//
// How: compute the transposed tile.  tcgen05.mma takes its A operand from tensor memory, so
//     Dt[128 out-cols, 128 edges] = Wt_r[128 out-cols, 256] * Ht[256, 128 edges]
//   A = Wt_r  : TMEM columns [256, 512)  (lane = output column n, column = k; written once per unit with tcgen05.st)
//   B = Ht    : the gathered rows exactly as before (edge-major, K contiguous, 128B swizzle) = a K-major B operand
//   D = Dt    : TMEM columns [0,128) / [128,256) (double buffered): lane = output column, column = edge
// The epilogue transposes 32-edge blocks through shared memory so that each warp still issues full 512 B
// red.global.add.v4.f32 rows into out[dst_e] (+ bias[r]).
//
// Warp roles (576 threads, 1 CTA / SM):
//   0-7 epilogue (2 groups) | 8-11 row-gather producers | 12 MMA issuer + TMEM allocator | 13 unit scheduler |
//   14-17 weight loaders
#include <cstdlib>

#include "common.cuh"
#include "mp.cuh"
#include "umma.cuh"

namespace ghf {
namespace {

using namespace ptx;

constexpr int kD = 128;
constexpr int kTile = 128;                 // edges per MMA tile (the N of the transposed product)
constexpr int kChunks = 2 * kD / 32;       // 8 K-chunks of 32 tf32 (128 B swizzle rows)
constexpr int kHalf = kChunks / 2;
constexpr int kSub = kTile * 128;          // one K-chunk of a tile: 128 rows x 128 B = 16 KiB
constexpr int kChunksPerStage = 2;         // a pipeline stage = 2 K-chunks: the MMA thread pays one mbarrier wait and
                                           // one commit per 8 MMAs (512 tensor cycles); per-chunk stages left the
                                           // issuing thread, not the tensor pipe, as the bottleneck
constexpr int kAStage = kChunksPerStage * kSub;  // 32 KiB
constexpr int kStages = 6;                 // 192 KiB of gathers in flight
constexpr int kEpiGroups = 2;              // 2 x 4 epilogue warps: the row scatter is latency-bound per warp
constexpr int kStaging = 32 * kD * 4;      // per group: one 32-edge x 128-column block, 16 KiB
constexpr int kBarBytes = 512;
constexpr int kSmem = 1024 + kStages * kAStage + kEpiGroups * kStaging + kBarBytes;
constexpr int kWarpProd = 4 * kEpiGroups;  // first producer warp
constexpr int kWarpMma = kWarpProd + 4, kWarpSched = kWarpMma + 1, kWarpLoad = kWarpSched + 1;
constexpr int kThreads = 32 * (kWarpLoad + 4);
constexpr int kQueueConsumers = 4 * kEpiGroups + 4 + 1 + 4;
constexpr int kQueue = 4;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kWCol = 256;            // first TMEM column of Wt
constexpr int kWBytes = kD * 2 * kD * 4;   // one relation's Wt image: [128 n][256 k] fp32 (tf32-rounded)

constexpr uint32_t kFlagSrcEvictFirst = 1u, kFlagDstEvictLast = 2u, kFlagRedEvictLast = 4u;
constexpr uint32_t kDefaultFlags = kFlagSrcEvictFirst | kFlagDstEvictLast | kFlagRedEvictLast;
// timing experiments only (results are wrong): drop the reductions / the gathers / the whole tile epilogue
constexpr uint32_t kDbgNoRed = 32u, kDbgNoGather = 64u, kDbgNoEpilogue = 128u, kDbgNoWeights = 256u, kDbgNoRing = 512u;

// Wt image of relation r, laid out for the weight loaders: element (n, k) of Wt_r (k < 128 -> W_msg[r][k][n],
// else W_self[r][k-128][n]) lives at float offset
//     piece c = k/32 | quarter q = n/32 | 16-byte group j = (k%32)/4 | lane l = n%32 | k%4
// so that load j of piece c by warp-quarter q is one contiguous 512 B line set (lane l reads 16 B at
// ((c*4 + q)*8 + j)*32 + l) and the 8 loads of a lane give its 32 consecutive k values.
__device__ __forceinline__ int wt_offset(int n, int k) {
  return ((((k >> 5) * 4 + (n >> 5)) * 8 + ((k & 31) >> 2)) * 32 + (n & 31)) * 4 + (k & 3);
}

__global__ void pack_wt_kernel(const float* __restrict__ W_msg, const float* __restrict__ W_self, int R,
                               float* __restrict__ pack) {
  // one thread per (r, k, n): reads contiguous in n; writes are 16-byte-group scattered (4 B each)
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)R * 2 * kD * kD) return;
  const int n = (int)(idx % kD);
  const int k = (int)((idx / kD) % (2 * kD));
  const int64_t r = idx / (2 * kD * kD);
  const float v = k < kD ? W_msg[(r * kD + k) * kD + n] : W_self[(r * kD + (k - kD)) * kD + n];
  pack[r * (2 * kD * kD) + wt_offset(n, k)] = to_tf32_rna(v);
}

__global__ void __launch_bounds__(kThreads, 1)
mp_umma_ts_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
                  const int32_t* __restrict__ unit_rel, int64_t num_units, const int32_t* __restrict__ src_sorted,
                  const int32_t* __restrict__ dst_sorted, const float* __restrict__ h, int64_t dst_lo,
                  const float* __restrict__ wpack, const float* __restrict__ bias, float* __restrict__ acc,
                  int* __restrict__ unit_counter, uint32_t flags) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sStg = sA + kStages * kAStage;
  const uint32_t sBar = sStg + kEpiGroups * kStaging;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  const uint32_t bar2 = sBar + 8u * (2 * kStages);
  auto acc_full = [&](int a) { return bar2 + 8u * a; };
  auto acc_empty = [&](int a) { return bar2 + 16u + 8u * a; };
  auto w_full = [&](int c) { return bar2 + 32u + 8u * c; };   // one per K-chunk of Wt
  const uint32_t w_empty = bar2 + 32u + 8u * kChunks;
  const uint32_t q_full0 = w_empty + 8u;
  const uint32_t q_empty0 = q_full0 + 8u * kQueue;
  const uint32_t q_slots = q_empty0 + 8u * kQueue;
  const uint32_t tmem_slot = q_slots + 4u * kQueue;
  volatile int32_t* q_slot_ptr = reinterpret_cast<volatile int32_t*>(smem_raw + (q_slots - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // unit queue: scheduler -> 4 epilogue warps + 4 producer warps + MMA thread + 4 weight-loader warps
  int q_idx = 0;
  uint32_t q_phase = 0;
  auto next_unit = [&](bool whole_warp) -> int {
    mbar_wait(q_full0 + 8u * q_idx, q_phase);
    const int u = q_slot_ptr[q_idx];
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(q_empty0 + 8u * q_idx);
    if (++q_idx == kQueue) { q_idx = 0; q_phase ^= 1u; }
    return u;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full(s), 128);
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 128 * kEpiGroups);
    }
    for (int c = 0; c < kChunks; ++c) mbar_init(w_full(c), 128);  // every weight-loader thread
    mbar_init(w_empty, 1);                                         // tcgen05.commit
    for (int q = 0; q < kQueue; ++q) {
      mbar_init(q_full0 + 8u * q, 1);
      mbar_init(q_empty0 + 8u * q, kQueueConsumers);
    }
    mbar_fence_init();
  }
  if (warp == kWarpMma) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < kWarpProd) {
    // ------------------------------------------------------------------ epilogue: Dt -> rows -> red.v4
    // Two groups of 4 warps; group g handles edges [64g, 64g+64) of every tile in two 32-edge blocks.  Within a
    // group, warp q reads TMEM lanes 32q..32q+31 (output columns), the block is transposed through the group's
    // 16 KiB staging buffer, and each warp emits 8 of the 32 edges as full 512 B rows (one red.v4 per lane).
    const int grp = warp >> 2, q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem_raw + (sStg - raw) + grp * kStaging);
    const float4* stg4 = reinterpret_cast<const float4*>(stg);
    const uint64_t pol_red = (flags & kFlagRedEvictLast) ? policy_evict_last() : policy_evict_normal();
    uint32_t it = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      const float4 b4 = *reinterpret_cast<const float4*>(bias + (int64_t)unit_rel[u] * kD + 4 * lane);
      for (int t0 = 0; t0 < count; t0 += kTile, ++it) {
        const int a = it & 1;
        const int rows = min(kTile, count - t0);
        // lane L keeps the destination of edge 64*grp + 32*(L/16) + 8*q + (L%8) for L%16 < 8
        const int my_edge = 64 * grp + 32 * (lane >> 4) + 8 * q + (lane & 7);
        const int my_dst = my_edge < rows ? dst_sorted[start + t0 + my_edge] : -1;
        mbar_wait(acc_full(a), (it >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int bb = 0; bb < ((flags & kDbgNoEpilogue) ? 0 : 2); ++bb) {
          const int e0 = 64 * grp + 32 * bb;  // first edge of this block
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * kTile + e0), r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) stg[e * kD + q * 32 + lane] = __uint_as_float(r[e]);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = 8 * q + i;
            const int dsti = __shfl_sync(0xffffffffu, my_dst, 16 * bb + i);
            float4 v = stg4[e * (kD / 4) + lane];
            if (dsti >= 0 && !(flags & kDbgNoRed)) {
              v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
              red_add_v4_hint(acc + (int64_t)dsti * kD + 4 * lane, v, pol_red);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        }
        tc_fence_before();
        mbar_arrive(acc_empty(a));
      }
    }
  } else if (warp < kWarpMma) {
    // ------------------------------------------------------------------ row-gather producers
    const int pw = warp - kWarpProd;
    const int cj = lane & 7;
    const uint64_t pol_src = (flags & kFlagSrcEvictFirst) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_dst = (flags & kFlagDstEvictLast) ? policy_evict_last() : policy_evict_normal();
    int stage = 0;
    uint32_t phase = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      for (int t0 = 0; t0 < count; t0 += kTile) {
        const int rows = min(kTile, count - t0);
        const int my_row = pw * 32 + lane;
        const bool ok = my_row < rows;
        const int64_t my_src = ok ? (int64_t)src_sorted[start + t0 + my_row] : -1;
        const int64_t my_dst = ok ? dst_lo + dst_sorted[start + t0 + my_row] : -1;
#pragma unroll 1
        for (int c0 = 0; c0 < ((flags & kDbgNoRing) ? 0 : kChunks); c0 += kChunksPerStage) {
          mbar_wait(empty(stage), phase ^ 1u);
#pragma unroll
          for (int cs = 0; cs < kChunksPerStage; ++cs) {
            const int c = c0 + cs;
            const bool from_src = c < kHalf;
            const int64_t mine = from_src ? my_src : my_dst;
            const int col = (from_src ? c : c - kHalf) * 32 + 4 * cj;
            const uint32_t dst_base = sA + stage * kAStage + cs * kSub;
            const uint64_t pol = from_src ? pol_src : pol_dst;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const int64_t idx = __shfl_sync(0xffffffffu, mine, rr);
              const int row = pw * 32 + rr;
              const uint32_t to = dst_base + row * 128 + ((cj ^ (row & 7)) << 4);
              if (idx >= 0 && !(flags & kDbgNoGather)) cp_async_16_hint(to, h + idx * kD + col, pol);
            }
          }
          cp_async_arrive_noinc(full(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop (uniform control flow: barrier addresses, descriptors and TMEM addresses
    // stay in uniform registers); one elected lane issues the tcgen05 instructions.
    {
      // M = 128 output columns (A = Wt in TMEM), N = 128 edges (B = gathered rows, K-major)
      constexpr uint32_t idesc = umma_idesc_tf32(kD, kTile);
      int stage = 0;
      uint32_t phase = 0, wphase = 0, it = 0;
      for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
        const int count = unit_count[u];
        bool first_tile = true;
        for (int t0 = 0; t0 < count; t0 += kTile, ++it) {
          const int a = it & 1;
          mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * kTile);
#pragma unroll 1
          for (int c0 = 0; c0 < kChunks; c0 += kChunksPerStage) {
            if (first_tile) {  // this unit's Wt columns for these K-chunks have landed in TMEM
#pragma unroll
              for (int cs = 0; cs < kChunksPerStage; ++cs) mbar_wait(w_full(c0 + cs), wphase);
            }
            if (!(flags & kDbgNoRing)) mbar_wait(full(stage), phase);
            tc_fence_after();
            const uint32_t stage_addr = sA + stage * kAStage;
            if (elect_one()) {
#pragma unroll
              for (int cs = 0; cs < kChunksPerStage; ++cs) {
                const uint64_t bdesc = umma_desc_k128(stage_addr + cs * kSub);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  umma_tf32_ts(d_tmem, tmem_base + kWCol + (uint32_t)(32 * (c0 + cs) + 8 * j), bdesc + 2 * j, idesc,
                               (uint32_t)((c0 + cs) | j));
              }
              if (!(flags & kDbgNoRing)) umma_commit(empty(stage));
              if (c0 + kChunksPerStage == kChunks) umma_commit(acc_full(a));
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          first_tile = false;
        }
        if (elect_one()) umma_commit(w_empty);  // every MMA that reads this unit's Wt has completed
        __syncwarp();
        wphase ^= 1u;
      }
    }
  } else if (warp == kWarpSched) {
    // ------------------------------------------------------------------ unit scheduler
    if (lane == 0) {
      uint32_t sphase = 0;
      int sq = 0;
      for (;;) {
        mbar_wait(q_empty0 + 8u * sq, sphase ^ 1u);
        const int64_t u = atomicAdd(unit_counter, 1);
        const bool done = u >= num_units;
        q_slot_ptr[sq] = done ? -1 : (int)u;
        mbar_arrive(q_full0 + 8u * sq);
        if (++sq == kQueue) { sq = 0; sphase ^= 1u; }
        if (done) break;
      }
    }
  } else {
    // ------------------------------------------------------------------ weight loaders: Wt_r -> TMEM
    // thread = output column n = 32*(warp%4) + lane = TMEM lane; 256 K-columns in 8 pieces of 32
    const int quarter = warp & 3;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + kWCol;
    uint32_t wphase = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const float* img = wpack + (int64_t)unit_rel[u] * (2 * kD * kD) + quarter * (8 * 32 * 4) + lane * 4;
      uint32_t r[2][32];
      auto fetch = [&](int slot, int piece) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = ldg_nc_v4(img + piece * (4 * 8 * 32 * 4) + j * (32 * 4));
          r[slot][4 * j] = v.x; r[slot][4 * j + 1] = v.y; r[slot][4 * j + 2] = v.z; r[slot][4 * j + 3] = v.w;
        }
      };
      if (flags & kDbgNoWeights) {  // timing experiment: pretend the weights are already in TMEM
        mbar_wait(w_empty, wphase ^ 1u);
        for (int c = 0; c < kChunks; ++c) mbar_arrive(w_full(c));
        wphase ^= 1u;
        continue;
      }
      // pull the whole 128 KiB image into L2 now (the previous unit is still computing), so that the loads
      // issued after w_empty are L2 hits; the first two pieces travel to registers right away
#pragma unroll
      for (int c = 2; c < kChunks; ++c) prefetch_l2(img + c * (4 * 8 * 32 * 4) + (lane & 7) * (32 * 4));
      fetch(0, 0); fetch(1, 1);
      mbar_wait(w_empty, wphase ^ 1u);
      tc_fence_after();
#pragma unroll
      for (int base = 0; base < kChunks; base += 2) {
        tmem_st_32x32(t_row + 32u * base, r[0]);
        tmem_st_32x32(t_row + 32u * (base + 1), r[1]);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(w_full(base));
        mbar_arrive(w_full(base + 1));
        if (base + 2 < kChunks) { fetch(0, base + 2); fetch(1, base + 3); }
      }
      wphase ^= 1u;
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

int configure() {
  static bool done = false;
  if (!done) {
    GHF_CUDA(cudaFuncSetAttribute(mp_umma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    done = true;
  }
  return 0;
}

}  // namespace

bool mp_ts_supported(int d) { return d == kD; }

int mp_ts_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream) {
  GHF_REQUIRE(g->hidden_dim == kD, "mp_ts: hidden_dim must be %d", kD);
  const int64_t total = (int64_t)g->num_rel * 2 * kD * kD;
  pack_wt_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(W_msg, W_self, g->num_rel,
                                                                reinterpret_cast<float*>(pack_scratch));
  GHF_LAUNCH_CHECK();
  return 0;
}

int mp_ts_launch(const ghf_graph* g, const float* h, const float* bias, float* acc, const void* pack_scratch,
                 int* unit_counter, cudaStream_t stream) {
  GHF_REQUIRE(g->unit_edges % kTile == 0, "mp_ts: unit_edges=%d must be a multiple of %d", g->unit_edges, kTile);
  if (int rc = configure()) return rc;
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  mp_umma_ts_kernel<<<(unsigned)grid, kThreads, kSmem, stream>>>(
      g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted, h, g->dst_lo,
      reinterpret_cast<const float*>(pack_scratch), bias, acc, unit_counter, env_flags());
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace ghf
