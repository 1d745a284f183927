// mp_wgrad_f16.cu — gradients of the generated relation matrices for hidden_dim 128 on tcgen05 (GHF_PREC_F16).
//
//   g_W_msg[r]  = sum over the edges e of relation r of  h_src(e)^T  g_acc_dst(e)        [128 x 128]
//   g_W_self[r] = sum over the edges e of relation r of  h_dst(e)^T  g_acc_dst(e)        [128 x 128]
//
// The contraction index is the EDGE: per unit (<= 1024 edges of one relation, graph.cu) two accumulators in tensor
// memory receive  D += A B  with A = [gathered feature rows]^T (M = feature, K = edge) and B = gathered gradient
// rows (N = output column, K = edge).  A gathered row is contiguous along M (resp. N), i.e. both operands are
// "MN-major": the very same shared-memory image the forward uses as a K-major operand - 64-feature chunks of
// [edge][128 B] with the 128-byte swizzle - read through MN-major descriptors (LBO = chunk stride, SBO = 1024 B per
// 8 edges; cute/atom/mma_traits_sm100.hpp, "make_umma_desc<Major::MN>").  Operands travel as fp16 (the shadow of h
// the forward already has, and g16 = fp16(g_acc * 2^k)), products accumulate in fp32, the epilogue undoes the scales.
//
// Stage = 64 edges x {h16[src], h16[dst], g16[dst]} = 48 KiB, ring of 4; the last stage of a unit is zero-filled to
// a multiple of 16 edges (one MMA K step), so stale shared memory never enters a product.
// Warp roles (448 threads, 1 CTA / SM, persistent, dynamic unit scheduler):
//   0-3   epilogue: TMEM -> 32x32 transposes through shared memory -> red.global.add.f32, one 128 B line per
//         instruction (a unit's two matrices are 128 KiB of reductions into g_W[r])
//   4     MMA issuer + TMEM allocator     5  scheduler (unit descriptors through a shared-memory queue)
//   6-13  row-gather producers (cp.async, zero-fill for the tail rows)
//   14    bias-gradient warp: column sums of the gathered gradient rows of every stage, read from shared memory
//         next to the tensor core (g_bias[r] = sum over the unit's edges of g_acc[dst]), one atomic per column and unit
// TMEM: 2 buffers x (msg 128 + self 128 columns): the epilogue of unit i overlaps the products of unit i + 1.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "mp.cuh"
#include "umma.cuh"

namespace ghf {
namespace {

using namespace ptx;

constexpr int kD = 128;
constexpr int kT = 64;                       // edges per stage
constexpr int kChunk = kT * 128;             // [64 edges][64 halfs]: 8 KiB
constexpr int kSet = 2 * kChunk;             // one gathered row set (128 halfs per edge)
constexpr int kStageBytes = 3 * kSet;        // h16[src] | h16[dst] | g16[dst]
constexpr int kStages = 4;
constexpr int kQueue = 8;
constexpr int kEpiWarps = 4, kProdWarps = 8;
constexpr int kWarpMma = kEpiWarps, kWarpSched = kWarpMma + 1, kWarpProd = kWarpSched + 1,
              kWarpBias = kWarpProd + kProdWarps;
constexpr int kThreads = 32 * (kWarpBias + 1);
constexpr int kRowsPerWarp = kT / kProdWarps;   // 8
constexpr uint32_t kTmemCols = 512;
constexpr int kXposeBytes = kEpiWarps * 32 * 33 * 4;
constexpr int kSmem = 1024 + kStages * kStageBytes + kXposeBytes + kQueue * 16 + 512;
// Optional L2 eviction-priority hints (GHF_WGRAD_FLAGS): source rows evict_first / destination rows evict_last / g_W
// reductions evict_last.  At c3 no combination moves the kernel outside run-to-run noise (tools/wgrad_sweep.py), so the
// default is none.  The kernel is HBM-bound (ncu, profiles/r01_ncu_backward_summary.txt): 5.4 GB of row gathers that
// miss L2 plus 2 x 3.5 GB for the read-modify-write of g_W lines - a relation's 128 KiB come back every ~50 units,
// after ~240 MB of row traffic has passed through L2.
constexpr uint32_t kFlagSrcEvictFirst = 1u, kFlagDstEvictLast = 2u, kFlagRedEvictLast = 4u;
constexpr uint32_t kDefaultFlags = 0u;

// kind::f16, D fp32, A and B fp16, BOTH MN-major ([15], [16]), N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(kD >> 3) << 17) | ((uint32_t)(kD >> 4) << 24);

// 16 bytes global -> shared, or 16 zero bytes when `bytes` == 0
__device__ __forceinline__ void cp_async_16_zfill(uint32_t dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
mp_wgrad_f16_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
                    const int32_t* __restrict__ unit_rel, int64_t num_units, const int32_t* __restrict__ src_sorted,
                    const int32_t* __restrict__ dst_sorted, const __half* __restrict__ h16, int64_t dst_lo,
                    const float* __restrict__ h_scale, const __half* __restrict__ g16,
                    const float* __restrict__ g_scale, float* __restrict__ gW_msg, float* __restrict__ gW_self,
                    float* __restrict__ gbias, int* __restrict__ unit_counter, uint32_t flags) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sA = (raw + 1023u) & ~1023u;
  const uint32_t sX = sA + kStages * kStageBytes;          // per-epilogue-warp 32 x 33 float transposes
  const uint32_t sQ = sX + kXposeBytes;
  const uint32_t sBar = sQ + kQueue * 16;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return sBar + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * kStages + 2 + a); };
  const uint32_t q_full0 = sBar + 8u * (2 * kStages + 4);
  const uint32_t q_empty0 = q_full0 + 8u * kQueue;
  const uint32_t tmem_slot = q_empty0 + 8u * kQueue;
  volatile int4* q_ptr = reinterpret_cast<volatile int4*>(smem_raw + (sQ - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kConsumers = kEpiWarps + 1 + kProdWarps + 1;

  auto q_acquire = [&](uint32_t idx) -> int4 {             // {first sorted edge (-1: done), edges, relation, -}
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
      mbar_init(full(s), 32 * kProdWarps);
      mbar_init(empty(s), 2);               // tcgen05.commit + the bias-gradient warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), kEpiWarps);
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
    // ------------------------------------------------------------------ epilogue
    const int q = warp;                                    // TMEM lanes [32q, 32q + 32) = matrix rows
    float* xp = reinterpret_cast<float*>(smem_raw + (sX - raw)) + warp * (32 * 33);
    const float scale = h_scale[0] * g_scale[0];
    const uint64_t pol_red = (flags & kFlagRedEvictLast) ? policy_evict_last() : policy_evict_normal();
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      if (t.x < 0) break;
      q_release(it);
      const int a = (int)(it & 1u);
      mbar_wait(acc_full(a), (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int mat = 0; mat < 2; ++mat) {
        float* out = (mat ? gW_self : gW_msg) + ((int64_t)t.z * kD + 32 * q) * kD + lane;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 256 + mat * kD + cb * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) xp[lane * 33 + j] = __uint_as_float(r[j]) * scale;   // row = lane
          __syncwarp();
#pragma unroll
          for (int rr = 0; rr < 32; ++rr)                 // one matrix row per instruction: a 128 B line
            red_add_f32_hint(out + (int64_t)rr * kD + cb * 32, xp[rr * 33 + lane], pol_red);
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(a));
    }
  } else if (warp >= kWarpProd && warp < kWarpBias) {
    // ------------------------------------------------------------------ row-gather producers
    const int pw = warp - kWarpProd;
    const int l16 = lane & 15, hi = lane >> 4;
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h16) + l16 * 16;
    const uint8_t* gb = reinterpret_cast<const uint8_t*>(g16) + l16 * 16;
    const uint64_t pol_src = (flags & kFlagSrcEvictFirst) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_dst = (flags & kFlagDstEvictLast) ? policy_evict_last() : policy_evict_normal();
    struct Ids { int src, dst; };
    auto ids_of = [&](const int4& t, int s) -> Ids {
      const int row = s * kT + kRowsPerWarp * pw + (lane & (kRowsPerWarp - 1));
      if (t.x < 0 || row >= t.y) return Ids{-1, -1};
      return Ids{src_sorted[t.x + row], dst_sorted[t.x + row]};
    };
    int stage = 0;
    uint32_t phase = 0, qi = 0;
    int4 t = q_acquire(0);
    int s = 0;
    Ids ids = ids_of(t, 0);
    while (t.x >= 0) {
      uint32_t nqi = qi;
      int4 nt = t;
      int ns = s + 1;
      if (ns * kT >= t.y) {
        nqi = qi + 1;
        nt = q_acquire(nqi);
        ns = 0;
      }
      const Ids nids = ids_of(nt, ns);                     // in flight while this stage's rows are issued
      const int rows = min(kT, t.y - s * kT);
      const int rows16 = (rows + 15) & ~15;
      mbar_wait(empty(stage), phase ^ 1u);
      const uint32_t base = sA + stage * kStageBytes + (l16 >> 3) * kChunk;
#pragma unroll
      for (int set = 0; set < 3; ++set) {
        const int mine = set == 0 ? ids.src : ids.dst;
        const uint8_t* table = set == 0 ? hb : (set == 1 ? hb + dst_lo * (kD * 2) : gb);
        const uint64_t pol = set == 0 ? pol_src : pol_dst;
#pragma unroll
        for (int i = 0; i < kRowsPerWarp / 2; ++i) {
          const int rl = 2 * i + hi;
          const int idx = __shfl_sync(0xffffffffu, mine, rl);
          const int row = kRowsPerWarp * pw + rl;
          const uint32_t to = base + set * kSet + row * 128 + (((l16 & 7) ^ (row & 7)) << 4);
          if (idx >= 0) cp_async_16_hint(to, table + (int64_t)idx * (kD * 2), pol);
          else if (row < rows16) cp_async_16_zfill(to, table, 0u);
        }
      }
      cp_async_arrive_noinc(full(stage));
      if (++stage == kStages) { stage = 0; phase ^= 1u; }
      if (nqi != qi) q_release(qi);
      qi = nqi; t = nt; s = ns; ids = nids;
    }
  } else if (warp == kWarpMma) {
    // ------------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      if (t.x < 0) break;
      q_release(it);
      const int a = (int)(it & 1u);
      mbar_wait(acc_empty(a), ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_msg = tmem_base + (uint32_t)(a * 256), d_self = d_msg + kD;
      const int nstages = (t.y + kT - 1) / kT;
#pragma unroll 1
      for (int s = 0; s < nstages; ++s) {
        const int rows = min(kT, t.y - s * kT);
        const int ksteps = (rows + 15) >> 4;
        mbar_wait(full(stage), phase);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t st = sA + stage * kStageBytes;
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t b = umma_desc_mn128(st + 2 * kSet + k * 2048, kChunk);
            const uint32_t accum = (uint32_t)(s | k);
            umma_f16_ss(d_msg, umma_desc_mn128(st + k * 2048, kChunk), b, kIdesc, accum);
            umma_f16_ss(d_self, umma_desc_mn128(st + kSet + k * 2048, kChunk), b, kIdesc, accum);
          }
          umma_commit(empty(stage));
          if (s == nstages - 1) umma_commit(acc_full(a));
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kWarpBias) {
    // ------------------------------------------------------------------ bias gradient: column sums of the g rows
    // lane l owns columns [4l, 4l + 4): chunk l / 16, 16-byte group (l % 16) / 2, half of the group l % 2
    const uint32_t lane_off = (uint32_t)(2 * kSet + (lane >> 4) * kChunk + 8 * (lane & 1));
    const int grp = (lane & 15) >> 1;
    const float gs = g_scale[0];
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t it = 0;; ++it) {
      const int4 t = q_acquire(it);
      if (t.x < 0) break;
      q_release(it);
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      const int nstages = (t.y + kT - 1) / kT;
#pragma unroll 1
      for (int s = 0; s < nstages; ++s) {
        const int rows = min(kT, t.y - s * kT);
        mbar_wait(full(stage), phase);
        const uint32_t base = sA + stage * kStageBytes + lane_off;
#pragma unroll 4
        for (int e = 0; e < rows; ++e) {
          uint32_t lo, hi;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(lo), "=r"(hi)
                       : "r"(base + e * 128 + ((grp ^ (e & 7)) << 4)));
          const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo));
          const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&hi));
          sum[0] += a.x; sum[1] += a.y; sum[2] += b.x; sum[3] += b.y;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty(stage));
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      float* out = gbias + (int64_t)t.z * kD + 4 * lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(out + j, sum[j] * gs);
    }
  } else if (warp == kWarpSched) {
    // ------------------------------------------------------------------ scheduler
    uint32_t qi = 0;
    auto publish = [&](int start, int rows, int rel) {
      if (lane == 0) {
        mbar_wait(q_empty0 + 8u * (qi % kQueue), ((qi / kQueue) & 1u) ^ 1u);
        volatile int4* p = q_ptr + (qi % kQueue);
        p->x = start; p->y = rows; p->z = rel; p->w = 0;
        mbar_arrive(q_full0 + 8u * (qi % kQueue));
      }
      ++qi;
    };
    auto draw = [&]() -> int64_t {
      int v = 0;
      if (lane == 0) v = atomicAdd(unit_counter, 1);
      return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    };
    int64_t u = draw();
    while (u < num_units) {
      const int start = unit_start[u], count = unit_count[u], rel = unit_rel[u];
      const int64_t u_next = draw();
      if (count > 0) publish(start, count, rel);
      u = u_next;
    }
    publish(-1, 0, 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) tmem_dealloc<kTmemCols>(tmem_base);
}

}  // namespace

// g_W_msg / g_W_self / g_bias must be zero at entry.  h16: [N,128] fp16 shadow of h (+ scale), g16: [local,128] fp16
// shadow of g_acc (+ scale).  `unit_counter`: one int of device memory.
int mp_wgrad_f16_launch(const ghf_graph* g, const void* h16, const float* h_scale, const void* g16,
                        const float* g_scale, float* gW_msg, float* gW_self, float* gbias, int* unit_counter,
                        cudaStream_t stream) {
  GHF_REQUIRE(g->hidden_dim == kD, "mp_wgrad_f16: hidden_dim must be %d", kD);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h16) | reinterpret_cast<uintptr_t>(g16) |
               reinterpret_cast<uintptr_t>(gW_msg) | reinterpret_cast<uintptr_t>(gW_self)) % 16 == 0,
              "mp_wgrad_f16: buffers must be 16-byte aligned");
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(mp_wgrad_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  }
  GHF_CUDA(cudaMemsetAsync(unit_counter, 0, sizeof(int), stream));
  const char* fenv = getenv("GHF_WGRAD_FLAGS");
  const uint32_t flags = fenv ? (uint32_t)atoi(fenv) : kDefaultFlags;
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  mp_wgrad_f16_kernel<<<(unsigned)grid, kThreads, kSmem, stream>>>(
      g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted,
      reinterpret_cast<const __half*>(h16), g->dst_lo, h_scale, reinterpret_cast<const __half*>(g16), g_scale, gW_msg,
      gW_self, gbias, unit_counter, flags);
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace ghf
