// mp_umma.cu — the tensor-core contraction of the message-passing layer (HG:201-204 + HG:217-228)
// as a persistent, warp-specialised tcgen05 kernel for sm_100a.
//
// Work unit = up to `unit_edges` consecutive edges of the sorted order that share one relation r.
// For every 128-edge tile of a unit:
//     D[128, d] = A[128, 2d] * B_r[2d, d],   A row e = [ h[src_e] | h[dst_e] ],  B_r = [W_msg[r]; W_self[r]]
// with A gathered row by row into 128B-swizzled shared memory (cp.async, 16 B per lane, 8 lanes per
// row chunk so every 128 B line is fetched whole), B_r resident in shared memory for the whole unit
// (one 1-D bulk copy of a pre-swizzled image), the accumulator in TMEM (double buffered), and an
// epilogue that transposes 32x32 blocks through shared memory so each warp issues full-line
// red.global.add.v4.f32 into acc[dst_e] (+ bias[r]).
//
// Warp roles (320 threads, 1 CTA / SM):
//   warps 0-3  epilogue   (TMEM lanes 32w..32w+31 -> registers -> smem transpose -> vector red)
//   warps 4-7  A producers (row gather, cp.async -> full[stage])
//   warp  8    MMA issuer (lane 0) + TMEM allocator
//   warp  9    B loader   (lane 0, cp.async.bulk -> b_full)
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mp.cuh"
#include "umma.cuh"

namespace ghf {
namespace {

using namespace ptx;

// L2 residency flags (GHF_MP_FLAGS): the destination super-block (h[dst] rows + accumulator rows) should
// stay in L2 while the source rows stream through once.
constexpr uint32_t kFlagSrcEvictFirst = 1u, kFlagDstEvictLast = 2u, kFlagRedEvictLast = 4u;
constexpr uint32_t kFlagStaticSchedule = 8u;
constexpr uint32_t kFlagWeightsEvictLast = 16u;  // keep the operand images resident across super-block phases  // units round-robin by CTA instead of the shared counter (experiments)
constexpr uint32_t kDefaultFlags = kFlagSrcEvictFirst | kFlagDstEvictLast | kFlagRedEvictLast;

template <int D>
struct Cfg {
  static constexpr int kTileM = 128;
  static constexpr int kK = 2 * D;
  static constexpr int kChunks = kK / 32;             // 32 tf32 = one 128 B swizzle row
  static constexpr int kHalf = kChunks / 2;           // chunks taken from h[src]; the rest from h[dst]
  static constexpr int kAStage = kTileM * 128;        // 16 KiB
  static constexpr int kBChunk = D * 128;
  static constexpr int kBBytes = kChunks * kBChunk;   // 2*D*D*4
  static constexpr int kStaging = 4 * 32 * 128;       // 4 epilogue warps x (32 rows x 32 fp32)
  static constexpr int kStages = D == 128 ? 5 : 8;
  static constexpr int kTmemCols = 2 * D;             // two accumulators; 64 / 128 / 256 (powers of two)
  static constexpr int kBarBytes = 512;
  static constexpr int kQueue = 4;                    // unit-id ring between the scheduler and the other roles
  static constexpr int kSmem = 1024 + kBBytes + kStages * kAStage + kStaging + kBarBytes;
  static constexpr int kThreads = 320;
};

// element (n, k) of relation r's operand image; k < D -> W_msg[r][k][n], else W_self[r][k-D][n]
__device__ __forceinline__ int64_t pack_offset_bytes(int D, int n, int k) {
  const int c = k >> 5, kk = k & 31;
  return (int64_t)c * (D * 128) + (int64_t)n * 128 + ((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
}

__global__ void pack_weights_kernel(const float* __restrict__ W_msg, const float* __restrict__ W_self, int R,
                                    int D, uint8_t* __restrict__ pack) {
  // one thread per (r, k, n): reads are contiguous in n, writes land in the swizzled image
  const int64_t total = (int64_t)R * 2 * D * D;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % D);
  const int k = (int)((idx / D) % (2 * D));
  const int64_t r = idx / ((int64_t)2 * D * D);
  const float v = k < D ? W_msg[(r * D + k) * D + n] : W_self[(r * D + (k - D)) * D + n];
  *reinterpret_cast<float*>(pack + r * ((int64_t)2 * D * D * 4) + pack_offset_bytes(D, n, k)) = to_tf32_rna(v);
}

template <int D>
__global__ void __launch_bounds__(Cfg<D>::kThreads, 1)
mp_umma_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
               const int32_t* __restrict__ unit_rel, int64_t num_units,
               const int32_t* __restrict__ src_sorted, const int32_t* __restrict__ dst_sorted,
               const float* __restrict__ h, int64_t dst_lo, const uint8_t* __restrict__ wpack,
               const float* __restrict__ bias, float* __restrict__ acc, int* __restrict__ unit_counter, uint32_t flags) {
  using C = Cfg<D>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sB = (raw + 1023u) & ~1023u;
  const uint32_t sA = sB + C::kBBytes;
  const uint32_t sStg = sA + C::kStages * C::kAStage;
  const uint32_t sBar = sStg + C::kStaging;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (C::kStages + s); };
  const uint32_t b_full = sBar + 8u * (2 * C::kStages);
  const uint32_t b_empty = b_full + 8u;
  auto acc_full = [&](int a) { return b_empty + 8u + 8u * a; };
  auto acc_empty = [&](int a) { return b_empty + 24u + 8u * a; };
  const uint32_t tmem_slot = b_empty + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // unit queue: the scheduler (warp 9) publishes unit ids; epilogue warps, producer warps and the MMA thread
  // each consume every entry (9 arrivals free a slot)
  const uint32_t q_full0 = b_empty + 48u;
  const uint32_t q_empty0 = q_full0 + 8u * C::kQueue;
  const uint32_t q_slots = q_empty0 + 8u * C::kQueue;
  volatile int32_t* q_slot_ptr = reinterpret_cast<volatile int32_t*>(smem_raw + (q_slots - raw));
  int q_idx = 0;
  uint32_t q_phase = 0;
  auto next_unit = [&](bool whole_warp) -> int {
    mbar_wait(q_full0 + 8u * q_idx, q_phase);
    const int u = q_slot_ptr[q_idx];
    if (whole_warp) __syncwarp();
    if (!whole_warp || lane == 0) mbar_arrive(q_empty0 + 8u * q_idx);
    if (++q_idx == C::kQueue) { q_idx = 0; q_phase ^= 1u; }
    return u;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full(s), 128);  // one cp.async-completion arrival per producer thread
      mbar_init(empty(s), 1);   // tcgen05.commit
    }
    mbar_init(b_full, 1);       // arrive.expect_tx by the loader + transaction bytes
    mbar_init(b_empty, 1);      // tcgen05.commit
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);     // tcgen05.commit
      mbar_init(acc_empty(a), 128);  // every epilogue thread
    }
    for (int q = 0; q < C::kQueue; ++q) {
      mbar_init(q_full0 + 8u * q, 1);
      mbar_init(q_empty0 + 8u * q, 9);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<C::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    float4* stg = reinterpret_cast<float4*>(smem_raw + (sStg - raw) + warp * 4096);
    const int cj = lane & 7;
    // one code path, policy chosen at run time (a branch between hinted / unhinted forms in the inner
    // loops cost 35% of the kernel when tried)
    const uint64_t pol_red = (flags & kFlagRedEvictLast) ? policy_evict_last() : policy_evict_normal();
    uint32_t it = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      const int64_t rel = unit_rel[u];
      float4 b4[D / 32];
#pragma unroll
      for (int cc = 0; cc < D / 32; ++cc)
        b4[cc] = *reinterpret_cast<const float4*>(bias + rel * D + cc * 32 + 4 * cj);
      for (int t0 = 0; t0 < count; t0 += C::kTileM, ++it) {
        const int a = it & 1;
        const int rows = min(C::kTileM, count - t0);
        const int my_row = warp * 32 + lane;
        const int my_dst = my_row < rows ? dst_sorted[start + t0 + my_row] : -1;
        mbar_wait(acc_full(a), (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < D / 32; ++cc) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * D + cc * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg[lane * 8 + (j ^ (lane & 7))] =
                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            const int dsti = __shfl_sync(0xffffffffu, my_dst, rr);
            float4 v = stg[rr * 8 + (cj ^ (rr & 7))];
            if (dsti >= 0) {
              v.x += b4[cc].x; v.y += b4[cc].y; v.z += b4[cc].z; v.w += b4[cc].w;
              float* p = acc + (int64_t)dsti * D + cc * 32 + 4 * cj;
              red_add_v4_hint(p, v, pol_red);
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(acc_empty(a));
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A producers
    const int pw = warp - 4;
    const int cj = lane & 7;
    const uint64_t pol_src = (flags & kFlagSrcEvictFirst) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_dst = (flags & kFlagDstEvictLast) ? policy_evict_last() : policy_evict_normal();
    int stage = 0;
    uint32_t phase = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      for (int t0 = 0; t0 < count; t0 += C::kTileM) {
        const int rows = min(C::kTileM, count - t0);
        const int my_row = pw * 32 + lane;
        const bool ok = my_row < rows;
        const int64_t my_src = ok ? (int64_t)src_sorted[start + t0 + my_row] : -1;
        const int64_t my_dst = ok ? dst_lo + dst_sorted[start + t0 + my_row] : -1;
#pragma unroll 1
        for (int c = 0; c < C::kChunks; ++c) {
          mbar_wait(empty(stage), phase ^ 1u);
          const bool from_src = c < C::kHalf;
          const int64_t mine = from_src ? my_src : my_dst;
          const int col = (from_src ? c : c - C::kHalf) * 32 + 4 * cj;
          const uint32_t dst_base = sA + stage * C::kAStage;
          const uint64_t pol = from_src ? pol_src : pol_dst;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            const int64_t idx = __shfl_sync(0xffffffffu, mine, rr);
            const int row = pw * 32 + rr;
            const uint32_t to = dst_base + row * 128 + ((cj ^ (row & 7)) << 4);
            if (idx >= 0) cp_async_16_hint(to, h + idx * D + col, pol);
          }
          cp_async_arrive_noinc(full(stage));
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(C::kTileM, D);
      int stage = 0;
      uint32_t phase = 0, bphase = 0, it = 0;
      for (int u = next_unit(false); u >= 0; u = next_unit(false)) {
        const int count = unit_count[u];
        mbar_wait(b_full, bphase);
        tc_fence_after();
        for (int t0 = 0; t0 < count; t0 += C::kTileM, ++it) {
          const int a = it & 1;
          mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * D);
#pragma unroll 1
          for (int c = 0; c < C::kChunks; ++c) {
            mbar_wait(full(stage), phase);
            fence_proxy_async();
            tc_fence_after();
            const uint64_t adesc = umma_desc_k128(sA + stage * C::kAStage);
            const uint64_t bdesc = umma_desc_k128(sB + c * C::kBChunk);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_tf32(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (c | j) != 0);
            umma_commit(empty(stage));
            if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(acc_full(a));
        }
        umma_commit(b_empty);
        bphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ unit scheduler + B loader
    if (lane == 0) {
      uint32_t bphase = 0, sphase = 0;
      int sq = 0;
      const uint64_t pol_w = (flags & kFlagWeightsEvictLast) ? policy_evict_last() : policy_evict_normal();
      int64_t static_next = blockIdx.x;
          for (;;) {
        mbar_wait(q_empty0 + 8u * sq, sphase ^ 1u);
        int64_t u;
        if (flags & kFlagStaticSchedule) { u = static_next; static_next += gridDim.x; }
        else u = atomicAdd(unit_counter, 1);
        const bool done = u >= num_units;
        q_slot_ptr[sq] = done ? -1 : (int)u;
        mbar_arrive(q_full0 + 8u * sq);  // release: the slot write is visible to the waiters
        if (++sq == C::kQueue) { sq = 0; sphase ^= 1u; }
        if (done) break;
        const int64_t rel = unit_rel[u];
        mbar_wait(b_empty, bphase ^ 1u);
        mbar_arrive_expect_tx(b_full, C::kBBytes);
        const uint8_t* src = wpack + rel * (int64_t)C::kBBytes;
        constexpr int kPiece = 16384;
#pragma unroll 1
        for (int off = 0; off < C::kBBytes; off += kPiece) {
          const int bytes = min(kPiece, C::kBBytes - off);
          bulk_g2s_hint(sB + off, src + off, bytes, b_full, pol_w);
        }
        bphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<C::kTmemCols>(tmem_base);
}

uint32_t env_flags() {
  const char* env = getenv("GHF_MP_FLAGS");
  return env ? (uint32_t)atoi(env) : kDefaultFlags;
}

template <int D>
int launch(const ghf_graph* g, const float* h, const uint8_t* wpack, const float* bias, float* acc,
           int* unit_counter, cudaStream_t stream) {
  using C = Cfg<D>;
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(mp_umma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
  }
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  mp_umma_kernel<D><<<(unsigned)grid, C::kThreads, C::kSmem, stream>>>(
      g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted, h, g->dst_lo, wpack,
      bias, acc, unit_counter, env_flags());
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool mp_umma_supported(int d) { return d == 32 || d == 64 || d == 128; }

int64_t mp_umma_pack_bytes(int num_rel, int d) { return align_up((int64_t)num_rel * 2 * d * d * 4, 256); }

int mp_umma_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch,
                 cudaStream_t stream) {
  const int d = g->hidden_dim;
  GHF_REQUIRE(reinterpret_cast<uintptr_t>(pack_scratch) % 16 == 0, "mp_umma: scratch must be 16-byte aligned");
  const int64_t total = (int64_t)g->num_rel * 2 * d * d;
  pack_weights_kernel<<<(unsigned)cdiv(total, 256), 256, 0, stream>>>(W_msg, W_self, g->num_rel, d,
                                                                    reinterpret_cast<uint8_t*>(pack_scratch));
  GHF_LAUNCH_CHECK();
  return 0;
}

int mp_umma_launch(const ghf_graph* g, const float* h, const float* bias, float* acc, const void* pack_scratch,
                   int* unit_counter, cudaStream_t stream) {
  const int d = g->hidden_dim;
  GHF_REQUIRE(g->unit_edges % 128 == 0, "mp_umma: unit_edges=%d must be a multiple of 128", g->unit_edges);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(acc) |
               reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_umma: h / acc / bias / scratch must be 16-byte aligned");
  const uint8_t* pack = reinterpret_cast<const uint8_t*>(pack_scratch);
  switch (d) {
    case 32: return launch<32>(g, h, pack, bias, acc, unit_counter, stream);
    case 64: return launch<64>(g, h, pack, bias, acc, unit_counter, stream);
    case 128: return launch<128>(g, h, pack, bias, acc, unit_counter, stream);
  }
  return fail("mp_umma: unsupported hidden_dim %d", d);
}

}  // namespace ghf
