// mp_f16_ss.cu — message-passing contraction on tcgen05 kind::f16 with STREAMED weights, hidden_dim 256 and 64
// (GHF_PREC_F16; BASELINE config 4: 20k relation texts, ~100 edges per relation, 10.5 GB of generated fp32 weights
// per layer - and config 5, hidden 64, where the point is the halved row-gather traffic).  Written for hidden 256
// (numbers below); hidden 64 is the same kernel with 2 K-chunks, 8 KiB weight chunks and 8 stages.
//
// At hidden 256 one relation's operand [W_msg[r]; W_self[r]] is 512 x 256 values: 256 KiB even as fp16 - more than
// shared memory, twice tensor memory.  And with ~100 edges per relation every tile needs a different one.  So the
// weights are not resident anywhere: for every 128-edge tile
//     D[128 edges, 256] = A[128, 512] * B_r[512, 256]      (fp16 operands, fp32 accumulation in TMEM)
// runs over 8 K-chunks of 64, and a pipeline stage carries BOTH operands of one chunk:
//     A chunk: 128 gathered half-rows of h16 (cp.async, 128 B each, 128B-swizzled)                     16 KiB
//     B chunk: 64 k x 256 n halfs of relation r's pre-swizzled image (cp.async.bulk, contiguous 32 KiB) 32 KiB
// 4 stages = 192 KiB.  The kernel is bound by the weight stream (256 KiB per tile): at c4 5.2 GB per layer instead of
// the 10.5 GB of fp32 weights the CUDA-core kernel reads - and on tensor cores instead of 524 GFLOP of FFMA.
// Scales as in mp_f16.cu: h16 = fp16(h * 2^k), image = fp16(W_r * 2^k_r), the epilogue multiplies by 2^-(k + k_r).
//
// Warp roles (320 threads, 1 CTA / SM, persistent; the layout of mp_umma.cu):
//   0-3  epilogue (TMEM -> 32x32 transposes through shared memory -> red.global.add.v4.f32 rows of acc[dst])
//   4-7  A producers (row gather)      8  MMA issuer + TMEM allocator
//   9    unit scheduler (one unit AHEAD of the loads) + B loader (bulk copies)
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"
#include "mp.cuh"
#include "umma.cuh"

namespace ghf {
namespace {

using namespace ptx;

constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;             // A chunk: 128 half-rows of 128 B = 16 KiB
constexpr int kStaging = 4 * 32 * 128;            // 4 epilogue warps x (32 rows x 32 fp32)
constexpr int kQueue = 4;
constexpr int kThreads = 320;

template <int D>
struct Cfg {
  static constexpr int kChunks = 2 * D / 64;              // K-chunks of 64 halfs (one 128 B swizzle row): 8 / 2
  static constexpr int kHalfChunks = kChunks / 2;         // chunks taken from h16[src]; the rest from h16[dst]
  static constexpr int kBBytes = D * 128;                 // B chunk: D rows of 128 B = 32 KiB / 8 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = D == 256 ? 4 : 8;
  static constexpr uint32_t kTmemCols = 2 * D;            // two accumulators
  static constexpr int kSmem = 1024 + kStages * kStageBytes + kStaging + 512;
  static constexpr int64_t kImageBytes = (int64_t)kChunks * kBBytes;   // 256 KiB / 16 KiB per relation
  // kind::f16: D fp32, A fp16 K-major (gathered rows), B fp16 MN-major ([16]: n is the contiguous index of a
  // generated matrix, W[k][n]), M = 128, N = D
  static constexpr uint32_t kIdesc = (1u << 4) | (1u << 16) | ((uint32_t)(D >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
};


// entry (k, n) of relation r's operand image (k < D -> W_msg[r][k][n], else W_self[r][k-D][n]): K-chunk k/64, inside it
// 64-wide n blocks of [64 k][128 B], 16-byte group (n%64)/8 swizzled by k%8 - the MN-major SWIZZLE_128B operand
// (LBO = 8 KiB between n blocks, SBO = 1 KiB per 8 k).  n is the contiguous index, as in W itself, which is what lets
// the generator's last Linear write images directly (linear_umma.cu, ImageOut - the same formula).
template <int D>
__device__ __forceinline__ int64_t image_offset_bytes(int n, int k) {
  return (int64_t)(k >> 6) * Cfg<D>::kBBytes + (int64_t)(n >> 6) * (64 * 128) + (k & 63) * 128 +
         (((((n & 63) >> 3) ^ (k & 7))) << 4) + (n & 7) * 2;
}

// One CTA per relation: max |W| -> power-of-two scale -> scaled fp16 image (see pack_f16_kernel in mp_f16.cu).
// A relation is 512 KiB of fp32 and is read twice (range, then conversion).  1024 threads and ONE CTA per SM (the
// launch asks for shared memory it does not use): 148 relations = 76 MB are in flight, so the second read hits L2 and
// HBM sees the weights once.
constexpr int kPackThreads = 1024;
constexpr int kPackSmem = 160 * 1024;
template <int kD>
__global__ void __launch_bounds__(kPackThreads, 1)
pack_f16_ss_kernel(const float* __restrict__ W_msg, const float* __restrict__ W_self, uint8_t* __restrict__ pack,
                   float* __restrict__ inv_scale, int num_rel) {
  __shared__ float s_max[kPackThreads / 32];
  __shared__ float s_scale;
  for (int64_t r = blockIdx.x; r < num_rel; r += gridDim.x) {
    float m = 0.f;
    for (int which = 0; which < 2; ++which) {
      const float4* w4 = reinterpret_cast<const float4*>((which ? W_self : W_msg) + r * kD * kD);
#pragma unroll 4
      for (int i = threadIdx.x; i < kD * kD / 4; i += kPackThreads) {
        const float4 a = __ldcg(w4 + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
    __syncthreads();                             // s_max / s_scale of the previous relation are no longer read
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = 0.f;
      for (int i = 0; i < kPackThreads / 32; ++i) mm = fmaxf(mm, s_max[i]);
      int e = 0;
      float scale = 1.f;
      if (mm > 0.f && isfinite(mm)) {
        frexpf(mm, &e);
        e = 15 - e;                                // mm * 2^(15-e) in [2^14, 2^15)
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        scale = ldexpf(1.f, e);
      }
      s_scale = scale;
      inv_scale[r] = 1.f / scale;
    }
    __syncthreads();
    const float scale = s_scale;
    uint8_t* img = pack + r * Cfg<kD>::kImageBytes;
    // thread = (one k, 8 consecutive n): 32 B read from one W row, one 16 B store
#pragma unroll 2
    for (int i = threadIdx.x; i < 2 * kD * (kD / 8); i += kPackThreads) {
      const int k = i / (kD / 8), n0 = (i % (kD / 8)) * 8;
      const float* src = k < kD ? W_msg + (r * kD + k) * kD + n0 : W_self + (r * kD + (k - kD)) * kD + n0;
      const float4 a = __ldcg(reinterpret_cast<const float4*>(src)), b = __ldcg(reinterpret_cast<const float4*>(src) + 1);
      const __half2 p0 = __floats2half2_rn(a.x * scale, a.y * scale), p1 = __floats2half2_rn(a.z * scale, a.w * scale);
      const __half2 p2 = __floats2half2_rn(b.x * scale, b.y * scale), p3 = __floats2half2_rn(b.z * scale, b.w * scale);
      *reinterpret_cast<uint4*>(img + image_offset_bytes<kD>(n0, k)) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1),
                     *reinterpret_cast<const uint32_t*>(&p2), *reinterpret_cast<const uint32_t*>(&p3));
    }
  }
}

template <int kD>
__global__ void __launch_bounds__(kThreads, 1)
mp_f16_ss_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
                 const int32_t* __restrict__ unit_rel, int64_t num_units, const int32_t* __restrict__ src_sorted,
                 const int32_t* __restrict__ dst_sorted, const __half* __restrict__ h16, int64_t dst_lo,
                 const float* __restrict__ h_scale, const uint8_t* __restrict__ wpack,
                 const float* __restrict__ w_inv_scale, const float* __restrict__ bias, float* __restrict__ acc,
                 int* __restrict__ unit_counter) {
  using C = Cfg<kD>;
  constexpr int kChunks = C::kChunks, kHalfChunks = C::kHalfChunks, kBBytes = C::kBBytes,
                kStageBytes = C::kStageBytes, kStages = C::kStages;
  constexpr uint32_t kTmemCols = C::kTmemCols, kIdesc = C::kIdesc;
  constexpr int64_t kImageBytes = C::kImageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sS = (raw + 1023u) & ~1023u;              // stages: [A chunk 16 KiB | B chunk]
  const uint32_t sStg = sS + kStages * kStageBytes;
  const uint32_t sBar = sStg + kStaging;
  auto full = [&](int s) { return sBar + 8u * s; };
  auto empty = [&](int s) { return sBar + 8u * (kStages + s); };
  auto acc_full = [&](int a) { return sBar + 8u * (2 * kStages + a); };
  auto acc_empty = [&](int a) { return sBar + 8u * (2 * kStages + 2 + a); };
  const uint32_t q_full0 = sBar + 8u * (2 * kStages + 4);
  const uint32_t q_empty0 = q_full0 + 8u * kQueue;
  const uint32_t q_slots = q_empty0 + 8u * kQueue;
  const uint32_t tmem_slot = q_slots + 4u * kQueue;
  volatile int32_t* q_slot_ptr = reinterpret_cast<volatile int32_t*>(smem_raw + (q_slots - raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // unit queue: the scheduler publishes unit ids; 4 epilogue warps + 4 producer warps + the MMA thread consume each
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
      mbar_init(full(s), 128 + 1);  // every producer thread (cp.async completion) + the loader's expect_tx arrival
      mbar_init(empty(s), 1);       // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), 128);
    }
    for (int q = 0; q < kQueue; ++q) {
      mbar_init(q_full0 + 8u * q, 1);
      mbar_init(q_empty0 + 8u * q, 9);
    }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    // ------------------------------------------------------------------ epilogue
    float4* stg = reinterpret_cast<float4*>(smem_raw + (sStg - raw) + warp * 4096);
    const int cj = lane & 7;
    const float h_inv = h_scale[0];
    uint32_t it = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      const int64_t rel = unit_rel[u];
      const float inv = w_inv_scale[rel] * h_inv;
      for (int t0 = 0; t0 < count; t0 += kTileM, ++it) {
        const int a = it & 1;
        const int rows = min(kTileM, count - t0);
        const int my_row = warp * 32 + lane;
        const int my_dst = my_row < rows ? dst_sorted[start + t0 + my_row] : -1;
        mbar_wait(acc_full(a), (it >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < kD / 32; ++cc) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + rel * kD + cc * 32 + 4 * cj);
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * kD + cc * 32), r);
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
              v.x = fmaf(v.x, inv, b4.x); v.y = fmaf(v.y, inv, b4.y);
              v.z = fmaf(v.z, inv, b4.z); v.w = fmaf(v.w, inv, b4.w);
              red_add_v4(acc + (int64_t)dsti * kD + cc * 32 + 4 * cj, v);
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(acc_empty(a));
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------------ A producers (half-rows of h16)
    const int pw = warp - 4;
    const int cj = lane & 7;
    const uint8_t* hb = reinterpret_cast<const uint8_t*>(h16) + cj * 16;
    int stage = 0;
    uint32_t phase = 0;
    for (int u = next_unit(true); u >= 0; u = next_unit(true)) {
      const int start = unit_start[u], count = unit_count[u];
      for (int t0 = 0; t0 < count; t0 += kTileM) {
        const int rows = min(kTileM, count - t0);
        const int my_row = pw * 32 + lane;
        const bool ok = my_row < rows;
        const int64_t my_src = ok ? (int64_t)src_sorted[start + t0 + my_row] : -1;
        const int64_t my_dst = ok ? dst_lo + dst_sorted[start + t0 + my_row] : -1;
#pragma unroll 1
        for (int c = 0; c < kChunks; ++c) {
          mbar_wait(empty(stage), phase ^ 1u);
          const bool from_src = c < kHalfChunks;
          const int64_t mine = from_src ? my_src : my_dst;
          const int col_bytes = (from_src ? c : c - kHalfChunks) * 128;
          const uint32_t dst_base = sS + stage * kStageBytes;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + (lane >> 3);
            const int64_t idx = __shfl_sync(0xffffffffu, mine, rr);
            const int row = pw * 32 + rr;
            const uint32_t to = dst_base + row * 128 + ((cj ^ (row & 7)) << 4);
            if (idx >= 0) cp_async_16(to, hb + idx * (kD * 2) + col_bytes);
          }
          cp_async_arrive_noinc(full(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int u = next_unit(false); u >= 0; u = next_unit(false)) {
        const int count = unit_count[u];
        for (int t0 = 0; t0 < count; t0 += kTileM, ++it) {
          const int a = it & 1;
          mbar_wait(acc_empty(a), ((it >> 1) & 1) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(a * kD);
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c) {
            mbar_wait(full(stage), phase);
            fence_proxy_async();
            tc_fence_after();
            const uint64_t adesc = umma_desc_k128(sS + stage * kStageBytes);
            const uint32_t b_addr = sS + stage * kStageBytes + kABytes;
#pragma unroll
            for (int j = 0; j < 4; ++j)          // K step of 16: +32 B in the K-major A rows, +2 KiB (16 k rows) in B
              umma_f16_ss(d_tmem, adesc + 2 * j, umma_desc_mn128(b_addr + j * 2048, 64 * 128), kIdesc, (c | j) != 0);
            umma_commit(empty(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(acc_full(a));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ unit scheduler + B loader
    if (lane == 0) {
      uint32_t sphase = 0, phase = 0;
      int sq = 0, stage = 0;
      auto publish = [&](int64_t u) {
        mbar_wait(q_empty0 + 8u * sq, sphase ^ 1u);
        q_slot_ptr[sq] = u >= num_units ? -1 : (int)u;
        mbar_arrive(q_full0 + 8u * sq);
        if (++sq == kQueue) { sq = 0; sphase ^= 1u; }
      };
      int64_t u = atomicAdd(unit_counter, 1);
      publish(u);
      while (u < num_units) {
        const int64_t u_next = atomicAdd(unit_counter, 1);   // the other roles learn the next unit while this one loads
        publish(u_next);
        const int count = unit_count[u];
        const uint8_t* img = wpack + (int64_t)unit_rel[u] * kImageBytes;
        for (int t0 = 0; t0 < count; t0 += kTileM) {
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c) {
            mbar_wait(empty(stage), phase ^ 1u);
            mbar_arrive_expect_tx(full(stage), kBBytes);
            const uint32_t to = sS + stage * kStageBytes + kABytes;
            constexpr int kPiece = kBBytes > 16384 ? 16384 : kBBytes;
#pragma unroll
            for (int off = 0; off < kBBytes; off += kPiece)
              bulk_g2s(to + off, img + (int64_t)c * kBBytes + off, kPiece, full(stage));
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
          }
        }
        u = u_next;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<kTmemCols>(tmem_base);
}

// Per-relation power-of-two scale of the operand images WITHOUT a pass over the generated values: an entry of
// W[r] is alpha * (z_r . w_j + b_j), so |entry| <= alpha * (|z_r|_1 * max|w| + max|b|) - z_r the (post-ReLU) input of
// the generator's last Linear, w / b its parameters.  The bound is loose by about |z|_1 / |z|_2 ~ 2^4, which only
// costs exponent headroom (fp16 keeps its 11-bit significand for anything within 2^25 of the scaled maximum).
// One scale per relation, shared by the W_msg and W_self halves (they meet in one accumulator).
__global__ void __launch_bounds__(256)
image_scale_kernel(const float* __restrict__ Zm, const float* __restrict__ Zs, int H, int64_t R,
                   const float* __restrict__ words, const float* __restrict__ ls_m, const float* __restrict__ ls_s,
                   float* __restrict__ scale, float* __restrict__ inv_scale) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  float l1m = 0.f, l1s = 0.f;
  for (int c = lane; c < H; c += 32) {
    l1m += fabsf(Zm[r * H + c]);
    l1s += fabsf(Zs[r * H + c]);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    l1m += __shfl_xor_sync(0xffffffffu, l1m, s);
    l1s += __shfl_xor_sync(0xffffffffu, l1s, s);
  }
  if (lane == 0) {
    // words: [2i + 1] = max |.| of W3_msg, b3_msg, W3_self, b3_self (the scale-word pairs of mp_f16_absmax)
    const float bm = expf(*ls_m) * (l1m * words[1] + words[3]);
    const float bs = expf(*ls_s) * (l1s * words[5] + words[7]);
    const float bound = fmaxf(bm, bs);
    float sc = 1.f;
    if (bound > 0.f && isfinite(bound)) {
      int e;
      frexpf(bound, &e);                          // bound < 2^e  ->  bound * 2^(15 - e) < 2^15
      e = 15 - e;
      sc = ldexpf(1.f, e > 100 ? 100 : (e < -100 ? -100 : e));
    }
    scale[r] = sc;
    inv_scale[r] = 1.f / sc;
  }
}

}  // namespace

bool mp_f16ss_supported(int d) { return d == 256 || d == 64; }

static int64_t image_bytes(int d) { return d == 256 ? Cfg<256>::kImageBytes : Cfg<64>::kImageBytes; }

int64_t mp_f16ss_pack_bytes(int num_rel, int d) {
  return align_up((int64_t)num_rel * image_bytes(d), 256) + align_up((int64_t)num_rel * 4, 256);
}

template <int D>
static int pack_impl(const ghf_graph* g, const float* W_msg, const float* W_self, uint8_t* img, float* inv,
                     cudaStream_t stream) {
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(pack_f16_ss_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackSmem));
  }
  const int grid = g->num_rel < sm_count() ? g->num_rel : sm_count();
  pack_f16_ss_kernel<D><<<(unsigned)grid, kPackThreads, kPackSmem, stream>>>(W_msg, W_self, img, inv, g->num_rel);
  GHF_LAUNCH_CHECK();
  return 0;
}

int mp_f16ss_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch,
                  cudaStream_t stream) {
  const int d = g->hidden_dim;
  GHF_REQUIRE(mp_f16ss_supported(d), "mp_f16_ss: hidden_dim must be 64 or 256, got %d", d);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(W_msg) | reinterpret_cast<uintptr_t>(W_self) |
               reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_f16_ss: W_msg / W_self / scratch must be 16-byte aligned");
  uint8_t* img = reinterpret_cast<uint8_t*>(pack_scratch);
  float* inv = reinterpret_cast<float*>(img + align_up((int64_t)g->num_rel * image_bytes(d), 256));
  return d == 256 ? pack_impl<256>(g, W_msg, W_self, img, inv, stream) : pack_impl<64>(g, W_msg, W_self, img, inv, stream);
}

template <int D>
static int launch_impl(const ghf_graph* g, const __half* h16, const float* h16_scale, const float* bias, float* acc,
                       const uint8_t* img, const float* inv, int* unit_counter, cudaStream_t stream) {
  static bool configured[64] = {false};
  if (first_use_on_device(configured)) {
    GHF_CUDA(cudaFuncSetAttribute(mp_f16_ss_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::kSmem));
  }
  const int64_t grid = g->num_units < sm_count() ? g->num_units : sm_count();
  mp_f16_ss_kernel<D><<<(unsigned)grid, kThreads, Cfg<D>::kSmem, stream>>>(
      g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted, h16, g->dst_lo, h16_scale,
      img, inv, bias, acc, unit_counter);
  GHF_LAUNCH_CHECK();
  return 0;
}

// acc (zero at entry) += the tiles' products; `unit_counter`: one zeroed int
int mp_f16ss_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias, float* acc,
                    const void* pack_scratch, int* unit_counter, cudaStream_t stream) {
  const int d = g->hidden_dim;
  GHF_REQUIRE(mp_f16ss_supported(d), "mp_f16_ss: hidden_dim must be 64 or 256, got %d", d);
  GHF_REQUIRE(h16_scale != nullptr, "mp_f16_ss: the fp16 shadow needs its scale words");
  GHF_REQUIRE(g->unit_edges % kTileM == 0, "mp_f16_ss: unit_edges=%d must be a multiple of %d", g->unit_edges, kTileM);
  GHF_REQUIRE((reinterpret_cast<uintptr_t>(h16) | reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(bias) |
               reinterpret_cast<uintptr_t>(pack_scratch)) % 16 == 0,
              "mp_f16_ss: h16 / acc / bias / scratch must be 16-byte aligned");
  const uint8_t* img = reinterpret_cast<const uint8_t*>(pack_scratch);
  const float* inv = reinterpret_cast<const float*>(img + align_up((int64_t)g->num_rel * image_bytes(d), 256));
  const __half* h = reinterpret_cast<const __half*>(h16);
  return d == 256 ? launch_impl<256>(g, h, h16_scale, bias, acc, img, inv, unit_counter, stream)
                  : launch_impl<64>(g, h, h16_scale, bias, acc, img, inv, unit_counter, stream);
}

// scales for generator-written images (see image_scale_kernel): fills scale[R] and the inverse scales behind the
// images (the layout mp_f16ss_launch reads).  `words`: 8 floats of scratch.
int mp_f16ss_image_scales(const float* Zm, const float* Zs, int H, int64_t R, const float* W3m, const float* b3m,
                          const float* W3s, const float* b3s, int d, const float* ls_m, const float* ls_s,
                          float* words, float* scale, void* images, cudaStream_t stream) {
  GHF_REQUIRE(mp_f16ss_supported(d), "mp_f16_ss: hidden_dim must be 64 or 256, got %d", d);
  const int64_t n_w = (int64_t)d * d * H, n_b = (int64_t)d * d;
  if (int rc = mp_f16_absmax(W3m, n_w, words + 0, stream)) return rc;
  if (int rc = mp_f16_absmax(b3m, n_b, words + 2, stream)) return rc;
  if (int rc = mp_f16_absmax(W3s, n_w, words + 4, stream)) return rc;
  if (int rc = mp_f16_absmax(b3s, n_b, words + 6, stream)) return rc;
  float* inv = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(images) + align_up(R * image_bytes(d), 256));
  if (R == 0) return 0;
  image_scale_kernel<<<(unsigned)cdiv(R * 32, 256), 256, 0, stream>>>(Zm, Zs, H, R, words, ls_m, ls_s, scale, inv);
  GHF_LAUNCH_CHECK();
  return 0;
}

int64_t mp_f16ss_image_bytes(int d) { return image_bytes(d); }

}  // namespace ghf
