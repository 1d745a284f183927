// mp.cu — ghf_mp_layer: one relation-typed message-passing layer (HG:160-230 + HG:289-296).
//
//   acc_v  = sum_{e:(u->v), r} ( [h_u | h_v] @ [W_msg[r] ; W_self[r]] + bias[r] )      (contraction)
//   upd_v  = acc_v / max(indeg_v, 1)
//   out_v  = LayerNorm( relu(upd_v + h_v) )                                            (epilogue)
//
// The self-loop of the reference, h_v @ mean_e(W_self[r_e]) (HG:217-228), is the same sum by
// linearity, which is why it rides along as the second half of K.  Two contraction engines:
//   GHF_PREC_FP32  this file: CUDA-core FFMA tiles, exact fp32;
//   GHF_PREC_TF32  mp_umma.cu: tcgen05 kind::tf32 with TMEM accumulators.
#include <cuda_fp16.h>

#include <cstdlib>
#include <vector>

#include "ffma_gemm.cuh"
#include "ghf_b200.h"
#include "graph.cuh"
#include "mp.cuh"
#include "mp_fuse.cuh"

namespace ghf {
namespace {

template <int BN, bool VEC>
__global__ void __launch_bounds__(kFfmaThreads)
mp_fp32_kernel(const int32_t* __restrict__ unit_start, const int32_t* __restrict__ unit_count,
               const int32_t* __restrict__ unit_rel, int64_t num_units,
               const int32_t* __restrict__ src_sorted, const int32_t* __restrict__ dst_sorted,
               const float* __restrict__ h, int64_t dst_lo, int d, const float* __restrict__ W_msg,
               const float* __restrict__ W_self, const float* __restrict__ bias, float* __restrict__ acc_out) {
  __shared__ FfmaSmem<BN> sm;
  __shared__ int32_t s_src[kFfmaBM], s_dst[kFfmaBM];
  constexpr int TN = BN / 16;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  for (int64_t u = blockIdx.x; u < num_units; u += gridDim.x) {
    const int start = unit_start[u], count = unit_count[u];
    const int64_t r = unit_rel[u];
    const float* Wm = W_msg + r * d * d;
    const float* Ws = W_self + r * d * d;
    for (int t0 = 0; t0 < count; t0 += kFfmaBM) {
      const int rows = min(kFfmaBM, count - t0);
      __syncthreads();
      if (threadIdx.x < kFfmaBM) {
        const bool ok = (int)threadIdx.x < rows;
        s_src[threadIdx.x] = ok ? src_sorted[start + t0 + threadIdx.x] : -1;
        s_dst[threadIdx.x] = ok ? dst_sorted[start + t0 + threadIdx.x] : -1;
      }
      __syncthreads();
      for (int n0 = 0; n0 < d; n0 += BN) {
        float acc[8][TN];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

        // A row = [ h[src] | h[dst] ]  (K = 2d)
        auto rowptr = [&](int row, int k) -> const float* {
          const int s = s_src[row];
          if (s < 0 || k >= 2 * d) return nullptr;
          return k < d ? h + (int64_t)s * d + k : h + (dst_lo + s_dst[row]) * d + (k - d);
        };
        auto loadA = [&](int row, int k) -> float {
          const float* p = rowptr(row, k);
          return p ? *p : 0.f;
        };
        auto loadA4 = [&](int row, int k) -> float4 {
          const float* p = rowptr(row, k);
          return p ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // B row k = W_msg[r][k][:] for k < d, W_self[r][k-d][:] after
        auto colptr = [&](int k, int n) -> const float* {
          const int nn = n0 + n;
          if (nn >= d || k >= 2 * d) return nullptr;
          return (k < d ? Wm + (int64_t)k * d : Ws + (int64_t)(k - d) * d) + nn;
        };
        auto loadB = [&](int k, int n) -> float {
          const float* p = colptr(k, n);
          return p ? *p : 0.f;
        };
        auto loadB4 = [&](int k, int n) -> float4 {
          const float* p = colptr(k, n);
          return p ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        ffma_mainloop<BN, VEC, /*B_KMAJOR=*/false>(sm, 2 * d, loadA, loadA4, loadB, loadB4, acc);

#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = ffma_row(ty, i);
          if (row >= rows) continue;
          float* out = acc_out + (int64_t)s_dst[row] * d;
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            const int n = n0 + ffma_col<BN>(tx, j);
            if (n < d) atomicAdd(out + n, acc[i][j] + bias[r * d + n]);
          }
        }
      }
    }
  }
}

// One warp per destination row: upd = acc / max(indeg,1); x = relu(upd + h); LayerNorm(x).
// MAXV values per lane live in registers (d <= 32*MAXV); larger d re-reads the row.
constexpr int kLnMaxV = 8;
__global__ void __launch_bounds__(256)
mp_epilogue_kernel(const float* __restrict__ acc, const int32_t* __restrict__ indeg,
                   const float* __restrict__ h, int64_t dst_lo, int64_t num_local, int d,
                   const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                   float* __restrict__ out, float* __restrict__ upd, const DropoutArgs da) {
  const int lane = threadIdx.x & 31;
  const int64_t v = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (v >= num_local) return;
  const uint64_t e_row = (uint64_t)(dst_lo + v) * (uint64_t)d;   // index of the row's first element in [num_nodes, d]
  const float inv = 1.f / (float)max(indeg[v], 1);
  const float* a = acc + v * d;
  const float* hv = h + (dst_lo + v) * d;
  float* o = out + v * d;
  float* up = upd ? upd + v * d : nullptr;
  float x[kLnMaxV];
  float sum = 0.f;
  const bool in_regs = d <= 32 * kLnMaxV;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      x[i] = 0.f;
      if (c < d) {
        const float u = a[c] * inv;
        if (up) up[c] = u;
        x[i] = fmaxf(u + hv[c], 0.f);
        if (da.keep > 0.f) x[i] *= fuse::dropout_mult1(da, e_row + c);
        sum += x[i];
      }
    }
  } else {
    for (int c = lane; c < d; c += 32) {
      const float u = a[c] * inv;
      if (up) up[c] = u;
      float xv = fmaxf(u + hv[c], 0.f);
      if (da.keep > 0.f) xv *= fuse::dropout_mult1(da, e_row + c);
      o[c] = xv;  // parked in the output row, normalised in place below
      sum += xv;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  const float mean = sum / (float)d;
  float var = 0.f;
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i)
      if (lane + 32 * i < d) var += (x[i] - mean) * (x[i] - mean);
  } else {
    for (int c = lane; c < d; c += 32) var += (o[c] - mean) * (o[c] - mean);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
  const float rstd = rsqrtf(var / (float)d + eps);
  if (in_regs) {
#pragma unroll
    for (int i = 0; i < kLnMaxV; ++i) {
      const int c = lane + 32 * i;
      if (c < d) o[c] = (x[i] - mean) * rstd * ln_w[c] + ln_b[c];
    }
  } else {
    for (int c = lane; c < d; c += 32) o[c] = (o[c] - mean) * rstd * ln_w[c] + ln_b[c];
  }
}

// The same epilogue for hidden_dim 32 / 64 / 128: a lane owns D/32 CONSECUTIVE columns (one vector access per row
// and operand), four rows in flight per warp, and an optional fp16 copy of the output row for the next layer's
// gathers (GHF_PREC_F16).
template <int D, bool DROP>
__global__ void __launch_bounds__(256)
mp_epilogue_vec_kernel(const float* __restrict__ acc, const int32_t* __restrict__ indeg,
                       const float* __restrict__ h, int64_t dst_lo, int64_t num_local,
                       const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                       float* __restrict__ out, float* __restrict__ upd, __half* __restrict__ out16,
                       float* __restrict__ out16_scale, const int32_t* __restrict__ det_words, const PeerPush push,
                       const DropoutArgs da) {
  using namespace fuse;
  const int det_eB = det_words ? det_words[2] : 0;     // deterministic mode: acc holds int32 fixed point (mp.cuh)
  constexpr int V = D / 32;
  constexpr int kRows = 4;
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float lw[V], lb[V];
  vload_nc<V>(lw, ln_w + lane * V);
  vload_nc<V>(lb, ln_b + lane * V);
  // fp16 shadow of the output: |LayerNorm(x)_c| <= sqrt(D-1) |w_c| + |b_c|, so a power-of-two scale that fits the
  // largest such bound fits every row - chosen here, identically by every warp (and every rank), from w and b alone
  float s16 = 1.f;
  if (out16) {
    float bound = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) bound = fmaxf(bound, sqrtf((float)(D - 1)) * fabsf(lw[j]) + fabsf(lb[j]));
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, s));
    s16 = f16_scale_for(bound);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      out16_scale[0] = 1.f / s16;
      out16_scale[1] = bound;
    }
  }
  // rows are walked from the END: the accumulator rows the contraction touched last are still in L2
  for (int64_t r0 = w0; r0 < num_local; r0 += kRows * warps) {
    float a[kRows][V], hv[kRows][V];
    int deg[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = num_local - 1 - (r0 + k * warps);
      if (r >= 0) {
        vload_cg<V>(a[k], acc + r * D + lane * V);
        vload_nc<V>(hv[k], h + (dst_lo + r) * D + lane * V);
        deg[k] = __ldg(indeg + r);
      }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
      const int64_t r = num_local - 1 - (r0 + k * warps);
      if (r < 0) break;
      float inv = 1.f / (float)max(deg[k], 1);
      if (det_words) {                                   // integers * 2^-k_v, then the mean
        const int kv = 30 - (32 - __clz(max(deg[k], 1))) - det_eB;
        const float unscale = __int_as_float((uint32_t)(127 - max(-126, min(127, kv))) << 23);
#pragma unroll
        for (int j = 0; j < V; ++j) a[k][j] = (float)__float_as_int(a[k][j]) * unscale;
      }
      float x[V], u[V], sum = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        u[j] = a[k][j] * inv;
        x[j] = fmaxf(u[j] + hv[k][j], 0.f);
      }
      if constexpr (DROP) {                              // training-mode dropout (HG:293-294), torch's Philox stream
        float m[V];
        dropout_multv<V>(da, (uint64_t)(dst_lo + r) * D + lane * V, m);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] *= m[j];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) sum += x[j];
      if (upd) vstore<V>(upd + r * D + lane * V, u);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
      const float mean = sum / (float)D;
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < V; ++j) var += (x[j] - mean) * (x[j] - mean);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) var += __shfl_xor_sync(0xffffffffu, var, s);
      const float rstd = rsqrtf(var / (float)D + eps);
      float y[V];
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = (x[j] - mean) * rstd * lw[j] + lb[j];
      vstore<V>(out + r * D + lane * V, y);
      if (out16) {
        __half* o = out16 + r * D + lane * V;
        if constexpr (V == 4) {
          const __half2 p0 = __floats2half2_rn(y[0] * s16, y[1] * s16), p1 = __floats2half2_rn(y[2] * s16, y[3] * s16);
          const uint2 packed = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
          *reinterpret_cast<uint2*>(o) = packed;
          if (push.mask)
            for (int q = 0; q < push.world; ++q)
              if (q != push.me && push.mask[q * push.mask_stride + r])
                *reinterpret_cast<uint2*>(push.tables[q] + (push.table_row0 + r) * D + lane * V) = packed;
        } else if constexpr (V == 2) {
          const __half2 packed = __floats2half2_rn(y[0] * s16, y[1] * s16);
          *reinterpret_cast<__half2*>(o) = packed;
          if (push.mask)
            for (int q = 0; q < push.world; ++q)
              if (q != push.me && push.mask[q * push.mask_stride + r])
                *reinterpret_cast<__half2*>(push.tables[q] + (push.table_row0 + r) * D + lane * V) = packed;
        } else {
          *o = __float2half_rn(y[0] * s16);
        }
      }
    }
  }
}

template <int BN>
int launch_mp_fp32(const ghf_graph* g, const float* h, const float* W_msg, const float* W_self,
                   const float* bias, float* acc, cudaStream_t stream) {
  const int d = g->hidden_dim;
  const bool vec = (d % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(W_msg) |
                     reinterpret_cast<uintptr_t>(W_self)) % 16 == 0);
  const int64_t grid = g->num_units < (int64_t)sm_count() * 4 ? g->num_units : (int64_t)sm_count() * 4;
  if (vec)
    mp_fp32_kernel<BN, true><<<(unsigned)grid, kFfmaThreads, 0, stream>>>(
        g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted, h, g->dst_lo, d,
        W_msg, W_self, bias, acc);
  else
    mp_fp32_kernel<BN, false><<<(unsigned)grid, kFfmaThreads, 0, stream>>>(
        g->unit_start, g->unit_count, g->unit_rel, g->num_units, g->src_sorted, g->dst_sorted, h, g->dst_lo, d,
        W_msg, W_self, bias, acc);
  GHF_LAUNCH_CHECK();
  return 0;
}

}  // namespace
}  // namespace ghf

using namespace ghf;

static bool mp_ts_enabled(int d) {  // weights in tensor memory (hidden_dim 128) unless GHF_MP_TS=0
  const char* env = getenv("GHF_MP_TS");
  return mp_ts_supported(d) && !(env && env[0] == '0');
}

// ---- optional per-kernel timing (bench.py roofline): event triples per layer call
namespace {
struct ProfRec { cudaEvent_t e[4]; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace

extern "C" int ghf_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}

extern "C" int ghf_profile_read(double ms[3], int64_t* launches) {
  GHF_REQUIRE(ms != nullptr, "ghf_profile_read: NULL");
  ms[0] = ms[1] = ms[2] = 0.0;
  for (auto& r : g_prof) {
    GHF_CUDA(cudaEventSynchronize(r.e[3]));
    float a = 0, b = 0, c = 0;
    GHF_CUDA(cudaEventElapsedTime(&c, r.e[0], r.e[1]));
    GHF_CUDA(cudaEventElapsedTime(&a, r.e[1], r.e[2]));
    GHF_CUDA(cudaEventElapsedTime(&b, r.e[2], r.e[3]));
    ms[0] += a; ms[1] += b; ms[2] += c;
    for (auto& e : r.e) cudaEventDestroy(e);
  }
  if (launches) *launches = (int64_t)g_prof.size();
  g_prof.clear();
  return 0;
}

// accumulator region of the workspace: [local rows, d] floats, or the fused kernel's ring when that is larger
static int64_t acc_region_bytes(const ghf_graph* g, int d, int precision) {
  int64_t rows = g->num_local;
  if (precision == GHF_PREC_F16 && mp_f16_supported(d) && mp_f16_fused_ring_rows(g) > rows) rows = mp_f16_fused_ring_rows(g);
  return align_up(rows * (int64_t)d * 4, 256);
}

extern "C" int64_t ghf_mp_workspace_bytes(const ghf_graph* g, int32_t hidden_dim, int precision) {
  if (!g) return -1;
  int64_t bytes = 256 /* work counter */ + acc_region_bytes(g, hidden_dim, precision);
  if (precision == GHF_PREC_TF32) bytes += mp_umma_pack_bytes(g->num_rel, hidden_dim);
  if (precision == GHF_PREC_F16)  // sync words, weight images, the fp16 copy of h made when the caller passes none
    bytes += mp_f16_sync_bytes(g) +
             (mp_f16ss_supported(hidden_dim) ? mp_f16ss_pack_bytes(g->num_rel, hidden_dim) : mp_f16_pack_bytes(g->num_rel)) +
             256 /* scale words */ + align_up(g->num_nodes * (int64_t)hidden_dim * 2, 256);
  return bytes + 256;
}

// rows [r0, r1) of the graph's local rows (all of them by default); acc / out / upd / out16 are indexed by local row
static int launch_epilogue(const ghf_graph* g_full, const float* acc, const float* d_h, const float* d_ln_w,
                           const float* d_ln_b, float eps, float* d_out, float* d_upd, void* d_out16,
                           float* d_out16_scale, cudaStream_t stream, int64_t r0 = 0, int64_t r1 = -1,
                           const int32_t* det_words = nullptr, PeerPush push = PeerPush(),
                           DropoutArgs da = DropoutArgs()) {
  const int d = g_full->hidden_dim;
  if (r1 < 0) r1 = g_full->num_local;
  if (r1 <= r0) return 0;
  ghf_graph view = *g_full;                              // the same tables, seen from row r0 on
  view.indeg = g_full->indeg + r0;
  view.dst_lo = g_full->dst_lo + r0;
  view.num_local = r1 - r0;
  const ghf_graph* g = &view;
  acc += r0 * d;
  d_out += r0 * d;
  if (d_upd) d_upd += r0 * d;
  if (d_out16) d_out16 = reinterpret_cast<char*>(d_out16) + r0 * d * 2;
  if (push.mask) {
    GHF_REQUIRE(d_out16 != nullptr && (d == 64 || d == 128), "ghf_mp_layer: the peer push needs the fp16 output, hidden 64/128");
    push.mask += r0;
    push.table_row0 = g->dst_lo;                         // the view's first row
  }
  const int64_t nl = g->num_local;
  const int threads = 256;
  const bool aligned = (reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(d_h) |
                        reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_upd) |
                        reinterpret_cast<uintptr_t>(d_ln_w) | reinterpret_cast<uintptr_t>(d_ln_b) |
                        reinterpret_cast<uintptr_t>(d_out16)) % 16 == 0;
  if (aligned && (d == 32 || d == 64 || d == 128)) {
    const int64_t want = cdiv(nl, (threads / 32) * 4);
    const int64_t cap = (int64_t)sm_count() * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    __half* o16 = reinterpret_cast<__half*>(d_out16);
#define GHF_EPI_VEC(D, DROP)                                                                                       \
  mp_epilogue_vec_kernel<D, DROP><<<grid, threads, 0, stream>>>(acc, g->indeg, d_h, g->dst_lo, nl, d_ln_w, d_ln_b, eps, \
                                                                d_out, d_upd, o16, d_out16_scale, det_words, push, da)
    const bool drop = da.keep > 0.f;   // a separate instantiation: the inference kernel keeps its registers
    if (d == 32) { if (drop) GHF_EPI_VEC(32, true); else GHF_EPI_VEC(32, false); }
    else if (d == 64) { if (drop) GHF_EPI_VEC(64, true); else GHF_EPI_VEC(64, false); }
    else { if (drop) GHF_EPI_VEC(128, true); else GHF_EPI_VEC(128, false); }
#undef GHF_EPI_VEC
  } else {
    GHF_REQUIRE(push.mask == nullptr, "ghf_mp_layer: the peer push needs hidden_dim 64/128 and aligned buffers");
    GHF_REQUIRE(det_words == nullptr, "ghf_mp_layer: the deterministic mode needs hidden_dim 128 and aligned buffers");
    GHF_REQUIRE(d_out16 == nullptr, "ghf_mp_layer: fp16 output needs hidden_dim 32/64/128 and 16-byte alignment");
    mp_epilogue_kernel<<<(unsigned)cdiv(nl * 32, threads), threads, 0, stream>>>(
        acc, g->indeg, d_h, g->dst_lo, nl, d, d_ln_w, d_ln_b, eps, d_out, d_upd, da);
  }
  GHF_LAUNCH_CHECK();
  return 0;
}

// what the fused layer kernel (mp_f16_fused.cu) needs beyond the contraction's own arguments
struct FusedEpilogue {
  const float *h, *ln_w, *ln_b;
  float eps;
  float *out, *upd;
  void* out16;
  float* out16_scale;
  const PeerPush* push = nullptr;   // multi-GPU: rows also go to the peers' tables (either kernel can do it)
};

// The contraction of one layer: acc[v] = sum over in-edges of [h_u | h_v] @ [W_msg[r] ; W_self[r]] + bias[r].
// `acc_ext` (optional) receives the sums; otherwise they stay in the workspace for the epilogue.  -> *acc_used.
static int run_contraction(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                           const float* d_W_msg, const float* d_W_self, const float* d_bias, int precision,
                           float* acc_ext, bool accumulate, bool transposed, void* d_workspace, cudaStream_t stream,
                           float** acc_used, ProfRec* rec, const void* prepacked = nullptr,
                           const FusedEpilogue* fe = nullptr, bool* fused_done = nullptr, int phase_lo = 0,
                           int phase_hi = -1, const int32_t** det_words_out = nullptr) {
  // A range of super-blocks [phase_lo, phase_hi): the units of those super-blocks are contiguous, so the kernels see
  // a view of the graph that starts at the first of them; accumulator rows keep their local row index.
  const ghf_graph* g_whole = g;
  ghf_graph view = *g;
  if (phase_hi < 0) phase_hi = (int)g->num_phases;
  const bool ranged = phase_lo != 0 || phase_hi != (int)g->num_phases;
  if (ranged) {
    GHF_REQUIRE(0 <= phase_lo && phase_lo <= phase_hi && phase_hi <= g->num_phases && g->h_phase_unit_begin,
                "ghf_mp_layer: bad super-block range [%d, %d) of %lld", phase_lo, phase_hi, (long long)g->num_phases);
    const int64_t u0 = g->h_phase_unit_begin[phase_lo], u1 = g->h_phase_unit_begin[phase_hi];
    view.unit_start += u0; view.unit_count += u0; view.unit_rel += u0; view.unit_phase += u0;
    view.num_units = u1 - u0;
    g = &view;
  }
  const int64_t row0 = (int64_t)phase_lo * g_whole->sb_nodes;
  const int64_t row1 = phase_hi * (int64_t)g_whole->sb_nodes < g_whole->num_local ? phase_hi * (int64_t)g_whole->sb_nodes
                                                                                    : g_whole->num_local;
  const int d = g->hidden_dim;
  const int64_t nl = g->num_local;
  // workspace: [work counter, 256 B][accumulator rows][operand images (tensor-core paths)][fp16 h (f16 path)]
  // (the f16 kernel clears the accumulator itself and keeps per-phase sync words next to the counter)
  const bool f16_ss = precision == GHF_PREC_F16 && mp_f16ss_supported(d);   // hidden 256: streamed weights
  // GHF_PUSH_FUSED=1: with a peer push, use the fused kernel - its row epilogue runs per super-block WHILE later
  // super-blocks are still being contracted, so the NVLink stores overlap the contraction.  Measured equal to the
  // separate epilogue kernel at c3 on 8 GPUs (3.82 ms both, profiles/r02_multi_gpu_8.txt), so it stays opt-in.
  const char* pf_env = getenv("GHF_PUSH_FUSED");
  const bool push_fused = fe != nullptr && fe->push != nullptr && fe->push->mask != nullptr && g->num_units > 0 &&
                          pf_env && pf_env[0] == '1';
  const bool fused = fe != nullptr && !ranged && precision == GHF_PREC_F16 && !f16_ss && mp_f16_supported(d) &&
                     (mp_f16_fused_enabled(g) || push_fused) && fe->ln_w != nullptr && fe->ln_b != nullptr &&
                     (reinterpret_cast<uintptr_t>(fe->ln_w) | reinterpret_cast<uintptr_t>(fe->ln_b) |
                      reinterpret_cast<uintptr_t>(fe->h) | reinterpret_cast<uintptr_t>(fe->out) |
                      reinterpret_cast<uintptr_t>(fe->upd) | reinterpret_cast<uintptr_t>(fe->out16)) % 16 == 0;
  const bool self_clearing = precision == GHF_PREC_F16 && !f16_ss && g->num_units > 0;
  int* counter = reinterpret_cast<int*>(align_up(reinterpret_cast<int64_t>(d_workspace), 256));
  float* acc_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(counter) +
                                           (precision == GHF_PREC_F16 ? mp_f16_sync_bytes(g_whole) : 256));
  const int64_t acc_bytes = acc_region_bytes(g_whole, d, precision);
  void* pack = reinterpret_cast<char*>(acc_ws) + acc_bytes;
  float* acc = acc_ext ? acc_ext : acc_ws;
  *acc_used = acc;
  if (rec) GHF_CUDA(cudaEventRecord(rec->e[0], stream));
  if (!self_clearing) {
    const size_t head = reinterpret_cast<char*>(acc_ws) - reinterpret_cast<char*>(counter);
    GHF_CUDA(cudaMemsetAsync(counter, 0, head, stream));
    if (!(acc_ext && accumulate) && row1 > row0)         // only the rows of the super-blocks in range
      GHF_CUDA(cudaMemsetAsync((acc_ext ? acc_ext : acc_ws) + row0 * d, 0, (size_t)(row1 - row0) * d * 4, stream));
  }
  // gradient contractions (ghf_mp_contract): transposed relation matrices and an absent (NULL) half are understood
  // by the f16 engine only; the host side materialises them for the other engines
  if (prepacked) pack = const_cast<void*>(prepacked);   // operand images built by the generator (hidden 64 / 256)
  const bool plain = prepacked || (!transposed && d_W_msg != nullptr && d_W_self != nullptr && d_bias != nullptr);
  GHF_REQUIRE(plain || (precision == GHF_PREC_F16 && mp_f16_supported(d) && (d_W_msg != nullptr || d_W_self != nullptr)),
              "ghf_mp_contract: transposed / NULL weight tensors need precision f16 and hidden_dim 128");
  const int skip_half = prepacked ? 0 : (d_W_msg == nullptr ? 1 : (d_W_self == nullptr ? 2 : 0));
  const bool ts = mp_ts_enabled(d);
  const void* h16 = d_h16;
  const float* h16_scale = d_h16_scale;
  if (precision == GHF_PREC_TF32 && g->num_units > 0) {
    GHF_REQUIRE(mp_umma_supported(d), "ghf_mp_layer: tf32 path supports hidden_dim in {32,64,128}, got %d", d);
    if (int rc = ts ? mp_ts_pack(g, d_W_msg, d_W_self, pack, stream) : mp_umma_pack(g, d_W_msg, d_W_self, pack, stream))
      return rc;
  } else if (precision == GHF_PREC_F16) {
    GHF_REQUIRE(mp_f16_supported(d) || f16_ss, "ghf_mp_layer: f16 path supports hidden_dim 64, 128 and 256, got %d", d);
    if (g->num_units > 0) {
      if (!prepacked)
        if (int rc = f16_ss ? mp_f16ss_pack(g, d_W_msg, d_W_self, pack, stream)
                            : mp_f16_pack(g, d_W_msg, d_W_self, pack, stream, transposed))
          return rc;
      if (h16 == nullptr) {  // no fp16 shadow of h from the previous layer: make one (scale words, then the rows)
        float* sc = reinterpret_cast<float*>(reinterpret_cast<char*>(acc_ws) + acc_bytes +
                                             (f16_ss ? mp_f16ss_pack_bytes(g->num_rel, d) : mp_f16_pack_bytes(g->num_rel)));
        void* conv = reinterpret_cast<char*>(sc) + 256;
        if (int rc = mp_f16_absmax(d_h, g->num_nodes * (int64_t)d, sc, stream)) return rc;
        if (int rc = mp_f16_convert(d_h, g->num_nodes * (int64_t)d, conv, sc, /*rescue=*/false, stream)) return rc;
        h16 = conv;
        h16_scale = sc;
      }
    }
  }
  if (rec) GHF_CUDA(cudaEventRecord(rec->e[1], stream));
  if (g->num_units > 0) {
    int rc;
    if (precision == GHF_PREC_TF32) {
      rc = ts ? mp_ts_launch(g, d_h, d_bias, acc, pack, counter, stream)
              : mp_umma_launch(g, d_h, d_bias, acc, pack, counter, stream);
    } else if (f16_ss) {
      rc = mp_f16ss_launch(g, h16, h16_scale, d_bias, acc, pack, counter, stream);
    } else if (fused && !(getenv("GHF_DETERMINISTIC") && getenv("GHF_DETERMINISTIC")[0] == '1')) {
      // contraction + row epilogue in one kernel; the accumulator region is its ring
      rc = mp_f16_fused_launch(g, h16, h16_scale, d_bias, pack, acc_ws, counter, fe->h, fe->ln_w, fe->ln_b, fe->eps,
                               fe->out, fe->upd, fe->out16, fe->out16_scale, stream, fe->push);
      if (fused_done) *fused_done = true;
    } else if (precision == GHF_PREC_F16) {
      // GHF_DETERMINISTIC=1 (layer entries only): fixed-point per-destination sums, bit-identical from run to run
      const char* denv = getenv("GHF_DETERMINISTIC");
      const bool det = det_words_out != nullptr && denv && denv[0] == '1';
      GHF_REQUIRE(!det || (d_W_msg && d_W_self && !transposed && !accumulate && !prepacked),
                  "GHF_DETERMINISTIC=1 needs the fp32 relation matrices of a plain layer");
      rc = mp_f16_launch(g, h16, h16_scale, d_bias, acc, pack, counter, stream, accumulate, skip_half, phase_lo,
                         phase_hi, det ? d_W_msg : nullptr, det ? d_W_self : nullptr);
      if (det) *det_words_out = mp_f16_det_words(counter);
    } else if (d <= 32) {
      rc = launch_mp_fp32<32>(g, d_h, d_W_msg, d_W_self, d_bias, acc, stream);
    } else if (d <= 64) {
      rc = launch_mp_fp32<64>(g, d_h, d_W_msg, d_W_self, d_bias, acc, stream);
    } else {
      rc = launch_mp_fp32<128>(g, d_h, d_W_msg, d_W_self, d_bias, acc, stream);
    }
    if (rc) return rc;
  }
  if (rec) GHF_CUDA(cudaEventRecord(rec->e[2], stream));
  return 0;
}

static int check_layer_args(const ghf_graph* g, const void* d_workspace, int precision, const void* d_h16,
                            const float* d_h16_scale) {
  GHF_REQUIRE(g != nullptr, "ghf_mp_layer: graph is NULL");
  GHF_REQUIRE(d_workspace != nullptr, "ghf_mp_layer: workspace is NULL");
  GHF_REQUIRE(precision == GHF_PREC_FP32 || precision == GHF_PREC_TF32 || precision == GHF_PREC_F16,
              "ghf_mp_layer: precision=%d", precision);
  GHF_REQUIRE(d_h16 == nullptr || d_h16_scale != nullptr, "ghf_mp_layer: d_h16 needs d_h16_scale (float[2])");
  return 0;
}

static int mp_layer_impl(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                         const float* d_W_msg, const float* d_W_self, const float* d_bias, const float* d_ln_w,
                         const float* d_ln_b, float eps, int precision, float* d_out, void* d_out16,
                         float* d_out16_scale, float* d_upd, void* d_workspace, int phase_lo, int phase_hi,
                         void* stream_, PeerPush push = PeerPush(), DropoutArgs da = DropoutArgs()) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_layer_args(g, d_workspace, precision, d_h16, d_h16_scale)) return rc;
  GHF_REQUIRE(d_out16 == nullptr || d_out16_scale != nullptr, "ghf_mp_layer: d_out16 needs d_out16_scale (float[2])");
  g->stream = stream_;
  if (g->num_local == 0) return 0;
  if (phase_hi < 0) phase_hi = (int)g->num_phases;
  GHF_REQUIRE(0 <= phase_lo && phase_lo <= phase_hi && phase_hi <= g->num_phases,
              "ghf_mp_layer: bad super-block range [%d, %d) of %lld", phase_lo, phase_hi, (long long)g->num_phases);
  if (phase_lo == phase_hi) return 0;
  ProfRec rec{};
  const bool prof = g_prof_on;
  if (prof)
    for (auto& e : rec.e) GHF_CUDA(cudaEventCreate(&e));
  float* acc = nullptr;
  const int32_t* det_words = nullptr;
  push.table_row0 = g->dst_lo;
  const FusedEpilogue fe{d_h, d_ln_w, d_ln_b, eps, d_out, d_upd, d_out16, d_out16_scale, push.mask ? &push : nullptr};
  bool fused_done = false;
  if (int rc = run_contraction(g, d_h, d_h16, d_h16_scale, d_W_msg, d_W_self, d_bias, precision, nullptr,
                               false, false, d_workspace, stream, &acc, prof ? &rec : nullptr, nullptr,
                               da.keep > 0.f ? nullptr : &fe,   // dropout lives in the separate row epilogue only
                               &fused_done, phase_lo, phase_hi, &det_words))
    return rc;
  if (!fused_done) {
    const int64_t r0 = (int64_t)phase_lo * g->sb_nodes;
    const int64_t r1 = phase_hi * (int64_t)g->sb_nodes < g->num_local ? phase_hi * (int64_t)g->sb_nodes : g->num_local;
    if (int rc = launch_epilogue(g, acc, d_h, d_ln_w, d_ln_b, eps, d_out, d_upd, d_out16, d_out16_scale, stream, r0, r1,
                                 det_words, push, da))
      return rc;
  }
  if (prof) {
    GHF_CUDA(cudaEventRecord(rec.e[3], stream));
    g_prof.push_back(rec);
  }
  return 0;
}

extern "C" int ghf_mp_layer_f16(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                                const float* d_W_msg, const float* d_W_self, const float* d_bias,
                                const float* d_ln_w, const float* d_ln_b, float eps, int precision, float* d_out,
                                void* d_out16, float* d_out16_scale, float* d_upd, void* d_workspace,
                                void* stream_) {
  return mp_layer_impl(g, d_h, d_h16, d_h16_scale, d_W_msg, d_W_self, d_bias, d_ln_w, d_ln_b, eps, precision, d_out,
                       d_out16, d_out16_scale, d_upd, d_workspace, 0, -1, stream_);
}

extern "C" int ghf_mp_layer_f16_range(const ghf_graph* g, const float* d_h, const void* d_h16,
                                      const float* d_h16_scale, const float* d_W_msg, const float* d_W_self,
                                      const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                                      int precision, float* d_out, void* d_out16, float* d_out16_scale, float* d_upd,
                                      void* d_workspace, int32_t phase_lo, int32_t phase_hi, void* stream_) {
  return mp_layer_impl(g, d_h, d_h16, d_h16_scale, d_W_msg, d_W_self, d_bias, d_ln_w, d_ln_b, eps, precision, d_out,
                       d_out16, d_out16_scale, d_upd, d_workspace, phase_lo, phase_hi, stream_);
}

extern "C" int ghf_mp_layer_f16_push(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                                     const float* d_W_msg, const float* d_W_self, const float* d_bias,
                                     const float* d_ln_w, const float* d_ln_b, float eps, int precision, float* d_out,
                                     void* d_out16, float* d_out16_scale, float* d_upd, void* d_workspace,
                                     int32_t phase_lo, int32_t phase_hi, const uint8_t* d_peer_mask,
                                     int64_t mask_stride, void* const* d_peer_tables, int32_t world, int32_t rank,
                                     void* stream_) {
  GHF_REQUIRE(g != nullptr, "ghf_mp_layer_f16_push: graph is NULL");
  GHF_REQUIRE(d_peer_mask && d_peer_tables && world >= 1 && rank >= 0 && rank < world && mask_stride >= g->num_local,
              "ghf_mp_layer_f16_push: bad peer arguments");
  PeerPush push;
  push.mask = d_peer_mask;
  push.mask_stride = mask_stride;
  push.tables = reinterpret_cast<__half* const*>(d_peer_tables);
  push.world = world;
  push.me = rank;
  return mp_layer_impl(g, d_h, d_h16, d_h16_scale, d_W_msg, d_W_self, d_bias, d_ln_w, d_ln_b, eps, precision, d_out,
                       d_out16, d_out16_scale, d_upd, d_workspace, phase_lo, phase_hi, stream_, push);
}

namespace ghf {
int make_dropout_args(float p, uint64_t seed, uint64_t offset, int64_t numel, DropoutArgs* out, int64_t* advance) {
  GHF_REQUIRE(p > 0.f && p < 1.f, "dropout probability must lie in (0, 1), got %g", (double)p);
  GHF_REQUIRE(numel > 0 && numel % 4 == 0 && offset % 4 == 0,
              "native dropout follows torch's vectorised kernel: numel %% 4 == 0 and a generator offset %% 4 == 0");
  int dev = 0, max_threads = 0;
  GHF_CUDA(cudaGetDevice(&dev));
  GHF_CUDA(cudaDeviceGetAttribute(&max_threads, cudaDevAttrMaxThreadsPerMultiProcessor, dev));
  // torch's launch for this tensor: blocks of 256 threads, min(SMs * (max threads per SM / 256), ceil(numel / 256))
  const int64_t blocks_full = (int64_t)sm_count() * (max_threads / 256);
  const int64_t blocks = blocks_full < cdiv(numel, 256) ? blocks_full : cdiv(numel, 256);
  const int64_t threads = 256 * blocks;
  if (out) {
    const double keep = 1.0 - (double)p;                 // torch: p1m = 1. - p (double), compared and inverted as float
    out->keep = (float)keep;
    out->scale = (float)(1.0 / (double)out->keep);
    out->key0 = (uint32_t)seed;
    out->key1 = (uint32_t)(seed >> 32);
    out->ctr0 = offset / 4;
    out->threads = (uint32_t)threads;
  }
  if (advance) *advance = 4 * cdiv(numel, 4 * threads);  // ((numel - 1) / (256 * blocks * 4) + 1) * 4
  return 0;
}
}  // namespace ghf

extern "C" int64_t ghf_dropout_offset_advance(int64_t numel) {
  int64_t adv = -1;
  if (numel <= 0 || numel % 4 != 0) return -1;
  if (make_dropout_args(0.5f, 0, 0, numel, nullptr, &adv)) return -1;
  return adv;
}

extern "C" int ghf_mp_layer_dropout(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                                    const float* d_W_msg, const float* d_W_self, const float* d_bias,
                                    const float* d_ln_w, const float* d_ln_b, float eps, int precision, float p_drop,
                                    uint64_t seed, uint64_t offset, float* d_out, void* d_out16, float* d_out16_scale,
                                    float* d_upd, void* d_workspace, void* stream_) {
  GHF_REQUIRE(g != nullptr, "ghf_mp_layer_dropout: graph is NULL");
  DropoutArgs da;
  if (int rc = make_dropout_args(p_drop, seed, offset, g->num_nodes * (int64_t)g->hidden_dim, &da, nullptr)) return rc;
  return mp_layer_impl(g, d_h, d_h16, d_h16_scale, d_W_msg, d_W_self, d_bias, d_ln_w, d_ln_b, eps, precision, d_out,
                       d_out16, d_out16_scale, d_upd, d_workspace, 0, -1, stream_, PeerPush(), da);
}

static __global__ void mark_rows_kernel(const int64_t* __restrict__ ids, const uint32_t* __restrict__ subset, int64_t n,
                                 int64_t num_rows, uint8_t* __restrict__ mask) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t v = ids[subset ? subset[j] : j];
  if (v >= 0 && v < num_rows) mask[v] = 1;               // every writer stores the same byte
}

extern "C" int ghf_mark_rows(const int64_t* d_ids, const uint32_t* d_subset, int64_t n, int64_t num_rows,
                             uint8_t* d_mask, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(n >= 0 && num_rows >= 0 && (n == 0 || (d_ids && d_mask)), "ghf_mark_rows: bad arguments");
  GHF_CUDA(cudaMemsetAsync(d_mask, 0, (size_t)num_rows, stream));
  if (n == 0) return 0;
  mark_rows_kernel<<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(d_ids, d_subset, n, num_rows, d_mask);
  GHF_LAUNCH_CHECK();
  return 0;
}

namespace ghf {
int mp_layer_prepacked(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                       const void* images, const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                       float* d_out, void* d_out16, float* d_out16_scale, void* d_workspace, cudaStream_t stream,
                       int phase_lo, int phase_hi) {
  if (int rc = check_layer_args(g, d_workspace, GHF_PREC_F16, d_h16, d_h16_scale)) return rc;
  GHF_REQUIRE(images != nullptr && (mp_f16ss_supported(g->hidden_dim) || mp_f16_supported(g->hidden_dim)),
              "mp_layer_prepacked: hidden_dim 64 / 128 / 256 only");
  g->stream = stream;
  if (g->num_local == 0) return 0;
  if (phase_hi < 0) phase_hi = (int)g->num_phases;
  GHF_REQUIRE(0 <= phase_lo && phase_lo <= phase_hi && phase_hi <= g->num_phases,
              "mp_layer_prepacked: bad super-block range [%d, %d) of %lld", phase_lo, phase_hi, (long long)g->num_phases);
  if (phase_lo == phase_hi) return 0;

  ProfRec rec{};
  const bool prof = g_prof_on;
  if (prof)
    for (auto& e : rec.e) GHF_CUDA(cudaEventCreate(&e));
  float* acc = nullptr;
  const FusedEpilogue fe{d_h, d_ln_w, d_ln_b, eps, d_out, nullptr, d_out16, d_out16_scale};
  bool fused_done = false;
  if (int rc = run_contraction(g, d_h, d_h16, d_h16_scale, nullptr, nullptr, d_bias, GHF_PREC_F16, nullptr, false,
                               false, d_workspace, stream, &acc, prof ? &rec : nullptr, images, &fe, &fused_done,
                               phase_lo, phase_hi))
    return rc;
  if (!fused_done) {
    const int64_t r0 = (int64_t)phase_lo * g->sb_nodes;
    const int64_t r1 = phase_hi * (int64_t)g->sb_nodes < g->num_local ? phase_hi * (int64_t)g->sb_nodes : g->num_local;
    if (int rc = launch_epilogue(g, acc, d_h, d_ln_w, d_ln_b, eps, d_out, nullptr, d_out16, d_out16_scale, stream, r0, r1))
      return rc;
  }
  if (prof) {
    GHF_CUDA(cudaEventRecord(rec.e[3], stream));
    g_prof.push_back(rec);
  }
  return 0;
}
}  // namespace ghf

extern "C" int ghf_mp_contract(const ghf_graph* g, const float* d_x, const void* d_x16, const float* d_x16_scale,
                               const float* d_W_msg, const float* d_W_self, const float* d_bias, int precision,
                               float* d_acc, int accumulate, int transposed, void* d_workspace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_layer_args(g, d_workspace, precision, d_x16, d_x16_scale)) return rc;
  GHF_REQUIRE(d_acc != nullptr && reinterpret_cast<uintptr_t>(d_acc) % 256 == 0,
              "ghf_mp_contract: d_acc must be a 256-byte aligned [local nodes, hidden] buffer");
  g->stream = stream_;
  if (g->num_local == 0) return 0;
  float* acc = nullptr;
  return run_contraction(g, d_x, d_x16, d_x16_scale, d_W_msg, d_W_self, d_bias, precision, d_acc, accumulate != 0,
                         transposed != 0, d_workspace, stream, &acc, nullptr);
}

extern "C" int ghf_absmax(const float* d_x, int64_t elems, float* d_scale, void* stream_) {
  GHF_REQUIRE(elems >= 0 && (d_x != nullptr || elems == 0) && d_scale != nullptr, "ghf_absmax: bad arguments");
  return mp_f16_absmax(d_x, elems, d_scale, (cudaStream_t)stream_);
}

extern "C" int ghf_convert_f16(const float* d_x, int64_t elems, void* d_y16, float* d_scale, int have_amax,
                               void* stream_) {
  GHF_REQUIRE(elems >= 0 && (d_x != nullptr || elems == 0) && (d_y16 != nullptr || elems == 0) && d_scale != nullptr,
              "ghf_convert_f16: bad arguments");
  if (!have_amax)
    if (int rc = mp_f16_absmax(d_x, elems, d_scale, (cudaStream_t)stream_)) return rc;
  return mp_f16_convert(d_x, elems, d_y16, d_scale, /*rescue=*/false, (cudaStream_t)stream_);
}

extern "C" int ghf_mp_layer(const ghf_graph* g, const float* d_h, const float* d_W_msg, const float* d_W_self,
                            const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                            int precision, float* d_out, float* d_upd, void* d_workspace, void* stream_) {
  return ghf_mp_layer_f16(g, d_h, nullptr, nullptr, d_W_msg, d_W_self, d_bias, d_ln_w, d_ln_b, eps, precision, d_out,
                          nullptr, nullptr, d_upd, d_workspace, stream_);
}
