// api.cu — library-wide state and ghf_hypergnn_forward_host, the end-to-end entry point that takes
// HOST buffers (the shape of the reference's HyperGNN.forward, HG:236-298, at a C boundary).
#include <array>
#include <functional>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ghf_b200.h"
#include "graph.cuh"
#include "mp.cuh"

namespace ghf {

std::atomic<int64_t> g_launches{0};

char* err_buf() {
  static thread_local char buf[kErrLen] = {0};
  return buf;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached[dev] = n;
    else
      return 148;
  }
  return cached[dev];
}

namespace {
thread_local int (*t_presync_fn)(void*) = nullptr;
thread_local void* t_presync_arg = nullptr;
}  // namespace

int readback_wait(cudaStream_t stream, const std::function<int()>* first) {
  static thread_local cudaEvent_t ev[64] = {nullptr};
  int dev = 0;
  GHF_CUDA(cudaGetDevice(&dev));
  GHF_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  if (!ev[dev]) GHF_CUDA(cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming));
  GHF_CUDA(cudaEventRecord(ev[dev], stream));            // the read-back copy is the last thing enqueued so far
  if (first && *first)
    if (int rc = (*first)()) return rc;
  if (t_presync_fn) {
    int (*fn)(void*) = t_presync_fn;
    void* arg = t_presync_arg;
    t_presync_fn = nullptr;                              // one shot: consumed by the first entry point that waits
    t_presync_arg = nullptr;
    const int rc = fn(arg);
    if (rc != 0) return fail("the pre-sync hook failed (%d)", rc);
  }
  GHF_CUDA(cudaEventSynchronize(ev[dev]));
  return 0;
}

}  // namespace ghf

using namespace ghf;

extern "C" int ghf_set_presync_hook(int (*fn)(void*), void* arg) {
  t_presync_fn = fn;
  t_presync_arg = arg;
  return 0;
}

extern "C" int ghf_abi_version(void) { return GHF_ABI_VERSION; }
extern "C" const char* ghf_last_error(void) { return err_buf(); }
extern "C" int64_t ghf_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

extern "C" int ghf_copy_async(void* d_dst, const void* d_src, int64_t bytes, void* stream_) {
  GHF_REQUIRE(bytes >= 0 && (bytes == 0 || (d_dst != nullptr && d_src != nullptr)), "ghf_copy_async: bad arguments");
  if (bytes == 0) return 0;
  GHF_CUDA(cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream_));
  return 0;
}

extern "C" int ghf_device_ok(void) {
  static bool checked[64] = {false};   // per device, once: the calls below are not allowed during stream capture
  int dev = 0, major = 0;
  GHF_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && checked[dev]) return 0;
  GHF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  GHF_REQUIRE(major == 10, "ghf_b200 needs a compute-capability 10.x device (sm_100a); current device is %d.x",
              major);
  // keep stream-ordered scratch cached in the pool between calls
  cudaMemPool_t pool;
  GHF_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  uint64_t keep = ~0ull;
  GHF_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  if (dev >= 0 && dev < 64) checked[dev] = true;
  return 0;
}

namespace {

// flat parameter list: 5 global tensors, then per layer
//   3 MLPs (W_msg, W_self, bias) x (depth+1) x {weight, bias}, 3 log_scales, LayerNorm {weight, bias}
struct LayerParams {
  std::vector<const float*> w[3], b[3];
  const float* log_scale[3];
  const float *ln_w, *ln_b;
};

// Grow-only device arenas for the scratch of the whole-forward entry points (per device, two generations: sizes
// known at entry, and sizes known after dedup + graph build).  The stream-ordered pool is fine for the small
// scratch inside the stages, but re-allocating gigabytes from it on every call costs milliseconds of host time
// with the GPU idle.  One forward at a time per device (the entry points hold a lock).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, used = 0;
  cudaError_t reserve(size_t bytes) {
    used = 0;
    if (bytes <= cap) return cudaSuccess;
    if (base) {
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) return e;
      cudaFree(base);
      base = nullptr;
      cap = 0;
    }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&base), want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T>
  T* take(size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
  static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
};
Arena g_arena[64][3];
std::mutex g_forward_lock;

// side stream + events of the whole-forward entry points: the weight generators of all layers depend only on the
// text embeddings, so they run beside graph build and input projection instead of between the layers
constexpr int kMaxSideLayers = 16;
struct SideStream {
  cudaStream_t stream = nullptr;        // weight generators (+ operand-image packing) of every layer
  cudaStream_t proj_stream = nullptr;   // input projection: independent of dedup and graph build
  cudaEvent_t text_ready = nullptr, weights_ready[kMaxSideLayers] = {}, entry = nullptr, proj_done = nullptr,
              forward_done = nullptr;
  bool forward_recorded = false;
  cudaError_t init() {
    if (stream) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&proj_stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&entry, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&proj_done, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&forward_done, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&text_ready, cudaEventDisableTiming)) != cudaSuccess) return e;
    for (auto& ev : weights_ready)
      if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    return cudaSuccess;
  }
};
SideStream g_side[64];

}  // namespace

// GHF_FWD_TRACE=1: CUDA events on the main stream between the stages of the whole-forward entry, printed per call
// (synchronises; diagnostics only).
struct StageTrace {
  bool on = getenv("GHF_FWD_TRACE") != nullptr;
  cudaStream_t stream = nullptr;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  void mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    marks.push_back({name, e});
  }
  ~StageTrace() {
    if (!on || marks.empty()) return;
    cudaEventSynchronize(marks.back().second);
    fprintf(stderr, "forward stages (ms):");
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      fprintf(stderr, " %s %.3f |", marks[i].first, ms);
    }
    float total = 0;
    cudaEventElapsedTime(&total, marks.front().second, marks.back().second);
    fprintf(stderr, " total %.3f\n", total);
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
};

// Host-buffer entry: the LAST layer runs in a few pieces of super-blocks and the finished rows of a piece start their
// way to the host (copy stream) while the next piece is computed - the device-to-host copy of the result (1.28 GB at
// c3, ~23 ms over PCIe) begins ~2 ms earlier than after a whole-layer launch.
struct OutputPipe {
  float* h_out;
  cudaStream_t copy_stream;
  cudaEvent_t* piece_done;   // kMaxOutPieces events
};
constexpr int kMaxOutPieces = 4;

// The whole forward on DEVICE buffers (HG:236-298); d_out receives the final embeddings [num_nodes, d].
static int forward_device_impl(const ghf_model_desc* desc, const float* const* d_params, int64_t n_params,
                               const float* d_x, int64_t num_nodes, const int64_t* d_ei, int64_t E,
                               const uint8_t* d_utf8, const int64_t* d_offs, float* d_out, cudaEvent_t x_ready,
                               cudaStream_t stream, const OutputPipe* pipe = nullptr) {
  GHF_REQUIRE(desc && d_params, "ghf_hypergnn_forward: NULL model");
  const int T = desc->text_dim, F = desc->node_feat_dim, d = desc->hidden_dim, L = desc->num_layers;
  const int C = desc->char_emb_dim, H = desc->gen_hidden, depth = desc->gen_depth;
  GHF_REQUIRE(L >= 1, "num_layers must be at least 1");
  GHF_REQUIRE(T > 0 && F > 0 && d > 0 && C > 0 && depth >= 0 && (H > 0 || depth == 0), "bad model dimensions");
  const int64_t per_layer = 3 * 2 * (depth + 1) + 3 + 2;
  GHF_REQUIRE(n_params == 5 + L * per_layer, "expected %lld parameter tensors, got %lld",
              (long long)(5 + L * per_layer), (long long)n_params);
  if (int rc = ghf_device_ok()) return rc;

  const float *emb = d_params[0], *Wp = d_params[1], *bp = d_params[2], *Win = d_params[3], *bin = d_params[4];
  std::vector<LayerParams> layers(L);
  {
    int64_t p = 5;
    for (int l = 0; l < L; ++l) {
      for (int m = 0; m < 3; ++m)
        for (int i = 0; i <= depth; ++i) {
          layers[l].w[m].push_back(d_params[p++]);
          layers[l].b[m].push_back(d_params[p++]);
        }
      for (int m = 0; m < 3; ++m) layers[l].log_scale[m] = d_params[p++];
      layers[l].ln_w = d_params[p++];
      layers[l].ln_b = d_params[p++];
    }
  }

  const bool want_f16 = desc->precision == GHF_PREC_F16 && (d == 128 || d == 64);   // shadows chained layer to layer
  std::lock_guard<std::mutex> lock(g_forward_lock);
  int dev = 0;
  GHF_CUDA(cudaGetDevice(&dev));
  GHF_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  Arena& A = g_arena[dev][0];
  Arena& B = g_arena[dev][1];
  const size_t hbytes = (size_t)num_nodes * d * 4;
  GHF_CUDA(A.reserve(Arena::padded(16) + Arena::padded(E * sizeof(int32_t)) + Arena::padded(E * sizeof(int64_t)) +
                     2 * Arena::padded(hbytes) + (want_f16 ? 2 * Arena::padded(hbytes / 2) : 0) + 4096));
  float* scales = A.take<float>(4);                     // scale words of the two fp16 shadows (float[2] each)
  int32_t* rel = A.take<int32_t>(E > 0 ? E : 1);
  int64_t* first = A.take<int64_t>(E > 0 ? E : 1);
  float* h0 = A.take<float>((size_t)num_nodes * d);
  float* h1 = A.take<float>((size_t)num_nodes * d);
  void* h16_0 = want_f16 ? A.take<uint16_t>((size_t)num_nodes * d) : nullptr;   // fp16 shadow buffers (ping-pong)
  void* h16_1 = want_f16 ? A.take<uint16_t>((size_t)num_nodes * d) : nullptr;

  // The arenas are shared by consecutive forwards on this device: whatever stream the previous one ran on, this one
  // starts after it (an event, not a host wait).
  SideStream& side = g_side[dev];
  const bool use_side = L <= kMaxSideLayers && side.init() == cudaSuccess;
  if (use_side && side.forward_recorded) GHF_CUDA(cudaStreamWaitEvent(stream, side.forward_done, 0));
  // Side work must be joined before an error return lets the caller free or reuse buffers.
  struct JoinOnError {
    SideStream* s;
    bool armed;
    ~JoinOnError() {
      if (armed && s->stream) {
        cudaStreamSynchronize(s->stream);
        cudaStreamSynchronize(s->proj_stream);
      }
    }
  } join{&side, use_side};

  // HG:261  h = relu(input_proj(x))  (+ the fp16 shadow of h on the f16 path): depends on nothing but x, so it runs on
  // its own stream beside dedup, text encoder and graph build (HBM-bound vs. sort passes: they overlap well)
  cudaStream_t proj_stream = use_side ? side.proj_stream : stream;
  auto project = [&]() -> int {
    if (x_ready) GHF_CUDA(cudaStreamWaitEvent(proj_stream, x_ready, 0));
    return ghf_linear_f16out(d_x, num_nodes, F, Win, bin, d, 1, nullptr, h0, h16_0, want_f16 ? scales : nullptr,
                             proj_stream);
  };
  // It is enqueued from inside dedup, right before the host waits for the number of distinct strings: the projection
  // (HBM-bound, persistent CTAs) then runs through that round trip and beside the small sort / scan kernels that
  // follow, instead of in front of the dedup kernels.
  const bool entry_early = getenv("GHF_PROJ_ENTRY_LATE") == nullptr;   // experiment knob
  if (use_side && entry_early) GHF_CUDA(cudaEventRecord(side.entry, stream));
  const std::function<int()> enqueue_projection = [&]() -> int {
    if (!use_side) return 0;
    if (!entry_early) GHF_CUDA(cudaEventRecord(side.entry, stream));
    GHF_CUDA(cudaStreamWaitEvent(proj_stream, side.entry, 0));
    if (int rc = project()) return rc;
    GHF_CUDA(cudaEventRecord(side.proj_done, proj_stream));
    return 0;
  };

  StageTrace trace;
  trace.stream = stream;
  trace.mark("start");
  // HG:264-268  dedup (first-occurrence order), HG:270 text encoder on the distinct strings
  int64_t U = 0;
  bool projected = false;
  const std::function<int()> hook = [&]() -> int {
    projected = true;
    return enqueue_projection();
  };
  if (int rc = dedup_texts_hooked(d_utf8, d_offs, E, nullptr, 0, rel, first, &U, stream, &hook)) return rc;
  if (!projected)                                        // no strings at all: dedup returned before its hook
    if (int rc = enqueue_projection()) return rc;
  trace.mark("dedup");
  // the same rule as the staged entries (ghf_mp_layer_f16): an engine that does not cover this hidden size is an
  // error, not a silent change of arithmetic
  GHF_REQUIRE(desc->precision != GHF_PREC_F16 || d == 64 || d == 128 || d == 256,
              "ghf_hypergnn_forward: the f16 engine covers hidden_dim 64, 128 and 256, got %d", d);
  GHF_REQUIRE(desc->precision != GHF_PREC_TF32 || d == 32 || d == 64 || d == 128,
              "ghf_hypergnn_forward: the tf32 engine covers hidden_dim 32, 64 and 128, got %d", d);
  GHF_REQUIRE(desc->precision == GHF_PREC_FP32 || desc->precision == GHF_PREC_TF32 || desc->precision == GHF_PREC_F16,
              "ghf_hypergnn_forward: precision=%d", desc->precision);
  const int prec = desc->precision;
  // second arena: everything whose size depends on the number of distinct relations
  const size_t Un = (size_t)(U > 0 ? U : 1), Hn = (size_t)(H > 0 ? H : 1);
  // Hidden 64 / 256 on the f16 engine: the generator's last Linear writes the fp16 operand images of the contraction
  // itself (linear_umma_to_images) - no fp32 W_msg / W_self, no packing pass.  Needs the tcgen05
  // Linear (its input width is 128) and enough relations to fill it.
  const bool fuse = prec == GHF_PREC_F16 && mp_f16ss_supported(d) && H == 128 && depth >= 1 && U >= 64 &&
                    (int64_t)U * d * d >= (1 << 21) && !getenv("GHF_NO_FUSED_GENERATOR");
  // Hidden 128 on the f16 engine: fp32 weights are generated (70 MB at c3), and their fp16 operand images are packed
  // on the generator's stream too, so the layers on the main stream start with their operands ready.
  const char* det_env = getenv("GHF_DETERMINISTIC");   // the deterministic layer wants the fp32 matrices (a bound)
  const bool prepack = !fuse && prec == GHF_PREC_F16 && mp_f16_supported(d) && U > 0 && !(det_env && det_env[0] == '1');
  const size_t img_bytes = fuse ? (size_t)mp_f16ss_pack_bytes((int)Un, d) : prepack ? (size_t)mp_f16_pack_bytes((int)Un) : 0;
  const size_t w_layer = fuse ? Arena::padded(img_bytes) + Arena::padded(Un * 4) + Arena::padded(Un * d * 4)
                              : 2 * Arena::padded(Un * d * d * 4) + Arena::padded(Un * d * 4);
  const size_t gen_scratch_bytes = (size_t)ghf_weight_generators_scratch_bytes((int64_t)Un, H, depth, L);
  GHF_CUDA(B.reserve(Arena::padded(Un * T * 4) + (size_t)L * (w_layer + (prepack ? Arena::padded(img_bytes) : 0)) +
                     Arena::padded(gen_scratch_bytes) + 8192));
  float* temb = B.take<float>(Un * T);
  char* gen_scratch = B.take<char>(gen_scratch_bytes);
  float* words = B.take<float>(16);
  std::vector<std::array<float*, 3>> outs(L);
  std::vector<void*> images(L, nullptr);
  std::vector<float*> img_scale(L, nullptr);
  for (int l = 0; l < L; ++l) {
    if (fuse) {
      images[l] = B.take<char>(img_bytes);
      img_scale[l] = B.take<float>(Un);
      outs[l] = {nullptr, nullptr, B.take<float>(Un * d)};
    } else {
      outs[l] = {B.take<float>(Un * d * d), B.take<float>(Un * d * d), B.take<float>(Un * d)};
      if (prepack) images[l] = B.take<char>(img_bytes);
    }
  }
  if (int rc = ghf_text_encode(d_utf8, d_offs, first, U, emb, C, Wp, bp, T, temb, stream)) return rc;

  // WG:137-141 for the U distinct relations, EVERY layer at once (ghf_weight_generators: the hidden Linears of all
  // 3 L MLPs are one grouped launch per depth level), on the side stream, beside graph build and projection
  cudaStream_t gen_stream = use_side ? side.stream : stream;
  if (use_side) {
    GHF_CUDA(cudaEventRecord(side.text_ready, stream));
    GHF_CUDA(cudaStreamWaitEvent(side.stream, side.text_ready, 0));
  }
  auto generate_all = [&]() -> int {
    if (U == 0) return 0;
    std::vector<const float*> gp, gls;
    std::vector<float*> gout;
    for (int l = 0; l < L; ++l)
      for (int m = 0; m < 3; ++m) {
        for (int i = 0; i <= depth; ++i) {
          gp.push_back(layers[l].w[m][i]);
          gp.push_back(layers[l].b[m][i]);
        }
        gls.push_back(layers[l].log_scale[m]);
        gout.push_back(outs[l][m]);
      }
    if (int rc = ghf_weight_generators(temb, U, T, H, depth, L, d, d, gp.data(), gls.data(), gout.data(), gen_scratch,
                                       fuse ? 1 : 0, gen_stream))
      return rc;
    for (int l = 0; l < L; ++l) {
      if (prepack)
        if (int rc = mp_f16_pack_rel((int)U, outs[l][0], outs[l][1], images[l], gen_stream, false)) return rc;
      if (fuse) {
        // inputs of the two big heads of generator l (layout documented in ghf_b200.h)
        const float* base = reinterpret_cast<const float*>(gen_scratch) + (size_t)((depth - 1) & 1) * 3 * L * U * H;
        const float* zm = base + (size_t)(3 * l + 0) * U * H;
        const float* zs = base + (size_t)(3 * l + 1) * U * H;
        if (int rc = mp_f16ss_image_scales(zm, zs, H, U, layers[l].w[0][depth], layers[l].b[0][depth],
                                           layers[l].w[1][depth], layers[l].b[1][depth], d, layers[l].log_scale[0],
                                           layers[l].log_scale[1], words, img_scale[l], images[l], gen_stream))
          return rc;
        const float* z[2] = {zm, zs};
        for (int m = 0; m < 2; ++m)
          if (int rc = linear_umma_to_images(z[m], U, layers[l].w[m][depth], layers[l].b[m][depth], d, m,
                                             layers[l].log_scale[m], img_scale[l], images[l],
                                             mp_f16ss_image_bytes(d), gen_stream))
            return rc;
      }
      if (use_side) GHF_CUDA(cudaEventRecord(side.weights_ready[l], side.stream));
    }
    return 0;
  };
  if (use_side)
    if (int rc = generate_all()) return rc;

  ghf_graph* g = nullptr;
  if (int rc = ghf_graph_build(d_ei, E, nullptr, 0, rel, num_nodes, (int32_t)(U > 0 ? U : 1), d, 0, num_nodes, 0, 0,
                               &g, stream))
    return rc;
  struct Guard {
    ghf_graph* g;
    ~Guard() { ghf_graph_free(g); }
  } guard{g};
  trace.mark("text+graph");
  Arena& Cws = g_arena[dev][2];                          // third arena: the layer workspace (sized by the graph tables)
  const size_t ws_bytes = (size_t)ghf_mp_workspace_bytes(g, d, prec);
  GHF_CUDA(Cws.reserve(ws_bytes + 4096));
  void* ws = Cws.take<char>(ws_bytes);

  if (use_side) {
    GHF_CUDA(cudaStreamWaitEvent(stream, side.proj_done, 0));
  } else if (int rc = project()) {
    return rc;
  }
  trace.mark("proj-join");
  if (!use_side)
    if (int rc = generate_all()) return rc;

  float* cur = h0;
  float* nxt = h1;
  void* cur16 = want_f16 ? h16_0 : nullptr;   // fp16 shadow of `cur` and its scale words (chained at hidden 128)
  void* nxt16 = h16_1;
  float* cur_sc = scales;
  float* nxt_sc = scales + 2;
  for (int l = 0; l < L; ++l) {
    if (use_side && U > 0) GHF_CUDA(cudaStreamWaitEvent(stream, side.weights_ready[l], 0));
    // HG:286-296
    void* out16 = (want_f16 && l + 1 < L) ? nxt16 : nullptr;
    float* dst = l + 1 < L ? nxt : d_out;                // the last layer writes the caller's buffer
    // pieces of the last layer when its rows go to the host (at least 4 super-blocks per piece: a piece boundary
    // drains the persistent kernel once)
    int pieces = 1;
    if (pipe && l + 1 == L && !getenv("GHF_NO_OUTPUT_PIPE")) {
      pieces = (int)(g->num_phases / 4 < kMaxOutPieces ? g->num_phases / 4 : kMaxOutPieces);
      if (pieces < 1) pieces = 1;
    }
    for (int pc = 0; pc < pieces; ++pc) {
      const int p_lo = (int)(g->num_phases * pc / pieces), p_hi = (int)(g->num_phases * (pc + 1) / pieces);
      if ((fuse || prepack) && U > 0) {
        if (int rc = mp_layer_prepacked(g, cur, cur16, cur16 ? cur_sc : nullptr, images[l], outs[l][2], layers[l].ln_w,
                                        layers[l].ln_b, desc->ln_eps, dst, out16, out16 ? nxt_sc : nullptr, ws, stream,
                                        p_lo, p_hi))
          return rc;
      } else if (int rc = ghf_mp_layer_f16_range(g, cur, cur16, cur16 ? cur_sc : nullptr, outs[l][0], outs[l][1],
                                                 outs[l][2], layers[l].ln_w, layers[l].ln_b, desc->ln_eps, prec, dst,
                                                 out16, out16 ? nxt_sc : nullptr, nullptr, ws, p_lo, p_hi, stream)) {
        return rc;
      }
      if (pieces > 1) {                                    // rows of this piece: on their way while the next one runs
        const int64_t r0 = (int64_t)p_lo * g->sb_nodes;
        const int64_t r1 = (int64_t)p_hi * g->sb_nodes < num_nodes ? (int64_t)p_hi * g->sb_nodes : num_nodes;
        GHF_CUDA(cudaEventRecord(pipe->piece_done[pc], stream));
        GHF_CUDA(cudaStreamWaitEvent(pipe->copy_stream, pipe->piece_done[pc], 0));
        if (r1 > r0)
          GHF_CUDA(cudaMemcpyAsync(pipe->h_out + r0 * d, d_out + r0 * d, (size_t)(r1 - r0) * d * 4,
                                   cudaMemcpyDeviceToHost, pipe->copy_stream));
      }
    }
    if (pipe && l + 1 == L && pieces == 1)                 // one piece (small graphs): the plain copy, in stream order
      GHF_CUDA(cudaMemcpyAsync(pipe->h_out, d_out, (size_t)num_nodes * d * 4, cudaMemcpyDeviceToHost, stream));
    trace.mark("layer");
    float* t = cur; cur = nxt; nxt = t;
    t = cur_sc; cur_sc = nxt_sc; nxt_sc = t;
    cur16 = out16;
    nxt16 = (nxt16 == h16_1) ? h16_0 : h16_1;
  }
  if (use_side) {
    GHF_CUDA(cudaEventRecord(side.forward_done, stream));
    side.forward_recorded = true;
  }
  join.armed = false;                                    // everything was joined into `stream` by events
  return 0;
}

extern "C" int ghf_hypergnn_forward_device(const ghf_model_desc* desc, const float* const* d_params,
                                           int64_t n_params, const float* d_node_features, int64_t num_nodes,
                                           const int64_t* d_edge_index, int64_t E, const uint8_t* d_utf8,
                                           const int64_t* d_offsets, float* d_out, void* stream_) {
  return forward_device_impl(desc, d_params, n_params, d_node_features, num_nodes, d_edge_index, E, d_utf8,
                             d_offsets, d_out, nullptr, (cudaStream_t)stream_);
}

extern "C" int ghf_hypergnn_forward_host(const ghf_model_desc* desc, const float* const* d_params,
                                         int64_t n_params, const float* h_node_features, int64_t num_nodes,
                                         const int64_t* h_edge_index, int64_t E, const uint8_t* h_utf8,
                                         const int64_t* h_offsets, float* h_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(desc && d_params, "ghf_hypergnn_forward_host: NULL model");
  GHF_REQUIRE(desc->node_feat_dim > 0 && desc->hidden_dim > 0, "bad model dimensions");
  const int F = desc->node_feat_dim, d = desc->hidden_dim;
  const int64_t text_bytes = E > 0 ? h_offsets[E] : 0;
  TempBuf x, ei, utf8, offs, out;
  GHF_CUDA(x.alloc(num_nodes * (size_t)F * 4, stream));
  GHF_CUDA(ei.alloc(2 * E * sizeof(int64_t), stream));
  GHF_CUDA(utf8.alloc(text_bytes, stream));
  GHF_CUDA(offs.alloc((E + 1) * sizeof(int64_t), stream));
  GHF_CUDA(out.alloc(num_nodes * (size_t)d * 4, stream));
  // edges and strings first (dedup and graph build need only them); the node features follow on a second stream
  // and are waited for right before the input projection
  static cudaStream_t copy_stream[64] = {nullptr};
  static cudaEvent_t x_ready[64] = {nullptr}, buffers_ready[64] = {nullptr}, piece_done[64][kMaxOutPieces] = {};
  int dev = 0;
  GHF_CUDA(cudaGetDevice(&dev));
  GHF_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
  if (!copy_stream[dev]) {
    GHF_CUDA(cudaStreamCreateWithFlags(&copy_stream[dev], cudaStreamNonBlocking));
    GHF_CUDA(cudaEventCreateWithFlags(&x_ready[dev], cudaEventDisableTiming));
    GHF_CUDA(cudaEventCreateWithFlags(&buffers_ready[dev], cudaEventDisableTiming));
    for (auto& ev : piece_done[dev]) GHF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  GHF_CUDA(cudaMemcpyAsync(ei.p, h_edge_index, 2 * E * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  GHF_CUDA(cudaMemcpyAsync(utf8.p, h_utf8, text_bytes, cudaMemcpyHostToDevice, stream));
  GHF_CUDA(cudaMemcpyAsync(offs.p, h_offsets, (E + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  // the feature copy starts when those three are through (the link is shared), in `stream` order after x.p exists
  GHF_CUDA(cudaEventRecord(buffers_ready[dev], stream));
  GHF_CUDA(cudaStreamWaitEvent(copy_stream[dev], buffers_ready[dev], 0));
  GHF_CUDA(cudaMemcpyAsync(x.p, h_node_features, num_nodes * (size_t)F * 4, cudaMemcpyHostToDevice, copy_stream[dev]));
  GHF_CUDA(cudaEventRecord(x_ready[dev], copy_stream[dev]));
  // the result leaves piece by piece from inside the last layer (OutputPipe); copy_stream is idle again by then
  // (the feature copy was waited for by the projection)
  const OutputPipe pipe{h_out, copy_stream[dev], piece_done[dev]};
  const int rc = forward_device_impl(desc, d_params, n_params, x.as<float>(), num_nodes, ei.as<int64_t>(), E,
                                     utf8.as<uint8_t>(), offs.as<int64_t>(), out.as<float>(), x_ready[dev], stream,
                                     &pipe);
  // both streams drain before the scratch is released and before the caller reads h_out, error or not
  const cudaError_t e1 = cudaStreamSynchronize(stream), e2 = cudaStreamSynchronize(copy_stream[dev]);
  if (rc) return rc;
  GHF_CUDA(e1);
  GHF_CUDA(e2);
  return 0;
}

// ---- generator -> operand-image fusion as C-ABI pieces (the whole-forward entry points use the same functions)
extern "C" int64_t ghf_weight_images_bytes(int64_t R, int32_t hidden_dim) {
  if (!mp_f16ss_supported(hidden_dim) || R < 0) return -1;
  return mp_f16ss_pack_bytes((int)(R > 0 ? R : 1), hidden_dim) + 256;
}

extern "C" int ghf_weight_images_f16(const float* d_Zm, const float* d_Zs, int64_t R, const float* d_W3m,
                                     const float* d_b3m, const float* d_lsm, const float* d_W3s, const float* d_b3s,
                                     const float* d_lss, int32_t hidden_dim, void* d_images, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(mp_f16ss_supported(hidden_dim), "ghf_weight_images_f16: hidden_dim 64 or 256, got %d", hidden_dim);
  GHF_REQUIRE(R >= 0 && d_Zm && d_Zs && d_W3m && d_b3m && d_lsm && d_W3s && d_b3s && d_lss && d_images,
              "ghf_weight_images_f16: NULL argument");
  GHF_REQUIRE(reinterpret_cast<uintptr_t>(d_images) % 1024 == 0, "ghf_weight_images_f16: d_images must be 1 KiB aligned");
  if (R == 0) return 0;
  TempBuf scratch;                                   // [8 floats of range words][R scales]
  GHF_CUDA(scratch.alloc(256 + (size_t)R * sizeof(float), stream));
  float* words = scratch.as<float>();
  float* scale = words + 64;
  if (int rc = mp_f16ss_image_scales(d_Zm, d_Zs, 128, R, d_W3m, d_b3m, d_W3s, d_b3s, hidden_dim, d_lsm, d_lss, words,
                                     scale, d_images, stream))
    return rc;
  if (int rc = linear_umma_to_images(d_Zm, R, d_W3m, d_b3m, hidden_dim, 0, d_lsm, scale, d_images,
                                     mp_f16ss_image_bytes(hidden_dim), stream))
    return rc;
  return linear_umma_to_images(d_Zs, R, d_W3s, d_b3s, hidden_dim, 1, d_lss, scale, d_images,
                               mp_f16ss_image_bytes(hidden_dim), stream);
}

extern "C" int ghf_mp_layer_images(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                                   const void* d_images, const float* d_bias, const float* d_ln_w, const float* d_ln_b,
                                   float eps, float* d_out, void* d_out16, float* d_out16_scale, void* d_workspace,
                                   void* stream_) {
  GHF_REQUIRE(d_out16 == nullptr || d_out16_scale != nullptr, "ghf_mp_layer_images: d_out16 needs d_out16_scale");
  return mp_layer_prepacked(g, d_h, d_h16, d_h16_scale, d_images, d_bias, d_ln_w, d_ln_b, eps, d_out, d_out16,
                            d_out16_scale, d_workspace, (cudaStream_t)stream_);
}
