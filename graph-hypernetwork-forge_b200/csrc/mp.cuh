// mp.cuh — interface between ghf_mp_layer (mp.cu) and the tcgen05 contraction engine (mp_umma.cu).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "graph.cuh"

namespace ghf {

// Multi-GPU: the fp16 rows the epilogue produces are also stored straight into the peers' copies of the shadow table
// (peer-mapped symmetric memory, NVLink stores) - but only into the copies of the ranks that READ the row: rank q
// gathers row v only if one of its edges has source v (`mask[q * mask_stride + local row]`, exchanged once per
// graph).  No staging copy, no collective kernel, ~half the bytes of an all-gather at in-degree 6.
struct PeerPush {
  const uint8_t* mask = nullptr;      // [world][mask_stride] bytes; nullptr: no push
  int64_t mask_stride = 0;
  __half* const* tables = nullptr;    // device array [world]: base of every rank's [N, D] fp16 table
  int world = 0, me = 0;
  int64_t table_row0 = 0;             // global row of local row 0
};

// Training-mode dropout between the ReLU and the LayerNorm of a layer (HG:293-294), applied inside the row epilogue on
// the SAME Philox stream torch's CUDA dropout would use for the [num_nodes, d] tensor (tools/dropout_stream_probe.py
// restates and checks the mapping): element i belongs to "thread" t = (i / 4) % threads at step s = (i / 4) / threads;
// it is kept iff uniform(Philox4x32-10(key = seed, counter = (offset / 4 + s, subsequence t))[i % 4]) < keep, and kept
// values are multiplied by `scale` = float(1 / keep).  keep == 0: no dropout.
struct DropoutArgs {
  float keep = 0.f, scale = 1.f;
  uint32_t key0 = 0, key1 = 0;   // the generator's seed
  uint64_t ctr0 = 0;             // the generator's offset / 4 at the time of the call
  uint32_t threads = 1;          // threads of the launch torch would make for this tensor
};
// fills `out` for dropout probability p on a tensor of `numel` elements (numel % 4 == 0) on the current device;
// *advance = how far the call moves the generator's offset
int make_dropout_args(float p, uint64_t seed, uint64_t offset, int64_t numel, DropoutArgs* out, int64_t* advance);

bool mp_umma_supported(int hidden_dim);
// bytes of scratch for the per-relation operand images [R][2d x d] (tf32, UMMA K-major SW128 layout)
int64_t mp_umma_pack_bytes(int num_rel, int hidden_dim);
// operand images: pack[r] = tf32([W_msg[r]; W_self[r]]) in the kernel's shared-memory layout
int mp_umma_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch,
                 cudaStream_t stream);
// acc[dst_local, :] += sum over edges of [h_src | h_dst] @ [W_msg; W_self][rel] + bias[rel]
// `unit_counter` is one zeroed int in device memory (the shared work counter of the persistent CTAs)
int mp_umma_launch(const ghf_graph* g, const float* h, const float* bias, float* acc, const void* pack_scratch,
                   int* unit_counter, cudaStream_t stream);

// hidden_dim 128 with the weights resident in tensor memory (mp_umma_ts.cu); same contracts as above
bool mp_ts_supported(int hidden_dim);
int mp_ts_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream);
int mp_ts_launch(const ghf_graph* g, const float* h, const float* bias, float* acc, const void* pack_scratch,
                 int* unit_counter, cudaStream_t stream);


// hidden_dim 128 with fp16 feature transport, kind::f16 MMA and double-buffered weights in TMEM (mp_f16.cu)
bool mp_f16_supported(int hidden_dim);
// scratch for the per-relation fp16 weight images [R][64 KiB] followed by their inverse power-of-two scales [R]
int64_t mp_f16_pack_bytes(int num_rel);
// W_msg or W_self may be NULL (zeros); `transposed`: the images are built from W[r]^T
int mp_f16_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream,
                bool transposed = false);
// the same without a graph (the whole-forward entry packs on the generator's stream, before the graph exists)
int mp_f16_pack_rel(int num_rel, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream,
                    bool transposed = false);
// fp16 shadow of h: (h16, scale[2]) with h = h16 * scale[0], scale[1] = max|h| (see common.cuh).
// absmax: scale[1] = max|x|.  convert: scale chosen from scale[1], scale[0] written, h16 = fp16(h * s); with
// `rescue` the shadow already holds fp16(h) and is rewritten only if the range demands a scale.
int mp_f16_absmax(const float* x, int64_t elems, float* scale, cudaStream_t stream);
int mp_f16_convert(const float* h, int64_t elems, void* h16, float* scale, bool rescue, cudaStream_t stream);
// acc[dst_local, :] = sum over edges of [h16_src | h16_dst] @ [W_msg; W_self][rel] + bias[rel]; h16 is [N, 128] fp16.
// The kernel clears acc itself (every local row, also those without in-edges) unless `keep_acc` (then it adds to
// what the rows hold).  sync_words: mp_f16_sync_bytes(g).
int64_t mp_f16_sync_bytes(const ghf_graph* g);
int mp_f16_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias, float* acc,
                  const void* pack_scratch, int* sync_words, cudaStream_t stream, bool keep_acc = false,
                  int skip_half = 0,    // skip_half: 1 = no source term (W_msg NULL), 2 = no destination term
                  int phase_lo = 0, int phase_hi = -1,    // super-blocks the given units lie in (clearing covers them)
                  const float* det_W_msg = nullptr, const float* det_W_self = nullptr);
// Deterministic mode (both det_W_* given: the fp32 relation matrices, for a bound on the messages): per-destination
// sums are accumulated as int32 fixed point - `acc` then holds integers, scaled per destination by 2^k_v with
// k_v = 30 - bits(indeg_v) - eB, eB = mp_f16_det_words(sync_words)[2] (device) - and the epilogue converts back.
const int32_t* mp_f16_det_words(const int* sync_words);

// The whole layer in one kernel (mp_f16_fused.cu, hidden_dim 128): the contraction reduces into a ring of L2-resident
// accumulator windows and the row epilogue (mean, residual, ReLU, LayerNorm, fp16 shadow) runs per super-block inside
// the same kernel.  `ring`: mp_f16_fused_ring_rows(g) x 128 floats; sync_words: mp_f16_sync_bytes(g); `upd`, `out16`
// optional.  GHF_MP_FUSED=0 falls back to mp_f16_launch + the separate epilogue kernel.
bool mp_f16_fused_enabled(const ghf_graph* g);
int64_t mp_f16_fused_ring_rows(const ghf_graph* g);
int mp_f16_fused_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias,
                        const void* pack_scratch, float* ring, int* sync_words, const float* h, const float* ln_w,
                        const float* ln_b, float eps, float* out, float* upd, void* out16, float* out16_scale,
                        cudaStream_t stream, const PeerPush* push = nullptr);

// hidden_dim 256 / 64 with streamed fp16 weights (mp_f16_ss.cu): operand images [R][256 KiB / 16 KiB] + inverse
// scales [R]; acc must be zero at entry, `unit_counter` one zeroed int
bool mp_f16ss_supported(int hidden_dim);
int64_t mp_f16ss_pack_bytes(int num_rel, int hidden_dim);
int mp_f16ss_pack(const ghf_graph* g, const float* W_msg, const float* W_self, void* pack_scratch, cudaStream_t stream);
int mp_f16ss_launch(const ghf_graph* g, const void* h16, const float* h16_scale, const float* bias, float* acc,
                    const void* pack_scratch, int* unit_counter, cudaStream_t stream);

// images written by the generator itself (linear_umma.cu: linear_umma_to_images): per-relation scales from an
// analytic bound, and the layer on pre-built images
int64_t mp_f16ss_image_bytes(int hidden_dim);
int mp_f16ss_image_scales(const float* Zm, const float* Zs, int H, int64_t R, const float* W3m, const float* b3m,
                          const float* W3s, const float* b3s, int d, const float* ls_m, const float* ls_s,
                          float* words, float* scale, void* images, cudaStream_t stream);
int linear_umma_to_images(const float* X, int64_t M, const float* W, const float* b, int d, int which,
                          const float* log_scale, const float* row_scale, void* images, int64_t image_bytes,
                          cudaStream_t stream);
// ghf_mp_layer_f16 for hidden 64 / 256 with the operand images already built (`images`: mp_f16ss_pack_bytes layout)
// phase_lo / phase_hi: only the super-blocks [phase_lo, phase_hi) (default: all), as ghf_mp_layer_f16_range
int mp_layer_prepacked(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                       const void* images, const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                       float* d_out, void* d_out16, float* d_out16_scale, void* d_workspace, cudaStream_t stream,
                       int phase_lo = 0, int phase_hi = -1);

// gradients of the generated relation tensors on tcgen05 (mp_wgrad_f16.cu, hidden_dim 128): g_W_msg[r] / g_W_self[r] /
// g_bias[r] += sums over the edges of r (buffers zero at entry); h16 / g16 are fp16 shadows with their scale words.
int mp_wgrad_f16_launch(const ghf_graph* g, const void* h16, const float* h_scale, const void* g16,
                        const float* g_scale, float* gW_msg, float* gW_self, float* gbias, int* unit_counter,
                        cudaStream_t stream);

}  // namespace ghf
