/* ghf_b200.h — C ABI of the B200-native HyperGNN forward path.
 *
 * The reference (danieleschmidt/Graph-Hypernetwork-Forge) has no FFI layer: its hot path is
 * Python calling stock ATen ops.  Each entry point below replaces one group of those call sites
 * ("HG" = graph_hypernetwork_forge/models/hypergnn.py, "WG" = .../models/weight_generator.py).
 * The Python drop-in (graph-hypernetwork-forge_b200/graph_hypernetwork_forge) binds them with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory on the current CUDA device, h_* is HOST memory;
 *   - tensors are dense row-major; float = IEEE fp32; ids are int32 unless stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - functions enqueue work on `stream` and return without synchronising unless stated;
 *   - return value 0 = success, non-zero = failure; ghf_last_error() describes the failure
 *     (thread-local).  There is no CPU fallback: without a usable sm_100 device every
 *     compute entry point fails.
 */
#ifndef GHF_B200_H
#define GHF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHF_ABI_VERSION 6

/* precision of the relation-typed contraction in ghf_mp_layer */
#define GHF_PREC_FP32 0 /* CUDA-core FFMA, fp32 end to end (rtol 1e-5 vs reference)          */
#define GHF_PREC_TF32 1 /* tcgen05 kind::tf32, fp32 accumulate in TMEM (tolerance: DESIGN.md) */
#define GHF_PREC_F16  2 /* fp16 feature/weight transport with exact power-of-two scaling, tcgen05 kind::f16,
                           fp32 accumulate in TMEM: the same 11-bit operand significand as TF32 at half the
                           bytes (hidden_dim 128; 64 and 256 with streamed weights; tolerance: DESIGN.md) */

int ghf_abi_version(void);
const char* ghf_last_error(void);
/* 0 when a compute-capability 10.x device is current and usable */
int ghf_device_ok(void);

/* ---- a4: relation dedup (HG:264-268) -------------------------------------------------------
 * Strings are packed as UTF-8 bytes + offsets[E+1] (string e = bytes [off[e], off[e+1])).
 * Work items: all E strings (d_subset NULL), or the strings d_subset[0..n_subset) (ascending ids; a rank's own
 * edges from ghf_select_edges).  d_rel_ids[j] = rank of work item j's string among the distinct strings of the
 * work items in FIRST-OCCURRENCE order; d_first_edge[u] = smallest string id with rank u (capacity: number of
 * work items).  The number of distinct strings is written to *h_num_unique after an internal stream synchronise. */
int ghf_dedup_texts(const uint8_t* d_utf8, const int64_t* d_offsets, int64_t E, const uint32_t* d_subset,
                    int64_t n_subset, int32_t* d_rel_ids, int64_t* d_first_edge, int64_t* h_num_unique,
                    void* stream);

/* Multi-GPU helper: the edges whose destination lies in [dst_lo, dst_hi), as ascending edge ids in d_edge_ids
 * (capacity E); their number in *h_count (synchronises).  Dedup and graph build then touch only those edges. */
int ghf_select_edges(const int64_t* d_edge_index, int64_t E, int64_t dst_lo, int64_t dst_hi,
                     uint32_t* d_edge_ids, int64_t* h_count, void* stream);

/* ---- a5+a6: TextEncoder (HG:66-81) ----------------------------------------------------------
 * out[u,:] = tanh( mean_i Emb[tok_i] @ Wp^T + bp ), tok = code points clamped to 127 recovered
 * from UTF-8 (lead byte >= 0x80 -> 127, continuation bytes skipped), "" -> [0].
 * Emb [128,C], Wp [T,C], bp [T], out [U,T].  Row u encodes string d_string_index[u] of the
 * packed set (pass the d_first_edge of ghf_dedup_texts to encode the distinct strings in
 * first-occurrence order), or string u itself when d_string_index is NULL. */
int ghf_text_encode(const uint8_t* d_utf8, const int64_t* d_offsets, const int64_t* d_string_index,
                    int64_t U, const float* d_emb, int C, const float* d_Wp, const float* d_bp, int T,
                    float* d_out, void* stream);

/* ---- a3 / a8: Linear (+ReLU) (+exp(log_scale)) (HG:261, WG:97-107, WG:138-140) -------------
 * Y[M,N] = alpha * act( X[M,K] @ W[N,K]^T + b[N] ), act = ReLU when relu != 0,
 * alpha = exp(*d_log_scale) when d_log_scale != NULL else 1.  fp32-grade: tcgen05 with a 3xTF32 split when K = 128,
 * N % 128 == 0 and M*N >= 2^21 (node projection, the generators' last Linear), fp32 FFMA tiles otherwise. */
int ghf_linear(const float* d_X, int64_t M, int K, const float* d_W, const float* d_b, int N,
               int relu, const float* d_log_scale, float* d_Y, void* stream);

/* fp16 SHADOW of a feature matrix (GHF_PREC_F16): a pair (d_y16, d_scale) with d_scale = float[2] in device
 * memory: y = y16 * d_scale[0] (an exact power of two, chosen on the device so that the largest magnitude lands in
 * [2^13, 2^14) - no host round trip, no dependence on the data range), d_scale[1] = max |y|.
 *
 * ghf_linear plus an fp16 shadow of Y in (d_Y16 [M,N], d_Y16_scale), both NULL for none: the node-feature
 * projection (HG:261) emits the shadow of h that GHF_PREC_F16 gathers from in the same pass (fused on the tcgen05
 * path: written unscaled with its range, rewritten by a rescue pass only if the range demands a scale; otherwise
 * a range pass and a conversion pass follow, which need M*N % 8 == 0 and 16-byte aligned pointers). */
int ghf_linear_f16out(const float* d_X, int64_t M, int K, const float* d_W, const float* d_b, int N,
                      int relu, const float* d_log_scale, float* d_Y, void* d_Y16, float* d_Y16_scale,
                      void* stream);

/* ---- f3: gradients of ghf_linear (what loss.backward() needs of HG:261, WG:97-107, WG:138-140) ------------------
 * For Y = alpha * act(X W^T + b) and gY = dL/dY, with g_pre = gY * [Y > 0] under ReLU (g_pre = gY otherwise):
 *   d_gX [M,K] = alpha * g_pre W          d_gW [N,K] = alpha * g_pre^T X
 *   d_gb [N]   = alpha * sum_m g_pre      d_gls [1]  = sum gY * Y   (the gradient of log_scale)
 * Any output may be NULL (not computed).  d_Y is needed under ReLU and for d_gls, d_X for d_gW, d_W for d_gX.
 * fp32 FFMA tiles; the ReLU mask is applied while tiles are loaded (no g_pre tensor), sums over rows and split
 * contractions are reduced with fp32 atomics (outputs are zeroed here).  Replaces the library GEMMs of round 1. */
int ghf_linear_backward(const float* d_X, int64_t M, int K, const float* d_W, int N, int relu,
                        const float* d_log_scale, const float* d_Y, const float* d_gY, float* d_gX, float* d_gW,
                        float* d_gb, float* d_gls, void* stream);

/* ---- overlap of host round trips ---------------------------------------------------------------------------------
 * ghf_select_edges, ghf_dedup_texts and ghf_graph_build each wait once for a size to come back from the device.
 * A hook set here is called ONCE by the next of those calls on this thread, after its kernels are enqueued and right
 * before it waits: whatever the hook enqueues (the input projection, the weight generators - on any stream) runs
 * through the round trip instead of after it.  The hook returns 0; anything else fails the entry point.  A hook
 * that is never consumed (an early error, an empty input) stays set: clear it with (NULL, NULL). */
int ghf_set_presync_hook(int (*fn)(void*), void* arg);

/* ---- graph preprocessing (replaces the per-call gathers HG:281-283, scatter index HG:207-219)
 * Builds, for destinations dst in [dst_lo, dst_hi):
 *   in-degree, dst-CSR rowptr, and the edge order (super-block of dst, relation, dst), stable,
 *   cut into work units of <= unit_edges edges that share one relation.
 * d_edge_index is the reference's [2,E] int64 tensor.  Edges with dst outside the range are
 * dropped (1-D destination partition for multi-GPU).  Work items: all E edges with d_rel_ids[E] (d_edge_ids
 * NULL), or the pre-selected edges d_edge_ids[0..n_subset) with d_rel_ids[n_subset] indexed like d_edge_ids.
 * sb_nodes <= 0 / unit_edges <= 0 pick defaults.  Synchronises `stream` once (to size the tables); fails when a
 * node id lies outside [0, num_nodes) or a relation id outside [0, num_rel). */
typedef struct ghf_graph ghf_graph;
int ghf_graph_build(const int64_t* d_edge_index, int64_t E, const uint32_t* d_edge_ids, int64_t n_subset,
                    const int32_t* d_rel_ids, int64_t num_nodes, int32_t num_rel, int32_t hidden_dim,
                    int64_t dst_lo, int64_t dst_hi, int32_t sb_nodes, int32_t unit_edges,
                    ghf_graph** out, void* stream);
void ghf_graph_free(ghf_graph* g);
/* sizes: [0]=kept edges, [1]=units, [2]=sb_nodes, [3]=unit_edges, [4]=local nodes, [5]=bytes held */
int ghf_graph_info(const ghf_graph* g, int64_t info[6]);
/* device views for parity checks: any output pointer may be NULL.  perm[kept] int64 (original edge id at each
 * sorted position; no kernel needs it, so it is recomputed here by re-sorting - pass the SAME d_edge_index,
 * d_edge_ids / n_subset and d_rel_ids the graph was built from; they may be NULL when d_perm is NULL),
 * indeg[local nodes] int32, rowptr[local nodes+1] int64, unit_start/unit_count/unit_rel [units] int32. */
int ghf_graph_export(const ghf_graph* g, const int64_t* d_edge_index, const uint32_t* d_edge_ids, int64_t n_subset,
                     const int32_t* d_rel_ids, int64_t* d_perm, int32_t* d_indeg, int64_t* d_rowptr,
                     int32_t* d_unit_start, int32_t* d_unit_count, int32_t* d_unit_rel, void* stream);

/* ---- a10-a13: one message-passing layer (HG:160-230, HG:289-296) ---------------------------
 *   upd_v = (1/c_v) sum_{e:(u->v)} ( h_u W_msg[r_e] + bias[r_e] + h_v W_self[r_e] ),  c_v = max(indeg,1)
 *   out_v = LayerNorm( relu(upd_v + h_v) ) * ln_w + ln_b            (eps, biased variance)
 * d_h [num_nodes,d] (all nodes: sources are arbitrary); W_msg/W_self [R,d,d] ([r][in][out]),
 * bias [R,d]; d_out / d_upd hold rows [dst_lo,dst_hi) only ([local nodes, d]); d_upd may be NULL.
 * d_workspace: ghf_mp_workspace_bytes(g, d) bytes of scratch (no alignment beyond 256 B). */
int64_t ghf_mp_workspace_bytes(const ghf_graph* g, int32_t hidden_dim, int precision);
int ghf_mp_layer(const ghf_graph* g, const float* d_h, const float* d_W_msg, const float* d_W_self,
                 const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                 int precision, float* d_out, float* d_upd, void* d_workspace, void* stream);

/* The same layer with the fp16 shadow of the node features made explicit (GHF_PREC_F16 chains it from layer to
 * layer instead of re-converting):
 *   (d_h16, d_h16_scale)       shadow of d_h ([num_nodes, d] fp16 + float[2]), or NULL/NULL (then it is made inside,
 *                              in the workspace).  When given, d_h is only read at rows [dst_lo, dst_hi)
 *                              (residual), so a multi-GPU caller exchanges only the fp16 rows between layers and may
 *                              keep just its own fp32 rows: d_h = (pointer to row dst_lo) - dst_lo * d floats.
 *   (d_out16, d_out16_scale)   shadow of d_out ([local nodes, d] fp16 + float[2]) for the next layer, or NULL/NULL
 *                              (hidden_dim 32/64/128).  Its scale follows from ln_w / ln_b alone
 *                              (|LayerNorm(x)_c| <= sqrt(d-1)|w_c| + |b_c|), so every rank picks the same one.
 * With precision FP32 / TF32 this is ghf_mp_layer plus the optional output shadow. */
int ghf_mp_layer_f16(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                     const float* d_W_msg, const float* d_W_self, const float* d_bias, const float* d_ln_w,
                     const float* d_ln_b, float eps, int precision, float* d_out, void* d_out16,
                     float* d_out16_scale, float* d_upd, void* d_workspace, void* stream);

/* The same layer restricted to the super-blocks [phase_lo, phase_hi) of the graph (ghf_graph_num_phases of them, each
 * ghf_graph_info()[2] = sb_nodes destination rows): only rows [phase_lo * sb_nodes, min(phase_hi * sb_nodes, local
 * nodes)) of d_out / d_out16 / d_upd are written (the pointers still address row 0 of the local range).  A multi-GPU
 * caller runs a rank's rows in a few such pieces and sends the finished rows of one piece to the peers while the next
 * one is computed ("chunk the shard and overlap gather of finished dst tiles with remaining compute"). */
int64_t ghf_graph_num_phases(const ghf_graph* g);
int ghf_mp_layer_f16_range(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                           const float* d_W_msg, const float* d_W_self, const float* d_bias, const float* d_ln_w,
                           const float* d_ln_b, float eps, int precision, float* d_out, void* d_out16,
                           float* d_out16_scale, float* d_upd, void* d_workspace, int32_t phase_lo, int32_t phase_hi,
                           void* stream);

/* Multi-GPU, fused with the exchange: the same ranged layer whose epilogue kernel ALSO stores each fp16 result row
 * straight into the peers' copies of the shadow table (peer-mapped symmetric memory, NVLink stores from the kernel),
 * but only for the peers that read the row.  d_peer_mask[q * mask_stride + r] != 0 <=> rank q has an edge whose
 * source is this rank's local row r (ghf_mark_rows on rank q + one all-to-all of the byte masks per graph);
 * d_peer_tables: device array of `world` pointers, entry q = base of rank q's [num_nodes, d] fp16 table; d_out16 is
 * this rank's own rows of its own table (as in ghf_mp_layer_f16).  hidden_dim 64 / 128. */
int ghf_mp_layer_f16_push(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                          const float* d_W_msg, const float* d_W_self, const float* d_bias, const float* d_ln_w,
                          const float* d_ln_b, float eps, int precision, float* d_out, void* d_out16,
                          float* d_out16_scale, float* d_upd, void* d_workspace, int32_t phase_lo, int32_t phase_hi,
                          const uint8_t* d_peer_mask, int64_t mask_stride, void* const* d_peer_tables, int32_t world,
                          int32_t rank, void* stream);
/* d_mask[v] = 1 for every v = d_ids[j] (j over all n entries, or over d_subset[0..n) when given), 0 elsewhere:
 * the rows of the node table a rank gathers from (sources of its edges). */
int ghf_mark_rows(const int64_t* d_ids, const uint32_t* d_subset, int64_t n, int64_t num_rows, uint8_t* d_mask,
                  void* stream);

/* Building a shadow by hand (layer 0 of a multi-GPU run, where max|h| must be agreed between ranks first):
 * ghf_absmax writes d_scale[1] = max |x|; ghf_convert_f16 picks the scale from d_scale[1] (computing it first when
 * have_amax == 0), writes d_scale[0] and d_y16 = fp16(x * 2^k).  elems % 8 == 0, 16-byte aligned pointers. */
int ghf_absmax(const float* d_x, int64_t elems, float* d_scale, void* stream);
int ghf_convert_f16(const float* d_x, int64_t elems, void* d_y16, float* d_scale, int have_amax, void* stream);

/* ---- gradients of the path (the reference trains through HyperGNN.forward:
 * tests/test_hypergnn.py:183-226, demo.py:79-101) ------------------------------------------------------------
 * ghf_mp_contract: only the contraction of a layer - d_acc[v] = sum over in-edges (u->v, r) of
 *   x_u W_msg[r] + x_v W_self[r] + bias[r]   (no mean, no epilogue; [local nodes, d], 256-byte aligned).
 * It is the forward kernel, and it is also the gradient w.r.t. h: with x = g_acc and the relation matrices
 * transposed, on the graph built from the REVERSED edge list (W_self = 0) it yields the message term, on the
 * graph itself (W_msg = 0) the self-loop term.  (d_x16, d_x16_scale) as in ghf_mp_layer_f16.  accumulate != 0: the
 * sums are ADDED to what d_acc holds (the three shares of dL/dh land in one buffer without extra passes).
 * With precision GHF_PREC_F16 (hidden 128): transposed != 0 reads the relation matrices as W[r]^T while packing (no
 * transposed copy), and d_W_msg, d_W_self or d_bias may be NULL (zeros) - the rows of an absent half are neither
 * gathered nor multiplied. */
int ghf_mp_contract(const ghf_graph* g, const float* d_x, const void* d_x16, const float* d_x16_scale,
                    const float* d_W_msg, const float* d_W_self, const float* d_bias, int precision,
                    float* d_acc, int accumulate, int transposed, void* d_workspace, void* stream);
/* Undo LayerNorm, ReLU and the mean (HG:212-213, 289-296) for rows [dst_lo, dst_hi): from d_g_out = dL/d out,
 * the saved pre-residual update d_upd (ghf_mp_layer's tap) and d_h:
 *   d_g_pre = dL/d(upd + h)  (also the residual's share of dL/dh),  d_g_acc = d_g_pre / max(indeg, 1),
 *   d_g_ln_w[d], d_g_ln_b[d] = LayerNorm parameter gradients (overwritten).  Any hidden_dim (rows beyond 256
 *   columns are staged in shared memory).
 *   d_g_acc_scale (float[2], optional, hidden_dim 32/64/128): [1] = max |g_acc|, ready for ghf_convert_f16 with
 *   have_amax = 1 - the fp16 shadow of g_acc the gradient contractions gather from. */
int ghf_mp_epilogue_backward(const ghf_graph* g, const float* d_g_out, const float* d_upd, const float* d_h,
                             const float* d_ln_w, float eps, float* d_g_pre, float* d_g_acc, float* d_g_ln_w,
                             float* d_g_ln_b, float* d_g_acc_scale, void* stream);
/* ---- f4: training-mode dropout inside the row epilogue (HG:293-294: F.dropout between the ReLU and the LayerNorm) --
 * ghf_mp_layer_f16 with x = dropout(relu(upd + h), p_drop) before the LayerNorm.  The mask is the one torch's CUDA
 * dropout draws for the [num_nodes, hidden] tensor from generator state (seed, offset): element i is kept iff
 * uniform(Philox4x32-10(seed; counter offset/4 + (i/4)/T, subsequence (i/4) % T)[i % 4]) < 1 - p_drop, with T the
 * thread count of torch's launch (tools/dropout_stream_probe.py checks this mapping against F.dropout); kept
 * values are scaled by 1 / (1 - p_drop).  Needs num_nodes * hidden % 4 == 0 and offset % 4 == 0 (torch's vectorised
 * kernel).  The caller advances its generator by ghf_dropout_offset_advance(num_nodes * hidden), as torch would.
 * ghf_mp_epilogue_backward_dropout regenerates the same mask from (seed, offset): no mask tensor is stored. */
int64_t ghf_dropout_offset_advance(int64_t numel);
int ghf_mp_layer_dropout(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                         const float* d_W_msg, const float* d_W_self, const float* d_bias, const float* d_ln_w,
                         const float* d_ln_b, float eps, int precision, float p_drop, uint64_t seed, uint64_t offset,
                         float* d_out, void* d_out16, float* d_out16_scale, float* d_upd, void* d_workspace,
                         void* stream);
int ghf_mp_epilogue_backward_dropout(const ghf_graph* g, const float* d_g_out, const float* d_upd, const float* d_h,
                                     const float* d_ln_w, float eps, float p_drop, uint64_t seed, uint64_t offset,
                                     float* d_g_pre, float* d_g_acc, float* d_g_ln_w, float* d_g_ln_b,
                                     float* d_g_acc_scale, void* stream);
/* Gradients of the generated relation tensors (overwritten): g_W_msg[r] = sum_{e in r} h_u^T g_acc_v,
 * g_W_self[r] = sum_{e in r} h_v^T g_acc_v, g_bias[r] = sum_{e in r} g_acc_v  ([R,d,d], [R,d,d], [R,d]).
 * (d_h16, d_h16_scale), (d_g16, d_g16_scale): optional fp16 shadows of h and g_acc for the tensor-core path
 * (precision GHF_PREC_F16, hidden 128; made inside, in the workspace, when NULL). */
int ghf_mp_weight_grad(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                       const float* d_g_acc, const void* d_g16, const float* d_g16_scale, int precision,
                       float* d_gW_msg, float* d_gW_self, float* d_gbias, void* d_workspace, void* stream);
/* Backward of ghf_text_encode: parameter gradients g_emb [128,C], g_Wp [T,C], g_bp [T] (overwritten) from
 * d_out (the forward result) and d_g_out. */
int ghf_text_encode_backward(const uint8_t* d_utf8, const int64_t* d_offsets, const int64_t* d_string_index,
                             int64_t U, const float* d_emb, int C, const float* d_Wp, int T, const float* d_out,
                             const float* d_g_out, float* d_g_emb, float* d_g_Wp, float* d_g_bp, void* stream);

/* ---- generator -> operand-image fusion (hidden_dim 64 / 256, precision GHF_PREC_F16) -----------
 * The last Linear of the W_msg and W_self generators (WG:138-140) written directly as the fp16 operand images of the
 * contraction: no fp32 [R,d,d] tensors, no packing pass.  d_Zm / d_Zs [R,128]: the inputs of those two Linears (the
 * outputs of the generators' last hidden layer); d_W3* [d*d,128], d_b3* [d*d], d_ls* [1]: their parameters and
 * log-scales.  d_images: ghf_weight_images_bytes(R, d) bytes.  The per-relation power-of-two scales come from an
 * analytic bound (no pass over the generated values); the products use TF32 operands (the result is fp16).
 * ghf_mp_layer_images is ghf_mp_layer_f16 on such images (d_bias [R,d] fp32 from the third generator). */
int64_t ghf_weight_images_bytes(int64_t R, int32_t hidden_dim);
int ghf_weight_images_f16(const float* d_Zm, const float* d_Zs, int64_t R, const float* d_W3m, const float* d_b3m,
                          const float* d_lsm, const float* d_W3s, const float* d_b3s, const float* d_lss,
                          int32_t hidden_dim, void* d_images, void* stream);
int ghf_mp_layer_images(const ghf_graph* g, const float* d_h, const void* d_h16, const float* d_h16_scale,
                        const void* d_images, const float* d_bias, const float* d_ln_w, const float* d_ln_b, float eps,
                        float* d_out, void* d_out16, float* d_out16_scale, void* d_workspace, void* stream);

/* ---- batched link scores fused with the row gather (HG:304-318 applied to `embs[heads]`, `embs[tails]`, as
 * demo.py:90-94 does): out[b] = <emb[heads[b]], emb[tails[b]]>, emb [N,d] fp32, ids int64 in [0,N) (checked;
 * synchronises for the check).  The backward overwrites d_g_emb [N,d] with the scattered gradient. */
int ghf_score_pairs(const float* d_emb, int64_t N, int d, const int64_t* d_heads, const int64_t* d_tails, int64_t B,
                    float* d_out, void* stream);
int ghf_score_pairs_backward(const float* d_emb, int64_t N, int d, const int64_t* d_heads, const int64_t* d_tails,
                             int64_t B, const float* d_g_out, float* d_g_emb, void* stream);

/* ---- whole forward from HOST buffers (HG:236-298): the end-to-end entry point --------------
 * Parameters are passed as one flat array of DEVICE pointers in reference state_dict order
 * (see INTEGRATION.md for the exact list); inputs and output are HOST
 * buffers, copied inside the call.  Synchronises. */
typedef struct {
  int32_t text_dim, node_feat_dim, hidden_dim, num_layers, char_emb_dim, gen_hidden, gen_depth;
  int32_t precision;
  float ln_eps;
} ghf_model_desc;
int ghf_hypergnn_forward_host(const ghf_model_desc* desc, const float* const* d_params,
                              int64_t n_params, const float* h_node_features, int64_t num_nodes,
                              const int64_t* h_edge_index, int64_t E, const uint8_t* h_utf8,
                              const int64_t* h_offsets, float* h_out, void* stream);
/* The same forward on DEVICE buffers (no copies; stream-ordered except for the two size read-backs of dedup and
 * graph build): one call per forward instead of ~60 from the host language. */
int ghf_hypergnn_forward_device(const ghf_model_desc* desc, const float* const* d_params,
                                int64_t n_params, const float* d_node_features, int64_t num_nodes,
                                const int64_t* d_edge_index, int64_t E, const uint8_t* d_utf8,
                                const int64_t* d_offsets, float* d_out, void* stream);

/* ---- a8 for several generators at once (WG:120-143; north_star item 2: "the three per-layer MLPs run as batched
 * GEMMs over the R unique relations") ---------------------------------------------------------------------------
 * The L generators of a HyperGNN depend on the text embeddings only, so their hidden Linears are ONE grouped launch
 * per depth level (3 n_gen equally shaped problems), the bias heads another one, and the two [U, d_in*d_out] heads of
 * each generator go through ghf_linear (tcgen05 when large enough).
 * h_params: HOST array of n_gen x 3 x (depth+1) x {weight, bias} DEVICE pointers (MLP order W_msg, W_self, bias - the
 * flat order of INTEGRATION.md); h_log_scales: n_gen x 3 device pointers; h_out: n_gen x 3 device output pointers
 * (W_msg [U,d_in,d_out], W_self likewise, bias [U,d_out]).  d_scratch: ghf_weight_generators_scratch_bytes bytes; it
 * holds the hidden activations, the input of head m of generator g starting at float offset
 * ((depth-1)&1) * 3 n_gen U H + (3 g + m) U H.  skip_big != 0: the W_msg / W_self heads are left to the caller (who
 * writes operand images from those inputs, ghf_weight_images_f16). */
int64_t ghf_weight_generators_scratch_bytes(int64_t U, int32_t H, int32_t depth, int32_t n_gen);
int ghf_weight_generators(const float* d_text_emb, int64_t U, int32_t T, int32_t H, int32_t depth, int32_t n_gen,
                          int32_t d_in, int32_t d_out, const float* const* h_params,
                          const float* const* h_log_scales, float* const* h_out, void* d_scratch, int32_t skip_big,
                          void* stream);

/* ---- multi-GPU plumbing --------------------------------------------------------------------------
 * Stream-ordered copy between device buffers that may live on different GPUs of one box (peer-mapped / symmetric
 * memory): a rank pushes the fp16 rows it has just computed into every peer's copy of h16 over NVLink with the copy
 * engines - no SM is taken from the contraction kernel that is still running on the next chunk of rows. */
int ghf_copy_async(void* d_dst, const void* d_src, int64_t bytes, void* stream);

/* counters for bench.py: kernels launched by this library since the last reset */
int64_t ghf_launch_count(int reset);

/* Per-kernel device timing for the roofline report.  While enabled, ghf_mp_layer brackets its
 * contraction kernel and its epilogue kernel with CUDA events on the launching stream.
 * ghf_profile_read synchronises those events, returns the summed milliseconds and launch count
 * since the last read ([0]=contraction ms, [1]=epilogue ms, [2]=operand-pack + clear ms) and
 * resets the sums. */
int ghf_profile_enable(int on);
int ghf_profile_read(double ms[3], int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* GHF_B200_H */
