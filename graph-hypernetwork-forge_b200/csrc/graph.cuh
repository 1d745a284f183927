// graph.cuh — device-resident result of ghf_graph_build, shared with the message-passing kernels.
#pragma once

#include <stdint.h>

struct ghf_graph {
  int64_t num_edges_in = 0;   // E of the edge_index the graph was built from
  int64_t num_kept = 0;       // edges whose destination lies in [dst_lo, dst_hi)
  int64_t num_nodes = 0;      // N (global)
  int64_t dst_lo = 0, dst_hi = 0;
  int64_t num_local = 0;      // dst_hi - dst_lo
  int64_t num_units = 0;
  int32_t num_rel = 0, hidden_dim = 0, sb_nodes = 0, unit_edges = 0;
  // all arrays below are device memory owned by the graph
  int32_t* src_sorted = nullptr;   // [kept]  global source id at each sorted position
  int32_t* dst_sorted = nullptr;   // [kept]  LOCAL destination id (dst - dst_lo)
  int32_t* indeg = nullptr;        // [local] in-degree (multi-edges counted)
  int64_t* rowptr = nullptr;       // [local+1] exclusive scan of indeg (dst-CSR row pointer)
  int32_t* unit_start = nullptr;   // [units] first sorted position of the unit
  int32_t* unit_count = nullptr;   // [units] edges in the unit (<= unit_edges)
  int32_t* unit_rel = nullptr;     // [units] the one relation all its edges share
  int32_t* unit_phase = nullptr;   // [units] super-block ("phase") of the unit's destinations
  int32_t* phase_units = nullptr;  // [phases] number of units per super-block
  int32_t* phase_tiles = nullptr;  // [phases] number of 128-edge tiles per super-block (sum over its units)
  int32_t* h_phase_unit_begin = nullptr;   // HOST, [phases + 1]: first unit of each super-block (units are ordered by
                                           // super-block, then relation); lets a layer run on a range of super-blocks
  int64_t num_phases = 0;          // ceil(num_local / sb_nodes), at least 1
  int64_t bytes = 0;
  mutable void* stream = nullptr;  // stream the tables were allocated on / last used on (freed there)
};

#ifdef __cplusplus
#include <cuda_runtime.h>

#include <functional>
namespace ghf {
// ghf_dedup_texts with a hook that runs after the first kernels are enqueued and before the host waits (text.cu)
int dedup_texts_hooked(const uint8_t* d_utf8, const int64_t* d_offsets, int64_t E, const uint32_t* d_subset,
                       int64_t n_subset, int32_t* d_rel_ids, int64_t* d_first_edge, int64_t* h_num_unique,
                       cudaStream_t stream, const std::function<int()>* before_sync);
}  // namespace ghf
#endif
