// graph.cu — ghf_graph_build: in-degree, dst-CSR row pointer and the relation-grouped edge order.
//
// The reference never builds a graph structure: it gathers per-edge weights (HG:281-283) and
// scatter-adds by destination (HG:207-219) on every call.  Here edges are sorted ONCE by
// (super-block of dst, relation, dst) so that (1) runs of edges share one generated weight matrix
// (tensor-core tiles), (2) all destinations touched at any moment fit in L2 (the scatter side and
// the h[dst] gathers stay on chip).  Integer results are bit-exact with oracle.edge_order().
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <cstdlib>

#include "common.cuh"
#include "ghf_b200.h"
#include "graph.cuh"

namespace ghf {
namespace {

struct ToI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};
// edges of a (super-block, relation) group -> work units of that group
struct UnitsOf {
  int32_t unit_edges;
  __host__ __device__ int32_t operator()(int32_t count) const { return (count + unit_edges - 1) / unit_edges; }
};

// Sort key of edge e: ((super-block of dst) * R + relation) * sb + dst % sb, or `invalid_key` (sorts last) when the
// destination lies outside [dst_lo, dst_hi).  Also counts in-degrees and the edges of every (super-block,
// relation) group, from which the unit table follows without touching the sorted edges again.
// Work item j is edge `edge_ids[j]` (a pre-selected subset, relation ids indexed by j) or edge j itself.
// The sort payload is the SOURCE id (src != NULL): the sorted payload then IS src_sorted, no random gather after
// the sort.  With src == NULL the payload is the edge id and nothing is counted (ghf_graph_export recomputes the
// permutation that way, for parity checks only).
template <class Key>
__global__ void keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                            const int32_t* __restrict__ rel, const uint32_t* __restrict__ edge_ids, int64_t n,
                            int64_t dst_lo, int64_t dst_hi, int64_t sb, int64_t R, uint64_t invalid_key,
                            Key* __restrict__ keys, uint32_t* __restrict__ vals, int32_t* __restrict__ indeg,
                            int32_t* __restrict__ group_count, int64_t num_nodes, int32_t* __restrict__ bad_ids) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t e = edge_ids ? edge_ids[j] : j;
  const int64_t v = dst[e];
  if (src) {  // node ids outside [0, N) would make the gathers read out of bounds: reported by ghf_graph_build
    const int64_t u = src[e];
    if (u < 0 || u >= num_nodes || v < 0 || v >= num_nodes) atomicOr(bad_ids, 1);
    if (rel[j] < 0 || rel[j] >= R) atomicOr(bad_ids, 2);
  }
  uint64_t key = invalid_key;
  if (v >= dst_lo && v < dst_hi && rel[j] >= 0 && rel[j] < R) {
    const int64_t dl = v - dst_lo;
    const int64_t grp = (dl / sb) * R + rel[j];
    key = (uint64_t)(grp * sb + dl % sb);
    if (src) {
      atomicAdd(&indeg[dl], 1);
      atomicAdd(&group_count[grp], 1);
    }
  }
  keys[j] = (Key)key;
  vals[j] = src ? (uint32_t)src[e] : (uint32_t)e;   // (an out-of-range source id never reaches a kernel: the build fails)
}

struct InDstRange {
  const int64_t* dst;
  int64_t lo, hi;
  __host__ __device__ bool operator()(uint32_t e) const { return dst[e] >= lo && dst[e] < hi; }
};

// sorted position i -> local destination id, straight from the key
template <class Key>
__global__ void dst_from_keys_kernel(const Key* __restrict__ keys, int64_t kept, int64_t sb, int64_t R,
                                     int32_t* __restrict__ dst_sorted) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= kept) return;
  const uint64_t key = keys[i];
  const uint64_t g = key / sb;
  dst_sorted[i] = (int32_t)((g / R) * sb + key % sb);
}

__global__ void widen_kernel(const uint32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// one thread per (super-block, relation) group: its units, in order
__global__ void unit_fill_kernel(const int32_t* __restrict__ group_count, const int32_t* __restrict__ group_start,
                                 const int32_t* __restrict__ unit_base, int64_t groups, int64_t R,
                                 int32_t unit_edges, int32_t* __restrict__ unit_start,
                                 int32_t* __restrict__ unit_count, int32_t* __restrict__ unit_rel,
                                 int32_t* __restrict__ unit_phase, int32_t* __restrict__ phase_units,
                                 int32_t* __restrict__ phase_tiles) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= groups) return;
  const int32_t cnt = group_count[g];
  if (cnt == 0) return;
  const int32_t gs = group_start[g], ub = unit_base[g];
  const int32_t n = (cnt + unit_edges - 1) / unit_edges;
  int32_t tiles = 0;
  for (int32_t j = 0; j < n; ++j) {
    unit_start[ub + j] = gs + j * unit_edges;
    unit_count[ub + j] = min(unit_edges, cnt - j * unit_edges);
    tiles += (min(unit_edges, cnt - j * unit_edges) + 127) / 128;
    unit_rel[ub + j] = (int32_t)(g % R);
    unit_phase[ub + j] = (int32_t)(g / R);
  }
  atomicAdd(&phase_units[g / R], n);
  atomicAdd(&phase_tiles[g / R], tiles);
}

int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

// graph tables come from the stream-ordered pool (cached between calls: cudaMalloc/cudaFree cost
// milliseconds per graph at 16M edges)
template <class T>
cudaError_t dmalloc(T** p, int64_t n, ghf_graph* g) {
  const size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(T);
  g->bytes += bytes;
  return cudaMallocAsync(reinterpret_cast<void**>(p), bytes, (cudaStream_t)g->stream);
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" void ghf_graph_free(ghf_graph* g) {
  if (!g) return;
  cudaStream_t s = (cudaStream_t)g->stream;  // the stream that last used the tables
  void* ptrs[] = {g->src_sorted, g->dst_sorted, g->indeg, g->rowptr, g->unit_start, g->unit_count,
                  g->unit_rel, g->unit_phase, g->phase_units, g->phase_tiles};
  for (void* p : ptrs)
    if (p) cudaFreeAsync(p, s);
  delete[] g->h_phase_unit_begin;
  delete g;
}

template <class Key>
static int graph_build_impl(ghf_graph* g, const int64_t* d_edge_index, const uint32_t* d_edge_ids, int64_t n_items,
                            const int32_t* d_rel_ids, cudaStream_t stream) {
  // E: edges of edge_index (src = edge_index[0..E), dst = edge_index[E..2E)); n: work items (all edges, or the subset)
  const int64_t E = g->num_edges_in, n = n_items, sb = g->sb_nodes, R = g->num_rel, nl = g->num_local;
  const int threads = 256;
  const int64_t n_sb = cdiv(nl > 0 ? nl : 1, sb);
  const int64_t groups = n_sb * R;
  const uint64_t invalid_key = (uint64_t)n_sb * R * sb;  // sorts after every valid key
  const int end_bit = bits_for(invalid_key);
  GHF_REQUIRE(groups < (int64_t)0x7FFFFFFF, "ghf_graph_build: %lld (super-block, relation) groups", (long long)groups);

  g->num_phases = n_sb;
  GHF_CUDA(dmalloc(&g->indeg, nl, g));
  GHF_CUDA(dmalloc(&g->rowptr, nl + 1, g));
  GHF_CUDA(dmalloc(&g->phase_units, n_sb, g));
  GHF_CUDA(cudaMemsetAsync(g->phase_units, 0, (size_t)n_sb * sizeof(int32_t), stream));
  GHF_CUDA(dmalloc(&g->phase_tiles, n_sb, g));
  GHF_CUDA(cudaMemsetAsync(g->phase_tiles, 0, (size_t)n_sb * sizeof(int32_t), stream));
  GHF_CUDA(cudaMemsetAsync(g->indeg, 0, (size_t)(nl > 0 ? nl : 1) * sizeof(int32_t), stream));

  TempBuf keys_a, keys_b, vals_a, vals_b, tmp, gcount, gstart, ubase;
  const int64_t En = n > 0 ? n : 1;
  GHF_CUDA(keys_a.alloc(En * sizeof(Key), stream));
  GHF_CUDA(keys_b.alloc(En * sizeof(Key), stream));
  GHF_CUDA(vals_a.alloc(En * sizeof(uint32_t), stream));   // the payload (source ids): whichever of the two buffers
  GHF_CUDA(vals_b.alloc(En * sizeof(uint32_t), stream));   // the sort ends in becomes the graph's src_sorted
  GHF_CUDA(gcount.alloc((groups + 1) * sizeof(int32_t), stream));   // one trailing zero: scans yield the totals
  GHF_CUDA(gstart.alloc((groups + 1) * sizeof(int32_t), stream));
  GHF_CUDA(ubase.alloc((groups + 1) * sizeof(int32_t), stream));
  GHF_CUDA(cudaMemsetAsync(gcount.p, 0, (size_t)(groups + 1) * sizeof(int32_t), stream));
  TempBuf bad;
  GHF_CUDA(bad.alloc(sizeof(int32_t), stream));
  GHF_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int32_t), stream));

  const int64_t* src = d_edge_index;
  const int64_t* dst = d_edge_index + E;
  if (n > 0) {
    keys_kernel<Key><<<(unsigned)cdiv(n, threads), threads, 0, stream>>>(
        src, dst, d_rel_ids, d_edge_ids, n, g->dst_lo, g->dst_hi, sb, R, invalid_key, keys_a.as<Key>(),
        vals_a.as<uint32_t>(), g->indeg, gcount.as<int32_t>(), g->num_nodes, bad.as<int32_t>());
    GHF_LAUNCH_CHECK();
  }
  // rowptr[0..local) = exclusive scan of in-degree (int64); rowptr[local] = kept is written below
  {
    cub::TransformInputIterator<int64_t, ToI64, const int32_t*> it(g->indeg, ToI64());
    size_t bytes = 0;
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, g->rowptr, (int)nl, stream));
    TempBuf t;
    GHF_CUDA(t.alloc(bytes, stream));
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(t.p, bytes, it, g->rowptr, (int)nl, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  // group_start = exclusive scan of the group sizes (last entry: kept edges); unit_base likewise over the unit
  // counts (last entry: number of units)
  {
    size_t b1 = 0, b2 = 0;
    cub::TransformInputIterator<int32_t, UnitsOf, const int32_t*> units_it(gcount.as<int32_t>(),
                                                                           UnitsOf{g->unit_edges});
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b1, gcount.as<int32_t>(), gstart.as<int32_t>(), (int)(groups + 1),
                                           stream));
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, units_it, ubase.as<int32_t>(), (int)(groups + 1), stream));
    TempBuf t;
    GHF_CUDA(t.alloc(b1 > b2 ? b1 : b2, stream));
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(t.p, b1, gcount.as<int32_t>(), gstart.as<int32_t>(), (int)(groups + 1),
                                           stream));
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(t.p, b2, units_it, ubase.as<int32_t>(), (int)(groups + 1), stream));
    g_launches.fetch_add(2, std::memory_order_relaxed);
  }
  int32_t totals[2] = {0, 0}, bad_ids = 0;  // kept edges, units, id check: the one host round trip of the build
  GHF_CUDA(cudaMemcpyAsync(&bad_ids, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GHF_CUDA(cudaMemcpyAsync(&totals[0], gstart.as<int32_t>() + groups, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GHF_CUDA(cudaMemcpyAsync(&totals[1], ubase.as<int32_t>() + groups, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  // first unit of every super-block = unit_base of its first group (groups are ordered super-block, relation)
  g->h_phase_unit_begin = new int32_t[n_sb + 1]();
  GHF_CUDA(cudaMemcpy2DAsync(g->h_phase_unit_begin, sizeof(int32_t), ubase.as<int32_t>(), (size_t)R * sizeof(int32_t),
                             sizeof(int32_t), (size_t)n_sb + 1, cudaMemcpyDeviceToHost, stream));

  cub::DoubleBuffer<Key> kbuf(keys_a.as<Key>(), keys_b.as<Key>());
  cub::DoubleBuffer<uint32_t> vbuf(vals_a.as<uint32_t>(), vals_b.as<uint32_t>());
  if (n > 0) {
    size_t bytes = 0;
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kbuf, vbuf, (int)n, 0, end_bit, stream));
    GHF_CUDA(tmp.alloc(bytes, stream));
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kbuf, vbuf, (int)n, 0, end_bit, stream));
    g_launches.fetch_add((end_bit + 7) / 8 + 2, std::memory_order_relaxed);
  }
  if (int rc = readback_wait(stream)) return rc;         // runs the caller's one-shot hook (ghf_set_presync_hook)
  GHF_REQUIRE((bad_ids & 1) == 0, "ghf_graph_build: edge_index holds node ids outside [0, %lld)",
              (long long)g->num_nodes);
  GHF_REQUIRE((bad_ids & 2) == 0, "ghf_graph_build: relation ids outside [0, %d)", g->num_rel);
  const int64_t kept = totals[0];
  g->num_kept = kept;
  g->num_units = totals[1];
  GHF_CUDA(cudaMemcpyAsync(g->rowptr + nl, &g->num_kept, sizeof(int64_t), cudaMemcpyHostToDevice, stream));

  {  // adopt the sorted payload as src_sorted (its first `kept` entries; the rest belongs to dropped edges)
    TempBuf& cur = vbuf.Current() == vals_a.as<uint32_t>() ? vals_a : vals_b;
    g->src_sorted = reinterpret_cast<int32_t*>(cur.p);
    g->bytes += (size_t)En * sizeof(uint32_t);
    cur.p = nullptr;   // ownership moved to the graph (freed in ghf_graph_free)
  }
  GHF_CUDA(dmalloc(&g->dst_sorted, kept, g));
  GHF_CUDA(dmalloc(&g->unit_start, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_count, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_rel, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_phase, g->num_units, g));
  if (kept == 0) return 0;
  dst_from_keys_kernel<Key><<<(unsigned)cdiv(kept, threads), threads, 0, stream>>>(kbuf.Current(), kept, sb, R,
                                                                                     g->dst_sorted);
  GHF_LAUNCH_CHECK();
  unit_fill_kernel<<<(unsigned)cdiv(groups, threads), threads, 0, stream>>>(
      gcount.as<int32_t>(), gstart.as<int32_t>(), ubase.as<int32_t>(), groups, R, g->unit_edges, g->unit_start,
      g->unit_count, g->unit_rel, g->unit_phase, g->phase_units, g->phase_tiles);
  GHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int ghf_select_edges(const int64_t* d_edge_index, int64_t E, int64_t dst_lo, int64_t dst_hi,
                                uint32_t* d_edge_ids, int64_t* h_count, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(E >= 0 && E < (int64_t)0x7FFFFFFF, "ghf_select_edges: E=%lld out of range", (long long)E);
  GHF_REQUIRE(h_count != nullptr, "ghf_select_edges: h_count is NULL");
  *h_count = 0;
  if (E == 0) return 0;
  TempBuf count, tmp;
  GHF_CUDA(count.alloc(sizeof(int64_t), stream));
  cub::CountingInputIterator<uint32_t> ids(0u);
  InDstRange pred{d_edge_index + E, dst_lo, dst_hi};
  size_t bytes = 0;
  GHF_CUDA(cub::DeviceSelect::If(nullptr, bytes, ids, d_edge_ids, count.as<int64_t>(), (int)E, pred, stream));
  GHF_CUDA(tmp.alloc(bytes, stream));
  GHF_CUDA(cub::DeviceSelect::If(tmp.p, bytes, ids, d_edge_ids, count.as<int64_t>(), (int)E, pred, stream));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  GHF_CUDA(cudaMemcpyAsync(h_count, count.p, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
  if (int rc = readback_wait(stream)) return rc;         // runs the caller's one-shot hook (ghf_set_presync_hook)
  return 0;
}

extern "C" int ghf_graph_build(const int64_t* d_edge_index, int64_t E, const uint32_t* d_edge_ids, int64_t n_subset,
                               const int32_t* d_rel_ids, int64_t num_nodes, int32_t num_rel, int32_t hidden_dim,
                               int64_t dst_lo, int64_t dst_hi, int32_t sb_nodes, int32_t unit_edges,
                               ghf_graph** out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(d_edge_ids == nullptr || (n_subset >= 0 && n_subset <= E), "ghf_graph_build: bad subset size");
  const int64_t n_items = d_edge_ids ? n_subset : E;
  GHF_REQUIRE(out != nullptr, "ghf_graph_build: out is NULL");
  GHF_REQUIRE(E >= 0 && E < (int64_t)0x7FFFFFFF, "ghf_graph_build: E=%lld out of range", (long long)E);
  GHF_REQUIRE(num_nodes >= 0 && num_nodes < (int64_t)0x7FFFFFFF, "ghf_graph_build: N=%lld out of range",
              (long long)num_nodes);
  GHF_REQUIRE(num_rel > 0 || E == 0, "ghf_graph_build: num_rel=%d", num_rel);
  GHF_REQUIRE(hidden_dim > 0, "ghf_graph_build: hidden_dim=%d", hidden_dim);
  GHF_REQUIRE(0 <= dst_lo && dst_lo <= dst_hi && dst_hi <= num_nodes,
              "ghf_graph_build: bad dst range [%lld,%lld) for N=%lld", (long long)dst_lo, (long long)dst_hi,
              (long long)num_nodes);
  ghf_graph* g = new ghf_graph();
  g->stream = stream_;
  g->num_edges_in = E; g->num_nodes = num_nodes; g->dst_lo = dst_lo; g->dst_hi = dst_hi;
  g->num_local = dst_hi - dst_lo; g->num_rel = num_rel > 0 ? num_rel : 1; g->hidden_dim = hidden_dim;
  if (sb_nodes <= 0) {
    if (const char* env = getenv("GHF_SB_NODES")) sb_nodes = atoi(env);   // tuning knob for the whole-forward entries
  }
  if (sb_nodes <= 0) {
    // Super-block size.  Small enough: the accumulator rows and the h[dst] rows of one super-block
    // (2 * d * 4 B per node) fit in ~48 MiB of L2, so reductions and destination gathers stay on chip.  But every
    // super-block re-reads the generated weights of the relations it touches: with many relations and few edges
    // per relation (BASELINE config 4: 20k relations, 10.5 GB of weights per layer) that re-read costs more HBM
    // traffic than letting the reductions go to HBM (a read-modify-write of 4d bytes per edge) in ONE block.
    const int64_t nl = dst_hi - dst_lo > 0 ? dst_hi - dst_lo : 1;
    int64_t s = ((int64_t)48 << 20) / (8 * (int64_t)hidden_dim);
    s = s < 1024 ? 1024 : s;
    const int64_t n_sb = cdiv(nl, s);
    const double w_bytes = (double)num_rel * 2.0 * hidden_dim * hidden_dim * 4.0;
    const bool weights_stay_in_l2 = w_bytes < 48.0 * (1 << 20);
    const double tiled = weights_stay_in_l2 ? w_bytes : (double)n_sb * w_bytes;
    const double single = w_bytes + 2.0 * (double)E * hidden_dim * 4.0;
    if (n_sb > 1 && single < tiled) s = nl;
    sb_nodes = (int32_t)(s > 0x7FFFFFF0 ? 0x7FFFFFF0 : s);
  }
  g->sb_nodes = sb_nodes;
  g->unit_edges = unit_edges > 0 ? unit_edges : 1024;
  // 32-bit sort keys whenever (super-blocks * relations * sb_nodes) fits: half the radix-sort traffic
  const uint64_t key_range = (uint64_t)cdiv(g->num_local > 0 ? g->num_local : 1, g->sb_nodes) * g->num_rel *
                             (uint64_t)g->sb_nodes;
  const int rc = key_range < 0xFFFFFFFFull
                     ? graph_build_impl<uint32_t>(g, d_edge_index, d_edge_ids, n_items, d_rel_ids, stream)
                     : graph_build_impl<uint64_t>(g, d_edge_index, d_edge_ids, n_items, d_rel_ids, stream);
  if (rc != 0) {
    ghf_graph_free(g);
    return rc;
  }
  *out = g;
  return 0;
}

extern "C" int64_t ghf_graph_num_phases(const ghf_graph* g) { return g ? g->num_phases : -1; }

extern "C" int ghf_graph_info(const ghf_graph* g, int64_t info[6]) {
  GHF_REQUIRE(g && info, "ghf_graph_info: NULL argument");
  info[0] = g->num_kept; info[1] = g->num_units; info[2] = g->sb_nodes; info[3] = g->unit_edges;
  info[4] = g->num_local; info[5] = g->bytes;
  return 0;
}

// The permutation (original edge id at each sorted position) is not needed by any kernel; parity checks ask for it
// through ghf_graph_export, which re-sorts (key, edge id) from the build inputs.
template <class Key>
static int recompute_perm(const ghf_graph* g, const int64_t* d_edge_index, const uint32_t* d_edge_ids, int64_t n,
                          const int32_t* d_rel_ids, int64_t* d_perm, cudaStream_t stream) {
  const int64_t E = g->num_edges_in, sb = g->sb_nodes, R = g->num_rel, nl = g->num_local;
  const int64_t n_sb = cdiv(nl > 0 ? nl : 1, sb);
  const uint64_t invalid_key = (uint64_t)n_sb * R * sb;
  const int end_bit = bits_for(invalid_key);
  if (n == 0 || g->num_kept == 0) return 0;
  TempBuf keys_a, keys_b, vals_a, vals_b, tmp;
  GHF_CUDA(keys_a.alloc(n * sizeof(Key), stream));
  GHF_CUDA(keys_b.alloc(n * sizeof(Key), stream));
  GHF_CUDA(vals_a.alloc(n * sizeof(uint32_t), stream));
  GHF_CUDA(vals_b.alloc(n * sizeof(uint32_t), stream));
  const int threads = 256;
  keys_kernel<Key><<<(unsigned)cdiv(n, threads), threads, 0, stream>>>(
      nullptr, d_edge_index + E, d_rel_ids, d_edge_ids, n, g->dst_lo, g->dst_hi, sb, R, invalid_key, keys_a.as<Key>(),
      vals_a.as<uint32_t>(), nullptr, nullptr, g->num_nodes, nullptr);
  GHF_LAUNCH_CHECK();
  cub::DoubleBuffer<Key> kbuf(keys_a.as<Key>(), keys_b.as<Key>());
  cub::DoubleBuffer<uint32_t> vbuf(vals_a.as<uint32_t>(), vals_b.as<uint32_t>());
  size_t bytes = 0;
  GHF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kbuf, vbuf, (int)n, 0, end_bit, stream));
  GHF_CUDA(tmp.alloc(bytes, stream));
  GHF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kbuf, vbuf, (int)n, 0, end_bit, stream));
  widen_kernel<<<(unsigned)cdiv(g->num_kept, threads), threads, 0, stream>>>(vbuf.Current(), g->num_kept, d_perm);
  GHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int ghf_graph_export(const ghf_graph* g, const int64_t* d_edge_index, const uint32_t* d_edge_ids,
                                int64_t n_subset, const int32_t* d_rel_ids, int64_t* d_perm, int32_t* d_indeg,
                                int64_t* d_rowptr, int32_t* d_unit_start, int32_t* d_unit_count,
                                int32_t* d_unit_rel, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(g != nullptr, "ghf_graph_export: graph is NULL");
  if (d_perm) {
    GHF_REQUIRE(d_edge_index != nullptr && d_rel_ids != nullptr,
                "ghf_graph_export: the permutation is recomputed from the build inputs (edge_index, rel_ids)");
    const int64_t n_items = d_edge_ids ? n_subset : g->num_edges_in;
    const uint64_t key_range = (uint64_t)cdiv(g->num_local > 0 ? g->num_local : 1, g->sb_nodes) * g->num_rel *
                               (uint64_t)g->sb_nodes;
    if (int rc = key_range < 0xFFFFFFFFull
                     ? recompute_perm<uint32_t>(g, d_edge_index, d_edge_ids, n_items, d_rel_ids, d_perm, stream)
                     : recompute_perm<uint64_t>(g, d_edge_index, d_edge_ids, n_items, d_rel_ids, d_perm, stream))
      return rc;
  }
  auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream) : cudaSuccess;
  };
  GHF_CUDA(cp(d_indeg, g->indeg, g->num_local * sizeof(int32_t)));
  GHF_CUDA(cp(d_rowptr, g->rowptr, (g->num_local + 1) * sizeof(int64_t)));
  GHF_CUDA(cp(d_unit_start, g->unit_start, g->num_units * sizeof(int32_t)));
  GHF_CUDA(cp(d_unit_count, g->unit_count, g->num_units * sizeof(int32_t)));
  GHF_CUDA(cp(d_unit_rel, g->unit_rel, g->num_units * sizeof(int32_t)));
  return 0;
}
