// graph.cu — ghf_graph_build: in-degree, dst-CSR row pointer and the relation-grouped edge order.
//
// The reference never builds a graph structure: it gathers per-edge weights (HG:281-283) and
// scatter-adds by destination (HG:207-219) on every call.  Here edges are sorted ONCE by
// (super-block of dst, relation, dst) so that (1) runs of edges share one generated weight matrix
// (tensor-core tiles), (2) all destinations touched at any moment fit in L2 (the scatter side and
// the h[dst] gathers stay on chip).  Integer results are bit-exact with oracle.edge_order().
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"
#include "ghf_b200.h"
#include "graph.cuh"

namespace ghf {
namespace {

struct ToI64 {
  __host__ __device__ int64_t operator()(int32_t v) const { return (int64_t)v; }
};
struct MaxOp {
  __host__ __device__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

__global__ void keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                            const int32_t* __restrict__ rel, int64_t E, int64_t dst_lo, int64_t dst_hi,
                            int64_t sb, int64_t R, uint64_t invalid_key, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals, int32_t* __restrict__ indeg) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t v = dst[e];
  uint64_t key = invalid_key;
  if (v >= dst_lo && v < dst_hi) {
    const int64_t dl = v - dst_lo;
    key = (uint64_t)(((dl / sb) * R + rel[e]) * sb + dl % sb);
    atomicAdd(&indeg[dl], 1);
  }
  keys[e] = key;
  vals[e] = (uint32_t)e;
}

// sorted position i -> (perm, src, local dst), plus "i if i starts a (super-block, relation) group"
__global__ void gather_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                              const int64_t* __restrict__ src, int64_t kept, int64_t sb, int64_t R,
                              int64_t* __restrict__ perm, int32_t* __restrict__ src_sorted,
                              int32_t* __restrict__ dst_sorted, int32_t* __restrict__ gstart_in) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= kept) return;
  const uint64_t key = keys[i];
  const uint32_t e = vals[i];
  const uint64_t g = key / sb;
  perm[i] = e;
  src_sorted[i] = (int32_t)src[e];
  dst_sorted[i] = (int32_t)((g / R) * sb + key % sb);
  const bool head = (i == 0) || (keys[i - 1] / sb != g);
  gstart_in[i] = head ? (int32_t)i : 0;
}

__global__ void unit_flag_kernel(const int32_t* __restrict__ gstart, int64_t kept, int32_t unit_edges,
                                 int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < kept) flag[i] = ((int32_t)i - gstart[i]) % unit_edges == 0;
}

__global__ void unit_fill_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ uidx,
                                 const uint64_t* __restrict__ keys, int64_t kept, int64_t sb, int64_t R,
                                 int32_t* __restrict__ unit_start, int32_t* __restrict__ unit_rel,
                                 int32_t* __restrict__ unit_phase, int32_t* __restrict__ phase_units) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= kept || !flag[i]) return;
  const int32_t u = uidx[i];
  const uint64_t g = keys[i] / sb;
  unit_start[u] = (int32_t)i;
  unit_rel[u] = (int32_t)(g % R);
  unit_phase[u] = (int32_t)(g / R);
  atomicAdd(&phase_units[g / R], 1);
}

__global__ void unit_count_kernel(const int32_t* __restrict__ unit_start, int64_t units, int64_t kept,
                                  int32_t* __restrict__ unit_count) {
  const int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (u < units) unit_count[u] = (u + 1 < units ? unit_start[u + 1] : (int32_t)kept) - unit_start[u];
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
  void* ptrs[] = {g->src_sorted, g->dst_sorted, g->perm, g->indeg, g->rowptr, g->unit_start, g->unit_count,
                  g->unit_rel, g->unit_phase, g->phase_units};
  for (void* p : ptrs)
    if (p) cudaFreeAsync(p, s);
  delete g;
}

static int graph_build_impl(ghf_graph* g, const int64_t* d_edge_index, const int32_t* d_rel_ids,
                            cudaStream_t stream) {
  const int64_t E = g->num_edges_in, sb = g->sb_nodes, R = g->num_rel, nl = g->num_local;
  const int threads = 256;
  const int64_t n_sb = cdiv(nl > 0 ? nl : 1, sb);
  const uint64_t invalid_key = (uint64_t)n_sb * R * sb;  // sorts after every valid key
  const int end_bit = bits_for(invalid_key);

  g->num_phases = n_sb;
  GHF_CUDA(dmalloc(&g->indeg, nl, g));
  GHF_CUDA(dmalloc(&g->rowptr, nl + 1, g));
  GHF_CUDA(dmalloc(&g->phase_units, n_sb, g));
  GHF_CUDA(cudaMemsetAsync(g->phase_units, 0, (size_t)n_sb * sizeof(int32_t), stream));
  GHF_CUDA(cudaMemsetAsync(g->indeg, 0, (size_t)(nl > 0 ? nl : 1) * sizeof(int32_t), stream));

  TempBuf keys_a, keys_b, vals_a, vals_b, tmp;
  const int64_t En = E > 0 ? E : 1;
  GHF_CUDA(keys_a.alloc(En * sizeof(uint64_t), stream));
  GHF_CUDA(keys_b.alloc(En * sizeof(uint64_t), stream));
  GHF_CUDA(vals_a.alloc(En * sizeof(uint32_t), stream));
  GHF_CUDA(vals_b.alloc(En * sizeof(uint32_t), stream));

  const int64_t* src = d_edge_index;
  const int64_t* dst = d_edge_index + E;
  if (E > 0) {
    keys_kernel<<<(unsigned)cdiv(E, threads), threads, 0, stream>>>(
        src, dst, d_rel_ids, E, g->dst_lo, g->dst_hi, sb, R, invalid_key, keys_a.as<uint64_t>(),
        vals_a.as<uint32_t>(), g->indeg);
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
  cub::DoubleBuffer<uint64_t> kbuf(keys_a.as<uint64_t>(), keys_b.as<uint64_t>());
  cub::DoubleBuffer<uint32_t> vbuf(vals_a.as<uint32_t>(), vals_b.as<uint32_t>());
  if (E > 0) {
    size_t bytes = 0;
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kbuf, vbuf, (int)E, 0, end_bit, stream));
    GHF_CUDA(tmp.alloc(bytes, stream));
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, kbuf, vbuf, (int)E, 0, end_bit, stream));
    g_launches.fetch_add((end_bit + 7) / 8 + 2, std::memory_order_relaxed);
  }
  // kept = sum of in-degrees: last rowptr entry.  rowptr[nl] = rowptr[nl-1] + indeg[nl-1].
  int64_t tail[2] = {0, 0};
  int32_t last_deg = 0;
  if (nl > 0) {
    GHF_CUDA(cudaMemcpyAsync(&tail[0], g->rowptr + (nl - 1), sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    GHF_CUDA(cudaMemcpyAsync(&last_deg, g->indeg + (nl - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  }
  GHF_CUDA(cudaStreamSynchronize(stream));
  const int64_t kept = tail[0] + last_deg;
  g->num_kept = kept;
  GHF_CUDA(cudaMemcpyAsync(g->rowptr + nl, &g->num_kept, sizeof(int64_t), cudaMemcpyHostToDevice, stream));

  GHF_CUDA(dmalloc(&g->src_sorted, kept, g));
  GHF_CUDA(dmalloc(&g->dst_sorted, kept, g));
  GHF_CUDA(dmalloc(&g->perm, kept, g));
  if (kept == 0) {
    g->num_units = 0;
    GHF_CUDA(dmalloc(&g->unit_start, 1, g));
    GHF_CUDA(dmalloc(&g->unit_count, 1, g));
    GHF_CUDA(dmalloc(&g->unit_rel, 1, g));
    GHF_CUDA(dmalloc(&g->unit_phase, 1, g));
    return 0;
  }
  TempBuf gstart, flag, uidx, total;
  GHF_CUDA(gstart.alloc(kept * sizeof(int32_t), stream));
  GHF_CUDA(flag.alloc(kept * sizeof(int32_t), stream));
  GHF_CUDA(uidx.alloc((kept + 1) * sizeof(int32_t), stream));
  const unsigned kblocks = (unsigned)cdiv(kept, threads);
  gather_kernel<<<kblocks, threads, 0, stream>>>(kbuf.Current(), vbuf.Current(), src, kept, sb, R, g->perm,
                                                 g->src_sorted, g->dst_sorted, gstart.as<int32_t>());
  GHF_LAUNCH_CHECK();
  {
    size_t bytes = 0;
    GHF_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, gstart.as<int32_t>(), gstart.as<int32_t>(), MaxOp(),
                                            (int)kept, stream));
    TempBuf t;
    GHF_CUDA(t.alloc(bytes, stream));
    GHF_CUDA(cub::DeviceScan::InclusiveScan(t.p, bytes, gstart.as<int32_t>(), gstart.as<int32_t>(), MaxOp(),
                                            (int)kept, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  unit_flag_kernel<<<kblocks, threads, 0, stream>>>(gstart.as<int32_t>(), kept, g->unit_edges, flag.as<int32_t>());
  GHF_LAUNCH_CHECK();
  {
    size_t bytes = 0;
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.as<int32_t>(), uidx.as<int32_t>(), (int)kept, stream));
    TempBuf t;
    GHF_CUDA(t.alloc(bytes, stream));
    GHF_CUDA(cub::DeviceScan::ExclusiveSum(t.p, bytes, flag.as<int32_t>(), uidx.as<int32_t>(), (int)kept, stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  int32_t last_idx = 0, last_flag = 0;
  GHF_CUDA(cudaMemcpyAsync(&last_idx, uidx.as<int32_t>() + (kept - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GHF_CUDA(cudaMemcpyAsync(&last_flag, flag.as<int32_t>() + (kept - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  GHF_CUDA(cudaStreamSynchronize(stream));
  g->num_units = (int64_t)last_idx + last_flag;
  GHF_CUDA(dmalloc(&g->unit_start, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_count, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_rel, g->num_units, g));
  GHF_CUDA(dmalloc(&g->unit_phase, g->num_units, g));
  unit_fill_kernel<<<kblocks, threads, 0, stream>>>(flag.as<int32_t>(), uidx.as<int32_t>(), kbuf.Current(), kept, sb,
                                                    R, g->unit_start, g->unit_rel, g->unit_phase, g->phase_units);
  GHF_LAUNCH_CHECK();
  unit_count_kernel<<<(unsigned)cdiv(g->num_units, threads), threads, 0, stream>>>(g->unit_start, g->num_units, kept,
                                                                                 g->unit_count);
  GHF_LAUNCH_CHECK();
  return 0;
}

extern "C" int ghf_graph_build(const int64_t* d_edge_index, const int32_t* d_rel_ids, int64_t E,
                               int64_t num_nodes, int32_t num_rel, int32_t hidden_dim, int64_t dst_lo,
                               int64_t dst_hi, int32_t sb_nodes, int32_t unit_edges, ghf_graph** out,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
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
    // destinations of one super-block: h[dst] rows + accumulator rows (2 * d * 4 B per node) ~ 48 MiB of L2
    int64_t s = ((int64_t)48 << 20) / (8 * (int64_t)hidden_dim);
    sb_nodes = (int32_t)(s < 1024 ? 1024 : s);
  }
  g->sb_nodes = sb_nodes;
  g->unit_edges = unit_edges > 0 ? unit_edges : 1024;
  const int rc = graph_build_impl(g, d_edge_index, d_rel_ids, stream);
  if (rc != 0) {
    ghf_graph_free(g);
    return rc;
  }
  *out = g;
  return 0;
}

extern "C" int ghf_graph_info(const ghf_graph* g, int64_t info[6]) {
  GHF_REQUIRE(g && info, "ghf_graph_info: NULL argument");
  info[0] = g->num_kept; info[1] = g->num_units; info[2] = g->sb_nodes; info[3] = g->unit_edges;
  info[4] = g->num_local; info[5] = g->bytes;
  return 0;
}

extern "C" int ghf_graph_export(const ghf_graph* g, int64_t* d_perm, int32_t* d_indeg, int64_t* d_rowptr,
                                int32_t* d_unit_start, int32_t* d_unit_count, int32_t* d_unit_rel,
                                void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(g != nullptr, "ghf_graph_export: graph is NULL");
  auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
    return (dst && bytes) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream) : cudaSuccess;
  };
  GHF_CUDA(cp(d_perm, g->perm, g->num_kept * sizeof(int64_t)));
  GHF_CUDA(cp(d_indeg, g->indeg, g->num_local * sizeof(int32_t)));
  GHF_CUDA(cp(d_rowptr, g->rowptr, (g->num_local + 1) * sizeof(int64_t)));
  GHF_CUDA(cp(d_unit_start, g->unit_start, g->num_units * sizeof(int32_t)));
  GHF_CUDA(cp(d_unit_count, g->unit_count, g->num_units * sizeof(int32_t)));
  GHF_CUDA(cp(d_unit_rel, g->unit_rel, g->num_units * sizeof(int32_t)));
  return 0;
}
