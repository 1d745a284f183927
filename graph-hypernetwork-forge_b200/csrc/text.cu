// text.cu — relation-string dedup (HG:264-268) and the character-bag text encoder (HG:66-81).
//
// Strings arrive packed: UTF-8 bytes + int64 offsets.  Dedup keys are whole byte strings (UTF-8 is
// injective on Python strings, so this equals the reference's dict-key equality); ids are ranks in
// first-occurrence order, bit-exact with `list(dict.fromkeys(edge_texts))`.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "ghf_b200.h"

namespace ghf {
namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

__device__ __forceinline__ uint64_t hash_bytes(const uint8_t* __restrict__ p, int64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;  // FNV-1a, then a 64-bit finaliser
  for (int64_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}

__device__ __forceinline__ bool same_string(const uint8_t* __restrict__ data,
                                            const int64_t* __restrict__ off, int64_t a, int64_t b) {
  const int64_t a0 = off[a], b0 = off[b];
  const int64_t n = off[a + 1] - a0;
  if (off[b + 1] - b0 != n) return false;
  for (int64_t i = 0; i < n; ++i)
    if (data[a0 + i] != data[b0 + i]) return false;
  return true;
}

// Open-addressing table of representative edge ids.  Equal strings meet in one slot; the slot keeps
// the smallest edge id that ever arrived (atomicMin), i.e. the first occurrence.
__global__ void dedup_insert_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ off,
                                    int64_t E, uint32_t* __restrict__ table, uint64_t mask,
                                    uint32_t* __restrict__ slot_of) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s0 = off[e];
  uint64_t slot = hash_bytes(data + s0, off[e + 1] - s0) & mask;
  for (;;) {
    uint32_t cur = table[slot];
    if (cur == kEmpty) {
      cur = atomicCAS(&table[slot], kEmpty, (uint32_t)e);
      if (cur == kEmpty) break;  // claimed
    }
    if (same_string(data, off, e, cur)) {
      atomicMin(&table[slot], (uint32_t)e);
      break;
    }
    slot = (slot + 1) & mask;
  }
  slot_of[e] = (uint32_t)slot;
}

__global__ void dedup_flag_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of,
                                  int64_t E, int32_t* __restrict__ flag) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e < E) flag[e] = table[slot_of[e]] == (uint32_t)e;
}

__global__ void dedup_assign_kernel(const uint32_t* __restrict__ table, const uint32_t* __restrict__ slot_of,
                                    const int32_t* __restrict__ rank, int64_t E,
                                    int32_t* __restrict__ rel_ids, int64_t* __restrict__ first_edge,
                                    int64_t* __restrict__ num_unique) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  const uint32_t rep = table[slot_of[e]];
  rel_ids[e] = rank[rep];
  if (rep == (uint32_t)e) {
    if (first_edge) first_edge[rank[e]] = e;
  }
  if (e == E - 1) *num_unique = rank[e] + (rep == (uint32_t)e ? 1 : 0);
}

// One block per unique string: mean-pool character embeddings, project, tanh.
__global__ void text_encode_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ off,
                                   const int64_t* __restrict__ index, const float* __restrict__ emb, int C, const float* __restrict__ Wp,
                                   const float* __restrict__ bp, int T, float* __restrict__ out) {
  extern __shared__ float pooled[];  // [C]
  const int64_t u = blockIdx.x;
  const int64_t str = index ? index[u] : u;
  const int64_t s0 = off[str], s1 = off[str + 1];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sum = 0.f;
    int len = 0;
    for (int64_t i = s0; i < s1; ++i) {
      const uint32_t b = data[i];
      if ((b & 0xC0u) == 0x80u) continue;        // UTF-8 continuation byte: same code point
      const uint32_t tok = b < 128u ? b : 127u;  // min(ord(c), 127)
      sum += emb[tok * C + c];
      ++len;
    }
    if (len == 0) { sum = emb[c]; len = 1; }     // "" -> [0]
    pooled[c] = sum / (float)len;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(pooled[c], Wp[(int64_t)t * C + c], acc);
    out[u * T + t] = tanhf(acc + bp[t]);
  }
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" int ghf_dedup_texts(const uint8_t* d_utf8, const int64_t* d_offsets, int64_t E,
                               int32_t* d_rel_ids, int64_t* d_first_edge, int64_t* h_num_unique,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(E >= 0 && E < (int64_t)0xFFFFFFFE, "ghf_dedup_texts: E=%lld out of range", (long long)E);
  if (E == 0) {
    if (h_num_unique) *h_num_unique = 0;
    return 0;
  }
  uint64_t cap = 64;
  while (cap < (uint64_t)E * 2) cap <<= 1;
  TempBuf table, slot_of, flag, rank, scan_tmp, count;
  GHF_CUDA(table.alloc(cap * sizeof(uint32_t), stream));
  GHF_CUDA(slot_of.alloc(E * sizeof(uint32_t), stream));
  GHF_CUDA(flag.alloc(E * sizeof(int32_t), stream));
  GHF_CUDA(rank.alloc(E * sizeof(int32_t), stream));
  GHF_CUDA(count.alloc(sizeof(int64_t), stream));
  GHF_CUDA(cudaMemsetAsync(table.p, 0xFF, cap * sizeof(uint32_t), stream));
  const int threads = 256;
  const unsigned blocks = (unsigned)cdiv(E, threads);
  dedup_insert_kernel<<<blocks, threads, 0, stream>>>(d_utf8, d_offsets, E, table.as<uint32_t>(),
                                                      cap - 1, slot_of.as<uint32_t>());
  GHF_LAUNCH_CHECK();
  dedup_flag_kernel<<<blocks, threads, 0, stream>>>(table.as<uint32_t>(), slot_of.as<uint32_t>(), E,
                                                    flag.as<int32_t>());
  GHF_LAUNCH_CHECK();
  size_t tmp_bytes = 0;
  GHF_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, flag.as<int32_t>(), rank.as<int32_t>(),
                                         (int)E, stream));
  GHF_CUDA(scan_tmp.alloc(tmp_bytes, stream));
  GHF_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, flag.as<int32_t>(), rank.as<int32_t>(),
                                         (int)E, stream));
  g_launches.fetch_add(2, std::memory_order_relaxed);
  dedup_assign_kernel<<<blocks, threads, 0, stream>>>(table.as<uint32_t>(), slot_of.as<uint32_t>(),
                                                      rank.as<int32_t>(), E, d_rel_ids, d_first_edge,
                                                      count.as<int64_t>());
  GHF_LAUNCH_CHECK();
  if (h_num_unique) {
    GHF_CUDA(cudaMemcpyAsync(h_num_unique, count.p, sizeof(int64_t), cudaMemcpyDeviceToHost, stream));
    GHF_CUDA(cudaStreamSynchronize(stream));
  }
  return 0;
}

extern "C" int ghf_text_encode(const uint8_t* d_utf8, const int64_t* d_offsets, const int64_t* d_string_index,
                               int64_t U, const float* d_emb, int C, const float* d_Wp, const float* d_bp, int T,
                               float* d_out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(C > 0 && T > 0 && U >= 0, "ghf_text_encode: bad dims U=%lld C=%d T=%d", (long long)U, C, T);
  if (U == 0) return 0;
  text_encode_kernel<<<(unsigned)U, 128, C * sizeof(float), stream>>>(d_utf8, d_offsets, d_string_index, d_emb, C,
                                                                     d_Wp, d_bp, T, d_out);
  GHF_LAUNCH_CHECK();
  return 0;
}
