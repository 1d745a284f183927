// text.cu — relation-string dedup (HG:264-268) and the character-bag text encoder (HG:66-81).
//
// Strings arrive packed: UTF-8 bytes + int64 offsets.  Dedup keys are whole byte strings (UTF-8 is
// injective on Python strings, so this equals the reference's dict-key equality); ids are ranks in
// first-occurrence order, bit-exact with `list(dict.fromkeys(edge_texts))`.
#include <cub/device/device_radix_sort.cuh>

#include <functional>

#include "common.cuh"
#include "ghf_b200.h"
#include "graph.cuh"

namespace ghf {
namespace {

constexpr uint32_t kEmpty = 0xFFFFFFFFu;

// Bytes [p, p + 4) of the packed strings as one little-endian word, from aligned 4-byte loads (the read-only
// path caches them: neighbouring threads hash neighbouring strings).  The caller guarantees p + 4 <= total bytes;
// `limit` = total rounded down to 4: words that would reach past it are assembled from byte loads.
__device__ __forceinline__ uint32_t load_word(const uint8_t* __restrict__ data, int64_t p, int64_t limit) {
  const int64_t a = p & ~(int64_t)3;
  if (a + 8 <= limit) {
    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(data + a));
    const uint32_t hi = __ldg(reinterpret_cast<const uint32_t*>(data + a + 4));
    return __funnelshift_r(lo, hi, 8 * (int)(p & 3));
  }
  uint32_t w = 0;
  for (int i = 0; i < 4; ++i) w |= (uint32_t)data[p + i] << (8 * i);
  return w;
}
// word i of string [s0, s0 + n): bytes beyond the string are zeroed
__device__ __forceinline__ uint32_t string_word(const uint8_t* __restrict__ data, int64_t s0, int64_t n, int64_t i,
                                                int64_t limit, int64_t total) {
  const int64_t p = s0 + 4 * i;
  uint32_t w;
  if (p + 4 <= total) {
    w = load_word(data, p, limit);
  } else {  // the last word of the buffer: byte loads only
    w = 0;
    for (int k = 0; k < 4; ++k)
      if (p + k < total) w |= (uint32_t)data[p + k] << (8 * k);
  }
  const int64_t rem = n - 4 * i;
  return rem >= 4 ? w : (w & ((1u << (8 * (int)rem)) - 1u));
}

__device__ __forceinline__ uint64_t hash_string(const uint8_t* __restrict__ data, int64_t s0, int64_t n,
                                                int64_t limit, int64_t total) {
  uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)n;  // word-wise FNV-style mix, then a 64-bit finaliser
  for (int64_t i = 0; 4 * i < n; ++i) h = (h ^ string_word(data, s0, n, i, limit, total)) * 0x100000001b3ull;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}

__device__ __forceinline__ bool same_string(const uint8_t* __restrict__ data,
                                            const int64_t* __restrict__ off, int64_t a, int64_t b, int64_t limit,
                                            int64_t total) {
  const int64_t a0 = off[a], b0 = off[b];
  const int64_t n = off[a + 1] - a0;
  if (off[b + 1] - b0 != n) return false;
  for (int64_t i = 0; 4 * i < n; ++i)
    if (string_word(data, a0, n, i, limit, total) != string_word(data, b0, n, i, limit, total)) return false;
  return true;
}

// Strings of at most 16 bytes (relation names usually are) are held in four registers: five aligned word loads cover
// [s0, s0 + 16) at any alignment, bytes beyond the string are zeroed.  -> false when the fast path does not apply.
__device__ __forceinline__ bool load_short(const uint8_t* __restrict__ data, int64_t s0, int64_t n, int64_t limit,
                                           uint32_t (&w)[4]) {
  const int64_t a = s0 & ~(int64_t)3;
  if (n > 16 || a + 20 > limit) return false;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(data + a);
  const uint32_t x0 = __ldg(p), x1 = __ldg(p + 1), x2 = __ldg(p + 2), x3 = __ldg(p + 3), x4 = __ldg(p + 4);
  const int sh = 8 * (int)(s0 & 3);
  w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh);
  w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rem = (int)n - 4 * i;
    w[i] = rem >= 4 ? w[i] : (rem <= 0 ? 0u : (w[i] & ((1u << (8 * rem)) - 1u)));
  }
  return true;
}
__device__ __forceinline__ uint64_t hash_short(const uint32_t (&w)[4], int64_t n) {   // == hash_string on the same bytes
  uint64_t h = 0xcbf29ce484222325ull ^ (uint64_t)n;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * i < n) h = (h ^ w[i]) * 0x100000001b3ull;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
  h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
  h ^= h >> 33;
  return h;
}

// Open-addressing table of representative edge ids.  Equal strings meet in one slot; the slot keeps
// the smallest edge id that ever arrived (atomicMin), i.e. the first occurrence.  A thread that already sees
// a smaller id in the slot skips the atomic: with few distinct strings almost every edge does.
// Work item j is string `subset[j]` (ascending ids) or string j itself; the table holds work-item indices, so
// "smallest index" is "first occurrence" either way.
// The table is SMALL by default (2^20 slots, 4 MiB: it lives in L2 and needs no 100 MB clear): a thread that probes
// more than `max_probes` slots raises `overflow` and the host retries with a table of 2 n slots.
__global__ void dedup_insert_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ off,
                                    int64_t E, const uint32_t* __restrict__ subset, int64_t n,
                                    uint32_t* __restrict__ table, uint64_t mask, uint32_t* __restrict__ slot_of,
                                    uint32_t max_probes, int32_t* __restrict__ overflow) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t e = subset ? subset[j] : j;
  const int64_t total = off[E], limit = total & ~(int64_t)3;
  const int64_t s0 = off[e], len = off[e + 1] - s0;
  uint32_t mine[4];
  const bool is_short = load_short(data, s0, len, limit, mine);
  uint64_t slot = (is_short ? hash_short(mine, len) : hash_string(data, s0, len, limit, total)) & mask;
  for (uint32_t probes = 0;; ++probes) {
    if (probes > max_probes) {
      *overflow = 1;
      break;
    }
    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&table[slot]);
    if (cur == kEmpty) {
      cur = atomicCAS(&table[slot], kEmpty, (uint32_t)j);
      if (cur == kEmpty) break;  // claimed
    }
    if (cur == (uint32_t)j) break;
    const int64_t other = subset ? subset[cur] : cur;
    bool same;
    uint32_t theirs[4];
    const int64_t o0 = off[other], olen = off[other + 1] - o0;
    if (is_short && olen == len && load_short(data, o0, olen, limit, theirs))
      same = mine[0] == theirs[0] && mine[1] == theirs[1] && mine[2] == theirs[2] && mine[3] == theirs[3];
    else
      same = same_string(data, off, e, other, limit, total);
    if (same) {
      if (cur > (uint32_t)j) atomicMin(&table[slot], (uint32_t)j);   // entries only ever decrease
      break;
    }
    slot = (slot + 1) & mask;
  }
  slot_of[j] = (uint32_t)slot;
}

// Occupied slots -> (representative work item, slot) pairs, in any order; *count = number of distinct strings.
__global__ void dedup_compact_kernel(const uint32_t* __restrict__ table, uint64_t cap, uint32_t* __restrict__ reps,
                                     uint32_t* __restrict__ slots, int32_t* __restrict__ count, int64_t max_out) {
  const uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const uint32_t rep = table[s];
  if (rep == kEmpty) return;
  const int i = atomicAdd(count, 1);
  if (i < max_out) {
    reps[i] = rep;
    slots[i] = (uint32_t)s;
  }
}

// Pairs sorted by representative = first-occurrence order: rank i goes to slot slots[i], first_edge[i] = that string.
__global__ void dedup_rank_kernel(const uint32_t* __restrict__ reps, const uint32_t* __restrict__ slots, int64_t U,
                                  const uint32_t* __restrict__ subset, int32_t* __restrict__ rank_of_slot,
                                  int64_t* __restrict__ first_edge) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= U) return;
  rank_of_slot[slots[i]] = (int32_t)i;
  if (first_edge) first_edge[i] = subset ? (int64_t)subset[reps[i]] : (int64_t)reps[i];
}

__global__ void dedup_assign_kernel(const uint32_t* __restrict__ slot_of, const int32_t* __restrict__ rank_of_slot,
                                    int64_t n, int32_t* __restrict__ rel_ids) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j < n) rel_ids[j] = rank_of_slot[slot_of[j]];
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

// Backward of text_encode_kernel: one CTA per string.  out = tanh(pooled Wp^T + bp), pooled = mean of the code
// points' embedding rows; the three parameter gradients are accumulated with atomics (U strings, a few thousand
// addresses - this stage is microseconds next to the layers).
__global__ void text_encode_bwd_kernel(const uint8_t* __restrict__ data, const int64_t* __restrict__ off,
                                       const int64_t* __restrict__ index, const float* __restrict__ emb, int C,
                                       const float* __restrict__ Wp, int T, const float* __restrict__ out,
                                       const float* __restrict__ g_out, float* __restrict__ g_emb,
                                       float* __restrict__ g_Wp, float* __restrict__ g_bp) {
  extern __shared__ float sh[];  // pooled[C] | gmean[C] | gpre[T]
  float* pooled = sh;
  float* gmean = sh + C;
  float* gpre = sh + 2 * C;
  __shared__ int s_len;
  const int64_t u = blockIdx.x;
  const int64_t str = index ? index[u] : u;
  const int64_t s0 = off[str], s1 = off[str + 1];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sum = 0.f;
    int len = 0;
    for (int64_t i = s0; i < s1; ++i) {
      const uint32_t b = data[i];
      if ((b & 0xC0u) == 0x80u) continue;
      sum += emb[(b < 128u ? b : 127u) * C + c];
      ++len;
    }
    if (len == 0) { sum = emb[c]; len = 1; }
    pooled[c] = sum / (float)len;
    if (c == 0) s_len = len;
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float y = out[u * T + t];
    const float gp = g_out[u * T + t] * (1.f - y * y);
    gpre[t] = gp;
    atomicAdd(g_bp + t, gp);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * C; idx += blockDim.x)
    atomicAdd(g_Wp + idx, gpre[idx / C] * pooled[idx % C]);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float gm = 0.f;
    for (int t = 0; t < T; ++t) gm = fmaf(gpre[t], Wp[(int64_t)t * C + c], gm);
    gmean[c] = gm / (float)s_len;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int seen = 0;
    for (int64_t i = s0; i < s1; ++i) {
      const uint32_t b = data[i];
      if ((b & 0xC0u) == 0x80u) continue;
      atomicAdd(g_emb + (b < 128u ? b : 127u) * C + c, gmean[c]);
      ++seen;
    }
    if (seen == 0) atomicAdd(g_emb + c, gmean[c]);  // "" -> [0]
  }
}

}  // namespace
}  // namespace ghf

using namespace ghf;

extern "C" int ghf_dedup_texts(const uint8_t* d_utf8, const int64_t* d_offsets, int64_t E,
                               const uint32_t* d_subset, int64_t n_subset, int32_t* d_rel_ids,
                               int64_t* d_first_edge, int64_t* h_num_unique, void* stream_) {
  return ghf::dedup_texts_hooked(d_utf8, d_offsets, E, d_subset, n_subset, d_rel_ids, d_first_edge, h_num_unique,
                                 (cudaStream_t)stream_, nullptr);
}

// `before_sync` (optional) is called once, after the insert + compaction kernels are enqueued and before the host
// waits for the number of distinct strings: work enqueued there (on another stream) fills the GPU during the round trip.
int ghf::dedup_texts_hooked(const uint8_t* d_utf8, const int64_t* d_offsets, int64_t E, const uint32_t* d_subset,
                            int64_t n_subset, int32_t* d_rel_ids, int64_t* d_first_edge, int64_t* h_num_unique,
                            cudaStream_t stream, const std::function<int()>* before_sync) {
  GHF_REQUIRE(E >= 0 && E <= (int64_t)0x7FFFFFFF, "ghf_dedup_texts: E=%lld out of range", (long long)E);
  GHF_REQUIRE(reinterpret_cast<uintptr_t>(d_utf8) % 4 == 0, "ghf_dedup_texts: d_utf8 must be 4-byte aligned");
  GHF_REQUIRE(d_subset == nullptr || (n_subset >= 0 && n_subset <= E), "ghf_dedup_texts: bad subset size");
  const int64_t n = d_subset ? n_subset : E;   // work items: the strings of the subset, or all of them
  if (n == 0) {
    if (h_num_unique) *h_num_unique = 0;
    return 0;
  }
  // Knowledge graphs have few distinct relation strings (BASELINE: 7 ... 20k), so the hash table starts small enough
  // to stay in L2; a graph with more than 2^19 distinct strings (e.g. the reference-shaped _message_passing entry,
  // where every edge is its own relation) is detected and redone with the worst-case table.
  uint64_t full = 64;
  while (full < (uint64_t)n * 2) full <<= 1;
  constexpr uint64_t kSmall = 1ull << 20;
  TempBuf slot_of, words;
  GHF_CUDA(slot_of.alloc(n * sizeof(uint32_t), stream));
  GHF_CUDA(words.alloc(2 * sizeof(int32_t), stream));          // [0] distinct strings, [1] overflow
  const int threads = 256;
  const unsigned blocks = (unsigned)cdiv(n, threads);
  for (uint64_t cap = full < kSmall ? full : kSmall;; cap = full) {
    const int64_t max_u = (int64_t)(cap / 2 < (uint64_t)n ? cap / 2 : (uint64_t)n);   // load factor <= 1/2
    TempBuf table, reps, slots, rank_of_slot;
    GHF_CUDA(table.alloc(cap * sizeof(uint32_t), stream));
    GHF_CUDA(reps.alloc(2 * max_u * sizeof(uint32_t), stream));    // double buffers of the pair sort
    GHF_CUDA(slots.alloc(2 * max_u * sizeof(uint32_t), stream));
    GHF_CUDA(rank_of_slot.alloc(cap * sizeof(int32_t), stream));
    GHF_CUDA(cudaMemsetAsync(table.p, 0xFF, cap * sizeof(uint32_t), stream));
    GHF_CUDA(cudaMemsetAsync(words.p, 0, 2 * sizeof(int32_t), stream));
    dedup_insert_kernel<<<blocks, threads, 0, stream>>>(d_utf8, d_offsets, E, d_subset, n, table.as<uint32_t>(),
                                                        cap - 1, slot_of.as<uint32_t>(),
                                                        cap == full ? 0xFFFFFFFFu : 4096u, words.as<int32_t>() + 1);
    GHF_LAUNCH_CHECK();
    dedup_compact_kernel<<<(unsigned)cdiv((int64_t)cap, threads), threads, 0, stream>>>(
        table.as<uint32_t>(), cap, reps.as<uint32_t>(), slots.as<uint32_t>(), words.as<int32_t>(), max_u);
    GHF_LAUNCH_CHECK();
    int32_t h_words[2] = {0, 0};
    GHF_CUDA(cudaMemcpyAsync(h_words, words.p, sizeof(h_words), cudaMemcpyDeviceToHost, stream));
    // wait for the two words only (an event after the copy), with the hooks run in between: `before_sync` and the
    // caller's one-shot hook (ghf_set_presync_hook) - both once, not again on the retry with the full table
    if (int rc = readback_wait(stream, before_sync)) return rc;
    before_sync = nullptr;
    if ((h_words[1] != 0 || h_words[0] > max_u) && cap != full) continue;   // too many distinct strings: full table
    const int64_t U = h_words[0];
    cub::DoubleBuffer<uint32_t> kbuf(reps.as<uint32_t>(), reps.as<uint32_t>() + max_u);
    cub::DoubleBuffer<uint32_t> vbuf(slots.as<uint32_t>(), slots.as<uint32_t>() + max_u);
    size_t tmp_bytes = 0;
    TempBuf tmp;
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kbuf, vbuf, (int)U, 0, 32, stream));
    GHF_CUDA(tmp.alloc(tmp_bytes, stream));
    GHF_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, kbuf, vbuf, (int)U, 0, 32, stream));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    dedup_rank_kernel<<<(unsigned)cdiv(U, threads), threads, 0, stream>>>(kbuf.Current(), vbuf.Current(), U, d_subset,
                                                                         rank_of_slot.as<int32_t>(), d_first_edge);
    GHF_LAUNCH_CHECK();
    dedup_assign_kernel<<<blocks, threads, 0, stream>>>(slot_of.as<uint32_t>(), rank_of_slot.as<int32_t>(), n,
                                                        d_rel_ids);
    GHF_LAUNCH_CHECK();
    if (h_num_unique) *h_num_unique = U;
    return 0;
  }
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

extern "C" int ghf_text_encode_backward(const uint8_t* d_utf8, const int64_t* d_offsets, const int64_t* d_string_index,
                                        int64_t U, const float* d_emb, int C, const float* d_Wp, int T,
                                        const float* d_out, const float* d_g_out, float* d_g_emb, float* d_g_Wp,
                                        float* d_g_bp, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  GHF_REQUIRE(C > 0 && T > 0 && U >= 0, "ghf_text_encode_backward: bad dims U=%lld C=%d T=%d", (long long)U, C, T);
  GHF_REQUIRE(d_g_emb && d_g_Wp && d_g_bp, "ghf_text_encode_backward: NULL gradient buffer");
  GHF_CUDA(cudaMemsetAsync(d_g_emb, 0, (size_t)128 * C * sizeof(float), stream));
  GHF_CUDA(cudaMemsetAsync(d_g_Wp, 0, (size_t)T * C * sizeof(float), stream));
  GHF_CUDA(cudaMemsetAsync(d_g_bp, 0, (size_t)T * sizeof(float), stream));
  if (U == 0) return 0;
  text_encode_bwd_kernel<<<(unsigned)U, 128, (2 * C + T) * sizeof(float), stream>>>(
      d_utf8, d_offsets, d_string_index, d_emb, C, d_Wp, T, d_out, d_g_out, d_g_emb, d_g_Wp, d_g_bp);
  GHF_LAUNCH_CHECK();
  return 0;
}
