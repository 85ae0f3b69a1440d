// Helpers shared by the two entropy-stream kernels (modular_stream.cu, entropy.cu): wrapping
// integer arithmetic, the shared-memory view of an entropy code and the symbol / hybrid-uint
// readers (crates/jxl-coding/src/{lib.rs:572-605, ans.rs:276-330, prefix.rs:335-357}).
#pragma once
#include "common.cuh"

namespace jxlb {
namespace {

__device__ __forceinline__ int32_t wadd(int32_t a, int32_t b) { return int32_t(uint32_t(a) + uint32_t(b)); }
__device__ __forceinline__ int32_t wsub(int32_t a, int32_t b) { return int32_t(uint32_t(a) - uint32_t(b)); }
__device__ __forceinline__ int32_t wmul(int32_t a, int32_t b) { return int32_t(uint32_t(a) * uint32_t(b)); }
__device__ __forceinline__ uint32_t abs_diff(int32_t a, int32_t b) {
  return a > b ? uint32_t(a) - uint32_t(b) : uint32_t(b) - uint32_t(a);
}
__device__ __forceinline__ int32_t grad_clamped(int32_t n, int32_t w, int32_t nw) {
  int32_t hi = max(n, w), lo = min(n, w);
  int64_t v = int64_t(lo) + int64_t(hi) - int64_t(nw);
  return int32_t(v < lo ? int64_t(lo) : (v > hi ? int64_t(hi) : v));
}
__device__ __forceinline__ uint32_t ilog2_u32(uint32_t v) { return 31u - uint32_t(__clz(int(v))); }
__device__ __forceinline__ int64_t abs64(int64_t v) { return v < 0 ? -v : v; }

// Shared-memory view of an entropy code; pointers may alias global memory when a table is too
// large to stage.
struct CodeView {
  const uint32_t* configs;
  const uint64_t* ans;
  const uint32_t* prefix;
  const uint32_t* prefix_meta;
  uint32_t log_alphabet_size, use_prefix;
};

// ANS symbol (ans.rs:276-330): alias-table lookup, state update, 16-bit refill.
template <typename BR>
__device__ __forceinline__ uint32_t cv_read_symbol_ans(const CodeView& c, uint32_t& ans_state, BR& br, uint32_t cluster) {
  const uint32_t log_bucket = 12 - c.log_alphabet_size;
  uint32_t state = ans_state;
  uint32_t idx = state & 0xfff;
  uint32_t i = idx >> log_bucket;
  uint32_t pos = idx & ((1u << log_bucket) - 1);
  uint64_t b = c.ans[(size_t(cluster) << c.log_alphabet_size) + i];
  uint32_t lo = uint32_t(b), hi32 = uint32_t(b >> 32);
  uint32_t alias_symbol = lo & 0xff;
  uint32_t alias_cutoff = (lo >> 8) & 0xff;
  uint32_t dist = lo >> 16;
  bool map_to_alias = pos >= alias_cutoff;
  uint32_t hi = map_to_alias ? hi32 : 0u;
  uint32_t offset = (hi & 0xffff) + pos;
  dist ^= hi >> 16;
  uint32_t symbol = map_to_alias ? alias_symbol : i;
  uint32_t next = (state >> 12) * dist + offset;
  if (next < (1u << 16)) next = (next << 16) | br.read(16);
  ans_state = next;
  return symbol;
}

template <typename BR>
__device__ __forceinline__ uint32_t cv_read_symbol(const CodeView& c, uint32_t& ans_state, BR& br, uint32_t cluster) {
  if (c.use_prefix) {  // prefix.rs:335-357
    uint32_t off = c.prefix_meta[cluster * 2], root_bits = c.prefix_meta[cluster * 2 + 1];
    uint32_t peeked = br.peek(15);
    uint32_t e = c.prefix[off + (peeked & ((1u << root_bits) - 1))];
    if (e & 0x80000000u) {
      uint32_t sb = (e >> 16) & 0xff;
      e = c.prefix[off + (1u << root_bits) + (e & 0xffff) + ((peeked >> root_bits) & ((1u << sb) - 1))];
    }
    br.consume((e >> 16) & 0xff);
    return e & 0xffff;
  }
  return cv_read_symbol_ans(c, ans_state, br, cluster);
}

template <typename BR>
__device__ __forceinline__ uint32_t cv_read_uint(BR& br, uint32_t cfg, uint32_t token) {
  uint32_t split_exponent = cfg & 0xff;
  uint32_t split = 1u << split_exponent;
  if (token < split) return token;
  uint32_t msb = (cfg >> 8) & 0xff, lsb = (cfg >> 16) & 0xff;
  uint32_t in_token = msb + lsb;
  uint32_t n = (split_exponent - in_token + ((token - split) >> in_token)) & 31;
  uint32_t rest = br.read(n);
  uint32_t low = token & ((1u << lsb) - 1);
  uint32_t t = (token >> lsb) & ((1u << msb) - 1);
  t |= 1u << msb;
  return uint32_t((((uint64_t(t) << n) | rest) << lsb) | low);
}

// cooperative copy global -> shared by the 32 lanes of a warp (word granularity)
__device__ __forceinline__ void warp_copy_words(uint32_t* dst, const uint32_t* src, uint32_t nwords, uint32_t lane) {
  for (uint32_t i = lane; i < nwords; i += 32) dst[i] = __ldg(src + i);
}

}  // namespace
}  // namespace jxlb
