// Device-side bit reader and entropy decoder (ANS / prefix / hybrid-uint / LZ77) shared by the
// Modular and HF-coefficient stream kernels. Table layouts are the ones built by
// host/entropy.cc; decoding semantics follow crates/jxl-coding/src/{lib.rs,ans.rs,prefix.rs}.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace jxlb {

enum DevStatus : int {
  kDevOk = 0,
  kDevBadStream = 1,   // ANS final state mismatch / invalid symbol
  kDevOverrun = 2,     // read past the end of the section
  kDevInvalid = 3,     // semantic validation failed (e.g. non_zeros too large)
};

struct DevEntropyCode {
  const uint8_t* cluster_map;
  const uint32_t* configs;      // packed HybridUintConfig per cluster
  const uint64_t* ans;          // num_clusters << log_alphabet_size buckets
  const uint32_t* prefix;       // concatenated LUTs
  const uint32_t* prefix_meta;  // per cluster: table_offset, root_bits
  uint32_t log_alphabet_size;
  uint32_t use_prefix;
  uint32_t num_clusters, cluster_map_size, prefix_table_size;
  uint32_t lz77_enabled, lz77_min_symbol, lz77_min_length, lz_len_conf, lz_dist_cluster;
};

// LSB-first bit reader over the (zero-padded) codestream in global memory. 64-bit buffer fed in
// aligned 32-bit words; the NEXT word is always already in flight (`ahead`), so the global-load
// latency of a refill is hidden behind the symbols decoded from the current buffer.
struct DevBitReader {
  const uint32_t* next_word;  // word after `ahead`
  uint64_t pos;               // absolute bit offset of the next unread bit
  uint64_t buf;
  uint32_t ahead;
  int nbits;

  __device__ __forceinline__ void init(const uint8_t* d, uint64_t bit_pos) {
    pos = bit_pos;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(d) + (bit_pos >> 5);
    const uint32_t skip = uint32_t(bit_pos & 31);
    buf = uint64_t(__ldg(w)) >> skip;
    nbits = 32 - int(skip);
    buf |= uint64_t(__ldg(w + 1)) << nbits;
    nbits += 32;
    ahead = __ldg(w + 2);
    next_word = w + 3;
  }
  __device__ __forceinline__ void refill() {  // requires nbits <= 32; afterwards nbits > 32
    buf |= uint64_t(ahead) << nbits;
    nbits += 32;
    ahead = __ldg(next_word);
    ++next_word;
  }
  __device__ __forceinline__ uint32_t peek(uint32_t n) {  // n <= 32
    if (nbits < 32) refill();
    return uint32_t(buf) & (n >= 32 ? 0xffffffffu : ((1u << n) - 1));
  }
  __device__ __forceinline__ void consume(uint32_t n) {
    buf >>= n;
    nbits -= int(n);
    pos += n;
  }
  __device__ __forceinline__ uint32_t read(uint32_t n) {
    uint32_t v = peek(n);
    consume(n);
    return v;
  }
};

struct DevEntropyState {
  uint32_t ans_state;
  // LZ77
  uint32_t* window;
  uint32_t num_to_copy, copy_pos, num_decoded;
};

__device__ __forceinline__ void entropy_begin(const DevEntropyCode& c, DevEntropyState& s, DevBitReader& br,
                                              uint32_t* window) {
  s.window = window;
  s.num_to_copy = s.copy_pos = s.num_decoded = 0;
  s.ans_state = c.use_prefix ? 0x130000u : br.read(32);
}

__device__ __forceinline__ bool entropy_final_ok(const DevEntropyCode& c, const DevEntropyState& s) {
  return c.use_prefix || s.ans_state == 0x130000u;
}

__device__ __forceinline__ uint32_t entropy_read_symbol(const DevEntropyCode& c, DevEntropyState& s, DevBitReader& br,
                                                        uint32_t cluster) {
  if (c.use_prefix) {
    uint32_t off = c.prefix_meta[cluster * 2], root_bits = c.prefix_meta[cluster * 2 + 1];
    uint32_t peeked = br.peek(15);
    uint32_t e = __ldg(c.prefix + off + (peeked & ((1u << root_bits) - 1)));
    if (e & 0x80000000u) {
      uint32_t sb = (e >> 16) & 0xff;
      e = __ldg(c.prefix + off + (1u << root_bits) + (e & 0xffff) + ((peeked >> root_bits) & ((1u << sb) - 1)));
    }
    br.consume((e >> 16) & 0xff);
    return e & 0xffff;
  }
  const uint32_t log_bucket = 12 - c.log_alphabet_size;
  uint32_t state = s.ans_state;
  uint32_t idx = state & 0xfff;
  uint32_t i = idx >> log_bucket;
  uint32_t pos = idx & ((1u << log_bucket) - 1);
  uint64_t b = __ldg(c.ans + (size_t(cluster) << c.log_alphabet_size) + i);
  uint32_t alias_symbol = uint32_t(b) & 0xff;
  uint32_t alias_cutoff = (uint32_t(b) >> 8) & 0xff;
  uint32_t dist = uint32_t(b) >> 16;
  bool map_to_alias = pos >= alias_cutoff;
  uint32_t hi = map_to_alias ? uint32_t(b >> 32) : 0u;
  uint32_t offset = (hi & 0xffff) + pos;
  dist ^= hi >> 16;
  uint32_t symbol = map_to_alias ? alias_symbol : i;
  uint32_t next = (state >> 12) * dist + offset;
  if (next < (1u << 16)) next = (next << 16) | br.read(16);
  s.ans_state = next;
  return symbol;
}

__device__ __forceinline__ uint32_t entropy_read_uint(DevBitReader& br, uint32_t cfg, uint32_t token) {
  uint32_t split_exponent = cfg & 0xff, msb = (cfg >> 8) & 0xff, lsb = (cfg >> 16) & 0xff;
  uint32_t split = 1u << split_exponent;
  if (token < split) return token;
  uint32_t in_token = msb + lsb;
  uint32_t n = (split_exponent - in_token + ((token - split) >> in_token)) & 31;
  uint32_t rest = br.read(n);
  uint32_t low = token & ((1u << lsb) - 1);
  uint32_t t = (token >> lsb) & ((1u << msb) - 1);
  t |= 1u << msb;
  return uint32_t((((uint64_t(t) << n) | rest) << lsb) | low);
}

__device__ __constant__ const int8_t kDevSpecialDistances[120][2] = {
    {0, 1},  {1, 0},  {1, 1},  {-1, 1}, {0, 2},  {2, 0},  {1, 2},  {-1, 2}, {2, 1},  {-2, 1},
    {2, 2},  {-2, 2}, {0, 3},  {3, 0},  {1, 3},  {-1, 3}, {3, 1},  {-3, 1}, {2, 3},  {-2, 3},
    {3, 2},  {-3, 2}, {0, 4},  {4, 0},  {1, 4},  {-1, 4}, {4, 1},  {-4, 1}, {3, 3},  {-3, 3},
    {2, 4},  {-2, 4}, {4, 2},  {-4, 2}, {0, 5},  {3, 4},  {-3, 4}, {4, 3},  {-4, 3}, {5, 0},
    {1, 5},  {-1, 5}, {5, 1},  {-5, 1}, {2, 5},  {-2, 5}, {5, 2},  {-5, 2}, {4, 4},  {-4, 4},
    {3, 5},  {-3, 5}, {5, 3},  {-5, 3}, {0, 6},  {6, 0},  {1, 6},  {-1, 6}, {6, 1},  {-6, 1},
    {2, 6},  {-2, 6}, {6, 2},  {-6, 2}, {4, 5},  {-4, 5}, {5, 4},  {-5, 4}, {3, 6},  {-3, 6},
    {6, 3},  {-6, 3}, {0, 7},  {7, 0},  {1, 7},  {-1, 7}, {5, 5},  {-5, 5}, {7, 1},  {-7, 1},
    {4, 6},  {-4, 6}, {6, 4},  {-6, 4}, {2, 7},  {-2, 7}, {7, 2},  {-7, 2}, {3, 7},  {-3, 7},
    {7, 3},  {-7, 3}, {5, 6},  {-5, 6}, {6, 5},  {-6, 5}, {8, 0},  {4, 7},  {-4, 7}, {7, 4},
    {-7, 4}, {8, 1},  {8, 2},  {6, 6},  {-6, 6}, {8, 3},  {5, 7},  {-5, 7}, {7, 5},  {-7, 5},
    {8, 4},  {6, 7},  {-6, 7}, {7, 6},  {-7, 6}, {8, 5},  {7, 7},  {-7, 7}, {8, 6},  {8, 7},
};

// read_varint_with_multiplier_clustered incl. LZ77 (lib.rs:476-569). `err` is set on a repeat
// before any symbol.
__device__ __forceinline__ uint32_t entropy_read_varint(const DevEntropyCode& c, DevEntropyState& s, DevBitReader& br,
                                                        uint32_t cluster, uint32_t dist_multiplier, int& err) {
  if (!c.lz77_enabled) {
    uint32_t token = entropy_read_symbol(c, s, br, cluster);
    return entropy_read_uint(br, __ldg(c.configs + cluster), token);
  }
  uint32_t r;
  if (s.num_to_copy > 0) {
    r = s.window[s.copy_pos & 0xfffff];
    ++s.copy_pos;
    --s.num_to_copy;
  } else {
    uint32_t token = entropy_read_symbol(c, s, br, cluster);
    if (token >= c.lz77_min_symbol) {
      if (s.num_decoded == 0) {
        err = kDevBadStream;
        return 0;
      }
      uint32_t n = entropy_read_uint(br, c.lz_len_conf, token - c.lz77_min_symbol);
      s.num_to_copy = n + c.lz77_min_length;
      uint32_t dtoken = entropy_read_symbol(c, s, br, c.lz_dist_cluster);
      uint32_t distance = entropy_read_uint(br, __ldg(c.configs + c.lz_dist_cluster), dtoken);
      if (dist_multiplier == 0) {
      } else if (distance < 120) {
        int32_t dd = int32_t(kDevSpecialDistances[distance][0]) + int32_t(dist_multiplier) * int32_t(kDevSpecialDistances[distance][1]);
        distance = uint32_t(max(dd - 1, 0));
      } else {
        distance -= 120;
      }
      distance = min(min((1u << 20) - 1, distance) + 1, s.num_decoded);
      s.copy_pos = s.num_decoded - distance;
      r = s.window[s.copy_pos & 0xfffff];
      ++s.copy_pos;
      --s.num_to_copy;
    } else {
      r = entropy_read_uint(br, __ldg(c.configs + cluster), token);
    }
  }
  s.window[s.num_decoded & 0xfffff] = r;
  ++s.num_decoded;
  return r;
}

__device__ __forceinline__ int32_t dev_unpack_signed(uint32_t x) { return int32_t((x >> 1) ^ (0u - (x & 1))); }

}  // namespace jxlb
