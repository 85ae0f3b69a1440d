// Optimised entropy-stream kernels: Modular channel decode and HF coefficient decode.
//
// Both are strictly serial per stream (ANS state / context chain), so the design goal is the
// shortest dependent-instruction chain per symbol:
//   * one warp per stream; all 32 lanes stage the stream's tables (MA tree or its flattened LUT,
//     ANS alias tables / prefix LUTs, hybrid-uint configs, cluster maps, WP error rows) into shared
//     memory, then lane 0 walks the chain with every dependent load hitting shared memory;
//   * neighbour samples are carried in registers and the next row positions are prefetched one
//     iteration ahead (they do not depend on the decoded value);
//   * the weighted predictor divides through the 65-entry reciprocal table of the reference.
// Integer semantics are those of crates/jxl-modular/src/{image.rs,predictor.rs,ma.rs} and
// crates/jxl-vardct/src/hf_coeff.rs (bit-exact, wrapping i32).
#include "kernels.h"

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

__device__ __forceinline__ uint32_t cv_read_symbol(const CodeView& c, uint32_t& ans_state, DevBitReader& br, uint32_t cluster) {
  if (c.use_prefix) {
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

__device__ __forceinline__ uint32_t cv_read_uint(DevBitReader& br, uint32_t cfg, uint32_t token) {
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

constexpr uint32_t kSmemTreeBytes = 32 * 1024;
constexpr uint32_t kSmemAnsBytes = 32 * 1024;
constexpr uint32_t kSmemPrefixBytes = 24 * 1024;
constexpr uint32_t kSmemLutBytes = 8 * 1024;
constexpr uint32_t kSmemWpMaxWidth = 1024;

struct SmemLayout {
  uint32_t tree, ans, prefix, prefix_meta, configs, luts, wp, div, total;
};

__host__ __device__ inline SmemLayout modular_layout(uint32_t num_nodes, const DevEntropyCode& code, uint32_t lut_total,
                                                      uint32_t use_wp, uint32_t max_w) {
  SmemLayout L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t o = off;
    off += (bytes + 15) & ~15u;
    return o;
  };
  L.div = take(65 * 4);
  uint32_t tree_bytes = num_nodes * 16;
  L.tree = (tree_bytes && tree_bytes <= kSmemTreeBytes) ? take(tree_bytes) : 0xffffffffu;
  L.configs = take(code.num_clusters * 4);
  if (code.use_prefix) {
    L.ans = 0xffffffffu;
    uint32_t pb = code.prefix_table_size * 4;
    L.prefix = (pb && pb <= kSmemPrefixBytes) ? take(pb) : 0xffffffffu;
    L.prefix_meta = take(code.num_clusters * 8);
  } else {
    uint32_t ab = (code.num_clusters << code.log_alphabet_size) * 8;
    L.ans = (ab <= kSmemAnsBytes) ? take(ab) : 0xffffffffu;
    L.prefix = 0xffffffffu;
    L.prefix_meta = 0xffffffffu;
  }
  uint32_t lb = lut_total * 2;
  L.luts = (lb && lb <= kSmemLutBytes) ? take(lb) : 0xffffffffu;
  L.wp = (use_wp && max_w <= kSmemWpMaxWidth) ? take(max_w * 5 * 4) : 0xffffffffu;
  L.total = off;
  return L;
}

// SelfCorrectingPredictor (predictor.rs:279-441) with row state in shared (or global) memory.
struct FastWp {
  uint32_t width, x, y;
  int32_t* true_err_row;
  uint32_t* sub_err_row;
  const uint32_t* div;  // DIV_LOOKUP (predictor.rs:150-160)
  uint32_t p1, p2, p3a, p3b, p3c, p3d, p3e, w0, w1, w2, w3;
  int32_t te_w, te_nw, te_n, te_ne;
  uint32_t a0, a1, a2, a3;  // subpred_err_nw_ww
  uint32_t b0, b1, b2, b3;  // subpred_err_n_w
  uint32_t c0, c1, c2, c3;  // subpred_err_ne
  int64_t prediction;
  int32_t max_error;
  int64_t s0, s1, s2, s3;

  __device__ __forceinline__ void reset(uint32_t width_, int32_t* rows, const uint32_t* hdr, const uint32_t* div_) {
    width = width_;
    x = y = 0;
    sub_err_row = reinterpret_cast<uint32_t*>(rows);  // 16-byte aligned: accessed as uint4
    true_err_row = rows + size_t(width_) * 4;
    div = div_;
    for (uint32_t i = 0; i < width_ * 5; ++i) rows[i] = 0;
    p1 = hdr[0], p2 = hdr[1], p3a = hdr[2], p3b = hdr[3], p3c = hdr[4], p3d = hdr[5], p3e = hdr[6];
    w0 = hdr[7], w1 = hdr[8], w2 = hdr[9], w3 = hdr[10];
    te_w = te_nw = te_n = te_ne = 0;
    a0 = a1 = a2 = a3 = b0 = b1 = b2 = b3 = c0 = c1 = c2 = c3 = 0;
    prediction = 0;
    max_error = 0;
  }
  __device__ __forceinline__ uint32_t weight_of(uint32_t err_sum, uint32_t maxweight) const {
    uint32_t t = uint32_t((uint64_t(err_sum) + 1) >> 5);
    uint32_t shift = t ? ilog2_u32(t) : 0;
    return 4 + ((maxweight * div[(err_sum >> shift) + 1]) >> shift);
  }
  __device__ __forceinline__ void predict(int32_t n, int32_t nw, int32_t ne, int32_t wv, int32_t nn) {
    int64_t tew = te_w, tenw = te_nw, ten = te_n, tene = te_ne;
    int64_t n3 = int64_t(n) << 3, nw3 = int64_t(nw) << 3, ne3 = int64_t(ne) << 3, w3_ = int64_t(wv) << 3,
            nn3 = int64_t(nn) << 3;
    s0 = w3_ + ne3 - n3;
    s1 = n3 - (((tew + ten + tene) * int64_t(p1)) >> 5);
    s2 = w3_ - (((tew + ten + tenw) * int64_t(p2)) >> 5);
    s3 = n3 - ((tenw * int64_t(p3a) + ten * int64_t(p3b) + tene * int64_t(p3c) + (nn3 - n3) * int64_t(p3d) +
                (nw3 - w3_) * int64_t(p3e)) >> 5);
    uint32_t g0 = weight_of(a0 + b0 + c0, w0), g1 = weight_of(a1 + b1 + c1, w1), g2 = weight_of(a2 + b2 + c2, w2),
             g3 = weight_of(a3 + b3 + c3, w3);
    uint32_t sum_weights = g0 + g1 + g2 + g3;
    uint32_t log_weight = ilog2_u32(sum_weights >> 4);
    g0 >>= log_weight, g1 >>= log_weight, g2 >>= log_weight, g3 >>= log_weight;
    sum_weights = g0 + g1 + g2 + g3;
    int64_t s = (int64_t(sum_weights) >> 1) - 1;
    s += s0 * int64_t(g0) + s1 * int64_t(g1) + s2 * int64_t(g2) + s3 * int64_t(g3);
    int64_t pred = (s * int64_t(div[sum_weights])) >> 24;
    if (((ten ^ tew) | (ten ^ tenw)) <= 0) {
      int64_t mn = min(min(n3, w3_), ne3), mx = max(max(n3, w3_), ne3);
      pred = min(max(pred, mn), mx);
    }
    int64_t me = tew;
    if (abs64(ten) > abs64(me)) me = ten;
    if (abs64(tenw) > abs64(me)) me = tenw;
    if (abs64(tene) > abs64(me)) me = tene;
    prediction = pred;
    max_error = int32_t(me);
  }
  __device__ __forceinline__ void record(int32_t sample_) {
    int64_t s8 = int64_t(sample_) << 3;
    int64_t true_err = prediction - s8;
    uint32_t e0 = uint32_t((uint64_t(abs64(s0 - s8)) + 3) >> 3), e1 = uint32_t((uint64_t(abs64(s1 - s8)) + 3) >> 3),
             e2 = uint32_t((uint64_t(abs64(s2 - s8)) + 3) >> 3), e3 = uint32_t((uint64_t(abs64(s3 - s8)) + 3) >> 3);
    true_err_row[x] = int32_t(true_err);
    *reinterpret_cast<uint4*>(sub_err_row + size_t(x) * 4) = make_uint4(e0, e1, e2, e3);
    ++x;
    if (x >= width) {
      ++y;
      x = 0;
      te_w = 0;
      te_n = true_err_row[0];
      te_nw = te_n;
      uint4 r = *reinterpret_cast<const uint4*>(sub_err_row);
      b0 = a0 = r.x, b1 = a1 = r.y, b2 = a2 = r.z, b3 = a3 = r.w;
      if (width <= 1) {
        te_ne = te_n;
        c0 = b0, c1 = b1, c2 = b2, c3 = b3;
      } else {
        te_ne = true_err_row[1];
        uint4 q = *reinterpret_cast<const uint4*>(sub_err_row + 4);
        c0 = q.x, c1 = q.y, c2 = q.z, c3 = q.w;
      }
    } else {
      te_w = int32_t(true_err);
      te_nw = te_n;
      te_n = te_ne;
      a0 = b0, a1 = b1, a2 = b2, a3 = b3;
      b0 = c0 + e0, b1 = c1 + e1, b2 = c2 + e2, b3 = c3 + e3;
      if (x + 1 >= width) {
        te_ne = te_n;
        c0 = b0, c1 = b1, c2 = b2, c3 = b3;
      } else if (y != 0) {
        te_ne = true_err_row[x + 1];
        uint4 q = *reinterpret_cast<const uint4*>(sub_err_row + size_t(x + 1) * 4);
        c0 = q.x, c1 = q.y, c2 = q.z, c3 = q.w;
      }
    }
  }
};

constexpr int kMaxPrev = 16;

__global__ void __launch_bounds__(32) modular_decode_fast_kernel(const uint8_t* __restrict__ cs,
                                                                 const DevModularJob* __restrict__ jobs,
                                                                 const DevChannel* __restrict__ channels,
                                                                 const DevChannelPlan* __restrict__ plans,
                                                                 uint64_t* __restrict__ end_bits, int* __restrict__ status,
                                                                 int num_jobs) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int job_idx = blockIdx.x;
  if (job_idx >= num_jobs) return;
  const uint32_t lane = threadIdx.x;
  const DevModularJob& job = jobs[job_idx];
  const DevEntropyCode& code = job.code;
  const DevChannel* chans = channels + job.first_channel;
  const DevChannelPlan* chplans = plans + job.first_channel;
  uint32_t max_w = 0;
  for (uint32_t ci = 0; ci < job.num_channels; ++ci) max_w = max(max_w, chans[ci].w);
  const SmemLayout L = modular_layout(job.num_tree_nodes, code, job.lut_total, job.use_wp, max_w);

  // ---- stage tables ----
  uint32_t* s_div = reinterpret_cast<uint32_t*>(smem + L.div);
  for (uint32_t i = lane; i < 65; i += 32) s_div[i] = i ? (1u << 24) / i : 0;
  const MaNode* tree = job.tree;
  if (L.tree != 0xffffffffu) {
    warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.tree), reinterpret_cast<const uint32_t*>(job.tree),
                    job.num_tree_nodes * 4, lane);
    tree = reinterpret_cast<const MaNode*>(smem + L.tree);
  }
  CodeView cv;
  cv.log_alphabet_size = code.log_alphabet_size;
  cv.use_prefix = code.use_prefix;
  warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.configs), code.configs, code.num_clusters, lane);
  cv.configs = reinterpret_cast<const uint32_t*>(smem + L.configs);
  cv.ans = code.ans;
  cv.prefix = code.prefix;
  cv.prefix_meta = code.prefix_meta;
  if (code.use_prefix) {
    warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.prefix_meta), code.prefix_meta, code.num_clusters * 2, lane);
    cv.prefix_meta = reinterpret_cast<const uint32_t*>(smem + L.prefix_meta);
    if (L.prefix != 0xffffffffu) {
      warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.prefix), code.prefix, code.prefix_table_size, lane);
      cv.prefix = reinterpret_cast<const uint32_t*>(smem + L.prefix);
    }
  } else if (L.ans != 0xffffffffu) {
    warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.ans), reinterpret_cast<const uint32_t*>(code.ans),
                    (code.num_clusters << code.log_alphabet_size) * 2, lane);
    cv.ans = reinterpret_cast<const uint64_t*>(smem + L.ans);
  }
  const uint16_t* luts = job.luts;
  if (L.luts != 0xffffffffu) {
    warp_copy_words(reinterpret_cast<uint32_t*>(smem + L.luts), reinterpret_cast<const uint32_t*>(job.luts),
                    (job.lut_total + 1) / 2, lane);
    luts = reinterpret_cast<const uint16_t*>(smem + L.luts);
  }
  int32_t* wp_rows = (L.wp != 0xffffffffu) ? reinterpret_cast<int32_t*>(smem + L.wp) : job.wp_scratch;
  __syncwarp();
  if (lane != 0) return;

  // ---- serial decode (lane 0) ----
  DevBitReader br;
  br.init(cs, job.bit_pos);
  uint32_t ans_state = code.use_prefix ? 0x130000u : br.read(32);
  // LZ77 state (lib.rs:346-352)
  uint32_t* window = job.lz_window;
  uint32_t lz_to_copy = 0, lz_copy_pos = 0, lz_decoded = 0;
  const bool lz77 = code.lz77_enabled != 0;
  int err = kDevOk;
  FastWp wp;

  for (uint32_t ci = 0; ci < job.num_channels && err == kDevOk; ++ci) {
    const DevChannel out = chans[ci];
    if (!out.w || !out.h) continue;
    const DevChannelPlan plan = chplans[ci];
    DevChannel prev[kMaxPrev];
    int nprev = 0;
    for (int pj = int(ci) - 1; pj >= 0 && nprev < kMaxPrev; --pj) {
      const DevChannel p = chans[pj];
      if (p.w == out.w && p.h == out.h && p.hshift == out.hshift && p.vshift == out.vshift && p.w && p.h) prev[nprev++] = p;
    }
    const uint32_t width = out.w;
    const bool use_wp = job.use_wp != 0;
    if (use_wp) wp.reset(width, wp_rows, job.wp, s_div);
    const uint16_t* lut = luts + plan.lut_offset;
    for (uint32_t y = 0; y < out.h && err == kDevOk; ++y) {
      int32_t* row = out.ptr + size_t(y) * out.stride;
      const int32_t* rn = y ? row - out.stride : nullptr;
      const int32_t* rnn = y >= 2 ? row - 2 * size_t(out.stride) : nullptr;
      int32_t w = 0, ww = 0;                       // samples at x-1 and x-2 of this row
      int32_t r_m1 = 0, r_0 = 0, r_1 = 0, r_2 = 0;  // previous row at x-1, x, x+1, x+2
      if (rn) {
        r_0 = rn[0];
        r_1 = width > 1 ? rn[1] : 0;
        r_2 = width > 2 ? rn[2] : 0;
      }
      int32_t nn_cur = rnn ? rnn[0] : 0;
      int32_t prev_grad = 0;
      for (uint32_t x = 0; x < width; ++x) {
        int32_t wv, n, nw;
        if (!rn) {
          wv = x ? w : 0;
          n = wv;
          nw = wv;
        } else if (x == 0) {
          n = r_0;
          wv = n;
          nw = n;
        } else {
          wv = w;
          n = r_0;
          nw = r_m1;
        }
        const int32_t ne = (!rn || x + 1 >= width) ? n : r_1;
        const int32_t nee = (!rn || x + 2 >= width) ? ne : r_2;
        const int32_t nn = rnn ? nn_cur : n;
        const int32_t wwv = x >= 2 ? ww : wv;
        // prefetch the next iteration's previous-row samples (independent of the decoded value)
        const int32_t r_3 = (rn && x + 3 < width) ? rn[x + 3] : 0;
        const int32_t nn_next = (rnn && x + 1 < width) ? rnn[x + 1] : 0;
        if (use_wp) wp.predict(n, nw, ne, wv, nn);
        const int32_t w_nw = wsub(wv, nw);
        const int32_t grad = wadd(w_nw, n);
        // ---- leaf selection ----
        uint32_t node_idx;
        if (plan.lut_prop >= 0) {
          int32_t v;
          switch (plan.lut_prop) {
            case 2: v = int32_t(y); break;
            case 3: v = int32_t(x); break;
            case 4: v = int32_t(n < 0 ? 0u - uint32_t(n) : uint32_t(n)); break;
            case 5: v = int32_t(wv < 0 ? 0u - uint32_t(wv) : uint32_t(wv)); break;
            case 6: v = n; break;
            case 7: v = wv; break;
            case 8: v = wsub(wv, prev_grad); break;
            case 9: v = grad; break;
            case 10: v = w_nw; break;
            case 11: v = wsub(nw, n); break;
            case 12: v = wsub(n, ne); break;
            case 13: v = wsub(n, nn); break;
            case 14: v = wsub(wv, wwv); break;
            default: v = use_wp ? wp.max_error : 0; break;
          }
          int64_t d = int64_t(v) - int64_t(plan.lut_base);
          uint32_t li = d < 0 ? 0u : (d >= int64_t(plan.lut_len) ? plan.lut_len - 1 : uint32_t(d));
          node_idx = lut[li];
        } else {
          node_idx = plan.root;
          for (;;) {
            const MaNode nd = tree[node_idx];
            if (nd.property < 0) break;
            int32_t v;
            switch (nd.property) {
              case 0: v = int32_t(ci); break;
              case 1: v = int32_t(job.stream_index); break;
              case 2: v = int32_t(y); break;
              case 3: v = int32_t(x); break;
              case 4: v = int32_t(n < 0 ? 0u - uint32_t(n) : uint32_t(n)); break;
              case 5: v = int32_t(wv < 0 ? 0u - uint32_t(wv) : uint32_t(wv)); break;
              case 6: v = n; break;
              case 7: v = wv; break;
              case 8: v = wsub(wv, prev_grad); break;
              case 9: v = grad; break;
              case 10: v = w_nw; break;
              case 11: v = wsub(nw, n); break;
              case 12: v = wsub(n, ne); break;
              case 13: v = wsub(n, nn); break;
              case 14: v = wsub(wv, wwv); break;
              case 15: v = use_wp ? wp.max_error : 0; break;
              default: {
                uint32_t e = uint32_t(nd.property - 16);
                uint32_t pidx = e >> 2, k = e & 3;
                if (int(pidx) >= nprev) {
                  v = 0;
                } else {
                  const DevChannel& pc = prev[pidx];
                  const int32_t* pr = pc.ptr + size_t(y) * pc.stride;
                  int32_t c = pr[x];
                  if (k == 0) v = c < 0 ? -c : c;
                  else if (k == 1) v = c;
                  else {
                    int32_t g;
                    if (x == 0 && y == 0) g = 0;
                    else if (x == 0) g = pr[-ptrdiff_t(pc.stride)];
                    else if (y == 0) g = pr[x - 1];
                    else g = grad_clamped(pr[ptrdiff_t(x) - ptrdiff_t(pc.stride)], pr[x - 1], pr[ptrdiff_t(x) - 1 - ptrdiff_t(pc.stride)]);
                    v = (k == 2) ? int32_t(abs_diff(c, g)) : wsub(c, g);
                  }
                }
              }
            }
            node_idx = v > nd.value ? nd.a : nd.b;
          }
        }
        const MaNode leaf = tree[node_idx];
        const uint32_t predictor = leaf.a & 0xff, cluster = leaf.a >> 8;
        // ---- entropy decode (lib.rs:476-605) ----
        uint32_t token_value;
        if (!lz77) {
          uint32_t token = cv_read_symbol(cv, ans_state, br, cluster);
          token_value = cv_read_uint(br, cv.configs[cluster], token);
        } else {
          if (lz_to_copy > 0) {
            token_value = window[lz_copy_pos & 0xfffff];
            ++lz_copy_pos;
            --lz_to_copy;
          } else {
            uint32_t token = cv_read_symbol(cv, ans_state, br, cluster);
            if (token >= code.lz77_min_symbol) {
              if (lz_decoded == 0) {
                err = kDevBadStream;
                break;
              }
              uint32_t nc = cv_read_uint(br, code.lz_len_conf, token - code.lz77_min_symbol);
              lz_to_copy = nc + code.lz77_min_length;
              uint32_t dtoken = cv_read_symbol(cv, ans_state, br, code.lz_dist_cluster);
              uint32_t distance = cv_read_uint(br, cv.configs[code.lz_dist_cluster], dtoken);
              if (job.dist_multiplier == 0) {
              } else if (distance < 120) {
                int32_t dd = int32_t(kDevSpecialDistances[distance][0]) +
                             int32_t(job.dist_multiplier) * int32_t(kDevSpecialDistances[distance][1]);
                distance = uint32_t(max(dd - 1, 0));
              } else {
                distance -= 120;
              }
              distance = min(min((1u << 20) - 1, distance) + 1, lz_decoded);
              lz_copy_pos = lz_decoded - distance;
              token_value = window[lz_copy_pos & 0xfffff];
              ++lz_copy_pos;
              --lz_to_copy;
            } else {
              token_value = cv_read_uint(br, cv.configs[cluster], token);
            }
          }
          window[lz_decoded & 0xfffff] = token_value;
          ++lz_decoded;
        }
        const int32_t diff = wadd(wmul(dev_unpack_signed(token_value), int32_t(leaf.b)), leaf.value);
        int32_t pred;
        switch (predictor) {
          case 0: pred = 0; break;
          case 1: pred = wv; break;
          case 2: pred = n; break;
          case 3: pred = int32_t((int64_t(wv) + int64_t(n)) / 2); break;
          case 4: pred = abs_diff(n, nw) < abs_diff(wv, nw) ? wv : n; break;
          case 5: pred = grad_clamped(n, wv, nw); break;
          case 6: pred = int32_t((wp.prediction + 3) >> 3); break;
          case 7: pred = ne; break;
          case 8: pred = nw; break;
          case 9: pred = wwv; break;
          case 10: pred = int32_t((int64_t(wv) + int64_t(nw)) / 2); break;
          case 11: pred = int32_t((int64_t(n) + int64_t(nw)) / 2); break;
          case 12: pred = int32_t((int64_t(n) + int64_t(ne)) / 2); break;
          default:
            pred = int32_t((6 * int64_t(n) - 2 * int64_t(nn) + 7 * int64_t(wv) + int64_t(wwv) + int64_t(nee) +
                            3 * int64_t(ne) + 8) / 16);
            break;
        }
        const int32_t value = wadd(diff, pred);
        row[x] = value;
        if (use_wp) wp.record(value);
        prev_grad = grad;
        ww = w;
        w = value;
        r_m1 = r_0;
        r_0 = r_1;
        r_1 = r_2;
        r_2 = r_3;
        nn_cur = nn_next;
      }
      if (br.pos > job.bit_limit) err = kDevOverrun;
    }
  }
  if (err == kDevOk && !code.use_prefix && ans_state != 0x130000u) err = kDevBadStream;
  if (err == kDevOk && br.pos > job.bit_limit) err = kDevOverrun;
  end_bits[job_idx] = br.pos;
  status[job_idx] = err;
}

// ---------------------------------------------------------------------------------------------
// HF coefficients (jxl-vardct/src/hf_coeff.rs:21-252)
#define JXLB_TABLE_QUAL __device__ __constant__ const
namespace hftab {
#include "../host/jxl_tables.inc"
}
#undef JXLB_TABLE_QUAL

__device__ __constant__ const uint8_t kTInfo[27][5] = {
    {1, 1, 0, 0, 1},  {1, 1, 1, 1, 0},  {1, 1, 2, 1, 0},   {1, 1, 3, 1, 0},    {2, 2, 4, 2, 1},   {4, 4, 5, 3, 1},
    {1, 2, 6, 4, 1},  {2, 1, 6, 4, 0},  {1, 4, 7, 5, 1},   {4, 1, 7, 5, 0},    {2, 4, 8, 6, 1},   {4, 2, 8, 6, 0},
    {1, 1, 9, 1, 0},  {1, 1, 9, 1, 0},  {1, 1, 10, 1, 0},  {1, 1, 10, 1, 0},   {1, 1, 10, 1, 0},  {1, 1, 10, 1, 0},
    {8, 8, 11, 7, 1}, {4, 8, 12, 8, 1}, {8, 4, 12, 8, 0},  {16, 16, 13, 9, 1}, {8, 16, 14, 10, 1}, {16, 8, 14, 10, 0},
    {32, 32, 15, 11, 1}, {16, 32, 16, 12, 1}, {32, 16, 16, 12, 0},
};

constexpr int kHfWarpsPerCta = 2;

struct HfSmem {
  uint32_t cmap, configs, ans, ctxlut, total;
};
__host__ __device__ inline HfSmem hf_layout(const DevHfParams& p) {
  HfSmem L;
  uint32_t off = 0;
  auto take = [&](uint32_t bytes) {
    uint32_t o = off;
    off += (bytes + 15) & ~15u;
    return o;
  };
  L.ctxlut = take(128);
  L.configs = take(p.code.num_clusters * 4);
  // the per-preset cluster map slice: shared only when a single preset exists
  L.cmap = (p.num_hf_presets == 1) ? take(495 * p.num_block_clusters) : 0xffffffffu;
  uint32_t ab = p.code.use_prefix ? 0 : (p.code.num_clusters << p.code.log_alphabet_size) * 8;
  L.ans = (!p.code.use_prefix && ab <= 48 * 1024) ? take(ab) : 0xffffffffu;
  L.total = off;
  return L;
}

__global__ void __launch_bounds__(kHfWarpsPerCta * 32) decode_hf_fast_kernel(const uint8_t* __restrict__ cs, DevFrame f,
                                                                             DevHfParams p,
                                                                             const DevHfJob* __restrict__ jobs,
                                                                             uint64_t* __restrict__ end_bits,
                                                                             int* __restrict__ status, int num_jobs,
                                                                             int first_pass) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_nz[kHfWarpsPerCta][3][32];
  const HfSmem L = hf_layout(p);
  const uint32_t tid = threadIdx.x, nthreads = blockDim.x;
  // ---- stage tables (whole CTA) ----
  uint8_t* s_ctx = smem + L.ctxlut;  // [0..63): freq ctx, [64..127): nonzero ctx
  for (uint32_t i = tid; i < 63; i += nthreads) {
    s_ctx[i] = hftab::kCoeffFreqContext[i];
    s_ctx[64 + i] = hftab::kCoeffNumNonzeroContext[i];
  }
  uint32_t* s_cfg = reinterpret_cast<uint32_t*>(smem + L.configs);
  for (uint32_t i = tid; i < p.code.num_clusters; i += nthreads) s_cfg[i] = __ldg(p.code.configs + i);
  const uint8_t* cmap_base = p.code.cluster_map;
  if (L.cmap != 0xffffffffu) {
    uint8_t* s_cmap = smem + L.cmap;
    for (uint32_t i = tid; i < 495 * p.num_block_clusters; i += nthreads) s_cmap[i] = __ldg(p.code.cluster_map + i);
    cmap_base = s_cmap;
  }
  CodeView cv;
  cv.log_alphabet_size = p.code.log_alphabet_size;
  cv.use_prefix = p.code.use_prefix;
  cv.configs = s_cfg;
  cv.ans = p.code.ans;
  cv.prefix = p.code.prefix;
  cv.prefix_meta = p.code.prefix_meta;
  if (L.ans != 0xffffffffu) {
    uint32_t* s_ans = reinterpret_cast<uint32_t*>(smem + L.ans);
    const uint32_t words = (p.code.num_clusters << p.code.log_alphabet_size) * 2;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.code.ans);
    for (uint32_t i = tid; i < words; i += nthreads) s_ans[i] = __ldg(src + i);
    cv.ans = reinterpret_cast<const uint64_t*>(s_ans);
  }
  __syncthreads();
  const int job_idx = blockIdx.x * kHfWarpsPerCta + int(tid >> 5);
  if (job_idx >= num_jobs || (tid & 31) != 0) return;

  const DevHfJob job = jobs[job_idx];
  const uint32_t nbc = p.num_block_clusters;
  const uint32_t lf_idx_mul = (p.num_lf_thr[0] + 1) * (p.num_lf_thr[1] + 1) * (p.num_lf_thr[2] + 1);
  const uint32_t hf_idx_mul = p.num_qf_thr + 1;
  DevBitReader br;
  br.init(cs, job.bit_pos);
  int err = kDevOk;
  uint32_t hfp_bits = 0;
  while ((1u << hfp_bits) < p.num_hf_presets) ++hfp_bits;
  const uint32_t hfp = br.read(hfp_bits);
  if (hfp >= p.num_hf_presets) err = kDevInvalid;
  const uint8_t* cluster_map = (L.cmap != 0xffffffffu) ? cmap_base : cmap_base + size_t(495) * nbc * (err ? 0 : hfp);
  uint32_t ans_state = p.code.use_prefix ? 0x130000u : br.read(32);

  const uint32_t gx = job.group_idx % p.groups_per_row, gy = job.group_idx / p.groups_per_row;
  const uint32_t gb = p.group_dim_blocks;
  const uint32_t bx0 = gx * gb, by0 = gy * gb;
  const uint32_t width = min(gb, f.bw - bx0), height = min(gb, f.bh - by0);
  uint32_t(*nz_row)[32] = s_nz[tid >> 5];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 32; ++i) nz_row[c][i] = 0;
  const int32_t* thr_base[3] = {p.lf_thresholds, p.lf_thresholds + p.num_lf_thr[0],
                                p.lf_thresholds + p.num_lf_thr[0] + p.num_lf_thr[1]};

  for (uint32_t y = 0; y < height && err == kDevOk; ++y)
    for (uint32_t x = 0; x < width && err == kDevOk; ++x) {
      const size_t gi = size_t(by0 + y) * f.bw + bx0 + x;
      const int32_t t = f.blk_type[gi];
      if (t < 0) continue;
      const int32_t qf = f.blk_mul[gi];
      const uint32_t w8 = kTInfo[t][0], h8 = kTInfo[t][1];
      const uint32_t order_id = kTInfo[t][3];
      const bool transpose = kTInfo[t][4] != 0;
      const uint32_t num_blocks = w8 * h8;
      const uint32_t num_blocks_log = 31u - uint32_t(__clz(int(num_blocks)));
      uint32_t lf_idx = 0;
      {
        const int cs3[3] = {0, 2, 1};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int c = cs3[k];
          lf_idx *= p.num_lf_thr[c] + 1;
          const int32_t q = f.lf_quant[c][gi];
          for (uint32_t i = 0; i < p.num_lf_thr[c]; ++i)
            if (q > thr_base[c][i]) ++lf_idx;
        }
      }
      uint32_t hf_idx = 0;
      for (uint32_t i = 0; i < p.num_qf_thr; ++i)
        if (qf > int32_t(p.qf_thresholds[i])) ++hf_idx;
#pragma unroll 1
      for (int ci = 0; ci < 3 && err == kDevOk; ++ci) {
        const uint32_t ch_idx = uint32_t(ci) * 13 + order_id;
        const int c = (ci == 0) ? 1 : (ci == 1 ? 0 : 2);
        const uint32_t idx = (ch_idx * hf_idx_mul + hf_idx) * lf_idx_mul + lf_idx;
        const uint32_t block_ctx = p.block_ctx_map[idx];
        uint32_t predicted;
        const uint32_t nz_here = nz_row[c][x];
        const uint32_t nz_left = x ? nz_row[c][x - 1] : 0;
        if (y == 0) predicted = x == 0 ? 32 : nz_left;
        else if (x == 0) predicted = nz_here;
        else predicted = (nz_here + nz_left + 1) >> 1;
        const uint32_t pidx = predicted >= 8 ? 4 + predicted / 2 : predicted;
        const uint32_t nz_ctx = block_ctx + pidx * nbc;
        uint32_t cl = cluster_map[nz_ctx];
        uint32_t non_zeros = cv_read_uint(br, s_cfg[cl], cv_read_symbol(cv, ans_state, br, cl));
        if (non_zeros > (63u << num_blocks_log)) {
          err = kDevInvalid;
          break;
        }
        const uint32_t nz_val = (non_zeros + num_blocks - 1) >> num_blocks_log;
        for (uint32_t dx = 0; dx < w8; ++dx) nz_row[c][x + dx] = nz_val;
        if (non_zeros == 0) continue;
        uint32_t prev_nonzero = (non_zeros <= num_blocks * 4) ? 1 : 0;
        const uint32_t* order = p.orders + p.order_offset[order_id * 3 + c];
        const uint32_t size = num_blocks * 64;
        const uint8_t* cmap = cluster_map + block_ctx * 458 + 37 * nbc;
        uint32_t* plane = f.coeff[c];
        const size_t base = (size_t(by0 + y) * 8) * f.cw + size_t(bx0 + x) * 8;
        uint32_t o_next = __ldg(order + num_blocks);
        for (uint32_t k = num_blocks, i = 0; k < size; ++k, ++i) {
          const uint32_t o = o_next;
          if (k + 1 < size) o_next = __ldg(order + k + 1);
          const uint32_t nzc = (non_zeros - 1) >> num_blocks_log;
          const uint32_t fi = i >> num_blocks_log;
          const uint32_t cctx = (uint32_t(s_ctx[64 + nzc]) + uint32_t(s_ctx[fi])) * 2 + prev_nonzero;
          if (cctx >= 458) {
            err = kDevInvalid;
            break;
          }
          cl = cmap[cctx];
          const uint32_t ucoeff = cv_read_uint(br, s_cfg[cl], cv_read_symbol(cv, ans_state, br, cl));
          if (ucoeff == 0) {
            prev_nonzero = 0;
            continue;
          }
          const uint32_t cvv = uint32_t(dev_unpack_signed(ucoeff)) << p.coeff_shift;
          uint32_t dx = o & 0xffff, dy = o >> 16;
          if (transpose) {
            const uint32_t tmp = dx;
            dx = dy;
            dy = tmp;
          }
          uint32_t* dst = plane + base + size_t(dy) * f.cw + dx;
          if (first_pass) *dst = cvv;
          else *dst += cvv;
          prev_nonzero = 1;
          if (--non_zeros == 0) break;
        }
        if (br.pos > job.bit_limit) err = kDevOverrun;
      }
    }
  if (err == kDevOk && !p.code.use_prefix && ans_state != 0x130000u) err = kDevBadStream;
  if (err == kDevOk && br.pos > job.bit_limit) err = kDevOverrun;
  end_bits[job_idx] = br.pos;
  status[job_idx] = err;
}

}  // namespace

size_t modular_job_smem_bytes(const DevModularJob& job, uint32_t max_width) {
  return modular_layout(job.num_tree_nodes, job.code, job.lut_total, job.use_wp, max_width).total;
}

void launch_modular_decode(const uint8_t* cs, const DevModularJob* jobs, const DevChannel* channels,
                           const DevChannelPlan* plans, uint64_t* end_bits, int* status, int num_jobs, size_t smem_bytes,
                           cudaStream_t stream) {
  if (num_jobs <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(modular_decode_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_set = true;
  }
  modular_decode_fast_kernel<<<num_jobs, 32, smem_bytes, stream>>>(cs, jobs, channels, plans, end_bits, status, num_jobs);
}

void launch_decode_hf(const uint8_t* cs, DevFrame f, DevHfParams p, const DevHfJob* jobs, uint64_t* end_bits, int* status,
                      int num_jobs, int first_pass, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(decode_hf_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_set = true;
  }
  const HfSmem L = hf_layout(p);
  const int ctas = (num_jobs + kHfWarpsPerCta - 1) / kHfWarpsPerCta;
  decode_hf_fast_kernel<<<ctas, kHfWarpsPerCta * 32, L.total, stream>>>(cs, f, p, jobs, end_bits, status, num_jobs, first_pass);
}

}  // namespace jxlb
