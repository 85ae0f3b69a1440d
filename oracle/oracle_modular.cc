// TEST INFRASTRUCTURE — see oracle_backend.h.
// Modular sample decode: MA-tree walk, 14 predictors, weighted (self-correcting) predictor,
// inverse Squeeze / RCT / Palette. Restates crates/jxl-modular/src/{image.rs,predictor.rs,ma.rs,
// transform/{squeeze,rct,palette}.rs} of the reference (i32 sample type).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>

#include "oracle_backend.h"

namespace jxlo {

namespace {

inline int32_t wadd(int32_t a, int32_t b) { return int32_t(uint32_t(a) + uint32_t(b)); }
inline int32_t wsub(int32_t a, int32_t b) { return int32_t(uint32_t(a) - uint32_t(b)); }
inline int32_t wmul(int32_t a, int32_t b) { return int32_t(uint32_t(a) * uint32_t(b)); }
inline uint32_t abs_diff(int32_t a, int32_t b) { return a > b ? uint32_t(a) - uint32_t(b) : uint32_t(b) - uint32_t(a); }

inline int32_t grad_clamped(int32_t n, int32_t w, int32_t nw) {  // sample.rs:130-138
  int64_t hi = std::max<int64_t>(n, w), lo = std::min<int64_t>(n, w);
  int64_t v = lo + hi - int64_t(nw);
  return int32_t(std::min(std::max(v, lo), hi));
}

inline uint32_t ilog2_u64(uint64_t v) {
  uint32_t r = 0;
  while (v >>= 1) ++r;
  return r;
}

struct ChannelRef {
  int32_t* base;
  size_t stride;
  uint32_t w, h;
  int32_t at(uint32_t x, uint32_t y) const { return base[size_t(y) * stride + x]; }
};

// SelfCorrectingPredictor, predictor.rs:279-441
struct WpState {
  uint32_t width = 0, x = 0, y = 0;
  std::vector<int32_t> true_err_row;
  std::vector<uint32_t> subpred_err_row;  // 4 per column
  WpHeader wp;
  int32_t true_err_w = 0, true_err_nw = 0, true_err_n = 0, true_err_ne = 0;
  uint32_t err_nw_ww[4] = {}, err_n_w[4] = {}, err_ne[4] = {};
  // last prediction
  int64_t prediction = 0;
  int32_t max_error = 0;
  int64_t subpred[4] = {};

  void reset(uint32_t w, const WpHeader& h) {
    *this = WpState();
    width = w;
    wp = h;
    true_err_row.assign(w, 0);
    subpred_err_row.assign(size_t(w) * 4, 0);
  }
  void predict(int32_t n, int32_t nw, int32_t ne, int32_t w, int32_t nn) {  // predictor.rs:312-394
    int64_t tew = true_err_w, tenw = true_err_nw, ten = true_err_n, tene = true_err_ne;
    // (x * 8: the reference's `x << 3` on i64, written without shifting a negative value)
    int64_t n3 = int64_t(n) * 8, nw3 = int64_t(nw) * 8, ne3 = int64_t(ne) * 8, w3 = int64_t(w) * 8, nn3 = int64_t(nn) * 8;
    subpred[0] = w3 + ne3 - n3;
    subpred[1] = n3 - (((tew + ten + tene) * int64_t(wp.p1)) >> 5);
    subpred[2] = w3 - (((tew + ten + tenw) * int64_t(wp.p2)) >> 5);
    subpred[3] = n3 - ((tenw * int64_t(wp.p3a) + ten * int64_t(wp.p3b) + tene * int64_t(wp.p3c) +
                        (nn3 - n3) * int64_t(wp.p3d) + (nw3 - w3) * int64_t(wp.p3e)) >> 5);
    uint32_t weight[4];
    for (int i = 0; i < 4; ++i) {
      uint32_t err_sum = err_nw_ww[i] + err_n_w[i] + err_ne[i];
      uint64_t t = (uint64_t(err_sum) + 1) >> 5;
      uint32_t shift = t ? ilog2_u64(t) : 0;
      uint32_t div = (1u << 24) / ((err_sum >> shift) + 1);
      weight[i] = 4 + ((wp.w[i] * div) >> shift);
    }
    uint32_t sum_weights = weight[0] + weight[1] + weight[2] + weight[3];
    uint32_t log_weight = ilog2_u64(uint64_t(sum_weights) >> 4);
    for (uint32_t& w_ : weight) w_ >>= log_weight;
    sum_weights = weight[0] + weight[1] + weight[2] + weight[3];
    int64_t s = (int64_t(sum_weights) >> 1) - 1;
    for (int i = 0; i < 4; ++i) s += subpred[i] * int64_t(weight[i]);
    int64_t pred = (s * int64_t((1u << 24) / sum_weights)) >> 24;
    if (((ten ^ tew) | (ten ^ tenw)) <= 0) {
      int64_t mn = std::min(std::min(n3, w3), ne3), mx = std::max(std::max(n3, w3), ne3);
      pred = std::min(std::max(pred, mn), mx);
    }
    int64_t me = tew;
    for (int64_t e : {ten, tenw, tene})
      if (std::llabs(e) > std::llabs(me)) me = e;
    prediction = pred;
    max_error = int32_t(me);
  }
  void record(int32_t sample_) {  // predictor.rs:396-441
    int64_t sample = sample_;
    int64_t true_err = prediction - sample * 8;
    uint32_t sub_err[4];
    for (int i = 0; i < 4; ++i) {
      int64_t d = subpred[i] - sample * 8;
      uint64_t ad = d < 0 ? uint64_t(-d) : uint64_t(d);
      sub_err[i] = uint32_t((ad + 3) >> 3);
    }
    true_err_row[x] = int32_t(true_err);
    for (int i = 0; i < 4; ++i) subpred_err_row[size_t(x) * 4 + i] = sub_err[i];
    ++x;
    if (x >= width) {
      ++y;
      x = 0;
      true_err_w = 0;
      true_err_n = true_err_row[0];
      true_err_nw = true_err_n;
      for (int i = 0; i < 4; ++i) err_n_w[i] = err_nw_ww[i] = subpred_err_row[i];
      if (width <= 1) {
        true_err_ne = true_err_n;
        for (int i = 0; i < 4; ++i) err_ne[i] = err_n_w[i];
      } else {
        true_err_ne = true_err_row[1];
        for (int i = 0; i < 4; ++i) err_ne[i] = subpred_err_row[4 + i];
      }
    } else {
      true_err_w = int32_t(true_err);
      true_err_nw = true_err_n;
      true_err_n = true_err_ne;
      for (int i = 0; i < 4; ++i) {
        err_nw_ww[i] = err_n_w[i];
        err_n_w[i] = err_ne[i] + sub_err[i];
      }
      if (x + 1 >= width) {
        true_err_ne = true_err_n;
        for (int i = 0; i < 4; ++i) err_ne[i] = err_n_w[i];
      } else if (y != 0) {
        true_err_ne = true_err_row[x + 1];
        for (int i = 0; i < 4; ++i) err_ne[i] = subpred_err_row[size_t(x + 1) * 4 + i];
      }
    }
  }
};

struct TreeUse {
  bool wp = false;
  uint32_t max_prev = 0;
};

TreeUse scan_tree(const MaTree& t) {
  TreeUse u;
  for (const MaNode& n : t.nodes) {
    if (n.property >= 0) {
      if (n.property == 15) u.wp = true;
      if (n.property >= 16) u.max_prev = std::max<uint32_t>(u.max_prev, uint32_t(n.property - 16) / 4 + 1);
    } else if ((n.a & 0xff) == 6) {
      u.wp = true;
    }
  }
  return u;
}

}  // namespace

void OracleBackend::decode_modular(std::vector<ModularStreamJob>& jobs) {
  parallel_for(jobs.size(), [&](size_t i) { decode_one_modular(jobs[i]); });
}

// TransformedModularSubimage::decode_inner (image.rs:456-593) via the general path `decode_slow`
// (image.rs:1169-1228); the reference's fast paths are result-identical specialisations.
void OracleBackend::decode_one_modular(ModularStreamJob& job) {
  const MaTree& tree = *job.tree;
  BitReader br(cs_, job.bit_limit / 8, job.bit_pos);
  EntropyReader dec(&tree.code);
  dec.begin(br);
  uint32_t dist_multiplier = 0;
  for (const auto& c : job.channels) dist_multiplier = std::max(dist_multiplier, c.view.w);
  const TreeUse use = scan_tree(tree);
  std::vector<ChannelRef> refs;
  for (const auto& c : job.channels) {
    ChannelRef r{nullptr, 0, c.view.w, c.view.h};
    if (c.view.w && c.view.h) {
      Plane& p = plane(c.view.plane);
      r.base = p.i32() + size_t(c.view.y0) * p.w + c.view.x0;
      r.stride = p.w;
    }
    refs.push_back(r);
  }
  WpState wp;
  for (size_t ci = 0; ci < job.channels.size(); ++ci) {
    const ModularChannelTarget& ct = job.channels[ci];
    const ChannelRef& out = refs[ci];
    if (!out.w || !out.h) continue;
    std::vector<ChannelRef> prev;  // most recent first (image.rs:548,590)
    for (size_t pj = ci; pj-- > 0;) {
      const ModularChannelTarget& pt = job.channels[pj];
      if (!pt.view.w || !pt.view.h) continue;
      if (pt.view.w == ct.view.w && pt.view.h == ct.view.h && pt.hshift == ct.hshift && pt.vshift == ct.vshift)
        prev.push_back(refs[pj]);
    }
    const uint32_t width = out.w;
    if (use.wp) wp.reset(width, job.wp);
    int32_t prev_grad = 0;
    int32_t props[16];
    props[0] = int32_t(ci);
    props[1] = int32_t(job.stream_index);
    for (uint32_t y = 0; y < out.h; ++y) {
      int32_t* row = out.base + size_t(y) * out.stride;
      const int32_t* rn = y ? row - out.stride : nullptr;
      const int32_t* rnn = y >= 2 ? row - 2 * out.stride : nullptr;
      for (uint32_t x = 0; x < width; ++x) {
        // PredictorState (predictor.rs:128-277, 536-577)
        int32_t w, n, nw;
        if (y == 0) {
          w = x ? row[x - 1] : 0;
          n = w;
          nw = w;
        } else if (x == 0) {
          n = rn[0];
          w = n;
          nw = n;
        } else {
          w = row[x - 1];
          n = rn[x];
          nw = rn[x - 1];
        }
        int32_t ne = (!rn || x + 1 >= width) ? n : rn[x + 1];
        int32_t nee = (!rn || x + 2 >= width) ? ne : rn[x + 2];
        int32_t nn = rnn ? rnn[x] : n;
        int32_t ww = x >= 2 ? row[x - 2] : w;
        if (x == 0) prev_grad = 0;
        if (use.wp) wp.predict(n, nw, ne, w, nn);
        int32_t w_nw = wsub(w, nw);
        props[2] = int32_t(y);
        props[3] = int32_t(x);
        props[4] = int32_t(n < 0 ? 0u - uint32_t(n) : uint32_t(n));
        props[5] = int32_t(w < 0 ? 0u - uint32_t(w) : uint32_t(w));
        props[6] = n;
        props[7] = w;
        props[8] = wsub(w, prev_grad);
        props[9] = wadd(w_nw, n);
        props[10] = w_nw;
        props[11] = wsub(nw, n);
        props[12] = wsub(n, ne);
        props[13] = wsub(n, nn);
        props[14] = wsub(w, ww);
        props[15] = use.wp ? wp.max_error : 0;
        // FlatMaTree::get_leaf (ma.rs:330-369) on the unflattened tree
        const MaNode* node = &tree.nodes[0];
        while (node->property >= 0) {
          int32_t v;
          if (node->property < 16) {
            v = props[node->property];
          } else {  // Properties::get_extra (predictor.rs:488-522)
            uint32_t e = uint32_t(node->property - 16);
            uint32_t pidx = e / 4, k = e % 4;
            if (pidx >= prev.size()) {
              v = 0;
            } else {
              const ChannelRef& pc = prev[pidx];
              int32_t c = pc.at(x, y);
              if (k == 0) v = std::abs(c);
              else if (k == 1) v = c;
              else {
                int32_t g;
                if (x == 0 && y == 0) g = 0;
                else if (x == 0) g = pc.at(0, y - 1);
                else if (y == 0) g = pc.at(x - 1, 0);
                else g = grad_clamped(pc.at(x, y - 1), pc.at(x - 1, y), pc.at(x - 1, y - 1));
                v = (k == 2) ? int32_t(abs_diff(c, g)) : wsub(c, g);
              }
            }
          }
          node = &tree.nodes[v > node->value ? node->a : node->b];
        }
        uint32_t predictor = node->a & 0xff, cluster = node->a >> 8;
        uint32_t token = dec.read_varint_clustered(br, cluster, dist_multiplier);
        int32_t diff = wadd(wmul(unpack_signed(token), int32_t(node->b)), node->value);
        int32_t pred;
        switch (predictor) {  // Predictor::predict (predictor.rs:74-126)
          case 0: pred = 0; break;
          case 1: pred = w; break;
          case 2: pred = n; break;
          case 3: pred = int32_t((int64_t(w) + int64_t(n)) / 2); break;
          case 4: pred = abs_diff(n, nw) < abs_diff(w, nw) ? w : n; break;
          case 5: {
            int64_t g = int64_t(n) + int64_t(w) - int64_t(nw);
            int64_t lo = std::min<int64_t>(w, n), hi = std::max<int64_t>(w, n);
            pred = int32_t(std::min(std::max(g, lo), hi));
            break;
          }
          case 6: pred = int32_t((wp.prediction + 3) >> 3); break;
          case 7: pred = ne; break;
          case 8: pred = nw; break;
          case 9: pred = ww; break;
          case 10: pred = int32_t((int64_t(w) + int64_t(nw)) / 2); break;
          case 11: pred = int32_t((int64_t(n) + int64_t(nw)) / 2); break;
          case 12: pred = int32_t((int64_t(n) + int64_t(ne)) / 2); break;
          default:
            pred = int32_t((6 * int64_t(n) - 2 * int64_t(nn) + 7 * int64_t(w) + int64_t(ww) + int64_t(nee) +
                            3 * int64_t(ne) + 8) / 16);
            break;
        }
        int32_t value = wadd(diff, pred);
        row[x] = value;
        if (use.wp) wp.record(value);
        prev_grad = props[9];
      }
      JXLB_CHECK(!br.overrun(), kErrEof, "modular stream truncated");
    }
  }
  JXLB_CHECK(dec.finalize_ok(), kErrBitstream, "invalid ANS final state (modular stream)");
  JXLB_CHECK(!br.overrun(), kErrEof, "modular stream truncated");
  job.end_bit = br.pos();
}

namespace {

int32_t tendency(int32_t a_, int32_t b_, int32_t c_) {  // squeeze.rs:1104-1137 (wrapping i32)
  int32_t a = a_, b = b_, c = c_;
  if (a >= b && b >= c) {
    int32_t x = wadd(wsub(wsub(wmul(4, a), wmul(3, c)), b), 6) / 12;
    if (wsub(x, x & 1) > wmul(2, wsub(a, b))) x = wadd(wmul(2, wsub(a, b)), 1);
    if (wadd(x, x & 1) > wmul(2, wsub(b, c))) x = wmul(2, wsub(b, c));
    return x;
  } else if (a <= b && b <= c) {
    int32_t x = wsub(wsub(wsub(wmul(4, a), wmul(3, c)), b), 6) / 12;
    if (wadd(x, x & 1) < wmul(2, wsub(a, b))) x = wsub(wmul(2, wsub(a, b)), 1);
    if (wsub(x, x & 1) < wmul(2, wsub(b, c))) x = wmul(2, wsub(b, c));
    return x;
  }
  return 0;
}

}  // namespace

int OracleBackend::squeeze_inverse(const View& avg, const View& res, bool horizontal) {
  // SqueezeParams::inverse + inverse_h/v_i32_base (transform.rs:457-493, squeeze.rs:59-90, 803-832)
  uint32_t ow = horizontal ? avg.w + res.w : avg.w;
  uint32_t oh = horizontal ? avg.h : avg.h + res.h;
  int id = alloc_plane(ow, oh, false);
  if (!ow || !oh) return id;
  Plane& out = plane(id);
  auto src = [&](const View& v, uint32_t x, uint32_t y) -> int32_t {
    Plane& p = plane(v.plane);
    return p.i32()[size_t(v.y0 + y) * p.w + v.x0 + x];
  };
  int32_t* o = out.i32();
  if (horizontal) {
    JXLB_CHECK(res.h == avg.h || res.w == 0, kErrBitstream, "squeeze residual size mismatch");
    for (uint32_t y = 0; y < oh; ++y) {
      int32_t a = src(avg, 0, y);
      int32_t left = a;
      for (uint32_t x = 0; x < res.w; ++x) {
        int32_t r = src(res, x, y);
        int32_t next_avg = (x + 1 < avg.w) ? src(avg, x + 1, y) : a;
        int32_t diff = wadd(r, tendency(left, a, next_avg));
        int32_t first = wadd(a, diff / 2);
        int32_t second = wsub(first, diff);
        o[size_t(y) * ow + 2 * x] = first;
        o[size_t(y) * ow + 2 * x + 1] = second;
        a = next_avg;
        left = second;
      }
      if (ow & 1) o[size_t(y) * ow + ow - 1] = src(avg, avg.w - 1, y);
    }
  } else {
    JXLB_CHECK(res.w == avg.w || res.h == 0, kErrBitstream, "squeeze residual size mismatch");
    for (uint32_t x = 0; x < ow; ++x) {
      int32_t a = src(avg, x, 0);
      int32_t top = a;
      for (uint32_t y = 0; y < res.h; ++y) {
        int32_t r = src(res, x, y);
        int32_t next_avg = (y + 1 < avg.h) ? src(avg, x, y + 1) : a;
        int32_t diff = wadd(r, tendency(top, a, next_avg));
        int32_t first = wadd(a, diff / 2);
        int32_t second = wsub(first, diff);
        o[size_t(2 * y) * ow + x] = first;
        o[size_t(2 * y + 1) * ow + x] = second;
        a = next_avg;
        top = second;
      }
      if (oh & 1) o[size_t(oh - 1) * ow + x] = src(avg, x, avg.h - 1);
    }
  }
  return id;
}

void OracleBackend::rct_inverse(const View v[3], uint32_t rct_type) {  // rct.rs:87-256
  uint32_t permutation = rct_type / 7, ty = rct_type % 7;
  int32_t* base[3];
  size_t stride[3];
  for (int c = 0; c < 3; ++c) {
    Plane& p = plane(v[c].plane);
    base[c] = p.i32() + size_t(v[c].y0) * p.w + v[c].x0;
    stride[c] = p.w;
  }
  for (uint32_t y = 0; y < v[0].h; ++y)
    for (uint32_t x = 0; x < v[0].w; ++x) {
      int32_t& ra = base[0][y * stride[0] + x];
      int32_t& rb = base[1][y * stride[1] + x];
      int32_t& rc = base[2][y * stride[2] + x];
      int32_t a = ra, b = rb, c = rc, d, e, f;
      if (ty == 6) {
        int32_t tmp = wsub(a, c >> 1);
        e = wadd(c, tmp);
        f = wsub(tmp, b >> 1);
        d = wadd(f, b);
      } else {
        d = a;
        f = (ty & 1) ? wadd(c, a) : c;
        if ((ty >> 1) == 1) e = wadd(b, a);
        else if ((ty >> 1) == 2) e = wadd(b, wadd(a, f) >> 1);
        else e = b;
      }
      // inverse_permute (rct.rs:232-256)
      switch (permutation) {
        case 1: ra = f, rb = d, rc = e; break;
        case 2: ra = e, rb = f, rc = d; break;
        case 3: ra = d, rb = f, rc = e; break;
        case 4: ra = e, rb = d, rc = f; break;
        case 5: ra = f, rb = e, rc = d; break;
        default: ra = d, rb = e, rc = f; break;
      }
    }
}

namespace {
const int16_t kDeltaPalette[72][3] = {  // palette.rs:10-24 (normative table)
#include "../jxl_oxide_b200/csrc/host/delta_palette.inc"
};
}

void OracleBackend::palette_inverse(const View& palette, const std::vector<View>& targets, const Transform& t,
                                    const WpHeader& wph, uint32_t bit_depth) {  // palette.rs:26-173
  const int32_t nb_deltas = int32_t(t.nb_deltas), nb_colors = int32_t(t.nb_colours);
  const uint32_t width = targets[0].w, height = targets[0].h;
  const size_t channels = targets.size();
  Plane* pp = palette.plane >= 0 ? &plane(palette.plane) : nullptr;  // absent when nb_colours == 0
  auto pal = [&](int32_t index, size_t c) { return pp->i32()[size_t(palette.y0 + c) * pp->w + palette.x0 + index]; };
  std::vector<int32_t*> base(channels);
  std::vector<size_t> stride(channels);
  for (size_t c = 0; c < channels; ++c) {
    Plane& p = plane(targets[c].plane);
    base[c] = p.i32() + size_t(targets[c].y0) * p.w + targets[c].x0;
    stride[c] = p.w;
  }
  std::vector<std::pair<uint32_t, uint32_t>> need_delta;
  for (uint32_t y = 0; y < height; ++y)
    for (uint32_t x = 0; x < width; ++x) {
      int32_t index = base[0][y * stride[0] + x];
      if (index < nb_deltas) need_delta.push_back({x, y});
      for (size_t c = 0; c < channels; ++c) {
        int32_t& sample = base[c][y * stride[c] + x];
        if (index >= 0 && index < nb_colors) {
          sample = pal(index, c);
        } else if (index >= nb_colors) {
          int32_t idx = index - nb_colors;
          if (idx < 64) {
            sample = ((idx >> (2 * c)) % 4) * ((1 << bit_depth) - 1) / 4 + (1 << (bit_depth > 3 ? bit_depth - 3 : 0));
          } else {
            int32_t k = idx - 64;
            for (size_t q = 0; q < c; ++q) k /= 5;
            sample = (k % 5) * ((1 << bit_depth) - 1) / 4;
          }
        } else {
          if (c >= 3) {
            sample = 0;
            continue;
          }
          int32_t i2 = -(index + 1);
          uint32_t ii = uint32_t(i2 % 143);
          int32_t ts = kDeltaPalette[(ii + 1) >> 1][c];
          if ((ii & 1) == 0) ts = -ts;
          if (bit_depth > 8) ts <<= std::min<uint32_t>(bit_depth, 24) - 8;
          sample = ts;
        }
      }
    }
  if (need_delta.empty()) return;
  // delta-palette prediction pass (palette.rs:120-152): properties::<true>() + Predictor::predict
  for (size_t c = 0; c < channels; ++c) {
    WpState wp;
    const bool use_wp = t.d_pred == 6;
    if (use_wp) wp.reset(width, wph);
    size_t idx = 0;
    bool done = false;
    for (uint32_t y = 0; y < height && !done; ++y) {
      int32_t* row = base[c] + size_t(y) * stride[c];
      const int32_t* rn = y ? row - stride[c] : nullptr;
      const int32_t* rnn = y >= 2 ? row - 2 * stride[c] : nullptr;
      for (uint32_t x = 0; x < width; ++x) {
        int32_t w, n, nw;
        if (y == 0) {
          w = x ? row[x - 1] : 0;
          n = w, nw = w;
        } else if (x == 0) {
          n = rn[0];
          w = n, nw = n;
        } else {
          w = row[x - 1], n = rn[x], nw = rn[x - 1];
        }
        int32_t ne = (!rn || x + 1 >= width) ? n : rn[x + 1];
        int32_t nee = (!rn || x + 2 >= width) ? ne : rn[x + 2];
        int32_t nn = rnn ? rnn[x] : n;
        int32_t ww = x >= 2 ? row[x - 2] : w;
        if (use_wp) wp.predict(n, nw, ne, w, nn);
        int32_t value = row[x];
        if (need_delta[idx] == std::make_pair(x, y)) {
          int32_t pred;
          switch (t.d_pred) {
            case 0: pred = 0; break;
            case 1: pred = w; break;
            case 2: pred = n; break;
            case 3: pred = int32_t((int64_t(w) + n) / 2); break;
            case 4: pred = abs_diff(n, nw) < abs_diff(w, nw) ? w : n; break;
            case 5: pred = grad_clamped(n, w, nw); break;
            case 6: pred = int32_t((wp.prediction + 3) >> 3); break;
            case 7: pred = ne; break;
            case 8: pred = nw; break;
            case 9: pred = ww; break;
            case 10: pred = int32_t((int64_t(w) + nw) / 2); break;
            case 11: pred = int32_t((int64_t(n) + nw) / 2); break;
            case 12: pred = int32_t((int64_t(n) + ne) / 2); break;
            default:
              pred = int32_t((6 * int64_t(n) - 2 * int64_t(nn) + 7 * int64_t(w) + ww + nee + 3 * int64_t(ne) + 8) / 16);
              break;
          }
          value = wadd(value, pred);
          row[x] = value;
          ++idx;
          if (idx >= need_delta.size()) {
            done = true;
            break;
          }
        }
        if (use_wp) wp.record(value);
      }
    }
  }
}

}  // namespace jxlo
