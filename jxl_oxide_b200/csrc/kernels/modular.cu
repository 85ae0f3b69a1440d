// Modular sub-codec on the device: entropy-coded channel decode (MA tree walk, 14 predictors,
// weighted predictor), inverse Squeeze, inverse RCT, simple palette, sample conversions.
// Integer arithmetic is bit-exact with crates/jxl-modular/src/{image.rs,predictor.rs,
// transform/{squeeze,rct,palette}.rs} (i32 samples, wrapping).
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
  int64_t hi = n > w ? n : w, lo = n > w ? w : n;
  int64_t v = lo + hi - int64_t(nw);
  return int32_t(v < lo ? lo : (v > hi ? hi : v));
}
__device__ __forceinline__ uint32_t ilog2_u32(uint32_t v) { return 31u - uint32_t(__clz(int(v))); }
__device__ __forceinline__ int64_t abs64(int64_t v) { return v < 0 ? -v : v; }

// SelfCorrectingPredictor (predictor.rs:279-441). Row state lives in global scratch:
// true_err_row[width] followed by subpred_err_row[width][4].
struct DevWp {
  uint32_t width, x, y;
  int32_t* true_err_row;
  uint32_t* sub_err_row;
  uint32_t p1, p2, p3a, p3b, p3c, p3d, p3e, w[4];
  int32_t te_w, te_nw, te_n, te_ne;
  uint32_t e_nw_ww[4], e_n_w[4], e_ne[4];
  int64_t prediction;
  int32_t max_error;
  int64_t subpred[4];

  __device__ void reset(uint32_t width_, int32_t* scratch, const uint32_t* hdr) {
    width = width_;
    x = y = 0;
    true_err_row = scratch;
    sub_err_row = reinterpret_cast<uint32_t*>(scratch + width_);
    for (uint32_t i = 0; i < width_ * 5; ++i) scratch[i] = 0;
    p1 = hdr[0], p2 = hdr[1], p3a = hdr[2], p3b = hdr[3], p3c = hdr[4], p3d = hdr[5], p3e = hdr[6];
    for (int i = 0; i < 4; ++i) w[i] = hdr[7 + i];
    te_w = te_nw = te_n = te_ne = 0;
    for (int i = 0; i < 4; ++i) e_nw_ww[i] = e_n_w[i] = e_ne[i] = 0;
    prediction = 0;
    max_error = 0;
  }
  __device__ void predict(int32_t n, int32_t nw, int32_t ne, int32_t wv, int32_t nn) {
    int64_t tew = te_w, tenw = te_nw, ten = te_n, tene = te_ne;
    int64_t n3 = int64_t(n) << 3, nw3 = int64_t(nw) << 3, ne3 = int64_t(ne) << 3, w3 = int64_t(wv) << 3,
            nn3 = int64_t(nn) << 3;
    subpred[0] = w3 + ne3 - n3;
    subpred[1] = n3 - (((tew + ten + tene) * int64_t(p1)) >> 5);
    subpred[2] = w3 - (((tew + ten + tenw) * int64_t(p2)) >> 5);
    subpred[3] = n3 - ((tenw * int64_t(p3a) + ten * int64_t(p3b) + tene * int64_t(p3c) + (nn3 - n3) * int64_t(p3d) +
                        (nw3 - w3) * int64_t(p3e)) >> 5);
    uint32_t weight[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t err_sum = e_nw_ww[i] + e_n_w[i] + e_ne[i];
      uint32_t t = uint32_t((uint64_t(err_sum) + 1) >> 5);
      uint32_t shift = t ? ilog2_u32(t) : 0;
      uint32_t div = (1u << 24) / ((err_sum >> shift) + 1);
      weight[i] = 4 + ((w[i] * div) >> shift);
    }
    uint32_t sum_weights = weight[0] + weight[1] + weight[2] + weight[3];
    uint32_t log_weight = ilog2_u32(sum_weights >> 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) weight[i] >>= log_weight;
    sum_weights = weight[0] + weight[1] + weight[2] + weight[3];
    int64_t s = (int64_t(sum_weights) >> 1) - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += subpred[i] * int64_t(weight[i]);
    int64_t pred = (s * int64_t((1u << 24) / sum_weights)) >> 24;
    if (((ten ^ tew) | (ten ^ tenw)) <= 0) {
      int64_t mn = min(min(n3, w3), ne3), mx = max(max(n3, w3), ne3);
      pred = min(max(pred, mn), mx);
    }
    int64_t me = tew;
    if (abs64(ten) > abs64(me)) me = ten;
    if (abs64(tenw) > abs64(me)) me = tenw;
    if (abs64(tene) > abs64(me)) me = tene;
    prediction = pred;
    max_error = int32_t(me);
  }
  __device__ void record(int32_t sample_) {
    int64_t sample = sample_;
    int64_t true_err = prediction - (sample << 3);
    uint32_t sub_err[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sub_err[i] = uint32_t((uint64_t(abs64(subpred[i] - (sample << 3))) + 3) >> 3);
    true_err_row[x] = int32_t(true_err);
#pragma unroll
    for (int i = 0; i < 4; ++i) sub_err_row[size_t(x) * 4 + i] = sub_err[i];
    ++x;
    if (x >= width) {
      ++y;
      x = 0;
      te_w = 0;
      te_n = true_err_row[0];
      te_nw = te_n;
#pragma unroll
      for (int i = 0; i < 4; ++i) e_n_w[i] = e_nw_ww[i] = sub_err_row[i];
      if (width <= 1) {
        te_ne = te_n;
#pragma unroll
        for (int i = 0; i < 4; ++i) e_ne[i] = e_n_w[i];
      } else {
        te_ne = true_err_row[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) e_ne[i] = sub_err_row[4 + i];
      }
    } else {
      te_w = int32_t(true_err);
      te_nw = te_n;
      te_n = te_ne;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        e_nw_ww[i] = e_n_w[i];
        e_n_w[i] = e_ne[i] + sub_err[i];
      }
      if (x + 1 >= width) {
        te_ne = te_n;
#pragma unroll
        for (int i = 0; i < 4; ++i) e_ne[i] = e_n_w[i];
      } else if (y != 0) {
        te_ne = true_err_row[x + 1];
#pragma unroll
        for (int i = 0; i < 4; ++i) e_ne[i] = sub_err_row[size_t(x + 1) * 4 + i];
      }
    }
  }
};

constexpr int kMaxPrevChannels = 16;

// One warp per stream; lane 0 walks the serial entropy/prediction chain (jxl-modular image.rs:
// 456-593, general path). The other lanes are idle in this first version.
__global__ void modular_decode_kernel(const uint8_t* __restrict__ cs, const DevModularJob* __restrict__ jobs,
                                      const DevChannel* __restrict__ channels, uint64_t* __restrict__ end_bits,
                                      int* __restrict__ status, int num_jobs) {
  int job_idx = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (job_idx >= num_jobs || (threadIdx.x & 31) != 0) return;
  const DevModularJob& job = jobs[job_idx];
  const DevEntropyCode& code = job.code;
  const MaNode* __restrict__ tree = job.tree;
  DevBitReader br;
  br.init(cs, job.bit_pos);
  DevEntropyState es;
  entropy_begin(code, es, br, job.lz_window);
  int err = kDevOk;
  DevWp wp;
  const DevChannel* chans = channels + job.first_channel;
  for (uint32_t ci = 0; ci < job.num_channels && err == kDevOk; ++ci) {
    const DevChannel out = chans[ci];
    if (!out.w || !out.h) continue;
    DevChannel prev[kMaxPrevChannels];
    int nprev = 0;
    for (int pj = int(ci) - 1; pj >= 0 && nprev < kMaxPrevChannels; --pj) {
      const DevChannel p = chans[pj];
      if (p.w == out.w && p.h == out.h && p.hshift == out.hshift && p.vshift == out.vshift && p.w && p.h) prev[nprev++] = p;
    }
    const uint32_t width = out.w;
    if (job.use_wp) wp.reset(width, job.wp_scratch, job.wp);
    int32_t prev_grad = 0;
    int32_t props[16];
    props[0] = int32_t(ci);
    props[1] = int32_t(job.stream_index);
    for (uint32_t y = 0; y < out.h && err == kDevOk; ++y) {
      int32_t* row = out.ptr + size_t(y) * out.stride;
      const int32_t* rn = y ? row - out.stride : nullptr;
      const int32_t* rnn = y >= 2 ? row - 2 * size_t(out.stride) : nullptr;
      for (uint32_t x = 0; x < width; ++x) {
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
        if (job.use_wp) wp.predict(n, nw, ne, w, nn);
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
        props[15] = job.use_wp ? wp.max_error : 0;
        MaNode node = tree[0];
        while (node.property >= 0) {
          int32_t v;
          if (node.property < 16) {
            v = props[node.property];
          } else {
            uint32_t e = uint32_t(node.property - 16);
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
          node = tree[v > node.value ? node.a : node.b];
        }
        uint32_t predictor = node.a & 0xff, cluster = node.a >> 8;
        uint32_t token = entropy_read_varint(code, es, br, cluster, job.dist_multiplier, err);
        int32_t diff = wadd(wmul(dev_unpack_signed(token), int32_t(node.b)), node.value);
        int32_t pred;
        switch (predictor) {
          case 0: pred = 0; break;
          case 1: pred = w; break;
          case 2: pred = n; break;
          case 3: pred = int32_t((int64_t(w) + int64_t(n)) / 2); break;
          case 4: pred = abs_diff(n, nw) < abs_diff(w, nw) ? w : n; break;
          case 5: pred = grad_clamped(n, w, nw); break;
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
        if (job.use_wp) wp.record(value);
        prev_grad = props[9];
      }
      if (br.pos > job.bit_limit) err = kDevOverrun;
    }
  }
  if (err == kDevOk && !entropy_final_ok(code, es)) err = kDevBadStream;
  if (err == kDevOk && br.pos > job.bit_limit) err = kDevOverrun;
  end_bits[job_idx] = br.pos;
  status[job_idx] = err;
}

__device__ __forceinline__ int32_t tendency(int32_t a, int32_t b, int32_t c) {  // squeeze.rs:1104-1137
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

// One thread per line: the recurrence along the line is serial (`left = second`,
// squeeze.rs:59-90, 803-832); lines are independent.
__global__ void squeeze_h_kernel(DevView avg, DevView res, DevView out) {
  uint32_t y = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= out.h) return;
  const int32_t* ar = static_cast<const int32_t*>(avg.ptr) + size_t(y) * avg.stride;
  const int32_t* rr = static_cast<const int32_t*>(res.ptr) + size_t(y) * res.stride;
  int32_t* o = static_cast<int32_t*>(out.ptr) + size_t(y) * out.stride;
  int32_t a = ar[0], left = a;
  for (uint32_t x = 0; x < res.w; ++x) {
    int32_t next_avg = (x + 1 < avg.w) ? ar[x + 1] : a;
    int32_t diff = wadd(rr[x], tendency(left, a, next_avg));
    int32_t first = wadd(a, diff / 2);
    int32_t second = wsub(first, diff);
    o[2 * x] = first;
    o[2 * x + 1] = second;
    a = next_avg;
    left = second;
  }
  if (out.w & 1) o[out.w - 1] = ar[avg.w - 1];
}

__global__ void squeeze_v_kernel(DevView avg, DevView res, DevView out) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= out.w) return;
  const int32_t* ap = static_cast<const int32_t*>(avg.ptr) + x;
  const int32_t* rp = static_cast<const int32_t*>(res.ptr) + x;
  int32_t* o = static_cast<int32_t*>(out.ptr) + x;
  int32_t a = ap[0], top = a;
  for (uint32_t y = 0; y < res.h; ++y) {
    int32_t next_avg = (y + 1 < avg.h) ? ap[size_t(y + 1) * avg.stride] : a;
    int32_t diff = wadd(rp[size_t(y) * res.stride], tendency(top, a, next_avg));
    int32_t first = wadd(a, diff / 2);
    int32_t second = wsub(first, diff);
    o[size_t(2 * y) * out.stride] = first;
    o[size_t(2 * y + 1) * out.stride] = second;
    a = next_avg;
    top = second;
  }
  if (out.h & 1) o[size_t(out.h - 1) * out.stride] = ap[size_t(avg.h - 1) * avg.stride];
}

__global__ void rct_kernel(DevView va, DevView vb, DevView vc, uint32_t rct_type) {  // rct.rs:154-256
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= va.w) return;
  int32_t* pa = static_cast<int32_t*>(va.ptr) + size_t(y) * va.stride + x;
  int32_t* pb = static_cast<int32_t*>(vb.ptr) + size_t(y) * vb.stride + x;
  int32_t* pc = static_cast<int32_t*>(vc.ptr) + size_t(y) * vc.stride + x;
  uint32_t permutation = rct_type / 7, ty = rct_type % 7;
  int32_t a = *pa, b = *pb, c = *pc, d, e, f;
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
  switch (permutation) {
    case 1: *pa = f, *pb = d, *pc = e; break;
    case 2: *pa = e, *pb = f, *pc = d; break;
    case 3: *pa = d, *pb = f, *pc = e; break;
    case 4: *pa = e, *pb = d, *pc = f; break;
    case 5: *pa = f, *pb = e, *pc = d; break;
    default: *pa = d, *pb = e, *pc = f; break;
  }
}

struct PaletteTargets {
  DevView v[4];
};
// Palette without delta entries (palette.rs:26-118 with need_delta empty): pure lookup plus the
// implicit colour cube for indices >= nb_colours. Indices < nb_deltas flag `status`.
__global__ void palette_kernel(DevView pal, PaletteTargets t, int num_c, int nb_colours, int bit_depth, int nb_deltas,
                               int* status) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= t.v[0].w) return;
  int32_t index = static_cast<int32_t*>(t.v[0].ptr)[size_t(y) * t.v[0].stride + x];
  if (index < nb_deltas) {
    atomicExch(status, kDevInvalid);
    return;
  }
  for (int c = 0; c < num_c; ++c) {
    int32_t sample;
    if (index < nb_colours) {
      sample = static_cast<const int32_t*>(pal.ptr)[size_t(c) * pal.stride + index];
    } else {
      int32_t idx = index - nb_colours;
      if (idx < 64) {
        sample = ((idx >> (2 * c)) % 4) * ((1 << bit_depth) - 1) / 4 + (1 << (bit_depth > 3 ? bit_depth - 3 : 0));
      } else {
        int32_t k = idx - 64;
        for (int q = 0; q < c; ++q) k /= 5;
        sample = (k % 5) * ((1 << bit_depth) - 1) / 4;
      }
    }
    static_cast<int32_t*>(t.v[c].ptr)[size_t(y) * t.v[c].stride + x] = sample;
  }
}

__global__ void int_to_float_kernel(DevView v, uint32_t bits, uint32_t exp_bits, int float_sample) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= v.w) return;
  uint32_t* p = static_cast<uint32_t*>(v.ptr) + size_t(y) * v.stride + x;
  int32_t s = int32_t(*p);
  float f;
  if (!float_sample) {
    int32_t div = int32_t((1u << bits) - 1);
    f = __fdiv_rn(float(s), float(div));
  } else {  // jxl-image/src/lib.rs:464-488
    uint32_t sample = uint32_t(s);
    uint32_t mantissa_bits = bits - exp_bits - 1;
    uint32_t mantissa_mask = (1u << mantissa_bits) - 1;
    uint32_t exp_mask = ((1u << (bits - 1)) - 1) ^ mantissa_mask;
    uint32_t sign = (sample >> (bits - 1)) & 1;
    uint32_t mantissa = sample & mantissa_mask;
    int32_t exp = int32_t((sample & exp_mask) >> mantissa_bits) - ((1 << (exp_bits - 1)) - 1);
    if (mantissa_bits < 23) mantissa <<= (23 - mantissa_bits);
    else if (mantissa_bits > 23) mantissa >>= (mantissa_bits - 23);
    f = __uint_as_float((sign << 31) | (uint32_t(exp + 127) << 23) | mantissa);
  }
  *p = __float_as_uint(f);
}

__global__ void modular_xyb_kernel(DevView vy, DevView vx, DevView vb, float mx, float my, float mb) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= vy.w) return;
  uint32_t* py = static_cast<uint32_t*>(vy.ptr) + size_t(y) * vy.stride + x;
  uint32_t* px = static_cast<uint32_t*>(vx.ptr) + size_t(y) * vx.stride + x;
  uint32_t* pb = static_cast<uint32_t*>(vb.ptr) + size_t(y) * vb.stride + x;
  int64_t bsum = int64_t(int32_t(*pb)) + int64_t(int32_t(*py));
  int32_t bi = int32_t(bsum < INT32_MIN ? INT32_MIN : (bsum > INT32_MAX ? INT32_MAX : bsum));
  float fy = float(int32_t(*py)), fx = float(int32_t(*px)), fb = float(bi);
  *py = __float_as_uint(__fmul_rn(fx, mx));
  *px = __float_as_uint(__fmul_rn(fy, my));
  *pb = __float_as_uint(__fmul_rn(fb, mb));
}

__global__ void fill_kernel(uint32_t* p, size_t n, uint32_t v) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

__global__ void copy_rect_kernel(DevView src, DevView dst) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= src.w) return;
  static_cast<uint32_t*>(dst.ptr)[size_t(y) * dst.stride + x] = static_cast<const uint32_t*>(src.ptr)[size_t(y) * src.stride + x];
}

inline dim3 grid2d(uint32_t w, uint32_t h, uint32_t bx = 128) { return dim3((w + bx - 1) / bx, h, 1); }

}  // namespace

// First (unoptimised, global-memory) version; kept as a debugging reference for entropy.cu.
void launch_modular_decode_v1(const uint8_t* cs, const DevModularJob* jobs, const DevChannel* channels, uint64_t* end_bits,
                              int* status, int num_jobs, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  const int warps_per_block = 1;
  modular_decode_kernel<<<(num_jobs + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      cs, jobs, channels, end_bits, status, num_jobs);
}

void launch_squeeze_inverse(DevView avg, DevView res, DevView out, bool horizontal, cudaStream_t stream) {
  if (!out.w || !out.h) return;
  if (horizontal) squeeze_h_kernel<<<(out.h + 63) / 64, 64, 0, stream>>>(avg, res, out);
  else squeeze_v_kernel<<<(out.w + 63) / 64, 64, 0, stream>>>(avg, res, out);
}

void launch_rct_inverse(DevView a, DevView b, DevView c, uint32_t rct_type, cudaStream_t stream) {
  if (!a.w || !a.h) return;
  rct_kernel<<<grid2d(a.w, a.h), 128, 0, stream>>>(a, b, c, rct_type);
}

void launch_palette_inverse_simple(DevView palette, const DevView* targets, int num_c, int nb_colours, int bit_depth,
                                   int nb_deltas, int* status, cudaStream_t stream) {
  PaletteTargets t;
  for (int i = 0; i < num_c && i < 4; ++i) t.v[i] = targets[i];
  if (!t.v[0].w || !t.v[0].h) return;
  palette_kernel<<<grid2d(t.v[0].w, t.v[0].h), 128, 0, stream>>>(palette, t, num_c, nb_colours, bit_depth, nb_deltas, status);
}

void launch_int_to_float(DevView v, uint32_t bits, uint32_t exp_bits, bool float_sample, cudaStream_t stream) {
  if (!v.w || !v.h) return;
  int_to_float_kernel<<<grid2d(v.w, v.h), 128, 0, stream>>>(v, bits, exp_bits, float_sample ? 1 : 0);
}

void launch_modular_xyb(DevView y, DevView x, DevView b, float mx, float my, float mb, cudaStream_t stream) {
  if (!y.w || !y.h) return;
  modular_xyb_kernel<<<grid2d(y.w, y.h), 128, 0, stream>>>(y, x, b, mx, my, mb);
}

void launch_fill_u32(uint32_t* p, size_t n, uint32_t value, cudaStream_t stream) {
  if (!n) return;
  size_t blocks = (n + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fill_kernel<<<unsigned(blocks), 256, 0, stream>>>(p, n, value);
}

void launch_copy_rect(DevView src, DevView dst, cudaStream_t stream) {
  if (!src.w || !src.h) return;
  copy_rect_kernel<<<grid2d(src.w, src.h), 128, 0, stream>>>(src, dst);
}

}  // namespace jxlb
