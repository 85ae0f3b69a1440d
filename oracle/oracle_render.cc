// TEST INFRASTRUCTURE — see oracle_backend.h.
// Plane management, sample conversions, Gaborish, edge-preserving filter, upsampling and
// XYB -> sRGB. Restates crates/jxl-render/src/filter/{gabor.rs,epf.rs},
// filter/impls/generic/{gabor.rs,epf.rs}, features/upsampling.rs, image.rs:93-189 and
// crates/jxl-color/src/{xyb.rs:35-60,ciexyz.rs:81-87,tf/srgb.rs:13-48}.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <type_traits>

#include "oracle_backend.h"

namespace jxlo {

int OracleBackend::alloc_plane(uint32_t w, uint32_t h, bool /*zero*/) {
  int id = next_id_++;
  Plane& p = planes_[id];
  p.w = w;
  p.h = h;
  p.data.assign(size_t(w) * h, 0);
  return id;
}
void OracleBackend::free_plane(int id) { planes_.erase(id); }

void OracleBackend::download_rect(const View& v, void* dst) {
  Plane& p = plane(v.plane);
  uint32_t* o = static_cast<uint32_t*>(dst);
  for (uint32_t y = 0; y < v.h; ++y)
    std::memcpy(o + size_t(y) * v.w, p.data.data() + size_t(v.y0 + y) * p.w + v.x0, size_t(v.w) * 4);
}

void OracleBackend::copy_rect(const View& src, const View& dst) {
  Plane& s = plane(src.plane);
  Plane& d = plane(dst.plane);
  for (uint32_t y = 0; y < src.h; ++y)
    std::memcpy(d.data.data() + size_t(dst.y0 + y) * d.w + dst.x0, s.data.data() + size_t(src.y0 + y) * s.w + src.x0,
                size_t(src.w) * 4);
}

void OracleBackend::stage_marker(const char* name, const View* views, int n) {
  if (!capture) return;
  auto& out = stages[name];
  auto& dims = stage_dims[name];
  out.clear();
  dims.clear();
  for (int i = 0; i < n; ++i) {
    std::vector<uint32_t> buf(size_t(views[i].w) * views[i].h);
    download_rect(views[i], buf.data());
    out.push_back(std::move(buf));
    dims.push_back({views[i].w, views[i].h});
  }
}

// BitDepth::parse_integer_sample (jxl-image/src/lib.rs:458-490)
void OracleBackend::int_to_float(const View& v, const BitDepth& d) {
  Plane& p = plane(v.plane);
  for (uint32_t y = 0; y < v.h; ++y) {
    uint32_t* row = p.data.data() + size_t(v.y0 + y) * p.w + v.x0;
    for (uint32_t x = 0; x < v.w; ++x) {
      int32_t s = int32_t(row[x]);
      float f;
      if (!d.float_sample) {
        int32_t div = int32_t((1u << d.bits_per_sample) - 1);
        f = float(s) / float(div);
      } else {
        uint32_t sample = uint32_t(s);
        uint32_t mantissa_bits = d.bits_per_sample - d.exp_bits - 1;
        uint32_t mantissa_mask = (1u << mantissa_bits) - 1;
        uint32_t exp_mask = ((1u << (d.bits_per_sample - 1)) - 1) ^ mantissa_mask;
        bool is_signed = (sample & (1u << (d.bits_per_sample - 1))) != 0;
        uint32_t mantissa = sample & mantissa_mask;
        int32_t exp = int32_t((sample & exp_mask) >> mantissa_bits);
        exp = exp - ((1 << (d.exp_bits - 1)) - 1);
        if (mantissa_bits < 23) mantissa <<= (23 - mantissa_bits);
        else if (mantissa_bits > 23) mantissa >>= (mantissa_bits - 23);
        uint32_t bits = (uint32_t(is_signed) << 31) | (uint32_t(exp + 127) << 23) | mantissa;
        std::memcpy(&f, &bits, 4);
      }
      std::memcpy(&row[x], &f, 4);
    }
  }
}

// ImageBuffer::convert_to_float_modular_xyb (jxl-render/src/image.rs:148-189)
void OracleBackend::modular_xyb_to_float(const View yxb[3], const float m[3]) {
  Plane& py = plane(yxb[0].plane);
  Plane& px = plane(yxb[1].plane);
  Plane& pb = plane(yxb[2].plane);
  for (uint32_t y = 0; y < yxb[0].h; ++y)
    for (uint32_t x = 0; x < yxb[0].w; ++x) {
      uint32_t& ry = py.data[size_t(yxb[0].y0 + y) * py.w + yxb[0].x0 + x];
      uint32_t& rx = px.data[size_t(yxb[1].y0 + y) * px.w + yxb[1].x0 + x];
      uint32_t& rb = pb.data[size_t(yxb[2].y0 + y) * pb.w + yxb[2].x0 + x];
      int64_t bsum = int64_t(int32_t(rb)) + int64_t(int32_t(ry));
      int32_t bi = int32_t(std::min<int64_t>(std::max<int64_t>(bsum, INT32_MIN), INT32_MAX));
      float fy = float(int32_t(ry)), fx = float(int32_t(rx)), fb = float(bi);
      float o0 = fx * m[0], o1 = fy * m[1], o2 = fb * m[2];
      std::memcpy(&ry, &o0, 4);
      std::memcpy(&rx, &o1, 4);
      std::memcpy(&rb, &o2, 4);
    }
}

// ---------------------------------------------------------------------------------------------
// Gaborish (filter/gabor.rs:59-116, impls/generic/gabor.rs)
namespace {

void gabor_row_edge(const float* row_c, const float* row_a, float* out, size_t width, float w0, float w1) {
  const float global_weight = 1.0f / (1.0f + w0 * 4.0f + w1 * 4.0f);
  if (row_a) {
    if (width == 1) {
      float u = row_a[0], c = row_c[0];
      out[0] = (c * (1.0f + 3.0f * w0 + 2.0f * w1) + u * (w0 + 2.0f * w1)) * global_weight;
      return;
    }
    {
      float a1 = row_a[0], a0 = row_a[1], c1 = row_c[0], c0 = row_c[1];
      out[0] = (c1 * (1.0f + 2.0f * w0 + w1) + (a1 + c0) * (w0 + w1) + a0 * w1) * global_weight;
    }
    for (size_t x = 1; x + 1 < width; ++x) {
      float a0 = row_a[x - 1], a1 = row_a[x], a2 = row_a[x + 1];
      float c0 = row_c[x - 1], c1 = row_c[x], c2 = row_c[x + 1];
      out[x] = (c1 + (a1 + c0 + c1 + c2) * w0 + (a0 + a2 + c0 + c2) * w1) * global_weight;
    }
    {
      float a0 = row_a[width - 2], a1 = row_a[width - 1], c0 = row_c[width - 2], c1 = row_c[width - 1];
      out[width - 1] = (c1 * (1.0f + 2.0f * w0 + w1) + (a1 + c0) * (w0 + w1) + a0 * w1) * global_weight;
    }
  } else {
    if (width == 1) {
      out[0] = row_c[0];
      return;
    }
    float merged_w0 = 1.0f + 2.0f + w0;
    float merged_w1 = w0 + 2.0f * w1;
    out[0] = (row_c[0] * (merged_w0 + merged_w1) + row_c[1] * merged_w1) * global_weight;
    for (size_t x = 1; x + 1 < width; ++x) out[x] = (row_c[x] * merged_w0 + (row_c[x - 1] + row_c[x + 1]) * merged_w1) * global_weight;
    out[width - 1] = (row_c[width - 1] * (merged_w0 + merged_w1) + row_c[width - 2] * merged_w1) * global_weight;
  }
}

void gabor_row(const float* t, const float* c, const float* b, float* out, size_t width, float w0, float w1) {
  if (width == 0) return;
  const float global_weight = 1.0f / (1.0f + w0 * 4.0f + w1 * 4.0f);
  if (width == 1) {
    float sum_side = t[0] + 2.0f * c[0] + b[0];
    float sum_diag = 2.0f * (t[0] + b[0]);
    out[0] = (c[0] + sum_side * w0 + sum_diag * w1) * global_weight;
    return;
  }
  {
    float t1 = t[0], c1 = c[0], b1 = b[0], t0 = t[1], c0 = c[1], b0 = b[1];
    float sum_side = t1 + c0 + c1 + b1, sum_diag = t0 + t1 + b0 + b1;
    out[0] = (c1 + sum_side * w0 + sum_diag * w1) * global_weight;
  }
  for (size_t x = 1; x + 1 < width; ++x) {
    float sum_side = t[x] + c[x - 1] + c[x + 1] + b[x];
    float sum_diag = t[x - 1] + t[x + 1] + b[x - 1] + b[x + 1];
    out[x] = (c[x] + sum_side * w0 + sum_diag * w1) * global_weight;
  }
  {
    float t1 = t[width - 1], c1 = c[width - 1], b1 = b[width - 1], t0 = t[width - 2], c0 = c[width - 2], b0 = b[width - 2];
    float sum_side = t1 + c0 + c1 + b1, sum_diag = t0 + t1 + b0 + b1;
    out[width - 1] = (c1 + sum_side * w0 + sum_diag * w1) * global_weight;
  }
}

}  // namespace

void OracleBackend::gaborish(const View v[3], const float weights[3][2]) {
  for (int c = 0; c < 3; ++c) {
    Plane& p = plane(v[c].plane);
    const size_t width = v[c].w, height = v[c].h, stride = p.w;
    float* base = p.f32() + size_t(v[c].y0) * stride + v[c].x0;
    std::vector<float> out(width * height);
    const float w0 = weights[c][0], w1 = weights[c][1];
    if (height == 1) {
      gabor_row_edge(base, nullptr, out.data(), width, w0, w1);
    } else {
      gabor_row_edge(base, base + stride, out.data(), width, w0, w1);
      parallel_for(height - 2, [&](size_t i) {
        size_t y = i + 1;
        gabor_row(base + (y - 1) * stride, base + y * stride, base + (y + 1) * stride, &out[y * width], width, w0, w1);
      });
      gabor_row_edge(base + (height - 1) * stride, base + (height - 2) * stride, &out[(height - 1) * width], width, w0, w1);
    }
    for (size_t y = 0; y < height; ++y) std::memcpy(base + y * stride, &out[y * width], width * 4);
  }
}

// ---------------------------------------------------------------------------------------------
// Edge-preserving filter (filter/epf.rs:10-291, impls/generic/epf.rs:3-210)
namespace {

inline size_t mirror(ptrdiff_t offset, size_t len) {  // util.rs:376-386
  for (;;) {
    if (offset < 0) offset = -(offset + 1);
    else if (size_t(offset) >= len) offset = ptrdiff_t(len * 2) - (offset + 1);
    else return size_t(offset);
  }
}

const int8_t kKernel1[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
const int8_t kKernel2[12][2] = {{0, -2}, {-1, -1}, {0, -1}, {1, -1}, {-2, 0}, {-1, 0}, {1, 0}, {2, 0}, {-1, 1}, {0, 1}, {1, 1}, {0, 2}};
const int8_t kDist0[5][2] = {{0, -1}, {1, 0}, {0, 0}, {-1, 0}, {0, 1}};
const int8_t kDist1[5][2] = {{0, -1}, {0, 0}, {0, 1}, {-1, 0}, {1, 0}};
const int8_t kDist2[1][2] = {{0, 0}};

}  // namespace

void OracleBackend::epf(const View v[3], const View& sigma, const EpfParams& p, bool sigma_is_constant) {
  const size_t width = v[0].w, height = v[0].h;
  auto T0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* n) { auto t = std::chrono::steady_clock::now(); if (std::getenv("JXLO_TRACE")) std::fprintf(stderr, "  epf %-10s %.1f ms\n", n, std::chrono::duration<double, std::milli>(t - T0).count()); T0 = t; };
  std::vector<float> bufs[2][3];
  for (int c = 0; c < 3; ++c) {
    bufs[0][c].resize(width * height);
    bufs[1][c].resize(width * height);
    Plane& pl = plane(v[c].plane);
    for (size_t y = 0; y < height; ++y)
      std::memcpy(&bufs[0][c][y * width], pl.f32() + (v[c].y0 + y) * size_t(pl.w) + v[c].x0, width * 4);
  }
  const float* sig = nullptr;
  size_t sig_stride = 0;
  if (!sigma_is_constant) {
    Plane& sp = plane(sigma.plane);
    sig = sp.f32();
    sig_stride = sp.w;
  }
  lap("copy_in");
  int cur = 0;
  auto run_step = [&](auto step_tag) {  // the step is a compile-time constant: fixed trip counts for the tap loops
    constexpr int step = decltype(step_tag)::value;
    const int8_t(*kernel)[2] = step == 0 ? kKernel2 : kKernel1;
    constexpr int nk = step == 0 ? 12 : 4;
    const int8_t(*dist)[2] = step == 0 ? kDist0 : (step == 1 ? kDist1 : kDist2);
    constexpr int nd = step == 2 ? 1 : 5;
    const float step_multiplier = step == 0 ? p.pass0_sigma_scale : (step == 2 ? p.pass2_sigma_scale : 1.0f);
    const float* in[3] = {bufs[cur][0].data(), bufs[cur][1].data(), bufs[cur][2].data()};
    float* out[3] = {bufs[cur ^ 1][0].data(), bufs[cur ^ 1][1].data(), bufs[cur ^ 1][2].data()};
    // Pixels whose whole stencil (kernel reach + distance-pattern reach) lies inside the image index the planes
    // directly; the mirrored coordinates of the general path are only needed in a border of `reach` pixels. Same
    // operations in the same order either way (this only keeps the CPU baseline from paying for mirror() per tap).
    const ptrdiff_t reach = step == 0 ? 3 : (step == 1 ? 2 : 1);
    ptrdiff_t koff[12], doff[5];
    for (int k = 0; k < nk; ++k) koff[k] = ptrdiff_t(kernel[k][1]) * ptrdiff_t(width) + kernel[k][0];
    for (int i = 0; i < nd; ++i) doff[i] = ptrdiff_t(dist[i][1]) * ptrdiff_t(width) + dist[i][0];
    parallel_for(height, [&](size_t y) {
      const bool is_y_border = ((y + 1) & 6) == 0;
      float sm[8];
      for (int i = 0; i < 8; ++i) sm[i] = is_y_border ? step_multiplier * p.border_sad_mul : step_multiplier;
      if (!is_y_border) {
        sm[0] *= p.border_sad_mul;
        sm[7] *= p.border_sad_mul;
      }
      const bool row_inside = ptrdiff_t(y) >= reach && ptrdiff_t(y) + reach < ptrdiff_t(height);
      // Interior of the row, array form: the same per-pixel operations in the same order, laid out so that the compiler
      // can run several pixels per instruction (what the reference's SIMD EPF does); pixels with sigma < 0.3 are
      // overwritten with the input afterwards, exactly as the scalar path would have left them.
      size_t xa = width, xb = width;  // [xa, xb): pixels done here
      if (row_inside && ptrdiff_t(width) > 2 * reach) {
        xa = size_t(reach), xb = width - size_t(reach);
        const size_t n = xb - xa, row = y * width + xa;
        std::vector<float> tmp(n * 8);
        float *nis = tmp.data(), *wsum = nis + n, *s0 = wsum + n, *s1 = s0 + n, *s2 = s1 + n, *dd = s2 + n, *acc = dd + n, *sg = acc + n;
        for (size_t i = 0; i < n; ++i) sg[i] = sigma_is_constant ? p.sigma_for_modular : sig[(y / 8) * sig_stride + (xa + i) / 8];
        for (size_t i = 0; i < n; ++i) nis[i] = 6.6f * (0.70710678118654752440f - 1.0f) / sg[i] * sm[(xa + i) & 7];
        for (size_t i = 0; i < n; ++i) wsum[i] = 1.0f, s0[i] = in[0][row + i], s1[i] = in[1][row + i], s2[i] = in[2][row + i];
        for (int k = 0; k < nk; ++k) {
          for (size_t i = 0; i < n; ++i) dd[i] = 0.0f;
          for (int c = 0; c < 3; ++c) {
            const float* pc = in[c] + row;
            const float scale = p.channel_scale[c];
            for (size_t i = 0; i < n; ++i) acc[i] = 0.0f;
            for (int t = 0; t < nd; ++t) {
              const float* pa = pc + koff[k] + doff[t];
              const float* pb = pc + doff[t];
              for (size_t i = 0; i < n; ++i) acc[i] += std::fabs(pa[i] - pb[i]);
            }
            for (size_t i = 0; i < n; ++i) dd[i] += scale * acc[i];
          }
          const float *q0 = in[0] + row + koff[k], *q1 = in[1] + row + koff[k], *q2 = in[2] + row + koff[k];
          for (size_t i = 0; i < n; ++i) {
            const float weight = std::max(1.0f + dd[i] * nis[i], 0.0f);
            wsum[i] += weight;
            s0[i] += weight * q0[i];
            s1[i] += weight * q1[i];
            s2[i] += weight * q2[i];
          }
        }
        for (size_t i = 0; i < n; ++i) {
          const bool keep = sg[i] < 0.3f;
          out[0][row + i] = keep ? in[0][row + i] : s0[i] / wsum[i];
          out[1][row + i] = keep ? in[1][row + i] : s1[i] / wsum[i];
          out[2][row + i] = keep ? in[2][row + i] : s2[i] / wsum[i];
        }
      }
      for (size_t dx = 0; dx < width; ++dx) {
        if (dx >= xa && dx < xb) {
          dx = xb - 1;
          continue;
        }
        float sigma_val = sigma_is_constant ? p.sigma_for_modular : sig[(y / 8) * sig_stride + dx / 8];
        if (sigma_val < 0.3f) {
          for (int c = 0; c < 3; ++c) out[c][y * width + dx] = in[c][y * width + dx];
          continue;
        }
        float sum_weights = 1.0f;
        float sum_channels[3] = {in[0][y * width + dx], in[1][y * width + dx], in[2][y * width + dx]};
        // weight() (impls/generic/epf.rs:205-210)
        const float neg_inv_sigma = 6.6f * (0.70710678118654752440f - 1.0f) / sigma_val * sm[dx & 7];
        if (row_inside && ptrdiff_t(dx) >= reach && ptrdiff_t(dx) + reach < ptrdiff_t(width)) {
          const size_t at = y * width + dx;
          for (int k = 0; k < nk; ++k) {
            float d = 0.0f;
            for (int c = 0; c < 3; ++c) {
              const float* pc = in[c] + at;
              float acc = 0.0f;
              for (int i = 0; i < nd; ++i) acc += std::fabs(pc[koff[k] + doff[i]] - pc[doff[i]]);
              d += p.channel_scale[c] * acc;
            }
            float weight = std::max(1.0f + d * neg_inv_sigma, 0.0f);
            sum_weights += weight;
            for (int c = 0; c < 3; ++c) sum_channels[c] += weight * in[c][at + koff[k]];
          }
        } else {
          for (int k = 0; k < nk; ++k) {
            ptrdiff_t kx = ptrdiff_t(dx) + kernel[k][0], ky = ptrdiff_t(y) + kernel[k][1];
            float d = 0.0f;
            for (int c = 0; c < 3; ++c) {
              float acc = 0.0f;
              for (int i = 0; i < nd; ++i) {
                size_t ay = mirror(ky + dist[i][1], height), ax = mirror(kx + dist[i][0], width);
                size_t by = mirror(ptrdiff_t(y) + dist[i][1], height), bx = mirror(ptrdiff_t(dx) + dist[i][0], width);
                acc += std::fabs(in[c][ay * width + ax] - in[c][by * width + bx]);
              }
              d += p.channel_scale[c] * acc;
            }
            float weight = std::max(1.0f + d * neg_inv_sigma, 0.0f);
            sum_weights += weight;
            size_t my = mirror(ky, height), mx = mirror(kx, width);
            for (int c = 0; c < 3; ++c) sum_channels[c] += weight * in[c][my * width + mx];
          }
        }
        for (int c = 0; c < 3; ++c) out[c][y * width + dx] = sum_channels[c] / sum_weights;
      }
    });
    cur ^= 1;
    lap("step");
  };
  if (p.iters == 3) run_step(std::integral_constant<int, 0>{});
  run_step(std::integral_constant<int, 1>{});
  if (p.iters >= 2) run_step(std::integral_constant<int, 2>{});
  for (int c = 0; c < 3; ++c) {
    Plane& pl = plane(v[c].plane);
    for (size_t y = 0; y < height; ++y)
      std::memcpy(pl.f32() + (v[c].y0 + y) * size_t(pl.w) + v[c].x0, &bufs[cur][c][y * width], width * 4);
  }
}

// Non-separable upsampling (crates/jxl-render/src/features/upsampling.rs:5-132): log2 factor 3 is
// one 8x pass, then at most one 2x / 4x pass; each pass is a 5x5 kernel per output phase built from
// the symmetric weight list, mirror padding of 2 (util.rs:423-454), clamped to the local min/max.
namespace {
std::vector<float> upsample_pass(const std::vector<float>& in, size_t gw, size_t gh, size_t k,
                                 const std::vector<float>& weights, OracleBackend* be) {
  const size_t pad = 2, pw = gw + 2 * pad, ph = gh + 2 * pad;
  std::vector<float> padded(pw * ph, 0.0f);
  for (size_t y = 0; y < gh; ++y) std::memcpy(&padded[(y + pad) * pw + pad], &in[y * gw], gw * 4);
  // mirror_edges_padding, literally
  for (size_t y = pad; y < gh + pad; ++y)
    for (size_t x = 0; x < pad; ++x) {
      padded[y * pw + x] = padded[y * pw + pad * 2 - x - 1];
      padded[(y + 1) * pw - x - 1] = padded[(y + 1) * pw - pad * 2 + x];
    }
  for (size_t r = 0; r < pad; ++r) {
    std::memcpy(&padded[r * pw], &padded[(2 * pad - 1 - r) * pw], pw * 4);
    std::memcpy(&padded[(gh + pad + r) * pw], &padded[(gh + pad - 1 - r) * pw], pw * 4);
  }
  const size_t mat_n = k / 2;
  std::vector<std::array<float, 25>> quarter(k * k / 4);
  for (auto& q : quarter) q.fill(0.0f);
  size_t weight_idx = 0;
  for (size_t y = 0; y < 5 * mat_n; ++y) {
    const size_t mat_y = y / 5, ky = y % 5;
    for (size_t x = y; x < 5 * mat_n; ++x) {
      const size_t mat_x = x / 5, kx = x % 5;
      const float w = weights[weight_idx++];
      quarter[mat_y * mat_n + mat_x][ky * 5 + kx] = w;
      quarter[mat_x * mat_n + mat_y][kx * 5 + ky] = w;
    }
  }
  const size_t fw = gw * k, fh = gh * k;
  std::vector<float> out(fw * fh);
  be->parallel_for(fh, [&](size_t y) {
    const size_t ref_y = y / k, mat_y = std::min(y % k, k - y % k - 1);
    const bool flip_v = y % k >= mat_n;
    for (size_t x = 0; x < fw; ++x) {
      const size_t ref_x = x / k, mat_x = std::min(x % k, k - x % k - 1);
      const bool flip_h = x % k >= mat_n;
      const std::array<float, 25>& kernel = quarter[mat_y * mat_n + mat_x];
      float sum = 0.0f, mn = INFINITY, mx = -INFINITY;
      for (size_t iy = 0; iy < 5; ++iy) {
        const size_t ky = flip_v ? 4 - iy : iy;
        for (size_t ix = 0; ix < 5; ++ix) {
          const size_t kx = flip_h ? 4 - ix : ix;
          const float sample = padded[(ref_y + iy) * pw + (ref_x + ix)];
          sum = sum + kernel[ky * 5 + kx] * sample;
          mn = std::fmin(mn, sample);  // f32::min / f32::max ignore NaN operands, like fmin / fmax
          mx = std::fmax(mx, sample);
        }
      }
      float r;
      if (!std::isfinite(mn)) r = NAN;
      else r = sum < mn ? mn : (sum > mx ? mx : sum);  // f32::clamp
      out[y * fw + x] = r;
    }
  });
  return out;
}
}  // namespace

int OracleBackend::upsample(const View& v, uint32_t factor_log2, const ImageHeader& ih) {
  Plane& src = plane(v.plane);
  size_t w = v.w, h = v.h;
  std::vector<float> cur(w * h);
  for (size_t y = 0; y < h; ++y) std::memcpy(&cur[y * w], src.f32() + (v.y0 + y) * size_t(src.w) + v.x0, w * 4);
  auto pass = [&](size_t k, const std::vector<float>& weights) {
    cur = upsample_pass(cur, w, h, k, weights, this);
    w *= k;
    h *= k;
  };
  for (uint32_t i = 0; i < factor_log2 / 3; ++i) pass(8, ih.up8_weight);
  if (factor_log2 % 3 == 1) pass(2, ih.up2_weight);
  if (factor_log2 % 3 == 2) pass(4, ih.up4_weight);
  int id = alloc_plane(uint32_t(w), uint32_t(h), false);
  std::memcpy(plane(id).f32(), cur.data(), cur.size() * 4);
  return id;
}

// ---------------------------------------------------------------------------------------------
// Patches: blend_single for the modes without alpha (crates/jxl-render/src/blend.rs:550-606)
void OracleBackend::blend_patches(const std::vector<PatchJob>& jobs) {
  auto clamp01 = [](float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); };  // f32::clamp keeps NaN
  for (const PatchJob& j : jobs) {
    Plane& sp = plane(j.src.plane);
    Plane& dp = plane(j.dst.plane);
    Plane* bap = j.base_alpha.plane >= 0 ? &plane(j.base_alpha.plane) : nullptr;
    Plane* nap = j.new_alpha.plane >= 0 ? &plane(j.new_alpha.plane) : nullptr;
    for (uint32_t y = 0; y < j.dst.h; ++y) {
      const float* s = sp.f32() + size_t(j.src.y0 + y) * sp.w + j.src.x0;
      float* d = dp.f32() + size_t(j.dst.y0 + y) * dp.w + j.dst.x0;
      const float* ba = bap ? bap->f32() + size_t(j.base_alpha.y0 + y) * bap->w + j.base_alpha.x0 : nullptr;
      const float* na = nap ? nap->f32() + size_t(j.new_alpha.y0 + y) * nap->w + j.new_alpha.x0 : nullptr;
      for (uint32_t x = 0; x < j.dst.w; ++x) {
        float v = s[x];
        switch (j.mode) {
          case 1: d[x] = v; break;
          case 2: d[x] = d[x] + v; break;
          case 3:
            if (j.clamp) v = clamp01(v);
            d[x] = d[x] * v;
            break;
          case 4: {  // blend.rs:607-660
            const float frame_alpha = ba ? ba[x] : 0.0f, patch_alpha = na ? na[x] : 0.0f;
            const float base_sample = j.swapped ? v : d[x], new_sample = j.swapped ? d[x] : v;
            const float base_alpha = j.swapped ? patch_alpha : frame_alpha;
            float new_alpha = j.swapped ? frame_alpha : patch_alpha;
            if (j.clamp) new_alpha = clamp01(new_alpha);
            if (j.premultiplied) {
              d[x] = new_sample + base_sample * (1.0f - new_alpha);
            } else {
              const float base_alpha_rev = 1.0f - base_alpha, new_alpha_rev = 1.0f - new_alpha;
              const float mixed_alpha = 1.0f - new_alpha_rev * base_alpha_rev;
              const float mixed_alpha_recip = mixed_alpha > 0.0f ? 1.0f / mixed_alpha : 0.0f;
              d[x] = (new_alpha * new_sample + base_alpha * base_sample * new_alpha_rev) * mixed_alpha_recip;
            }
            break;
          }
          case 5: {  // blend.rs:662-700
            const float base_sample = j.swapped ? v : d[x], new_sample = j.swapped ? d[x] : v;
            float new_alpha = j.swapped ? (ba ? ba[x] : 0.0f) : (na ? na[x] : 0.0f);
            if (j.clamp) new_alpha = clamp01(new_alpha);
            d[x] = base_sample + new_alpha * new_sample;
            break;
          }
          default: {  // 6, MixAlpha: blend.rs:702-723
            float base = j.swapped ? v : d[x], nw = j.swapped ? d[x] : v;
            if (j.clamp) nw = clamp01(nw);
            d[x] = base + nw * (1.0f - base);
            break;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Splines: the splat loop of render_spline and its erf (crates/jxl-render/src/features/spline.rs:218-252, 314-331)
namespace {
float spline_erf(float x) {
  const float ax = std::fabs(x);
  const float denom1 = ax * 7.77394369e-02f + 2.05260015e-04f;
  const float denom2 = denom1 * ax + 2.32120216e-01f;
  const float denom3 = denom2 * ax + 2.77820801e-01f;
  const float denom4 = denom3 * ax + 1.0f;
  const float denom5 = denom4 * denom4;
  const float inv_denom5 = 1.0f / denom5;
  const float result = -inv_denom5 * inv_denom5 + 1.0f;
  return x < 0.0f ? -result : result;
}
}  // namespace

void OracleBackend::splat_splines(const View v[3], const std::vector<SplineArc>& arcs) {
  for (const SplineArc& a : arcs)
    for (int c = 0; c < 3; ++c) {
      Plane& p = plane(v[c].plane);
      for (int32_t y = a.ybegin; y < a.yend; ++y) {
        if (uint32_t(y) >= v[c].h) break;
        float* row = p.f32() + size_t(v[c].y0 + y) * p.w + v[c].x0;
        for (int32_t x = a.xbegin; x < a.xend; ++x) {
          if (uint32_t(x) >= v[c].w) break;
          const float dx = float(x) - a.x, dy = float(y) - a.y;
          const float distance = std::sqrt(dx * dx + dy * dy);
          const float factor = spline_erf((0.5f * distance + 0.35355338f) * a.inv_sigma) -
                               spline_erf((0.5f * distance - 0.35355338f) * a.inv_sigma);
          const float extra = 0.25f * a.value[c] * a.sigma * factor * factor;
          row[x] = row[x] + extra;
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------
// Noise synthesis (crates/jxl-render/src/features/noise.rs)
namespace {
uint64_t split_mix_64(uint64_t z) {  // noise.rs:454-458
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
}  // namespace

void OracleBackend::add_noise(const View v[3], const float lut8[8], uint32_t group_dim, uint64_t seed0, float corr_x,
                              float corr_b) {
  const size_t width = v[0].w, height = v[0].h;
  JXLB_CHECK(width >= 2 && height >= 2, kErrUnsupported, "noise on frames narrower than 2 samples is not supported");
  const size_t gpr = (width + group_dim - 1) / group_dim, gpc = (height + group_dim - 1) / group_dim;
  // 1. the random field, group by group, channel after channel from one generator (noise.rs:199-235)
  std::vector<float> field[3];
  for (auto& f : field) f.assign(width * height, 0.0f);
  for (size_t gy = 0; gy < gpc; ++gy)
    for (size_t gx = 0; gx < gpr; ++gx) {
      const size_t x0 = gx * group_dim, y0 = gy * group_dim;
      const size_t gw = std::min<size_t>(group_dim, width - x0), gh = std::min<size_t>(group_dim, height - y0);
      const uint64_t seed1 = (uint64_t(x0) << 32) + uint64_t(y0);
      uint64_t s0[8], s1[8];
      s0[0] = split_mix_64(seed0 + 0x9E3779B97F4A7C15ull);
      s1[0] = split_mix_64(seed1 + 0x9E3779B97F4A7C15ull);
      for (int i = 1; i < 8; ++i) {
        s0[i] = split_mix_64(s0[i - 1]);
        s1[i] = split_mix_64(s1[i - 1]);
      }
      const size_t width_n2 = (gw + 15) / 16;
      for (int c = 0; c < 3; ++c)
        for (size_t it = 0; it < width_n2 * gh; ++it) {
          const size_t row = it / width_n2, colb = (it % width_n2) * 16;
          for (int i = 0; i < 8; ++i) {  // XorShift128Plus::fill_batch (noise.rs:439-451)
            uint64_t a = s0[i];
            const uint64_t b = s1[i];
            const uint64_t ret = a + b;
            s0[i] = b;
            a ^= a << 23;
            s1[i] = a ^ (b ^ (a >> 18) ^ (b >> 5));
            const uint32_t bits[2] = {uint32_t(ret), uint32_t(ret >> 32)};
            for (int k = 0; k < 2; ++k) {
              const size_t x = colb + size_t(2 * i + k);
              if (x >= gw) continue;
              const uint32_t fb = (bits[k] >> 9) | 0x3f800000u;
              float fv;
              std::memcpy(&fv, &fb, 4);
              field[c][(y0 + row) * width + x0 + x] = fv;
            }
          }
        }
    }
  auto mirror = [](ptrdiff_t p, ptrdiff_t n) { return p < 0 ? -p - 1 : (p >= n ? 2 * n - p - 1 : p); };
  float lut[9];
  for (int i = 0; i < 8; ++i) lut[i] = lut8[i];
  lut[8] = lut8[7];
  Plane* pl[3] = {&plane(v[0].plane), &plane(v[1].plane), &plane(v[2].plane)};
  // 2 + 3. 5x5 high-pass (rows summed in the order of the reference's 5-row ring buffer, which depends on
  // the row's position inside its group: noise.rs:297-320) and application (noise.rs:45-83)
  parallel_for(height, [&](size_t y) {
    const ptrdiff_t ly = ptrdiff_t(y % group_dim);
    float* rx = pl[0]->f32() + (v[0].y0 + y) * size_t(pl[0]->w) + v[0].x0;
    float* ry = pl[1]->f32() + (v[1].y0 + y) * size_t(pl[1]->w) + v[1].x0;
    float* rb = pl[2]->f32() + (v[2].y0 + y) * size_t(pl[2]->w) + v[2].x0;
    for (size_t x = 0; x < width; ++x) {
      float n[3];
      for (int c = 0; c < 3; ++c) {
        float sum = 0.0f;
        for (ptrdiff_t s = 0; s < 5; ++s) {
          const ptrdiff_t r = ly - 2 + (((s - ly) % 5) + 5) % 5;  // local row held by ring slot s
          const ptrdiff_t sy = mirror(ptrdiff_t(y) - ly + r, ptrdiff_t(height));
          for (ptrdiff_t dx = 0; dx < 5; ++dx) {
            const ptrdiff_t sx = mirror(ptrdiff_t(x) + dx - 2, ptrdiff_t(width));
            sum = sum + field[c][size_t(sy) * width + size_t(sx)] * 0.16f;
          }
        }
        n[c] = sum - field[c][y * width + x] * 4.0f;
      }
      const float grid_x = rx[x], grid_y = ry[x];
      const float in_x = grid_x + grid_y, in_y = grid_y - grid_x;
      const float in_scaled_x = std::fmax(0.0f, in_x * 3.0f), in_scaled_y = std::fmax(0.0f, in_y * 3.0f);
      const size_t in_x_int = std::min<size_t>(size_t(in_scaled_x), 7), in_y_int = std::min<size_t>(size_t(in_scaled_y), 7);
      const float in_x_frac = in_scaled_x - float(in_x_int), in_y_frac = in_scaled_y - float(in_y_int);
      const float sx = (lut[in_x_int + 1] - lut[in_x_int]) * in_x_frac + lut[in_x_int];
      const float sy = (lut[in_y_int + 1] - lut[in_y_int]) * in_y_frac + lut[in_y_int];
      const float nx = 0.22f * sx * (0.0078125f * n[0] + 0.9921875f * n[2]);
      const float ny = 0.22f * sy * (0.0078125f * n[1] + 0.9921875f * n[2]);
      rx[x] = rx[x] + (corr_x * (nx + ny) + nx - ny);
      ry[x] = ry[x] + (nx + ny);
      rb[x] = rb[x] + corr_b * (nx + ny);
    }
  });
}

// ---------------------------------------------------------------------------------------------
// XYB -> linear sRGB (-> sRGB), jxl-color/src/{xyb.rs:35-60, ciexyz.rs:81-87, tf/srgb.rs:13-48}
// apply_jpeg_upsampling_single (jxl-render/src/filter/ycbcr.rs:6-78): horizontal pass into the output rows, then the
// vertical pass over those rows (the reference runs it in place bottom-to-top; the values read are the horizontal
// pass's, so a second buffer gives the same result).
int OracleBackend::upsample_jpeg(const View& v, bool horizontal, bool vertical, uint32_t out_w, uint32_t out_h) {
  Plane& src = plane(v.plane);
  std::vector<float> tmp(size_t(out_w) * v.h);
  for (uint32_t y = 0; y < v.h; ++y) {
    const float* row = src.f32() + size_t(v.y0 + y) * src.w + v.x0;
    float* o = tmp.data() + size_t(y) * out_w;
    if (!horizontal) {
      for (uint32_t x = 0; x < out_w; ++x) o[x] = row[x];
      continue;
    }
    for (uint32_t i = 0; i < v.w; ++i) {
      const float prev = row[i ? i - 1 : 0], cur = row[i], next = row[i + 1 < v.w ? i + 1 : v.w - 1];
      if (2 * i < out_w) o[2 * i] = 0.25f * prev + 0.75f * cur;
      if (2 * i + 1 < out_w) o[2 * i + 1] = 0.75f * cur + 0.25f * next;
    }
  }
  const int id = alloc_plane(out_w, out_h, false);
  Plane& dst = plane(id);
  for (uint32_t y = 0; y < v.h; ++y) {
    const float* cur = tmp.data() + size_t(y) * out_w;
    if (!vertical) {
      if (y < out_h) std::memcpy(dst.f32() + size_t(y) * dst.w, cur, size_t(out_w) * 4);
      continue;
    }
    const float* above = tmp.data() + size_t(y ? y - 1 : 0) * out_w;
    const float* below = tmp.data() + size_t(y + 1 < v.h ? y + 1 : v.h - 1) * out_w;
    for (uint32_t x = 0; x < out_w; ++x) {
      if (2 * y < out_h) dst.f32()[size_t(2 * y) * dst.w + x] = 0.75f * cur[x] + 0.25f * above[x];
      if (2 * y + 1 < out_h) dst.f32()[size_t(2 * y + 1) * dst.w + x] = 0.25f * below[x] + 0.75f * cur[x];
    }
  }
  return id;
}

// jxl-color/src/ycbcr.rs:40-56 (mul_add = fused)
void OracleBackend::ycbcr_to_rgb(const View v[3], const YcbcrParams& p) {
  Plane* pl[3] = {&plane(v[0].plane), &plane(v[1].plane), &plane(v[2].plane)};
  parallel_for(v[0].h, [&](size_t y) {
    float* r[3];
    for (int c = 0; c < 3; ++c) r[c] = pl[c]->f32() + (v[c].y0 + y) * size_t(pl[c]->w) + v[c].x0;
    for (uint32_t x = 0; x < v[0].w; ++x) {
      const float cb = r[0][x], yy = r[1][x] + p.y_offset, cr = r[2][x];
      r[0][x] = std::fmaf(cr, p.cr_to_r, yy);
      r[1][x] = std::fmaf(cb, p.cb_to_g, std::fmaf(cr, p.cr_to_g, yy));
      r[2][x] = std::fmaf(cb, p.cb_to_b, yy);
    }
  });
}

// linear_to_pq_generic (tf/pq.rs:126-142) with rational_poly::eval_generic (Horner from the last coefficient)
float linear_to_pq(float s, float intensity_target) {
  static const float kP[5] = {1.351392e-2f, -1.095778f, 5.522776e1f, 1.492516e2f, 4.838434e1f};
  static const float kQ[5] = {1.012416f, 2.016708e1f, 9.26371e1f, 1.120607e2f, 2.590418e1f};
  static const float kPSmall[5] = {9.863406e-6f, 3.881234e-1f, 1.352821e2f, 6.889862e4f, -2.864824e5f};
  static const float kQSmall[5] = {3.371868e1f, 1.477719e3f, 1.608477e4f, -4.389884e4f, -2.072546e5f};
  const float y_mult = intensity_target / 10000.0f;
  const float a0 = std::fabs(s);
  const float a_1_4 = std::sqrt(std::sqrt(a0 * y_mult));
  const float* pp = a0 < 1e-4f ? kPSmall : kP;
  const float* qq = a0 < 1e-4f ? kQSmall : kQ;
  float yp = pp[4], yq = qq[4];
  for (int k = 3; k >= 0; --k) yp = yp * a_1_4 + pp[k], yq = yq * a_1_4 + qq[k];
  return std::copysign(yp / yq, s);
}

void OracleBackend::xyb_to_rgb(const View v[3], const ColorParams& p) {
  static const uint8_t kPowUpper[16] = {0x00, 0x0a, 0x19, 0x26, 0x32, 0x41, 0x4d, 0x5c, 0x68, 0x75, 0x83, 0x8f, 0xa0, 0xaa, 0xb9, 0xc6};
  static const uint8_t kPowLower[16] = {0x00, 0xb7, 0x04, 0x0d, 0xcb, 0xe7, 0x41, 0x68, 0x51, 0xd1, 0xeb, 0xf2, 0x00, 0xb7, 0x04, 0x0d};
  Plane* pl[3] = {&plane(v[0].plane), &plane(v[1].plane), &plane(v[2].plane)};
  parallel_for(v[0].h, [&](size_t y) {
    float* r[3];
    for (int c = 0; c < 3; ++c) r[c] = pl[c]->f32() + (v[c].y0 + y) * size_t(pl[c]->w) + v[c].x0;
    for (uint32_t x = 0; x < v[0].w; ++x) {
      float xx = r[0][x], yy = r[1][x], bb = r[2][x];
      float g_l = yy + xx, g_m = yy - xx, g_s = bb;
      g_l = g_l - p.cbrt_opsin_bias[0];
      g_m = g_m - p.cbrt_opsin_bias[1];
      g_s = g_s - p.cbrt_opsin_bias[2];
      float a = std::fmaf(g_l * g_l, g_l, p.opsin_bias[0]) * p.itscale;
      float b = std::fmaf(g_m * g_m, g_m, p.opsin_bias[1]) * p.itscale;
      float c = std::fmaf(g_s * g_s, g_s, p.opsin_bias[2]) * p.itscale;
      const float* m = p.matrix;
      float o[3] = {m[0] * a + m[1] * b + m[2] * c, m[3] * a + m[4] * b + m[5] * c, m[6] * a + m[7] * b + m[8] * c};
      if (p.second_stage) {
        {  // map_gamut_generic (jxl-color/src/gamut.rs:4-46), saturation factor 0.3
          const float yl = o[0] * p.luminances[0] + o[1] * p.luminances[1] + o[2] * p.luminances[2];
          float gray_saturation = 0.0f, gray_luminance = 0.0f;
          for (float vv : o) {
            const float v_sub_y = vv - yl;
            const float inv = 1.0f / (v_sub_y == 0.0f ? 1.0f : v_sub_y);
            const float v_over = vv * inv;
            if (!(v_sub_y >= 0.0f)) gray_saturation = std::fmax(gray_saturation, v_over);
            gray_luminance = std::fmax(v_sub_y <= 0.0f ? gray_saturation : v_over - inv, gray_luminance);
          }
          float gray_mix = 0.3f * (gray_saturation - gray_luminance) + gray_luminance;
          gray_mix = gray_mix < 0.0f ? 0.0f : (gray_mix > 1.0f ? 1.0f : gray_mix);  // f32::clamp (NaN passes through)
          float max_colour = 1.0f;
          for (float vv : o) max_colour = std::fmax(vv, max_colour);
          for (float& vv : o) vv = (gray_mix * (yl - vv) + vv) / max_colour;
        }
        const float* m2 = p.matrix2;
        const float t0 = m2[0] * o[0] + m2[1] * o[1] + m2[2] * o[2], t1 = m2[3] * o[0] + m2[4] * o[1] + m2[5] * o[2],
                    t2 = m2[6] * o[0] + m2[7] * o[1] + m2[8] * o[2];
        o[0] = p.to_luma ? t1 : t0, o[1] = t1, o[2] = t2;
      }
      if (p.pq_intensity_target > 0.0f) {
        for (float& sref : o) sref = linear_to_pq(sref, p.pq_intensity_target);
      } else if (p.gamma > 0.0f) {  // apply_gamma (tf.rs:62-69) with fast_powf_generic (fastmath/powf.rs)
        for (float& sref : o) {
          const float a0 = sref;
          if (a0 <= 1e-7f) {
            sref = 0.0f;
            continue;
          }
          int32_t x_bits;
          std::memcpy(&x_bits, &a0, 4);
          const int32_t exp_shifted = (x_bits - 0x3f2aaaab) >> 23;
          const int32_t mb = x_bits - exp_shifted * (1 << 23);
          float mantissa;
          std::memcpy(&mantissa, &mb, 4);
          const float xx2 = mantissa - 1.0f;
          const float yp = (7.4245873327820566e-1f * xx2 + 1.4287160470083755f) * xx2 + -1.8503833400518310e-6f;
          const float yq = (1.7409343003366853e-1f * xx2 + 1.0096718572241148f) * xx2 + 9.9032814277590719e-1f;
          const float e = (yp / yq + float(exp_shifted)) * p.gamma;
          const float x_floor = std::floor(e);
          // `x_floor as i32` saturates (NaN -> 0); the sum wraps like release-mode Rust
          const int32_t xi = x_floor != x_floor ? 0 : (x_floor >= 2147483648.0f ? INT32_MAX : (x_floor <= -2147483648.0f ? INT32_MIN : int32_t(x_floor)));
          const uint32_t eb = (uint32_t(xi) + 127u) << 23;
          float ex;
          std::memcpy(&ex, &eb, 4);
          const float frac = e - x_floor;
          float num = frac + 1.01749063e1f;
          num = num * frac + 4.88687798e1f;
          num = num * frac + 9.85506591e1f;
          num = num * ex;
          float den = 2.10242958e-1f * frac + -2.22328856e-2f;
          den = den * frac + -1.94414990e1f;
          den = den * frac + 9.85506633e1f;
          sref = num / den;
        }
      } else if (p.apply_srgb_tf) {
        for (float& s : o) {
          uint32_t bits;
          std::memcpy(&bits, &s, 4);
          uint32_t vb = bits & 0x7fffffffu;
          uint32_t adj = (vb | 0x3e800000u) & 0x3effffffu;
          float v_adj;
          std::memcpy(&v_adj, &adj, 4);
          float pow = 0.059914046f;
          pow = pow * v_adj - 0.10889456f;
          pow = pow * v_adj + 0.107963754f;
          pow = pow * v_adj + 0.018092343f;
          uint32_t idx = ((vb >> 23) - 118) & 0xf;
          uint32_t mulb = 0x40000000u | (uint32_t(kPowUpper[idx]) << 18) | (uint32_t(kPowLower[idx]) << 10);
          float mul, av;
          std::memcpy(&mul, &mulb, 4);
          std::memcpy(&av, &vb, 4);
          float small = av * 12.92f;
          float acc = pow * mul - 0.055f;
          float res = av <= 0.0031308f ? small : acc;
          s = std::copysign(res, s);
        }
      } else if (p.apply_bt709_tf) {
        // linear_to_bt709 (jxl-color/src/tf/bt709.rs:61-68) with fast_powf_generic
        // (fastmath/powf.rs:7-22, 147-156, 242-244; rational_poly.rs:2-6)
        for (float& sref : o) {
          const float a = sref;
          if (a <= 0.018f) {
            sref = 4.5f * a;
            continue;
          }
          int32_t x_bits;
          std::memcpy(&x_bits, &a, 4);
          const int32_t exp_shifted = (x_bits - 0x3f2aaaab) >> 23;
          const int32_t mb = x_bits - exp_shifted * (1 << 23);
          float mantissa;
          std::memcpy(&mantissa, &mb, 4);
          const float exp_val = float(exp_shifted);
          const float xx = mantissa - 1.0f;
          const float yp = (7.4245873327820566e-1f * xx + 1.4287160470083755f) * xx + -1.8503833400518310e-6f;
          const float yq = (1.7409343003366853e-1f * xx + 1.0096718572241148f) * xx + 9.9032814277590719e-1f;
          const float l2 = yp / yq + exp_val;
          const float e = l2 * 0.45f;
          const float x_floor = std::floor(e);
          const uint32_t eb = uint32_t(int32_t(x_floor) + 127) << 23;
          float ex;
          std::memcpy(&ex, &eb, 4);
          const float frac = e - x_floor;
          float num = frac + 1.01749063e1f;
          num = num * frac + 4.88687798e1f;
          num = num * frac + 9.85506591e1f;
          num = num * ex;
          float den = 2.10242958e-1f * frac + -2.22328856e-2f;
          den = den * frac + -1.94414990e1f;
          den = den * frac + 9.85506633e1f;
          sref = std::fmaf(num / den, 1.099f, -0.099f);
        }
      }
      r[0][x] = o[0];
      r[1][x] = o[1];
      r[2][x] = o[2];
    }
  });
}

}  // namespace jxlo
