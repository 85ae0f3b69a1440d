// Restoration filters and colour on the device: Gaborish 3x3, edge-preserving filter (steps
// 0/1/2), XYB -> linear sRGB (-> sRGB). Float op order follows the reference's generic path:
// crates/jxl-render/src/filter/{gabor.rs,epf.rs}, filter/impls/generic/{gabor.rs,epf.rs},
// crates/jxl-color/src/{xyb.rs:35-60, ciexyz.rs:81-87, tf/srgb.rs:13-48}. Compiled with
// -fmad=false; fused multiply-add only where the reference uses mul_add.
#include "kernels.h"

namespace jxlb {

namespace {

__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// impls/generic/gabor.rs: the formulas differ between interior rows, top/bottom rows, first/last
// columns and the degenerate 1-row / 1-column cases; each is reproduced literally.
__global__ void gaborish_kernel(DevView in, DevView out, float w0, float w1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  const int width = int(in.w), height = int(in.h);
  if (x >= width) return;
  const float* base = static_cast<const float*>(in.ptr);
  const float gw = fdiv(1.0f, fadd(fadd(1.0f, fmul(w0, 4.0f)), fmul(w1, 4.0f)));
  auto at = [&](int xx, int yy) { return base[size_t(yy) * in.stride + xx]; };
  float res;
  if (height == 1) {
    if (width == 1) {
      res = at(0, 0);
    } else {
      const float merged_w0 = fadd(fadd(1.0f, 2.0f), w0);
      const float merged_w1 = fadd(w0, fmul(2.0f, w1));
      if (x == 0) res = fmul(fadd(fmul(at(0, 0), fadd(merged_w0, merged_w1)), fmul(at(1, 0), merged_w1)), gw);
      else if (x == width - 1)
        res = fmul(fadd(fmul(at(width - 1, 0), fadd(merged_w0, merged_w1)), fmul(at(width - 2, 0), merged_w1)), gw);
      else res = fmul(fadd(fmul(at(x, 0), merged_w0), fmul(fadd(at(x - 1, 0), at(x + 1, 0)), merged_w1)), gw);
    }
  } else if (y == 0 || y == height - 1) {
    const int ya = (y == 0) ? 1 : height - 2;  // the one adjacent row
    if (width == 1) {
      float u = at(0, ya), c = at(0, y);
      res = fmul(fadd(fmul(c, fadd(fadd(1.0f, fmul(3.0f, w0)), fmul(2.0f, w1))), fmul(u, fadd(w0, fmul(2.0f, w1)))), gw);
    } else if (x == 0 || x == width - 1) {
      const int xo = (x == 0) ? 1 : width - 2;
      float a1 = at(x, ya), a0 = at(xo, ya), c1 = at(x, y), c0 = at(xo, y);
      res = fmul(fadd(fadd(fmul(c1, fadd(fadd(1.0f, fmul(2.0f, w0)), w1)), fmul(fadd(a1, c0), fadd(w0, w1))), fmul(a0, w1)), gw);
    } else {
      float a0 = at(x - 1, ya), a1 = at(x, ya), a2 = at(x + 1, ya);
      float c0 = at(x - 1, y), c1 = at(x, y), c2 = at(x + 1, y);
      res = fmul(fadd(fadd(c1, fmul(fadd(fadd(fadd(a1, c0), c1), c2), w0)), fmul(fadd(fadd(fadd(a0, a2), c0), c2), w1)), gw);
    }
  } else {
    if (width == 1) {
      float t = at(0, y - 1), c = at(0, y), b = at(0, y + 1);
      float sum_side = fadd(fadd(t, fmul(2.0f, c)), b);
      float sum_diag = fmul(2.0f, fadd(t, b));
      res = fmul(fadd(fadd(c, fmul(sum_side, w0)), fmul(sum_diag, w1)), gw);
    } else if (x == 0 || x == width - 1) {
      const int xo = (x == 0) ? 1 : width - 2;
      float t1 = at(x, y - 1), c1 = at(x, y), b1 = at(x, y + 1);
      float t0 = at(xo, y - 1), c0 = at(xo, y), b0 = at(xo, y + 1);
      float sum_side = fadd(fadd(fadd(t1, c0), c1), b1);
      float sum_diag = fadd(fadd(fadd(t0, t1), b0), b1);
      res = fmul(fadd(fadd(c1, fmul(sum_side, w0)), fmul(sum_diag, w1)), gw);
    } else {
      float sum_side = fadd(fadd(fadd(at(x, y - 1), at(x - 1, y)), at(x + 1, y)), at(x, y + 1));
      float sum_diag = fadd(fadd(fadd(at(x - 1, y - 1), at(x + 1, y - 1)), at(x - 1, y + 1)), at(x + 1, y + 1));
      res = fmul(fadd(fadd(at(x, y), fmul(sum_side, w0)), fmul(sum_diag, w1)), gw);
    }
  }
  static_cast<float*>(out.ptr)[size_t(y) * out.stride + x] = res;
}

__device__ __forceinline__ int mirror(int offset, int len) {  // util.rs:376-386
  for (;;) {
    if (offset < 0) offset = -(offset + 1);
    else if (offset >= len) offset = 2 * len - (offset + 1);
    else return offset;
  }
}

__device__ __constant__ const int8_t kKernel1[4][2] = {{0, -1}, {0, 1}, {-1, 0}, {1, 0}};
__device__ __constant__ const int8_t kKernel2[12][2] = {{0, -2}, {-1, -1}, {0, -1}, {1, -1}, {-2, 0}, {-1, 0},
                                                        {1, 0},  {2, 0},   {-1, 1}, {0, 1},  {1, 1},  {0, 2}};
__device__ __constant__ const int8_t kDist0[5][2] = {{0, -1}, {1, 0}, {0, 0}, {-1, 0}, {0, 1}};
__device__ __constant__ const int8_t kDist1[5][2] = {{0, -1}, {0, 0}, {0, 1}, {-1, 0}, {1, 0}};

struct View3 {
  DevView v[3];
};

// epf_row<STEP> (impls/generic/epf.rs:3-204): one thread per pixel, all three channels.
template <int STEP>
__global__ void epf_kernel(View3 in, View3 out, const float* __restrict__ sigma, uint32_t sigma_stride, DevEpfParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int width = int(in.v[0].w), height = int(in.v[0].h);
  if (x >= width || y >= height) return;
  const float* ch[3] = {static_cast<const float*>(in.v[0].ptr), static_cast<const float*>(in.v[1].ptr),
                        static_cast<const float*>(in.v[2].ptr)};
  const size_t stride[3] = {in.v[0].stride, in.v[1].stride, in.v[2].stride};
  const float sigma_val = sigma ? sigma[size_t(y >> 3) * sigma_stride + (x >> 3)] : p.sigma_for_modular;
  float o[3];
  if (sigma_val < 0.3f) {
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = ch[c][size_t(y) * stride[c] + x];
  } else {
    const float step_multiplier = STEP == 0 ? p.pass0_sigma_scale : (STEP == 2 ? p.pass2_sigma_scale : 1.0f);
    const bool is_y_border = ((y + 1) & 6) == 0;
    float sm;
    if (is_y_border) sm = fmul(step_multiplier, p.border_sad_mul);
    else sm = ((x & 7) == 0 || (x & 7) == 7) ? fmul(step_multiplier, p.border_sad_mul) : step_multiplier;
    const float neg_inv_sigma = fmul(fdiv(fmul(6.6f, fsub(0.70710678118654752440f, 1.0f)), sigma_val), sm);
    float sum_weights = 1.0f;
    float sum_channels[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sum_channels[c] = ch[c][size_t(y) * stride[c] + x];
    constexpr int NK = STEP == 0 ? 12 : 4;
    constexpr int ND = STEP == 2 ? 1 : 5;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      const int kx = x + (STEP == 0 ? kKernel2[k][0] : kKernel1[k][0]);
      const int ky = y + (STEP == 0 ? kKernel2[k][1] : kKernel1[k][1]);
      float dist = 0.0f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = 0.0f;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
          const int ox = STEP == 2 ? 0 : (STEP == 0 ? kDist0[i][0] : kDist1[i][0]);
          const int oy = STEP == 2 ? 0 : (STEP == 0 ? kDist0[i][1] : kDist1[i][1]);
          const int ay = mirror(ky + oy, height), ax = mirror(kx + ox, width);
          const int by = mirror(y + oy, height), bx = mirror(x + ox, width);
          acc = fadd(acc, fabsf(fsub(ch[c][size_t(ay) * stride[c] + ax], ch[c][size_t(by) * stride[c] + bx])));
        }
        dist = fadd(dist, fmul(p.channel_scale[c], acc));
      }
      const float weight = fmaxf(fadd(1.0f, fmul(dist, neg_inv_sigma)), 0.0f);
      sum_weights = fadd(sum_weights, weight);
      const int my = mirror(ky, height), mx = mirror(kx, width);
#pragma unroll
      for (int c = 0; c < 3; ++c) sum_channels[c] = fadd(sum_channels[c], fmul(weight, ch[c][size_t(my) * stride[c] + mx]));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = fdiv(sum_channels[c], sum_weights);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) static_cast<float*>(out.v[c].ptr)[size_t(y) * out.v[c].stride + x] = o[c];
}

__device__ __constant__ const uint8_t kPowUpper[16] = {0x00, 0x0a, 0x19, 0x26, 0x32, 0x41, 0x4d, 0x5c,
                                                       0x68, 0x75, 0x83, 0x8f, 0xa0, 0xaa, 0xb9, 0xc6};
__device__ __constant__ const uint8_t kPowLower[16] = {0x00, 0xb7, 0x04, 0x0d, 0xcb, 0xe7, 0x41, 0x68,
                                                       0x51, 0xd1, 0xeb, 0xf2, 0x00, 0xb7, 0x04, 0x0d};

__device__ __forceinline__ float linear_to_srgb(float s) {  // tf/srgb.rs:28-47 (scalar path)
  uint32_t bits = __float_as_uint(s);
  uint32_t vb = bits & 0x7fffffffu;
  float v_adj = __uint_as_float((vb | 0x3e800000u) & 0x3effffffu);
  float pow = 0.059914046f;
  pow = fsub(fmul(pow, v_adj), 0.10889456f);
  pow = fadd(fmul(pow, v_adj), 0.107963754f);
  pow = fadd(fmul(pow, v_adj), 0.018092343f);
  uint32_t idx = ((vb >> 23) - 118) & 0xf;
  float mul = __uint_as_float(0x40000000u | (uint32_t(kPowUpper[idx]) << 18) | (uint32_t(kPowLower[idx]) << 10));
  float av = __uint_as_float(vb);
  float small = fmul(av, 12.92f);
  float acc = fsub(fmul(pow, mul), 0.055f);
  float res = av <= 0.0031308f ? small : acc;
  return copysignf(res, s);
}


// BT.709 OETF exactly as the reference's generic path evaluates it (jxl-color/src/tf/bt709.rs:61-68 with
// fastmath/powf.rs:7-22, 147-156 and rational_poly.rs:2-6): rational-polynomial log2 / pow2, un-fused
// except for the final mul_add.
__device__ __forceinline__ float linear_to_bt709(float a) {
  if (a <= 0.018f) return fmul(4.5f, a);
  const int32_t x_bits = __float_as_int(a);
  const int32_t exp_shifted = (x_bits - 0x3f2aaaab) >> 23;
  const float mantissa = __int_as_float(x_bits - (exp_shifted << 23));
  const float exp_val = float(exp_shifted);
  const float x = fsub(mantissa, 1.0f);
  const float yp = fadd(fmul(fadd(fmul(7.4245873327820566e-1f, x), 1.4287160470083755f), x), -1.8503833400518310e-6f);
  const float yq = fadd(fmul(fadd(fmul(1.7409343003366853e-1f, x), 1.0096718572241148f), x), 9.9032814277590719e-1f);
  const float l2 = fadd(fdiv(yp, yq), exp_val);
  const float e = fmul(l2, 0.45f);
  const float x_floor = floorf(e);
  const float ex = __int_as_float(int32_t(uint32_t(int32_t(x_floor) + 127) << 23));
  const float frac = fsub(e, x_floor);
  float num = fadd(frac, 1.01749063e1f);
  num = fadd(fmul(num, frac), 4.88687798e1f);
  num = fadd(fmul(num, frac), 9.85506591e1f);
  num = fmul(num, ex);
  float den = fadd(fmul(2.10242958e-1f, frac), -2.22328856e-2f);
  den = fadd(fmul(den, frac), -1.94414990e1f);
  den = fadd(fmul(den, frac), 9.85506633e1f);
  return __fmaf_rn(fdiv(num, den), 1.099f, -0.099f);
}

// apply_jpeg_upsampling_single (jxl-render/src/filter/ycbcr.rs:6-78): the horizontal pass, then the vertical pass over
// its output, both computed per output sample (edges replicate).
__device__ __forceinline__ float jpeg_hsample(const float* row, int x, int in_w, int horizontal) {
  if (!horizontal) return row[x];
  const int i = x >> 1;
  const float cur = row[i];
  if (x & 1) return fadd(fmul(0.75f, cur), fmul(0.25f, row[min(i + 1, in_w - 1)]));
  return fadd(fmul(0.25f, row[max(i - 1, 0)]), fmul(0.75f, cur));
}

__global__ void upsample_jpeg_kernel(DevView in, DevView out, int horizontal, int vertical) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= int(out.w)) return;
  const float* src = static_cast<const float*>(in.ptr);
  const int in_w = int(in.w), in_h = int(in.h);
  float v;
  if (!vertical) {
    v = jpeg_hsample(src + size_t(y) * in.stride, x, in_w, horizontal);
  } else {
    const int r = y >> 1;
    const float cur = jpeg_hsample(src + size_t(r) * in.stride, x, in_w, horizontal);
    if (y & 1) {
      const float below = jpeg_hsample(src + size_t(min(r + 1, in_h - 1)) * in.stride, x, in_w, horizontal);
      v = fadd(fmul(0.25f, below), fmul(0.75f, cur));
    } else {
      const float above = jpeg_hsample(src + size_t(max(r - 1, 0)) * in.stride, x, in_w, horizontal);
      v = fadd(fmul(0.75f, cur), fmul(0.25f, above));
    }
  }
  static_cast<float*>(out.ptr)[size_t(y) * out.stride + x] = v;
}

__global__ void ycbcr_to_rgb_kernel(DevView vcb, DevView vy, DevView vcr, DevYcbcrParams p) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= vy.w) return;
  float* pcb = static_cast<float*>(vcb.ptr) + size_t(y) * vcb.stride + x;
  float* py = static_cast<float*>(vy.ptr) + size_t(y) * vy.stride + x;
  float* pcr = static_cast<float*>(vcr.ptr) + size_t(y) * vcr.stride + x;
  const float cb = *pcb, yy = fadd(*py, p.y_offset), cr = *pcr;
  *pcb = __fmaf_rn(cr, p.cr_to_r, yy);
  *py = __fmaf_rn(cb, p.cb_to_g, __fmaf_rn(cr, p.cr_to_g, yy));
  *pcr = __fmaf_rn(cb, p.cb_to_b, yy);
}

// linear_to_pq_generic (jxl-color/src/tf/pq.rs:126-142): fourth root, then a 4/4 rational polynomial (Horner, un-fused)
__device__ __forceinline__ float pq_tf(float s, float y_mult) {
  const float a = fabsf(s);
  const float a_1_4 = __fsqrt_rn(__fsqrt_rn(fmul(a, y_mult)));
  float yp, yq;
  if (a < 1e-4f) {
    yp = fadd(fmul(fadd(fmul(fadd(fmul(fadd(fmul(-2.864824e5f, a_1_4), 6.889862e4f), a_1_4), 1.352821e2f), a_1_4), 3.881234e-1f), a_1_4), 9.863406e-6f);
    yq = fadd(fmul(fadd(fmul(fadd(fmul(fadd(fmul(-2.072546e5f, a_1_4), -4.389884e4f), a_1_4), 1.608477e4f), a_1_4), 1.477719e3f), a_1_4), 3.371868e1f);
  } else {
    yp = fadd(fmul(fadd(fmul(fadd(fmul(fadd(fmul(4.838434e1f, a_1_4), 1.492516e2f), a_1_4), 5.522776e1f), a_1_4), -1.095778f), a_1_4), 1.351392e-2f);
    yq = fadd(fmul(fadd(fmul(fadd(fmul(fadd(fmul(2.590418e1f, a_1_4), 1.120607e2f), a_1_4), 9.26371e1f), a_1_4), 2.016708e1f), a_1_4), 1.012416f);
  }
  return copysignf(fdiv(yp, yq), s);
}

// apply_gamma's scalar tail (jxl-color/src/tf.rs:62-69): v <= 1e-7 ? 0 : fast_powf_generic(v, gamma)
__device__ __forceinline__ float gamma_tf(float a, float gamma) {
  if (a <= 1e-7f) return 0.0f;
  const int32_t x_bits = __float_as_int(a);
  const int32_t exp_shifted = (x_bits - 0x3f2aaaab) >> 23;
  const float mantissa = __int_as_float(x_bits - (exp_shifted << 23));
  const float x = fsub(mantissa, 1.0f);
  const float yp = fadd(fmul(fadd(fmul(7.4245873327820566e-1f, x), 1.4287160470083755f), x), -1.8503833400518310e-6f);
  const float yq = fadd(fmul(fadd(fmul(1.7409343003366853e-1f, x), 1.0096718572241148f), x), 9.9032814277590719e-1f);
  const float e = fmul(fadd(fdiv(yp, yq), float(exp_shifted)), gamma);
  const float x_floor = floorf(e);
  const float ex = __int_as_float(int32_t(uint32_t(__float2int_rz(x_floor) + 127) << 23));  // saturating, NaN -> 0
  const float frac = fsub(e, x_floor);
  float num = fadd(frac, 1.01749063e1f);
  num = fadd(fmul(num, frac), 4.88687798e1f);
  num = fadd(fmul(num, frac), 9.85506591e1f);
  num = fmul(num, ex);
  float den = fadd(fmul(2.10242958e-1f, frac), -2.22328856e-2f);
  den = fadd(fmul(den, frac), -1.94414990e1f);
  den = fadd(fmul(den, frac), 9.85506633e1f);
  return fdiv(num, den);
}

// map_gamut_generic (jxl-color/src/gamut.rs:4-46) followed by the merged target matrix (convert.rs:397-466)
__device__ __forceinline__ void second_colour_stage(const DevColorParams& p, float& o0, float& o1, float& o2) {
  float o[3] = {o0, o1, o2};
  const float yl = fadd(fadd(fmul(o[0], p.luminances[0]), fmul(o[1], p.luminances[1])), fmul(o[2], p.luminances[2]));
  float gray_saturation = 0.0f, gray_luminance = 0.0f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float v_sub_y = fsub(o[i], yl);
    const float inv = fdiv(1.0f, v_sub_y == 0.0f ? 1.0f : v_sub_y);
    const float v_over = fmul(o[i], inv);
    if (!(v_sub_y >= 0.0f)) gray_saturation = fmaxf(gray_saturation, v_over);
    gray_luminance = fmaxf(v_sub_y <= 0.0f ? gray_saturation : fsub(v_over, inv), gray_luminance);
  }
  float gray_mix = fadd(fmul(0.3f, fsub(gray_saturation, gray_luminance)), gray_luminance);
  gray_mix = gray_mix < 0.0f ? 0.0f : (gray_mix > 1.0f ? 1.0f : gray_mix);
  const float max_colour = fmaxf(o[2], fmaxf(o[1], fmaxf(o[0], 1.0f)));
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = fdiv(fadd(fmul(gray_mix, fsub(yl, o[i])), o[i]), max_colour);
  const float* m = p.matrix2;
  const float t0 = fadd(fadd(fmul(m[0], o[0]), fmul(m[1], o[1])), fmul(m[2], o[2]));
  const float t1 = fadd(fadd(fmul(m[3], o[0]), fmul(m[4], o[1])), fmul(m[5], o[2]));
  const float t2 = fadd(fadd(fmul(m[6], o[0]), fmul(m[7], o[1])), fmul(m[8], o[2]));
  o0 = p.to_luma ? t1 : t0;
  o1 = t1;
  o2 = t2;
}

__global__ void xyb_to_rgb_kernel(DevView vx, DevView vy, DevView vb, DevColorParams p) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= vx.w) return;
  float* px = static_cast<float*>(vx.ptr) + size_t(y) * vx.stride + x;
  float* py = static_cast<float*>(vy.ptr) + size_t(y) * vy.stride + x;
  float* pb = static_cast<float*>(vb.ptr) + size_t(y) * vb.stride + x;
  float xx = *px, yy = *py, bb = *pb;
  float g_l = fsub(fadd(yy, xx), p.cbrt_opsin_bias[0]);
  float g_m = fsub(fsub(yy, xx), p.cbrt_opsin_bias[1]);
  float g_s = fsub(bb, p.cbrt_opsin_bias[2]);
  float a = fmul(__fmaf_rn(fmul(g_l, g_l), g_l, p.opsin_bias[0]), p.itscale);
  float b = fmul(__fmaf_rn(fmul(g_m, g_m), g_m, p.opsin_bias[1]), p.itscale);
  float c = fmul(__fmaf_rn(fmul(g_s, g_s), g_s, p.opsin_bias[2]), p.itscale);
  const float* m = p.matrix;
  float o0 = fadd(fadd(fmul(m[0], a), fmul(m[1], b)), fmul(m[2], c));
  float o1 = fadd(fadd(fmul(m[3], a), fmul(m[4], b)), fmul(m[5], c));
  float o2 = fadd(fadd(fmul(m[6], a), fmul(m[7], b)), fmul(m[8], c));
  if (p.second_stage) second_colour_stage(p, o0, o1, o2);
  if (p.pq_intensity_target > 0.0f) {
    const float y_mult = fdiv(p.pq_intensity_target, 10000.0f);
    o0 = pq_tf(o0, y_mult);
    o1 = pq_tf(o1, y_mult);
    o2 = pq_tf(o2, y_mult);
  } else if (p.gamma > 0.0f) {
    o0 = gamma_tf(o0, p.gamma);
    o1 = gamma_tf(o1, p.gamma);
    o2 = gamma_tf(o2, p.gamma);
  } else if (p.apply_srgb_tf) {
    o0 = linear_to_srgb(o0);
    o1 = linear_to_srgb(o1);
    o2 = linear_to_srgb(o2);
  } else if (p.apply_bt709_tf) {
    o0 = linear_to_bt709(o0);
    o1 = linear_to_bt709(o1);
    o2 = linear_to_bt709(o2);
  }
  *px = o0;
  *py = o1;
  *pb = o2;
}

// upsample_inner<K, NW> (features/upsampling.rs:45-132): one thread per output sample; the 5x5
// neighbourhood is read with mirrored coordinates (== the reference's mirror-padded copy).
__global__ void upsample_kernel(DevView in, DevView out, int k, const float* __restrict__ quarter) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= int(out.w) || y >= int(out.h)) return;
  const int gw = int(in.w), gh = int(in.h), mat_n = k / 2;
  const int ref_x = x / k, ref_y = y / k, px = x % k, py = y % k;
  const int mat_x = min(px, k - px - 1), mat_y = min(py, k - py - 1);
  const bool flip_h = px >= mat_n, flip_v = py >= mat_n;
  const float* kernel = quarter + (mat_y * mat_n + mat_x) * 25;
  const float* src = static_cast<const float*>(in.ptr);
  float sum = 0.0f, mn = INFINITY, mx = -INFINITY;
#pragma unroll
  for (int iy = 0; iy < 5; ++iy) {
    const int ky = flip_v ? 4 - iy : iy;
    // One-sample-wide / -high images: the reference fills its padding in place (util.rs:423-454), which leaves
    // zeros two samples to the left and right of a single column and two rows above a single row.
    const int sy = gh == 1 ? 0 : mirror(ref_y + iy - 2, gh);
    const bool zero_row = gh == 1 && iy == 0;
#pragma unroll
    for (int ix = 0; ix < 5; ++ix) {
      const int kx = flip_h ? 4 - ix : ix;
      const int sx = gw == 1 ? 0 : mirror(ref_x + ix - 2, gw);
      const bool zero = zero_row || (gw == 1 && (ix == 0 || ix == 4));
      const float sample = zero ? 0.0f : src[size_t(sy) * in.stride + sx];
      sum = fadd(sum, fmul(__ldg(kernel + ky * 5 + kx), sample));
      mn = fminf(mn, sample);
      mx = fmaxf(mx, sample);
    }
  }
  float r;
  if (!isfinite(mn)) r = __int_as_float(0x7fc00000);
  else r = sum < mn ? mn : (sum > mx ? mx : sum);
  static_cast<float*>(out.ptr)[size_t(y) * out.stride + x] = r;
}

// One thread per output pixel; (x, y) are coordinates in the oriented output image
// (fb.rs:387-401 to_original_coord, sample rules fb.rs:436-520).
__global__ void pack_interleaved_kernel(DevPackParams p, void* out) {
  const uint32_t ow = p.orientation >= 5 ? p.height : p.width, oh = p.orientation >= 5 ? p.width : p.height;
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  uint32_t sx, sy;
  switch (p.orientation) {
    case 2: sx = ow - x - 1, sy = y; break;
    case 3: sx = ow - x - 1, sy = oh - y - 1; break;
    case 4: sx = x, sy = oh - y - 1; break;
    case 5: sx = y, sy = x; break;
    case 6: sx = y, sy = ow - x - 1; break;
    case 7: sx = oh - y - 1, sy = ow - x - 1; break;
    case 8: sx = oh - y - 1, sy = x; break;
    default: sx = x, sy = y; break;
  }
  const size_t o = (size_t(y) * ow + x) * p.num_channels;
  for (uint32_t c = 0; c < p.num_channels; ++c) {
    float v = p.planes[c][size_t(sy) * p.strides[c] + sx];
    if (c < 3)
      for (uint32_t s = 0; s < p.num_spots; ++s) {
        const float mix = fmul(p.spot_planes[s][size_t(sy) * p.spot_strides[s] + sx], p.spot_solidity[s]);
        v = fadd(fmul(p.spot_rgb[s][c], mix), fmul(v, fsub(1.0f, mix)));
      }
    if (p.sample_type == 2) {
      static_cast<float*>(out)[o + c] = v;
    } else {
      const float hi = p.sample_type == 0 ? 255.0f : 65535.0f;
      float t = fadd(fmul(v, hi), 0.5f);
      t = t < 0.0f ? 0.0f : (t > hi ? hi : t);  // f32::clamp; NaN falls through and casts to 0
      const uint32_t q = (t == t) ? uint32_t(t) : 0u;
      if (p.sample_type == 0) static_cast<uint8_t*>(out)[o + c] = uint8_t(q);
      else static_cast<uint16_t*>(out)[o + c] = uint16_t(q);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Patches without alpha (blend.rs:550-606). Jobs of one launch may overlap in `dst` only if the bitstream
// makes patches overlap; they are then applied in job order by launching overlapping batches separately
// (the host splits the list), so inside a launch every destination sample has one writer.
__global__ void blend_patches_kernel(const DevPatchJob* __restrict__ jobs) {
  const DevPatchJob j = jobs[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < j.w * j.h; i += blockDim.x) {
    const uint32_t x = i % j.w, y = i / j.w;
    float v = j.src[size_t(y) * j.src_stride + x];
    float* d = j.dst + size_t(y) * j.dst_stride + x;
    const float base = *d;
    float r;
    if (j.mode == 1) {
      r = v;
    } else if (j.mode == 2) {
      r = fadd(base, v);
    } else if (j.mode == 3) {
      if (j.clamp) v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
      r = fmul(base, v);
    } else if (j.mode == 4) {
      const float frame_alpha = j.base_alpha ? j.base_alpha[size_t(y) * j.base_alpha_stride + x] : 0.0f;
      const float patch_alpha = j.new_alpha ? j.new_alpha[size_t(y) * j.new_alpha_stride + x] : 0.0f;
      const float base_sample = j.swapped ? v : base, new_sample = j.swapped ? base : v;
      const float base_alpha = j.swapped ? patch_alpha : frame_alpha;
      float new_alpha = j.swapped ? frame_alpha : patch_alpha;
      if (j.clamp) new_alpha = new_alpha < 0.0f ? 0.0f : (new_alpha > 1.0f ? 1.0f : new_alpha);
      if (j.premultiplied) {
        r = fadd(new_sample, fmul(base_sample, fsub(1.0f, new_alpha)));
      } else {
        const float base_alpha_rev = fsub(1.0f, base_alpha), new_alpha_rev = fsub(1.0f, new_alpha);
        const float mixed_alpha = fsub(1.0f, fmul(new_alpha_rev, base_alpha_rev));
        const float mixed_alpha_recip = mixed_alpha > 0.0f ? fdiv(1.0f, mixed_alpha) : 0.0f;
        r = fmul(fadd(fmul(new_alpha, new_sample), fmul(fmul(base_alpha, base_sample), new_alpha_rev)), mixed_alpha_recip);
      }
    } else if (j.mode == 5) {
      const float base_sample = j.swapped ? v : base, new_sample = j.swapped ? base : v;
      float new_alpha;
      if (j.swapped) new_alpha = j.base_alpha ? j.base_alpha[size_t(y) * j.base_alpha_stride + x] : 0.0f;
      else new_alpha = j.new_alpha ? j.new_alpha[size_t(y) * j.new_alpha_stride + x] : 0.0f;
      if (j.clamp) new_alpha = new_alpha < 0.0f ? 0.0f : (new_alpha > 1.0f ? 1.0f : new_alpha);
      r = fadd(base_sample, fmul(new_alpha, new_sample));
    } else {
      const float b0 = j.swapped ? v : base;
      float n0 = j.swapped ? base : v;
      if (j.clamp) n0 = n0 < 0.0f ? 0.0f : (n0 > 1.0f ? 1.0f : n0);
      r = fadd(b0, fmul(n0, fsub(1.0f, b0)));
    }
    *d = r;
  }
}

// ---------------------------------------------------------------------------------------------
// Splines (features/spline.rs:218-252, erf :314-331)
__device__ __forceinline__ float spline_erf(float x) {
  const float ax = fabsf(x);
  const float denom1 = fadd(fmul(ax, 7.77394369e-02f), 2.05260015e-04f);
  const float denom2 = fadd(fmul(denom1, ax), 2.32120216e-01f);
  const float denom3 = fadd(fmul(denom2, ax), 2.77820801e-01f);
  const float denom4 = fadd(fmul(denom3, ax), 1.0f);
  const float denom5 = fmul(denom4, denom4);
  const float inv_denom5 = fdiv(1.0f, denom5);
  const float result = fadd(fmul(-inv_denom5, inv_denom5), 1.0f);
  return x < 0.0f ? -result : result;
}

// One thread per pixel, 32x8 tiles. The arc list is walked in chunks of 256: every thread tests one arc's bounding box
// against the tile and the hits are compacted IN LIST ORDER into shared memory (float addition order is part of the
// result), then every pixel accumulates the surviving arcs.
__global__ void __launch_bounds__(256) splat_splines_kernel(DevView v0, DevView v1, DevView v2, const DevSplineArc* __restrict__ arcs,
                                                            int num_arcs) {
  __shared__ DevSplineArc hit[256];
  __shared__ int warp_count[8];
  const int tid = threadIdx.y * 32 + threadIdx.x, lane = threadIdx.x, warp = threadIdx.y;
  const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 8;
  const int x = tx0 + lane, y = ty0 + warp;
  const bool inside = x < int(v0.w) && y < int(v0.h);
  float* p[3] = {static_cast<float*>(v0.ptr) + size_t(y) * v0.stride + x, static_cast<float*>(v1.ptr) + size_t(y) * v1.stride + x,
                 static_cast<float*>(v2.ptr) + size_t(y) * v2.stride + x};
  float acc[3] = {0.0f, 0.0f, 0.0f};
  if (inside) acc[0] = *p[0], acc[1] = *p[1], acc[2] = *p[2];
  bool touched = false;
  for (int base = 0; base < num_arcs; base += 256) {
    DevSplineArc a;
    bool h = false;
    if (base + tid < num_arcs) {
      a = arcs[base + tid];
      h = a.xbegin < tx0 + 32 && a.xend > tx0 && a.ybegin < ty0 + 8 && a.yend > ty0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, h);
    if (lane == 0) warp_count[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) before += warp_count[w];
      total += warp_count[w];
    }
    if (h) hit[before + __popc(m & ((1u << lane) - 1u))] = a;
    __syncthreads();
    if (inside)
      for (int i = 0; i < total; ++i) {
        const DevSplineArc& q = hit[i];
        if (x < q.xbegin || x >= q.xend || y < q.ybegin || y >= q.yend) continue;
        const float dx = fsub(float(x), q.x), dy = fsub(float(y), q.y);
        const float distance = __fsqrt_rn(fadd(fmul(dx, dx), fmul(dy, dy)));
        const float factor = fsub(spline_erf(fmul(fadd(fmul(0.5f, distance), 0.35355338f), q.inv_sigma)),
                                  spline_erf(fmul(fsub(fmul(0.5f, distance), 0.35355338f), q.inv_sigma)));
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = fadd(acc[c], fmul(fmul(fmul(fmul(0.25f, q.value[c]), q.sigma), factor), factor));
        touched = true;
      }
    __syncthreads();
  }
  if (touched) {
    *p[0] = acc[0];
    *p[1] = acc[1];
    *p[2] = acc[2];
  }
}

// ---------------------------------------------------------------------------------------------
// Noise synthesis (features/noise.rs)
__device__ __forceinline__ unsigned long long split_mix_64(unsigned long long z) {  // noise.rs:454-458
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// The reference's generator is eight independent xorshift128+ streams (noise.rs:405-451) whose outputs are
// interleaved into batches of 16 floats; one thread per (group, stream) walks its stream through the
// group's three channels and stores the two floats each step contributes.
__global__ void noise_field_kernel(float* f0, float* f1, float* f2, uint32_t width, uint32_t height, uint32_t group_dim,
                                   unsigned long long seed0) {
  const uint32_t gpr = (width + group_dim - 1) / group_dim, gpc = (height + group_dim - 1) / group_dim;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = t >> 3, lane = t & 7;
  if (g >= gpr * gpc) return;
  const uint32_t x0 = (g % gpr) * group_dim, y0 = (g / gpr) * group_dim;
  const uint32_t gw = min(group_dim, width - x0), gh = min(group_dim, height - y0);
  const unsigned long long seed1 = ((unsigned long long)(x0) << 32) + (unsigned long long)(y0);
  unsigned long long s0 = split_mix_64(seed0 + 0x9E3779B97F4A7C15ull), s1 = split_mix_64(seed1 + 0x9E3779B97F4A7C15ull);
  for (uint32_t i = 0; i < lane; ++i) {
    s0 = split_mix_64(s0);
    s1 = split_mix_64(s1);
  }
  const uint32_t width_n2 = (gw + 15) / 16;
  float* planes[3] = {f0, f1, f2};
  for (int c = 0; c < 3; ++c) {
    float* dst = planes[c];
    for (uint32_t row = 0; row < gh; ++row)
      for (uint32_t cb = 0; cb < width_n2; ++cb) {
        unsigned long long a = s0;
        const unsigned long long b = s1;
        const unsigned long long ret = a + b;
        s0 = b;
        a ^= a << 23;
        s1 = a ^ (b ^ (a >> 18) ^ (b >> 5));
        const uint32_t x = cb * 16 + 2 * lane;
        float* o = dst + size_t(y0 + row) * width + x0 + x;
        if (x < gw) o[0] = __uint_as_float((uint32_t(ret) >> 9) | 0x3f800000u);
        if (x + 1 < gw) o[1] = __uint_as_float((uint32_t(ret >> 32) >> 9) | 0x3f800000u);
      }
  }
}

// 5x5 high-pass of the field (rows added in the order of the reference's 5-row ring buffer, which depends
// on the row's position inside its group: noise.rs:297-320) and application to the XYB planes (noise.rs:45-83).
__global__ void noise_apply_kernel(DevView vx, DevView vy, DevView vb, const float* __restrict__ f0,
                                   const float* __restrict__ f1, const float* __restrict__ f2, DevNoiseParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  const int width = int(vx.w), height = int(vx.h);
  if (x >= width || y >= height) return;
  const int ly = y % int(p.group_dim);
  const float* fields[3] = {f0, f1, f2};
  float n[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* f = fields[c];
    float sum = 0.0f;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
      const int r = ly - 2 + (((s - ly) % 5) + 5) % 5;  // local row held by ring slot s
      const int sy = mirror(y - ly + r, height);
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const int sx = mirror(x + dx - 2, width);
        sum = fadd(sum, fmul(f[size_t(sy) * width + sx], 0.16f));
      }
    }
    n[c] = fsub(sum, fmul(f[size_t(y) * width + x], 4.0f));
  }
  float* px = static_cast<float*>(vx.ptr) + size_t(y) * vx.stride + x;
  float* py = static_cast<float*>(vy.ptr) + size_t(y) * vy.stride + x;
  float* pb = static_cast<float*>(vb.ptr) + size_t(y) * vb.stride + x;
  const float grid_x = *px, grid_y = *py;
  const float in_x = fadd(grid_x, grid_y), in_y = fsub(grid_y, grid_x);
  const float in_scaled_x = fmaxf(0.0f, fmul(in_x, 3.0f)), in_scaled_y = fmaxf(0.0f, fmul(in_y, 3.0f));
  // `as usize` saturates (and maps NaN to 0) before the min with 7
  const uint32_t in_x_int = min(uint32_t(__float2uint_rz(in_scaled_x)), 7u), in_y_int = min(uint32_t(__float2uint_rz(in_scaled_y)), 7u);
  const float in_x_frac = fsub(in_scaled_x, float(in_x_int)), in_y_frac = fsub(in_scaled_y, float(in_y_int));
  const float sx = fadd(fmul(fsub(p.lut[in_x_int + 1], p.lut[in_x_int]), in_x_frac), p.lut[in_x_int]);
  const float sy = fadd(fmul(fsub(p.lut[in_y_int + 1], p.lut[in_y_int]), in_y_frac), p.lut[in_y_int]);
  const float nx = fmul(fmul(0.22f, sx), fadd(fmul(0.0078125f, n[0]), fmul(0.9921875f, n[2])));
  const float ny = fmul(fmul(0.22f, sy), fadd(fmul(0.0078125f, n[1]), fmul(0.9921875f, n[2])));
  const float nsum = fadd(nx, ny);
  *px = fadd(*px, fsub(fadd(fmul(p.corr_x, nsum), nx), ny));
  *py = fadd(*py, nsum);
  *pb = fadd(*pb, fmul(p.corr_b, nsum));
}

}  // namespace

void launch_blend_patches(const DevPatchJob* jobs, int num_jobs, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  blend_patches_kernel<<<num_jobs, 128, 0, stream>>>(jobs);
}

void launch_splat_splines(const DevView v[3], const DevSplineArc* arcs, int num_arcs, cudaStream_t stream) {
  if (num_arcs <= 0 || !v[0].w || !v[0].h) return;
  dim3 block(32, 8);
  dim3 grid((v[0].w + 31) / 32, (v[0].h + 7) / 8);
  splat_splines_kernel<<<grid, block, 0, stream>>>(v[0], v[1], v[2], arcs, num_arcs);
}

void launch_add_noise(const DevView v[3], float* const field[3], DevNoiseParams p, cudaStream_t stream) {
  const uint32_t width = v[0].w, height = v[0].h;
  if (!width || !height) return;
  const uint32_t groups = ((width + p.group_dim - 1) / p.group_dim) * ((height + p.group_dim - 1) / p.group_dim);
  noise_field_kernel<<<(groups * 8 + 63) / 64, 64, 0, stream>>>(field[0], field[1], field[2], width, height, p.group_dim, p.seed0);
  dim3 block(32, 8);
  dim3 grid((width + 31) / 32, (height + 7) / 8);
  noise_apply_kernel<<<grid, block, 0, stream>>>(v[0], v[1], v[2], field[0], field[1], field[2], p);
}

void launch_pack_interleaved(DevPackParams p, void* out, cudaStream_t stream) {
  if (!p.width || !p.height || !p.num_channels) return;
  const uint32_t ow = p.orientation >= 5 ? p.height : p.width, oh = p.orientation >= 5 ? p.width : p.height;
  dim3 block(32, 8);
  dim3 grid((ow + 31) / 32, (oh + 7) / 8);
  pack_interleaved_kernel<<<grid, block, 0, stream>>>(p, out);
}

void launch_upsample(DevView in, DevView out, int k, const float* quarter, cudaStream_t stream) {
  if (!out.w || !out.h) return;
  dim3 block(32, 8);
  dim3 grid((out.w + 31) / 32, (out.h + 7) / 8);
  upsample_kernel<<<grid, block, 0, stream>>>(in, out, k, quarter);
}

void launch_gaborish(DevView in, DevView out, float w0, float w1, cudaStream_t stream) {
  if (!in.w || !in.h) return;
  dim3 grid((in.w + 127) / 128, in.h);
  gaborish_kernel<<<grid, 128, 0, stream>>>(in, out, w0, w1);
}

void launch_epf_step(const DevView in[3], const DevView out[3], const float* sigma, uint32_t sigma_stride, DevEpfParams p,
                     int step, cudaStream_t stream) {
  View3 vi, vo;
  for (int c = 0; c < 3; ++c) {
    vi.v[c] = in[c];
    vo.v[c] = out[c];
  }
  if (!vi.v[0].w || !vi.v[0].h) return;
  dim3 block(32, 8);
  dim3 grid((vi.v[0].w + 31) / 32, (vi.v[0].h + 7) / 8);
  if (step == 0) epf_kernel<0><<<grid, block, 0, stream>>>(vi, vo, sigma, sigma_stride, p);
  else if (step == 1) epf_kernel<1><<<grid, block, 0, stream>>>(vi, vo, sigma, sigma_stride, p);
  else epf_kernel<2><<<grid, block, 0, stream>>>(vi, vo, sigma, sigma_stride, p);
}

void launch_upsample_jpeg(DevView in, DevView out, int horizontal, int vertical, cudaStream_t stream) {
  if (!out.w || !out.h) return;
  dim3 grid((out.w + 127) / 128, out.h);
  upsample_jpeg_kernel<<<grid, 128, 0, stream>>>(in, out, horizontal, vertical);
}

void launch_ycbcr_to_rgb(DevView cb, DevView y, DevView cr, DevYcbcrParams p, cudaStream_t stream) {
  if (!y.w || !y.h) return;
  dim3 grid((y.w + 127) / 128, y.h);
  ycbcr_to_rgb_kernel<<<grid, 128, 0, stream>>>(cb, y, cr, p);
}

void launch_xyb_to_rgb(DevView x, DevView y, DevView b, DevColorParams p, cudaStream_t stream) {
  if (!x.w || !x.h) return;
  dim3 grid((x.w + 127) / 128, x.h);
  xyb_to_rgb_kernel<<<grid, 128, 0, stream>>>(x, y, b, p);
}

}  // namespace jxlb
