// VarDCT stages on the device: HfMetadata placement scan, HF coefficient entropy decode, LF
// dequant / chroma-from-luma / adaptive smoothing, HF dequant + chroma-from-luma, LLF insertion
// and the 27 inverse transforms. Float op order follows the reference's generic code path
// (crates/jxl-render/src/vardct/{mod.rs,transform_common.rs,generic/*.rs}); this file is
// compiled with -fmad=false and fuses only where the reference calls mul_add.
#include "kernels.h"

namespace jxlb {

#define JXLB_TABLE_QUAL __device__ __constant__ const
#include "../host/jxl_tables.inc"
#undef JXLB_TABLE_QUAL

// TransformTypeInfo (host/frame_syntax.cc kTransformInfo): w8, h8, param, order, transpose
__device__ __constant__ const uint8_t kDevTransformInfo[27][5] = {
    {1, 1, 0, 0, 1},  {1, 1, 1, 1, 0},  {1, 1, 2, 1, 0},   {1, 1, 3, 1, 0},    {2, 2, 4, 2, 1},   {4, 4, 5, 3, 1},
    {1, 2, 6, 4, 1},  {2, 1, 6, 4, 0},  {1, 4, 7, 5, 1},   {4, 1, 7, 5, 0},    {2, 4, 8, 6, 1},   {4, 2, 8, 6, 0},
    {1, 1, 9, 1, 0},  {1, 1, 9, 1, 0},  {1, 1, 10, 1, 0},  {1, 1, 10, 1, 0},   {1, 1, 10, 1, 0},  {1, 1, 10, 1, 0},
    {8, 8, 11, 7, 1}, {4, 8, 12, 8, 1}, {8, 4, 12, 8, 0},  {16, 16, 13, 9, 1}, {8, 16, 14, 10, 1}, {16, 8, 14, 10, 0},
    {32, 32, 15, 11, 1}, {16, 32, 16, 12, 1}, {32, 16, 16, 12, 0},
};

// sec_half tables for n = 64, 128, 256 (computed on the host with cosf, dct_common.rs:57-67)
__device__ __constant__ float kSecLarge[32 + 64 + 128];
void upload_sec_large(const float* host224) { cudaMemcpyToSymbol(kSecLarge, host224, sizeof(float) * 224); }

namespace {

__device__ __forceinline__ const float* sec_half(int n) {
  switch (n) {
    case 4: return kSecHalf4;
    case 8: return kSecHalf8;
    case 16: return kSecHalf16;
    case 32: return kSecHalf32;
    case 64: return kSecLarge;
    case 128: return kSecLarge + 32;
    default: return kSecLarge + 96;
  }
}

// ---------------------------------------------------------------------------------------------
// HfMetadata::parse placement scan (hf_metadata.rs:99-230): one thread per LF group.
__global__ void build_block_info_kernel(DevFrame f, const DevBlockInfoJob* jobs, int num_jobs, float quant_mul_base,
                                        const float* sharp_lut, int has_epf, int* status) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num_jobs) return;
  const DevBlockInfoJob job = jobs[j];
  const DevLfGroupRect rc = job.rect;
  const int32_t kUninit = INT32_MIN;
  for (uint32_t y = 0; y < rc.bh; ++y)
    for (uint32_t x = 0; x < rc.bw; ++x) f.blk_type[size_t(rc.by0 + y) * f.bw + rc.bx0 + x] = kUninit;
  uint32_t data_idx = 0;
  for (uint32_t y = 0; y < rc.bh; ++y) {
    for (uint32_t x = 0; x < rc.bw;) {
      if (f.blk_type[size_t(rc.by0 + y) * f.bw + rc.bx0 + x] != kUninit) {
        ++x;
        continue;
      }
      if (data_idx >= job.nb_blocks) {
        status[j] = kDevInvalid;
        return;
      }
      int32_t dct_select = job.raw[data_idx];
      int32_t hf_mul = job.raw[job.raw_stride + data_idx] + 1;
      if (dct_select < 0 || dct_select >= 27 || hf_mul <= 0) {
        status[j] = kDevInvalid;
        return;
      }
      uint32_t dw = kDevTransformInfo[dct_select][0], dh = kDevTransformInfo[dct_select][1];
      if ((x % 32) + dw > 32 || (y % 32) + dh > 32 || x + dw > rc.bw || y + dh > rc.bh) {
        status[j] = kDevInvalid;
        return;
      }
      float sigma_q = __fdiv_rn(quant_mul_base, float(hf_mul));
      for (uint32_t dy = 0; dy < dh; ++dy)
        for (uint32_t dx = 0; dx < dw; ++dx) {
          size_t gi = size_t(rc.by0 + y + dy) * f.bw + rc.bx0 + x + dx;
          if (f.blk_type[gi] != kUninit) {
            status[j] = kDevInvalid;
            return;
          }
          f.blk_type[gi] = (dx == 0 && dy == 0) ? dct_select : -int32_t(1 + dx + 32 * dy);
          f.blk_mul[gi] = hf_mul;
          if (has_epf) {
            int32_t s = f.sharpness[gi];
            if (s < 0 || s >= 8) {
              status[j] = kDevInvalid;
              return;
            }
            f.epf_sigma[gi] = __fmul_rn(sigma_q, sharp_lut[s]);
          }
        }
      ++data_idx;
      x += dw;
    }
  }
  status[j] = kDevOk;
}

// ---------------------------------------------------------------------------------------------
// write_hf_coeff (jxl-vardct/src/hf_coeff.rs:21-252): one warp per (pass, 256x256 group); lane 0
// runs the serial ANS / context chain.
__global__ void decode_hf_kernel(const uint8_t* __restrict__ cs, DevFrame f, DevHfParams p,
                                 const DevHfJob* __restrict__ jobs, uint64_t* __restrict__ end_bits,
                                 int* __restrict__ status, int num_jobs) {
  int job_idx = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
  if (job_idx >= num_jobs || (threadIdx.x & 31) != 0) return;
  const DevHfJob job = jobs[job_idx];
  const uint32_t nbc = p.num_block_clusters;
  const uint32_t lf_idx_mul = (p.num_lf_thr[0] + 1) * (p.num_lf_thr[1] + 1) * (p.num_lf_thr[2] + 1);
  const uint32_t hf_idx_mul = p.num_qf_thr + 1;
  DevBitReader br;
  br.init(cs, job.bit_pos);
  int err = kDevOk;
  uint32_t hfp_bits = 0;
  while ((1u << hfp_bits) < p.num_hf_presets) ++hfp_bits;
  uint32_t hfp = br.read(hfp_bits);
  if (hfp >= p.num_hf_presets) err = kDevInvalid;
  const uint8_t* cluster_map = p.code.cluster_map + size_t(495) * nbc * (err ? 0 : hfp);
  DevEntropyState es;
  entropy_begin(p.code, es, br, nullptr);

  const uint32_t gx = job.group_idx % p.groups_per_row, gy = job.group_idx / p.groups_per_row;
  const uint32_t gb = p.group_dim_blocks;
  const uint32_t bx0 = gx * gb, by0 = gy * gb;
  const uint32_t width = min(gb, f.bw - bx0), height = min(gb, f.bh - by0);
  uint32_t nz_row[3][32];
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < 32; ++i) nz_row[c][i] = 0;
  const int32_t* thr_base[3] = {p.lf_thresholds, p.lf_thresholds + p.num_lf_thr[0],
                                p.lf_thresholds + p.num_lf_thr[0] + p.num_lf_thr[1]};

  for (uint32_t y = 0; y < height && err == kDevOk; ++y)
    for (uint32_t x = 0; x < width && err == kDevOk; ++x) {
      size_t gi = size_t(by0 + y) * f.bw + bx0 + x;
      int32_t t = f.blk_type[gi];
      if (t < 0) continue;
      int32_t qf = f.blk_mul[gi];
      const uint32_t w8 = kDevTransformInfo[t][0], h8 = kDevTransformInfo[t][1];
      const uint32_t order_id = kDevTransformInfo[t][3];
      const bool transpose = kDevTransformInfo[t][4] != 0;
      const uint32_t num_blocks = w8 * h8;
      const uint32_t num_blocks_log = 31u - uint32_t(__clz(int(num_blocks)));
      uint32_t lf_idx = 0;
      {
        const int cs3[3] = {0, 2, 1};
        for (int k = 0; k < 3; ++k) {
          int c = cs3[k];
          lf_idx *= p.num_lf_thr[c] + 1;
          int32_t q = f.lf_quant[c][gi];
          for (uint32_t i = 0; i < p.num_lf_thr[c]; ++i)
            if (q > thr_base[c][i]) ++lf_idx;
        }
      }
      uint32_t hf_idx = 0;
      for (uint32_t i = 0; i < p.num_qf_thr; ++i)
        if (qf > int32_t(p.qf_thresholds[i])) ++hf_idx;
      for (int ci = 0; ci < 3 && err == kDevOk; ++ci) {
        const uint32_t ch_idx = uint32_t(ci) * 13 + order_id;
        const int c = (ci == 0) ? 1 : (ci == 1 ? 0 : 2);
        const uint32_t idx = (ch_idx * hf_idx_mul + hf_idx) * lf_idx_mul + lf_idx;
        const uint32_t block_ctx = p.block_ctx_map[idx];
        uint32_t predicted;
        if (y == 0) predicted = x == 0 ? 32 : nz_row[c][x - 1];
        else if (x == 0) predicted = nz_row[c][x];
        else predicted = (nz_row[c][x] + nz_row[c][x - 1] + 1) >> 1;
        const uint32_t pidx = predicted >= 8 ? 4 + predicted / 2 : predicted;
        const uint32_t nz_ctx = block_ctx + pidx * nbc;
        uint32_t non_zeros = entropy_read_varint(p.code, es, br, cluster_map[nz_ctx], 0, err);
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
        for (uint32_t k = num_blocks, i = 0; k < size; ++k, ++i) {
          const uint32_t nzc = (non_zeros - 1) >> num_blocks_log;
          const uint32_t fi = i >> num_blocks_log;
          const uint32_t cctx = (uint32_t(kCoeffNumNonzeroContext[nzc]) + uint32_t(kCoeffFreqContext[fi])) * 2 + prev_nonzero;
          if (cctx >= 458) {
            err = kDevInvalid;
            break;
          }
          const uint32_t ucoeff = entropy_read_varint(p.code, es, br, cmap[cctx], 0, err);
          if (ucoeff == 0) {
            prev_nonzero = 0;
            continue;
          }
          const uint32_t cv = uint32_t(dev_unpack_signed(ucoeff)) << p.coeff_shift;
          const uint32_t o = __ldg(order + k);
          uint32_t dx = o & 0xffff, dy = o >> 16;
          if (transpose) {
            uint32_t tmp = dx;
            dx = dy;
            dy = tmp;
          }
          const size_t px = size_t(bx0 + x) * 8 + dx, py = size_t(by0 + y) * 8 + dy;
          plane[py * f.cw + px] += cv;
          prev_nonzero = 1;
          if (--non_zeros == 0) break;
        }
        if (br.pos > job.bit_limit) err = kDevOverrun;
      }
    }
  if (err == kDevOk && !entropy_final_ok(p.code, es)) err = kDevBadStream;
  if (err == kDevOk && br.pos > job.bit_limit) err = kDevOverrun;
  end_bits[job_idx] = br.pos;
  status[job_idx] = err;
}

// ---------------------------------------------------------------------------------------------
// LF (vardct/mod.rs:387-412, 544-568; generic/mod.rs:11-103)
__global__ void lf_dequant_kernel(DevFrame f, const DevLfDequantJob* jobs) {
  const DevLfDequantJob j = jobs[blockIdx.z];
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= j.rect.bw || y >= j.rect.bh) return;
  size_t i = size_t(j.rect.by0 + y) * f.bw + j.rect.bx0 + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) f.lf[c][i] = __fmul_rn(float(f.lf_quant[c][i]), j.scale[c]);
}

__global__ void lf_cfl_kernel(DevFrame f, float kx, float kb) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= size_t(f.bw) * f.bh) return;
  float y = f.lf[1][i];
  f.lf[0][i] = __fadd_rn(f.lf[0][i], __fmul_rn(kx, y));
  f.lf[2][i] = __fadd_rn(f.lf[2][i], __fmul_rn(kb, y));
}

// Reads the original planes `f.lf`, writes `out` (the reference updates in place but only ever
// reads original values: left neighbour is saved, up/down sums are precomputed).
__global__ void lf_smooth_kernel(DevFrame f, float* out0, float* out1, float* out2, float lf_x, float lf_y, float lf_b) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= f.bw) return;
  const size_t w = f.bw;
  const size_t i = size_t(y) * w + x;
  float* out[3] = {out0, out1, out2};
  const float lfs[3] = {lf_x, lf_y, lf_b};
  if (f.bw <= 2 || f.bh <= 2 || x == 0 || y == 0 || x + 1 >= f.bw || y + 1 >= f.bh) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c][i] = f.lf[c][i];
    return;
  }
  const float kSelf = 0.052262735f, kSide = 0.2034514f, kDiag = 0.03348292f;
  float self[3], wa[3], gap = 0.5f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = f.lf[c];
    self[c] = p[i];
    float ud_c = __fadd_rn(p[i - w], p[i + w]);
    float ud_l = __fadd_rn(p[i - w - 1], p[i + w - 1]);
    float ud_r = __fadd_rn(p[i - w + 1], p[i + w + 1]);
    float side = __fadd_rn(__fadd_rn(p[i - 1], p[i + 1]), ud_c);
    float diag = __fadd_rn(ud_l, ud_r);
    wa[c] = __fadd_rn(__fadd_rn(__fmul_rn(self[c], kSelf), __fmul_rn(side, kSide)), __fmul_rn(diag, kDiag));
    float gap_t = __fdiv_rn(fabsf(__fsub_rn(wa[c], self[c])), lfs[c]);
    gap = fmaxf(gap, gap_t);
  }
  float gap_scale = fmaxf(__fsub_rn(3.0f, __fmul_rn(4.0f, gap)), 0.0f);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[c][i] = __fadd_rn(__fmul_rn(__fsub_rn(wa[c], self[c]), gap_scale), self[c]);
}

// ---------------------------------------------------------------------------------------------
// dequant_hf_varblock_grouped + chroma_from_luma_hf_grouped (vardct/mod.rs:442-542, 570-603),
// one thread per coefficient position, all three channels.
__global__ void hf_dequant_cfl_kernel(DevFrame f, DevDequantParams p) {
  uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= f.cw || y >= f.ch) return;
  uint32_t bx = x >> 3, by = y >> 3;
  int32_t t = f.blk_type[size_t(by) * f.bw + bx];
  uint32_t ox = bx, oy = by;
  if (t < 0) {
    uint32_t code = uint32_t(-t - 1);
    ox = bx - (code & 31);
    oy = by - (code >> 5);
    t = f.blk_type[size_t(oy) * f.bw + ox];
  }
  const uint32_t w = uint32_t(kDevTransformInfo[t][0]) * 8;
  const uint32_t set = kDevTransformInfo[t][2], tr = kDevTransformInfo[t][4];
  const uint32_t ix = x - ox * 8, iy = y - oy * 8;
  const float hf_mul = float(f.blk_mul[size_t(oy) * f.bw + ox]);
  const size_t i = size_t(y) * f.cw + x;
  float v[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float mul = __fmul_rn(__fdiv_rn(65536.0f, __fmul_rn(p.global_scale, hf_mul)), p.qm_scale[c]);
    float m = __ldg(p.matrices + p.matrix_offset[(set * 3 + c) * 2 + tr] + iy * w + ix);
    float q = float(int32_t(f.coeff[c][i]));
    if (fabsf(q) <= 1.0f) q = __fmul_rn(q, p.quant_bias[c]);
    else q = __fsub_rn(q, __fdiv_rn(p.quant_bias_numerator, q));
    q = __fmul_rn(q, m);
    q = __fmul_rn(q, mul);
    v[c] = q;
  }
  size_t ti = size_t(y >> 6) * f.w64 + (x >> 6);
  float kx = __fadd_rn(p.base_correlation_x, __fdiv_rn(float(f.x_from_y[ti]), p.colour_factor));
  float kb = __fadd_rn(p.base_correlation_b, __fdiv_rn(float(f.b_from_y[ti]), p.colour_factor));
  v[0] = __fadd_rn(v[0], __fmul_rn(kx, v[1]));
  v[2] = __fadd_rn(v[2], __fmul_rn(kb, v[1]));
#pragma unroll
  for (int c = 0; c < 3; ++c) f.coeff[c][i] = __float_as_uint(v[c]);
}

// ---------------------------------------------------------------------------------------------
// 1-D DCT (generic/dct.rs:143-293). `io` and `scratch` hold N floats each.
#define SQRT2F 1.41421356237309504880f

__device__ __forceinline__ void dct4(float* io, bool forward) {
  const float sec0 = 0.5411961f, sec1 = 1.306563f;
  float i0 = io[0], i1 = io[1], i2 = io[2], i3 = io[3];
  if (forward) {
    float sum03 = __fadd_rn(i0, i3), sum12 = __fadd_rn(i1, i2);
    float tmp0 = __fmul_rn(__fsub_rn(i0, i3), sec0), tmp1 = __fmul_rn(__fsub_rn(i1, i2), sec1);
    float out0 = __fdiv_rn(__fadd_rn(tmp0, tmp1), 4.0f), out1 = __fdiv_rn(__fsub_rn(tmp0, tmp1), 4.0f);
    io[0] = __fdiv_rn(__fadd_rn(sum03, sum12), 4.0f);
    io[1] = __fadd_rn(__fmul_rn(out0, SQRT2F), out1);
    io[2] = __fdiv_rn(__fsub_rn(sum03, sum12), 4.0f);
    io[3] = out1;
  } else {
    float tmp0 = __fmul_rn(i1, SQRT2F), tmp1 = __fadd_rn(i1, i3);
    float out0 = __fmul_rn(__fadd_rn(tmp0, tmp1), sec0), out1 = __fmul_rn(__fsub_rn(tmp0, tmp1), sec1);
    float sum02 = __fadd_rn(i0, i2), sub02 = __fsub_rn(i0, i2);
    io[0] = __fadd_rn(sum02, out0);
    io[1] = __fadd_rn(sub02, out1);
    io[2] = __fsub_rn(sub02, out1);
    io[3] = __fsub_rn(sum02, out0);
  }
}

template <int N>
struct Dct1D {
  static __device__ __noinline__ void run(float* io, float* scratch, bool forward) {
    constexpr int h = N / 2;
    float* in0 = scratch;
    float* in1 = scratch + h;
    const float* sec = sec_half(N);
    if (forward) {
      for (int i = 0; i < h; ++i) {
        in0[i] = __fdiv_rn(__fadd_rn(io[i], io[N - i - 1]), 2.0f);
        in1[i] = __fdiv_rn(__fsub_rn(io[i], io[N - i - 1]), 2.0f);
      }
      for (int i = 0; i < h; ++i) in1[i] = __fmul_rn(in1[i], sec[i]);
      Dct1D<h>::run(in0, io, true);
      Dct1D<h>::run(in1, io + h, true);
      in1[0] = __fmul_rn(in1[0], SQRT2F);
      for (int i = 0; i + 1 < h; ++i) in1[i] = __fadd_rn(in1[i], in1[i + 1]);
      for (int i = 0; i < h; ++i) io[i * 2] = in0[i];
      for (int i = 0; i < h; ++i) io[i * 2 + 1] = in1[i];
    } else {
      for (int i = 0; i < h; ++i) {
        in0[i] = io[i * 2];
        in1[i] = io[i * 2 + 1];
      }
      for (int i = 1; i < h; ++i) in1[h - i] = __fadd_rn(in1[h - i], in1[h - i - 1]);
      in1[0] = __fmul_rn(in1[0], SQRT2F);
      Dct1D<h>::run(in0, io, false);
      Dct1D<h>::run(in1, io + h, false);
      for (int i = 0; i < h; ++i) in1[i] = __fmul_rn(in1[i], sec[i]);
      for (int i = 0; i < h; ++i) {
        float a = scratch[i], b = scratch[i + h];
        io[i] = __fadd_rn(a, b);
        io[N - i - 1] = __fsub_rn(a, b);
      }
    }
  }
};

template <>
struct Dct1D<8> {
  static __device__ __forceinline__ void run(float* io, float*, bool forward) {
    const float* sec = kSecHalf8;
    if (forward) {
      float in0[4], in1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        in0[i] = __fdiv_rn(__fadd_rn(io[i], io[7 - i]), 2.0f);
        in1[i] = __fdiv_rn(__fmul_rn(__fsub_rn(io[i], io[7 - i]), sec[i]), 2.0f);
      }
      dct4(in0, true);
#pragma unroll
      for (int i = 0; i < 4; ++i) io[i * 2] = in0[i];
      dct4(in1, true);
      in1[0] = __fmul_rn(in1[0], SQRT2F);
#pragma unroll
      for (int i = 0; i < 3; ++i) io[i * 2 + 1] = __fadd_rn(in1[i], in1[i + 1]);
      io[7] = in1[3];
    } else {
      float in0[4] = {io[0], io[2], io[4], io[6]};
      float in1[4] = {__fmul_rn(io[1], SQRT2F), __fadd_rn(io[3], io[1]), __fadd_rn(io[5], io[3]), __fadd_rn(io[7], io[5])};
      dct4(in0, false);
      dct4(in1, false);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float r = __fmul_rn(in1[i], sec[i]);
        io[i] = __fadd_rn(in0[i], r);
        io[7 - i] = __fsub_rn(in0[i], r);
      }
    }
  }
};

__device__ void dct1d(float* io, float* scratch, int n, bool forward) {
  switch (n) {
    case 1: return;
    case 2: {
      float t0 = __fadd_rn(io[0], io[1]), t1 = __fsub_rn(io[0], io[1]);
      if (forward) {
        io[0] = __fdiv_rn(t0, 2.0f);
        io[1] = __fdiv_rn(t1, 2.0f);
      } else {
        io[0] = t0;
        io[1] = t1;
      }
      return;
    }
    case 4: dct4(io, forward); return;
    case 8: Dct1D<8>::run(io, scratch, forward); return;
    case 16: Dct1D<16>::run(io, scratch, forward); return;
    case 32: Dct1D<32>::run(io, scratch, forward); return;
    case 64: Dct1D<64>::run(io, scratch, forward); return;
    case 128: Dct1D<128>::run(io, scratch, forward); return;
    default: Dct1D<256>::run(io, scratch, forward); return;
  }
}

struct Grid {
  float* p;
  int stride, w, h;
  __device__ __forceinline__ float& at(int x, int y) { return p[y * stride + x]; }
};

// dct_2d (generic/dct.rs:5-141), executed by a single thread on a small grid. `tmp` must hold
// 3 * max(w, h) floats.
__device__ void dct_2d_serial(Grid io, bool forward, float* tmp) {
  const int width = io.w, height = io.h;
  if (width * height <= 1) return;
  const float mul = forward ? 0.5f : 1.0f;
  if (width == 2 && height == 1) {
    float v0 = io.at(0, 0), v1 = io.at(1, 0);
    io.at(0, 0) = __fmul_rn(__fadd_rn(v0, v1), mul);
    io.at(1, 0) = __fmul_rn(__fsub_rn(v0, v1), mul);
    return;
  }
  if (width == 1 && height == 2) {
    float v0 = io.at(0, 0), v1 = io.at(0, 1);
    io.at(0, 0) = __fmul_rn(__fadd_rn(v0, v1), mul);
    io.at(0, 1) = __fmul_rn(__fsub_rn(v0, v1), mul);
    return;
  }
  if (width == 2 && height == 2) {
    float v00 = io.at(0, 0), v01 = io.at(1, 0), v10 = io.at(0, 1), v11 = io.at(1, 1);
    io.at(0, 0) = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v00, v01), v10), v11), mul), mul);
    io.at(1, 0) = __fmul_rn(__fmul_rn(__fsub_rn(__fadd_rn(__fsub_rn(v00, v01), v10), v11), mul), mul);
    io.at(0, 1) = __fmul_rn(__fmul_rn(__fsub_rn(__fsub_rn(__fadd_rn(v00, v01), v10), v11), mul), mul);
    io.at(1, 1) = __fmul_rn(__fmul_rn(__fadd_rn(__fsub_rn(__fsub_rn(v00, v01), v10), v11), mul), mul);
    return;
  }
  float* line = tmp;
  float* scratch = tmp + (width > height ? width : height);
  if (height == 1) {
    for (int x = 0; x < width; ++x) line[x] = io.at(x, 0);
    dct1d(line, scratch, width, forward);
    for (int x = 0; x < width; ++x) io.at(x, 0) = line[x];
    return;
  }
  if (width == 1) {
    for (int y = 0; y < height; ++y) line[y] = io.at(0, y);
    dct1d(line, scratch, height, forward);
    for (int y = 0; y < height; ++y) io.at(0, y) = line[y];
    return;
  }
  if (height == 2) {
    for (int x = 0; x < width; ++x) {
      float t0 = io.at(x, 0), t1 = io.at(x, 1);
      io.at(x, 0) = __fmul_rn(__fadd_rn(t0, t1), mul);
      io.at(x, 1) = __fmul_rn(__fsub_rn(t0, t1), mul);
    }
    for (int r = 0; r < 2; ++r) {
      for (int x = 0; x < width; ++x) line[x] = io.at(x, r);
      dct1d(line, scratch, width, forward);
      for (int x = 0; x < width; ++x) io.at(x, r) = line[x];
    }
    return;
  }
  if (width == 2) {
    for (int y = 0; y < height; ++y) {
      float v0 = io.at(0, y), v1 = io.at(1, y);
      io.at(0, y) = __fmul_rn(__fadd_rn(v0, v1), mul);
      io.at(1, y) = __fmul_rn(__fsub_rn(v0, v1), mul);
    }
    for (int c = 0; c < 2; ++c) {
      for (int y = 0; y < height; ++y) line[y] = io.at(c, y);
      dct1d(line, scratch, height, forward);
      for (int y = 0; y < height; ++y) io.at(c, y) = line[y];
    }
    return;
  }
  for (int y = 0; y < height; ++y) {
    for (int x = 0; x < width; ++x) line[x] = io.at(x, y);
    dct1d(line, scratch, width, forward);
    for (int x = 0; x < width; ++x) io.at(x, y) = line[x];
  }
  for (int x = 0; x < width; ++x) {
    for (int y = 0; y < height; ++y) line[y] = io.at(x, y);
    dct1d(line, scratch, height, forward);
    for (int y = 0; y < height; ++y) io.at(x, y) = line[y];
  }
}

// generic/transform.rs -------------------------------------------------------------------------
__device__ void aux_idct2(Grid b, int size, float* s /* size*size */) {
  const int n = size / 2;
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      float c00 = b.at(x, y), c01 = b.at(x + n, y), c10 = b.at(x, y + n), c11 = b.at(x + n, y + n);
      s[(2 * y) * size + 2 * x] = __fadd_rn(__fadd_rn(__fadd_rn(c00, c01), c10), c11);
      s[(2 * y) * size + 2 * x + 1] = __fsub_rn(__fsub_rn(__fadd_rn(c00, c01), c10), c11);
      s[(2 * y + 1) * size + 2 * x] = __fsub_rn(__fadd_rn(__fsub_rn(c00, c01), c10), c11);
      s[(2 * y + 1) * size + 2 * x + 1] = __fadd_rn(__fsub_rn(__fsub_rn(c00, c01), c10), c11);
    }
  for (int y = 0; y < size; ++y)
    for (int x = 0; x < size; ++x) b.at(x, y) = s[y * size + x];
}

// All 8x8 "special" transforms, executed by one thread on an 8x8 grid in shared memory.
__device__ void transform_special(Grid c, int type, float* scratch /* 64 + 48 floats */) {
  float* tmp = scratch + 64;
  if (type == 2) {  // Dct2
    aux_idct2(c, 2, scratch);
    aux_idct2(c, 4, scratch);
    aux_idct2(c, 8, scratch);
  } else if (type == 3) {  // Dct4
    aux_idct2(c, 2, scratch);
    for (int y = 0; y < 2; ++y)
      for (int x = 0; x < 2; ++x) {
        Grid s{scratch + (y * 2 + x) * 16, 4, 4, 4};
        for (int iy = 0; iy < 4; ++iy)
          for (int ix = 0; ix < 4; ++ix) s.at(iy, ix) = c.at(x + ix * 2, y + iy * 2);
      }
    for (int k = 0; k < 4; ++k) dct_2d_serial(Grid{scratch + k * 16, 4, 4, 4}, false, tmp);
    for (int y = 0; y < 2; ++y)
      for (int x = 0; x < 2; ++x)
        for (int iy = 0; iy < 4; ++iy)
          for (int ix = 0; ix < 4; ++ix) c.at(x * 4 + ix, y * 4 + iy) = scratch[(y * 2 + x) * 16 + iy * 4 + ix];
  } else if (type == 1) {  // Hornuss
    aux_idct2(c, 2, scratch);
    for (int y = 0; y < 2; ++y)
      for (int x = 0; x < 2; ++x) {
        float* s = scratch + (y * 2 + x) * 16;
        for (int iy = 0; iy < 4; ++iy)
          for (int ix = 0; ix < 4; ++ix) s[iy * 4 + ix] = c.at(x + ix * 2, y + iy * 2);
        float residual_sum = 0.0f;
        for (int i = 1; i < 16; ++i) residual_sum = __fadd_rn(residual_sum, s[i]);
        float avg = __fsub_rn(s[0], __fdiv_rn(residual_sum, 16.0f));
        s[0] = s[5];
        s[5] = 0.0f;
        for (int i = 0; i < 16; ++i) s[i] = __fadd_rn(s[i], avg);
      }
    for (int y = 0; y < 2; ++y)
      for (int x = 0; x < 2; ++x)
        for (int iy = 0; iy < 4; ++iy)
          for (int ix = 0; ix < 4; ++ix) c.at(x * 4 + ix, y * 4 + iy) = scratch[(y * 2 + x) * 16 + iy * 4 + ix];
  } else if (type == 12 || type == 13) {  // Dct4x8 / Dct8x4
    float coeff0 = c.at(0, 0), coeff1 = c.at(0, 1);
    c.at(0, 0) = __fadd_rn(coeff0, coeff1);
    c.at(0, 1) = __fsub_rn(coeff0, coeff1);
    for (int idx = 0; idx < 2; ++idx) {
      Grid s{scratch + idx * 32, 8, 8, 4};
      for (int iy = 0; iy < 4; ++iy)
        for (int ix = 0; ix < 8; ++ix) s.at(ix, iy) = c.at(ix, iy * 2 + idx);
      dct_2d_serial(s, false, tmp);
    }
    if (type == 13) {
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) c.at(y, x) = scratch[y * 8 + x];
    } else {
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) c.at(x, y) = scratch[y * 8 + x];
    }
  } else {  // Afv0..3
    const int n = type - 14;
    const int flip_x = n % 2, flip_y = n / 2;
    float* coeff_afv = scratch;        // 16
    float* samples_afv = scratch + 16; // 16
    float* s4x4 = scratch + 32;        // 16
    float* s4x8 = scratch + 48;        // 32
    float* tmp2 = scratch + 80;        // 24 (3 * 8)
    coeff_afv[0] = __fmul_rn(__fadd_rn(__fadd_rn(c.at(0, 0), c.at(1, 0)), c.at(0, 1)), 4.0f);
    for (int idx = 1; idx < 16; ++idx) coeff_afv[idx] = c.at(2 * (idx % 4), 2 * (idx / 4));
    for (int j = 0; j < 16; ++j) samples_afv[j] = 0.0f;
    for (int i = 0; i < 16; ++i)
      for (int j = 0; j < 16; ++j) samples_afv[j] = __fmaf_rn(coeff_afv[i], kAfvBasis[i][j], samples_afv[j]);
    for (int i = 0; i < 16; ++i) s4x4[i] = 0.0f;
    for (int i = 0; i < 32; ++i) s4x8[i] = 0.0f;
    s4x4[0] = __fadd_rn(__fsub_rn(c.at(0, 0), c.at(1, 0)), c.at(0, 1));
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 4; ++ix) {
        if ((ix | iy) == 0) continue;
        s4x4[ix * 4 + iy] = c.at(2 * ix + 1, 2 * iy);
      }
    dct_2d_serial(Grid{s4x4, 4, 4, 4}, false, tmp2);
    s4x8[0] = __fsub_rn(c.at(0, 0), c.at(0, 1));
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 8; ++ix) {
        if ((ix | iy) == 0) continue;
        s4x8[iy * 8 + ix] = c.at(ix, 2 * iy + 1);
      }
    dct_2d_serial(Grid{s4x8, 8, 8, 4}, false, tmp2);
    for (int iy = 0; iy < 4; ++iy) {
      int afv_y = flip_y == 0 ? iy : 3 - iy;
      for (int ix = 0; ix < 4; ++ix) {
        int afv_x = flip_x == 0 ? ix : 3 - ix;
        c.at(flip_x * 4 + ix, flip_y * 4 + iy) = samples_afv[afv_y * 4 + afv_x];
      }
    }
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 4; ++ix) c.at((1 - flip_x) * 4 + ix, flip_y * 4 + iy) = s4x4[iy * 4 + ix];
    for (int iy = 0; iy < 4; ++iy)
      for (int ix = 0; ix < 8; ++ix) c.at(ix, (1 - flip_y) * 4 + iy) = s4x8[iy * 8 + ix];
  }
}

// transform_varblocks_inner (transform_common.rs:11-75): one CTA per (8x8 cell, channel); cells
// that are not a varblock origin exit immediately. First, correctness-oriented version: rows then
// columns straight on the coefficient plane (L1/L2 resident), one thread per line.
constexpr int kTransformThreads = 64;
__global__ void __launch_bounds__(kTransformThreads) hf_transform_kernel(DevFrame f) {
  const uint32_t bx = blockIdx.x, by = blockIdx.y, c = blockIdx.z;
  const int32_t t = f.blk_type[size_t(by) * f.bw + bx];
  if (t < 0) return;
  const int bw = kDevTransformInfo[t][0], bh = kDevTransformInfo[t][1];
  const int w = bw * 8, h = bh * 8;
  float* plane = reinterpret_cast<float*>(f.coeff[c]);
  float* block = plane + size_t(by) * 8 * f.cw + size_t(bx) * 8;
  const int stride = int(f.cw);
  __shared__ float smem[1024 + 3 * 32 + 160];
  float* llf = smem;            // up to 32 x 32
  float* tmp = smem + 1024;     // 96
  float* special = smem + 1024 + 96;  // 160
  const float* lf = f.lf[c];
  if (threadIdx.x == 0) {
    if (bw * bh == 1) {
      llf[0] = lf[size_t(by) * f.bw + bx];
    } else {
      for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x) llf[y * bw + x] = lf[size_t(by + y) * f.bw + bx + x];
      dct_2d_serial(Grid{llf, bw, bw, bh}, true, tmp);
      int logbw = 31 - __clz(bw), logbh = 31 - __clz(bh);
      for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x)
          llf[y * bw + x] = __fdiv_rn(llf[y * bw + x], __fmul_rn(kScaleF[y << (5 - logbh)], kScaleF[x << (5 - logbw)]));
    }
  }
  __syncthreads();
  const bool is_special = (t == 1 || t == 2 || t == 3 || (t >= 12 && t <= 17));
  if (is_special) {
    if (threadIdx.x == 0) {
      float* g = special;  // 8x8 copy + scratch behind it
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) g[y * 8 + x] = block[y * stride + x];
      g[0] = llf[0];
      __shared__ float sscratch[128];
      transform_special(Grid{g, 8, 8, 8}, t, sscratch);
      for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) block[y * stride + x] = g[y * 8 + x];
    }
    return;
  }
  // rows
  float line[256];
  float scratch[256];
  for (int r = threadIdx.x; r < h; r += kTransformThreads) {
    float* row = block + size_t(r) * stride;
    for (int x = 0; x < w; ++x) line[x] = row[x];
    if (r < bh)
      for (int x = 0; x < bw; ++x) line[x] = llf[r * bw + x];
    dct1d(line, scratch, w, false);
    for (int x = 0; x < w; ++x) row[x] = line[x];
  }
  __syncthreads();
  for (int col = threadIdx.x; col < w; col += kTransformThreads) {
    float* cp = block + col;
    for (int y = 0; y < h; ++y) line[y] = cp[size_t(y) * stride];
    dct1d(line, scratch, h, false);
    for (int y = 0; y < h; ++y) cp[size_t(y) * stride] = line[y];
  }
}

}  // namespace

// First (single-thread, global-memory) version; superseded by kernels/blockinfo.cu.
void launch_build_block_info_v1(DevFrame f, const DevBlockInfoJob* jobs, int num_jobs, float quant_mul_base,
                                const float* sharp_lut8, int has_epf, int* status, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  build_block_info_kernel<<<(num_jobs + 31) / 32, 32, 0, stream>>>(f, jobs, num_jobs, quant_mul_base, sharp_lut8, has_epf, status);
}

// First (unoptimised, global-memory) version; kept as a debugging reference for entropy.cu.
void launch_decode_hf_v1(const uint8_t* cs, DevFrame f, DevHfParams p, const DevHfJob* jobs, uint64_t* end_bits, int* status,
                         int num_jobs, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  decode_hf_kernel<<<num_jobs, 32, 0, stream>>>(cs, f, p, jobs, end_bits, status, num_jobs);
}

void launch_lf_dequant(DevFrame f, const DevLfDequantJob* jobs, int num_jobs, cudaStream_t stream) {
  if (num_jobs <= 0) return;
  dim3 grid(2, 256, num_jobs);  // LF groups are at most 256 x 256 blocks
  lf_dequant_kernel<<<grid, 128, 0, stream>>>(f, jobs);
}

void launch_lf_cfl(DevFrame f, float kx, float kb, cudaStream_t stream) {
  size_t n = size_t(f.bw) * f.bh;
  lf_cfl_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(f, kx, kb);
}

void launch_lf_smooth(DevFrame f, float* tmp[3], float lf_x, float lf_y, float lf_b, cudaStream_t stream) {
  dim3 grid((f.bw + 127) / 128, f.bh);
  lf_smooth_kernel<<<grid, 128, 0, stream>>>(f, tmp[0], tmp[1], tmp[2], lf_x, lf_y, lf_b);
}

void launch_hf_dequant_cfl(DevFrame f, DevDequantParams p, cudaStream_t stream) {
  dim3 block(64, 4);
  dim3 grid((f.cw + 63) / 64, (f.ch + 3) / 4);
  hf_dequant_cfl_kernel<<<grid, block, 0, stream>>>(f, p);
}

void launch_hf_transform(DevFrame f, const float*, float*, cudaStream_t stream) {
  dim3 grid(f.bw, f.bh, 3);
  hf_transform_kernel<<<grid, kTransformThreads, 0, stream>>>(f);
}

}  // namespace jxlb
