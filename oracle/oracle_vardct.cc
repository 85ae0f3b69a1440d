// TEST INFRASTRUCTURE — see oracle_backend.h.
// VarDCT stages: HfMetadata post-processing, HF coefficient decode, LF dequant / chroma-from-luma /
// adaptive smoothing, HF dequant + chroma-from-luma, LLF insertion and the 27 inverse transforms.
// Restates crates/jxl-vardct/src/{hf_metadata.rs,hf_coeff.rs} and
// crates/jxl-render/src/vardct/{mod.rs,transform_common.rs,dct_common.rs,generic/*.rs}
// (the *generic* float code path; compile with -ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "oracle_backend.h"
#include "oracle_tables.h"

namespace jxlo {

// ---------------------------------------------------------------------------------------------
// HfMetadata::parse placement scan (hf_metadata.rs:99-230)
void OracleBackend::build_block_info(VarDctState& st, const std::vector<BlockInfoJob>& jobs) {
  Plane& type = plane(st.blk_type);
  Plane& mul = plane(st.blk_mul);
  Plane& sig = plane(st.epf_sigma);
  Plane& sharp = plane(st.sharpness);
  const EpfParams& epf = st.fh->restoration_filter.epf;
  const bool has_epf = epf.iters > 0;
  for (const BlockInfoJob& job : jobs) {
    const LfGroupRect& rc = job.rect;
    Plane& raw = plane(job.raw_plane);
    for (uint32_t y = 0; y < rc.bh; ++y)
      for (uint32_t x = 0; x < rc.bw; ++x) type.i32()[size_t(rc.by0 + y) * type.w + rc.bx0 + x] = INT32_MIN;  // Uninit
    const float quant_mul_base = epf.quant_mul * 65536.0f / float(st.lfg->global_scale);
    uint32_t data_idx = 0;
    for (uint32_t y = 0; y < rc.bh; ++y) {
      for (uint32_t x = 0; x < rc.bw;) {
        int32_t& cell = type.i32()[size_t(rc.by0 + y) * type.w + rc.bx0 + x];
        if (cell != INT32_MIN) {
          ++x;
          continue;
        }
        JXLB_CHECK(data_idx < job.nb_blocks, kErrBitstream, "BlockInfo doesn't fill LF group");
        int32_t dct_select = raw.i32()[data_idx];
        JXLB_CHECK(dct_select >= 0 && dct_select < kNumTransformTypes, kErrBitstream, "invalid dct_select");
        int32_t hf_mul = raw.i32()[raw.w + data_idx] + 1;
        JXLB_CHECK(hf_mul > 0, kErrBitstream, "non-positive HfMul");
        uint32_t dw = kTransformInfo[dct_select].w8, dh = kTransformInfo[dct_select].h8;
        JXLB_CHECK((x % 32) + dw <= 32 && (y % 32) + dh <= 32, kErrBitstream, "varblock crosses group border");
        float sigma_q = quant_mul_base / float(hf_mul);
        for (uint32_t dy = 0; dy < dh; ++dy)
          for (uint32_t dx = 0; dx < dw; ++dx) {
            JXLB_CHECK(x + dx < rc.bw && y + dy < rc.bh, kErrBitstream, "varblock doesn't fit in LF group");
            size_t gi = size_t(rc.by0 + y + dy) * type.w + rc.bx0 + x + dx;
            JXLB_CHECK(type.i32()[gi] == INT32_MIN, kErrBitstream, "varblocks overlap");
            type.i32()[gi] = (dx == 0 && dy == 0) ? dct_select : -int32_t(1 + dx + 32 * dy);
            mul.i32()[gi] = hf_mul;
            if (has_epf) {
              int32_t s = sharp.i32()[gi];
              JXLB_CHECK(s >= 0 && s < 8, kErrBitstream, "invalid EPF sharpness value");
              sig.f32()[gi] = sigma_q * epf.sharp_lut[s];
            }
          }
        ++data_idx;
        x += dw;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// write_hf_coeff (hf_coeff.rs:21-252)
void OracleBackend::decode_hf(VarDctState& st, std::vector<HfGroupJob>& jobs) {
  parallel_for(jobs.size(), [&](size_t i) { decode_one_hf(st, jobs[i]); });
}

void OracleBackend::decode_one_hf(VarDctState& st, HfGroupJob& job) {
  const HfBlockContext& hbc = st.lfg->hf_block_ctx;
  const HfGlobalSyntax& hfg = *st.hfg;
  const HfPassSyntax& pass = hfg.passes[job.pass_idx];
  const uint32_t nbc = hbc.num_block_clusters;
  const uint32_t coeff_shift = job.pass_idx < st.fh->passes.shift.size() ? st.fh->passes.shift[job.pass_idx] : 0;
  const size_t lf_idx_mul = (hbc.lf_thresholds[0].size() + 1) * (hbc.lf_thresholds[1].size() + 1) * (hbc.lf_thresholds[2].size() + 1);
  const size_t hf_idx_mul = hbc.qf_thresholds.size() + 1;

  BitReader br(cs_, job.bit_limit / 8, job.bit_pos);
  uint32_t hfp = br.read(ceil_log2_nonzero(hfg.num_hf_presets));
  JXLB_CHECK(hfp < hfg.num_hf_presets, kErrBitstream, "selected HF preset out of bounds");
  const uint32_t ctx_size = 495 * nbc;
  const uint8_t* cluster_map = pass.code.cluster_map.data() + size_t(ctx_size) * hfp;
  EntropyReader dec(&pass.code);
  dec.begin(br);

  const uint32_t gx = job.group_idx % st.groups_per_row, gy = job.group_idx / st.groups_per_row;
  const uint32_t gb = st.group_dim / 8;
  const uint32_t bx0 = gx * gb, by0 = gy * gb;
  const uint32_t width = std::min(gb, st.bw - bx0), height = std::min(gb, st.bh - by0);
  Plane& type = plane(st.blk_type);
  Plane& mul = plane(st.blk_mul);
  Plane* lfq[3] = {&plane(st.lf_quant[0]), &plane(st.lf_quant[1]), &plane(st.lf_quant[2])};
  Plane* coeff[3] = {&plane(st.coeff[0]), &plane(st.coeff[1]), &plane(st.coeff[2])};
  std::vector<uint32_t> nz_row[3];
  for (auto& v : nz_row) v.assign(width, 0);
  const uint32_t* hs = st.hshift;
  const uint32_t* vs = st.vshift;

  for (uint32_t y = 0; y < height; ++y)
    for (uint32_t x = 0; x < width; ++x) {
      size_t gi = size_t(by0 + y) * type.w + bx0 + x;
      int32_t t = type.i32()[gi];
      if (t < 0) continue;
      const TransformTypeInfo& ti = kTransformInfo[t];
      int32_t qf = mul.i32()[gi];
      uint32_t w8 = ti.w8, h8 = ti.h8;
      uint32_t num_blocks = w8 * h8, num_blocks_log = ceil_log2_nonzero(num_blocks);
      uint32_t order_id = ti.order_id;
      size_t lf_idx = 0;
      if (!st.use_lf_frame) {
        for (int c : {0, 2, 1}) {
          const auto& thr = hbc.lf_thresholds[c];
          lf_idx *= thr.size() + 1;
          int32_t q = lfq[c]->i32()[size_t((by0 + y) >> vs[c]) * type.w + ((bx0 + x) >> hs[c])];
          for (int32_t th : thr)
            if (q > th) ++lf_idx;
        }
      }
      size_t hf_idx = 0;
      for (uint32_t th : hbc.qf_thresholds)
        if (qf > int32_t(th)) ++hf_idx;
      for (int ci = 0; ci < 3; ++ci) {
        size_t ch_idx = size_t(ci) * 13 + order_id;
        int c = (ci == 0) ? 1 : (ci == 1 ? 0 : 2);  // y, x, b
        // a subsampled channel codes only the blocks aligned to its grid, at the shifted position (hf_coeff.rs:143-155)
        const uint32_t sx = x >> hs[c], sy = y >> vs[c];
        if (hs[c] || vs[c]) {
          if ((sx << hs[c]) != x || (sy << vs[c]) != y) continue;
          if (type.i32()[size_t(by0 + sy) * type.w + bx0 + sx] < 0) continue;
          JXLB_CHECK(num_blocks == 1, kErrUnsupported, "chroma subsampling with varblocks larger than 8x8 is not supported");
        }
        size_t idx = (ch_idx * hf_idx_mul + hf_idx) * lf_idx_mul + lf_idx;
        JXLB_CHECK(idx < hbc.block_ctx_map.size(), kErrBitstream, "block context out of range");
        uint32_t block_ctx = hbc.block_ctx_map[idx];
        uint32_t predicted;
        if (sy == 0) predicted = sx == 0 ? 32 : nz_row[c][sx - 1];
        else if (sx == 0) predicted = nz_row[c][sx];
        else predicted = (nz_row[c][sx] + nz_row[c][sx - 1] + 1) >> 1;
        uint32_t pidx = predicted >= 8 ? 4 + predicted / 2 : predicted;
        uint32_t nz_ctx = block_ctx + pidx * nbc;
        uint32_t non_zeros = dec.read_varint_clustered(br, cluster_map[nz_ctx], 0);
        JXLB_CHECK(non_zeros <= (63u << num_blocks_log), kErrBitstream, "non_zeros too large");
        uint32_t nz_val = (non_zeros + num_blocks - 1) >> num_blocks_log;
        for (uint32_t dx = 0; dx < w8; ++dx) nz_row[c][sx + dx] = nz_val;
        if (non_zeros == 0) continue;
        uint32_t prev_nonzero = (non_zeros <= num_blocks * 4) ? 1 : 0;
        const std::vector<uint32_t>& custom = pass.order[order_id][c];
        const std::vector<uint32_t>& order = custom.empty() ? natural_order_cached(order_id) : custom;
        const uint32_t coeff_ctx_base = block_ctx * 458 + 37 * nbc;
        const uint8_t* cmap = cluster_map + coeff_ctx_base;
        Plane& cp = *coeff[c];
        for (size_t k = num_blocks, i = 0; k < order.size(); ++k, ++i) {
          uint32_t nzc = (non_zeros - 1) >> num_blocks_log;
          uint32_t fi = uint32_t(i >> num_blocks_log);
          uint32_t cctx = (kCoeffNumNonzeroContext[nzc] + kCoeffFreqContext[fi]) * 2 + prev_nonzero;
          JXLB_CHECK(cctx < 458, kErrBitstream, "too many zeros in varblock HF coefficient");
          uint32_t ucoeff = dec.read_varint_clustered(br, cmap[cctx], 0);
          if (ucoeff == 0) {
            prev_nonzero = 0;
            continue;
          }
          int32_t cv = int32_t(uint32_t(unpack_signed(ucoeff)) << coeff_shift);
          uint32_t dx = order[k] & 0xffff, dy = order[k] >> 16;
          if (ti.transpose) std::swap(dx, dy);
          size_t px = size_t((bx0 >> hs[c]) + sx) * 8 + dx, py = size_t((by0 >> vs[c]) + sy) * 8 + dy;
          int32_t& dst = cp.i32()[py * cp.w + px];
          dst = int32_t(uint32_t(dst) + uint32_t(cv));
          prev_nonzero = 1;
          if (--non_zeros == 0) break;
        }
        JXLB_CHECK(!br.overrun(), kErrEof, "HF stream truncated");
      }
    }
  JXLB_CHECK(dec.finalize_ok(), kErrBitstream, "invalid ANS final state (HF coefficients)");
  JXLB_CHECK(!br.overrun(), kErrEof, "HF stream truncated");
  job.end_bit = br.pos();
}

// ---------------------------------------------------------------------------------------------
// LF (vardct/mod.rs:387-412, 544-568; generic/mod.rs:11-103)
void OracleBackend::lf_dequant(VarDctState& st, const std::vector<LfDequantJob>& jobs) {
  for (const LfDequantJob& j : jobs)
    for (int c = 0; c < 3; ++c) {
      Plane& q = plane(st.lf_quant[c]);
      Plane& o = plane(st.lf[c]);
      const LfGroupRect rc = shifted_rect(j.rect, st.hshift[c], st.vshift[c]);
      for (uint32_t y = 0; y < rc.bh; ++y)
        for (uint32_t x = 0; x < rc.bw; ++x) {
          size_t i = size_t(rc.by0 + y) * q.w + rc.bx0 + x;
          o.f32()[i] = float(q.i32()[i]) * j.scale[c];
        }
    }
}

void OracleBackend::lf_chroma_from_luma(VarDctState& st) {
  const LfGlobalSyntax& g = *st.lfg;
  int32_t x_factor = int32_t(g.x_factor_lf) - 128, b_factor = int32_t(g.b_factor_lf) - 128;
  float kx = g.base_correlation_x + (float(x_factor) / float(g.colour_factor));
  float kb = g.base_correlation_b + (float(b_factor) / float(g.colour_factor));
  float *x = plane(st.lf[0]).f32(), *y = plane(st.lf[1]).f32(), *b = plane(st.lf[2]).f32();
  size_t n = size_t(st.bw) * st.bh;
  for (size_t i = 0; i < n; ++i) {
    float yy = y[i];
    x[i] += kx * yy;
    b[i] += kb * yy;
  }
}

void OracleBackend::lf_adaptive_smoothing(VarDctState& st) {
  const LfGlobalSyntax& g = *st.lfg;
  uint64_t scale_inv = uint64_t(g.global_scale) * g.quant_lf;
  const float lf_scale[3] = {float(512.0 * double(g.m_x_lf) / double(scale_inv)),
                             float(512.0 * double(g.m_y_lf) / double(scale_inv)),
                             float(512.0 * double(g.m_b_lf) / double(scale_inv))};
  const size_t width = st.bw, height = st.bh;
  if (width <= 2 || height <= 2) return;
  const float kSelf = 0.052262735f, kSide = 0.2034514f, kDiag = 0.03348292f;
  float* in[3] = {plane(st.lf[0]).f32(), plane(st.lf[1]).f32(), plane(st.lf[2]).f32()};
  std::vector<float> udsum[3];
  for (int c = 0; c < 3; ++c) {
    udsum[c].resize(width * (height - 2));
    for (size_t y = 0; y + 2 < height; ++y)
      for (size_t x = 0; x < width; ++x) udsum[c][y * width + x] = in[c][y * width + x] + in[c][(y + 2) * width + x];
  }
  for (size_t y = 1; y + 1 < height; ++y) {
    float* row[3] = {in[0] + y * width, in[1] + y * width, in[2] + y * width};
    const float* ud[3] = {&udsum[0][(y - 1) * width], &udsum[1][(y - 1) * width], &udsum[2][(y - 1) * width]};
    float prev[3] = {row[0][0], row[1][0], row[2][0]};
    for (size_t x = 1; x + 1 < width; ++x) {
      float self[3], wa[3], gap_t[3];
      for (int c = 0; c < 3; ++c) {
        self[c] = row[c][x];
        float side = prev[c] + row[c][x + 1] + ud[c][x];
        float diag = ud[c][x - 1] + ud[c][x + 1];
        wa[c] = self[c] * kSelf + side * kSide + diag * kDiag;
        gap_t[c] = std::fabs(wa[c] - self[c]) / lf_scale[c];
      }
      float gap = std::max(std::max(std::max(0.5f, gap_t[0]), gap_t[1]), gap_t[2]);
      float gap_scale = std::max(3.0f - 4.0f * gap, 0.0f);
      for (int c = 0; c < 3; ++c) {
        row[c][x] = (wa[c] - self[c]) * gap_scale + self[c];
        prev[c] = self[c];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dequant_hf_varblock_grouped + chroma_from_luma_hf_grouped (vardct/mod.rs:442-542, 570-603)
namespace {
// for_each_varblocks (vardct/mod.rs:693-730): where channel c keeps the varblock that starts at (bx, by), or false
// when a subsampled channel skips it. The second look-up is group-local, like the reference's.
bool shifted_block(const VarDctState& st, Plane& type, int c, uint32_t bx, uint32_t by, size_t* sbx, size_t* sby) {
  const uint32_t hs = st.hshift[c], vs = st.vshift[c];
  if (!hs && !vs) return true;
  const uint32_t gb = st.group_dim / 8;
  const uint32_t gx0 = bx / gb * gb, gy0 = by / gb * gb;
  const uint32_t lx = bx - gx0, ly = by - gy0;
  if (((lx >> hs) << hs) != lx || ((ly >> vs) << vs) != ly) return false;
  if (type.i32()[size_t(gy0 + (ly >> vs)) * type.w + gx0 + (lx >> hs)] < 0) return false;
  *sbx = (gx0 >> hs) + (lx >> hs);
  *sby = (gy0 >> vs) + (ly >> vs);
  return true;
}
}  // namespace

void OracleBackend::hf_dequant_cfl(VarDctState& st) {
  const OpsinInverseMatrix& oim = st.ih->opsin_inverse_matrix;
  const LfGlobalSyntax& g = *st.lfg;
  const float qm_scale[3] = {powi_f32(0.8f, int32_t(st.fh->x_qm_scale) - 2), 1.0f, powi_f32(0.8f, int32_t(st.fh->b_qm_scale) - 2)};
  Plane& type = plane(st.blk_type);
  Plane& mulp = plane(st.blk_mul);
  for (int c = 0; c < 3; ++c) {
    Plane& cp = plane(st.coeff[c]);
    const float quant_bias = oim.quant_bias[c];
    parallel_for(st.bh, [&](size_t by) {
      for (uint32_t bx = 0; bx < st.bw; ++bx) {
        size_t gi = by * type.w + bx;
        int32_t t = type.i32()[gi];
        if (t < 0) continue;
        size_t sbx = bx, sby = by;
        if (!shifted_block(st, type, c, bx, uint32_t(by), &sbx, &sby)) continue;
        const TransformTypeInfo& ti = kTransformInfo[t];
        uint32_t w = ti.w8 * 8u, h = ti.h8 * 8u;
        float mul = 65536.0f / (float(g.global_scale) * float(mulp.i32()[gi])) * qm_scale[c];
        const std::vector<float>& m = ti.transpose ? st.hfg->dequant->matrices_tr[ti.param_index][c]
                                                   : st.hfg->dequant->matrices[ti.param_index][c];
        for (uint32_t y = 0; y < h; ++y) {
          uint32_t* row = cp.data.data() + (sby * 8 + y) * cp.w + sbx * 8;
          for (uint32_t x = 0; x < w; ++x) {
            float q = float(int32_t(row[x]));
            if (std::fabs(q) <= 1.0f) q *= quant_bias;
            else q -= oim.quant_bias_numerator / q;
            q *= m[size_t(y) * w + x];
            q *= mul;
            std::memcpy(&row[x], &q, 4);
          }
        }
      }
    });
  }
  if (st.subsampled) return;  // no chroma from luma between planes of different sizes (vardct/mod.rs:353)
  // chroma from luma on coefficients, per 64x64 tile
  Plane& xfy = plane(st.x_from_y);
  Plane& bfy = plane(st.b_from_y);
  float* cx = plane(st.coeff[0]).f32();
  const float* cy = plane(st.coeff[1]).f32();
  float* cb = plane(st.coeff[2]).f32();
  const size_t W = size_t(st.bw) * 8, H = size_t(st.bh) * 8;
  parallel_for(H, [&](size_t y) {
    for (size_t x = 0; x < W; ++x) {
      size_t ti = (y / 64) * xfy.w + x / 64;
      float kx = g.base_correlation_x + (float(xfy.i32()[ti]) / float(g.colour_factor));
      float kb = g.base_correlation_b + (float(bfy.i32()[ti]) / float(g.colour_factor));
      float yy = cy[y * W + x];
      cx[y * W + x] += kx * yy;
      cb[y * W + x] += kb * yy;
    }
  });
}

// ---------------------------------------------------------------------------------------------
// DCT (generic/dct.rs)
namespace {

const float kSqrt2 = 1.41421356237309504880f;

void dct4(float* io, bool forward) {  // generic/dct.rs:143-172
  const float sec0 = 0.5411961f, sec1 = 1.306563f;
  float i0 = io[0], i1 = io[1], i2 = io[2], i3 = io[3];
  if (forward) {
    float sum03 = i0 + i3, sum12 = i1 + i2;
    float tmp0 = (i0 - i3) * sec0, tmp1 = (i1 - i2) * sec1;
    float out0 = (tmp0 + tmp1) / 4.0f, out1 = (tmp0 - tmp1) / 4.0f;
    io[0] = (sum03 + sum12) / 4.0f;
    io[1] = out0 * kSqrt2 + out1;
    io[2] = (sum03 - sum12) / 4.0f;
    io[3] = out1;
  } else {
    float tmp0 = i1 * kSqrt2, tmp1 = i1 + i3;
    float out0 = (tmp0 + tmp1) * sec0, out1 = (tmp0 - tmp1) * sec1;
    float sum02 = i0 + i2, sub02 = i0 - i2;
    io[0] = sum02 + out0;
    io[1] = sub02 + out1;
    io[2] = sub02 - out1;
    io[3] = sum02 - out0;
  }
}

void dct1d(float* io, float* scratch, size_t n, bool forward) {  // generic/dct.rs:174-293
  if (n <= 1) return;
  if (n == 2) {
    float t0 = io[0] + io[1], t1 = io[0] - io[1];
    if (forward) {
      io[0] = t0 / 2.0f;
      io[1] = t1 / 2.0f;
    } else {
      io[0] = t0;
      io[1] = t1;
    }
    return;
  }
  if (n == 4) {
    dct4(io, forward);
    return;
  }
  if (n == 8) {
    const float* sec = sec_half(8);
    if (forward) {
      float in0[4] = {(io[0] + io[7]) / 2.0f, (io[1] + io[6]) / 2.0f, (io[2] + io[5]) / 2.0f, (io[3] + io[4]) / 2.0f};
      float in1[4] = {(io[0] - io[7]) * sec[0] / 2.0f, (io[1] - io[6]) * sec[1] / 2.0f, (io[2] - io[5]) * sec[2] / 2.0f,
                      (io[3] - io[4]) * sec[3] / 2.0f};
      dct4(in0, true);
      for (int i = 0; i < 4; ++i) io[i * 2] = in0[i];
      dct4(in1, true);
      in1[0] *= kSqrt2;
      for (int i = 0; i < 3; ++i) io[i * 2 + 1] = in1[i] + in1[i + 1];
      io[7] = in1[3];
    } else {
      float in0[4] = {io[0], io[2], io[4], io[6]};
      float in1[4] = {io[1] * kSqrt2, io[3] + io[1], io[5] + io[3], io[7] + io[5]};
      dct4(in0, false);
      dct4(in1, false);
      for (int i = 0; i < 4; ++i) {
        float r = in1[i] * sec[i];
        io[i] = in0[i] + r;
        io[7 - i] = in0[i] - r;
      }
    }
    return;
  }
  const size_t h = n / 2;
  float* in0 = scratch;
  float* in1 = scratch + h;
  const float* sec = sec_half(n);
  if (forward) {
    for (size_t i = 0; i < h; ++i) {
      in0[i] = (io[i] + io[n - i - 1]) / 2.0f;
      in1[i] = (io[i] - io[n - i - 1]) / 2.0f;
    }
    for (size_t i = 0; i < h; ++i) in1[i] *= sec[i];
    dct1d(in0, io, h, true);
    dct1d(in1, io + h, h, true);
    in1[0] *= kSqrt2;
    for (size_t i = 0; i + 1 < h; ++i) in1[i] += in1[i + 1];
    for (size_t i = 0; i < h; ++i) io[i * 2] = in0[i];
    for (size_t i = 0; i < h; ++i) io[i * 2 + 1] = in1[i];
  } else {
    for (size_t i = 0; i < h; ++i) {
      in0[i] = io[i * 2];
      in1[i] = io[i * 2 + 1];
    }
    for (size_t i = 1; i < h; ++i) in1[h - i] += in1[h - i - 1];
    in1[0] *= kSqrt2;
    dct1d(in0, io, h, false);
    dct1d(in1, io + h, h, false);
    for (size_t i = 0; i < h; ++i) in1[i] *= sec[i];
    for (size_t i = 0; i < h; ++i) {
      io[i] = scratch[i] + scratch[i + h];
      io[n - i - 1] = scratch[i] - scratch[i + h];
    }
  }
}

struct Grid {
  float* p;
  size_t stride, w, h;
  float& at(size_t x, size_t y) { return p[y * stride + x]; }
};

void dct_2d(Grid io, bool forward) {  // generic/dct.rs:5-141
  const size_t width = io.w, height = io.h;
  if (width * height <= 1) return;
  const float mul = forward ? 0.5f : 1.0f;
  if (width == 2 && height == 1) {
    float v0 = io.at(0, 0), v1 = io.at(1, 0);
    io.at(0, 0) = (v0 + v1) * mul;
    io.at(1, 0) = (v0 - v1) * mul;
    return;
  }
  if (width == 1 && height == 2) {
    float v0 = io.at(0, 0), v1 = io.at(0, 1);
    io.at(0, 0) = (v0 + v1) * mul;
    io.at(0, 1) = (v0 - v1) * mul;
    return;
  }
  if (width == 2 && height == 2) {
    float v00 = io.at(0, 0), v01 = io.at(1, 0), v10 = io.at(0, 1), v11 = io.at(1, 1);
    io.at(0, 0) = (v00 + v01 + v10 + v11) * mul * mul;
    io.at(1, 0) = (v00 - v01 + v10 - v11) * mul * mul;
    io.at(0, 1) = (v00 + v01 - v10 - v11) * mul * mul;
    io.at(1, 1) = (v00 - v01 - v10 + v11) * mul * mul;
    return;
  }
  std::vector<float> buf(std::max(width, height));
  if (height == 1) {
    dct1d(&io.at(0, 0), buf.data(), width, forward);
    return;
  }
  if (width == 1) {
    std::vector<float> row(height);
    for (size_t y = 0; y < height; ++y) row[y] = io.at(0, y);
    dct1d(row.data(), buf.data(), height, forward);
    for (size_t y = 0; y < height; ++y) io.at(0, y) = row[y];
    return;
  }
  if (height == 2) {
    for (size_t x = 0; x < width; ++x) {
      float t0 = io.at(x, 0), t1 = io.at(x, 1);
      io.at(x, 0) = (t0 + t1) * mul;
      io.at(x, 1) = (t0 - t1) * mul;
    }
    dct1d(&io.at(0, 0), buf.data(), width, forward);
    dct1d(&io.at(0, 1), buf.data(), width, forward);
    return;
  }
  if (width == 2) {
    std::vector<float> row(height * 2);
    float *r0 = row.data(), *r1 = row.data() + height;
    for (size_t y = 0; y < height; ++y) {
      float v0 = io.at(0, y), v1 = io.at(1, y);
      r0[y] = (v0 + v1) * mul;
      r1[y] = (v0 - v1) * mul;
    }
    dct1d(r0, buf.data(), height, forward);
    dct1d(r1, buf.data(), height, forward);
    for (size_t y = 0; y < height; ++y) {
      io.at(0, y) = r0[y];
      io.at(1, y) = r1[y];
    }
    return;
  }
  for (size_t y = 0; y < height; ++y) dct1d(&io.at(0, y), buf.data(), width, forward);
  // column pass: the reference transposes min(w,h)-square sub-blocks, runs row DCTs and
  // transposes back, which is the 1-D transform of every column.
  std::vector<float> col(height);
  for (size_t x = 0; x < width; ++x) {
    for (size_t y = 0; y < height; ++y) col[y] = io.at(x, y);
    dct1d(col.data(), buf.data(), height, forward);
    for (size_t y = 0; y < height; ++y) io.at(x, y) = col[y];
  }
}

// generic/transform.rs ------------------------------------------------------------------------
template <size_t SIZE>
void aux_idct2_in_place(Grid b) {  // transform.rs:28-48
  const size_t n = SIZE / 2;
  float s[SIZE][SIZE];
  for (size_t y = 0; y < n; ++y)
    for (size_t x = 0; x < n; ++x) {
      float c00 = b.at(x, y), c01 = b.at(x + n, y), c10 = b.at(x, y + n), c11 = b.at(x + n, y + n);
      s[2 * y][2 * x] = c00 + c01 + c10 + c11;
      s[2 * y][2 * x + 1] = c00 + c01 - c10 - c11;
      s[2 * y + 1][2 * x] = c00 - c01 + c10 - c11;
      s[2 * y + 1][2 * x + 1] = c00 - c01 - c10 + c11;
    }
  for (size_t y = 0; y < SIZE; ++y)
    for (size_t x = 0; x < SIZE; ++x) b.at(x, y) = s[y][x];
}

void transform_dct2(Grid c) {
  aux_idct2_in_place<2>(c);
  aux_idct2_in_place<4>(c);
  aux_idct2_in_place<8>(c);
}

void transform_dct4(Grid c) {  // transform.rs:56-82
  aux_idct2_in_place<2>(c);
  float scratch[64] = {};
  for (size_t y = 0; y < 2; ++y)
    for (size_t x = 0; x < 2; ++x) {
      Grid s{scratch + (y * 2 + x) * 16, 4, 4, 4};
      for (size_t iy = 0; iy < 4; ++iy)
        for (size_t ix = 0; ix < 4; ++ix) s.at(iy, ix) = c.at(x + ix * 2, y + iy * 2);
      dct_2d(s, false);
    }
  for (size_t y = 0; y < 2; ++y)
    for (size_t x = 0; x < 2; ++x) {
      const float* s = scratch + (y * 2 + x) * 16;
      for (size_t iy = 0; iy < 4; ++iy)
        for (size_t ix = 0; ix < 4; ++ix) c.at(x * 4 + ix, y * 4 + iy) = s[iy * 4 + ix];
    }
}

void transform_hornuss(Grid c) {  // transform.rs:84-116
  aux_idct2_in_place<2>(c);
  float scratch[64] = {};
  for (size_t y = 0; y < 2; ++y)
    for (size_t x = 0; x < 2; ++x) {
      float* s = scratch + (y * 2 + x) * 16;
      for (size_t iy = 0; iy < 4; ++iy)
        for (size_t ix = 0; ix < 4; ++ix) s[iy * 4 + ix] = c.at(x + ix * 2, y + iy * 2);
      float residual_sum = 0.0f;
      for (size_t i = 1; i < 16; ++i) residual_sum += s[i];
      float avg = s[0] - residual_sum / 16.0f;
      s[0] = s[5];
      s[5] = 0.0f;
      for (size_t i = 0; i < 16; ++i) s[i] += avg;
    }
  for (size_t y = 0; y < 2; ++y)
    for (size_t x = 0; x < 2; ++x) {
      const float* s = scratch + (y * 2 + x) * 16;
      for (size_t iy = 0; iy < 4; ++iy)
        for (size_t ix = 0; ix < 4; ++ix) c.at(x * 4 + ix, y * 4 + iy) = s[iy * 4 + ix];
    }
}

void transform_dct4x8(Grid c, bool tr) {  // transform.rs:118-146
  float coeff0 = c.at(0, 0), coeff1 = c.at(0, 1);
  c.at(0, 0) = coeff0 + coeff1;
  c.at(0, 1) = coeff0 - coeff1;
  float scratch[64] = {};
  for (size_t idx = 0; idx < 2; ++idx) {
    Grid s{scratch + idx * 32, 8, 8, 4};
    for (size_t iy = 0; iy < 4; ++iy)
      for (size_t ix = 0; ix < 8; ++ix) s.at(ix, iy) = c.at(ix, iy * 2 + idx);
    dct_2d(s, false);
  }
  if (tr) {
    for (size_t y = 0; y < 8; ++y)
      for (size_t x = 0; x < 8; ++x) c.at(y, x) = scratch[y * 8 + x];
  } else {
    for (size_t y = 0; y < 8; ++y)
      for (size_t x = 0; x < 8; ++x) c.at(x, y) = scratch[y * 8 + x];
  }
}

void transform_afv(Grid c, int n) {  // transform.rs:148-222
  const size_t flip_x = n % 2, flip_y = n / 2;
  float coeff_afv[16];
  coeff_afv[0] = (c.at(0, 0) + c.at(1, 0) + c.at(0, 1)) * 4.0f;
  for (size_t idx = 1; idx < 16; ++idx) coeff_afv[idx] = c.at(2 * (idx % 4), 2 * (idx / 4));
  float samples_afv[16] = {};
  for (size_t i = 0; i < 16; ++i)
    for (size_t j = 0; j < 16; ++j) samples_afv[j] = std::fmaf(coeff_afv[i], kAfvBasis[i][j], samples_afv[j]);
  float s4x4[16] = {}, s4x8[32] = {};
  s4x4[0] = c.at(0, 0) - c.at(1, 0) + c.at(0, 1);
  for (size_t iy = 0; iy < 4; ++iy)
    for (size_t ix = 0; ix < 4; ++ix) {
      if ((ix | iy) == 0) continue;
      s4x4[ix * 4 + iy] = c.at(2 * ix + 1, 2 * iy);
    }
  dct_2d(Grid{s4x4, 4, 4, 4}, false);
  s4x8[0] = c.at(0, 0) - c.at(0, 1);
  for (size_t iy = 0; iy < 4; ++iy)
    for (size_t ix = 0; ix < 8; ++ix) {
      if ((ix | iy) == 0) continue;
      s4x8[iy * 8 + ix] = c.at(ix, 2 * iy + 1);
    }
  dct_2d(Grid{s4x8, 8, 8, 4}, false);
  for (size_t iy = 0; iy < 4; ++iy) {
    size_t afv_y = flip_y == 0 ? iy : 3 - iy;
    for (size_t ix = 0; ix < 4; ++ix) {
      size_t afv_x = flip_x == 0 ? ix : 3 - ix;
      c.at(flip_x * 4 + ix, flip_y * 4 + iy) = samples_afv[afv_y * 4 + afv_x];
    }
  }
  for (size_t iy = 0; iy < 4; ++iy)
    for (size_t ix = 0; ix < 4; ++ix) c.at((1 - flip_x) * 4 + ix, flip_y * 4 + iy) = s4x4[iy * 4 + ix];
  for (size_t iy = 0; iy < 4; ++iy)
    for (size_t ix = 0; ix < 8; ++ix) c.at(ix, (1 - flip_y) * 4 + iy) = s4x8[iy * 8 + ix];
}

}  // namespace

// transform_varblocks_inner (transform_common.rs:11-75)
void OracleBackend::hf_transform(VarDctState& st) {
  Plane& type = plane(st.blk_type);
  for (int c = 0; c < 3; ++c) {
    Plane& cp = plane(st.coeff[c]);
    Plane& lf = plane(st.lf[c]);
    parallel_for(st.bh, [&](size_t by) {
      for (uint32_t bx = 0; bx < st.bw; ++bx) {
        int32_t t = type.i32()[by * type.w + bx];
        if (t < 0) continue;
        size_t sbx = bx, sby = by;
        if (!shifted_block(st, type, c, bx, uint32_t(by), &sbx, &sby)) continue;
        const TransformTypeInfo& ti = kTransformInfo[t];
        const size_t bw = ti.w8, bh = ti.h8;
        Grid llf{cp.f32() + (sby * 8) * cp.w + sbx * 8, cp.w, bw, bh};
        if (bw * bh == 1) {
          llf.at(0, 0) = lf.f32()[sby * lf.w + sbx];
        } else {
          for (size_t y = 0; y < bh; ++y)
            for (size_t x = 0; x < bw; ++x) llf.at(x, y) = lf.f32()[(sby + y) * lf.w + sbx + x];
          dct_2d(llf, true);
          size_t logbw = ceil_log2_nonzero(uint32_t(bw)), logbh = ceil_log2_nonzero(uint32_t(bh));
          for (size_t y = 0; y < bh; ++y)
            for (size_t x = 0; x < bw; ++x) llf.at(x, y) /= kScaleF[y << (5 - logbh)] * kScaleF[x << (5 - logbw)];
        }
        Grid block{cp.f32() + (sby * 8) * cp.w + sbx * 8, cp.w, bw * 8, bh * 8};
        switch (t) {
          case kDct2: transform_dct2(block); break;
          case kDct4: transform_dct4(block); break;
          case kHornuss: transform_hornuss(block); break;
          case kDct4x8: transform_dct4x8(block, false); break;
          case kDct8x4: transform_dct4x8(block, true); break;
          case kAfv0: transform_afv(block, 0); break;
          case kAfv1: transform_afv(block, 1); break;
          case kAfv2: transform_afv(block, 2); break;
          case kAfv3: transform_afv(block, 3); break;
          default: dct_2d(block, false); break;
        }
      }
    });
  }
}

}  // namespace jxlo

// Known-answer-test hook (tests/test_oracle_golden.py): the reference's own DCT unit tests
// (crates/jxl-render/src/vardct/generic/dct.rs:295-436) are replayed against this entry point.
extern "C" void jxlo_dct_2d(float* data, uint32_t width, uint32_t height, int forward) {
  jxlo::dct_2d(jxlo::Grid{data, width, width, height}, forward != 0);
}
