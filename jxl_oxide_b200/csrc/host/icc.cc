// Embedded ICC profiles: reconstruction of the profile bytes from the codestream's compressed form and
// recognition of profiles that an enum colour encoding describes exactly (so that an XYB image can be rendered
// into that encoding by the colour kernel, no CMS involved). Restates crates/jxl-color/src/icc/decode.rs and
// icc/parse.rs; profiles the reference cannot map (tabulated curves, LUT-based, CMYK) stay opaque.
#include "icc.h"

#include <cmath>
#include <cstring>

namespace jxlb {

namespace {

// LEB128-style varint of the ICC command stream (decode.rs:108-123)
bool read_varint(const std::vector<uint8_t>& s, size_t* pos, uint64_t* out) {
  uint64_t value = 0;
  for (uint32_t shift = 0; shift < 63; shift += 7) {
    if (*pos >= s.size()) return false;
    const uint8_t b = s[(*pos)++];
    value |= uint64_t(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
  }
  *out = value;
  return true;
}

// The predicted value of header byte idx (decode.rs:125-151)
uint8_t predict_header(size_t idx, uint32_t output_size, const uint8_t* header) {
  static const char kMntr[] = "mntrRGB XYZ ";
  static const char kAcsp[] = "acsp";
  if (idx <= 3) return uint8_t(output_size >> (8 * (3 - idx)));
  if (idx == 8) return 4;
  if (idx >= 12 && idx <= 23) return uint8_t(kMntr[idx - 12]);
  if (idx >= 36 && idx <= 39) return uint8_t(kAcsp[idx - 36]);
  if (idx >= 41 && idx <= 43) {
    const uint8_t h40 = header[40], h41 = header[41];
    if (h40 == 'A') return idx == 43 ? 'L' : 'P';  // APPL
    if (h40 == 'M') return idx == 41 ? 'S' : (idx == 42 ? 'F' : 'T');  // MSFT
    if (idx >= 42 && h40 == 'S' && h41 == 'G') return idx == 42 ? 'I' : ' ';  // SGI_
    if (idx >= 42 && h40 == 'S' && h41 == 'U') return idx == 42 ? 'N' : 'W';  // SUNW
    return 0;
  }
  switch (idx) {
    case 70: return 246;
    case 71: return 214;
    case 73: return 1;
    case 78: return 211;
    case 79: return 45;
    default: break;
  }
  if (idx >= 80 && idx <= 83) return header[4 + idx - 80];
  return 0;
}

// Byte de-interleaving of 16- and 32-bit big-endian arrays (decode.rs:153-190)
std::vector<uint8_t> unshuffle(const uint8_t* bytes, size_t len, size_t width) {
  std::vector<uint8_t> out;
  out.reserve(len);
  if (width == 2) {
    const size_t height = len / 2, odd = len % 2;
    for (size_t i = 0; i < height; ++i) {
      out.push_back(bytes[i]);
      out.push_back(bytes[i + height + odd]);
    }
    if (odd) out.push_back(bytes[height]);
  } else {
    const size_t step = len / 4, wide = len % 4;
    for (size_t i = 0; i < step; ++i) {
      size_t base = i;
      for (size_t k = 0; k < wide; ++k) {
        out.push_back(bytes[base]);
        base += step + 1;
      }
      for (size_t k = wide; k < 4; ++k) {
        out.push_back(bytes[base]);
        base += step;
      }
    }
    for (size_t i = 1; i <= wide; ++i) out.push_back(bytes[(step + 1) * i - 1]);
  }
  return out;
}

void push_be32(std::vector<uint8_t>& out, uint32_t v) {
  for (int s = 24; s >= 0; s -= 8) out.push_back(uint8_t(v >> s));
}
void push_tag(std::vector<uint8_t>& out, const char* tag, uint32_t start, uint32_t size) {
  out.insert(out.end(), tag, tag + 4);
  push_be32(out, start);
  push_be32(out, size);
}

}  // namespace

// decode_icc (decode.rs:192-423)
std::vector<uint8_t> decode_icc_stream(const std::vector<uint8_t>& stream) {
  static const char* kCommonTags[19] = {"rTRC", "rXYZ", "cprt", "wtpt", "bkpt", "rXYZ", "gXYZ", "bXYZ", "kXYZ", "rTRC",
                                        "gTRC", "bTRC", "kTRC", "chad", "desc", "chrm", "dmnd", "dmdd", "lumi"};
  static const char* kCommonData[8] = {"XYZ ", "desc", "text", "mluc", "para", "curv", "sf32", "gbd "};
  size_t pos = 0;
  uint64_t output_size = 0, commands_size = 0;
  JXLB_CHECK(read_varint(stream, &pos, &output_size) && read_varint(stream, &pos, &commands_size), kErrBitstream,
             "ICC stream is too short");
  JXLB_CHECK(commands_size <= stream.size() && pos + commands_size <= stream.size(), kErrBitstream, "invalid ICC commands_size");
  JXLB_CHECK(output_size <= (1u << 28), kErrBitstream, "ICC output_size too large");
  const std::vector<uint8_t> commands(stream.begin() + pos, stream.begin() + pos + commands_size);
  const uint8_t* data = stream.data() + pos + commands_size;
  size_t data_left = stream.size() - pos - size_t(commands_size);
  const size_t header_size = size_t(std::min<uint64_t>(output_size, 128));
  JXLB_CHECK(data_left >= header_size, kErrBitstream, "invalid ICC output_size");
  std::vector<uint8_t> out;
  out.reserve(size_t(output_size));
  {
    uint8_t header[128] = {};
    std::memcpy(header, data, header_size);
    for (size_t i = 0; i < header_size; ++i) out.push_back(uint8_t(predict_header(i, uint32_t(output_size), header) + header[i]));
    data += header_size;
    data_left -= header_size;
  }
  if (output_size <= 128) return out;
  auto take = [&](size_t n) {
    JXLB_CHECK(n <= data_left, kErrBitstream, "ICC data stream is too short");
    const uint8_t* p = data;
    data += n;
    data_left -= n;
    return p;
  };
  size_t cpos = 0;
  auto cvarint = [&]() {
    uint64_t v = 0;
    JXLB_CHECK(read_varint(commands, &cpos, &v), kErrBitstream, "ICC command stream is too short");
    return v;
  };

  // tag list
  const uint64_t v = cvarint();
  if (v >= 1) {
    const uint64_t num_tags64 = v - 1;
    JXLB_CHECK((output_size - 128) / 12 >= num_tags64, kErrBitstream, "ICC num_tags too large");
    const uint32_t num_tags = uint32_t(num_tags64);
    push_be32(out, num_tags);
    uint32_t prev_tagstart = num_tags * 12 + 128, prev_tagsize = 0;
    for (;;) {
      if (cpos >= commands.size()) return out;
      const uint8_t command = commands[cpos++];
      const uint32_t tagcode = command & 63;
      char tag[4];
      if (tagcode == 0) break;
      if (tagcode == 1) {
        std::memcpy(tag, take(4), 4);
      } else {
        JXLB_CHECK(tagcode <= 20, kErrBitstream, "invalid ICC tagcode");
        std::memcpy(tag, kCommonTags[tagcode - 2], 4);
      }
      const uint32_t tagstart = (command & 64) ? uint32_t(cvarint()) : prev_tagstart + prev_tagsize;
      uint32_t tagsize = prev_tagsize;
      if (command & 128) {
        tagsize = uint32_t(cvarint());
      } else {
        for (const char* fixed : {"rXYZ", "gXYZ", "bXYZ", "kXYZ", "wtpt", "bkpt", "lumi"})
          if (!std::memcmp(tag, fixed, 4)) tagsize = 20;
      }
      JXLB_CHECK(uint64_t(tagstart) + tagsize <= output_size, kErrBitstream, "ICC profile size mismatch");
      prev_tagstart = tagstart;
      prev_tagsize = tagsize;
      push_tag(out, tag, tagstart, tagsize);
      if (tagcode == 2) {
        push_tag(out, "gTRC", tagstart, tagsize);
        push_tag(out, "bTRC", tagstart, tagsize);
      } else if (tagcode == 3) {
        push_tag(out, "gXYZ", tagstart + tagsize, tagsize);
        push_tag(out, "bXYZ", tagstart + tagsize * 2, tagsize);
      }
    }
  }

  // main content
  while (cpos < commands.size()) {
    const uint8_t command = commands[cpos++];
    if (command == 1) {
      const size_t num = size_t(cvarint());
      const uint8_t* b = take(num);
      out.insert(out.end(), b, b + num);
    } else if (command == 2 || command == 3) {
      const size_t num = size_t(cvarint());
      const uint8_t* b = take(num);
      const std::vector<uint8_t> s = unshuffle(b, num, command == 2 ? 2 : 4);
      out.insert(out.end(), s.begin(), s.end());
    } else if (command == 4) {  // Nth-order prediction of big-endian integers
      JXLB_CHECK(cpos < commands.size(), kErrBitstream, "ICC command stream is too short");
      const uint8_t flags = commands[cpos++];
      const size_t width = (flags & 3) + 1;
      const uint32_t order = (flags >> 2) & 3;
      JXLB_CHECK(width != 3 && order != 3, kErrBitstream, "invalid ICC predictor");
      size_t stride = width;
      if (flags & 16) {
        const uint64_t s = cvarint();
        JXLB_CHECK(s >= width && s < (uint64_t(1) << 40), kErrBitstream, "invalid ICC predictor stride");
        stride = size_t(s);
      }
      JXLB_CHECK(stride * 4 < out.size(), kErrBitstream, "ICC predictor stride beyond the decoded part");
      const size_t num = size_t(cvarint());
      const uint8_t* raw = take(num);
      std::vector<uint8_t> shuffled;
      const uint8_t* bytes = raw;
      if (width != 1) {
        shuffled = unshuffle(raw, num, width);
        bytes = shuffled.data();
      }
      for (size_t i = 0; i < num; i += width) {
        uint32_t prev[3] = {0, 0, 0};
        for (uint32_t j = 0; j <= order; ++j) {
          const size_t offset = out.size() - stride * (j + 1);
          uint32_t val = 0;
          for (size_t k = 0; k < width; ++k) val = (val << 8) | out[offset + k];
          prev[j] = val;
        }
        uint32_t p;  // wrapping arithmetic
        if (order == 0) p = prev[0];
        else if (order == 1) p = 2u * prev[0] - prev[1];
        else p = 3u * (prev[0] - prev[1]) + prev[2];
        for (size_t j = 0; j < std::min(width, num - i); ++j) out.push_back(uint8_t(uint32_t(bytes[i + j]) + (p >> (8 * (width - 1 - j)))));
      }
    } else if (command == 10) {
      static const uint8_t kXyz[8] = {'X', 'Y', 'Z', ' ', 0, 0, 0, 0};
      out.insert(out.end(), kXyz, kXyz + 8);
      const uint8_t* b = take(12);
      out.insert(out.end(), b, b + 12);
    } else if (command >= 16 && command <= 23) {
      const char* d = kCommonData[command - 16];
      out.insert(out.end(), d, d + 4);
      out.insert(out.end(), 4, uint8_t(0));
    } else {
      fail(kErrBitstream, "invalid ICC command");
    }
    JXLB_CHECK(out.size() <= output_size, kErrBitstream, "decoded ICC profile size mismatch");
  }
  JXLB_CHECK(out.size() == output_size, kErrBitstream, "decoded ICC profile size mismatch");
  return out;
}

namespace {

// transfer curves the reference recognises (parse.rs:170-210, 300-397)
struct KnownTrc {
  enum Kind { kNone, kGamma, kLinear, kSrgb, kBt709 } kind = kNone;
  uint32_t gamma = 0;  // s15Fixed16
  bool operator==(const KnownTrc& o) const { return kind == o.kind && gamma == o.gamma; }
};
KnownTrc trc_from_gamma(int32_t g) {
  KnownTrc t;
  if (g <= 65535) return t;
  if (g == 65536) t.kind = KnownTrc::kLinear;
  else t.kind = KnownTrc::kGamma, t.gamma = uint32_t(g);
  return t;
}
int32_t be_i32(const uint8_t* p) { return int32_t((uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]); }
uint32_t be_u32(const uint8_t* p) { return uint32_t(be_i32(p)); }

typedef float Mat3[9];
void matinv3(const Mat3 m, Mat3 out) {  // ciexyz.rs:90-105
  const float det = m[0] * (m[4] * m[8] - m[5] * m[7]) + m[1] * (m[5] * m[6] - m[3] * m[8]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
  out[0] = (m[4] * m[8] - m[5] * m[7]) / det, out[1] = (m[7] * m[2] - m[8] * m[1]) / det, out[2] = (m[1] * m[5] - m[2] * m[4]) / det;
  out[3] = (m[5] * m[6] - m[3] * m[8]) / det, out[4] = (m[8] * m[0] - m[6] * m[2]) / det, out[5] = (m[2] * m[3] - m[0] * m[5]) / det;
  out[6] = (m[3] * m[7] - m[4] * m[6]) / det, out[7] = (m[6] * m[1] - m[7] * m[0]) / det, out[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}
bool xyz_valid(const int32_t xyz[3]) {  // validate_xyz
  const float f[3] = {float(xyz[0]) / 65536.0f, float(xyz[1]) / 65536.0f, float(xyz[2]) / 65536.0f};
  const float sum = f[0] + f[1] + f[2];
  for (float v : f)
    if (!std::isfinite(v / sum)) return false;
  return true;
}

}  // namespace

// detect_profile_info + parse_icc (parse.rs:229-560)
IccStatus icc_to_enum(const std::vector<uint8_t>& profile, IccInfo* info) {
  *info = IccInfo();
  if (profile.size() < 128) return IccStatus::kMalformed;
  const uint32_t size = be_u32(&profile[0]);
  if (profile.size() != size) return IccStatus::kMalformed;
  const uint8_t* cs = &profile[0x10];
  info->is_gray = !std::memcmp(cs, "GRAY", 4);
  info->is_cmyk = !std::memcmp(cs, "CMYK", 4);
  const bool is_rgb = !std::memcmp(cs, "RGB ", 4);
  const uint8_t intent = profile[0x43];
  if (intent > 3) return IccStatus::kMalformed;
  struct Tag {
    const uint8_t* name;
    const uint8_t* data;
    size_t len;
  };
  std::vector<Tag> tags;
  if (size >= 0x84) {
    const uint32_t tag_count = be_u32(&profile[0x80]);
    if (uint64_t(size) < 0x84 + 12 * uint64_t(tag_count)) return IccStatus::kMalformed;
    for (uint32_t i = 0; i < tag_count; ++i) {
      const uint8_t* raw = &profile[0x84 + 12 * size_t(i)];
      const uint64_t offset = be_u32(raw + 4), tag_size = be_u32(raw + 8);
      if (offset + tag_size > size) return IccStatus::kMalformed;
      tags.push_back({raw, profile.data() + offset, size_t(tag_size)});
    }
  }
  int32_t wtpt[3] = {0xf6d6, 0x10000, 0xd32d};  // D50
  int32_t chad[9] = {65536, 0, 0, 0, 65536, 0, 0, 0, 65536};
  KnownTrc trcs[4];
  bool have_trc[4] = {false, false, false, false};
  int32_t xyzs[3][3] = {};
  bool have_xyz[3] = {false, false, false};
  bool have_cicp = false;
  uint8_t cicp[4] = {};
  for (const Tag& t : tags) {
    const uint8_t* d = t.data;
    const size_t n = t.len;
    if (n < 4) continue;
    const uint8_t* tag = t.name;
    if (!std::memcmp(tag + 1, "TRC", 3)) {
      int index;
      switch (tag[0]) {
        case 'r': index = 0; break;
        case 'g': index = 1; break;
        case 'b': index = 2; break;
        case 'k': index = 3; break;
        default: continue;
      }
      KnownTrc tf;
      if (!std::memcmp(d, "para", 4)) {
        if (n < 12) continue;
        const uint32_t curve_type = (uint32_t(d[8]) << 8) | d[9];
        const size_t parameters = (n - 12) / 4;
        if (curve_type == 0) {
          if (parameters != 1) return IccStatus::kMalformed;
          tf = trc_from_gamma(be_i32(d + 12));
          if (tf.kind == KnownTrc::kNone) continue;
        } else if (curve_type == 3) {
          if (parameters != 5) return IccStatus::kMalformed;
          int32_t p[5];
          for (int k = 0; k < 5; ++k) p[k] = be_i32(d + 12 + 4 * k);
          const int32_t k709[5] = {(65536 * 20 + 4) / 9, (65536 * 1000 + 549) / 1099, (65536 * 99 + 549) / 1099, (65536 * 10 + 22) / 45,
                                   (65536 * 81 + 500) / 1000};
          const int32_t ksrgb[5] = {(65536 * 24 + 5) / 10, (65536 * 1000 + 527) / 1055, (65536 * 55 + 527) / 1055,
                                    (65536 * 100 + 646) / 1292, int32_t((int64_t(65536) * 4045 + 50000) / 100000)};
          if (!std::memcmp(p, k709, sizeof(p))) {
            tf.kind = KnownTrc::kBt709;
          } else if (!std::memcmp(p, ksrgb, sizeof(p))) {
            tf.kind = KnownTrc::kSrgb;
          } else if (p[1] == 65536 && p[2] == 0 && p[3] == 65536 && p[4] == 0) {
            tf = trc_from_gamma(p[0]);
            if (tf.kind == KnownTrc::kNone) continue;
          } else {
            continue;
          }
        } else {
          continue;
        }
      } else if (n == 12 && !std::memcmp(d, "curv\0\0\0\0\0\0\0\0", 12)) {
        tf.kind = KnownTrc::kLinear;
      } else if (n == 14 && !std::memcmp(d, "curv\0\0\0\0\0\0\0\1", 12)) {
        tf.kind = KnownTrc::kGamma;
        tf.gamma = (uint32_t(d[12]) << 16) | (uint32_t(d[13]) << 8);
      } else {
        continue;
      }
      trcs[index] = tf;
      have_trc[index] = true;
    } else if (!std::memcmp(tag + 1, "XYZ", 3)) {
      int index;
      switch (tag[0]) {
        case 'r': index = 0; break;
        case 'g': index = 1; break;
        case 'b': index = 2; break;
        default: continue;
      }
      if (std::memcmp(d, "XYZ ", 4) || n < 20) return IccStatus::kMalformed;
      for (int k = 0; k < 3; ++k) xyzs[index][k] = be_i32(d + 8 + 4 * k);
      if (!xyz_valid(xyzs[index])) return IccStatus::kMalformed;
      have_xyz[index] = true;
    } else if (!std::memcmp(tag, "chad", 4)) {
      if (std::memcmp(d, "sf32", 4) || n < 44) return IccStatus::kMalformed;
      for (int k = 0; k < 9; ++k) chad[k] = be_i32(d + 8 + 4 * k);
      Mat3 m, inv;
      for (int k = 0; k < 9; ++k) m[k] = float(chad[k]) / 65536.0f;
      matinv3(m, inv);
      for (float x : inv)
        if (!std::isfinite(x)) return IccStatus::kMalformed;
    } else if (!std::memcmp(tag, "wtpt", 4)) {
      if (std::memcmp(d, "XYZ ", 4) || n < 20) return IccStatus::kMalformed;
      for (int k = 0; k < 3; ++k) wtpt[k] = be_i32(d + 8 + 4 * k);
      if (!xyz_valid(wtpt)) return IccStatus::kMalformed;
    } else if (((tag[0] == 'A' || tag[0] == 'D') && tag[1] == '2' && tag[2] == 'B' && tag[3] >= '0' && tag[3] <= '3') ||
               (tag[0] == 'B' && tag[1] == '2' && (tag[2] == 'A' || tag[2] == 'D') && tag[3] >= '0' && tag[3] <= '3') ||
               (!std::memcmp(tag, "pre", 3) && tag[3] >= '0' && tag[3] <= '2')) {
      return IccStatus::kUnsupported;
    } else if (!std::memcmp(tag, "chrm", 4) || !std::memcmp(tag, "clro", 4) || !std::memcmp(tag, "clrt", 4) ||
               !std::memcmp(tag, "clot", 4) || !std::memcmp(tag, "ciis", 4) || !std::memcmp(tag, "lumi", 4) ||
               !std::memcmp(tag, "meas", 4) || !std::memcmp(tag, "ncl2", 4) || !std::memcmp(tag, "resp", 4) ||
               !std::memcmp(tag, "view", 4)) {
      return IccStatus::kUnsupported;
    } else if (!std::memcmp(tag, "cicp", 4)) {
      std::memcpy(cicp, d, 4);
      have_cicp = true;
    }
  }
  // cicp overrides the curves for PQ / HLG (parse.rs:463-470); note that the four bytes are the tag data's first four
  TransferFunctionKind override_tf = TransferFunctionKind::kUnknown;
  if (have_cicp && cicp[1] == 16) override_tf = TransferFunctionKind::kPq;
  if (have_cicp && cicp[1] == 18) override_tf = TransferFunctionKind::kHlg;

  auto to_tf = [&](const KnownTrc& t, ColourEncoding* ce) {
    ce->gamma_inverted = true;
    if (override_tf != TransferFunctionKind::kUnknown) {
      ce->tf = override_tf;
      return;
    }
    switch (t.kind) {
      case KnownTrc::kLinear: ce->tf = TransferFunctionKind::kLinear; break;
      case KnownTrc::kSrgb: ce->tf = TransferFunctionKind::kSrgb; break;
      case KnownTrc::kBt709: ce->tf = TransferFunctionKind::kBt709; break;
      default:  // ParametricGamma -> Gamma { g * 1e7 / 65536 rounded, inverted: false }
        ce->tf = TransferFunctionKind::kGamma;
        ce->gamma = uint32_t((uint64_t(t.gamma) * 10000000 + 32768) / 65536);
        ce->gamma_inverted = false;
        break;
    }
  };

  // chad^-1, used to undo the D50 adaptation of the colorant and white point tags (parse.rs:52-165)
  Mat3 chad_f, chad_inv;
  for (int k = 0; k < 9; ++k) chad_f[k] = float(chad[k]) / 65536.0f;
  matinv3(chad_f, chad_inv);
  auto white_point = [&](ColourEncoding* ce) {
    const float w[3] = {float(wtpt[0]) / 65536.0f, float(wtpt[1]) / 65536.0f, float(wtpt[2]) / 65536.0f};
    const float ill[3] = {chad_inv[0] * w[0] + chad_inv[1] * w[1] + chad_inv[2] * w[2], chad_inv[3] * w[0] + chad_inv[4] * w[1] + chad_inv[5] * w[2],
                          chad_inv[6] * w[0] + chad_inv[7] * w[1] + chad_inv[8] * w[2]};
    const float sum = ill[0] + ill[1] + ill[2];
    const float xy[2] = {ill[0] / sum, ill[1] / sum};
    struct Known {
      float xy[2];
      WhitePointKind kind;
    };
    static const Known kKnown[3] = {{{0.3127f, 0.329f}, WhitePointKind::kD65}, {{0.314f, 0.351f}, WhitePointKind::kDci},
                                    {{1.0f / 3.0f, 1.0f / 3.0f}, WhitePointKind::kE}};
    for (const Known& k : kKnown)
      if (std::fabs(xy[0] - k.xy[0]) < 1e-4f && std::fabs(xy[1] - k.xy[1]) < 1e-4f) {
        ce->white_point = k.kind;
        return;
      }
    ce->white_point = WhitePointKind::kCustom;
    ce->white_xy[0] = int32_t(xy[0] * 1e6f + 0.5f);
    ce->white_xy[1] = int32_t(xy[1] * 1e6f + 0.5f);
  };

  ColourEncoding ce;
  ce.want_icc = false;
  ce.rendering_intent = intent;
  if (info->is_cmyk) return IccStatus::kUnsupported;
  if (info->is_gray) {
    if (!have_trc[3] && override_tf == TransferFunctionKind::kUnknown) return IccStatus::kUnsupported;
    if (!have_trc[3]) return IccStatus::kUnsupported;  // trc_k.map(...): the override applies only to an existing curve
    ce.colour_space = ColourSpace::kGrey;
    ce.primaries = PrimariesKind::kSrgb;
    to_tf(trcs[3], &ce);
    white_point(&ce);
  } else if (is_rgb) {
    if (!(have_trc[0] && have_trc[1] && have_trc[2])) return IccStatus::kUnsupported;
    if (override_tf == TransferFunctionKind::kUnknown && !(trcs[0] == trcs[1] && trcs[1] == trcs[2])) return IccStatus::kUnsupported;
    if (!(have_xyz[0] && have_xyz[1] && have_xyz[2])) return IccStatus::kUnsupported;
    ce.colour_space = ColourSpace::kRgb;
    to_tf(trcs[0], &ce);
    // primaries(): colorants as columns, un-adapted, normalised to chromaticities
    Mat3 m, a;
    for (int idx = 0; idx < 9; ++idx) m[idx] = float(xyzs[idx % 3][idx / 3]) / 65536.0f;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) a[i * 3 + j] = chad_inv[i * 3] * m[j] + chad_inv[i * 3 + 1] * m[3 + j] + chad_inv[i * 3 + 2] * m[6 + j];
    const float sum[3] = {a[0] + a[3] + a[6], a[1] + a[4] + a[7], a[2] + a[5] + a[8]};
    const float prim[3][2] = {{a[0] / sum[0], a[3] / sum[0]}, {a[1] / sum[1], a[4] / sum[1]}, {a[2] / sum[2], a[5] / sum[2]}};
    struct Known {
      float p[3][2];
      PrimariesKind kind;
    };
    static const Known kKnown[3] = {
        {{{0.639998686f, 0.330010138f}, {0.300003784f, 0.600003357f}, {0.150002046f, 0.059997204f}}, PrimariesKind::kSrgb},
        {{{0.680f, 0.320f}, {0.265f, 0.690f}, {0.150f, 0.060f}}, PrimariesKind::kP3},
        {{{0.708f, 0.292f}, {0.170f, 0.797f}, {0.131f, 0.046f}}, PrimariesKind::kBt2100}};
    ce.primaries = PrimariesKind::kCustom;
    for (const Known& k : kKnown) {
      bool match = true;
      for (int y = 0; y < 3; ++y)
        for (int x = 0; x < 2; ++x) match &= std::fabs(prim[y][x] - k.p[y][x]) < 1e-4f;
      if (match) {
        ce.primaries = k.kind;
        break;
      }
    }
    if (ce.primaries == PrimariesKind::kCustom)
      for (int y = 0; y < 3; ++y)
        for (int x = 0; x < 2; ++x) ce.primaries_xy[y][x] = int32_t(prim[y][x] * 1e6f + 0.5f);
    white_point(&ce);
  } else {
    return IccStatus::kUnsupported;
  }
  info->encoding = ce;
  return IccStatus::kEnum;
}

}  // namespace jxlb
