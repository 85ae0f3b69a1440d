// Modular sub-bitstream *syntax*: MA tree, per-image header (weighted-predictor params,
// transform list) and the channel bookkeeping that transforms imply. No sample is decoded here.
//
// Reference: crates/jxl-modular/src/{lib.rs,ma.rs,param.rs,predictor.rs:8-25,transform.rs}.
#pragma once
#include <cstdint>
#include <vector>

#include "bitreader.h"
#include "entropy.h"

namespace jxlb {

struct WpHeader {  // predictor.rs:8-25
  uint32_t p1 = 16, p2 = 10, p3a = 7, p3b = 7, p3c = 7, p3d = 0, p3e = 0;
  uint32_t w[4] = {13, 12, 12, 12};
};

// MA tree node, 16 bytes, uploaded verbatim to the device.
//   decision: property >= 0, value, a = index of the `property > value` child, b = other child
//   leaf:     property = -1, value = offset, a = predictor | cluster << 8, b = multiplier
struct MaNode {
  int32_t property;
  int32_t value;
  uint32_t a;
  uint32_t b;
};

struct MaTree {  // ma.rs:16-226
  std::vector<MaNode> nodes;
  EntropyCode code;
  uint32_t num_leaves = 0;
};

MaTree parse_ma_tree(BitReader& br, size_t node_limit, size_t depth_limit = 2048);

struct ChannelInfo {  // lib.rs:143-190 (ModularChannelInfo)
  uint32_t width = 0, height = 0;
  int32_t hshift = 0, vshift = 0;
};

struct SqueezeStep {  // transform.rs:130-136
  bool horizontal = false, in_place = false;
  uint32_t begin_c = 0, num_c = 0;
};

struct Transform {
  enum Kind : uint32_t { kRct = 0, kPalette = 1, kSqueeze = 2 } kind = kRct;
  uint32_t begin_c = 0;
  uint32_t rct_type = 0;                                   // Rct
  uint32_t num_c = 0, nb_colours = 0, nb_deltas = 0, d_pred = 0;  // Palette
  std::vector<SqueezeStep> squeeze;                        // Squeeze (defaults resolved)
};

struct ModularHeader {  // lib.rs:117-125
  bool use_global_tree = false;
  WpHeader wp;
  std::vector<Transform> transforms;
};

// Parses the header, applies the transforms' channel bookkeeping to `channels`
// (`prepare_transform_info`, transform.rs:25-40) and, when the stream carries its own tree,
// parses it into `local_tree`. lib.rs:192-245.
struct ModularStreamSyntax {
  ModularHeader header;
  std::vector<ChannelInfo> channels;  // coded channels, after all forward transforms
  uint32_t nb_meta_channels = 0;
  bool has_local_tree = false;
  MaTree local_tree;
};
ModularStreamSyntax parse_modular_stream_header(BitReader& br, const std::vector<ChannelInfo>& image_channels,
                                                bool global_tree_available);

// ChannelShift helpers (param.rs:108-175)
struct JpegShift {
  int32_t hshift, vshift;
  uint32_t width, height;
};

}  // namespace jxlb
