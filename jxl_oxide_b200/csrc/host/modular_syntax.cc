// See modular_syntax.h.
#include "modular_syntax.h"

#include <algorithm>

namespace jxlb {

MaTree parse_ma_tree(BitReader& br, size_t node_limit, size_t depth_limit) {  // ma.rs:68-226
  MaTree tree;
  EntropyCode tree_code = parse_entropy_code(br, 6);
  {  // is_infinite_tree_dist (ma.rs:228-239)
    int32_t tok = tree_code.single_token(tree_code.cluster_map[1]);
    JXLB_CHECK(!(tok > 0), kErrBitstream, "infinite MA tree");
  }
  EntropyReader dec(&tree_code);
  dec.begin(br);
  size_t nodes_left = 1;
  uint32_t ctx = 0;
  uint32_t next_child = 1;
  std::vector<uint32_t> leaf_ctx;  // per node: ctx for leaves
  while (nodes_left > 0) {
    JXLB_CHECK(tree.nodes.size() < (1u << 26), kErrBitstream, "invalid MA tree");
    JXLB_CHECK(tree.nodes.size() <= node_limit, kErrBitstream, "MA tree node limit exceeded");
    --nodes_left;
    uint32_t property = dec.read_varint(br, 1);
    MaNode node;
    if (property > 0) {
      // Properties beyond the last previous channel a stream can have (16 + 4 * 16) evaluate to 0 whatever their
      // number (predictor.rs:495-528), so they share one value: a huge property must never look like a leaf (< 0).
      node.property = int32_t(std::min<uint32_t>(property - 1, 16 + 4 * 16));
      node.value = unpack_signed(dec.read_varint(br, 0));
      node.a = next_child;
      node.b = next_child + 1;
      next_child += 2;
      nodes_left += 2;
      leaf_ctx.push_back(0);
    } else {
      uint32_t predictor = dec.read_varint(br, 2);
      JXLB_CHECK(predictor <= 13, kErrBitstream, "invalid MA predictor");
      int32_t offset = unpack_signed(dec.read_varint(br, 3));
      uint32_t mul_log = dec.read_varint(br, 4);
      JXLB_CHECK(mul_log <= 30, kErrBitstream, "invalid MA tree");
      uint32_t mul_bits = dec.read_varint(br, 5);
      JXLB_CHECK(mul_bits <= (1u << (31 - mul_log)) - 2, kErrBitstream, "invalid MA tree");
      node.property = -1;
      node.value = offset;
      node.a = predictor;
      node.b = (mul_bits + 1) << mul_log;
      leaf_ctx.push_back(ctx++);
    }
    tree.nodes.push_back(node);
    br.check();
  }
  JXLB_CHECK(dec.finalize_ok(), kErrBitstream, "invalid ANS stream (MA tree)");
  tree.num_leaves = ctx;
  tree.code = parse_entropy_code(br, ctx);
  // depth check and leaf clustering
  std::vector<uint32_t> depth(tree.nodes.size(), 0);
  for (size_t i = tree.nodes.size(); i-- > 0;) {
    MaNode& n = tree.nodes[i];
    if (n.property >= 0) {
      depth[i] = std::max(depth[n.a], depth[n.b]) + 1;
      JXLB_CHECK(depth[i] <= depth_limit, kErrBitstream, "MA tree too deep");
    } else {
      n.a |= uint32_t(tree.code.cluster_map[leaf_ctx[i]]) << 8;
    }
  }
  return tree;
}

namespace {

using D = BitReader::U32Dist;

uint32_t read_begin_c(BitReader& br) { return br.read_u32({0, 3}, {8, 6}, {72, 10}, {1096, 13}); }

Transform parse_transform(BitReader& br) {  // transform.rs:103-147
  Transform t;
  uint32_t tr = br.read(2);
  if (tr == 0) {
    t.kind = Transform::kRct;
    t.begin_c = read_begin_c(br);
    t.rct_type = br.read_u32({6, 0}, {0, 2}, {2, 4}, {10, 6});
  } else if (tr == 1) {
    t.kind = Transform::kPalette;
    t.begin_c = read_begin_c(br);
    t.num_c = br.read_u32({1, 0}, {3, 0}, {4, 0}, {1, 13});
    t.nb_colours = br.read_u32({0, 8}, {256, 10}, {1280, 12}, {5376, 16});
    t.nb_deltas = br.read_u32({0, 0}, {1, 8}, {257, 10}, {1281, 16});
    t.d_pred = br.read(4);
    JXLB_CHECK(t.d_pred <= 13, kErrBitstream, "invalid palette predictor");
  } else if (tr == 2) {
    t.kind = Transform::kSqueeze;
    uint32_t num_sq = br.read_u32({0, 0}, {1, 4}, {9, 6}, {41, 8});
    for (uint32_t i = 0; i < num_sq; ++i) {
      SqueezeStep s;
      s.horizontal = br.read_bool();
      s.in_place = br.read_bool();
      s.begin_c = read_begin_c(br);
      s.num_c = br.read_u32({1, 0}, {2, 0}, {3, 0}, {4, 4});
      t.squeeze.push_back(s);
    }
  } else {
    fail(kErrBitstream, "invalid transform id");
  }
  br.check();
  return t;
}

void default_squeeze(Transform& t, const std::vector<ChannelInfo>& ch, uint32_t nb_meta) {  // transform.rs:285-341
  if (!t.squeeze.empty()) return;
  uint32_t first = nb_meta;
  JXLB_CHECK(first < ch.size(), kErrBitstream, "invalid squeeze params");
  uint32_t w = ch[first].width, h = ch[first].height;
  if (ch.size() - first >= 3) {
    const ChannelInfo& next = ch[first + 1];
    if (next.width == w && next.height == h) {
      t.squeeze.push_back({true, false, first + 1, 2});
      t.squeeze.push_back({false, false, first + 1, 2});
    }
  }
  uint32_t num_c = uint32_t(ch.size()) - first;
  if (h >= w && h > 8) {
    t.squeeze.push_back({false, true, first, num_c});
    h = (h + 1) / 2;
  }
  while (w > 8 || h > 8) {
    if (w > 8) {
      t.squeeze.push_back({true, true, first, num_c});
      w = (w + 1) / 2;
    }
    if (h > 8) {
      t.squeeze.push_back({false, true, first, num_c});
      h = (h + 1) / 2;
    }
  }
}

void apply_transform_info(Transform& t, std::vector<ChannelInfo>& ch, uint32_t& nb_meta) {
  if (t.kind == Transform::kRct) {  // transform.rs:175-191
    uint64_t end_c = uint64_t(t.begin_c) + 3;
    JXLB_CHECK(end_c <= ch.size(), kErrBitstream, "invalid RCT params");
    for (uint32_t i = t.begin_c + 1; i < end_c; ++i)
      JXLB_CHECK(ch[i].width == ch[t.begin_c].width && ch[i].height == ch[t.begin_c].height, kErrBitstream,
                 "invalid RCT params");
  } else if (t.kind == Transform::kPalette) {  // transform.rs:218-263
    uint64_t begin_c = t.begin_c, end_c = begin_c + t.num_c;
    JXLB_CHECK(end_c <= ch.size(), kErrBitstream, "invalid palette params");
    if (begin_c < nb_meta) {
      JXLB_CHECK(end_c <= nb_meta, kErrBitstream, "invalid palette params");
      nb_meta = nb_meta + 2 - t.num_c;
    } else {
      nb_meta += 1;
    }
    for (uint64_t i = begin_c + 1; i < end_c; ++i)
      JXLB_CHECK(ch[i].width == ch[begin_c].width && ch[i].height == ch[begin_c].height, kErrBitstream,
                 "invalid palette params");
    ch.erase(ch.begin() + begin_c + 1, ch.begin() + end_c);
    ChannelInfo pal;
    pal.width = t.nb_colours;
    pal.height = t.num_c;
    pal.hshift = pal.vshift = -1;
    ch.insert(ch.begin(), pal);
  } else {  // transform.rs:343-437
    default_squeeze(t, ch, nb_meta);
    for (const SqueezeStep& sp : t.squeeze) {
      uint64_t begin = sp.begin_c, end = begin + sp.num_c;
      JXLB_CHECK(end <= ch.size(), kErrBitstream, "invalid squeeze params");
      if (begin < nb_meta) {
        JXLB_CHECK(sp.in_place && end <= nb_meta, kErrBitstream, "invalid squeeze params");
        nb_meta += sp.num_c;
      }
      std::vector<ChannelInfo> residu;
      for (uint64_t i = begin; i < end; ++i) {
        ChannelInfo& c = ch[i];
        ChannelInfo r = c;
        JXLB_CHECK(c.width != 0 && c.height != 0, kErrBitstream, "cannot squeeze zero-sized channel");
        JXLB_CHECK(c.hshift <= 30 && c.vshift <= 30, kErrBitstream, "channel squeezed too much");
        if (sp.horizontal) {
          uint32_t len = c.width;
          c.width = (len + 1) / 2;
          r.width = len / 2;
          if (c.hshift >= 0) {
            c.hshift += 1;
            r.hshift += 1;
          }
        } else {
          uint32_t len = c.height;
          c.height = (len + 1) / 2;
          r.height = len / 2;
          if (c.vshift >= 0) {
            c.vshift += 1;
            r.vshift += 1;
          }
        }
        residu.push_back(r);
      }
      if (sp.in_place) ch.insert(ch.begin() + end, residu.begin(), residu.end());
      else ch.insert(ch.end(), residu.begin(), residu.end());
    }
  }
}

}  // namespace

ModularStreamSyntax parse_modular_stream_header(BitReader& br, const std::vector<ChannelInfo>& image_channels,
                                                bool global_tree_available) {
  ModularStreamSyntax s;
  ModularHeader& h = s.header;
  h.use_global_tree = br.read_bool();
  if (!br.read_bool()) {  // !default_wp
    WpHeader& w = h.wp;
    w.p1 = br.read(5), w.p2 = br.read(5);
    w.p3a = br.read(5), w.p3b = br.read(5), w.p3c = br.read(5), w.p3d = br.read(5), w.p3e = br.read(5);
    for (uint32_t& v : w.w) v = br.read(4);
  }
  uint32_t nb_transforms = br.read_u32({0, 0}, {1, 0}, {2, 4}, {18, 8});
  for (uint32_t i = 0; i < nb_transforms; ++i) h.transforms.push_back(parse_transform(br));
  br.check();
  JXLB_CHECK(nb_transforms <= 512, kErrBitstream, "nb_transforms too large");
  s.channels = image_channels;
  s.nb_meta_channels = 0;
  for (Transform& t : h.transforms) apply_transform_info(t, s.channels, s.nb_meta_channels);
  JXLB_CHECK(s.channels.size() <= (1u << 16), kErrBitstream, "too many channels after transforms");
  if (h.use_global_tree) {
    JXLB_CHECK(global_tree_available, kErrBitstream, "global MA tree requested but not available");
  } else {
    uint64_t samples = 0;
    for (const ChannelInfo& c : s.channels) samples += uint64_t(c.width) * c.height;
    size_t limit = size_t(std::min<uint64_t>(1024 + samples, 1u << 20));
    s.local_tree = parse_ma_tree(br, limit);
    s.has_local_tree = true;
  }
  return s;
}

}  // namespace jxlb
