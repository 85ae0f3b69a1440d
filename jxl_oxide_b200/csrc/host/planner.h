// Host-side frame planner: walks the codestream syntax (image header, frame header, TOC,
// LfGlobal, LfGroups, HfGlobal, PassGroups), builds the tables the sample-level stages need and
// drives a `Backend` through the stages in the order of the reference's render_frame
// (crates/jxl-render/src/render.rs:14-156, vardct/mod.rs:48-385, modular.rs:6-147,
// lib.rs:925-998). It never touches a sample itself.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "backend.h"

namespace jxlb {

struct DecodeOptions {
  // Output colour: 0 = image's signalled encoding (sRGB transfer for the supported set),
  // 1 = linear sRGB, 2 = leave XYB.
  int output_colour = 0;
  uint32_t max_frames = 0xffffffffu;
};

struct DecodedFrame {
  bool internal = false;  // an LF frame: consumed by later frames, never shown (jxl-render/src/lib.rs:294-318)
  uint32_t width = 0, height = 0;
  uint32_t num_color = 0;
  std::vector<View> channels;  // f32 planes: colour channels then extra channels
  FrameHeader header;
};

// What ImageStream::from_render puts into an interleaved buffer (jxl-oxide/src/fb.rs:184-283): the colour channels,
// then the first alpha channel; spot-colour channels are mixed into a three-channel colour image while it is written
// (fb.rs:335-362) unless the image is grayscale (lib.rs:416). The black channel of a CMYK image (recognised by its
// ICC profile's data colour space) follows the colour channels; the samples stay CMYK - no CMS here.
struct StreamSpot {
  size_t channel;  // index into DecodedFrame::channels
  float rgb[3];
  float solidity;
};
struct StreamLayout {
  std::vector<size_t> channels;  // indices into DecodedFrame::channels, in output order
  std::vector<StreamSpot> spots;
};
StreamLayout stream_layout(const ImageHeader& ih, const DecodedFrame& f);

struct DecodeResult {
  ImageHeader image_header;
  std::vector<DecodedFrame> frames;
};

// Strips an ISOBMFF container if present (crates/jxl-bitstream/src/container*). Returns the bare
// codestream (a copy when boxes had to be concatenated).
std::vector<uint8_t> extract_codestream(const uint8_t* data, size_t size);

// Decodes every keyframe of `codestream` (bare) through `be`. Planes referenced by the result
// stay alive in the backend until the caller frees them.
DecodeResult decode_codestream(Backend& be, const uint8_t* codestream, size_t size, const DecodeOptions& opt);

}  // namespace jxlb
