// TEST INFRASTRUCTURE — see oracle_backend.h. C entry points for ctypes (tests/, bench.py).
#include <cstring>
#include <memory>
#include <string>

#include "../jxl_oxide_b200/csrc/host/planner.h"
#include "oracle_backend.h"

namespace {
struct Handle {
  std::unique_ptr<jxlo::OracleBackend> be;
  jxlb::DecodeResult res;
  std::vector<uint8_t> codestream;
};
void set_err(char* err, size_t n, const std::string& s) {
  if (err && n) {
    std::strncpy(err, s.c_str(), n - 1);
    err[n - 1] = 0;
  }
}
}  // namespace

extern "C" {

void* jxlo_decode(const uint8_t* data, size_t size, int output_colour, int threads, int capture, int* status,
                  char* err, size_t errlen) {
  auto h = std::make_unique<Handle>();
  try {
    h->codestream = jxlb::extract_codestream(data, size);
    h->be.reset(new jxlo::OracleBackend(threads));
    h->be->capture = capture != 0;
    jxlb::DecodeOptions opt;
    opt.output_colour = output_colour;
    h->res = jxlb::decode_codestream(*h->be, h->codestream.data(), h->codestream.size(), opt);
    if (status) *status = 0;
    return h.release();
  } catch (const jxlb::Error& e) {
    if (status) *status = e.code;
    set_err(err, errlen, e.what());
  } catch (const std::exception& e) {
    if (status) *status = -1;
    set_err(err, errlen, e.what());
  }
  return nullptr;
}

int jxlo_num_frames(void* hp) { return int(static_cast<Handle*>(hp)->res.frames.size()); }

void jxlo_image_info(void* hp, uint32_t* width, uint32_t* height, uint32_t* bits, uint32_t* num_extra,
                     uint32_t* xyb, uint32_t* gray) {
  const jxlb::ImageHeader& ih = static_cast<Handle*>(hp)->res.image_header;
  *width = ih.width, *height = ih.height, *bits = ih.bit_depth.bits_per_sample;
  *num_extra = uint32_t(ih.ec_info.size()), *xyb = ih.xyb_encoded, *gray = ih.grayscale();
}

void jxlo_frame_info(void* hp, int frame, uint32_t* width, uint32_t* height, uint32_t* num_channels,
                     uint32_t* num_color, uint32_t* is_vardct) {
  const jxlb::DecodedFrame& f = static_cast<Handle*>(hp)->res.frames.at(frame);
  *width = f.width, *height = f.height, *num_channels = uint32_t(f.channels.size()), *num_color = f.num_color;
  *is_vardct = f.header.encoding == jxlb::Encoding::kVarDct;
}

void jxlo_frame_channel(void* hp, int frame, int channel, float* out) {
  Handle* h = static_cast<Handle*>(hp);
  const jxlb::DecodedFrame& f = h->res.frames.at(frame);
  h->be->download_rect(f.channels.at(channel), out);
}

// Returns the number of planes captured for `name` (0 if absent). With out != NULL copies plane idx.
int jxlo_stage(void* hp, const char* name, int idx, uint32_t* w, uint32_t* hgt, uint32_t* out) {
  Handle* h = static_cast<Handle*>(hp);
  auto it = h->be->stages.find(name);
  if (it == h->be->stages.end()) return 0;
  if (idx >= 0 && idx < int(it->second.size())) {
    *w = h->be->stage_dims[name][idx].first;
    *hgt = h->be->stage_dims[name][idx].second;
    if (out) std::memcpy(out, it->second[idx].data(), it->second[idx].size() * 4);
  }
  return int(it->second.size());
}

void jxlo_free(void* hp) { delete static_cast<Handle*>(hp); }

}  // extern "C"
