// Embedded ICC profiles (crates/jxl-color/src/icc/{decode,parse}.rs): the profile's bytes and, where an enum colour
// encoding describes the profile exactly, that encoding.
#pragma once

#include <cstdint>
#include <vector>

#include "bitreader.h"
#include "headers.h"

namespace jxlb {

// Reconstructs the ICC profile from the entropy-decoded stream read_icc_stream() returns (decode.rs:192-423).
std::vector<uint8_t> decode_icc_stream(const std::vector<uint8_t>& encoded);

enum class IccStatus {
  kEnum,         // `encoding` is equivalent to the profile (parse_icc succeeded)
  kUnsupported,  // a valid profile no enum encoding describes (Error::UnsupportedIccProfile)
  kMalformed,    // Error::IccParseFailure
};
struct IccInfo {
  bool is_gray = false, is_cmyk = false;  // the profile's data colour space
  ColourEncoding encoding;                // valid for IccStatus::kEnum
};
IccStatus icc_to_enum(const std::vector<uint8_t>& profile, IccInfo* info);

}  // namespace jxlb
