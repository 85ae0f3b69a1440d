// TEST INFRASTRUCTURE — see oracle_backend.h. Constant tables + lazily built helpers.
#pragma once
#include <cmath>
#include <cstdint>
#include <mutex>
#include <vector>

#include "../jxl_oxide_b200/csrc/host/frame_syntax.h"

namespace jxlo {

#define JXLB_TABLE_QUAL static const
#include "../jxl_oxide_b200/csrc/host/jxl_tables.inc"
#undef JXLB_TABLE_QUAL

// sec_half (dct_common.rs:45-69): n >= 64 tables are computed at run time in f32.
inline const float* sec_half(size_t n) {
  switch (n) {
    case 4: return kSecHalf4;
    case 8: return kSecHalf8;
    case 16: return kSecHalf16;
    case 32: return kSecHalf32;
    default: break;
  }
  static std::once_flag once;
  static std::vector<float> large[3];  // 64, 128, 256
  std::call_once(once, [] {
    for (int i = 0; i < 3; ++i) {
      size_t nn = size_t(64) << i;
      large[i].resize(nn / 2);
      for (size_t k = 0; k < nn / 2; ++k) {
        float theta = float(2 * k + 1) / float(2 * nn) * 3.14159265358979323846f;
        large[i][k] = (1.0f / cosf(theta)) / 2.0f;
      }
    }
  });
  return large[n == 64 ? 0 : (n == 128 ? 1 : 2)].data();
}

inline const std::vector<uint32_t>& natural_order_cached(uint32_t order_id) {
  static std::once_flag once;
  static std::vector<uint32_t> orders[13];
  std::call_once(once, [] {
    for (uint32_t i = 0; i < 13; ++i) orders[i] = jxlb::natural_order(i);
  });
  return orders[order_id];
}

}  // namespace jxlo
