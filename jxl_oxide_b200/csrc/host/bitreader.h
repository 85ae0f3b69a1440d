// LSB-first bit reader used by the host-side syntax parser.
//
// Semantics follow the reference `Bitstream` (crates/jxl-bitstream/src/bitstream.rs:9-207):
// bits are consumed LSB-first from little-endian bytes; reading past the end of the
// buffer yields zero bits and is reported as an error when `check()` is called
// (the reference raises UnexpectedEof from `consume_bits`, bitstream.rs:133-141).
// The position-based formulation here is equivalent for every in-bounds read.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>

namespace jxlb {

enum ErrorCode : int {
  kOk = 0,
  kErrBitstream = 1,      // malformed codestream
  kErrUnsupported = 2,    // valid but outside the implemented hot path
  kErrEof = 3,            // truncated input
  kErrCuda = 4,
  kErrInvalidArg = 5,
  kErrDeviceDecode = 6,   // a device-side stream decoder flagged an invalid stream
  kErrOutOfMemory = 7,    // the decoder's allocation budget (jxlb_decoder_create_ex) would be exceeded
};

struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Error(code, msg); }
#define JXLB_CHECK(cond, code, msg) \
  do {                              \
    if (!(cond)) ::jxlb::fail((code), (msg)); \
  } while (0)

class BitReader {
 public:
  BitReader() : data_(nullptr), size_(0), pos_(0) {}
  BitReader(const uint8_t* data, size_t size_bytes, size_t pos_bits = 0)
      : data_(data), size_(size_bytes), pos_(pos_bits) {}

  size_t pos() const { return pos_; }
  size_t size_bits() const { return size_ * 8; }
  const uint8_t* data() const { return data_; }
  size_t size_bytes() const { return size_; }
  bool overrun() const { return pos_ > size_ * 8; }
  void check() const { JXLB_CHECK(!overrun(), kErrEof, "unexpected end of bitstream"); }

  // Up to 56 valid bits starting at the current position (zero past the end).
  inline uint64_t peek64() const {
    size_t byte = pos_ >> 3;
    uint64_t v = 0;
    if (byte + 8 <= size_) {
      std::memcpy(&v, data_ + byte, 8);
    } else if (byte < size_) {
      std::memcpy(&v, data_ + byte, size_ - byte);
    }
    return v >> (pos_ & 7);
  }
  inline uint32_t peek(uint32_t n) const {  // n <= 32
    return static_cast<uint32_t>(peek64() & ((n >= 32) ? 0xffffffffull : ((1ull << n) - 1)));
  }
  inline void consume(uint32_t n) { pos_ += n; }
  inline uint32_t read(uint32_t n) {
    uint32_t v = peek(n);
    pos_ += n;
    return v;
  }
  inline bool read_bool() { return read(1) != 0; }
  void skip(size_t n) { pos_ += n; }
  void seek(size_t bit) { pos_ = bit; }

  // ZeroPadToByte (bitstream.rs:198-206)
  void zero_pad_to_byte() {
    uint32_t n = static_cast<uint32_t>((8 - (pos_ & 7)) & 7);
    uint32_t v = read(n);
    JXLB_CHECK(v == 0, kErrBitstream, "non-zero padding bits");
  }

  // U32(d0,d1,d2,d3): each distribution is (offset, nbits); bitstream.rs:223-244
  struct U32Dist {
    uint32_t offset;
    uint32_t bits;
  };
  uint32_t read_u32(U32Dist d0, U32Dist d1, U32Dist d2, U32Dist d3) {
    U32Dist d[4] = {d0, d1, d2, d3};
    uint32_t sel = read(2);
    return d[sel].offset + (d[sel].bits ? read(d[sel].bits) : 0);  // wrapping add
  }
  // bitstream.rs:247-267
  uint64_t read_u64() {
    uint32_t sel = read(2);
    switch (sel) {
      case 0: return 0;
      case 1: return uint64_t(read(4)) + 1;
      case 2: return uint64_t(read(8)) + 17;
      default: {
        uint64_t value = read(12);
        uint32_t shift = 12;
        while (read(1) == 1) {
          if (shift == 60) {
            value |= uint64_t(read(4)) << shift;
            break;
          }
          value |= uint64_t(read(8)) << shift;
          shift += 8;
          if (overrun()) break;
        }
        return value;
      }
    }
  }
  // bitstream.rs:279-305
  float read_f16() {
    uint32_t v = read(16);
    uint32_t neg = (v & 0x8000u) << 16;
    if ((v & 0x7fff) == 0) {
      float f;
      std::memcpy(&f, &neg, 4);
      return f;
    }
    uint32_t mantissa = v & 0x3ff;
    uint32_t exponent = (v >> 10) & 0x1f;
    JXLB_CHECK(exponent != 0x1f, kErrBitstream, "F16 is NaN or infinity");
    if (exponent == 0) {
      float val = (1.0f / 16384.0f) * (float(mantissa) / 1024.0f);
      return neg ? -val : val;
    }
    uint32_t bits = (mantissa << 13) | ((exponent + 112) << 23) | neg;
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
  }
  // bitstream.rs:308-314
  uint32_t read_enum() { return read_u32({0, 0}, {1, 0}, {2, 4}, {18, 6}); }

 private:
  const uint8_t* data_;
  size_t size_;
  size_t pos_;
};

inline int32_t unpack_signed(uint32_t x) {  // jxl-bitstream/src/lib.rs:24-29
  return static_cast<int32_t>((x >> 1) ^ (0u - (x & 1)));
}

inline uint32_t ceil_log2_nonzero(uint32_t x) {  // next_power_of_two().trailing_zeros()
  uint32_t r = 0;
  while ((1ull << r) < x) ++r;
  return r;
}

}  // namespace jxlb
