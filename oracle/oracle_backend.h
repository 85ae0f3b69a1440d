// TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.
//
// CPU oracle: a scalar restatement of jxl-oxide's sample-level decode stages (the *generic*,
// portable code paths — SURVEY.md §7 "Hard parts"), used only by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs to check and
// time-compare the CUDA path. The product library never links or calls anything in oracle/.
//
// It plugs into the product's host-side planner (jxl_oxide_b200/csrc/host/planner.h) through the
// `Backend` seam, so syntax parsing is shared while every sample is computed here on the CPU.
// Parity status: integer/Modular path pinned by the reference's own fixtures
// (crates/jxl-oxide-tests/decode/{issue_311,squeeze_edge,grayalpha}); VarDCT float path pinned
// only at 8/16-bit by conformance ref.png — see DESIGN.md "Oracle".
#pragma once
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <exception>
#include <map>
#include <thread>
#include <string>
#include <vector>

#include "../jxl_oxide_b200/csrc/host/backend.h"

namespace jxlo {

using namespace jxlb;

struct Plane {
  uint32_t w = 0, h = 0;
  std::vector<uint32_t> data;  // i32 or f32 bit patterns
  int32_t* i32() { return reinterpret_cast<int32_t*>(data.data()); }
  float* f32() { return reinterpret_cast<float*>(data.data()); }
};

class OracleBackend : public Backend {
 public:
  explicit OracleBackend(int num_threads = 1) : threads_(num_threads < 1 ? 1 : num_threads) {}
  void set_codestream(const uint8_t* data, size_t size) override {
    cs_ = data;
    cs_size_ = size;
  }
  int alloc_plane(uint32_t w, uint32_t h, bool zero) override;
  void free_plane(int id) override;
  void download_rect(const View& v, void* dst) override;
  void copy_rect(const View& src, const View& dst) override;
  void decode_modular(std::vector<ModularStreamJob>& jobs) override;
  int squeeze_inverse(const View& avg, const View& residual, bool horizontal) override;
  void rct_inverse(const View v[3], uint32_t rct_type) override;
  void palette_inverse(const View& palette, const std::vector<View>& targets, const Transform& t,
                       const WpHeader& wp, uint32_t bit_depth) override;
  void int_to_float(const View& v, const BitDepth& depth) override;
  void modular_xyb_to_float(const View yxb[3], const float m_lf_unscaled[3]) override;
  void build_block_info(VarDctState& st, const std::vector<BlockInfoJob>& jobs) override;
  void decode_hf(VarDctState& st, std::vector<HfGroupJob>& jobs) override;
  void lf_dequant(VarDctState& st, const std::vector<LfDequantJob>& jobs) override;
  void lf_chroma_from_luma(VarDctState& st) override;
  void lf_adaptive_smoothing(VarDctState& st) override;
  void hf_dequant_cfl(VarDctState& st) override;
  void hf_transform(VarDctState& st) override;
  void gaborish(const View v[3], const float weights[3][2]) override;
  void epf(const View v[3], const View& sigma, const EpfParams& p, bool sigma_is_constant) override;
  int upsample(const View& v, uint32_t factor_log2, const ImageHeader& ih) override;
  int upsample_jpeg(const View& v, bool horizontal, bool vertical, uint32_t out_w, uint32_t out_h) override;
  void blend_patches(const std::vector<PatchJob>& jobs) override;
  void splat_splines(const View v[3], const std::vector<SplineArc>& arcs) override;
  void add_noise(const View v[3], const float lut[8], uint32_t group_dim, uint64_t seed0, float corr_x, float corr_b) override;
  void xyb_to_rgb(const View v[3], const ColorParams& p) override;
  void ycbcr_to_rgb(const View v[3], const YcbcrParams& p) override;
  void stage_marker(const char* name, const View* views, int n) override;
  // JXLO_TRACE=1: wall time of every host phase of the planner on stderr (where the CPU time of a frame goes)
  void phase_mark(const char* name) override {
    static const bool on = std::getenv("JXLO_TRACE") != nullptr;
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    if (name) std::fprintf(stderr, "[oracle] %-20s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(now - phase_t0_).count());
    phase_t0_ = now;
  }

  Plane& plane(int id) { return planes_.at(id); }
  // Stage snapshots (tightly packed rects), filled when capture is on.
  bool capture = false;
  std::map<std::string, std::vector<std::vector<uint32_t>>> stages;
  std::map<std::string, std::vector<std::pair<uint32_t, uint32_t>>> stage_dims;

  template <typename F>
  void parallel_for(size_t n, F f) {
    size_t nt = size_t(threads_) < n ? size_t(threads_) : n;
    if (nt <= 1) {
      for (size_t i = 0; i < n; ++i) f(i);
      return;
    }
    std::vector<std::thread> pool;
    std::vector<std::exception_ptr> errs(nt);
    std::atomic<size_t> next(0);
    for (size_t t = 0; t < nt; ++t)
      pool.emplace_back([&, t] {
        try {
          for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) break;
            f(i);
          }
        } catch (...) {
          errs[t] = std::current_exception();
        }
      });
    for (auto& th : pool) th.join();
    for (auto& e : errs)
      if (e) std::rethrow_exception(e);
  }

 private:
  void decode_one_modular(ModularStreamJob& job);
  void decode_one_hf(VarDctState& st, HfGroupJob& job);
  const uint8_t* cs_ = nullptr;
  size_t cs_size_ = 0;
  std::map<int, Plane> planes_;
  int next_id_ = 0;
  int threads_;
  std::chrono::steady_clock::time_point phase_t0_ = std::chrono::steady_clock::now();
};

float linear_to_pq(float s, float intensity_target);  // oracle_render.cc

}  // namespace jxlo
