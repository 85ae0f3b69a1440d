// TEST INFRASTRUCTURE. Lets the per-stream device code of the entropy kernels (plain integer C++ with no
// cross-lane traffic: kernels/common.cuh, stream_common.cuh, hf_lanes.cuh) compile for the host, so that its
// logic can be checked against the oracle in the GPU-less container. <cuda_runtime.h> already turns
// __device__ / __constant__ / __forceinline__ into host-side no-ops under g++; only the intrinsics are missing.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

template <typename T>
static inline T __ldg(const T* p) {
  return *p;
}
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz(static_cast<unsigned>(v)); }
static inline int32_t max(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t min(int32_t a, int32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
