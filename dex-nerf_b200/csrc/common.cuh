// Shared helpers for the dexnerf CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dexnerf.h"

namespace dexnerf {

void set_error(const char* fmt, ...);

#define DN_REQUIRE(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      ::dexnerf::set_error(__VA_ARGS__);       \
      return DEXNERF_E_INVALID;                \
    }                                          \
  } while (0)

#define DN_CHECK_LAUNCH(name)                                                           \
  do {                                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      ::dexnerf::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));      \
      return DEXNERF_E_CUDA;                                                            \
    }                                                                                   \
  } while (0)

#define DN_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ::dexnerf::set_error("%s failed: %s", #call, cudaGetErrorString(e_));             \
      return DEXNERF_E_CUDA;                                                            \
    }                                                                                   \
  } while (0)

constexpr int kNumSMs = 148;  // B200

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// torch.linspace(start, end, steps)[i] in fp32 as ATen's CPU kernel evaluates it: fp32 step,
// symmetric about the middle, one fused rounding per element (checked bit-for-bit against torch
// in tests/test_host_logic.py).
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
  if (steps == 1) return start;
  const float step = __fdiv_rn(__fsub_rn(end, start), (float)(steps - 1));
  return (i < steps / 2) ? __fmaf_rn(step, (float)i, start)
                         : __fmaf_rn(-step, (float)(steps - 1 - i), end);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// positional_encoding column c for a 3-vector x (nerf/nerf_helpers.py:115-159):
//   [x(3)] if include_input, then for band b: sin(f_b x)(3), cos(f_b x)(3).
__device__ __forceinline__ float pe_frequency(int band, int L, int log_sampling) {
  if (log_sampling) return (float)(1u << band);                       // 2 ** linspace(0, L-1, L)
  return linspace_at(1.0f, (float)(1u << (L - 1)), L, band);          // linspace(1, 2^(L-1), L)
}

__device__ __forceinline__ float pe_column(const float x[3], int c, int L, int include_input,
                                           int log_sampling) {
  if (include_input) {
    if (c < 3) return x[c];
    c -= 3;
  }
  const int band = c / 6, rem = c - band * 6;
  const float arg = __fmul_rn(x[rem % 3], pe_frequency(band, L, log_sampling));
  return rem < 3 ? sinf(arg) : cosf(arg);
}

}  // namespace dexnerf
