// Shared helpers for the dexnerf CUDA translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dexnerf.h"

namespace dexnerf {

void set_error(const char* fmt, ...);

// composite.cu: dexnerf_volume_render with an explicit stride between the threshold planes of the Dex outputs
int volume_render_impl(const float* rf, const float* z, const float* rd, const float* noise, int64_t n, int S,
                       int white_background, const float* thresholds, int T, float* rgb, float* disp, float* acc,
                       float* weights, float* depth, float* dex_depth, int64_t* dex_index, int64_t dex_stride,
                       void* stream);

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


// ---------------------------------------------------------------- camera (get_ray_bundle, nerf/nerf_helpers.py:67-112)
__device__ inline void invert4x4_f64(const float* T, double inv[16]) {
  double a[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      a[r][c] = (double)T[r * 4 + c];
      a[r][c + 4] = (r == c) ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; ++col) {
    int piv = col;
    double best = fabs(a[col][col]);
    for (int r = col + 1; r < 4; ++r)
      if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
    if (piv != col)
      for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
    const double d = 1.0 / a[col][col];
    for (int c = 0; c < 8; ++c) a[col][c] *= d;
    for (int r = 0; r < 4; ++r)
      if (r != col) {
        const double f = a[r][col];
        for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
      }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[r * 4 + c] = a[r][c + 4];
}

__device__ inline void invert3x3_f64(const float* T /*4x4, top-left block*/, double inv[9]) {
  const double a = T[0], b = T[1], c = T[2], d = T[4], e = T[5], f = T[6], g = T[8], h = T[9],
               i = T[10];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  const double r = 1.0 / det;
  inv[0] = A * r; inv[1] = -(b * i - c * h) * r; inv[2] = (b * f - c * e) * r;
  inv[3] = B * r; inv[4] = (a * i - c * g) * r;  inv[5] = -(a * f - c * d) * r;
  inv[6] = C * r; inv[7] = -(a * h - b * g) * r; inv[8] = (a * e - b * d) * r;
}


// rinv = inverse of the rotation block, org = inverse(T)[:3, 3] (both through fp64, rounded once)
__device__ inline void camera_inverse(const float* T, float rinv[9], float org[3]) {
  double inv4[16], inv3[9];
  invert4x4_f64(T, inv4);
  invert3x3_f64(T, inv3);
  for (int k = 0; k < 9; ++k) rinv[k] = (float)inv3[k];
  for (int k = 0; k < 3; ++k) org[k] = (float)inv4[k * 4 + 3];
}

// world direction of pixel (row r, column c): R^-1 [(c - cx) / fx, (r - cy) / fx, 1] - the reference divides
// BOTH axes by fx (nerf_helpers.py:100-101)
__device__ __forceinline__ void pixel_direction(int r, int c, float fx, float cx, float cy, const float* rinv,
                                                float d[3]) {
  const float dx = __fdiv_rn(__fsub_rn((float)c, cx), fx);
  const float dy = __fdiv_rn(__fsub_rn((float)r, cy), fx);
#pragma unroll
  for (int a = 0; a < 3; ++a)
    d[a] = __fadd_rn(__fadd_rn(__fmul_rn(dx, rinv[a * 3 + 0]), __fmul_rn(dy, rinv[a * 3 + 1])), rinv[a * 3 + 2]);
}

// ---------------------------------------------------------------- stratified depths (nerf/train_utils.py:104-133)
__device__ __forceinline__ float coarse_depth(float near, float far, int Nc, int lindisp, int i) {
  const float t = linspace_at(0.0f, 1.0f, Nc, i);
  const float omt = __fsub_rn(1.0f, t);
  if (!lindisp) return __fadd_rn(__fmul_rn(near, omt), __fmul_rn(far, t));
  return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near), omt),
                                   __fmul_rn(__fdiv_rn(1.0f, far), t)));
}


// z_i of a ray: the linspace depth, or with `perturbed` the jittered depth lower + (upper - lower) * t
// between the mid-points around it (train_utils.py:126-133)
__device__ __forceinline__ float stratified_depth(float near, float far, int Nc, int lindisp, int i, bool perturbed,
                                                  float t) {
  float v = coarse_depth(near, far, Nc, lindisp, i);
  if (perturbed) {
    const float prev = i > 0 ? coarse_depth(near, far, Nc, lindisp, i - 1) : v;
    const float next = i < Nc - 1 ? coarse_depth(near, far, Nc, lindisp, i + 1) : v;
    const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(v, prev)) : v;
    const float upper = i < Nc - 1 ? __fmul_rn(0.5f, __fadd_rn(next, v)) : v;
    v = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
  }
  return v;
}

}  // namespace dexnerf
