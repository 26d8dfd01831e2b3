// Validation-time depth error metrics over the Dex-NeRF threshold planes, on the device.
//
// Reference: nerf/train_utils.py:9-30 compute_err_metric and its use in the validation block of
// train_dexnerf_rgb.py:391-404: for every threshold candidate the predicted depth plane is moved to
// the host and reduced there (T round trips of H*W floats); the threshold with the smallest mean
// absolute error wins.  Here ONE pass over the (T, n) planes produces, per threshold,
//   [ mean |pred - gt| in millimetres,  fraction > 2 mm,  fraction > 4 mm,  fraction > 8 mm ]
// over the masked pixels, so that only 4*T floats ever leave the GPU.
//
// The mask is either explicit (uint8, n) or the reference's rule (gt > 0) & (gt < 1.25)
// (train_dexnerf_rgb.py:392).  HBM-bound: 4*T + 4 (+1) bytes per pixel.
#include "common.cuh"

namespace dexnerf {

constexpr int kMetricThreads = 256;

// partial[t][0..3] accumulated with fp64 atomics: sum |diff| * 1000, count > 2mm, > 4mm, > 8mm; partial[T][0] = pixels
__global__ void __launch_bounds__(kMetricThreads)
depth_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const uint8_t* __restrict__ mask,
                     int64_t n, int T, double* __restrict__ partial) {
  __shared__ double s_red[kMetricThreads / 32][4];
  const int t = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sum = 0.0;
  unsigned c2 = 0, c4 = 0, c8 = 0, cnt = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = gt[i];
    const bool m = mask ? (mask[i] != 0) : (g > 0.0f && g < 1.25f);
    if (m) {
      const float p = pred[(int64_t)t * n + i];
      // F.l1_loss(pred * 1000, gt * 1000): the products are rounded to fp32 first
      sum += (double)fabsf(__fsub_rn(__fmul_rn(p, 1000.0f), __fmul_rn(g, 1000.0f)));
      const float d = fabsf(__fsub_rn(g, p));
      c2 += d > 2e-3f; c4 += d > 4e-3f; c8 += d > 8e-3f; ++cnt;
    }
  }
  double v[4] = {sum, (double)c2, (double)c4, (double)c8};
  double vc = (double)cnt;
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = warp_sum_f64(v[k]);
  vc = warp_sum_f64(vc);
  if (lane == 0) { for (int k = 0; k < 4; ++k) s_red[warp][k] = v[k]; }
  __shared__ double s_cnt[kMetricThreads / 32];
  if (lane == 0) s_cnt[warp] = vc;
  __syncthreads();
  if (threadIdx.x < 4) {
    double a = 0.0;
    for (int w = 0; w < kMetricThreads / 32; ++w) a += s_red[w][threadIdx.x];
    atomicAdd(&partial[t * 4 + threadIdx.x], a);
  }
  if (threadIdx.x == 4 && t == 0) {
    double a = 0.0;
    for (int w = 0; w < kMetricThreads / 32; ++w) a += s_cnt[w];
    atomicAdd(&partial[T * 4], a);
  }
}

__global__ void depth_metrics_finish_kernel(const double* __restrict__ partial, int T, float* __restrict__ out,
                                            int32_t* __restrict__ best) {
  // one warp: normalise and pick the first threshold with the smallest mean abs error (strict <, as the
  // reference's loop at train_dexnerf_rgb.py:396-404)
  const double cnt = partial[T * 4];
  float best_err = 1000.0f;    // min_abs_err starts at 1000. (train_dexnerf_rgb.py:394)
  int best_t = -1;
  for (int t = 0; t < T; ++t) {
    const float abs_err = (float)(partial[t * 4] / cnt);
    if (threadIdx.x == 0) {
      out[t * 4 + 0] = abs_err;
      out[t * 4 + 1] = (float)(partial[t * 4 + 1] / cnt);
      out[t * 4 + 2] = (float)(partial[t * 4 + 2] / cnt);
      out[t * 4 + 3] = (float)(partial[t * 4 + 3] / cnt);
    }
    if (abs_err < best_err) { best_err = abs_err; best_t = t; }
  }
  if (threadIdx.x == 0 && best) *best = best_t;
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_depth_error_metrics(const float* pred, const float* gt, const uint8_t* mask,
                                                       int64_t n, int T, float* out, int32_t* best,
                                                       void* workspace, void* stream) {
  DN_REQUIRE(pred && gt && out && workspace, "depth_error_metrics: null pointer");
  DN_REQUIRE(T >= 1 && n >= 1, "depth_error_metrics: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  DN_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * (size_t)(T * 4 + 1), st));
  int64_t blocks = ceil_div64(n, kMetricThreads * 4);
  const int64_t cap = (int64_t)kNumSMs * 8 / (T < 8 ? T : 8) + 1;
  if (blocks > cap) blocks = cap;
  depth_metrics_kernel<<<dim3((unsigned)blocks, (unsigned)T), kMetricThreads, 0, st>>>(
      pred, gt, mask, n, T, reinterpret_cast<double*>(workspace));
  DN_CHECK_LAUNCH("depth_metrics");
  depth_metrics_finish_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const double*>(workspace), T, out, best);
  DN_CHECK_LAUNCH("depth_metrics_finish");
  return 0;
}
