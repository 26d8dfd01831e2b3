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

// Depth-error colour image of the validation block (nerf/train_utils.py:31-70 gen_error_colormap_depth +
// depth_error_img): error = |gt - est| / abs_thres on the masked pixels, coloured by eleven buckets
// [0, 1e-5), [1e-5, 2000/2^10), [2000/2^10, 2000/2^9), ... [2000/2^2, inf); unmasked pixels black; the
// colour legend in the top-left corner (10 rows, 20 columns per bucket).  One thread per pixel, 12 B written.
__constant__ float kErrEdges[12] = {0.0f, 0.00001f, 2000.0f / 1024, 2000.0f / 512, 2000.0f / 256, 2000.0f / 128,
                                    2000.0f / 64, 2000.0f / 32, 2000.0f / 16, 2000.0f / 8, 2000.0f / 4, 0.0f};
__constant__ float kErrRgb[11][3] = {{0, 0, 0},       {49, 54, 149},   {69, 117, 180}, {116, 173, 209},
                                     {171, 217, 233}, {224, 243, 248}, {254, 224, 144}, {253, 174, 97},
                                     {244, 109, 67},  {215, 48, 39},   {165, 0, 38}};

__global__ void __launch_bounds__(256)
depth_error_image_kernel(const float* __restrict__ est, const float* __restrict__ gt, const uint8_t* __restrict__ mask,
                         int H, int W, float abs_thres, float* __restrict__ out) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / W), c = (int)(i - (int64_t)r * W);
    int bucket = -1;
    if (r < 10 && c < 11 * 20) {
      bucket = c / 20;                                   // the legend overwrites the image
    } else if (mask[i]) {
      const float e = __fdiv_rn(fabsf(__fsub_rn(gt[i], est[i])), abs_thres);
      for (int b = 0; b < 11; ++b) {
        const bool below_hi = (b == 10) ? (e < __int_as_float(0x7f800000)) : (e < kErrEdges[b + 1]);
        if (e >= kErrEdges[b] && below_hi) bucket = b;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch)
      out[i * 3 + ch] = bucket < 0 ? 0.0f : __fdiv_rn(kErrRgb[bucket][ch], 255.0f);
  }
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_depth_error_image(const float* est, const float* gt, const uint8_t* mask, int H,
                                                     int W, float abs_thres, float* out, void* stream) {
  if (H <= 0 || W <= 0) return 0;
  DN_REQUIRE(est && gt && mask && out, "depth_error_image: null pointer");
  int64_t blocks = ceil_div64((int64_t)H * W, 256);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  depth_error_image_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(est, gt, mask, H, W, abs_thres, out);
  DN_CHECK_LAUNCH("depth_error_image");
  return 0;
}

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
