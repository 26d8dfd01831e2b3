// Ray generation, NDC warp, positional encoding, stratified depths.
// HBM-bound elementwise kernels: one thread per output element, coalesced stores, grids sized in
// whole waves of the 148 SMs.  fp32 arithmetic is written with explicit rounding intrinsics so
// that it follows the operation order of the oracle (oracle/nerf_oracle.py) and is not
// re-associated or FMA-contracted by the compiler.
#include "common.cuh"

namespace dexnerf {

// ---------------------------------------------------------------- get_ray_bundle
// Reference: nerf/nerf_helpers.py:67-112 (+ meshgrid_xy :28-40).
__global__ void __launch_bounds__(256) ray_bundle_kernel(const float* __restrict__ T,
                                                         const float* __restrict__ K, int W,
                                                         int row0, int64_t n_pix,
                                                         float* __restrict__ ro,
                                                         float* __restrict__ rd) {
  __shared__ float s_rinv[9];
  __shared__ float s_org[3];
  if (threadIdx.x == 0) camera_inverse(T, s_rinv, s_org);
  __syncthreads();
  const float fx = K[0], cx = K[2], cy = K[5];
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix;
       p += (int64_t)gridDim.x * blockDim.x) {
    float d[3];
    pixel_direction((int)(p / W) + row0, (int)(p % W), fx, cx, cy, s_rinv, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      rd[p * 3 + a] = d[a];
      ro[p * 3 + a] = s_org[a];
    }
  }
}

// ---------------------------------------------------------------- ndc_rays
// Reference: nerf/nerf_helpers.py:172-199.
__global__ void __launch_bounds__(256) ndc_kernel(const float* __restrict__ ro,
                                                  const float* __restrict__ rd, int64_t n, float sx,
                                                  float sy, float near, float* __restrict__ oo,
                                                  float* __restrict__ od) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float ox = ro[i * 3], oy = ro[i * 3 + 1], oz = ro[i * 3 + 2];
    const float dx = rd[i * 3], dy = rd[i * 3 + 1], dz = rd[i * 3 + 2];
    const float t = __fdiv_rn(-__fadd_rn(near, oz), dz);
    const float px = __fadd_rn(ox, __fmul_rn(t, dx));
    const float py = __fadd_rn(oy, __fmul_rn(t, dy));
    const float pz = __fadd_rn(oz, __fmul_rn(t, dz));
    oo[i * 3 + 0] = __fdiv_rn(__fmul_rn(sx, px), pz);
    oo[i * 3 + 1] = __fdiv_rn(__fmul_rn(sy, py), pz);
    oo[i * 3 + 2] = __fadd_rn(1.0f, __fdiv_rn(__fmul_rn(2.0f, near), pz));
    od[i * 3 + 0] = __fmul_rn(sx, __fsub_rn(__fdiv_rn(dx, dz), __fdiv_rn(px, pz)));
    od[i * 3 + 1] = __fmul_rn(sy, __fsub_rn(__fdiv_rn(dy, dz), __fdiv_rn(py, pz)));
    od[i * 3 + 2] = __fdiv_rn(__fmul_rn(-2.0f, near), pz);
  }
}

// ---------------------------------------------------------------- positional_encoding
// Reference: nerf/nerf_helpers.py:115-159.  Column c of the output:
//   [x(3)] if include_input, then for band b: sin(f_b x)(3), cos(f_b x)(3).
__global__ void __launch_bounds__(256) posenc_kernel(const float* __restrict__ x, int64_t M, int L,
                                                     int include_input, int log_sampling, int D,
                                                     float* __restrict__ out) {
  const int64_t total = M * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = e / D;
    const int c = (int)(e - m * D);
    const float v[3] = {x[m * 3], x[m * 3 + 1], x[m * 3 + 2]};
    out[e] = pe_column(v, c, L, include_input, log_sampling);
  }
}

// ---------------------------------------------------------------- stratified depths
// Reference: nerf/train_utils.py:104-133.
__global__ void __launch_bounds__(256) stratified_kernel(int64_t n, int Nc, float near_s, float far_s,
                                                         const float* __restrict__ near_a,
                                                         const float* __restrict__ far_a, int lindisp,
                                                         const float* __restrict__ t_rand,
                                                         float* __restrict__ z) {
  const int64_t total = n * Nc;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / Nc;
    const int i = (int)(e - r * Nc);
    const float near = near_a ? near_a[r] : near_s;
    const float far = far_a ? far_a[r] : far_s;
    z[e] = stratified_depth(near, far, Nc, lindisp, i, t_rand != nullptr, t_rand ? t_rand[e] : 0.0f);
  }
}

static inline int elementwise_grid(int64_t total) {
  int64_t blocks = ceil_div64(total, 256);
  const int64_t wave = (int64_t)kNumSMs * 8;  // 8 resident 256-thread CTAs per SM
  if (blocks > wave) blocks = wave * ((blocks + wave - 1) / wave > 4 ? 4 : (blocks + wave - 1) / wave);
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_ray_bundle(const float* T_w2c, const float* K, int H, int W, int row0,
                                  int rows, float* ro, float* rd, void* stream) {
  DN_REQUIRE(T_w2c && K && ro && rd, "ray_bundle: null pointer");
  DN_REQUIRE(H > 0 && W > 0 && row0 >= 0 && rows >= 0 && row0 + rows <= H, "ray_bundle: bad rows");
  const int64_t n = (int64_t)rows * W;
  if (n == 0) return 0;
  ray_bundle_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(T_w2c, K, W, row0, n, ro, rd);
  DN_CHECK_LAUNCH("ray_bundle");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_ndc_rays(const float* ro, const float* rd, int64_t n, int H, int W,
                                float focal, float near, float* ro_out, float* rd_out,
                                void* stream) {
  DN_REQUIRE(ro && rd && ro_out && rd_out, "ndc_rays: null pointer");
  if (n <= 0) return 0;
  // -1 / (W / (2 focal)) evaluated in double like the Python scalar expression, then rounded
  const float sx = (float)(-1.0 / ((double)W / (2.0 * (double)focal)));
  const float sy = (float)(-1.0 / ((double)H / (2.0 * (double)focal)));
  ndc_kernel<<<elementwise_grid(n), 256, 0, (cudaStream_t)stream>>>(ro, rd, n, sx, sy, near, ro_out, rd_out);
  DN_CHECK_LAUNCH("ndc_rays");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_positional_encoding(const float* x, int64_t M, int L, int include_input,
                                           int log_sampling, float* out, void* stream) {
  if (M <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(x && out, "positional_encoding: null pointer");
  DN_REQUIRE(L >= 0 && L <= 31, "positional_encoding: L out of range");
  const int D = (include_input ? 3 : 0) + 6 * L;
  DN_REQUIRE(D > 0, "positional_encoding: empty encoding");
  posenc_kernel<<<elementwise_grid(M * D), 256, 0, (cudaStream_t)stream>>>(x, M, L, include_input,
                                                                         log_sampling, D, out);
  DN_CHECK_LAUNCH("positional_encoding");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_stratified_z(int64_t n, int Nc, float near, float far, const float* near_arr,
                                    const float* far_arr, int lindisp, const float* t_rand, float* z,
                                    void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(z, "stratified_z: null output");
  DN_REQUIRE(Nc >= 1, "stratified_z: Nc < 1");
  stratified_kernel<<<elementwise_grid(n * Nc), 256, 0, (cudaStream_t)stream>>>(
      n, Nc, near, far, near_arr, far_arr, lindisp, t_rand, z);
  DN_CHECK_LAUNCH("stratified_z");
  return 0;
}
