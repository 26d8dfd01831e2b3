// Training-loop glue on the device: the MSE loss with its gradient
// (nerf/nerf_helpers.py:9-10 img2mse + autograd, train_dexnerf_rgb.py:264-277) and one fused Adam
// step over the FLAT parameter buffer of both networks (torch.optim.Adam with its defaults,
// train_dexnerf_rgb.py:142-148, 278-289) - the same buffer the weight-gradient GEMM reduces into
// and NCCL all-reduces, so a training iteration touches the 2 x 595 844 parameters in exactly three
// launches (memset, all-reduce, Adam) instead of ~100 per-tensor ones.
#include "common.cuh"

namespace dexnerf {

// loss += mean((pred - target)^2) over n*3 values;  g = 2 (pred - target) / (3 n)  (d loss / d pred)
__global__ void __launch_bounds__(256) mse_loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                            int64_t count, float inv_count, float* __restrict__ g,
                                                            float* __restrict__ loss) {
  double part = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(pred[i], target[i]);
    part += (double)__fmul_rn(d, d);
    g[i] = __fmul_rn(__fmul_rn(2.0f, d), inv_count);
  }
  part = warp_sum_f64(part);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += s[w];
    atomicAdd(loss, (float)(a * (double)inv_count));
  }
}

// the training loss of train_dexnerf_rgb.py:264-277 in one launch: blockIdx.y = 0 / 1 -> coarse / fine prediction against
// the same target; loss3 = [total, coarse, fine] accumulates (chunked batches), g_c / g_f = d loss / d pred
__global__ void __launch_bounds__(256) mse_loss_pair_kernel(const float* __restrict__ pred_c, const float* __restrict__ pred_f,
                                                            const float* __restrict__ target, int64_t count, float inv_count,
                                                            float* __restrict__ g_c, float* __restrict__ g_f,
                                                            float* __restrict__ loss3) {
  const float* pred = blockIdx.y ? pred_f : pred_c;
  float* g = blockIdx.y ? g_f : g_c;
  double part = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(pred[i], target[i]);
    part += (double)__fmul_rn(d, d);
    g[i] = __fmul_rn(__fmul_rn(2.0f, d), inv_count);
  }
  part = warp_sum_f64(part);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < 8; ++w) a += s[w];
    const float v = (float)(a * (double)inv_count);
    atomicAdd(loss3 + 1 + blockIdx.y, v);
    atomicAdd(loss3, v);
  }
}

// torch.optim.Adam (amsgrad = False, weight_decay = 0, maximize = False), element for element:
//   exp_avg = beta1 exp_avg + (1 - beta1) g;  exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) g g
//   denom = sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps;  p -= (lr / (1 - beta1^t)) exp_avg / denom
// grad_scale folds the 1 / world of the data-parallel mean into the same pass.
// kZero: the gradient buffer is cleared as it is consumed, so the next iteration's weight-gradient GEMMs (which
// ACCUMULATE with red.global.add) need no separate memset launch.
template <bool kZero>
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        float beta1, float beta2, float eps, float step_size,
                                                        float bc2_sqrt, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = __fmul_rn(g[i], grad_scale);
    if (kZero) g[i] = 0.0f;
    // exp_avg.lerp_(grad, 1 - beta1) == exp_avg + (1 - beta1) (grad - exp_avg)
    const float mi = __fmaf_rn(1.0f - beta1, __fsub_rn(gi, m[i]), m[i]);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float vi = __fmaf_rn(__fmul_rn(1.0f - beta2, gi), gi, __fmul_rn(v[i], beta2));
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
    p[i] = __fsub_rn(p[i], __fmul_rn(step_size, __fdiv_rn(mi, denom)));
    m[i] = mi;
    v[i] = vi;
  }
}

// The same update on four consecutive parameters per thread (16-byte loads and stores; the flat buffers are 16-byte
// aligned and a multiple of four long): the scalar form ran at 0.5 TB/s (68 us for the 1.19 M parameters of the two
// 8x256 networks - more than the compositing and resampling kernels of a training iteration together).
template <bool kZero>
__global__ void __launch_bounds__(256) adam_step_vec4_kernel(float4* __restrict__ p, float4* __restrict__ g,
                                                             float4* __restrict__ m, float4* __restrict__ v, int64_t n4,
                                                             float beta1, float beta2, float eps, float step_size,
                                                             float bc2_sqrt, float grad_scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gg = g[i];
    float4 pp = p[i], mm = m[i], vv = v[i];
    if (kZero) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float gs[4] = {gg.x, gg.y, gg.z, gg.w};
    float* ps[4] = {&pp.x, &pp.y, &pp.z, &pp.w};
    float* ms[4] = {&mm.x, &mm.y, &mm.z, &mm.w};
    float* vs[4] = {&vv.x, &vv.y, &vv.z, &vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gi = __fmul_rn(gs[k], grad_scale);
      const float mi = __fmaf_rn(1.0f - beta1, __fsub_rn(gi, *ms[k]), *ms[k]);
      const float vi = __fmaf_rn(__fmul_rn(1.0f - beta2, gi), gi, __fmul_rn(*vs[k], beta2));
      const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
      *ps[k] = __fsub_rn(*ps[k], __fmul_rn(step_size, __fdiv_rn(mi, denom)));
      *ms[k] = mi;
      *vs[k] = vi;
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_mse_loss_grad(const float* pred, const float* target, int64_t count,
                                                 int64_t total_count, float* grad, float* loss_accum, void* stream) {
  if (count <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(pred && target && grad && loss_accum, "mse_loss_grad: null pointer");
  DN_REQUIRE(total_count >= count, "mse_loss_grad: total_count < count");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
  mse_loss_grad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, count, 1.0f / (float)total_count, grad,
                                                                      loss_accum);
  DN_CHECK_LAUNCH("mse_loss_grad");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_mse_loss_pair(const float* pred_coarse, const float* pred_fine, const float* target,
                                                 int64_t count, int64_t total_count, float* grad_coarse,
                                                 float* grad_fine, float* loss3, void* stream) {
  if (count <= 0) return 0;
  DN_REQUIRE(pred_coarse && pred_fine && target && grad_coarse && grad_fine && loss3, "mse_loss_pair: null pointer");
  DN_REQUIRE(total_count >= count, "mse_loss_pair: total_count < count");
  int64_t blocks = ceil_div64(count, 256 * 4);
  if (blocks > kNumSMs * 2) blocks = kNumSMs * 2;
  mse_loss_pair_kernel<<<dim3((unsigned)blocks, 2), 256, 0, (cudaStream_t)stream>>>(
      pred_coarse, pred_fine, target, count, 1.0f / (float)total_count, grad_coarse, grad_fine, loss3);
  DN_CHECK_LAUNCH("mse_loss_pair");
  return 0;
}

static int adam_impl(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step, float grad_scale, bool zero, void* stream);

extern "C" DEXNERF_API int dexnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                             int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                                             float grad_scale, void* stream) {
  return adam_impl(params, const_cast<float*>(grads), exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale,
                   false, stream);
}

extern "C" DEXNERF_API int dexnerf_adam_step_zero_grad(float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                                       int64_t n, float lr, float beta1, float beta2, float eps,
                                                       int64_t step, float grad_scale, void* stream) {
  return adam_impl(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, true, stream);
}

static int adam_impl(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, int64_t step, float grad_scale, bool zero, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(params && grads && exp_avg && exp_avg_sq, "adam_step: null pointer");
  DN_REQUIRE(step >= 1, "adam_step: step counts from 1");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const bool vec = n % 4 == 0 && ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                                   reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
  if (vec) {
    int64_t vb = ceil_div64(n / 4, 256);
    if (vb > kNumSMs * 8) vb = kNumSMs * 8;
    auto p4 = reinterpret_cast<float4*>(params), g4 = reinterpret_cast<float4*>(grads);
    auto m4 = reinterpret_cast<float4*>(exp_avg), v4 = reinterpret_cast<float4*>(exp_avg_sq);
    if (zero)
      adam_step_vec4_kernel<true><<<(int)vb, 256, 0, (cudaStream_t)stream>>>(p4, g4, m4, v4, n / 4, beta1, beta2, eps,
                                                                            step_size, bc2_sqrt, grad_scale);
    else
      adam_step_vec4_kernel<false><<<(int)vb, 256, 0, (cudaStream_t)stream>>>(p4, g4, m4, v4, n / 4, beta1, beta2, eps,
                                                                             step_size, bc2_sqrt, grad_scale);
    DN_CHECK_LAUNCH("adam_step");
    return 0;
  }
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (zero)
    adam_step_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, beta1,
                                                                          beta2, eps, step_size, bc2_sqrt, grad_scale);
  else
    adam_step_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, beta1,
                                                                           beta2, eps, step_size, bc2_sqrt, grad_scale);
  DN_CHECK_LAUNCH("adam_step");
  return 0;
}

// ---- nn.Linear parameters -> the flat program-layout buffer, one launch.
// The drop-in training loop (torch.optim over nn.Parameters) changes every weight each iteration;
// re-packing them with per-tensor transposes + cat costs ~100 small launches.  Here a device table of
// (weight pointer, bias pointer) per layer is walked by one kernel: Wt[in][out] = W[out][in].
namespace dexnerf {
struct PackParamsTable {
  int n_ops;
  int in_dim[DEXNERF_MAX_OPS], out_dim[DEXNERF_MAX_OPS];
  int64_t w_off[DEXNERF_MAX_OPS], b_off[DEXNERF_MAX_OPS];
};

__global__ void __launch_bounds__(256) pack_params_kernel(const float* const* __restrict__ ptrs,
                                                          const __grid_constant__ PackParamsTable Q,
                                                          float* __restrict__ flat) {
  __shared__ float tile[32][33];
  for (int op = blockIdx.y; op < Q.n_ops; op += gridDim.y) {
    const float* W = ptrs[2 * op];
    const float* b = ptrs[2 * op + 1];
    const int fin = Q.in_dim[op], fout = Q.out_dim[op];
    float* Wt = flat + Q.w_off[op];
    // 32x32 tiles through shared memory: coalesced reads of W rows and writes of Wt rows
    const int tiles_i = (fin + 31) / 32, tiles_o = (fout + 31) / 32;
    for (int t = blockIdx.x; t < tiles_i * tiles_o; t += gridDim.x) {
      const int ti = t % tiles_i, to = t / tiles_i;
      const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
      for (int r = ty; r < 32; r += 8) {
        const int o = to * 32 + r, i = ti * 32 + tx;
        tile[r][tx] = (o < fout && i < fin) ? W[(int64_t)o * fin + i] : 0.0f;
      }
      __syncthreads();
      for (int r = ty; r < 32; r += 8) {
        const int i = ti * 32 + r, o = to * 32 + tx;
        if (i < fin && o < fout) Wt[(int64_t)i * fout + o] = tile[tx][r];
      }
      __syncthreads();
    }
    if (blockIdx.x == 0)
      for (int o = threadIdx.x; o < fout; o += blockDim.x) flat[Q.b_off[op] + o] = b[o];
  }
}
}  // namespace dexnerf

extern "C" DEXNERF_API int dexnerf_pack_params(const dexnerf_mlp_program* prog, const void* ptr_table, float* flat,
                                               void* stream) {
  DN_REQUIRE(prog && ptr_table && flat, "pack_params: null pointer");
  DN_REQUIRE(prog->n_ops >= 1 && prog->n_ops <= DEXNERF_MAX_OPS, "pack_params: bad op count %d", prog->n_ops);
  dexnerf::PackParamsTable Q{};
  Q.n_ops = prog->n_ops;
  for (int i = 0; i < prog->n_ops; ++i) {
    Q.in_dim[i] = prog->ops[i].src0_dim + prog->ops[i].src1_dim;
    Q.out_dim[i] = prog->ops[i].out_dim;
    Q.w_off[i] = prog->ops[i].w_off;
    Q.b_off[i] = prog->ops[i].b_off;
  }
  dexnerf::pack_params_kernel<<<dim3(24, prog->n_ops), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float* const*>(ptr_table), Q, flat);
  DN_CHECK_LAUNCH("pack_params");
  return 0;
}
