// The training step's one real exchange - the data-parallel mean of the 2 x 595 844 parameter gradients
// (train_dexnerf_rgb.py:278 loss.backward() on every rank, then optimizer.step()) - fused with its consumer over
// NVLink peer memory: ONE kernel per step reads every rank's gradient buffer directly (coalesced 16-byte P2P loads
// through NVSwitch), sums them in rank order, applies torch.optim.Adam's update to the local parameters and clears
// the OTHER gradient buffer for the next step.  No NCCL launch, no extra pass over the gradients, no stream
// hand-over between the backward and the optimizer; every rank computes the same sum in the same order, so the
// replicas stay bit-identical.
//
// Protocol (one process per GPU; buffers come from dexnerf_p2p_alloc = cudaMalloc, exported / opened with CUDA IPC):
//   * gradient buffers ping-pong: step s accumulates into G[s & 1]; the fused kernel of step s reads the peers' G[s & 1]
//     and zeroes the local G[(s + 1) & 1], which the peers finished reading in step s - 1 - they could not have posted
//     their step-s flag otherwise - so one barrier per step is enough;
//   * barrier: rank r stores the step token into slot r of EVERY rank's flag array (st.release.sys after a system
//     fence: its weight-gradient GEMMs finished earlier on the same stream) and polls its own local array until every
//     slot holds a token >= the step's (ld.acquire.sys); bounded, traps instead of hanging;
//   * peer gradients are read with ld.volatile (never from a stale L1 line of the step before last).
#include "common.cuh"

namespace dexnerf {

constexpr int kMaxPeers = 8;

struct PeerTable {
  const float* grads[kMaxPeers];     // every rank's gradient buffer of THIS step (device pointers valid on this GPU)
  uint32_t* flags[kMaxPeers];        // every rank's flag array [world]
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void adam_update(float gi, float& p, float& m, float& v, float beta1, float beta2, float eps,
                                            float step_size, float bc2_sqrt) {
  // torch.optim.Adam element for element (csrc/optim.cu adam_step_kernel)
  const float mi = __fmaf_rn(1.0f - beta1, __fsub_rn(gi, m), m);
  const float vi = __fmaf_rn(__fmul_rn(1.0f - beta2, gi), gi, __fmul_rn(v, beta2));
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), eps);
  p = __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(mi, denom)));
  m = mi;
  v = vi;
}

__global__ void __launch_bounds__(256) adam_allreduce_kernel(float* __restrict__ p, float* __restrict__ m,
                                                             float* __restrict__ v, float* __restrict__ zero_next,
                                                             int64_t n, const PeerTable peers, int rank, int world,
                                                             uint32_t token, float beta1, float beta2, float eps,
                                                             float step_size, float bc2_sqrt, float grad_scale) {
  // ---- barrier: my gradients are complete (stream order) -> tell everyone; wait until everyone told me
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.flags[threadIdx.x] + rank, token);
  }
  if (threadIdx.x < world) {
    const uint32_t* mine = peers.flags[rank] + threadIdx.x;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - token) < 0) {
      if (++spins > (1u << 26)) {
        printf("dexnerf adam_allreduce: rank %d still waits for rank %d at step token %u\n", rank, (int)threadIdx.x, token);
        __trap();
      }
    }
  }
  __syncthreads();
  // ---- sum over ranks in rank order (identical on every rank), Adam, clear the other buffer
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 g = ld_volatile_f4(peers.grads[0] + 4 * i);
    for (int r = 1; r < world; ++r) {
      const float4 o = ld_volatile_f4(peers.grads[r] + 4 * i);
      g.x = __fadd_rn(g.x, o.x); g.y = __fadd_rn(g.y, o.y); g.z = __fadd_rn(g.z, o.z); g.w = __fadd_rn(g.w, o.w);
    }
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    adam_update(__fmul_rn(g.x, grad_scale), pp.x, mm.x, vv.x, beta1, beta2, eps, step_size, bc2_sqrt);
    adam_update(__fmul_rn(g.y, grad_scale), pp.y, mm.y, vv.y, beta1, beta2, eps, step_size, bc2_sqrt);
    adam_update(__fmul_rn(g.z, grad_scale), pp.z, mm.z, vv.z, beta1, beta2, eps, step_size, bc2_sqrt);
    adam_update(__fmul_rn(g.w, grad_scale), pp.w, mm.w, vv.w, beta1, beta2, eps, step_size, bc2_sqrt);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(zero_next)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_p2p_alloc(int64_t bytes, void** ptr) {
  DN_REQUIRE(ptr && bytes > 0, "p2p_alloc: bad arguments");
  DN_CUDA(cudaMalloc(ptr, (size_t)bytes));
  DN_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  return 0;
}

extern "C" DEXNERF_API int dexnerf_p2p_free(void* ptr) {
  if (ptr) DN_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" DEXNERF_API int dexnerf_p2p_export(void* ptr, void* handle64) {
  DN_REQUIRE(ptr && handle64, "p2p_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DN_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return 0;
}

extern "C" DEXNERF_API int dexnerf_p2p_open(const void* handle64, void** ptr) {
  DN_REQUIRE(handle64 && ptr, "p2p_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" DEXNERF_API int dexnerf_p2p_close(void* ptr) {
  if (ptr) DN_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" DEXNERF_API int dexnerf_adam_step_allreduce(float* params, float* exp_avg, float* exp_avg_sq, float* zero_next,
                                                       int64_t n, const void* const* peer_grads,
                                                       void* const* peer_flags, int rank, int world, uint32_t token,
                                                       float lr, float beta1, float beta2, float eps, int64_t step,
                                                       float grad_scale, void* stream) {
  DN_REQUIRE(params && exp_avg && exp_avg_sq && zero_next && peer_grads && peer_flags, "adam_step_allreduce: null pointer");
  DN_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, "adam_step_allreduce: world must be 1..%d", kMaxPeers);
  DN_REQUIRE(n > 0 && n % 4 == 0, "adam_step_allreduce: the parameter count must be a multiple of 4");
  DN_REQUIRE(step >= 1, "adam_step_allreduce: step counts from 1");
  PeerTable t{};
  for (int r = 0; r < world; ++r) {
    DN_REQUIRE(peer_grads[r] && peer_flags[r], "adam_step_allreduce: null peer pointer");
    t.grads[r] = reinterpret_cast<const float*>(peer_grads[r]);
    t.flags[r] = reinterpret_cast<uint32_t*>(peer_flags[r]);
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
  int64_t blocks = ceil_div64(n / 4, 256);
  if (blocks > kNumSMs) blocks = kNumSMs;        // every CTA is resident: the in-kernel barrier cannot starve
  adam_allreduce_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(params, exp_avg, exp_avg_sq, zero_next, n, t, rank,
                                                                      world, token, beta1, beta2, eps, step_size, bc2_sqrt,
                                                                      grad_scale);
  DN_CHECK_LAUNCH("adam_step_allreduce");
  return 0;
}
