// Micro-benchmark: execution rate of tcgen05.mma (kind::f16, M=128, cta_group::1) for the operand
// sources / shapes / shared-memory layouts the MLP kernel can choose between.  Timing only: the
// operand contents are arbitrary.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_microbench tc_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_noswz(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint64_t desc_swz128(uint32_t addr) {   // K-major, 128B swizzle: SBO = 1024
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}

// variant: 0 SS noswz, 1 TS noswz-B, 2 SS swz128, 3 TS swz128-B
__global__ void __launch_bounds__(128, 1) bench(int variant, int N, int reps, int ksteps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x < 32) {
    const uint32_t a_smem = smem_u32(smem);                 // A tile: 128 rows x (ksteps*16) K
    const uint32_t b_smem = smem_u32(smem) + 64 * 1024;     // B tile: N rows x (ksteps*16) K
    const uint32_t id = idesc(N);
    uint32_t parity = 0;
    long long best = 1ll << 60;
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      const uint64_t hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      const uint32_t a_lo = ((a_smem >> 4) & 0x3FFF) | (128u << 16);
      const uint32_t b_lo = ((b_smem >> 4) & 0x3FFF) | ((uint32_t)N << 16);
      if (elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll 1
          for (int c = 0; c < ksteps / 4; ++c) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const int ks = c * 4 + k4;
              const uint32_t acc = (r | ks) ? 1u : 0u;
              if (variant == 0) mma_ss(tm + 256, hi | (uint64_t)(a_lo + ks * 256), hi | (uint64_t)(b_lo + ks * 2 * N), id, acc);
              else if (variant == 1) mma_ts(tm + 256, tm + ks * 8, hi | (uint64_t)(b_lo + ks * 2 * N), id, acc);
              else if (variant == 2) mma_ss(tm + 256, desc_swz128(a_smem + (ks / 4) * 16384 + (ks % 4) * 32), desc_swz128(b_smem + (ks / 4) * N * 128 + (ks % 4) * 32), id, acc);
              else mma_ts(tm + 256, tm + ks * 8, desc_swz128(b_smem + (ks / 4) * N * 128 + (ks % 4) * 32), id, acc);
            }
          }
        }
      }
      __syncwarp();
      if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      while (!try_wait(smem_u32(&bar), parity)) {}
      parity ^= 1;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"SS  B no-swizzle", "TS  B no-swizzle", "SS  B 128B-swizzle", "TS  B 128B-swizzle"};
  for (int grid : {1, 148}) {
    for (int N : {64, 128, 256}) {
      for (int v = 0; v < 4; ++v) {
        const int reps = 8, ksteps = 16;
        bench<<<grid, 128, 200 * 1024>>>(v, N, reps, ksteps, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d N %d: %s\n", v, N, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double per = (double)mx / (reps * ksteps);
        printf("grid %3d  N %3d  %-20s: %7lld cycles for %d MMAs = %6.1f cyc/MMA  (floor %d)  %.0f%% of peak\n", grid, N,
               names[v], mx, reps * ksteps, per, N / 2, 100.0 * (N / 2) / per);
      }
    }
  }
  return 0;
}
