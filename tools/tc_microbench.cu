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
#define LD16(taddr, r)                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                        \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"               \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
               : "r"(taddr) : "memory")
__device__ volatile int g_stop;
__global__ void __launch_bounds__(512, 1) bench(int variant, int N, int reps, int ksteps, long long* out, int n_ld, long long* ld_out, unsigned* sink) {
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
  __shared__ volatile int s_stop;
  if (threadIdx.x == 0) s_stop = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp >= 4 && warp < 4 + n_ld) {
    // concurrent accumulator readers (columns 0..127 of this warp's lane quarter)
    const uint32_t base = tm + (((uint32_t)(warp & 3) * 32) << 16);
    uint32_t r[16]; unsigned acc = 0; long long n = 0;
    const long long t0 = clock64();
    while (!s_stop) {
      for (int c = 0; c < 8; ++c) {
        LD16(base + c * 16, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= r[j];
      }
      ++n;
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { ld_out[(warp - 4) * 2] = t1 - t0; ld_out[(warp - 4) * 2 + 1] = n; }
    sink[threadIdx.x] = acc;
  }
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
              else if (variant == 4) {
                // both operands MN-major, no swizzle, as the weight-gradient GEMM reads the tape images:
                // LBO = 128 B (K-adjacent core matrices), SBO = 1024 B (MN-adjacent), a K step = 256 B
                const uint64_t hi_mn = ((uint64_t)(1024 >> 4) << 32) | (1ull << 46);
                const uint32_t id_mn = id | (1u << 15) | (1u << 16);
                const uint32_t a_mn = ((a_smem >> 4) & 0x3FFF) | ((128u >> 4) << 16);
                const uint32_t b_mn = ((b_smem >> 4) & 0x3FFF) | ((128u >> 4) << 16);
                mma_ss(tm + 256, hi_mn | (uint64_t)(a_mn + (ks % 4) * 16), hi_mn | (uint64_t)(b_mn + (ks % 4) * 16), id_mn, acc);
              }
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
    if (threadIdx.x == 0) { out[blockIdx.x] = best; s_stop = 1; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long *out, *ld_out; unsigned* sink;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaMalloc(&ld_out, 64 * sizeof(long long));
  cudaMalloc(&sink, 512 * 4);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[5] = {"SS  B no-swizzle", "TS  B no-swizzle", "SS  B 128B-swizzle", "TS  B 128B-swizzle",
                          "SS  A,B MN-major (dW)"};
  for (int n_ld : {0, 4, 8}) {
    for (int N : {128, 256}) {
      for (int v : {0, 1, 4}) {
        const int reps = 32, ksteps = 16;
        cudaMemset(ld_out, 0, 64 * sizeof(long long));
        bench<<<1, 512, 200 * 1024>>>(v, N, reps, ksteps, out, n_ld, ld_out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d N %d: %s\n", v, N, cudaGetErrorString(e)); return 1; }
        long long h[1], l[32];
        cudaMemcpy(h, out, sizeof(long long), cudaMemcpyDeviceToHost);
        cudaMemcpy(l, ld_out, 32 * sizeof(long long), cudaMemcpyDeviceToHost);
        const double per = (double)h[0] / (reps * ksteps);
        printf("readers %d  N %3d  %-18s: %6.1f cyc/MMA (%.0f%% of peak)", n_ld, N, names[v], per, 100.0 * (N / 2) / per);
        if (n_ld) printf(" | reader warp 0: %.0f cycles per 128-col pass (%lld passes)", (double)l[0] / (l[1] ? l[1] : 1), l[1]);
        printf("\n");
      }
    }
  }
  return 0;
}
