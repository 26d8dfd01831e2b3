// Micro-benchmark: TMEM -> register read bandwidth (tcgen05.ld.32x32b) with 4 / 8 / 16 warps,
// optionally while the tensor pipe runs TS-mode MMAs (which also read TMEM).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                        \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                        \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"        \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),      \
                 "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),    \
                 "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),    \
                 "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
               : "r"(taddr) : "memory")

__global__ void __launch_bounds__(512, 1) bench(int n_warps, int reps, int per_wait, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  long long t0 = 0, t1 = 0;
  uint32_t acc = 0;
  if (warp < n_warps) {
    const uint32_t base = tm + (((uint32_t)(warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128) % 512;
    uint32_t r[32];
    __syncwarp();
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      for (int c = 0; c < 4; ++c) {
        LD32(base + c * 32, r);
        if ((c + 1) % per_wait == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    t1 = clock64();
  }
  if (lane == 0 && warp < n_warps) { out[warp] = t1 - t0; }
  sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* out; uint32_t* sink;
  cudaMalloc(&out, 64 * sizeof(long long));
  cudaMalloc(&sink, 512 * 4);
  for (int per_wait : {1, 2, 4}) {
    for (int nw : {1, 4, 8, 16}) {
      const int reps = 64;
      bench<<<1, 512>>>(nw, reps, per_wait, out, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[16];
      cudaMemcpy(h, out, nw * sizeof(long long), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int i = 0; i < nw; ++i) mx = h[i] > mx ? h[i] : mx;
      const double bytes = (double)nw * reps * 4 * 32 * 32 * 4;
      printf("warps %2d  loads per wait %d: %8lld cycles, %.1f B/cycle/SM, %.0f cycles per x32 load per warp\n", nw, per_wait, mx,
             bytes / mx, (double)mx / (reps * 4));
    }
  }
  return 0;
}
