// Per-SM streaming bandwidth from HBM into shared memory: cp.async.bulk (TMA bulk copy) ring vs LDG.128 with many
// loads in flight, for 1 ... 148 CTAs (one per SM).  Question it answers: can HALF the SMs
// pull the forward tape at the full HBM rate, i.e. what is the per-SM ceiling of each load path?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/sm_stream_bench tools/sm_stream_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// each CTA streams its own contiguous `bytes_per_cta` region in `chunk`-byte bulk copies through `stages` slots
__global__ void __launch_bounds__(128, 1) tma_stream(const uint8_t* src, size_t bytes_per_cta, int chunk, int stages) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  const uint32_t sb = smem_u32(smem), bb = smem_u32(bars);
  if (threadIdx.x == 0) { for (int s = 0; s < stages; ++s) mbar_init(bb + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* p = src + (size_t)blockIdx.x * bytes_per_cta;
    const size_t n = bytes_per_cta / chunk;
    for (size_t i = 0; i < n + stages; ++i) {
      const int st = (int)(i % stages);
      if (i >= (size_t)stages) mbar_wait(bb + 8 * st, (uint32_t)(((i / stages) - 1) & 1));   // previous copy into this slot landed
      if (i < n) { mbar_expect(bb + 8 * st, chunk); bulk_g2s(sb + st * chunk, p + i * chunk, chunk, bb + 8 * st); }
    }
  }
}


__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// The weight-gradient GEMM's ring without the GEMM: a producer warp, `n_cons` consumer warps that wait for a stage and
// hand the slot back (full -> consumer -> empty -> producer: two polling hops per slot instead of the one above), two
// source streams per stage (A and G images, `chunk` bytes each).
__global__ void __launch_bounds__(256, 1) tma_ring(const uint8_t* src, size_t bytes_per_cta, int chunk, int stages, int n_cons,
                                                   int two_streams) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[32];
  const uint32_t sb = smem_u32(smem), bb = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bb + 8 * s, 1); mbar_init(bb + 128 + 8 * s, n_cons * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t per_stage = two_streams ? 2 * (size_t)chunk : (size_t)chunk;
  const size_t n = bytes_per_cta / per_stage;
  const uint8_t* p = src + (size_t)blockIdx.x * bytes_per_cta;
  const uint8_t* q = p + bytes_per_cta / 2;
  if (warp == 0) {
    for (size_t i = 0; i < n; ++i) {
      const int st = (int)(i % stages);
      const uint32_t ph = (uint32_t)((i / stages) & 1);
      mbar_wait(bb + 128 + 8 * st, ph ^ 1);
      if (lane == 0) {
        mbar_expect(bb + 8 * st, (uint32_t)per_stage);
        if (two_streams) {
          bulk_g2s(sb + st * 2 * chunk, p + i * chunk, chunk, bb + 8 * st);
          bulk_g2s(sb + st * 2 * chunk + chunk, q + i * chunk, chunk, bb + 8 * st);
        } else {
          bulk_g2s(sb + st * chunk, p + i * chunk, chunk, bb + 8 * st);
        }
      }
      __syncwarp();
    }
  } else if (warp <= n_cons) {
    for (size_t i = 0; i < n; ++i) {
      const int st = (int)(i % stages);
      const uint32_t ph = (uint32_t)((i / stages) & 1);
      mbar_wait(bb + 8 * st, ph);
      mbar_arrive(bb + 128 + 8 * st);
    }
  }
}

// each thread keeps `kDepth` 16-byte loads in flight; results are stored to shared memory (like a real loader)
template <int kDepth>
__global__ void __launch_bounds__(1024, 1) ldg_stream(const uint4* src, size_t vec_per_cta, uint32_t* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint4* s4 = reinterpret_cast<uint4*>(smem);
  const uint4* p = src + (size_t)blockIdx.x * vec_per_cta;
  uint32_t acc = 0;
  for (size_t base = 0; base + (size_t)kDepth * blockDim.x <= vec_per_cta; base += (size_t)kDepth * blockDim.x) {
    uint4 v[kDepth];
#pragma unroll
    for (int k = 0; k < kDepth; ++k) v[k] = __ldcs(p + base + (size_t)k * blockDim.x + threadIdx.x);
#pragma unroll
    for (int k = 0; k < kDepth; ++k) { s4[(k * blockDim.x + threadIdx.x) & 4095] = v[k]; acc ^= v[k].x; }
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void fill_random(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    p[i] = make_uint4(h, h * 3266489917u, h ^ 0x9E3779B9u, h * 668265263u);
  }
}

int main() {
  const size_t total = (size_t)8 << 30;
  uint8_t* buf; uint32_t* sink;
  cudaMalloc(&buf, total); cudaMalloc(&sink, 4); cudaMemset(buf, 1, total);
  fill_random<<<148 * 8, 256>>>((uint4*)buf, total / 16);   // (constant data would flatter the memory system)
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int ctas[] = {1, 8, 37, 74, 111, 148};
  printf("%-34s", "path");
  for (int c : ctas) printf("  %4d CTAs (GB/s/SM | TB/s)", c);
  printf("\n");
  auto report = [&](const char* name, auto launch) {
    printf("%-34s", name);
    for (int c : ctas) {
      const size_t per = ((size_t)48 << 20);           // 48 MB per CTA (>> L2 share)
      launch(c, per);                                   // warm-up
      cudaEventRecord(e0); launch(c, per); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double gbs = (double)per * c / (ms * 1e-3) / 1e9;
      printf("  %12.1f | %6.2f      ", gbs / c, gbs / 1e3);
    }
    cudaError_t err = cudaGetLastError();
    printf("%s\n", err == cudaSuccess ? "" : cudaGetErrorString(err));
  };
  cudaFuncSetAttribute(tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  report("TMA bulk 32 KB x 3 stages", [&](int c, size_t per) { tma_stream<<<c, 128, 96 * 1024 + 1024>>>(buf, per, 32768, 3); });
  report("TMA bulk 32 KB x 6 stages", [&](int c, size_t per) { tma_stream<<<c, 128, 192 * 1024 + 1024>>>(buf, per, 32768, 6); });
  report("TMA bulk 16 KB x 12 stages", [&](int c, size_t per) { tma_stream<<<c, 128, 192 * 1024 + 1024>>>(buf, per, 16384, 12); });
  report("TMA bulk 4 KB x 48 stages (16 bars)", [&](int c, size_t per) { tma_stream<<<c, 128, 64 * 1024 + 1024>>>(buf, per, 4096, 16); });
  cudaFuncSetAttribute(tma_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  report("ring 64 KB x 3, 1 consumer warp", [&](int c, size_t per) { tma_ring<<<c, 256, 192 * 1024 + 1024>>>(buf, per, 65536, 3, 1, 0); });
  report("ring 64 KB x 3, 5 consumer warps", [&](int c, size_t per) { tma_ring<<<c, 256, 192 * 1024 + 1024>>>(buf, per, 65536, 3, 5, 0); });
  report("ring (32+32) KB x 3, 5 cons. warps", [&](int c, size_t per) { tma_ring<<<c, 256, 192 * 1024 + 1024>>>(buf, per, 32768, 3, 5, 1); });
  report("ring (16+16) KB x 6, 5 cons. warps", [&](int c, size_t per) { tma_ring<<<c, 256, 192 * 1024 + 1024>>>(buf, per, 16384, 6, 5, 1); });
  cudaFuncSetAttribute(ldg_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(ldg_stream<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(ldg_stream<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  report("LDG.128 x4 in flight, 384 thr", [&](int c, size_t per) { ldg_stream<4><<<c, 384, 65536>>>((const uint4*)buf, per / 16, sink); });
  report("LDG.128 x8 in flight, 384 thr", [&](int c, size_t per) { ldg_stream<8><<<c, 384, 65536>>>((const uint4*)buf, per / 16, sink); });
  report("LDG.128 x16 in flight, 384 thr", [&](int c, size_t per) { ldg_stream<16><<<c, 384, 65536>>>((const uint4*)buf, per / 16, sink); });
  report("LDG.128 x8 in flight, 1024 thr", [&](int c, size_t per) { ldg_stream<8><<<c, 1024, 65536>>>((const uint4*)buf, per / 16, sink); });
  return 0;
}
