// Hierarchical resampling: sample_pdf (inverse-CDF sampling) and the fused
// mid-points -> sample_pdf -> cat -> sort step of the fine pass.
//
// Reference: nerf/nerf_helpers.py:262-304 (sample_pdf_2, which is what `nerf.sample_pdf` resolves
// to), the external torchsearchsorted.searchsorted(side="right") it calls (:290), and
// nerf/train_utils.py:163-173.  Arithmetic contract: oracle/nerf_oracle.py pdf_to_cdf / sample_pdf /
// merge_fine - the normaliser and the cdf are accumulated in fp64 and rounded once per element.
// For weights in [0, 1] (+1e-5) every fp64 partial sum is exact, so the warp-parallel scan below
// gives the same bits as a sequential one and the searchsorted indices are bit-exact.
//
// Layout: one warp per ray.  The ray's cdf and bins live in shared memory; each lane inverts the
// cdf for samples lane, lane+32, ... by binary search (the warp-cooperative replacement for the
// reference's one-thread-per-query extension), and the Nc+Nf depths are sorted by a warp-level
// bitonic network in shared memory.  HBM-bound: 8*S + 4*Nf (+4*Nf when u is supplied) bytes/ray.
#include "common.cuh"

namespace dexnerf {

__device__ __forceinline__ double warp_inclusive_sum_f64(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// Build cdf[0..B-1] (cdf[0] = 0) in shared memory from B-1 weights read through `wsrc(j)`.
template <typename WeightFn>
__device__ __forceinline__ void build_cdf(WeightFn wsrc, int B, float* s_cdf, int lane) {
  const int nw = B - 1;
  double part = 0.0;
  for (int j = lane; j < nw; j += 32) part += (double)__fadd_rn(wsrc(j), 1e-5f);
  const float total = (float)warp_sum_f64(part);
  double carry = 0.0;
  if (lane == 0) s_cdf[0] = 0.0f;
  for (int c0 = 0; c0 < nw; c0 += 32) {
    const int j = c0 + lane;
    const double pdf = j < nw ? (double)__fdiv_rn(__fadd_rn(wsrc(j), 1e-5f), total) : 0.0;
    const double incl = warp_inclusive_sum_f64(pdf, lane);
    if (j < nw) s_cdf[j + 1] = (float)(carry + incl);
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
}

// searchsorted(cdf, u, side="right"): first index with cdf[idx] > u, B if none.
__device__ __forceinline__ int upper_bound(const float* s_cdf, int B, float u) {
  int lo = 0, hi = B;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (s_cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float invert_cdf(const float* s_cdf, const float* s_bins, int B, float u,
                                            int* ind_out) {
  const int ind = upper_bound(s_cdf, B, u);
  const int below = ind - 1 < 0 ? 0 : ind - 1;
  const int above = ind > B - 1 ? B - 1 : ind;
  const float cb = s_cdf[below], ca = s_cdf[above];
  const float bb = s_bins[below], ba = s_bins[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < 1e-5f) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  *ind_out = ind;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

// ---- stand-alone sample_pdf: bins (n,B), weights (n,B-1) -> samples (n,Nf), inds (n,Nf)
__global__ void __launch_bounds__(256) sample_pdf_kernel(const float* __restrict__ bins,
                                                         const float* __restrict__ weights,
                                                         int64_t n, int B, int Nf,
                                                         const float* __restrict__ u,
                                                         float* __restrict__ samples,
                                                         int64_t* __restrict__ inds) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* s_cdf = smem + (size_t)warp * 2 * B;
  float* s_bins = s_cdf + B;
  for (int64_t ray = (int64_t)blockIdx.x * wpc + warp; ray < n; ray += (int64_t)gridDim.x * wpc) {
    const float* w_row = weights + ray * (B - 1);
    for (int j = lane; j < B; j += 32) s_bins[j] = bins[ray * B + j];
    build_cdf([&](int j) { return w_row[j]; }, B, s_cdf, lane);
    for (int s = lane; s < Nf; s += 32) {
      const float uu = u ? u[ray * Nf + s] : linspace_at(0.0f, 1.0f, Nf, s);
      int ind;
      samples[ray * Nf + s] = invert_cdf(s_cdf, s_bins, B, uu, &ind);
      if (inds) inds[ray * Nf + s] = ind;
    }
    __syncwarp();
  }
}

// ---- fused fine-sample construction: z_coarse (n,Nc), weights (n,Nc) -> sorted z_fine (n,Nc+Nf)
__global__ void __launch_bounds__(256) resample_merge_kernel(const float* __restrict__ zc,
                                                             const float* __restrict__ weights,
                                                             int64_t n, int Nc, int Nf, int P,
                                                             const float* __restrict__ u,
                                                             float* __restrict__ z_fine) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  const int B = Nc - 1;
  const int St = Nc + Nf;
  float* s_cdf = smem + (size_t)warp * (2 * B + P + St);
  float* s_bins = s_cdf + B;
  float* s_sort = s_bins + B;
  float* s_merge = s_sort + P;     // destination of the merge
  for (int64_t ray = (int64_t)blockIdx.x * wpc + warp; ray < n; ray += (int64_t)gridDim.x * wpc) {
    const float* z_row = zc + ray * Nc;
    const float* w_row = weights + ray * Nc + 1;  // weights[..., 1:-1]
    for (int j = lane; j < Nc; j += 32) s_sort[j] = z_row[j];
    for (int j = St + lane; j < P; j += 32) s_sort[j] = __int_as_float(0x7f800000);  // +inf pad
    __syncwarp();
    for (int j = lane; j < B; j += 32) s_bins[j] = __fmul_rn(0.5f, __fadd_rn(s_sort[j + 1], s_sort[j]));
    build_cdf([&](int j) { return w_row[j]; }, B, s_cdf, lane);
    for (int s = lane; s < Nf; s += 32) {
      const float uu = u ? u[ray * Nf + s] : linspace_at(0.0f, 1.0f, Nf, s);
      int ind;
      s_sort[Nc + s] = invert_cdf(s_cdf, s_bins, B, uu, &ind);
    }
    __syncwarp();
    // cat + sort (train_utils.py:173).  The coarse depths are sorted; with sorted u (validation:
    // u = linspace) the new samples are sorted too, and the sort is a MERGE: every element's final
    // position is its own index plus its rank in the other list (two binary searches per lane and
    // element instead of a 36-stage bitonic network).  Ties: a coarse depth goes before equal
    // samples, so the positions are a permutation.  Unsorted samples (random u) take the network.
    bool sorted = true;
    for (int s = lane; s < Nf - 1; s += 32) sorted = sorted && (s_sort[Nc + s] <= s_sort[Nc + s + 1]);
    for (int j = lane; j < Nc - 1; j += 32) sorted = sorted && (s_sort[j] <= s_sort[j + 1]);
    if (__all_sync(0xffffffffu, sorted)) {
      const float* A = s_sort;         // Nc coarse depths
      const float* Sm = s_sort + Nc;   // Nf samples
      for (int i = lane; i < Nc; i += 32) {
        const float v = A[i];
        int lo = 0, hi = Nf;           // samples strictly below v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (Sm[mid] < v) lo = mid + 1; else hi = mid; }
        s_merge[i + lo] = v;
      }
      for (int i = lane; i < Nf; i += 32) {
        const float v = Sm[i];
        int lo = 0, hi = Nc;           // coarse depths at or below v
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (A[mid] <= v) lo = mid + 1; else hi = mid; }
        s_merge[i + lo] = v;
      }
      __syncwarp();
      for (int j = lane; j < St; j += 32) z_fine[ray * St + j] = s_merge[j];
      __syncwarp();
      continue;
    }
    // bitonic sort of P (power of two) values by one warp
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          const int p = i ^ j;
          if (p > i) {
            const float a = s_sort[i], b = s_sort[p];
            const bool asc = (i & k) == 0;
            if ((a > b) == asc) { s_sort[i] = b; s_sort[p] = a; }
          }
        }
        __syncwarp();
      }
    }
    for (int j = lane; j < St; j += 32) z_fine[ray * St + j] = s_sort[j];
    __syncwarp();
  }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int B, int Nf,
                                  const float* u, float* samples, int64_t* inds, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(bins && weights && samples, "sample_pdf: null pointer");
  DN_REQUIRE(B >= 2 && Nf >= 1, "sample_pdf: need at least 2 bins and 1 sample");
  const size_t per_warp = sizeof(float) * 2 * (size_t)B;
  int wpc = 8;
  while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
  DN_REQUIRE(per_warp * wpc <= 200 * 1024, "sample_pdf: %d bins do not fit in shared memory", B);
  const size_t smem = per_warp * wpc;
  if (smem > 48 * 1024)
    DN_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div64(n, wpc);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  sample_pdf_kernel<<<(int)blocks, wpc * 32, smem, (cudaStream_t)stream>>>(bins, weights, n, B, Nf, u,
                                                                          samples, inds);
  DN_CHECK_LAUNCH("sample_pdf");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_resample_merge(const float* z_coarse, const float* weights, int64_t n, int Nc,
                                      int Nf, const float* u, float* z_fine, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(z_coarse && weights && z_fine, "resample_merge: null pointer");
  DN_REQUIRE(Nc >= 3 && Nf >= 1, "resample_merge: need Nc >= 3 and Nf >= 1");
  const int P = next_pow2(Nc + Nf);
  const size_t per_warp = sizeof(float) * (2 * (size_t)(Nc - 1) + P + (size_t)(Nc + Nf));
  int wpc = 8;
  while (wpc > 1 && per_warp * wpc > 96 * 1024) wpc >>= 1;
  DN_REQUIRE(per_warp * wpc <= 200 * 1024, "resample_merge: %d+%d samples do not fit in shared memory", Nc, Nf);
  const size_t smem = per_warp * wpc;
  if (smem > 48 * 1024)
    DN_CUDA(cudaFuncSetAttribute(resample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = ceil_div64(n, wpc);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  resample_merge_kernel<<<(int)blocks, wpc * 32, smem, (cudaStream_t)stream>>>(z_coarse, weights, n, Nc,
                                                                              Nf, P, u, z_fine);
  DN_CHECK_LAUNCH("resample_merge");
  return 0;
}
