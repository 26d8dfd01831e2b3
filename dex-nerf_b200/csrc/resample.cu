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

// Branch-free binary searches over a sorted shared-memory row of `len` floats: the number of
// entries that compare true against `key` (kOrEqual: s[i] <= key, else s[i] < key) = the insertion
// index.  One LDS + two compares + one predicated add per step, fully unrolled (kLog2 steps cover
// len < 2^kLog2); the reference's per-query loop with data-dependent exits (torchsearchsorted) and the
// first version of this file (lo/hi/mid loop: ~15 instructions per step) both spent most of their
// instructions on index bookkeeping.
template <bool kOrEqual, int kLog2>
__device__ __forceinline__ int count_below_fixed(const float* s, int len, float key) {
  int pos = 0;
#pragma unroll
  for (int b = kLog2 - 1; b >= 0; --b) {
    const int nxt = pos + (1 << b);
    if (nxt <= len) {
      const float c = s[nxt - 1];
      if (kOrEqual ? (c <= key) : (c < key)) pos = nxt;
    }
  }
  return pos;
}

template <bool kOrEqual>
__device__ __forceinline__ int count_below(const float* s, int len, int log2, float key) {
  switch (log2) {   // warp-uniform
    case 5: return count_below_fixed<kOrEqual, 5>(s, len, key);
    case 6: return count_below_fixed<kOrEqual, 6>(s, len, key);
    case 7: return count_below_fixed<kOrEqual, 7>(s, len, key);
    case 8: return count_below_fixed<kOrEqual, 8>(s, len, key);
    case 9: return count_below_fixed<kOrEqual, 9>(s, len, key);
    default: return count_below_fixed<kOrEqual, 12>(s, len, key);   // len <= 4095 (shared-memory limit is lower)
  }
}

// smallest L >= 5 with len < 2^L
__host__ __device__ inline int search_log2(int len) {
  int l = 5;
  while ((1 << l) <= len) ++l;
  return l;
}

// searchsorted(cdf, u, side="right"): first index with cdf[idx] > u, B if none.
__device__ __forceinline__ int upper_bound(const float* s_cdf, int B, int log2B, float u) {
  return count_below<true>(s_cdf, B, log2B, u);
}

__device__ __forceinline__ float invert_cdf(const float* s_cdf, const float* s_bins, int B, int log2B,
                                            float u, int* ind_out) {
  const int ind = upper_bound(s_cdf, B, log2B, u);
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
  const int log2B = search_log2(B);
  for (int64_t ray = (int64_t)blockIdx.x * wpc + warp; ray < n; ray += (int64_t)gridDim.x * wpc) {
    const float* w_row = weights + ray * (B - 1);
    for (int j = lane; j < B; j += 32) s_bins[j] = bins[ray * B + j];
    build_cdf([&](int j) { return w_row[j]; }, B, s_cdf, lane);
    for (int s = lane; s < Nf; s += 32) {
      const float uu = u ? u[ray * Nf + s] : linspace_at(0.0f, 1.0f, Nf, s);
      int ind;
      samples[ray * Nf + s] = invert_cdf(s_cdf, s_bins, B, log2B, uu, &ind);
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
  const int log2B = search_log2(B), log2Nf = search_log2(Nf);
  for (int64_t ray = (int64_t)blockIdx.x * wpc + warp; ray < n; ray += (int64_t)gridDim.x * wpc) {
    const float* z_row = zc + ray * Nc;
    const float* w_row = weights + ray * Nc + 1;  // weights[..., 1:-1]
    for (int j = lane; j < Nc; j += 32) s_sort[j] = z_row[j];
    for (int j = St + lane; j < P; j += 32) s_sort[j] = __int_as_float(0x7f800000);  // +inf pad
    __syncwarp();
    for (int j = lane; j < B; j += 32) s_bins[j] = __fmul_rn(0.5f, __fadd_rn(s_sort[j + 1], s_sort[j]));
    build_cdf([&](int j) { return w_row[j]; }, B, s_cdf, lane);
    // cat + sort (train_utils.py:173).  The coarse depths are sorted; with sorted u (validation:
    // u = linspace) the new samples are sorted too, and the sort is a MERGE: every element's final
    // position is its own index plus its rank in the other list.  Ties: a coarse depth goes before
    // equal samples, so the positions are a permutation.  A sample's rank among the coarse depths
    // needs no search: it was drawn from the bin [mid(z[ind-1], z[ind]), mid(z[ind], z[ind+1])], so
    // the number of coarse depths <= it is ind or ind + 1; the walk below starts at ind and moves
    // while the neighbours say so (valid for any start, 2-3 compares in practice).  The merged row is
    // written speculatively; unsorted samples (random u) discard it and take the network.
    bool sorted = true;
    for (int s = lane; s < Nf; s += 32) {
      const float uu = u ? u[ray * Nf + s] : linspace_at(0.0f, 1.0f, Nf, s);
      int ind;
      const float v = invert_cdf(s_cdf, s_bins, B, log2B, uu, &ind);
      s_sort[Nc + s] = v;
      int r = ind;                    // ind <= B = Nc - 1
      while (r < Nc && s_sort[r] <= v) ++r;
      while (r > 0 && s_sort[r - 1] > v) --r;
      s_merge[s + r] = v;
    }
    __syncwarp();
    for (int s = lane; s < Nf - 1; s += 32) sorted = sorted && (s_sort[Nc + s] <= s_sort[Nc + s + 1]);
    for (int j = lane; j < Nc - 1; j += 32) sorted = sorted && (s_sort[j] <= s_sort[j + 1]);
    if (__all_sync(0xffffffffu, sorted)) {
      const float* Sm = s_sort + Nc;   // Nf samples
      for (int i = lane; i < Nc; i += 32) {
        const float v = s_sort[i];
        s_merge[i + count_below<false>(Sm, Nf, log2Nf, v)] = v;   // samples strictly below v
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

// ---- the same step with the sample counts as template parameters (powers of two: the shipped
// configurations are 64+64, 64+128 and 128+256).  Everything a lane owns stays in registers, all
// loops unroll, and the searches need no bounds: the cdf has 2^L - 1 entries, so an L-step
// branch-free search covers it exactly.  ~4x fewer instructions per ray than the generic kernel.
template <int kLog2>
__device__ __forceinline__ int count_le_full(const float* s, float key) {   // over 2^kLog2 - 1 entries
  const float* p = s;
#pragma unroll
  for (int b = kLog2 - 1; b >= 0; --b)
    if (p[(1 << b) - 1] <= key) p += 1 << b;
  return (int)(p - s);
}

// x / y, correctly rounded, for x >= 0: a zero numerator would send the whole warp through the
// division's slow path (~20 instructions); the quotient is 0 anyway.
__device__ __forceinline__ float fdiv_rn_nonneg(float x, float y) {
  const float q = __fdiv_rn(x == 0.0f ? 1.0f : x, y);
  return x == 0.0f ? 0.0f : q;
}

template <int N> struct Log2Of { static constexpr int value = 1 + Log2Of<N / 2>::value; };
template <> struct Log2Of<1> { static constexpr int value = 0; };

template <int NC, int NF>
__global__ void __launch_bounds__(256) resample_merge_fixed_kernel(const float* __restrict__ zc,
                                                                   const float* __restrict__ weights,
                                                                   int64_t n, const float* __restrict__ u,
                                                                   float* __restrict__ z_fine) {
  constexpr int KC = NC / 32, KF = NF / 32, B = NC - 1, ST = NC + NF;
  constexpr int LC = Log2Of<NC>::value;
  constexpr int P = 1 << (Log2Of<ST - 1>::value + 1);        // power of two >= ST (bitonic fallback)
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
  float* s_sort = smem + (size_t)warp * (P + ST + 2 * NC);   // [coarse NC | samples NF | pad]
  float* s_merge = s_sort + P;                               // ST; scratch for the pdf transpose first
  float* s_cdf = s_merge + ST;                               // NC (B used)
  float* s_bins = s_cdf + NC;                                // NC (B used)
  float* s_samp = s_sort + NC;
  // torch.linspace(0, 1, NF) (common.cuh linspace_at) with the step hoisted out of the ray loop
  const float lin_step = __fdiv_rn(1.0f, (float)(NF - 1));
  for (int64_t ray = (int64_t)blockIdx.x * wpc + warp; ray < n; ray += (int64_t)gridDim.x * wpc) {
    const float* z_row = zc + ray * NC;
    const float* w_row = weights + ray * NC + 1;             // weights[..., 1:-1]
    float zk[KC], wk[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int j = lane + 32 * k;
      zk[k] = z_row[j];
      wk[k] = (j < NC - 2) ? __fadd_rn(w_row[j], 1e-5f) : 0.0f;
      s_sort[j] = zk[k];
    }
    // mid-points (train_utils.py:163), coarse sortedness, normaliser
    bool sorted = true;
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      float nxt = __shfl_down_sync(0xffffffffu, zk[k], 1);
      if (k + 1 < KC) {
        const float first_next = __shfl_sync(0xffffffffu, zk[k + 1 < KC ? k + 1 : k], 0);
        if (lane == 31) nxt = first_next;
      }
      const int j = lane + 32 * k;
      if (j < B) {
        s_bins[j] = __fmul_rn(0.5f, __fadd_rn(nxt, zk[k]));
        sorted = sorted && (zk[k] <= nxt);
      }
      part += (double)wk[k];
    }
    const float total = (float)warp_sum_f64(part);
    // pdf -> cdf.  The fp64 partial sums are exact (see the header), so their association is free:
    // the pdf is transposed through shared memory to KC consecutive entries per lane and ONE warp
    // scan of the lane totals gives every prefix.
#pragma unroll
    for (int k = 0; k < KC; ++k) s_merge[lane + 32 * k] = fdiv_rn_nonneg(wk[k], total);
    __syncwarp();
    double pre[KC];
    {
      double run = 0.0;
#pragma unroll
      for (int c = 0; c < KC; ++c) { run += (double)s_merge[lane * KC + c]; pre[c] = run; }
      const double incl = warp_inclusive_sum_f64(run, lane);
      double excl = __shfl_up_sync(0xffffffffu, incl, 1);
      if (lane == 0) { excl = 0.0; s_cdf[0] = 0.0f; }
#pragma unroll
      for (int c = 0; c < KC; ++c) {
        const int j = lane * KC + c;                          // weight index; cdf[j + 1]
        if (j + 1 <= B - 1) s_cdf[j + 1] = (float)(excl + pre[c]);
      }
    }
    __syncwarp();
    // Inverse-cdf samples, each with its rank among the coarse depths.  A sample drawn with index ind
    // lies in [mid(z[ind-1], z[ind]), mid(z[ind], z[ind+1])] (t is in [0, 1] and every step rounds
    // monotonically), so at least `ind` sorted coarse depths are <= it and normally ind or ind + 1; the
    // loop only runs on when rounding pushed the sample an ulp past a run of tied depths.
    float vk[KF];
    int rk[KF];
#pragma unroll
    for (int k = 0; k < KF; ++k) {
      const int s = lane + 32 * k;
      const float uu = u ? u[ray * NF + s]
                         : (s < NF / 2 ? __fmaf_rn(lin_step, (float)s, 0.0f)
                                       : __fmaf_rn(-lin_step, (float)(NF - 1 - s), 1.0f));
      const int ind = count_le_full<LC>(s_cdf, uu);           // searchsorted(cdf, u, side="right"), in [0, B]
      const int below = ind - 1 < 0 ? 0 : ind - 1;
      const int above = ind > B - 1 ? B - 1 : ind;
      const float cb = s_cdf[below], ca = s_cdf[above];
      const float bb = s_bins[below], ba = s_bins[above];
      float denom = __fsub_rn(ca, cb);
      if (denom < 1e-5f) denom = 1.0f;
      const float num = __fsub_rn(uu, cb);
      const float t = (num >= 0.0f) ? fdiv_rn_nonneg(num, denom) : __fdiv_rn(num, denom);
      const float v = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      vk[k] = v;
      s_samp[s] = v;
      int r = ind;                                            // <= B = NC - 1: s_sort[r] is a coarse depth
      r += (s_sort[r] <= v) ? 1 : 0;
      while (r < NC && s_sort[r] <= v) ++r;
      rk[k] = r;
      s_merge[s + r] = v;                                     // speculative: used if everything is sorted
    }
#pragma unroll
    for (int k = 0; k < KF; ++k) {
      float nxt = __shfl_down_sync(0xffffffffu, vk[k], 1);
      if (k + 1 < KF) {
        const float first_next = __shfl_sync(0xffffffffu, vk[k + 1 < KF ? k + 1 : k], 0);
        if (lane == 31) nxt = first_next;
      }
      if (lane + 32 * k < NF - 1) sorted = sorted && (vk[k] <= nxt);
    }
    const bool all_sorted = __all_sync(0xffffffffu, sorted);
    __syncwarp();
    if (!all_sorted) {
      // random u (training) or unsorted input: full bitonic network over cat(coarse, samples)
      for (int j = ST + lane; j < P; j += 32) s_sort[j] = __int_as_float(0x7f800000);
      __syncwarp();
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
#pragma unroll
      for (int k = 0; k < KC + KF; ++k) z_fine[ray * ST + lane + 32 * k] = s_sort[lane + 32 * k];
      __syncwarp();
      continue;
    }
    // Coarse depths: position = index + number of samples strictly below (ties: coarse first).  No
    // search either: the sample ranks r_j are non-decreasing, so the coarse depths i in
    // [r_{j-1}, r_j) have exactly j samples below them.  Sample j places them (0.5 per sample on
    // average; gaps longer than two - empty stretches of a peaked pdf - are done by the whole warp);
    // the depths before r_0 and from r_{NF-1} on are placed by index.
    const int r_first = __shfl_sync(0xffffffffu, rk[0], 0);
    const int r_last = __shfl_sync(0xffffffffu, rk[KF - 1], 31);
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int i = lane + 32 * k;
      if (i < r_first) s_merge[i] = zk[k];
      if (i >= r_last) s_merge[i + NF] = zk[k];
    }
#pragma unroll
    for (int k = 0; k < KF; ++k) {
      const int j = lane + 32 * k;
      int r_prev = __shfl_up_sync(0xffffffffu, rk[k], 1);
      if (k > 0) {
        const int last_prev = __shfl_sync(0xffffffffu, rk[k > 0 ? k - 1 : 0], 31);
        if (lane == 0) r_prev = last_prev;
      } else if (lane == 0) {
        r_prev = rk[0];                                       // j = 0: the head is placed above
      }
      const int gap = rk[k] - r_prev;
      if (gap >= 1) s_merge[r_prev + j] = s_sort[r_prev];
      if (gap >= 2) s_merge[r_prev + 1 + j] = s_sort[r_prev + 1];
      unsigned big = __ballot_sync(0xffffffffu, gap > 2);
      while (big) {                                           // warp-uniform
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const int lo = __shfl_sync(0xffffffffu, r_prev, src) + 2;
        const int hi = __shfl_sync(0xffffffffu, rk[k], src);
        const int jj = src + 32 * k;
        for (int i = lo + lane; i < hi; i += 32) s_merge[i + jj] = s_sort[i];
      }
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < KC + KF; ++k) z_fine[ray * ST + lane + 32 * k] = s_merge[lane + 32 * k];
    __syncwarp();
  }
}

template <int NC, int NF>
static int launch_resample_fixed(const float* zc, const float* weights, int64_t n, const float* u, float* z_fine,
                                 cudaStream_t stream) {
  constexpr int ST = NC + NF;
  constexpr int P = 1 << (Log2Of<ST - 1>::value + 1);
  constexpr int wpc = 8;
  const size_t smem = sizeof(float) * wpc * (size_t)(P + ST + 2 * NC);
  static_assert(sizeof(float) * wpc * (size_t)(P + ST + 2 * NC) <= 48 * 1024, "fits the default shared-memory limit");
  int64_t blocks = ceil_div64(n, wpc);
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (blocks > cap) blocks = cap;
  resample_merge_fixed_kernel<NC, NF><<<(int)blocks, wpc * 32, smem, stream>>>(zc, weights, n, u, z_fine);
  return 0;
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
  {
    bool fixed = true;
    cudaStream_t st = (cudaStream_t)stream;
    if (Nc == 64 && Nf == 128) launch_resample_fixed<64, 128>(z_coarse, weights, n, u, z_fine, st);
    else if (Nc == 64 && Nf == 64) launch_resample_fixed<64, 64>(z_coarse, weights, n, u, z_fine, st);
    else if (Nc == 128 && Nf == 256) launch_resample_fixed<128, 256>(z_coarse, weights, n, u, z_fine, st);
    else if (Nc == 128 && Nf == 128) launch_resample_fixed<128, 128>(z_coarse, weights, n, u, z_fine, st);
    else fixed = false;
    if (fixed) {
      DN_CHECK_LAUNCH("resample_merge");
      return 0;
    }
  }
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
