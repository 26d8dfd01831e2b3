// Alpha compositing with the Dex-NeRF first-crossing depth, and the stand-alone
// cumprod_exclusive.
//
// Reference: nerf/volume_rendering_utils.py:6-70, nerf/nerf_helpers.py:43-64.
// Arithmetic contract: oracle/nerf_oracle.py volume_render_radiance_field / cumprod_exclusive
// (fp32 element ops in the reference's order; transmittance product and the per-ray sums
// accumulated in fp64 and rounded once).
//
// Layout: one warp per ray, lane l owns samples l, l+32, l+64, ... so that every global access of
// the (n,S,4) field, the (n,S) depths and the (n,S) weights is a fully coalesced 128/512-byte
// warp transaction.  The transmittance T_i = prod_{j<i}(1-alpha_j+1e-10) is a warp-level
// exclusive product scan (__shfl_up_sync) with a running carry between 32-sample chunks - this
// replaces the reference's cumprod -> roll -> overwrite sequence.  The first sample with
// sigma > m comes from __ballot_sync + __ffs per threshold.  A CTA of 8 warps covers 32
// consecutive rays; per-ray results are staged in shared memory and written as 128-byte rows.
// HBM-bound: 24*S + 36 + 4*T algorithmic bytes per ray.
#include "common.cuh"

namespace dexnerf {

constexpr int kCompositeWarps = 8;
constexpr int kRaysPerCta = 32;
constexpr int kRaysPerWarp = kRaysPerCta / kCompositeWarps;
constexpr int kMaxThresholds = 64;

__device__ __forceinline__ double shfl_up_f64(double v, int d) {
  return __shfl_up_sync(0xffffffffu, v, d);
}

// Inclusive product scan over the 32 lanes.
__device__ __forceinline__ double warp_inclusive_product(double v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = shfl_up_f64(v, d);
    if (lane >= d) v *= o;
  }
  return v;
}

// sigmoid(x) = 1 / (1 + e^-x) on the MUFU unit: ex2.approx (2 + |1.16 x| ulp) and rcp.approx (1 ulp) - a
// few 1e-7 relative on a colour in [0, 1], far inside the 1e-5 parity bar of the colour sums.  (The
// correctly rounded reciprocal cost 10 instructions per colour; compositing is bound by instruction issue.)
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__fadd_rn(1.0f, __expf(-x))));
  return r;
}

// kSpl = consecutive samples per lane and chunk; kFull = S is a multiple of 32 * kSpl (no bounds
// predicates: true for the shipped 64 / 128 / 192 / 384 samples per ray).
template <int kSpl, bool kFull>
__global__ void __launch_bounds__(kCompositeWarps * 32, kSpl >= 6 ? 4 : 1)
composite_kernel(const float4* __restrict__ rf, const float* __restrict__ z,
                 const float* __restrict__ rd, const float* __restrict__ noise, int64_t n, int S,
                 int white_background, const float* __restrict__ thresholds, int T,
                 float* __restrict__ rgb_out, float* __restrict__ disp_out,
                 float* __restrict__ acc_out, float* __restrict__ weights_out,
                 float* __restrict__ depth_out, float* __restrict__ dex_depth,
                 int64_t* __restrict__ dex_index, int64_t dex_stride) {
  extern __shared__ float smem[];
  // staging: [6 + 2T][32]: rgb0 rgb1 rgb2 disp acc depth | dex depth (T) | dex index (T, as int)
  // Thresholds are ranked ascending once per CTA: a sample that exceeds the k-th smallest threshold
  // exceeds all smaller ones, so "how many thresholds does sigma exceed" (a 6-step binary search)
  // replaces T compares, and the first crossings of a whole run of thresholds come out of ONE
  // ballot (see the level loop below).
  float* s_thr = smem;                          // T ascending, then +inf up to 2 * kMaxThresholds
  float* s_out = smem + 2 * kMaxThresholds;     // (6 + 2T) * 32
  __shared__ int s_orig[kMaxThresholds];        // sorted rank -> caller's threshold index
  __shared__ int s_found[kCompositeWarps][kMaxThresholds];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = T + threadIdx.x; t < 2 * kMaxThresholds; t += blockDim.x) s_thr[t] = __int_as_float(0x7f800000);
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float v = thresholds[t];
    int rank = 0;
    for (int o = 0; o < T; ++o) {
      const float w = thresholds[o];
      rank += (w < v) || (w == v && o < t);
    }
    s_thr[rank] = v;
    s_orig[rank] = t;
  }
  __syncthreads();

  for (int64_t base = (int64_t)blockIdx.x * kRaysPerCta; base < n;
       base += (int64_t)gridDim.x * kRaysPerCta) {
    // |rd| of the warp's rays, one per lane (the fp64 square root is ~40 instructions: once, not per ray)
    float norm_l = 0.0f;
    {
      const int64_t ray = base + warp * kRaysPerWarp + lane;
      if (lane < kRaysPerWarp && ray < n) {
        const float dx = rd[ray * 3], dy = rd[ray * 3 + 1], dz = rd[ray * 3 + 2];
        norm_l = (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy + (double)dz * (double)dz);
      }
    }
    for (int rr = 0; rr < kRaysPerWarp; ++rr) {
      const int slot = warp * kRaysPerWarp + rr;
      const int64_t ray = base + slot;
      if (ray >= n) break;  // warp-uniform
      const float norm = __shfl_sync(0xffffffffu, norm_l, rr);
      const float4* rf_row = rf + ray * S;
      const float* z_row = z + ray * S;
      double carry = 1.0;
      // per-lane partial sums in fp32 (6-12 terms each), then a warp tree: the reference's own torch.sum is an
      // fp32 cascade; only the transmittance product needs the fp64 the reference's cumprod accumulates in
      float s_acc = 0.f, s_depth = 0.f, s_r = 0.f, s_g = 0.f, s_b = 0.f;
      int level = 0;   // thresholds (in ascending rank) whose first crossing is already known
      // A lane owns kSpl CONSECUTIVE samples of every 32*kSpl-sample chunk: its loads are 16*kSpl
      // contiguous bytes, the transmittance needs one warp scan per chunk (of the lanes' local
      // products) instead of one per 32 samples, and sample order = (lane, k) order.
      for (int c0 = 0; c0 < S; c0 += 32 * kSpl) {
        const int j0 = c0 + lane * kSpl;
        float4 v[kSpl];
        float zj[kSpl], sigma[kSpl], alpha[kSpl], xk[kSpl];
        bool valid[kSpl];
        float z_next = 0.f;            // depth of the sample after this lane's last one
#pragma unroll
        for (int k = 0; k < kSpl; ++k) {
          valid[k] = kFull || (j0 + k < S);
          v[k] = valid[k] ? rf_row[j0 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
          zj[k] = valid[k] ? z_row[j0 + k] : 0.f;
        }
        z_next = __shfl_down_sync(0xffffffffu, zj[0], 1);
        if (lane == 31 && j0 + kSpl < S) z_next = z_row[j0 + kSpl];
        double local = 1.0;            // product of this lane's x, in sample order
        double before[kSpl];           // ... of the samples before sample k within the lane
#pragma unroll
        for (int k = 0; k < kSpl; ++k) {
          const int j = j0 + k;
          const float zn = (k + 1 < kSpl) ? zj[k + 1] : z_next;
          float nz = 0.f;
          if (valid[k] && noise) nz = noise[ray * S + j];
          float dist = (j + 1 < S) ? __fsub_rn(zn, zj[k]) : 1e10f;
          dist = __fmul_rn(dist, norm);
          sigma[k] = fmaxf(__fadd_rn(v[k].w, nz), 0.0f);
          alpha[k] = valid[k] ? __fsub_rn(1.0f, expf(-__fmul_rn(sigma[k], dist))) : 0.0f;
          xk[k] = __fadd_rn(__fsub_rn(1.0f, alpha[k]), 1e-10f);
          before[k] = local;
          if (valid[k]) local *= (double)xk[k];
        }
        const double incl = warp_inclusive_product(local, lane);
        double excl = shfl_up_f64(incl, 1);
        if (lane == 0) excl = 1.0;
        const double lane_base = carry * excl;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int k = 0; k < kSpl; ++k) {
          const float trans = (float)(lane_base * before[k]);
          const float w = __fmul_rn(alpha[k], trans);
          if (valid[k]) {
            if (weights_out) weights_out[ray * S + j0 + k] = w;
            s_acc += w;
            s_depth = fmaf(w, zj[k], s_depth);
            s_r = fmaf(w, sigmoid_mufu(v[k].x), s_r);
            s_g = fmaf(w, sigmoid_mufu(v[k].y), s_g);
            s_b = fmaf(w, sigmoid_mufu(v[k].z), s_b);
          }
        }
        // ---- Dex-NeRF first crossings.  A chunk can only add crossings if some sigma exceeds the
        // smallest threshold still open; then each sample counts the thresholds it exceeds (upper
        // bound in the ascending table) and the open thresholds are closed in (lane, k) order.
        if (level < T) {
          float smax = 0.0f;
#pragma unroll
          for (int k = 0; k < kSpl; ++k) smax = fmaxf(smax, valid[k] ? sigma[k] : 0.0f);
          if (__any_sync(0xffffffffu, smax > s_thr[level])) {   // warp-uniform
            // cnt = thresholds a sample exceeds (branch-free search of the ascending table, padded with
            // +inf to 127 entries).  Threshold rank q is first crossed by the first sample, in (lane, k)
            // order, with cnt > q - i.e. by the samples that raise the running maximum of cnt.  The
            // exclusive running maximum is a lane-local pass plus one warp max-scan, and every record
            // sample then writes the ranks it is the first to cross: no serial loop over the thresholds.
            int cnt[kSpl], before[kSpl];
            int run = 0;
#pragma unroll
            for (int k = 0; k < kSpl; ++k) {
              const float sg = valid[k] ? sigma[k] : 0.0f;
              const float* p = s_thr;
#pragma unroll
              for (int b = 6; b >= 0; --b)
                if (p[(1 << b) - 1] < sg) p += 1 << b;
              cnt[k] = (int)(p - s_thr);
              before[k] = run;
              run = max(run, cnt[k]);
            }
            int incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int o = __shfl_up_sync(0xffffffffu, incl, d);
              if (lane >= d) incl = max(incl, o);
            }
            int excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0;
            excl = max(excl, level);
#pragma unroll
            for (int k = 0; k < kSpl; ++k) {
              const int from = max(excl, before[k]);
              for (int q = from; q < cnt[k]; ++q) s_found[warp][q] = c0 + lane * kSpl + k;
            }
            level = max(level, __shfl_sync(0xffffffffu, incl, 31));
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_acc += __shfl_xor_sync(0xffffffffu, s_acc, o);
        s_depth += __shfl_xor_sync(0xffffffffu, s_depth, o);
        s_r += __shfl_xor_sync(0xffffffffu, s_r, o);
        s_g += __shfl_xor_sync(0xffffffffu, s_g, o);
        s_b += __shfl_xor_sync(0xffffffffu, s_b, o);
      }
      if (lane == 0) {
        const float acc = s_acc, depth = s_depth;
        float r = s_r, g = s_g, b = s_b;
        const float q = __fdiv_rn(depth, acc);
        const float m = (q != q) ? q : fmaxf(1e-10f, q);  // torch.max propagates NaN (acc == 0)
        if (white_background) {
          const float bg = __fsub_rn(1.0f, acc);
          r = __fadd_rn(r, bg); g = __fadd_rn(g, bg); b = __fadd_rn(b, bg);
        }
        s_out[0 * 32 + slot] = r;
        s_out[1 * 32 + slot] = g;
        s_out[2 * 32 + slot] = b;
        s_out[3 * 32 + slot] = __fdiv_rn(1.0f, m);
        s_out[4 * 32 + slot] = acc;
        s_out[5 * 32 + slot] = depth;
      }
      __syncwarp();
      for (int k = lane; k < T; k += 32) {
        const int idx = (k < level) ? s_found[warp][k] : 0;   // argmax of an all-zero row is 0 (volume_rendering_utils.py:56)
        const int t = s_orig[k];
        s_out[(6 + t) * 32 + slot] = z_row[idx];
        reinterpret_cast<int*>(s_out)[(6 + T + t) * 32 + slot] = idx;
      }
      __syncwarp();
    }
    __syncthreads();
    // coalesced write-back of the 32-ray block
    const int64_t left = n - base;
    const int cnt = left < kRaysPerCta ? (int)left : kRaysPerCta;
    for (int e = threadIdx.x; e < (6 + 2 * T) * 32; e += blockDim.x) {
      const int plane = e >> 5, slot = e & 31;
      if (slot >= cnt) continue;
      const int64_t ray = base + slot;
      const float val = s_out[e];
      if (plane < 3) { if (rgb_out) rgb_out[ray * 3 + plane] = val; }
      else if (plane == 3) { if (disp_out) disp_out[ray] = val; }
      else if (plane == 4) { if (acc_out) acc_out[ray] = val; }
      else if (plane == 5) { if (depth_out) depth_out[ray] = val; }
      else if (plane < 6 + T) { if (dex_depth) dex_depth[(int64_t)(plane - 6) * dex_stride + ray] = val; }
      else if (dex_index) dex_index[(int64_t)(plane - 6 - T) * dex_stride + ray] =
          (int64_t) reinterpret_cast<const int*>(s_out)[e];
    }
    __syncthreads();
  }
}

// Backward of the compositing (training, BASELINE config 4): given dL/d(rgb_map), dL/d(depth_map),
// dL/d(acc_map) per ray, produce dL/d(radiance_field) (n,S,4).  Same warp-per-ray layout as the
// forward.  With gw_i = dL/dw_i = g_rgb.c_i + g_depth z_i + g_acc (- sum g_rgb if white bg):
//   dL/dalpha_i = gw_i T_i - (sum_{k>i} gw_k w_k) / x_i        x_i = 1 - alpha_i + 1e-10
//   dL/dsigma_i = dL/dalpha_i * dist_i * exp(-sigma_i dist_i),  masked by sigma_i > 0 (ReLU)
//   dL/drgb_raw = w_i g_rgb c_i (1 - c_i)                        c_i = sigmoid(rgb_raw_i)
// (what autograd derives for volume_rendering_utils.py:17-48 + cumprod_exclusive).  The forward
// quantities are recomputed: pass 1 collects the transmittance carried into every 32-sample chunk,
// pass 2 walks the chunks backwards with a warp-level suffix scan of gw_k w_k in fp64.
constexpr int kMaxChunks = 64;   // S <= 2048

__global__ void __launch_bounds__(kCompositeWarps * 32)
composite_bwd_kernel(const float4* __restrict__ rf, const float* __restrict__ z, const float* __restrict__ rd,
                     const float* __restrict__ noise, int64_t n, int S, int white_background,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
                     const float* __restrict__ g_acc, float4* __restrict__ d_rf) {
  __shared__ double s_carry[kCompositeWarps][kMaxChunks];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * kCompositeWarps;
  const int n_chunks = (S + 31) / 32;
  for (int64_t ray = (int64_t)blockIdx.x * kCompositeWarps + warp; ray < n; ray += warps) {
    const float dx = rd[ray * 3], dy = rd[ray * 3 + 1], dz = rd[ray * 3 + 2];
    const float norm =
        (float)sqrt((double)dx * (double)dx + (double)dy * (double)dy + (double)dz * (double)dz);
    const float gr = g_rgb ? g_rgb[ray * 3] : 0.f, gg = g_rgb ? g_rgb[ray * 3 + 1] : 0.f,
                gb = g_rgb ? g_rgb[ray * 3 + 2] : 0.f;
    const float gd = g_depth ? g_depth[ray] : 0.f;
    float ga = g_acc ? g_acc[ray] : 0.f;
    if (white_background) ga -= (gr + gg + gb);   // rgb_map += 1 - acc_map
    const float4* rf_row = rf + ray * S;
    const float* z_row = z + ray * S;
    auto sample = [&](int j, float4& v, float& zj, float& dist, float& sigma, float& e) {
      const bool valid = j < S;
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      zj = 0.f;
      float zn = 0.f, nz = 0.f;
      if (valid) {
        v = rf_row[j];
        zj = z_row[j];
        if (j + 1 < S) zn = z_row[j + 1];
        if (noise) nz = noise[ray * S + j];
      }
      dist = (j + 1 < S) ? __fsub_rn(zn, zj) : 1e10f;
      dist = __fmul_rn(dist, norm);
      sigma = fmaxf(__fadd_rn(v.w, nz), 0.0f);
      e = expf(-__fmul_rn(sigma, dist));
      return valid;
    };
    // pass 1: transmittance carried into each chunk
    double carry = 1.0;
    for (int c = 0; c < n_chunks; ++c) {
      float4 v; float zj, dist, sigma, e;
      const bool valid = sample(c * 32 + lane, v, zj, dist, sigma, e);
      const float alpha = valid ? __fsub_rn(1.0f, e) : 0.0f;
      const float x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      const double incl = warp_inclusive_product(valid ? (double)x : 1.0, lane);
      if (lane == 0) s_carry[warp][c] = carry;
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // pass 2: chunks in reverse
    double suffix = 0.0;
    for (int c = n_chunks - 1; c >= 0; --c) {
      const int j = c * 32 + lane;
      float4 v; float zj, dist, sigma, e;
      const bool valid = sample(j, v, zj, dist, sigma, e);
      const float alpha = valid ? __fsub_rn(1.0f, e) : 0.0f;
      const float x = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      const double incl = warp_inclusive_product(valid ? (double)x : 1.0, lane);
      double excl = shfl_up_f64(incl, 1);
      if (lane == 0) excl = 1.0;
      const float trans = (float)(s_carry[warp][c] * excl);
      const float w = __fmul_rn(alpha, trans);
      const float cr = 1.0f / (1.0f + expf(-v.x)), cg = 1.0f / (1.0f + expf(-v.y)),
                  cb = 1.0f / (1.0f + expf(-v.z));
      const float gw = gr * cr + gg * cg + gb * cb + gd * zj + ga;
      const double p = valid ? (double)gw * (double)w : 0.0;
      double rev = p;   // inclusive suffix sum over the lanes
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_down_sync(0xffffffffu, rev, d);
        if (lane + d < 32) rev += o;
      }
      double after = __shfl_down_sync(0xffffffffu, rev, 1);
      if (lane == 31) after = 0.0;
      const double R = after + suffix;
      suffix += __shfl_sync(0xffffffffu, rev, 0);
      if (valid) {
        const float dalpha = (float)((double)gw * (double)trans - R / (double)x);
        const float dsigma = (sigma > 0.0f) ? dalpha * dist * e : 0.0f;
        float4 o;
        o.x = w * gr * cr * (1.0f - cr);
        o.y = w * gg * cg * (1.0f - cg);
        o.z = w * gb * cb * (1.0f - cb);
        o.w = dsigma;
        d_rf[ray * S + j] = o;
      }
    }
    __syncwarp();
  }
}

// Stand-alone cumprod_exclusive (nerf/nerf_helpers.py:43-64): one warp per row.
__global__ void __launch_bounds__(256) cumprod_exclusive_kernel(const float* __restrict__ x,
                                                                int64_t n, int S,
                                                                float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n;
       row += warps) {
    double carry = 1.0;
    for (int c0 = 0; c0 < S; c0 += 32) {
      const int j = c0 + lane;
      const bool valid = j < S;
      const double incl = warp_inclusive_product(valid ? (double)x[row * S + j] : 1.0, lane);
      double excl = shfl_up_f64(incl, 1);
      if (lane == 0) excl = 1.0;
      if (valid) out[row * S + j] = (float)(carry * excl);
      carry *= __shfl_sync(0xffffffffu, incl, 31);
    }
  }
}

}  // namespace dexnerf

using namespace dexnerf;

// dex_stride: distance (in elements) between two threshold planes of dex_depth / dex_index - n for a stand-alone
// call, the full ray count when a chunk writes its columns of a (T, n_total) output (render.cu).
int dexnerf::volume_render_impl(const float* rf, const float* z, const float* rd,
                                     const float* noise, int64_t n, int S, int white_background,
                                     const float* thresholds, int T, float* rgb, float* disp,
                                     float* acc, float* weights, float* depth, float* dex_depth,
                                     int64_t* dex_index, int64_t dex_stride, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(rf && z && rd, "volume_render: null input");
  DN_REQUIRE(S >= 1, "volume_render: S < 1");
  DN_REQUIRE(T >= 0 && T <= kMaxThresholds, "volume_render: at most %d thresholds", kMaxThresholds);
  DN_REQUIRE(T == 0 || thresholds, "volume_render: thresholds is null");
  DN_REQUIRE((reinterpret_cast<uintptr_t>(rf) & 15) == 0, "volume_render: rf must be 16-byte aligned");
  const size_t smem = sizeof(float) * (2 * kMaxThresholds + (6 + 2 * (size_t)T) * 32);
  int64_t blocks = ceil_div64(n, kRaysPerCta);
  const int64_t cap = (int64_t)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  // consecutive samples per lane: a ray of 64 / 128 / 192 samples is ONE chunk (2 / 4 / 6 per lane: one
  // transmittance scan and one threshold vote per ray), 384 (C5) is two chunks of 6; other sizes take the
  // predicated 2-per-lane form (4 beyond 256 samples)
  auto launch = [&](auto kernel) {
    kernel<<<(int)blocks, kCompositeWarps * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(rf), z, rd, noise, n, S, white_background, thresholds, T, rgb,
        disp, acc, weights, depth, dex_depth, dex_index, dex_stride);
  };
  if (S <= 32) launch(composite_kernel<1, false>);
  else if (S % 192 == 0) launch(composite_kernel<6, true>);
  else if (S % 128 == 0) launch(composite_kernel<4, true>);
  else if (S % 64 == 0) launch(composite_kernel<2, true>);
  else if (S <= 256) launch(composite_kernel<2, false>);
  else launch(composite_kernel<4, false>);
  DN_CHECK_LAUNCH("volume_render");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_volume_render(const float* rf, const float* z, const float* rd,
                                     const float* noise, int64_t n, int S, int white_background,
                                     const float* thresholds, int T, float* rgb, float* disp,
                                     float* acc, float* weights, float* depth, float* dex_depth,
                                     int64_t* dex_index, void* stream) {
  return volume_render_impl(rf, z, rd, noise, n, S, white_background, thresholds, T, rgb, disp, acc, weights, depth,
                            dex_depth, dex_index, n, stream);
}

extern "C" DEXNERF_API int dexnerf_cumprod_exclusive(const float* x, int64_t n, int S, float* out, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(x && out, "cumprod_exclusive: null pointer");
  DN_REQUIRE(S >= 1, "cumprod_exclusive: S < 1");
  int64_t blocks = ceil_div64(n, 8);
  const int64_t cap = (int64_t)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  cumprod_exclusive_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, S, out);
  DN_CHECK_LAUNCH("cumprod_exclusive");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_volume_render_backward(const float* rf, const float* z, const float* rd,
                                                          const float* noise, int64_t n, int S,
                                                          int white_background, const float* g_rgb,
                                                          const float* g_depth, const float* g_acc,
                                                          float* d_rf, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  DN_REQUIRE(rf && z && rd && d_rf, "volume_render_backward: null pointer");
  DN_REQUIRE(S >= 1 && S <= 32 * kMaxChunks, "volume_render_backward: S must be in 1..%d", 32 * kMaxChunks);
  DN_REQUIRE((reinterpret_cast<uintptr_t>(rf) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_rf) & 15) == 0,
             "volume_render_backward: rf and d_rf must be 16-byte aligned");
  int64_t blocks = ceil_div64(n, kCompositeWarps);
  const int64_t cap = (int64_t)kNumSMs * 8 * 4;
  if (blocks > cap) blocks = cap;
  composite_bwd_kernel<<<(int)blocks, kCompositeWarps * 32, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(rf), z, rd, noise, n, S, white_background, g_rgb, g_depth, g_acc,
      reinterpret_cast<float4*>(d_rf));
  DN_CHECK_LAUNCH("volume_render_backward");
  return 0;
}
