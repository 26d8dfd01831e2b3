// Pieces shared by the tcgen05 forward kernels: mlp_tc.cu (pairs of tiles; hidden 256, training, debug taps)
// and mlp_tc3.cu (three tiles in flight; hidden 128 inference): kernel parameters, the in-register positional
// encoder, the hidden-128 epilogue pass.
#pragma once
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace dexnerf {
namespace tc {

constexpr int kSlotBytes = 16384;
constexpr int kNumSlots = 9;
constexpr int kThreads = 768;   // 4 control warps + 16 epilogue warps + 4 encoder warps
constexpr int kEpiThreadsPerTile = 256;
constexpr int kPeXyzBytes = kTileM * 64 * 2;  // 16 KB, K padded to 64
constexpr int kPeDirBytes = kTileM * 32 * 2;  // 8 KB,  K padded to 32
constexpr int kBiasImgBytes = kTileM * 16;    // 2 KB: one 16-byte row (8 bf16, the first two used) per output

struct TcParams {
  const uint8_t* weights;  // chunk images, consumption order
  const float* consts;     // biases | w_alpha | b_alpha | W_rgb | b_rgb
  const uint8_t* bias_img; // per layer 2 KB: [128 outputs][8 bf16] = (hi, lo, 0 ...) of the bias (mlp_tc3.cu: bias as an MMA)
  const float* ro; const float* rd; const float* vd; const float* z;
  float* rf;
  float* dbg;              // optional: raw accumulator dump of (dbg_layer, dbg_pass), [tile][128][128]
  int64_t m_total;
  int S;
  int n_layers, hidden, n_const, last_xyz_layer;
  int off_walpha, off_balpha, off_wrgb, off_brgb;
  int Lx, Ld, include_xyz, include_dir, log_xyz, log_dir, dim_xyz, dim_dir;
  int dbg_layer, dbg_pass;
  // training tape (kTape kernels only): bf16 operand images of every layer input, in the
  // MN-major half-tile layout the weight-gradient GEMM consumes ([tile][half][fg][64 rows][8]),
  // plus one ReLU bit per activation ([tile][slot][128 rows] x 64 bits)
  uint8_t* tape;
  int64_t tape_xyz, tape_dir;            // byte offsets of the encoding images
  int64_t tape_act[kMaxLayers];          // ... of each tensor-core layer's OUTPUT image
  int64_t tape_mask[kMaxLayers];         // ... of its ReLU mask (unused for layers without ReLU)
  TcLayer layers[kMaxLayers];
};

// sin and cos of an fp32 argument of any magnitude the encodings reach (|arg| < ~1e4): two-term
// Cody-Waite reduction by 2*pi (exact product in the FMA, one rounding) to [-pi, pi], then the
// MUFU approximations, whose absolute error there is < 2^-21 - three orders of magnitude below
// the bf16 rounding the operand gets next.  ~8 instructions instead of the ~100 of sinf + cosf.
__device__ __forceinline__ void sincos_reduced(float arg, float& s, float& c) {
#ifdef DEXNERF_ENC_CHEAP      // experiment (results invalid): what the encoders cost
  s = arg; c = arg; return;
#endif
  const float n = rintf(arg * 0.15915494309189535f);
  float r = fmaf(n, -6.2831854820251465f, arg);
  r = fmaf(n, 1.7484556000744883e-7f, r);
  s = __sinf(r);
  c = __cosf(r);
}

// One sample row of a positional-encoding operand tile in the standard configuration
// (include_input, log sampling: columns [x(3), sin(2^b x)(3), cos(2^b x)(3), ...]), written as
// bf16 UMMA core-matrix rows (16 bytes = 8 columns each, kTileM*16 bytes apart).  Fully unrolled:
// every column's (band, axis, sin|cos) is a compile-time constant and only one band's six values
// are live at a time.
template <int kGroups>
__device__ __forceinline__ void encode_row_std(const float (&x)[3], int dim, bool valid, uint8_t* dst,
                                               uint8_t* tape_dst) {
  float sv[3] = {0.f, 0.f, 0.f}, cv[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k8 = 0; k8 < kGroups; ++k8) {
    uint32_t w4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = k8 * 8 + 2 * j + h;
        float v;
        if (c < 3) {
          v = x[c];
        } else {
          const int band = (c - 3) / 6, rem = (c - 3) % 6;
          if (rem == 0 && c < dim) {
            const float f = (float)(1u << band);
#pragma unroll
            for (int a = 0; a < 3; ++a) sincos_reduced(__fmul_rn(x[a], f), sv[a], cv[a]);
          }
          v = rem < 3 ? sv[rem] : cv[rem - 3];
        }
        e[h] = (valid && c < dim) ? v : 0.f;
      }
      w4[j] = pack_bf16(e[0], e[1], false);
    }
    *reinterpret_cast<uint4*>(dst + k8 * (kTileM * 16)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    if (tape_dst) *reinterpret_cast<uint4*>(tape_dst + k8 * 1024) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
  }
}

__device__ __forceinline__ void stg128(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
#ifdef DEXNERF_EXP_NO_TAPE_STG      // experiment (tools/README.md): the forward kernel with tape minus its activation stores
  if (a == 0x12345678u && b == 0x9ABCDEF0u)
#endif
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

__device__ __forceinline__ int chunks_in_pass(const TcLayer& L) { return L.k_main / 64 + (L.smem_src ? 1 : 0); }

// The same pass for hidden = 128 kernels (one pass per layer, nothing held).  tcgen05.wait::ld waits for EVERY
// outstanding load of the thread, so the 16-column double buffering above is a chain of four ~200-cycle round trips
// (850 - 1 100 cycles per pass, longer than the 8-MMA pass of the other tile, which made the MMA -> epilogue -> MMA
// chain of a tile the limiter at hidden 128).  Here the warp's 64 columns are requested with
// fewer, deeper round trips - the hidden-128 kernels have the registers (no held[] array):
//   DEXNERF_WIDE_EPI == 1: both 32-column halves at once, ONE round trip, accumulator released before the arithmetic;
//   DEXNERF_WIDE_EPI == 2: three 16-column buffers, two loads in flight while a third slice is processed.
// The sigma head accumulates in four independent partial sums (a 64-deep dependent FMA chain otherwise).
#ifndef DEXNERF_WIDE_EPI
#define DEXNERF_WIDE_EPI 2
#endif
template <bool kRelu, bool kSig, bool kDbg, bool kTape, bool kBias = true, int kEpi = DEXNERF_WIDE_EPI>
__device__ __forceinline__ void epilogue_pass_wide(uint32_t d_tmem, uint32_t a_store, uint32_t bias, uint32_t wa,
                                                   float& sigma, uint32_t dfree_bar, float* dbg_dst,
                                                   uint8_t* tape_row, uint2* tape_mask) {
  uint32_t mbits[2] = {0u, 0u};
  float sg[4] = {0.f, 0.f, 0.f, 0.f};
  // 16 columns [c0, c0 + 16) of this warp's 64: + bias, head, ReLU, pack -> pk[0..8)
  auto slice16 = [&](const uint32_t* v, int c0, uint32_t* pk) {
    if (kDbg && dbg_dst) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dbg_dst[c0 + i] = __uint_as_float(v[i]);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float x0, x1, x2, x3;
      if (!kBias) {      // the accumulator already holds the bias (mlp_tc3.cu adds it as an MMA)
        x0 = __uint_as_float(v[i]); x1 = __uint_as_float(v[i + 1]); x2 = __uint_as_float(v[i + 2]); x3 = __uint_as_float(v[i + 3]);
      } else {
        const float4 b4 = lds128(bias + (uint32_t)((c0 + i) * 4));
        add_f32x2(v[i], v[i + 1], b4.x, b4.y, x0, x1);
        add_f32x2(v[i + 2], v[i + 3], b4.z, b4.w, x2, x3);
      }
      if (kSig) {
        const float4 w4 = lds128(wa + (uint32_t)((c0 + i) * 4));
        sg[0] = fmaf(kRelu ? fmaxf(x0, 0.0f) : x0, w4.x, sg[0]);
        sg[1] = fmaf(kRelu ? fmaxf(x1, 0.0f) : x1, w4.y, sg[1]);
        sg[2] = fmaf(kRelu ? fmaxf(x2, 0.0f) : x2, w4.z, sg[2]);
        sg[3] = fmaf(kRelu ? fmaxf(x3, 0.0f) : x3, w4.w, sg[3]);
      }
      pk[i / 2] = pack_bf16(x0, x1, kRelu);
      pk[i / 2 + 1] = pack_bf16(x2, x3, kRelu);
    }
    tmem_st8(a_store + (uint32_t)(c0 / 2), pk);
    if (kTape) {
      stg128(tape_row + (c0 / 8) * 1024, pk[0], pk[1], pk[2], pk[3]);
      stg128(tape_row + (c0 / 8 + 1) * 1024, pk[4], pk[5], pk[6], pk[7]);
      if (kRelu) mbits[c0 >> 5] |= relu_bits16(pk) << (c0 & 16);
    }
  };
  if constexpr (kEpi == 1) {
  uint32_t v[2][32];
  tmem_ld32_issue(d_tmem, v[0]);
  tmem_ld32_issue(d_tmem + 32, v[1]);
  tmem_ld32_wait(v[0]);
  tmem_ld32_tie(v[1]);
  tc_fence_before();
  mbar_arrive(dfree_bar);                 // all 64 columns are in registers
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t pk[8];
    slice16(&v[h][0], h * 32, pk);
    slice16(&v[h][16], h * 32 + 16, pk);
  }
  } else {
  uint32_t v[3][16];
  tmem_ld16_issue(d_tmem, v[0]);
  tmem_ld16_issue(d_tmem + 16, v[1]);
  tmem_ld16_wait(v[0]);
  tmem_ld16_tie(v[1]);
  tmem_ld16_issue(d_tmem + 32, v[2]);
  uint32_t pk[8];
  slice16(v[0], 0, pk);
  tmem_ld16_issue_tied(d_tmem + 48, v[0], v[1]);
  slice16(v[1], 16, pk);
  tmem_ld16_wait(v[2]);
  tmem_ld16_tie(v[0]);
  tc_fence_before();
  mbar_arrive(dfree_bar);
  slice16(v[2], 32, pk);
  slice16(v[0], 48, pk);
  }
  if (kSig) sigma += (sg[0] + sg[1]) + (sg[2] + sg[3]);
  if (kTape && kRelu) *tape_mask = make_uint2(mbits[0], mbits[1]);
}

// mlp_tc3.cu: the three-tile kernel for hidden-128 inference (FlexibleNeRFModel with or without view directions)
int launch_mlp_tc3(const TcParams& P, int64_t n_tiles, cudaStream_t st);

}  // namespace tc
}  // namespace dexnerf
