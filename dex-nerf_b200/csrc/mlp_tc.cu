// tcgen05 tensor-core evaluation of FlexibleNeRFModel (nerf/models.py:185-256, repaired forward)
// fused with the point construction and positional encoding of run_network
// (nerf/train_utils.py:72-89, :136).  bf16 operands, fp32 accumulation in TMEM.
//
// One persistent CTA per SM processes PAIRS of 128-sample tiles (tile 0 / tile 1) so that the
// tensor pipe always has an MMA pass of one tile to run while the other tile is in its epilogue:
//
//   warp 0        weight producer: streams pre-packed bf16 UMMA weight images (<=16 KB chunks,
//                 128 out-features x 64 in-features) from L2 into a 9-slot shared-memory ring with
//                 cp.async.bulk (TMA bulk copy, mbarrier complete_tx); each chunk is consumed by
//                 both tiles before the slot is recycled.
//   warp 1        MMA issuer: one thread issues tcgen05.mma (M=128, N=128, K=16).  Hidden
//                 activations are the A operand READ FROM TENSOR MEMORY (bf16, 2 per column);
//                 the positional encodings (layer 1, skip layer, view-direction layer) are A
//                 operands read from shared memory.  A 256-wide layer runs as two N=128 passes
//                 into a 128-column fp32 accumulator, so a tile needs 128 (A) + 128 (D) TMEM
//                 columns and two tiles fill the 512 columns.
//   warp 2        TMEM allocator.
//   warps 4-11    epilogue, 4 warps per tile, one thread per sample row: tcgen05.ld the accumulator,
//                 + bias, ReLU, pack to bf16 and tcgen05.st it back as the next layer's A operand
//                 (first-pass results wait in registers until the layer's second pass has
//                 finished reading the old A).  fc_alpha and fc_rgb (1 and 3 outputs) are fp32
//                 dot products on the CUDA cores inside the epilogue; the only HBM write is the
//                 final (r,g,b,sigma) float4 per sample.
//   warps 12-15   encoders: one thread per sample row of the NEXT tile pair computes
//                 pts = ro + rd*z and the sin/cos encodings in registers and writes them as
//                 bf16 UMMA core matrices into shared memory.
//
// Weight-image layout (no swizzle, K-major "interleaved" canonical layout): an operand tile is a
// grid of 8-row x 16-byte core matrices, each 128 contiguous bytes; byte offset of element
// (row r, k) = (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2 with SBO = 128 and LBO = rows*16.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dexnerf {
namespace tc {

constexpr int kTileM = 128;
constexpr int kSlotBytes = 16384;
constexpr int kNumSlots = 9;
constexpr int kMaxLayers = 16;
constexpr int kThreads = 768;   // 4 control warps + 16 epilogue warps + 4 encoder warps
constexpr int kEpiThreadsPerTile = 256;
constexpr int kPeXyzBytes = kTileM * 64 * 2;  // 16 KB, K padded to 64
constexpr int kPeDirBytes = kTileM * 32 * 2;  // 8 KB,  K padded to 32
constexpr int kMaxConstFloats = 4096;
constexpr uint32_t kSpinLimit = 1u << 22;   // ~ a second of polling, then trap

struct TcLayer {
  int k_main;    // K read from the TMEM-resident activations (0 for layer1)
  int smem_src;  // 0 none, 1 xyz encoding, 2 dir encoding
  int k_smem;    // padded K of the shared-memory operand
  int n_out;     // output features
  int n_pass;    // passes of 128 (or n_out when smaller)
  int relu;
  int head;      // 1: sigma head is evaluated in this layer's epilogue, 2: rgb head + final store
  int bias_off;  // into the const block (floats)
};

struct TcParams {
  const uint8_t* weights;  // chunk images, consumption order
  const float* consts;     // biases | w_alpha | b_alpha | W_rgb | b_rgb
  const float* ro; const float* rd; const float* vd; const float* z;
  float* rf;
  float* dbg;              // optional: raw accumulator dump of (dbg_layer, dbg_pass), [tile][128][128]
  int64_t m_total;
  int S;
  int n_layers, hidden, n_const, last_xyz_layer;
  int off_walpha, off_balpha, off_wrgb, off_brgb;
  int Lx, Ld, include_xyz, include_dir, log_xyz, log_dir, dim_xyz, dim_dir;
  int dbg_layer, dbg_pass;
  TcLayer layers[kMaxLayers];
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end in a trap (launch failure), never in a hung GPU.
__device__ __noinline__ void barrier_timeout(int who) {
  printf("dexnerf mlp_tc: barrier timeout (wait site %d, block %d, thread %d)\n", who, blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) barrier_timeout(who);
  }
}
// Four barrier polls issued back to back (their ~100-cycle latencies overlap); falls back to the
// bounded sequential wait when any of them is not complete yet.
__device__ __forceinline__ void mbar_wait4(uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, uint32_t b2,
                                           uint32_t p2, uint32_t b3, uint32_t p3, int who) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred q0, q1, q2, q3;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
      "and.pred q0, q0, q1;\n\tand.pred q2, q2, q3;\n\tand.pred q0, q0, q2;\n\t"
      "selp.u32 %0, 1, 0, q0;\n\t}"
      : "=r"(ok) : "r"(b0), "r"(p0), "r"(b1), "r"(p1), "r"(b2), "r"(p2), "r"(b3), "r"(p3) : "memory");
  if (!ok) { mbar_wait(b0, p0, who); mbar_wait(b1, p1, who); mbar_wait(b2, p2, who); mbar_wait(b3, p3, who); }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// split-phase accumulator load: issue now, make the registers valid later with tmem_ld32_wait
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// the registers are listed as read-write so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :: "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// Same, but also lists the slice that is about to be processed as read-write: every use of `cur`
// is then ordered AFTER this load has been issued, so its latency overlaps the arithmetic on `cur`
// (the compiler would otherwise sink the issue below the arithmetic to save registers).
__device__ __forceinline__ void tmem_ld16_issue_tied(uint32_t taddr, uint32_t (&r)[16], uint32_t (&cur)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "+r"(cur[0]), "+r"(cur[1]), "+r"(cur[2]), "+r"(cur[3]), "+r"(cur[4]), "+r"(cur[5]), "+r"(cur[6]), "+r"(cur[7]),
        "+r"(cur[8]), "+r"(cur[9]), "+r"(cur[10]), "+r"(cur[11]), "+r"(cur[12]), "+r"(cur[13]), "+r"(cur[14]), "+r"(cur[15])
      : "r"(taddr));
}
// read-only constants (biases, head weights) straight from shared memory; not volatile, so the
// compiler may hoist these loads above the TMEM waits
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// two fp32 -> packed bf16x2 (lo = first K element), optional ReLU
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi, bool relu) {
  uint32_t d;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// sin and cos of an fp32 argument of any magnitude the encodings reach (|arg| < ~1e4): two-term
// Cody-Waite reduction by 2*pi (exact product in the FMA, one rounding) to [-pi, pi], then the
// MUFU approximations, whose absolute error there is < 2^-21 - three orders of magnitude below
// the bf16 rounding the operand gets next.  ~8 instructions instead of the ~100 of sinf + cosf.
__device__ __forceinline__ void sincos_reduced(float arg, float& s, float& c) {
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
__device__ __forceinline__ void encode_row_std(const float (&x)[3], int dim, bool valid, uint8_t* dst) {
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
  }
}

// ------------------------------------------------------------------ shared-memory map
struct Smem {
  // offsets from the 1024-aligned base
  static constexpr int w_slots = 0;
  static constexpr int pe_xyz = w_slots + kNumSlots * kSlotBytes;        // [tile]
  static constexpr int pe_dir = pe_xyz + 2 * kPeXyzBytes;                // [tile]
  static constexpr int consts = pe_dir + 2 * kPeDirBytes;
  static constexpr int xchg = consts + kMaxConstFloats * 4;             // [tile][row] float4 head partials
  static constexpr int bars = xchg + 2 * kTileM * 16;
  static constexpr int n_bars = 2 * kNumSlots + 4 + 4 + 2 + 2 + 2;
  static constexpr int tmem_ptr = bars + n_bars * 8;
  static constexpr int total = tmem_ptr + 16;
};
// barrier indices
__device__ __forceinline__ int B_wfull(int s) { return s; }
__device__ __forceinline__ int B_wempty(int s) { return kNumSlots + s; }
__device__ __forceinline__ int B_xyzfull(int t) { return 2 * kNumSlots + t; }
__device__ __forceinline__ int B_xyzempty(int t) { return 2 * kNumSlots + 2 + t; }
__device__ __forceinline__ int B_dirfull(int t) { return 2 * kNumSlots + 4 + t; }
__device__ __forceinline__ int B_dirempty(int t) { return 2 * kNumSlots + 6 + t; }
__device__ __forceinline__ int B_aready(int t) { return 2 * kNumSlots + 8 + t; }
__device__ __forceinline__ int B_dfull(int t) { return 2 * kNumSlots + 10 + t; }
__device__ __forceinline__ int B_dfree(int t) { return 2 * kNumSlots + 12 + t; }

__device__ __forceinline__ int chunks_in_pass(const TcLayer& L) { return L.k_main / 64 + (L.smem_src ? 1 : 0); }

// One epilogue pass of one warp over ITS 64 of the 128 accumulator columns of a tile (one thread per
// sample row; two warps per 32-row quarter split the columns): + bias, optional sigma head,
// optional ReLU, pack to bf16, then either keep the packed words in registers (first pass of a
// 256-wide layer) or store them as the next layer's A operand.
template <bool kRelu, bool kSig, bool kHold, bool kPark, bool kDbg>
__device__ __forceinline__ void epilogue_pass(uint32_t d_tmem, uint32_t a_park, uint32_t a_store, uint32_t bias,
                                              uint32_t wa, float& sigma, uint32_t (&held)[32],
                                              uint32_t dfree_bar, float* dbg_dst, long long* stamps) {
  uint32_t v[2][16];      // 16-column slices of the accumulator
  if (kDbg && stamps) stamps[0] = clock64();
  tmem_ld16_issue(d_tmem, v[0]);
  if (kPark) {
    // the old A is dead now: park the first pass's share of the new A
    tmem_st16(a_park, &held[0]);
    tmem_st16(a_park + 16, &held[16]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {     // 4 x 16 columns; unrolled so held[] has static indices
    tmem_ld16_wait(v[c & 1]);
    if (c + 1 < 4) tmem_ld16_issue(d_tmem + (uint32_t)((c + 1) * 16), v[(c + 1) & 1]);
    if (kDbg && dbg_dst) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dbg_dst[c * 16 + i] = __uint_as_float(v[c & 1][i]);
    }
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b4 = lds128(bias + (uint32_t)((c * 16 + i) * 4));
      const float x0 = __uint_as_float(v[c & 1][i]) + b4.x;
      const float x1 = __uint_as_float(v[c & 1][i + 1]) + b4.y;
      const float x2 = __uint_as_float(v[c & 1][i + 2]) + b4.z;
      const float x3 = __uint_as_float(v[c & 1][i + 3]) + b4.w;
      if (kSig) {   // fc_alpha on the rectified, unrounded trunk output
        const float4 w4 = lds128(wa + (uint32_t)((c * 16 + i) * 4));
        sigma = fmaf(fmaxf(x0, 0.0f), w4.x, sigma);
        sigma = fmaf(fmaxf(x1, 0.0f), w4.y, sigma);
        sigma = fmaf(fmaxf(x2, 0.0f), w4.z, sigma);
        sigma = fmaf(fmaxf(x3, 0.0f), w4.w, sigma);
      }
      if (kHold) {
        held[c * 8 + i / 2] = pack_bf16(x0, x1, kRelu);
        held[c * 8 + i / 2 + 1] = pack_bf16(x2, x3, kRelu);
      } else {
        pk[i / 2] = pack_bf16(x0, x1, kRelu);
        pk[i / 2 + 1] = pack_bf16(x2, x3, kRelu);
      }
    }
    if (!kHold) tmem_st8(a_store + (uint32_t)(c * 8), pk);
    if (kDbg && stamps) stamps[1 + c] = clock64();
    if (c == 3) {
      // this warp's columns are fully read (the last wait::ld covered them)
      tc_fence_before();
      mbar_arrive(dfree_bar);
    }
  }
}

// ------------------------------------------------------------------ the kernel
template <int H, bool kDbg>
__global__ void __launch_bounds__(kThreads, 1) mlp_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + Smem::bars;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (P.m_total + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < kNumSlots; ++s) { mbar_init(bar(B_wfull(s)), 1); mbar_init(bar(B_wempty(s)), 2); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(bar(B_xyzfull(t)), 128); mbar_init(bar(B_xyzempty(t)), 1);
      mbar_init(bar(B_dirfull(t)), 128); mbar_init(bar(B_dirempty(t)), 1);
      mbar_init(bar(B_aready(t)), kEpiThreadsPerTile);
      mbar_init(bar(B_dfull(t)), 1);
      mbar_init(bar(B_dfree(t)), kEpiThreadsPerTile);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::tmem_ptr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    float* c = reinterpret_cast<float*>(smem + Smem::consts);
    for (int i = threadIdx.x; i < P.n_const; i += kThreads) c[i] = P.consts[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + Smem::tmem_ptr);
  const float* s_const = reinterpret_cast<const float*>(smem + Smem::consts);

  // register budget: 768 threads x 80 registers, no re-balancing (setmaxnreg.inc can only draw
  // from registers other warpgroups of the CTA released with setmaxnreg.dec).
  if (warp < 4) {
  if (warp == 0) {
    // =============================== weight producer ===============================
    const bool leader = elect_one();
    uint32_t cnt = 0;
#pragma unroll 1
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const uint8_t* src = P.weights;
#pragma unroll 1
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const int np = L.n_out < 128 ? L.n_out : 128;
        const int nc = chunks_in_pass(L);
#pragma unroll 1
        for (int p = 0; p < L.n_pass; ++p) {
#pragma unroll 1
          for (int c = 0; c < nc; ++c, ++cnt) {
            const int kc = (c < L.k_main / 64) ? 64 : L.k_smem;
            uint32_t bytes = (uint32_t)(np * kc * 2);
            if (kDbg && P.dbg_layer == -3) bytes = 1024;   // experiment: timing without the weight traffic
            const uint32_t slot = cnt % kNumSlots, ph = (cnt / kNumSlots) & 1;
            mbar_wait(bar(B_wempty(slot)), ph ^ 1, 0);
            if (leader) {
              mbar_arrive_expect_tx(bar(B_wfull(slot)), bytes);
              bulk_g2s(sbase + Smem::w_slots + slot * kSlotBytes, src, bytes, bar(B_wfull(slot)));
            }
            __syncwarp();
            src += bytes;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the (warp-uniform) control flow so that descriptors stay cheap to form;
    // only the asynchronous instructions themselves are issued by one elected lane.  Tile 1 reuses
    // the weight chunks tile 0 has just waited for (they stay resident until both tiles have
    // committed), so only tile 0 polls the weight barriers.
    const bool leader = elect_one();
    uint32_t w_slot = 0, w_phase = 0;           // ring cursor of tile 0 (tile 1 trails by one pass)
    uint32_t ph_dfree[2] = {0, 0}, ph_aready[2] = {0, 0};
    uint32_t it = 0;
    const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);   // SBO = 128 B, version 1
    const uint32_t slot0_lo = ((sbase + Smem::w_slots) >> 4) & 0x3FFF;      // 16-byte units
    constexpr int kMain = H / 64;               // 64-wide K chunks of a hidden-activation operand
#pragma unroll 1
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const uint32_t pe_ph = it & 1;
#pragma unroll 1
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const int np = L.n_out < 128 ? L.n_out : 128;
        const uint32_t idesc = instr_desc(np);
        const bool has_main = L.k_main != 0;
        const uint32_t b_lbo16 = (uint32_t)np;            // LBO = np*16 bytes -> np in 16-byte units
        const bool timing = kDbg && P.dbg_layer == -2 && blockIdx.x == 0 && it == 0 && leader;
#pragma unroll 1
        for (int p = 0; p < L.n_pass; ++p) {
          const uint32_t slot_p = w_slot, phase_p = w_phase;   // first chunk of this pass
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t a_tmem = tmem_base + (uint32_t)(t * 256);
            const uint32_t d_tmem = a_tmem + 128;
            // Gate: the accumulator must have been drained by this tile's previous epilogue pass and,
            // for the first pass of a layer, the new A operand must be in place.  a_ready is
            // arrived after d_free by every epilogue thread, so one poll covers both.
            if (p == 0 && l > 0) {
              mbar_wait(bar(B_aready(t)), ph_aready[t], 3);
              ph_aready[t] ^= 1;
            } else {
              mbar_wait(bar(B_dfree(t)), ph_dfree[t] ^ 1, 1);
              if (l == 0) mbar_wait(bar(B_xyzfull(t)), pe_ph, 2);
            }
            ph_dfree[t] ^= 1;
            uint32_t slot = slot_p, phase = phase_p;
            if (t == 0 && has_main) {
              // all main chunks of the pass are normally resident already (the ring runs ahead)
              uint32_t sl[4], ph4[4];
              uint32_t s2 = slot, p2 = phase;
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                sl[c] = bar(B_wfull(s2)); ph4[c] = p2;
                if (c + 1 < kMain) { if (++s2 == kNumSlots) { s2 = 0; p2 ^= 1; } }
              }
              if (kMain == 4) mbar_wait4(sl[0], ph4[0], sl[1], ph4[1], sl[2], ph4[2], sl[3], ph4[3], 4);
              else mbar_wait4(sl[0], ph4[0], sl[1], ph4[1], sl[1], ph4[1], sl[1], ph4[1], 4);
            }
            tc_fence_after();
            if (timing) reinterpret_cast<long long*>(P.dbg)[((l * 2 + p) * 2 + t) * 2] = clock64();
            if (has_main) {
#pragma unroll 1
              for (int c = 0; c < kMain; ++c) {
                // descriptor low word: address (16-byte units) | LBO << 16; K-step advance = 2*LBO
                const uint32_t b_lo = (slot0_lo + slot * (kSlotBytes >> 4)) | (b_lbo16 << 16);
                if (leader) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    mma_ts(d_tmem, a_tmem + (uint32_t)(c * 32 + ks * 8),
                           desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 2 * b_lbo16), idesc, (c | ks) ? 1u : 0u);
                  tc_commit(bar(B_wempty(slot)));   // slot is refilled once both tiles' MMAs retire
                }
                __syncwarp();
                if (++slot == kNumSlots) { slot = 0; phase ^= 1; }
              }
            }
            if (L.smem_src) {
              if (t == 0) { mbar_wait(bar(B_wfull(slot)), phase, 4); tc_fence_after(); }
              const uint32_t b_lo = (slot0_lo + slot * (kSlotBytes >> 4)) | (b_lbo16 << 16);
              if (L.smem_src == 2 && p == 0) mbar_wait(bar(B_dirfull(t)), pe_ph, 2);
              const uint32_t a_addr = (L.smem_src == 1)
                  ? sbase + Smem::pe_xyz + (uint32_t)t * kPeXyzBytes
                  : sbase + Smem::pe_dir + (uint32_t)t * kPeDirBytes;
              const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | ((uint32_t)kTileM << 16);   // LBO = 128*16 B
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  if (ks * 16 < L.k_smem)
                    mma_ss(d_tmem, desc_hi | (uint64_t)(a_lo + (uint32_t)ks * 2 * kTileM),
                           desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 2 * b_lbo16), idesc,
                           (has_main || ks) ? 1u : 0u);
                }
                tc_commit(bar(B_wempty(slot)));
              }
              __syncwarp();
              if (++slot == kNumSlots) { slot = 0; phase ^= 1; }
            }
            if (leader) {
              tc_commit(bar(B_dfull(t)));
              if (p == L.n_pass - 1) {
                if (l == P.last_xyz_layer) tc_commit(bar(B_xyzempty(t)));   // encoders may refill
                if (L.smem_src == 2) tc_commit(bar(B_dirempty(t)));
              }
            }
            __syncwarp();
            if (timing) reinterpret_cast<long long*>(P.dbg)[((l * 2 + p) * 2 + t) * 2 + 1] = clock64();
            if (t == 1) { w_slot = slot; w_phase = phase; }
          }
        }
      }
    }
  }
  } else if (warp < 20) {
    // =============================== epilogue ===============================
    // 16 warps: tile t = e / 8, column half hs = (e / 4) % 2, lane quarter q = warp % 4 (a warp can
    // only touch the 32 TMEM lanes of its quarter).  Two warps share each 32-row quarter so that
    // one computes while the other waits on its tcgen05.ld.
    const int e = warp - 4;
    const int t = e >> 3, hs = (e >> 2) & 1, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_tmem = tmem_base + (uint32_t)(t * 256) + lane_base;
    const uint32_t d_tmem = a_tmem + 128 + (uint32_t)(hs * 64);
    const bool timing = kDbg && P.dbg_layer == -2 && blockIdx.x == 0 && q == 0 && hs == 0 && lane == 0;
    long long* tl = reinterpret_cast<long long*>(P.dbg);
    float4* xchg = reinterpret_cast<float4*>(smem + Smem::xchg) + t * kTileM + row;
    const int pair_bar = 1 + t * 4 + q;            // named barrier of the two warps of this quarter
    uint32_t ph_dfull = 0;
    uint32_t held[32];   // first-pass results of a 2-pass layer (this warp's 64 outputs as bf16 pairs)
#pragma unroll 1
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int64_t g = (pair * 2 + t) * kTileM + row;
      float sigma = 0.0f;
#pragma unroll 1
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        if (L.head == 2) {
          // ---- last layer: ReLU, fc_rgb on the CUDA cores, final (r,g,b,sigma) store
          mbar_wait(bar(B_dfull(t)), ph_dfull, 5);
          ph_dfull ^= 1;
          tc_fence_after();
          if (timing && pair == blockIdx.x) tl[256 + ((l * 2) * 2 + t) * 2] = clock64();
          float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
          constexpr int hw = H / 2;           // outputs of the dir layer
          constexpr int mine = hw / 2;        // columns this warp reduces
          const uint32_t col0 = (uint32_t)(hs * mine);
          const uint32_t bias = sbase + Smem::consts + ((uint32_t)L.bias_off + col0) * 4;
          const uint32_t wr = sbase + Smem::consts + ((uint32_t)P.off_wrgb + col0) * 4;
          const uint32_t d_last = a_tmem + 128 + col0;
          uint32_t v[2][16];
          tmem_ld16_issue(d_last, v[0]);
#pragma unroll
          for (int c = 0; c < mine / 16; ++c) {
            tmem_ld16_wait(v[c & 1]);
            if (c + 1 < mine / 16) tmem_ld16_issue(d_last + (uint32_t)((c + 1) * 16), v[(c + 1) & 1]);
            if (kDbg && l == P.dbg_layer && g < P.m_total) {
              float* dst = P.dbg + g * 128 + col0 + c * 16;
#pragma unroll
              for (int i = 0; i < 16; ++i) dst[i] = __uint_as_float(v[c & 1][i]);
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 b4 = lds128(bias + (uint32_t)((c * 16 + i) * 4));
              const float4 r4 = lds128(wr + (uint32_t)((c * 16 + i) * 4));
              const float4 g4 = lds128(wr + (uint32_t)((hw + c * 16 + i) * 4));
              const float4 u4 = lds128(wr + (uint32_t)((2 * hw + c * 16 + i) * 4));
              const float x0 = fmaxf(__uint_as_float(v[c & 1][i]) + b4.x, 0.0f);
              const float x1 = fmaxf(__uint_as_float(v[c & 1][i + 1]) + b4.y, 0.0f);
              const float x2 = fmaxf(__uint_as_float(v[c & 1][i + 2]) + b4.z, 0.0f);
              const float x3 = fmaxf(__uint_as_float(v[c & 1][i + 3]) + b4.w, 0.0f);
              rgb0 = fmaf(x0, r4.x, fmaf(x1, r4.y, fmaf(x2, r4.z, fmaf(x3, r4.w, rgb0))));
              rgb1 = fmaf(x0, g4.x, fmaf(x1, g4.y, fmaf(x2, g4.z, fmaf(x3, g4.w, rgb1))));
              rgb2 = fmaf(x0, u4.x, fmaf(x1, u4.y, fmaf(x2, u4.z, fmaf(x3, u4.w, rgb2))));
            }
          }
          tc_fence_before();
          mbar_arrive(bar(B_dfree(t)));
          if (timing && pair == blockIdx.x) tl[256 + ((l * 2) * 2 + t) * 2 + 1] = clock64();
          // combine the two column halves: hs == 1 hands its partial sums to hs == 0
          if (hs == 1) *xchg = make_float4(rgb0, rgb1, rgb2, sigma);
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          if (hs == 0 && g < P.m_total) {
            const float4 o2 = *xchg;
            const float* br = s_const + P.off_brgb;
            float4 o;
            o.x = rgb0 + o2.x + br[0]; o.y = rgb1 + o2.y + br[1]; o.z = rgb2 + o2.z + br[2];
            o.w = sigma + o2.w + s_const[P.off_balpha];
            reinterpret_cast<float4*>(P.rf)[g] = o;
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // xchg may be rewritten
        } else {
          const uint32_t bias = sbase + Smem::consts + ((uint32_t)L.bias_off + hs * 64) * 4;
          const uint32_t wa = sbase + Smem::consts + ((uint32_t)P.off_walpha + hs * 64) * 4;
          const int kind = L.relu ? (L.head == 1 ? 2 : 1) : 0;
#pragma unroll
          for (int p = 0; p < H / 128; ++p) {
            mbar_wait(bar(B_dfull(t)), ph_dfull, 5);
            ph_dfull ^= 1;
            tc_fence_after();
            if (timing && pair == blockIdx.x) tl[256 + ((l * 2 + p) * 2 + t) * 2] = clock64();
            constexpr bool kTwoPass = (H == 256);
            const bool hold = kTwoPass && p == 0;
            float* dbg_dst = (kDbg && P.dbg_layer == l && P.dbg_pass == p && g < P.m_total)
                                 ? P.dbg + g * 128 + hs * 64 : nullptr;
            const uint32_t bp = bias + (uint32_t)(p * 128 * 4);
            const uint32_t wp = wa + (uint32_t)(p * 128 * 4);
            const uint32_t dfree = bar(B_dfree(t));
            const uint32_t a_park = a_tmem + (uint32_t)(hs * 32);                 // pass-0 outputs: K [hs*64, +64)
            const uint32_t a_store = a_tmem + (uint32_t)(p * 64 + hs * 32);       // pass-p outputs
            long long* stamps = (timing && pair == blockIdx.x && l == 2) ? tl + 600 + (p * 2 + t) * 16 : nullptr;
            if (hold) {
              if (kind == 0) epilogue_pass<false, false, true, false, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
              else if (kind == 1) epilogue_pass<true, false, true, false, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
              else epilogue_pass<true, true, true, false, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
            } else {
              constexpr bool park = kTwoPass;
              if (kind == 0) epilogue_pass<false, false, false, park, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
              else if (kind == 1) epilogue_pass<true, false, false, park, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
              else epilogue_pass<true, true, false, park, kDbg>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps);
              if (kDbg && stamps) stamps[9] = clock64();
              tmem_wait_st();
              if (kDbg && stamps) stamps[10] = clock64();
              tc_fence_before();
              mbar_arrive(bar(B_aready(t)));
            }
            if (timing && pair == blockIdx.x) tl[256 + ((l * 2 + p) * 2 + t) * 2 + 1] = clock64();
          }
        }
      }
    }
  } else {
    // =============================== encoders ===============================
    const int row = (warp - 20) * 32 + lane;
    const bool std_xyz = P.include_xyz && P.log_xyz, std_dir = P.include_dir && P.log_dir;
    uint32_t it = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      // xyz tiles first (needed by the pair's very first layer), then the dir tiles (only needed
      // by its last layer, and only free once the previous pair has completely retired)
      for (int t = 0; t < 2; ++t) {
        const int64_t g = (pair * 2 + t) * kTileM + row;
        float pt[3] = {0.f, 0.f, 0.f};
        const bool valid = g < P.m_total;
        if (valid) {
          const int64_t ray = g / P.S;
          const float zz = P.z[g];
#pragma unroll
          for (int a = 0; a < 3; ++a) pt[a] = __fadd_rn(P.ro[ray * 3 + a], __fmul_rn(P.rd[ray * 3 + a], zz));
        }
        mbar_wait(bar(B_xyzempty(t)), ph ^ 1, 6);
        uint8_t* xyz = smem + Smem::pe_xyz + t * kPeXyzBytes + row * 16;
        if (std_xyz) {
          encode_row_std<8>(pt, P.dim_xyz, valid, xyz);
        } else {
          for (int k8 = 0; k8 < 8; ++k8) {   // 8 encoding columns = one 16-byte core-matrix row
            uint32_t w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c0 = k8 * 8 + 2 * j;
              const float e0 = (valid && c0 < P.dim_xyz) ? pe_column(pt, c0, P.Lx, P.include_xyz, P.log_xyz) : 0.f;
              const float e1 = (valid && c0 + 1 < P.dim_xyz) ? pe_column(pt, c0 + 1, P.Lx, P.include_xyz, P.log_xyz) : 0.f;
              w4[j] = pack_bf16(e0, e1, false);
            }
            *reinterpret_cast<uint4*>(xyz + k8 * (kTileM * 16)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar(B_xyzfull(t)));
      }
      for (int t = 0; t < 2; ++t) {
        const int64_t g = (pair * 2 + t) * kTileM + row;
        float dir[3] = {0.f, 0.f, 0.f};
        const bool valid = g < P.m_total;
        if (valid) {
          const int64_t ray = g / P.S;
#pragma unroll
          for (int a = 0; a < 3; ++a) dir[a] = P.vd[ray * 3 + a];
        }
        mbar_wait(bar(B_dirempty(t)), ph ^ 1, 7);
        uint8_t* dr = smem + Smem::pe_dir + t * kPeDirBytes + row * 16;
        if (std_dir) {
          encode_row_std<4>(dir, P.dim_dir, valid, dr);
        } else {
          for (int k8 = 0; k8 < 4; ++k8) {
            uint32_t w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c0 = k8 * 8 + 2 * j;
              const float e0 = (valid && c0 < P.dim_dir) ? pe_column(dir, c0, P.Ld, P.include_dir, P.log_dir) : 0.f;
              const float e1 = (valid && c0 + 1 < P.dim_dir) ? pe_column(dir, c0 + 1, P.Ld, P.include_dir, P.log_dir) : 0.f;
              w4[j] = pack_bf16(e0, e1, false);
            }
            *reinterpret_cast<uint4*>(dr + k8 * (kTileM * 16)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
        fence_proxy_async();
        mbar_arrive(bar(B_dirfull(t)));
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------ host: layer table + packing
struct HostLayer { TcLayer tc; int prog_op; };

struct Plan {
  int n_layers = 0;
  HostLayer layers[kMaxLayers];
  int op_alpha = -1, op_rgb = -1;
  int off_walpha = 0, off_balpha = 0, off_wrgb = 0, off_brgb = 0, n_const = 0;
  int64_t weight_bytes = 0;
  int kx = 0, kd = 0;
};

static int pad16(int v) { return (v + 15) / 16 * 16; }

static int make_plan(const dexnerf_flexible_spec* s, Plan* plan) {
  DN_REQUIRE(s, "tc: null spec");
  DN_REQUIRE(s->hidden == 256 || s->hidden == 128, "tc: hidden must be 128 or 256 (got %d)", s->hidden);
  DN_REQUIRE(s->n_trunk >= 1 && s->n_trunk + 3 <= kMaxLayers, "tc: unsupported trunk depth %d", s->n_trunk);
  DN_REQUIRE(s->skip_every >= 1, "tc: skip_every < 1");
  DN_REQUIRE(s->dim_xyz >= 1 && s->dim_xyz <= 64, "tc: dim_xyz must be <= 64 (got %d)", s->dim_xyz);
  DN_REQUIRE(s->dim_dir >= 1 && s->dim_dir <= 32, "tc: dim_dir must be in 1..32 (got %d)", s->dim_dir);
  Plan& P = *plan;
  const int H = s->hidden;
  P.kx = pad16(s->dim_xyz);
  P.kd = pad16(s->dim_dir);
  int bias = 0, op = 0;
  auto add = [&](int k_main, int src, int k_smem, int n_out, int relu, int head, int prog_op) {
    HostLayer& L = P.layers[P.n_layers++];
    L.tc.k_main = k_main; L.tc.smem_src = src; L.tc.k_smem = k_smem; L.tc.n_out = n_out;
    L.tc.n_pass = (n_out + 127) / 128; L.tc.relu = relu; L.tc.head = head; L.tc.bias_off = bias;
    L.prog_op = prog_op;
    bias += (n_out + 127) / 128 * 128;
    const int np = n_out < 128 ? n_out : 128;
    P.weight_bytes += (int64_t)L.tc.n_pass * np * (k_main + (src ? k_smem : 0)) * 2;
  };
  add(0, 1, P.kx, H, 0, 0, op++);                                    // layer1 (no ReLU)
  for (int i = 0; i < s->n_trunk; ++i) {
    const bool skip = (i % s->skip_every == 0) && i > 0;
    add(H, skip ? 1 : 0, skip ? P.kx : 0, H, 1, i == s->n_trunk - 1 ? 1 : 0, op++);
  }
  P.op_alpha = op++;
  add(H, 0, 0, H, 1, 0, op++);                                        // fc_feat
  add(H, 2, P.kd, H / 2, 1, 2, op++);                                 // layers_dir[0] (+ fc_rgb head)
  P.op_rgb = op++;
  P.off_walpha = bias; bias += H;
  P.off_balpha = bias; bias += 4;
  P.off_wrgb = bias; bias += 3 * (H / 2);
  P.off_brgb = bias; bias += 4;
  P.n_const = bias;
  DN_REQUIRE(P.n_const <= kMaxConstFloats, "tc: const block too large");
  return 0;
}

static int64_t blob_bytes(const Plan& P) { return (int64_t)kMaxConstFloats * 4 + P.weight_bytes; }

struct PackChunk { int64_t dst; int w_off; int n_out; int n0; int np; int k0; int kc; int k_valid; };

__global__ void pack_weights_kernel(const float* __restrict__ params, const PackChunk* __restrict__ chunks,
                                    int n_chunks, uint8_t* __restrict__ blob) {
  for (int ci = blockIdx.y; ci < n_chunks; ci += gridDim.y) {
    const PackChunk c = chunks[ci];
    const int total = c.np * c.kc;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int k = e / c.np, n = e - k * c.np;          // n fastest: coalesced reads of Wt[k][n]
      const int ks = c.k0 + k;
      const float w = (k < c.k_valid) ? params[c.w_off + (int64_t)ks * c.n_out + c.n0 + n] : 0.0f;
      const int64_t off = c.dst + (int64_t)(k >> 3) * (c.np * 16) + n * 16 + (k & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(blob + off) = __float2bfloat16_rn(w);
    }
  }
}

__global__ void pack_consts_kernel(const float* __restrict__ params, const int4* __restrict__ moves, int n_moves,
                                   float* __restrict__ consts) {
  // moves: (dst, src, count, transpose_cols) ; transpose_cols > 0: src is Wt[count/cols... ] see host
  for (int m = blockIdx.x; m < n_moves; m += gridDim.x) {
    const int4 mv = moves[m];
    for (int i = threadIdx.x; i < mv.z; i += blockDim.x) {
      int src = mv.y + i;
      if (mv.w > 0) {               // W_rgb: consts[c*hw + k] = Wt[k*3 + c]
        const int hw = mv.w, cch = i / hw, k = i - cch * hw;
        src = mv.y + k * 3 + cch;
      }
      consts[mv.x + i] = params[src];
    }
  }
}

}  // namespace tc
}  // namespace dexnerf

using namespace dexnerf;
using namespace dexnerf::tc;

extern "C" DEXNERF_API int64_t dexnerf_tc_packed_bytes(const dexnerf_flexible_spec* spec) {
  Plan plan;
  if (make_plan(spec, &plan)) return -1;
  return blob_bytes(plan);
}

extern "C" DEXNERF_API int dexnerf_tc_pack(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                                           const float* params, void* packed, void* workspace, void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(prog && params && packed && workspace, "tc_pack: null pointer");
  DN_REQUIRE(prog->n_ops == plan.n_layers + 2, "tc_pack: program has %d ops, expected %d", prog->n_ops, plan.n_layers + 2);
  cudaStream_t st = (cudaStream_t)stream;
  // chunk table (host) -> workspace (device)
  static thread_local PackChunk h_chunks[256];
  static thread_local int4 h_moves[32];
  int nch = 0, nmv = 0;
  int64_t dst = (int64_t)kMaxConstFloats * 4;
  for (int l = 0; l < plan.n_layers; ++l) {
    const TcLayer& L = plan.layers[l].tc;
    const dexnerf_op& op = prog->ops[plan.layers[l].prog_op];
    const int real_smem = L.smem_src == 1 ? spec->dim_xyz : (L.smem_src == 2 ? spec->dim_dir : 0);
    DN_REQUIRE(op.out_dim == L.n_out && op.src0_dim + op.src1_dim == L.k_main + real_smem,
               "tc_pack: program op %d does not match the Flexible layer table", plan.layers[l].prog_op);
    const int np = L.n_out < 128 ? L.n_out : 128;
    for (int p = 0; p < L.n_pass; ++p) {
      for (int c = 0; c < L.k_main / 64 + (L.smem_src ? 1 : 0); ++c) {
        DN_REQUIRE(nch < 256, "tc_pack: too many chunks");
        PackChunk& pc = h_chunks[nch++];
        const bool main = c < L.k_main / 64;
        pc.dst = dst; pc.w_off = (int)op.w_off; pc.n_out = L.n_out; pc.n0 = p * 128; pc.np = np;
        // layer1: the smem operand is the whole input; skip/dir layers: it follows the k_main hidden inputs
        pc.k0 = main ? c * 64 : L.k_main;
        pc.kc = main ? 64 : L.k_smem;
        pc.k_valid = main ? 64 : real_smem;
        dst += (int64_t)np * pc.kc * 2;
      }
    }
    h_moves[nmv++] = make_int4(L.bias_off, (int)op.b_off, L.n_out, 0);
  }
  const int H = spec->hidden;
  const dexnerf_op& oa = prog->ops[plan.op_alpha];
  const dexnerf_op& orgb = prog->ops[plan.op_rgb];
  DN_REQUIRE(oa.out_dim == 1 && oa.src0_dim == H && orgb.out_dim == 3 && orgb.src0_dim == H / 2,
             "tc_pack: head ops do not match");
  h_moves[nmv++] = make_int4(plan.off_walpha, (int)oa.w_off, H, 0);        // Wt[k][0]
  h_moves[nmv++] = make_int4(plan.off_balpha, (int)oa.b_off, 1, 0);
  h_moves[nmv++] = make_int4(plan.off_wrgb, (int)orgb.w_off, 3 * (H / 2), H / 2);
  h_moves[nmv++] = make_int4(plan.off_brgb, (int)orgb.b_off, 3, 0);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  DN_CUDA(cudaMemcpyAsync(ws, h_chunks, sizeof(PackChunk) * nch, cudaMemcpyHostToDevice, st));
  DN_CUDA(cudaMemcpyAsync(ws + sizeof(PackChunk) * 256, h_moves, sizeof(int4) * nmv, cudaMemcpyHostToDevice, st));
  DN_CUDA(cudaStreamSynchronize(st));   // h_* are reused by the next call
  DN_CUDA(cudaMemsetAsync(packed, 0, (size_t)kMaxConstFloats * 4, st));
  pack_weights_kernel<<<dim3(8, nch), 256, 0, st>>>(params, reinterpret_cast<const PackChunk*>(ws), nch,
                                                    reinterpret_cast<uint8_t*>(packed));
  DN_CHECK_LAUNCH("pack_weights");
  pack_consts_kernel<<<nmv, 128, 0, st>>>(params, reinterpret_cast<const int4*>(ws + sizeof(PackChunk) * 256), nmv,
                                          reinterpret_cast<float*>(packed));
  DN_CHECK_LAUNCH("pack_consts");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_tc_query(const dexnerf_flexible_spec* spec, const void* packed, const float* ro,
                                            const float* rd, const float* viewdirs, const float* z, int64_t n,
                                            int S, float* rf, float* dbg, int dbg_layer, int dbg_pass,
                                            void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(packed && ro && rd && viewdirs && z && rf, "tc_query: null pointer");
  DN_REQUIRE(S >= 1, "tc_query: S < 1");
  DN_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(rf) & 15) == 0,
             "tc_query: packed weights and rf must be 16-byte aligned");
  if (n <= 0) return 0;
  TcParams P{};
  P.consts = reinterpret_cast<const float*>(packed);
  P.weights = reinterpret_cast<const uint8_t*>(packed) + (size_t)kMaxConstFloats * 4;
  P.ro = ro; P.rd = rd; P.vd = viewdirs; P.z = z; P.rf = rf; P.dbg = dbg;
  P.m_total = n * (int64_t)S; P.S = S;
  P.n_layers = plan.n_layers; P.hidden = spec->hidden; P.n_const = plan.n_const;
  P.off_walpha = plan.off_walpha; P.off_balpha = plan.off_balpha; P.off_wrgb = plan.off_wrgb; P.off_brgb = plan.off_brgb;
  P.Lx = spec->Lx; P.Ld = spec->Ld; P.include_xyz = spec->include_xyz; P.include_dir = spec->include_dir;
  P.log_xyz = spec->log_xyz; P.log_dir = spec->log_dir; P.dim_xyz = spec->dim_xyz; P.dim_dir = spec->dim_dir;
  P.dbg_layer = dbg_layer; P.dbg_pass = dbg_pass;
  for (int l = 0; l < plan.n_layers; ++l) {
    P.layers[l] = plan.layers[l].tc;
    if (plan.layers[l].tc.smem_src == 1) P.last_xyz_layer = l;
  }
  const int64_t n_tiles = (P.m_total + kTileM - 1) / kTileM, n_pairs = (n_tiles + 1) / 2;
  const int grid = (int)(n_pairs < kNumSMs ? n_pairs : kNumSMs);
  const size_t smem = Smem::total + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kernel) -> int {
    DN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kThreads, smem, st>>>(P);
    return 0;
  };
  int rc;
  if (spec->hidden == 256) rc = dbg ? launch(mlp_tc_kernel<256, true>) : launch(mlp_tc_kernel<256, false>);
  else rc = dbg ? launch(mlp_tc_kernel<128, true>) : launch(mlp_tc_kernel<128, false>);
  if (rc) return rc;
  DN_CHECK_LAUNCH("mlp_tc");
  return 0;
}
