// PTX wrappers shared by the tcgen05 kernels (mlp_tc.cu forward, mlp_tc_bwd.cu backward):
// mbarriers with bounded waits, bulk TMA copies, tcgen05 MMA / TMEM load-store, UMMA descriptors.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace dexnerf {
namespace tc {

constexpr int kTileM = 128;
constexpr uint32_t kSpinLimit = 1u << 22;   // ~ a second of polling, then trap

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
static __device__ __noinline__ void barrier_timeout(int who) {
  printf("dexnerf tcgen05 kernel: barrier timeout (wait site %d, block %d, thread %d)\n", who, blockIdx.x, threadIdx.x);
  __trap();
}
// DEXNERF_TIGHT_POLL = 1: the poll loop as hand-written PTX (four polls per bound check, 2.75 instructions per failed
// poll instead of the 9 of the compiled loop).  Measured on a B200 and NOT used: the faster loop polls more often, and
// every kernel got slower (8x256 render 108 -> 118.6 ms, C4 training 5.13 -> 5.38 ms) - waiting warps take issue slots
// and shared-memory bandwidth in proportion to their poll RATE, so the slow compiled loop is the better citizen.
#ifndef DEXNERF_TIGHT_POLL
#define DEXNERF_TIGHT_POLL 0
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
#if DEXNERF_TIGHT_POLL
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .u32 n;\n\t"
      "mov.u32 n, 0;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra LAB_DONE;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra LAB_DONE;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra LAB_DONE;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t@p bra LAB_DONE;\n\t"
      "add.u32 n, n, 4;\n\t"
      "setp.lt.u32 p, n, %3;\n\t"
      "@p bra LAB_WAIT;\n\t"
      "mov.u32 %0, 0;\n\tbra LAB_END;\n\t"
      "LAB_DONE:\n\tmov.u32 %0, 1;\n\t"
      "LAB_END:\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSpinLimit) : "memory");
  if (!ok) barrier_timeout(who);
#else
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) barrier_timeout(who);
  }
#endif
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
// Two barrier polls issued back to back (their latencies overlap); bounded sequential waits when either is pending.
__device__ __forceinline__ void mbar_wait2(uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, int who) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred q0, q1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
      "and.pred q0, q0, q1;\n\t"
      "selp.u32 %0, 1, 0, q0;\n\t}"
      : "=r"(ok) : "r"(b0), "r"(p0), "r"(b1), "r"(p1) : "memory");
  if (!ok) { mbar_wait(b0, p0, who); mbar_wait(b1, p1, who); }
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
// order every later use of `r` after the preceding (volatile) wait::ld without emitting an instruction: for a
// second buffer whose load was issued together with the first one's
__device__ __forceinline__ void tmem_ld32_tie(uint32_t (&r)[32]) {
  asm volatile(
      ""
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :: "memory");
}
__device__ __forceinline__ void tmem_ld16_tie(uint32_t (&r)[16]) {
  asm volatile(
      ""
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
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
// (a0, a1) + (b0, b1) as ONE packed fp32x2 add (sm_100 FADD2; same round-to-nearest results as two scalar
// adds): the epilogue warps share the SM's issue slots with everything else, and at hidden = 128 - where a
// pass is only 8 MMAs long - their instruction count is what the tensor pipe ends up waiting for.
__device__ __forceinline__ void add_f32x2(uint32_t a0, uint32_t a1, float b0, float b1, float& x0, float& x1) {
  uint64_t a, b, d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(d));
}
// two fp32 -> packed bf16x2 (lo = first K element), optional ReLU
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi, bool relu) {
  uint32_t d;
  if (relu) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// ReLU bits of eight packed bf16 pairs w[0..8) = 16 columns: column 2 j (the low half of w[j]) -> bit 7 - j, column
// 2 j + 1 (the high half) -> bit 15 - j; relu_bit_pos(i) is the position of column i.  min(x, 1) clamped at 0 on the
// signed halves (one VIMNMX.S16x2.RELU per pair, -0 included) is the flag of a pair, the flags of the eight pairs are
// shifted together and the two flag bytes compacted: 20 instructions per 16 columns, where testing and inserting the
// bits one by one took 40 - a third of all instructions of the forward kernel with tape (ncu source view, round 2).
__device__ __forceinline__ uint32_t relu_bits16(const uint32_t* w) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc = acc * 2u + __vimin_s16x2_relu(w[i], 0x00010001u);
  return __byte_perm(acc, 0u, 0x4420);
}
__host__ __device__ constexpr int relu_bit_pos(int i) { return (i & 1) * 8 + 7 - (i >> 1); }
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

}  // namespace tc
}  // namespace dexnerf
