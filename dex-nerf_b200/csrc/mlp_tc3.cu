// tcgen05 evaluation of hidden-128 FlexibleNeRFModels (nerf/models.py:185-256 - the width every script of the
// reference instantiates and all shipped checkpoints have) with THREE 128-sample tiles in
// flight per SM.  Inference only; same arithmetic contract, weight images, layer table and encoders as mlp_tc.cu
// (one difference: the layer bias enters the accumulator as an MMA, see below).
//
// Why a second kernel.  At hidden 128 an MMA pass of a tile is 8 instructions (~600 cycles), its epilogue ~1 100, and
// the hand-offs between them another ~500, so with two tiles in flight (mlp_tc.cu) the tensor pipe waits for the
// MMA -> epilogue -> MMA chain of a tile two thirds of the time (tensor pipe active 33 - 45 %).  A tile needs 64 TMEM
// columns for its bf16 A operand and 128 for the fp32 accumulator; three private accumulators do not fit next to
// three A operands (576 > 512 columns).  But an accumulator is only occupied from the first MMA of a pass until the
// epilogue has LOADED it, so here the three tiles share TWO accumulators:
//
//   TMEM   columns [0, 192): A operand of tile 0 / 1 / 2;  [256, 384): accumulator D0;  [384, 512): accumulator D1
//   pass k (k = 0, 1, 2, ... over (tile group, layer, tile)) computes tile k % 3 into accumulator k & 1
//
//   warp 0        weight producer: 6-slot ring of 16 KB UMMA images, a chunk is consumed by the three tiles
//   warps 1-3     MMA issuers, one per tile (warp 2 is also the TMEM allocator): each does the waiting for its own
//                 pass (weights, encodings, then the gate: accumulator drained by the epilogue of pass k - 2, A operand
//                 stored by the tile's own previous epilogue) and then takes its TURN from the issuer of pass k - 1,
//                 so that the passes enter the tensor pipe in k order
//   warps 4-11    epilogue team 0: serves accumulator D0, i.e. the even passes, whichever tile they belong to
//   warps 12-19   epilogue team 1: serves accumulator D1 (odd passes)
//                 (8 warps per pass as in mlp_tc.cu: two per 32-lane quarter, 64 columns each; consecutive layers
//                 of a tile alternate between the teams, so the one per-row value that outlives a layer - the
//                 partial sigma dot product - travels through shared memory)
//   warps 20-23   encoders for the NEXT group of three tiles
//
// The layer biases are MMA operands: D = ones[128 x 16] x B_l[n_out x 16]^T is the first MMA of every pass, with
// B_l[n][0:2] = (bf16(b), bf16(b - bf16(b))) - both products are exact in the fp32 accumulator, the bias loses less than
// 2^-17 of its value - so the epilogue is load, ReLU + pack, store (no 16 shared-memory loads and 32 packed adds per warp
// and pass).  One K group of 8 per operand (2 KB); the second K group of both is a shared block of zeros that the
// descriptor's leading-dimension offset points to.
//
// Measured on a B200 (tools/mlp_sweep.py, 640 000 x 192 samples): 8 x 128 skip 3 36.7 ms = 1 108 TFLOP/s (pair kernel
// 39.7 - 40.4), 4 x 128 23.6 ms = 875 TFLOP/s (pair kernel 23.7 - 24.4).  What bounds it now (tools/tc3_timeline.py): two
// passes on the same accumulator are issue (~700 cycles: 9 MMAs) + completion seen by the epilogue (~290) + drain of the
// accumulator (~850) + gate -> turn -> first MMA (~250) apart, i.e. ~1 050 cycles per pass against 576 of tensor work,
// and the SM's issue slots are nearly all taken (ncu: 2.1 of 4 instructions per cycle with every warp a dependent
// chain).  Tried on top and NOT kept: one issuer + scout (the control warp itself needs ~1 000 cycles per pass), two
// issuers alternating accumulators (same), a hand-written tight barrier-poll loop (more polls per second, everything
// slower), the whole 64-column share in registers for an early release (needs 88 registers; setmaxnreg is accepted but
// ptxas keeps allocating 80 and spills: 40.9 ms), all sixteen epilogue warps on every pass with 32 columns each
// (shorter drain on paper, but twice the per-pass overhead instructions: 41.5 ms), one polling warp per epilogue team
// and per encoder group with the others blocked on a named barrier (bar.sync costs no issue slots; no gain: 38.4 / 23.5
// ms - the spinning warps take slots nobody else wanted, the limit is the latency of the dependent chains), and
// HALF passes: every layer as two N = 64 passes into five rotating 64-column accumulators (192 + 320 = 512 columns: five
// half passes between two uses of an accumulator instead of two full ones).  Built and run: 50.3 / 29.6 ms - a TMEM-A
// tcgen05.mma costs ~65-75 cycles whether N is 64 or 128 (the 128 x 16 A operand is read per instruction), so halving N
// doubles the tensor time.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "mlp_tc_shared.cuh"

namespace dexnerf {
namespace tc {
namespace three {

constexpr int kNT = 3;                 // tiles in flight
constexpr int kSlots = 6;              // weight ring
constexpr int kTeamThreads = 256;      // 8 warps per epilogue team

constexpr int kHeadFloats = 1024;      // fc_alpha / fc_rgb / fc_out weights and biases (<= 4 x 128 + 4 floats)
struct Smem {
  static constexpr int w_slots = 0;
  static constexpr int pe_xyz = w_slots + kSlots * kSlotBytes;           // [tile]
  static constexpr int pe_dir = pe_xyz + kNT * kPeXyzBytes;              // [tile]
  static constexpr int heads = pe_dir + kNT * kPeDirBytes;               // the const block from off_walpha on
  static constexpr int sig = heads + kHeadFloats * 4;                    // [tile][half][row] partial sigma
  static constexpr int xchg = sig + kNT * 2 * kTileM * 4;                // [tile][row] float4 head partials of half 1
  static constexpr int bias = xchg + kNT * kTileM * 16;                  // [layer] bias images (B operands, K group 0)
  static constexpr int ones = bias + kMaxLayers * kBiasImgBytes;         // A operand of the bias MMA (K group 0)
  static constexpr int zero = ones + kBiasImgBytes;                      // K group 1 of both: zeros
  static constexpr int bars = zero + kBiasImgBytes;
  static constexpr int n_bars = 2 * kSlots + 6 * kNT + 2 + 2 * kNT;
  static constexpr int tmem_ptr = bars + n_bars * 8;
  static constexpr int total = tmem_ptr + 16;
};
static_assert(Smem::total + 1024 <= 227 * 1024, "shared memory budget");
__device__ __forceinline__ int B_wfull(int s) { return s; }
__device__ __forceinline__ int B_wempty(int s) { return kSlots + s; }
__device__ __forceinline__ int B_xyzfull(int t) { return 2 * kSlots + t; }
__device__ __forceinline__ int B_xyzempty(int t) { return 2 * kSlots + kNT + t; }
__device__ __forceinline__ int B_dirfull(int t) { return 2 * kSlots + 2 * kNT + t; }
__device__ __forceinline__ int B_dirempty(int t) { return 2 * kSlots + 3 * kNT + t; }
__device__ __forceinline__ int B_aready(int t) { return 2 * kSlots + 4 * kNT + t; }
__device__ __forceinline__ int B_dfull(int b) { return 2 * kSlots + 5 * kNT + b; }
__device__ __forceinline__ int B_turn(int t) { return 2 * kSlots + 5 * kNT + 2 + t; }
__device__ __forceinline__ int B_dfree(int b, int t) { return 2 * kSlots + 6 * kNT + 2 + b * kNT + t; }   // drained by tile t's epilogue

constexpr int kH = 128;
#ifndef DEXNERF_TC3_TURN_FIRST
#define DEXNERF_TC3_TURN_FIRST 1
#endif
#ifndef DEXNERF_TC3_EPI
#define DEXNERF_TC3_EPI 2
#endif
constexpr int kEpi = DEXNERF_TC3_EPI;   // epilogue_pass_wide variant (2: three 16-column buffers; 1 needs 88 registers, see below)
// Bring-up aids (compile with -DDEXNERF_TC3_BRINGUP): a role mask (DEXNERF_TC3_STAGE), a pass limit
// (DEXNERF_TC3_PASSES) and a clock-stamped trace of CTA 0 (DEXNERF_TC3_TRACE = address of a device or pinned-host buffer -
// the latter survives a faulting kernel; DEXNERF_TC3_WINDOW = "lo,hi" passes): trace[role][0] = count, then pairs of
// (code << 16 | pass - lo, clock).  tools/tc3_timeline.py prints the timeline, tools/tc3_check.py the progress.
#ifdef DEXNERF_TC3_BRINGUP
constexpr bool kBringUp = true;
#else
constexpr bool kBringUp = false;
#endif
constexpr uint32_t kTmemA = 0, kTmemD = 256;   // A operands in [0, 192), the two accumulators at 256 and 384

__device__ __forceinline__ void trace(const TcParams& P, int role, uint32_t code, uint32_t kk, uint32_t& n) {
  if (!kBringUp) return;
  uint32_t* tr = reinterpret_cast<uint32_t*>(P.tape_mask[0]);
  const uint32_t lo = (uint32_t)P.tape_mask[1], hi = (uint32_t)P.tape_mask[2];     // window of passes to record
  if (!tr || blockIdx.x != 0 || (threadIdx.x & 31) != 0 || kk < lo || kk >= hi) return;
  uint32_t* r = tr + role * 512;      // the count lives in a register of the caller: stores only, no global loads
  if (n < 250) { r[1 + 2 * n] = (code << 16) | (kk - lo); r[2 + 2 * n] = (uint32_t)clock64(); r[0] = ++n; }
}

__global__ void __launch_bounds__(kThreads, 1) mlp_tc3_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + Smem::bars;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (P.m_total + kTileM - 1) / kTileM;
  const int64_t n_groups = (n_tiles + kNT - 1) / kNT;
  const int nl = P.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(bar(B_wfull(s)), 1); mbar_init(bar(B_wempty(s)), kNT); }
    for (int t = 0; t < kNT; ++t) {
      mbar_init(bar(B_xyzfull(t)), 128); mbar_init(bar(B_xyzempty(t)), 1);
      mbar_init(bar(B_dirfull(t)), 128); mbar_init(bar(B_dirempty(t)), 1);
      mbar_init(bar(B_aready(t)), kTeamThreads);
      mbar_init(bar(B_turn(t)), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(B_dfull(b)), 1);
      for (int t = 0; t < kNT; ++t) mbar_init(bar(B_dfree(b, t)), kTeamThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + Smem::tmem_ptr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  {
    // head weights (the tail of the const block); the layers' biases are MMA operands here:
    //   D = ones[128 x 16] * B_l[n_out x 16]^T as the first MMA of every pass, ones[:, 0:2] = 1, B_l[n][0:2] = (hi, lo) of
    //   the bias.  Both operands are ONE K group of 8 (2 KB: 16-byte rows) whose second K group is a shared block of zeros -
    //   the descriptor's leading-dimension offset points there.
    float* c = reinterpret_cast<float*>(smem + Smem::heads);
    for (int i = threadIdx.x; i < P.n_const - P.off_walpha; i += kThreads) c[i] = P.consts[P.off_walpha + i];
    uint4* b = reinterpret_cast<uint4*>(smem + Smem::bias);
    const uint4* src = reinterpret_cast<const uint4*>(P.bias_img);
    for (int i = threadIdx.x; i < nl * (kBiasImgBytes / 16); i += kThreads) b[i] = src[i];
    uint4* o = reinterpret_cast<uint4*>(smem + Smem::ones);
    for (int i = threadIdx.x; i < 2 * (kBiasImgBytes / 16); i += kThreads)      // ones, then zero
      o[i] = i < kBiasImgBytes / 16 ? make_uint4(0x3F803F80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + Smem::tmem_ptr);
  const float* s_head = reinterpret_cast<const float*>(smem + Smem::heads);   // index: const offset - off_walpha

  const int stage = kBringUp ? P.dbg_layer : 7;   // bring-up: which roles run (DEXNERF_TC3_STAGE)
  // Register re-balancing (DEXNERF_TC3_EPI = 1, an experiment that did not pay): the drain of a SHARED accumulator is on
  // the critical loop (issue -> complete -> drain -> next pass on the same accumulator), and with its whole 64-column
  // share in registers (one tcgen05.ld round trip) the epilogue could release it ~300 instead of ~850 cycles after the
  // pass completes.  That needs ~88 registers; setmaxnreg (control warpgroup - 32, epilogue warpgroups + 8) is accepted,
  // but this ptxas keeps allocating at most the launch bound's 80 and spills instead: 37.5 -> 40.9 ms at 8 x 128.
  constexpr bool kRebalance = DEXNERF_TC3_EPI == 1;
  if (warp == 0) {
    // =============================== weight producer ===============================
    if (kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const bool leader = elect_one();
    uint32_t cnt = 0, tn = 0, pk0 = 0;        // pk0: first pass of the group (bring-up trace only)
#pragma unroll 1
    for (int64_t grp = blockIdx.x; grp < n_groups && stage >= 2; grp += gridDim.x, pk0 += (uint32_t)(nl * kNT)) {
      const uint8_t* src = P.weights;
#pragma unroll 1
      for (int l = 0; l < nl; ++l) {
        const TcLayer& L = P.layers[l];
        const int np = L.n_out < 128 ? L.n_out : 128;
        const int nc = chunks_in_pass(L);
#pragma unroll 1
        for (int c = 0; c < nc; ++c, ++cnt) {
          const int kc = (c < L.k_main / 64) ? 64 : L.k_smem;
          const uint32_t bytes = (uint32_t)(np * kc * 2);
          const uint32_t slot = cnt % kSlots, ph = (cnt / kSlots) & 1;
          trace(P, 0, (uint32_t)(c * 2), pk0 + (uint32_t)(l * kNT), tn);
          mbar_wait(bar(B_wempty(slot)), ph ^ 1, 30);
          trace(P, 0, (uint32_t)(c * 2 + 1), pk0 + (uint32_t)(l * kNT), tn);
          if (leader) {
            mbar_arrive_expect_tx(bar(B_wfull(slot)), bytes);
            bulk_g2s(sbase + Smem::w_slots + slot * kSlotBytes, src, bytes, bar(B_wfull(slot)));
          }
          __syncwarp();
          src += bytes;
        }
      }
    }
  } else if (warp < 4) {
    // =============================== MMA issuers: one per tile ===============================
    // At hidden 128 a pass is 8 MMAs = 512 tensor-pipe cycles, and a control warp is ONE dependent instruction stream
    // that gets through a barrier poll or a dozen instructions in ~100 - 300 cycles (it shares its scheduler with four
    // epilogue warps and an encoder warp).  A single issuer (+ a scout doing its waiting, as in mlp_tc.cu) needs
    // ~1 000 - 1 200 cycles per pass for its polls, descriptor arithmetic and the issue itself - measured with the
    // clock-stamped trace below - and sets the pace, whatever the epilogues do.  So every tile has its OWN issuer
    // (warps 1, 2, 3; warp 2 is also the TMEM allocator), which does all the waiting for its pass - the layer's weight
    // chunks, the tile's encodings, then the gate: accumulator drained by the epilogue of pass kk - 2 and A operand
    // stored by the tile's previous epilogue, ONE double poll - while the other two tiles' passes are being issued,
    // and then waits for its TURN: one arrival from the issuer of pass kk - 1 once that pass is completely issued, so
    // that the passes enter the tensor pipe strictly in order (interleaved, two passes would both finish late, and the
    // MMA -> epilogue -> MMA chain of a tile is what matters).  Power-of-two ring (slot = counter & 7), one elected
    // block per pass, straight-line code for plain H -> H layers.
    const int t = warp - 1;                   // my tile
    if (kRebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    const bool leader = elect_one();
    uint32_t tn = 0;
    uint32_t slot = 0, wph = 0;               // ring cursor of the current layer: slot and phase of its first chunk
    const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);   // SBO = 128 B, version 1
    const uint32_t slot0_lo = ((sbase + Smem::w_slots) >> 4) & 0x3FFF;
    const uint32_t wfull0 = bar(B_wfull(0)), wempty0 = bar(B_wempty(0));
    const uint32_t a_tmem = tmem_base + kTmemA + (uint32_t)(t * (kH / 2));
    const uint32_t my_turn = bar(B_turn(t)), next_turn = bar(B_turn(t + 1 == kNT ? 0 : t + 1));
    const uint32_t aready_bar = bar(B_aready(t));
    // bias MMA operands: K group 0 = the 2 KB image, K group 1 = the zero block (leading-dimension offset = distance)
    const uint64_t ones_desc = desc_hi | (uint64_t)((((sbase + Smem::ones) >> 4) & 0x3FFF) |
                                                    ((uint32_t)((Smem::zero - Smem::ones) >> 4) << 16));
    const int tp = t + 1 == kNT ? 0 : t + 1;  // the tile of pass kk - 2
    uint32_t kk = (uint32_t)t, j = 0, it = 0;  // pass counter, the tile's pass index (kk = 3 j + t), group index
    TcLayer L = P.layers[0];
#pragma unroll 1
    for (int64_t grp = blockIdx.x; grp < n_groups && stage >= 4; grp += gridDim.x, ++it) {
#pragma unroll 1
      for (int l = 0; l < nl; ++l, kk += kNT, ++j) {
        if (kBringUp && P.dbg_pass && kk >= (uint32_t)P.dbg_pass) break;
        const TcLayer Lnext = P.layers[l + 1 < nl ? l + 1 : 0];
        const uint32_t b = kk & 1;
        const uint32_t d_tmem = tmem_base + kTmemD + b * 128;
        const int n_chunks = chunks_in_pass(L);
        const uint64_t bias_desc = desc_hi | (uint64_t)((((sbase + Smem::bias + (uint32_t)l * kBiasImgBytes) >> 4) & 0x3FFF) |
                                                        ((uint32_t)((Smem::zero - Smem::bias - l * kBiasImgBytes) >> 4) << 16));
        trace(P, 1 + t, 0, kk, tn);
        // what the pass needs besides its gate (nearly always there already)
        {
          uint32_t sl = slot, ph = wph;
          for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(wfull0 + 8 * sl, ph, 32);
            if (++sl == kSlots) { sl = 0; ph ^= 1; }
          }
        }
        if (L.smem_src == 2) mbar_wait(bar(B_dirfull(t)), it & 1, 33);
        if (l == 0) mbar_wait(bar(B_xyzfull(t)), it & 1, 34);
        trace(P, 1 + t, 4, kk, tn);
        // the gate
        // (accumulator b was last used by pass kk - 2 = pass jp of tile tp; d_free is kept per (accumulator, tile): a
        // parity wait must never be more than one phase ahead of its barrier, and with one barrier per accumulator
        // an issuer polling early could still be TWO drains ahead - a false "complete", found the hard way)
        const uint32_t jp = (t == kNT - 1) ? j : j - 1;
        const uint32_t dfree_bar = bar(B_dfree((int)b, tp));
#if DEXNERF_TC3_TURN_FIRST
        // my turn first (pass kk - 1 is issued one pass before the accumulator of pass kk - 2 is drained, so it is
        // normally there), the gate last: nothing but the fence stands between the drain and the first MMA
        if (kk) mbar_wait(my_turn, (t ? j : j - 1) & 1, 40);
        trace(P, 1 + t, 2, kk, tn);
        if (j) mbar_wait2(dfree_bar, (jp >> 1) & 1, aready_bar, (j - 1) & 1, 35);
        else if (kk >= 2) mbar_wait(dfree_bar, (jp >> 1) & 1, 35);
#else
        if (j) mbar_wait2(dfree_bar, (jp >> 1) & 1, aready_bar, (j - 1) & 1, 35);
        else if (kk >= 2) mbar_wait(dfree_bar, (jp >> 1) & 1, 35);
        trace(P, 1 + t, 2, kk, tn);
        // my turn: pass kk - 1 has been issued
        if (kk) mbar_wait(my_turn, (t ? j : j - 1) & 1, 40);
#endif
        tc_fence_after();
        trace(P, 1 + t, 3, kk, tn);
        if (L.k_main == kH && L.smem_src == 0 && L.n_out == kH) {
          // ---- plain hidden layer: two 64-wide K chunks from the TMEM-resident activations
          const uint32_t s0 = slot, s1 = slot + 1 == kSlots ? 0 : slot + 1;
          const uint32_t b0 = (slot0_lo + s0 * (kSlotBytes >> 4)) | (128u << 16);
          const uint32_t b1 = (slot0_lo + s1 * (kSlotBytes >> 4)) | (128u << 16);
          const uint32_t idesc = instr_desc(128);
          if (leader) {
            mma_ss(d_tmem, ones_desc, bias_desc, idesc, 0u);            // D = bias
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(d_tmem, a_tmem + (uint32_t)(ks * 8), desc_hi | (uint64_t)(b0 + (uint32_t)ks * 256), idesc, 1u);
            tc_commit(wempty0 + 8 * s0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(d_tmem, a_tmem + (uint32_t)(32 + ks * 8), desc_hi | (uint64_t)(b1 + (uint32_t)ks * 256), idesc, 1u);
            mbar_arrive(next_turn);
            tc_commit(wempty0 + 8 * s1);
            tc_commit(bar(B_dfull(b)));
          }
          __syncwarp();
        } else {
          // ---- layer with a shared-memory operand (layer 1, skip layers, the view-direction layer)
          const int np = L.n_out < 128 ? L.n_out : 128;
          const uint32_t idesc = instr_desc(np);
          const int n_main = L.k_main / 64;
          const uint32_t b_lbo16 = (uint32_t)np;
          if (leader) {
            uint32_t sl = slot;
            mma_ss(d_tmem, ones_desc, bias_desc, idesc, 0u);            // D = bias
            for (int m = 0; m < n_main; ++m) {
              const uint32_t b_lo = (slot0_lo + sl * (kSlotBytes >> 4)) | (b_lbo16 << 16);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_ts(d_tmem, a_tmem + (uint32_t)(m * 32 + ks * 8),
                       desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 2 * b_lbo16), idesc, 1u);
              tc_commit(wempty0 + 8 * sl);
              if (++sl == kSlots) sl = 0;
            }
            if (L.smem_src) {
              const uint32_t b_lo = (slot0_lo + sl * (kSlotBytes >> 4)) | (b_lbo16 << 16);
              const uint32_t a_addr = (L.smem_src == 1)
                  ? sbase + Smem::pe_xyz + (uint32_t)t * kPeXyzBytes
                  : sbase + Smem::pe_dir + (uint32_t)t * kPeDirBytes;
              const uint32_t a_lo = ((a_addr >> 4) & 0x3FFF) | ((uint32_t)kTileM << 16);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                if (ks * 16 < L.k_smem)
                  mma_ss(d_tmem, desc_hi | (uint64_t)(a_lo + (uint32_t)ks * 2 * kTileM),
                         desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 2 * b_lbo16), idesc, 1u);
              }
              tc_commit(wempty0 + 8 * sl);
            }
            mbar_arrive(next_turn);
            tc_commit(bar(B_dfull(b)));
            if (l == P.last_xyz_layer) tc_commit(bar(B_xyzempty(t)));
            if (L.smem_src == 2) tc_commit(bar(B_dirempty(t)));
          }
          __syncwarp();
        }
        trace(P, 1 + t, 1, kk, tn);
        slot += (uint32_t)n_chunks;
        if (slot >= kSlots) { slot -= kSlots; wph ^= 1; }
        L = Lnext;
      }
    }
  } else if (warp >= 4 && warp < 20) {
    // =============================== epilogue teams ===============================
    if (kRebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
    const int e = warp - 4;
    const int team = e >> 3, hs = (e >> 2) & 1, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t d_base = tmem_base + kTmemD + (uint32_t)(team * 128) + lane_base;
    const int pair_bar = 1 + team * 4 + q;          // named barrier of the two warps of this quarter
    float* s_sig = reinterpret_cast<float*>(smem + Smem::sig);
    float4* s_xchg = reinterpret_cast<float4*>(smem + Smem::xchg);
    const int hw = kH / 2;
    uint32_t ph_dfull = 0;
    // pass kk = team, team + 2, ... decoded incrementally into (group, layer, tile)
    int t = team, l = 0;
    int64_t grp = blockIdx.x;
    uint32_t ekk = (uint32_t)team, tn = 0;    // bring-up trace only
#pragma unroll 1
    while (grp < n_groups && stage >= 5) {
      const TcLayer& L = P.layers[l];
      const int64_t g = (grp * kNT + t) * kTileM + row;
      const uint32_t a_tmem = tmem_base + kTmemA + (uint32_t)(t * (kH / 2)) + lane_base;
      trace(P, 4 + (warp - 4), 0, ekk, tn);
      const uint32_t dfree = bar(B_dfree(team, t));
      mbar_wait(bar(B_dfull(team)), ph_dfull, 37);
      ph_dfull ^= 1;
      tc_fence_after();
      trace(P, 4 + (warp - 4), 1, ekk, tn);
      if (stage < 6 || (stage < 7 && L.head >= 2)) {     // bring-up: protocol only
        tc_fence_before();
        mbar_arrive(dfree);
        mbar_arrive(bar(B_aready(t)));
      } else if (L.head == 2) {
        // ---- last layer: ReLU, fc_rgb on the CUDA cores over this warp's 32 of the 64 outputs, final store
        float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
        const uint32_t col0 = (uint32_t)(hs * (hw / 2));
        const uint32_t wr = sbase + Smem::heads + ((uint32_t)(P.off_wrgb - P.off_walpha) + col0) * 4;
        uint32_t v[2][16];
        tmem_ld16_issue(d_base + col0, v[0]);
        tmem_ld16_issue(d_base + col0 + 16, v[1]);
        tmem_ld16_wait(v[0]);
        tmem_ld16_tie(v[1]);
        tc_fence_before();
        mbar_arrive(dfree);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 r4 = lds128(wr + (uint32_t)((c * 16 + i) * 4));
            const float4 g4 = lds128(wr + (uint32_t)((hw + c * 16 + i) * 4));
            const float4 u4 = lds128(wr + (uint32_t)((2 * hw + c * 16 + i) * 4));
            const float x0 = fmaxf(__uint_as_float(v[c][i]), 0.0f);          // the bias is in the accumulator
            const float x1 = fmaxf(__uint_as_float(v[c][i + 1]), 0.0f);
            const float x2 = fmaxf(__uint_as_float(v[c][i + 2]), 0.0f);
            const float x3 = fmaxf(__uint_as_float(v[c][i + 3]), 0.0f);
            rgb0 = fmaf(x0, r4.x, fmaf(x1, r4.y, fmaf(x2, r4.z, fmaf(x3, r4.w, rgb0))));
            rgb1 = fmaf(x0, g4.x, fmaf(x1, g4.y, fmaf(x2, g4.z, fmaf(x3, g4.w, rgb1))));
            rgb2 = fmaf(x0, u4.x, fmaf(x1, u4.y, fmaf(x2, u4.z, fmaf(x3, u4.w, rgb2))));
          }
        }
        if (hs == 1) s_xchg[t * kTileM + row] = make_float4(rgb0, rgb1, rgb2, 0.f);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        if (hs == 0 && g < P.m_total) {
          const float4 o2 = s_xchg[t * kTileM + row];
          const float* br = s_head + (P.off_brgb - P.off_walpha);
          const float sg = s_sig[(t * 2) * kTileM + row] + s_sig[(t * 2 + 1) * kTileM + row];
          float4 o;
          o.x = rgb0 + o2.x + br[0]; o.y = rgb1 + o2.y + br[1]; o.z = rgb2 + o2.z + br[2];
          o.w = sg + s_head[P.off_balpha - P.off_walpha];
          reinterpret_cast<float4*>(P.rf)[g] = o;
        }
        mbar_arrive(bar(B_aready(t)));            // the row is finished: the tile may start its next group
      } else if (L.head == 3) {
        // ---- last trunk layer of a model without view directions: ReLU, fc_out (4 outputs) on the CUDA cores
        float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
        const uint32_t col0 = (uint32_t)(hs * 64);
        const uint32_t wo = sbase + Smem::heads + ((uint32_t)(P.off_wrgb - P.off_walpha) + col0) * 4;
        uint32_t v[2][16];
        tmem_ld16_issue(d_base + col0, v[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld16_wait(v[c & 1]);
          if (c + 1 < 4) tmem_ld16_issue(d_base + col0 + (uint32_t)((c + 1) * 16), v[(c + 1) & 1]);
          if (c == 3) {
            tc_fence_before();
            mbar_arrive(dfree);
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w0 = lds128(wo + (uint32_t)((c * 16 + i) * 4));
            const float4 w1 = lds128(wo + (uint32_t)((kH + c * 16 + i) * 4));
            const float4 w2 = lds128(wo + (uint32_t)((2 * kH + c * 16 + i) * 4));
            const float4 w3 = lds128(wo + (uint32_t)((3 * kH + c * 16 + i) * 4));
            const float x0 = fmaxf(__uint_as_float(v[c & 1][i]), 0.0f);
            const float x1 = fmaxf(__uint_as_float(v[c & 1][i + 1]), 0.0f);
            const float x2 = fmaxf(__uint_as_float(v[c & 1][i + 2]), 0.0f);
            const float x3 = fmaxf(__uint_as_float(v[c & 1][i + 3]), 0.0f);
            o0 = fmaf(x0, w0.x, fmaf(x1, w0.y, fmaf(x2, w0.z, fmaf(x3, w0.w, o0))));
            o1 = fmaf(x0, w1.x, fmaf(x1, w1.y, fmaf(x2, w1.z, fmaf(x3, w1.w, o1))));
            o2 = fmaf(x0, w2.x, fmaf(x1, w2.y, fmaf(x2, w2.z, fmaf(x3, w2.w, o2))));
            o3 = fmaf(x0, w3.x, fmaf(x1, w3.y, fmaf(x2, w3.z, fmaf(x3, w3.w, o3))));
          }
        }
        if (hs == 1) s_xchg[t * kTileM + row] = make_float4(o0, o1, o2, o3);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
        if (hs == 0 && g < P.m_total) {
          const float4 p2 = s_xchg[t * kTileM + row];
          const float* bo = s_head + (P.off_brgb - P.off_walpha);
          float4 o;
          o.x = o0 + p2.x + bo[0]; o.y = o1 + p2.y + bo[1]; o.z = o2 + p2.z + bo[2]; o.w = o3 + p2.w + bo[3];
          reinterpret_cast<float4*>(P.rf)[g] = o;
        }
        mbar_arrive(bar(B_aready(t)));
      } else {
        const uint32_t bias = 0;                // unused: the bias is in the accumulator
        const uint32_t wa = sbase + Smem::heads + (uint32_t)(hs * 64) * 4;
        const uint32_t d_tmem = d_base + (uint32_t)(hs * 64);
        const uint32_t a_store = a_tmem + (uint32_t)(hs * 32);
        float sigma = 0.0f;
        if (!L.relu) epilogue_pass_wide<false, false, false, false, false, kEpi>(d_tmem, a_store, bias, wa, sigma, dfree, nullptr, nullptr, nullptr);
        else if (L.head != 1) epilogue_pass_wide<true, false, false, false, false, kEpi>(d_tmem, a_store, bias, wa, sigma, dfree, nullptr, nullptr, nullptr);
        else {
          epilogue_pass_wide<true, true, false, false, false, kEpi>(d_tmem, a_store, bias, wa, sigma, dfree, nullptr, nullptr, nullptr);
          s_sig[(t * 2 + hs) * kTileM + row] = sigma;     // read by the final layer's epilogue (either team)
        }
        trace(P, 4 + (warp - 4), 2, ekk, tn);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(B_aready(t)));
      }
      trace(P, 4 + (warp - 4), 3, ekk, tn);
      if (kBringUp) ekk += 2;
      // next pass of this team
      t += 2;
      if (t >= kNT) {
        t -= kNT;
        if (++l == nl) { l = 0; grp += gridDim.x; }
      }
    }
  } else if (warp >= 20) {
    // =============================== encoders ===============================
    const int row = (warp - 20) * 32 + lane;
    const bool std_xyz = P.include_xyz && P.log_xyz, std_dir = P.include_dir && P.log_dir;
    uint32_t it = 0;
    for (int64_t grp = blockIdx.x; grp < n_groups && stage >= 3; grp += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      for (int t = 0; t < kNT; ++t) {
        const int64_t g = (grp * kNT + t) * kTileM + row;
        float pt[3] = {0.f, 0.f, 0.f};
        const bool valid = g < P.m_total;
        if (valid) {
          const int64_t ray = g / P.S;
          const float zz = P.z[g];
#pragma unroll
          for (int a = 0; a < 3; ++a) pt[a] = __fadd_rn(P.ro[ray * 3 + a], __fmul_rn(P.rd[ray * 3 + a], zz));
        }
        mbar_wait(bar(B_xyzempty(t)), ph ^ 1, 38);
        uint8_t* xyz = smem + Smem::pe_xyz + t * kPeXyzBytes + row * 16;
        if (std_xyz) {
          encode_row_std<8>(pt, P.dim_xyz, valid, xyz, nullptr);
        } else {
          for (int k8 = 0; k8 < 8; ++k8) {
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
        fence_proxy_async();
        mbar_arrive(bar(B_xyzfull(t)));
      }
      for (int t = 0; t < kNT && P.dim_dir > 0; ++t) {
        const int64_t g = (grp * kNT + t) * kTileM + row;
        float dir[3] = {0.f, 0.f, 0.f};
        const bool valid = g < P.m_total;
        if (valid) {
          const int64_t ray = g / P.S;
#pragma unroll
          for (int a = 0; a < 3; ++a) dir[a] = P.vd[ray * 3 + a];
        }
        mbar_wait(bar(B_dirempty(t)), ph ^ 1, 39);
        uint8_t* dr = smem + Smem::pe_dir + t * kPeDirBytes + row * 16;
        if (std_dir) {
          encode_row_std<4>(dir, P.dim_dir, valid, dr, nullptr);
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

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace three

int launch_mlp_tc3(const TcParams& P, int64_t n_tiles, cudaStream_t st) {
  const int64_t n_groups = (n_tiles + three::kNT - 1) / three::kNT;
  const int grid = (int)(n_groups < kNumSMs ? n_groups : kNumSMs);
  const size_t smem = three::Smem::total + 1024;
  DN_CUDA(cudaFuncSetAttribute(three::mlp_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TcParams Q = P;
  if (three::kBringUp) {
    const char* e = getenv("DEXNERF_TC3_STAGE");
    Q.dbg_layer = e ? atoi(e) : 7;
    const char* np = getenv("DEXNERF_TC3_PASSES");
    Q.dbg_pass = np ? atoi(np) : 0;
    const char* tr = getenv("DEXNERF_TC3_TRACE");
    Q.tape_mask[0] = tr ? (int64_t)strtoull(tr, nullptr, 0) : 0;
    const char* w = getenv("DEXNERF_TC3_WINDOW");        // "lo,hi": passes of CTA 0 to record
    int lo = 0, hi = 1 << 30;
    if (w) sscanf(w, "%d,%d", &lo, &hi);
    Q.tape_mask[1] = lo; Q.tape_mask[2] = hi;
  }
  three::mlp_tc3_kernel<<<grid, kThreads, smem, st>>>(Q);
  return 0;
}

}  // namespace tc
}  // namespace dexnerf
