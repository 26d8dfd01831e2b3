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
//   warp 1        MMA issuer: one elected thread issues tcgen05.mma (M=128, N=128, K=16), the
//                 passes of the two tiles alternating strictly; hidden H -> H layers run a pass
//                 specialised at compile time (nothing to compute between two passes).  Hidden activations are the A
//                 operand READ FROM TENSOR MEMORY (bf16, 2 per column); the positional encodings
//                 (layer 1, skip layer, view-direction layer) are A operands read from shared
//                 memory.  A 256-wide layer runs as two N=128 passes into a 128-column fp32
//                 accumulator, so a tile needs 128 (A) + 128 (D) TMEM columns and two tiles fill
//                 the 512 columns.
//   warp 2        TMEM allocator.
//   warp 3        scout: walks the pass sequence one step ahead of the issuer and does all the
//                 waiting for it - the weight chunks and the encodings first, the tile's epilogue
//                 (the gate on the critical path) last - then releases the pass with ONE arrival
//                 on ready[tile].
//   warps 4-19    epilogue, 8 warps per tile (two per 32-lane quarter, 64 columns each), one thread
//                 per sample row: tcgen05.ld the accumulator, + bias, ReLU, pack to bf16 and
//                 tcgen05.st it back as the next layer's A operand (first-pass results wait in
//                 registers until the layer's second pass has finished reading the old A).
//                 fc_alpha and fc_rgb (1 and 3 outputs) are fp32 dot products on the CUDA cores
//                 inside the epilogue; the only HBM write of the inference variant is the final
//                 (r,g,b,sigma) float4 per sample.  The training variant (kTape) also writes every
//                 layer's bf16 operand image and ReLU bits - the tape the backward consumes.
//   warps 20-23   encoders: one thread per sample row of the NEXT tile pair computes
//                 pts = ro + rd*z and the sin/cos encodings in registers (Cody-Waite reduction +
//                 MUFU) and writes them as bf16 UMMA core matrices into shared memory.
//
// Weight-image layout (no swizzle, K-major "interleaved" canonical layout): an operand tile is a
// grid of 8-row x 16-byte core matrices, each 128 contiguous bytes; byte offset of element
// (row r, k) = (k/8)*LBO + (r/8)*SBO + (r%8)*16 + (k%8)*2 with SBO = 128 and LBO = rows*16.
#include <type_traits>

#include "mlp_tc_shared.cuh"

namespace dexnerf {
namespace tc {

// ------------------------------------------------------------------ shared-memory map
struct Smem {
  // offsets from the 1024-aligned base
  static constexpr int w_slots = 0;
  static constexpr int pe_xyz = w_slots + kNumSlots * kSlotBytes;        // [tile]
  static constexpr int pe_dir = pe_xyz + 2 * kPeXyzBytes;                // [tile]
  static constexpr int consts = pe_dir + 2 * kPeDirBytes;
  static constexpr int xchg = consts + kMaxConstFloats * 4;             // [tile][row] float4 head partials
  static constexpr int bars = xchg + 2 * kTileM * 16;
  static constexpr int n_bars = 2 * kNumSlots + 4 + 4 + 2 + 2 + 2 + 2;
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
__device__ __forceinline__ int B_ready(int t) { return 2 * kNumSlots + 14 + t; }

// One epilogue pass of one warp over ITS 64 of the 128 accumulator columns of a tile (one thread per
// sample row; two warps per 32-row quarter split the columns): + bias, optional sigma head,
// optional ReLU, pack to bf16, then either keep the packed words in registers (first pass of a
// 256-wide layer) or store them as the next layer's A operand.
template <bool kRelu, bool kSig, bool kHold, bool kPark, bool kDbg, bool kTape>
__device__ __forceinline__ void epilogue_pass(uint32_t d_tmem, uint32_t a_park, uint32_t a_store, uint32_t bias,
                                              uint32_t wa, float& sigma, uint32_t (&held)[32],
                                              uint32_t dfree_bar, float* dbg_dst, long long* stamps,
                                              uint8_t* tape_row, uint2* tape_mask) {
  uint32_t mbits[2] = {0u, 0u};
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
      float x0, x1, x2, x3;
      add_f32x2(v[c & 1][i], v[c & 1][i + 1], b4.x, b4.y, x0, x1);
      add_f32x2(v[c & 1][i + 2], v[c & 1][i + 3], b4.z, b4.w, x2, x3);
      if (kSig) {   // fc_alpha on the unrounded layer output: rectified (Flexible: trunk output) or not (Paper: feat)
        const float4 w4 = lds128(wa + (uint32_t)((c * 16 + i) * 4));
        sigma = fmaf(kRelu ? fmaxf(x0, 0.0f) : x0, w4.x, sigma);
        sigma = fmaf(kRelu ? fmaxf(x1, 0.0f) : x1, w4.y, sigma);
        sigma = fmaf(kRelu ? fmaxf(x2, 0.0f) : x2, w4.z, sigma);
        sigma = fmaf(kRelu ? fmaxf(x3, 0.0f) : x3, w4.w, sigma);
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
    if (kTape) {
      // this thread's 16 outputs of the slice = two 16-byte feature groups of its row
      const uint32_t* wd = kHold ? &held[c * 8] : pk;
      stg128(tape_row + (2 * c) * 1024, wd[0], wd[1], wd[2], wd[3]);
      stg128(tape_row + (2 * c + 1) * 1024, wd[4], wd[5], wd[6], wd[7]);
      if (kRelu) mbits[c >> 1] |= relu_bits16(wd) << ((c & 1) * 16);
    }
    if (kDbg && stamps) stamps[1 + c] = clock64();
    if (c == 3) {
      // this warp's columns are fully read (the last wait::ld covered them)
      tc_fence_before();
      mbar_arrive(dfree_bar);
    }
  }
  if (kTape && kRelu) *tape_mask = make_uint2(mbits[0], mbits[1]);
}

// ------------------------------------------------------------------ the kernel
// kPaper: PaperNeRFModel's layer table (tc_plan.cuh) - layers narrower than H (128 -> 128 in an H = 256 kernel),
// a sigma head on an unrectified layer; a separate instantiation so that the Flexible kernels stay as they are.
template <int H, bool kDbg, bool kTape, bool kPaper = false>
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
      mbar_init(bar(B_ready(t)), 1);
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
    // only the asynchronous instructions themselves are issued by one elected lane.  The passes of the
    // two tiles alternate strictly (a tile's epilogue hides behind the other tile's pass); all waiting
    // is done by the scout warp below, which releases a pass with one arrival on ready[tile].
    const bool leader = elect_one();
    uint32_t w_slot = 0, w_phase = 0;           // ring cursor of tile 0 (tile 1 trails by one pass)
    uint32_t ph_ready[2] = {0, 0};
    uint32_t it = 0;
    const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);   // SBO = 128 B, version 1
    const uint32_t slot0_lo = ((sbase + Smem::w_slots) >> 4) & 0x3FFF;      // 16-byte units
    constexpr int kMain = H / 64;               // 64-wide K chunks of a hidden-activation operand
    // The layer record of the NEXT layer is fetched (indexed constant loads, a few hundred cycles of
    // dependent latency) while the last pass of the current layer is being issued, not between layers.
    TcLayer L = P.layers[0];
    // One pass (both tiles) of layer l.  kPlain = a hidden layer H -> H without a shared-memory operand (most
    // layers): N, the instruction descriptor and the chunk structure are then compile-time constants and the
    // issuing warp - a single dependent instruction stream that shares its scheduler with five other warps,
    // ~10 cycles per instruction - has next to nothing to compute between two passes.
    auto issue_pass = [&](auto plain_tag, int l, int p, bool last_pass) {
      constexpr bool kPlain = decltype(plain_tag)::value;
      const int np = kPlain ? 128 : (L.n_out < 128 ? L.n_out : 128);
      const uint32_t idesc = instr_desc(np);
      const bool has_main = kPlain || L.k_main != 0;
      const uint32_t b_lbo16 = (uint32_t)np;            // LBO = np*16 bytes -> np in 16-byte units
      const bool timing = kDbg && P.dbg_layer == -2 && blockIdx.x == 0 && it == (uint32_t)P.dbg_pass && leader;
      const uint32_t slot_p = w_slot, phase_p = w_phase;   // first chunk of this pass
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const uint32_t a_tmem = tmem_base + (uint32_t)(t * 256);
        const uint32_t d_tmem = a_tmem + 128;
        // Everything this pass waits for (its tile's epilogue, the weight chunks, the encoders) is
        // awaited by the SCOUT warp (warp 3, below), which then arrives on ready[t]: the issuer's
        // own critical path between two passes is one barrier poll.
        if (timing) reinterpret_cast<long long*>(P.dbg)[900 + ((l * 2 + p) * 2 + t) * 3] = clock64();
        mbar_wait(bar(B_ready(t)), ph_ready[t], 1);
        ph_ready[t] ^= 1;
        if (timing) reinterpret_cast<long long*>(P.dbg)[900 + ((l * 2 + p) * 2 + t) * 3 + 1] = clock64();
        uint32_t slot = slot_p, phase = phase_p;
        tc_fence_after();
        if (timing) reinterpret_cast<long long*>(P.dbg)[((l * 2 + p) * 2 + t) * 2] = clock64();
        if (has_main) {
          const int n_main = (kPaper && !kPlain) ? L.k_main / 64 : kMain;
#pragma unroll 1
          for (int c = 0; c < n_main; ++c) {
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
        if (!kPlain && L.smem_src) {
          const uint32_t b_lo = (slot0_lo + slot * (kSlotBytes >> 4)) | (b_lbo16 << 16);
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
          if (!kPlain && last_pass) {
            if (l == P.last_xyz_layer) tc_commit(bar(B_xyzempty(t)));   // encoders may refill
            if (L.smem_src == 2) tc_commit(bar(B_dirempty(t)));
          }
        }
        __syncwarp();
        if (timing) reinterpret_cast<long long*>(P.dbg)[((l * 2 + p) * 2 + t) * 2 + 1] = clock64();
        if (t == 1) { w_slot = slot; w_phase = phase; }
      }
    };
#pragma unroll 1
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
#pragma unroll 1
      for (int l = 0; l < P.n_layers; ++l) {
        TcLayer Lnext = L;
        const int l_next = l + 1 < P.n_layers ? l + 1 : 0;
        if (L.k_main == H && L.smem_src == 0 && L.n_out == H) {
#pragma unroll
          for (int p = 0; p < H / 128; ++p) {
            if (p == H / 128 - 1) Lnext = P.layers[l_next];
            issue_pass(std::true_type{}, l, p, p == H / 128 - 1);
          }
        } else {
#pragma unroll 1
          for (int p = 0; p < L.n_pass; ++p) {
            if (p == L.n_pass - 1) Lnext = P.layers[l_next];
            issue_pass(std::false_type{}, l, p, p == L.n_pass - 1);
          }
        }
        L = Lnext;
      }
    }
  } else if (warp == 3) {
    // =============================== scout ===============================
    // Walks the same pass sequence as the issuer, one step ahead of it: waits for the gate of the
    // next pass (that tile's previous epilogue: d_free, or a_ready for the first pass of a layer),
    // for its weight chunks (tile 0 only; tile 1 finds them resident) and for the encodings, then
    // hands the pass to the issuer with ONE arrival.  It cannot run more than one pass of a tile
    // ahead: the next gate of that tile needs the epilogue of the pass it has just released.
    const bool leader = elect_one();
    uint32_t w_slot = 0, w_phase = 0;
    uint32_t ph_dfree[2] = {0, 0}, ph_aready[2] = {0, 0};
    uint32_t it = 0;
#pragma unroll 1
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x, ++it) {
      const uint32_t pe_ph = it & 1;
#pragma unroll 1
      for (int l = 0; l < P.n_layers; ++l) {
        const TcLayer& L = P.layers[l];
        const int n_chunks = chunks_in_pass(L);
#pragma unroll 1
        for (int p = 0; p < L.n_pass; ++p) {
          const uint32_t slot_p = w_slot, phase_p = w_phase;
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            // The weights and the encodings come first: they arrive independently of this tile's epilogue
            // (the producer and the encoders run ahead) and are normally complete already, but each poll
            // costs ~100 cycles.  The gate - the one wait that is on the tile's MMA -> epilogue -> MMA
            // critical path - comes last, so that nothing but one fence and one arrival follows it.
            uint32_t slot = slot_p, phase = phase_p;
            for (int c = 0; c < n_chunks; ++c) {
              if (t == 0) mbar_wait(bar(B_wfull(slot)), phase, 4);
              if (++slot == kNumSlots) { slot = 0; phase ^= 1; }
            }
            if (L.smem_src == 2 && p == 0) mbar_wait(bar(B_dirfull(t)), pe_ph, 2);
            if (l == 0 && p == 0) mbar_wait(bar(B_xyzfull(t)), pe_ph, 2);
            if (p == 0 && l > 0) {
              mbar_wait(bar(B_aready(t)), ph_aready[t], 3);
              ph_aready[t] ^= 1;
            } else {
              mbar_wait(bar(B_dfree(t)), ph_dfree[t] ^ 1, 1);
            }
            ph_dfree[t] ^= 1;
            tc_fence_after();
            tc_fence_before();
            if (leader) mbar_arrive(bar(B_ready(t)));
            __syncwarp();
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
    const bool timing_all = kDbg && P.dbg_layer == -2 && blockIdx.x == 0 && lane == 0;   // every epilogue warp: latest end
    long long* tl = reinterpret_cast<long long*>(P.dbg);
    const int64_t tap_pair = (int64_t)blockIdx.x + (int64_t)P.dbg_pass * gridDim.x;   // which pair of this CTA the timing tap records
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
          if (timing && pair == tap_pair) tl[256 + ((l * 2) * 2 + t) * 2] = clock64();
          float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
          constexpr int hw = H / 2;           // outputs of the dir layer
          constexpr int mine = hw / 2;        // columns this warp reduces
          const uint32_t col0 = (uint32_t)(hs * mine);
          const uint32_t bias = sbase + Smem::consts + ((uint32_t)L.bias_off + col0) * 4;
          const uint32_t wr = sbase + Smem::consts + ((uint32_t)P.off_wrgb + col0) * 4;
          const uint32_t d_last = a_tmem + 128 + col0;
          uint32_t v[2][16];
          tmem_ld16_issue(d_last, v[0]);
          uint32_t ymask[2] = {0u, 0u};
          uint8_t* y_row = nullptr;
          if (kTape)
            y_row = P.tape + P.tape_act[l] + (pair * 2 + t) * (int64_t)(hw * 256) + (row >> 6) * (hw * 128) +
                    (col0 / 8) * 1024 + (row & 63) * 16;
#pragma unroll
          for (int c = 0; c < mine / 16; ++c) {
            uint32_t ypk[8];
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
              if (kTape) { ypk[i / 2] = pack_bf16(x0, x1, false); ypk[i / 2 + 1] = pack_bf16(x2, x3, false); }
              rgb0 = fmaf(x0, r4.x, fmaf(x1, r4.y, fmaf(x2, r4.z, fmaf(x3, r4.w, rgb0))));
              rgb1 = fmaf(x0, g4.x, fmaf(x1, g4.y, fmaf(x2, g4.z, fmaf(x3, g4.w, rgb1))));
              rgb2 = fmaf(x0, u4.x, fmaf(x1, u4.y, fmaf(x2, u4.z, fmaf(x3, u4.w, rgb2))));
            }
            if (kTape) {
              stg128(y_row + (2 * c) * 1024, ypk[0], ypk[1], ypk[2], ypk[3]);
              stg128(y_row + (2 * c + 1) * 1024, ypk[4], ypk[5], ypk[6], ypk[7]);
              ymask[c >> 1] |= relu_bits16(ypk) << ((c & 1) * 16);
            }
          }
          if (kTape)
            reinterpret_cast<uint2*>(P.tape + P.tape_mask[l] + (pair * 2 + t) * (int64_t)2048 + hs * 1024)[row] =
                make_uint2(ymask[0], ymask[1]);
          tc_fence_before();
          mbar_arrive(bar(B_dfree(t)));
          if (timing && pair == tap_pair) tl[256 + ((l * 2) * 2 + t) * 2 + 1] = clock64();
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
        } else if (kPaper && L.head == 3) {
          // ---- last trunk layer of a model without view directions: ReLU, fc_out (4 outputs) on the CUDA
          // cores over this warp's 64 columns of every pass, final (r,g,b,sigma) store; no A operand follows
          float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
          for (int p = 0; p < H / 128; ++p) {
            mbar_wait(bar(B_dfull(t)), ph_dfull, 5);
            ph_dfull ^= 1;
            tc_fence_after();
            const uint32_t col0 = (uint32_t)(p * 128 + hs * 64);
            const uint32_t bias = sbase + Smem::consts + ((uint32_t)L.bias_off + col0) * 4;
            const uint32_t wo = sbase + Smem::consts + ((uint32_t)P.off_wrgb + col0) * 4;
            uint32_t v[2][16];
            tmem_ld16_issue(d_tmem, v[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tmem_ld16_wait(v[c & 1]);
              if (c + 1 < 4) tmem_ld16_issue(d_tmem + (uint32_t)((c + 1) * 16), v[(c + 1) & 1]);
              if (c == 3) {            // this warp's columns of the accumulator are in registers
                tc_fence_before();
                mbar_arrive(bar(B_dfree(t)));
              }
              if (kDbg && P.dbg_layer == l && P.dbg_pass == p && g < P.m_total) {
                float* dst = P.dbg + g * 128 + hs * 64 + c * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[i] = __uint_as_float(v[c & 1][i]);
              }
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 b4 = lds128(bias + (uint32_t)((c * 16 + i) * 4));
                const float4 w0 = lds128(wo + (uint32_t)((c * 16 + i) * 4));
                const float4 w1 = lds128(wo + (uint32_t)((H + c * 16 + i) * 4));
                const float4 w2 = lds128(wo + (uint32_t)((2 * H + c * 16 + i) * 4));
                const float4 w3 = lds128(wo + (uint32_t)((3 * H + c * 16 + i) * 4));
                const float x0 = fmaxf(__uint_as_float(v[c & 1][i]) + b4.x, 0.0f);
                const float x1 = fmaxf(__uint_as_float(v[c & 1][i + 1]) + b4.y, 0.0f);
                const float x2 = fmaxf(__uint_as_float(v[c & 1][i + 2]) + b4.z, 0.0f);
                const float x3 = fmaxf(__uint_as_float(v[c & 1][i + 3]) + b4.w, 0.0f);
                o0 = fmaf(x0, w0.x, fmaf(x1, w0.y, fmaf(x2, w0.z, fmaf(x3, w0.w, o0))));
                o1 = fmaf(x0, w1.x, fmaf(x1, w1.y, fmaf(x2, w1.z, fmaf(x3, w1.w, o1))));
                o2 = fmaf(x0, w2.x, fmaf(x1, w2.y, fmaf(x2, w2.z, fmaf(x3, w2.w, o2))));
                o3 = fmaf(x0, w3.x, fmaf(x1, w3.y, fmaf(x2, w3.z, fmaf(x3, w3.w, o3))));
              }
            }
          }
          if (hs == 1) *xchg = make_float4(o0, o1, o2, o3);
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          if (hs == 0 && g < P.m_total) {
            const float4 o2nd = *xchg;
            const float* bo = s_const + P.off_brgb;
            float4 o;
            o.x = o0 + o2nd.x + bo[0]; o.y = o1 + o2nd.y + bo[1]; o.z = o2 + o2nd.z + bo[2]; o.w = o3 + o2nd.w + bo[3];
            reinterpret_cast<float4*>(P.rf)[g] = o;
          }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // xchg may be rewritten
        } else {
          const uint32_t bias = sbase + Smem::consts + ((uint32_t)L.bias_off + hs * 64) * 4;
          const uint32_t wa = sbase + Smem::consts + ((uint32_t)P.off_walpha + hs * 64) * 4;
          const int kind = L.relu ? (L.head == 1 ? 2 : 1) : (kPaper && L.head == 1 ? 3 : 0);
#pragma unroll
          for (int p = 0; p < H / 128; ++p) {
            if (kPaper && p >= L.n_pass) break;     // a 128-wide layer of the H = 256 Paper kernel: one pass
            mbar_wait(bar(B_dfull(t)), ph_dfull, 5);
            ph_dfull ^= 1;
            tc_fence_after();
            if (timing && pair == tap_pair) tl[256 + ((l * 2 + p) * 2 + t) * 2] = clock64();
            constexpr bool kTwoPass = (H == 256);
            const bool two_pass = kPaper ? L.n_pass == 2 : kTwoPass;
            const bool hold = two_pass && p == 0;
            float* dbg_dst = (kDbg && P.dbg_layer == l && P.dbg_pass == p && g < P.m_total)
                                 ? P.dbg + g * 128 + hs * 64 : nullptr;
            const uint32_t bp = bias + (uint32_t)(p * 128 * 4);
            const uint32_t wp = wa + (uint32_t)(p * 128 * 4);
            const uint32_t dfree = bar(B_dfree(t));
            const uint32_t a_park = a_tmem + (uint32_t)(hs * 32);                 // pass-0 outputs: K [hs*64, +64)
            const uint32_t a_store = a_tmem + (uint32_t)(p * 64 + hs * 32);       // pass-p outputs
            long long* stamps = (timing && pair == tap_pair && l == 2) ? tl + 600 + (p * 2 + t) * 16 : nullptr;
            uint8_t* tape_row = nullptr;
            uint2* tape_mask = nullptr;
            if (kTape) {
              const int64_t tile = pair * 2 + t;
              tape_row = P.tape + P.tape_act[l] + tile * (int64_t)(H * 256) + (row >> 6) * (H * 128) +
                         (p * 16 + hs * 8) * 1024 + (row & 63) * 16;
              tape_mask = reinterpret_cast<uint2*>(P.tape + P.tape_mask[l] + tile * (int64_t)((H / 128) * 2 * 1024) +
                                                   (p * 2 + hs) * 1024) + row;
            }
            if (hold) {
              if (kind == 0) epilogue_pass<false, false, true, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else if (kind == 1) epilogue_pass<true, false, true, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else if (kPaper && kind == 3) epilogue_pass<false, true, true, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else epilogue_pass<true, true, true, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
            } else if (kPaper && !two_pass) {
              // single-pass layer (dir branch of the Paper model): nothing held, nothing to park
              if (kind == 0) epilogue_pass<false, false, false, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else epilogue_pass<true, false, false, false, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              tmem_wait_st();
              tc_fence_before();
              mbar_arrive(bar(B_aready(t)));
            } else if (H == 128 && DEXNERF_WIDE_EPI != 0) {
              // hidden 128: one pass per layer; the accumulator is fetched with deeper round trips (see above)
              if (kind == 0) epilogue_pass_wide<false, false, kDbg, kTape>(d_tmem, a_store, bp, wp, sigma, dfree, dbg_dst, tape_row, tape_mask);
              else if (kind == 1) epilogue_pass_wide<true, false, kDbg, kTape>(d_tmem, a_store, bp, wp, sigma, dfree, dbg_dst, tape_row, tape_mask);
              else if (kPaper && kind == 3) epilogue_pass_wide<false, true, kDbg, kTape>(d_tmem, a_store, bp, wp, sigma, dfree, dbg_dst, tape_row, tape_mask);
              else epilogue_pass_wide<true, true, kDbg, kTape>(d_tmem, a_store, bp, wp, sigma, dfree, dbg_dst, tape_row, tape_mask);
              tmem_wait_st();
              tc_fence_before();
              mbar_arrive(bar(B_aready(t)));
            } else {
              constexpr bool park = kTwoPass;
              if (kind == 0) epilogue_pass<false, false, false, park, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else if (kind == 1) epilogue_pass<true, false, false, park, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else if (kPaper && kind == 3) epilogue_pass<false, true, false, park, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              else epilogue_pass<true, true, false, park, kDbg, kTape>(d_tmem, a_park, a_store, bp, wp, sigma, held, dfree, dbg_dst, stamps, tape_row, tape_mask);
              if (kDbg && stamps) stamps[9] = clock64();
              tmem_wait_st();
              if (kDbg && stamps) stamps[10] = clock64();
              tc_fence_before();
              mbar_arrive(bar(B_aready(t)));
            }
            if (timing && pair == tap_pair) tl[256 + ((l * 2 + p) * 2 + t) * 2 + 1] = clock64();
            if (timing_all && pair == tap_pair)
              atomicMax(reinterpret_cast<unsigned long long*>(tl) + 800 + (l * 2 + p) * 2 + t, (unsigned long long)clock64());
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
        uint8_t* xyz_tape = nullptr;
        if (kTape)
          xyz_tape = P.tape + P.tape_xyz + (pair * 2 + t) * (int64_t)16384 + (row >> 6) * 8192 + (row & 63) * 16;
        if (std_xyz) {
          encode_row_std<8>(pt, P.dim_xyz, valid, xyz, xyz_tape);
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
            if (kTape) stg128(xyz_tape + k8 * 1024, w4[0], w4[1], w4[2], w4[3]);
          }
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(bar(B_xyzfull(t)));
      }
      for (int t = 0; t < 2 && (!kPaper || P.dim_dir > 0); ++t) {
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
        uint8_t* dir_tape = nullptr;
        if (kTape)
          dir_tape = P.tape + P.tape_dir + (pair * 2 + t) * (int64_t)8192 + (row >> 6) * 4096 + (row & 63) * 16;
        if (std_dir) {
          encode_row_std<4>(dir, P.dim_dir, valid, dr, dir_tape);
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
            if (kTape) stg128(dir_tape + k8 * 1024, w4[0], w4[1], w4[2], w4[3]);
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

// ------------------------------------------------------------------ host: packing
static int64_t blob_bytes(const Plan& P) { return (int64_t)kMaxConstFloats * 4 + P.weight_bytes + (int64_t)P.n_layers * kBiasImgBytes; }

struct PackChunk { int64_t dst; int w_off; int n_out; int n0; int np; int k0; int kc; int k_valid; };

constexpr int kMaxPackChunks = 176;
struct PackTables {     // passed by value (__grid_constant__): no host->device copy, no synchronisation
  int n_chunks, n_moves;
  PackChunk chunks[kMaxPackChunks];
  int4 moves[32];
};

__global__ void pack_weights_kernel(const float* __restrict__ params, const __grid_constant__ PackTables Q,
                                    uint8_t* __restrict__ blob) {
  for (int ci = blockIdx.y; ci < Q.n_chunks; ci += gridDim.y) {
    const PackChunk c = Q.chunks[ci];
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

// Bias images of the three-tile hidden-128 kernel (mlp_tc3.cu): the bias enters the accumulator as the first MMA of a
// pass, ones[128 x 16] x B_l[n_out x 16], with B_l[n][0] = bf16(b[n]) and B_l[n][1] = bf16(b[n] - B_l[n][0]) (the two
// products are exact in the fp32 accumulator; what is lost of the bias is below 2^-17 of its value), zeros elsewhere.
__global__ void pack_bias_kernel(const float* __restrict__ params, const __grid_constant__ PackTables Q,
                                 uint8_t* __restrict__ img) {
  const int4 mv = Q.moves[blockIdx.x];                 // (const offset, b_off, n_out, 0): one per tensor-core layer
  for (int n = threadIdx.x; n < kTileM; n += blockDim.x) {
    uint4 row = make_uint4(0u, 0u, 0u, 0u);
    if (n < mv.z) {
      const float b = params[mv.y + n];
      const __nv_bfloat16 hi = __float2bfloat16_rn(b);
      const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
      row.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    }
    reinterpret_cast<uint4*>(img + (size_t)blockIdx.x * kBiasImgBytes)[n] = row;
  }
}

__global__ void pack_consts_kernel(const float* __restrict__ params, const __grid_constant__ PackTables Q,
                                   float* __restrict__ consts) {
  // moves: (dst, src, count, transpose_cols) ; transpose_cols > 0: src is Wt[count/cols... ] see host
  for (int m = blockIdx.x; m < Q.n_moves; m += gridDim.x) {
    const int4 mv = Q.moves[m];
    for (int i = threadIdx.x; i < mv.z; i += blockDim.x) {
      int src = mv.y + i;
      if (mv.w > 0) {               // head weights: consts[c*hw + k] = Wt[k*nch + c] (fc_rgb: 3 channels, fc_out: 4)
        const int hw = mv.w, nch = mv.z / hw, cch = i / hw, k = i - cch * hw;
        src = mv.y + k * nch + cch;
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
                                           const float* params, void* packed, void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(prog && params && packed, "tc_pack: null pointer");
  const int n_head_ops = spec->arch == 2 ? 1 : 2;
  DN_REQUIRE(prog->n_ops == plan.n_layers + n_head_ops, "tc_pack: program has %d ops, expected %d", prog->n_ops,
             plan.n_layers + n_head_ops);
  cudaStream_t st = (cudaStream_t)stream;
  // chunk table (host) -> workspace (device)
  static thread_local PackTables tables;
  PackChunk* h_chunks = tables.chunks;
  int4* h_moves = tables.moves;
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
        DN_REQUIRE(nch < kMaxPackChunks, "tc_pack: too many chunks");
        PackChunk& pc = h_chunks[nch++];
        const bool main = c < L.k_main / 64;
        pc.dst = dst; pc.w_off = (int)op.w_off; pc.n_out = L.n_out; pc.n0 = p * 128; pc.np = np;
        // layer1: the smem operand is the whole input; skip/dir layers: it follows the k_main hidden inputs -
        // except in the Paper model's skip layer, cat((xyz, x)), where the encoding comes first
        pc.k0 = plan.layers[l].smem_first ? (main ? real_smem + c * 64 : 0) : (main ? c * 64 : L.k_main);
        pc.kc = main ? 64 : L.k_smem;
        pc.k_valid = main ? 64 : real_smem;
        dst += (int64_t)np * pc.kc * 2;
      }
    }
    h_moves[nmv++] = make_int4(L.bias_off, (int)op.b_off, L.n_out, 0);
  }
  const int H = spec->hidden;
  const dexnerf_op& orgb = prog->ops[plan.op_rgb];
  if (spec->arch == 2) {
    DN_REQUIRE(orgb.out_dim == 4 && orgb.src0_dim == H, "tc_pack: fc_out does not match");
    h_moves[nmv++] = make_int4(plan.off_wrgb, (int)orgb.w_off, 4 * H, H);
    h_moves[nmv++] = make_int4(plan.off_brgb, (int)orgb.b_off, 4, 0);
  } else {
    const dexnerf_op& oa = prog->ops[plan.op_alpha];
    DN_REQUIRE(oa.out_dim == 1 && oa.src0_dim == H && orgb.out_dim == 3 && orgb.src0_dim == H / 2,
               "tc_pack: head ops do not match");
    h_moves[nmv++] = make_int4(plan.off_walpha, (int)oa.w_off, H, 0);        // Wt[k][0]
    h_moves[nmv++] = make_int4(plan.off_balpha, (int)oa.b_off, 1, 0);
    h_moves[nmv++] = make_int4(plan.off_wrgb, (int)orgb.w_off, 3 * (H / 2), H / 2);
    h_moves[nmv++] = make_int4(plan.off_brgb, (int)orgb.b_off, 3, 0);
  }
  tables.n_chunks = nch;
  tables.n_moves = nmv;
  DN_CUDA(cudaMemsetAsync(packed, 0, (size_t)kMaxConstFloats * 4, st));
  pack_weights_kernel<<<dim3(8, nch), 256, 0, st>>>(params, tables, reinterpret_cast<uint8_t*>(packed));
  DN_CHECK_LAUNCH("pack_weights");
  pack_consts_kernel<<<nmv, 128, 0, st>>>(params, tables, reinterpret_cast<float*>(packed));
  DN_CHECK_LAUNCH("pack_consts");
  // the first n_layers moves are the layers' biases, in layer order
  pack_bias_kernel<<<plan.n_layers, 128, 0, st>>>(params, tables, reinterpret_cast<uint8_t*>(packed) +
                                                  (size_t)kMaxConstFloats * 4 + plan.weight_bytes);
  DN_CHECK_LAUNCH("pack_bias");
  return 0;
}

static int tc_query_impl(const dexnerf_flexible_spec* spec, const void* packed, const float* ro, const float* rd,
                         const float* viewdirs, const float* z, int64_t n, int S, float* rf, void* tape,
                         float* dbg, int dbg_layer, int dbg_pass, void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(packed && ro && rd && z && rf && (viewdirs || spec->arch == 2), "tc_query: null pointer");
  DN_REQUIRE(S >= 1, "tc_query: S < 1");
  DN_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0 && (reinterpret_cast<uintptr_t>(rf) & 15) == 0,
             "tc_query: packed weights and rf must be 16-byte aligned");
  DN_REQUIRE((reinterpret_cast<uintptr_t>(tape) & 127) == 0, "tc_query: the tape must be 128-byte aligned");
  DN_REQUIRE(!(tape && dbg), "tc_query: the debug taps are not available in the training variant");
  if (n <= 0) return 0;
  TcParams P{};
  P.consts = reinterpret_cast<const float*>(packed);
  P.weights = reinterpret_cast<const uint8_t*>(packed) + (size_t)kMaxConstFloats * 4;
  P.bias_img = P.weights + plan.weight_bytes;
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
  if (tape) {
    // the kernel always processes whole tile PAIRS, so the tape is laid out for an even tile count
    TapeLayout T;
    make_tape_layout(plan, n_pairs * 2, &T);
    P.tape = reinterpret_cast<uint8_t*>(tape);
    P.tape_xyz = T.xyz; P.tape_dir = T.dir;
    for (int l = 0; l < plan.n_layers; ++l) { P.tape_act[l] = T.act[l]; P.tape_mask[l] = T.mask[l]; }
  }
  const int grid = (int)(n_pairs < kNumSMs ? n_pairs : kNumSMs);
  const size_t smem = Smem::total + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kernel) -> int {
    DN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kThreads, smem, st>>>(P);
    return 0;
  };
  int rc;
  DN_REQUIRE(!(tape && spec->arch != 0), "tc_query_train: the training variant exists for FlexibleNeRFModel only");
  // hidden-128 inference: three tiles in flight (mlp_tc3.cu); DEXNERF_TC3=0 keeps the pair kernel for A/B timing
  static const bool use_three = [] { const char* e = getenv("DEXNERF_TC3"); return !(e && e[0] == '0'); }();
  if (spec->hidden == 128 && !tape && !dbg && use_three) {
    if (int rc3 = launch_mlp_tc3(P, n_tiles, st)) return rc3;
    DN_CHECK_LAUNCH("mlp_tc3");
    return 0;
  }
  if (spec->arch != 0 && spec->hidden == 256)
    rc = dbg ? launch(mlp_tc_kernel<256, true, false, true>) : launch(mlp_tc_kernel<256, false, false, true>);
  else if (spec->arch != 0)
    rc = dbg ? launch(mlp_tc_kernel<128, true, false, true>) : launch(mlp_tc_kernel<128, false, false, true>);
  else if (spec->hidden == 256)
    rc = tape ? launch(mlp_tc_kernel<256, false, true>)
              : (dbg ? launch(mlp_tc_kernel<256, true, false>) : launch(mlp_tc_kernel<256, false, false>));
  else
    rc = tape ? launch(mlp_tc_kernel<128, false, true>)
              : (dbg ? launch(mlp_tc_kernel<128, true, false>) : launch(mlp_tc_kernel<128, false, false>));
  if (rc) return rc;
  DN_CHECK_LAUNCH("mlp_tc");
  return 0;
}

extern "C" DEXNERF_API int dexnerf_tc_query(const dexnerf_flexible_spec* spec, const void* packed, const float* ro,
                                            const float* rd, const float* viewdirs, const float* z, int64_t n,
                                            int S, float* rf, float* dbg, int dbg_layer, int dbg_pass,
                                            void* stream) {
  return tc_query_impl(spec, packed, ro, rd, viewdirs, z, n, S, rf, nullptr, dbg, dbg_layer, dbg_pass, stream);
}

extern "C" DEXNERF_API int64_t dexnerf_tc_tape_bytes(const dexnerf_flexible_spec* spec, int64_t n_samples) {
  Plan plan;
  if (make_plan(spec, &plan)) return -1;
  if (spec->arch != 0) { set_error("tc_tape_bytes: the training variant exists for FlexibleNeRFModel only"); return -1; }
  if (n_samples < 0) { set_error("tc_tape_bytes: negative sample count"); return -1; }
  const int64_t n_pairs = ((n_samples + kTileM - 1) / kTileM + 1) / 2;
  TapeLayout T;
  make_tape_layout(plan, n_pairs * 2, &T);
  return T.total;
}

extern "C" DEXNERF_API int dexnerf_tc_query_train(const dexnerf_flexible_spec* spec, const void* packed,
                                                  const float* ro, const float* rd, const float* viewdirs,
                                                  const float* z, int64_t n, int S, float* rf, void* tape,
                                                  void* stream) {
  DN_REQUIRE(tape, "tc_query_train: null tape");
  return tc_query_impl(spec, packed, ro, rd, viewdirs, z, n, S, rf, tape, nullptr, -1, 0, stream);
}
