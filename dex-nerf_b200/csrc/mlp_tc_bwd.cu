// tcgen05 backward of FlexibleNeRFModel (nerf/models.py:185-256, repaired forward) for the
// training step of BASELINE config 4 (train_dexnerf_rgb.py:264-281: loss.backward()).
//
// The forward (mlp_tc.cu, kTape) leaves a TAPE in HBM: the bf16 input image of every layer and
// one ReLU bit per activation (tc_plan.cuh, TapeLayout).  The backward is two kernels:
//
//   mlp_tc_bwd_dx_kernel   the activation-gradient chain.  Same machine as the forward: one
//       persistent CTA per SM, pairs of 128-sample tiles, gradients G_l = dL/d(pre-activation_l)
//       resident in TENSOR MEMORY as the bf16 A operand, TRANSPOSED weight images streamed by bulk
//       TMA into a shared-memory ring, D = G_{l+1} . W_l[:, :H] accumulated in TMEM, epilogue =
//       (+ d sigma * w_alpha) -> ReLU mask from the tape -> bf16 -> next A operand.  The first
//       operand (G of layers_dir[0]) comes from d rgb through fc_rgb on the CUDA cores.  Every G is
//       also written to the tape as an MN-major bf16 image.
//   mlp_tc_bwd_dw_kernel   the weight gradients dWt_l[in][out] = X_l^T . G_l: a split-K GEMM whose
//       K dimension is the SAMPLES.  Both operands are the tape images (MN-major UMMA operands,
//       loaded by 1-D bulk copies, 3-stage mbarrier ring of 64-sample half-tiles), accumulators for
//       up to 2 x 128 input features x 256 outputs fill the 512 TMEM columns, CTAs split the tiles
//       of a layer in proportion to its bytes and reduce with red.global.add.f32.  Bias gradients
//       are column sums of G taken from the shared-memory image by the otherwise idle warps.
//
//   mlp_tc_bwd_fused_kernel   BOTH of the above in ONE launch (the training path): the first n_chain CTAs run the
//       activation-gradient chain, the others the weight-gradient GEMM, and the G images never visit HBM: a chain
//       CTA releases a per-(layer, tile) flag in global memory when an image is complete (dirty in the 126 MB L2),
//       the GEMM CTA of that (layer, tile) acquires it, bulk-copies the image out of L2 and then DISCARDS its lines
//       (discard.global.L2: dropped without write-back).  The chain never waits for the GEMM (no back-pressure, so
//       no deadlock); a GEMM CTA that falls behind only lets lines spill to HBM.  The tile split of the GEMM CTAs is
//       interleaved (tile = split + k * n_cta) so that they consume in the order the chain produces.
//
// Arithmetic contract: bf16-rounded G and activations as tensor-core operands, fp32 accumulation;
// fc_alpha / fc_rgb gradients use the same bf16 images.  oracle.train_step(bf16=True) is the
// statement of this contract; the fp32 reference gradients are matched to bf16 tolerance.
#include "tc_plan.cuh"
#include "tc_ptx.cuh"

namespace dexnerf {
namespace tc {

// =====================================================================================
//                                  dX chain kernel
// =====================================================================================
constexpr int kBSlotBytes = 16384;   // one transposed chunk: 128 in-features x 64 out-features
constexpr int kBSlots = 11;
constexpr int kBThreads = 640;       // 4 control warps + 16 epilogue warps
constexpr int kBEpiThreads = 256;    // per tile

struct BwdParams {
  const uint8_t* weights_t;   // transposed chunk images in consumption order
  const float* consts;        // forward const block (w_alpha, W_rgb live there)
  const float4* d_rf;         // dL/d(rgb, sigma) per sample
  uint8_t* tape;
  int64_t m_total;
  int nl;                     // tensor-core layers of the forward (layer1, trunk.., fc_feat, dir)
  int n_const, off_walpha, off_wrgb;
  int64_t mask_off[kMaxLayers];
  int64_t grad_off[kMaxLayers];
  int64_t ghead_off;
  int relu[kMaxLayers];       // forward ReLU flag of each layer (its G needs the mask)
  uint32_t* ready;            // fused launch: [kMaxLayers + 1][n_tiles] completion counters of the G images (else NULL)
  int64_t n_tiles_pad;        // tiles the tape is laid out for (an even count)
  // back-pressure (shared-SM kernel): a chain may start tile t only when the weight-gradient side has finished
  // throttle_items * (t - throttle_window + 1) (item, tile) units, so that the gradient images in flight stay
  // L2-resident; NULL = no throttle
  const uint32_t* consumed_total;
  int throttle_items, throttle_window;
};

// ---- cross-CTA hand-off of a finished G image (fused launch).  Producer: the warp's stores are ordered before
// lane 0's gpu-scope release by __syncwarp + the fence (cumulativity); the image is read by the consumer's bulk copy
// (async proxy), hence the proxy fence on both sides.
__device__ __forceinline__ void signal_image(uint32_t* flag, int lane) {
  __syncwarp();
  if (lane == 0) {
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
  }
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
static __device__ __noinline__ void flag_timeout(int who) {
  printf("dexnerf fused backward: flag wait timeout (site %d, block %d, thread %d)\n", who, blockIdx.x, threadIdx.x);
  __trap();
}
// soft (bring-up, variant bit 6): give up after a short while and carry on with whatever the tape holds
__device__ __forceinline__ void wait_image(const uint32_t* flag, uint32_t expected, int who, bool soft) {
  uint32_t spins = 0;
  while (ld_acquire_u32(flag) < expected) {
    __nanosleep(64);
    ++spins;
    if (soft && spins > (1u << 14)) return;
    if (spins > (1u << 22)) flag_timeout(who);     // seconds: a protocol bug must trap, not hang the GPU
  }
}

// The backward reads only the head weights of the const block, [off_walpha, n_const) = H + 4 + 3 H / 2 + 4 floats; the
// shared-SM kernel keeps just those (kCompact) and spends the shared memory on weight slots instead.
constexpr int kCompactConstFloats = 1024;
template <int kSlots, bool kCompact = false>
struct BSmemT {
  static constexpr int w_slots = 0;
  static constexpr int consts = w_slots + kSlots * kBSlotBytes;
  static constexpr int bars = consts + (kCompact ? kCompactConstFloats : kMaxConstFloats) * 4;
  static constexpr int n_bars = 2 * kSlots + 6;
  static constexpr int tmem_ptr = bars + n_bars * 8;
  static constexpr int img_cnt = tmem_ptr + 16;      // fused: uint32[2], epilogue warps that have stored their pass
  static constexpr int total = (img_cnt + 16 + 127) / 128 * 128;
};
using BSmem = BSmemT<kBSlots>;
constexpr int kBSlotsShared = 7;      // weight ring of the chain when it shares the SM with the weight-gradient GEMM
// whole-CTA barrier with an explicit thread count: the role groups of the shared-SM kernel reach it from different
// call sites
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
// fused launch: epilogue warp -> signaller hand-off through a MONOTONIC shared-memory counter (an mbarrier's
// parity wait cannot tell phase k from phase k + 2, and the signaller, which pays a gpu-scope fence per image, may
// lag the epilogue by several passes)
__device__ __forceinline__ void img_stored(uint32_t cnt_addr, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(cnt_addr) : "memory");
}
__device__ __forceinline__ void img_wait(uint32_t cnt_addr, uint32_t need, int who) {
  uint32_t v, spins = 0;
  while (true) {
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(cnt_addr) : "memory");
    if (v >= need) break;
    if (++spins > kSpinLimit) barrier_timeout(who);
  }
}

// One backward epilogue pass of one warp over its 64 accumulator columns: optional rank-1 term
// d_sigma * w_alpha, optional ReLU mask, bf16 pack, A-operand store (held / parked / direct) and
// the MN-major tape image of G.
template <bool kAdd, bool kHold, bool kPark, bool kStore>
__device__ __forceinline__ void bwd_epilogue_pass(uint32_t d_tmem, uint32_t a_park, uint32_t a_store, bool use_mask,
                                                  uint2 mask, float dsig, uint32_t wa, uint32_t (&held)[32],
                                                  uint32_t dfree_bar, uint8_t* tape_row) {
  uint32_t v[2][16];
  tmem_ld16_issue(d_tmem, v[0]);
  if (kPark) {
    tmem_st16(a_park, &held[0]);
    tmem_st16(a_park + 16, &held[16]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    tmem_ld16_wait(v[c & 1]);
    if (c + 1 < 4) tmem_ld16_issue(d_tmem + (uint32_t)((c + 1) * 16), v[(c + 1) & 1]);
    const uint32_t bits = ((c < 2 ? mask.x : mask.y) >> ((c & 1) * 16)) & 0xFFFFu;
    uint32_t pk[8];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      float x0 = __uint_as_float(v[c & 1][i]), x1 = __uint_as_float(v[c & 1][i + 1]);
      float x2 = __uint_as_float(v[c & 1][i + 2]), x3 = __uint_as_float(v[c & 1][i + 3]);
      if (kAdd) {
        const float4 w4 = lds128(wa + (uint32_t)((c * 16 + i) * 4));
        x0 = fmaf(dsig, w4.x, x0); x1 = fmaf(dsig, w4.y, x1);
        x2 = fmaf(dsig, w4.z, x2); x3 = fmaf(dsig, w4.w, x3);
      }
      if (use_mask) {
        x0 = (bits >> relu_bit_pos(i)) & 1u ? x0 : 0.0f;     x1 = (bits >> relu_bit_pos(i + 1)) & 1u ? x1 : 0.0f;
        x2 = (bits >> relu_bit_pos(i + 2)) & 1u ? x2 : 0.0f; x3 = (bits >> relu_bit_pos(i + 3)) & 1u ? x3 : 0.0f;
      }
      if (kHold) {
        held[c * 8 + i / 2] = pack_bf16(x0, x1, false);
        held[c * 8 + i / 2 + 1] = pack_bf16(x2, x3, false);
      } else {
        pk[i / 2] = pack_bf16(x0, x1, false);
        pk[i / 2 + 1] = pack_bf16(x2, x3, false);
      }
    }
    const uint32_t* wd = kHold ? &held[c * 8] : pk;
    if (kStore) tmem_st8(a_store + (uint32_t)(c * 8), wd);
    *reinterpret_cast<uint4*>(tape_row + (2 * c) * 1024) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    *reinterpret_cast<uint4*>(tape_row + (2 * c + 1) * 1024) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
    if (c == 3) {
      tc_fence_before();
      mbar_arrive(dfree_bar);
    }
  }
}

// cta / n_cta: this CTA's index among the chain CTAs (the whole grid in the stand-alone kernel).
// kTiles: 128-sample tiles in flight per CTA - 2 (a pair: one tile's MMA pass hides the other's epilogue; all 512
// TMEM columns) or 1 (256 columns: the shared-SM kernel, where the weight-gradient GEMM of the same CTA keeps the
// tensor pipe busy during the epilogue).  kSlots: weight ring depth.  Threads of warps >= 4 + 8 * kTiles belong to
// somebody else and only take part in the two CTA-wide barriers.
template <int H, bool kFused, int kTiles, int kSlots>
__device__ __forceinline__ void chain_body(const BwdParams& P, uint8_t* smem, const int cta, const int n_cta) {
  constexpr bool kCompact = kSlots != kBSlots;       // (the shared-SM instantiation)
  using BSmem = BSmemT<kSlots, kCompact>;
  const int c_base = kCompact ? P.off_walpha : 0;    // first float of the const block that is kept in shared memory
  auto BB_wfull = [](int s) { return s; };
  auto BB_wempty = [](int s) { return kSlots + s; };
  auto BB_aready = [](int t) { return 2 * kSlots + t; };
  auto BB_dfull = [](int t) { return 2 * kSlots + 2 + t; };
  auto BB_dfree = [](int t) { return 2 * kSlots + 4 + t; };
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + BSmem::bars;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fused launches walk every tile the tape is laid out for (an even count): the weight-gradient side waits for
  // the flags of the padding tile too (its gradients are zero)
  const int64_t n_tiles = kFused ? P.n_tiles_pad : (P.m_total + kTileM - 1) / kTileM;
  const int64_t n_pairs = (n_tiles + kTiles - 1) / kTiles;     // work units: tile pairs, or single tiles
  constexpr int kPass = H / 128;          // N = 128 passes of a layer
  const int n_steps = P.nl - 1;           // MMA layers: step j consumes G[nl-1-j], produces G[nl-2-j]
  constexpr int kChainThreads = (4 + 8 * kTiles) * 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) { mbar_init(bar(BB_wfull(s)), 1); mbar_init(bar(BB_wempty(s)), kTiles); }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(bar(BB_aready(t)), kBEpiThreads);
      mbar_init(bar(BB_dfull(t)), 1);
      mbar_init(bar(BB_dfree(t)), kBEpiThreads);
    }
    reinterpret_cast<volatile uint32_t*>(smem + BSmem::img_cnt)[0] = 0u;
    reinterpret_cast<volatile uint32_t*>(smem + BSmem::img_cnt)[1] = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + BSmem::tmem_ptr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x < kChainThreads) {
    float* c = reinterpret_cast<float*>(smem + BSmem::consts);
    for (int i = c_base + threadIdx.x; i < P.n_const; i += kChainThreads) c[i - c_base] = P.consts[i];
  }
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + BSmem::tmem_ptr);

  if (warp == 0) {
    // =============================== weight producer ===============================
    const bool leader = elect_one();
    uint32_t cnt = 0;
#pragma unroll 1
    for (int64_t pair = cta; pair < n_pairs; pair += n_cta) {
      const uint8_t* src = P.weights_t;
      if (kFused && P.consumed_total) {
        // back-pressure: the first weight chunk of a tile is not requested (so the issuer does not start it) until
        // the weight-gradient side is within `throttle_window` tiles of it
        const int64_t need = (int64_t)P.throttle_items * (pair * kTiles - P.throttle_window + 1);
        if (need > 0 && lane == 0) {
          uint32_t spins = 0;
          while ((int64_t)ld_acquire_u32(P.consumed_total) < need) {
            __nanosleep(256);
            if (++spins > (1u << 22)) flag_timeout(25);
          }
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int j = 0; j < n_steps; ++j) {
        const int nc = (j == 0 ? H / 2 : H) / 64;
#pragma unroll 1
        for (int pc = 0; pc < kPass * nc; ++pc, ++cnt) {
          const uint32_t slot = cnt % kSlots, ph = (cnt / kSlots) & 1;
          mbar_wait(bar(BB_wempty(slot)), ph ^ 1, 10);
          if (leader) {
            mbar_arrive_expect_tx(bar(BB_wfull(slot)), kBSlotBytes);
            bulk_g2s(sbase + BSmem::w_slots + slot * kBSlotBytes, src, kBSlotBytes, bar(BB_wfull(slot)));
          }
          __syncwarp();
          src += kBSlotBytes;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = elect_one();
    uint32_t w_slot = 0, w_phase = 0;
    uint32_t ph_dfree[2] = {0, 0}, ph_aready[2] = {0, 0};
    const uint64_t desc_hi = ((uint64_t)(128 >> 4) << 32) | (1ull << 46);   // SBO = 128 B, version 1
    const uint32_t slot0_lo = ((sbase + BSmem::w_slots) >> 4) & 0x3FFF;
    const uint32_t idesc = instr_desc(128);
    constexpr uint32_t b_lbo16 = 128;      // LBO = 128 rows * 16 B, in 16-byte units
#pragma unroll 1
    for (int64_t pair = cta; pair < n_pairs; pair += n_cta) {
#pragma unroll 1
      for (int j = 0; j < n_steps; ++j) {
        const int nc = (j == 0 ? H / 2 : H) / 64;
#pragma unroll 1
        for (int p = 0; p < kPass; ++p) {
          const uint32_t slot_p = w_slot, phase_p = w_phase;
#pragma unroll
          for (int t = 0; t < kTiles; ++t) {
            const uint32_t a_tmem = tmem_base + (uint32_t)(t * 256);
            const uint32_t d_tmem = a_tmem + 128;
            if (p == 0) {
              mbar_wait(bar(BB_aready(t)), ph_aready[t], 11);
              ph_aready[t] ^= 1;
            } else {
              mbar_wait(bar(BB_dfree(t)), ph_dfree[t] ^ 1, 12);
            }
            ph_dfree[t] ^= 1;
            uint32_t slot = slot_p, phase = phase_p;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < nc; ++c) {
              if (t == 0) { mbar_wait(bar(BB_wfull(slot)), phase, 13); tc_fence_after(); }
              const uint32_t b_lo = (slot0_lo + slot * (kBSlotBytes >> 4)) | (b_lbo16 << 16);
              if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma_ts(d_tmem, a_tmem + (uint32_t)(c * 32 + ks * 8),
                         desc_hi | (uint64_t)(b_lo + (uint32_t)ks * 2 * b_lbo16), idesc, (c | ks) ? 1u : 0u);
                tc_commit(bar(BB_wempty(slot)));
              }
              __syncwarp();
              if (++slot == kSlots) { slot = 0; phase ^= 1; }
            }
            if (leader) tc_commit(bar(BB_dfull(t)));
            __syncwarp();
            if (t == kTiles - 1) { w_slot = slot; w_phase = phase; }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + 8 * kTiles) {
    // =============================== epilogue ===============================
    const int e = warp - 4;
    const int t = e >> 3, hs = (e >> 2) & 1, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_tmem = tmem_base + (uint32_t)(t * 256) + lane_base;
    const uint32_t d_tmem = a_tmem + 128 + (uint32_t)(hs * 64);
    const float* s_const = reinterpret_cast<const float*>(smem + BSmem::consts);
    uint32_t ph_dfull = 0;
    uint32_t held[32];
    const int half = row >> 6, r64 = row & 63;
#pragma unroll 1
    for (int64_t pair = cta; pair < n_pairs; pair += n_cta) {
      const int64_t tile = pair * kTiles + t;
      const int64_t g = tile * kTileM + row;
      const float4 d = (g < P.m_total) ? P.d_rf[g] : make_float4(0.f, 0.f, 0.f, 0.f);
      // ---- G of layers_dir[0]: (W_rgb^T d_rgb) masked by y > 0, on the CUDA cores
      {
        constexpr int hw = H / 2, mine = hw / 2;
        const int col0 = hs * mine;
        const int ld = P.nl - 1;
        const uint2 ym = reinterpret_cast<const uint2*>(P.tape + P.mask_off[ld] + tile * 2048 + hs * 1024)[row];
        const float* wr = s_const + (P.off_wrgb - c_base) + col0;
        uint8_t* trow = P.tape + P.grad_off[ld] + tile * (int64_t)(hw * 256) + half * (hw * 128) +
                        (col0 / 8) * 1024 + r64 * 16;
#pragma unroll
        for (int c = 0; c < mine / 16; ++c) {
          const uint32_t bits = ((c < 2 ? ym.x : ym.y) >> ((c & 1) * 16)) & 0xFFFFu;
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const int k = c * 16 + i;
            float x0 = fmaf(d.x, wr[k], fmaf(d.y, wr[hw + k], d.z * wr[2 * hw + k]));
            float x1 = fmaf(d.x, wr[k + 1], fmaf(d.y, wr[hw + k + 1], d.z * wr[2 * hw + k + 1]));
            x0 = (bits >> relu_bit_pos(i)) & 1u ? x0 : 0.0f;
            x1 = (bits >> relu_bit_pos(i + 1)) & 1u ? x1 : 0.0f;
            pk[i / 2] = pack_bf16(x0, x1, false);
          }
          tmem_st8(a_tmem + (uint32_t)(col0 / 2 + c * 8), pk);
          *reinterpret_cast<uint4*>(trow + (2 * c) * 1024) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(trow + (2 * c + 1) * 1024) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        if (hs == 0) {   // the head operand [d rgb, d sigma, 0 ...] (16 features) for dW of fc_rgb / fc_alpha
          uint8_t* hrow = P.tape + P.ghead_off + tile * 4096 + half * 2048 + r64 * 16;
          *reinterpret_cast<uint4*>(hrow) = make_uint4(pack_bf16(d.x, d.y, false), pack_bf16(d.z, d.w, false), 0u, 0u);
          *reinterpret_cast<uint4*>(hrow + 1024) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (kFused) img_stored(sbase + BSmem::img_cnt + 4u * (uint32_t)t, lane);   // G of the dir layer (+ the head operand) is stored
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(bar(BB_aready(t)));
      }
      // the bf16-rounded d sigma, as the tensor-core operand of dW(fc_alpha) sees it
      const float dsig = __bfloat162float(__float2bfloat16_rn(d.w));
#pragma unroll 1
      for (int j = 0; j < n_steps; ++j) {
        const int dst = P.nl - 2 - j;
        const bool last = (j == n_steps - 1);
        const bool use_mask = P.relu[dst] != 0;
        const bool add = (j == 1);
#pragma unroll
        for (int p = 0; p < kPass; ++p) {
          uint2 mask = make_uint2(0u, 0u);
          if (use_mask)
            mask = reinterpret_cast<const uint2*>(P.tape + P.mask_off[dst] + tile * (int64_t)(kPass * 2 * 1024) +
                                                  (p * 2 + hs) * 1024)[row];
          uint8_t* trow = P.tape + P.grad_off[dst] + tile * (int64_t)(H * 256) + half * (H * 128) +
                          (p * 16 + hs * 8) * 1024 + r64 * 16;
          const uint32_t wa = sbase + BSmem::consts + (uint32_t)(P.off_walpha - c_base + p * 128 + hs * 64) * 4;
          const uint32_t a_park = a_tmem + (uint32_t)(hs * 32);
          const uint32_t a_store = a_tmem + (uint32_t)(p * 64 + hs * 32);
          const uint32_t dfree = bar(BB_dfree(t));
          mbar_wait(bar(BB_dfull(t)), ph_dfull, 14);
          ph_dfull ^= 1;
          tc_fence_after();
          if (last) {
            bwd_epilogue_pass<false, false, false, false>(d_tmem, a_park, a_store, use_mask, mask, dsig, wa, held, dfree, trow);
          } else if (kPass == 2 && p == 0) {
            if (add) bwd_epilogue_pass<true, true, false, false>(d_tmem, a_park, a_store, use_mask, mask, dsig, wa, held, dfree, trow);
            else bwd_epilogue_pass<false, true, false, false>(d_tmem, a_park, a_store, use_mask, mask, dsig, wa, held, dfree, trow);
          } else {
            constexpr bool park = (kPass == 2);
            if (add) bwd_epilogue_pass<true, false, park, true>(d_tmem, a_park, a_store, use_mask, mask, dsig, wa, held, dfree, trow);
            else bwd_epilogue_pass<false, false, park, true>(d_tmem, a_park, a_store, use_mask, mask, dsig, wa, held, dfree, trow);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(bar(BB_aready(t)));
          }
          if (kFused) img_stored(sbase + BSmem::img_cnt + 4u * (uint32_t)t, lane);   // this pass's rows of G[dst] are stored
        }
      }
    }
  } else if (kFused && warp >= 2 && warp < 2 + kTiles) {
    // =============================== signallers (fused launch; warp 2: tile 0, warp 3: tile 1) ===============
    // The 8 epilogue warps of a tile only bump a shared-memory counter after their stores (release.cta); THIS warp
    // pays the gpu-scope fence (~1 us) and publishes the image to the weight-gradient CTAs, off the
    // MMA -> epilogue -> MMA chain.  By cumulativity the fence orders the stores it has synchronised with before
    // the flag.
    const int t = warp - 2;
    const uint32_t cnt = sbase + BSmem::img_cnt + 4u * (uint32_t)t;
    uint32_t need = 0;            // epilogue-warp arrivals so far: 8 per pass
#pragma unroll 1
    for (int64_t pair = cta; pair < n_pairs; pair += n_cta) {
      const int64_t tile = pair * kTiles + t;
      need += 8;
      img_wait(cnt, need, 15);
      signal_image(P.ready + (int64_t)(P.nl - 1) * P.n_tiles_pad + tile, lane);
      signal_image(P.ready + (int64_t)kMaxLayers * P.n_tiles_pad + tile, lane);
#pragma unroll 1
      for (int j = 0; j < n_steps; ++j) {
        need += 8 * kPass;
        img_wait(cnt, need, 15);
        signal_image(P.ready + (int64_t)(P.nl - 2 - j) * P.n_tiles_pad + tile, lane);
      }
    }
  }

  tc_fence_before();
  cta_sync();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

template <int H>
__global__ void __launch_bounds__(kBThreads, 1) mlp_tc_bwd_dx_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  chain_body<H, false, 2, kBSlots>(P, smem, (int)blockIdx.x, (int)gridDim.x);
}

// ---- transposed weight images for the dX chain (one 16 KB chunk = 128 in-features x 64 out-features)
struct BwdPackChunk { int w_off, ld, n0, k0; };
struct BwdPackParams { int n_chunks; BwdPackChunk chunks[200]; };

__global__ void pack_weights_t_kernel(const float* __restrict__ params, const __grid_constant__ BwdPackParams Q,
                                      uint8_t* __restrict__ blob) {
  for (int ci = blockIdx.y; ci < Q.n_chunks; ci += gridDim.y) {
    const BwdPackChunk c = Q.chunks[ci];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 128 * 64; e += gridDim.x * blockDim.x) {
      const int n = e >> 6, k = e & 63;          // k fastest: coalesced reads of Wt[n][k0 + k]
      const float w = params[c.w_off + (int64_t)(c.n0 + n) * c.ld + c.k0 + k];
      const int64_t off = (int64_t)ci * kBSlotBytes + (k >> 3) * (128 * 16) + n * 16 + (k & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(blob + off) = __float2bfloat16_rn(w);
    }
  }
}

// =====================================================================================
//                              weight-gradient GEMM kernel
// =====================================================================================
constexpr int kWStageA = 32768;     // up to 2 M-blocks x 16 feature groups x 64 samples x 16 B
constexpr int kWStageG = 32768;     // up to 32 feature groups (N = 256)
constexpr int kWStages = 3;
constexpr int kWThreads = 256;      // producer, MMA issuer, TMEM allocator, (idle), 4 reducer warps
constexpr int kMaxDwItems = 40;     // <= 16 layers (x 2 M blocks in the shared-SM kernel) + encoding parts + heads

struct DwItem {
  int64_t a_off;     // tape offset of the A image array
  int64_t g_off;     // tape offset of the G image array
  int64_t w_out;     // float offset into `grads` of dWt[first in-feature of this item][0]
  int64_t b_out;     // float offset of the bias gradient, -1: none
  int a_fg;          // feature groups of the whole A image (per-tile stride a_fg * 2048)
  int a_fg0;         // first feature group covered by this item
  int a_fgs;         // feature groups copied per half-tile (<= 32)
  int n_mblk;        // M blocks of 128 in-features (1 or 2)
  int a_rows;        // valid in-features (rows flushed)
  int g_fg;          // feature groups of G; N = 8 * g_fg
  int col0, n_cols;  // output columns flushed: D[:, col0 : col0 + n_cols] -> dWt[:, 0 : n_cols]
  int ld;            // row stride of dWt (out features of the layer)
  int cta0, n_cta;   // CTAs [cta0, cta0 + n_cta) split this item's tiles
  int flag_row;      // fused launch: row of the `ready` table that announces this item's G image
  int n_consumers;   // ... and how many items read that image (the last reader discards it from L2)
};

struct DwParams {
  const uint8_t* tape;
  float* grads;
  int64_t n_tiles;
  int n_items;
  int variant;       // bring-up knob: bit 0 swaps the LBO / SBO fields of the MN-major descriptors
  int n_cta_total;         // CTAs that own an item (the shared-SM kernel launches one CTA per SM regardless)
  const uint32_t* ready;   // fused launch: [kMaxLayers + 1][n_tiles] image-complete counters written by the chain CTAs
  uint32_t* consumed;      // fused launch: [kMaxLayers + 1][n_tiles][2] readers done with a half image
  uint32_t* consumed_total;   // shared-SM kernel: (item, tile) units finished - the chains' back-pressure signal
  // Stand-alone kernel: fc_alpha's gradient rides on the item of the layer that reads the same activations (the
  // feature layer): dW_alpha[k] = sum_s X[s][k] d_sigma[s] is a WEIGHTED column sum of the A image, taken by the
  // column-sum warps with mma.sync (d_sigma as the 16 identical rows of the first operand).  A separate item would
  // stream that image - 4.5 % of the kernel's bytes - a second time.
  int head_item;              // item that carries the head, -1: none
  int head_rows;              // in-features of fc_alpha
  int64_t head_g_off;         // tape offset of the [d rgb, d sigma, 0 ...] image (2 feature groups per tile half)
  int64_t head_w, head_b;     // float offsets of fc_alpha's weight / bias gradient
  DwItem items[kMaxDwItems];
};

template <int kStages, int kStageA, int kStageG>
struct WSmemT {
  static constexpr int a = 0;
  static constexpr int g = a + kStages * kStageA;
  static constexpr int bars = g + kStages * kStageG;
  static constexpr int n_bars = 2 * kStages + 1;
  static constexpr int tmem_ptr = bars + n_bars * 8;
  static constexpr int total = tmem_ptr + 16;
};
using WSmem = WSmemT<kWStages, kWStageA, kWStageG>;
// The stand-alone weight-gradient GEMM packs its ring per item: a stage is as large as the item's operands
// (n_mblk x 16 KB of A + g_fg KB of G) and the 192 KB hold as many stages as fit, up to 8.  An item with small
// operands (the heads: 18 KB per stage; the direction encoding: 32 KB) is otherwise bound by the latency of three
// small copies in flight, and the kernel ends with its slowest CTA.
constexpr int kWPackedStages = 8;
constexpr int kWHeadBytes = 1024;    // feature group 0 of a [d rgb, d sigma, ...] half image (see DwParams::head_item)
constexpr int kWRingBytes = kWStages * (kWStageA + kWStageG + 2048);
using WSmemPacked = WSmemT<kWPackedStages, kWRingBytes / kWPackedStages / 2, kWRingBytes / kWPackedStages / 2>;
static_assert(WSmemPacked::bars == kWRingBytes, "the packed ring is the same 192 KB");
// the weight-gradient GEMM when it shares the SM with the chain: ONE M block per item (256 TMEM columns), so a stage
// is 16 KB of A (16 feature groups x 64 samples) + 32 KB of G, two stages
constexpr int kWStagesShared = 2;
constexpr int kWStageAShared = 16384;
using WSmemShared = WSmemT<kWStagesShared, kWStageAShared, kWStageG>;

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// drop 128-byte lines from L2 without writing them back (the data is dead: every reader has its copy)
__device__ __forceinline__ void discard_l2(const uint8_t* p) {
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
}

// cta: this CTA's index among the weight-gradient CTAs.  kFused: the G images come from chain CTAs of the same launch
// (wait for their flags, interleaved tile split, discard after use).
// kWarp0 = 0: the CTA is all ours (8 working warps: producer, issuer, TMEM allocator, -, 4 reducers; further warps
// idle).  kWarp0 > 0 (shared-SM kernel): our 8 warps start at warp kWarp0, the TMEM allocation belongs to the chain
// group (*shared_tmem holds its base) and we use the columns from kCol0 on; the other warps of the CTA only meet us
// at the two CTA-wide barriers.
// kPacked: the ring geometry follows the item (WSmemPacked above); otherwise kStages stages of kStageA + kWStageG bytes.
template <bool kFused, int kWarp0, int kCol0, typename WS, int kStages, int kStageA, bool kPacked = false>
__device__ __forceinline__ void dw_body(const DwParams& P, uint8_t* smem, const int cta, const uint8_t* shared_tmem) {
  using WSmem = WS;
  constexpr int kWStages = kStages;        // (kPacked: the maximum)
  constexpr int kWStageA = kStageA;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + WSmem::bars;
  auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
  auto empty = [&](int s) { return bars + 8u * (uint32_t)(kWStages + s); };
  const uint32_t acc_bar = bars + 8u * (uint32_t)(2 * kWStages);
  const int warp = (int)(threadIdx.x >> 5) - kWarp0, lane = threadIdx.x & 31;
  const int tid = (int)threadIdx.x - kWarp0 * 32;                 // thread index within the group
  const int n_group = kWarp0 ? 256 : (int)blockDim.x;

  // which item / which slice of its tiles
  int it = 0;
  while (it + 1 < P.n_items && cta >= P.items[it].cta0 + P.items[it].n_cta) ++it;
  const DwItem& I = P.items[it];
  const int split = cta - I.cta0;
  // stand-alone: a contiguous range of tiles; fused: every n_cta-th tile, in the order the chain produces them
  const int64_t tile_begin = kFused ? split : P.n_tiles * split / I.n_cta;
  const int64_t tile_end = kFused ? P.n_tiles : P.n_tiles * (split + 1) / I.n_cta;
  const int64_t tile_step = kFused ? I.n_cta : 1;
  const bool has_item = !kWarp0 || cta < P.n_cta_total;
  const int64_t n_my_tiles = (has_item && tile_end > tile_begin) ? (tile_end - tile_begin + tile_step - 1) / tile_step : 0;
  const int64_t n_stage_total = n_my_tiles * 2;     // half-tiles
  const uint32_t a_bytes = (uint32_t)I.a_fgs * 1024u, g_bytes = (uint32_t)I.g_fg * 1024u;
  // ring geometry: stage st holds A at a_base + st * a_stride and G at g_base + st * g_stride
  const uint32_t a_slot = (uint32_t)I.n_mblk * 16384u;
  const bool with_head = kPacked && it == P.head_item;       // CTA-uniform
  const uint32_t packed_stage = a_slot + g_bytes + (with_head ? (uint32_t)kWHeadBytes : 0u);
  const uint32_t h_base = sbase + WSmem::a + a_slot + g_bytes;      // (with_head) + st * packed_stage
  int n_st = kWStages;
  if (kPacked) { n_st = (int)((uint32_t)kWRingBytes / packed_stage); n_st = n_st > kWStages ? kWStages : n_st; }
  const uint32_t a_stride = kPacked ? packed_stage : (uint32_t)kWStageA, g_stride = kPacked ? packed_stage : (uint32_t)kWStageG;
  const uint32_t a_base = sbase + WSmem::a, g_base = kPacked ? a_base + a_slot : sbase + WSmem::g;
  constexpr int kZeroBytes = kPacked ? kWRingBytes : kWStages * kWStageA;

  const long long t_body = (P.variant & 128) ? clock64() : 0;
  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1 + 128); }
    mbar_init(acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (kWarp0 == 0 && warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + WSmem::tmem_ptr), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the MMA reads 16 feature groups per M block even when the image has fewer: keep the tail finite
  for (int i = tid; i < kZeroBytes / 16; i += n_group)
    reinterpret_cast<uint4*>(smem + WSmem::a)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  tc_fence_before();
  cta_sync();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const volatile uint32_t*>(kWarp0 ? shared_tmem : smem + WSmem::tmem_ptr) +
                             (uint32_t)kCol0;

  if (warp == 0) {
    // =============================== producer ===============================
    const bool leader = elect_one();
    auto g_image = [&](int64_t s) {
      const int64_t tile = tile_begin + (s >> 1) * tile_step;
      return P.tape + I.g_off + tile * (int64_t)(I.g_fg * 2048) + (int64_t)(s & 1) * (I.g_fg * 1024);
    };
    // The producer's critical path per stage is: slot free -> issue the two bulk copies.  Everything else of the
    // fused hand-off is done AFTER the issue: dropping the previous occupant's lines from L2 and - one tile ahead -
    // the acquire of the next tile's flag (a ~1 us global round trip even when the flag is long set).
    auto acquire_tile = [&](int64_t s) {
      const int64_t tile = tile_begin + (s >> 1) * tile_step;
      if (lane == 0 && !(P.variant & 16))
        wait_image(P.ready + (int64_t)I.flag_row * P.n_tiles + tile, 1u, 24, (P.variant & 64) != 0);
      __syncwarp();
    };
    const bool prof = (P.variant & 128) != 0;          // bring-up: where does the producer's time go?
    long long t_empty = 0, t_flag = 0, t_disc = 0, t_all = prof ? clock64() : 0;
    if (kFused && n_stage_total > 0) acquire_tile(0);
    // fused: kWStages extra rounds drain the ring so that the last images are discarded too
    int st = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int64_t s = 0; s < n_stage_total + (kFused ? kWStages : 0); ++s) {
      long long c0 = prof ? clock64() : 0;
      mbar_wait(empty(st), ph ^ 1, 20);
      if (prof) { t_empty += clock64() - c0; c0 = clock64(); }
      if (s < n_stage_total && leader) {
        const int64_t tile = tile_begin + (s >> 1) * tile_step;
        const int half = (int)(s & 1);
        const uint8_t* a_src = P.tape + I.a_off + tile * (int64_t)(I.a_fg * 2048) + (int64_t)half * (I.a_fg * 1024) +
                               (int64_t)I.a_fg0 * 1024;
        mbar_arrive_expect_tx(full(st), a_bytes + g_bytes + (with_head ? (uint32_t)kWHeadBytes : 0u));
        bulk_g2s(a_base + st * a_stride, a_src, a_bytes, full(st));
        bulk_g2s(g_base + st * g_stride, g_image(s), g_bytes, full(st));
        if (with_head)
          bulk_g2s(h_base + st * packed_stage, P.tape + P.head_g_off + tile * 4096 + (int64_t)half * 2048, kWHeadBytes, full(st));
      }
      __syncwarp();
      if (kFused && s >= kWStages) {
        // (item, tile) finished: the chains' back-pressure signal
        const int64_t sp = s - kWStages;
        if (P.consumed_total && (sp & 1) == 1 && lane == 0) atomicAdd(P.consumed_total, 1u);
      }
      if (prof) { t_disc += clock64() - c0; c0 = clock64(); }
      if (kFused && (s & 1) == 1 && s + 1 < n_stage_total) acquire_tile(s + 1);
      if (prof) t_flag += clock64() - c0;
      if (++st == n_st) { st = 0; ph ^= 1u; }
    }
    if (prof && lane == 0 && (cta % 37) == 0)
      printf("dw producer cta %d item %d (%d ctas, mblk %d, gfg %d): stages %lld total %lld cyc | wait-empty %lld issue+discard %lld wait-flag %lld\n",
             cta, it, I.n_cta, I.n_mblk, I.g_fg, (long long)n_stage_total, clock64() - t_all, t_empty, t_disc, t_flag);
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = elect_one();
    const int N = I.g_fg * 8;
    // bf16 x bf16 -> fp32, M = 128, both operands MN-major (bits 15 / 16)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
    // canonical MN-major no-swizzle layout: 8 K-rows x 16 B core matrices; K-adjacent core matrices
    // 128 B apart (leading byte offset), MN-adjacent ones 1024 B apart (stride byte offset)
    uint32_t lbo = 128 >> 4, sbo = 1024 >> 4;
    if (P.variant & 1) { const uint32_t x = lbo; lbo = sbo; sbo = x; }
    const uint64_t desc_hi = ((uint64_t)sbo << 32) | (1ull << 46);
    const bool prof = (P.variant & 128) != 0;
    long long t_full = 0;
    int st = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int64_t s = 0; s < n_stage_total; ++s) {
      const long long c0 = prof ? clock64() : 0;
      mbar_wait(full(st), ph, 21);
      if (prof) t_full += clock64() - c0;
      tc_fence_after();
      const uint32_t a_lo = (((a_base + st * a_stride) >> 4) & 0x3FFF) | (lbo << 16);
      const uint32_t g_lo = (((g_base + st * g_stride) >> 4) & 0x3FFF) | (lbo << 16);
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < ((P.variant & 4) ? 0 : 4); ++ks) {        // 64 samples = 4 x K16; a K step = 2 core matrices = 256 B
          for (int mb = 0; mb < I.n_mblk; ++mb)
            mma_ss(tmem_base + (uint32_t)(mb * 256),
                   desc_hi | (uint64_t)(a_lo + (uint32_t)(mb * 16384 + ks * 256) / 16),   // (one M block in the shared-SM kernel)
                   desc_hi | (uint64_t)(g_lo + (uint32_t)(ks * 256) / 16), idesc, (s | ks) ? 1u : 0u);
        }
        tc_commit(empty(st));
      }
      __syncwarp();
      if (++st == n_st) { st = 0; ph ^= 1u; }
    }
    if (leader) tc_commit(acc_bar);
    __syncwarp();
    if (prof && lane == 0 && (cta % 37) == 0) printf("dw issuer   cta %d: wait-full %lld cyc\n", cta, t_full);
  } else if (warp >= 4 && warp < 8) {
    // =============================== bias column sums + final reduction ===============================
    // Column sums on the legacy tensor path: ones[16 x 16] x G[16 samples x 8 features] with mma.sync m16n8k16 gives,
    // in every row of the result, the sum over the 16 samples - four of them per feature group and 64-sample stage,
    // the B fragments straight from the image with ldmatrix.trans (a row of the image = one sample's 8 features =
    // 16 bytes).  Warp w owns the feature groups [8 w, 8 w + 8): 16 LDSM + 32 HMMA per stage instead of the ~270
    // load / unpack / add instructions per thread the CUDA-core form took (it held the stage's slot longer than the MMAs
    // did and cost the kernel 8 %).
    const int j = tid - 128;                    // 0..127
    __shared__ uint32_t s_last;
    const bool has_bias = I.b_out >= 0;
    const int bw = warp - 4;                    // 0..3
    float acc[4][4], hacc[4][4];                // column sums of G; (with_head) fc_alpha's weight gradient
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      hacc[i][0] = hacc[i][1] = hacc[i][2] = hacc[i][3] = 0.f;
    }
    float hbias = 0.f;                          // ... and its bias gradient (lanes 0-3 of the first warp)
    int st = 0;
    uint32_t ph = 0;
#pragma unroll 1
    for (int64_t s = 0; s < n_stage_total; ++s) {
      mbar_wait(full(st), ph, 22);
      if (kFused && !(P.variant & 8)) {
        // The half image is in shared memory now; its global copy is dead once every item that reads it has its own.
        // The LAST reader drops the lines from L2 - they are dirty there and would be written back to HBM otherwise.
        // (Done by these 128 threads, 2 lines each, not by the producer: a discard is ~100s of cycles.)
        const int64_t tile = tile_begin + (s >> 1) * tile_step;
        bool last = true;
        if (I.n_consumers > 1) {
          uint32_t old = 0;
          if (lane == 0 && warp == 4)
            old = atomicAdd(P.consumed + ((int64_t)I.flag_row * P.n_tiles + tile) * 2 + (s & 1), 1u);
          if (warp == 4) s_last = (__shfl_sync(0xffffffffu, old, 0) + 1 == (uint32_t)I.n_consumers) ? 1u : 0u;
          asm volatile("bar.sync 2, 128;" ::: "memory");
          last = s_last != 0;
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
        if (last) {
          const uint8_t* g = P.tape + I.g_off + tile * (int64_t)(I.g_fg * 2048) + (int64_t)(s & 1) * (I.g_fg * 1024);
          for (uint32_t off = (uint32_t)j * 128u; off < g_bytes; off += 128u * 128u) discard_l2(g + off);
        }
      }
      // One mma.sync m16n8k16 per 16 FEATURES and 16 samples: the first operand is the transposed image slice
      // [16 features x 16 samples] (ldmatrix.x4.trans: feature groups 2 j and 2 j + 1, two runs of 8 samples), the second
      // [16 samples x 8] is all ones for the column sums (every column of the result is the sum) or the head image
      // [d rgb, d sigma, 0 ...] for fc_alpha (column 3 of the result).  Warp w owns the feature-group pairs
      // [4 w, 4 w + 4): 16 LDSM + 16 HMMA per stage and image.  A legacy HMMA keeps the SM sub-partition's tensor pipe
      // for ~32 cycles (measured: 64 per warp and stage already exceed the ~2 400 cycles a stage takes to stream), so the
      // form with the features on N (8 per instruction) was twice too expensive to carry the head as well.
      const uint32_t lm_off = (uint32_t)((lane >> 3) & 1) * 1024u + (uint32_t)((lane >> 4) * 8 + (lane & 7)) * 16u;
      auto column_mma = [&](float (&d)[4][4], const uint32_t img, const int n_fg, const bool head, const uint32_t himg) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t b0 = 0x3F803F80u, b1 = 0x3F803F80u;      // (1.0, 1.0) in bf16
          if (head)
            asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                         : "=r"(b0), "=r"(b1) : "r"(himg + (uint32_t)(ks * 16 + (lane & 15)) * 16u));
          uint32_t a[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if ((bw * 4 + i) * 2 < n_fg)              // warp-uniform
              asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                           : "=r"(a[i][0]), "=r"(a[i][1]), "=r"(a[i][2]), "=r"(a[i][3])
                           : "r"(img + (uint32_t)(bw * 4 + i) * 2048u + (uint32_t)ks * 256u + lm_off));
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if ((bw * 4 + i) * 2 < n_fg)
              asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                           "{%0, %1, %2, %3};"
                           : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                           : "r"(a[i][0]), "r"(a[i][1]), "r"(a[i][2]), "r"(a[i][3]), "r"(b0), "r"(b1));
        }
      };
      if (has_bias && !(P.variant & 2)) column_mma(acc, g_base + st * g_stride, I.g_fg, false, 0u);
      if (with_head) {
        const uint32_t himg = h_base + st * packed_stage;
        column_mma(hacc, a_base + st * a_stride, I.a_fgs, true, himg);
        if (bw == 0) {       // fc_alpha's bias gradient: the sum of d sigma (column 3: bytes 6-7 of a sample's 16), two samples per lane
          uint32_t v0, v1;
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v0) : "r"(himg + (uint32_t)lane * 16u + 6u));
          asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v1) : "r"(himg + (uint32_t)(lane + 32) * 16u + 6u));
          hbias += __uint_as_float(v0 << 16) + __uint_as_float(v1 << 16);
        }
      }
      mbar_arrive(empty(st));
      if (++st == n_st) { st = 0; ph ^= 1u; }
    }
    const long long t_loop_end = (P.variant & 128) ? clock64() : 0;
    // result rows g = lane / 4 and g + 8 of feature-group pair j: feature g of groups 2 j and 2 j + 1
    if (has_bias && (lane & 3) == 0) {           // every column holds the sum: take column 0
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int fg = (bw * 4 + i) * 2;
        if (fg < I.g_fg) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int cc = (fg + h) * 8 + (lane >> 2) - I.col0;
            if (cc >= 0 && cc < I.n_cols) red_add_f32(P.grads + I.b_out + cc, acc[i][2 * h]);
          }
        }
      }
    }
    if (with_head) {
      if ((lane & 3) == 1) {                     // column 3 = d sigma: the second value of the lanes with lane % 4 == 1
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int fg = (bw * 4 + i) * 2;
          if (fg < I.a_fgs) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int k = (fg + h) * 8 + (lane >> 2);
              if (k < P.head_rows) red_add_f32(P.grads + P.head_w + k, hacc[i][2 * h + 1]);
            }
          }
        }
      }
      if (bw == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hbias += __shfl_xor_sync(0xffffffffu, hbias, o);
        if (lane == 0) red_add_f32(P.grads + P.head_b, hbias);
      }
    }
    // ---- accumulators -> global (one thread per in-feature row, TMEM lane = row within the M block)
    if (n_stage_total > 0) {
      mbar_wait(acc_bar, 0, 23);
      tc_fence_after();
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const uint32_t lane_base = (uint32_t)(q * 32) << 16;
      for (int mb = 0; mb < I.n_mblk; ++mb) {
        const int r = mb * 128 + row;
        float* dst = P.grads + I.w_out + (int64_t)r * I.ld;
        for (int c = 0; c < I.g_fg * 8; c += 16) {
          float v[16];
          tmem_ld16(tmem_base + lane_base + (uint32_t)(mb * 256 + c), v);      // warp-collective
          if (r < I.a_rows) {
            const int cc0 = c - I.col0;
            if (cc0 >= 0 && cc0 + 16 <= I.n_cols && ((reinterpret_cast<uintptr_t>(dst + cc0) & 15) == 0)) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) red_add_f32x4(dst + cc0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int cc = cc0 + i;
                if (cc >= 0 && cc < I.n_cols) red_add_f32(dst + cc, v[i]);
              }
            }
          }
        }
      }
    }
    if ((P.variant & 128) && tid == 128 && (cta % 37) == 0)
      printf("dw cta %d: start -> stream loop done %lld cyc, column sums + accumulator reduction %lld cyc\n", cta,
             t_loop_end - t_body, clock64() - t_loop_end);
  }

  tc_fence_before();
  cta_sync();
  if ((P.variant & 128) && tid == 128)
    printf("dw cta %3d item %2d (%2d ctas, A %2d fg, G %2d fg): %5lld stages, body %lld cyc\n", cta, it, I.n_cta, I.a_fgs, I.g_fg,
           (long long)n_stage_total, clock64() - t_body);
  if (kWarp0 == 0 && warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

__global__ void __launch_bounds__(kWThreads, 1) mlp_tc_bwd_dw_kernel(const __grid_constant__ DwParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  dw_body<false, 0, 0, WSmemPacked, kWPackedStages, kWStageA, true>(P, smem, (int)blockIdx.x, nullptr);
}

// The training path: chain CTAs [0, n_chain) and weight-gradient CTAs [n_chain, gridDim.x) in ONE launch
// (see the file header).  Chain CTAs carry the low block indices so that they are scheduled first.
template <int H>
__global__ void __launch_bounds__(kBThreads, 1)
mlp_tc_bwd_fused_kernel(const __grid_constant__ BwdParams B, const __grid_constant__ DwParams W, const int n_chain) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if ((int)blockIdx.x < n_chain) chain_body<H, true, 2, kBSlots>(B, smem, (int)blockIdx.x, n_chain);
  else dw_body<true, 0, 0, WSmem, kWStages, kWStageA>(W, smem, (int)blockIdx.x - n_chain, nullptr);
}

// The SAME two roles on EVERY SM (one CTA per SM, 640 threads): warps 0-11 are a one-tile-at-a-time chain (TMEM
// columns 0-255, a 5-slot weight ring), warps 12-19 a weight-gradient GEMM with one M block per item (TMEM columns
// 256-511, two 48 KB stages).  Why: streaming the forward tape for the GEMM needs the bulk-copy engines of ALL SMs
// (a GEMM CTA sustains ~60-80 GB/s, tools/sm_stream_bench.cu; the tape is 4.8 GB per fine pass), and the GEMM's MMAs
// fill the tensor pipe while the chain's single tile is in its epilogue.  Gradient images still go from the chain of
// one CTA to the GEMM of another (an item needs every tile) through L2 with the same flags.
constexpr int kSharedChainWarps = 12;
template <int H>
__global__ void __launch_bounds__(kBThreads, 1)
mlp_tc_bwd_shared_kernel(const __grid_constant__ BwdParams B, const __grid_constant__ DwParams W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using CS = BSmemT<kBSlotsShared, true>;
  uint8_t* smem_dw = smem + (CS::total + 1023) / 1024 * 1024;
  if ((int)(threadIdx.x >> 5) < kSharedChainWarps)
    chain_body<H, true, 1, kBSlotsShared>(B, smem, (int)blockIdx.x, (int)gridDim.x);
  else
    dw_body<true, kSharedChainWarps, 256, WSmemShared, kWStagesShared, kWStageAShared>(W, smem_dw, (int)blockIdx.x,
                                                                                       smem + CS::tmem_ptr);
}

}  // namespace tc
}  // namespace dexnerf

using namespace dexnerf;
using namespace dexnerf::tc;

// chunks of the dX chain in consumption order: step j uses forward layer nl-1-j, N passes of 128
// in-features, K chunks of 64 out-features
static int bwd_chunk_count(const Plan& plan, int H) {
  int n = 0;
  for (int j = 0; j < plan.n_layers - 1; ++j) n += (H / 128) * ((j == 0 ? H / 2 : H) / 64);
  return n;
}

extern "C" DEXNERF_API int64_t dexnerf_tc_packed_bwd_bytes(const dexnerf_flexible_spec* spec) {
  Plan plan;
  if (make_plan(spec, &plan)) return -1;
  return (int64_t)bwd_chunk_count(plan, spec->hidden) * kBSlotBytes;
}

extern "C" DEXNERF_API int dexnerf_tc_pack_bwd(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                                               const float* params, void* packed_t, void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(spec->arch == 0, "tc: the backward kernels exist for FlexibleNeRFModel only");
  DN_REQUIRE(prog && params && packed_t, "tc_pack_bwd: null pointer");
  DN_REQUIRE(prog->n_ops == plan.n_layers + 2, "tc_pack_bwd: program has %d ops, expected %d", prog->n_ops,
             plan.n_layers + 2);
  const int H = spec->hidden;
  BwdPackParams Q{};
  for (int j = 0; j < plan.n_layers - 1; ++j) {
    const int l = plan.n_layers - 1 - j;
    const dexnerf_op& op = prog->ops[plan.layers[l].prog_op];
    const int K = (j == 0) ? H / 2 : H;          // out features of forward layer l
    DN_REQUIRE(op.out_dim == K && op.src0_dim == H, "tc_pack_bwd: op %d is not a hidden layer", plan.layers[l].prog_op);
    for (int p = 0; p < H / 128; ++p)
      for (int c = 0; c < K / 64; ++c) {
        DN_REQUIRE(Q.n_chunks < 200, "tc_pack_bwd: too many chunks");
        BwdPackChunk& pc = Q.chunks[Q.n_chunks++];
        pc.w_off = (int)op.w_off; pc.ld = op.out_dim; pc.n0 = p * 128; pc.k0 = c * 64;
      }
  }
  pack_weights_t_kernel<<<dim3(4, Q.n_chunks), 256, 0, (cudaStream_t)stream>>>(params, Q, reinterpret_cast<uint8_t*>(packed_t));
  DN_CHECK_LAUNCH("pack_weights_t");
  return 0;
}

// Tape layout for `n_samples` samples as 64-bit byte offsets (host array, 4 + 3 * 16 + 3 entries):
// out[0] = tensor-core layer count nl, out[1] = tile count, out[2] = total bytes, out[3] = ghead,
// out[4] = xyz, out[5] = dir, out[6 + l] = act[l], out[22 + l] = mask[l], out[38 + l] = grad[l].
extern "C" DEXNERF_API int dexnerf_tc_tape_layout(const dexnerf_flexible_spec* spec, int64_t n_samples, int64_t* out) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(spec->arch == 0, "tc: the backward kernels exist for FlexibleNeRFModel only");
  DN_REQUIRE(out && n_samples >= 0, "tc_tape_layout: bad argument");
  const int64_t n_pairs = ((n_samples + kTileM - 1) / kTileM + 1) / 2;
  TapeLayout T;
  make_tape_layout(plan, n_pairs * 2, &T);
  out[0] = plan.n_layers; out[1] = T.n_tiles; out[2] = T.total; out[3] = T.ghead; out[4] = T.xyz; out[5] = T.dir;
  for (int l = 0; l < kMaxLayers; ++l) { out[6 + l] = T.act[l]; out[22 + l] = T.mask[l]; out[38 + l] = T.grad[l]; }
  return 0;
}

// what: bit 0 = run the dX chain, bit 1 = run the weight-gradient GEMM (both for a training step)
// variant: bits 0-7 bring-up switches of the GEMM, bits 8-15 the number of SMs the GEMM may use (0 = all; in the
// single-launch kernels the chain's share / the throttle window), bits 16-23 the number of SMs of the chain (0 = all) -
// the training driver runs the fine network's GEMM and the coarse network's chain side by side with them.
extern "C" DEXNERF_API int dexnerf_tc_backward(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                                               const void* packed, const void* packed_t, void* tape,
                                               const float* d_rf, int64_t n, int S, float* grads, int what,
                                               int variant, void* stream) {
  Plan plan;
  if (int rc = make_plan(spec, &plan)) return rc;
  DN_REQUIRE(spec->arch == 0, "tc: the backward kernels exist for FlexibleNeRFModel only");
  DN_REQUIRE(prog && packed && packed_t && tape && d_rf && grads, "tc_backward: null pointer");
  DN_REQUIRE(prog->n_ops == plan.n_layers + 2, "tc_backward: program has %d ops, expected %d", prog->n_ops,
             plan.n_layers + 2);
  DN_REQUIRE((reinterpret_cast<uintptr_t>(tape) & 127) == 0 && (reinterpret_cast<uintptr_t>(d_rf) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(packed_t) & 15) == 0, "tc_backward: misaligned pointer");
  DN_REQUIRE(S >= 1, "tc_backward: S < 1");
  if (n <= 0) return 0;
  const int H = spec->hidden;
  const int nl = plan.n_layers;
  const int64_t m_total = n * (int64_t)S;
  const int64_t n_tiles = (m_total + kTileM - 1) / kTileM, n_pairs = (n_tiles + 1) / 2;
  TapeLayout T;
  make_tape_layout(plan, n_pairs * 2, &T);
  cudaStream_t st = (cudaStream_t)stream;

  const bool shared_sm = what == 8;         // chain + GEMM in every CTA (one CTA per SM)
  const bool fused = (what & 4) != 0 || shared_sm;
  DN_REQUIRE(!fused || what == 4 || what == 8, "tc_backward: what = 4 / 8 (the fused launches) exclude the other bits");
  BwdParams P{};
  if ((what & 1) || fused) {
    P.weights_t = reinterpret_cast<const uint8_t*>(packed_t);
    P.consts = reinterpret_cast<const float*>(packed);
    P.d_rf = reinterpret_cast<const float4*>(d_rf);
    P.tape = reinterpret_cast<uint8_t*>(tape);
    P.m_total = m_total;
    P.nl = nl; P.n_const = plan.n_const; P.off_walpha = plan.off_walpha; P.off_wrgb = plan.off_wrgb;
    for (int l = 0; l < nl; ++l) {
      P.mask_off[l] = T.mask[l]; P.grad_off[l] = T.grad[l]; P.relu[l] = plan.layers[l].tc.relu;
    }
    P.ghead_off = T.ghead;
    P.n_tiles_pad = n_pairs * 2;
    P.ready = fused ? reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(tape) + T.flags) : nullptr;
  }
  if (what & 1) {
    int sms = (variant >> 16) & 0xFF;
    if (sms <= 0 || sms > kNumSMs) sms = kNumSMs;
    const int grid = (int)(n_pairs < sms ? n_pairs : sms);
    const size_t smem = BSmem::total + 1024;
    auto launch = [&](auto kernel) -> int {
      DN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kernel<<<grid, kBThreads, smem, st>>>(P);
      return 0;
    };
    const int rc = (H == 256) ? launch(mlp_tc_bwd_dx_kernel<256>) : launch(mlp_tc_bwd_dx_kernel<128>);
    if (rc) return rc;
    DN_CHECK_LAUNCH("mlp_tc_bwd_dx");
  }

  if ((what & 2) || fused) {
    DwParams W{};
    W.tape = reinterpret_cast<const uint8_t*>(tape);
    W.grads = grads;
    W.n_tiles = n_pairs * 2;
    W.variant = variant & 0xFF;
    double cost[kMaxDwItems];
    auto add1 = [&](int64_t a_off, int a_fg, int a_fg0, int a_fgs, int a_rows, int64_t g_off, int g_fg, int64_t w_out,
                    int64_t b_out, int col0, int n_cols, int ld, int flag_row) {
      if (W.n_items >= kMaxDwItems) { ++W.n_items; return; }
      DwItem& I = W.items[W.n_items];
      I.flag_row = flag_row; I.n_consumers = 1;
      I.a_off = a_off; I.a_fg = a_fg; I.a_fg0 = a_fg0; I.a_fgs = a_fgs; I.n_mblk = (a_fgs + 15) / 16;
      I.a_rows = a_rows; I.g_off = g_off; I.g_fg = g_fg; I.w_out = w_out; I.b_out = b_out;
      I.col0 = col0; I.n_cols = n_cols; I.ld = ld;
      // stand-alone kernel: bytes streamed per tile (HBM-bound); fused launch: tensor-pipe cycles per tile
      // (n_mblk M blocks x 8 K steps, an N-wide MMA takes ~max(N, 64) / 2 cycles) - the G half comes out of L2
      cost[W.n_items] = fused ? (double)(I.n_mblk * 8 * ((g_fg * 8 > 64 ? g_fg * 8 : 64) / 2)) + 200.0
                              : (double)(a_fgs + g_fg);
      ++W.n_items;
    };
    // the shared-SM kernel accumulates ONE M block (128 in-features, 256 TMEM columns) per item: wider operands
    // become two items that read the same G image and halves of the A image
    auto add = [&](int64_t a_off, int a_fg, int a_fg0, int a_fgs, int a_rows, int64_t g_off, int g_fg, int64_t w_out,
                   int64_t b_out, int col0, int n_cols, int ld, int flag_row) {
      if (!shared_sm || a_fgs <= 16) {
        add1(a_off, a_fg, a_fg0, a_fgs, a_rows, g_off, g_fg, w_out, b_out, col0, n_cols, ld, flag_row);
        return;
      }
      add1(a_off, a_fg, a_fg0, 16, a_rows < 128 ? a_rows : 128, g_off, g_fg, w_out, b_out, col0, n_cols, ld, flag_row);
      add1(a_off, a_fg, a_fg0 + 16, a_fgs - 16, a_rows - 128, g_off, g_fg, w_out + (int64_t)128 * ld, -1, col0, n_cols, ld,
           flag_row);
    };
    const int hfg = H / 8;
    W.head_item = -1;
    // stand-alone kernel: fc_alpha rides on the item that streams the same activations (DwParams::head_item)
    const bool merge_alpha = !fused && !(variant & 32);
    for (int l = 0; l < nl; ++l) {
      const TcLayer& L = plan.layers[l].tc;
      const dexnerf_op& op = prog->ops[plan.layers[l].prog_op];
      const int gfg = L.n_out / 8;
      int64_t w_row = op.w_off;
      bool bias_done = false;
      if (L.k_main) {      // hidden-input part: rows [0, H) of Wt
        if (merge_alpha && l - 1 == nl - 3 && W.head_item < 0 && W.n_items < kMaxDwItems) W.head_item = W.n_items;
        add(T.act[l - 1], hfg, 0, hfg, H, T.grad[l], gfg, w_row, op.b_off, 0, L.n_out, L.n_out, l);
        bias_done = true;
        w_row += (int64_t)H * L.n_out;
      }
      if (L.smem_src == 1)
        add(T.xyz, 8, 0, 8, spec->dim_xyz, T.grad[l], gfg, w_row, bias_done ? -1 : op.b_off, 0, L.n_out, L.n_out, l);
      else if (L.smem_src == 2)
        add(T.dir, 4, 0, 4, spec->dim_dir, T.grad[l], gfg, w_row, bias_done ? -1 : op.b_off, 0, L.n_out, L.n_out, l);
    }
    {   // heads: fc_alpha reads the last trunk output, fc_rgb the dir-layer output; G = [d rgb, d sigma, 0..]
      const dexnerf_op& oa = prog->ops[plan.op_alpha];
      const dexnerf_op& orgb = prog->ops[plan.op_rgb];
      if (W.head_item >= 0) {
        W.head_rows = H; W.head_g_off = T.ghead; W.head_w = oa.w_off; W.head_b = oa.b_off;
        // its CTAs take ~2 700 cycles per stage where the others stream theirs in ~2 250 (per-CTA clocks, tools/dw_profile.py):
        // the 32 legacy HMMAs per warp and stage of the column-sum warps are what paces them
        cost[W.head_item] = (cost[W.head_item] + 1.0) * 1.2;
      } else {
        add(T.act[nl - 3], hfg, 0, hfg, H, T.ghead, 2, oa.w_off, oa.b_off, 3, 1, 1, kMaxLayers);
      }
      add(T.act[nl - 1], hfg / 2, 0, hfg / 2, H / 2, T.ghead, 2, orgb.w_off, orgb.b_off, 0, 3, 3, kMaxLayers);
    }
    DN_REQUIRE(W.n_items <= kMaxDwItems, "tc_backward: too many weight-gradient items");
    for (int i = 0; i < W.n_items; ++i) {
      int readers = 0;
      for (int k = 0; k < W.n_items; ++k) readers += W.items[k].flag_row == W.items[i].flag_row;
      W.items[i].n_consumers = readers;
    }
    // ONE wave of CTAs (one per SM).  Every CTA of item i streams cost[i] * n_tiles / share[i] bytes;
    // the kernel ends with the slowest CTA, so shares are chosen greedily to minimise that maximum:
    // start with one CTA per item and keep giving a CTA to the currently most loaded item.
    // Fused launch: the SMs are first divided between the chain and the GEMM in proportion to their tensor-pipe
    // cycles per tile (variant >> 8 overrides the chain's share for experiments).
    int n_chain = 0;
    if (fused && !shared_sm) {
      double chain_cost = 0.0, dw_cost = 0.0;
      for (int j = 0; j < nl - 1; ++j) chain_cost += (double)(H / 128) * ((j == 0 ? H / 2 : H) / 64) * 4 * 64 + 150.0 * (H / 128);
      for (int i = 0; i < W.n_items; ++i) dw_cost += cost[i];
      n_chain = (int)(kNumSMs * chain_cost / (chain_cost + dw_cost) + 0.5);
      if (((variant >> 8) & 0xFF) > 0) n_chain = (variant >> 8) & 0xFF;
      if (n_chain > kNumSMs - W.n_items) n_chain = kNumSMs - W.n_items;
      if (n_chain > n_pairs) n_chain = (int)n_pairs;
      if (n_chain < 1) n_chain = 1;
    }
    int64_t budget = kNumSMs - n_chain;      // (all SMs in the shared-SM kernel: n_chain stays 0 there)
    if (!fused && ((variant >> 8) & 0xFF) > 0) budget = (variant >> 8) & 0xFF;      // the stand-alone GEMM on fewer SMs
    if (budget < W.n_items) budget = W.n_items;
    int share[kMaxDwItems];
    for (int i = 0; i < W.n_items; ++i) share[i] = 1;
    for (int64_t used = W.n_items; used < budget; ++used) {
      int best = -1;
      for (int i = 0; i < W.n_items; ++i)
        if (share[i] < W.n_tiles && (best < 0 || cost[i] / share[i] > cost[best] / share[best])) best = i;
      if (best < 0) break;
      ++share[best];
    }
    int cta = 0;
    for (int i = 0; i < W.n_items; ++i) {
      int64_t k = share[i];
      if (k > W.n_tiles) k = W.n_tiles;
      W.items[i].cta0 = cta; W.items[i].n_cta = (int)k;
      cta += (int)k;
    }
    W.n_cta_total = cta;
    if (fused) {
      uint8_t* flags = reinterpret_cast<uint8_t*>(tape) + T.flags;
      W.ready = reinterpret_cast<const uint32_t*>(flags);
      W.consumed = reinterpret_cast<uint32_t*>(flags) + (int64_t)(kMaxLayers + 1) * W.n_tiles;
      W.variant = variant & 0xFF;
      DN_CUDA(cudaMemsetAsync(flags, 0, (size_t)T.flag_bytes, st));
      if (shared_sm && !(variant & 32)) {        // (variant bit 5: no back-pressure, for experiments)
        W.consumed_total = reinterpret_cast<uint32_t*>(flags + T.flag_bytes - 128);
        P.consumed_total = W.consumed_total;
        P.throttle_items = W.n_items;
        P.throttle_window = ((variant >> 8) & 0xFF) > 0 ? ((variant >> 8) & 0xFF) : kNumSMs + 24;
      }
      if (shared_sm) {
        using CS = BSmemT<kBSlotsShared, true>;
        const size_t smem = (CS::total + 1023) / 1024 * 1024 + WSmemShared::total + 1024;
        static_assert((BSmemT<kBSlotsShared, true>::total + 1023) / 1024 * 1024 + WSmemShared::total + 1024 <= 227 * 1024,
                      "the two role groups must fit one SM");
        auto launch = [&](auto kernel) -> int {
          DN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          kernel<<<kNumSMs, kBThreads, smem, st>>>(P, W);
          return 0;
        };
        const int rc = (H == 256) ? launch(mlp_tc_bwd_shared_kernel<256>) : launch(mlp_tc_bwd_shared_kernel<128>);
        if (rc) return rc;
        DN_CHECK_LAUNCH("mlp_tc_bwd_shared");
        return 0;
      }
      const size_t smem = (WSmem::total > BSmem::total ? WSmem::total : BSmem::total) + 1024;
      auto launch = [&](auto kernel) -> int {
        DN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<n_chain + cta, kBThreads, smem, st>>>(P, W, n_chain);
        return 0;
      };
      const int rc = (H == 256) ? launch(mlp_tc_bwd_fused_kernel<256>) : launch(mlp_tc_bwd_fused_kernel<128>);
      if (rc) return rc;
      DN_CHECK_LAUNCH("mlp_tc_bwd_fused");
      return 0;
    }
    const size_t smem = WSmemPacked::total + 1024;
    DN_CUDA(cudaFuncSetAttribute(mlp_tc_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_tc_bwd_dw_kernel<<<cta, kWThreads, smem, st>>>(W);
    DN_CHECK_LAUNCH("mlp_tc_bwd_dw");
  }
  return 0;
}
