// Layer table of the tensor-core FlexibleNeRFModel kernels and the layout of the training tape,
// shared by mlp_tc.cu (forward) and mlp_tc_bwd.cu (backward).
#pragma once
#include "common.cuh"

namespace dexnerf {
namespace tc {

constexpr int kMaxLayers = 16;
constexpr int kMaxConstFloats = 4096;

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

struct HostLayer { TcLayer tc; int prog_op; int smem_first; };   // smem_first: the encoding precedes the hidden inputs in the weight's K order

struct Plan {
  int n_layers = 0;
  HostLayer layers[kMaxLayers];
  int op_alpha = -1, op_rgb = -1;
  int off_walpha = 0, off_balpha = 0, off_wrgb = 0, off_brgb = 0, n_const = 0;
  int64_t weight_bytes = 0;
  int kx = 0, kd = 0;
};

inline int pad16(int v) { return (v + 15) / 16 * 16; }

inline int make_plan(const dexnerf_flexible_spec* s, Plan* plan) {
  DN_REQUIRE(s, "tc: null spec");
  DN_REQUIRE(s->arch >= 0 && s->arch <= 2, "tc: unknown architecture %d", s->arch);
  DN_REQUIRE(s->hidden == 256 || (s->hidden == 128 && s->arch != 1), "tc: hidden must be 128 or 256 (got %d)", s->hidden);
  DN_REQUIRE(s->arch == 1 || (s->n_trunk >= 1 && s->n_trunk + 3 <= kMaxLayers), "tc: unsupported trunk depth %d", s->n_trunk);
  DN_REQUIRE(s->arch == 1 || s->skip_every >= 1, "tc: skip_every < 1");
  DN_REQUIRE(s->dim_xyz >= 1 && s->dim_xyz <= 64, "tc: dim_xyz must be <= 64 (got %d)", s->dim_xyz);
  DN_REQUIRE(s->arch == 2 ? s->dim_dir == 0 : (s->dim_dir >= 1 && s->dim_dir <= 32),
             "tc: dim_dir must be in 1..32 (0 without view directions), got %d", s->dim_dir);
  Plan& P = *plan;
  const int H = s->hidden;
  P.kx = pad16(s->dim_xyz);
  P.kd = pad16(s->dim_dir);
  int bias = 0, op = 0;
  auto add = [&](int k_main, int src, int k_smem, int n_out, int relu, int head, int prog_op, int smem_first = 0) {
    HostLayer& L = P.layers[P.n_layers++];
    L.tc.k_main = k_main; L.tc.smem_src = src; L.tc.k_smem = k_smem; L.tc.n_out = n_out;
    L.tc.n_pass = (n_out + 127) / 128; L.tc.relu = relu; L.tc.head = head; L.tc.bias_off = bias;
    L.prog_op = prog_op;
    L.smem_first = smem_first;
    bias += (n_out + 127) / 128 * 128;
    const int np = n_out < 128 ? n_out : 128;
    P.weight_bytes += (int64_t)L.tc.n_pass * np * (k_main + (src ? k_smem : 0)) * 2;
  };
  if (s->arch == 1) {
    // PaperNeRFModel (nerf/models.py:123-182, repaired forward): 8 x 256 trunk from xyz, cat((xyz, x)) into
    // layer 4, feat = fc_feat(x) WITHOUT ReLU, alpha = fc_alpha(feat), then cat((feat, dirs)) ->
    // 128 -> 128 -> 128 (ReLU each) -> fc_rgb.  Program ops: xyz0..7, fc_feat, fc_alpha, dir0..2, fc_rgb.
    add(0, 1, P.kx, 256, 1, 0, op++);                                 // layers_xyz[0] (with ReLU)
    for (int i = 1; i < 8; ++i) add(256, i == 4 ? 1 : 0, i == 4 ? P.kx : 0, 256, 1, 0, op++, i == 4);
    add(256, 0, 0, 256, 0, 1, op++);                                  // fc_feat (no ReLU) + sigma head on feat
    P.op_alpha = op++;
    add(256, 2, P.kd, 128, 1, 0, op++);                               // layers_dir[0]
    add(128, 0, 0, 128, 1, 0, op++);                                  // layers_dir[1]
    add(128, 0, 0, 128, 1, 2, op++);                                  // layers_dir[2] (+ fc_rgb head)
    P.op_rgb = op++;
  } else if (s->arch == 2) {
    // FlexibleNeRFModel(use_viewdirs=False) (nerf/models.py:250-256): trunk, then fc_out (H -> 4 = rgb, sigma)
    // evaluated on the CUDA cores in the last trunk layer's epilogue (head 3).  Program ops: layer1, trunk, fc_out.
    add(0, 1, P.kx, H, 0, 0, op++);
    for (int i = 0; i < s->n_trunk; ++i) {
      const bool skip = (i % s->skip_every == 0) && i > 0;
      add(H, skip ? 1 : 0, skip ? P.kx : 0, H, 1, i == s->n_trunk - 1 ? 3 : 0, op++);
    }
    P.op_alpha = -1;
    P.op_rgb = op++;                                                   // fc_out
    P.off_walpha = bias;
    P.off_balpha = bias;
    P.off_wrgb = bias; bias += 4 * H;                                  // W_out as [4][H]
    P.off_brgb = bias; bias += 4;
    P.n_const = bias;
    DN_REQUIRE(P.n_const <= kMaxConstFloats, "tc: const block too large");
    return 0;
  } else {
    add(0, 1, P.kx, H, 0, 0, op++);                                    // layer1 (no ReLU)
    for (int i = 0; i < s->n_trunk; ++i) {
      const bool skip = (i % s->skip_every == 0) && i > 0;
      add(H, skip ? 1 : 0, skip ? P.kx : 0, H, 1, i == s->n_trunk - 1 ? 1 : 0, op++);
    }
    P.op_alpha = op++;
    add(H, 0, 0, H, 1, 0, op++);                                        // fc_feat
    add(H, 2, P.kd, H / 2, 1, 2, op++);                                 // layers_dir[0] (+ fc_rgb head)
    P.op_rgb = op++;
  }
  P.off_walpha = bias; bias += H;
  P.off_balpha = bias; bias += 4;
  P.off_wrgb = bias; bias += 3 * (H / 2);
  P.off_brgb = bias; bias += 4;
  P.n_const = bias;
  DN_REQUIRE(P.n_const <= kMaxConstFloats, "tc: const block too large");
  return 0;
}


// Training tape of one query of n_tiles 128-sample tiles.  Every IMAGE is a bf16 matrix
// [128 samples x F features] per tile stored as [half (2)][F/8 feature groups][64 samples][8
// features] - 16-byte rows, 128-byte core matrices: the MN-major no-swizzle UMMA operand layout
// with the samples as the K dimension (LBO = 128 B between K core matrices, SBO = 1024 B between
// feature groups), so the weight-gradient GEMM streams half-tiles with plain bulk copies.
// Every MASK is [tile][slot][128 samples] x 64 ReLU bits (slot = 64-column block of the layer).
struct TapeLayout {
  int64_t n_tiles = 0;
  int64_t xyz = 0, dir = 0;            // encodings: 64 and 32 features (zero padded)
  int64_t act[kMaxLayers] = {};        // output of tensor-core layer l
  int64_t mask[kMaxLayers] = {};       // its ReLU bits
  int64_t grad[kMaxLayers] = {};       // dL/d(pre-activation) of layer l (written by the backward)
  int64_t ghead = 0;                   // [d rgb (3), d sigma, 0 x 12]: 16 features
  int64_t flags = 0;                   // fused backward: uint32 ready[kMaxLayers + 1][n_tiles], consumed[..][n_tiles][2]
  int64_t flag_bytes = 0;
  int64_t total = 0;
};

inline void make_tape_layout(const Plan& plan, int64_t n_tiles, TapeLayout* out) {
  TapeLayout& T = *out;
  T.n_tiles = n_tiles;
  int64_t off = 0;
  auto take = [&](int64_t bytes_per_tile) { const int64_t o = off; off += bytes_per_tile * n_tiles; return o; };
  T.xyz = take(64 * 256);
  T.dir = take(32 * 256);
  for (int l = 0; l < plan.n_layers; ++l) {
    const TcLayer& L = plan.layers[l].tc;
    T.act[l] = take((int64_t)L.n_out * 256);
    T.mask[l] = take((int64_t)L.n_pass * 2 * 1024);
    T.grad[l] = take((int64_t)L.n_out * 256);
  }
  T.ghead = take(16 * 256);
  T.flag_bytes = (int64_t)(kMaxLayers + 1) * n_tiles * 3 * 4 + 128;    // + the consumed-total counter (last 128 B)
  T.flags = off;
  off += (T.flag_bytes + 127) / 128 * 128;
  T.total = off;
}

}  // namespace tc
}  // namespace dexnerf
