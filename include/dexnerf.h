/* dexnerf.h - C ABI of the B200-native Dex-NeRF ray-render hot path.
 *
 * The reference (edwardyang12/Dex-NERF, nerf-pytorch/) is pure Python/PyTorch and has NO FFI,
 * plugin or operator interface of its own: its boundary is the Python namespace `nerf`
 * (nerf/__init__.py:1-8).  This header is therefore the interface a maintainer would bind
 * UNDERNEATH those Python functions (ctypes stub in INTEGRATION.md); each entry point names the
 * reference function (file:line under nerf-pytorch/) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 (int64 where stated) unless
 *     the parameter is documented as host;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, nothing
 *     synchronises, nothing allocates (callers pass outputs and workspaces);
 *   - return value 0 on success, negative DEXNERF_E_* otherwise; dexnerf_last_error() returns a
 *     thread-local message for the last failure;
 *   - there is no CPU implementation behind any entry point.
 */
#ifndef DEXNERF_H_
#define DEXNERF_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DEXNERF_API __attribute__((visibility("default")))
#else
#define DEXNERF_API
#endif

#define DEXNERF_ABI_VERSION 2
#define DEXNERF_E_INVALID (-1) /* bad argument (null pointer, unsupported size) */
#define DEXNERF_E_CUDA (-2)    /* a CUDA runtime call or launch failed */
#define DEXNERF_E_UNSUPPORTED (-3)

DEXNERF_API int dexnerf_abi_version(void);
DEXNERF_API const char* dexnerf_last_error(void);

/* ---- a-1  get_ray_bundle (nerf/nerf_helpers.py:67-112) + meshgrid_xy (:28-40)
 * T_w2c: 4x4 world->cam (inverted in-kernel), K: 3x3.  Both pixel axes are divided by K[0][0]
 * (reference quirk, :100-101).  Produces rows [row0, row0+rows) of the H x W bundle:
 * ro, rd: (rows, W, 3). */
DEXNERF_API int dexnerf_ray_bundle(const float* T_w2c, const float* K, int H, int W, int row0, int rows,
                       float* ro, float* rd, void* stream);

/* ---- ndc_rays (nerf/nerf_helpers.py:172-199); n rays, in/out (n,3). */
DEXNERF_API int dexnerf_ndc_rays(const float* ro, const float* rd, int64_t n, int H, int W, float focal,
                     float near, float* ro_out, float* rd_out, void* stream);

/* ---- a-2  positional_encoding (nerf/nerf_helpers.py:115-159); x (M,3) -> out (M, 3*inc + 6L) */
DEXNERF_API int dexnerf_positional_encoding(const float* x, int64_t M, int L, int include_input,
                                int log_sampling, float* out, void* stream);

/* ---- a-4  stratified depths (nerf/train_utils.py:104-133).  near/far per ray come from
 * rays[:, 6:8] in the reference; here they are scalars or per-ray arrays (near_arr/far_arr may be
 * NULL -> scalars used).  t_rand (n,Nc) NULL -> no perturbation.  z: (n,Nc). */
DEXNERF_API int dexnerf_stratified_z(int64_t n, int Nc, float near, float far, const float* near_arr,
                         const float* far_arr, int lindisp, const float* t_rand, float* z,
                         void* stream);

/* ---- a-5  cumprod_exclusive (nerf/nerf_helpers.py:43-64); x,out (n,S). */
DEXNERF_API int dexnerf_cumprod_exclusive(const float* x, int64_t n, int S, float* out, void* stream);

/* ---- a-6  volume_render_radiance_field (nerf/volume_rendering_utils.py:6-70), including the
 * Dex-NeRF first-crossing depth (:51-58).
 * rf (n,S,4) raw network output, z (n,S), rd (n,3), noise (n,S) pre-scaled N(0,std) or NULL,
 * thresholds: T floats (device).  Outputs (any may be NULL): rgb (n,3), disp (n), acc (n),
 * weights (n,S), depth (n), dex_depth (T,n), dex_index (T,n) int64. */
DEXNERF_API int dexnerf_volume_render(const float* rf, const float* z, const float* rd, const float* noise,
                          int64_t n, int S, int white_background, const float* thresholds, int T,
                          float* rgb, float* disp, float* acc, float* weights, float* depth,
                          float* dex_depth, int64_t* dex_index, void* stream);

/* ---- a-7  sample_pdf == sample_pdf_2 (nerf/nerf_helpers.py:262-304) with the external
 * torchsearchsorted.searchsorted(side="right") (requirements.txt:9) folded in.
 * bins (n,B), weights (n,B-1), u (n,Nf) or NULL (=> det: linspace(0,1,Nf)).
 * samples (n,Nf); inds (n,Nf) int64 or NULL. */
DEXNERF_API int dexnerf_sample_pdf(const float* bins, const float* weights, int64_t n, int B, int Nf,
                       const float* u, float* samples, int64_t* inds, void* stream);

/* ---- a-7 + a-8 fused: mid-points, weights[...,1:-1], sample_pdf, cat + sort
 * (nerf/train_utils.py:163-173).  z_coarse (n,Nc), weights (n,Nc) -> z_fine (n,Nc+Nf) sorted. */
DEXNERF_API int dexnerf_resample_merge(const float* z_coarse, const float* weights, int64_t n, int Nc, int Nf,
                           const float* u, float* z_fine, void* stream);

/* ---- a-3  the MLPs (nerf/models.py) as a small layer program.
 * Each op computes dst = act(W . cat(src0, src1) + b).  Weights are packed by the host as
 * Wt[in][out] fp32 (transposed nn.Linear.weight) followed anywhere by the bias. */
enum {
  DEXNERF_ENC_XYZ = 0, DEXNERF_ENC_DIR = 1, DEXNERF_BUF_A = 2, DEXNERF_BUF_B = 3,
  DEXNERF_OUT_RGB = 4,   /* 3 channels -> out[:, 0:3] */
  DEXNERF_OUT_SIGMA = 5, /* 1 channel  -> out[:, 3]   */
  DEXNERF_OUT_ALL = 6,   /* 4 channels -> out[:, 0:4] */
  DEXNERF_NONE = -1
};
#define DEXNERF_MAX_OPS 16
typedef struct {
  int32_t src0, src0_dim, src1, src1_dim; /* src1 = DEXNERF_NONE when there is no concat */
  int32_t dst, out_dim, relu, pad_;
  int64_t w_off, b_off; /* offsets in floats into `params` */
} dexnerf_op;
typedef struct {
  int32_t n_ops, dim_xyz, dim_dir, max_width;
  int32_t Lx, Ld, include_xyz, include_dir, log_xyz, log_dir; /* encoders, query mode */
  int32_t pad_[2];
  dexnerf_op ops[DEXNERF_MAX_OPS];
} dexnerf_mlp_program;

/* model(x): x (M, dim_xyz + dim_dir) already encoded -> out (M,4).  fp32 CUDA-core path. */
DEXNERF_API int dexnerf_mlp_forward(const dexnerf_mlp_program* prog /*host*/, const float* params,
                        const float* x, int64_t M, float* out, void* stream);

/* run_network (nerf/train_utils.py:72-89) fused with the point construction of :136/:177:
 * pts = ro + rd * z, encode (xyz and broadcast view dir), MLP.  ro, rd (n,3), viewdirs (n,3)
 * or NULL, z (n,S) -> rf (n,S,4).  precision 0: fp32 CUDA cores. */
DEXNERF_API int dexnerf_mlp_query(const dexnerf_mlp_program* prog /*host*/, const float* params,
                      const float* ro, const float* rd, const float* viewdirs, const float* z,
                      int64_t n, int S, float* rf, void* stream);

/* ---- tensor-core path (sm_100a tcgen05): FlexibleNeRFModel (nerf/models.py:185-256, inference and training)
 * and PaperNeRFModel (nerf/models.py:123-182, arch = 1: hidden 256, inference only).
 * Packing converts nn.Linear weights into the bf16 UMMA shared-memory images the kernel streams
 * with bulk TMA.  See dex-nerf_b200/csrc/mlp_tc.cu. */
typedef struct {
  int32_t hidden;     /* 128 or 256 */
  int32_t n_trunk;    /* number of layers_xyz (num_layers - 1) */
  int32_t skip_every; /* skip_connect_every */
  int32_t dim_xyz, dim_dir, Lx, Ld, include_xyz, include_dir, log_xyz, log_dir;
  int32_t arch;       /* 0: FlexibleNeRFModel; 1: PaperNeRFModel (hidden must be 256; n_trunk / skip_every ignored);
                         2: FlexibleNeRFModel(use_viewdirs=False): dim_dir = 0, viewdirs may be NULL */
} dexnerf_flexible_spec;
/* size in bytes of the packed weight blob for `spec` (negative on error) */
DEXNERF_API int64_t dexnerf_tc_packed_bytes(const dexnerf_flexible_spec* spec /*host*/);
/* params: the fp32 program-layout buffer of the same model (as for dexnerf_mlp_query);
 * prog: the program it was built from.  packed: device blob of dexnerf_tc_packed_bytes().  Packing is
 * stream-ordered like every other entry point (the layout tables travel as kernel parameters; ABI 2 dropped the
 * unused workspace argument of ABI 1). */
DEXNERF_API int dexnerf_tc_pack(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                    const float* params, void* packed, void* stream);
/* run_network on tensor cores: ro, rd, viewdirs (n,3), z (n,S) -> rf (n,S,4).
 * dbg (optional, may be NULL): raw fp32 accumulator of (dbg_layer, dbg_pass), [n*S rounded up to
 * 128][128] floats - used by the kernel's own tests. */
DEXNERF_API int dexnerf_tc_query(const dexnerf_flexible_spec* spec, const void* packed, const float* ro,
                     const float* rd, const float* viewdirs, const float* z, int64_t n, int S,
                     float* rf, float* dbg, int dbg_layer, int dbg_pass, void* stream);

/* ---- training (BASELINE config 4; train_dexnerf_rgb.py:246-281 = forward, loss.backward()).
 *
 * dexnerf_volume_render_backward: backward of volume_render_radiance_field w.r.t. the radiance
 * field.  g_rgb (n,3), g_depth (n), g_acc (n) are dL/d(rgb_map, depth_map, acc_map) (any may be
 * NULL = zero); d_rf (n,S,4) receives dL/d(radiance_field).  rf, z, rd, noise, white_background as
 * in the forward call. */
DEXNERF_API int dexnerf_volume_render_backward(const float* rf, const float* z, const float* rd,
                                   const float* noise, int64_t n, int S, int white_background,
                                   const float* g_rgb, const float* g_depth, const float* g_acc,
                                   float* d_rf, void* stream);

/* Tensor-core training path of FlexibleNeRFModel.  The forward variant records a TAPE (bf16
 * operand images of every layer + ReLU bits, dexnerf_tc_tape_bytes() bytes, 128-byte aligned);
 * the backward consumes it.  See dex-nerf_b200/csrc/mlp_tc_bwd.cu. */
DEXNERF_API int64_t dexnerf_tc_tape_bytes(const dexnerf_flexible_spec* spec /*host*/, int64_t n_samples);
/* host-side description of the tape (offsets in bytes; see mlp_tc_bwd.cu); out: 54 int64 */
DEXNERF_API int dexnerf_tc_tape_layout(const dexnerf_flexible_spec* spec /*host*/, int64_t n_samples,
                           int64_t* out /*host*/);
DEXNERF_API int dexnerf_tc_query_train(const dexnerf_flexible_spec* spec, const void* packed, const float* ro,
                           const float* rd, const float* viewdirs, const float* z, int64_t n, int S,
                           float* rf, void* tape, void* stream);
/* transposed bf16 weight images for the activation-gradient chain */
DEXNERF_API int64_t dexnerf_tc_packed_bwd_bytes(const dexnerf_flexible_spec* spec /*host*/);
DEXNERF_API int dexnerf_tc_pack_bwd(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                        const float* params, void* packed_t, void* stream);
/* d_rf (n,S,4) = dL/d(raw rgb, sigma) per sample -> grads: fp32 buffer in the program layout of
 * `params` (Wt[in][out] | bias per op), ACCUMULATED into (zero it first).  packed = forward blob,
 * packed_t = dexnerf_tc_pack_bwd blob.  what: bit 0 activation-gradient chain, bit 1 weight-gradient GEMM (3 = a
 * full backward: the two kernels one after the other, the gradient images make one HBM round trip - the training
 * path).  Two single-launch forms in which the images go from the chain to the GEMM through L2 with release /
 * acquire flags (same results; slower than 3 on a B200 today, kept selectable - DESIGN.md section 3.2): 4 = chain and
 * GEMM on disjoint SMs, 8 = chain and GEMM as two warp groups of every CTA (one CTA per SM, with back-pressure).
 * variant: 0.  Bits 8-15: SMs the weight-gradient GEMM may use (what = 2; 0 = all), bits 16-23: SMs of the chain
 * (what = 1; 0 = all) - the results do not depend on either (the chain's images bit for bit, the GEMM's sums up to
 * their fp32 order).  (Bring-up / experiment bits: 1 swaps the GEMM's descriptor strides, 2 skips the bias column
 * sums, 4 skips the MMAs - results invalid; 32 (what = 2) keeps fc_alpha's gradient a work item of its own instead
 * of folding it into the feature layer's; 128 prints per-CTA clocks of the GEMM; single-launch forms: 8 no L2
 * discard, 16 no flag waits - results invalid, 32 no back-pressure, bits 8-15 chain CTAs (what = 4) / the
 * back-pressure window in tiles (what = 8).) */
DEXNERF_API int dexnerf_tc_backward(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                        const void* packed, const void* packed_t, void* tape, const float* d_rf,
                        int64_t n, int S, float* grads, int what, int variant, void* stream);

/* ---- a-9  the render driver as ONE call per ray chunk (nerf/train_utils.py:92-202 predict_and_render_radiance
 * driven by :205-288 run_one_iter_of_nerf).  The reference runs ~200 eager kernels per chunk; here the host side
 * of the library sequences the kernels of this header on `stream` with no torch glue in between:
 *
 *   ray setup (one launch: optional ray generation from a camera, view directions rd/|rd| (:222-226), the
 *     stratified depths (:111-133) and - in train mode - the Philox draws t_rand / u / sigma noise)
 *   [ndc_rays first when `ndc` is set (:238-242; near is 1.0 as in the reference)]
 *   -> coarse MLP query -> coarse compositing -> resample + merge -> fine MLP query -> fine compositing (+ Dex depth)
 *
 * Rays: explicit arrays ro, rd (n,3) - or ro == NULL and a camera (T_w2c, K, H, W, row0, rows; n = rows * W), in
 * which case get_ray_bundle (a-1) is part of the setup launch and a frame is 6 launches in total.
 * RNG: t_rand (n,Nc), u (n,Nf), noise_coarse (n,Nc), noise_fine (n,Nc+Nf) replay given draws (parity tests; the
 * noises are pre-scaled by the std); a NULL pointer means "draw in the setup kernel" with Philox4x32-10
 * keyed by (seed, offset) when perturb / noise_std ask for it (perturb: jittered depths AND random u,
 * train_utils.py:126-133,169).
 * Models: `spec`/`packed` select the tensor-core path (bf16 operands), spec == NULL the fp32 CUDA-core path
 * (`prog`/`params`).  tape_coarse / tape_fine (dexnerf_tc_tape_bytes each; tensor-core path only) make the two
 * queries record the training tape; NULL = inference.
 * Workspace: dexnerf_render_workspace_bytes(n, Nc, Nf) bytes, 256-byte aligned; dexnerf_render_workspace_layout
 * describes where the intermediates (rays, view directions, depths, radiance fields, weights, noises) live, for
 * dexnerf_render_fused_bwd and for tests.
 * Outputs (device, any may be NULL): rgb (n,3), depth (n) [the EXPECTED depth, train_utils.py:201], acc (n) of both
 * passes; dex_fine: T planes, plane t at dex_fine + t * dex_stride (dex_stride >= n lets a chunk write its columns
 * of a (T, n_total) tensor). */
typedef struct {
  const dexnerf_mlp_program* prog;   /* host */
  const float* params;               /* device: fp32 program-layout buffer (fp32 path; also the source of `packed`) */
  const dexnerf_flexible_spec* spec; /* host; NULL = fp32 CUDA-core path */
  const void* packed;                /* device: dexnerf_tc_pack blob (tensor-core path) */
  const void* packed_t;              /* device: dexnerf_tc_pack_bwd blob (dexnerf_render_fused_bwd only) */
} dexnerf_model_ref;

typedef struct {
  int64_t n;
  const float* ro; const float* rd;          /* (n,3) each, or NULL with a camera */
  const float* T_w2c; const float* K;        /* device 4x4 / 3x3, camera form */
  int32_t H, W, row0, rows;
  int32_t use_viewdirs, ndc;
  float focal, near, far;
  int32_t Nc, Nf, lindisp, perturb, white_background;
  float noise_std;
  const float* thresholds; int32_t T; int32_t pad0_;
  const float* t_rand; const float* u; const float* noise_coarse; const float* noise_fine;
  uint64_t seed, offset;
  dexnerf_model_ref coarse, fine;
  void* tape_coarse; void* tape_fine;
  void* workspace; int64_t workspace_bytes;
  float* rgb_coarse; float* depth_coarse; float* acc_coarse;
  float* rgb_fine; float* depth_fine; float* acc_fine;
  float* dex_fine; int64_t dex_stride;
  /* optional per-launch timing (benchmarks): host array of 2 * DEXNERF_RENDER_LAUNCHES events made by
   * dexnerf_event_create; events[2k] / events[2k+1] are recorded before / after launch k (order: setup, ndc,
   * coarse query, coarse compositing, resample, fine query, fine compositing).  NULL = no events. */
  void** events;
} dexnerf_render_params;
#define DEXNERF_RENDER_LAUNCHES 7
DEXNERF_API void* dexnerf_event_create(void);                 /* cudaEvent_t with timing, NULL on failure */
DEXNERF_API void dexnerf_event_destroy(void* ev);
DEXNERF_API float dexnerf_event_elapsed_ms(void* start, void* end); /* < 0 when not complete / never recorded */

/* byte offsets into the workspace: [0] total, [1] ro, [2] rd, [3] viewdirs, [4] z_coarse, [5] rf_coarse,
 * [6] weights_coarse, [7] z_fine, [8] rf_fine, [9] t_rand, [10] u, [11] noise_coarse, [12] noise_fine,
 * [13] ro_raw, [14] rd_raw (the camera's rays before ndc_rays) */
#define DEXNERF_RENDER_WS_SLOTS 16
DEXNERF_API int64_t dexnerf_render_workspace_bytes(int64_t n, int Nc, int Nf);
DEXNERF_API int dexnerf_render_workspace_layout(int64_t n, int Nc, int Nf, int64_t* out /*host, 16*/);
/* the setup launch alone (what the training paths of nerf/training.py share with the fused call): fills the
 * workspace slots ro / rd / viewdirs / z_coarse / t_rand / u / noise_* according to `p` */
DEXNERF_API int dexnerf_ray_setup(const dexnerf_render_params* p /*host*/, void* stream);
DEXNERF_API int dexnerf_render_fused_fwd(const dexnerf_render_params* p /*host*/, void* stream);
/* backward of a dexnerf_render_fused_fwd call that recorded tapes (train_dexnerf_rgb.py:278 loss.backward()):
 * g_rgb_coarse, g_rgb_fine (n,3) = dL/d(rgb maps) -> grads_coarse / grads_fine, fp32 buffers in the program layout,
 * ACCUMULATED into.  d_rf_scratch: (n, Nc+Nf, 4) floats.  which: bit 0 = the fine network's chain, bit 1 = the
 * coarse network's (3 = both, fine first: its gradient buffer is complete first, so a data-parallel caller can
 * issue the two halves separately and start reducing the fine gradients while the coarse chain runs).
 * 3 launches per network: compositing backward, activation-gradient chain, weight-gradient GEMM
 * (DEXNERF_BWD=fused / shared select the single-launch MLP backwards of dexnerf_tc_backward instead;
 * DEXNERF_BWD_SPLIT=k with which = 3 runs the coarse network's compositing backward + chain on k SMs and a second
 * stream next to the fine network's GEMM on the other 148 - k: same results, measured slower on a B200, off). */
DEXNERF_API int dexnerf_render_fused_bwd(const dexnerf_render_params* p /*host*/, const float* g_rgb_coarse,
                                         const float* g_rgb_fine, float* d_rf_scratch, float* grads_coarse,
                                         float* grads_fine, int which, void* stream);

/* ---- validation depth metrics (nerf/train_utils.py:9-30 compute_err_metric, looped over the
 * threshold candidates at train_dexnerf_rgb.py:391-404).  pred (T,n) threshold depth planes, gt (n),
 * mask (n) uint8 or NULL (= the reference's (gt > 0) & (gt < 1.25)).  out (T,4) =
 * [mean abs err in mm, fraction > 2 mm, > 4 mm, > 8 mm]; best (int32, may be NULL) = index of the
 * first threshold with the smallest abs err (< 1000, else -1).  workspace: (4T + 1) doubles. */
DEXNERF_API int dexnerf_depth_error_metrics(const float* pred, const float* gt, const uint8_t* mask, int64_t n,
                                int T, float* out, int32_t* best, void* workspace, void* stream);

/* ---- depth_error_img (nerf/train_utils.py:31-70): the colour-coded depth error image of the validation
 * block (train_dexnerf_rgb.py:415).  est, gt: (H, W) floats of ONE image (the reference returns image 0 of
 * its batch), mask (H, W) uint8, out (H, W, 3). */
DEXNERF_API int dexnerf_depth_error_image(const float* est, const float* gt, const uint8_t* mask, int H, int W,
                              float abs_thres, float* out, void* stream);

/* ---- training-loop glue on the flat parameter buffer (train_dexnerf_rgb.py:264-289).
 * dexnerf_mse_loss_grad: *loss_accum += sum((pred - target)^2) / total_count over `count` floats
 * (img2mse, nerf_helpers.py:9-10; total_count = count unless the batch is processed in chunks) and
 * grad = 2 (pred - target) / total_count.
 * dexnerf_adam_step: torch.optim.Adam's update (defaults: no weight decay, no amsgrad) on n floats;
 * `step` counts from 1; grads are multiplied by grad_scale first (1 / world for data parallel). */
DEXNERF_API int dexnerf_mse_loss_grad(const float* pred, const float* target, int64_t count, int64_t total_count,
                          float* grad, float* loss_accum, void* stream);
/* both loss terms of train_dexnerf_rgb.py:264-277 in one launch: loss3 = [total, coarse, fine] ACCUMULATES
 * sum((pred - target)^2) / total_count of each prediction (total = coarse + fine); grads as above. */
DEXNERF_API int dexnerf_mse_loss_pair(const float* pred_coarse, const float* pred_fine, const float* target,
                          int64_t count, int64_t total_count, float* grad_coarse, float* grad_fine, float* loss3,
                          void* stream);
/* nn.Linear tensors -> the flat program-layout buffer in one launch.  ptr_table: DEVICE array of 2*n_ops
 * pointers (weight (out,in) row-major, bias) in op order; flat: the buffer dexnerf_mlp_query / tc_pack read. */
DEXNERF_API int dexnerf_pack_params(const dexnerf_mlp_program* prog /*host*/, const void* ptr_table, float* flat,
                        void* stream);
DEXNERF_API int dexnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                      void* stream);

/* dexnerf_adam_step that also clears `grads` as it consumes them (the weight-gradient GEMMs accumulate, so the
 * next iteration needs no memset launch) */
DEXNERF_API int dexnerf_adam_step_zero_grad(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                                void* stream);

/* ---- Data-parallel training over NVLink peer memory (one process per GPU).
 * Replaces, for the flat-buffer trainer, the pair "ncclAllReduce of the gradients, then optimizer.step()"
 * (train_dexnerf_rgb.py:278-281 on every rank of a data-parallel run) by ONE kernel that reads every rank's gradient
 * buffer directly (P2P loads), sums in rank order, applies torch.optim.Adam's update and clears the next step's buffer.
 * dexnerf_p2p_alloc: a zeroed cudaMalloc allocation that CUDA IPC can export; _export writes the 64-byte
 * cudaIpcMemHandle_t, _open maps a peer's handle into this process (peer access enabled lazily), _close unmaps it. */
DEXNERF_API int dexnerf_p2p_alloc(int64_t bytes, void** ptr);
DEXNERF_API int dexnerf_p2p_free(void* ptr);
DEXNERF_API int dexnerf_p2p_export(void* ptr, void* handle64);
DEXNERF_API int dexnerf_p2p_open(const void* handle64, void** ptr);
DEXNERF_API int dexnerf_p2p_close(void* ptr);
/* peer_grads[r] / peer_flags[r] (HOST arrays of `world` device pointers valid on this GPU; entry `rank` is the local
 * buffer): every rank's gradient buffer of THIS step (n floats) and every rank's flag array (`world` uint32, from
 * dexnerf_p2p_alloc).  The gradient buffers ping-pong between steps: `zero_next` is the LOCAL buffer the next step
 * accumulates into; it is cleared here.  `token` must grow by one per call, the same on every rank, starting at 1.
 * n must be a multiple of 4; `step` is Adam's step count (from 1); grad_scale = 1 / world for the mean. */
DEXNERF_API int dexnerf_adam_step_allreduce(float* params, float* exp_avg, float* exp_avg_sq, float* zero_next, int64_t n,
                                const void* const* peer_grads, void* const* peer_flags, int rank, int world,
                                uint32_t token, float lr, float beta1, float beta2, float eps, int64_t step,
                                float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEXNERF_H_ */
