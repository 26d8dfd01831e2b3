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

#define DEXNERF_ABI_VERSION 1
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
#define DEXNERF_TC_PACK_WORKSPACE_BYTES 16384
/* params: the fp32 program-layout buffer of the same model (as for dexnerf_mlp_query);
 * prog: the program it was built from.  packed: device blob of dexnerf_tc_packed_bytes();
 * workspace: unused since ABI revision 1.1 (may be NULL); packing is stream-ordered like every other
 * entry point (the layout tables travel as kernel parameters). */
DEXNERF_API int dexnerf_tc_pack(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                    const float* params, void* packed, void* workspace, void* stream);
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
 * packed_t = dexnerf_tc_pack_bwd blob.  what: bit 0 activation-gradient chain, bit 1
 * weight-gradient GEMM (3 = a full backward); variant: 0.  (Bring-up / experiment bits of the GEMM:
 * 1 swaps the descriptor strides, 2 skips the bias column sums, 4 skips the MMAs - results invalid.) */
DEXNERF_API int dexnerf_tc_backward(const dexnerf_flexible_spec* spec, const dexnerf_mlp_program* prog,
                        const void* packed, const void* packed_t, void* tape, const float* d_rf,
                        int64_t n, int S, float* grads, int what, int variant, void* stream);

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
/* nn.Linear tensors -> the flat program-layout buffer in one launch.  ptr_table: DEVICE array of 2*n_ops
 * pointers (weight (out,in) row-major, bias) in op order; flat: the buffer dexnerf_mlp_query / tc_pack read. */
DEXNERF_API int dexnerf_pack_params(const dexnerf_mlp_program* prog /*host*/, const void* ptr_table, float* flat,
                        void* stream);
DEXNERF_API int dexnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEXNERF_H_ */
