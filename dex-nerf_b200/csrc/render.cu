// The render driver as one C-ABI call per ray chunk (reference: nerf/train_utils.py:92-202
// predict_and_render_radiance, :205-288 run_one_iter_of_nerf).
//
// dexnerf_render_fused_fwd sequences the kernels of this library on the caller's stream:
//
//   ray_setup_kernel   ONE launch for everything the reference does with ~25 eager kernels before the first
//                      network call: (camera form) get_ray_bundle, view directions rd / |rd| (:222-226), the
//                      stratified depths (:111-133) and, in train mode, the random draws (jitter t_rand, the
//                      inverse-CDF u, the sigma noises) from a counter-based Philox4x32-10 generator - or the
//                      replayed draws of a parity test
//   [ndc_kernel]       only when the YAML says no_ndc: False (:238-242)
//   coarse MLP query -> compositing -> resample + merge -> fine MLP query -> compositing + Dex depth
//
// i.e. 6 launches per chunk, no per-ray near / far / ones arrays, no (n, 11) ray matrix, no torch glue.
// Intermediates live in a caller-provided workspace (dexnerf_render_workspace_layout).
#include <stdlib.h>

#include "common.cuh"

namespace dexnerf {

// ------------------------------------------------------------------ Philox4x32-10 (Salmon et al., SC'11)
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint64_t index, uint32_t stream, uint64_t offset) const {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32) ^ (uint32_t)(offset >> 32);
    uint32_t c2 = stream, c3 = (uint32_t)offset;
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// [0, 1) with 24 random bits, like torch.rand on fp32
__device__ __forceinline__ float uniform01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }
// standard normal (Box-Muller on two 24-bit uniforms; the first one in (0, 1])
__device__ __forceinline__ float normal01(uint32_t x, uint32_t y) {
  const float u1 = (float)((x >> 8) + 1u) * 5.9604644775390625e-8f;
  const float u2 = (float)(y >> 8) * 5.9604644775390625e-8f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

struct SetupArgs {
  int64_t n;
  const float* ro_in; const float* rd_in;       // explicit rays (NULL in camera form)
  const float* T; const float* K; int W, row0;  // camera form
  float* ro_out; float* rd_out;                 // written in camera form only
  float* vd;                                    // view directions or NULL
  float* z; int Nc; float near, far; int lindisp;
  const float* t_rand_in; int perturb;          // replayed jitter, or Philox when perturb && !t_rand_in
  float* u_out; int Nf;                         // Philox u (NULL: not wanted / replayed by the caller)
  float* noise_c_out; float* noise_f_out; float noise_std;
  uint64_t seed, offset;
};

constexpr int kSetupTable = 1024;     // coarse samples per ray the depth tables hold

__global__ void __launch_bounds__(256) ray_setup_kernel(const SetupArgs a) {
  __shared__ float s_rinv[9];
  __shared__ float s_org[3];
  __shared__ float s_lower[kSetupTable], s_diff[kSetupTable];
  const bool camera = a.ro_in == nullptr;
  if (camera) {
    if (threadIdx.x == 0) camera_inverse(a.T, s_rinv, s_org);
    __syncthreads();
  }
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  // ---- per ray: (camera) origin + direction, view direction
  if (camera || a.vd) {
    float fx = 0.f, cx = 0.f, cy = 0.f;
    if (camera) { fx = a.K[0]; cx = a.K[2]; cy = a.K[5]; }
    for (int64_t r = tid; r < a.n; r += nth) {
      float d[3];
      if (camera) {
        pixel_direction((int)(r / a.W) + a.row0, (int)(r % a.W), fx, cx, cy, s_rinv, d);
#pragma unroll
        for (int k = 0; k < 3; ++k) { a.rd_out[r * 3 + k] = d[k]; a.ro_out[r * 3 + k] = s_org[k]; }
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = a.rd_in[r * 3 + k];
      }
      if (a.vd) {
        // rd / rd.norm(p=2, dim=-1): the norm accumulated in fp64 and rounded once, as ATen's CPU kernel does
        const float nrm = (float)sqrt((double)d[0] * (double)d[0] + (double)d[1] * (double)d[1] +
                                      (double)d[2] * (double)d[2]);
#pragma unroll
        for (int k = 0; k < 3; ++k) a.vd[r * 3 + k] = __fdiv_rn(d[k], nrm);
      }
    }
  }
  const Philox rng(a.seed);
  // ---- per coarse sample: depth (+ jitter), coarse sigma noise.  The linspace depth, and with jitter the lower bound
  // and the width of the stratum, depend on the sample index only: one table per CTA (the same operations, hence the
  // same bits, as stratified_depth), then four consecutive samples of a ray per thread and 16-byte stores with
  // 32-bit index arithmetic - the first version spent its time in a 64-bit modulo per sample (0.22 ms for the 41 M
  // depths of a C2 frame, 0.13 of the HBM peak).
  const int64_t total_c = a.n * a.Nc;
  const bool draw = (a.perturb && !a.t_rand_in) || a.noise_c_out;
  if (a.Nc <= kSetupTable && a.Nc % 4 == 0 && total_c < (int64_t)1 << 32) {
    for (int i = threadIdx.x; i < a.Nc; i += blockDim.x) {
      const float v = coarse_depth(a.near, a.far, a.Nc, a.lindisp, i);
      float lower = v, diff = 0.0f;
      if (a.perturb) {
        const float prev = i > 0 ? coarse_depth(a.near, a.far, a.Nc, a.lindisp, i - 1) : v;
        const float next = i < a.Nc - 1 ? coarse_depth(a.near, a.far, a.Nc, a.lindisp, i + 1) : v;
        lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(v, prev)) : v;
        const float upper = i < a.Nc - 1 ? __fmul_rn(0.5f, __fadd_rn(next, v)) : v;
        diff = __fsub_rn(upper, lower);
      }
      s_lower[i] = lower; s_diff[i] = diff;
    }
    __syncthreads();
    const uint32_t quads = (uint32_t)(total_c >> 2), step = (uint32_t)nth, nc = (uint32_t)a.Nc;
    for (uint32_t q = (uint32_t)tid; q < quads; q += step) {
      const uint32_t e0 = q << 2, i0 = e0 % nc;
      float zz[4], nz[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float t = 0.0f;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (draw) x = rng((uint64_t)(e0 + k), 0u, a.offset);
        if (a.perturb) t = a.t_rand_in ? a.t_rand_in[e0 + k] : uniform01(x.x);
        zz[k] = a.perturb ? __fadd_rn(s_lower[i0 + k], __fmul_rn(s_diff[i0 + k], t)) : s_lower[i0 + k];
        nz[k] = a.noise_c_out ? normal01(x.y, x.z) * a.noise_std : 0.0f;
      }
      reinterpret_cast<float4*>(a.z)[q] = make_float4(zz[0], zz[1], zz[2], zz[3]);
      if (a.noise_c_out) reinterpret_cast<float4*>(a.noise_c_out)[q] = make_float4(nz[0], nz[1], nz[2], nz[3]);
    }
  } else {
    for (int64_t e = tid; e < total_c; e += nth) {
      const int i = (int)(e % a.Nc);
      float t = 0.0f;
      uint4 x = make_uint4(0, 0, 0, 0);
      if (draw) x = rng((uint64_t)e, 0u, a.offset);
      if (a.perturb) t = a.t_rand_in ? a.t_rand_in[e] : uniform01(x.x);
      a.z[e] = stratified_depth(a.near, a.far, a.Nc, a.lindisp, i, a.perturb != 0, t);
      if (a.noise_c_out) a.noise_c_out[e] = normal01(x.y, x.z) * a.noise_std;
    }
  }
  // ---- per fine draw: u
  if (a.u_out) {
    const int64_t total_u = a.n * a.Nf;
    for (int64_t e = tid; e < total_u; e += nth) a.u_out[e] = uniform01(rng((uint64_t)e, 1u, a.offset).x);
  }
  // ---- per fine sample: sigma noise
  if (a.noise_f_out) {
    const int64_t total_f = a.n * (a.Nc + a.Nf);
    for (int64_t e = tid; e < total_f; e += nth) {
      const uint4 x = rng((uint64_t)e, 2u, a.offset);
      a.noise_f_out[e] = normal01(x.x, x.y) * a.noise_std;
    }
  }
}

// ------------------------------------------------------------------ workspace
struct WsLayout {
  int64_t total, ro, rd, vd, z_c, rf_c, w_c, z_f, rf_f, t_rand, u, noise_c, noise_f, ro_raw, rd_raw;
};
static WsLayout ws_layout(int64_t n, int Nc, int Nf) {
  WsLayout L{};
  int64_t off = 0;
  auto take = [&](int64_t bytes) { const int64_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const int64_t Sf = (int64_t)Nc + Nf;
  L.ro = take(n * 12); L.rd = take(n * 12); L.vd = take(n * 12);
  L.z_c = take(n * Nc * 4); L.rf_c = take(n * Nc * 16); L.w_c = take(n * Nc * 4);
  L.z_f = take(n * Sf * 4); L.rf_f = take(n * Sf * 16);
  L.t_rand = take(0); L.u = take(n * Nf * 4);
  L.noise_c = take(n * Nc * 4); L.noise_f = take(n * Sf * 4);
  L.ro_raw = take(n * 12); L.rd_raw = take(n * 12);
  L.total = off;
  return L;
}

struct Resolved {     // pointers of one call, after the setup decisions
  const float *ro, *rd, *vd, *z_c, *noise_c, *noise_f, *u;
  float *rf_c, *w_c, *z_f, *rf_f;
};

static int check_params(const dexnerf_render_params* p) {
  DN_REQUIRE(p, "render: null params");
  DN_REQUIRE(p->n >= 0 && p->Nc >= 3 && p->Nf >= 1, "render: need n >= 0, Nc >= 3, Nf >= 1");
  DN_REQUIRE(p->workspace && (reinterpret_cast<uintptr_t>(p->workspace) & 255) == 0,
             "render: the workspace must be 256-byte aligned");
  DN_REQUIRE(p->workspace_bytes >= ws_layout(p->n, p->Nc, p->Nf).total, "render: workspace too small (%lld < %lld)",
             (long long)p->workspace_bytes, (long long)ws_layout(p->n, p->Nc, p->Nf).total);
  if (!p->ro) {
    DN_REQUIRE(p->T_w2c && p->K && !p->rd, "render: give either ro and rd or a camera (T_w2c, K)");
    DN_REQUIRE(p->H > 0 && p->W > 0 && p->row0 >= 0 && p->rows >= 0 && p->row0 + p->rows <= p->H &&
               (int64_t)p->rows * p->W == p->n, "render: camera rows do not match n");
  } else {
    DN_REQUIRE(p->rd, "render: rd is null");
  }
  DN_REQUIRE(p->T >= 0 && (p->T == 0 || p->thresholds), "render: thresholds is null");
  return 0;
}

// optional event bracket around launch k of the fused call
struct Bracket {
  void** ev; int k; cudaStream_t st;
  Bracket(void** ev_, int k_, cudaStream_t st_) : ev(ev_), k(k_), st(st_) { if (ev) cudaEventRecord((cudaEvent_t)ev[2 * k], st); }
  ~Bracket() { if (ev) cudaEventRecord((cudaEvent_t)ev[2 * k + 1], st); }
};

// the setup launch (+ ndc); fills `r`
static int run_setup(const dexnerf_render_params* p, Resolved* r, cudaStream_t st) {
  const WsLayout L = ws_layout(p->n, p->Nc, p->Nf);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  auto f = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const bool camera = p->ro == nullptr;
  SetupArgs a{};
  a.n = p->n;
  a.ro_in = p->ro; a.rd_in = p->rd; a.T = p->T_w2c; a.K = p->K; a.W = p->W; a.row0 = p->row0;
  // with ndc the camera's rays are the RAW rays: the warped ones go to the ro / rd slots
  a.ro_out = camera ? f(p->ndc ? L.ro_raw : L.ro) : nullptr;
  a.rd_out = camera ? f(p->ndc ? L.rd_raw : L.rd) : nullptr;
  a.vd = p->use_viewdirs ? f(L.vd) : nullptr;
  a.z = f(L.z_c); a.Nc = p->Nc; a.near = p->near; a.far = p->far; a.lindisp = p->lindisp;
  a.t_rand_in = p->t_rand; a.perturb = p->perturb != 0;
  a.Nf = p->Nf;
  a.u_out = (p->perturb && !p->u) ? f(L.u) : nullptr;
  const bool noisy = p->noise_std > 0.0f;
  a.noise_c_out = (noisy && !p->noise_coarse) ? f(L.noise_c) : nullptr;
  a.noise_f_out = (noisy && !p->noise_fine) ? f(L.noise_f) : nullptr;
  a.noise_std = p->noise_std; a.seed = p->seed; a.offset = p->offset;
  const int64_t work = p->n * ((int64_t)p->Nc + (a.noise_f_out ? p->Nc + p->Nf : 0));
  int64_t blocks = ceil_div64(work > p->n ? work : p->n, 256);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  {
    Bracket b(p->events, 0, st);
    ray_setup_kernel<<<(int)blocks, 256, 0, st>>>(a);
  }
  DN_CHECK_LAUNCH("ray_setup");
  const float* ro = camera ? a.ro_out : p->ro;
  const float* rd = camera ? a.rd_out : p->rd;
  if (p->ndc) {
    Bracket b(p->events, 1, st);
    if (int rc = dexnerf_ndc_rays(ro, rd, p->n, p->H, p->W, p->focal, 1.0f, f(L.ro), f(L.rd), st)) return rc;
    ro = f(L.ro); rd = f(L.rd);
  }
  r->ro = ro; r->rd = rd; r->vd = a.vd; r->z_c = a.z;
  r->noise_c = p->noise_coarse ? p->noise_coarse : a.noise_c_out;
  r->noise_f = p->noise_fine ? p->noise_fine : a.noise_f_out;
  r->u = p->u ? p->u : a.u_out;
  r->rf_c = f(L.rf_c); r->w_c = f(L.w_c); r->z_f = f(L.z_f); r->rf_f = f(L.rf_f);
  return 0;
}

// pointers as run_setup left them, without launching anything (the backward of a recorded forward)
static void resolve_only(const dexnerf_render_params* p, Resolved* r) {
  const WsLayout L = ws_layout(p->n, p->Nc, p->Nf);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p->workspace);
  auto f = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const bool camera = p->ro == nullptr;
  const bool noisy = p->noise_std > 0.0f;
  r->ro = (camera || p->ndc) ? f(L.ro) : p->ro;
  r->rd = (camera || p->ndc) ? f(L.rd) : p->rd;
  r->vd = p->use_viewdirs ? f(L.vd) : nullptr;
  r->z_c = f(L.z_c);
  r->noise_c = p->noise_coarse ? p->noise_coarse : (noisy ? f(L.noise_c) : nullptr);
  r->noise_f = p->noise_fine ? p->noise_fine : (noisy ? f(L.noise_f) : nullptr);
  r->u = p->u ? p->u : (p->perturb ? f(L.u) : nullptr);
  r->rf_c = f(L.rf_c); r->w_c = f(L.w_c); r->z_f = f(L.z_f); r->rf_f = f(L.rf_f);
}

// Which MLP backward the training path launches (DESIGN.md section 3.2 has the measurements):
//   default / DEXNERF_BWD=split   activation-gradient chain, then weight-gradient GEMM (two launches, the gradient
//                                 images make one HBM round trip) - the fastest of the three on a B200 today;
//   DEXNERF_BWD=fused             one launch, chain and GEMM on disjoint SMs, images handed over through L2;
//   DEXNERF_BWD=shared            one launch, chain and GEMM in every CTA (one per SM).
static int backward_mode() {
  static const int v = [] {
    const char* e = getenv("DEXNERF_BWD");
    if (!e) return 3;
    if (e[0] == 'f') return 4;
    if (e[0] == 's' && e[1] == 'h') return 8;
    return 3;
  }();
  return v;
}
// tuning knob of the fused launches: DEXNERF_BWD_VARIANT=<n> is passed through as `variant`
static int variant_override() {
  static const int v = [] { const char* e = getenv("DEXNERF_BWD_VARIANT"); return e ? atoi(e) : 0; }();
  return v;
}

// The two networks' backwards are independent (the resampled depths carry no gradient, train_utils.py:166), and their
// kernels look complementary: the weight-gradient GEMM reads (tensor pipe 24 % busy, nothing written), the
// activation-gradient chain computes and writes.  With DEXNERF_BWD_SPLIT=k the driver runs the FINE network's GEMM on
// 148 - k SMs and, on a second stream, the COARSE network's compositing backward + chain on the other k at the same
// time; the coarse GEMM follows on all SMs.  Measured on a B200 (C4, ms per iteration): off 4.62, k = 32 4.69, 40 4.87,
// 48 4.94, 56 4.94 - the GEMM slows down by as much as the chain takes (1.47 -> 1.71 ms at k = 32), because both are
// bound by the same HBM (10.2 GB in that window = 6 TB/s of mixed traffic).  Hence OFF by default (k = 0).
static int overlap_split() {
  static const int v = [] {
    const char* e = getenv("DEXNERF_BWD_SPLIT");
    const int k = e ? atoi(e) : 0;
    return (k < 0 || k > kNumSMs - 16) ? 0 : k;
  }();
  return v;
}
struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
static SideStream* side_stream() {
  static SideStream per_device[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_device[dev];
  if (!s.st) {
    if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) { s.st = nullptr; return nullptr; }
    cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming);
  }
  return &s;
}

static int query(const dexnerf_model_ref& m, const Resolved& r, const float* z, int64_t n, int S, float* rf,
                 void* tape, cudaStream_t st) {
  DN_REQUIRE(m.prog, "render: model program is null");
  const float* vd = m.prog->dim_dir ? r.vd : nullptr;
  DN_REQUIRE(m.prog->dim_dir == 0 || vd, "render: the model takes view directions but use_viewdirs is 0");
  if (m.spec) {
    DN_REQUIRE(m.packed, "render: tensor-core path without packed weights");
    if (tape) return dexnerf_tc_query_train(m.spec, m.packed, r.ro, r.rd, vd, z, n, S, rf, tape, st);
    return dexnerf_tc_query(m.spec, m.packed, r.ro, r.rd, vd, z, n, S, rf, nullptr, -1, 0, st);
  }
  DN_REQUIRE(!tape, "render: the training tape exists on the tensor-core path only");
  DN_REQUIRE(m.params, "render: fp32 path without parameters");
  return dexnerf_mlp_query(m.prog, m.params, r.ro, r.rd, vd, z, n, S, rf, st);
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int64_t dexnerf_render_workspace_bytes(int64_t n, int Nc, int Nf) {
  if (n < 0 || Nc < 1 || Nf < 0) { set_error("render_workspace_bytes: bad sizes"); return -1; }
  return ws_layout(n, Nc, Nf).total;
}

extern "C" DEXNERF_API int dexnerf_render_workspace_layout(int64_t n, int Nc, int Nf, int64_t* out) {
  DN_REQUIRE(out && n >= 0 && Nc >= 1 && Nf >= 0, "render_workspace_layout: bad arguments");
  const WsLayout L = ws_layout(n, Nc, Nf);
  const int64_t v[DEXNERF_RENDER_WS_SLOTS] = {L.total, L.ro, L.rd, L.vd, L.z_c, L.rf_c, L.w_c, L.z_f, L.rf_f,
                                              L.t_rand, L.u, L.noise_c, L.noise_f, L.ro_raw, L.rd_raw, 0};
  for (int i = 0; i < DEXNERF_RENDER_WS_SLOTS; ++i) out[i] = v[i];
  return 0;
}

extern "C" DEXNERF_API void* dexnerf_event_create(void) {
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) { set_error("event_create: cudaEventCreate failed"); return nullptr; }
  return e;
}
extern "C" DEXNERF_API void dexnerf_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
extern "C" DEXNERF_API float dexnerf_event_elapsed_ms(void* start, void* end) {
  float ms = -1.0f;
  if (!start || !end || cudaEventElapsedTime(&ms, (cudaEvent_t)start, (cudaEvent_t)end) != cudaSuccess) {
    cudaGetLastError();
    return -1.0f;
  }
  return ms;
}

extern "C" DEXNERF_API int dexnerf_ray_setup(const dexnerf_render_params* p, void* stream) {
  if (int rc = check_params(p)) return rc;
  if (p->n == 0) return 0;
  Resolved r;
  return run_setup(p, &r, (cudaStream_t)stream);
}

extern "C" DEXNERF_API int dexnerf_render_fused_fwd(const dexnerf_render_params* p, void* stream) {
  if (int rc = check_params(p)) return rc;
  if (p->n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = p->n;
  const int Nc = p->Nc, Sf = p->Nc + p->Nf;
  Resolved r;
  if (int rc = run_setup(p, &r, st)) return rc;
  {
    Bracket b(p->events, 2, st);
    if (int rc = query(p->coarse, r, r.z_c, n, Nc, r.rf_c, p->tape_coarse, st)) return rc;
  }
  {
    // the coarse pass keeps its weights (the resampler's pdf) and drops its Dex depths (train_utils.py:157)
    Bracket b(p->events, 3, st);
    if (int rc = volume_render_impl(r.rf_c, r.z_c, r.rd, r.noise_c, n, Nc, p->white_background, nullptr, 0,
                                    p->rgb_coarse, nullptr, p->acc_coarse, r.w_c, p->depth_coarse, nullptr, nullptr,
                                    n, st)) return rc;
  }
  {
    Bracket b(p->events, 4, st);
    if (int rc = dexnerf_resample_merge(r.z_c, r.w_c, n, Nc, p->Nf, r.u, r.z_f, st)) return rc;
  }
  {
    Bracket b(p->events, 5, st);
    if (int rc = query(p->fine, r, r.z_f, n, Sf, r.rf_f, p->tape_fine, st)) return rc;
  }
  Bracket b(p->events, 6, st);
  return volume_render_impl(r.rf_f, r.z_f, r.rd, r.noise_f, n, Sf, p->white_background, p->thresholds,
                            p->dex_fine ? p->T : 0, p->rgb_fine, nullptr, p->acc_fine, nullptr, p->depth_fine,
                            p->dex_fine, nullptr, p->dex_stride >= n ? p->dex_stride : n, st);
}

extern "C" DEXNERF_API int dexnerf_render_fused_bwd(const dexnerf_render_params* p, const float* g_rgb_coarse,
                                                    const float* g_rgb_fine, float* d_rf_scratch,
                                                    float* grads_coarse, float* grads_fine, int which,
                                                    void* stream) {
  if (int rc = check_params(p)) return rc;
  if (p->n == 0) return 0;
  DN_REQUIRE(p->tape_coarse && p->tape_fine && p->coarse.spec && p->fine.spec && p->coarse.packed_t && p->fine.packed_t,
             "render_fused_bwd: needs the tapes of a tensor-core forward and the transposed weight images");
  DN_REQUIRE(d_rf_scratch && (which & 3), "render_fused_bwd: null scratch / nothing to do");
  const int64_t n = p->n;
  const int Nc = p->Nc, Sf = p->Nc + p->Nf;
  Resolved r;
  resolve_only(p, &r);
  cudaStream_t st = (cudaStream_t)stream;
  // events (optional): 0 compositing bwd fine, 1 MLP backward fine (dX when split), 2 (dW when split), 3 compositing bwd
  // coarse, 4 MLP backward coarse (dX when split), 5 (dW when split)
  auto chain = [&](const dexnerf_model_ref& m, void* tape, const float* rf, const float* z, const float* noise, int S,
                   const float* g_rgb, float* grads, int ev0) -> int {
    DN_REQUIRE(g_rgb && grads, "render_fused_bwd: null gradient pointer");
    {
      Bracket b(p->events, ev0, st);
      if (int rc = dexnerf_volume_render_backward(rf, z, r.rd, noise, n, S, p->white_background, g_rgb, nullptr,
                                                  nullptr, d_rf_scratch, stream)) return rc;
    }
    if (backward_mode() == 3) {
      for (int bit = 1; bit <= 2; ++bit) {
        Bracket b(p->events, ev0 + bit, st);
        if (int rc = dexnerf_tc_backward(m.spec, m.prog, m.packed, m.packed_t, tape, d_rf_scratch, n, S, grads, bit, 0,
                                         stream)) return rc;
      }
      return 0;
    }
    Bracket b(p->events, ev0 + 1, st);
    return dexnerf_tc_backward(m.spec, m.prog, m.packed, m.packed_t, tape, d_rf_scratch, n, S, grads, backward_mode(),
                               variant_override(), stream);
  };
  const int k = overlap_split();
  SideStream* side = (which == 3 && backward_mode() == 3 && k > 0 && n * (int64_t)Nc >= (int64_t)k * 256) ? side_stream() : nullptr;
  if (side) {
    // main stream: compositing backward (fine), chain (fine), GEMM (fine, 148 - k SMs), [join], GEMM (coarse)
    // side stream: [fork after the fine chain - it has read the shared d_rf scratch], compositing backward (coarse),
    //              chain (coarse, k SMs)
    DN_REQUIRE(g_rgb_fine && g_rgb_coarse && grads_fine && grads_coarse, "render_fused_bwd: null gradient pointer");
    {
      Bracket b(p->events, 0, st);
      if (int rc = dexnerf_volume_render_backward(r.rf_f, r.z_f, r.rd, r.noise_f, n, Sf, p->white_background, g_rgb_fine,
                                                  nullptr, nullptr, d_rf_scratch, stream)) return rc;
    }
    {
      Bracket b(p->events, 1, st);
      if (int rc = dexnerf_tc_backward(p->fine.spec, p->fine.prog, p->fine.packed, p->fine.packed_t, p->tape_fine,
                                       d_rf_scratch, n, Sf, grads_fine, 1, 0, stream)) return rc;
    }
    DN_CUDA(cudaEventRecord(side->fork, st));
    DN_CUDA(cudaStreamWaitEvent(side->st, side->fork, 0));
    {
      Bracket b(p->events, 2, st);
      if (int rc = dexnerf_tc_backward(p->fine.spec, p->fine.prog, p->fine.packed, p->fine.packed_t, p->tape_fine,
                                       d_rf_scratch, n, Sf, grads_fine, 2, (kNumSMs - k) << 8, stream)) return rc;
    }
    {
      Bracket b(p->events, 3, side->st);
      if (int rc = dexnerf_volume_render_backward(r.rf_c, r.z_c, r.rd, r.noise_c, n, Nc, p->white_background,
                                                  g_rgb_coarse, nullptr, nullptr, d_rf_scratch, side->st)) return rc;
    }
    {
      Bracket b(p->events, 4, side->st);
      if (int rc = dexnerf_tc_backward(p->coarse.spec, p->coarse.prog, p->coarse.packed, p->coarse.packed_t,
                                       p->tape_coarse, d_rf_scratch, n, Nc, grads_coarse, 1, k << 16, side->st)) return rc;
    }
    DN_CUDA(cudaEventRecord(side->join, side->st));
    DN_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    Bracket b(p->events, 5, st);
    return dexnerf_tc_backward(p->coarse.spec, p->coarse.prog, p->coarse.packed, p->coarse.packed_t, p->tape_coarse,
                               d_rf_scratch, n, Nc, grads_coarse, 2, 0, stream);
  }
  if (which & 1)
    if (int rc = chain(p->fine, p->tape_fine, r.rf_f, r.z_f, r.noise_f, Sf, g_rgb_fine, grads_fine, 0)) return rc;
  if (which & 2)
    if (int rc = chain(p->coarse, p->tape_coarse, r.rf_c, r.z_c, r.noise_c, Nc, g_rgb_coarse, grads_coarse, 3)) return rc;
  return 0;
}
