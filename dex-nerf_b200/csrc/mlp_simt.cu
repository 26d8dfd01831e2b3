// fp32 CUDA-core evaluation of the NeRF MLPs (nerf/models.py) as a small layer program, with the
// point construction and positional encoding of run_network (nerf/train_utils.py:72-89, :136)
// fused in front of it.  This is the full-precision path (bit-level behaviour close to the
// reference's fp32 addmm); the throughput path is the tcgen05 kernel in mlp_tc.cu.
//
// One CTA owns a tile of 64 samples.  The encodings and two ping-pong activation buffers live in
// shared memory; each of the 256 threads accumulates a 4-row x (out/16)-column register tile,
// reading activations as shared-memory broadcasts and the transposed weights Wt[in][out] either from
// two cp.async-fed shared-memory stages (layers whose width is a multiple of 64: 128-bit loads) or as
// coalesced, L1/L2-resident global loads (heads, odd widths).  Nothing per-sample except the final (r,g,b,sigma)
// is written to HBM.
#include "common.cuh"

namespace dexnerf {

constexpr int kTileM = 64;
constexpr int kThreads = 256;
constexpr int kMaxOut = 256;

constexpr int kStageK = 16;                         // weight rows (input features) per staged chunk

struct SimtLayout {
  int ld_xyz, ld_dir, ld_buf;
  int off_xyz, off_dir, off_a, off_b, off_out;  // float offsets
  int off_stage, stage_floats;                  // two weight stages (16-byte aligned, stage_floats each), or -1 when they do not fit
  int total_floats;
};

static SimtLayout make_layout(const dexnerf_mlp_program& p) {
  SimtLayout L;
  L.ld_xyz = (p.dim_xyz > 0 ? p.dim_xyz : 1) | 1;
  L.ld_dir = (p.dim_dir > 0 ? p.dim_dir : 1) | 1;
  L.ld_buf = (p.max_width > 0 ? p.max_width : 1) | 1;
  L.off_xyz = 0;
  L.off_dir = L.off_xyz + kTileM * L.ld_xyz;
  L.off_a = L.off_dir + kTileM * L.ld_dir;
  L.off_b = L.off_a + kTileM * L.ld_buf;
  L.off_out = L.off_b + kTileM * L.ld_buf;
  L.total_floats = L.off_out + kTileM * 4;
  // two weight stages sized for the widest layer that takes the staged path
  L.off_stage = -1;
  L.stage_floats = 0;
  int widest = 0;
  for (int i = 0; i < p.n_ops; ++i)
    if ((p.ops[i].out_dim & 63) == 0 && p.ops[i].out_dim <= kMaxOut && p.ops[i].out_dim > widest) widest = p.ops[i].out_dim;
  const int stage0 = (L.total_floats + 3) & ~3;
  if (widest > 0 && (size_t)(stage0 + 2 * kStageK * widest) * sizeof(float) <= 227 * 1024) {
    L.off_stage = stage0;
    L.stage_floats = kStageK * widest;
    L.total_floats = stage0 + 2 * L.stage_floats;
  }
  return L;
}

__device__ __forceinline__ void src_lookup(const SimtLayout& L, float* smem, int id, float** base,
                                           int* ld) {
  switch (id) {
    case DEXNERF_ENC_XYZ: *base = smem + L.off_xyz; *ld = L.ld_xyz; break;
    case DEXNERF_ENC_DIR: *base = smem + L.off_dir; *ld = L.ld_dir; break;
    case DEXNERF_BUF_A: *base = smem + L.off_a; *ld = L.ld_buf; break;
    default: *base = smem + L.off_b; *ld = L.ld_buf; break;
  }
}

// acc[r][j] += sum_k src[(row0+r)*ld + k] * Wt[(k0+k)*N + tx + 16 j]
template <int NJ>
__device__ __forceinline__ void accumulate(float (&acc)[4][NJ], const float* __restrict__ src, int ld,
                                           int K, const float* __restrict__ Wt, int N, int tx,
                                           int row0) {
  const float* a0 = src + (row0 + 0) * ld;
  const float* a1 = src + (row0 + 1) * ld;
  const float* a2 = src + (row0 + 2) * ld;
  const float* a3 = src + (row0 + 3) * ld;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
    const float* wrow = Wt + (size_t)k * N + tx;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int nidx = tx + 16 * j;
      const float w = nidx < N ? __ldg(wrow + 16 * j) : 0.0f;
      acc[0][j] = fmaf(x0, w, acc[0][j]);
      acc[1][j] = fmaf(x1, w, acc[1][j]);
      acc[2][j] = fmaf(x2, w, acc[2][j]);
      acc[3][j] = fmaf(x3, w, acc[3][j]);
    }
  }
}

// ---- wide layers (N a multiple of 64): the transposed weights are streamed through shared memory in chunks
// of 16 input features (cp.async, two stages), so that the inner loop reads them as 128-bit shared loads
// (4 + J loads for 16 J FMAs per k instead of 4 + 4 J scalar global loads).  A thread owns 4 rows x 4 J columns,
// columns 4 tx + 64 j + (0..3).  The accumulation order per output is unchanged (k ascending), so the results
// are bit-identical to the direct path below.
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ src, int n_floats, int tid) {
  for (int e = tid * 4; e < n_floats; e += kThreads * 4) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst + e);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + e) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int J>
__device__ __forceinline__ void accumulate_staged(float (&acc)[4][4 * J], const float* __restrict__ src, int ld,
                                                  int K, const float* __restrict__ Wt, float* stage,
                                                  int stage_floats, int tx, int row0, int tid) {
  constexpr int N = 64 * J;
  const float* a0 = src + (row0 + 0) * ld;
  const float* a1 = src + (row0 + 1) * ld;
  const float* a2 = src + (row0 + 2) * ld;
  const float* a3 = src + (row0 + 3) * ld;
  const int n_chunks = (K + kStageK - 1) / kStageK;
  auto rows_of = [&](int c) { return K - c * kStageK < kStageK ? K - c * kStageK : kStageK; };
  stage_rows(stage, Wt, rows_of(0) * N, tid);
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) {
      stage_rows(stage + ((c + 1) & 1) * stage_floats, Wt + (size_t)(c + 1) * kStageK * N, rows_of(c + 1) * N, tid);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* w = stage + (c & 1) * stage_floats + 4 * tx;
    const int rows = rows_of(c), k0 = c * kStageK;
#pragma unroll 4
    for (int kk = 0; kk < rows; ++kk) {
      const float x0 = a0[k0 + kk], x1 = a1[k0 + kk], x2 = a2[k0 + kk], x3 = a3[k0 + kk];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + kk * N + 64 * j);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[0][4 * j + i] = fmaf(x0, wv[i], acc[0][4 * j + i]);
          acc[1][4 * j + i] = fmaf(x1, wv[i], acc[1][4 * j + i]);
          acc[2][4 * j + i] = fmaf(x2, wv[i], acc[2][4 * j + i]);
          acc[3][4 * j + i] = fmaf(x3, wv[i], acc[3][4 * j + i]);
        }
      }
    }
    __syncthreads();     // the stage is refilled two chunks later
  }
}

template <int J>
__device__ __forceinline__ void run_op_staged(const dexnerf_op& op, const SimtLayout& L, float* smem,
                                              const float* __restrict__ params, int tx, int ty, int tid) {
  constexpr int N = 64 * J;
  float acc[4][4 * J];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4 * J; ++c) acc[r][c] = 0.0f;
  const int row0 = ty * 4;
  const float* Wt = params + op.w_off;
  float* stage = smem + L.off_stage;
  float* src; int ld;
  src_lookup(L, smem, op.src0, &src, &ld);
  accumulate_staged<J>(acc, src, ld, op.src0_dim, Wt, stage, L.stage_floats, tx, row0, tid);
  if (op.src1 != DEXNERF_NONE && op.src1_dim > 0) {
    src_lookup(L, smem, op.src1, &src, &ld);
    accumulate_staged<J>(acc, src, ld, op.src1_dim, Wt + (size_t)op.src0_dim * N, stage, L.stage_floats, tx, row0, tid);
  }
  const float* bias = params + op.b_off;
  float* dst = smem + (op.dst == DEXNERF_BUF_A ? L.off_a : L.off_b);
#pragma unroll
  for (int j = 0; j < J; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nidx = 4 * tx + 64 * j + i;
      const float b = __ldg(bias + nidx);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float v = acc[r][4 * j + i] + b;
        if (op.relu) v = fmaxf(v, 0.0f);
        dst[(row0 + r) * L.ld_buf + nidx] = v;
      }
    }
  }
}

template <int NJ>
__device__ __forceinline__ void run_op(const dexnerf_op& op, const SimtLayout& L, float* smem,
                                       const float* __restrict__ params, int tx, int ty) {
  float acc[4][NJ];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[r][j] = 0.0f;
  const int N = op.out_dim, row0 = ty * 4;
  const float* Wt = params + op.w_off;
  float* src; int ld;
  src_lookup(L, smem, op.src0, &src, &ld);
  accumulate<NJ>(acc, src, ld, op.src0_dim, Wt, N, tx, row0);
  if (op.src1 != DEXNERF_NONE && op.src1_dim > 0) {
    src_lookup(L, smem, op.src1, &src, &ld);
    accumulate<NJ>(acc, src, ld, op.src1_dim, Wt + (size_t)op.src0_dim * N, N, tx, row0);
  }
  const float* bias = params + op.b_off;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int nidx = tx + 16 * j;
    if (nidx >= N) continue;
    const float b = __ldg(bias + nidx);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v = acc[r][j] + b;
      if (op.relu) v = fmaxf(v, 0.0f);
      const int row = row0 + r;
      if (op.dst == DEXNERF_BUF_A) smem[L.off_a + row * L.ld_buf + nidx] = v;
      else if (op.dst == DEXNERF_BUF_B) smem[L.off_b + row * L.ld_buf + nidx] = v;
      else if (op.dst == DEXNERF_OUT_SIGMA) smem[L.off_out + row * 4 + 3] = v;
      else smem[L.off_out + row * 4 + nidx] = v;  // OUT_RGB (3) or OUT_ALL (4)
    }
  }
}

struct SimtArgs {
  const float* params;
  // mode 0: encoded inputs
  const float* x;
  // mode 1: query
  const float* ro; const float* rd; const float* viewdirs; const float* z;
  int S;
  int64_t M;  // total samples
  float* out;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads)
mlp_simt_kernel(const __grid_constant__ dexnerf_mlp_program prog, const SimtLayout L, const SimtArgs a) {
  extern __shared__ float smem[];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n_tiles = ceil_div64(a.M, kTileM);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t g0 = tile * kTileM;
    // ---- stage the encoded inputs of this tile
    if (MODE == 0) {
      const int D = prog.dim_xyz + prog.dim_dir;
      for (int e = tid; e < kTileM * D; e += kThreads) {
        const int m = e / D, c = e - m * D;
        const int64_t g = g0 + m;
        const float v = g < a.M ? a.x[g * D + c] : 0.0f;
        if (c < prog.dim_xyz) smem[L.off_xyz + m * L.ld_xyz + c] = v;
        else smem[L.off_dir + m * L.ld_dir + (c - prog.dim_xyz)] = v;
      }
    } else {
      for (int e = tid; e < kTileM * prog.dim_xyz; e += kThreads) {
        const int m = e / prog.dim_xyz, c = e - m * prog.dim_xyz;
        const int64_t g = g0 + m;
        float v = 0.0f;
        if (g < a.M) {
          const int64_t ray = g / a.S;
          const float zz = a.z[g];
          const float p[3] = {__fadd_rn(a.ro[ray * 3 + 0], __fmul_rn(a.rd[ray * 3 + 0], zz)),
                              __fadd_rn(a.ro[ray * 3 + 1], __fmul_rn(a.rd[ray * 3 + 1], zz)),
                              __fadd_rn(a.ro[ray * 3 + 2], __fmul_rn(a.rd[ray * 3 + 2], zz))};
          v = pe_column(p, c, prog.Lx, prog.include_xyz, prog.log_xyz);
        }
        smem[L.off_xyz + m * L.ld_xyz + c] = v;
      }
      for (int e = tid; e < kTileM * prog.dim_dir; e += kThreads) {
        const int m = e / prog.dim_dir, c = e - m * prog.dim_dir;
        const int64_t g = g0 + m;
        float v = 0.0f;
        if (g < a.M) {
          const int64_t ray = g / a.S;
          const float d[3] = {a.viewdirs[ray * 3], a.viewdirs[ray * 3 + 1], a.viewdirs[ray * 3 + 2]};
          v = pe_column(d, c, prog.Ld, prog.include_dir, prog.log_dir);
        }
        smem[L.off_dir + m * L.ld_dir + c] = v;
      }
    }
    __syncthreads();
    // ---- run the layer program
    for (int i = 0; i < prog.n_ops; ++i) {
      const dexnerf_op& op = prog.ops[i];
      // wide hidden layers whose weight block is 16-byte aligned take the staged path (block-uniform choice)
      const bool staged = L.off_stage >= 0 && (op.out_dim & 63) == 0 && (op.w_off & 3) == 0 &&
                          (op.dst == DEXNERF_BUF_A || op.dst == DEXNERF_BUF_B) &&
                          ((reinterpret_cast<uintptr_t>(a.params) & 15) == 0);
      if (staged && op.out_dim == 256) run_op_staged<4>(op, L, smem, a.params, tx, ty, tid);
      else if (staged && op.out_dim == 192) run_op_staged<3>(op, L, smem, a.params, tx, ty, tid);
      else if (staged && op.out_dim == 128) run_op_staged<2>(op, L, smem, a.params, tx, ty, tid);
      else if (staged && op.out_dim == 64) run_op_staged<1>(op, L, smem, a.params, tx, ty, tid);
      else if (op.out_dim <= 16) run_op<1>(op, L, smem, a.params, tx, ty);
      else if (op.out_dim <= 64) run_op<4>(op, L, smem, a.params, tx, ty);
      else if (op.out_dim <= 128) run_op<8>(op, L, smem, a.params, tx, ty);
      else run_op<16>(op, L, smem, a.params, tx, ty);
      __syncthreads();
    }
    // ---- (r, g, b, sigma) per sample
    for (int e = tid; e < kTileM * 4; e += kThreads) {
      const int64_t g = g0 + (e >> 2);
      if (g < a.M) a.out[g * 4 + (e & 3)] = smem[L.off_out + e];
    }
    __syncthreads();
  }
}

static int validate_program(const dexnerf_mlp_program* p, bool query) {
  DN_REQUIRE(p, "mlp: null program");
  DN_REQUIRE(p->n_ops >= 1 && p->n_ops <= DEXNERF_MAX_OPS, "mlp: bad op count %d", p->n_ops);
  DN_REQUIRE(p->dim_xyz >= 1 && p->dim_dir >= 0, "mlp: bad input dims");
  DN_REQUIRE(p->max_width >= 1 && p->max_width <= 1024, "mlp: bad max_width");
  for (int i = 0; i < p->n_ops; ++i) {
    const dexnerf_op& op = p->ops[i];
    DN_REQUIRE(op.out_dim >= 1 && op.out_dim <= kMaxOut, "mlp: op %d out_dim %d unsupported", i, op.out_dim);
    DN_REQUIRE(op.dst != op.src0 && op.dst != op.src1, "mlp: op %d writes its own source", i);
    DN_REQUIRE(op.dst >= DEXNERF_BUF_A && op.dst <= DEXNERF_OUT_ALL, "mlp: op %d bad dst", i);
    if (op.dst == DEXNERF_BUF_A || op.dst == DEXNERF_BUF_B)
      DN_REQUIRE(op.out_dim <= p->max_width, "mlp: op %d wider than max_width", i);
    if (op.dst == DEXNERF_OUT_RGB) DN_REQUIRE(op.out_dim == 3, "mlp: OUT_RGB needs 3 channels");
    if (op.dst == DEXNERF_OUT_SIGMA) DN_REQUIRE(op.out_dim == 1, "mlp: OUT_SIGMA needs 1 channel");
    if (op.dst == DEXNERF_OUT_ALL) DN_REQUIRE(op.out_dim == 4, "mlp: OUT_ALL needs 4 channels");
  }
  if (query) {
    const int dx = (p->include_xyz ? 3 : 0) + 6 * p->Lx, dd = (p->include_dir ? 3 : 0) + 6 * p->Ld;
    DN_REQUIRE(dx == p->dim_xyz, "mlp_query: dim_xyz %d does not match the encoder (%d)", p->dim_xyz, dx);
    DN_REQUIRE(p->dim_dir == 0 || dd == p->dim_dir, "mlp_query: dim_dir mismatch");
  }
  return 0;
}

template <int MODE>
static int launch(const dexnerf_mlp_program* prog, const SimtArgs& a, cudaStream_t stream) {
  const SimtLayout L = make_layout(*prog);
  const size_t smem = sizeof(float) * (size_t)L.total_floats;
  DN_REQUIRE(smem <= 227 * 1024, "mlp: activations (%zu B) exceed shared memory", smem);
  DN_CUDA(cudaFuncSetAttribute(mlp_simt_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_sm = smem > 0 ? (int)((227 * 1024) / smem) : 1;
  int64_t tiles = ceil_div64(a.M, kTileM);
  int64_t grid = (int64_t)kNumSMs * (per_sm < 1 ? 1 : per_sm);
  if (grid > tiles) grid = tiles;
  mlp_simt_kernel<MODE><<<(int)grid, kThreads, smem, stream>>>(*prog, L, a);
  DN_CHECK_LAUNCH("mlp_simt");
  return 0;
}

}  // namespace dexnerf

using namespace dexnerf;

extern "C" DEXNERF_API int dexnerf_mlp_forward(const dexnerf_mlp_program* prog, const float* params,
                                   const float* x, int64_t M, float* out, void* stream) {
  if (M <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  if (int rc = validate_program(prog, false)) return rc;
  DN_REQUIRE(params && x && out, "mlp_forward: null pointer");
  SimtArgs a{};
  a.params = params; a.x = x; a.M = M; a.out = out; a.S = 1;
  return launch<0>(prog, a, (cudaStream_t)stream);
}

extern "C" DEXNERF_API int dexnerf_mlp_query(const dexnerf_mlp_program* prog, const float* params,
                                 const float* ro, const float* rd, const float* viewdirs,
                                 const float* z, int64_t n, int S, float* rf, void* stream) {
  if (n <= 0) return 0;   // an empty batch (null data pointers) is a no-op
  if (int rc = validate_program(prog, true)) return rc;
  DN_REQUIRE(params && ro && rd && z && rf, "mlp_query: null pointer");
  DN_REQUIRE(prog->dim_dir == 0 || viewdirs, "mlp_query: the model takes view directions but viewdirs is null");
  DN_REQUIRE(S >= 1, "mlp_query: S < 1");
  SimtArgs a{};
  a.params = params; a.ro = ro; a.rd = rd; a.viewdirs = viewdirs; a.z = z; a.S = S;
  a.M = n * (int64_t)S; a.out = rf;
  return launch<1>(prog, a, (cudaStream_t)stream);
}
