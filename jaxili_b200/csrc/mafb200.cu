// mafb200.cu -- C ABI (include/mafb200.h), planner and launchers of the B200-native MAF hot path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include "../../include/mafb200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "flow_kernel.cuh"
#include "chain_kernel.cuh"
#include "opt_kernels.cuh"
#include "sample_kernel.cuh"
#include "dp_kernel.cuh"
#include "hmc_kernels.cuh"
#include "wide_kernels.cuh"
#include "wide_tc.cuh"

using namespace mafb;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

constexpr int kSmemLimit = 232448;    // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int kSmemLimit2 = 115712;   // two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2
constexpr int kFlowThreads = 512;     // one CTA per SM, 32-row tiles
constexpr int kFlowThreads2 = 256;    // two CTAs per SM, 16-row tiles

inline int r4(int v) { return (v + 3) & ~3; }

}  // namespace

struct mafb200_handle_s {
  MafDesc desc{};
  int device = 0, n_sm = 0;
  Plan plan{};                  // one CTA per SM
  Plan plan2{};                 // two CTAs per SM (plan2.P == 0 when it does not fit)
  int occ_mode = 0;             // 0 auto, 1 / 2 forced (env MAFB200_OCC)
  int flow_kernel = 0;          // 0 auto (chain for training, phased otherwise), 1 phased always, 2 chain wherever it applies
  SamplePlan splan{};
  SamplePackArgs spack{};
  int dims[MAFB_MAX_LIN + 1]{};
  std::vector<uint8_t> mask_host;
  // device buffers
  uint8_t* mask = nullptr;
  float* wimg = nullptr;
  float* simg = nullptr;
  int* spos = nullptr;
  float* stdc = nullptr;        // [shift D][inv_scale D][mean C][std C][scale D]
  float* partial = nullptr;     // [n_sm][P]
  float* partial_loss = nullptr;
  float* normpart = nullptr;
  float* normsum = nullptr;     // [1] sum of the partials (large models: summed once, not by every block)
  unsigned int* done_counter = nullptr;
  unsigned long long* grid_bar = nullptr;   // step_tail_kernel: [0] launches so far, [1 + b] norm word of block b
  int tail_ok = -1;             // step_tail_kernel co-residency (-1: not probed yet)
  float logdet_const = 0.f;
  int n_norm = 0;
  long long* trace = nullptr;   // debug: device buffer of 1001 int64 (set by mafb200_debug_trace)
  // data parallel over NVLink peer memory (dp_kernel.cuh)
  int dp_rank = -1, dp_world = 0;
  const int* dp_step = nullptr;                    // device optax count shared by the step's kernels
  float* dp_sym = nullptr;                         // own: [2][r4(P) + 4] gradient slots, then flags
  unsigned int* dp_flags = nullptr;                // own flag array (inside dp_sym allocation)
  float* dp_peer_sym[DP_MAX_WORLD] = {};           // mapped peers (own pointer at [rank])
  unsigned int* dp_error = nullptr;
  // wide (layer-by-layer) path: used when the MADE does not fit shared memory
  bool wide = false;
  int wide_core = 1;            // 1: tcgen05 TF32 GEMM core, 0: FP32 FFMA core (exact)
  float* wm = nullptr;          // mask-folded copy of the parameters (flat layout)
  long long wcap = 0;           // rows the scratch below is sized for
  float* w_a0 = nullptr;        // [L][cap][K0]   per-layer MADE input [x_k | y']
  float* w_h = nullptr;         // [L][sum H][cap] hidden activations, level-major inside a layer
  float* w_d = nullptr;         // same shape: act'(a) (non-ReLU only)
  float* w_s = nullptr;         // [L][cap][D]
  float* w_v = nullptr;         // [L][cap][D]
  float* w_o = nullptr;         // [cap][2D]
  float* w_u = nullptr;         // [cap][D]  base-space point, then grad wrt layer output
  float* w_gx = nullptr;        // [cap][D]
  float* w_ld = nullptr;        // [cap]
  float* w_g[2] = {nullptr, nullptr};   // [cap][max H] hidden-gradient ping-pong
  float* w_g0 = nullptr;        // [cap][K0] full input gradient of a layer's first linear (conditioner-gradient mode)
  float* w_zero = nullptr;      // [cap][K0] zeros (addend of that product)
  float* w_dy = nullptr;        // [cap][C]  conditioner gradient summed over the layers
  long long wcap_dy = 0;
  float* eval_lp = nullptr;     // log_prob scratch of mafb200_eval_accumulate
  long long eval_cap = 0;
  // optional CUDA-event bracket around the flow kernel alone (roofline measurement in bench.py)
  bool timing = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double flow_ms_sum = 0.0;
  long long flow_launches = 0;
  bool ev_pending = false;
};

namespace {

// ------------------------------------------------------------------ masks from degrees
// jaxili/model.py:621-657.  deg_h[l] = degrees of hidden level l+1; input degrees arange(D).
void build_masks(const MafDesc& d, const int32_t* degrees, std::vector<uint8_t>& mask, const int* dims,
                 int n_lin, int ppl) {
  const int D = d.n_in, L = d.n_layers;
  int hsum = 0;
  for (int i = 0; i < d.n_hidden; ++i) hsum += d.hidden[i];
  mask.assign((size_t)L * ppl, 1);
  for (int k = 0; k < L; ++k) {
    const int32_t* dk = degrees + (size_t)k * hsum;
    std::vector<std::vector<int>> deg(n_lin + 1);
    deg[0].resize(D);
    for (int j = 0; j < D; ++j) deg[0][j] = j;
    int off = 0;
    for (int l = 1; l < n_lin; ++l) {
      deg[l].assign(dk + off, dk + off + d.hidden[l - 1]);
      off += d.hidden[l - 1];
    }
    deg[n_lin] = deg[0];
    size_t p = (size_t)k * ppl;
    for (int i = 0; i < n_lin; ++i) {
      const int K = dims[i], N = dims[i + 1];
      for (int a = 0; a < K; ++a)
        for (int b = 0; b < N; ++b) {
          uint8_t m;
          if (i == n_lin - 1) m = deg[i][a] < deg[n_lin][b % D];          // strict, duplicated (mu|logp)
          else if (i == 0) m = (a >= D) ? 1 : (deg[0][a] <= deg[1][b]);   // conditioning rows all ones
          else m = deg[i][a] <= deg[i + 1][b];
          mask[p + (size_t)a * N + b] = m;
        }
      p += (size_t)K * N + N;
    }
  }
}

// ------------------------------------------------------------------ flow-kernel planner
bool plan_layout(Plan& P, int R, bool stash_all, int n_wbuf, int limit) {
  P.R = R;
  P.RP = R + 4;
  P.stash_all = stash_all ? 1 : 0;
  P.n_wbuf = n_wbuf;
  const int RP = P.RP;
  int o = 0;
  P.wbuf_stride = stash_all ? std::max(P.fwd_floats, P.bwd_floats) : P.img_floats;
  P.o_wbuf = o;
  o += n_wbuf * P.wbuf_stride;
  P.stage_stride = r4(R * (P.D + P.C));
  P.o_stage = o;
  o += 2 * P.stage_stride;
  P.xin_stride = (r4(P.K0) + 4) * RP;   // + the ones row (bias gradient)
  P.o_xin = o;
  o += P.L * P.xin_stride;
  int rows = 0, maxhp = 4;
  for (int i = 1; i < P.n_lin; ++i) {
    P.h_off[i] = rows * RP;
    rows += P.lin[i].KP + 4;               // + the ones row
    maxhp = std::max(maxhp, P.lin[i].KP);
  }
  for (int i = 1; i < P.n_lin; ++i) {
    P.d_off[i] = P.need_d ? rows * RP : 0;
    if (P.need_d) rows += P.lin[i].KP;
  }
  P.s_off = rows * RP;
  rows += P.DP;
  P.v_off = rows * RP;
  rows += P.DP;
  P.slot_stride = rows * RP;
  P.o_slot = o;
  o += (stash_all ? P.L : 1) * P.slot_stride;
  P.o_gu = o;
  o += P.DP * RP;
  P.o_gx = o;
  o += 2 * P.DP * RP;                     // ping-pong by layer parity (fused reverse elementwise step)
  P.o_ld = o;
  o += P.DP * RP;
  P.o_cst = o;
  o += r4(2 * P.D + 2 * P.C);
  P.o_red = o;
  o += 32;
  P.o_gy = o;
  o += (r4(P.K0) - P.DP + 4) * RP;        // conditioner rows of the padded MADE input (dy mode)
  // gradient buffers: [obuf | hidden-gradient buffers], twice for the chain kernel (set 1 only when every layer's
  // activations are stashed -- the chain kernel's reverse pass needs that anyway)
  P.o_obuf = o;
  P.o_gb = o;
  const int out_rows = r4(2 * P.D);
  P.o_g[0] = o + out_rows * RP;
  P.o_g[1] = P.o_g[0] + maxhp * RP;
  int grows = out_rows;
  for (int i = 0; i < P.n_lin; ++i) {
    if (i == P.n_lin - 1) P.g_off[i] = 0;
    else { P.g_off[i] = grows * RP; grows += P.lin[i].NP; }
  }
  grows = std::max(grows, out_rows + 2 * maxhp);
  P.gb_stride = grows * RP;
  o += (stash_all ? 2 : 1) * P.gb_stride;
  o += 8 * RP;                            // slack: strided tile loads of the chain kernel may run past a buffer's pad rows
  P.smem_bytes = MAFB_SMEM_HDR + o * 4;
  return P.smem_bytes <= limit;
}

// occ = CTAs per SM the plan is sized for (1: 512 threads, tiles up to 32 rows; 2: 256 threads,
// tiles up to 16 rows -- two independent CTAs hide each other's barrier and latency bubbles)
bool make_plan(const MafDesc& d, const int* dims, Plan& P, int occ) {
  std::memset(&P, 0, sizeof(P));
  P.D = d.n_in;
  P.C = d.n_cond;
  P.K0 = d.n_in + d.n_cond;
  P.L = d.n_layers;
  P.n_lin = d.n_hidden + 1;
  P.act = d.activation;
  P.reverse = d.use_reverse ? 1 : 0;
  P.DP = r4(P.D);
  P.need_d = d.activation != MAFB200_ACT_RELU;
  P.threads = occ == 2 ? kFlowThreads2 : kFlowThreads;
  int fo = 0, bo = 0, po = 0;
  for (int i = 0; i < P.n_lin; ++i) {
    LinDesc& ln = P.lin[i];
    ln.K = dims[i];
    ln.N = dims[i + 1];
    ln.KP = r4(ln.K);
    ln.NP = r4(ln.N);
    ln.w_off = fo;
    fo += ln.KP * ln.NP;          // zero rows K..KP-1: the chain kernel's reduction loops run over KP
    ln.b_off = fo;
    fo += ln.NP;
    ln.wt_off = bo;
    bo += ln.NP * ln.KP;          // zero rows N..NP-1
    ln.pw_off = po;
    po += ln.K * ln.N;
    ln.pb_off = po;
    po += ln.N;
  }
  P.fwd_floats = r4(fo);
  P.bwd_floats = r4(bo);
  P.img_floats = P.fwd_floats + P.bwd_floats;
  P.ppl = po;
  P.P = po * P.L;
  P.pgs = r4(P.P);
  static const int Rs[] = {32, 16, 8, 4};
  for (int R : Rs) {
    if (occ == 2 && R > 16) continue;
    for (int stash = 1; stash >= 0; --stash)
      for (int nb = 2; nb >= 1; --nb)
        if (plan_layout(P, R, stash != 0, nb, occ == 2 ? kSmemLimit2 : kSmemLimit)) return true;
  }
  return false;
}

// model description only (no shared-memory layout): what the wide path needs of Plan
void make_plan_dims_only(const MafDesc& d, const int* dims, Plan& P) {
  std::memset(&P, 0, sizeof(P));
  P.D = d.n_in;
  P.C = d.n_cond;
  P.K0 = d.n_in + d.n_cond;
  P.L = d.n_layers;
  P.n_lin = d.n_hidden + 1;
  P.act = d.activation;
  P.reverse = d.use_reverse ? 1 : 0;
  P.DP = r4(P.D);
  P.need_d = d.activation != MAFB200_ACT_RELU;
  int po = 0;
  for (int i = 0; i < P.n_lin; ++i) {
    LinDesc& ln = P.lin[i];
    ln.K = dims[i];
    ln.N = dims[i + 1];
    ln.KP = ln.K;
    ln.NP = ln.N;
    ln.pw_off = po;
    po += ln.K * ln.N;
    ln.pb_off = po;
    po += ln.N;
  }
  P.ppl = po;
  P.P = po * P.L;
  P.pgs = r4(P.P);
}

// ------------------------------------------------------------------ sampler planner + sorted positions
bool make_sample_plan(const MafDesc& d, const int* dims, SamplePlan& S) {
  std::memset(&S, 0, sizeof(S));
  S.D = d.n_in;
  S.C = d.n_cond;
  S.L = d.n_layers;
  S.n_lin = d.n_hidden + 1;
  S.act = d.activation;
  S.reverse = d.use_reverse ? 1 : 0;
  if (S.D > MAFB_MAX_DEG) return false;
  S.width[0] = dims[0];
  for (int l = 1; l < S.n_lin; ++l) S.width[l] = r4(dims[l] + S.D);
  S.width[S.n_lin] = r4(2 * S.D);
  int o = 0;
  for (int i = 0; i < S.n_lin; ++i) {
    S.w_off[i] = o;
    o += S.width[i] * S.width[i + 1];
    o = r4(o);
    S.b_off[i] = o;
    o += S.width[i + 1];
    o = r4(o);
  }
  S.tab_off = o;
  o += r4((S.n_lin - 1) * (S.D + 1));
  S.img_floats = o;
  S.rows_per_warp = 32;
  S.RP = 36;
  int w = 0;
  for (int l = 1; l < S.n_lin; ++l) {
    S.lvl_off[l] = w;
    w += S.width[l] * S.RP;
  }
  S.y_off = w;
  w += std::max(4, r4(S.C)) * S.RP;
  S.u_off = w;
  w += r4(S.D) * S.RP;
  S.x_off = w;
  w += r4(S.D) * S.RP;
  S.warp_stride = w;
  // prefer many warps (latency hiding) over double-buffered images
  static const int kWarps[] = {8, 6, 4, 3, 2, 1};
  for (int warps : kWarps)
    for (int n_img = 2; n_img >= 1; --n_img) {
      S.n_img = n_img;
      S.o_img = 0;
      S.o_warp = n_img * S.img_floats;
      S.warps = warps;
      S.smem_bytes = 64 + (S.o_warp + S.warps * S.warp_stride) * 4;
      if (S.smem_bytes <= kSmemLimit) return true;
    }
  return false;
}

}  // namespace

// ~20 s at 1.9 GHz: long enough for a rank that is busy on the host (checkpoint, validation) to arrive, short
// enough that a dead peer ends in an error word + NaN loss instead of a hung GPU
constexpr long long kDpTimeoutCycles = 40000000000LL;
static size_t dp_slot_floats(const Plan& P) { return (size_t)((P.P + 3) & ~3) + 4; }
// symmetric buffer (floats): [2][sf] own slots (pull protocol) | [8] flags | [2][8][sf] 64-bit {value, epoch}
// words = receive area of the push protocol of step_tail_kernel
static size_t dp_recv_off(const Plan& P) { return 2 * dp_slot_floats(P) + 8; }
static size_t dp_total_floats(const Plan& P) { return dp_recv_off(P) + 2 * (2 * DP_MAX_WORLD * dp_slot_floats(P)); }

// Which fused flow kernel serves a launch (mode: 0 auto, 1 barrier-phased always, 2 chain wherever it applies;
// mafb200_set_flow_kernel / env MAFB200_FLOW=auto|phased|chain).  The chain kernel needs every layer's
// activations stashed for the reverse pass; the conditioner-gradient mode stays on the barrier-phased kernel.
// auto: chain for the training pass, barrier-phased for forward-only launches -- measured on B200
// (tests/ab_flow.py): log_prob is 10-15 % faster on the phased kernel (16 warps on the products, no
// weight-gradient work to overlap), the training pass 7-9 % faster on the chain kernel.
// Programmatic dependent launch between the two kernels of a training step (and into the next step's flow kernel):
// the dependent grid's CTAs are placed as the preceding grid's CTAs retire and run their independent prologue
// (shared-memory initialisation; the tail's operand loads) under its drain.  Measured on B200 (c2, graph replay):
// 57.32 us per step with, 57.37 us without -- the 185 KB / 61 k-register flow CTA and the tail blocks cannot share
// an SM, and the flow CTAs still enter 3.8 us after the tail's last block has left (4.0 us without): nothing
// overlaps; kept behind MAFB200_PDL=1 (default off).
static bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MAFB200_PDL"); return e && atoi(e) != 0; }();
  return on;
}

static bool use_chain(int mode, const Plan& P, const FlowArgs& A, bool bwd) {
  if (mode == 1 || P.threads != kFlowThreads) return false;
  if (!bwd) return mode == 2;
  return P.stash_all && !A.out_dy;
}

template <bool BWD>
static void launch_flow(int mode, const Plan& P, const FlowArgs& A, int grid, cudaStream_t st) {
  const bool relu = P.act == MAFB200_ACT_RELU;
  if (use_chain(mode, P, A, BWD)) {
    constexpr int nt = BWD ? CH_NT_BWD : CH_NT_FWD;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(nt);
    cfg.dynamicSmemBytes = P.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (P.n_lin == 3) {
      if (relu) cudaLaunchKernelEx(&cfg, maf_chain_kernel<BWD, true, 3>, P, A);
      else cudaLaunchKernelEx(&cfg, maf_chain_kernel<BWD, false, 3>, P, A);
    } else {
      if (relu) cudaLaunchKernelEx(&cfg, maf_chain_kernel<BWD, true, 0>, P, A);
      else cudaLaunchKernelEx(&cfg, maf_chain_kernel<BWD, false, 0>, P, A);
    }
    return;
  }
#define LAUNCH(NT_, R_, L_) maf_flow_kernel<BWD, R_, NT_, L_><<<grid, NT_, P.smem_bytes, st>>>(P, A)
#define PICK_L(NT_, R_) do { if (P.n_lin == 3) LAUNCH(NT_, R_, 3); else LAUNCH(NT_, R_, 0); } while (0)
  if (P.threads == kFlowThreads2) {
    if (relu) PICK_L(kFlowThreads2, true); else PICK_L(kFlowThreads2, false);
  } else {
    if (relu) PICK_L(kFlowThreads, true); else PICK_L(kFlowThreads, false);
  }
#undef PICK_L
#undef LAUNCH
}


// ====================================================================== wide path orchestration
namespace {

constexpr long long kWideChunk = 65536;   // rows per pass through the layer chain

int wide_ensure_scratch(mafb200_handle h, long long rows) {
  const long long cap = std::min<long long>(kWideChunk, (rows + 127) / 128 * 128);
  if (cap <= h->wcap) return 0;
  const Plan& P = h->plan;
  int hsum = 0, hmax = 4;
  for (int i = 1; i < P.n_lin; ++i) { hsum += P.lin[i].K; hmax = std::max(hmax, P.lin[i].K); }
  float** bufs[] = {&h->w_a0, &h->w_h, &h->w_d, &h->w_s, &h->w_v, &h->w_o, &h->w_u, &h->w_gx, &h->w_ld,
                    &h->w_g[0], &h->w_g[1]};
  for (float** b : bufs) { cudaFree(*b); *b = nullptr; }
  h->wcap = 0;
  const size_t c = (size_t)cap;
  CU(cudaMalloc(&h->w_a0, (size_t)P.L * c * P.K0 * 4));
  CU(cudaMalloc(&h->w_h, (size_t)P.L * c * hsum * 4));
  if (P.need_d) CU(cudaMalloc(&h->w_d, (size_t)P.L * c * hsum * 4));
  CU(cudaMalloc(&h->w_s, (size_t)P.L * c * P.D * 4));
  CU(cudaMalloc(&h->w_v, (size_t)P.L * c * P.D * 4));
  CU(cudaMalloc(&h->w_o, c * 2 * P.D * 4));
  CU(cudaMalloc(&h->w_u, c * P.D * 4));
  CU(cudaMalloc(&h->w_gx, c * P.D * 4));
  CU(cudaMalloc(&h->w_ld, c * 4));
  CU(cudaMalloc(&h->w_g[0], c * hmax * 4));
  CU(cudaMalloc(&h->w_g[1], c * hmax * 4));
  h->wcap = cap;
  return 0;
}

template <bool A_KFAST, bool B_KFAST, int EPI>
int wide_launch_gemm(const GemmArgs& g, bool relu, int splits, cudaStream_t st) {
  const dim3 block(256);
#define WG(BN_)                                                                                          \
  do {                                                                                                   \
    const dim3 grid((g.N + BN_ - 1) / BN_, (g.M + 127) / 128, splits);                                   \
    if (relu) wide_gemm_kernel<BN_, A_KFAST, B_KFAST, EPI, true><<<grid, block, 0, st>>>(g);             \
    else wide_gemm_kernel<BN_, A_KFAST, B_KFAST, EPI, false><<<grid, block, 0, st>>>(g);                 \
  } while (0)
  if (g.N > 64) WG(128);
  else if (g.N > 32) WG(64);
  else WG(32);
#undef WG
  CU(cudaGetLastError());
  return 0;
}


// ---------------------------------------------------------------- tcgen05 core: tensor maps + launch
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// FP32 matrix [outer][inner] with row stride ld floats; box = box_inner x box_outer, 128-byte swizzle
int make_tmap(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_inner,
              int box_outer, bool mn_major) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return 0;
}

// A_MN / B_MN: operand stored with the OUTPUT index contiguous ([reduction][output] row-major)
template <bool A_MN, bool B_MN, int EPI>
int wide_launch_tc(const GemmArgs& g, bool relu, int splits, cudaStream_t st) {
  CUtensorMap ta, tb;
  // K-major: inner = reduction (g.K), outer = output rows; MN-major: inner = output, outer = reduction
#define WTC(BM_, BN_)                                                                                      \
  do {                                                                                                     \
    if (A_MN ? make_tmap(&ta, g.A, g.M, g.K, g.lda, 32, 32, true) : make_tmap(&ta, g.A, g.K, g.M, g.lda, 32, 128, false)) \
      return 1;                                                                                            \
    if (B_MN ? make_tmap(&tb, g.B, g.N, g.K, g.ldb, 32, 32, true) : make_tmap(&tb, g.B, g.K, g.N, g.ldb, 32, BN_, false)) \
      return 1;                                                                                            \
    using Cfg = TcCfg<BM_, BN_>;                                                                           \
    const dim3 grid((g.N + BN_ - 1) / BN_, (g.M + BM_ - 1) / BM_, splits);                                 \
    static bool attr_t = false, attr_f = false;                                                            \
    if (relu) {                                                                                            \
      auto kfn = wide_tc_gemm_kernel<BM_, BN_, A_MN, B_MN, EPI, true>;                                     \
      if (!attr_t) { CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); attr_t = true; } \
      kfn<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, g);                                         \
    } else {                                                                                               \
      auto kfn = wide_tc_gemm_kernel<BM_, BN_, A_MN, B_MN, EPI, false>;                                    \
      if (!attr_f) { CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)); attr_f = true; } \
      kfn<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, g);                                         \
    }                                                                                                      \
  } while (0)
  static const int tile_mode = [] { const char* e = getenv("MAFB200_TC_TILE"); return e ? atoi(e) : 0; }();
  if (g.N >= 256 && g.M >= 256 && tile_mode == 1) WTC(128, 256);
  else if (g.N >= 256 && g.M >= 256) WTC(256, 256);
  else if (g.N > 64) WTC(128, 128);
  else if (g.N > 32) WTC(128, 64);
  else WTC(128, 32);
#undef WTC
  CU(cudaGetLastError());
  return 0;
}

// dispatch: <A reduction-contiguous?, B reduction-contiguous?> of the FFMA core <-> <A_MN, B_MN> of the TC core
template <bool A_KFAST, bool B_KFAST, int EPI>
int wide_gemm(mafb200_handle h, const GemmArgs& g, bool relu, int splits, cudaStream_t st) {
  if (h->wide_core == 1) return wide_launch_tc<!A_KFAST, !B_KFAST, EPI>(g, relu, splits, st);
  return wide_launch_gemm<A_KFAST, B_KFAST, EPI>(g, relu, splits, st);
}

// one chunk of rows through the whole flow; bwd: also the reverse pass accumulating into grad / loss
// dy != nullptr: conditioner-gradient mode (jaxili/posterior/mcmc_posterior.py:139-159): reverse pass without weight
// gradients, the first linear's input gradient covers all K0 columns, its y-part is summed over the layers
int wide_chunk(mafb200_handle h, const float* x, const float* y, long long rows, float inv_b, bool bwd,
               float* out_logp, float* out_u, float* out_logdet, float* loss, float* grad, cudaStream_t st,
               float* dy = nullptr) {
  const Plan& P = h->plan;
  const bool relu = P.act == MAFB200_ACT_RELU;
  const int D = P.D, K0 = P.K0, L = P.L, n_lin = P.n_lin;
  const size_t cap = (size_t)h->wcap;
  int hsum = 0;
  for (int i = 1; i < n_lin; ++i) hsum += P.lin[i].K;
  WideCommon w{D, P.C, K0, P.reverse, rows, inv_b, h->logdet_const};
  const bool stash = bwd;
  auto a0 = [&](int k) { return h->w_a0 + (size_t)(stash ? k : (k & 1)) * cap * K0; };
  auto hbuf = [&](float* base, int k, int lvl) {   // hidden level lvl (1..n_hidden) of layer k
    size_t off = (size_t)(stash ? k : 0) * cap * hsum;
    for (int i = 1; i < lvl; ++i) off += cap * P.lin[i].K;
    return base + off;
  };
  const long long a0_stride = (long long)cap * K0;
  {
    const long long n = rows * K0;
    wide_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, x, y, h->stdc, h->w_a0, a0_stride,
                                                                  stash ? L : 2, h->w_ld);
    CU(cudaGetLastError());
  }
  // ---------------- forward
  for (int k = 0; k < L; ++k) {
    const float* wl = h->wm + (size_t)k * P.ppl;
    const float* in = a0(k);
    int ld_in = K0;
    for (int i = 0; i < n_lin; ++i) {
      const LinDesc& ln = P.lin[i];
      GemmArgs g{};
      g.A = in; g.lda = ld_in;
      g.B = wl + ln.pw_off; g.ldb = ln.N;
      g.M = (int)rows; g.N = ln.N; g.K = ln.K;
      g.bias = wl + ln.pb_off;
      g.act = P.act;
      if (i < n_lin - 1) {
        g.C = hbuf(h->w_h, k, i + 1); g.ldc = ln.N;
        if (P.need_d && bwd) { g.aux_out = hbuf(h->w_d, k, i + 1); g.ldaux_out = ln.N; }
        if (wide_gemm<true, false, EPI_BIAS_ACT>(h, g, relu, 1, st)) return 1;
        in = g.C; ld_in = ln.N;
      } else {
        g.C = h->w_o; g.ldc = ln.N;
        if (wide_gemm<true, false, EPI_BIAS>(h, g, relu, 1, st)) return 1;
      }
    }
    float* xnext = (k + 1 < L) ? a0(k + 1) : h->w_u;
    const int xld = (k + 1 < L) ? K0 : D;
    wide_affine_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(
        w, h->w_o, a0(k), bwd ? h->w_s + (size_t)k * cap * D : nullptr, bwd ? h->w_v + (size_t)k * cap * D : nullptr,
        xnext, xld, h->w_ld);
    CU(cudaGetLastError());
  }
  if (out_u) CU(cudaMemcpyAsync(out_u, h->w_u, (size_t)rows * D * 4, cudaMemcpyDeviceToDevice, st));
  if (out_logdet) CU(cudaMemcpyAsync(out_logdet, h->w_ld, (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
  wide_final_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w, h->w_u, h->w_ld, out_logp, (bwd && !dy) ? loss : nullptr,
                                                                bwd ? h->w_u : nullptr);
  CU(cudaGetLastError());
  if (!bwd) return 0;
  // ---------------- reverse pass
  const int splits = (int)std::max<long long>(1, std::min<long long>(64, rows / 1024));
  const int k_chunk = (int)(((rows + splits - 1) / splits + 31) / 32 * 32);
  for (int k = L - 1; k >= 0; --k) {
    const float* wl = h->wm + (size_t)k * P.ppl;
    float* gl = grad + (size_t)k * P.ppl;
    const unsigned char* ml = h->mask + (size_t)k * P.ppl;
    {
      const long long n = rows * D;
      wide_bwd_elem_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
          w, h->w_u, h->w_s + (size_t)k * cap * D, h->w_v + (size_t)k * cap * D, h->w_o, h->w_gx);
      CU(cudaGetLastError());
    }
    const float* G = h->w_o;
    for (int i = n_lin - 1; i >= 0; --i) {
      const LinDesc& ln = P.lin[i];
      const float* Ain = (i == 0) ? a0(k) : hbuf(h->w_h, k, i);
      if (!dy) {  // weight gradient: dW = M * (A^T G), rows split over blockIdx.z
        GemmArgs g{};
        g.A = Ain; g.lda = ln.K;
        g.B = G; g.ldb = ln.N;
        g.C = gl + ln.pw_off; g.ldc = ln.N;
        g.M = ln.K; g.N = ln.N; g.K = (int)rows;
        g.mask = ml + ln.pw_off;
        g.k_chunk = k_chunk;
        if (wide_gemm<false, false, EPI_ATOMIC_MASK>(h, g, relu, (int)((rows + k_chunk - 1) / k_chunk), st)) return 1;
        if (i == n_lin - 1) {   // hidden levels get their bias gradient from the producing GEMM's epilogue
          const int rpb = 256;
          wide_colsum_kernel<<<dim3((ln.N + 31) / 32, (unsigned)((rows + rpb - 1) / rpb)), 256, 0, st>>>(
              G, rows, ln.N, rpb, gl + ln.pb_off);
          CU(cudaGetLastError());
        }
      }
      {  // input gradient
        GemmArgs g{};
        g.A = G; g.lda = ln.N;
        g.B = wl + ln.pw_off; g.ldb = ln.N;
        g.M = (int)rows; g.K = ln.N;
        g.act = P.act;
        if (i > 0) {
          g.N = ln.K;
          g.C = h->w_g[i & 1]; g.ldc = ln.K;
          g.aux = relu ? hbuf(h->w_h, k, i) : hbuf(h->w_d, k, i); g.ldaux = ln.K;
          g.colsum = dy ? nullptr : gl + P.lin[i - 1].pb_off;      // db_{i-1} = column sums of g_a_{i-1}
          if (wide_gemm<true, true, EPI_MUL_DERIV>(h, g, relu, 1, st)) return 1;
          G = g.C;
        } else if (!dy) {
          g.N = D;                       // only the x part of the input carries gradient further
          g.C = h->w_u; g.ldc = D;
          g.aux = h->w_gx; g.ldaux = D;
          if (wide_gemm<true, true, EPI_ADD>(h, g, relu, 1, st)) return 1;
        } else {
          g.N = K0;                      // all columns: [d/dx | d/dy'] of this layer's MADE
          g.C = h->w_g0; g.ldc = K0;
          g.aux = h->w_zero; g.ldaux = K0;
          if (wide_gemm<true, true, EPI_ADD>(h, g, relu, 1, st)) return 1;
          const long long nx = rows * D, ny = rows * P.C;
          wide_dx_pick_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, st>>>(w, h->w_g0, h->w_gx, h->w_u);
          wide_dy_accum_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(w, h->w_g0, h->w_dy, k == L - 1 ? 1 : 0);
          CU(cudaGetLastError());
        }
      }
    }
  }
  if (dy) {
    const long long ny = rows * P.C;
    wide_dy_final_kernel<<<(unsigned)((ny + 255) / 256), 256, 0, st>>>(w, h->w_dy, h->stdc, dy);
    CU(cudaGetLastError());
  }
  return 0;
}

int wide_ensure_dy_scratch(mafb200_handle h) {
  if (h->wcap_dy >= h->wcap && h->w_g0) return 0;
  const Plan& P = h->plan;
  cudaFree(h->w_g0); cudaFree(h->w_zero); cudaFree(h->w_dy);
  h->w_g0 = h->w_zero = h->w_dy = nullptr;
  h->wcap_dy = 0;
  const size_t c = (size_t)h->wcap;
  CU(cudaMalloc(&h->w_g0, c * P.K0 * 4));
  CU(cudaMalloc(&h->w_zero, c * P.K0 * 4));
  CU(cudaMemset(h->w_zero, 0, c * P.K0 * 4));
  CU(cudaMalloc(&h->w_dy, c * std::max(P.C, 1) * 4));
  h->wcap_dy = h->wcap;
  return 0;
}

int wide_log_prob_grad_y(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                         float* out_logp, float* out_dy, int32_t flags, cudaStream_t st) {
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, st)) return 1;
  if (wide_ensure_scratch(h, B)) return 1;
  if (wide_ensure_dy_scratch(h)) return 1;
  const Plan& P = h->plan;
  for (int64_t r0 = 0; r0 < B; r0 += h->wcap) {
    const long long rows = std::min<long long>(h->wcap, B - r0);
    // 1/B := -1: the "loss" is sum_b log p_b, whose y-gradient is per row (as on the fused path)
    if (wide_chunk(h, x + r0 * P.D, y + r0 * P.C, rows, -1.f, true, out_logp ? out_logp + r0 : nullptr, nullptr,
                   nullptr, nullptr, nullptr, st, out_dy + r0 * P.C))
      return 1;
  }
  return 0;
}

// inverse flow of the wide path: per layer D sequential passes through the MADE's GEMM chain
int wide_sample(mafb200_handle h, const float* params, const float* u, uint64_t seed, uint64_t row_offset,
                const float* y_obs, int64_t n_obs, int64_t rows_per_obs, int64_t N, float* out, float* out_logdet,
                int32_t flags, cudaStream_t st) {
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, st)) return 1;
  if (wide_ensure_scratch(h, N)) return 1;
  const Plan& P = h->plan;
  const bool relu = P.act == MAFB200_ACT_RELU;
  const int D = P.D, K0 = P.K0, L = P.L, n_lin = P.n_lin;
  const size_t cap = (size_t)h->wcap;
  for (int64_t r0 = 0; r0 < N; r0 += h->wcap) {
    const long long rows = std::min<long long>(h->wcap, N - r0);
    WideCommon w{D, P.C, K0, P.reverse, rows, 1.f, h->logdet_const};
    WideInvArgs a{u ? u + r0 * D : nullptr, seed, row_offset, y_obs, n_obs, rows_per_obs, r0};
    const unsigned gr = (unsigned)((rows + 255) / 256), ge = (unsigned)((rows * D + 255) / 256);
    wide_inv_prep_kernel<<<gr, 256, 0, st>>>(w, a, h->stdc, h->w_a0, h->w_u, h->w_ld);
    CU(cudaGetLastError());
    for (int k = L - 1; k >= 0; --k) {
      const float* wl = h->wm + (size_t)k * P.ppl;
      for (int d = 0; d < D; ++d) {
        const float* in = h->w_a0;
        int ld_in = K0;
        size_t hoff = 0;
        for (int i = 0; i < n_lin; ++i) {
          const LinDesc& ln = P.lin[i];
          GemmArgs g{};
          g.A = in; g.lda = ld_in;
          g.B = wl + ln.pw_off; g.ldb = ln.N;
          g.M = (int)rows; g.N = ln.N; g.K = ln.K;
          g.bias = wl + ln.pb_off;
          g.act = P.act;
          if (i < n_lin - 1) {
            g.C = h->w_h + hoff; g.ldc = ln.N;
            hoff += cap * ln.N;
            if (wide_gemm<true, false, EPI_BIAS_ACT>(h, g, relu, 1, st)) return 1;
            in = g.C; ld_in = ln.N;
          } else {
            g.C = h->w_o; g.ldc = ln.N;
            if (wide_gemm<true, false, EPI_BIAS>(h, g, relu, 1, st)) return 1;
          }
        }
        wide_inv_update_kernel<<<gr, 256, 0, st>>>(w, d, h->w_o, h->w_u, h->w_a0, h->w_ld);
        CU(cudaGetLastError());
      }
      wide_inv_next_kernel<<<ge, 256, 0, st>>>(w, h->w_a0, h->w_u);
      CU(cudaGetLastError());
    }
    wide_inv_final_kernel<<<ge, 256, 0, st>>>(w, h->w_u, h->w_ld, h->stdc, out + r0 * D,
                                               out_logdet ? out_logdet + r0 : nullptr);
    CU(cudaGetLastError());
  }
  return 0;
}

int wide_forward(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                 float* out_logp, float* out_u, float* out_logdet, int32_t flags, cudaStream_t st) {
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, st)) return 1;
  if (wide_ensure_scratch(h, B)) return 1;
  const Plan& P = h->plan;
  for (int64_t r0 = 0; r0 < B; r0 += h->wcap) {
    const long long rows = std::min<long long>(h->wcap, B - r0);
    if (wide_chunk(h, x + r0 * P.D, y + r0 * P.C, rows, 1.f, false, out_logp ? out_logp + r0 : nullptr,
                   out_u ? out_u + r0 * P.D : nullptr, out_logdet ? out_logdet + r0 : nullptr, nullptr, nullptr, st))
      return 1;
  }
  return 0;
}

int wide_loss_grad(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                   float inv_global_B, float* out_loss, float* out_grad, int32_t flags, cudaStream_t st) {
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, st)) return 1;
  if (wide_ensure_scratch(h, B)) return 1;
  const Plan& P = h->plan;
  CU(cudaMemsetAsync(out_grad, 0, (size_t)P.P * 4, st));
  CU(cudaMemsetAsync(out_loss, 0, 4, st));
  for (int64_t r0 = 0; r0 < B; r0 += h->wcap) {
    const long long rows = std::min<long long>(h->wcap, B - r0);
    if (wide_chunk(h, x + r0 * P.D, y + r0 * P.C, rows, inv_global_B, true, nullptr, nullptr, nullptr, out_loss,
                   out_grad, st))
      return 1;
  }
  return 0;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

const char* mafb200_last_error(void) { return g_err.c_str(); }
int mafb200_version(void) { return 100; }

int mafb200_set_standardization(mafb200_handle h, const float* shift, const float* scale,
                                const float* mean, const float* stdv) {
  if (!h) return fail("null handle");
  const int D = h->desc.n_in, C = h->desc.n_cond;
  std::vector<float> c(3 * D + 2 * C);
  float ld = 0.f;
  for (int j = 0; j < D; ++j) {
    const float sh = shift ? shift[j] : 0.f, sc = scale ? scale[j] : 1.f;
    c[j] = sh;
    c[D + j] = 1.0f / sc;                 // distrax.ScalarAffine: inv_scale = 1/scale
    c[2 * D + 2 * C + j] = sc;
    ld += -logf(fabsf(sc));               // inverse_log_det_jacobian summed (model.py:1094-1095)
  }
  for (int j = 0; j < C; ++j) {
    c[2 * D + j] = mean ? mean[j] : 0.f;
    c[2 * D + C + j] = stdv ? stdv[j] : 1.f;
  }
  h->logdet_const = ld;
  CU(cudaSetDevice(h->device));
  CU(cudaMemcpy(h->stdc, c.data(), c.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int mafb200_create(const MafDesc* desc, const int32_t* degrees, const float* shift, const float* scale,
                   const float* mean, const float* stdv, int device, mafb200_handle* out) {
  if (!desc || !degrees || !out) return fail("mafb200_create: null argument");
  const MafDesc& d = *desc;
  if (d.n_in < 2) return fail("n_in must be >= 2 (jaxili's randint(0, n_in-1) needs it)");
  if (d.n_cond < 0 || d.n_layers < 1) return fail("bad n_cond / n_layers");
  if (d.n_hidden < 1 || d.n_hidden > MAFB200_MAX_HIDDEN) return fail("layers must have 1..5 entries");
  if (d.activation < 0 || d.activation > MAFB200_ACT_LEAKY_RELU) return fail("unsupported activation");
  for (int i = 0; i < d.n_hidden; ++i)
    if (d.hidden[i] < 1) return fail("hidden sizes must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("no CUDA device: the MAF hot path has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("bad device index");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail("this library is built for sm_100a (B200) only");

  mafb200_handle h = new mafb200_handle_s();
  h->desc = d;
  h->device = device;
  h->n_sm = prop.multiProcessorCount;
  h->dims[0] = d.n_in + d.n_cond;
  for (int i = 0; i < d.n_hidden; ++i) h->dims[i + 1] = d.hidden[i];
  h->dims[d.n_hidden + 1] = 2 * d.n_in;
  const char* force_wide = getenv("MAFB200_FORCE_WIDE");
  if (!make_plan(d, h->dims, h->plan, 1) || (force_wide && atoi(force_wide))) {
    // MADE too large for the shared-memory-resident kernel: layer-by-layer GEMM path
    for (int i = 0; i <= d.n_hidden + 1; ++i)
      if (h->dims[i] % 4) {
        delete h;
        return fail("wide path needs n_in, n_cond and all hidden sizes to be multiples of 4");
      }
    make_plan_dims_only(d, h->dims, h->plan);
    h->wide = true;
    if (const char* e = getenv("MAFB200_WIDE_CORE")) h->wide_core = (strcmp(e, "ffma") == 0) ? 0 : 1;
  }
  if (h->wide || !make_plan(d, h->dims, h->plan2, 2) || !h->plan2.stash_all || h->plan2.R < 16) h->plan2.P = 0;
  if (const char* e = getenv("MAFB200_OCC")) h->occ_mode = atoi(e);
  if (const char* e = getenv("MAFB200_FLOW")) h->flow_kernel = strcmp(e, "phased") == 0 ? 1 : strcmp(e, "chain") == 0 ? 2 : 0;
  const bool have_sampler = !h->wide && make_sample_plan(d, h->dims, h->splan);
  if (!have_sampler) h->splan.img_floats = 0;
  const Plan& P = h->plan;
  build_masks(d, degrees, h->mask_host, h->dims, P.n_lin, P.ppl);

#define CUH(call)                                                           \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      mafb200_destroy(h);                                                   \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));      \
    }                                                                       \
  } while (0)
  CUH(cudaMalloc(&h->mask, P.P));
  CUH(cudaMemcpy(h->mask, h->mask_host.data(), P.P, cudaMemcpyHostToDevice));
  if (h->wide) {
    CUH(cudaMalloc(&h->wm, (size_t)P.P * 4));
  } else {
    CUH(cudaMalloc(&h->wimg, (size_t)P.L * P.img_floats * 4));
    CUH(cudaMemset(h->wimg, 0, (size_t)P.L * P.img_floats * 4));
  }
  CUH(cudaMalloc(&h->stdc, (3 * P.D + 2 * P.C) * 4));
  if (!h->wide) {
    CUH(cudaMalloc(&h->partial, (size_t)2 * h->n_sm * P.pgs * 4));
    CUH(cudaMemset(h->partial, 0, (size_t)2 * h->n_sm * P.pgs * 4));
  }
  CUH(cudaMalloc(&h->partial_loss, 2 * h->n_sm * 4));
  h->n_norm = (P.P + RED_IDX - 1) / RED_IDX;
  CUH(cudaMalloc(&h->normpart, h->n_norm * 4));
  CUH(cudaMemset(h->normpart, 0, h->n_norm * 4));
  CUH(cudaMalloc(&h->normsum, 4));
  CUH(cudaMalloc(&h->done_counter, 4));
  CUH(cudaMemset(h->done_counter, 0, 4));
  CUH(cudaMalloc(&h->grid_bar, (size_t)(h->n_norm + 2) * 8));
  CUH(cudaMemset(h->grid_bar, 0, (size_t)(h->n_norm + 2) * 8));

  if (have_sampler) {
    // degree-sorted positions and group tables
    const SamplePlan& S = h->splan;
    const int n_lin = S.n_lin, D = S.D;
    int hsum = 0, tot = 0;
    for (int i = 0; i < d.n_hidden; ++i) hsum += d.hidden[i];
    for (int l = 0; l <= n_lin; ++l) {
      h->spack.lvl_start[l] = tot;
      tot += h->dims[l];
    }
    h->spack.spos_per_layer = tot;
    std::vector<int> spos((size_t)P.L * tot);
    std::vector<float> simg((size_t)P.L * S.img_floats, 0.f);
    for (int k = 0; k < P.L; ++k) {
      int* pos = spos.data() + (size_t)k * tot;
      for (int a = 0; a < h->dims[0]; ++a) pos[a] = a;
      const int32_t* dk = degrees + (size_t)k * hsum;
      int off = 0;
      int* tab = reinterpret_cast<int*>(simg.data() + (size_t)k * S.img_floats + S.tab_off);
      for (int l = 1; l < n_lin; ++l) {
        const int H = h->dims[l];
        int* pl = pos + h->spack.lvl_start[l];
        int* T = tab + (l - 1) * (D + 1);
        int col = 0;
        T[0] = 0;
        for (int m = 0; m < D; ++m) {
          for (int a = 0; a < H; ++a)
            if (dk[off + a] == m) pl[a] = col++;
          col = (col + SN_ - 1) / SN_ * SN_;
          T[m + 1] = col;
        }
        if (col > S.width[l]) {
          mafb200_destroy(h);
          return fail("internal: sorted width overflow");
        }
        for (int a = 0; a < H; ++a)
          if (dk[off + a] < 0 || dk[off + a] >= D) {
            mafb200_destroy(h);
            return fail("degree out of range");
          }
        off += H;
      }
      int* po = pos + h->spack.lvl_start[n_lin];
      for (int n = 0; n < 2 * D; ++n) po[n] = n < D ? 2 * n : 2 * (n - D) + 1;
    }
    CUH(cudaMalloc(&h->spos, spos.size() * sizeof(int)));
    CUH(cudaMemcpy(h->spos, spos.data(), spos.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUH(cudaMalloc(&h->simg, simg.size() * 4));
    CUH(cudaMemcpy(h->simg, simg.data(), simg.size() * 4, cudaMemcpyHostToDevice));
  }

  // opt in to the large dynamic shared memory carve-out for every kernel we may launch
  if (!h->wide) {
    const int sb = P.smem_bytes, sb2 = h->plan2.P ? h->plan2.smem_bytes : 0;
#define SET_FLOW(B_, R_, L_)                                                                              \
  CUH(cudaFuncSetAttribute(maf_flow_kernel<B_, R_, kFlowThreads, L_>,                                     \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, sb));                             \
  if (sb2)                                                                                                \
    CUH(cudaFuncSetAttribute(maf_flow_kernel<B_, R_, kFlowThreads2, L_>,                                  \
                             cudaFuncAttributeMaxDynamicSharedMemorySize, sb2));
    SET_FLOW(true, true, 3) SET_FLOW(true, false, 3) SET_FLOW(false, true, 3) SET_FLOW(false, false, 3)
    SET_FLOW(true, true, 0) SET_FLOW(true, false, 0) SET_FLOW(false, true, 0) SET_FLOW(false, false, 0)
#undef SET_FLOW
#define SET_CHAIN(B_, R_, L_) \
  CUH(cudaFuncSetAttribute(maf_chain_kernel<B_, R_, L_>, cudaFuncAttributeMaxDynamicSharedMemorySize, sb));
    SET_CHAIN(true, true, 3) SET_CHAIN(true, false, 3) SET_CHAIN(false, true, 3) SET_CHAIN(false, false, 3)
    SET_CHAIN(true, true, 0) SET_CHAIN(true, false, 0) SET_CHAIN(false, true, 0) SET_CHAIN(false, false, 0)
#undef SET_CHAIN
  }
  if (have_sampler) {
    CUH(cudaFuncSetAttribute(maf_sample_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->splan.smem_bytes));
    CUH(cudaFuncSetAttribute(maf_sample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->splan.smem_bytes));
  }
#undef CUH
  *out = h;
  if (mafb200_set_standardization(h, shift, scale, mean, stdv)) {
    mafb200_destroy(h);
    *out = nullptr;
    return 1;
  }
  return 0;
}

int mafb200_destroy(mafb200_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->mask);
  cudaFree(h->wimg);
  cudaFree(h->simg);
  cudaFree(h->spos);
  cudaFree(h->stdc);
  cudaFree(h->partial);
  cudaFree(h->partial_loss);
  cudaFree(h->normpart);
  cudaFree(h->normsum);
  cudaFree(h->done_counter);
  cudaFree(h->grid_bar);
  for (int r = 0; r < h->dp_world; ++r)
    if (r != h->dp_rank && h->dp_peer_sym[r]) cudaIpcCloseMemHandle(h->dp_peer_sym[r]);
  cudaFree(h->dp_sym);
  cudaFree(h->dp_error);
  cudaFree(h->wm);
  cudaFree(h->w_a0); cudaFree(h->w_h); cudaFree(h->w_d); cudaFree(h->w_s); cudaFree(h->w_v);
  cudaFree(h->w_o); cudaFree(h->w_u); cudaFree(h->w_gx); cudaFree(h->w_ld);
  cudaFree(h->w_g[0]); cudaFree(h->w_g[1]);
  cudaFree(h->w_g0); cudaFree(h->w_zero); cudaFree(h->w_dy);
  cudaFree(h->eval_lp);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return 0;
}

int mafb200_debug_trace(mafb200_handle h, long long* trace_dev) {
  if (!h) return fail("null handle");
  h->trace = trace_dev;
  return 0;
}

int mafb200_flow_timing(mafb200_handle h, int32_t enable) {
  if (!h) return fail("null handle");
  CU(cudaSetDevice(h->device));
  if (enable && !h->ev0) {
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
  }
  h->timing = enable != 0;
  h->ev_pending = false;
  h->flow_ms_sum = 0.0;
  h->flow_launches = 0;
  return 0;
}

int mafb200_flow_elapsed(mafb200_handle h, double* total_ms, int64_t* launches) {
  if (!h || !total_ms || !launches) return fail("null argument");
  if (h->ev_pending) {
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->flow_ms_sum += ms;
    h->flow_launches += 1;
    h->ev_pending = false;
  }
  *total_ms = h->flow_ms_sum;
  *launches = h->flow_launches;
  return 0;
}

// debug / test hook: one GEMM of the wide path.  variant 0: C = A[MxK] . B[KxN] (+bias 0);
// 1: C = A[MxK] . B[NxK]^T; 2: C += A[KxM]^T . B[KxN] (split over K, mask all ones expected in `mask`).
int mafb200_debug_gemm(int32_t variant, int32_t core, const float* A, const float* B, float* C, int32_t M,
                       int32_t N, int32_t K, const float* zeros_n, const float* zeros_mn,
                       const unsigned char* mask, void* stream) {
  mafb200_handle_s fake;
  fake.wide_core = core;
  cudaStream_t st = (cudaStream_t)stream;
  GemmArgs g{};
  g.A = A; g.B = B; g.C = C; g.M = M; g.N = N; g.K = K; g.act = 0;
  if (variant == 0) {
    g.lda = K; g.ldb = N; g.ldc = N; g.bias = zeros_n;
    return wide_gemm<true, false, EPI_BIAS>(&fake, g, true, 1, st);
  } else if (variant == 1) {
    g.lda = K; g.ldb = K; g.ldc = N; g.aux = zeros_mn; g.ldaux = N;
    return wide_gemm<true, true, EPI_ADD>(&fake, g, true, 1, st);
  } else {
    g.lda = M; g.ldb = N; g.ldc = N; g.mask = mask;
    const int splits = std::max(1, std::min(8, K / 256));
    g.k_chunk = ((K + splits - 1) / splits + 31) / 32 * 32;
    return wide_gemm<false, false, EPI_ATOMIC_MASK>(&fake, g, true, (K + g.k_chunk - 1) / g.k_chunk, st);
  }
}

// ---- data parallel over peer memory ------------------------------------------------------------

int mafb200_dp_init(mafb200_handle h, int32_t rank, int32_t world, void* ipc_handle_out) {
  if (!h || !ipc_handle_out) return fail("null argument");
  if (world < 1 || world > DP_MAX_WORLD || rank < 0 || rank >= world) return fail("bad rank / world (max 8)");
  if (h->dp_sym) return fail("dp already initialised");
  CU(cudaSetDevice(h->device));
  const size_t bytes = dp_total_floats(h->plan) * 4;
  CU(cudaMalloc(&h->dp_sym, bytes));
  CU(cudaMemset(h->dp_sym, 0, bytes));
  CU(cudaMalloc(&h->dp_error, 4));
  CU(cudaMemset(h->dp_error, 0, 4));
  h->dp_flags = reinterpret_cast<unsigned int*>(h->dp_sym + 2 * dp_slot_floats(h->plan));
  h->dp_rank = rank;
  h->dp_world = world;
  h->dp_peer_sym[rank] = h->dp_sym;
  cudaIpcMemHandle_t hd;
  CU(cudaIpcGetMemHandle(&hd, h->dp_sym));
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  return 0;
}

int mafb200_dp_connect(mafb200_handle h, const void* all_handles) {
  if (!h || !all_handles || !h->dp_sym) return fail("dp not initialised");
  CU(cudaSetDevice(h->device));
  for (int r = 0; r < h->dp_world; ++r) {
    if (r == h->dp_rank) continue;
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, static_cast<const char*>(all_handles) + 64 * r, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    h->dp_peer_sym[r] = static_cast<float*>(p);
  }
  return 0;
}

// sums the ranks' reduced gradients (written by mafb200_loss_grad with MAFB200_DP_SLOT) into
// grad_out / loss_out and leaves the squared-norm partials current for mafb200_clip_adam
int mafb200_dp_allreduce(mafb200_handle h, float* grad_out, float* loss_out, void* stream) {
  if (!h || !grad_out || !loss_out) return fail("null argument");
  if (!h->dp_sym) return fail("dp not initialised");
  for (int r = 0; r < h->dp_world; ++r)
    if (!h->dp_peer_sym[r]) return fail("dp peers not connected");
  if (!h->dp_step) return fail("mafb200_dp_bind_step was not called");
  const Plan& P = h->plan;
  DpArgs a{};
  a.rank = h->dp_rank;
  a.world = h->dp_world;
  a.P = P.P;
  a.step = h->dp_step;
  const size_t sf = dp_slot_floats(P);
  a.slot_floats = (long long)sf;
  for (int r = 0; r < h->dp_world; ++r) {
    a.sym[r] = h->dp_peer_sym[r];
    a.flags_of[r] = reinterpret_cast<unsigned int*>(h->dp_peer_sym[r] + 2 * sf);
  }
  a.grad_out = grad_out;
  a.loss_out = loss_out;
  a.normpart = h->normpart;
  a.error = h->dp_error;
  a.timeout_cycles = kDpTimeoutCycles;
  dp_allreduce_kernel<<<h->n_norm, RED_IDX, 0, (cudaStream_t)stream>>>(a);
  CU(cudaGetLastError());
  return 0;
}

// the device int32 optax count (the `step` of mafb200_clip_adam) doubles as the exchange epoch
int mafb200_dp_bind_step(mafb200_handle h, const int32_t* step) {
  if (!h || !step) return fail("null argument");
  h->dp_step = step;
  return 0;
}

// forget all epochs (call on every rank, between host barriers, whenever the step counter restarts)
int mafb200_dp_reset(mafb200_handle h) {
  if (!h || !h->dp_sym) return fail("dp not initialised");
  CU(cudaSetDevice(h->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemset(h->dp_flags, 0, DP_MAX_WORLD * sizeof(unsigned int)));
  // receive area too: a rewound step counter must never match the epoch of a value left by an earlier run
  CU(cudaMemset(h->dp_sym + dp_recv_off(h->plan), 0,
                (dp_total_floats(h->plan) - dp_recv_off(h->plan)) * sizeof(float)));
  CU(cudaMemset(h->dp_error, 0, 4));
  CU(cudaDeviceSynchronize());
  return 0;
}

int mafb200_dp_error(mafb200_handle h, int32_t* out) {
  if (!h || !out) return fail("null argument");
  *out = 0;
  if (h->dp_error) {
    unsigned int e = 0;
    CU(cudaMemcpy(&e, h->dp_error, 4, cudaMemcpyDeviceToHost));
    *out = (int32_t)e;
  }
  return 0;
}

int mafb200_set_wide_core(mafb200_handle h, int32_t core) {
  if (!h) return fail("null handle");
  if (core != 0 && core != 1) return fail("core must be 0 (FP32 FFMA) or 1 (tcgen05 TF32)");
  h->wide_core = core;
  return 0;
}

int mafb200_set_flow_kernel(mafb200_handle h, int32_t mode) {
  if (!h) return fail("null handle");
  if (mode < 0 || mode > 2) return fail("mode must be 0 (auto), 1 (barrier-phased) or 2 (chain)");
  h->flow_kernel = mode;
  return 0;
}

int64_t mafb200_param_count(mafb200_handle h) { return h ? h->plan.P : -1; }

int mafb200_get_mask(mafb200_handle h, uint8_t* mask_host) {
  if (!h || !mask_host) return fail("null argument");
  std::memcpy(mask_host, h->mask_host.data(), h->mask_host.size());
  return 0;
}

int mafb200_query(mafb200_handle h, int32_t* smem_bytes, int32_t* tile_rows, int32_t* max_ctas,
                  int32_t* stash_all) {
  if (!h) return fail("null handle");
  if (smem_bytes) *smem_bytes = h->wide ? 0 : h->plan.smem_bytes;
  if (tile_rows) *tile_rows = h->plan.R;
  if (max_ctas) *max_ctas = h->n_sm;
  if (stash_all) *stash_all = h->plan.stash_all;
  return 0;
}

int mafb200_pack(mafb200_handle h, const float* params, void* stream) {
  if (!h || !params) return fail("null argument");
  const Plan& P = h->plan;
  if (h->wide)
    wide_pack_kernel<<<(P.P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(params, h->mask, h->wm, P.P);
  else
    pack_kernel<<<(P.P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, params, h->mask, h->wimg);
  CU(cudaGetLastError());
  return 0;
}

// Picks the plan (CTAs per SM) and the tile geometry for a batch of B rows.  Rows are spread as
// evenly as possible: every CTA gets ceil(rows/CTAs) rows cut into equal tiles of Rt <= R rows.
static int launch_geometry(mafb200_handle h, int64_t B, const Plan*& plan, int& Rt, int& n_tiles, int& grid) {
  int occ = 1;
  if (h->plan2.P) {
    // measured on B200 (tests/occ_bench.py): one 512-thread CTA per SM beats two 256-thread CTAs on
    // c1-c3 at every batch size tried; the second plan stays selectable for experiments
    if (h->occ_mode == 2) occ = 2;
  }
  plan = occ == 2 ? &h->plan2 : &h->plan;
  const int R = plan->R;
  const int64_t slots = (int64_t)h->n_sm * occ;
  const int64_t rows_per_cta = (B + slots - 1) / slots;
  const int64_t n_t = (rows_per_cta + R - 1) / R;
  const int64_t rt = (rows_per_cta + n_t - 1) / n_t;
  Rt = (int)std::min<int64_t>(R, (rt + 3) & ~3LL);
  const int64_t T = (B + Rt - 1) / Rt;
  if (T > 0x7fffffff) return fail("batch too large");
  n_tiles = (int)T;
  grid = (int)std::min<int64_t>(slots, T);
  return 0;
}

// fold the previous launch's event pair into the running sum, then open a new bracket
static int flow_timing_begin(mafb200_handle h, cudaStream_t st) {
  if (h->ev_pending) {
    CU(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->flow_ms_sum += ms;
    h->flow_launches += 1;
  }
  CU(cudaEventRecord(h->ev0, st));
  h->ev_pending = true;
  return 0;
}

// work-list geometry of one launch (see FlowArgs::Phase)
static void fill_phases(const Plan& P, FlowArgs& A, bool want_dy = false, bool skip_wgrad = false) {
  const int nRB = A.Rt / 4;
  int s = 0;
  while ((1 << s) < nRB) ++s;          // nRB <= 8 -> s <= 3
  A.rb_shift = s;
  const int cpw = 16 >> s;             // column blocks per warp in the outer-form products
  for (int i = 0; i < P.n_lin; ++i) {
    const LinDesc& ln = P.lin[i];
    FlowArgs::Phase& ph = A.ph[i];
    ph.fB = (ln.NP / 4 + cpw - 1) / cpw;
    ph.xC = (i == 0 ? (want_dy ? ln.KP : P.DP) : ln.KP) / 4;   // dy mode: also the conditioner columns
    ph.xB = (ph.xC + cpw - 1) / cpw;
    ph.nKB = (ln.K + 1 + 3) / 4;
    ph.nNB = (ln.N + 3) / 4;
    int nbs = 0;
    while ((1 << nbs) < ph.nNB && nbs < 3) ++nbs;   // at most 8 column blocks per half-warp
    ph.nb_shift = nbs;
    ph.nbg = (ph.nNB + (1 << nbs) - 1) >> nbs;
    const int kpw = 16 >> nbs;
    ph.nKG = (ph.nKB + kpw - 1) / kpw;
    ph.bB = skip_wgrad ? 0 : ph.nKG * ph.nbg;
  }
  // chain kernel: reduction steps (4 parts) and the column split between the two warps of a row pair
  for (int i = 0; i < P.n_lin; ++i) {
    const LinDesc& ln = P.lin[i];
    A.cf[i].nsteps = ln.KP / 4;
    A.cf[i].nCB = ln.NP / 4;
    A.cf[i].split = A.cf[i].nCB > 8;
    A.cx[i].nsteps = ln.NP / 4;
    A.cx[i].nCB = (i == 0 ? P.DP : ln.KP) / 4;
    A.cx[i].split = A.cx[i].nCB > 8;
  }
}

int mafb200_forward(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                    float* out_logp, float* out_u, float* out_logdet, int32_t flags, void* stream) {
  if (!h || !params || !x || (h->desc.n_cond && !y)) return fail("null argument");
  if (!out_logp && !out_u && !out_logdet) return fail("no output requested");
  if (B <= 0) return B == 0 ? 0 : fail("negative batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->wide) return wide_forward(h, params, x, y, B, out_logp, out_u, out_logdet, flags, st);
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, stream)) return 1;
  FlowArgs A{};
  A.wimg = h->wimg;
  A.x = x;
  A.y = y;
  A.B = B;
  A.batch_index = nullptr;
  A.n_batches = 1;
  A.stdc = h->stdc;
  A.logdet_const = h->logdet_const;
  A.inv_b = 1.f;
  A.trace = h->trace;
  A.out_logp = out_logp;
  A.out_u = out_u;
  A.out_logdet = out_logdet;
  int grid;
  const Plan* pp;
  if (launch_geometry(h, B, pp, A.Rt, A.n_tiles, grid)) return 1;
  fill_phases(*pp, A);
  const Plan& P = *pp;
  if (h->timing && flow_timing_begin(h, st)) return 1;
  launch_flow<false>(h->flow_kernel, P, A, grid, st);
  CU(cudaGetLastError());
  if (h->timing) CU(cudaEventRecord(h->ev1, st));
  return 0;
}

int mafb200_log_prob(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                     float* out_logp, int32_t flags, void* stream) {
  if (!out_logp) return fail("null argument");
  return mafb200_forward(h, params, x, y, B, out_logp, nullptr, nullptr, flags, stream);
}

// eval_step + the accumulation of eval_model (jaxili/train.py:337-340, 453-487) without leaving the device:
// acc[0] += -sum_b log_prob(x_b | y_b), acc[1] += B.  The size-weighted mean over a loader is acc[0] / acc[1],
// read once at the end of the loop.  The log_prob scratch grows on demand (not during graph capture).
int mafb200_eval_accumulate(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                            float* acc, int32_t flags, void* stream) {
  if (!h || !acc) return fail("null argument");
  if (B <= 0) return B == 0 ? 0 : fail("negative batch");
  CU(cudaSetDevice(h->device));
  if (B > h->eval_cap) {
    cudaFree(h->eval_lp);
    h->eval_lp = nullptr;
    h->eval_cap = 0;
    const long long cap = std::max<long long>(B, 4096);
    CU(cudaMalloc(&h->eval_lp, (size_t)cap * 4));
    h->eval_cap = cap;
  }
  if (mafb200_forward(h, params, x, y, B, h->eval_lp, nullptr, nullptr, flags, stream)) return 1;
  eval_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(h->eval_lp, B, acc);
  CU(cudaGetLastError());
  return 0;
}

int mafb200_loss_grad(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                      float inv_global_B, const int32_t* batch_index, int64_t n_batches, float* out_loss,
                      float* out_grad, int32_t flags, void* stream) {
  if (!h || !params || !x || (h->desc.n_cond && !y)) return fail("null argument");
  if (!(flags & MAFB200_DP_SLOT) && (!out_loss || !out_grad)) return fail("null argument");
  if (B <= 0) return fail("loss_grad needs a non-empty batch");
  if (batch_index && n_batches <= 0) return fail("n_batches must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  const bool dp_slot = (flags & MAFB200_DP_SLOT) != 0;
  if (dp_slot) {
    if (!h->dp_sym || !h->dp_step) return fail("MAFB200_DP_SLOT needs mafb200_dp_init and mafb200_dp_bind_step");
    if (h->wide) return fail("the peer-memory all-reduce serves the fused narrow path; use NCCL for wide models");
  }
  if (h->wide) {
    if (batch_index) return fail("device-side batch selection is not available on the wide path");
    if (wide_loss_grad(h, params, x, y, B, inv_global_B, out_loss, out_grad, flags, st)) return 1;
    sqnorm_kernel<<<h->n_norm, RED_IDX, 0, st>>>(out_grad, h->plan.P, h->normpart);
    CU(cudaGetLastError());
    return 0;
  }
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, stream)) return 1;
  FlowArgs A{};
  A.wimg = h->wimg;
  A.x = x;
  A.y = y;
  A.B = B;
  A.batch_index = batch_index;
  A.n_batches = batch_index ? n_batches : 1;
  A.stdc = h->stdc;
  A.logdet_const = h->logdet_const;
  A.inv_b = inv_global_B;
  A.trace = h->trace;
  A.partial_grad = h->partial;
  A.partial_loss = h->partial_loss;
  int grid;
  const Plan* pp;
  if (launch_geometry(h, B, pp, A.Rt, A.n_tiles, grid)) return 1;
  fill_phases(*pp, A);
  const Plan& P = *pp;
  if (h->timing && flow_timing_begin(h, st)) return 1;
  launch_flow<true>(h->flow_kernel, P, A, grid, st);
  CU(cudaGetLastError());
  if (h->timing) CU(cudaEventRecord(h->ev1, st));
  if (flags & MAFB200_FLOW_ONLY) return 0;
  const size_t sf = dp_slot_floats(P);
  reduce_kernel<<<h->n_norm, RED_THREADS, 0, st>>>(h->partial, P.pgs, grid, P.P, h->mask, out_grad,
                                                         h->partial_loss, out_loss, h->normpart,
                                                         dp_slot ? h->dp_step : nullptr, h->dp_sym,
                                                         h->dp_sym ? h->dp_sym + sf : nullptr);
  CU(cudaGetLastError());
  return 0;
}

// out_logp[b] = log p(x_b | y_b) and out_dy[b][c] = d log p(x_b | y_b) / d y_b[c]: the gradient with
// respect to the CONDITIONING input, i.e. what HMC/NUTS over an NLE likelihood differentiates
// (jaxili/posterior/mcmc_posterior.py:139-159, 194-201).  Same fused kernel as loss_grad run with
// 1/B := -1 (so the "loss" is sum_b log p_b, whose y-gradient is per row), with the weight-gradient
// work list switched off and the input-gradient product extended to the conditioner columns.
int mafb200_log_prob_grad_y(mafb200_handle h, const float* params, const float* x, const float* y, int64_t B,
                            float* out_logp, float* out_dy, int32_t flags, void* stream) {
  if (!h || !params || !x || !y || !out_dy) return fail("null argument");
  if (h->desc.n_cond == 0) return fail("the model has no conditioning input");
  if (B <= 0) return B == 0 ? 0 : fail("negative batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->wide) return wide_log_prob_grad_y(h, params, x, y, B, out_logp, out_dy, flags, st);
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, stream)) return 1;
  FlowArgs A{};
  A.wimg = h->wimg;
  A.x = x;
  A.y = y;
  A.B = B;
  A.n_batches = 1;
  A.stdc = h->stdc;
  A.logdet_const = h->logdet_const;
  A.inv_b = -1.f;
  A.out_logp = out_logp;
  A.out_dy = out_dy;
  A.partial_grad = h->partial;        // unused: no weight-gradient tasks are scheduled
  A.partial_loss = h->partial_loss;
  A.trace = nullptr;
  int grid;
  const Plan* pp;
  if (launch_geometry(h, B, pp, A.Rt, A.n_tiles, grid)) return 1;
  fill_phases(*pp, A, /*want_dy=*/true, /*skip_wgrad=*/true);
  launch_flow<true>(h->flow_kernel, *pp, A, grid, st);
  CU(cudaGetLastError());
  return 0;
}

int mafb200_hmc_trajectory(mafb200_handle h, const float* params, const float* x, int64_t n,
                           const MafHmcState* s, int32_t n_steps, int32_t flags, void* stream) {
  if (!h || !params || !x || !s) return fail("null argument");
  if (!s->eps || !s->inv_mass || !s->prior_kind || !s->lo || !s->hi || !s->z || !s->p || !s->g || !s->logdens ||
      !s->theta || !s->scratch)
    return fail("null pointer in MafHmcState");
  if (n <= 0 || n_steps < 0) return fail("hmc_trajectory needs n > 0 and n_steps >= 0");
  if (h->desc.n_cond == 0) return fail("the model has no conditioning input");
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, stream)) return 1;
  const int C = h->desc.n_cond;
  HmcArgs A{};
  A.eps = s->eps;
  A.inv_mass = s->inv_mass;
  A.kind = s->prior_kind;
  A.lo = s->lo;
  A.hi = s->hi;
  A.z = s->z;
  A.p = s->p;
  A.g = s->g;
  A.theta = s->theta;
  A.dll = s->scratch;
  A.ll = s->scratch + n * C;
  A.logdens = s->logdens;
  A.n = n;
  A.C = C;
  const int bd = (int)std::min<int64_t>((n * C + 255) / 256, 4 * (int64_t)h->n_sm);
  const int bk = (int)std::min<int64_t>((n + 7) / 8, 4 * (int64_t)h->n_sm);
  const int steps = n_steps > 0 ? n_steps : 1;
  for (int i = 0; i < steps; ++i) {
    A.kick = n_steps > 0 ? 0.5f : 0.f;
    A.drift = n_steps > 0 ? 1.f : 0.f;
    hmc_drift_kernel<<<bd, 256, 0, st>>>(A);
    CU(cudaGetLastError());
    if (mafb200_log_prob_grad_y(h, params, x, s->theta, n, s->scratch + n * C, s->scratch,
                                MAFB200_PACKED_CURRENT, stream))
      return 1;
    hmc_kick_kernel<<<bk, 256, 0, st>>>(A);
    CU(cudaGetLastError());
  }
  return 0;
}

// TrainerModule.train_step as one call (jaxili/train.py:330-335: value_and_grad + apply_gradients) for a
// single rank.  Narrow models: the flow kernel followed by ONE cooperative tail kernel (reduction of the
// partial gradients, global norm, clip, schedule, Adam, re-pack).  Wide models and anything the tail cannot
// serve fall through to loss_grad + clip_adam, same results.
int mafb200_train_step(mafb200_handle h, float* params, const float* x, const float* y, int64_t B,
                       float inv_global_B, const int32_t* batch_index, int64_t n_batches, float* out_loss,
                       float* out_grad, float* m, float* v, int32_t* step, const MafOptDesc* opt, int32_t flags,
                       void* stream) {
  if (!h || !params || !x || (h->desc.n_cond && !y) || !out_loss || !out_grad || !step || !opt)
    return fail("null argument");
  const bool dp = (flags & MAFB200_DP_SLOT) != 0;
  if (dp) {
    if (!h->dp_sym || !h->dp_step) return fail("MAFB200_DP_SLOT needs mafb200_dp_init and mafb200_dp_bind_step");
    if (h->dp_step != step) return fail("MAFB200_DP_SLOT: step must be the counter bound with mafb200_dp_bind_step");
    if (h->wide) return fail("the peer-memory all-reduce serves the fused narrow path; use NCCL for wide models");
    for (int r = 0; r < h->dp_world; ++r)
      if (!h->dp_peer_sym[r]) return fail("dp peers not connected");
  }
  if (opt->kind < 0 || opt->kind > 2) return fail("unknown optimizer kind");
  if (opt->kind != 1 && (!m || !v)) return fail("adam needs m and v");
  if (opt->decay_steps - opt->warmup <= 0.f) return fail("cosine_decay_schedule requires positive decay_steps");
  if (!h->wide && h->tail_ok < 0) {
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_tail_kernel<true>, RED_THREADS, 0));
    int coop = 0;
    CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
    h->tail_ok = (coop && (long long)per_sm * h->n_sm >= h->n_norm) ? 1 : 0;
  }
  if (dp && h->tail_ok != 1) {
    if (mafb200_loss_grad(h, params, x, y, B, inv_global_B, batch_index, n_batches, nullptr, nullptr, flags, stream))
      return 1;
    if (mafb200_dp_allreduce(h, out_grad, out_loss, stream)) return 1;
    return mafb200_clip_adam(h, params, out_grad, m, v, step, opt, MAFB200_NORM_CURRENT, stream);
  }
  if (h->wide || h->tail_ok != 1) {
    if (mafb200_loss_grad(h, params, x, y, B, inv_global_B, batch_index, n_batches, out_loss, out_grad, flags, stream))
      return 1;
    return mafb200_clip_adam(h, params, out_grad, m, v, step, opt, MAFB200_NORM_CURRENT, stream);
  }
  if (B <= 0) return fail("train_step needs a non-empty batch");
  if (batch_index && n_batches <= 0) return fail("n_batches must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (!(flags & MAFB200_PACKED_CURRENT) && mafb200_pack(h, params, stream)) return 1;
  FlowArgs A{};
  A.wimg = h->wimg;
  A.x = x;
  A.y = y;
  A.B = B;
  A.batch_index = batch_index;
  A.n_batches = batch_index ? n_batches : 1;
  A.stdc = h->stdc;
  A.logdet_const = h->logdet_const;
  A.inv_b = inv_global_B;
  A.trace = h->trace;
  A.partial_grad = h->partial;
  A.partial_loss = h->partial_loss;
  int grid;
  const Plan* pp;
  if (launch_geometry(h, B, pp, A.Rt, A.n_tiles, grid)) return 1;
  fill_phases(*pp, A);
  if (h->timing && flow_timing_begin(h, st)) return 1;
  launch_flow<true>(h->flow_kernel, *pp, A, grid, st);
  CU(cudaGetLastError());
  if (h->timing) CU(cudaEventRecord(h->ev1, st));
  OptArgs O{opt->lr, opt->warmup, opt->decay_steps, opt->init_factor, opt->end_factor, opt->b1, opt->b2,
            opt->eps, opt->clip, opt->weight_decay, opt->kind};
  Plan Pk = h->plan;           // image layout and parameter count (shared by both occupancy plans)
  const float* partial = h->partial;
  const unsigned char* mask = h->mask;
  const float* ploss = h->partial_loss;
  float* normpart = h->normpart;
  float* wimg = h->wimg;
  unsigned long long* bar = h->grid_bar;
  DpArgs a{};
  if (dp) {
    a.rank = h->dp_rank;
    a.world = h->dp_world;
    a.P = Pk.P;
    a.step = h->dp_step;
    const size_t sf = dp_slot_floats(Pk);
    a.slot_floats = (long long)sf;
    for (int r = 0; r < h->dp_world; ++r) {
      a.sym[r] = h->dp_peer_sym[r];
      a.flags_of[r] = reinterpret_cast<unsigned int*>(h->dp_peer_sym[r] + 2 * sf);
    }
    a.error = h->dp_error;
    a.timeout_cycles = kDpTimeoutCycles;
    a.recv_off = (long long)dp_recv_off(Pk);
  }
  a.trace = h->trace ? h->trace + 900 : nullptr;
  void* args[] = {&Pk, &O, &partial, &grid, &mask, &out_grad, &ploss, &out_loss, &normpart, &params,
                  &m, &v, &step, &wimg, &bar, &a};
  const void* fn = dp ? (const void*)step_tail_kernel<true> : (const void*)step_tail_kernel<false>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(h->n_norm);
  cfg.blockDim = dim3(RED_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  CU(cudaLaunchKernelExC(&cfg, fn, args));
  return 0;
}

int mafb200_clip_adam(mafb200_handle h, float* params, const float* grad, float* m, float* v, int32_t* step,
                      const MafOptDesc* opt, int32_t flags, void* stream) {
  if (!h || !params || !grad || !step || !opt) return fail("null argument");
  if (opt->kind < 0 || opt->kind > 2) return fail("unknown optimizer kind");
  if (opt->kind != 1 && (!m || !v)) return fail("adam needs m and v");
  if (opt->decay_steps - opt->warmup <= 0.f) return fail("cosine_decay_schedule requires positive decay_steps");
  cudaStream_t st = (cudaStream_t)stream;
  const Plan& P = h->plan;
  if (!(flags & MAFB200_NORM_CURRENT)) {
    sqnorm_kernel<<<h->n_norm, RED_IDX, 0, st>>>(grad, P.P, h->normpart);
    CU(cudaGetLastError());
  }
  OptArgs O{opt->lr, opt->warmup, opt->decay_steps, opt->init_factor, opt->end_factor, opt->b1, opt->b2,
            opt->eps, opt->clip, opt->weight_decay, opt->kind};
  const float* npart = h->normpart;
  int n_part = h->n_norm;
  if (n_part > 2048) {   // wide models: tens of thousands of partials -- sum them once
    sum_partials_kernel<<<1, 1024, 0, st>>>(h->normpart, h->n_norm, h->normsum);
    CU(cudaGetLastError());
    npart = h->normsum;
    n_part = 1;
  }
  clip_adam_kernel<<<h->n_norm, RED_IDX, 0, st>>>(P, O, params, grad, m, v, step, npart, n_part,
                                                  h->mask, h->wide ? h->wm : h->wimg, h->wide ? 1 : 0,
                                                  h->done_counter, nullptr);
  CU(cudaGetLastError());
  return 0;
}

#include "pipeline.inc"

static int sample_common(mafb200_handle h, const float* params, const float* u, uint64_t seed,
                         uint64_t row_offset, const float* y_obs, int64_t n_obs, int64_t rows_per_obs,
                         int64_t N, float* out, float* out_logdet, int32_t flags, void* stream) {
  if (!h || !params || !out || (h->desc.n_cond && !y_obs)) return fail("null argument");
  if (N <= 0) return N == 0 ? 0 : fail("negative sample count");
  if (n_obs <= 0 || rows_per_obs <= 0 || n_obs * rows_per_obs < N)
    return fail("need n_obs * rows_per_obs >= N");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->wide)   // models that do not fit shared memory: D sequential passes through the GEMM chain per layer
    return wide_sample(h, params, u, seed, row_offset, y_obs, n_obs, rows_per_obs, N, out, out_logdet, flags, st);
  if (h->splan.img_floats == 0) return fail("sampler not available for this configuration");
  const Plan& P = h->plan;
  const SamplePlan& S = h->splan;
  if (!(flags & MAFB200_PACKED_CURRENT)) {
    pack_sample_kernel<<<(P.P + 255) / 256, 256, 0, st>>>(P, S, h->spack, params, h->mask, h->spos, h->simg);
    CU(cudaGetLastError());
  }
  SampleArgs A{};
  A.simg = h->simg;
  A.u = u;
  A.seed = seed;
  A.row_offset = row_offset;
  A.y_obs = y_obs;
  A.n_obs = n_obs;
  A.rows_per_obs = rows_per_obs;
  A.N = N;
  A.stdc = h->stdc;
  A.out = out;
  A.out_logdet = out_logdet;
  const int64_t tiles = (N + 31) / 32;
  const int grid = (int)std::min<int64_t>(h->n_sm, (tiles + S.warps - 1) / S.warps);
  if (S.act == MAFB200_ACT_RELU)
    maf_sample_kernel<true><<<grid, S.warps * 32, S.smem_bytes, st>>>(S, A);
  else
    maf_sample_kernel<false><<<grid, S.warps * 32, S.smem_bytes, st>>>(S, A);
  CU(cudaGetLastError());
  return 0;
}

int mafb200_sample_from_noise(mafb200_handle h, const float* params, const float* u, const float* y_obs,
                              int64_t n_obs, int64_t rows_per_obs, int64_t N, float* out, int32_t flags,
                              void* stream) {
  if (!u) return fail("null noise pointer");
  return sample_common(h, params, u, 0, 0, y_obs, n_obs, rows_per_obs, N, out, nullptr, flags, stream);
}

int mafb200_inverse(mafb200_handle h, const float* params, const float* u, const float* y, int64_t N,
                    float* out_x, float* out_logdet, int32_t flags, void* stream) {
  if (!u) return fail("null noise pointer");
  return sample_common(h, params, u, 0, 0, y, N > 0 ? N : 1, 1, N, out_x, out_logdet, flags, stream);
}

int mafb200_sample_philox(mafb200_handle h, const float* params, uint64_t seed, uint64_t row_offset,
                          const float* y_obs, int64_t n_obs, int64_t rows_per_obs, int64_t N, float* out,
                          int32_t flags, void* stream) {
  return sample_common(h, params, nullptr, seed, row_offset, y_obs, n_obs, rows_per_obs, N, out, nullptr, flags,
                       stream);
}

// FP32 FMA-pipe peak measured live: the 4x4 register-tile outer product of the flow kernels' inner loops with
// all-register operands, as plain FFMA (packed = 0) or packed FFMA2 (packed = 1, what the kernels issue)
int mafb200_fp32_peak(int device, int32_t iters, int32_t packed, float* out_tflops) {
  if (!out_tflops || iters <= 0) return fail("bad argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  float* buf = nullptr;
  CU(cudaMalloc(&buf, (size_t)blocks * threads * 4));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  auto launch = [&](int n) {
    if (packed) ffma2_peak_kernel<<<blocks, threads>>>(buf, n, 1e-3f);
    else ffma_peak_kernel<<<blocks, threads>>>(buf, n, 1e-3f);
  };
  launch(iters / 4 + 1);   // warm-up
  float best = 0.f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0));
    launch(iters);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * threads * (double)iters * 8 * 16 * 2;
    best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *out_tflops = best;
  return 0;
}

int mafb200_ffma_peak(int device, int32_t iters, float* out_tflops) {
  return mafb200_fp32_peak(device, iters, 0, out_tflops);
}

// TF32 tcgen05 tensor-pipe peak measured live (tf32_peak_kernel): burst figure of a kernel timed alone
int mafb200_tf32_peak(int device, int32_t iters, float* out_tflops) {
  if (!out_tflops || iters <= 0) return fail("bad argument");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail("this library is built for sm_100a (B200) only");
  const int blocks = prop.multiProcessorCount;
  const int smem = (128 + 256) * 32 * 4 + 1024;
  CU(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  float* buf = nullptr;
  CU(cudaMalloc(&buf, blocks * 4));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  tf32_peak_kernel<<<blocks, 128, smem>>>(iters / 8 + 1, buf);   // warm-up
  float best = 0.f;
  for (int rep = 0; rep < 5; ++rep) {
    CU(cudaEventRecord(e0));
    tf32_peak_kernel<<<blocks, 128, smem>>>(iters, buf);
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = (double)blocks * (double)iters * 4 * 2.0 * 128 * 256 * 8;
    best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  CU(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *out_tflops = best;
  return 0;
}

}  // extern "C"
