// plan.h -- host/device shared description of one ConditionalMAF and of the shared-memory
// layout the fused flow kernels use.  Built once per handle by the planner in mafb200.cu.
#pragma once
#include <stdint.h>

#define MAFB_MAX_LIN 6      // masked linears per MADE = n_hidden + 1
#define MAFB_MAX_DEG 64     // n_in limit of the staircase sampler tables
#define MAFB_SMEM_HDR 128   // bytes in front of the float layout: mbarriers, loss accumulators

// One masked linear of the MADE (jaxili/model.py:517-544).  Feature counts are padded to
// multiples of 4 so every shared-memory row can be read with 128-bit loads.
struct LinDesc {
  int K, N;      // in, out
  int KP, NP;    // padded
  int w_off;     // forward image: W[KP][NP] (mask folded, rows K..KP-1 zero), floats from image start
  int b_off;     // forward image: bias[NP]
  int wt_off;    // backward image: Wt[NP][KP] (rows N..NP-1 zero), floats from backward-image start
  int pw_off;    // kernel offset inside one layer of the flat parameter vector
  int pb_off;    // bias offset inside one layer of the flat parameter vector
};

struct Plan {
  // model
  int D, C, K0, L, n_lin, act, reverse;
  int DP;                       // round4(D)
  LinDesc lin[MAFB_MAX_LIN];
  int fwd_floats, bwd_floats;   // per-layer image parts (multiples of 4 floats)
  int img_floats;               // fwd_floats + bwd_floats
  int ppl, P;                   // params per layer, total
  int pgs;                      // row stride of the per-CTA partial gradients (P rounded up to 4: 128-bit reads)
  // tiling
  int R, RP;                    // rows per tile, row pitch of feature-major buffers
  int threads;
  int n_wbuf;                   // weight-image buffers (1 or 2)
  int stash_all;                // 1: per-layer activation slots kept for backward; 0: recompute
  int need_d;                   // activation derivative stored separately (non-ReLU)
  // shared-memory layout, offsets in floats from the dynamic smem base
  int o_wbuf, wbuf_stride;
  int o_stage, stage_stride;    // raw [R][D]+[R][C] input tiles (2 buffers)
  int o_xin, xin_stride;        // per-layer MADE input [K0P][RP]: x rows then y rows
  int o_slot, slot_stride;      // per-layer activation slot
  int h_off[MAFB_MAX_LIN];      // slot-relative offset of hidden level i (1..n_hidden)
  int d_off[MAFB_MAX_LIN];      // slot-relative offset of act'(a_i) (need_d only)
  int s_off, v_off;             // slot-relative exp(logp/2) and v=(x-mu)s, [DP][RP] each
  int o_obuf;                   // [round4(2D)][RP]: MADE output, then grad wrt it
  int o_gu, o_gx, o_ld;         // [DP][RP] each (o_gx: two of them, ping-pong by layer parity)
  int o_cst;                    // standardisation constants [2D + 2C]
  int o_red;                    // [32] per-warp loss contributions of the finalize step
  int o_gy;                     // [round4(C)][RP]: gradient w.r.t. the standardised conditioner (dy mode)
  int o_g[2];                   // ping-pong [max HP][RP] hidden-gradient buffers (barrier-phased kernel)
  // chain kernel: two gradient-buffer sets (alternating per layer step); set b starts at o_gb + b * gb_stride,
  // the gradient w.r.t. the output of linear i sits at g_off[i] inside a set (g_off[n_lin-1] == 0).
  // Set 0 is the region [o_obuf, o_g[0], o_g[1]] of the barrier-phased kernel.
  int o_gb, gb_stride;
  int g_off[MAFB_MAX_LIN];
  int smem_bytes;
};

// staircase sampler: per-layer image layout (floats from image start)
struct SamplePlan {
  int D, C, L, n_lin, act, reverse;
  int width[MAFB_MAX_LIN];      // padded sorted width of level i (level 0 = K0, last = 2*D)
  int w_off[MAFB_MAX_LIN];      // W_i[rows = width[i]][cols = width[i+1]]
  int b_off[MAFB_MAX_LIN];
  int tab_off;                  // int tables: for level i=1..n_hidden, degree m: start,end (padded cols)
  int img_floats;
  int rows_per_warp, RP;        // 32, 36
  int warps;                    // warps per CTA
  int o_img, o_warp, warp_stride;     // smem offsets (floats)
  int lvl_off[MAFB_MAX_LIN];    // per-warp offset of level i activations (i = 1..n_hidden)
  int y_off, u_off, x_off;      // per-warp y'[CP][RP], u[DP][RP], x[DP][RP]
  int n_img;                    // image buffers in smem (1 or 2)
  int smem_bytes;
};
