// Fused data-path adjoint of one SVGP layer in V-form (the adjoint of fused_forward_kernel; replaces three GEMM launches and the
// RBF adjoint kernel of the unfused pipeline). Reference operation being differentiated: utils/layers.py:243-278 via
// tape.gradient (models/dgp.py:272-275); formulas: SURVEY.md §9 / DESIGN.md §4.
//
// One persistent CTA per SM owns tiles of PT point-samples. Per tile, with V = Lu^-1 Kuf and T_d = C_d V from the training stash,
// the upstream gradients Gm [P][D], Gv^T [D][Pp], gq = -sum_d Gv_d:
//   0. dV = sum_d C_d^T (2 Gv_d o T_d) + beta Gm^T + 2 V diag(gq)             accumulators in registers over all d; C_d^T (lower)
//      panels AND the T_d k-slabs [16 x PT] stream through one mbarrier ring (bulk copies, SASS UBLKCP); the column scale
//      2 Gv_d is applied to the B fragments as they are loaded; the rank-D_out and the diagonal terms are added at the block end.
//      dV goes to the resident shared-memory tile and to HBM (the parameter contraction tril(dV V^T) reads it).
//   1. K-bar = Lu^-T dV     in place on the resident tile (upper operator, row blocks ascending)
//   2. kernel adjoint on the resident tile: Gbar = K-bar o (-2 dk/dr2) with Kuf rebuilt from the inputs (FP64 exp), written to HBM
//      for the H = Gbar [X, 1] contraction; dX[p] = sum_m Gbar (z - x) / l^2 (+ mean-function path); partial sums for dl, ds2;
//      XaugPad rows.
// dV and K-bar make no HBM round trip between the steps; the stash is read once (1 + 1/(2 nb) times for nb > 1 row blocks).
#pragma once
#include "common.cuh"
#include "fused.cuh"

namespace dgp {

struct FusedBwdArgs {
  const double* stream;                 // packed operator panels in consumption order (pack_bwd_stream_kernel)
  const double* V; const double* T;     // stash: [Mp][Pp], [D_out][Mp][Pp]
  const double* GvT; const double* gq;  // [D_out][Pp], [Pp]
  const double* Gm; int gm_ld;          // [Pp][gm_ld] upstream mean gradient (rows >= P zero)
  const double* beta;                   // [Mp][32]  Lu^-1 q_mu (whitened: q_mu)
  const double* Zs; const double* ls; const double* var;   // [M][D_in] scaled inducing inputs, [D_in], [1]
  const double* Xin; long xmod; int D_in;
  const double* mfW; int mean_kind; int kind;
  int M, Mp, D_out; long P, Pp;
  double* dV; double* Gbar;             // [Mp][Pp] out
  double* dXin;                         // [P][D_in] out or null
  double* XaugPad;                      // [Pp][32] out: [x, 1, 0...]
  double* part;                         // [tiles * WN][D_in + 1] out: partial sums for dl_j, ds2
  int warp_major_groups, group_skew;    // see FusedFwdArgs
  const double* zz;                     // [Mp] |Zs[m]|^2 (zz_kernel)
};

// Inside a row block the k-panels are visited long, short, long, short ...: in natural order the panels of a diagonal block get
// shorter and shorter (4, 4, 3, 3, 2, 2, 1, 1 active m-tiles per warp) and the last ones hold less DMMA work than one ring refill
// takes, so the consumers ran dry there (ring waits 11% -> see profiles/). The order of a k-sum only changes rounding.
__host__ __device__ __forceinline__ int zigzag(int idx, int n) { return (idx & 1) ? n - 1 - (idx >> 1) : (idx >> 1); }

// Panel order: pass 0: row block i = 0..nb-1, output d = 0..D-1, k-panels 0..(i+1)*BM/16-1 of C_d^T; pass 1: row block i,
// k-panels i*BM/16..Mp/16-1 of Lu^-T. One CTA per panel.
template <int BM>
__global__ void __launch_bounds__(256) pack_bwd_stream_kernel(const double* __restrict__ Cmat, const double* __restrict__ LinvT,
                                                              int Mp, int D, double* __restrict__ stream) {
  const int nb = Mp / BM, kpb = BM / kPanelK;
  int q = blockIdx.x, pass = 0, i = 0, d = 0, ks = 0;
  const int NP0 = D * kpb * nb * (nb + 1) / 2;
  if (q < NP0) {
    for (i = 0;; ++i) { const int cnt = D * (i + 1) * kpb; if (q < cnt) break; q -= cnt; }
    d = q / ((i + 1) * kpb); ks = zigzag(q % ((i + 1) * kpb), (i + 1) * kpb);
  } else {
    q -= NP0; pass = 1;
    for (i = 0;; ++i) { const int cnt = (nb - i) * kpb; if (q < cnt) break; q -= cnt; }
    ks = i * kpb + zigzag(q, (nb - i) * kpb);
  }
  double* dst = stream + (long)blockIdx.x * BM * kPanelK;
  for (int idx = threadIdx.x; idx < BM * kPanelK; idx += blockDim.x) {
    const int r = idx % BM, k = idx / BM;            // consecutive threads walk the rows: coalesced reads of C_d (row k of C_d = column k of C_d^T)
    const int row = i * BM + r, col = ks * kPanelK + k;
    dst[panel_swz(r, k)] = pass == 0 ? Cmat[((long)d * Mp + col) * Mp + row] : LinvT[(long)row * Mp + col];
  }
}

// panel_block of fused.cuh with the B fragments multiplied by a per-column scale as they are loaded (pass 0: 2 Gv_d)
template <int LO, int HI, int TM, int TN, int WM, int LDT, bool SC>
__device__ __forceinline__ void bwd_panel_block(double (&c0)[TM][TN], double (&c1)[TM][TN], const double* __restrict__ pan,
                                                const double* __restrict__ bt, const double (&sc)[TN], int wm, int g8, int t4) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double bv[TN], av[HI - LO > 0 ? HI - LO : 1];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = bt[kk * 4 * LDT + j * 8];
#pragma unroll
    for (int i = LO; i < HI; ++i) {
      const int row = i * 8 * WM + wm * 8 + g8;
      av[i - LO] = pan[row * kPanelK + (((kk ^ (row & 3)) << 2) | t4)];
    }
    if (SC) {
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] *= sc[j];
    }
#pragma unroll
    for (int i = LO; i < HI; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) dmma884(c0[i][j], c1[i][j], av[i - LO], bv[j]);
  }
}
template <int TM, int TN, int WM, int LDT, bool SC>
__device__ __forceinline__ void bwd_panel_dispatch(double (&c0)[TM][TN], double (&c1)[TM][TN], const double* pan, const double* bt,
                                                   const double (&sc)[TN], int wm, int g8, int t4, int lo, int hi) {
  static_assert(TM <= 4, "dispatch covers TM <= 4");
  if (lo == 0) {
    if (hi == TM) { bwd_panel_block<0, TM, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; }
    if constexpr (TM >= 2) { if (hi == 1) { bwd_panel_block<0, 1, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
    if constexpr (TM >= 3) { if (hi == 2) { bwd_panel_block<0, 2, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
    if constexpr (TM >= 4) { if (hi == 3) { bwd_panel_block<0, 3, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
    return;
  }
  if constexpr (TM >= 2) { if (lo == TM - 1) { bwd_panel_block<TM - 1, TM, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
  if constexpr (TM >= 3) { if (lo == TM - 2) { bwd_panel_block<TM - 2, TM, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
  if constexpr (TM >= 4) { if (lo == TM - 3) { bwd_panel_block<TM - 3, TM, TM, TN, WM, LDT, SC>(c0, c1, pan, bt, sc, wm, g8, t4); return; } }
}



template <int BM, int PT, int WM, int WN>
struct FusedBwdCfg {
  static_assert(WM * WN == 8, "8 consumer warps");
  static constexpr int THREADS = 384;   // 8 consumer warps + a warp group whose first warp is the producer (see regs_shrink / regs_grow)
  static constexpr int GT = WM * 32, GC = PT / WN;
  static constexpr int TM = BM / (8 * WM), TN = GC / 8;
  static constexpr int LDT = PT + 4;
  static constexpr int PANEL = BM * kPanelK, SLAB = kPanelK * LDT, STAGE = PANEL + SLAB;
  static constexpr int STAGES = 3;
  static size_t smem_bytes(int Mp, int D_in, int D_out) {
    return ((size_t)STAGES * STAGE + (size_t)Mp * LDT + (size_t)D_in * PT + (size_t)D_out * PT + (size_t)PT + 2 * STAGES + 2 + (size_t)WN * WM * (kMaxD + 1) + (size_t)PT * (D_in + 1)) * sizeof(double);
  }
};

template <int BM, int PT, int WM, int WN, int DMAX>
__global__ void __launch_bounds__(384, 1) fused_backward_kernel(FusedBwdArgs a) {
  using Cfg = FusedBwdCfg<BM, PT, WM, WN>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, LDT = Cfg::LDT, PANEL = Cfg::PANEL, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES;
  constexpr int GT = Cfg::GT, GC = Cfg::GC, KPB = BM / kPanelK;
  constexpr int XSW = GC == 32 ? 1 : 0;   // 8-column blocks of xs row j rotated by j & 3 (conflict-free B fragments, see fused.cuh)
  extern __shared__ __align__(128) double bsmem[];
  double* pbuf = bsmem;                                   // [STAGES][PANEL | SLAB]
  double* tile = pbuf + STAGES * STAGE;                   // [Mp][LDT]   dV -> K-bar -> Gbar
  double* xs_all = tile + (size_t)a.Mp * LDT;             // [WN][D_in][GC] scaled inputs
  double* gv2_all = xs_all + a.D_in * PT;                 // [WN][D_out][GC] 2 Gv
  double* gq_all = gv2_all + a.D_out * PT;                // [WN][GC]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(gq_all + PT);
  unsigned long long* empty = full + STAGES;
  unsigned long long* vbar = empty + STAGES;                      // the tile's V rows have landed in the resident tile
  double* red_all = reinterpret_cast<double*>(vbar + 2);          // [WN][WM][kMaxD + 1] scratch of the per-tile group reductions
  double* us_all = red_all + WN * WM * (kMaxD + 1);               // [WN][GC][D_in + 1]  U = Gbar^T [Zs, 1] of the group's columns

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.Mp / BM, kt = a.Mp / kPanelK;
  const int ntiles = (int)(a.Pp / PT);
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    mbar_init(vbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= 8) {
    regs_shrink();
    if (warp > 8) return;
    // ---- producer: per stage one operator panel, and in pass 0 the 16 row segments of the T_d slab it multiplies ----
    if (lane == 0) {
      int st = 0;
      unsigned ph = 0;
      // The T_d slabs come from HBM (the stash is far larger than L2) and the ring is only STAGES deep, so a second iterator runs
      // kAhead stages ahead of the copies and pulls the slabs into L2 (cp.async.bulk.prefetch.L2): the copy itself then hits L2.
#ifndef DGP_BWD_AHEAD
#define DGP_BWD_AHEAD 8
#endif
      constexpr int kAhead = DGP_BWD_AHEAD;
      int f_tl = 0, f_i = 0, f_d = 0, f_ks = 0;
      auto prefetch_next = [&]() {
        if (f_tl >= my_tiles) return;
        const double* trow = a.T + ((long)f_d * a.Mp + (long)zigzag(f_ks, (f_i + 1) * KPB) * kPanelK) * a.Pp + (long)(blockIdx.x + f_tl * gridDim.x) * PT;
#pragma unroll 4
        for (int r = 0; r < kPanelK; ++r)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(trow + (long)r * a.Pp), "r"(PT * 8) : "memory");
        if (++f_ks == (f_i + 1) * KPB) { f_ks = 0; if (++f_d == a.D_out) { f_d = 0; if (++f_i == nb) { f_i = 0; ++f_tl; } } }
      };
      for (int q = 0; q < kAhead; ++q) prefetch_next();
      for (int tl = 0; tl < my_tiles; ++tl) {
        const long p0 = (long)(blockIdx.x + tl * gridDim.x) * PT;
        const double* src = a.stream;
        int issued = 0;
        for (int i = 0; i < nb; ++i)
          for (int d = 0; d < a.D_out; ++d)
            for (int ksi = 0; ksi < (i + 1) * KPB; ++ksi) {
              const int ks = zigzag(ksi, (i + 1) * KPB);
              mbar_wait(empty + st, ph ^ 1);
              if (issued++ == STAGES) {
                // All 8 consumer warps have released the first panel of THIS tile, so the previous tile's epilogue is over and
                // the resident tile is write-only until the block ends of pass 0: park the tile's V rows there (the 2 V diag(gq)
                // term reads them at the block end, just before dV overwrites them).
                mbar_arrive_expect_tx(vbar, (unsigned)(a.Mp * PT * 8));
                for (int r = 0; r < a.Mp; ++r) bulk_g2s(tile + r * LDT, a.V + (long)r * a.Pp + p0, PT * 8, vbar);
              }
              double* dst = pbuf + st * STAGE;
#ifdef DGP_BWD_EXPERIMENT_NO_SLABS   // measurement only (wrong results): how much of the ring wait is the stash traffic?
              mbar_arrive_expect_tx(full + st, PANEL * 8);
              bulk_g2s(dst, src, PANEL * 8, full + st);
#else
              // diagonal block of the lower operator: rows above the panel's k-range are never read (see fused_forward_kernel)
              const int r0 = max(0, ks * kPanelK - i * BM);
              mbar_arrive_expect_tx(full + st, (unsigned)(((BM - r0) * kPanelK + kPanelK * PT) * 8));
              bulk_g2s(dst + r0 * kPanelK, src + r0 * kPanelK, (unsigned)((BM - r0) * kPanelK * 8), full + st);
              const double* trow = a.T + ((long)d * a.Mp + (long)ks * kPanelK) * a.Pp + p0;
#pragma unroll 4
              for (int r = 0; r < kPanelK; ++r) bulk_g2s(dst + PANEL + r * LDT, trow + (long)r * a.Pp, PT * 8, full + st);
#endif
              prefetch_next();
              src += PANEL;
              if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        for (int i = 0; i < nb; ++i)
          for (int ksi = 0; ksi < kt - i * KPB; ++ksi) {
            const int ks = i * KPB + zigzag(ksi, kt - i * KPB);
            mbar_wait(empty + st, ph ^ 1);
            const int nr = min(BM, ks * kPanelK - i * BM + kPanelK);   // upper operator: rows below the panel's k-range are dead
            mbar_arrive_expect_tx(full + st, (unsigned)(nr * kPanelK * 8));
            bulk_g2s(pbuf + st * STAGE, src, (unsigned)(nr * kPanelK * 8), full + st);
            src += PANEL;
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
      }
    }
    return;
  }

  // ---- consumers ----
  regs_grow();
  const int g8 = lane >> 2, t4 = lane & 3;
  const int wm = a.warp_major_groups ? warp % WM : warp / WN, wn = a.warp_major_groups ? warp / WM : warp % WN;
  const int tg = wm * 32 + lane;
  const int col0 = wn * GC;
  double* xs = xs_all + wn * a.D_in * GC;
  double* gv2 = gv2_all + wn * a.D_out * GC;
  double* gqs = gq_all + wn * GC;
  double* red = red_all + wn * WM * (kMaxD + 1);
  double* us = us_all + wn * GC * (a.D_in + 1);
  const int bar_id = 1 + wn;
  const double s2 = a.var[0];
  int cst = 0;
  unsigned cph = 0;
  double c0[TM][TN], c1[TM][TN];

  if (a.group_skew > 0 && wn > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)a.group_skew * wn) {}
  }
#ifdef DGP_DEBUG_PHASECLK
  long long ph_pro = 0, ph_p0 = 0, ph_be = 0, ph_p1 = 0, ph_row = 0, ph_red = 0, ph_out = 0, ph_wait = 0;
  const long long ph_t0 = clock64();
#endif
  for (int tl = 0; tl < my_tiles; ++tl) {
    const long tile_index = (long)blockIdx.x + (long)tl * gridDim.x;
    const long p0 = tile_index * PT + col0;
    PH_MARK(ph_a);
    group_sync(bar_id, GT);   // the previous tile's epilogue is done with the group's columns
    for (int idx = tg; idx < a.D_in * GC; idx += GT) {
      const int j = idx / GC, c = idx % GC;
      const long p = p0 + c;
      xs[j * GC + (c ^ (XSW * ((j & 3) << 3)))] = (p < a.P) ? a.Xin[(p % a.xmod) * a.D_in + j] * (1.0 / a.ls[j]) : 0.0;   // layout: fused.cuh
    }
    for (int idx = tg; idx < a.D_out * GC; idx += GT) {
      const int d = idx / GC, c = idx % GC;
      gv2[idx] = 2.0 * a.GvT[(long)d * a.Pp + p0 + c];
    }
    if (tg < GC) gqs[tg] = a.gq[p0 + tg];
    group_sync(bar_id, GT);

    PH_ADD(ph_pro, ph_a);
    // ---- pass 0: dV ----
    for (int i = 0; i < nb; ++i) {
      PH_MARK(ph_b);
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int j = 0; j < TN; ++j) { c0[ti][j] = 0.0; c1[ti][j] = 0.0; }
      for (int d = 0; d < a.D_out; ++d) {
        double sc[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) sc[j] = gv2[d * GC + j * 8 + g8];
        for (int ksi = 0; ksi < (i + 1) * KPB; ++ksi) {
          const int ks = zigzag(ksi, (i + 1) * KPB);
          const int num = ks * kPanelK - i * BM - 7 - wm * 8;   // lower operator: row r needs k <= r
          const int lo = num > 0 ? (num + 8 * WM - 1) / (8 * WM) : 0;
          const double* stage = pbuf + cst * STAGE;
          PH_MARK(ph_w);
          mbar_wait(full + cst, cph);
          PH_ADD(ph_wait, ph_w);
          bwd_panel_dispatch<TM, TN, WM, LDT, true>(c0, c1, stage, stage + PANEL + t4 * LDT + col0 + g8, sc, wm, g8, t4, lo < TM ? lo : 0, lo < TM ? TM : 0);
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + cst);
          if (++cst == STAGES) { cst = 0; cph ^= 1; }
        }
      }
      PH_ADD(ph_p0, ph_b);
      PH_MARK(ph_c);
      // block end: + beta Gm^T + 2 V diag(gq); to the resident tile and to HBM
      for (int d = 0; d < a.D_out; d += 2) {   // two outputs per round: 2 (2 TN + TM) independent loads in flight (L2 hits)
        const bool two = d + 1 < a.D_out;
        double gm0[2][TN], gm1[2][TN], bq[2][TM];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const double* g = a.Gm + (p0 + j * 8 + 2 * t4) * a.gm_ld + d;
          gm0[0][j] = g[0]; gm1[0][j] = g[a.gm_ld];
          gm0[1][j] = two ? g[1] : 0.0; gm1[1][j] = two ? g[a.gm_ld + 1] : 0.0;
        }
#pragma unroll
        for (int ti = 0; ti < TM; ++ti) {
          const double* b = a.beta + (long)(i * BM + ti * 8 * WM + wm * 8 + g8) * 32 + d;
          bq[0][ti] = b[0]; bq[1][ti] = two ? b[1] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int ti = 0; ti < TM; ++ti)
#pragma unroll
            for (int j = 0; j < TN; ++j) { c0[ti][j] = fma(bq[u][ti], gm0[u][j], c0[ti][j]); c1[ti][j] = fma(bq[u][ti], gm1[u][j], c1[ti][j]); }
      }
      if (i == 0) mbar_wait(vbar, (unsigned)(tl & 1));   // the V rows parked by the producer
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) {
        const int row = i * BM + ti * 8 * WM + wm * 8 + g8;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int cl = j * 8 + 2 * t4;                      // column inside the group
          double2* tp = reinterpret_cast<double2*>(tile + row * LDT + col0 + cl);
          const double2 v = *tp;
          const double r0 = fma(2.0 * v.x, gqs[cl], c0[ti][j]), r1 = fma(2.0 * v.y, gqs[cl + 1], c1[ti][j]);
          *tp = make_double2(r0, r1);
          if (a.dV) *reinterpret_cast<double2*>(a.dV + (long)row * a.Pp + p0 + cl) = make_double2(r0, r1);
        }
      }
      PH_ADD(ph_be, ph_c);
    }
    PH_MARK(ph_d);
    group_sync(bar_id, GT);   // the group's columns of dV are complete

    // ---- pass 1: K-bar = Lu^-T dV, in place, row blocks ascending ----
    const double sc1[TN] = {};
    for (int i = 0; i < nb; ++i) {
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int j = 0; j < TN; ++j) { c0[ti][j] = 0.0; c1[ti][j] = 0.0; }
      for (int ksi = 0; ksi < kt - i * KPB; ++ksi) {
        const int ks = i * KPB + zigzag(ksi, kt - i * KPB);
        const int num = ks * kPanelK - i * BM + kPanelK - 1 - wm * 8;   // upper operator: row r needs k >= r
        const int hi = num >= 0 ? min(TM, num / (8 * WM) + 1) : 0;
        const double* stage = pbuf + cst * STAGE;
        mbar_wait(full + cst, cph);
        bwd_panel_dispatch<TM, TN, WM, LDT, false>(c0, c1, stage, tile + (ks * kPanelK + t4) * LDT + col0 + g8, sc1, wm, g8, t4, 0, hi);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + cst);
        if (++cst == STAGES) { cst = 0; cph ^= 1; }
      }
      group_sync(bar_id, GT);   // every warp of the group has read the rows this block overwrites
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int row = i * BM + ti * 8 * WM + wm * 8 + g8, col = col0 + j * 8 + 2 * t4;
          *reinterpret_cast<double2*>(tile + row * LDT + col) = make_double2(c0[ti][j], c1[ti][j]);
        }
      group_sync(bar_id, GT);
    }

    PH_ADD(ph_p1, ph_d);
    PH_MARK(ph_e);
    // ---- kernel adjoint on the resident K-bar tile, in accumulator-fragment layout ----
    // The FP64 (non-tensor) pipe runs 64 FMA per clock and SM, and a row-per-thread sweep (distances by differences, dl_j += gb t_j^2,
    // then a second sweep for the input gradient) needed ~85 FP64 instructions per element of the tile. Here the distances come from
    // the expanded square with the dot products on the tensor pipe (as in the forward kernel), and everything that is a contraction of
    // Gbar moves there too:  U = Gbar^T [Zs, 1]  (input gradient and column sums T0),  S = Gbar [Xs, 1]  (row sums S0 and S1), with
    //   dX[c][j] = (U[c][j] - x_cj T0_c) / l_j,      dl_j = (sum_m z_mj^2 S0_m - 2 z_mj S1_mj + sum_c x_cj^2 T0_c) / l_j.
    constexpr int TNG = GC / 8, KS = DMAX / 4, NTJ = DMAX / 8 + 1, KC = GC / 4, MB = 4;
    const int J = a.D_in + 1, nmt = a.Mp / 8;
    double ds2 = 0.0, dla[NTJ - 1][2];
#pragma unroll
    for (int jt = 0; jt < NTJ - 1; ++jt) { dla[jt][0] = 0.0; dla[jt][1] = 0.0; }
    {
      double xx0[TNG], xx1[TNG], bx[KS][TNG];
#pragma unroll
      for (int nt = 0; nt < TNG; ++nt) {
        const int col = nt * 8 + 2 * t4;
        double q0 = 0.0, q1 = 0.0;
        for (int j = 0; j < a.D_in; ++j) {
          const int sw = XSW * ((j & 3) << 3);
          const double v0 = xs[j * GC + (col ^ sw)], v1 = xs[j * GC + ((col + 1) ^ sw)];
          q0 = fma(v0, v0, q0); q1 = fma(v1, v1, q1);
        }
        xx0[nt] = q0; xx1[nt] = q1;
      }
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const int j = kk * 4 + t4;
#pragma unroll
        for (int nt = 0; nt < TNG; ++nt) bx[kk][nt] = j < a.D_in ? xs[j * GC + ((nt * 8 + g8) ^ (XSW * (t4 << 3)))] : 0.0;
      }
      auto sweep = [&](auto kconst) {
        constexpr int KIND = decltype(kconst)::value;
        for (int mt0 = wm; mt0 < nmt; mt0 += WM * MB) {
          double av[MB][KS], zzv[MB];
#pragma unroll
          for (int u = 0; u < MB; ++u) {
            const int m = (mt0 + u * WM) * 8 + g8;
            const bool ml = mt0 + u * WM < nmt && m < a.M;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
              const int j = kk * 4 + t4;
              av[u][kk] = (ml && j < a.D_in) ? a.Zs[(long)m * a.D_in + j] : 0.0;
            }
            zzv[u] = ml ? a.zz[m] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < MB; ++u) {
            const int mt = mt0 + u * WM;
            if (mt >= nmt) break;
            const int m = mt * 8 + g8;
            const bool ml = m < a.M;
            double e0[TNG], e1[TNG];
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) { e0[nt] = 0.0; e1[nt] = 0.0; }
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
#pragma unroll
              for (int nt = 0; nt < TNG; ++nt) dmma884(e0[nt], e1[nt], av[u][kk], bx[kk][nt]);
            double ka[TNG], kb[TNG], ga[TNG], gb[TNG];
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) {
              kernel_eval(KIND, fmax(fma(-2.0, e0[nt], zzv[u] + xx0[nt]), 0.0), s2, ka[nt], ga[nt]);
              kernel_eval(KIND, fmax(fma(-2.0, e1[nt], zzv[u] + xx1[nt]), 0.0), s2, kb[nt], gb[nt]);
            }
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) {
              double2* tp = reinterpret_cast<double2*>(tile + m * LDT + col0 + nt * 8 + 2 * t4);
              const double2 kbar = *tp;
              if (ml) { ds2 = fma(kbar.x, ka[nt], ds2); ds2 = fma(kbar.y, kb[nt], ds2); }
              *tp = ml ? make_double2(kbar.x * ga[nt], kbar.y * gb[nt]) : make_double2(0.0, 0.0);
            }
          }
        }
      };
      if (a.kind == 0) sweep(std::integral_constant<int, 0>{});
      else if (a.kind == 1) sweep(std::integral_constant<int, 1>{});
      else sweep(std::integral_constant<int, 2>{});
    }
    PH_ADD(ph_row, ph_e);
    PH_MARK(ph_f);
    group_sync(bar_id, GT);   // Gbar rows of the group's columns are complete
    // ---- U = Gbar^T [Zs, 1]: one column tile per warp, 8 k-steps per round (their inducing-input loads in flight together),
    // even / odd k-steps in separate accumulators (a dependent DMMA chain issues every 26 clocks) ----
    for (int ct = wm; ct < TNG; ct += WM) {
      double u0[2][NTJ], u1[2][NTJ];
#pragma unroll
      for (int jt = 0; jt < NTJ; ++jt) { u0[0][jt] = u0[1][jt] = u1[0][jt] = u1[1][jt] = 0.0; }
      const double* ta = tile + t4 * LDT + col0 + ct * 8 + g8;
      for (int m0 = 0; m0 < a.Mp; m0 += 32) {
        double av[8], bz[8][NTJ];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int m = m0 + q * 4 + t4;
          av[q] = ta[(m0 + q * 4) * LDT];
#pragma unroll
          for (int jt = 0; jt < NTJ; ++jt) {
            const int j = jt * 8 + g8;
            bz[q][jt] = m < a.M ? (j < a.D_in ? a.Zs[(long)m * a.D_in + j] : (j == a.D_in ? 1.0 : 0.0)) : 0.0;
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int jt = 0; jt < NTJ; ++jt) dmma884(u0[q & 1][jt], u1[q & 1][jt], av[q], bz[q][jt]);
      }
#pragma unroll
      for (int jt = 0; jt < NTJ; ++jt) {
        const int j = jt * 8 + 2 * t4, c = ct * 8 + g8;
        if (j < J) us[c * J + j] = u0[0][jt] + u0[1][jt];
        if (j + 1 < J) us[c * J + j + 1] = u1[0][jt] + u1[1][jt];
      }
    }
    // ---- S = Gbar [Xs, 1], contracted on the fly with the inducing inputs ----
    {
      double bxs[KC][NTJ];
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        const int c = kk * 4 + t4;
#pragma unroll
        for (int jt = 0; jt < NTJ; ++jt) {
          const int j = jt * 8 + g8;
          bxs[kk][jt] = j < a.D_in ? xs[j * GC + (c ^ (XSW * ((j & 3) << 3)))] : (j == a.D_in ? 1.0 : 0.0);
        }
      }
      const int jt_one = a.D_in >> 3, src = (lane & ~3) | ((a.D_in & 7) >> 1), comp = a.D_in & 1;   // where column D_in (the ones) lands
      for (int mt0 = wm; mt0 < nmt; mt0 += WM * MB) {
        double zv[MB][NTJ - 1][2];
#pragma unroll
        for (int u = 0; u < MB; ++u) {
          const int m = (mt0 + u * WM) * 8 + g8;
          const bool ml = mt0 + u * WM < nmt && m < a.M;
#pragma unroll
          for (int jt = 0; jt < NTJ - 1; ++jt) {
            const int j = jt * 8 + 2 * t4;
            zv[u][jt][0] = (ml && j < a.D_in) ? a.Zs[(long)m * a.D_in + j] : 0.0;
            zv[u][jt][1] = (ml && j + 1 < a.D_in) ? a.Zs[(long)m * a.D_in + j + 1] : 0.0;
          }
        }
        double s0[MB][NTJ], s1[MB][NTJ];
#pragma unroll
        for (int u = 0; u < MB; ++u)
#pragma unroll
          for (int jt = 0; jt < NTJ; ++jt) { s0[u][jt] = 0.0; s1[u][jt] = 0.0; }
#pragma unroll
        for (int kk = 0; kk < KC; ++kk)
#pragma unroll
          for (int u = 0; u < MB; ++u) {
            const int mt = mt0 + u * WM;
            const double avv = mt < nmt ? tile[(mt * 8 + g8) * LDT + col0 + kk * 4 + t4] : 0.0;
#pragma unroll
            for (int jt = 0; jt < NTJ; ++jt) dmma884(s0[u][jt], s1[u][jt], avv, bxs[kk][jt]);
          }
#pragma unroll
        for (int u = 0; u < MB; ++u) {
          double cand = 0.0;
#pragma unroll
          for (int jt = 0; jt < NTJ; ++jt)
            if (jt == jt_one) cand = comp ? s1[u][jt] : s0[u][jt];
          const double S0 = __shfl_sync(0xffffffffu, cand, src);
#pragma unroll
          for (int jt = 0; jt < NTJ - 1; ++jt) {
            dla[jt][0] = fma(zv[u][jt][0], fma(zv[u][jt][0], S0, -2.0 * s0[u][jt]), dla[jt][0]);
            dla[jt][1] = fma(zv[u][jt][1], fma(zv[u][jt][1], S0, -2.0 * s1[u][jt]), dla[jt][1]);
          }
        }
      }
    }
    // per-tile, per-group partial sums in a fixed order: lanes of equal t4 hold the same j's, so a shuffle tree over g8, then the
    // WM warps through shared memory
#pragma unroll
    for (int jt = 0; jt < NTJ - 1; ++jt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double v = dla[jt][h];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int j = jt * 8 + 2 * lane + h;
        if (lane < 4 && j < a.D_in) red[wm * (kMaxD + 1) + j] = v;
      }
    {
      const double v = warp_sum(ds2 / s2);
      if (lane == 0) red[wm * (kMaxD + 1) + a.D_in] = v;
    }
    group_sync(bar_id, GT);   // U and the warps' partial sums are in shared memory
    if (tg <= a.D_in) {
      double sacc = 0.0;
#pragma unroll
      for (int w = 0; w < WM; ++w) sacc += red[w * (kMaxD + 1) + tg];
      if (tg < a.D_in) {
        double s2j = 0.0;
        const int sw = XSW * ((tg & 3) << 3);
        for (int c = 0; c < GC; ++c) {
          const double xv = xs[tg * GC + (c ^ sw)];
          s2j = fma(xv * xv, us[c * J + a.D_in], s2j);
        }
        sacc = (sacc + s2j) * (1.0 / a.ls[tg]);
      }
      a.part[(tile_index * WN + wn) * (a.D_in + 1) + tg] = sacc;
    }
    PH_ADD(ph_red, ph_f);
    PH_MARK(ph_g);
    // Gbar tile -> HBM (row segments of GC doubles)
    for (int idx = tg; idx < a.Mp * (GC / 2); idx += GT) {
      const int m = idx / (GC / 2), c2 = (idx % (GC / 2)) * 2;
      *reinterpret_cast<double2*>(a.Gbar + (long)m * a.Pp + p0 + c2) = *reinterpret_cast<const double2*>(tile + m * LDT + col0 + c2);
    }
    // input gradient: dX[p][j] = (U[c][j] - x_cj T0_c) / l_j  (+ mean-function path)
    if (a.dXin) {
      for (int o = tg; o < GC * a.D_in; o += GT) {
        const int c = o % GC, j = o / GC;
        const long p = p0 + c;
        if (p < a.P) {
          const double xv = xs[j * GC + (c ^ (XSW * ((j & 3) << 3)))];
          double v = (us[c * J + j] - xv * us[c * J + a.D_in]) * (1.0 / a.ls[j]);
          const double* gm = a.Gm + p * a.gm_ld;
          if (a.mean_kind == 1) v += gm[j];
          else if (a.mean_kind == 2)
            for (int d = 0; d < a.D_out; ++d) v = fma(gm[d], a.mfW[j * a.D_out + d], v);
          a.dXin[p * a.D_in + j] = v;
        }
      }
    }
    // XaugPad rows [x, 1, 0 ...] (unscaled inputs)
    for (int idx = tg; idx < GC * 32; idx += GT) {
      const int c = idx / 32, jj = idx % 32;
      const long p = p0 + c;
      double v = 0.0;
      if (p < a.P) v = jj < a.D_in ? a.Xin[(p % a.xmod) * a.D_in + jj] : (jj == a.D_in ? 1.0 : 0.0);
      a.XaugPad[p * 32 + jj] = v;
    }
    PH_ADD(ph_out, ph_g);
  }
#ifdef DGP_DEBUG_PHASECLK
  if (lane == 0 && blockIdx.x == 1 && my_tiles > 10 && (warp == 0 || warp == 5)) {
    const double tot = (double)(clock64() - ph_t0);
    printf("fused_bwd cta %d warp %d (D_out %d, BM %d, tiles %d): total %.0f clk/tile | prologue %.1f%% pass0 %.1f%% (slab/panel waits %.1f%%) block-end %.1f%% pass1 %.1f%% row sweep %.1f%% reductions %.1f%% stores+dX %.1f%%\n",
           blockIdx.x, warp, a.D_out, BM, my_tiles, tot / my_tiles, 100.0 * ph_pro / tot, 100.0 * ph_p0 / tot, 100.0 * ph_wait / tot, 100.0 * ph_be / tot,
           100.0 * ph_p1 / tot, 100.0 * ph_row / tot, 100.0 * ph_red / tot, 100.0 * ph_out / tot);
  }
#endif
}

}  // namespace dgp
