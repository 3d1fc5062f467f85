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
};

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
    d = q / ((i + 1) * kpb); ks = q % ((i + 1) * kpb);
  } else {
    q -= NP0; pass = 1;
    for (i = 0;; ++i) { const int cnt = (nb - i) * kpb; if (q < cnt) break; q -= cnt; }
    ks = i * kpb + q;
  }
  double* dst = stream + (long)blockIdx.x * BM * kPanelK;
  for (int idx = threadIdx.x; idx < BM * kPanelK; idx += blockDim.x) {
    const int r = idx % BM, k = idx / BM;            // consecutive threads walk the rows: coalesced reads of C_d (row k of C_d = column k of C_d^T)
    const int row = i * BM + r, col = ks * kPanelK + k;
    dst[panel_swz(r, k)] = pass == 0 ? Cmat[((long)d * Mp + col) * Mp + row] : LinvT[(long)row * Mp + col];
  }
}

template <int BM, int PT, int WM, int WN>
struct FusedBwdCfg {
  static_assert(WM * WN == 8, "8 consumer warps");
  static constexpr int THREADS = 288;
  static constexpr int GT = WM * 32, GC = PT / WN;
  static constexpr int TM = BM / (8 * WM), TN = GC / 8;
  static constexpr int LDT = PT + 4;
  static constexpr int PANEL = BM * kPanelK, SLAB = kPanelK * LDT, STAGE = PANEL + SLAB;
  static constexpr int STAGES = BM >= 256 ? 2 : 3;
  static size_t smem_bytes(int Mp, int D_in, int D_out) {
    return ((size_t)STAGES * STAGE + (size_t)Mp * LDT + (size_t)D_in * PT + (size_t)D_out * PT + (size_t)PT + 2 * STAGES + 2 * WN * WM) * sizeof(double);
  }
};

template <int BM, int PT, int WM, int WN, int DMAX>
__global__ void __launch_bounds__(288, 1) fused_backward_kernel(FusedBwdArgs a) {
  using Cfg = FusedBwdCfg<BM, PT, WM, WN>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, LDT = Cfg::LDT, PANEL = Cfg::PANEL, STAGE = Cfg::STAGE, STAGES = Cfg::STAGES;
  constexpr int GT = Cfg::GT, GC = Cfg::GC, KPB = BM / kPanelK;
  extern __shared__ __align__(128) double bsmem[];
  double* pbuf = bsmem;                                   // [STAGES][PANEL | SLAB]
  double* tile = pbuf + STAGES * STAGE;                   // [Mp][LDT]   dV -> K-bar -> Gbar
  double* xs_all = tile + (size_t)a.Mp * LDT;             // [WN][D_in][GC] scaled inputs
  double* gv2_all = xs_all + a.D_in * PT;                 // [WN][D_out][GC] 2 Gv
  double* gq_all = gv2_all + a.D_out * PT;                // [WN][GC]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(gq_all + PT);
  unsigned long long* empty = full + STAGES;
  double* red_all = reinterpret_cast<double*>(empty + STAGES);   // [WN][WM][2] scratch of the per-tile group reductions

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.Mp / BM, kt = a.Mp / kPanelK;
  const int ntiles = (int)(a.Pp / PT);
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 8) {
    // ---- producer: per stage one operator panel, and in pass 0 the 16 row segments of the T_d slab it multiplies ----
    if (lane == 0) {
      int st = 0;
      unsigned ph = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const long p0 = (long)(blockIdx.x + tl * gridDim.x) * PT;
        const double* src = a.stream;
        for (int i = 0; i < nb; ++i)
          for (int d = 0; d < a.D_out; ++d)
            for (int ks = 0; ks < (i + 1) * KPB; ++ks) {
              mbar_wait(empty + st, ph ^ 1);
              double* dst = pbuf + st * STAGE;
              mbar_arrive_expect_tx(full + st, (PANEL + kPanelK * PT) * 8);
              bulk_g2s(dst, src, PANEL * 8, full + st);
              const double* trow = a.T + ((long)d * a.Mp + (long)ks * kPanelK) * a.Pp + p0;
#pragma unroll 4
              for (int r = 0; r < kPanelK; ++r) bulk_g2s(dst + PANEL + r * LDT, trow + (long)r * a.Pp, PT * 8, full + st);
              src += PANEL;
              if (++st == STAGES) { st = 0; ph ^= 1; }
            }
        for (int i = 0; i < nb; ++i)
          for (int ks = i * KPB; ks < kt; ++ks) {
            mbar_wait(empty + st, ph ^ 1);
            mbar_arrive_expect_tx(full + st, PANEL * 8);
            bulk_g2s(pbuf + st * STAGE, src, PANEL * 8, full + st);
            src += PANEL;
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
      }
    }
    return;
  }

  // ---- consumers ----
  const int g8 = lane >> 2, t4 = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int tg = wm * 32 + lane;
  const int col0 = wn * GC;
  double* xs = xs_all + wn * a.D_in * GC;
  double* gv2 = gv2_all + wn * a.D_out * GC;
  double* gqs = gq_all + wn * GC;
  double* red = red_all + wn * WM * 2;
  const int bar_id = 1 + wn;
  const double s2 = a.var[0];
  int cst = 0;
  unsigned cph = 0;
  double c0[TM][TN], c1[TM][TN];

  // one k-panel: A fragments from the swizzled panel, B fragments from `bt` (row stride LDT), m-tiles [lo, hi) only
  auto run_panel = [&](const double* pan, const double* bt, int lo, int hi, const double* sc) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      double bv[TN], av[TM];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = bt[kk * 4 * LDT + j * 8];
      if (sc) {
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[j] *= sc[j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int row = i * 8 * WM + wm * 8 + g8;
        av[i] = pan[row * kPanelK + (((kk ^ (row & 3)) << 2) | t4)];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
        if (i >= lo && i < hi) {   // warp-uniform: a predicated-off DMMA would still occupy the pipe
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(c0[i][j], c1[i][j], av[i], bv[j]);
        }
    }
  };

  for (int tl = 0; tl < my_tiles; ++tl) {
    const long tile_index = (long)blockIdx.x + (long)tl * gridDim.x;
    const long p0 = tile_index * PT + col0;
    group_sync(bar_id, GT);   // the previous tile's epilogue is done with the group's columns
    for (int idx = tg; idx < a.D_in * GC; idx += GT) {
      const int j = idx / GC, c = idx % GC;
      const long p = p0 + c;
      xs[idx] = (p < a.P) ? a.Xin[(p % a.xmod) * a.D_in + j] * (1.0 / a.ls[j]) : 0.0;
    }
    for (int idx = tg; idx < a.D_out * GC; idx += GT) {
      const int d = idx / GC, c = idx % GC;
      gv2[idx] = 2.0 * a.GvT[(long)d * a.Pp + p0 + c];
    }
    if (tg < GC) gqs[tg] = a.gq[p0 + tg];
    group_sync(bar_id, GT);

    // ---- pass 0: dV ----
    for (int i = 0; i < nb; ++i) {
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int j = 0; j < TN; ++j) { c0[ti][j] = 0.0; c1[ti][j] = 0.0; }
      for (int d = 0; d < a.D_out; ++d) {
        double sc[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) sc[j] = gv2[d * GC + j * 8 + g8];
        for (int ks = 0; ks < (i + 1) * KPB; ++ks) {
          const int num = ks * kPanelK - i * BM - 7 - wm * 8;   // lower operator: row r needs k <= r
          const int lo = num > 0 ? (num + 8 * WM - 1) / (8 * WM) : 0;
          const double* stage = pbuf + cst * STAGE;
          mbar_wait(full + cst, cph);
          if (lo < TM) run_panel(stage, stage + PANEL + t4 * LDT + col0 + g8, lo, TM, sc);
          __syncwarp();
          if (lane == 0) mbar_arrive(empty + cst);
          if (++cst == STAGES) { cst = 0; cph ^= 1; }
        }
      }
      // block end: + beta Gm^T + 2 V diag(gq); to the resident tile and to HBM
      for (int d = 0; d < a.D_out; ++d) {
        double gm0[TN], gm1[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const double* g = a.Gm + (p0 + j * 8 + 2 * t4) * a.gm_ld + d;
          gm0[j] = g[0]; gm1[j] = g[a.gm_ld];
        }
#pragma unroll
        for (int ti = 0; ti < TM; ++ti) {
          const double b = a.beta[(long)(i * BM + ti * 8 * WM + wm * 8 + g8) * 32 + d];
#pragma unroll
          for (int j = 0; j < TN; ++j) { c0[ti][j] = fma(b, gm0[j], c0[ti][j]); c1[ti][j] = fma(b, gm1[j], c1[ti][j]); }
        }
      }
#pragma unroll
      for (int ti = 0; ti < TM; ++ti) {
        const int row = i * BM + ti * 8 * WM + wm * 8 + g8;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int cl = j * 8 + 2 * t4;                      // column inside the group
          const long p = p0 + cl;
          const double2 v = *reinterpret_cast<const double2*>(a.V + (long)row * a.Pp + p);
          const double r0 = fma(2.0 * v.x, gqs[cl], c0[ti][j]), r1 = fma(2.0 * v.y, gqs[cl + 1], c1[ti][j]);
          *reinterpret_cast<double2*>(tile + row * LDT + col0 + cl) = make_double2(r0, r1);
          *reinterpret_cast<double2*>(a.dV + (long)row * a.Pp + p) = make_double2(r0, r1);
        }
      }
    }
    group_sync(bar_id, GT);   // the group's columns of dV are complete

    // ---- pass 1: K-bar = Lu^-T dV, in place, row blocks ascending ----
    for (int i = 0; i < nb; ++i) {
#pragma unroll
      for (int ti = 0; ti < TM; ++ti)
#pragma unroll
        for (int j = 0; j < TN; ++j) { c0[ti][j] = 0.0; c1[ti][j] = 0.0; }
      for (int ks = i * KPB; ks < kt; ++ks) {
        const int num = ks * kPanelK - i * BM + kPanelK - 1 - wm * 8;   // upper operator: row r needs k >= r
        const int hi = num >= 0 ? min(TM, num / (8 * WM) + 1) : 0;
        const double* stage = pbuf + cst * STAGE;
        mbar_wait(full + cst, cph);
        if (hi > 0) run_panel(stage, tile + (ks * kPanelK + t4) * LDT + col0 + g8, 0, hi, nullptr);
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

    // ---- kernel adjoint on the resident K-bar tile ----
    // row sweep: one inducing row per thread; Gbar = K-bar * (-2 dk/dr2) in place, row-local sums for dl_j and ds2
    double dl[DMAX], ds2 = 0.0;
#pragma unroll
    for (int j = 0; j < DMAX; ++j) dl[j] = 0.0;
    for (int m = tg; m < a.Mp; m += GT) {
      double* trow = tile + m * LDT + col0;
      if (m < a.M) {
        double zr[DMAX];
        const double* zg = a.Zs + (long)m * a.D_in;
#pragma unroll
        for (int j = 0; j < DMAX; ++j) zr[j] = j < a.D_in ? zg[j] : 0.0;
        for (int c = 0; c < GC; ++c) {
          double r2 = 0.0, t[DMAX];
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            if (j < a.D_in) {
              t[j] = zr[j] - xs[j * GC + c];
              r2 = fma(t[j], t[j], r2);
            }
          double k, gf;
          kernel_eval(a.kind, r2, s2, k, gf);
          const double kbar = trow[c];
          const double gb = kbar * gf;
          trow[c] = gb;
          ds2 = fma(kbar, k, ds2);
#pragma unroll
          for (int j = 0; j < DMAX; ++j)
            if (j < a.D_in) dl[j] = fma(gb * t[j], t[j], dl[j]);
        }
      } else {
        for (int c = 0; c < GC; ++c) trow[c] = 0.0;
      }
    }
    // per-tile, per-group partial sums in a fixed order: warp shuffle tree, then the WM warps through shared memory
    {
      double* pout = a.part + (tile_index * WN + wn) * (a.D_in + 1);
      for (int j = 0; j <= a.D_in; ++j) {
        double v = ds2 / s2;
        if (j < a.D_in) {
#pragma unroll
          for (int jj = 0; jj < DMAX; ++jj)
            if (jj == j) v = dl[jj] * (1.0 / a.ls[jj]);
        }
        v = warp_sum(v);
        if (lane == 0) red[wm * 2 + (j & 1)] = v;
        group_sync(bar_id, GT);
        if (tg == 0) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < WM; ++w) s += red[w * 2 + (j & 1)];
          pout[j] = s;
        }
      }
    }
    group_sync(bar_id, GT);   // Gbar rows of the group's columns are complete
    // Gbar tile -> HBM (row segments of GC doubles)
    for (int idx = tg; idx < a.Mp * (GC / 2); idx += GT) {
      const int m = idx / (GC / 2), c2 = (idx % (GC / 2)) * 2;
      *reinterpret_cast<double2*>(a.Gbar + (long)m * a.Pp + p0 + c2) = *reinterpret_cast<const double2*>(tile + m * LDT + col0 + c2);
    }
    // input gradient: dX[p][j] = (1/l_j) sum_m Gbar[m][p] (zs[m][j] - xs[p][j])  (+ mean-function path)
    if (a.dXin) {
      for (int idx = tg; idx < GC * a.D_in; idx += GT) {
        const int c = idx % GC, j = idx / GC;
        const long p = p0 + c;
        const double xv = xs[j * GC + c];
        const double* tc = tile + col0 + c;
        const double* zc = a.Zs + j;
        double s0 = 0.0, s1 = 0.0;
        int m = 0;
        for (; m + 2 <= a.M; m += 2) {
          s0 = fma(tc[m * LDT], zc[(long)m * a.D_in] - xv, s0);
          s1 = fma(tc[(m + 1) * LDT], zc[(long)(m + 1) * a.D_in] - xv, s1);
        }
        for (; m < a.M; ++m) s0 = fma(tc[m * LDT], zc[(long)m * a.D_in] - xv, s0);
        if (p < a.P) {
          double v = (s0 + s1) * (1.0 / a.ls[j]);
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
  }
}

}  // namespace dgp
