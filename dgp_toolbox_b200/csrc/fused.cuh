// Fused SVGP-layer conditional + sample kernel (the north-star kernel).
//
// One persistent CTA per SM owns tiles of PT point-samples. For each tile, entirely on chip:
//   1. Kuf tile [Mp x PT] built in shared memory from the layer input rows (RBF/ARD, FP64 exp)         (covs.Kuf, layers.py:243)
//   2. V = Lu^-1 Kuf      in place, row blocks descending   (FP64 DMMA, operator panels streamed from L2) (layers.py:245)
//   3. A = Lu^-T V        in place, row blocks ascending                                                   (layers.py:247)
//   4. T_d = q_sqrt_d^T A for every output d: only column sums of squares are kept                         (layers.py:257-271)
//   5. mean = A^T q_mu + mf(x), var = s2 - |V|^2 + |T_d|^2, z (Philox or supplied), F = mean + z sqrt(var + jitter)
//      written as [P][D_out]                                                     (layers.py:249,272-278; utils/utils.py:40-41)
// Kuf, V never touch HBM; A and T_d are written to HBM only when the adjoint will need them (training stash).
//
// The triangular operators (Lu^-1 lower, Lu^-T upper, q_sqrt_d^T upper) are pre-packed once per step by
// pack_stream_kernel into ONE linear stream of [BM x 16] panels in exactly the order the tile loop consumes them, each panel
// already in the XOR-swizzled shared-memory layout, so the producer side is a linear cp.async copy through a STAGES-deep
// ring and the consumer side is a flat loop over a small panel schedule. Inside diagonal blocks each warp skips the
// k-steps above/below its own 8-row m-tiles, so executed DMMA work stays within a few percent of the triangular minimum.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace dgp {

struct PanelDesc {
  int kind;    // 0: V = Linv * tile (lower), 1: A = LinvT * tile (upper), 2: T_d = RpT_d * tile (upper)
  int d;       // output index for kind 2
  int i;       // row block
  int k0;      // first k of the panel (absolute row of the resident tile)
  int flags;   // see below
  int pad;
};
constexpr int kPanelFirst = 1, kPanelLast = 2, kPanelClip = 4, kPanelStageEnd = 8;
constexpr int kPanelK = 16;

__device__ __forceinline__ int panel_swz(int row, int k) { return row * kPanelK + ((((k >> 2) ^ (row & 3)) << 2) | (k & 3)); }

// The panel schedule as an iterator (same order as build_schedule() on the host, which feeds pack_stream_kernel):
//   pass 0:        V  = Linv  * tile, row blocks i = nb-1 .. 0, k-panels 0 .. (i+1)*BM/16 - 1          (lower operator)
//   pass 1:        A  = LinvT * tile, row blocks i = 0 .. nb-1, k-panels i*BM/16 .. Mp/16 - 1          (upper operator)
//   pass 2 + d:    T_d = RpT_d * tile, same block order as pass 1
// Kept in registers: no global loads inside the panel loop (an LDG there shares the scoreboard the cp.async ring uses
// and would drain the whole ring every iteration).
template <int BM>
struct PanelIter {
  int pass, i, ks, ks_end, nb, kt;
  __device__ __forceinline__ void init(int Mp) {
    nb = Mp / BM; kt = Mp / kPanelK;
    pass = 0; i = nb - 1; ks = 0; ks_end = nb * (BM / kPanelK);
  }
  __device__ __forceinline__ PanelDesc get() const {
    PanelDesc e;
    e.kind = pass == 0 ? 0 : (pass == 1 ? 1 : 2);
    e.d = pass >= 2 ? pass - 2 : 0;
    e.i = i; e.k0 = ks * kPanelK; e.pad = 0;
    const bool last = ks == ks_end - 1;
    int fl = last ? kPanelLast : 0;
    if (pass == 0) {
      if (ks == 0) fl |= kPanelFirst;
      if (e.k0 >= i * BM) fl |= kPanelClip;
      if (last && i == 0) fl |= kPanelStageEnd;
    } else {
      if (ks == i * (BM / kPanelK)) fl |= kPanelFirst;
      if (e.k0 < (i + 1) * BM) fl |= kPanelClip;
      if (last && i == nb - 1) fl |= kPanelStageEnd;
    }
    e.flags = fl;
    return e;
  }
  __device__ __forceinline__ void next() {
    if (++ks < ks_end) return;
    if (pass == 0) {
      if (--i >= 0) { ks = 0; ks_end = (i + 1) * (BM / kPanelK); return; }
      pass = 1; i = 0;
    } else {
      if (++i >= nb) { ++pass; i = 0; }
    }
    ks = i * (BM / kPanelK); ks_end = kt;
  }
};

// One CTA per panel: copies the [BM x 16] block of the source operator into the stream, swizzled.
template <int BM>
__global__ void __launch_bounds__(256) pack_stream_kernel(const PanelDesc* __restrict__ sched, const double* __restrict__ Linv,
                                                          const double* __restrict__ LinvT, const double* __restrict__ RpT, int Mp,
                                                          double* __restrict__ stream) {
  const PanelDesc e = sched[blockIdx.x];
  const double* src = e.kind == 0 ? Linv : (e.kind == 1 ? LinvT : RpT + (long)e.d * Mp * Mp);
  double* dst = stream + (long)blockIdx.x * BM * kPanelK;
  for (int idx = threadIdx.x; idx < BM * kPanelK; idx += blockDim.x) {
    const int r = idx / kPanelK, k = idx % kPanelK;
    dst[panel_swz(r, k)] = src[(long)(e.i * BM + r) * Mp + e.k0 + k];
  }
}

// Zs[m][j] = Z[m][j] * (1 / l_j)
__global__ void scale_z_kernel(const double* __restrict__ Z, const double* __restrict__ ls, int M, int D, double* __restrict__ Zs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < M * D) Zs[idx] = Z[idx] * (1.0 / ls[idx % D]);
}

struct FusedFwdArgs {
  const double* stream; const PanelDesc* sched; int NP;
  const double* Zs;                                     // [M][D_in] inducing inputs / lengthscales
  const double* ls; const double* var;                  // [D_in], [1]
  const double* qmu;                                    // [M][D_out]
  const double* Xin; long xmod; int D_in;               // layer input
  const double* mfW; const double* mfb; int mean_kind;
  const double* z_in;                                   // caller [S][N_total][D_out] or null -> Philox
  unsigned long long seed; int layer; long Nc; long N_total; long n0; long n_offset;
  int M, Mp, D_out; long P, Pp;
  double jitter;
  double* Fmean; double* Fvar; double* F; double* z;    // chunk-local [P][D_out]; F / z may be null
  double* xFmean; double* xFvar; double* xF;            // caller-visible, any may be null
  double* stashA; double* stashT;                       // [Mp][Pp], [D_out][Mp][Pp] or null
};

template <int BM, int PT, int WM, int WN>
struct FusedCfg {
  static_assert(WM * WN == 8, "8 warps");
  static constexpr int THREADS = 256;
  static constexpr int TM = BM / (8 * WM), TN = PT / (8 * WN);
  static constexpr int LDT = PT + 4;
  static constexpr int PANEL = BM * kPanelK;
  static constexpr int STAGES = 4;
  static size_t smem_bytes(int Mp, int D_in, int D_out) {
    return ((size_t)Mp * LDT + (size_t)STAGES * PANEL + (size_t)D_in * PT + (size_t)(1 + D_out) * PT + (size_t)WM * PT) * sizeof(double);
  }
};

template <int BM, int PT, int WM, int WN>
__global__ void __launch_bounds__(256, 1) fused_forward_kernel(FusedFwdArgs a) {
  using Cfg = FusedCfg<BM, PT, WM, WN>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, LDT = Cfg::LDT, PANEL = Cfg::PANEL, STAGES = Cfg::STAGES;
  extern __shared__ __align__(16) double smem[];
  double* tile = smem;                                   // [Mp][LDT]   Kuf -> V -> A
  double* pbuf = tile + (size_t)a.Mp * LDT;              // [STAGES][PANEL]
  double* xs = pbuf + STAGES * PANEL;                    // [D_in][PT]  scaled inputs
  double* colsum = xs + a.D_in * PT;                     // [1 + D_out][PT]: |V|^2, |T_d|^2
  double* part = colsum + (1 + a.D_out) * PT;            // [WM][PT]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int wm = warp / WN, wn = warp % WN;
  const int ntiles = (int)(a.Pp / PT);   // padded tiles too: the stash planes must be fully written (zeros beyond P)
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const double s2 = a.var[0];

  // producer side of the ring: panel `iq` of the stream goes to stage `ist`; `ileft` panels remain for this CTA
  int iq = 0, ist = 0;
  long ileft = (long)my_tiles * a.NP;
  const double* isrc = a.stream + tid * 2;
  auto issue = [&]() {
    if (ileft > 0) {
      double* dst = pbuf + ist * PANEL + tid * 2;
#pragma unroll
      for (int c = 0; c < PANEL / 2 / 256; ++c) cp_async16(dst + c * 512, isrc + c * 512);
      --ileft;
      isrc += PANEL;
      if (++iq == a.NP) { iq = 0; isrc = a.stream + tid * 2; }
      if (++ist == STAGES) ist = 0;
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue();

  int cst = 0;   // consumer stage
  for (int tl = 0; tl < my_tiles; ++tl) {
    const long p0 = (long)(blockIdx.x + tl * gridDim.x) * PT;
    __syncthreads();   // previous tile's epilogue is done with tile / colsum
    // ---- stage 1: scaled inputs, then the Kuf tile ----
    for (int idx = tid; idx < a.D_in * PT; idx += 256) {
      const int j = idx / PT, c = idx % PT;
      const long p = p0 + c;
      xs[idx] = (p < a.P) ? a.Xin[(p % a.xmod) * a.D_in + j] * (1.0 / a.ls[j]) : 0.0;
    }
    __syncthreads();
    {
      const int c = tid % PT, mg = tid / PT;
      constexpr int MG = 256 / PT;
      const bool live = p0 + c < a.P;
      for (int m = mg; m < a.Mp; m += MG) {
        double k = 0.0;
        if (m < a.M && live) {
          double r2 = 0.0;
          const double* zr = a.Zs + (long)m * a.D_in;
          for (int j = 0; j < a.D_in; ++j) {
            const double t = zr[j] - xs[j * PT + c];
            r2 = fma(t, t, r2);
          }
          k = s2 * exp(-0.5 * r2);
        }
        tile[m * LDT + c] = k;
      }
    }
    // ---- stages 2-4: flat loop over the operator panels ----
    double c0[TM][TN], c1[TM][TN], sq0[TN], sq1[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { sq0[j] = 0.0; sq1[j] = 0.0; }
    PanelIter<BM> it;
    it.init(a.Mp);
    for (int q = 0; q < a.NP; ++q, it.next()) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue();
      const PanelDesc e = it.get();
      if (e.flags & kPanelFirst) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) { c0[i][j] = 0.0; c1[i][j] = 0.0; }
      }
      const double* pan = pbuf + cst * PANEL;
      if (++cst == STAGES) cst = 0;
      const double* bt = tile + (e.k0 + t4) * LDT + wn * TN * 8 + g8;
      // m-tiles of this warp that intersect the operator's triangle inside this panel: [imin, imax). The rest of the
      // panel is zero for them, so whole (panel, m-tile) pairs are skipped with a real (warp-uniform) branch.
      int imin = 0, imax = TM;
      if (e.flags & kPanelClip) {
        const int krel = e.k0 - e.i * BM;   // k of the panel relative to the diagonal block
        if (e.kind == 0) {                  // lower operator: row r needs k <= r  ->  rt + 7 >= krel
          const int num = krel - 7 - wm * 8;
          imin = num > 0 ? (num + 8 * WM - 1) / (8 * WM) : 0;
        } else {                            // upper operator: k >= r  ->  krel + 15 >= rt
          const int num = krel + kPanelK - 1 - wm * 8;
          imax = num >= 0 ? min(TM, num / (8 * WM) + 1) : 0;
        }
      }
      double bv[4][TN];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int j = 0; j < TN; ++j) bv[kk][j] = bt[kk * 4 * LDT + j * 8];
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        if (i >= imin && i < imax) {
          const int row = i * 8 * WM + wm * 8 + g8;
          const double* pr = pan + row * kPanelK + t4;
          double av[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) av[kk] = pr[(kk ^ (row & 3)) << 2];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int j = 0; j < TN; ++j) dmma884(c0[i][j], c1[i][j], av[kk], bv[kk][j]);
        }
      }
      if (e.flags & kPanelLast) {
        if (e.kind != 1) {   // column sums of squares of V / T_d
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { sq0[j] = fma(c0[i][j], c0[i][j], sq0[j]); sq1[j] = fma(c1[i][j], c1[i][j], sq1[j]); }
        }
        if (e.kind != 2) {   // in-place update of the resident tile
          __syncthreads();   // every warp has finished reading the rows this block overwrites
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
              const int row = e.i * BM + i * 8 * WM + wm * 8 + g8, col = wn * TN * 8 + j * 8 + 2 * t4;
              *reinterpret_cast<double2*>(tile + (long)row * LDT + col) = make_double2(c0[i][j], c1[i][j]);
            }
        }
        double* st = e.kind == 1 ? a.stashA : (e.kind == 2 && a.stashT ? a.stashT + (long)e.d * a.Mp * a.Pp : nullptr);
        if (st) {
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) {
              const int row = e.i * BM + i * 8 * WM + wm * 8 + g8, col = wn * TN * 8 + j * 8 + 2 * t4;
              *reinterpret_cast<double2*>(st + (long)row * a.Pp + p0 + col) = make_double2(c0[i][j], c1[i][j]);
            }
        }
        if ((e.flags & kPanelStageEnd) && e.kind != 1) {
          // reduce the per-thread partial sums over the 8 row lanes, then over the WM warps (fixed order)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              sq0[j] += __shfl_xor_sync(0xffffffffu, sq0[j], o);
              sq1[j] += __shfl_xor_sync(0xffffffffu, sq1[j], o);
            }
            if (g8 == 0) {
              const int col = wn * TN * 8 + j * 8 + 2 * t4;
              part[wm * PT + col] = sq0[j];
              part[wm * PT + col + 1] = sq1[j];
            }
            sq0[j] = 0.0; sq1[j] = 0.0;
          }
          __syncthreads();
          if (tid < PT) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < WM; ++w) s += part[w * PT + tid];
            colsum[(e.kind == 0 ? 0 : 1 + e.d) * PT + tid] = s;
          }
        }
      }
    }
    __syncthreads();
    // ---- stage 5: moments, sample, outputs ----
    for (int idx = tid; idx < PT * a.D_out; idx += 256) {
      const int c = idx / a.D_out, d = idx % a.D_out;
      const long p = p0 + c;
      if (p >= a.P) continue;
      double mean = 0.0;
      for (int m = 0; m < a.M; ++m) mean = fma(tile[m * LDT + c], a.qmu[m * a.D_out + d], mean);
      const double* x = a.Xin + (p % a.xmod) * a.D_in;
      double mf = 0.0;
      if (a.mean_kind == 1) mf = x[d];
      else if (a.mean_kind == 2) {
        for (int j = 0; j < a.D_in; ++j) mf = fma(x[j], a.mfW[j * a.D_out + d], mf);
        if (a.mfb) mf += a.mfb[d];
      }
      const double mu = mean + mf;
      const double var = s2 - colsum[c] + colsum[(1 + d) * PT + c];
      const long s = p / a.Nc, n = p % a.Nc;
      const long xrow = s * a.N_total + a.n0 + n;
      a.Fmean[p * a.D_out + d] = mu;
      a.Fvar[p * a.D_out + d] = var;
      if (a.xFmean) a.xFmean[xrow * a.D_out + d] = mu;
      if (a.xFvar) a.xFvar[xrow * a.D_out + d] = var;
      if (a.F) {
        const double z = a.z_in ? a.z_in[xrow * a.D_out + d]
                                : philox_normal(a.seed, (uint32_t)a.layer, (uint32_t)s, (uint32_t)(a.n0 + n + a.n_offset), (uint32_t)d);
        if (a.z) a.z[p * a.D_out + d] = z;
        const double f = mu + z * sqrt(var + a.jitter);
        a.F[p * a.D_out + d] = f;
        if (a.xF) a.xF[xrow * a.D_out + d] = f;
      }
    }
  }
  cp_async_wait<0>();
}

}  // namespace dgp
