// Fused SVGP-layer conditional + sample kernel (the north-star kernel).
//
// One persistent CTA per SM owns tiles of PT point-samples. For each tile, entirely on chip:
//   1. Kuf tile [Mp x PT] built in shared memory from the layer input rows (RBF/ARD, FP64 exp)         (covs.Kuf, layers.py:243)
//   2. V = Lu^-1 Kuf      in place, row blocks descending   (FP64 DMMA, operator panels streamed from L2) (layers.py:245)
//   3. A = Lu^-T V        in place, row blocks ascending                                                   (layers.py:247)
//   4. T_d = q_sqrt_d^T A for every output d: only column sums of squares are kept                         (layers.py:257-271)
//   5. mean = A^T q_mu + mf(x), var = s2 - |V|^2 + |T_d|^2, z (Philox or supplied), F = mean + z sqrt(var + jitter)
//      written as [P][D_out]                                                     (layers.py:249,272-278; utils/utils.py:40-41)
// Kuf never touches HBM; the operand of the T_d passes and T_d itself are written to HBM only when the adjoint will need them
// (training stash, streaming stores).
// V-form (default): C_d = q_sqrt_d^T Lu^-T and beta = Lu^-1 q_mu are folded once per step, T_d = C_d V, mean = V^T beta, and pass 3
// disappears ((1 + D_out) M^2 instead of (2 + D_out) M^2 flops per point-sample, same values to rounding); V is stashed.
//
// The triangular operators (Lu^-1 lower, Lu^-T upper, q_sqrt_d^T or C_d upper) are pre-packed once per step by
// pack_stream_kernel into ONE linear stream of [BM x 16] panels in exactly the order the tile loop consumes them, each panel
// already in the XOR-swizzled shared-memory layout. A producer warp moves one panel per bulk copy (cp.async.bulk -> SASS
// UBLKCP, completion on an mbarrier) through a STAGES-deep ring; 8 consumer warps in WN column groups wait on the full
// barriers, run straight-line DMMA blocks specialised on the active m-tile range, and release the stage on the empty
// barriers. Inside diagonal blocks whole (panel, m-tile) pairs that hold only zeros are skipped with warp-uniform
// branches (a predicated-off DMMA still occupies the pipe), so executed DMMA work stays within ~6% of the triangular minimum.
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

// zz[m] = |Zs[m]|^2 (rows >= M: 0), the inducing-point half of the expanded squared distance
__global__ void zz_kernel(const double* __restrict__ Zs, int M, int Mp, int D, double* __restrict__ zz) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mp) return;
  double s = 0.0;
  if (m < M)
    for (int j = 0; j < D; ++j) s = fma(Zs[(long)m * D + j], Zs[(long)m * D + j], s);
  zz[m] = s;
}

struct FusedFwdArgs {
  const double* stream; const PanelDesc* sched; int NP;
  const double* Zs;                                     // [M][D_in] inducing inputs / lengthscales
  const double* zz;                                     // [Mp] |Zs[m]|^2
  const double* ls; const double* var;                  // [D_in], [1]
  const double* qmu; int qmu_ld;                        // [M][qmu_ld] weights of the mean: q_mu (ld = D_out), or beta = Lu^-1 q_mu in V-form
  int vform;                                            // 1: skip the A pass, T_d = C_d V with C_d = q_sqrt_d^T Lu^-T packed in the stream
  const double* Xin; long xmod; int D_in;               // layer input
  const double* mfW; const double* mfb; int mean_kind;
  int kind;                                             // kernel kind (common.cuh: kernel_eval)
  const double* z_in;                                   // caller [S][N_total][D_out] or null -> Philox
  unsigned long long seed; const unsigned long long* seed_ptr; int layer; long Nc; long N_total; long n0; long n_offset;
  int M, Mp, D_out; long P, Pp;
  double jitter;
  double* Fmean; double* Fvar; double* F; double* z;    // chunk-local [P][D_out]; F / z may be null
  double* xFmean; double* xFvar; double* xF;            // caller-visible, any may be null
  double* stashA; double* stashT;                       // [Mp][Pp], [D_out][Mp][Pp] or null
  int warp_major_groups;                                // 1: warps [g*WM, (g+1)*WM) form column group g; 0: group = warp % WN (default)
  int group_skew;                                       // measurement hook: clocks by which column group g starts late (g * group_skew)
};

// ---- mbarrier / bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA engine, no tensor map needed for a linear panel); completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Register re-partitioning between warp groups (setmaxnreg, sm_90a+): a CTA of 12 warps is launched with 168 registers per thread
// (3 warps per SM sub-partition x 32 x 168 fills its 16 K registers); the warp group that only issues bulk copies shrinks to 40
// and the two consumer groups grow to 232, which is what lets the straight-line DMMA blocks keep their accumulators, both operand
// fragment sets and the addressing in registers without spills.
constexpr int kRegsProducer = 40, kRegsConsumer = 232;
__device__ __forceinline__ void regs_shrink() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer)); }
__device__ __forceinline__ void regs_grow() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsConsumer)); }

__device__ __forceinline__ void group_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Straight-line DMMA block for one operator panel restricted to the warp's m-tiles [LO, HI): k-step outer so that only
// TN + (HI - LO) operand registers are live and the loads of k-step kk+1 overlap the DMMAs of k-step kk.
template <int LO, int HI, int TM, int TN, int WM, int LDT>
__device__ __forceinline__ void panel_block(double (&c0)[TM][TN], double (&c1)[TM][TN], const double* __restrict__ pan,
                                            const double* __restrict__ bt, int wm, int g8, int t4) {
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
#pragma unroll
    for (int i = LO; i < HI; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) dmma884(c0[i][j], c1[i][j], av[i - LO], bv[j]);
  }
}

// dispatch on the active m-tile range (a prefix [0, hi) for upper operators, a suffix [lo, TM) for lower ones)
template <int TM, int TN, int WM, int LDT>
__device__ __forceinline__ void panel_dispatch(double (&c0)[TM][TN], double (&c1)[TM][TN], const double* pan, const double* bt,
                                               int wm, int g8, int t4, int lo, int hi) {
  if (lo == 0) {
    if (hi == TM) { panel_block<0, TM, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; }
    if constexpr (TM >= 2) { if (hi == 1) { panel_block<0, 1, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
    if constexpr (TM >= 3) { if (hi == 2) { panel_block<0, 2, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
    if constexpr (TM >= 4) { if (hi == 3) { panel_block<0, 3, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
    return;   // hi == 0
  }
  if constexpr (TM >= 2) { if (lo == TM - 1) { panel_block<TM - 1, TM, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
  if constexpr (TM >= 3) { if (lo == TM - 2) { panel_block<TM - 2, TM, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
  if constexpr (TM >= 4) { if (lo == TM - 3) { panel_block<TM - 3, TM, TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4); return; } }
}

// Two panels per iteration that multiply the SAME rows of the resident tile: the T_d passes of outputs d and d + 1 (same row
// block, same k-range, hence the same active m-tile range). One set of B fragments, one descriptor decode, one dispatch and one
// block end for twice the DMMAs: the per-panel boundary code (~640 clocks per warp, profiles/r02v_fused_fwd_lines.txt) is paid once
// per pair. Needs a second set of accumulators, which the 232-register consumer budget allows.
template <int LO, int HI, int TM, int TN, int WM, int LDT>
__device__ __forceinline__ void panel_block2(double (&a0)[TM][TN], double (&a1)[TM][TN], double (&b0)[TM][TN], double (&b1)[TM][TN],
                                             const double* __restrict__ panA, const double* __restrict__ panB,
                                             const double* __restrict__ bt, int wm, int g8, int t4) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    double bv[TN], avA[HI - LO > 0 ? HI - LO : 1], avB[HI - LO > 0 ? HI - LO : 1];
#pragma unroll
    for (int j = 0; j < TN; ++j) bv[j] = bt[kk * 4 * LDT + j * 8];
#pragma unroll
    for (int i = LO; i < HI; ++i) {
      const int row = i * 8 * WM + wm * 8 + g8;
      const int off = row * kPanelK + (((kk ^ (row & 3)) << 2) | t4);
      avA[i - LO] = panA[off];
      avB[i - LO] = panB[off];
    }
#pragma unroll
    for (int i = LO; i < HI; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        dmma884(a0[i][j], a1[i][j], avA[i - LO], bv[j]);
        dmma884(b0[i][j], b1[i][j], avB[i - LO], bv[j]);
      }
  }
}

template <int TM, int TN, int WM, int LDT>
__device__ __forceinline__ void panel_dispatch2(double (&a0)[TM][TN], double (&a1)[TM][TN], double (&b0)[TM][TN], double (&b1)[TM][TN],
                                                const double* panA, const double* panB, const double* bt, int wm, int g8, int t4, int hi) {
  // upper operators only: the active m-tiles are a prefix [0, hi)
  if (hi == TM) { panel_block2<0, TM, TM, TN, WM, LDT>(a0, a1, b0, b1, panA, panB, bt, wm, g8, t4); return; }
  if constexpr (TM >= 2) { if (hi == 1) { panel_block2<0, 1, TM, TN, WM, LDT>(a0, a1, b0, b1, panA, panB, bt, wm, g8, t4); return; } }
  if constexpr (TM >= 3) { if (hi == 2) { panel_block2<0, 2, TM, TN, WM, LDT>(a0, a1, b0, b1, panA, panB, bt, wm, g8, t4); return; } }
  if constexpr (TM >= 4) { if (hi == 3) { panel_block2<0, 3, TM, TN, WM, LDT>(a0, a1, b0, b1, panA, panB, bt, wm, g8, t4); return; } }
}

// Thread layout: WM*WN = 8 consumer warps + 1 producer warp. The WN consumer GROUPS (WM warps each) own disjoint column
// ranges of the tile and run decoupled from each other: they share only the operator-panel ring, which the producer warp
// fills with one bulk copy per panel (full/empty mbarriers). While one group is in per-panel bookkeeping, block-end
// write-back, the Kuf build or the epilogue, the other groups keep the DMMA pipe busy.
template <int BM, int PT, int WM, int WN>
struct FusedCfg {
  static_assert(WM * WN == 8, "8 consumer warps");
  static constexpr bool REGSPLIT = !(BM == 64 && PT == 32);   // the 2-CTA/SM configuration keeps 9 warps at equal registers
  static constexpr int THREADS = REGSPLIT ? 384 : 288;
  static constexpr int GT = WM * 32;          // threads per consumer group
  static constexpr int GC = PT / WN;          // tile columns per group
  static constexpr int TM = BM / (8 * WM), TN = GC / 8;
  static constexpr int LDT = PT + 4;
  static constexpr int PANEL = BM * kPanelK;
  static constexpr int STAGES = 4;
  static size_t smem_bytes(int Mp, int D_in, int D_out) {
    return ((size_t)Mp * LDT + (size_t)STAGES * PANEL + (size_t)D_in * PT + (size_t)(1 + D_out) * PT + (size_t)2 * WM * PT + (size_t)D_out * PT + 2 * STAGES + (size_t)(Mp / BM) * (Mp / BM + 1) * (BM / kPanelK)) * sizeof(double);
  }
};

template <int BM, int PT, int WM, int WN>
__global__ void __launch_bounds__((FusedCfg<BM, PT, WM, WN>::THREADS), (BM == 64 && PT == 32) ? 2 : 1) fused_forward_kernel(FusedFwdArgs a) {
  using Cfg = FusedCfg<BM, PT, WM, WN>;
  constexpr int TM = Cfg::TM, TN = Cfg::TN, LDT = Cfg::LDT, PANEL = Cfg::PANEL, STAGES = Cfg::STAGES, GT = Cfg::GT, GC = Cfg::GC;
  extern __shared__ __align__(128) double fsmem[];
  double* pbuf = fsmem;                                  // [STAGES][PANEL]   (first: 128-byte aligned for the bulk copies)
  double* tile = pbuf + STAGES * PANEL;                  // [Mp][LDT]   Kuf -> V -> A
  double* xs_all = tile + (size_t)a.Mp * LDT;            // [WN][D_in][GC]  scaled inputs
  double* colsum_all = xs_all + a.D_in * PT;             // [WN][1 + D_out][GC]: |V|^2, |T_d|^2
  double* part_all = colsum_all + (1 + a.D_out) * PT;    // [WN][WM][GC]
  double* part2_all = part_all + WM * PT;                // the same for the second output of a pair of T_d passes
  double* mean_all = part2_all + WM * PT;                // [WN][GC][D_out]  tile^T * weights (stage 5)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(mean_all + a.D_out * PT);   // [STAGES]
  unsigned long long* empty = full + STAGES;                                               // [STAGES]
  // per-panel descriptors of the V pass and of one upper pass (A and every T_d share it), unpacked with a handful of
  // integer ops per panel instead of re-deriving block / clip state each time
  uint2* ptab = reinterpret_cast<uint2*>(empty + STAGES);                                   // [NPv + NPa]
  const int NPv = (a.Mp / BM) * (a.Mp / BM + 1) / 2 * (BM / kPanelK), NPa = NPv;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ntiles = (int)(a.Pp / PT);   // padded tiles too: the stash planes must be fully written (zeros beyond P)
  const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 32) {
    PanelIter<BM> it;
    it.init(a.Mp);
    for (int q = 0; q < NPv + NPa; ++q, it.next()) {
      const PanelDesc e = it.get();
      unsigned y = 0;
      for (int w = 0; w < WM; ++w) {
        int lo = 0, hi = BM / (8 * WM);
        if (e.flags & kPanelClip) {
          const int krel = e.k0 - e.i * BM;   // k of the panel relative to the diagonal block
          if (e.kind == 0) {                  // lower operator: row r needs k <= r  ->  rt + 7 >= krel
            const int num = krel - 7 - w * 8;
            lo = num > 0 ? (num + 8 * WM - 1) / (8 * WM) : 0;
          } else {                            // upper operator: k >= r  ->  krel + 15 >= rt
            const int num = krel + kPanelK - 1 - w * 8;
            hi = num >= 0 ? min(hi, num / (8 * WM) + 1) : 0;
          }
        }
        y |= (unsigned)(lo | (hi << 3)) << (6 * w);
      }
      ptab[q] = make_uint2((unsigned)e.k0 | ((unsigned)e.i << 12) | ((unsigned)e.flags << 18), y);
    }
  }
  __syncthreads();

  if (warp >= 8) {
    if constexpr (Cfg::REGSPLIT) regs_shrink();
    if (warp > 8) return;
    // ---- producer: one bulk copy per panel, in stream order, round after round ----
    if (lane == 0) {
      int st = 0;
      unsigned ph = 0;
      const double* src = a.stream;
      // Inside a diagonal block only the rows on the operator's side of the diagonal are read by the consumers (a suffix for
      // the lower operator, a prefix for the upper ones; same rule as their m-tile ranges), and a row range is a byte range
      // of the packed panel: copy just that. 2/3 of the panels are diagonal-block panels, half of whose rows are dead.
      auto issue = [&](unsigned x, bool lower) {
        int r0 = 0, nr = BM;
        if ((x >> 18) & kPanelClip) {
          const int krel = (int)(x & 0xfff) - (int)((x >> 12) & 63) * BM;
          if (lower) { r0 = krel; nr = BM - krel; }
          else nr = min(BM, krel + kPanelK);
        }
        mbar_wait(empty + st, ph ^ 1);
        mbar_arrive_expect_tx(full + st, (unsigned)(nr * kPanelK * 8));
        bulk_g2s(pbuf + st * PANEL + r0 * kPanelK, src + r0 * kPanelK, (unsigned)(nr * kPanelK * 8), full + st);
        src += PANEL;
        if (++st == STAGES) { st = 0; ph ^= 1; }
      };
      const int pairs = a.D_out >> 1;
      for (int tl = 0; tl < my_tiles; ++tl) {   // same order as build_schedule() on the host
        src = a.stream;
        for (int qq = 0; qq < NPv; ++qq) issue(ptab[qq].x, true);
        if (!a.vform)
          for (int qq = 0; qq < NPa; ++qq) issue(ptab[NPv + qq].x, false);
        for (int dp = 0; dp < pairs; ++dp)
          for (int qq = 0; qq < NPa; ++qq) { issue(ptab[NPv + qq].x, false); issue(ptab[NPv + qq].x, false); }
        if (a.D_out & 1)
          for (int qq = 0; qq < NPa; ++qq) issue(ptab[NPv + qq].x, false);
      }
    }
    return;
  }

  // ---- consumers ----
  if constexpr (Cfg::REGSPLIT) regs_grow();
  const int g8 = lane >> 2, t4 = lane & 3;
  // Warps w and w + 4 share an SM sub-partition (and its FP64 pipe). With group = warp / WM the two warps of a sub-partition belong
  // to DIFFERENT column groups, which run decoupled (they share only the panel ring): while one group is in a block-end phase
  // (barriers, write-back, stash stores) or in the Kuf / epilogue stages, the other keeps that sub-partition's DMMA pipe busy.
  const int wm = a.warp_major_groups ? warp % WM : warp / WN, wn = a.warp_major_groups ? warp / WM : warp % WN;       // group = wn
  const int tg = wm * 32 + lane;                  // thread index inside the group
  const int col0 = wn * GC;                       // first tile column of the group
  double* xs = xs_all + wn * a.D_in * GC;
  double* colsum = colsum_all + wn * (1 + a.D_out) * GC;
  double* part = part_all + wn * WM * GC;
  double* part2 = part2_all + wn * WM * GC;
  double* meanbuf = mean_all + wn * a.D_out * GC;
  const int bar_id = 1 + wn;
  const double s2 = a.var[0];
  int cst = 0;
  unsigned cph = 0;
#ifdef DGP_DEBUG_WAITCLK
  long long wait_clk = 0;
  const long long t_k0 = clock64();
#endif
  for (int idx = tg; idx < WM * GC; idx += GT) { part[idx] = 0.0; part2[idx] = 0.0; }
  if (a.group_skew > 0 && wn > 0) {
    const long long t0 = clock64();
    while (clock64() - t0 < (long long)a.group_skew * wn) {}
  }

#ifdef DGP_DEBUG_PHASECLK
  long long ph_kuf = 0, ph_loop = 0, ph_end = 0, ph_epi = 0, ph_wait = 0, ph_dot = 0, ph_k0 = 0, ph_k1 = 0, ph_k2 = 0;
  const long long ph_t0 = clock64();
#define PH_MARK(var) const long long var = clock64()
#define PH_ADD(acc, from) acc += clock64() - (from)
#else
#define PH_MARK(var)
#define PH_ADD(acc, from)
#endif
  for (int tl = 0; tl < my_tiles; ++tl) {
    const long p0 = (long)(blockIdx.x + tl * gridDim.x) * PT + col0;   // first point-sample of this group's columns
    PH_MARK(ph_a);
    group_sync(bar_id, GT);   // previous tile's epilogue is done with the group's columns / colsum
    PH_ADD(ph_k0, ph_a);
    PH_MARK(ph_a1);
    // ---- stage 1: scaled inputs, then the Kuf tile ----
    // xs[j][c] is stored with the 8-column blocks of row j rotated by j & 3 (GC = 32): the B-fragment reads below (four rows j,
    // eight columns) then touch 32 different banks.
    constexpr int XSW = GC == 32 ? 1 : 0;
    for (int idx = tg; idx < a.D_in * GC; idx += GT) {
      const int j = idx / GC, c = idx % GC;
      const long p = p0 + c;
      xs[j * GC + (c ^ (XSW * ((j & 3) << 3)))] = (p < a.P) ? a.Xin[(p % a.xmod) * a.D_in + j] * (1.0 / a.ls[j]) : 0.0;
    }
    group_sync(bar_id, GT);
    PH_ADD(ph_k1, ph_a1);
    PH_MARK(ph_a2);
    // Kuf through the expanded square (the reference's own form, gpflow square_distance): r2 = |z|^2 + |x|^2 - 2 z.x with the
    // [Mp x D_in] . [D_in x GC] dot products on the tensor pipe (one 8 x 8 unit per (m-tile, n-tile), D_in / 4 k-steps), then the
    // kernel function on the accumulator fragments: 8 independent exp chains per m-tile and thread, half the FP64 instructions
    // of a difference-based sweep.
    {
      constexpr int TNG = GC / 8;
      double xx0[TNG], xx1[TNG];
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
      // MB m-tiles per round with all their inducing-input loads issued up front: these come from L2 (the shared-memory carve-out
      // leaves ~12 KB of L1), and one m-tile at a time the stage was bound by that latency (2.6 k clocks per m-tile, measured)
      auto sweep = [&](auto kconst, auto ksconst) {
        constexpr int KIND = decltype(kconst)::value;
        constexpr int KS = decltype(ksconst)::value;   // k-steps of 4 input dimensions, zero-padded beyond D_in
        constexpr int MB = 4;                          // operand registers: MB * KS + KS * TNG doubles
        const int nmt = a.Mp / 8;
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
          double bv[KS][TNG];
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            const int j = kk * 4 + t4;
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) bv[kk][nt] = j < a.D_in ? xs[j * GC + ((nt * 8 + g8) ^ (XSW * (t4 << 3)))] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < MB; ++u) {
            const int mt = mt0 + u * WM;
            if (mt >= nmt) break;
            const int m = mt * 8 + g8;
            double e0[TNG], e1[TNG];
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) { e0[nt] = 0.0; e1[nt] = 0.0; }
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
#pragma unroll
              for (int nt = 0; nt < TNG; ++nt) dmma884(e0[nt], e1[nt], av[u][kk], bv[kk][nt]);
            const bool ml = m < a.M;
            double k0[TNG], k1[TNG];
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) {
              k0[nt] = kernel_value(KIND, fmax(fma(-2.0, e0[nt], zzv[u] + xx0[nt]), 0.0), s2);
              k1[nt] = kernel_value(KIND, fmax(fma(-2.0, e1[nt], zzv[u] + xx1[nt]), 0.0), s2);
            }
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) {
              const int col = nt * 8 + 2 * t4;
              *reinterpret_cast<double2*>(tile + m * LDT + col0 + col) =
                  make_double2((ml && p0 + col < a.P) ? k0[nt] : 0.0, (ml && p0 + col + 1 < a.P) ? k1[nt] : 0.0);
            }
          }
        }
      };
      // wider layers: one m-tile at a time, k-steps in a run-time loop (keeps the register footprint of this stage small)
      auto sweep_wide = [&](auto kconst) {
        constexpr int KIND = decltype(kconst)::value;
        const int ksteps = (a.D_in + 3) >> 2;
        for (int mt = wm; mt < a.Mp / 8; mt += WM) {
          const int m = mt * 8 + g8;
          const bool ml = m < a.M;
          double e0[TNG], e1[TNG];
#pragma unroll
          for (int nt = 0; nt < TNG; ++nt) { e0[nt] = 0.0; e1[nt] = 0.0; }
          for (int kk = 0; kk < ksteps; ++kk) {
            const int j = kk * 4 + t4;
            const bool jl = j < a.D_in;
            const double avv = (jl && ml) ? a.Zs[(long)m * a.D_in + j] : 0.0;
#pragma unroll
            for (int nt = 0; nt < TNG; ++nt) {
              const double bvv = jl ? xs[j * GC + ((nt * 8 + g8) ^ (XSW * (t4 << 3)))] : 0.0;
              dmma884(e0[nt], e1[nt], avv, bvv);
            }
          }
          const double zzm = a.zz[m];
#pragma unroll
          for (int nt = 0; nt < TNG; ++nt) {
            const int col = nt * 8 + 2 * t4;
            const double ka = kernel_value(KIND, fmax(fma(-2.0, e0[nt], zzm + xx0[nt]), 0.0), s2);
            const double kb = kernel_value(KIND, fmax(fma(-2.0, e1[nt], zzm + xx1[nt]), 0.0), s2);
            *reinterpret_cast<double2*>(tile + m * LDT + col0 + col) =
                make_double2((ml && p0 + col < a.P) ? ka : 0.0, (ml && p0 + col + 1 < a.P) ? kb : 0.0);
          }
        }
      };
      auto sweep_k = [&](auto kconst) {
        if (a.D_in <= 8) sweep(kconst, std::integral_constant<int, 2>{});
        else sweep_wide(kconst);
      };
      if (a.kind == 0) sweep_k(std::integral_constant<int, 0>{});
      else if (a.kind == 1) sweep_k(std::integral_constant<int, 1>{});
      else sweep_k(std::integral_constant<int, 2>{});
    }
    PH_ADD(ph_k2, ph_a2);
    group_sync(bar_id, GT);
    PH_ADD(ph_kuf, ph_a);
    PH_MARK(ph_b);
    // ---- stages 2-4: the operator panels: V pass, (A pass), then the T_d passes two outputs at a time ----
    double c0[TM][TN], c1[TM][TN], f0[TM][TN], f1[TM][TN];
    static_assert(TM <= 4, "panel_dispatch covers TM <= 4");
    // end of a row block: column sums of squares, in-place update / stash of the block's rows, stage totals
    auto block_end = [&](double (&x0)[TM][TN], double (&x1)[TM][TN], int kind, int d, int bi, int flags, double* pt, bool totals = true) {
      PH_MARK(ph_e);
      if (kind != 1) {   // column sums of squares of V / T_d: reduce over this warp's rows, accumulate in pt[wm][col]
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          double q0 = 0.0, q1 = 0.0;
#pragma unroll
          for (int i = 0; i < TM; ++i) { q0 = fma(x0[i][j], x0[i][j], q0); q1 = fma(x1[i][j], x1[i][j], q1); }
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            q0 += __shfl_xor_sync(0xffffffffu, q0, o);
            q1 += __shfl_xor_sync(0xffffffffu, q1, o);
          }
          if (g8 == 0) {
            const int col = j * 8 + 2 * t4;
            pt[wm * GC + col] += q0;
            pt[wm * GC + col + 1] += q1;
          }
        }
      }
      if (kind != 2) {   // in-place update of the resident tile (group-local hazard: same columns, all rows)
        group_sync(bar_id, GT);   // every warp of the group has finished reading the rows this block overwrites
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            const int row = bi * BM + i * 8 * WM + wm * 8 + g8, col = col0 + j * 8 + 2 * t4;
            *reinterpret_cast<double2*>(tile + row * LDT + col) = make_double2(x0[i][j], x1[i][j]);
          }
      }
      // stash of the resident operand the T_d passes read (A, or V in V-form) and of T_d itself
      double* st = kind == (a.vform ? 0 : 1) ? a.stashA : (kind == 2 && a.stashT ? a.stashT + (long)d * a.Mp * a.Pp : nullptr);
      if (st) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            const int row = bi * BM + i * 8 * WM + wm * 8 + g8, col = j * 8 + 2 * t4;
            __stcs(reinterpret_cast<double2*>(st + (long)row * a.Pp + p0 + col), make_double2(x0[i][j], x1[i][j]));   // streaming: read back only by the adjoint
          }
      }
      if (kind != 2) group_sync(bar_id, GT);   // the new rows are visible before the next block reads them
      if (totals && (flags & kPanelStageEnd) && kind != 1) {
        // sum the WM per-warp partials in a fixed order, then clear them for the next stage
        group_sync(bar_id, GT);
        if (tg < GC) {
          double sacc = 0.0;
#pragma unroll
          for (int w = 0; w < WM; ++w) { sacc += pt[w * GC + tg]; pt[w * GC + tg] = 0.0; }
          colsum[(kind == 0 ? 0 : 1 + d) * GC + tg] = sacc;
        }
        group_sync(bar_id, GT);
      }
      PH_ADD(ph_end, ph_e);
    };
    // stage totals of a pair of T_d passes under one pair of group barriers
    auto pair_totals = [&](int d) {
      PH_MARK(ph_e);
      group_sync(bar_id, GT);
      if (tg < GC) {
        double sa = 0.0, sb = 0.0;
#pragma unroll
        for (int w = 0; w < WM; ++w) {
          sa += part[w * GC + tg]; part[w * GC + tg] = 0.0;
          sb += part2[w * GC + tg]; part2[w * GC + tg] = 0.0;
        }
        colsum[(1 + d) * GC + tg] = sa;
        colsum[(2 + d) * GC + tg] = sb;
      }
      group_sync(bar_id, GT);
      PH_ADD(ph_end, ph_e);
    };
    // one pass of single panels: descriptors ptab[base .. base + count)
    auto run_single = [&](int kind, int d, int base, int count) {
      for (int qq = 0; qq < count; ++qq) {
        const uint2 tq = ptab[base + qq];
        const int k0 = tq.x & 0xfff, bi = (tq.x >> 12) & 63, flags = (tq.x >> 18) & 15;
        const int imin = (tq.y >> (6 * wm)) & 7, imax = (tq.y >> (6 * wm + 3)) & 7;   // this warp's active m-tiles [imin, imax)
        if (flags & kPanelFirst) {
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { c0[i][j] = 0.0; c1[i][j] = 0.0; }
        }
        const double* bt = tile + (k0 + t4) * LDT + col0 + g8;
        const double* pan = pbuf + cst * PANEL;
#ifdef DGP_DEBUG_WAITCLK
        const long long t_w0 = clock64();
#endif
        PH_MARK(ph_w);
        mbar_wait(full + cst, cph);           // the panel's bytes have landed
        PH_ADD(ph_wait, ph_w);
#ifdef DGP_DEBUG_WAITCLK
        wait_clk += clock64() - t_w0;
#endif
        panel_dispatch<TM, TN, WM, LDT>(c0, c1, pan, bt, wm, g8, t4, imin, imax);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + cst);   // this warp is done with the stage
        if (++cst == STAGES) { cst = 0; cph ^= 1; }
        if (flags & kPanelLast) block_end(c0, c1, kind, d, bi, flags, part);
      }
    };
    // one pass over the upper-operator descriptors with TWO panels per iteration: outputs d and d + 1 (stream order d, d + 1)
    auto run_pair = [&](int d) {
      for (int qq = 0; qq < NPa; ++qq) {
        const uint2 tq = ptab[NPv + qq];
        const int k0 = tq.x & 0xfff, bi = (tq.x >> 12) & 63, flags = (tq.x >> 18) & 15;
        const int imax = (tq.y >> (6 * wm + 3)) & 7;
        if (flags & kPanelFirst) {
#pragma unroll
          for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { c0[i][j] = 0.0; c1[i][j] = 0.0; f0[i][j] = 0.0; f1[i][j] = 0.0; }
        }
        const double* bt = tile + (k0 + t4) * LDT + col0 + g8;
        const int st1 = cst + 1 == STAGES ? 0 : cst + 1;
        const unsigned ph1 = cst + 1 == STAGES ? cph ^ 1 : cph;
        PH_MARK(ph_w);
        mbar_wait(full + cst, cph);
        mbar_wait(full + st1, ph1);
        PH_ADD(ph_wait, ph_w);
        panel_dispatch2<TM, TN, WM, LDT>(c0, c1, f0, f1, pbuf + cst * PANEL, pbuf + st1 * PANEL, bt, wm, g8, t4, imax);
        __syncwarp();
        if (lane == 0) { mbar_arrive(empty + cst); mbar_arrive(empty + st1); }
        cst = st1; cph = ph1;
        if (++cst == STAGES) { cst = 0; cph ^= 1; }
        if (flags & kPanelLast) {
          block_end(c0, c1, 2, d, bi, flags, part, false);
          block_end(f0, f1, 2, d + 1, bi, flags, part2, false);
          if (flags & kPanelStageEnd) pair_totals(d);
        }
      }
    };
    run_single(0, 0, 0, NPv);
    if (!a.vform) run_single(1, 0, NPv, NPa);
    for (int dp = 0; dp < (a.D_out >> 1); ++dp) run_pair(2 * dp);
    if (a.D_out & 1) run_single(2, a.D_out - 1, NPv, NPa);
    PH_ADD(ph_loop, ph_b);
    PH_MARK(ph_c);
    // ---- stage 5: moments, sample, outputs ----
    // mean^T [D_out x GC] = weights^T [D_out x M] * tile [M x GC] on the tensor pipe: one 8 x 8 output unit per (warp, step), the 64+
    // k-steps spread over four independent accumulator pairs (a dependent DMMA chain would run at 1/26 of the issue rate)
    {
#ifdef DGP_DEBUG_PHASECLK
      const long long ph_c2 = clock64();
#endif
      const int mts = (a.D_out + 7) >> 3;
      for (int u0 = wm; u0 < mts * (GC / 8); u0 += WM) {
        const int mt = u0 / (GC / 8), nt = u0 % (GC / 8);
        const int d = mt * 8 + g8;
        const bool dlive = d < a.D_out;
        const double* qa = a.qmu + (dlive ? d : 0);
        const double* tb = tile + t4 * LDT + col0 + nt * 8 + g8;
        double e0[4] = {0.0, 0.0, 0.0, 0.0}, e1[4] = {0.0, 0.0, 0.0, 0.0};
        for (int k0 = 0; k0 < a.Mp; k0 += 32) {
          double av[8], bv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int k = k0 + u * 4 + t4;
            av[u] = (dlive && k < a.M) ? qa[(long)k * a.qmu_ld] : 0.0;
            bv[u] = tb[(k0 + u * 4) * LDT];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) dmma884(e0[u & 3], e1[u & 3], av[u], bv[u]);
        }
        if (dlive) {
          const int col = nt * 8 + 2 * t4;
          meanbuf[col * a.D_out + d] = (e0[0] + e0[1]) + (e0[2] + e0[3]);
          meanbuf[(col + 1) * a.D_out + d] = (e1[0] + e1[1]) + (e1[2] + e1[3]);
        }
      }
      group_sync(bar_id, GT);
#ifdef DGP_DEBUG_PHASECLK
      ph_dot += clock64() - ph_c2;
#endif
    }
    for (int idx = tg; idx < GC * a.D_out; idx += GT) {
      const int c = idx / a.D_out, d = idx % a.D_out;
      const long p = p0 + c;
      if (p >= a.P) continue;
      const double mean = meanbuf[idx];
      const double* x = a.Xin + (p % a.xmod) * a.D_in;
      double mf = 0.0;
      if (a.mean_kind == 1) mf = x[d];
      else if (a.mean_kind == 2) {
        for (int j = 0; j < a.D_in; ++j) mf = fma(x[j], a.mfW[j * a.D_out + d], mf);
        if (a.mfb) mf += a.mfb[d];
      }
      const double mu = mean + mf;
      const double var = s2 - colsum[c] + colsum[(1 + d) * GC + c];
      const long s = p / a.Nc, n = p % a.Nc;
      const long xrow = s * a.N_total + a.n0 + n;
      a.Fmean[p * a.D_out + d] = mu;
      a.Fvar[p * a.D_out + d] = var;
      if (a.xFmean) a.xFmean[xrow * a.D_out + d] = mu;
      if (a.xFvar) a.xFvar[xrow * a.D_out + d] = var;
      if (a.F) {
        const double z = a.z_in ? a.z_in[xrow * a.D_out + d]
                                : philox_normal(a.seed_ptr ? *a.seed_ptr : a.seed, (uint32_t)a.layer, (uint32_t)s, (uint32_t)(a.n0 + n + a.n_offset), (uint32_t)d);
        if (a.z) a.z[p * a.D_out + d] = z;
        const double f = mu + z * sqrt(var + a.jitter);
        a.F[p * a.D_out + d] = f;
        if (a.xF) a.xF[xrow * a.D_out + d] = f;
      }
    }
    PH_ADD(ph_epi, ph_c);
  }
#ifdef DGP_DEBUG_PHASECLK
  if (lane == 0 && blockIdx.x == 1 && my_tiles > 10 && (warp == 0 || warp == 5)) {
    const double tot = (double)(clock64() - ph_t0);
    printf("fused_fwd cta %d warp %d (D_out %d, tiles %d): total %.0f clk/tile | kuf %.1f%% loop %.1f%% (of which block-ends %.1f%%, panel waits %.1f%%) epilogue %.1f%% (dot products %.1f%%) | kuf split: first sync %.1f%% inputs %.1f%% compute %.1f%%\n",
           blockIdx.x, warp, a.D_out, my_tiles, tot / my_tiles, 100.0 * ph_kuf / tot, 100.0 * ph_loop / tot, 100.0 * ph_end / tot, 100.0 * ph_wait / tot, 100.0 * ph_epi / tot, 100.0 * ph_dot / tot, 100.0 * ph_k0 / tot, 100.0 * ph_k1 / tot, 100.0 * ph_k2 / tot);
  }
#endif
#ifdef DGP_DEBUG_WAITCLK
  if (lane == 0 && blockIdx.x < 2 && my_tiles > 10) printf("cta %d warp %d: waited %.1f%% of %lld clk for operator panels\n", blockIdx.x, warp, 100.0 * wait_clk / (double)(clock64() - t_k0), clock64() - t_k0);
#endif
}

}  // namespace dgp
