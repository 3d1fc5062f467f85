// SVGP-layer conditional and its adjoint on CALLER-SUPPLIED kernel matrices (Kuu + jitter, Kuf, K_diag): the contraction part of
// utils/layers.py:237-308 for layers whose kernel is not one of the built-in stationary ones -- the multi-fidelity composite
// k_corr (k_prev + Linear) + k_in + White of MF_DGP.py:262-290 (compkern.cuh evaluates it and its adjoint). Small streaming
// kernels only; the M^2 P work runs in the DMMA GEMM engine.
#pragma once
#include "common.cuh"

namespace dgp {

// out [Mp][Mp] = in [M][M] on the leading block, identity on the padding
__global__ void pad_square_identity_kernel(const double* __restrict__ in, int M, int Mp, double* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Mp) return;
  const int i = (int)(idx / Mp), j = (int)(idx % Mp);
  out[idx] = (i < M && j < M) ? in[(long)i * M + j] : (i == j ? 1.0 : 0.0);
}

// out [Mp][Pp] = in [M][P], zero padded
__global__ void pad_plane_kernel(const double* __restrict__ in, int M, int Mp, long P, long Pp, double* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Pp) return;
  const int m = (int)(idx / Pp);
  const long p = idx % Pp;
  out[idx] = (m < M && p < P) ? in[(long)m * P + p] : 0.0;
}

// mean[p][d] = sum_m A[m][p] q_mu[m][d];  var[p][d] = Kdiag[p] - sum_m V[m][p]^2 + sum_m T_d[m][p]^2   (white: A = V)
__global__ void __launch_bounds__(128) moments_ext_kernel(const double* __restrict__ V, const double* __restrict__ A,
                                                          const double* __restrict__ T, const double* __restrict__ qmuP,
                                                          const double* __restrict__ kdiag, int M, int Mp, int D, long P, long Pp,
                                                          double* __restrict__ mean, double* __restrict__ var) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long plane = (long)Mp * Pp;
  double v2 = 0.0;
  for (int m = 0; m < M; ++m) { const double v = V[(long)m * Pp + p]; v2 = fma(v, v, v2); }
  for (int d = 0; d < D; ++d) {
    double mu = 0.0, del = 0.0;
    for (int m = 0; m < M; ++m) {
      const long off = (long)m * Pp + p;
      const double t = T[(long)d * plane + off];
      mu = fma(A[off], qmuP[m * 32 + d], mu);
      del = fma(t, t, del);
    }
    mean[p * D + d] = mu;
    var[p * D + d] = kdiag[p] - v2 + del;
  }
}

// caller's upstream gradients [P][D] -> the layouts the adjoint contractions read
__global__ void __launch_bounds__(128) upstream_ext_kernel(const double* __restrict__ Gm_in, const double* __restrict__ Gv_in, long P, long Pp,
                                                           int D, double* __restrict__ GvT, double* __restrict__ GmPad,
                                                           double* __restrict__ gq, double* __restrict__ dKdiag) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Pp) return;
  double s = 0.0;
  for (int d = 0; d < D; ++d) {
    const double gm = p < P ? Gm_in[p * D + d] : 0.0, gv = p < P ? Gv_in[p * D + d] : 0.0;
    GvT[(long)d * Pp + p] = gv;
    GmPad[p * 32 + d] = gm;
    s += gv;
  }
  for (int d = D; d < 32; ++d) GmPad[p * 32 + d] = 0.0;
  gq[p] = -s;
  if (p < P) dKdiag[p] = s;      // d var / d Kdiag = 1 for every output
}

// Kbar_uf = W + 2 A diag(gq) (caller layout [M][P]);  W <- Wg = W + A diag(gq) in place (feeds dKu = -Wg A^T)
__global__ void kbar_ext_kernel(double* __restrict__ W, const double* __restrict__ A, const double* __restrict__ gq, int M, long P,
                                long Pp, double* __restrict__ dKuf) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * Pp) return;
  const int m = (int)(idx / Pp);
  const long p = idx % Pp;
  const double w = W[idx], a = A[idx], g = gq[p];
  W[idx] = w + a * g;
  if (p < P) dKuf[(long)m * P + p] = w + 2.0 * a * g;
}

}  // namespace dgp
