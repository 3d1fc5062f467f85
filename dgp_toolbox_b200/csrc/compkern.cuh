// Composite multi-fidelity kernel of the reference's MF / MO models (MF_DGP.py:262-290, MO_DGP.py likewise) and its adjoint:
//   k(x, y) = k_corr(x_a, y_a) * (k_prev(x_b, y_b) + s_l^2 <x_b, y_b>) + k_in(x_a, y_a)      (+ s_w^2 on the diagonal of K(X, X) / K_diag)
// a = the first Da columns (the model input), b = the remaining columns (the previous fidelity's output, augmented by the
// model: MF_DGP.py:125). k_corr, k_prev: isotropic SquaredExponential (GPflow RBF(active_dims=..., variance=1.0) has one
// lengthscale); k_in: SquaredExponential, ARD or isotropic. has_prod = 0 leaves k_in (+ White) alone: the first fidelity's kernel.
// GPflow semantics: RBF = s^2 exp(-r^2/2) on active_dims, Linear = s^2 <x, y> on active_dims, White = s^2 I for K(X) / K_diag only.
#pragma once
#include "common.cuh"

namespace dgp {

struct CompK {
  int D, Da, has_prod, has_linear, in_ard;
  const double *in_var, *in_ls, *corr_var, *corr_ls, *prev_var, *prev_ls, *lin_var, *white_var;
};
constexpr int kCompTheta = 7;   // in_var, corr_var, corr_ls, prev_var, prev_ls, lin_var, white_var; then in_ls[Da]

struct CompParts { double ki, kc, kp, kl, r2a, r2b, dot; };

__device__ __forceinline__ double compk_eval(const CompK& k, const double* __restrict__ x, const double* __restrict__ y, CompParts& q) {
  double ri = 0.0, r2a = 0.0;
  for (int j = 0; j < k.Da; ++j) {
    const double t = x[j] - y[j];
    const double il = 1.0 / k.in_ls[k.in_ard ? j : 0];
    ri = fma(t * il, t * il, ri);
    r2a = fma(t, t, r2a);
  }
  q.ki = k.in_var[0] * exp(-0.5 * ri);
  q.r2a = r2a; q.r2b = 0.0; q.dot = 0.0; q.kc = 0.0; q.kp = 0.0; q.kl = 0.0;
  if (!k.has_prod) return q.ki;
  for (int j = k.Da; j < k.D; ++j) {
    const double t = x[j] - y[j];
    q.r2b = fma(t, t, q.r2b);
    q.dot = fma(x[j], y[j], q.dot);
  }
  const double lc = k.corr_ls[0], lp = k.prev_ls[0];
  q.kc = k.corr_var[0] * exp(-0.5 * r2a / (lc * lc));
  q.kp = k.prev_var[0] * exp(-0.5 * q.r2b / (lp * lp));
  q.kl = k.has_linear ? k.lin_var[0] * q.dot : 0.0;
  return q.kc * (q.kp + q.kl) + q.ki;
}

// K [P][P2]; X2 == null: K(X, X) with the White variance on the diagonal
__global__ void compk_K_kernel(CompK k, const double* __restrict__ X, long P, const double* __restrict__ X2, long P2, double* __restrict__ K) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * P2) return;
  const long i = idx / P2, j = idx % P2;
  CompParts q;
  double v = compk_eval(k, X + i * k.D, (X2 ? X2 : X) + j * k.D, q);
  if (!X2 && i == j && k.white_var) v += k.white_var[0];
  K[idx] = v;
}

__global__ void compk_Kdiag_kernel(CompK k, const double* __restrict__ X, long P, double* __restrict__ out) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double v = k.in_var[0] + (k.white_var ? k.white_var[0] : 0.0);
  if (k.has_prod) {
    double n2 = 0.0;
    for (int j = k.Da; j < k.D; ++j) n2 = fma(X[p * k.D + j], X[p * k.D + j], n2);
    v += k.corr_var[0] * (k.prev_var[0] + (k.has_linear ? k.lin_var[0] * n2 : 0.0));
  }
  out[p] = v;
}

// accumulate d k(x, y) / d x (gx, may be null), / d y (gy, may be null) and / d theta (gt, may be null), each times `w`
__device__ __forceinline__ void compk_accumulate(const CompK& k, const double* __restrict__ x, const double* __restrict__ y,
                                                 const CompParts& q, double w, double* gx, double* gy, double* gt) {
  const double sum = q.kp + q.kl;
  for (int j = 0; j < k.Da; ++j) {
    const double t = x[j] - y[j];
    const double il = 1.0 / k.in_ls[k.in_ard ? j : 0];
    double g = -q.ki * t * il * il;
    if (k.has_prod) g -= sum * q.kc * t / (k.corr_ls[0] * k.corr_ls[0]);
    if (gx) gx[j] = fma(w, g, gx[j]);
    if (gy) gy[j] = fma(-w, g, gy[j]);
    if (gt) gt[kCompTheta + (k.in_ard ? j : 0)] = fma(w, q.ki * t * t * il * il * il, gt[kCompTheta + (k.in_ard ? j : 0)]);
  }
  if (gt) gt[0] = fma(w, q.ki / k.in_var[0], gt[0]);
  if (!k.has_prod) return;
  const double lp2 = k.prev_ls[0] * k.prev_ls[0];
  for (int j = k.Da; j < k.D; ++j) {
    const double t = x[j] - y[j];
    const double sl = k.has_linear ? k.lin_var[0] : 0.0;
    if (gx) gx[j] = fma(w, q.kc * (-q.kp * t / lp2 + sl * y[j]), gx[j]);
    if (gy) gy[j] = fma(w, q.kc * (q.kp * t / lp2 + sl * x[j]), gy[j]);
  }
  if (gt) {
    const double lc = k.corr_ls[0];
    gt[1] = fma(w, q.kc * sum / k.corr_var[0], gt[1]);
    gt[2] = fma(w, q.kc * sum * q.r2a / (lc * lc * lc), gt[2]);
    gt[3] = fma(w, q.kc * q.kp / k.prev_var[0], gt[3]);
    gt[4] = fma(w, q.kc * q.kp * q.r2b / (lp2 * k.prev_ls[0]), gt[4]);
    if (k.has_linear) gt[5] = fma(w, q.kc * q.dot, gt[5]);
  }
}

constexpr int kCompMaxD = 32;

// One block per row i of X: dX[i][:] = sum_j Kbar[i][j] dk(x_i, y_j)/dx_i (+ sum_j Kbar[j][i] dk(x_j, x_i)/dx_i when X2 == null),
// part[i][:] = this row's share of d/d theta (reduced by reduce_partials_kernel).
__global__ void __launch_bounds__(128) compk_grad_rows_kernel(CompK k, const double* __restrict__ X, long P, const double* __restrict__ X2,
                                                              long P2, const double* __restrict__ Kbar, double* __restrict__ dX,
                                                              double* __restrict__ part) {
  __shared__ double red[32];
  const long i = blockIdx.x;
  const int nt = kCompTheta + (k.in_ard ? k.Da : 1);
  double gx[kCompMaxD], gt[kCompTheta + kCompMaxD];
  for (int j = 0; j < k.D; ++j) gx[j] = 0.0;
  for (int j = 0; j < nt; ++j) gt[j] = 0.0;
  const double* xi = X + i * k.D;
  const double* Y = X2 ? X2 : X;
  for (long j = threadIdx.x; j < P2; j += blockDim.x) {
    CompParts q;
    compk_eval(k, xi, Y + j * k.D, q);
    const double w = Kbar[i * P2 + j];
    compk_accumulate(k, xi, Y + j * k.D, q, w, gx, nullptr, gt);
    if (!X2) {
      compk_accumulate(k, xi, Y + j * k.D, q, Kbar[j * P2 + i], gx, nullptr, nullptr);   // k is symmetric: d k(x_j, x_i)/d x_i = d k(x_i, x_j)/d x_i
      if (i == j && k.white_var) gt[6] += w;
    }
  }
  for (int j = 0; j < k.D; ++j) {
    const double r = block_sum(gx[j], red);
    if (threadIdx.x == 0) dX[i * k.D + j] = r;
  }
  for (int j = 0; j < nt; ++j) {
    const double r = block_sum(gt[j], red);
    if (threadIdx.x == 0) part[i * nt + j] = r;
  }
}

// One thread per row j of X2: dX2[j][:] = sum_i Kbar[i][j] dk(x_i, y_j)/dy_j
__global__ void __launch_bounds__(128) compk_grad_cols_kernel(CompK k, const double* __restrict__ X, long P, const double* __restrict__ X2,
                                                              long P2, const double* __restrict__ Kbar, double* __restrict__ dX2) {
  const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= P2) return;
  double gy[kCompMaxD], y[kCompMaxD];
  for (int d = 0; d < k.D; ++d) { gy[d] = 0.0; y[d] = X2[j * k.D + d]; }
  for (long i = 0; i < P; ++i) {
    CompParts q;
    compk_eval(k, X + i * k.D, y, q);
    compk_accumulate(k, X + i * k.D, y, q, Kbar[i * P2 + j], nullptr, gy, nullptr);
  }
  for (int d = 0; d < k.D; ++d) dX2[j * k.D + d] = gy[d];
}

// K_diag adjoint: dX[p][b] = g[p] * 2 s_c^2 s_l^2 x_pb; per-block partial sums of d/d theta
__global__ void __launch_bounds__(128) compk_Kdiag_grad_kernel(CompK k, const double* __restrict__ X, long P, const double* __restrict__ g,
                                                               double* __restrict__ dX, double* __restrict__ part) {
  __shared__ double red[32];
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double t[kCompTheta];
  for (int j = 0; j < kCompTheta; ++j) t[j] = 0.0;
  if (p < P) {
    const double w = g[p];
    for (int j = 0; j < k.Da; ++j) dX[p * k.D + j] = 0.0;
    double n2 = 0.0;
    for (int j = k.Da; j < k.D; ++j) {
      const double x = X[p * k.D + j];
      n2 = fma(x, x, n2);
      dX[p * k.D + j] = (k.has_prod && k.has_linear) ? w * 2.0 * k.corr_var[0] * k.lin_var[0] * x : 0.0;
    }
    t[0] = w;
    t[6] = k.white_var ? w : 0.0;
    if (k.has_prod) {
      const double sl = k.has_linear ? k.lin_var[0] : 0.0;
      t[1] = w * (k.prev_var[0] + sl * n2);
      t[3] = w * k.corr_var[0];
      t[5] = k.has_linear ? w * k.corr_var[0] * n2 : 0.0;
    }
  }
  const int nt = kCompTheta + (k.in_ard ? k.Da : 1);
  for (int j = 0; j < nt; ++j) {
    const double r = block_sum(j < kCompTheta ? t[j] : 0.0, red);
    if (threadIdx.x == 0) part[(long)blockIdx.x * nt + j] = r;
  }
}

}  // namespace dgp
