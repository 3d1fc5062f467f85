// M^2 / M^3-class replicated kernels: parameter padding, KL (utils/layers.py:280-308), its adjoint, the RBF adjoint
// on Kuu and the final gradient assembly.
#pragma once
#include "common.cuh"

namespace dgp {

// RpT[d][i][j] = q_sqrt[d][j][i] (upper triangular, A-operand of T_d = R_d^T A); Rcat[i][d*Mp + j] = q_sqrt[d][i][j]
// (lower-triangular blocks side by side, A-operand of the K-concatenated products); qmuP[Mp][32]. Zero on the padding.
__global__ void pad_params_kernel(const double* __restrict__ q_sqrt, const double* __restrict__ q_mu, int M, int Mp, int D,
                                  double* __restrict__ RpT, double* __restrict__ Rcat, double* __restrict__ qmuP) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long nR = (long)D * Mp * Mp;
  if (idx < nR) {
    int j = (int)(idx % Mp);
    long r = idx / Mp;
    int i = (int)(r % Mp), d = (int)(r / Mp);
    RpT[idx] = (j < M && i <= j) ? q_sqrt[((long)d * M + j) * M + i] : 0.0;
    Rcat[(long)i * D * Mp + (long)d * Mp + j] = (i < M && j <= i) ? q_sqrt[((long)d * M + i) * M + j] : 0.0;
  }
  if (idx < (long)Mp * 32) {
    int d = (int)(idx % 32), m = (int)(idx / 32);
    qmuP[idx] = (m < M && d < D) ? q_mu[(long)m * D + d] : 0.0;
  }
}

// KL = -0.5 D M - sum log|R_ii| + D sum log L_ii + 0.5 |Linv R|_F^2 + 0.5 sum q_mu o alpha     (non-white, layers.py:280-308)
// LR = Linv Rcat [Mp][D*Mp]; alpha = Kinv q_mu [Mp][32]. Single block, fixed reduction tree.
__global__ void __launch_bounds__(1024) kl_kernel(const double* __restrict__ L, const double* __restrict__ Rcat,
                                                  const double* __restrict__ LR, const double* __restrict__ qmuP,
                                                  const double* __restrict__ alpha, int M, int Mp, int D, double* __restrict__ kl) {
  __shared__ double red[32];
  double s = 0.0;
  const long nLR = (long)D * Mp * Mp;
  for (long i = threadIdx.x; i < nLR; i += blockDim.x) { double v = LR[i]; s = fma(0.5 * v, v, s); }
  for (long i = threadIdx.x; i < (long)Mp * 32; i += blockDim.x) s = fma(0.5 * qmuP[i], alpha[i], s);
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    s += (double)D * log(L[(long)i * Mp + i]);
    for (int d = 0; d < D; ++d) { double r = Rcat[(long)i * D * Mp + (long)d * Mp + i]; s -= 0.5 * log(r * r); }
  }
  double r = block_sum(s, red);
  if (threadIdx.x == 0) kl[0] = r - 0.5 * (double)D * (double)M;
}

// Whitened layer: KL(N(q_mu, R R^T) || N(0, I)) = -0.5 D M - sum log|R_ii| + 0.5 |R|_F^2 + 0.5 |q_mu|_F^2   (layers.py:296-303)
__global__ void __launch_bounds__(1024) kl_white_kernel(const double* __restrict__ Rcat, const double* __restrict__ qmuP, int M, int Mp,
                                                        int D, double* __restrict__ kl) {
  __shared__ double red[32];
  double s = 0.0;
  const long nR = (long)D * Mp * Mp;
  for (long i = threadIdx.x; i < nR; i += blockDim.x) { double v = Rcat[i]; s = fma(0.5 * v, v, s); }
  for (long i = threadIdx.x; i < (long)Mp * 32; i += blockDim.x) s = fma(0.5 * qmuP[i], qmuP[i], s);
  for (int i = threadIdx.x; i < M; i += blockDim.x)
    for (int d = 0; d < D; ++d) { double r = Rcat[(long)i * D * Mp + (long)d * Mp + i]; s -= 0.5 * log(r * r); }
  double r = block_sum(s, red);
  if (threadIdx.x == 0) kl[0] = r - 0.5 * (double)D * (double)M;
}

// dKu = dKu_data - klw * (0.5 D Kinv - 0.5 Kinv Ssum Kinv - 0.5 alpha alpha^T), zero on the padding (in place)
__global__ void dku_assemble_kernel(double* __restrict__ dKu, const double* __restrict__ Kinv, const double* __restrict__ KSK,
                                    const double* __restrict__ alpha, const double* __restrict__ Knj, int M, int Mp, int D,
                                    double klw, int have_data) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Mp) return;
  int i = (int)(idx / Mp), j = (int)(idx % Mp);
  double v = 0.0;
  if (i < M && j < M) {
    double aa = 0.0;
    if (klw != 0.0)   // klw == 0: a whitened layer, whose KL does not depend on Ku (Kinv / KSK / alpha are not computed)
      for (int d = 0; d < D; ++d) aa = fma(alpha[i * 32 + d], alpha[j * 32 + d], aa);
    v = (have_data ? dKu[idx] : 0.0);
    if (klw != 0.0) v -= klw * (0.5 * D * Kinv[idx] - 0.5 * KSK[idx] - 0.5 * aa);
  }
  dKu[idx] = v;
}

// Kernel adjoint on Kuu, one block per inducing row a, with Kbar = dKu, k and g = -2 dk/d(r2) re-evaluated per pair:
//   dZ[a][j] = -sum_b (Kbar[a][b] + Kbar[b][a]) g_ab (z_a - z_b)_j / l_j^2 ; part[a][j] = sum_b Kbar[a][b] g_ab (z_a - z_b)_j^2 / l_j^3 ;
//   part[a][D] = sum_b Kbar[a][b] k_ab / s2
__global__ void __launch_bounds__(128) kuu_bwd_kernel(const double* __restrict__ G, const double* __restrict__ Z,
                                                      const double* __restrict__ ls, const double* __restrict__ var, int M, int Mp,
                                                      int D, double* __restrict__ dZk, double* __restrict__ part, int kind) {
  __shared__ double red[32];
  const int a = blockIdx.x;
  double accz[kMaxD], accl[kMaxD], accs = 0.0;
  for (int j = 0; j < D; ++j) { accz[j] = 0.0; accl[j] = 0.0; }
  const double s2 = var[0];
  for (int b = threadIdx.x; b < M; b += blockDim.x) {
    double r2 = 0.0;
    for (int j = 0; j < D; ++j) {
      const double t = (Z[(long)a * D + j] - Z[(long)b * D + j]) / ls[j];
      r2 = fma(t, t, r2);
    }
    double k, gf;
    kernel_eval(kind, r2, s2, k, gf);
    const double kab = G[(long)a * Mp + b], kba = G[(long)b * Mp + a];
    const double gab = kab * gf, gba = kba * gf;
    accs += kab * k;
    for (int j = 0; j < D; ++j) {
      const double il = 1.0 / ls[j];
      const double t = (Z[(long)a * D + j] - Z[(long)b * D + j]) * il;
      accz[j] -= (gab + gba) * t * il;
      accl[j] += gab * t * t * il;
    }
  }
  for (int j = 0; j < D; ++j) {
    double rz = block_sum(accz[j], red);
    double rl = block_sum(accl[j], red);
    if (threadIdx.x == 0) { dZk[(long)a * D + j] = rz; part[(long)a * (D + 1) + j] = rl; }
  }
  double rs = block_sum(accs / s2, red);
  if (threadIdx.x == 0) part[(long)a * (D + 1) + D] = rs;
}

struct FinalizeArgs {
  // inputs
  const double* Gd;       // [D_out][Mp][Mp]  tril(A dT_d^T), or [Mp][D_out*Mp] when gd_cat (null when no data term)
  int gd_cat;
  const double* KR;       // [Mp][D_out*Mp]   Kinv Rcat
  const double* Rcat;     // [Mp][D_out*Mp]
  const double* dqmu;     // [Mp][32] data part (null when no data term)
  const double* alpha;    // [Mp][32]
  const double* H;        // [Mp][32] Gbar [X,1]  (null when no data term)
  const double* dZk;      // [M][D_in] Kuu part
  const double* rbf_red;  // [D_in+1] data-part partial sums (dl_j..., ds2)  (null when no data term)
  const double* kuu_red;  // [D_in+1]
  const double* sgv;      // [3] -> [2] = sum Gv (K_diag term)   (null when no data term)
  const double* Z; const double* ls;
  int M, Mp, D_in, D_out;
  double klw;
  int white; const double* qmuP;   // whitened layer: d KL / d q_sqrt = R - diag(1 / R_ii), d KL / d q_mu = q_mu
  // outputs (unpadded, inside the flat gradient buffer)
  double* dZ; double* dls; double* dvar; double* dq_mu; double* dq_sqrt;
};

__global__ void finalize_layer_kernel(FinalizeArgs a) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long nq = (long)a.D_out * a.M * a.M;
  if (idx < nq) {
    int j = (int)(idx % a.M);
    long r = idx / a.M;
    int i = (int)(r % a.M), d = (int)(r / a.M);
    double v = 0.0;
    if (j <= i) {
      const long off = ((long)d * a.Mp + i) * a.Mp + j;
      const long offc = (long)i * a.D_out * a.Mp + (long)d * a.Mp + j;
      v = (a.Gd ? a.Gd[a.gd_cat ? offc : off] : 0.0) - a.klw * ((a.white ? a.Rcat[offc] : a.KR[offc]) - (i == j ? 1.0 / a.Rcat[offc] : 0.0));
    }
    a.dq_sqrt[idx] = v;
  }
  if (idx < (long)a.M * a.D_out) {
    int d = (int)(idx % a.D_out), m = (int)(idx / a.D_out);
    a.dq_mu[idx] = (a.dqmu ? a.dqmu[m * 32 + d] : 0.0) - a.klw * (a.white ? a.qmuP[m * 32 + d] : a.alpha[m * 32 + d]);
  }
  if (idx < (long)a.M * a.D_in) {
    int j = (int)(idx % a.D_in), m = (int)(idx / a.D_in);
    double v = a.dZk[idx];
    if (a.H) {
      const double il = 1.0 / a.ls[j];
      v -= (a.Z[idx] * a.H[m * 32 + a.D_in] - a.H[m * 32 + j]) * il * il;
    }
    a.dZ[idx] = v;
  }
  if (idx < a.D_in) a.dls[idx] = (a.rbf_red ? a.rbf_red[idx] : 0.0) + a.kuu_red[idx];
  if (idx == 0) a.dvar[0] = (a.rbf_red ? a.rbf_red[a.D_in] : 0.0) + a.kuu_red[a.D_in] + (a.sgv ? a.sgv[2] : 0.0);
}

// ---- V-form adjoint glue (M^2-class, once per step and layer) ----
// CTcat[i][d*Mp + j] = Cmat[d][j][i]  (C_d^T, lower-triangular blocks side by side).
__global__ void vform_transpose_kernel(const double* __restrict__ Cmat, int Mp, int D, double* __restrict__ CTcat) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)D * Mp * Mp) return;
  const int j = (int)(idx % Mp);
  const long r = idx / Mp;
  const int i = (int)(r % Mp), d = (int)(r / Mp);
  CTcat[(long)i * D * Mp + (long)d * Mp + j] = Cmat[((long)d * Mp + j) * Mp + i];
}

// In place on `nblk` square blocks laid side by side in a [Mp][ld] matrix: keep the lower triangle (scaled), zero the rest.
__global__ void tril_scale_kernel(double* __restrict__ X, int Mp, long ld, int nblk, double scale) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nblk * Mp * Mp) return;
  const int j = (int)(idx % Mp);
  const long r = idx / Mp;
  const int i = (int)(r % Mp), b = (int)(r / Mp);
  double* p = X + (long)i * ld + (long)b * Mp + j;
  *p = (j <= i) ? scale * *p : 0.0;
}

// In place on `nblk` square blocks laid side by side in a [Mp][ld] matrix: mirror the lower triangle into the upper one.
__global__ void sym_fill_kernel(double* __restrict__ X, int Mp, long ld, int nblk) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)nblk * Mp * Mp) return;
  const int j = (int)(idx % Mp);
  const long r = idx / Mp;
  const int i = (int)(r % Mp), b = (int)(r / Mp);
  if (j > i) X[(long)i * ld + (long)b * Mp + j] = X[(long)j * ld + (long)b * Mp + i];
}

// G1 = tril(G1 - sum_d W_d), W_d side by side in a [Mp][nblk * Mp] matrix (V-form adjoint, see backward_layer).
__global__ void g1_finish_kernel(double* __restrict__ G1, const double* __restrict__ W, int Mp, int nblk) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Mp) return;
  const int i = (int)(idx / Mp), j = (int)(idx % Mp);
  double s = 0.0;
  if (j <= i) {
    for (int b = 0; b < nblk; ++b) s += W[(long)i * nblk * Mp + (long)b * Mp + j];
    s = G1[idx] - s;
  }
  G1[idx] = s;
}

// Cholesky adjoint core: out = sym(Phi(P)), Phi = lower triangle with halved diagonal, sym(Q) = (Q + Q^T) / 2.
__global__ void phi_sym_kernel(const double* __restrict__ P, int Mp, double* __restrict__ out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Mp) return;
  const int i = (int)(idx / Mp), j = (int)(idx % Mp);
  const int a = max(i, j), b = min(i, j);
  const double v = P[(long)a * Mp + b];
  out[idx] = 0.5 * v;   // off-diagonal: (Phi + Phi^T)/2 = P_lower / 2; diagonal: (P/2 + P/2) / 2 = P / 2
}

__global__ void scale_copy_kernel(const double* in, double s, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = s * in[0];
}

__global__ void scale_copy_diff_kernel(const double* flat, double* out) {   // ELBO estimate = data term - KL
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = flat[0] - flat[1];
}

// Copies the leading [M][M] block of a padded [Mp][Mp] matrix.
__global__ void unpad_square_kernel(const double* __restrict__ in, int M, int Mp, double* __restrict__ out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * M) return;
  int i = (int)(idx / M), j = (int)(idx % M);
  out[idx] = in[(long)i * Mp + j];
}

// ---------------------------------------------------------------------------------------------------------
// Adam on the unconstrained variables (models/dgp.py:132-154: tf.optimizers.Adam applied to the GPflow variables).
// The model keeps constrained values, so one thread per entry maps value -> unconstrained u through the parameter's
// bijector, applies the chain rule to the constrained-space ELBO gradient, updates (m, v, u) and writes bijector(u) back.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAdamMaxParams = 48;
struct AdamSeg {
  double* value; double* mirror;   // parameter [count]; optional broadcast copy [mirror_count] (scalar lengthscale -> [D_in])
  long start;                      // prefix offset into the (m, v) state
  long count, grad_offset, grad_count, mirror_count;
  int transform, M;                // 0 identity, 1 softplus, 2 softplus + 1e-6, 3 lower triangles of [count / M^2][M][M]
};
struct AdamTable { int n; long total; AdamSeg seg[kAdamMaxParams]; };

__global__ void __launch_bounds__(256) adam_kernel(AdamTable t, const double* __restrict__ grad, double* __restrict__ m_state,
                                                   double* __restrict__ v_state, double lr_t, double beta1, double beta2, double eps,
                                                   double* __restrict__ trace, const int* __restrict__ chol_failed) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && trace) *trace = grad[0] - grad[1];   // the step's ELBO estimate: data term - KL
  if (idx >= t.total) return;
  // a non-positive-definite Kuu makes this step's gradient NaN: keep the parameters at their last good values (the flag is sticky
  // until dgp_check reports it), as the reference does by raising at the failing step
  if (chol_failed && *chol_failed) return;
  int k = 0;
  while (k + 1 < t.n && idx >= t.seg[k + 1].start) ++k;
  const AdamSeg& sg = t.seg[k];
  const long i = idx - sg.start;
  if (sg.transform == 3) {
    const long r = (i / sg.M) % sg.M, col = i % sg.M;
    if (col > r) return;   // FillTriangular has no variable above the diagonal
  }
  double g = 0.0;   // d ELBO / d value
  if (sg.grad_count == sg.count) g = grad[sg.grad_offset + i];
  else for (long j = 0; j < sg.grad_count; ++j) g += grad[sg.grad_offset + j];   // one scalar shared by grad_count entries
  g = -g;           // the optimiser minimises -ELBO
  const double c = sg.value[i];
  double u = c;
  const double lower = sg.transform == 2 ? 1e-6 : 0.0;
  if (sg.transform == 1 || sg.transform == 2) {
    const double y = c - lower;
    const double sig = -expm1(-y);   // sigmoid(u) = 1 - exp(-softplus(u))
    u = y + log(sig);                // softplus^-1
    g *= sig;
  }
  const double m = beta1 * m_state[idx] + (1.0 - beta1) * g;
  const double v = beta2 * v_state[idx] + (1.0 - beta2) * g * g;
  m_state[idx] = m;
  v_state[idx] = v;
  u -= lr_t * m / (sqrt(v) + eps);
  double out = u;
  if (sg.transform == 1 || sg.transform == 2) out = (u > 0.0 ? u + log1p(exp(-u)) : log1p(exp(u))) + lower;
  sg.value[i] = out;
  if (sg.mirror)
    for (long j = 0; j < sg.mirror_count; ++j) sg.mirror[j] = out;
}


// ---------------------------------------------------------------------------------------------------------
// GPflow NaturalGradient(gamma) step on (q_mu, q_sqrt) pairs, XiNat parameterisation (models/dgp.py:188,218,312,343), in
// collapsed form: with R = q_sqrt_d, G_R / G_mu the gradients of -ELBO, T = tril(R^T G_R),
//   B = I + gamma (T + T^T - diag T) = U U^T (reverse Cholesky, U upper),  q_sqrt_new = R U^-T,  mu_new = mu - gamma C C^T G_mu.
// One "output" = one (layer, d) pair; all outputs of a call are batched. The reverse Cholesky is an ordinary one of the
// flipped matrix: with J the index reversal, J B J = Lf Lf^T, U = J Lf J, U^-T = J Lf^-T J.
// ---------------------------------------------------------------------------------------------------------
struct NatOut {              // one (layer, d) pair
  const double* q_sqrt;      // [M][M] block d of the layer's q_sqrt (lower)
  const double* g_sqrt;      // [M][M] d ELBO / d q_sqrt_d inside the flat gradient buffer
  const double* g_mu;        // d ELBO / d q_mu, [M][D] (column d used)
  double* q_mu;              // [M][D] updated in place (column d)
  double* q_sqrt_out;        // == q_sqrt, updated in place
  int M, Mp, D, d;
  long off;                  // offset of this output's [Mp][Mp] matrices inside the batched work arrays
};

// RT = R^T (upper, padded), R (lower, padded), GR = -tril(dELBO/dq_sqrt) (padded)
__global__ void natgrad_prep_kernel(const NatOut* __restrict__ outs, double* __restrict__ R, double* __restrict__ RT,
                                    double* __restrict__ GR) {
  const NatOut o = outs[blockIdx.y];
  const long mm = (long)o.Mp * o.Mp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < mm; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / o.Mp), j = (int)(idx % o.Mp);
    const bool in = i < o.M && j < o.M && j <= i;
    const double r = in ? o.q_sqrt[(long)i * o.M + j] : 0.0;
    R[o.off + idx] = r;
    RT[o.off + (long)j * o.Mp + i] = r;
    GR[o.off + idx] = in ? -o.g_sqrt[(long)i * o.M + j] : 0.0;
  }
}

// Bf = J (I + gamma (T_low + T_low^T - diag T)) J on the leading M x M block, identity on the padding
__global__ void natgrad_b_kernel(const NatOut* __restrict__ outs, const double* __restrict__ T, double gamma, double* __restrict__ Bf) {
  const NatOut o = outs[blockIdx.y];
  const long mm = (long)o.Mp * o.Mp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < mm; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / o.Mp), j = (int)(idx % o.Mp);
    double v = i == j ? 1.0 : 0.0;
    if (i < o.M && j < o.M) {
      const int a = o.M - 1 - i, b = o.M - 1 - j;          // un-flipped indices
      const int hi = a > b ? a : b, lo = a > b ? b : a;
      v += gamma * T[o.off + (long)hi * o.Mp + lo];         // lower triangle of T, mirrored; the diagonal once
    }
    Bf[o.off + idx] = v;
  }
}

// UinvT = J Lf^-T J (lower) on the leading block, zero elsewhere
__global__ void natgrad_flip_kernel(const NatOut* __restrict__ outs, const double* __restrict__ LfinvT, double* __restrict__ UinvT) {
  const NatOut o = outs[blockIdx.y];
  const long mm = (long)o.Mp * o.Mp;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < mm; idx += (long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / o.Mp), j = (int)(idx % o.Mp);
    // only the upper triangle of Lf^-T is defined (tri_inv_kernel leaves the rest of the plane untouched)
    UinvT[o.off + idx] = (i < o.M && j <= i) ? LfinvT[o.off + (long)(o.M - 1 - i) * o.Mp + (o.M - 1 - j)] : 0.0;
  }
}

// one CTA per output: w = C^T g (g = -dELBO/dq_mu[:, d]), q_mu[:, d] -= gamma C w, q_sqrt_d = tril(C)
__global__ void __launch_bounds__(256) natgrad_finalize_kernel(const NatOut* __restrict__ outs, const double* __restrict__ C, double gamma,
                                                               const int* __restrict__ failed) {
  extern __shared__ double nsh[];   // g [M], w [M]
  const NatOut o = outs[blockIdx.x];
  if (failed && *failed) return;    // a factorisation failed (here or in the ELBO): keep the last good values
  double* g = nsh;
  double* w = nsh + o.M;
  const double* Cm = C + o.off;
  for (int i = threadIdx.x; i < o.M; i += blockDim.x) g[i] = -o.g_mu[(long)i * o.D + o.d];
  __syncthreads();
  for (int j = threadIdx.x; j < o.M; j += blockDim.x) {      // w_j = sum_{i >= j} C[i][j] g_i   (C lower)
    double s = 0.0;
    for (int i = j; i < o.M; ++i) s = fma(Cm[(long)i * o.Mp + j], g[i], s);
    w[j] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < o.M; i += blockDim.x) {      // (C w)_i = sum_{j <= i} C[i][j] w_j
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s = fma(Cm[(long)i * o.Mp + j], w[j], s);
    o.q_mu[(long)i * o.D + o.d] -= gamma * s;
  }
  for (long idx = threadIdx.x; idx < (long)o.M * o.M; idx += blockDim.x) {
    const int i = (int)(idx / o.M), j = (int)(idx % o.M);
    o.q_sqrt_out[idx] = j <= i ? Cm[(long)i * o.Mp + j] : 0.0;
  }
}

}  // namespace dgp
