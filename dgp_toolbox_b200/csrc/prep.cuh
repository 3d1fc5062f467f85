// Per-layer, per-step replicated work: Kuu build, blocked Cholesky, explicit triangular inverse.
// One CTA per layer (all layers of the model in one launch); matrices live in L2, panels in shared memory.
// Reference semantics: dgp_dace/utils/layers.py:227-234 (Ku = Kuu + jitter I, Lu = chol(Ku)).
#pragma once
#include "common.cuh"

namespace dgp {

// Ku[i][j] = s2 * exp(-0.5 * sum_j ((z_i - z_j)/l)^2) + jitter * (i==j) for i,j < M; identity on the padding.
// Knj receives the same matrix without jitter (zero on the padding) for the RBF backward of Kuu.
__global__ void kuu_build_kernel(const double* __restrict__ Z, const double* __restrict__ ls, const double* __restrict__ var,
                                 int M, int Mp, int D, double jitter, double* __restrict__ Ku, double* __restrict__ Knj, int kind) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)Mp * Mp) return;
  int i = (int)(idx / Mp), j = (int)(idx % Mp);
  double k = 0.0, kj = 0.0;
  if (i < M && j < M) {
    double r2 = 0.0;
    for (int d = 0; d < D; ++d) {
      double t = (Z[(long)i * D + d] - Z[(long)j * D + d]) / ls[d];
      r2 = fma(t, t, r2);
    }
    k = kernel_value(kind, r2, var[0]);
    kj = k + (i == j ? jitter : 0.0);
  } else {
    kj = (i == j) ? 1.0 : 0.0;
  }
  Ku[idx] = kj;
  Knj[idx] = k;
}

struct CholArgs {
  const double* Ku;  // [Mp][Mp]
  double* L;         // [Mp][Mp] lower, zero above the diagonal
  double* Linv;      // [Mp][Mp] lower
  double* LinvT;     // [Mp][Mp] upper
  int Mp;
  int* info;         // set to 1 when a pivot is not positive
};

constexpr int kCholNB = 32;
constexpr int kCholThreads = 512;

// dynamic shared memory: D[32][33] + Dinv[32] + panel (transposed [32][Mp+4]; reused as the G block [32][Mp+1] of the inverse)
inline size_t chol_smem_bytes(int Mp) { return (size_t)(32 * 33 + 32 + (size_t)Mp * 33 + 128) * sizeof(double); }

__global__ void __launch_bounds__(kCholThreads) chol_inv_kernel(const CholArgs* __restrict__ args) {
  const CholArgs a = args[blockIdx.x];
  const int n = a.Mp, tid = threadIdx.x, nth = blockDim.x;
  extern __shared__ __align__(16) double sm[];
  double* D = sm;              // [32][33]
  double* Dinv = sm + 32 * 33; // [32] reciprocals of the current diagonal block's pivots
  double* Pn = Dinv + 32;      // panel, transposed: [32][n + 4]  (the inverse path reuses it as [32][n + 1])
  const int ldp = n + 4;
  double* L = a.L;

  for (long idx = tid; idx < (long)n * n; idx += nth) {
    int i = (int)(idx / n), j = (int)(idx % n);
    L[idx] = (j <= i) ? a.Ku[idx] : 0.0;
  }
  __syncthreads();

  const int nb = n / kCholNB;
  for (int kb = 0; kb < nb; ++kb) {
    const int k0 = kb * kCholNB;
    // 1. diagonal block -> smem, unblocked Cholesky
    for (int idx = tid; idx < 1024; idx += nth) {
      int r = idx >> 5, c = idx & 31;
      D[r * 33 + c] = L[(long)(k0 + r) * n + k0 + c];
    }
    __syncthreads();
    // right-looking, all threads on the block in shared memory: column j applies its rank-1 update
    // a_rc -= a_rj a_cj / d_jj to the columns right of it (one barrier per column), the scaling by 1 / sqrt(d_jj) comes last
    for (int j = 0; j < 32; ++j) {
      const double djj = D[j * 33 + j];
      const double inv = rsqrt(djj);
      if (tid == 0) {
        if (!(djj > 0.0)) *a.info = 1;
        Dinv[j] = inv;
      }
      const double w = inv * inv;
      for (int e = tid; e < 1024; e += nth) {
        const int r = e >> 5, c = e & 31;
        if (c > j && r >= c) D[r * 33 + c] = fma(-(D[r * 33 + j] * w), D[c * 33 + j], D[r * 33 + c]);
      }
      __syncthreads();
    }
    for (int e = tid; e < 1024; e += nth) {
      const int r = e >> 5, c = e & 31;
      D[r * 33 + c] = (r >= c) ? D[r * 33 + c] * Dinv[c] : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < 1024; idx += nth) {
      int r = idx >> 5, c = idx & 31;
      L[(long)(k0 + r) * n + k0 + c] = (c <= r) ? D[r * 33 + c] : 0.0;
    }
    // 2. panel solve: X * Lkk^T = A  (one thread per row below the block)
    const int nt = n - k0 - kCholNB;
    for (int r = tid; r < nt; r += nth) {
      double x[32];
      const double* arow = L + (long)(k0 + kCholNB + r) * n + k0;
#pragma unroll
      for (int c = 0; c < 32; ++c) x[c] = arow[c];
#pragma unroll
      for (int c = 0; c < 32; ++c) {   // right-looking: the dependency chain is one multiply + one fma per column
        x[c] *= Dinv[c];
#pragma unroll
        for (int c2 = 0; c2 < 32; ++c2)
          if (c2 > c) x[c2] = fma(-x[c], D[c2 * 33 + c], x[c2]);
      }
      double* orow = L + (long)(k0 + kCholNB + r) * n + k0;
#pragma unroll
      for (int c = 0; c < 32; ++c) { orow[c] = x[c]; Pn[c * ldp + r] = x[c]; }
    }
    __syncthreads();
    // 3. trailing update on the lower triangle, 4x4 micro-tiles dealt out over the triangle (consecutive threads: same row
    // strip, consecutive column strips, so the transposed panel is read as one broadcast + one contiguous run per k)
    const int nt4 = nt / 4;
    const int ntri = nt4 * (nt4 + 1) / 2;
    for (int tix = tid; tix < ntri; tix += nth) {
      int ti = (int)((sqrtf(8.f * (float)tix + 1.f) - 1.f) * 0.5f);
      while (ti * (ti + 1) / 2 > tix) --ti;
      while ((ti + 1) * (ti + 2) / 2 <= tix) ++ti;
      const int tj = tix - ti * (ti + 1) / 2;
      double acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = 0.0;
      const double* pi = Pn + ti * 4;
      const double* pj = Pn + tj * 4;
#pragma unroll 4
      for (int c = 0; c < 32; ++c) {
        const double2 a01 = *reinterpret_cast<const double2*>(pi + c * ldp), a23 = *reinterpret_cast<const double2*>(pi + c * ldp + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(pj + c * ldp), b23 = *reinterpret_cast<const double2*>(pj + c * ldp + 2);
        const double ai[4] = {a01.x, a01.y, a23.x, a23.y}, bj[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma(ai[p], bj[q], acc[p][q]);
      }
      const int r0 = k0 + kCholNB + ti * 4, c0 = k0 + kCholNB + tj * 4;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q <= r0 + p) L[(long)(r0 + p) * n + c0 + q] -= acc[p][q];
    }
    __syncthreads();
  }

  if (a.Linv == nullptr) return;   // the inverse is computed by tri_inv_kernel (one CTA per 32-column block)
  // ---- explicit inverse, block row by block row: X_i = Lii^{-1} (E_i - L[i, 0:i0] X[0:i0, :]) ----
  double* X = a.Linv;
  double* G = Pn;  // [32][n+1]
  const int ldg = n + 1;
  for (int ib = 0; ib < nb; ++ib) {
    const int i0 = ib * kCholNB;
    for (int idx = tid; idx < 1024; idx += nth) {
      int r = idx >> 5, c = idx & 31;
      D[r * 33 + c] = L[(long)(i0 + r) * n + i0 + c];
    }
    const int ncols = i0 + kCholNB;
    for (int idx = tid; idx < 32 * ncols; idx += nth) {
      int r = idx / ncols, c = idx % ncols;
      double s;
      if (c >= i0) {
        s = (c - i0 == r) ? 1.0 : 0.0;
      } else {
        s = 0.0;
        const double* lrow = L + (long)(i0 + r) * n;
        for (int k = c; k < i0; ++k) s = fma(-lrow[k], X[(long)k * n + c], s);
      }
      G[r * ldg + c] = s;
    }
    __syncthreads();
    for (int c = tid; c < ncols; c += nth) {
      double x[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        double s = G[r * ldg + c];
#pragma unroll
        for (int k = 0; k < r; ++k) s -= D[r * 33 + k] * x[k];
        x[r] = s / D[r * 33 + r];
      }
#pragma unroll
      for (int r = 0; r < 32; ++r) X[(long)(i0 + r) * n + c] = x[r];
    }
    for (int idx = tid; idx < 32 * (n - ncols); idx += nth) {
      int r = idx / (n - ncols), c = ncols + idx % (n - ncols);
      X[(long)(i0 + r) * n + c] = 0.0;
    }
    __syncthreads();
  }
  for (long idx = tid; idx < (long)n * n; idx += nth) {
    int i = (int)(idx / n), j = (int)(idx % n);
    a.LinvT[(long)j * n + i] = X[idx];
  }
}

// ---------------------------------------------------------------------------------------------------------
// Explicit inverse of the Cholesky factors, X = L^-1, by block columns: CTA (jb, layer) owns the 32 columns
// [32 jb, 32 jb + 32) and walks down the block rows i >= jb:
//   X_jj = L_jj^-1,   X_ij = L_ii^-1 ( - sum_{k=jb}^{i-1} L_ik X_kj )
// The block column lives in shared memory, L is read in coalesced 32x32 chunks. Block columns are independent, so the
// nb * layers CTAs run concurrently (the single-CTA version inside chol_inv_kernel took 0.6 ms for M = 256).
// ---------------------------------------------------------------------------------------------------------
inline size_t tri_inv_smem_bytes(int Mp) { return ((size_t)Mp * 33 + 3 * 32 * 33) * sizeof(double); }

__global__ void __launch_bounds__(256) tri_inv_kernel(const CholArgs* __restrict__ args, double* const* __restrict__ Linv_out,
                                                      double* const* __restrict__ LinvT_out) {
  const CholArgs a = args[blockIdx.y];
  const int n = a.Mp, nb = n / 32, jb = blockIdx.x;
  if (jb >= nb) return;
  extern __shared__ __align__(16) double sm[];
  double* Xc = sm;                  // [n][33]   block column of X (rows >= 32 jb are used)
  double* Lc = Xc + (size_t)n * 33; // [32][33]  chunk of L
  double* Ld = Lc + 32 * 33;        // [32][33]  diagonal block L_ii
  double* G = Ld + 32 * 33;         // [32][33]  right-hand side of the block solve
  const double* L = a.L;
  double* X = Linv_out[blockIdx.y];
  double* XT = LinvT_out[blockIdx.y];
  const int tid = threadIdx.x, j0 = jb * 32;
  const int r = tid >> 3, cs = (tid & 7) * 4;   // this thread's 1 x 4 strip of a 32 x 32 block

  for (int ib = jb; ib < nb; ++ib) {
    const int i0 = ib * 32;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    // the 32 x 32 chunks of L travel global -> registers -> shared memory, one chunk ahead of the products that consume them
    double pre[4], preD[4];
    auto fetch = [&](double* dst, int col0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int idx = tid + u * 256; dst[u] = L[(long)(i0 + (idx >> 5)) * n + col0 + (idx & 31)]; }
    };
    fetch(preD, i0);
    if (jb < ib) fetch(pre, jb * 32);
    for (int kb = jb; kb < ib; ++kb) {          // G = - sum_k L_ik X_kj
      __syncthreads();
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int idx = tid + u * 256; Lc[(idx >> 5) * 33 + (idx & 31)] = pre[u]; }
      __syncthreads();
      if (kb + 1 < ib) fetch(pre, (kb + 1) * 32);
      const double* xk = Xc + (size_t)(kb * 32) * 33 + cs;
#pragma unroll 8
      for (int k = 0; k < 32; ++k) {
        const double l = Lc[r * 33 + k];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = fma(-l, xk[k * 33 + q], acc[q]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int idx = tid + u * 256; Ld[(idx >> 5) * 33 + (idx & 31)] = preD[u]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) G[r * 33 + cs + q] = (ib == jb) ? ((r == cs + q) ? 1.0 : 0.0) : acc[q];
    __syncthreads();
    if (tid < 32) {                              // forward substitution, one column per thread
      // right-looking with the pivots' reciprocals: the dependency chain is one multiply + one fma per row, and the constant
      // loop bounds let both loops unroll completely (x[] in registers)
      const double myinv = 1.0 / Ld[tid * 33 + tid];
      double x[32];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) x[rr] = G[rr * 33 + tid];
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) {
        x[rr] *= __shfl_sync(0xffffffffu, myinv, rr);
#pragma unroll
        for (int r2 = 0; r2 < 32; ++r2)
          if (r2 > rr) x[r2] = fma(-Ld[r2 * 33 + rr], x[rr], x[r2]);
      }
#pragma unroll
      for (int rr = 0; rr < 32; ++rr) Xc[(size_t)(i0 + rr) * 33 + tid] = x[rr];
    }
    __syncthreads();
  }
  // write the block column (zeros above it) and its transpose
  for (int idx = tid; idx < n * 32; idx += 256) {
    const int i = idx >> 5, c = idx & 31;
    const double v = (i >= j0) ? Xc[(size_t)i * 33 + c] : 0.0;
    X[(long)i * n + j0 + c] = v;
  }
  for (int idx = tid; idx < n * 32; idx += 256) {
    const int c = idx / n, i = idx % n;
    XT[(long)(j0 + c) * n + i] = (i >= j0) ? Xc[(size_t)i * 33 + c] : 0.0;
  }
}

}  // namespace dgp
