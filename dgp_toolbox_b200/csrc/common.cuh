// Shared device/host helpers for the DGP hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <type_traits>
#include <cstdio>
#include <string>

#define DGP_OK 0
#define DGP_ERR_CUDA -1
#define DGP_ERR_ARG -2
#define DGP_ERR_UNSUPPORTED -3
#define DGP_ERR_NUMERIC -4

namespace dgp {

constexpr int kTileM = 64;     // inducing dimension is padded to a multiple of this
constexpr int kTileP = 128;    // point-sample dimension is padded to a multiple of this
constexpr int kMaxD = 32;      // maximum layer width (D_in, D_out) supported by the skinny kernels

__host__ __device__ inline long round_up(long x, long m) { return (x + m - 1) / m * m; }

// FP64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col). SASS: DMMA.8x8x4 (the only FP64 MMA on sm_100a).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Stationary kernels of the layers, as functions of the scaled squared distance r2 = sum_j ((x_j - x'_j) / l_j)^2
// (GPflow semantics; call sites dgp_dace/utils/layers.py:221,230,243 and BO/SO_BO.py:190-197,237-244):
//   kind 0 SquaredExponential: s2 exp(-r2 / 2)
//   kind 1 Matern32:           s2 (1 + sqrt3 r) exp(-sqrt3 r),               r = sqrt(max(r2, 1e-36))
//   kind 2 Matern52:           s2 (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r)
// kernel_gfac = -2 dk/d(r2), the factor the adjoint multiplies K-bar with (for the SquaredExponential it equals k itself);
// it is finite at r = 0 and the coordinate differences it multiplies vanish there, matching autodiff through the clamp.
// exp(x) for x <= 0 without control flow. libdevice's exp() carries a branch for its special cases, which keeps the compiler from
// interleaving independent evaluations: the kernel sweeps of the fused kernels ran one ~25-deep dependent FP64 chain at a time
// (≈ 400 clocks per value, measured). Cody-Waite reduction x = n ln2 + r, |r| <= ln2 / 2, degree-13 Taylor polynomial in Horner form
// (truncation r^14 / 14! < 5e-18), 2^n applied to the exponent field; arguments below -700 are clamped (the true value is < 1e-304),
// exp(0) = 1 exactly, NaN propagates. Error <= 1.5 ulp (tests/: against exp() over the kernels' argument range).
__device__ __forceinline__ double exp_nonpos(double x) {
  const double xc = fmax(x, -700.0);
  double t = fma(xc, 1.4426950408889634074, 6755399441055744.0);   // 1.5 * 2^52: the integer n lands in the low word
  const int n = __double2loint(t);
  t -= 6755399441055744.0;
  double r = fma(t, -6.93147180369123816490e-01, xc);               // ln2 split hi / lo (hi has 21 trailing zero bits: t * hi is exact)
  r = fma(t, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;                                // 1 / 13!
  p = fma(p, r, 2.08767569878681e-09);                              // 1 / 12!
  p = fma(p, r, 2.505210838544172e-08);                             // 1 / 11!
  p = fma(p, r, 2.755731922398589e-07);                             // 1 / 10!
  p = fma(p, r, 2.7557319223985893e-06);                            // 1 / 9!
  p = fma(p, r, 2.48015873015873e-05);                              // 1 / 8!
  p = fma(p, r, 1.984126984126984e-04);                             // 1 / 7!
  p = fma(p, r, 1.388888888888889e-03);                             // 1 / 6!
  p = fma(p, r, 8.333333333333333e-03);                             // 1 / 5!
  p = fma(p, r, 4.1666666666666664e-02);                            // 1 / 4!
  p = fma(p, r, 1.6666666666666666e-01);                            // 1 / 3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double y = __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  return x != x ? x : y;
}

__device__ __forceinline__ void kernel_eval(int kind, double r2, double s2, double& k, double& gfac) {
  if (kind == 0) {
    k = s2 * exp_nonpos(-0.5 * r2);
    gfac = k;
  } else if (kind == 1) {
    const double a = 1.7320508075688772935;
    const double r = sqrt(fmax(r2, 1e-36)), e = exp_nonpos(-a * r);
    k = s2 * (1.0 + a * r) * e;
    gfac = 3.0 * s2 * e;
  } else {
    const double a = 2.2360679774997896964;
    const double r = sqrt(fmax(r2, 1e-36)), e = exp_nonpos(-a * r);
    k = s2 * (1.0 + a * r + (5.0 / 3.0) * r * r) * e;
    gfac = (5.0 / 3.0) * s2 * (1.0 + a * r) * e;
  }
}
__device__ __forceinline__ double kernel_value(int kind, double r2, double s2) {
  double k, g;
  kernel_eval(kind, r2, s2, k, g);
  return k;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block reduction (fixed tree), result valid in thread 0. `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace dgp
