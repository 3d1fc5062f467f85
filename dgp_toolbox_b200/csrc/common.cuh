// Shared device/host helpers for the DGP hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
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
