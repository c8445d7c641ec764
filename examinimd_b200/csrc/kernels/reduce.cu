// reduce.cu -- thermo reductions.
// Replaces the Temperature / KinE functor sum_i m(type_i)*|v_i|^2 (src/property_temperature.h:55-57,
// src/property_kine.h) and the parallel_reduce that drives it (property_temperature.cpp:49).
// Two-stage deterministic reduction (fixed grid, fixed tree): 28 B/atom read.
#include "common.cuh"

using namespace emd;

namespace emd { int device_sum_partials(emd_ctx *ctx, const double *d_partial, int n, double *h_out); }

namespace {
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) mv2_kernel(const double *__restrict__ v, const int *__restrict__ type,
                                                       const double *__restrict__ mass, int n, double *__restrict__ partial) {
  __shared__ double sm[kThreads / 32];
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double vx = v[3 * (size_t)i], vy = v[3 * (size_t)i + 1], vz = v[3 * (size_t)i + 2];
    s += (vx * vx + vy * vy + vz * vz) * mass[type[i]];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (kThreads / 32) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (lane == 0) partial[blockIdx.x] = t;
  }
}
} // namespace

extern "C" int emd_reduce_mv2(emd_ctx *ctx, const double *d_v, const int *d_type, const double *d_mass, int n_local,
                              double *h_sum) {
  const int grid = max(1, min(grid_for(n_local, kThreads), ctx->num_sms * 8));
  if (ctx->s_c.ensure(sizeof(double) * ((size_t)grid + 8))) return 1;
  double *partial = ctx->s_c.as<double>() + 8;
  EMD_LAUNCH(ctx, mv2_kernel, grid, kThreads, 0, d_v, d_type, d_mass, n_local, partial);
  return device_sum_partials(ctx, partial, grid, h_sum);
}
