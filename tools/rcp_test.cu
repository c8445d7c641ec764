// rcp_test.cu -- accuracy of the force kernel's reciprocal (MUFU.RCP64H seed + one cubic step) against IEEE division.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rcp_test rcp_test.cu && ./rcp_test
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
__device__ __forceinline__ double seed_rcp(double a) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a)); return y; }
__global__ void k(double *out, int n) {
  double max_seed = 0, max_cubic = 0, bias = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double a = 0.5 + 7.5 * (i + 0.37) / n; // rsq range of an LJ liquid
    const double y0 = seed_rcp(a), ex = 1.0 / a;
    const double e = fma(-a, y0, 1.0), t = fma(e, e, e), y = fma(y0, t, y0);
    max_seed = fmax(max_seed, fabs(y0 - ex) / ex);
    max_cubic = fmax(max_cubic, fabs(y - ex) / ex);
    bias += (y - ex) / ex;
  }
  atomicMax((unsigned long long *)&out[0], __double_as_longlong(max_seed));
  atomicMax((unsigned long long *)&out[1], __double_as_longlong(max_cubic));
  atomicAdd(&out[2], bias);
}
int main() {
  double *d, h[3] = {0, 0, 0};
  cudaMalloc(&d, 24); cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice);
  const int n = 1 << 26;
  k<<<1184, 256>>>(d, n);
  cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  printf("{\"rcp_seed_max_rel_err\": %.3e, \"rcp_cubic_max_rel_err\": %.3e, \"rcp_cubic_mean_rel_err\": %.3e}\n", h[0], h[1], h[2] / n);
  return 0;
}
