// microbench.cu -- B200 hardware numbers the kernel designs in DESIGN.md rest on.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu && ./microbench
// Prints one JSON object: FP64 FMA peak, HBM copy bandwidth, global f64 reduction (REDG) rate
// under an MD-like scatter, shared-memory f64 atomic (CAS loop) rate, gather bandwidth.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double *out, double a, double b, int iters) {
  double acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += acc[k];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

// each thread (atom i) issues K pairs x 3 REDG to f[3*j+c], j = i + off[k] (cell-sorted locality)
__global__ void __launch_bounds__(128) red_kernel(double *f, const int *__restrict__ offs, int n, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < K; k++) {
    int j = i + offs[(i * 7 + k) & 1023];
    j = j < 0 ? j + n : (j >= n ? j - n : j);
    atomicAdd(&f[3 * (size_t)j], 1.0);
    atomicAdd(&f[3 * (size_t)j + 1], 1.0);
    atomicAdd(&f[3 * (size_t)j + 2], 1.0);
  }
}

// same traffic but through shared memory (CAS-loop atomics) on a 2048-atom tile
__global__ void __launch_bounds__(256) smem_atomic_kernel(double *f, const int *__restrict__ offs, int K) {
  __shared__ double s[2048 * 3];
  for (int t = threadIdx.x; t < 2048 * 3; t += blockDim.x) s[t] = 0;
  __syncthreads();
  for (int rep = 0; rep < 8; rep++) {
    const int i = rep * 256 + threadIdx.x;
    for (int k = 0; k < K; k++) {
      const int j = (i + offs[(i * 7 + k) & 1023]) & 2047;
      atomicAdd(&s[3 * j], 1.0);
      atomicAdd(&s[3 * j + 1], 1.0);
      atomicAdd(&s[3 * j + 2], 1.0);
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2048 * 3; t += blockDim.x) f[(size_t)blockIdx.x * 2048 * 3 + t] = s[t];
}

// gather 24 B rows x[j] with local offsets, sum into a register
__global__ void __launch_bounds__(128) gather_kernel(const double *__restrict__ x, const int *__restrict__ offs, double *out, int n, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0;
  for (int k = 0; k < K; k++) {
    int j = i + offs[(i * 7 + k) & 1023];
    j = j < 0 ? j + n : (j >= n ? j - n : j);
    s += x[3 * (size_t)j] + x[3 * (size_t)j + 1] + x[3 * (size_t)j + 2];
  }
  out[i] = s;
}

template <class F>
float time_ms(F f, int reps = 10) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    best = std::min(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double *d_out; CK(cudaMalloc(&d_out, 1 << 20));
  // --- DFMA
  const int iters = 4096;
  constexpr int ILP = 8;
  float ms = time_ms([&] { dfma_kernel<ILP><<<sms * 8, 256>>>(d_out, 1.0000001, 1e-9, iters); });
  const double dfma_tf = 2.0 * sms * 8 * 256.0 * ILP * iters / (ms * 1e-3) / 1e12;
  // sustained: 2 seconds back to back
  {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    int n = 0;
    for (; n < (int)(2000.0 / ms); n++) dfma_kernel<ILP><<<sms * 8, 256>>>(d_out, 1.0000001, 1e-9, iters);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float t; CK(cudaEventElapsedTime(&t, a, b));
    printf("{\"sms\": %d, \"l2_bytes\": %d, \"fp64_tflops_burst\": %.2f, \"fp64_tflops_sustained\": %.2f", sms, p.l2CacheSize, dfma_tf,
           2.0 * sms * 8 * 256.0 * ILP * iters * n / (t * 1e-3) / 1e12);
  }
  // --- copy
  const size_t nb = (size_t)1 << 30;
  double2 *a, *b; CK(cudaMalloc(&a, nb)); CK(cudaMalloc(&b, nb));
  CK(cudaMemset(a, 1, nb));
  ms = time_ms([&] { copy_kernel<<<sms * 16, 256>>>(a, b, nb / 16); });
  printf(", \"copy_gbs\": %.1f", 2.0 * nb / (ms * 1e-3) / 1e9);
  ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, nb, cudaMemcpyDeviceToDevice)); });
  printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * nb / (ms * 1e-3) / 1e9);
  // --- REDG f64, MD-like locality: offsets within +-600 atoms (a 27-cell stencil of 21-atom cells in a z-pencil order
  // is really three clusters; take a mix of near (+-60) and pencil-distance (+-1000, +-47000) offsets
  const int n = 2048000, K = 27;
  std::vector<int> offs(1024);
  srand(1);
  for (int k = 0; k < 1024; k++) {
    const int base[9] = {0, 987, -987, 46389, -46389, 47376, -47376, 45402, -45402};
    offs[k] = base[rand() % 9] + (rand() % 63) - 31;
  }
  int *d_offs; CK(cudaMalloc(&d_offs, 4096)); CK(cudaMemcpy(d_offs, offs.data(), 4096, cudaMemcpyHostToDevice));
  double *f; CK(cudaMalloc(&f, sizeof(double) * 3 * n)); CK(cudaMemset(f, 0, sizeof(double) * 3 * n));
  ms = time_ms([&] { red_kernel<<<(n + 127) / 128, 128>>>(f, d_offs, n, K); });
  printf(", \"redg_f64_gops\": %.1f, \"redg_ms_2M_x81\": %.3f", 3.0 * n * K / (ms * 1e-3) / 1e9, ms);
  // --- shared atomics
  const int nblk = n / 2048;
  ms = time_ms([&] { smem_atomic_kernel<<<nblk, 256>>>(f, d_offs, K); });
  printf(", \"smem_cas_f64_gops\": %.1f, \"smem_ms_2M_x81\": %.3f", 3.0 * nblk * 2048.0 * K / (ms * 1e-3) / 1e9, ms);
  // --- gather
  double *o; CK(cudaMalloc(&o, sizeof(double) * n));
  ms = time_ms([&] { gather_kernel<<<(n + 127) / 128, 128>>>(f, d_offs, o, n, 39); });
  printf(", \"gather_rows_g_per_s\": %.1f, \"gather_gbs\": %.1f, \"gather_ms_2M_x39\": %.3f}\n", 39.0 * n / (ms * 1e-3) / 1e9,
         24.0 * 39.0 * n / (ms * 1e-3) / 1e9, ms);
  return 0;
}
