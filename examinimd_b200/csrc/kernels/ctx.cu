// ctx.cu -- context lifetime, error string, memory helpers, exclusive scan.
#include "common.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <cstdlib>
#include <cstdio>
#include <cstdarg>

namespace emd {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------ exclusive scan (int)
// Tile = 2048 ints per CTA (256 threads x 8).  Pass 1 reduces each tile, pass 2 scans the tile
// sums inside one CTA (serial carry over 2048-wide chunks), pass 3 rescans each tile with its
// offset.  2*n*4 B read + n*4 B written; used only on rebuild steps (bins, rows, halo flags).
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *smem /*>=kScanThreads/32+1*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (kScanThreads / 32) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < (kScanThreads / 32)) smem[lane] = winc - w; // exclusive warp offsets
    if (lane == (kScanThreads / 32) - 1) smem[kScanThreads / 32] = winc;
  }
  __syncthreads();
  int excl = inc - v + smem[warp];
  *total = smem[kScanThreads / 32];
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_reduce(const int *__restrict__ in, int n, int *__restrict__ tile_sums) {
  __shared__ int sm[kScanThreads / 32 + 1];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
    if (base + k < n) s += in[base + k];
  int total;
  block_exclusive_scan(s, &total, sm);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(int *tile_sums, int ntiles, int *d_total) {
  __shared__ int sm[kScanThreads / 32 + 1];
  int carry = 0;
  for (int base = 0; base < ntiles; base += kScanThreads) {
    int idx = base + threadIdx.x;
    int v = idx < ntiles ? tile_sums[idx] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total, sm);
    if (idx < ntiles) tile_sums[idx] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0 && d_total) *d_total = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_final(const int *in, int *out, int n, const int *__restrict__ tile_sums) {
  __shared__ int sm[kScanThreads / 32 + 1];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int ex = block_exclusive_scan(s, &total, sm) + tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
  }
}

int exclusive_scan_int(emd_ctx *ctx, const int *d_in, int *d_out, int n, int *d_total) {
  if (n <= 0) {
    if (d_total) EMD_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int), ctx->stream));
    return 0;
  }
  const int ntiles = (n + kScanTile - 1) / kScanTile;
  if (ctx->s_scan.ensure(sizeof(int) * (size_t)ntiles)) return 1;
  int *sums = ctx->s_scan.as<int>();
  EMD_LAUNCH(ctx, scan_tile_reduce, ntiles, kScanThreads, 0, d_in, n, sums);
  EMD_LAUNCH(ctx, scan_tile_sums, 1, kScanThreads, 0, sums, ntiles, d_total);
  EMD_LAUNCH(ctx, scan_tile_final, ntiles, kScanThreads, 0, d_in, d_out, n, sums);
  return 0;
}
} // namespace emd

using namespace emd;

namespace {
// FP64 FMA peak of this device (the SNAP roofline's denominator; MEASURED_PEAKS.json carries HBM and bf16 only):
// 8 independent DFMA chains per thread, 8 CTAs of 256 threads per SM, best of `reps` launches.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, double a, double b, int iters) {
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; k++) acc[k] = threadIdx.x * 1e-9 + k;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = fma(acc[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += acc[k];
  if (s == 12345.678) out[0] = s;
}
} // namespace

extern "C" {

int emd_microbench_fp64(emd_ctx *ctx, int reps, double *h_tflops_best, double *h_tflops_mean) {
  if (ctx->s_c.ensure(64)) return 1;
  const int iters = 4096, grid = ctx->num_sms * 8;
  const double flop = 2.0 * grid * 256.0 * 8 * iters;
  cudaEvent_t a, b;
  EMD_CUDA(cudaEventCreate(&a)); EMD_CUDA(cudaEventCreate(&b));
  float best = 1e30f, sum = 0.f;
  for (int r = 0; r < reps + 2; r++) {
    EMD_CUDA(cudaEventRecord(a, ctx->stream));
    dfma_peak_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->s_c.as<double>(), 1.0000001, 1e-9, iters);
    EMD_CUDA(cudaEventRecord(b, ctx->stream));
    EMD_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    EMD_CUDA(cudaEventElapsedTime(&ms, a, b));
    if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (h_tflops_best) *h_tflops_best = flop / (best * 1e-3) / 1e12;
  if (h_tflops_mean) *h_tflops_mean = flop / (sum / reps * 1e-3) / 1e12;
  return 0;
}

const char *emd_last_error(void) { return emd::g_err; }
int emd_abi_version(void) { return EMD_ABI_VERSION; }

int emd_ctx_create(emd_ctx **out, int device, void *stream) {
  if (!out) { set_error("emd_ctx_create: out == NULL"); return 1; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("emd_ctx_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return 1;
  }
  if (device < 0 || device >= ndev) { set_error("emd_ctx_create: device %d out of range (%d devices)", device, ndev); return 1; }
  EMD_CUDA(cudaSetDevice(device));
  emd_ctx *c = new emd_ctx();
  c->device = device;
  if (stream) { c->stream = (cudaStream_t)stream; c->own_stream = false; }
  else { EMD_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
  EMD_CUDA(cudaEventCreate(&c->ev0));
  EMD_CUDA(cudaEventCreate(&c->ev1));
  EMD_CUDA(cudaMallocHost((void **)&c->h_pinned, 64 * sizeof(double)));
  cudaDeviceProp prop;
  EMD_CUDA(cudaGetDeviceProperties(&prop, device));
  c->num_sms = prop.multiProcessorCount;
  *out = c;
  return 0;
}

void emd_ctx_destroy(emd_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->s_a.release(); c->s_b.release(); c->s_c.release(); c->s_scan.release();
  if (c->d_lj_tables) cudaFree(c->d_lj_tables);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  if (c->side_stream) cudaStreamDestroy(c->side_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

void *emd_ctx_stream(emd_ctx *c) { return (void *)c->stream; }
int emd_ctx_sync(emd_ctx *c) { EMD_CUDA(cudaStreamSynchronize(c->stream)); return 0; }

namespace {
__global__ void halo_gate_wait_kernel(const int *flags, int seq, int mask) {
  const int p = threadIdx.x;
  if (p < 6 && ((mask >> p) & 1)) {
    int v;
    do { asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory"); if (v < seq) __nanosleep(200); } while (v < seq);
  }
  __threadfence_system();
}
} // namespace
int emd_ctx_set_halo_gate(emd_ctx *c, const int *d_arrived6, int seq, int phase_mask) {
  c->gate_flags = d_arrived6; c->gate_seq = seq; c->gate_mask = phase_mask; c->gate_pending = d_arrived6 && phase_mask;
  return 0;
}
int emd_ctx_halo_gate_wait(emd_ctx *c) {
  if (!c->gate_pending) return 0;
  c->gate_pending = false;
  EMD_LAUNCH(c, halo_gate_wait_kernel, 1, 32, 0, c->gate_flags, c->gate_seq, c->gate_mask);
  return 0;
}
int emd_ctx_halo_gate_pending(const emd_ctx *c) { return c->gate_pending ? 1 : 0; }
unsigned long long emd_ctx_launch_count(emd_ctx *c) { return c->launches; }
// The side stream lives in a GREEN CONTEXT that owns all but a few SMs (driver API, CUDA >= 12.4, resolved at run time so the
// library has no link-time dependency on libcuda): the persistent force kernel launched on it fills ITS SMs, and the pack and
// NCCL kernels of the halo exchange, on the module stream, always find the remaining SMs empty.  Measured on 2 B200s: with a
// plain low-priority stream the exchange kernels queue behind the persistent CTAs (stream priority does not preempt) and the
// overlap hides a quarter of the exchange; with 20 SMs set aside the 2-GPU step is 5.6 % shorter than without overlap
// (gpurun r01p).  EMD_OVERLAP_SMS = SMs left to the exchange (default 20; 0 = plain stream).
static cudaStream_t partition_stream(emd_ctx *c, int comm_sms, int priority, int *sms_out) {
  void *h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  typedef CUresult (*f_devget)(CUdevice *, int);
  typedef CUresult (*f_getres)(CUdevice, CUdevResource *, CUdevResourceType);
  typedef CUresult (*f_split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
  typedef CUresult (*f_desc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
  typedef CUresult (*f_gcreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
  typedef CUresult (*f_gstream)(CUstream *, CUgreenCtx, unsigned int, int);
  f_devget devget = (f_devget)dlsym(h, "cuDeviceGet");
  f_getres getres = (f_getres)dlsym(h, "cuDeviceGetDevResource");
  f_split split = (f_split)dlsym(h, "cuDevSmResourceSplitByCount");
  f_desc gendesc = (f_desc)dlsym(h, "cuDevResourceGenerateDesc");
  f_gcreate gcreate = (f_gcreate)dlsym(h, "cuGreenCtxCreate");
  f_gstream gstream = (f_gstream)dlsym(h, "cuGreenCtxStreamCreate");
  if (!devget || !getres || !split || !gendesc || !gcreate || !gstream) return nullptr;
  CUdevice dev;
  CUdevResource sm, grp, rem;
  if (devget(&dev, c->device) != CUDA_SUCCESS || getres(dev, &sm, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
  const int total = (int)sm.sm.smCount;
  if (comm_sms <= 0 || comm_sms >= total) return nullptr;
  unsigned int ngroups = 1;
  if (split(&grp, &ngroups, &sm, &rem, 0, (unsigned)(total - comm_sms)) != CUDA_SUCCESS || ngroups != 1) return nullptr;
  if ((int)grp.sm.smCount >= total) return nullptr; // the granularity left nothing for the exchange
  CUdevResourceDesc desc;
  CUgreenCtx g;
  CUstream st;
  if (gendesc(&desc, &grp, 1) != CUDA_SUCCESS || gcreate(&g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
  if (gstream(&st, g, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS) return nullptr;
  *sms_out = (int)grp.sm.smCount;
  return (cudaStream_t)st; // the green context lives as long as the process
}

static int side_init(emd_ctx *c) {
  if (c->side_stream) return 0;
  int lo = 0, hi = 0;
  EMD_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi)); // lo = least priority: the exchange on the module stream goes first
  const char *e = getenv("EMD_OVERLAP_SMS");
  const int comm_sms = e ? atoi(e) : 20;
  c->side_sms = 0;
  c->side_stream = partition_stream(c, comm_sms, lo, &c->side_sms);
  if (!c->side_stream) {
    c->side_sms = 0;
    EMD_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, lo));
  }
  if (getenv("EMD_VERBOSE")) fprintf(stderr, "emd: side stream on %d SMs (%s)\n", c->side_sms ? c->side_sms : c->num_sms, c->side_sms ? "green context" : "shared");
  EMD_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  EMD_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  return 0;
}
int emd_ctx_side_sms(const emd_ctx *c) { return (c->main_stream && c->side_sms) ? c->side_sms : c->num_sms; }
int emd_ctx_side_mark(emd_ctx *c) {
  if (c->main_stream) { set_error("emd_ctx_side_mark: already on the side stream"); return 1; }
  if (side_init(c)) return 1;
  EMD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  c->fork_marked = true;
  return 0;
}
int emd_ctx_side_begin(emd_ctx *c) {
  if (c->main_stream) { set_error("emd_ctx_side_begin: already on the side stream"); return 1; }
  if (side_init(c)) return 1;
  if (!c->fork_marked) EMD_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  c->fork_marked = false;
  EMD_CUDA(cudaStreamWaitEvent(c->side_stream, c->ev_fork, 0));
  c->main_stream = c->stream;
  c->stream = c->side_stream;
  return 0;
}
int emd_ctx_side_end(emd_ctx *c) {
  if (!c->main_stream) { set_error("emd_ctx_side_end: not on the side stream"); return 1; }
  EMD_CUDA(cudaEventRecord(c->ev_join, c->side_stream));
  c->stream = c->main_stream;
  c->main_stream = nullptr;
  return 0;
}
int emd_ctx_side_join(emd_ctx *c) {
  if (c->main_stream || !c->side_stream) { set_error("emd_ctx_side_join: no side work"); return 1; }
  EMD_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  return 0;
}
int emd_ctx_tic(emd_ctx *c) { EMD_CUDA(cudaEventRecord(c->ev0, c->stream)); return 0; }
int emd_ctx_toc(emd_ctx *c, float *h_ms) {
  EMD_CUDA(cudaEventRecord(c->ev1, c->stream));
  EMD_CUDA(cudaEventSynchronize(c->ev1));
  EMD_CUDA(cudaEventElapsedTime(h_ms, c->ev0, c->ev1));
  return 0;
}

int emd_event_create(void **ev) {
  cudaEvent_t e;
  EMD_CUDA(cudaEventCreate(&e));
  *ev = (void *)e;
  return 0;
}
int emd_event_destroy(void *ev) { if (ev) EMD_CUDA(cudaEventDestroy((cudaEvent_t)ev)); return 0; }
int emd_event_record(emd_ctx *c, void *ev) { EMD_CUDA(cudaEventRecord((cudaEvent_t)ev, c->stream)); return 0; }
int emd_event_elapsed_ms(void *a, void *b, float *h_ms) {
  EMD_CUDA(cudaEventSynchronize((cudaEvent_t)b));
  EMD_CUDA(cudaEventElapsedTime(h_ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return 0;
}

int emd_malloc(void **d_ptr, unsigned long long bytes) { EMD_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1)); return 0; }
int emd_free(void *d_ptr) { if (d_ptr) EMD_CUDA(cudaFree(d_ptr)); return 0; }
int emd_memcpy_h2d(emd_ctx *c, void *d, const void *h, unsigned long long bytes) {
  EMD_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream));
  return 0;
}
int emd_memcpy_d2h(emd_ctx *c, void *h, const void *d, unsigned long long bytes) {
  EMD_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream));
  EMD_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
int emd_memcpy_d2d(emd_ctx *c, void *d, const void *s, unsigned long long bytes) {
  EMD_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
int emd_memset_zero(emd_ctx *c, void *d, unsigned long long bytes) {
  EMD_CUDA(cudaMemsetAsync(d, 0, bytes, c->stream));
  return 0;
}

} // extern "C"
