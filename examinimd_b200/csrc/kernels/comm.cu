// comm.cu -- single-process periodic communication kernels.
// Replaces CommSerial::{exchange, exchange_halo, update_halo, update_force}
// (src/comm_types/comm_serial.cpp:47-127; functors src/comm_types/comm_serial.h:94-213).
//
// Ghost creation is a STABLE stream compaction (flag -> exclusive scan -> scatter) instead of the
// reference's atomic counter, so ghosts appear in ascending source index = the reference's
// 1-thread arrival order, and a run is reproducible.  All kernels are HBM-bound byte movers:
//   wrap   : 24 B/atom read (+ rare writes)
//   phase  : 8 B/scanned atom (predicate) + 4+4 B flags/offsets + 72 B read / 72 B written per ghost
//   update : 72+4 B read, 72 B written per ghost (the reference re-copies the whole particle)
//   fold   : 4 + 24 + 24 B read, 24 B written per ghost
#include "common.cuh"

using namespace emd;

namespace {

__global__ void __launch_bounds__(256) wrap_kernel(double *__restrict__ x, long long n3, double Lx, double Ly, double Lz) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int d = (int)(e % 3);
  const double L = d == 0 ? Lx : (d == 1 ? Ly : Lz);
  const double xo = x[e]; // both tests use the OLD coordinate, comm_serial.h:97-99
  double xn = xo;
  if (xo > L) xn -= L;
  if (xo < 0) xn += L;
  if (xn != xo) x[e] = xn;
}

__global__ void __launch_bounds__(256) halo_flag_kernel(const double *__restrict__ x, int n_scan, int dim, int upper,
                                                        double thr, int *__restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_scan) return;
  const double xi = x[3 * (size_t)i + dim];
  flags[i] = upper ? (xi >= thr) : (xi <= thr); // comm_serial.h:113,124,...
}

__device__ __forceinline__ void copy_particle(double *x, double *v, double *q, int *id, int *type, size_t dst, size_t src,
                                              int dim, double shift) {
  double p[3] = {x[3 * src], x[3 * src + 1], x[3 * src + 2]};
  p[dim] += shift; // p.x -= domain_x  ==  p.x + (-domain_x), exactly
  x[3 * dst] = p[0]; x[3 * dst + 1] = p[1]; x[3 * dst + 2] = p[2];
  v[3 * dst] = v[3 * src]; v[3 * dst + 1] = v[3 * src + 1]; v[3 * dst + 2] = v[3 * src + 2];
  q[dst] = q[src]; id[dst] = id[src]; type[dst] = type[src];
}

__global__ void __launch_bounds__(256) halo_scatter_kernel(double *x, double *v, double *q, int *id, int *type, int n_scan,
                                                           const int *__restrict__ flags, const int *__restrict__ offsets,
                                                           int ghost_begin, int *__restrict__ pack, int dim, double shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_scan || !flags[i]) return;
  const int slot = offsets[i];
  pack[slot] = i;
  copy_particle(x, v, q, id, type, (size_t)ghost_begin + slot, (size_t)i, dim, shift);
}

__global__ void __launch_bounds__(256) halo_update_kernel(double *x, double *v, double *q, int *id, int *type,
                                                          const int *__restrict__ pack, int count, int ghost_begin, int dim,
                                                          double shift) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= count) return;
  copy_particle(x, v, q, id, type, (size_t)ghost_begin + ii, (size_t)pack[ii], dim, shift);
}

__global__ void __launch_bounds__(256) force_fold_kernel(double *f, const int *__restrict__ pack, int count, int ghost_begin) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= count) return;
  const size_t i = (size_t)pack[ii], g = (size_t)ghost_begin + ii; // a source index appears once per phase
  f[3 * i] += f[3 * g];
  f[3 * i + 1] += f[3 * g + 1];
  f[3 * i + 2] += f[3 * g + 2];
}

// ---- fused per-step refresh --------------------------------------------------------------------------------
// A ghost made in phase p copies its source shifted by -/+L in dimension p/2; the source may itself be a ghost of an
// earlier dimension (comm_serial.cpp:62-66), which is why the reference refreshes phase by phase.  Following that chain
// ONCE per ghost build gives every ghost its owned root atom and its total shift (at most one shift per dimension: an
// odd phase never re-scans the ghosts of its even twin), after which the whole refresh is one kernel with no
// dependencies instead of six dependent launches: x[ghost] = x[root] + shift, bit-identical to the phase-wise copy.
struct HaloLists { const int *pack[6]; int begin[7]; };

__global__ void __launch_bounds__(256) halo_resolve_kernel(HaloLists h, int n_local, int n_ghost, double Lx, double Ly, double Lz,
                                                           int *__restrict__ root, double *__restrict__ shift) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_ghost) return;
  const double L[3] = {Lx, Ly, Lz};
  double s[3] = {0.0, 0.0, 0.0};
  int i = n_local + k;
  while (i >= n_local) {
    int p = 0;
#pragma unroll
    for (int q = 1; q < 6; q++) p += (i >= h.begin[q]);
    s[p / 2] += (p % 2 == 0) ? -L[p / 2] : L[p / 2];
    i = h.pack[p][i - h.begin[p]];
  }
  root[k] = i;
  shift[3 * (size_t)k] = s[0]; shift[3 * (size_t)k + 1] = s[1]; shift[3 * (size_t)k + 2] = s[2];
}

__global__ void __launch_bounds__(256) halo_refresh_kernel(double *__restrict__ x, int n_local, long long n3, const int *__restrict__ root,
                                                           const double *__restrict__ shift) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const long long k = e / 3;
  const int d = (int)(e - 3 * k);
  x[3 * (size_t)n_local + e] = x[3 * (size_t)root[k] + d] + shift[e];
}

} // namespace

extern "C" {

int emd_comm_halo_resolve(emd_ctx *ctx, const int *const d_pack_indicies[6], const int counts[6], int n_local, const double domain[3],
                          int *d_root, double *d_shift) {
  HaloLists h;
  int n_ghost = 0;
  for (int p = 0; p < 6; p++) { h.pack[p] = d_pack_indicies[p]; h.begin[p] = n_local + n_ghost; n_ghost += counts[p]; }
  h.begin[6] = n_local + n_ghost;
  if (n_ghost <= 0) return 0;
  EMD_LAUNCH(ctx, halo_resolve_kernel, grid_for(n_ghost, 256), 256, 0, h, n_local, n_ghost, domain[0], domain[1], domain[2], d_root, d_shift);
  return 0;
}

int emd_comm_halo_refresh(emd_ctx *ctx, double *d_x, int n_local, int n_ghost, const int *d_root, const double *d_shift) {
  if (n_ghost <= 0) return 0;
  const long long n3 = 3LL * n_ghost;
  EMD_LAUNCH(ctx, halo_refresh_kernel, grid_for(n3, 256), 256, 0, d_x, n_local, n3, d_root, d_shift);
  return 0;
}

int emd_comm_wrap(emd_ctx *ctx, double *d_x, int n_local, const double domain[3]) {
  if (n_local <= 0) return 0;
  const long long n3 = 3LL * n_local;
  EMD_LAUNCH(ctx, wrap_kernel, grid_for(n3, 256), 256, 0, d_x, n3, domain[0], domain[1], domain[2]);
  return 0;
}

int emd_comm_halo_phase(emd_ctx *ctx, int phase, double *d_x, double *d_v, double *d_q, int *d_id, int *d_type, int n_scan,
                        int ghost_begin, int capacity, int *d_pack, int pack_capacity, const double domain[3],
                        const double lo[3], const double hi[3], double depth, int *h_count) {
  if (phase < 0 || phase > 5) { set_error("emd_comm_halo_phase: bad phase %d", phase); return 1; }
  const int dim = phase / 2, upper = (phase % 2 == 0);
  const double thr = upper ? hi[dim] - depth : lo[dim] + depth;
  const double shift = upper ? -domain[dim] : domain[dim];
  int count = 0;
  if (n_scan > 0) {
    if (ctx->s_a.ensure(sizeof(int) * (2 * (size_t)n_scan + 2))) return 1;
    int *flags = ctx->s_a.as<int>();
    int *offsets = flags + n_scan;
    int *d_total = offsets + n_scan;
    EMD_LAUNCH(ctx, halo_flag_kernel, grid_for(n_scan, 256), 256, 0, d_x, n_scan, dim, upper, thr, flags);
    if (exclusive_scan_int(ctx, flags, offsets, n_scan, d_total)) return 1;
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_total, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream)); // comm_serial.cpp:73
    count = ctx->h_pinned[0];
    // the reference writes what fits, grows, and redoes the phase (:74-90); we write nothing
    // unless everything fits -- the caller grows and calls again, ending in the same state
    if (count <= pack_capacity && ghost_begin + count <= capacity && count > 0)
      EMD_LAUNCH(ctx, halo_scatter_kernel, grid_for(n_scan, 256), 256, 0, d_x, d_v, d_q, d_id, d_type, n_scan, flags, offsets,
                 ghost_begin, d_pack, dim, shift);
  }
  if (h_count) *h_count = count;
  return 0;
}

int emd_comm_halo_update_phase(emd_ctx *ctx, int phase, double *d_x, double *d_v, double *d_q, int *d_id, int *d_type,
                               const int *d_pack, int count, int ghost_begin, const double domain[3]) {
  if (phase < 0 || phase > 5) { set_error("emd_comm_halo_update_phase: bad phase %d", phase); return 1; }
  if (count <= 0) return 0;
  const int dim = phase / 2;
  const double shift = (phase % 2 == 0) ? -domain[dim] : domain[dim];
  EMD_LAUNCH(ctx, halo_update_kernel, grid_for(count, 256), 256, 0, d_x, d_v, d_q, d_id, d_type, d_pack, count, ghost_begin,
             dim, shift);
  return 0;
}

int emd_comm_force_fold_phase(emd_ctx *ctx, double *d_f, const int *d_pack, int count, int ghost_begin) {
  if (count <= 0) return 0;
  EMD_LAUNCH(ctx, force_fold_kernel, grid_for(count, 256), 256, 0, d_f, d_pack, count, ghost_begin);
  return 0;
}

} // extern "C"
