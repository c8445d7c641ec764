// force_lj.cu -- Lennard-Jones pair force and energy over a neighbor list.
// Replaces ForceLJNeigh<>::compute / compute_energy (src/force_types/force_lj_neigh_impl.h:100-156;
// functors TagFullNeigh :161-206, TagHalfNeigh :208-254, TagFullNeighPE :256-296,
// TagHalfNeighPE :298-343).
//
// v1: one thread per local atom, rows read through the duck-typed list (CSR or 2D).  Positions
// of the ~40-78 listed neighbours are gathered from L1/L2 (atoms are cell-sorted, so a warp's
// rows overlap heavily); f_i is accumulated in registers and written once; in half mode f_j is
// scattered with red.global.add.f64 exactly like the reference's atomic view (:245-247).
// The pair arithmetic keeps the reference's expression order; only the row summation order is
// the list order (same as the serial reference) and FMA contraction differs (<= 1 ulp/term).
#include "common.cuh"

using namespace emd;

namespace {

struct LJTable { // by-value kernel argument (<= 12 types, like the reference's stack params)
  double lj1[kMaxTypesConst * kMaxTypesConst];
  double lj2[kMaxTypesConst * kMaxTypesConst];
  double cutsq[kMaxTypesConst * kMaxTypesConst];
  int ntypes;
};

__device__ __forceinline__ void row_of(const emd_neigh_list &l, int i, const int *&row, int &n) {
  if (l.d_row_map) {
    const int b = l.d_row_map[i];
    n = l.d_row_map[i + 1] - b;
    row = l.d_neighs + b;
  } else {
    n = l.d_num_neighs[i];
    row = l.d_neighs + (size_t)i * l.stride;
  }
}

template <bool HALF, bool ONETYPE>
__global__ void __launch_bounds__(128) lj_force_kernel(const double *__restrict__ x, const int *__restrict__ type,
                                                        double *f, int n_local, emd_neigh_list list,
                                                        const LJTable tab, int overwrite) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const double x_i = x[3 * (size_t)i], y_i = x[3 * (size_t)i + 1], z_i = x[3 * (size_t)i + 2];
  const int type_i = ONETYPE ? 0 : type[i];
  const int *row;
  int num_neighs;
  row_of(list, i, row, num_neighs);

  double fxi = 0.0, fyi = 0.0, fzi = 0.0;
  for (int jj = 0; jj < num_neighs; jj++) {
    const int j = row[jj];
    const double dx = x_i - x[3 * (size_t)j];
    const double dy = y_i - x[3 * (size_t)j + 1];
    const double dz = z_i - x[3 * (size_t)j + 2];
    const int tij = ONETYPE ? 0 : type_i * tab.ntypes + type[j];
    const double rsq = dx * dx + dy * dy + dz * dz;
    if (rsq < tab.cutsq[tij]) {
      const double r2inv = 1.0 / rsq;
      const double r6inv = r2inv * r2inv * r2inv;
      const double fpair = (r6inv * (tab.lj1[tij] * r6inv - tab.lj2[tij])) * r2inv;
      fxi += dx * fpair;
      fyi += dy * fpair;
      fzi += dz * fpair;
      if (HALF) {
        atomicAdd(&f[3 * (size_t)j], -(dx * fpair));
        atomicAdd(&f[3 * (size_t)j + 1], -(dy * fpair));
        atomicAdd(&f[3 * (size_t)j + 2], -(dz * fpair));
      }
    }
  }
  if (HALF) {
    atomicAdd(&f[3 * (size_t)i], fxi);
    atomicAdd(&f[3 * (size_t)i + 1], fyi);
    atomicAdd(&f[3 * (size_t)i + 2], fzi);
  } else if (overwrite) {
    f[3 * (size_t)i] = fxi; f[3 * (size_t)i + 1] = fyi; f[3 * (size_t)i + 2] = fzi;
  } else {
    f[3 * (size_t)i] += fxi; f[3 * (size_t)i + 1] += fyi; f[3 * (size_t)i + 2] += fzi;
  }
}


// ForceLJIDialNeigh (pair_style lj/cut/idial): the same pair force accumulated `intensity` times, each term divided by
// `intensity` -- a compute-intensity dial (src/force_types/force_lj_idial_neigh_impl.h:113-163 full, :165-213 half).  The
// repeat loop is kept as arithmetic (the asm statement makes rsq opaque per iteration, so the divide and the five multiplies
// are really issued `intensity` times and the rounding sequence of fpair += term / intensity is the reference's).  Half
// lists subtract from j only when j is an owned atom (:203-207), whatever `newton` says.
struct IDialTable { double intensity[kMaxTypesConst * kMaxTypesConst]; };

template <bool HALF>
__global__ void __launch_bounds__(128) lj_idial_force_kernel(const double *__restrict__ x, const int *__restrict__ type, double *f,
                                                              int n_local, emd_neigh_list list, const LJTable tab, const IDialTable dial,
                                                              int overwrite) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const double x_i = x[3 * (size_t)i], y_i = x[3 * (size_t)i + 1], z_i = x[3 * (size_t)i + 2];
  const int type_i = type[i];
  const int *row;
  int num_neighs;
  row_of(list, i, row, num_neighs);
  double fxi = 0.0, fyi = 0.0, fzi = 0.0;
  for (int jj = 0; jj < num_neighs; jj++) {
    const int j = row[jj];
    const double dx = x_i - x[3 * (size_t)j], dy = y_i - x[3 * (size_t)j + 1], dz = z_i - x[3 * (size_t)j + 2];
    const int tij = type_i * tab.ntypes + type[j];
    double rsq = dx * dx + dy * dy + dz * dz;
    if (rsq < tab.cutsq[tij]) {
      const double lj1 = tab.lj1[tij], lj2 = tab.lj2[tij], inten = dial.intensity[tij];
      double fpair = 0.0;
      for (int repeat = 0; repeat < inten; repeat++) {
        asm volatile("" : "+d"(rsq));
        const double r2inv = 1.0 / rsq;
        const double r6inv = r2inv * r2inv * r2inv;
        fpair += (r6inv * (lj1 * r6inv - lj2)) * r2inv / inten;
      }
      fxi += dx * fpair; fyi += dy * fpair; fzi += dz * fpair;
      if (HALF && j < n_local) {
        atomicAdd(&f[3 * (size_t)j], -(dx * fpair));
        atomicAdd(&f[3 * (size_t)j + 1], -(dy * fpair));
        atomicAdd(&f[3 * (size_t)j + 2], -(dz * fpair));
      }
    }
  }
  if (HALF) {
    atomicAdd(&f[3 * (size_t)i], fxi); atomicAdd(&f[3 * (size_t)i + 1], fyi); atomicAdd(&f[3 * (size_t)i + 2], fzi);
  } else if (overwrite) {
    f[3 * (size_t)i] = fxi; f[3 * (size_t)i + 1] = fyi; f[3 * (size_t)i + 2] = fzi;
  } else {
    f[3 * (size_t)i] += fxi; f[3 * (size_t)i + 1] += fyi; f[3 * (size_t)i + 2] += fzi;
  }
}

constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double *sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (kRedThreads / 32) ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  return t; // valid in thread 0
}

template <bool HALF, bool ONETYPE>
__global__ void __launch_bounds__(kRedThreads) lj_energy_kernel(const double *__restrict__ x, const int *__restrict__ type,
                                                                int n_local, emd_neigh_list list, const LJTable tab,
                                                                double *__restrict__ partial) {
  __shared__ double sm[kRedThreads / 32];
  double PE = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += gridDim.x * blockDim.x) {
    const double x_i = x[3 * (size_t)i], y_i = x[3 * (size_t)i + 1], z_i = x[3 * (size_t)i + 2];
    const int type_i = ONETYPE ? 0 : type[i];
    const int *row;
    int num_neighs;
    row_of(list, i, row, num_neighs);
    for (int jj = 0; jj < num_neighs; jj++) {
      const int j = row[jj];
      const double dx = x_i - x[3 * (size_t)j], dy = y_i - x[3 * (size_t)j + 1], dz = z_i - x[3 * (size_t)j + 2];
      const int tij = ONETYPE ? 0 : type_i * tab.ntypes + type[j];
      const double rsq = dx * dx + dy * dy + dz * dz;
      const double cutsq_ij = tab.cutsq[tij];
      if (rsq < cutsq_ij) {
        const double lj1_ij = tab.lj1[tij], lj2_ij = tab.lj2[tij];
        const double r2inv = 1.0 / rsq, r6inv = r2inv * r2inv * r2inv;
        const double r2invc = 1.0 / cutsq_ij, r6invc = r2invc * r2invc * r2invc;
        const double fac = HALF ? ((j < n_local) ? 1.0 : 0.5) : 0.5; // :271, :330-332
        PE += fac * r6inv * (0.5 * lj1_ij * r6inv - lj2_ij) / 6.0;
        PE -= fac * r6invc * (0.5 * lj1_ij * r6invc - lj2_ij) / 6.0; // shift_flag, :274-278
      }
    }
  }
  const double t = block_sum(PE, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kRedThreads) final_sum_kernel(const double *__restrict__ partial, int n, double *out) {
  __shared__ double sm[kRedThreads / 32];
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += kRedThreads) s += partial[k];
  const double t = block_sum(s, sm);
  if (threadIdx.x == 0) *out = t;
}

__global__ void __launch_bounds__(256) zero_rows_kernel(double *f, long long begin, long long end) {
  const long long e = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < end) f[e] = 0.0;
}

void fill_table(const emd_ctx *ctx, LJTable &t) {
  t.ntypes = ctx->lj.ntypes;
  memcpy(t.lj1, ctx->lj.lj1, sizeof t.lj1);
  memcpy(t.lj2, ctx->lj.lj2, sizeof t.lj2);
  memcpy(t.cutsq, ctx->lj.cutsq, sizeof t.cutsq);
}

} // namespace

namespace emd {
// shared with reduce.cu
int device_sum_partials(emd_ctx *ctx, const double *d_partial, int n, double *h_out) {
  double *d_out = reinterpret_cast<double *>(ctx->s_c.p);
  EMD_LAUNCH(ctx, final_sum_kernel, 1, kRedThreads, 0, d_partial, n, d_out);
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(h_out, ctx->h_pinned, sizeof(double));
  return 0;
}
} // namespace emd

extern "C" {

int emd_force_lj_set_params(emd_ctx *ctx, int ntypes, const double *h_lj1, const double *h_lj2, const double *h_cutsq) {
  if (ntypes < 1 || ntypes > kMaxTypesConst) {
    set_error("emd_force_lj_set_params: ntypes=%d unsupported (1..%d)", ntypes, kMaxTypesConst);
    return 1;
  }
  ctx->lj.ntypes = ntypes;
  ctx->lj_version++;
  memset(ctx->lj.lj1, 0, sizeof ctx->lj.lj1);
  memset(ctx->lj.lj2, 0, sizeof ctx->lj.lj2);
  memset(ctx->lj.cutsq, 0, sizeof ctx->lj.cutsq);
  memcpy(ctx->lj.lj1, h_lj1, sizeof(double) * ntypes * ntypes);
  memcpy(ctx->lj.lj2, h_lj2, sizeof(double) * ntypes * ntypes);
  memcpy(ctx->lj.cutsq, h_cutsq, sizeof(double) * ntypes * ntypes);
  return 0;
}

int emd_force_lj_compute(emd_ctx *ctx, const double *d_x, const int *d_type, double *d_f, int n_local, int n_all,
                         const emd_neigh_list *list, int half, int zero_f) {
  if (ctx->lj.ntypes == 0) { set_error("emd_force_lj_compute: parameters not set"); return 1; }
  if (n_local <= 0) return 0;
  LJTable tab;
  fill_table(ctx, tab);
  const bool one = ctx->lj.ntypes == 1;
  if (zero_f) {
    // half: every row receives scattered contributions -> zero all; full: rows [0,n_local) are overwritten
    const long long b = half ? 0 : 3LL * n_local, e = 3LL * n_all;
    if (e > b) EMD_LAUNCH(ctx, zero_rows_kernel, grid_for(e - b, 256), 256, 0, d_f, b, e);
  }
  const int grid = grid_for(n_local, 128);
  if (half) {
    if (one) EMD_LAUNCH(ctx, (lj_force_kernel<true, true>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, 0);
    else EMD_LAUNCH(ctx, (lj_force_kernel<true, false>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, 0);
  } else {
    if (one) EMD_LAUNCH(ctx, (lj_force_kernel<false, true>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, zero_f);
    else EMD_LAUNCH(ctx, (lj_force_kernel<false, false>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, zero_f);
  }
  return 0;
}

int emd_force_lj_idial_compute(emd_ctx *ctx, const double *d_x, const int *d_type, double *d_f, int n_local, int n_all,
                               const emd_neigh_list *list, int half, int zero_f, const double *h_intensity) {
  if (ctx->lj.ntypes == 0) { set_error("emd_force_lj_idial_compute: parameters not set"); return 1; }
  if (!h_intensity) { set_error("emd_force_lj_idial_compute: intensity table missing"); return 1; }
  if (n_local <= 0) return 0;
  LJTable tab;
  fill_table(ctx, tab);
  IDialTable dial;
  memset(&dial, 0, sizeof dial);
  memcpy(dial.intensity, h_intensity, sizeof(double) * ctx->lj.ntypes * ctx->lj.ntypes);
  if (zero_f) {
    const long long b = half ? 0 : 3LL * n_local, e = 3LL * n_all;
    if (e > b) EMD_LAUNCH(ctx, zero_rows_kernel, grid_for(e - b, 256), 256, 0, d_f, b, e);
  }
  const int grid = grid_for(n_local, 128);
  if (half) EMD_LAUNCH(ctx, (lj_idial_force_kernel<true>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, dial, 0);
  else EMD_LAUNCH(ctx, (lj_idial_force_kernel<false>), grid, 128, 0, d_x, d_type, d_f, n_local, *list, tab, dial, zero_f);
  return 0;
}

int emd_force_lj_energy(emd_ctx *ctx, const double *d_x, const int *d_type, int n_local, const emd_neigh_list *list,
                        int half, double *h_pe) {
  if (ctx->lj.ntypes == 0) { set_error("emd_force_lj_energy: parameters not set"); return 1; }
  LJTable tab;
  fill_table(ctx, tab);
  const bool one = ctx->lj.ntypes == 1;
  const int grid = max(1, min(grid_for(n_local, kRedThreads), ctx->num_sms * 8));
  if (ctx->s_c.ensure(sizeof(double) * ((size_t)grid + 8))) return 1;
  double *partial = ctx->s_c.as<double>() + 8;
  if (half) {
    if (one) EMD_LAUNCH(ctx, (lj_energy_kernel<true, true>), grid, kRedThreads, 0, d_x, d_type, n_local, *list, tab, partial);
    else EMD_LAUNCH(ctx, (lj_energy_kernel<true, false>), grid, kRedThreads, 0, d_x, d_type, n_local, *list, tab, partial);
  } else {
    if (one) EMD_LAUNCH(ctx, (lj_energy_kernel<false, true>), grid, kRedThreads, 0, d_x, d_type, n_local, *list, tab, partial);
    else EMD_LAUNCH(ctx, (lj_energy_kernel<false, false>), grid, kRedThreads, 0, d_x, d_type, n_local, *list, tab, partial);
  }
  return device_sum_partials(ctx, partial, grid, h_pe);
}

} // extern "C"
