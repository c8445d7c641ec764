// binning.cu -- cell binning as a device counting sort.
// Replaces BinningKKSort::create_binning (src/binning_types/binning_kksort.cpp:71-140), i.e.
// Kokkos::BinSort<t_x_const, BinOp3D>::create_permute_vector + ::sort (kokkos/kokkos >= 3.0,
// algorithms/src/Kokkos_Sort.hpp).
//
// Kernels (n = atoms in range, B = nbinx*nbiny*nbinz bins):
//   bin_count   : bin id per atom (stored as int) + int atomic histogram      24n R, 4n W, n atomics
//   scan        : exclusive scan of the histogram (ctx.cu)                    ~12B bytes
//   bin_place   : claim a slot in the atom's bin with an int atomic           4n R, 4n W (scattered)
//   bin_order   : per-bin insertion sort of the ~20 claimed indices -> ascending index order,
//                 which is the reference's 1-thread arrival order, so the permutation is
//                 deterministic and bit-identical to the serial reference    ~8n bytes
//   permute     : one gather for x,v,f,type,id,q (88 B/atom read + 88 B/atom written)
// All are HBM/L2-bound integer/byte kernels; there is nothing GEMM-shaped here.
#include "common.cuh"

using namespace emd;

namespace {

struct BinOp {
  double mul[3], mn[3];
  int nb[3];
};

// BinOp3D::bin: ((int(mul0*(x-min0))*nb1 + int(mul1*(y-min1)))*nb2) + int(mul2*(z-min2)).
// (x-min) then *mul cannot be contracted into an FMA, so the result equals the CPU's.
__device__ __forceinline__ int bin_of(const BinOp &op, double x, double y, double z, bool *ok) {
  const int ix = (int)(op.mul[0] * (x - op.mn[0]));
  const int iy = (int)(op.mul[1] * (y - op.mn[1]));
  const int iz = (int)(op.mul[2] * (z - op.mn[2]));
  *ok = ix >= 0 && ix < op.nb[0] && iy >= 0 && iy < op.nb[1] && iz >= 0 && iz < op.nb[2];
  return (ix * op.nb[1] + iy) * op.nb[2] + iz;
}

__global__ void __launch_bounds__(256) bin_count_kernel(const double *__restrict__ x, int n, BinOp op,
                                                        int *__restrict__ binid, int *__restrict__ bincount,
                                                        int *__restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok;
  const int b = bin_of(op, x[3 * (size_t)i], x[3 * (size_t)i + 1], x[3 * (size_t)i + 2], &ok);
  if (!ok) { atomicExch(err, 1 + i); binid[i] = -1; return; }
  binid[i] = b;
  atomicAdd(&bincount[b], 1);
}

__global__ void __launch_bounds__(256) bin_place_kernel(const int *__restrict__ binid, int n,
                                                        const int *__restrict__ binoffsets,
                                                        int *__restrict__ cursor, int *__restrict__ permute) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = binid[i];
  if (b < 0) return;
  const int slot = atomicAdd(&cursor[b], 1);
  permute[binoffsets[b] + slot] = i;
}

// warp per bin: the bin's slice of the permute vector in ascending atom index (= the stable order of the reference's
// one-thread arrival).  The keys are distinct, so the rank of a key is the number of smaller keys: every lane holds up to
// kOrderKeys keys in registers, all keys pass by through shuffles, and each key is stored at its rank (no divergence,
// O(n^2 / 32) per bin instead of a one-thread insertion sort).  Bins beyond 32 * kOrderKeys atoms (never in the decks:
// ~20 atoms per bin) are sorted by lane 0.
constexpr int kOrderKeys = 8;
__global__ void __launch_bounds__(256) bin_order_kernel(int nbins, const int *__restrict__ bincount,
                                                        const int *__restrict__ binoffsets, int *permute) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= nbins) return;
  const int cnt = bincount[b];
  if (cnt < 2) return;
  int *p = permute + binoffsets[b];
  if (cnt > 32 * kOrderKeys) {
    if (lane == 0)
      for (int a = 1; a < cnt; a++) {
        const int key = p[a];
        int k = a - 1;
        while (k >= 0 && p[k] > key) { p[k + 1] = p[k]; k--; }
        p[k + 1] = key;
      }
    return;
  }
  if (cnt <= 32) { // the usual bin: one key per lane
    const int key = lane < cnt ? p[lane] : 0x7fffffff;
    int rank = 0;
#pragma unroll
    for (int src = 0; src < 32; src++) rank += __shfl_sync(0xffffffffu, key, src) < key;
    __syncwarp();
    if (lane < cnt) p[rank] = key;
    return;
  }
  const int nk = (cnt + 31) >> 5;
  int key[kOrderKeys], rank[kOrderKeys];
#pragma unroll
  for (int q = 0; q < kOrderKeys; q++) {
    key[q] = (q < nk && lane + 32 * q < cnt) ? p[lane + 32 * q] : 0x7fffffff;
    rank[q] = 0;
  }
#pragma unroll
  for (int q2 = 0; q2 < kOrderKeys; q2++) {
    if (q2 < nk) { // warp-uniform
      for (int src = 0; src < 32; src++) {
        const int other = __shfl_sync(0xffffffffu, key[q2], src);
#pragma unroll
        for (int q = 0; q < kOrderKeys; q++) rank[q] += other < key[q];
      }
    }
  }
  __syncwarp(); // every key has been read
#pragma unroll
  for (int q = 0; q < kOrderKeys; q++)
    if (q < nk && lane + 32 * q < cnt) p[rank[q]] = key[q];
}

__global__ void __launch_bounds__(256) permute_kernel(const int *__restrict__ permute, int n,
                                                      const double *__restrict__ x_in, const double *__restrict__ v_in,
                                                      const double *__restrict__ f_in, const int *__restrict__ type_in,
                                                      const int *__restrict__ id_in, const double *__restrict__ q_in,
                                                      double *__restrict__ x_out, double *__restrict__ v_out,
                                                      double *__restrict__ f_out, int *__restrict__ type_out,
                                                      int *__restrict__ id_out, double *__restrict__ q_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t s = (size_t)permute[i], d = (size_t)i;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    x_out[3 * d + k] = x_in[3 * s + k];
    v_out[3 * d + k] = v_in[3 * s + k];
    f_out[3 * d + k] = f_in[3 * s + k];
  }
  type_out[d] = type_in[s];
  id_out[d] = id_in[s];
  q_out[d] = q_in[s];
}

} // namespace

extern "C" {

// binning_kksort.cpp:77-99 -- the same double expressions in the same order
int emd_binning_geometry(const double sub[3], const double lo[3], const double hi[3], double dx_in, double dy_in,
                         double dz_in, int halo_depth, emd_bin_geom *g) {
  if (!g) { set_error("emd_binning_geometry: out == NULL"); return 1; }
  g->nhalo = halo_depth;
  g->nbinx = (int)(sub[0] / dx_in);
  g->nbiny = (int)(sub[1] / dy_in);
  g->nbinz = (int)(sub[2] / dz_in);
  if (g->nbinx == 0) g->nbinx = 1;
  if (g->nbiny == 0) g->nbiny = 1;
  if (g->nbinz == 0) g->nbinz = 1;
  const double dx = sub[0] / g->nbinx, dy = sub[1] / g->nbiny, dz = sub[2] / g->nbinz;
  g->nbinx += 2 * halo_depth;
  g->nbiny += 2 * halo_depth;
  g->nbinz += 2 * halo_depth;
  const double eps = dx / 1000;
  g->minx = -dx * halo_depth - eps + lo[0];
  g->maxx = dx * halo_depth + eps + hi[0];
  g->miny = -dy * halo_depth - eps + lo[1];
  g->maxy = dy * halo_depth + eps + hi[1];
  g->minz = -dz * halo_depth - eps + lo[2];
  g->maxz = dz * halo_depth + eps + hi[2];
  return 0;
}

int emd_binning_build(emd_ctx *ctx, const double *d_x, int n, const emd_bin_geom *g, int *d_bincount, int *d_binoffsets,
                      int *d_permute) {
  const long long nbins_ll = (long long)g->nbinx * g->nbiny * g->nbinz;
  if (nbins_ll <= 0 || nbins_ll > 0x7fffffffLL) { set_error("emd_binning_build: bad bin grid"); return 1; }
  const int nbins = (int)nbins_ll;
  BinOp op;
  op.nb[0] = g->nbinx; op.nb[1] = g->nbiny; op.nb[2] = g->nbinz;
  op.mn[0] = g->minx; op.mn[1] = g->miny; op.mn[2] = g->minz;
  op.mul[0] = (double)g->nbinx / (g->maxx - g->minx); // BinOp3D ctor
  op.mul[1] = (double)g->nbiny / (g->maxy - g->miny);
  op.mul[2] = (double)g->nbinz / (g->maxz - g->minz);

  // scratch: binid[n] | cursor[nbins] | err[1]
  if (ctx->s_a.ensure(sizeof(int) * ((size_t)n + (size_t)nbins + 1))) return 1;
  int *binid = ctx->s_a.as<int>();
  int *cursor = binid + n;
  int *err = cursor + nbins;
  EMD_CUDA(cudaMemsetAsync(d_bincount, 0, sizeof(int) * (size_t)nbins, ctx->stream));
  EMD_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int) * ((size_t)nbins + 1), ctx->stream));
  if (n > 0) EMD_LAUNCH(ctx, bin_count_kernel, grid_for(n, 256), 256, 0, d_x, n, op, binid, d_bincount, err);
  if (exclusive_scan_int(ctx, d_bincount, d_binoffsets, nbins, nullptr)) return 1;
  if (n > 0) {
    EMD_LAUNCH(ctx, bin_place_kernel, grid_for(n, 256), 256, 0, binid, n, d_binoffsets, cursor, d_permute);
    EMD_LAUNCH(ctx, bin_order_kernel, grid_for((size_t)nbins * 32, 256), 256, 0, nbins, d_bincount, d_binoffsets, d_permute);
  }
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->h_pinned[0] != 0) {
    set_error("emd_binning_build: atom %d lies outside the bin grid (lost atom)", ctx->h_pinned[0] - 1);
    return 2;
  }
  return 0;
}

int emd_binning_permute(emd_ctx *ctx, const int *d_permute, int n, const double *d_x_in, const double *d_v_in,
                        const double *d_f_in, const int *d_type_in, const int *d_id_in, const double *d_q_in,
                        double *d_x_out, double *d_v_out, double *d_f_out, int *d_type_out, int *d_id_out,
                        double *d_q_out) {
  if (n <= 0) return 0;
  EMD_LAUNCH(ctx, permute_kernel, grid_for(n, 256), 256, 0, d_permute, n, d_x_in, d_v_in, d_f_in, d_type_in, d_id_in,
             d_q_in, d_x_out, d_v_out, d_f_out, d_type_out, d_id_out, d_q_out);
  return 0;
}

} // extern "C"
