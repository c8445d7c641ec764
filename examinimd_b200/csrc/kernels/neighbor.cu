// neighbor.cu -- neighbor-list builds (CSR and 2D, full and half).
// Replaces NeighborCSR<>::create_neigh_list (src/neighbor_types/neighbor_csr.h:370-435, functors
// :176-368) and Neighbor2D<>::create_neigh_list (src/neighbor_types/neighbor_2d.h:280-331,
// functors :175-278).
//
// One CTA (4 warps) per interior bin.  The coordinates and indices of the bin's 27-bin stencil
// (~570 candidates at the LJ density) are staged ONCE in shared memory in the reference's
// serial traversal order (bx-1..bx+1, by, bz; permute order inside a bin); each warp then owns
// atoms of the centre bin and sweeps the staged candidates 32 at a time: distance test in FP64
// with separate multiplies/adds (bit-identical to the CPU reference's inclusion decision),
// warp ballot, popcount for the count pass, ballot-prefix compaction for the fill pass.  Rows are
// therefore written in ascending candidate order = the 1-thread reference's row order.
// Global traffic per local atom: 28 B staged once per 27 bins (vs 2 x 540 x 28 B of gathers in
// the reference), + 4 B/entry written.  Bound: shared-memory bandwidth + FP64 compares.
#include "common.cuh"

using namespace emd;

namespace {

constexpr int kNeighThreads = 128;
constexpr int kNeighWarps = kNeighThreads / 32;
constexpr int kStageCap = 1024; // candidates per staging tile (28 KB); typical stencil has ~570

enum { MODE_COUNT = 0, MODE_FILL_CSR = 1, MODE_FILL_2D = 2 };

struct NeighArgs {
  const double *x;
  int n_local;
  int nbx, nby, nbz, nhalo; // full grid incl. halo
  const int *bincount, *binoffsets, *permute;
  double cutsq;
  int newton;
  // outputs
  int *counts;        // COUNT: counts[i]; FILL_2D: num_neighs[i]
  const int *row_map; // FILL_CSR
  int *entries;       // FILL_CSR: entries; FILL_2D: neighs2d
  int maxneighs;      // FILL_2D row stride / capacity
  int *max_count;     // FILL_2D: atomicMax of row counts
};

template <bool HALF, int MODE>
__global__ void __launch_bounds__(kNeighThreads) neigh_kernel(NeighArgs a) {
  __shared__ double sx[kStageCap], sy[kStageCap], sz[kStageCap];
  __shared__ int sj[kStageCap];
  __shared__ int s_cnt[27], s_off[27], s_pre[28];

  const int nix = a.nbx - 2 * a.nhalo, niy = a.nby - 2 * a.nhalo, niz = a.nbz - 2 * a.nhalo;
  (void)nix;
  const int lr = blockIdx.x; // league rank over interior bins, neighbor_csr.h:178-180
  const int bx = lr / (niy * niz) + a.nhalo;
  const int by = (lr / niz) % niy + a.nhalo;
  const int bz = lr % niz + a.nhalo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (threadIdx.x < 27) {
    const int k = threadIdx.x;
    const int bxj = bx - 1 + k / 9, byj = by - 1 + (k / 3) % 3, bzj = bz - 1 + k % 3;
    const int c = (bxj * a.nby + byj) * a.nbz + bzj;
    s_cnt[k] = a.bincount[c];
    s_off[k] = a.binoffsets[c];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = 0; k < 27; k++) { s_pre[k] = acc; acc += s_cnt[k]; }
    s_pre[27] = acc;
  }
  __syncthreads();
  const int M = s_pre[27];
  const int ni = s_cnt[13];     // atoms of the centre bin
  const int i_off = s_off[13];
  if (ni == 0) return;

  for (int tile0 = 0; tile0 < M; tile0 += kStageCap) {
    const int tile1 = min(M, tile0 + kStageCap);
    if (tile0 > 0) __syncthreads();
    // stage candidates [tile0,tile1): warp w takes stencil bins w, w+4, ...
    for (int k = warp; k < 27; k += kNeighWarps) {
      const int lo = max(s_pre[k], tile0), hi = min(s_pre[k + 1], tile1);
      for (int c = lo + lane; c < hi; c += 32) {
        const int j = a.permute[s_off[k] + (c - s_pre[k])];
        const int s = c - tile0;
        sj[s] = j;
        sx[s] = a.x[3 * (size_t)j];
        sy[s] = a.x[3 * (size_t)j + 1];
        sz[s] = a.x[3 * (size_t)j + 2];
      }
    }
    __syncthreads();
    const int tm = tile1 - tile0;

    for (int bi = warp; bi < ni; bi += kNeighWarps) {
      const int i = a.permute[i_off + bi];
      if (i >= a.n_local) continue; // neighbor_csr.h:184
      const double x_i = a.x[3 * (size_t)i], y_i = a.x[3 * (size_t)i + 1], z_i = a.x[3 * (size_t)i + 2];
      // running row length; carried through a.counts[i] when a stencil needs more than one tile
      int count = (tile0 > 0) ? a.counts[i] : 0;
      const int base = (MODE == MODE_FILL_CSR) ? a.row_map[i] : 0;
      for (int c0 = 0; c0 < tm; c0 += 32) {
        const int c = c0 + lane;
        bool hit = false;
        int j = 0;
        if (c < tm) {
          j = sj[c];
          const double x_j = sx[c], y_j = sy[c], z_j = sz[c];
          bool skip;
          if (HALF) // neighbor_csr.h:290-291
            skip = ((j == i) || (j < a.n_local || a.newton)) &&
                   !((x_j > x_i) || ((x_j == x_i) && ((y_j > y_i) || ((y_j == y_i) && (z_j > z_i)))));
          else
            skip = (i == j); // :206
          if (!skip) {
            const double dx = x_i - x_j, dy = y_i - y_j, dz = z_i - z_j;
            // dx*dx + dy*dy + dz*dz with no FMA contraction (matches the x86-64 reference build)
            const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            hit = rsq <= a.cutsq;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (MODE != MODE_COUNT && hit) {
          const int pos = count + __popc(m & ((1u << lane) - 1u));
          if (MODE == MODE_FILL_CSR) a.entries[(size_t)base + pos] = j;
          else if (pos < a.maxneighs) a.entries[(size_t)i * a.maxneighs + pos] = j; // neighbor_2d.h:207-208
        }
        count += __popc(m);
      }
      if (lane == 0) {
        if (MODE == MODE_FILL_CSR) { if (tile1 < M) a.counts[i] = count; }
        else {
          a.counts[i] = count;
          if (MODE == MODE_FILL_2D && tile1 == M) atomicMax(a.max_count, count);
        }
      }
    }
  }
}

template <int MODE>
int launch_neigh(emd_ctx *ctx, const NeighArgs &a, int half) {
  const int nbins = (a.nbx - 2 * a.nhalo) * (a.nby - 2 * a.nhalo) * (a.nbz - 2 * a.nhalo);
  if (nbins <= 0) return 0;
  if (half) EMD_LAUNCH(ctx, (neigh_kernel<true, MODE>), nbins, kNeighThreads, 0, a);
  else EMD_LAUNCH(ctx, (neigh_kernel<false, MODE>), nbins, kNeighThreads, 0, a);
  return 0;
}

NeighArgs make_args(const double *d_x, int n_local, const emd_bin_geom *g, const int *bc, const int *bo, const int *pv,
                    double cut, int newton) {
  NeighArgs a;
  memset(&a, 0, sizeof a);
  a.x = d_x; a.n_local = n_local;
  a.nbx = g->nbinx; a.nby = g->nbiny; a.nbz = g->nbinz; a.nhalo = g->nhalo;
  a.bincount = bc; a.binoffsets = bo; a.permute = pv;
  a.cutsq = cut * cut; // neigh_cut*neigh_cut, neighbor_csr.h:206
  a.newton = newton;
  return a;
}

} // namespace

extern "C" {

int emd_neigh_csr_count(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *g, const int *d_bincount,
                        const int *d_binoffsets, const int *d_permute, double neigh_cut, int half, int newton,
                        int *d_row_map, int *h_total) {
  NeighArgs a = make_args(d_x, n_local, g, d_bincount, d_binoffsets, d_permute, neigh_cut, newton);
  a.counts = d_row_map;
  EMD_CUDA(cudaMemsetAsync(d_row_map, 0, sizeof(int) * ((size_t)n_local + 1), ctx->stream)); // neighbor_csr.h:387
  if (launch_neigh<MODE_COUNT>(ctx, a, half)) return 1;
  // create_offsets (:359-368): exclusive scan over n_local+1 so that row_map[n_local] = total
  if (exclusive_scan_int(ctx, d_row_map, d_row_map, n_local + 1, nullptr)) return 1;
  if (h_total) { // :413-414
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_row_map + n_local, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_total = ctx->h_pinned[0];
    if (*h_total < 0) { set_error("emd_neigh_csr_count: neighbor count overflows 32-bit row_map"); return 2; }
  }
  return 0;
}

int emd_neigh_csr_fill(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *g, const int *d_bincount,
                       const int *d_binoffsets, const int *d_permute, double neigh_cut, int half, int newton,
                       const int *d_row_map, int *d_entries) {
  NeighArgs a = make_args(d_x, n_local, g, d_bincount, d_binoffsets, d_permute, neigh_cut, newton);
  a.row_map = d_row_map;
  a.entries = d_entries;
  // per-row cursor, only touched when a stencil holds more than kStageCap candidates
  if (ctx->s_b.ensure(sizeof(int) * ((size_t)n_local + 1))) return 1;
  a.counts = ctx->s_b.as<int>();
  return launch_neigh<MODE_FILL_CSR>(ctx, a, half);
}

int emd_neigh_2d_fill(emd_ctx *ctx, const double *d_x, int n_local, const emd_bin_geom *g, const int *d_bincount,
                      const int *d_binoffsets, const int *d_permute, double neigh_cut, int half, int newton,
                      int maxneighs, int *d_num_neighs, int *d_neighs, int *h_max_count) {
  NeighArgs a = make_args(d_x, n_local, g, d_bincount, d_binoffsets, d_permute, neigh_cut, newton);
  a.counts = d_num_neighs;
  a.entries = d_neighs;
  a.maxneighs = maxneighs;
  if (ctx->s_b.ensure(sizeof(int))) return 1;
  a.max_count = ctx->s_b.as<int>();
  EMD_CUDA(cudaMemsetAsync(a.max_count, 0, sizeof(int), ctx->stream));
  EMD_CUDA(cudaMemsetAsync(d_num_neighs, 0, sizeof(int) * ((size_t)n_local + 1), ctx->stream)); // neighbor_2d.h:310
  if (launch_neigh<MODE_FILL_2D>(ctx, a, half)) return 1;
  if (h_max_count) {
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, a.max_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_max_count = ctx->h_pinned[0];
  }
  return 0;
}

} // extern "C"
