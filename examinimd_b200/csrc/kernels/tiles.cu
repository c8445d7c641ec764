// tiles.cu -- the B200 fast path of the neighbor build and the LJ force: cell TILES staged in
// shared memory.
//
// Why (measured on B200, profiles/r01_microbench_b200.json): a thread-per-atom walk over a CSR
// list is bound by the L1/L2 gather of x[j] (184 G rows/s) and, for Newton-3 half lists, by
// REDG.F64 (224 Gop/s; shared-memory FP64 atomics are CAS loops) -- 14 % FP64-pipe use at
// 0.96 ms for 2 M atoms.  FP64 FMA issue (36.8 TFLOP/s) is the cheap resource on this part, so
// the fast path recomputes each pair from both sides instead of scattering -f to j, and serves
// every x[j] from shared memory:
//
//   * the interior bins of BinningKKSort's grid are grouped into tiles of tx*ty*tz cells; one
//     CTA owns one tile and stages the coordinates of the (tx+2)(ty+2)(tz+2) cells around it
//     ONCE (coalesced: atoms are cell-sorted), ~6x re-read instead of 78 gathers per atom;
//   * the neighbor build emits, next to the reference's CSR/2D list, a tile-local FULL adjacency
//     in ELL layout (uint16 slot numbers into the staged array, 4 entries packed per 8-byte
//     word, atom-major so a warp reads 256 contiguous bytes per 4 neighbors).  It is filled by
//     an FP32 pre-filter with a conservative radius (warp = one cell, broadcast LDS of the
//     candidate, 7 FP32 ops per pair instead of 8 FP64); the exact FP64 inclusion test of the
//     reference (rsq <= cut*cut, no FMA contraction; half-list owner rule) is applied when the
//     CSR/2D list is emitted from it, so the API-visible list is bit-identical to the reference
//     and the ELL list is a superset that differs only by pairs within 1e-4 of the list radius
//     (which the force kernel's own cutoff test rejects);
//   * the force kernel walks the ELL rows with x[j] from shared memory, f_i in registers, one
//     coalesced store per atom, no atomics, no zero-f pass; the summation order is the
//     reference's serial row order.
//
// Replaces, when applicable (list radius <= bin width, tile fits in shared memory):
//   NeighborCSR/2D::create_neigh_list  src/neighbor_types/neighbor_csr.h:370-435, neighbor_2d.h:280-331
//   ForceLJNeigh::compute/_energy      src/force_types/force_lj_neigh_impl.h:100-156
// The generic kernels (neighbor.cu, force_lj.cu) remain the path for half lists with newton on
// (ghost forces are needed there) and for tiles that do not fit.
#include "common.cuh"
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>

using namespace emd;

namespace {

constexpr int kMaxStagedCells = 400;
constexpr int kMaxInteriorCells = 128;
constexpr int kFilterThreads = 256;

struct TileArgs {
  int nbx, nby, nbz, nhalo; // bin grid incl. halo bins
  int tx, ty, tz;           // interior cells per tile
  int ntx, nty, ntz;        // tiles per dimension
  int stride;               // atom slots (= threads of the per-atom kernels) per tile, multiple of 32
  int maxrow;               // ELL row capacity, multiple of 8
  int cap;                  // staged-atom capacity of the shared-memory arrays
  int n_local;
  const int *bincount, *binoffsets, *permute;
  const double *x;
  const int *type;
  double ox, oy, oz; // bin grid origin
  double wx, wy, wz; // bin widths
  unsigned short *ell; // [ntiles][maxrow/8][stride][8]: 8 entries of one atom = one 16-byte word
  int *nell;           // [ntiles][stride]
  // the same adjacency re-ordered for the force kernel (tiles_schedule_kernel): every warp's rows are laid
  // out in COLUMNS that a half-warp can read from shared memory without bank conflicts.  Entries with bit 15
  // set are padding (the slot in the low bits is one another lane of the half-warp reads in that column).
  unsigned short *ell_s; // [ntiles][maxrow_s/8][stride][8]
  int *nell_s;           // [ntiles][stride]: columns of the thread's warp, multiple of 8
  int maxrow_s;          // column capacity, multiple of 8
  // per-tile staging tables written once per build (tiles_tables_kernel), so that the per-step
  // force kernel needs no cell arithmetic: global index of every staged slot, and for every
  // dense atom slot its staged slot / global index
  int *stg_j;                // [ntiles][cap]
  int *stg_n;                // [ntiles]  staged atoms
  unsigned short *int_slot;  // [ntiles][stride]
  int *int_glob;             // [ntiles][stride]  (>= n_local or 0x7fffffff: no row)
  // tiles whose staged cells hold only owned atoms come first in `order` (their forces do not depend on the halo, so a
  // decomposed run computes them while the halo exchange is in flight); has_ghost[tile] is the classification
  int *order;          // [ntiles] tile numbers, halo-independent tiles first (stable)
  int *has_ghost;      // [ntiles]
  int *flags;          // [0] overflow bits (1 staged, 2 interior, 4 row)  [1] max row  [2] max staged  [3] max interior
};

struct TileCtx {
  int tile;
  int bx0, by0, bz0;  // first interior cell of the tile (grid coordinates)
  int sxn, syn, szn;  // staged cells per dimension
  int ncs, nci;       // staged / interior cell counts
  int total, n_int;   // staged atoms / atoms in interior cells
};

// in-place exclusive scan of arr[0..n) by warp 0; arr[n] = total
__device__ __forceinline__ void warp0_exclusive_scan(int *arr, int n) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const int chunk = (n + 31) / 32;
    const int b = lane * chunk, e = min(n, b + chunk);
    int s = 0;
    for (int k = b; k < e; k++) s += arr[k];
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int run = inc - s;
    for (int k = b; k < e; k++) { const int v = arr[k]; arr[k] = run; run += v; }
    if (lane == 31) arr[n] = inc;
  }
}

__device__ __forceinline__ int staged_of_interior(const TileCtx &t, const TileArgs &a, int ci) {
  const int cz = ci % a.tz, cy = (ci / a.tz) % a.ty, cx = ci / (a.tz * a.ty);
  return ((cx + 1) * t.syn + (cy + 1)) * t.szn + (cz + 1);
}

// cell tables of this CTA's tile: s_start[c] = first staged slot of staged cell c, s_goff[c] =
// its offset in the permute vector, s_ibase[ci] = first dense atom slot of interior cell ci
__device__ __forceinline__ void tile_setup(const TileArgs &a, TileCtx &t, int *s_start, int *s_goff, int *s_ibase) {
  t.tile = blockIdx.x;
  const int tZ = t.tile % a.ntz, tY = (t.tile / a.ntz) % a.nty, tX = t.tile / (a.ntz * a.nty);
  t.bx0 = a.nhalo + tX * a.tx; t.by0 = a.nhalo + tY * a.ty; t.bz0 = a.nhalo + tZ * a.tz;
  t.sxn = a.tx + 2; t.syn = a.ty + 2; t.szn = a.tz + 2;
  t.ncs = t.sxn * t.syn * t.szn;
  t.nci = a.tx * a.ty * a.tz;
  for (int c = threadIdx.x; c < t.ncs; c += blockDim.x) {
    const int cz = c % t.szn, cy = (c / t.szn) % t.syn, cx = c / (t.szn * t.syn);
    const int gx = t.bx0 - 1 + cx, gy = t.by0 - 1 + cy, gz = t.bz0 - 1 + cz;
    const bool valid = gx >= 0 && gx < a.nbx && gy >= 0 && gy < a.nby && gz >= 0 && gz < a.nbz;
    const int bin = (gx * a.nby + gy) * a.nbz + gz;
    s_start[c] = valid ? a.bincount[bin] : 0;
    s_goff[c] = valid ? a.binoffsets[bin] : 0;
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < t.nci; ci += blockDim.x) {
    const int cz = ci % a.tz, cy = (ci / a.tz) % a.ty, cx = ci / (a.tz * a.ty);
    const int gx = t.bx0 + cx, gy = t.by0 + cy, gz = t.bz0 + cz;
    const bool interior = gx < a.nbx - a.nhalo && gy < a.nby - a.nhalo && gz < a.nbz - a.nhalo;
    s_ibase[ci] = interior ? s_start[staged_of_interior(t, a, ci)] : 0;
  }
  __syncthreads();
  warp0_exclusive_scan(s_start, t.ncs);
  __syncthreads();
  warp0_exclusive_scan(s_ibase, t.nci);
  __syncthreads();
  t.total = s_start[t.ncs];
  t.n_int = s_ibase[t.nci];
}

enum { ST_F32 = 1, ST_F64 = 2, ST_J = 4, ST_TYPE = 8, ST_AOS = 16 }; // ST_AOS: FP64 coordinates as sx[3*slot + {0,1,2}]

template <int WHAT>
__device__ __forceinline__ void tile_stage(const TileArgs &a, const TileCtx &t, const int *s_start, const int *s_goff,
                                           float4 *sf, double *sx, double *sy, double *sz, int *sj, unsigned char *st) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // FP32 coordinates are relative to the staged region's low corner
  const double cx0 = a.ox + (t.bx0 - 1) * a.wx, cy0 = a.oy + (t.by0 - 1) * a.wy, cz0 = a.oz + (t.bz0 - 1) * a.wz;
  for (int c = warp; c < t.ncs; c += nwarps) {
    const int base = s_start[c], n = s_start[c + 1] - base, goff = s_goff[c];
    for (int k = lane; k < n; k += 32) {
      const int j = a.permute[goff + k];
      const double xj = a.x[3 * (size_t)j], yj = a.x[3 * (size_t)j + 1], zj = a.x[3 * (size_t)j + 2];
      const int s = base + k;
      if (WHAT & ST_F32) sf[s] = make_float4((float)(xj - cx0), (float)(yj - cy0), (float)(zj - cz0), 0.f);
      if (WHAT & ST_F64) { sx[s] = xj; sy[s] = yj; sz[s] = zj; }
      if (WHAT & ST_AOS) { sx[3 * s] = xj; sx[3 * s + 1] = yj; sx[3 * s + 2] = zj; }
      if (WHAT & ST_J) sj[s] = j;
      if (WHAT & ST_TYPE) st[s] = (unsigned char)a.type[j];
    }
  }
}

// dense atom slot t -> staged slot / global index of the interior atoms
__device__ __forceinline__ void tile_interior_table(const TileArgs &a, const TileCtx &t, const int *s_start, const int *s_goff,
                                                    const int *s_ibase, unsigned short *s_islot, int *s_iglob) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int ci = warp; ci < t.nci; ci += nwarps) {
    const int n = s_ibase[ci + 1] - s_ibase[ci];
    if (n == 0) continue;
    const int c = staged_of_interior(t, a, ci);
    for (int k = lane; k < n; k += 32) {
      s_islot[s_ibase[ci] + k] = (unsigned short)(s_start[c] + k);
      if (s_iglob) s_iglob[s_ibase[ci] + k] = a.permute[s_goff[c] + k];
    }
  }
}

__device__ __forceinline__ size_t ell_index(const TileArgs &a, int tile, int q, int t) {
  return (((size_t)tile * (a.maxrow >> 3) + (q >> 3)) * a.stride + t) * 8 + (q & 7);
}

// ---------------------------------------------------------------------------- FP32 pre-filter
// warp = one interior cell (lane = atom), candidates broadcast from shared memory in the
// reference's stencil order (bx-1..bx+1, by, bz; permute order inside a bin)
__global__ void __launch_bounds__(kFilterThreads) tiles_filter_kernel(TileArgs a, float cutf2) {
  __shared__ int s_start[kMaxStagedCells + 1], s_goff[kMaxStagedCells], s_ibase[kMaxInteriorCells + 1];
  extern __shared__ __align__(16) unsigned char dyn[];
  float4 *sf = reinterpret_cast<float4 *>(dyn);
  TileCtx t;
  tile_setup(a, t, s_start, s_goff, s_ibase);
  if (threadIdx.x == 0) {
    atomicMax(&a.flags[2], t.total);
    atomicMax(&a.flags[3], t.n_int);
    if (t.total > a.cap || t.total > 65535) atomicOr(&a.flags[0], 1);
    if (t.n_int > a.stride) atomicOr(&a.flags[0], 2);
  }
  if (t.total > a.cap || t.total > 65535 || t.n_int > a.stride) return;
  for (int k = t.n_int + threadIdx.x; k < a.stride; k += blockDim.x) a.nell[(size_t)t.tile * a.stride + k] = 0;
  tile_stage<ST_F32>(a, t, s_start, s_goff, sf, nullptr, nullptr, nullptr, nullptr, nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int ci = warp; ci < t.nci; ci += nwarps) {
    const int cnt = s_ibase[ci + 1] - s_ibase[ci];
    if (cnt == 0) continue;
    const int c_i = staged_of_interior(t, a, ci);
    for (int k0 = 0; k0 < cnt; k0 += 32) {
      const int k = k0 + lane;
      const bool in_cell = k < cnt;
      const int i_glob = in_cell ? a.permute[s_goff[c_i] + k] : 0x7fffffff;
      const bool active = i_glob < a.n_local; // neighbor_csr.h:184 (ghosts inside interior bins get no row)
      const int own = s_start[c_i] + k;
      const float4 p = in_cell ? sf[own] : make_float4(0.f, 0.f, 0.f, 0.f);
      const int tslot = s_ibase[ci] + k;
      // Accepted slots are shifted into a 64-bit accumulator from the top (after four of them the first sits in the low 16
      // bits) and leave as whole 16-byte words: one store per 8 entries instead of eight 2-byte stores with their index math.
      int q = 0;
      unsigned long long acc = 0ull, lo = 0ull;
      uint4 *const wrow = reinterpret_cast<uint4 *>(a.ell) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + tslot;
      for (int sc = 0; sc < 27; sc++) {
        const int c = c_i + ((sc / 9 - 1) * t.syn + ((sc / 3) % 3 - 1)) * t.szn + (sc % 3 - 1);
        const int e = s_start[c + 1];
#pragma unroll 4
        for (int s = s_start[c]; s < e; s++) {
          const float4 pj = sf[s]; // same address in every lane: broadcast
          const float dx = p.x - pj.x, dy = p.y - pj.y, dz = p.z - pj.z;
          const float r2 = dx * dx + dy * dy + dz * dz;
          if (active && r2 <= cutf2 && s != own) {
            acc = (acc >> 16) | ((unsigned long long)s << 48);
            q++;
            if ((q & 3) == 0) {
              if (q & 4) lo = acc;
              else if (q <= a.maxrow) wrow[(size_t)((q >> 3) - 1) * a.stride] = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)acc, (unsigned)(acc >> 32));
            }
          }
        }
      }
      if (in_cell) {
        const int rem = q & 7;
        if (rem && q < a.maxrow) { // the last, partial word (entries beyond the row length are never read)
          unsigned long long hi = 0ull;
          if (rem < 4) lo = acc >> (16 * (4 - rem));
          else if (rem > 4) hi = acc >> (16 * (8 - rem));
          wrow[(size_t)(q >> 3) * a.stride] = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
        }
        a.nell[(size_t)t.tile * a.stride + tslot] = min(q, a.maxrow);
        if (q > a.maxrow) { atomicOr(&a.flags[0], 4); atomicMax(&a.flags[1], q); }
      }
    }
  }
}

// per-tile staging tables for the force kernel (see TileArgs)
__global__ void __launch_bounds__(kFilterThreads) tiles_tables_kernel(TileArgs a) {
  __shared__ int s_start[kMaxStagedCells + 1], s_goff[kMaxStagedCells], s_ibase[kMaxInteriorCells + 1];
  TileCtx t;
  tile_setup(a, t, s_start, s_goff, s_ibase);
  if (t.total > a.cap || t.n_int > a.stride) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (threadIdx.x == 0) a.stg_n[t.tile] = t.total;
  int ghost = 0, owned_rows = 0;
  for (int c = warp; c < t.ncs; c += nwarps) {
    const int base = s_start[c], n = s_start[c + 1] - base, goff = s_goff[c];
    for (int k = lane; k < n; k += 32) {
      const int j = a.permute[goff + k];
      a.stg_j[(size_t)t.tile * a.cap + base + k] = j;
      ghost |= (j >= a.n_local);
    }
  }
  ghost = __syncthreads_or(ghost);
  if (threadIdx.x == 0) a.has_ghost[t.tile] = ghost ? 1 : 0;
  for (int k = t.n_int + threadIdx.x; k < a.stride; k += blockDim.x) {
    a.int_slot[(size_t)t.tile * a.stride + k] = 0;
    a.int_glob[(size_t)t.tile * a.stride + k] = 0x7fffffff;
  }
  for (int ci = warp; ci < t.nci; ci += nwarps) {
    const int n = s_ibase[ci + 1] - s_ibase[ci];
    if (n == 0) continue;
    const int c = staged_of_interior(t, a, ci);
    for (int k = lane; k < n; k += 32) {
      const int ig = a.permute[s_goff[c] + k];
      a.int_slot[(size_t)t.tile * a.stride + s_ibase[ci] + k] = (unsigned short)(s_start[c] + k);
      a.int_glob[(size_t)t.tile * a.stride + s_ibase[ci] + k] = ig;
      owned_rows += (ig < a.n_local);
    }
  }
  // flags[4]: owned atoms that have a row (= n_local unless an owned atom sits outside the interior bins, which the
  // reference leaves without neighbors, neighbor_csr.h:184; the fused force + integrator launch requires equality)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) owned_rows += __shfl_down_sync(0xffffffffu, owned_rows, o);
  if (lane == 0 && owned_rows) atomicAdd(&a.flags[4], owned_rows);
}

// order[]: halo-independent tiles first, then the others, each group in tile order.  One block: every thread owns a
// contiguous chunk of tiles; flags[2] receives the number of halo-independent tiles.
__global__ void __launch_bounds__(1024) tiles_order_kernel(TileArgs a, int ntiles) {
  __shared__ int s_cnt[1024];
  const int chunk = (ntiles + 1023) / 1024;
  const int b = min(ntiles, (int)threadIdx.x * chunk), e = min(ntiles, b + chunk);
  int mine = 0;
  for (int k = b; k < e; k++) mine += a.has_ghost[k] ? 0 : 1;
  s_cnt[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) { // inclusive scan
    const int v = threadIdx.x >= o ? s_cnt[threadIdx.x - o] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  const int total_free = s_cnt[1023];
  int free_before = s_cnt[threadIdx.x] - mine;
  for (int k = b; k < e; k++) {
    if (a.has_ghost[k]) a.order[total_free + (k - free_before)] = k;
    else a.order[free_before++] = k;
  }
  if (threadIdx.x == 0) a.flags[2] = total_free;
}

// ------------------------------------------------------- bank-conflict-free column schedule
// The force kernel reads x[j] of 32 different neighbors per warp instruction from shared memory
// (LDS.64, served per half-warp: 16 lanes x 8 bytes = one wavefront if the 16 words lie in 16
// different bank pairs or are the same word).  In list order the 16 rows of a half-warp hit
// ~2.5 words per bank pair (ncu: 57 % of the kernel's shared-memory wavefronts were conflicts
// and the LSU pipe, not FP64, bounded it).  Rows are therefore re-ordered once per build:
// column q of a half-warp holds, for every lane, an entry whose bank (slot mod 16: the AoS
// stride 3 is odd, so x, y and z all map slot -> bank pair bijectively) is different from
// every other lane's, or the SAME slot as another lane's (broadcast).  A lane with no such
// entry left idles in that column (a flagged copy of a slot another lane reads).
//
// Greedy, lane after lane (lane 0 of the half-warp picks first): join a slot already taken in
// this column if it is the head of one of my buckets (atoms of one cell share most neighbors),
// else take, among my non-empty buckets whose bank is free, one that holds more than its share
// of what I have left (the bottleneck banks drain first), searching from bank (q + lane) mod 16.
// The lane-serial dependency is a systolic pipeline: at step t lane l works on column t - l and
// reads the state its predecessors left for that column in a shared-memory ring, so a warp
// schedules its two half-warps in (columns + 16) steps.  tools/sim_lds_conflicts.py models both
// re-orderings on a reference liquid state (variants E and F): per 32 pairs and LDS.64, 5.6 wavefronts in list
// order, 3.5 with the per-lane rotation, 2.15 (2 = conflict-free) with this schedule for 1 % more
// columns than the longest row of the warp; the GPU test measures the same on the real lists.
constexpr int kSchedWarps = 2;
constexpr int kSchedMaxRow = 248;        // bucket positions and prefix sums are bytes
constexpr unsigned kNoSlot = 0xffffu;    // head of an empty bucket
constexpr unsigned kFreeBank = 0xfffeu;  // bank not taken in this column

__device__ __forceinline__ unsigned nib_of_bytes(unsigned x) { // byte i non-zero (0xff) -> bit i
  return ((x & 0x08040201u) * 0x01010101u) >> 24;
}
__device__ __forceinline__ unsigned half_eq_mask(const uint4 &h0, const uint4 &h1, const uint4 &t0, const uint4 &t1) {
  // bit b set iff 16-bit element b of h equals element b of t
  const unsigned e0 = __vcmpeq2(h0.x, t0.x), e1 = __vcmpeq2(h0.y, t0.y), e2 = __vcmpeq2(h0.z, t0.z), e3 = __vcmpeq2(h0.w, t0.w);
  const unsigned e4 = __vcmpeq2(h1.x, t1.x), e5 = __vcmpeq2(h1.y, t1.y), e6 = __vcmpeq2(h1.z, t1.z), e7 = __vcmpeq2(h1.w, t1.w);
  return nib_of_bytes(__byte_perm(e0, e1, 0x6420)) | (nib_of_bytes(__byte_perm(e2, e3, 0x6420)) << 4) |
         (nib_of_bytes(__byte_perm(e4, e5, 0x6420)) << 8) | (nib_of_bytes(__byte_perm(e6, e7, 0x6420)) << 12);
}
__device__ __forceinline__ unsigned bytes_gt_mask(const uint4 &c, unsigned th) { // bit b set iff byte b of c > th
  const unsigned t4 = th * 0x01010101u;
  return nib_of_bytes(__vcmpgtu4(c.x, t4)) | (nib_of_bytes(__vcmpgtu4(c.y, t4)) << 4) | (nib_of_bytes(__vcmpgtu4(c.z, t4)) << 8) |
         (nib_of_bytes(__vcmpgtu4(c.w, t4)) << 12);
}

size_t sched_warp_smem(int maxrow, int maxrow_s) {
  return (size_t)32 * maxrow * 2 + (size_t)32 * maxrow_s * 2 + 32 * 16 * 2 /*hs*/ + 32 * 16 /*cnt*/ + 32 * 16 /*hp*/ + 2 * 32 * 16 * 2 /*ring*/ +
         2 * 32 * 2 /*ring masks*/ + 128;
}

// MODE: SCHED_FULL = the conflict-free column schedule above; SCHED_ROT = every lane re-orders its own row so that at
// column q it reads bank (q + lane) mod 16 whenever it still has an entry there (no negotiation between lanes: conflicts
// remain where a bucket ran dry, 3.2 wavefronts per LDS.64 instead of 5.0 / 1.93, at a tenth of the scheduling cost);
// SCHED_NONE = list order (rows longer than kSchedMaxRow), only the flagged padding.
enum { SCHED_FULL = 0, SCHED_NONE = 1, SCHED_ROT = 2 };
template <int MODE>
__global__ void __launch_bounds__(32 * kSchedWarps) tiles_schedule_kernel(TileArgs a, int ntiles, unsigned warp_smem) {
  extern __shared__ __align__(16) unsigned char dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wrows = a.stride >> 5;
  const long long gw = (long long)blockIdx.x * kSchedWarps + warp;
  if (gw >= (long long)ntiles * wrows) return; // whole warps leave; no block-wide barrier below
  const int tile = (int)(gw / wrows), ts = (int)(gw % wrows) * 32 + lane;
  unsigned char *base = dyn + (size_t)warp * warp_smem;
  unsigned short *out = reinterpret_cast<unsigned short *>(base);                 // [lane][maxrow_s]
  unsigned short *srow = out + 32 * a.maxrow_s;                                    // [pos][lane], bank-sorted
  unsigned short *hs = srow + 32 * a.maxrow;                                       // [lane][16] slot at the head of every bucket
  unsigned short *ring = hs + 32 * 16;                                             // [half][32 columns][16] slot taken per bank
  unsigned short *rmask = ring + 2 * 32 * 16;                                      // [half][32] taken banks
  unsigned char *cnt = reinterpret_cast<unsigned char *>(rmask + 2 * 32);          // [lane][16]
  unsigned char *hp = cnt + 32 * 16;                                               // [lane][16] position of the bucket head in srow

  const int n = a.nell[(size_t)tile * a.stride + ts];
  const unsigned own = a.int_slot[(size_t)tile * a.stride + ts];
  const uint4 *row = reinterpret_cast<const uint4 *>(a.ell) + ((size_t)tile * (a.maxrow >> 3)) * a.stride + ts;
  { // every column starts as padding
    const unsigned pad2 = (0x8000u | own) * 0x00010001u;
    uint4 *o = reinterpret_cast<uint4 *>(out + (size_t)lane * a.maxrow_s);
    for (int c = 0; c < (a.maxrow_s >> 3); c++) o[c] = make_uint4(pad2, pad2, pad2, pad2);
  }
  int last = 0;      // columns this lane really uses
  bool ovf = false;
  int need = 0;
  if (MODE == SCHED_NONE) {
    if (n > a.maxrow_s) { ovf = true; need = n; }
    else {
      for (int c = 0; c * 8 < n; c++) {
        const uint4 w = row[(size_t)c * a.stride];
        const unsigned e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (c * 8 + k < n) out[(size_t)lane * a.maxrow_s + c * 8 + k] = (unsigned short)((e[k >> 1] >> (16 * (k & 1))) & 0xffffu);
      }
      last = n;
    }
  } else {
    // ---- counting sort of the row by bank (stable: buckets stay in ascending slot order)
    unsigned long long c_lo = 0, c_hi = 0; // bucket sizes, one byte per bank
    for (int c = 0; c * 8 < n; c++) {
      const uint4 w = row[(size_t)c * a.stride];
      const unsigned e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (c * 8 + k < n) {
          const unsigned b = (e[k >> 1] >> (16 * (k & 1))) & 15u;
          if (b < 8) c_lo += 1ull << (8 * b); else c_hi += 1ull << (8 * (b - 8));
        }
    }
    const unsigned long long M = 0x0101010101010101ull;
    const unsigned long long inc_lo = c_lo * M; // byte k = c_lo[0] + ... + c_lo[k]  (n <= 248: no carries)
    const unsigned long long tot_lo = inc_lo >> 56;
    const unsigned long long inc_hi = c_hi * M + tot_lo * M;
    const unsigned long long st_lo = inc_lo << 8, st_hi = (inc_hi << 8) | tot_lo; // first position of every bucket
    unsigned long long p_lo = st_lo, p_hi = st_hi;
    for (int c = 0; c * 8 < n; c++) {
      const uint4 w = row[(size_t)c * a.stride];
      const unsigned e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (c * 8 + k < n) {
          const unsigned s = (e[k >> 1] >> (16 * (k & 1))) & 0xffffu;
          const unsigned b = s & 15u;
          unsigned pos;
          if (b < 8) { pos = (unsigned)(p_lo >> (8 * b)) & 0xffu; p_lo += 1ull << (8 * b); }
          else { pos = (unsigned)(p_hi >> (8 * (b - 8))) & 0xffu; p_hi += 1ull << (8 * (b - 8)); }
          srow[pos * 32 + lane] = (unsigned short)s;
        }
    }
    unsigned N = 0; // non-empty buckets
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const unsigned cb = (unsigned)((b < 8 ? c_lo >> (8 * b) : c_hi >> (8 * (b - 8))) & 0xffu);
      const unsigned sb = (unsigned)((b < 8 ? st_lo >> (8 * b) : st_hi >> (8 * (b - 8))) & 0xffu);
      cnt[lane * 16 + b] = (unsigned char)cb;
      hp[lane * 16 + b] = (unsigned char)sb;
      hs[lane * 16 + b] = cb ? srow[sb * 32 + lane] : (unsigned short)kNoSlot;
      if (cb) N |= 1u << b;
    }
    int rem = n;
    const int half = lane >> 4, hl = lane & 15;
    unsigned short *rg = ring + half * 32 * 16;
    unsigned short *rm = rmask + half * 32;
    unsigned big = 0;   // banks holding more than th entries
    int th = -1;
    __syncwarp();
    for (int t = 0;; t++) {
      if (!__any_sync(0xffffffffu, rem > 0)) break;
      const int q = t - hl;
      if (q >= 0) {
        unsigned short *tq = rg + (q & 31) * 16;
        if (hl == 0) { // first lane of the half-warp opens the column
          const unsigned f2 = kFreeBank * 0x00010001u;
          reinterpret_cast<uint4 *>(tq)[0] = make_uint4(f2, f2, f2, f2);
          reinterpret_cast<uint4 *>(tq)[1] = make_uint4(f2, f2, f2, f2);
          rm[q & 31] = 0;
        }
        if (rem > 0 && q >= a.maxrow_s) { ovf = true; need = q + rem; rem = 0; }
        if (q < a.maxrow_s) {
          const unsigned tm = rm[q & 31];
          int b = -1;
          bool join = false;
          if (rem > 0) {
            if (tm) {
              const uint4 t0 = reinterpret_cast<const uint4 *>(tq)[0], t1 = reinterpret_cast<const uint4 *>(tq)[1];
              const uint4 h0 = reinterpret_cast<const uint4 *>(hs + lane * 16)[0], h1 = reinterpret_cast<const uint4 *>(hs + lane * 16)[1];
              const unsigned jm = half_eq_mask(h0, h1, t0, t1);
              if (jm) { b = __ffs(jm) - 1; join = true; }
            }
            if (!join) {
              const unsigned m = N & ~tm;
              if (m) {
                const int th_now = (rem + 15) >> 4;
                if (th_now != th) { th = th_now; big = bytes_gt_mask(reinterpret_cast<const uint4 *>(cnt + lane * 16)[0], (unsigned)th); }
                const unsigned mb = m & big;
                const unsigned sel = mb ? mb : m;
                const unsigned r = (unsigned)(q + hl) & 15u;
                const unsigned rot = ((sel >> r) | (sel << (16 - r))) & 0xffffu;
                b = (int)((__ffs(rot) - 1 + r) & 15u);
              }
            }
          }
          if (b >= 0) {
            const unsigned s = hs[lane * 16 + b];
            out[(size_t)lane * a.maxrow_s + q] = (unsigned short)s;
            const unsigned c = (unsigned)cnt[lane * 16 + b] - 1u, p = (unsigned)hp[lane * 16 + b] + 1u;
            cnt[lane * 16 + b] = (unsigned char)c;
            hp[lane * 16 + b] = (unsigned char)p;
            hs[lane * 16 + b] = c ? srow[p * 32 + lane] : (unsigned short)kNoSlot;
            if (!c) N &= ~(1u << b);
            if ((int)c <= th) big &= ~(1u << b);
            rem--;
            last = q + 1;
            if (!join) { tq[b] = (unsigned short)s; rm[q & 31] = (unsigned short)(tm | (1u << b)); }
          } else if (tm) {
            // idle: read what another lane of the half-warp reads in this column (no extra wavefront)
            out[(size_t)lane * a.maxrow_s + q] = (unsigned short)(0x8000u | tq[__ffs(tm) - 1]);
          }
        }
      }
      __syncwarp();
    }
  }
  const int cols = __reduce_max_sync(0xffffffffu, last);
  if (__any_sync(0xffffffffu, ovf)) {
    if (ovf) { atomicOr(&a.flags[0], 8); atomicMax(&a.flags[1], need); }
    return;
  }
  const int c8 = (cols + 7) & ~7;
  __syncwarp();
  uint4 *dst = reinterpret_cast<uint4 *>(a.ell_s) + ((size_t)tile * (a.maxrow_s >> 3)) * a.stride + ts;
  const uint4 *o = reinterpret_cast<const uint4 *>(out + (size_t)lane * a.maxrow_s);
  for (int c = 0; c < (c8 >> 3); c++) dst[(size_t)c * a.stride] = o[c];
  a.nell_s[(size_t)tile * a.stride + ts] = c8;
}

// The default re-ordering (SCHED_ROT), lean version: no negotiation between lanes, so no per-column state.  Per lane: the
// bank-sorted row ([pos][lane]) and one word per bucket ([bank][lane]: head position | entries left << 8) in shared memory,
// both laid out so that a warp access never has a bank conflict; 8 columns are emitted as one coalesced 16-byte word.
constexpr int kRotWarps = 4;
size_t rot_warp_smem(int maxrow) { return (size_t)32 * maxrow * sizeof(unsigned short) + 16 * 32 * sizeof(unsigned); }

__global__ void __launch_bounds__(32 * kRotWarps) tiles_rotate_kernel(TileArgs a, int ntiles, unsigned warp_smem) {
  extern __shared__ __align__(16) unsigned char dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wrows = a.stride >> 5;
  const long long gw = (long long)blockIdx.x * kRotWarps + warp;
  if (gw >= (long long)ntiles * wrows) return;
  const int tile = (int)(gw / wrows), ts = (int)(gw % wrows) * 32 + lane;
  unsigned *state = reinterpret_cast<unsigned *>(dyn + (size_t)warp * warp_smem) + lane;       // state[bank * 32]
  unsigned short *srow = reinterpret_cast<unsigned short *>(dyn + (size_t)warp * warp_smem + 16 * 32 * sizeof(unsigned)) + lane; // srow[pos * 32]
  const int n = a.nell[(size_t)tile * a.stride + ts];
  const unsigned own = a.int_slot[(size_t)tile * a.stride + ts];
  const uint4 *__restrict__ row = reinterpret_cast<const uint4 *>(a.ell) + ((size_t)tile * (a.maxrow >> 3)) * a.stride + ts;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  if (nmax > a.maxrow_s) {
    if (lane == 0) { atomicOr(&a.flags[0], 8); atomicMax(&a.flags[1], nmax); }
    return;
  }
  // counting sort by bank (stable): sizes -> starts -> scatter
#pragma unroll
  for (int b = 0; b < 16; b++) state[b * 32] = 0;
  // (both passes request word c+1 before they work on word c: the row comes from DRAM)
  uint4 wnext = n > 0 ? __ldg(row) : make_uint4(0, 0, 0, 0);
  for (int c = 0; c * 8 < n; c++) {
    const uint4 w = wnext;
    if ((c + 1) * 8 < n) wnext = __ldg(row + (size_t)(c + 1) * a.stride);
    const unsigned e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c * 8 + k < n) state[((e[k >> 1] >> (16 * (k & 1))) & 15u) * 32] += 1;
  }
  unsigned N = 0;
  {
    unsigned run = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const unsigned cb = state[b * 32];
      state[b * 32] = run | (cb << 8);
      if (cb) N |= 1u << b;
      run += cb;
    }
  }
  wnext = n > 0 ? __ldg(row) : make_uint4(0, 0, 0, 0);
  for (int c = 0; c * 8 < n; c++) {
    const uint4 w = wnext;
    if ((c + 1) * 8 < n) wnext = __ldg(row + (size_t)(c + 1) * a.stride);
    const unsigned e[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (c * 8 + k < n) {
        const unsigned s = (e[k >> 1] >> (16 * (k & 1))) & 0xffffu;
        const unsigned st = state[(s & 15u) * 32];
        srow[(st & 0xffu) * 32] = (unsigned short)s;
        state[(s & 15u) * 32] = st + 1;
      }
  }
#pragma unroll
  for (int b = 0; b < 16; b++) { // heads back to the bucket starts
    const unsigned st = state[b * 32];
    state[b * 32] = st - (st >> 8);
  }
  const int hl = lane & 15;
  const unsigned pad = 0x8000u | own;
  unsigned big = 0;
  int th = -1;
  const int c8 = (nmax + 7) & ~7;
  uint4 *dst = reinterpret_cast<uint4 *>(a.ell_s) + ((size_t)tile * (a.maxrow_s >> 3)) * a.stride + ts;
  for (int q0 = 0; q0 < c8; q0 += 8) {
    unsigned wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int q = q0 + k;
      unsigned s = pad;
      if (q < n) {
        const unsigned pref = (unsigned)(q + hl) & 15u;
        unsigned b = pref;
        if (!((N >> pref) & 1u)) { // my bucket for this column's bank ran dry: take from one that holds more than its share
          const int th_now = (n - q + 15) >> 4;
          if (th_now != th) {
            th = th_now;
            big = 0;
#pragma unroll
            for (int bb = 0; bb < 16; bb++) if ((int)(state[bb * 32] >> 8) > th) big |= 1u << bb;
          }
          const unsigned mb = N & big;
          const unsigned sel = mb ? mb : N;
          const unsigned rot = ((sel >> pref) | (sel << (16 - pref))) & 0xffffu;
          b = (unsigned)(__ffs(rot) - 1 + pref) & 15u;
        }
        const unsigned st = state[b * 32];
        s = srow[(st & 0xffu) * 32];
        const unsigned left = (st >> 8) - 1u;
        state[b * 32] = ((st + 1u) & 0xffu) | (left << 8);
        if (!left) N &= ~(1u << b);
        if ((int)left <= th) big &= ~(1u << b);
      }
      wv[k >> 1] |= s << (16 * (k & 1));
    }
    dst[(size_t)(q0 >> 3) * a.stride] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
  }
  a.nell_s[(size_t)tile * a.stride + ts] = c8;
}

// ------------------------------------------------------------ exact lists from the ELL superset
enum { EMIT_COUNT = 0, EMIT_CSR = 1, EMIT_2D = 2 };

struct EmitArgs {
  double cutsq;
  int newton;
  int *counts;        // COUNT: counts[i]; 2D: num_neighs[i]
  const int *row_map; // CSR
  int *entries;       // CSR entries / 2D table
  int maxneighs;      // 2D
  int *max_count;     // 2D
};

template <bool HALF, int MODE>
__global__ void __launch_bounds__(512) tiles_emit_kernel(TileArgs a, EmitArgs e) {
  __shared__ int s_start[kMaxStagedCells + 1], s_goff[kMaxStagedCells], s_ibase[kMaxInteriorCells + 1];
  extern __shared__ __align__(16) unsigned char dyn[];
  double *sx = reinterpret_cast<double *>(dyn), *sy = sx + a.cap, *sz = sy + a.cap;
  int *sj = reinterpret_cast<int *>(sz + a.cap);
  unsigned short *s_islot = reinterpret_cast<unsigned short *>(sj + a.cap);
  TileCtx t;
  tile_setup(a, t, s_start, s_goff, s_ibase);
  if (t.total > a.cap || t.n_int > a.stride) return; // cannot happen after a successful filter pass
  tile_stage<ST_F64 | ST_J>(a, t, s_start, s_goff, nullptr, sx, sy, sz, sj, nullptr);
  tile_interior_table(a, t, s_start, s_goff, s_ibase, s_islot, nullptr);
  __syncthreads();
  const int ts = threadIdx.x;
  if (ts >= t.n_int) return;
  const int own = s_islot[ts];
  const int i = sj[own];
  if (i >= a.n_local) return;
  const int n = a.nell[(size_t)t.tile * a.stride + ts];
  const double x_i = sx[own], y_i = sy[own], z_i = sz[own];
  const size_t base = (MODE == EMIT_CSR) ? (size_t)e.row_map[i] : (size_t)i * e.maxneighs;
  int count = 0;
  // one 16-byte word = 8 entries of the row; the next word is requested before the current one is used
  const uint4 *row = reinterpret_cast<const uint4 *>(a.ell) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + ts;
  const int nchunk = (n + 7) >> 3;
  uint4 cur = nchunk > 0 ? row[0] : make_uint4(0, 0, 0, 0);
  for (int c = 0; c < nchunk; c++) {
    const uint4 nxt = (c + 1 < nchunk) ? row[(size_t)(c + 1) * a.stride] : make_uint4(0, 0, 0, 0);
    const unsigned w4[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (c * 8 + k >= n) break;
      const int s = (int)((w4[k >> 1] >> (16 * (k & 1))) & 0xffffu);
      const int j = sj[s];
      const double x_j = sx[s], y_j = sy[s], z_j = sz[s];
      if (HALF) { // neighbor_csr.h:290-291 (j != i by construction)
        const bool skip = (j < a.n_local || e.newton) &&
                          !((x_j > x_i) || ((x_j == x_i) && ((y_j > y_i) || ((y_j == y_i) && (z_j > z_i)))));
        if (skip) continue;
      }
      const double dx = x_i - x_j, dy = y_i - y_j, dz = z_i - z_j;
      const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      if (rsq <= e.cutsq) { // neighbor_csr.h:206,299
        if (MODE == EMIT_CSR) e.entries[base + count] = j;
        if (MODE == EMIT_2D && count < e.maxneighs) e.entries[base + count] = j; // neighbor_2d.h:207-208
        count++;
      }
    }
    cur = nxt;
  }
  if (MODE == EMIT_COUNT) e.counts[i] = count;
  if (MODE == EMIT_2D) { e.counts[i] = count; atomicMax(e.max_count, count); }
}

// ------------------------------------------------------------------------------ LJ force
struct LJOne { double lj1, lj2, cutsq; };
struct LJTab {
  double lj1[kMaxTypesConst * kMaxTypesConst], lj2[kMaxTypesConst * kMaxTypesConst], cutsq[kMaxTypesConst * kMaxTypesConst];
  int ntypes;
};

// 1/a to <= 1 ulp: MUFU.RCP64H seed y0 (~2^-20) and one cubic step y0 (1 + e + e^2), e = 1 - a y0: three
// DFMA (error e^3 ~ 2^-60) instead of the four of two Newton steps; the library division adds range
// fix-ups that rsq in (0, cutsq) never needs
__device__ __forceinline__ double fast_rcp(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  const double e = fma(-a, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// Four pairs (one ELL word) at a time, branch-free and written stage by stage so that the four
// ~20-instruction FP64 dependency chains are interleaved by the scheduler.
template <bool ONETYPE, bool ENERGY>
__device__ __forceinline__ void lj_quad(const double *__restrict__ sp, const int *__restrict__ st, const unsigned w01, const unsigned w23,
                                        double x_i, double y_i, double z_i, int type_i, const LJOne &one, const LJTab *__restrict__ tab,
                                        double &fx, double &fy, double &fz, double &pe) {
  const unsigned raw[4] = {w01 & 0xffffu, w01 >> 16, w23 & 0xffffu, w23 >> 16};
  double dx[4], dy[4], dz[4], rsq[4], lj1[4], lj2[4], cutsq[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const unsigned sl = raw[u] & 0x7fffu;
    const double *p = sp + 3 * sl;
    dx[u] = x_i - p[0]; dy[u] = y_i - p[1]; dz[u] = z_i - p[2];
    if (ONETYPE) { lj1[u] = one.lj1; lj2[u] = one.lj2; cutsq[u] = one.cutsq; }
    else { const int tij = type_i * tab->ntypes + st[sl]; lj1[u] = tab->lj1[tij]; lj2[u] = tab->lj2[tij]; cutsq[u] = tab->cutsq[tij]; }
  }
#pragma unroll
  for (int u = 0; u < 4; u++) rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
  bool in[4];
  double r2inv[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    // force_lj_neigh_impl.h:189 (strict).  rsq and cutsq are non-negative, so the IEEE order is the order of the bit
    // patterns: the test runs on the integer pipe and leaves the FP64 pipe to the arithmetic.  Bit 15 marks padding.
    in[u] = !(raw[u] & 0x8000u) && __double_as_longlong(rsq[u]) < __double_as_longlong(cutsq[u]);
    r2inv[u] = fast_rcp(rsq[u]); // a padding entry that is the atom itself (rsq = 0) gives inf/NaN below, discarded by the select
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const double r6inv = r2inv[u] * r2inv[u] * r2inv[u];
    const double fpair = in[u] ? (r6inv * (lj1[u] * r6inv - lj2[u])) * r2inv[u] : 0.0;
    fx += dx[u] * fpair; fy += dy[u] * fpair; fz += dz[u] * fpair;
    if (ENERGY) { // force_lj_neigh_impl.h:271-278 with fac = 0.5 (every pair is seen from both sides)
      const double r2invc = 1.0 / cutsq[u], r6invc = r2invc * r2invc * r2invc;
      const double e = 0.5 * r6inv * (0.5 * lj1[u] * r6inv - lj2[u]) / 6.0 - 0.5 * r6invc * (0.5 * lj1[u] * r6invc - lj2[u]) / 6.0;
      pe += in[u] ? e : 0.0;
    }
  }
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
// The ELL words of a row are read once per step, one 16-byte word per 8 pairs.  The register budget (80 at two 384-thread
// CTAs per SM) makes the compiler place the load of word c+1 at the END of iteration c, next to its first use (as a plain
// load AND as volatile asm), and 26 % of the kernel's stall samples were warps waiting for it (profiles/, round 1).  A
// prefetch holds no register: word c+2 is requested into L1 at the top of iteration c, so the late load hits L1.
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kForceThreads = 384;
constexpr int kStagePerThread = 8; // cap <= kForceThreads * kStagePerThread

// Persistent CTAs (2 per SM), each walking tiles blockIdx.x, +gridDim.x, ...  The coordinates of
// tile n+1 are copied global->shared with cp.async (LDGSTS) into the second buffer while tile n
// is computed, and the staging indices of tile n+2 are prefetched into registers, so no global
// latency is exposed between tiles.
// The CTA walks the tiles a.order[first + blockIdx.x], [first + blockIdx.x + gridDim.x], ... below first + ntiles.
// RING: the ELL words travel global -> shared with cp.async into one 16-byte slot per thread (no register is held while
// the copy is in flight, so -- unlike a register prefetch, which ptxas sinks to the end of the loop body at this register
// budget -- the request really is issued a whole 8-pair iteration before its use); word 0 of the next tile's row is
// requested during the last iteration of the current one.  Needs two CTAs to still fit on an SM with the extra 6 KB.
// An experiment that did not pay (see lj_tiles_launch): kept behind EMD_TILES_RING=1.
// FUSE: the epilogue also applies IntegratorNVE::final_integrate of this step and initial_integrate of the next one to the
// atom whose force was just accumulated (src/integrator_nve.cpp:47-74, 87-112: same operations, same order, so v and x are
// bit-identical to the separate kernels): v is updated in place, the new position goes to a SECOND position array because
// other CTAs still stage the old coordinates (the host swaps the two arrays after the launch).
struct NveFuse { double *v; double *x_new; const double *mass; double dtf, dtv; };

template <bool ONETYPE, bool ENERGY, bool RING, bool FUSE>
__global__ void __launch_bounds__(kForceThreads, 2) lj_tiles_kernel(TileArgs a, int first, int ntiles, LJOne one, const LJTab *__restrict__ tab,
                                                                   double *__restrict__ f, double *__restrict__ pe_partial, unsigned ring_offset,
                                                                   NveFuse nve) {
  __shared__ double s_red[kForceThreads / 32];
  extern __shared__ __align__(16) unsigned char dyn[];
  double *const sp0 = reinterpret_cast<double *>(dyn); // 2 x [cap][3]: x,y,z of a staged atom adjacent (one address computation per pair)
  int *const st0 = reinterpret_cast<int *>(sp0 + 6 * (size_t)a.cap); // 2 x [cap] types (multi-type systems only)
  uint4 *const ering = reinterpret_cast<uint4 *>(dyn + ring_offset) + threadIdx.x; // RING: this thread's slot
  const int ts = threadIdx.x;
  const int G = gridDim.x;
  int jreg[kStagePerThread];

  auto tile_at = [&](int pos) { return pos < ntiles ? a.order[first + pos] : -1; }; // pos-th tile of this launch's range
  auto load_j = [&](int tile) { // staging indices of `tile` into registers (coalesced)
    if (tile >= 0) {
      const int n = a.stg_n[tile];
      const int *src = a.stg_j + (size_t)tile * a.cap;
#pragma unroll
      for (int k = 0; k < kStagePerThread; k++) { const int s = ts + k * kForceThreads; jreg[k] = s < n ? src[s] : -1; }
    } else {
#pragma unroll
      for (int k = 0; k < kStagePerThread; k++) jreg[k] = -1;
    }
  };
  auto issue_copies = [&](int buf) {
#pragma unroll
    for (int k = 0; k < kStagePerThread; k++) {
      const int j = jreg[k];
      if (j >= 0) {
        const int s = ts + k * kForceThreads;
        const double *src = a.x + 3 * (size_t)j;
        double *dst = sp0 + (size_t)buf * 3 * a.cap + 3 * s;
        cp_async8(dst, src); cp_async8(dst + 1, src + 1); cp_async8(dst + 2, src + 2);
        if (!ONETYPE) cp_async4(st0 + (size_t)buf * a.cap + s, a.type + j);
      }
    }
  };

  int pos = blockIdx.x;
  int tile = tile_at(pos), tile_nxt = tile_at(pos + G);
  load_j(tile);
  issue_copies(0);
  load_j(tile_nxt);
  // per-thread row descriptors of the current tile
  int i_cur = 0x7fffffff, own_cur = 0, n_cur = 0;
  if (tile >= 0) {
    i_cur = a.int_glob[(size_t)tile * a.stride + ts];
    own_cur = a.int_slot[(size_t)tile * a.stride + ts];
    n_cur = a.nell_s[(size_t)tile * a.stride + ts];
  }
  auto row_of_tile = [&](int tl) { return reinterpret_cast<const uint4 *>(a.ell_s) + ((size_t)tl * (a.maxrow_s >> 3)) * a.stride + ts; };
  if (RING && tile >= 0 && i_cur < a.n_local && n_cur > 0) cp_async16(ering, row_of_tile(tile)); // word 0 of the first row
  double pe = 0.0;
  int buf = 0;
  for (; pos < ntiles; pos += G, buf ^= 1) {
    cp_async_wait_all();
    __syncthreads(); // buffer `buf` is complete; every thread is done with buffer `buf^1`
    issue_copies(buf ^ 1);  // tile_nxt
    const int tile_nn = tile_at(pos + 2 * G);
    load_j(tile_nn);
    int i_nxt = 0x7fffffff, own_nxt = 0, n_nxt = 0;
    if (tile_nxt >= 0) {
      i_nxt = a.int_glob[(size_t)tile_nxt * a.stride + ts];
      own_nxt = a.int_slot[(size_t)tile_nxt * a.stride + ts];
      n_nxt = a.nell_s[(size_t)tile_nxt * a.stride + ts];
    }
    const double *sp = sp0 + (size_t)buf * 3 * a.cap;
    const int *st = st0 + (size_t)buf * a.cap;
    if (i_cur < a.n_local) {
      const double x_i = sp[3 * own_cur], y_i = sp[3 * own_cur + 1], z_i = sp[3 * own_cur + 2];
      const int type_i = ONETYPE ? 0 : st[own_cur];
      double fx = 0.0, fy = 0.0, fz = 0.0;
      // one 16-byte word = 8 columns of the warp's schedule (n_cur is the same multiple of 8 in every lane of the warp);
      // the next word is requested before the current one is used
      const uint4 *row = row_of_tile(tile);
      const int nchunk = n_cur >> 3;
      if (RING) {
        // word 0 arrived with the tile's coordinates (wait + barrier above); the row of the next tile, if this thread has one
        const uint4 *row_nxt = (tile_nxt >= 0 && i_nxt < a.n_local && n_nxt > 0) ? row_of_tile(tile_nxt) : nullptr;
        for (int c = 0; c < nchunk; c++) {
          if (c > 0) cp_async_wait_all(); // the word requested one iteration ago (and, long since, the next tile's coordinates)
          const uint4 cur = *ering;
          const uint4 *nextp = (c + 1 < nchunk) ? row + (size_t)(c + 1) * a.stride : row_nxt;
          if (nextp) cp_async16(ering, nextp); // same thread: the read above precedes the asynchronous write
          lj_quad<ONETYPE, ENERGY>(sp, st, cur.x, cur.y, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
          lj_quad<ONETYPE, ENERGY>(sp, st, cur.z, cur.w, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
        }
      } else {
        uint4 cur = make_uint4(0, 0, 0, 0);
        if (nchunk > 0) cur = ldg_nc_v4(row);
        if (nchunk > 1) prefetch_l1(row + a.stride);
        for (int c = 0; c < nchunk; c++) {
          if (c + 2 < nchunk) prefetch_l1(row + (size_t)(c + 2) * a.stride);
          uint4 nxt = make_uint4(0, 0, 0, 0);
          if (c + 1 < nchunk) nxt = ldg_nc_v4(row + (size_t)(c + 1) * a.stride);
          lj_quad<ONETYPE, ENERGY>(sp, st, cur.x, cur.y, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
          lj_quad<ONETYPE, ENERGY>(sp, st, cur.z, cur.w, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
          cur = nxt;
        }
      }
      if (!ENERGY) { f[3 * (size_t)i_cur] = fx; f[3 * (size_t)i_cur + 1] = fy; f[3 * (size_t)i_cur + 2] = fz; }
      if (FUSE) {
        const double dtfm = nve.dtf / nve.mass[ONETYPE ? a.type[i_cur] : type_i]; // integrator_nve.cpp:67,106
        double *vp = nve.v + 3 * (size_t)i_cur, *xp = nve.x_new + 3 * (size_t)i_cur;
        const double fi[3] = {fx, fy, fz}, xi[3] = {x_i, y_i, z_i};
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double kick = __dmul_rn(dtfm, fi[d]);
          const double v1 = __dadd_rn(vp[d], kick);            // final_integrate :107-109
          const double v2 = __dadd_rn(v1, kick);               // initial_integrate of the next step :68-70
          vp[d] = v2;
          xp[d] = __dadd_rn(xi[d], __dmul_rn(nve.dtv, v2));    // :71-73
        }
      }
    }
    if (RING && !(i_cur < a.n_local && n_cur > 0) && tile_nxt >= 0 && i_nxt < a.n_local && n_nxt > 0)
      cp_async16(ering, row_of_tile(tile_nxt)); // no row here, one in the next tile: its word 0
    i_cur = i_nxt; own_cur = own_nxt; n_cur = n_nxt;
    tile = tile_nxt; tile_nxt = tile_nn;
  }
  cp_async_wait_all();
  if (ENERGY) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pe += __shfl_down_sync(0xffffffffu, pe, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = pe;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += s_red[w];
      pe_partial[blockIdx.x] = s;
    }
  }
}

size_t emit_smem(int cap, int stride) { return (size_t)cap * (3 * sizeof(double) + sizeof(int)) + (size_t)stride * sizeof(unsigned short); }
size_t force_smem(int cap, bool types) { return 2 * ((size_t)cap * 3 * sizeof(double) + (types ? (size_t)cap * sizeof(int) : 0)); }

} // namespace

namespace emd { int device_sum_partials(emd_ctx *ctx, const double *d_partial, int n, double *h_out); }

struct emd_tiles {
  TileArgs a;
  int ntiles = 0;
  bool valid = false;
  unsigned short *d_ell = nullptr; size_t ell_cap = 0;
  int *d_nell = nullptr; size_t nell_cap = 0;
  unsigned short *d_ell_s = nullptr; size_t ell_s_cap = 0;
  int *d_nell_s = nullptr; size_t nell_s_cap = 0;
  int *d_stg_j = nullptr; size_t stg_j_cap = 0;
  int *d_stg_n = nullptr; size_t stg_n_cap = 0;
  unsigned short *d_int_slot = nullptr; size_t int_slot_cap = 0;
  int *d_int_glob = nullptr; size_t int_glob_cap = 0;
  int *d_order = nullptr; size_t order_cap = 0;
  int *d_has_ghost = nullptr; size_t has_ghost_cap = 0;
  int n_free_tiles = 0;  // tiles that do not read the halo (first in d_order)
  bool all_owned_have_rows = false; // every owned atom is listed (precondition of the fused force + integrator launch)
  int *d_flags = nullptr;
  int num_sms = 148;
  LJTab *d_tab = nullptr;
  double neigh_cut = 0.0;
  float cutf2 = 0.f;
  int max_smem_optin = 0;
  int max_smem_sm = 0;
};

namespace {

template <class K>
int set_smem(K kernel, size_t bytes) {
  EMD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

int ensure_bytes(void **p, size_t *cap, size_t bytes) {
  if (bytes <= *cap) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  const size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) { set_error("tiles: cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e)); return 1; }
  *cap = want;
  return 0;
}

} // namespace

extern "C" {

int emd_tiles_create(emd_tiles **out) {
  if (!out) { set_error("emd_tiles_create: out == NULL"); return 1; }
  emd_tiles *t = new emd_tiles();
  memset(&t->a, 0, sizeof t->a);
  EMD_CUDA(cudaMalloc((void **)&t->d_flags, 8 * sizeof(int)));
  EMD_CUDA(cudaMalloc((void **)&t->d_tab, sizeof(LJTab)));
  int dev = 0;
  EMD_CUDA(cudaGetDevice(&dev));
  EMD_CUDA(cudaDeviceGetAttribute(&t->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  EMD_CUDA(cudaDeviceGetAttribute(&t->num_sms, cudaDevAttrMultiProcessorCount, dev));
  EMD_CUDA(cudaDeviceGetAttribute(&t->max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
  *out = t;
  return 0;
}

void emd_tiles_destroy(emd_tiles *t) {
  if (!t) return;
  if (t->d_ell) cudaFree(t->d_ell);
  if (t->d_nell) cudaFree(t->d_nell);
  if (t->d_ell_s) cudaFree(t->d_ell_s);
  if (t->d_nell_s) cudaFree(t->d_nell_s);
  if (t->d_stg_j) cudaFree(t->d_stg_j);
  if (t->d_stg_n) cudaFree(t->d_stg_n);
  if (t->d_int_slot) cudaFree(t->d_int_slot);
  if (t->d_int_glob) cudaFree(t->d_int_glob);
  if (t->d_order) cudaFree(t->d_order);
  if (t->d_has_ghost) cudaFree(t->d_has_ghost);
  if (t->d_flags) cudaFree(t->d_flags);
  if (t->d_tab) cudaFree(t->d_tab);
  delete t;
}

int emd_tiles_valid(const emd_tiles *t) { return t && t->valid; }
void emd_tiles_invalidate(emd_tiles *t) { if (t) t->valid = false; }

int emd_tiles_info(const emd_tiles *t, int *tile_dims, int *ntiles, int *stride, int *maxrow, int *cap) {
  if (!t) return 1;
  if (tile_dims) { tile_dims[0] = t->a.tx; tile_dims[1] = t->a.ty; tile_dims[2] = t->a.tz; }
  if (ntiles) *ntiles = t->ntiles;
  if (stride) *stride = t->a.stride;
  if (maxrow) *maxrow = t->a.maxrow;
  if (cap) *cap = t->a.cap;
  return 0;
}

int emd_tiles_lists(const emd_tiles *t, const unsigned short **d_ell, const int **d_nell, int *maxrow_s, const unsigned short **d_ell_s,
                    const int **d_nell_s, const unsigned short **d_int_slot) {
  if (!t || !t->valid) { set_error("emd_tiles_lists: tiles not built"); return 1; }
  if (d_ell) *d_ell = t->a.ell;
  if (d_nell) *d_nell = t->a.nell;
  if (maxrow_s) *maxrow_s = t->a.maxrow_s;
  if (d_ell_s) *d_ell_s = t->a.ell_s;
  if (d_nell_s) *d_nell_s = t->a.nell_s;
  if (d_int_slot) *d_int_slot = t->a.int_slot;
  return 0;
}

// Build the tile-local full adjacency.  Returns 0 on success, 3 if the fast path does not apply
// to this configuration (the caller then uses emd_neigh_csr_* / emd_neigh_2d_fill), 1 on error.
int emd_neigh_tiles_build(emd_ctx *ctx, emd_tiles *t, const double *d_x, int n_local, int n_all, const emd_bin_geom *g,
                          const int *d_bincount, const int *d_binoffsets, const int *d_permute, double neigh_cut) {
  t->valid = false;
  const int nix = g->nbinx - 2 * g->nhalo, niy = g->nbiny - 2 * g->nhalo, niz = g->nbinz - 2 * g->nhalo;
  if (nix <= 0 || niy <= 0 || niz <= 0 || n_local <= 0) return 3;
  const double wx = (g->maxx - g->minx) / g->nbinx, wy = (g->maxy - g->miny) / g->nbiny, wz = (g->maxz - g->minz) / g->nbinz;
  // the 27-bin stencil must cover the list radius (true for every reference configuration:
  // bins are at least neigh_cut wide, binning_kksort.cpp:77-87); the reference itself simply
  // misses pairs otherwise, and the generic kernels reproduce that
  if (neigh_cut > wx || neigh_cut > wy || neigh_cut > wz) return 3;
  TileArgs &a = t->a;
  a.nbx = g->nbinx; a.nby = g->nbiny; a.nbz = g->nbinz; a.nhalo = g->nhalo;
  a.n_local = n_local;
  a.bincount = d_bincount; a.binoffsets = d_binoffsets; a.permute = d_permute;
  a.x = d_x; a.type = nullptr;
  a.ox = g->minx; a.oy = g->miny; a.oz = g->minz;
  a.wx = wx; a.wy = wy; a.wz = wz;
  a.stride = kForceThreads;
  // mean atoms per cell -> tile shape with ~0.9*stride atoms whose halo fits in shared memory
  const double m = std::max(1e-3, (double)n_all / ((double)g->nbinx * g->nbiny * g->nbinz));
  const int cap_max = 3000; // 72 KB of FP64 coordinates: three force CTAs per SM
  static const int shapes[][3] = {{4, 4, 8}, {4, 4, 4}, {2, 4, 4}, {2, 2, 8}, {2, 2, 4}, {2, 2, 2}, {1, 2, 2}, {1, 1, 2}, {1, 1, 1}};
  int pick = -1;
  for (int s = 0; s < (int)(sizeof shapes / sizeof shapes[0]); s++) {
    const int *d = shapes[s];
    const double atoms = m * d[0] * d[1] * d[2], staged = m * (d[0] + 2) * (d[1] + 2) * (d[2] + 2);
    if ((d[0] + 2) * (d[1] + 2) * (d[2] + 2) > kMaxStagedCells || d[0] * d[1] * d[2] > kMaxInteriorCells) continue;
    if (atoms <= 0.9 * a.stride && staged * 1.2 + 64 <= cap_max) { pick = s; break; }
  }
  if (pick < 0) return 3;
  a.tx = std::min(shapes[pick][0], nix); a.ty = std::min(shapes[pick][1], niy); a.tz = std::min(shapes[pick][2], niz);
  a.ntx = (nix + a.tx - 1) / a.tx; a.nty = (niy + a.ty - 1) / a.ty; a.ntz = (niz + a.tz - 1) / a.tz;
  const long long ntiles_ll = (long long)a.ntx * a.nty * a.ntz;
  if (ntiles_ll > 0x7fffffffLL) return 3;
  t->ntiles = (int)ntiles_ll;
  const int staged_cells = (a.tx + 2) * (a.ty + 2) * (a.tz + 2);
  a.cap = std::min(cap_max, ((int)(m * staged_cells * 1.25) + 64 + 31) / 32 * 32);
  // expected full-list row: density * sphere volume, +35 % head room
  const double rho = m / (wx * wy * wz);
  int maxrow = (int)(rho * 4.18879020478639 * neigh_cut * neigh_cut * neigh_cut * 1.35) + 8;
  maxrow = std::max(16, (maxrow + 7) / 8 * 8);
  // conservative FP32 radius: coordinates are relative to the staged region (extent E), so the
  // FP32 distance is off by < 8*E*2^-24; take 32*E*2^-23 + 2^-20 relative as the margin
  const double E = std::max({(a.tx + 2) * wx, (a.ty + 2) * wy, (a.tz + 2) * wz});
  const double thr = neigh_cut * (1.0 + 1.0 / 1048576.0) + 32.0 * E / 8388608.0;
  t->cutf2 = nextafterf((float)(thr * thr), INFINITY);
  t->neigh_cut = neigh_cut;

  for (int attempt = 0; attempt < 4; attempt++) {
    a.maxrow = maxrow;
    const size_t ell_bytes = (size_t)t->ntiles * a.maxrow * a.stride * sizeof(unsigned short);
    if (ensure_bytes((void **)&t->d_ell, &t->ell_cap, ell_bytes)) return 1;
    if (ensure_bytes((void **)&t->d_nell, &t->nell_cap, (size_t)t->ntiles * a.stride * sizeof(int))) return 1;
    a.ell = t->d_ell; a.nell = t->d_nell; a.flags = t->d_flags;
    EMD_CUDA(cudaMemsetAsync(t->d_flags, 0, 4 * sizeof(int), ctx->stream));
    const size_t smem = (size_t)a.cap * sizeof(float4);
    if (set_smem(tiles_filter_kernel, smem)) return 1;
    EMD_LAUNCH(ctx, tiles_filter_kernel, t->ntiles, kFilterThreads, smem, a, t->cutf2);
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, t->d_flags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    const int bits = ctx->h_pinned[0], need_row = ctx->h_pinned[1], need_cap = ctx->h_pinned[2], need_int = ctx->h_pinned[3];
    if (bits == 0) {
      // the slot numbering does not depend on cap: shrink it to the largest tile, so that the force
      // kernel's two coordinate buffers leave room for two CTAs per SM
      a.cap = std::max(32, (need_cap + 31) / 32 * 32);
      if (a.cap > kForceThreads * kStagePerThread) return 3;
      if (ensure_bytes((void **)&t->d_stg_j, &t->stg_j_cap, (size_t)t->ntiles * a.cap * sizeof(int))) return 1;
      if (ensure_bytes((void **)&t->d_stg_n, &t->stg_n_cap, (size_t)t->ntiles * sizeof(int))) return 1;
      if (ensure_bytes((void **)&t->d_int_slot, &t->int_slot_cap, (size_t)t->ntiles * a.stride * sizeof(unsigned short))) return 1;
      if (ensure_bytes((void **)&t->d_int_glob, &t->int_glob_cap, (size_t)t->ntiles * a.stride * sizeof(int))) return 1;
      if (ensure_bytes((void **)&t->d_order, &t->order_cap, (size_t)t->ntiles * sizeof(int))) return 1;
      if (ensure_bytes((void **)&t->d_has_ghost, &t->has_ghost_cap, (size_t)t->ntiles * sizeof(int))) return 1;
      a.stg_j = t->d_stg_j; a.stg_n = t->d_stg_n; a.int_slot = t->d_int_slot; a.int_glob = t->d_int_glob;
      a.order = t->d_order; a.has_ghost = t->d_has_ghost;
      EMD_CUDA(cudaMemsetAsync(t->d_flags + 4, 0, 4 * sizeof(int), ctx->stream));
      EMD_LAUNCH(ctx, tiles_tables_kernel, t->ntiles, kFilterThreads, 0, a);
      // the force kernel's copy of the adjacency: bank-conflict-free columns (tiles_schedule_kernel)
      if (ensure_bytes((void **)&t->d_nell_s, &t->nell_s_cap, (size_t)t->ntiles * a.stride * sizeof(int))) return 1;
      a.nell_s = t->d_nell_s;
      // EMD_TILES_SCHED = full | rot | none (measurement switch; default below)
      int mode = SCHED_ROT;
      if (const char *e = getenv("EMD_TILES_SCHED")) mode = !strcmp(e, "full") ? SCHED_FULL : !strcmp(e, "none") ? SCHED_NONE : SCHED_ROT;
      if (a.maxrow > kSchedMaxRow) mode = SCHED_NONE;
      int maxrow_s = a.maxrow + 8;
      const long long warps = (long long)t->ntiles * (a.stride / 32);
      const int sgrid = (int)((warps + kSchedWarps - 1) / kSchedWarps);
      for (int sa = 0; sa < 4; sa++) {
        a.maxrow_s = maxrow_s;
        if (ensure_bytes((void **)&t->d_ell_s, &t->ell_s_cap, (size_t)t->ntiles * a.maxrow_s * a.stride * sizeof(unsigned short))) return 1;
        a.ell_s = t->d_ell_s;
        const size_t wsm = sched_warp_smem(a.maxrow, a.maxrow_s), ssm = wsm * kSchedWarps;
        if (ssm > (size_t)t->max_smem_optin) return 3;
        EMD_CUDA(cudaMemsetAsync(t->d_flags, 0, 4 * sizeof(int), ctx->stream));
        EMD_LAUNCH(ctx, tiles_order_kernel, 1, 1024, 0, a, t->ntiles); // writes flags[2]
#define EMD_SCHED(M)                                                                                                        \
  do {                                                                                                                     \
    if (set_smem(tiles_schedule_kernel<M>, ssm)) return 1;                                                                 \
    EMD_LAUNCH(ctx, tiles_schedule_kernel<M>, sgrid, 32 * kSchedWarps, ssm, a, t->ntiles, (unsigned)wsm);                  \
  } while (0)
        if (mode == SCHED_ROT) {
          const size_t rwsm = rot_warp_smem(a.maxrow), rsm = rwsm * kRotWarps;
          if (set_smem(tiles_rotate_kernel, rsm)) return 1;
          EMD_LAUNCH(ctx, tiles_rotate_kernel, (int)((warps + kRotWarps - 1) / kRotWarps), 32 * kRotWarps, rsm, a, t->ntiles, (unsigned)rwsm);
        } else if (mode == SCHED_FULL) EMD_SCHED(SCHED_FULL);
        else EMD_SCHED(SCHED_NONE);
#undef EMD_SCHED
        EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, t->d_flags, 8 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        EMD_CUDA(cudaStreamSynchronize(ctx->stream));
        if (!(ctx->h_pinned[0] & 8)) {
          t->n_free_tiles = ctx->h_pinned[2];
          t->all_owned_have_rows = ctx->h_pinned[4] == n_local;
          t->valid = true;
          return 0;
        }
        maxrow_s = (ctx->h_pinned[1] + ctx->h_pinned[1] / 8 + 15) / 8 * 8;
      }
      return 3;
    }
    if (bits & 2) { (void)need_int; return 3; } // a tile holds more atoms than threads: density far from the estimate
    if (bits & 1) {
      if (need_cap > cap_max || need_cap > 65535) return 3;
      a.cap = (need_cap + need_cap / 16 + 31) / 32 * 32;
      if (a.cap > cap_max) a.cap = cap_max;
    }
    if (bits & 4) maxrow = (need_row + need_row / 8 + 7) / 8 * 8;
  }
  return 3;
}

int emd_neigh_tiles_count(emd_ctx *ctx, emd_tiles *t, int half, int newton, int *d_row_map, int *h_total) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_count: tiles not built"); return 1; }
  TileArgs &a = t->a;
  EmitArgs e;
  memset(&e, 0, sizeof e);
  e.cutsq = t->neigh_cut * t->neigh_cut; e.newton = newton; e.counts = d_row_map;
  EMD_CUDA(cudaMemsetAsync(d_row_map, 0, sizeof(int) * ((size_t)a.n_local + 1), ctx->stream));
  const size_t smem = emit_smem(a.cap, a.stride);
  if (half) { if (set_smem(tiles_emit_kernel<true, EMIT_COUNT>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<true, EMIT_COUNT>), t->ntiles, a.stride, smem, a, e); }
  else { if (set_smem(tiles_emit_kernel<false, EMIT_COUNT>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<false, EMIT_COUNT>), t->ntiles, a.stride, smem, a, e); }
  if (exclusive_scan_int(ctx, d_row_map, d_row_map, a.n_local + 1, nullptr)) return 1;
  if (h_total) {
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, d_row_map + a.n_local, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_total = ctx->h_pinned[0];
    if (*h_total < 0) { set_error("emd_neigh_tiles_count: neighbor count overflows 32-bit row_map"); return 2; }
  }
  return 0;
}

int emd_neigh_tiles_fill_csr(emd_ctx *ctx, emd_tiles *t, int half, int newton, const int *d_row_map, int *d_entries) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_fill_csr: tiles not built"); return 1; }
  TileArgs &a = t->a;
  EmitArgs e;
  memset(&e, 0, sizeof e);
  e.cutsq = t->neigh_cut * t->neigh_cut; e.newton = newton; e.row_map = d_row_map; e.entries = d_entries;
  const size_t smem = emit_smem(a.cap, a.stride);
  if (half) { if (set_smem(tiles_emit_kernel<true, EMIT_CSR>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<true, EMIT_CSR>), t->ntiles, a.stride, smem, a, e); }
  else { if (set_smem(tiles_emit_kernel<false, EMIT_CSR>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<false, EMIT_CSR>), t->ntiles, a.stride, smem, a, e); }
  return 0;
}

int emd_neigh_tiles_fill_2d(emd_ctx *ctx, emd_tiles *t, int half, int newton, int maxneighs, int *d_num_neighs, int *d_neighs,
                            int *h_max_count) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_fill_2d: tiles not built"); return 1; }
  TileArgs &a = t->a;
  EmitArgs e;
  memset(&e, 0, sizeof e);
  e.cutsq = t->neigh_cut * t->neigh_cut; e.newton = newton; e.counts = d_num_neighs; e.entries = d_neighs; e.maxneighs = maxneighs;
  e.max_count = t->d_flags + 1;
  EMD_CUDA(cudaMemsetAsync(e.max_count, 0, sizeof(int), ctx->stream));
  EMD_CUDA(cudaMemsetAsync(d_num_neighs, 0, sizeof(int) * ((size_t)a.n_local + 1), ctx->stream));
  const size_t smem = emit_smem(a.cap, a.stride);
  if (half) { if (set_smem(tiles_emit_kernel<true, EMIT_2D>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<true, EMIT_2D>), t->ntiles, a.stride, smem, a, e); }
  else { if (set_smem(tiles_emit_kernel<false, EMIT_2D>, smem)) return 1; EMD_LAUNCH(ctx, (tiles_emit_kernel<false, EMIT_2D>), t->ntiles, a.stride, smem, a, e); }
  if (h_max_count) {
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, e.max_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_max_count = ctx->h_pinned[0];
  }
  return 0;
}

// ForceLJNeigh::compute / compute_energy on the tile lists.  d_x/d_type are the CURRENT arrays
// (atoms keep their indices between rebuilds; the binning arrays captured at build time must
// still be alive).  With h_pe != NULL only the energy is computed (forces untouched).
// part: 0 = every tile; 1 = only the tiles that do not read the halo (their forces are final before the halo exchange of
// this step has landed); 2 = the rest.  reserve_ctas > 0 leaves that many CTA slots of the persistent grid free, so that
// the pack and NCCL kernels of a concurrent halo exchange find room on the SMs.
static int lj_tiles_launch(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *h_pe, int part,
                           int reserve_ctas, const NveFuse *fuse = nullptr) {
  if (!t || !t->valid) { set_error("emd_force_lj_compute_tiles: tiles not built"); return 1; }
  if (ctx->lj.ntypes == 0) { set_error("emd_force_lj_compute_tiles: parameters not set"); return 1; }
  if (part < 0 || part > 2 || (h_pe && part != 0)) { set_error("emd_force_lj_compute_tiles: bad part"); return 1; }
  if (fuse && h_pe) { set_error("emd_force_lj_compute_tiles: the energy launch cannot carry the integrator"); return 1; }
  const NveFuse nve = fuse ? *fuse : NveFuse{nullptr, nullptr, nullptr, 0.0, 0.0};
  TileArgs a = t->a;
  a.x = d_x; a.type = d_type;
  const bool one = ctx->lj.ntypes == 1;
  LJOne p1 = {ctx->lj.lj1[0], ctx->lj.lj2[0], ctx->lj.cutsq[0]};
  if (!one) {
    LJTab h;
    h.ntypes = ctx->lj.ntypes;
    memcpy(h.lj1, ctx->lj.lj1, sizeof h.lj1); memcpy(h.lj2, ctx->lj.lj2, sizeof h.lj2); memcpy(h.cutsq, ctx->lj.cutsq, sizeof h.cutsq);
    EMD_CUDA(cudaMemcpyAsync(t->d_tab, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream)); // h is a stack object
  }
  const size_t base_smem = force_smem(a.cap, !one);
  if (base_smem > (size_t)t->max_smem_optin) { set_error("emd_force_lj_compute_tiles: tile does not fit in shared memory"); return 1; }
  // the ELL ring (16 bytes per thread) only if two CTAs still fit on one SM (1 KB per CTA is reserved by the system).
  // Measured at 2 M atoms (gpurun r01t): the ring halves the long-scoreboard stalls (18 % -> 10 % of samples) but the
  // per-iteration cp.async wait and the extra LDS.128 cost more than that buys: 0.386 ms against 0.367 ms with the L1
  // prefetch.  Off unless EMD_TILES_RING=1.
  const bool ring_allowed = getenv("EMD_TILES_RING") && atoi(getenv("EMD_TILES_RING"));
  const size_t ring_smem = base_smem + (size_t)kForceThreads * sizeof(uint4);
  const bool ring = !fuse && ring_allowed && 2 * (ring_smem + 1024) <= (size_t)t->max_smem_sm && 2 * (base_smem + 1024) <= (size_t)t->max_smem_sm;
  const size_t smem = ring ? ring_smem : base_smem;
  const int first = part == 2 ? t->n_free_tiles : 0;
  const int count = part == 0 ? t->ntiles : part == 1 ? t->n_free_tiles : t->ntiles - t->n_free_tiles;
  if (count <= 0) return 0;
  const int grid = std::max(1, std::min(count, 2 * emd_ctx_side_sms(ctx) - std::max(0, reserve_ctas)));
  double *partial = nullptr;
  if (h_pe) {
    if (ctx->s_c.ensure(sizeof(double) * ((size_t)grid + 8))) return 1;
    partial = ctx->s_c.as<double>() + 8;
  }
#define EMD_LJ_TILES(ONE, EN, RG, FU)                                                                                      \
  do {                                                                                                                     \
    if (set_smem(lj_tiles_kernel<ONE, EN, RG, FU>, smem)) return 1;                                                        \
    EMD_LAUNCH(ctx, (lj_tiles_kernel<ONE, EN, RG, FU>), grid, kForceThreads, smem, a, first, count, p1, t->d_tab, d_f,     \
               partial, (unsigned)base_smem, nve);                                                                         \
  } while (0)
#define EMD_LJ_TILES2(ONE, EN) do { if (ring) EMD_LJ_TILES(ONE, EN, true, false); else EMD_LJ_TILES(ONE, EN, false, false); } while (0)
  if (fuse) { if (one) EMD_LJ_TILES(true, false, false, true); else EMD_LJ_TILES(false, false, false, true); }
  else if (h_pe) { if (one) EMD_LJ_TILES2(true, true); else EMD_LJ_TILES2(false, true); }
  else { if (one) EMD_LJ_TILES2(true, false); else EMD_LJ_TILES2(false, false); }
#undef EMD_LJ_TILES2
#undef EMD_LJ_TILES
  if (h_pe) return device_sum_partials(ctx, partial, grid, h_pe);
  return 0;
}

int emd_force_lj_compute_tiles(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *h_pe) {
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, h_pe, 0, 0);
}

int emd_force_lj_compute_tiles_part(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, int part,
                                    int reserve_ctas) {
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, part, reserve_ctas);
}

int emd_force_lj_compute_tiles_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *d_v,
                                   double *d_x_new, const double *d_mass, double dtf, double dtv) {
  if (t && t->valid && !t->all_owned_have_rows) return 3; // an owned atom has no row: its position would not be advanced
  if (!d_v || !d_x_new || !d_mass || d_x_new == d_x) { set_error("emd_force_lj_compute_tiles_nve: v, mass and a second position array are required"); return 1; }
  const NveFuse nve = {d_v, d_x_new, d_mass, dtf, dtv};
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, 0, 0, &nve);
}

int emd_force_lj_compute_tiles_part_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, int part,
                                        int reserve_ctas, double *d_v, double *d_x_new, const double *d_mass, double dtf, double dtv) {
  if (t && t->valid && !t->all_owned_have_rows) return 3;
  if (!d_v || !d_x_new || !d_mass || d_x_new == d_x) { set_error("emd_force_lj_compute_tiles_part_nve: v, mass and a second position array are required"); return 1; }
  const NveFuse nve = {d_v, d_x_new, d_mass, dtf, dtv};
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, part, reserve_ctas, &nve);
}

int emd_tiles_complete(const emd_tiles *t, int *all_owned_have_rows) {
  if (!t || !t->valid) { set_error("emd_tiles_complete: tiles not built"); return 1; }
  if (all_owned_have_rows) *all_owned_have_rows = t->all_owned_have_rows ? 1 : 0;
  return 0;
}

int emd_tiles_halo_split(const emd_tiles *t, int *n_free, int *n_halo) {
  if (!t || !t->valid) { set_error("emd_tiles_halo_split: tiles not built"); return 1; }
  if (n_free) *n_free = t->n_free_tiles;
  if (n_halo) *n_halo = t->ntiles - t->n_free_tiles;
  return 0;
}

} // extern "C"
