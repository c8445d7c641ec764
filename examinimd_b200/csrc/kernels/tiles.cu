// tiles.cu -- the B200 fast path of the neighbor build and the LJ force: cell TILES staged in
// shared memory.
//
// Why (measured on B200, profiles/r01_microbench_b200.json): a thread-per-atom walk over a CSR
// list is bound by the L1/L2 gather of x[j] (184 G rows/s) and, for Newton-3 half lists, by
// REDG.F64 (224 Gop/s; shared-memory FP64 atomics are CAS loops).  FP64 FMA issue (36.8 TFLOP/s)
// is the cheap resource on this part, so the fast path recomputes each pair from both sides
// instead of scattering -f to j, and serves every x[j] from shared memory:
//
//   * the interior bins of BinningKKSort's grid are grouped into tiles of tx*ty*tz cells; one
//     CTA owns one tile and stages the coordinates of the (tx+2)(ty+2)(tz+2) cells around it
//     ONCE (coalesced: atoms are cell-sorted);
//   * the neighbor build is three kernels per re-neighboring:
//       tiles_search_kernel  warp = one interior cell, lane = atom; every candidate of the 27-bin
//                            stencil costs 6 instructions (broadcast LDS.128 of a pre-scaled FP32
//                            record, 3 FFMA, FADD, funnel shift of the sign bit) and leaves ONE BIT
//                            in a mask word -- a conservative FP32 superset of the list;
//       tiles_lists_kernel   thread = atom row: pops the mask bits, applies the reference's exact
//                            FP64 inclusion test (rsq <= cut*cut, no FMA contraction; half-list
//                            owner rule) and leaves the exact rows as 16-bit staged-slot numbers
//                            in the reference's order + their lengths (= CSR row counts); in the
//                            same pass the FP32 superset is re-ordered in shared memory into the
//                            force kernel's bank-conflict-free column layout and written ONCE;
//       tiles_fill_kernel    CSR entries / 2D table = slot numbers -> atom indices at the offsets
//                            of the scanned counts: no coordinates, no arithmetic.
//     The API-visible list is bit-identical to the reference's; the force kernel's list is a
//     superset that differs only by pairs within ~1e-4 of the list radius (which the force
//     kernel's own cutoff test rejects);
//   * the force kernel walks its rows with x[j] from shared memory, f_i in registers, one
//     coalesced store per atom, no atomics, no zero-f pass.  Warps are decoupled: coordinate
//     buffers are handed from the producer warp to the consumers and back through mbarriers
//     (cp.async completion -> `full`, one arrival per warp -> `empty`); there is no CTA-wide barrier in the
//     tile loop, and a warp may run up to two tiles ahead of the slowest one.
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
constexpr int kMaxStagedRows = 36; // (tx+2)(ty+2) of the largest tile shape
constexpr int kSearchThreads = 256;
constexpr int kRowThreads = 384;  // dense atom slots (rows) per tile = threads of the lists / fill / force kernels
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kDummySlots = 16;   // the force kernel's coordinate buffers start with 16 far-away atoms (one per bank), the targets of padding entries
constexpr int kCapMax = 2688;     // (kDummySlots + cap) * 24 must fit in 16 bits
constexpr int kMaxRowLimit = 248; // bucket heads and counts of the bank sort are bytes
constexpr double kFarAway = 1e100;

enum { OVF_CAP = 1, OVF_INT = 2, OVF_ROW = 4, OVF_WORDS = 8 };
enum { FL_BITS = 0, FL_NEED_ROW = 1, FL_NEED_CAP = 2, FL_NEED_INT = 3, FL_OWNED_ROWS = 4, FL_NFREE = 5, FL_NEED_WORDS = 6, FL_MAX2D = 7, FL_COUNT = 16 };

struct TileArgs {
  int nbx, nby, nbz, nhalo; // bin grid incl. halo bins
  int tx, ty, tz;           // interior cells per tile
  int ntx, nty, ntz;        // tiles per dimension
  int stride;               // row slots per tile (= kRowThreads)
  int maxrow;               // row capacity (entries), multiple of 8, <= kMaxRowLimit
  int cap;                  // staged-atom capacity of the shared-memory arrays, multiple of 32
  int fcap;                 // the force kernel's coordinate buffers: the largest tile of this build, multiple of 32 (<= cap)
  int n_local;
  const int *bincount, *binoffsets, *permute;
  const double *x;
  const int *type;
  double ox, oy, oz; // bin grid origin
  double wx, wy, wz; // bin widths
  // search result: the FP32 superset rows in list order (the stencil is 9 runs of 3 z-adjacent bins = 9 contiguous
  // staged-slot ranges, walked in the reference's order), staged-slot numbers, 8 per 16-byte word
  unsigned short *ell; // [ntiles][maxrow/8][stride][8]
  int *nell;           // [ntiles][stride]
  // exact rows (reference order, reference inclusion rules) as staged-slot numbers, 8 per 16-byte word
  unsigned short *csr16; // [ntiles][maxrow/8][stride][8]
  int *ncsr;             // [ntiles][stride] exact row lengths
  // the force kernel's rows: the FP32 superset in COLUMNS that a half-warp reads from shared memory without bank
  // conflicts; an entry is the BYTE offset 24*slot of the neighbor's coordinates; padding entries point at a dummy slot
  unsigned short *ell_s; // [ntiles][maxrow/8][stride][8]
  int *nell_s;           // [ntiles][stride]: columns of the thread's warp, multiple of 8
  // per-tile staging tables, so that the per-step force kernel needs no cell arithmetic
  int *stg_j;                // [ntiles][cap]  global index of every staged slot
  int *stg_n;                // [ntiles]  staged atoms
  // staged z-rows ((tz+2) z-adjacent bins = one contiguous slot range): {first slot, first atom index or -1, atoms, 0}.  Owned atoms
  // are cell-sorted, so a row without ghosts is one contiguous piece of x and is copied without the slot -> atom table
  int4 *rowdesc;             // [ntiles][kMaxStagedRows]
  unsigned short *int_slot;  // [ntiles][stride]  staged slot of the row's own atom
  int *int_glob;             // [ntiles][stride]  its global index (>= n_local or 0x7fffffff: no row)
  // tiles whose staged cells hold only owned atoms come first in `order` (their forces do not depend on the halo, so a
  // decomposed run computes them while the halo exchange is in flight); has_ghost[tile] is the classification
  int *order;          // [ntiles]
  int *has_ghost;      // [ntiles]
  int *flags;          // FL_*
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

// cell tables of this CTA's tile: s_start[c] = first staged slot of staged cell c (cells z-fastest, so the three
// z-adjacent cells of a stencil run are one contiguous slot range), s_goff[c] = its offset in the permute vector,
// s_ibase[ci] = first dense atom slot of interior cell ci
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

// stencil run r (0..8) of the interior cell whose staged index is c_i: staged slots [lo, hi)
__device__ __forceinline__ void stencil_run(const TileCtx &t, const int *s_start, int c_i, int r, int &lo, int &hi) {
  const int c0 = c_i + ((r / 3 - 1) * t.syn + (r % 3 - 1)) * t.szn - 1;
  lo = s_start[c0];
  hi = s_start[c0 + 3];
}

// --------------------------------------------------------------------------------- search
// FP32 record of a staged atom, relative to the centre of the staged region: (-2x, -2y, -2z, x^2+y^2+z^2).
// r^2(i,j) = |p_i|^2 + (w_j + x_i X_j + y_i Y_j + z_i Z_j): three FFMA per candidate; the candidate is kept when
// r^2 - thr2 < 0, whose sign bit is shifted into the mask word.  thr2 carries the rounding margin of this form
// (|p| <= half the staged extent), so the mask is a superset of the exact list.
__global__ void __launch_bounds__(kSearchThreads) tiles_search_kernel(TileArgs a, float thr2) {
  __shared__ int s_start[kMaxStagedCells + 4], s_goff[kMaxStagedCells], s_ibase[kMaxInteriorCells + 1];
  extern __shared__ __align__(16) unsigned char dyn[];
  float4 *sf = reinterpret_cast<float4 *>(dyn);
  TileCtx t;
  tile_setup(a, t, s_start, s_goff, s_ibase);
  if (threadIdx.x == 0) {
    atomicMax(&a.flags[FL_NEED_CAP], t.total);
    atomicMax(&a.flags[FL_NEED_INT], t.n_int);
    if (t.total > a.cap) atomicOr(&a.flags[FL_BITS], OVF_CAP);
    if (t.n_int > a.stride) atomicOr(&a.flags[FL_BITS], OVF_INT);
    a.stg_n[t.tile] = min(t.total, a.cap);
  }
  if (t.total > a.cap || t.n_int > a.stride) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const double cx0 = a.ox + (t.bx0 - 1 + 0.5 * t.sxn) * a.wx, cy0 = a.oy + (t.by0 - 1 + 0.5 * t.syn) * a.wy,
               cz0 = a.oz + (t.bz0 - 1 + 0.5 * t.szn) * a.wz;
  // stage the records and write the staging table of the force kernel
  int ghost = 0;
  for (int c = warp; c < t.ncs; c += nwarps) {
    const int base = s_start[c], n = s_start[c + 1] - base, goff = s_goff[c];
    for (int k = lane; k < n; k += 32) {
      const int j = a.permute[goff + k];
      const float xr = (float)(a.x[3 * (size_t)j] - cx0), yr = (float)(a.x[3 * (size_t)j + 1] - cy0), zr = (float)(a.x[3 * (size_t)j + 2] - cz0);
      sf[base + k] = make_float4(-2.f * xr, -2.f * yr, -2.f * zr, fmaf(zr, zr, fmaf(yr, yr, xr * xr)));
      a.stg_j[(size_t)t.tile * a.cap + base + k] = j;
      ghost |= (j >= a.n_local);
    }
  }
  ghost = __syncthreads_or(ghost);
  if (threadIdx.x == 0) a.has_ghost[t.tile] = ghost ? 1 : 0;
  for (int r = warp; r < t.sxn * t.syn; r += nwarps) { // row descriptors: is the row one run of consecutive atom indices?
    const int c0 = r * t.szn;
    const int sb = s_start[c0], n = s_start[c0 + t.szn] - sb;
    const int j0 = n > 0 ? a.stg_j[(size_t)t.tile * a.cap + sb] : -1; // (written above by this CTA; visible after the barrier)
    bool ok = j0 >= 0 && j0 + n <= a.n_local;
    for (int k = lane; k < n && ok; k += 32) ok = a.stg_j[(size_t)t.tile * a.cap + sb + k] == j0 + k;
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) a.rowdesc[(size_t)t.tile * kMaxStagedRows + r] = make_int4(sb, ok ? j0 : -1, n, 0);
  }
  for (int k = t.n_int + threadIdx.x; k < a.stride; k += blockDim.x) {
    a.nell[(size_t)t.tile * a.stride + k] = 0;
    a.int_slot[(size_t)t.tile * a.stride + k] = 0;
    a.int_glob[(size_t)t.tile * a.stride + k] = 0x7fffffff;
  }
  int owned_rows = 0;
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
      const int tslot = s_ibase[ci] + k;
      if (in_cell) {
        a.int_slot[(size_t)t.tile * a.stride + tslot] = (unsigned short)own;
        a.int_glob[(size_t)t.tile * a.stride + tslot] = i_glob;
        owned_rows += active;
      }
      const float4 me = in_cell ? sf[own] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float xi = -0.5f * me.x, yi = -0.5f * me.y, zi = -0.5f * me.z, ci_thr = thr2 - me.w;
      // Accepted slots are shifted into a 64-bit accumulator from the top (after four of them the first sits in the low 16
      // bits) and leave as whole 16-byte words: one store per 8 entries.
      int q = 0;
      unsigned long long acc = 0ull, lo64 = 0ull;
      uint4 *const wrow = reinterpret_cast<uint4 *>(a.ell) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + tslot;
      for (int r = 0; r < 9; r++) {
        int lo, hi;
        stencil_run(t, s_start, c_i, r, lo, hi);
        for (int s = lo; s < hi; s += 32) {
          const int n = min(32, hi - s);
          unsigned m = 0u;
          const float4 *cand = sf + s; // same address in every lane: broadcast
          if (n == 32) {
#pragma unroll
            for (int u = 0; u < 32; u++) {
              const float4 pj = cand[u];
              const float d = fmaf(zi, pj.z, fmaf(yi, pj.y, fmaf(xi, pj.x, pj.w))) - ci_thr;
              m = __funnelshift_l(__float_as_uint(d), m, 1);
            }
          } else {
#pragma unroll 4
            for (int u = 0; u < n; u++) {
              const float4 pj = cand[u];
              const float d = fmaf(zi, pj.z, fmaf(yi, pj.y, fmaf(xi, pj.x, pj.w))) - ci_thr;
              m = __funnelshift_l(__float_as_uint(d), m, 1);
            }
          }
          unsigned word = __brev(m) >> (32 - n); // bit u = candidate s + u
          if ((unsigned)(own - s) < (unsigned)n) word &= ~(1u << (own - s));
          if (!active) word = 0u;
          while (word) { // list order = ascending slot
            const unsigned slot = (unsigned)(s + __ffs(word) - 1);
            word &= word - 1;
            acc = (acc >> 16) | ((unsigned long long)slot << 48);
            q++;
            if ((q & 3) == 0) {
              if (q & 4) lo64 = acc;
              else if (q <= a.maxrow) wrow[(size_t)((q >> 3) - 1) * a.stride] = make_uint4((unsigned)lo64, (unsigned)(lo64 >> 32), (unsigned)acc, (unsigned)(acc >> 32));
            }
          }
        }
      }
      if (in_cell) {
        const int rem = q & 7;
        if (rem && q < a.maxrow) { // the last, partial word (entries beyond the row length are never read)
          unsigned long long hi64 = 0ull;
          if (rem < 4) lo64 = acc >> (16 * (4 - rem));
          else if (rem > 4) hi64 = acc >> (16 * (8 - rem));
          wrow[(size_t)(q >> 3) * a.stride] = make_uint4((unsigned)lo64, (unsigned)(lo64 >> 32), (unsigned)hi64, (unsigned)(hi64 >> 32));
        }
        a.nell[(size_t)t.tile * a.stride + tslot] = min(q, a.maxrow);
        if (q > a.maxrow) { atomicOr(&a.flags[FL_BITS], OVF_ROW); atomicMax(&a.flags[FL_NEED_ROW], q); }
      }
    }
  }
  // flags[FL_OWNED_ROWS]: owned atoms that have a row (= n_local unless an owned atom sits outside the interior bins, which
  // the reference leaves without neighbors, neighbor_csr.h:184; the fused force + integrator launch requires equality)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) owned_rows += __shfl_down_sync(0xffffffffu, owned_rows, o);
  if (lane == 0 && owned_rows) atomicAdd(&a.flags[FL_OWNED_ROWS], owned_rows);
}

// order[]: halo-independent tiles first, then the others, each group in tile order.  One block: every thread owns a
// contiguous chunk of tiles; flags[FL_NFREE] receives the number of halo-independent tiles.
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
  if (threadIdx.x == 0) a.flags[FL_NFREE] = total_free;
}

// ------------------------------------------------------------------ exact rows + force rows
// Bank-conflict-free columns.  The force kernel reads x[j] of 32 different neighbors per warp instruction from shared
// memory (LDS.64, served per half-warp: one wavefront if the 16 words lie in 16 different bank pairs or are the same
// word).  In list order the 16 rows of a half-warp hit ~2.5 words per bank pair, and the shared-memory pipe, not FP64,
// bounds the kernel (profiles/r01e).  Rows are therefore re-ordered once per build: counting sort of the row by bank
// (slot mod 16: the AoS stride 3 is odd, so x, y and z all map slot -> bank pair bijectively), then lane l takes, at
// column q, an entry of bank (q + l) mod 16 while it has one, else of a bank that still holds more than its share.
// 1.1 instead of 3.6 extra wavefronts per half-warp column (test_tiles_force_rows_are_a_conflict_light_permutation).
struct ListArgs {
  double cutsq;
  int half, newton;
  int *counts; // counts[i] = exact row length (the CSR row_map before its scan), or nullptr
};

// thread = row.  Pass A walks the row's words (8 entries each, independent of one another): bank histogram, exact FP64 test
// of the reference (+ half-list owner rule) -> exact row (csr16) and its length.  Pass B scatters the entries into 16 bank
// buckets in shared memory (the coordinates are dead by then: same memory).  Then the columns of the force row: round t of
// a lane (hl = lane & 15) = the 16 columns 16t .. 16t+15, in which it visits the banks hl, hl+1, ... once each: column q
// takes the (q >> 4)-th entry of bank (q + hl) & 15 (slot mod 16: the AoS stride 3 is odd, so x, y and z all map slot ->
// bank pair bijectively) if the bucket is that deep.  Entries whose column does not exist (beyond the row length) are the
// "excess"; they fill, in order, the columns whose bucket was too shallow.  As many excess entries as holes, so every
// lane is done after n columns.  In list order the 16 rows of a half-warp hit ~2.5 words per bank pair and the
// shared-memory pipe, not FP64, bounds the force kernel (profiles/r01e); with these columns every bank pair is hit by the
// two lanes l, l+16 except where a hole was filled (test_tiles_force_rows_are_a_conflict_light_permutation).
constexpr int kExcessRoom = 40; // excess entries of a row (typically < 25)

template <bool EXACT>
__global__ void __launch_bounds__(kRowThreads, 2) tiles_lists_kernel(TileArgs a, ListArgs e, unsigned coords_bytes, int rowcap) {
  __shared__ int s_start[kMaxStagedCells + 4], s_goff[kMaxStagedCells], s_ibase[kMaxInteriorCells + 1];
  extern __shared__ __align__(16) unsigned char dyn[];
  // region A: FP64 coordinates + global indices while the exact rows are made, then the bank-sorted rows
  double *sx = reinterpret_cast<double *>(dyn);                 // [cap][3]
  int *sj = reinterpret_cast<int *>(dyn + (size_t)a.cap * 24);  // [cap]
  unsigned char *srow = reinterpret_cast<unsigned char *>(dyn) + threadIdx.x;                    // srow[pos * kRowThreads]: slot >> 4 (the bucket is the bank)
  unsigned short *sexc = reinterpret_cast<unsigned short *>(dyn + (size_t)a.maxrow * kRowThreads) + threadIdx.x; // sexc[k * kRowThreads]: excess entries (slots)
  unsigned short *state = reinterpret_cast<unsigned short *>(dyn + coords_bytes) + threadIdx.x;  // state[bank * kRowThreads]
  TileCtx t;
  tile_setup(a, t, s_start, s_goff, s_ibase);
  if (t.total > a.cap || t.n_int > a.stride) return; // flagged by the search kernel
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool ghost_rule = EXACT && e.half && !e.newton && a.has_ghost[t.tile]; // only then does j < n_local matter below
  if (EXACT) {
    for (int c = warp; c < t.ncs; c += kRowWarps) {
      const int base = s_start[c], n = s_start[c + 1] - base, goff = s_goff[c];
      for (int k = lane; k < n; k += 32) {
        const int j = a.permute[goff + k];
        const int s = base + k;
        sx[3 * s] = a.x[3 * (size_t)j]; sx[3 * s + 1] = a.x[3 * (size_t)j + 1]; sx[3 * s + 2] = a.x[3 * (size_t)j + 2];
        if (ghost_rule) sj[s] = j;
      }
    }
    __syncthreads();
  }
  const int ts = threadIdx.x;
  const size_t rbase = (size_t)t.tile * a.stride + ts;
  const int i = ts < t.n_int ? a.int_glob[rbase] : 0x7fffffff;
  const bool active = i < a.n_local;
  const int own = active ? a.int_slot[rbase] : 0;
  const int n = active ? a.nell[rbase] : 0;
  const uint4 *const row = reinterpret_cast<const uint4 *>(a.ell) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + ts;
  const int nchunk = (n + 7) >> 3;
  // ---- pass A
  int count = 0;
  unsigned long long cnt0 = 0ull, cnt1 = 0ull; // 16 bank counters, one byte each (a row holds <= 248 entries)
  if (n > 0) {
    double x_i = 0.0, y_i = 0.0, z_i = 0.0;
    if (EXACT) { x_i = sx[3 * own]; y_i = sx[3 * own + 1]; z_i = sx[3 * own + 2]; }
    const long long cutsq_bits = __double_as_longlong(e.cutsq);
    unsigned long long acc = 0ull, lo = 0ull;
    uint4 *const wrow = reinterpret_cast<uint4 *>(a.csr16) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + ts;
    uint4 cur = row[0];
    for (int c = 0; c < nchunk; c++) {
      const uint4 nxt = (c + 1 < nchunk) ? row[(size_t)(c + 1) * a.stride] : make_uint4(0, 0, 0, 0);
      const unsigned w4[4] = {cur.x, cur.y, cur.z, cur.w};
      bool keep[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int slot = (int)((w4[k >> 1] >> (16 * (k & 1))) & 0xffffu);
        const bool valid = c * 8 + k < n;
        if (valid) { // bank histogram
          const unsigned long long inc = 1ull << ((slot & 7) * 8);
          if (slot & 8) cnt1 += inc; else cnt0 += inc;
        }
        keep[k] = false;
        if (EXACT) {
          const double dx = x_i - sx[3 * slot], dy = y_i - sx[3 * slot + 1], dz = z_i - sx[3 * slot + 2];
          bool kp = valid;
          if (e.half) {
            // neighbor_csr.h:290-291 (j != i by construction): j stays if it is a ghost without newton, or "greater" than i:
            // x_j > x_i || (x_j == x_i && (y_j > y_i || (y_j == y_i && z_j > z_i))).  x_j > x_i <=> dx < 0 and x_j == x_i <=> dx == +0
            // exactly (IEEE subtraction), so the rule is read off the sign / zero bits of the differences on the integer pipe.
            const bool owned_j = ghost_rule ? sj[slot] < a.n_local : true; // newton on, or no ghost staged: the rule applies to every j
            const long long bx = __double_as_longlong(dx), by = __double_as_longlong(dy), bz = __double_as_longlong(dz);
            const bool greater = bx < 0 || (bx == 0 && (by < 0 || (by == 0 && bz < 0)));
            kp = valid && (!owned_j || greater);
          }
          const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
          keep[k] = kp && __double_as_longlong(rsq) <= cutsq_bits; // rsq <= cutsq, neighbor_csr.h:206,299 (both non-negative: bit order)
        }
      }
      if (EXACT) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (keep[k]) {
            const unsigned long long slot = (w4[k >> 1] >> (16 * (k & 1))) & 0xffffu;
            acc = (acc >> 16) | (slot << 48);
            count++;
            if ((count & 3) == 0) {
              if (count & 4) lo = acc;
              else wrow[(size_t)((count >> 3) - 1) * a.stride] = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)acc, (unsigned)(acc >> 32));
            }
          }
        }
      }
      cur = nxt;
    }
    if (EXACT) {
      const int rem = count & 7;
      if (rem) { // the last, partial word (count <= n <= maxrow)
        unsigned long long hi = 0ull;
        if (rem < 4) lo = acc >> (16 * (4 - rem));
        else if (rem > 4) hi = acc >> (16 * (8 - rem));
        wrow[(size_t)(count >> 3) * a.stride] = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
      }
      if (e.counts) e.counts[i] = count;
    }
  }
  if (EXACT) a.ncsr[rbase] = count;
  const int nmax = __reduce_max_sync(0xffffffffu, n);
  __syncthreads(); // every thread is done with the coordinates: region A becomes the bank-sorted rows
  // ---- bucket heads from the histogram: state[b] = head position | entries << 8
  {
    unsigned run = 0u;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const unsigned cb = (unsigned)((b < 8 ? cnt0 : cnt1) >> ((b & 7) * 8)) & 0xffu;
      state[b * kRowThreads] = (unsigned short)(run | (cb << 8));
      run += cb;
    }
  }
  // ---- pass B: scatter into the buckets (stable: list order inside a bank)
  if (n > 0) {
    uint4 cur = row[0];
    for (int c = 0; c < nchunk; c++) {
      const uint4 nxt = (c + 1 < nchunk) ? row[(size_t)(c + 1) * a.stride] : make_uint4(0, 0, 0, 0);
      const unsigned w4[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (c * 8 + k < n) {
          const unsigned slot = (w4[k >> 1] >> (16 * (k & 1))) & 0xffffu;
          const unsigned st = state[(slot & 15u) * kRowThreads];
          srow[(st & 0xffu) * kRowThreads] = (unsigned char)(slot >> 4);
          state[(slot & 15u) * kRowThreads] = (unsigned short)(st + 1u);
        }
      }
      cur = nxt;
    }
#pragma unroll
    for (int b = 0; b < 16; b++) { // heads back to the bucket starts
      const unsigned st = state[b * kRowThreads];
      state[b * kRowThreads] = (unsigned short)(st - (st >> 8));
    }
  }
  // ---- columns: 8 at a time as one coalesced 16-byte word
  const int hl = lane & 15;
  const int T = (n + 15) >> 4, last = n - 16 * (T - 1);
  int nexc = 0;
#pragma unroll
  for (int b = 0; b < 16; b++) {
    const int c = state[b * kRowThreads] >> 8;
    const int fe = T - ((((b - hl) & 15) >= last) ? 1 : 0); // first round of this bank that has no column
    nexc += max(0, c - fe);
  }
  const bool plain = nexc > rowcap; // (pathological rows: bank-sorted order as it is)
  if (!plain && nexc > 0) {
    int k = 0;
#pragma unroll
    for (int b = 0; b < 16; b++) {
      const unsigned st = state[b * kRowThreads];
      const int S = st & 0xffu, c = st >> 8;
      const int fe = T - ((((b - hl) & 15) >= last) ? 1 : 0);
      for (int tt = fe; tt < c; tt++) sexc[(k++) * kRowThreads] = (unsigned short)(((unsigned)srow[(S + tt) * kRowThreads] << 4) | (unsigned)b);
    }
  }
  int cursor = 0, pb = 0; // next excess entry; plain rows: the bucket that holds position q
  const int c8 = (nmax + 7) & ~7;
  uint4 *dst = reinterpret_cast<uint4 *>(a.ell_s) + ((size_t)t.tile * (a.maxrow >> 3)) * a.stride + ts;
  for (int q0 = 0; q0 < c8; q0 += 8) {
    unsigned wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int q = q0 + k;
      const unsigned b = (unsigned)(q + hl) & 15u;
      unsigned ent = b * 24u; // padding: the dummy atom of this lane's bank in this column
      if (q < n) {
        unsigned slot;
        if (!plain) {
          const unsigned st = state[b * kRowThreads];
          const int tt = q >> 4;
          if (tt < (int)(st >> 8)) slot = ((unsigned)srow[((st & 0xffu) + tt) * kRowThreads] << 4) | b;
          else slot = sexc[(cursor++) * kRowThreads];
        } else {
          while (q >= (int)((state[pb * kRowThreads] & 0xffu) + (state[pb * kRowThreads] >> 8))) pb++;
          slot = ((unsigned)srow[q * kRowThreads] << 4) | (unsigned)pb;
        }
        ent = (slot + kDummySlots) * 24u;
      }
      wv[k >> 1] |= ent << (16 * (k & 1));
    }
    dst[(size_t)(q0 >> 3) * a.stride] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
  }
  a.nell_s[rbase] = c8;
}

// counts[i] from the stored exact row lengths (a second list type asked of the same build)
__global__ void __launch_bounds__(kRowThreads) tiles_counts_kernel(TileArgs a, int *counts) {
  const size_t rbase = (size_t)blockIdx.x * a.stride + threadIdx.x;
  const int i = a.int_glob[rbase];
  if (i < a.n_local) counts[i] = a.ncsr[rbase];
}

// CSR entries / 2D table from the exact rows: slot number -> atom index, nothing else
enum { FILL_CSR = 0, FILL_2D = 1 };
struct FillArgs {
  const int *row_map; // CSR
  int *entries;       // CSR entries / 2D table
  int *num_neighs;    // 2D
  int maxneighs;      // 2D
};

// warp = 32 consecutive rows = (cell-sorted atoms) a contiguous piece of the CSR entries: the warp copies row after row,
// lane = entry, so the stores are coalesced
template <int MODE>
__global__ void __launch_bounds__(kRowThreads) tiles_fill_kernel(TileArgs a, FillArgs e) {
  const int tile = blockIdx.x, ts = threadIdx.x, lane = ts & 31;
  const size_t rbase = (size_t)tile * a.stride + ts;
  const int i = a.int_glob[rbase];
  const bool active = i < a.n_local;
  const int count = active ? a.ncsr[rbase] : 0;
  long long base = 0;
  if (active) base = (MODE == FILL_CSR) ? (long long)e.row_map[i] : (long long)i * e.maxneighs;
  const int nw = (MODE == FILL_2D) ? min(count, e.maxneighs) : count; // neighbor_2d.h:207-208
  if (MODE == FILL_2D && active) { e.num_neighs[i] = count; atomicMax(&a.flags[FL_MAX2D], count); }
  const unsigned short *tile16 = a.csr16 + ((size_t)tile * (a.maxrow >> 3)) * a.stride * 8;
  const int *__restrict__ jmap = a.stg_j + (size_t)tile * a.cap;
  const unsigned rows = __ballot_sync(0xffffffffu, nw > 0);
  for (unsigned m = rows; m; m &= m - 1) {
    const int src = __ffs(m) - 1;
    const int nw_r = __shfl_sync(0xffffffffu, nw, src);
    const long long base_r = __shfl_sync(0xffffffffu, base, src);
    const int ts_r = (ts & ~31) + src;
    for (int q = lane; q < nw_r; q += 32) {
      const unsigned slot = tile16[((size_t)(q >> 3) * a.stride + ts_r) * 8 + (q & 7)];
      e.entries[base_r + q] = __ldg(jmap + slot);
    }
  }
}

// ------------------------------------------------------------------------------ LJ force
struct LJOne { double lj1, lj2, cutsq, e1, e2, eshift; };
struct LJTab {
  double lj1[kMaxTypesConst * kMaxTypesConst], lj2[kMaxTypesConst * kMaxTypesConst], cutsq[kMaxTypesConst * kMaxTypesConst];
  double e1[kMaxTypesConst * kMaxTypesConst], e2[kMaxTypesConst * kMaxTypesConst], eshift[kMaxTypesConst * kMaxTypesConst];
  int ntypes;
};
// energy of a pair seen from one side (force_lj_neigh_impl.h:271-278 with fac = 0.5): 0.5 (r6inv (0.5 lj1 r6inv - lj2) / 6 - the
// same at the cutoff) = r6inv (e1 r6inv - e2) - eshift
inline void lj_energy_consts(double lj1, double lj2, double cutsq, double &e1, double &e2, double &eshift) {
  e1 = lj1 / 24.0; e2 = lj2 / 12.0;
  const double r2invc = 1.0 / cutsq, r6invc = r2invc * r2invc * r2invc;
  eshift = r6invc * (e1 * r6invc - e2);
}

// Four pairs at a time, branch-free and written stage by stage so that the four FP64 dependency chains are interleaved
// by the scheduler.  17 FP64 instructions per pair: 3 DADD, DMUL + 2 DFMA (rsq), 3 DFMA (reciprocal: MUFU.RCP64H seed
// y0 ~2^-20 and one cubic step y0 (1 + e + e^2), e = 1 - a y0, error e^3 ~ 2^-60), 2 DMUL (r6inv), DFMA + 2 DMUL (fpair),
// 3 DFMA (f).  Everything else is kept off the pair: an entry IS the byte offset of the neighbor's coordinates; the strict
// cutoff test (force_lj_neigh_impl.h:189) runs on the integer pipe (rsq and cutsq are non-negative, so the IEEE order is
// the order of the bit patterns) and its ONE consequence is a select on the high word of the reciprocal seed: a zero seed
// gives r2inv = 0 exactly and the pair contributes +-0.  Padding entries point at a far-away dummy atom and fail the
// same test; no entry is the atom itself.
template <bool ONETYPE, bool ENERGY>
__device__ __forceinline__ void lj_quad(const unsigned char *__restrict__ spb, const int *__restrict__ st, const unsigned w01, const unsigned w23,
                                        double x_i, double y_i, double z_i, int type_i, const LJOne &one, const LJTab *__restrict__ tab,
                                        double &fx, double &fy, double &fz, double &pe) {
  const unsigned off[4] = {w01 & 0xffffu, w01 >> 16, w23 & 0xffffu, w23 >> 16};
  double dx[4], dy[4], dz[4], rsq[4], lj1[4], lj2[4], cutsq[4];
  int tij[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const double *p = reinterpret_cast<const double *>(spb + off[u]);
    dx[u] = x_i - p[0]; dy[u] = y_i - p[1]; dz[u] = z_i - p[2];
    if (ONETYPE) { lj1[u] = one.lj1; lj2[u] = one.lj2; cutsq[u] = one.cutsq; tij[u] = 0; }
    else { tij[u] = type_i * tab->ntypes + st[off[u] / 24u]; lj1[u] = tab->lj1[tij[u]]; lj2[u] = tab->lj2[tij[u]]; cutsq[u] = tab->cutsq[tij[u]]; }
  }
#pragma unroll
  for (int u = 0; u < 4; u++) rsq[u] = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
  bool in[4];
  double r2inv[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    in[u] = __double_as_longlong(rsq[u]) < __double_as_longlong(cutsq[u]);
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(rsq[u]));
    y0 = __hiloint2double(in[u] ? __double2hiint(y0) : 0, 0);
    const double e = fma(-rsq[u], y0, 1.0);
    const double t = fma(e, e, e);
    r2inv[u] = fma(y0, t, y0);
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const double r6inv = r2inv[u] * r2inv[u] * r2inv[u];
    const double fpair = (r6inv * (lj1[u] * r6inv - lj2[u])) * r2inv[u];
    fx += dx[u] * fpair; fy += dy[u] * fpair; fz += dz[u] * fpair;
    if (ENERGY) { // force_lj_neigh_impl.h:271-278 with fac = 0.5 (every pair is seen from both sides); r6inv = 0 outside the cutoff
      const double e1 = ONETYPE ? one.e1 : tab->e1[tij[u]], e2 = ONETYPE ? one.e2 : tab->e2[tij[u]], es = ONETYPE ? one.eshift : tab->eshift[tij[u]];
      pe += fma(r6inv, fma(e1, r6inv, -e2), in[u] ? -es : 0.0);
    }
  }
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ldg_nc_v4(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// mbarriers (shared::cta).  `full`: the 32 lanes of the producer warp attach their cp.async groups (arrive.noinc: the
// arrival happens when the lane's copies have landed).  `empty`: one arrival per warp.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_on_copies(unsigned long long *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(addr), "r"(parity) : "memory");
}

constexpr int kForceThreads = kRowThreads;
constexpr int kForceWarps = kForceThreads / 32;
constexpr int kMaxBuf = 3;
constexpr int kProdUnroll = 16;

// Persistent CTAs (2 per SM), each walking the tiles a.order[first + blockIdx.x], [first + blockIdx.x + gridDim.x], ...
// below first + ntiles.  NBUF coordinate buffers; tile k of the CTA lives in buffer k mod NBUF.
//   producer = the last warp (its rows are the sparse tail of the tile): before it works on tile k it waits until every warp
//     has released the buffer of tile k - 1 (`empty`), then issues the copies of tile k + NBUF - 1 into it (LDGSTS through the
//     staging table; completion is attached to `full`);
//   every warp: waits for `full` of tile k, walks its 32 rows, releases the buffer.  No CTA-wide barrier.
// MODE_NVE: the epilogue also applies IntegratorNVE::final_integrate of this step and initial_integrate of the next one to the
// atom whose force was just accumulated (src/integrator_nve.cpp:47-74, 87-112: same operations, same order, so v and x are
// bit-identical to the separate kernels): v is updated in place, the new position goes to a SECOND position array because
// other CTAs still stage the old coordinates (the host swaps the two arrays after the launch).
struct NveFuse { double *v; double *x_new; const double *mass; double dtf, dtv; double *mv2_partial; };
// MODE_NVE_THERMO: MODE_NVE on a thermo step -- the pass also returns the potential energy (positions of this step) and
// sum m v^2 of the velocities between the two kicks, i.e. what Temperature / PotE / KinE (property_*.cpp) read after
// final_integrate, so that a thermo step keeps the fused integrator and needs no reduction pass of its own
enum { MODE_FORCE = 0, MODE_ENERGY = 1, MODE_NVE = 2, MODE_FORCE_ENERGY = 3, MODE_NVE_THERMO = 4 };

struct HaloGate { const int *flags; int seq, mask; }; // common.cuh: the neighbours' ghost stores of this step (comm_peer.cu)

template <bool ONETYPE, int MODE>
__global__ void __launch_bounds__(kForceThreads, 2) lj_tiles_kernel(TileArgs a, int first, int ntiles, int nbuf, unsigned ring_off, LJOne one, const LJTab *__restrict__ tab,
                                                                   double *__restrict__ f, double *__restrict__ pe_partial, NveFuse nve, HaloGate gate) {
  constexpr bool ENERGY = MODE == MODE_ENERGY || MODE == MODE_FORCE_ENERGY || MODE == MODE_NVE_THERMO;
  constexpr bool NVE = MODE == MODE_NVE || MODE == MODE_NVE_THERMO;
  __shared__ double s_red[kForceWarps], s_red2[kForceWarps];
  __shared__ unsigned long long s_full[kMaxBuf], s_empty[kMaxBuf];
  extern __shared__ __align__(16) unsigned char dyn[];
  const unsigned buf_bytes = (unsigned)(a.fcap + kDummySlots) * 24u;           // [16 dummy atoms + fcap][3] doubles: x,y,z of a staged atom adjacent
  int *const st0 = reinterpret_cast<int *>(dyn + (size_t)nbuf * buf_bytes);    // nbuf x [16 + fcap] types (multi-type systems only)
  const int tstride = a.fcap + kDummySlots;
  const int ts = threadIdx.x, lane = ts & 31, warp = ts >> 5;
  const int G = gridDim.x;
  const bool producer = warp == kForceWarps - 1;

  if (ts == 0) {
    for (int b = 0; b < nbuf; b++) { mbar_init(&s_full[b], 32u); mbar_init(&s_empty[b], (unsigned)kForceWarps); }
  }
  if (ts < kDummySlots * 3) // the dummy atoms of every buffer
    for (int b = 0; b < nbuf; b++) {
      reinterpret_cast<double *>(dyn + (size_t)b * buf_bytes)[ts] = kFarAway;
      if (!ONETYPE && ts < kDummySlots) st0[(size_t)b * tstride + ts] = 0;
    }
  __syncthreads();

  const int my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + G - 1) / G : 0;
  auto tile_of = [&](int k) { return a.order[first + blockIdx.x + k * G]; };
  // producer: copies of the CTA's kk-th tile (the caller has made sure that its buffer is free)
  bool halo_ok = gate.mask == 0; // producer warp: the ghost rows of this step have landed
  auto produce = [&](int kk) {
    const int b = kk % nbuf;
    if (kk < my_tiles) {
      const int tl = tile_of(kk);
      if (!halo_ok && a.has_ghost[tl]) {
        // The tiles that read no ghost come first in a.order, so the neighbours' stores (issued before this launch began) have
        // normally landed long before the first tile that needs them: the wait costs one flag read per phase, once per CTA.
        if (lane < 6 && ((gate.mask >> lane) & 1)) {
          int v;
          do { asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(gate.flags + lane) : "memory"); if (v < gate.seq) __nanosleep(100); } while (v < gate.seq);
        }
        __threadfence_system();
        __syncwarp();
        halo_ok = true;
      }
      const int *__restrict__ src = a.stg_j + (size_t)tl * a.cap;
      double *dstb = reinterpret_cast<double *>(dyn + (size_t)b * buf_bytes) + 3 * kDummySlots;
      int *dstt = st0 + (size_t)b * tstride + kDummySlots;
      // One descriptor per staged z-row, read with one coalesced load: in a tile without ghosts every row is a run of
      // consecutive atoms of x (owned atoms are cell-sorted) and is copied as consecutive doubles: one round of global latency
      // for the whole tile.  A tile that stages a ghost (or has one inside an interior bin) goes through the slot -> atom
      // table: dependent rounds of 16 indices per lane.
      const int nrows = (a.tx + 2) * (a.ty + 2);
      const int4 mine = lane < nrows ? a.rowdesc[(size_t)tl * kMaxStagedRows + lane] : make_int4(0, 0, 0, 0);
      const int4 mine2 = lane + 32 < nrows ? a.rowdesc[(size_t)tl * kMaxStagedRows + 32 + lane] : make_int4(0, 0, 0, 0);
      if (!__any_sync(0xffffffffu, (mine.z > 0 && mine.y < 0) || (mine2.z > 0 && mine2.y < 0))) {
        for (int r = 0; r < nrows; r++) {
          const int sl = r & 31;
          const int sb = __shfl_sync(0xffffffffu, r < 32 ? mine.x : mine2.x, sl), j0 = __shfl_sync(0xffffffffu, r < 32 ? mine.y : mine2.y, sl),
                    n = __shfl_sync(0xffffffffu, r < 32 ? mine.z : mine2.z, sl);
          const double *g = a.x + 3 * (size_t)j0;
          double *d = dstb + 3 * sb;
          for (int e = lane; e < 3 * n; e += 32) cp_async8(d + e, g + e);
          if (!ONETYPE) for (int e = lane; e < n; e += 32) cp_async4(dstt + sb + e, a.type + j0 + e);
        }
      } else {
        const int n = a.stg_n[tl];
        for (int s0 = 0; s0 < n; s0 += 32 * kProdUnroll) {
          int j[kProdUnroll];
#pragma unroll
          for (int u = 0; u < kProdUnroll; u++) { const int s = s0 + u * 32 + lane; j[u] = s < n ? __ldg(src + s) : -1; }
#pragma unroll
          for (int u = 0; u < kProdUnroll; u++) {
            if (j[u] >= 0) {
              const int s = s0 + u * 32 + lane;
              const double *g = a.x + 3 * (size_t)j[u];
              cp_async8(dstb + 3 * s, g); cp_async8(dstb + 3 * s + 1, g + 1); cp_async8(dstb + 3 * s + 2, g + 2);
              if (!ONETYPE) cp_async4(dstt + s, a.type + j[u]);
            }
          }
        }
      }
    }
    mbar_arrive_on_copies(&s_full[b]); // also for a tile that does not exist: nobody waits for it
  };

  if (producer)
    for (int kk = 0; kk < nbuf - 1; kk++) produce(kk);

  // per-thread row descriptors, one tile ahead
  int i_nxt = 0x7fffffff, own_nxt = 0, n_nxt = 0, tile_nxt = -1;
  auto load_desc = [&](int k) {
    i_nxt = 0x7fffffff; own_nxt = 0; n_nxt = 0; tile_nxt = -1;
    if (k < my_tiles) {
      tile_nxt = tile_of(k);
      const size_t rb = (size_t)tile_nxt * a.stride + ts;
      i_nxt = a.int_glob[rb];
      own_nxt = a.int_slot[rb];
      n_nxt = a.nell_s[rb];
    }
  };
  // The row words (16 bytes = 8 entries per thread and step) are read once per step, through two 16-byte shared-memory slots
  // per thread filled by cp.async (LDGSTS): word c+2 is requested right after word c was taken, no register is held while it
  // is in flight (at this register budget ptxas sinks a register load to the end of the loop body, next to its use), and
  // cp.async.wait_group 1 leaves word c+1 in flight.  Measured (2 M atoms, profiles/README.md, round 2): register load +
  // L1 prefetch two words ahead 0.320 ms (17 % of the stall samples on the first use of the word; only 47 % of those loads
  // hit the 23 KB of L1 left next to the coordinate buffers); two-word register look-ahead 0.329 ms; a per-warp ring of
  // 512-byte cp.async.bulk copies with one mbarrier per slot 0.42 ms at depth 3 (one CTA per SM) / 0.37 ms at depth 2
  // (+30 % instructions: elected-lane issue, try_wait, __syncwarp); this per-thread ring 0.316 ms (long-scoreboard 4.9 -> 2.8).
  auto row_of_tile = [&](int tl) { return reinterpret_cast<const uint4 *>(a.ell_s) + ((size_t)tl * (a.maxrow >> 3)) * a.stride + ts; };
  uint4 *const ering = reinterpret_cast<uint4 *>(dyn + ring_off) + ts; // two 16-byte slots per thread: ering[0], ering[kForceThreads]
  load_desc(0);
  double pe = 0.0;
  int b = 0;
  unsigned parity = 0u;
  double mv2 = 0.0;
  const double dtfm1 = (NVE && ONETYPE) ? nve.dtf / nve.mass[0] : 0.0; // integrator_nve.cpp:67,106 (one type: one mass)
  for (int k = 0; k < my_tiles; k++) {
    const int i_cur = i_nxt, own_cur = own_nxt, n_cur = n_nxt, tile = tile_nxt;
    const bool has_row = i_cur < a.n_local;
    const int nchunk = n_cur >> 3;
    const uint4 *row = row_of_tile(tile);
    const size_t rstep = (size_t)a.stride;
    // row words 0 and 1 are in flight while the warp waits for the coordinates
    if (has_row) {
      if (nchunk > 0) cp_async16(ering, row);
      cp_async_commit();
      if (nchunk > 1) cp_async16(ering + kForceThreads, row + rstep);
      cp_async_commit();
    }
    load_desc(k + 1);
    if (NVE && has_row) { prefetch_l1(nve.v + 3 * (size_t)i_cur); prefetch_l1(nve.v + 3 * (size_t)i_cur + 2); } // the epilogue's v
    if (producer) {
      // buffer (k + nbuf - 1) mod nbuf held tile k - 1
      if (k > 0) {
        const int kb = (k - 1) % nbuf;
        mbar_wait(&s_empty[kb], (unsigned)(((k - 1) / nbuf) & 1));
      }
      produce(k + nbuf - 1);
    }
    mbar_wait(&s_full[b], parity);
    const unsigned char *spb = dyn + (size_t)b * buf_bytes;
    const int *st = st0 + (size_t)b * tstride;
    const double *me = reinterpret_cast<const double *>(spb) + 3 * (own_cur + kDummySlots);
    const double x_i = me[0], y_i = me[1], z_i = me[2];
    const int type_i = ONETYPE ? 0 : st[own_cur + kDummySlots];
    double fx = 0.0, fy = 0.0, fz = 0.0;
    // one 16-byte word = 8 columns of the warp's schedule (n_cur is the same multiple of 8 in every lane of the warp)
    if (has_row) {
      for (int c = 0; c < nchunk; c++, row += rstep) {
        cp_async_wait_1(); // word c has landed (word c+1 may still be in flight)
        uint4 *slot = ering + (c & 1) * kForceThreads;
        const uint4 cur = *slot;
        if (c + 2 < nchunk) cp_async16(slot, row + 2 * rstep); // same thread: the read above precedes the asynchronous write
        cp_async_commit();
        lj_quad<ONETYPE, ENERGY>(spb, st, cur.x, cur.y, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
        lj_quad<ONETYPE, ENERGY>(spb, st, cur.z, cur.w, x_i, y_i, z_i, type_i, one, tab, fx, fy, fz, pe);
      }
    }
    if (has_row) {
      if (MODE != MODE_ENERGY) { f[3 * (size_t)i_cur] = fx; f[3 * (size_t)i_cur + 1] = fy; f[3 * (size_t)i_cur + 2] = fz; }
      if (NVE) {
        const double dtfm = ONETYPE ? dtfm1 : nve.dtf / nve.mass[type_i];
        double *vp = nve.v + 3 * (size_t)i_cur, *xp = nve.x_new + 3 * (size_t)i_cur;
        const double fi[3] = {fx, fy, fz}, xi[3] = {x_i, y_i, z_i};
        double vv = 0.0;
#pragma unroll
        for (int d = 0; d < 3; d++) {
          const double kick = __dmul_rn(dtfm, fi[d]);
          const double v1 = __dadd_rn(vp[d], kick);            // final_integrate :107-109
          const double v2 = __dadd_rn(v1, kick);               // initial_integrate of the next step :68-70
          vp[d] = v2;
          xp[d] = __dadd_rn(xi[d], __dmul_rn(nve.dtv, v2));    // :71-73
          if (MODE == MODE_NVE_THERMO) vv += v1 * v1;          // property_temperature.cpp:52-56 on the velocities thermo sees
        }
        if (MODE == MODE_NVE_THERMO) mv2 += vv * nve.mass[ONETYPE ? 0 : type_i];
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[b]);
    if (++b == nbuf) { b = 0; parity ^= 1u; }
  }
  if (producer) asm volatile("cp.async.wait_all;" ::: "memory"); // copies of tiles that do not exist were never issued; be tidy anyway
  if (ENERGY) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pe += __shfl_down_sync(0xffffffffu, pe, o);
    if (lane == 0) s_red[warp] = pe;
    if (MODE == MODE_NVE_THERMO) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mv2 += __shfl_down_sync(0xffffffffu, mv2, o);
      if (lane == 0) s_red2[warp] = mv2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < kForceWarps; w++) s += s_red[w];
      pe_partial[blockIdx.x] = s;
      if (MODE == MODE_NVE_THERMO) {
        double s2 = 0.0;
        for (int w = 0; w < kForceWarps; w++) s2 += s_red2[w];
        nve.mv2_partial[blockIdx.x] = s2;
      }
    }
  }
}

size_t lists_coords_bytes(int cap, int maxrow) { // region A of tiles_lists_kernel
  const size_t coords = (size_t)cap * 28, rows = (size_t)maxrow * kRowThreads + (size_t)kExcessRoom * kRowThreads * sizeof(unsigned short);
  return (std::max(coords, rows) + 15) / 16 * 16;
}
size_t lists_smem(int cap, int maxrow) { return lists_coords_bytes(cap, maxrow) + (size_t)16 * kRowThreads * sizeof(unsigned short); }
size_t force_coord_smem(int fcap, bool types, int nbuf) { return ((size_t)nbuf * ((size_t)(fcap + kDummySlots) * 24 + (types ? (size_t)(fcap + kDummySlots) * sizeof(int) : 0)) + 127) / 128 * 128; }
size_t force_smem(int fcap, bool types, int nbuf) { return force_coord_smem(fcap, types, nbuf) + (size_t)2 * kForceThreads * sizeof(uint4); }

} // namespace

namespace emd { int device_sum_partials(emd_ctx *ctx, const double *d_partial, int n, double *h_out); }

struct emd_tiles {
  TileArgs a;
  int ntiles = 0;
  bool valid = false;       // search done (masks + tables), parameters below describe it
  bool checked = false;     // its overflow flags have been read back
  bool rows_ready = false;  // ell_s of this build written
  int exact_key = -1;       // half | newton << 1 of the exact rows (csr16 / ncsr) of this build, -1: none
  unsigned short *d_ell = nullptr; size_t ell_cap = 0;
  int *d_nell = nullptr; size_t nell_cap = 0;
  unsigned short *d_csr16 = nullptr; size_t csr16_cap = 0;
  int *d_ncsr = nullptr; size_t ncsr_cap = 0;
  unsigned short *d_ell_s = nullptr; size_t ell_s_cap = 0;
  int *d_nell_s = nullptr; size_t nell_s_cap = 0;
  int *d_stg_j = nullptr; size_t stg_j_cap = 0;
  int *d_stg_n = nullptr; size_t stg_n_cap = 0;
  int4 *d_rowdesc = nullptr; size_t rowdesc_cap = 0;
  unsigned short *d_int_slot = nullptr; size_t int_slot_cap = 0;
  int *d_int_glob = nullptr; size_t int_glob_cap = 0;
  int *d_order = nullptr; size_t order_cap = 0;
  int *d_has_ghost = nullptr; size_t has_ghost_cap = 0;
  int n_free_tiles = 0;  // tiles that do not read the halo (first in d_order)
  bool all_owned_have_rows = false; // every owned atom is listed (precondition of the fused force + integrator launch)
  int *d_flags = nullptr;
  emd_ctx *ctx = nullptr; // the context of the last build (the lazy list / flag check of the const accessors runs on it)
  int num_sms = 148;
  LJTab *d_tab = nullptr, *h_tab = nullptr; // device copy of the pair table and its pinned staging copy
  unsigned long long tab_version = ~0ull; // ctx->lj version the device table was uploaded for
  double neigh_cut = 0.0;
  float thr2 = 0.f;
  int max_smem_optin = 0;
  int max_smem_sm = 0;
  // the inputs of the last build (a regrow re-runs the search)
  double m_density = 0.0;
  int staged_cells = 0;
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

// buffers for the current parameters (grow-only: a steady-state re-neighboring allocates nothing) and the search launches
int launch_search(emd_ctx *ctx, emd_tiles *t) {
  TileArgs &a = t->a;
  const size_t rows = (size_t)t->ntiles * a.stride;
  if (ensure_bytes((void **)&t->d_ell, &t->ell_cap, rows * a.maxrow * sizeof(unsigned short))) return 1;
  if (ensure_bytes((void **)&t->d_nell, &t->nell_cap, rows * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_csr16, &t->csr16_cap, rows * a.maxrow * sizeof(unsigned short))) return 1;
  if (ensure_bytes((void **)&t->d_ell_s, &t->ell_s_cap, rows * a.maxrow * sizeof(unsigned short))) return 1;
  if (ensure_bytes((void **)&t->d_ncsr, &t->ncsr_cap, rows * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_nell_s, &t->nell_s_cap, rows * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_stg_j, &t->stg_j_cap, (size_t)t->ntiles * a.cap * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_stg_n, &t->stg_n_cap, (size_t)t->ntiles * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_rowdesc, &t->rowdesc_cap, (size_t)t->ntiles * kMaxStagedRows * sizeof(int4))) return 1;
  if (ensure_bytes((void **)&t->d_int_slot, &t->int_slot_cap, rows * sizeof(unsigned short))) return 1;
  if (ensure_bytes((void **)&t->d_int_glob, &t->int_glob_cap, rows * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_order, &t->order_cap, (size_t)t->ntiles * sizeof(int))) return 1;
  if (ensure_bytes((void **)&t->d_has_ghost, &t->has_ghost_cap, (size_t)t->ntiles * sizeof(int))) return 1;
  a.ell = t->d_ell; a.nell = t->d_nell; a.csr16 = t->d_csr16; a.ncsr = t->d_ncsr; a.ell_s = t->d_ell_s; a.nell_s = t->d_nell_s;
  a.stg_j = t->d_stg_j; a.stg_n = t->d_stg_n; a.rowdesc = t->d_rowdesc; a.int_slot = t->d_int_slot; a.int_glob = t->d_int_glob;
  a.order = t->d_order; a.has_ghost = t->d_has_ghost; a.flags = t->d_flags;
  EMD_CUDA(cudaMemsetAsync(t->d_flags, 0, FL_COUNT * sizeof(int), ctx->stream));
  const size_t smem = (size_t)a.cap * sizeof(float4);
  if (smem > (size_t)t->max_smem_optin) return 3;
  if (set_smem(tiles_search_kernel, smem)) return 1;
  EMD_LAUNCH(ctx, tiles_search_kernel, t->ntiles, kSearchThreads, smem, a, t->thr2);
  EMD_LAUNCH(ctx, tiles_order_kernel, 1, 1024, 0, a, t->ntiles);
  t->valid = true; t->checked = false; t->rows_ready = false; t->exact_key = -1;
  return 0;
}

int launch_lists(emd_ctx *ctx, emd_tiles *t, bool exact, int half, int newton, int *d_counts) {
  TileArgs &a = t->a;
  ListArgs e;
  e.cutsq = t->neigh_cut * t->neigh_cut; e.half = half; e.newton = newton; e.counts = d_counts;
  const size_t cb = lists_coords_bytes(a.cap, a.maxrow), smem = lists_smem(a.cap, a.maxrow);
  if (smem > (size_t)t->max_smem_optin) return 3;
  if (exact) { if (set_smem(tiles_lists_kernel<true>, smem)) return 1; EMD_LAUNCH(ctx, tiles_lists_kernel<true>, t->ntiles, kRowThreads, smem, a, e, (unsigned)cb, kExcessRoom); }
  else { if (set_smem(tiles_lists_kernel<false>, smem)) return 1; EMD_LAUNCH(ctx, tiles_lists_kernel<false>, t->ntiles, kRowThreads, smem, a, e, (unsigned)cb, kExcessRoom); }
  t->rows_ready = true;
  if (exact) t->exact_key = (half ? 1 : 0) | (newton ? 2 : 0);
  return 0;
}

// Reads the flags of the search (+ lists) launches back -- the ONE host synchronisation of a re-neighboring -- and, if a
// capacity was exceeded, grows it and reports 2 (the caller re-runs the launches); 3: the fast path does not apply.
int check_flags(emd_ctx *ctx, emd_tiles *t, int n_local, const int *d_extra, int *h_extra) {
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, t->d_flags, FL_COUNT * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (d_extra) EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned + FL_COUNT, d_extra, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream));
  const int *h = ctx->h_pinned;
  if (h_extra && d_extra) *h_extra = h[FL_COUNT];
  const int bits = h[FL_BITS];
  TileArgs &a = t->a;
  if (bits == 0) {
    a.fcap = std::min(a.cap, std::max(32, (h[FL_NEED_CAP] + 31) / 32 * 32));
    t->n_free_tiles = h[FL_NFREE];
    t->all_owned_have_rows = h[FL_OWNED_ROWS] == n_local;
    t->checked = true;
    return 0;
  }
  t->valid = false;
  if (bits & OVF_INT) return 3; // a tile holds more atoms than rows: density far from the estimate
  if (bits & OVF_CAP) {
    if (h[FL_NEED_CAP] > kCapMax) return 3;
    a.cap = std::min(kCapMax, (h[FL_NEED_CAP] + h[FL_NEED_CAP] / 8 + 31) / 32 * 32);
  }
  if (bits & OVF_ROW) {
    a.maxrow = (h[FL_NEED_ROW] + h[FL_NEED_ROW] / 8 + 7) / 8 * 8;
    if (a.maxrow > kMaxRowLimit) return 3;
  }
  return 2;
}

// makes sure that the force rows (and, if asked, the exact rows of this list type) of the current build exist and that the
// build has been checked; d_counts (optional) receives the exact row lengths by atom.  0, 1 (error) or 3 (not applicable).
int ensure_lists(emd_ctx *ctx, emd_tiles *t, bool exact, int half, int newton, int *d_counts) {
  const int key = (half ? 1 : 0) | (newton ? 2 : 0);
  for (int attempt = 0; attempt < 5; attempt++) {
    if (!t->valid) { const int rc = launch_search(ctx, t); if (rc) return rc; }
    bool launched = false;
    if (!t->rows_ready || (exact && t->exact_key != key)) {
      const int rc = launch_lists(ctx, t, exact, half, newton, d_counts);
      if (rc) return rc;
      launched = true;
    } else if (exact && d_counts) {
      EMD_LAUNCH(ctx, tiles_counts_kernel, t->ntiles, kRowThreads, 0, t->a, d_counts);
    }
    if (t->checked && !launched) return 0;
    const int rc = check_flags(ctx, t, t->a.n_local, nullptr, nullptr);
    if (rc == 0) return 0;
    if (rc != 2) return rc;
  }
  return 3;
}

} // namespace

extern "C" {

// CUDA loads a kernel lazily at its first launch (~1 ms each): the variants that a run meets late (the thermo step's force +
// energy launch, the split launches of a decomposed run) are loaded here instead of inside somebody's timed region
static void warm_kernels() {
  static bool done = false;
  if (done) return;
  done = true;
  cudaFuncAttributes at;
#define EMD_WARM(K) (void)cudaFuncGetAttributes(&at, K)
  EMD_WARM((lj_tiles_kernel<true, MODE_FORCE>)); EMD_WARM((lj_tiles_kernel<true, MODE_ENERGY>)); EMD_WARM((lj_tiles_kernel<true, MODE_NVE>));
  EMD_WARM((lj_tiles_kernel<true, MODE_FORCE_ENERGY>)); EMD_WARM((lj_tiles_kernel<false, MODE_FORCE>)); EMD_WARM((lj_tiles_kernel<false, MODE_ENERGY>));
  EMD_WARM((lj_tiles_kernel<false, MODE_NVE>)); EMD_WARM((lj_tiles_kernel<false, MODE_FORCE_ENERGY>));
  EMD_WARM((lj_tiles_kernel<true, MODE_NVE_THERMO>)); EMD_WARM((lj_tiles_kernel<false, MODE_NVE_THERMO>));
  EMD_WARM(tiles_search_kernel); EMD_WARM(tiles_order_kernel); EMD_WARM(tiles_lists_kernel<true>); EMD_WARM(tiles_lists_kernel<false>);
  EMD_WARM(tiles_counts_kernel); EMD_WARM(tiles_fill_kernel<FILL_CSR>); EMD_WARM(tiles_fill_kernel<FILL_2D>);
#undef EMD_WARM
  (void)cudaGetLastError();
}

int emd_tiles_create(emd_tiles **out) {
  if (!out) { set_error("emd_tiles_create: out == NULL"); return 1; }
  warm_kernels();
  emd_tiles *t = new emd_tiles();
  memset(&t->a, 0, sizeof t->a);
  EMD_CUDA(cudaMalloc((void **)&t->d_flags, FL_COUNT * sizeof(int)));
  EMD_CUDA(cudaMalloc((void **)&t->d_tab, sizeof(LJTab)));
  EMD_CUDA(cudaMallocHost((void **)&t->h_tab, sizeof(LJTab)));
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
  void *bufs[] = {t->d_ell, t->d_nell, t->d_csr16, t->d_ncsr, t->d_ell_s, t->d_nell_s, t->d_stg_j, t->d_stg_n, t->d_rowdesc, t->d_int_slot, t->d_int_glob,
                  t->d_order, t->d_has_ghost, t->d_flags, t->d_tab};
  for (void *p : bufs) if (p) cudaFree(p);
  if (t->h_tab) cudaFreeHost(t->h_tab);
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

int emd_tiles_lists(const emd_tiles *t, const unsigned short **d_csr16, const int **d_ncsr, const unsigned short **d_ell_s,
                    const int **d_nell_s, const unsigned short **d_int_slot, const int **d_stg_j) {
  if (!t || !t->valid || !t->rows_ready) { set_error("emd_tiles_lists: lists not built"); return 1; }
  if (d_csr16) *d_csr16 = t->exact_key >= 0 ? t->a.csr16 : nullptr;
  if (d_ncsr) *d_ncsr = t->exact_key >= 0 ? t->a.ncsr : nullptr;
  if (d_ell_s) *d_ell_s = t->a.ell_s;
  if (d_nell_s) *d_nell_s = t->a.nell_s;
  if (d_int_slot) *d_int_slot = t->a.int_slot;
  if (d_stg_j) *d_stg_j = t->a.stg_j;
  return 0;
}

// Starts a build: tile shape and capacities from the mean density, then the search launches (no host synchronisation:
// the overflow flags are read back together with the row total in emd_neigh_tiles_count, or before the first force launch).
// Returns 0 on success, 3 if the fast path does not apply to this configuration (the caller then uses
// emd_neigh_csr_* / emd_neigh_2d_fill), 1 on error.
int emd_neigh_tiles_build(emd_ctx *ctx, emd_tiles *t, const double *d_x, int n_local, int n_all, const emd_bin_geom *g,
                          const int *d_bincount, const int *d_binoffsets, const int *d_permute, double neigh_cut) {
  t->valid = false;
  t->ctx = ctx;
  const int nix = g->nbinx - 2 * g->nhalo, niy = g->nbiny - 2 * g->nhalo, niz = g->nbinz - 2 * g->nhalo;
  if (nix <= 0 || niy <= 0 || niz <= 0 || n_local <= 0) return 3;
  const double wx = (g->maxx - g->minx) / g->nbinx, wy = (g->maxy - g->miny) / g->nbiny, wz = (g->maxz - g->minz) / g->nbinz;
  // the 27-bin stencil must cover the list radius (true for every reference configuration:
  // bins are at least neigh_cut wide, binning_kksort.cpp:77-87); the reference itself simply
  // misses pairs otherwise, and the generic kernels reproduce that
  if (neigh_cut > wx || neigh_cut > wy || neigh_cut > wz) return 3;
  TileArgs &a = t->a;
  const bool same_grid = a.nbx == g->nbinx && a.nby == g->nbiny && a.nbz == g->nbinz && a.nhalo == g->nhalo && t->neigh_cut == neigh_cut;
  a.nbx = g->nbinx; a.nby = g->nbiny; a.nbz = g->nbinz; a.nhalo = g->nhalo;
  a.n_local = n_local;
  a.bincount = d_bincount; a.binoffsets = d_binoffsets; a.permute = d_permute;
  a.x = d_x; a.type = nullptr;
  a.ox = g->minx; a.oy = g->miny; a.oz = g->minz;
  a.wx = wx; a.wy = wy; a.wz = wz;
  a.stride = kRowThreads;
  // mean atoms per cell -> tile shape with ~0.9*stride atoms whose halo fits in shared memory
  const double m = std::max(1e-3, (double)n_all / ((double)g->nbinx * g->nbiny * g->nbinz));
  static const int shapes[][3] = {{4, 4, 8}, {4, 4, 4}, {2, 4, 4}, {2, 2, 8}, {2, 2, 4}, {2, 2, 2}, {1, 2, 2}, {1, 1, 2}, {1, 1, 1}};
  int pick = -1;
  for (int s = 0; s < (int)(sizeof shapes / sizeof shapes[0]); s++) {
    const int *d = shapes[s];
    const double atoms = m * d[0] * d[1] * d[2], staged = m * (d[0] + 2) * (d[1] + 2) * (d[2] + 2);
    if ((d[0] + 2) * (d[1] + 2) * (d[2] + 2) > kMaxStagedCells || d[0] * d[1] * d[2] > kMaxInteriorCells) continue;
    if (atoms <= 0.9 * a.stride && staged * 1.2 + 64 <= kCapMax) { pick = s; break; }
  }
  if (pick < 0) return 3;
  const int tx = std::min(shapes[pick][0], nix), ty = std::min(shapes[pick][1], niy), tz = std::min(shapes[pick][2], niz);
  const bool same_shape = same_grid && tx == a.tx && ty == a.ty && tz == a.tz;
  a.tx = tx; a.ty = ty; a.tz = tz;
  a.ntx = (nix + a.tx - 1) / a.tx; a.nty = (niy + a.ty - 1) / a.ty; a.ntz = (niz + a.tz - 1) / a.tz;
  const long long ntiles_ll = (long long)a.ntx * a.nty * a.ntz;
  if (ntiles_ll > 0x7fffffffLL) return 3;
  t->ntiles = (int)ntiles_ll;
  const int staged_cells = (a.tx + 2) * (a.ty + 2) * (a.tz + 2);
  // capacities: from the density with head room; a build on the same grid keeps what an earlier build had to grow to
  const int cap = std::min(kCapMax, ((int)(m * staged_cells * 1.25) + 64 + 31) / 32 * 32);
  const double rho = m / (wx * wy * wz);
  int maxrow = (int)(rho * 4.18879020478639 * neigh_cut * neigh_cut * neigh_cut * 1.35) + 8; // expected full-list row + 35 %
  maxrow = std::min(kMaxRowLimit, std::max(16, (maxrow + 7) / 8 * 8));
  if (same_shape) { a.cap = std::max(a.cap, cap); a.maxrow = std::max(a.maxrow, maxrow); }
  else { a.cap = cap; a.maxrow = maxrow; }
  // Rounding margin of the FP32 search: coordinates are relative to the centre of the staged region, |p|^2 <= R2; the
  // expanded form w_j + p_i.P_j carries ~4 roundings of magnitude <= 2 R2 plus the input rounding of p (2 r |dp|).
  const double R2 = 0.25 * ((a.tx + 2) * wx * (a.tx + 2) * wx + (a.ty + 2) * wy * (a.ty + 2) * wy + (a.tz + 2) * wz * (a.tz + 2) * wz);
  const double thr2 = neigh_cut * neigh_cut * (1.0 + 1e-5) + 64.0 * R2 / 8388608.0;
  t->thr2 = nextafterf((float)thr2, INFINITY);
  t->neigh_cut = neigh_cut;
  return launch_search(ctx, t);
}

int emd_neigh_tiles_count(emd_ctx *ctx, emd_tiles *t, int half, int newton, int *d_row_map, int *h_total) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_count: tiles not built"); return 1; }
  for (int attempt = 0; attempt < 5; attempt++) {
    TileArgs &a = t->a;
    EMD_CUDA(cudaMemsetAsync(d_row_map, 0, sizeof(int) * ((size_t)a.n_local + 1), ctx->stream));
    if (!t->valid) { const int rc = launch_search(ctx, t); if (rc) return rc; }
    const int key = (half ? 1 : 0) | (newton ? 2 : 0);
    if (!t->rows_ready || t->exact_key != key) { const int rc = launch_lists(ctx, t, true, half, newton, d_row_map); if (rc) return rc; }
    else EMD_LAUNCH(ctx, tiles_counts_kernel, t->ntiles, kRowThreads, 0, a, d_row_map);
    if (exclusive_scan_int(ctx, d_row_map, d_row_map, a.n_local + 1, nullptr)) return 1;
    int total = 0;
    const int rc = check_flags(ctx, t, a.n_local, d_row_map + a.n_local, &total);
    if (rc == 2) continue; // a capacity grew: search and lists again
    if (rc) return rc;
    if (h_total) *h_total = total;
    if (total < 0) { set_error("emd_neigh_tiles_count: neighbor count overflows 32-bit row_map"); return 2; }
    return 0;
  }
  return 3;
}

int emd_neigh_tiles_fill_csr(emd_ctx *ctx, emd_tiles *t, int half, int newton, const int *d_row_map, int *d_entries) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_fill_csr: tiles not built"); return 1; }
  const int rc = ensure_lists(ctx, t, true, half, newton, nullptr);
  if (rc) { if (rc == 3) set_error("emd_neigh_tiles_fill_csr: tile lists not available"); return rc; }
  FillArgs e;
  memset(&e, 0, sizeof e);
  e.row_map = d_row_map; e.entries = d_entries;
  EMD_LAUNCH(ctx, tiles_fill_kernel<FILL_CSR>, t->ntiles, kRowThreads, 0, t->a, e);
  return 0;
}

int emd_neigh_tiles_fill_2d(emd_ctx *ctx, emd_tiles *t, int half, int newton, int maxneighs, int *d_num_neighs, int *d_neighs,
                            int *h_max_count) {
  if (!t || !t->valid) { set_error("emd_neigh_tiles_fill_2d: tiles not built"); return 1; }
  const int rc = ensure_lists(ctx, t, true, half, newton, nullptr);
  if (rc) { if (rc == 3) set_error("emd_neigh_tiles_fill_2d: tile lists not available"); return rc; }
  TileArgs &a = t->a;
  FillArgs e;
  memset(&e, 0, sizeof e);
  e.entries = d_neighs; e.num_neighs = d_num_neighs; e.maxneighs = maxneighs;
  EMD_CUDA(cudaMemsetAsync(t->d_flags + FL_MAX2D, 0, sizeof(int), ctx->stream));
  EMD_CUDA(cudaMemsetAsync(d_num_neighs, 0, sizeof(int) * ((size_t)a.n_local + 1), ctx->stream));
  EMD_LAUNCH(ctx, tiles_fill_kernel<FILL_2D>, t->ntiles, kRowThreads, 0, a, e);
  if (h_max_count) {
    EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, t->d_flags + FL_MAX2D, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    *h_max_count = ctx->h_pinned[0];
  }
  return 0;
}

} // extern "C"

// ForceLJNeigh::compute / compute_energy on the tile lists.  d_x/d_type are the CURRENT arrays
// (atoms keep their indices between rebuilds; the binning arrays captured at build time must
// still be alive).  With h_pe != NULL only the energy is computed (forces untouched).
// part: 0 = every tile; 1 = only the tiles that do not read the halo (their forces are final before the halo exchange of
// this step has landed); 2 = the rest.  reserve_ctas > 0 leaves that many CTA slots of the persistent grid free, so that
// the pack and transport kernels of a concurrent halo exchange find room on the SMs.
static int lj_tiles_launch(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *h_pe, int part,
                           int reserve_ctas, const NveFuse *fuse = nullptr, bool force_too = false, double *h_mv2 = nullptr) {
  if (!t || !t->valid) { set_error("emd_force_lj_compute_tiles: tiles not built"); return 1; }
  if (ctx->lj.ntypes == 0) { set_error("emd_force_lj_compute_tiles: parameters not set"); return 1; }
  if (part < 0 || part > 2 || (h_pe && part != 0)) { set_error("emd_force_lj_compute_tiles: bad part"); return 1; }
  if (fuse && (h_pe != nullptr) != (h_mv2 != nullptr)) { set_error("emd_force_lj_compute_tiles: the fused thermo launch returns both sums"); return 1; }
  if (!t->rows_ready || !t->checked) { // a build that nobody asked a CSR / 2D list of
    const int rc = ensure_lists(ctx, t, false, 0, 0, nullptr);
    if (rc) { if (rc == 3) set_error("emd_force_lj_compute_tiles: tile lists not available"); return rc == 3 ? 1 : rc; }
  }
  // an owned atom outside the interior bins has no row (the reference leaves it without neighbors, neighbor_csr.h:184):
  // its force is the zero of the reference's deep_copy(f, 0)
  if (!t->all_owned_have_rows && (!h_pe || force_too) && part == 0) EMD_CUDA(cudaMemsetAsync(d_f, 0, sizeof(double) * 3 * (size_t)t->a.n_local, ctx->stream));
  NveFuse nve = fuse ? *fuse : NveFuse{nullptr, nullptr, nullptr, 0.0, 0.0, nullptr};
  TileArgs a = t->a;
  a.x = d_x; a.type = d_type;
  const bool one = ctx->lj.ntypes == 1;
  LJOne p1 = {ctx->lj.lj1[0], ctx->lj.lj2[0], ctx->lj.cutsq[0], 0.0, 0.0, 0.0};
  lj_energy_consts(p1.lj1, p1.lj2, p1.cutsq, p1.e1, p1.e2, p1.eshift);
  if (!one && t->tab_version != ctx->lj_version) { // the pair table travels once per emd_force_lj_set_params
    LJTab *h = t->h_tab;
    h->ntypes = ctx->lj.ntypes;
    memcpy(h->lj1, ctx->lj.lj1, sizeof h->lj1); memcpy(h->lj2, ctx->lj.lj2, sizeof h->lj2); memcpy(h->cutsq, ctx->lj.cutsq, sizeof h->cutsq);
    for (int k = 0; k < h->ntypes * h->ntypes; k++) lj_energy_consts(h->lj1[k], h->lj2[k], h->cutsq[k], h->e1[k], h->e2[k], h->eshift[k]);
    EMD_CUDA(cudaMemcpyAsync(t->d_tab, h, sizeof *h, cudaMemcpyHostToDevice, ctx->stream));
    EMD_CUDA(cudaStreamSynchronize(ctx->stream));
    t->tab_version = ctx->lj_version;
  }
  // three coordinate buffers if two CTAs still fit on one SM (1 KB per CTA is reserved by the system), else two
  int nbuf = kMaxBuf;
  if (2 * (force_smem(a.fcap, !one, 3) + 1024) > (size_t)t->max_smem_sm) nbuf = 2;
  const size_t smem = force_smem(a.fcap, !one, nbuf);
  if (smem > (size_t)t->max_smem_optin) { set_error("emd_force_lj_compute_tiles: tile does not fit in shared memory"); return 1; }
  // a pending halo gate travels with the single launch; part 1 reads no ghost; part 2 has to wait on the stream first
  HaloGate gate = {nullptr, 0, 0};
  if (ctx->gate_pending && part == 0) { gate.flags = ctx->gate_flags; gate.seq = ctx->gate_seq; gate.mask = ctx->gate_mask; ctx->gate_pending = false; }
  else if (ctx->gate_pending && part == 2) { if (emd_ctx_halo_gate_wait(ctx)) return 1; }
  const int first = part == 2 ? t->n_free_tiles : 0;
  const int count = part == 0 ? t->ntiles : part == 1 ? t->n_free_tiles : t->ntiles - t->n_free_tiles;
  if (count <= 0) return 0;
  const int grid = std::max(1, std::min(count, 2 * emd_ctx_side_sms(ctx) - std::max(0, reserve_ctas)));
  double *partial = nullptr;
  if (h_pe) {
    if (ctx->s_c.ensure(sizeof(double) * (2 * (size_t)grid + 16))) return 1;
    partial = ctx->s_c.as<double>() + 8;
    nve.mv2_partial = partial + grid + 8;
  }
#define EMD_LJ_TILES(ONE, MD)                                                                                              \
  do {                                                                                                                     \
    if (set_smem(lj_tiles_kernel<ONE, MD>, smem)) return 1;                                                                \
    EMD_LAUNCH(ctx, (lj_tiles_kernel<ONE, MD>), grid, kForceThreads, smem, a, first, count, nbuf, (unsigned)force_coord_smem(a.fcap, !one, nbuf), p1, t->d_tab, d_f, partial, nve, gate); \
  } while (0)
  if (fuse && h_pe) { if (one) EMD_LJ_TILES(true, MODE_NVE_THERMO); else EMD_LJ_TILES(false, MODE_NVE_THERMO); }
  else if (fuse) { if (one) EMD_LJ_TILES(true, MODE_NVE); else EMD_LJ_TILES(false, MODE_NVE); }
  else if (h_pe && force_too) { if (one) EMD_LJ_TILES(true, MODE_FORCE_ENERGY); else EMD_LJ_TILES(false, MODE_FORCE_ENERGY); }
  else if (h_pe) { if (one) EMD_LJ_TILES(true, MODE_ENERGY); else EMD_LJ_TILES(false, MODE_ENERGY); }
  else { if (one) EMD_LJ_TILES(true, MODE_FORCE); else EMD_LJ_TILES(false, MODE_FORCE); }
#undef EMD_LJ_TILES
  if (h_pe && h_mv2) { if (int rc = device_sum_partials(ctx, nve.mv2_partial, grid, h_mv2)) return rc; }
  if (h_pe) return device_sum_partials(ctx, partial, grid, h_pe);
  return 0;
}

extern "C" {

int emd_force_lj_compute_tiles(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *h_pe) {
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, h_pe, 0, 0);
}

int emd_force_lj_compute_tiles_with_energy(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *h_pe) {
  if (!h_pe) { set_error("emd_force_lj_compute_tiles_with_energy: h_pe == NULL"); return 1; }
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, h_pe, 0, 0, nullptr, true);
}

int emd_force_lj_compute_tiles_part(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, int part,
                                    int reserve_ctas) {
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, part, reserve_ctas);
}

int emd_force_lj_compute_tiles_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *d_v,
                                   double *d_x_new, const double *d_mass, double dtf, double dtv) {
  if (t && t->valid && t->checked && !t->all_owned_have_rows) return 3; // an owned atom has no row: its position would not be advanced
  if (!d_v || !d_x_new || !d_mass || d_x_new == d_x) { set_error("emd_force_lj_compute_tiles_nve: v, mass and a second position array are required"); return 1; }
  const NveFuse nve = {d_v, d_x_new, d_mass, dtf, dtv, nullptr};
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, 0, 0, &nve);
}

int emd_force_lj_compute_tiles_nve_thermo(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, double *d_v,
                                          double *d_x_new, const double *d_mass, double dtf, double dtv, double *h_pe, double *h_mv2) {
  if (t && t->valid && t->checked && !t->all_owned_have_rows) return 3;
  if (!d_v || !d_x_new || !d_mass || d_x_new == d_x || !h_pe || !h_mv2) { set_error("emd_force_lj_compute_tiles_nve_thermo: v, mass, a second position array and both results are required"); return 1; }
  const NveFuse nve = {d_v, d_x_new, d_mass, dtf, dtv, nullptr};
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, h_pe, 0, 0, &nve, true, h_mv2);
}

int emd_force_lj_compute_tiles_part_nve(emd_ctx *ctx, emd_tiles *t, const double *d_x, const int *d_type, double *d_f, int part,
                                        int reserve_ctas, double *d_v, double *d_x_new, const double *d_mass, double dtf, double dtv) {
  if (t && t->valid && t->checked && !t->all_owned_have_rows) return 3;
  if (!d_v || !d_x_new || !d_mass || d_x_new == d_x) { set_error("emd_force_lj_compute_tiles_part_nve: v, mass and a second position array are required"); return 1; }
  const NveFuse nve = {d_v, d_x_new, d_mass, dtf, dtv, nullptr};
  return lj_tiles_launch(ctx, t, d_x, d_type, d_f, nullptr, part, reserve_ctas, &nve);
}

// the build is asynchronous: the first question about its outcome makes the force rows and reads the flags back
static int settle(const emd_tiles *tc, const char *who) {
  emd_tiles *t = const_cast<emd_tiles *>(tc);
  if (!t || !t->valid) { set_error("%s: tiles not built", who); return 1; }
  if (t->checked && t->rows_ready) return 0;
  const int rc = ensure_lists(t->ctx, t, false, 0, 0, nullptr);
  if (rc == 3) set_error("%s: tile lists not available", who);
  return rc ? 1 : 0;
}

int emd_tiles_complete(const emd_tiles *t, int *all_owned_have_rows) {
  if (settle(t, "emd_tiles_complete")) return 1;
  if (all_owned_have_rows) *all_owned_have_rows = t->all_owned_have_rows ? 1 : 0;
  return 0;
}

int emd_tiles_halo_split(const emd_tiles *t, int *n_free, int *n_halo) {
  if (settle(t, "emd_tiles_halo_split")) return 1;
  // an incomplete build is not split: the single launch zeroes the rows that no tile writes
  if (n_free) *n_free = t->all_owned_have_rows ? t->n_free_tiles : 0;
  if (n_halo) *n_halo = t->all_owned_have_rows ? t->ntiles - t->n_free_tiles : t->ntiles;
  return 0;
}

} // extern "C"
