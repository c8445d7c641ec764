// snap.cu -- SNAP bispectrum force on B200 (FP64).
//
// Replaces ForceSNAP<>::compute (src/force_types/force_snap_neigh_impl.h:159-208, team kernel
// :589-725) and the SNA class it drives (src/force_types/sna_impl.hpp).  The reference evaluates,
// per atom, U_tot -> Z(j1,j2,j) -> and per neighbor dU -> dB(j1,j2,j) -> sum_k beta_k dB_k, with
// 0.94 MB of global scratch per atom in flight.  Here the same force is evaluated in the adjoint
// order (the one BASELINE.json names: ui / yi / duidrj / deidrj):
//
//   Y(j)      = sum_{(j1>=j2)} betaj(j1,j2,j) Z(j1,j2,j)          (beta folded in once per atom)
//   F_ij      = 2 sum_{j,mb<=j/2,ma} w(j,mb,ma) Re( conj(dU_j(mb,ma)) Y_j(mb,ma) ) - 1.5e6 rij/r^14
//
// which is the reference's sum regrouped (compute_dbidrj's three conj(dU).Z sums, sna_impl.hpp:
// 393-527, each land on the Z block whose LAST index is the dU level); betaj carries the
// (j+1)/(j1+1), (j+1)/(j2+1) factors and w the half-column weights (1, 1/2 on the diagonal of the
// middle column, 0 below it).  No per-neighbor dB and no Z array exist: per atom the state is
// U_tot and Y on the half range mb <= j/2 (155 complex numbers each at 2J=8 instead of 2 x 9^5).
//
// Kernels (FP64 FMA pipe is the bound; nothing here is a dense contraction, so no tensor cores):
//   snap_pairs_*   in-cutoff pair list (count, scan, fill): lanes stay dense in the two pair kernels
//   snap_ui        lanes = 32 atoms, one warp per column mb of the Wigner recursion (VMK 4.8.2,
//                  compute_uarray :641-720); a column is independent of the others except for its
//                  first level, which each warp re-derives (levels 2k,2k+1 of columns k<mb: every
//                  warp ends up with the same 45 element updates per neighbor at 2J=8); U_tot is
//                  accumulated in shared memory without atomics or cross-lane reductions
//   snap_yi        lanes = 32 atoms, U_tot of the batch expanded to the full (ma,mb) range in
//                  146 KB of shared memory ([element][lane], conflict-free); warps take output
//                  strips (j, ma) from a cost-sorted queue and run the Clebsch-Gordan double sum
//                  (compute_zi :196-283) for every mb of the strip, accumulating beta*Z in registers
//   snap_deidrj    lanes = in-cutoff pairs; the dU recursion (compute_duarray :728-893) runs
//                  column by column IN PLACE in registers (9 + 3x9 complex) and every finished
//                  level is contracted with Y on the fly; f_i += F_ij, f_j -= F_ij with RED.F64
//                  (57 per atom-step: three orders of magnitude below the RED rate of the part)
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

using namespace emd;

namespace emd { int device_sum_partials(emd_ctx *ctx, const double *d_partial, int n, double *h_out); }

namespace {

constexpr int kMaxJ = 8;              // twojmax <= 8
constexpr int kMaxCol = kMaxJ / 2 + 1;
constexpr int kMaxTriples = 125;
constexpr int kMaxHalf = 155;       // U elements (j, mb <= j/2, ma) at twojmax = 8
constexpr int kMaxItems = 64;      // snap_yi work items: (group of one or two output rows) x (half of the blocks of its level)
constexpr int kRootDim = kMaxJ + 2;   // rootpq[p][q], p,q in 0..twojmax+1
constexpr double kPi = 3.14159265358979323846;

struct Triple { short j1, j2, j, pad; int cgoff; };

struct SnapTab {
  int twojmax, ncol, nuh, nuf, ntriples, nitems, ncg, ntypes, nelements, switchflag;
  double rcutfac, rfac0, rmin0, wself, cutsq;
  double unit_min_ar;        // snap_deidrj: smallest |a_r| for the unit-tangent recursion (EMD_SNAP_DEIDRJ_DIRECT=1: never)
  int uh_block[kMaxJ + 1];   // half layout: (j,mb,ma), mb <= j/2, at uh_block[j] + mb*(j+1) + ma
  int uf_block[kMaxJ + 1];   // full layout: (j,ma,mb) at uf_block[j] + ma*(j+1) + mb  (the reference's u(j,ma,mb))
  double rootpq[kRootDim * kRootDim];
  Triple triple[kMaxTriples];         // sorted by j
  int tri_begin[kMaxJ + 2];           // blocks of output level j: [tri_begin[j], tri_begin[j+1]) (twojmax = 8 list)
  // snap_yi work items, most expensive first: output rows (j, ma) [and (j, ma+1) if rows == 2], nmb outputs each, written to
  // Y array `half`; segments (YiSeg) [seg[0],seg[1]) advance both rows at once, [seg[1],seg[2]) only the first, [seg[2],seg[3]) only the second
  struct YiItem { short j, ma, rows, nmb, half, pad; int seg[4]; } item[kMaxItems];
  // snap_yi's expansion of the half range to the full one: half element e goes to full element exp_dst[e] and, unless it lies
  // on the middle column, its inversion image (conjugated, sign exp_img[e] < 0 ? -1 : +1) to |exp_img[e]| - 1
  short exp_dst[kMaxHalf], exp_img[kMaxHalf];
  int elem_of_type[kMaxTypesConst];
  double radelem[kMaxTypesConst], wjelem[kMaxTypesConst];
};

// ------------------------------------------------------------------ host: index lists and tables
int imin(int a, int b) { return a < b ? a : b; }
int imax(int a, int b) { return a > b ? a : b; }

double factorial(int n) { // sna_impl.hpp:962-968
  double r = 1.0;
  for (int i = 1; i <= n; i++) r *= 1.0 * i;
  return r;
}
double deltacg(int j1, int j2, int j) { // sna_impl.hpp:975-981
  const double sfaccg = factorial((j1 + j2 + j) / 2 + 1);
  return sqrt(factorial((j1 + j2 - j) / 2) * factorial((j1 - j2 + j) / 2) * factorial((-j1 + j2 + j) / 2) / sfaccg);
}
// Clebsch-Gordan coefficient cgarray(j1,j2,j,m1,m2), quasi-binomial formula VMK 8.2.1(3) (sna_impl.hpp:991-1046)
double clebsch_gordan(int j1, int j2, int j, int m1, int m2) {
  const int aa2 = 2 * m1 - j1, bb2 = 2 * m2 - j2;
  const int m = (aa2 + bb2 + j) / 2;
  if (m < 0 || m > j) return 0.0;
  double sum = 0.0;
  for (int z = imax(0, imax(-(j - j2 + aa2) / 2, -(j - j1 - bb2) / 2)); z <= imin((j1 + j2 - j) / 2, imin((j1 - aa2) / 2, (j2 + bb2) / 2)); z++) {
    const int ifac = z % 2 ? -1 : 1;
    sum += ifac / (factorial(z) * factorial((j1 + j2 - j) / 2 - z) * factorial((j1 - aa2) / 2 - z) * factorial((j2 + bb2) / 2 - z) *
                   factorial((j - j2 + aa2) / 2 + z) * factorial((j - j1 - bb2) / 2 + z));
  }
  const int cc2 = 2 * m - j;
  const double dcg = deltacg(j1, j2, j);
  const double sfaccg = sqrt(factorial((j1 + aa2) / 2) * factorial((j1 - aa2) / 2) * factorial((j2 + bb2) / 2) * factorial((j2 - bb2) / 2) *
                             factorial((j + cc2) / 2) * factorial((j - cc2) / 2) * (j + 1));
  return sum * dcg * sfaccg;
}

} // namespace

struct emd_snap {
  SnapTab h;              // host copy of the tables
  SnapTab *d_tab = nullptr;
  double *d_betaj = nullptr;  // [nelements][ntriples]
  double *d_betaj_e = nullptr; // the same for the energy: beta_k of the bispectrum blocks alone
  double *d_e0 = nullptr;      // [nelements][2]: beta_0, beta_0 - sum_k beta_k bzero_k
  int4 *d_segs = nullptr;     // YiSeg descriptors of snap_yi, strip after strip
  double *d_steptab = nullptr; // Clebsch-Gordan factors of snap_yi's steps: [block][mb2][k], rows padded to an even length
  int ntab = 0, nsegs = 0;
  int ncoeff = 0;
  // work arrays (grow-only)
  Scratch ulist, ylist, cnt, pair_i, pair_j, queue;
  int *d_direct = nullptr;    // snap_deidrj: a pair needs the direct recursion (set by the first launch, read by the second)
  int ucap = 0;           // atoms the U/Y arrays are sized for (= row stride of ulist)
  int npairs = 0;
  int max_smem_optin = 0;
};

namespace {

// ------------------------------------------------------------------------------ device helpers
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

// ------------------------------------------------------------------------------ in-cutoff pairs
// force_snap_neigh_impl.h:612-656: neighbors with rsq < rcutmax^2 (strict), in list order
template <bool FILL>
__global__ void __launch_bounds__(128) snap_pairs_kernel(const double *__restrict__ x, int n_local, emd_neigh_list list, double cutsq,
                                                         int *__restrict__ cnt, int *__restrict__ pair_i, int *__restrict__ pair_j) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  const double x_i = x[3 * (size_t)i], y_i = x[3 * (size_t)i + 1], z_i = x[3 * (size_t)i + 2];
  const int *row;
  int n;
  row_of(list, i, row, n);
  int c = FILL ? cnt[i] : 0;
  for (int jj = 0; jj < n; jj++) {
    const int j = row[jj];
    const double dx = x[3 * (size_t)j] - x_i, dy = x[3 * (size_t)j + 1] - y_i, dz = x[3 * (size_t)j + 2] - z_i;
    const double rsq = dx * dx + dy * dy + dz * dz;
    if (rsq < cutsq) {
      if (FILL) { pair_i[c] = i; pair_j[c] = j; }
      c++;
    }
  }
  if (!FILL) cnt[i] = c;
}

// Cayley-Klein parameters and switching function of one pair (compute_ui :166-181, compute_uarray :650-656,
// compute_sfac :1098-1111)
struct PairGeom {
  double a_r, a_i, b_r, b_i, sfac;
};

__device__ __forceinline__ double sfac_of(const SnapTab &t, double r, double rcut) {
  if (t.switchflag == 0) return 1.0;
  if (r <= t.rmin0) return 1.0;
  if (r > rcut) return 0.0;
  const double rcutfac = kPi / (rcut - t.rmin0);
  return 0.5 * (cos((r - t.rmin0) * rcutfac) + 1.0);
}
__device__ __forceinline__ double dsfac_of(const SnapTab &t, double r, double rcut) {
  if (t.switchflag == 0) return 0.0;
  if (r <= t.rmin0) return 0.0;
  if (r > rcut) return 0.0;
  const double rcutfac = kPi / (rcut - t.rmin0);
  return -0.5 * sin((r - t.rmin0) * rcutfac) * rcutfac;
}

// ------------------------------------------------------------------------------------ snap_ui
// block = 32 atoms (lanes) x ncol warps (warp = column mb).  Dynamic shared memory:
//   acc  [nuh][32] double2   U_tot accumulators of the batch (each warp touches only its column)
// A thread runs the recursion of TWO neighbors at a time (independent dependency chains: the kernel is bound by the
// latency of the FP64 pipe at ~2.5 warps per scheduler, profiles/r02s1_snap_ncu.csv: `wait` 2.7 warps per issue) and
// adds both to an accumulator with one read-modify-write.  The column loop is unrolled, so the inversion-symmetry image
// that starts the next column is a reversal of the register array with static indices (no scratch memory).
struct UiGeom { double a_r, a_i, b_r, b_i, sfac; };
constexpr int kUiGeomIt = 10; // neighbor-pair iterations whose geometry is staged in shared memory at a time

__device__ __forceinline__ UiGeom ui_geom(const SnapTab &t, const double *__restrict__ x, const int *__restrict__ type, int j, bool act,
                                          double x_i, double y_i, double z_i, double rad_i) {
  double dx = x[3 * (size_t)j] - x_i, dy = x[3 * (size_t)j + 1] - y_i, dz = x[3 * (size_t)j + 2] - z_i;
  if (!act) { dx = 1.0; dy = 0.0; dz = 0.0; }
  const int elem_j = t.elem_of_type[type[j]];
  const double rcut = (rad_i + t.radelem[elem_j]) * t.rcutfac;
  const double rsq = dx * dx + dy * dy + dz * dz;
  const double r = sqrt(rsq);
  const double theta0 = (r - t.rmin0) * t.rfac0 * kPi / (rcut - t.rmin0);
  double sn, cs;
  sincos(theta0, &sn, &cs);
  const double z0 = r * cs / sn; // = r / tan(theta0), compute_ui :178
  const double r0inv = 1.0 / sqrt(r * r + z0 * z0);
  UiGeom g;
  g.a_r = r0inv * z0; g.a_i = -r0inv * dz; g.b_r = r0inv * dy; g.b_i = -r0inv * dx;
  g.sfac = act ? sfac_of(t, r, rcut) * t.wjelem[elem_j] : 0.0;
  return g;
}

// inversion symmetry VMK 4.4(2) (:697-717): u(J, J-ma, J-mb) = (-1)^(ma+mb) conj(u(J, ma, mb)), J = 2c+1, mb = c: level J of
// column c, reversed in place, is level J of column c+1
template <int C>
__device__ __forceinline__ void u_image(double2 (&u)[kMaxJ + 1]) {
  constexpr int J = 2 * C + 1;
#pragma unroll
  for (int s = 0; s <= J / 2; s++) {
    const double sg = ((s + C) & 1) ? -1.0 : 1.0, sh = ((J - s + C) & 1) ? -1.0 : 1.0;
    const double2 lo = u[s], hi = u[J - s];
    u[J - s] = make_double2(sg * lo.x, -sg * lo.y);
    u[s] = make_double2(sh * hi.x, -sh * hi.y);
  }
}

// One element of a level of the Wigner-U recursion for column mb (sna_impl.hpp:667-692), evaluated for both neighbors inside one
// guarded block (two independent chains):
//   u_j(ma) = rootpq(j-ma, j-mb) conj(a) u_{j-1}(ma) - rootpq(ma, j-mb) conj(b) u_{j-1}(ma-1)
// in place, walked from ma = j down so that u_{j-1}(ma-1) is still the old value.
__device__ __forceinline__ double2 u_elem(const double2 (&u)[kMaxJ + 1], int ma, int j, double c1, double c2, const UiGeom &g) {
  // c1 = rootpq(0, .) = 0 at the top element (ma == j), where u[ma] still holds a finite value of an earlier level or column
  double nr = c1 * (g.a_r * u[ma].x + g.a_i * u[ma].y);
  double ni = c1 * (g.a_r * u[ma].y - g.a_i * u[ma].x);
  if (ma > 0) {
    nr -= c2 * (g.b_r * u[ma - 1].x + g.b_i * u[ma - 1].y);
    ni -= c2 * (g.b_r * u[ma - 1].y - g.b_i * u[ma - 1].x);
  }
  return make_double2(nr, ni);
}

// column C: one copy of the level code for all columns (five unrolled copies left the kernel waiting for instructions:
// 1.4 warps per issue in no_instruction); only the register reversal that starts the next column needs C at compile time
__device__ __forceinline__ void ui_column(int C, const SnapTab &t, double2 *__restrict__ acc, const double *__restrict__ s_rootpq, int col, int lane,
                                          double2 (&uA)[kMaxJ + 1], double2 (&uB)[kMaxJ + 1], const UiGeom &gA, const UiGeom &gB) {
  const bool own = C == col;
  const int jend = own ? t.twojmax : 2 * C + 1;
  for (int jl = (C == 0 ? 1 : 2 * C); jl <= jend; jl++) {
    double2 *q = acc + (size_t)(t.uh_block[jl] + col * (jl + 1)) * 32 + lane;
    const double *rq = s_rootpq + (jl - C);
    // one element: both neighbors, then (own column) the accumulator; elements are taken downwards so that u_{j-1}(ma-1) is
    // still the old value, two per guarded block (eight independent FP64 chains instead of four)
    auto elem = [&](int ma, double2 &nA, double2 &nB) {
      const double c1 = rq[(jl - ma) * kRootDim], c2 = ma > 0 ? rq[ma * kRootDim] : 0.0;
      nA = u_elem(uA, ma, jl, c1, c2, gA);
      nB = u_elem(uB, ma, jl, c1, c2, gB);
    };
    auto commit = [&](int ma, const double2 &nA, const double2 &nB) {
      uA[ma] = nA;
      uB[ma] = nB;
      if (own) {
        double2 v = q[ma * 32];
        v.x += gA.sfac * nA.x + gB.sfac * nB.x;
        v.y += gA.sfac * nA.y + gB.sfac * nB.y;
        q[ma * 32] = v;
      }
    };
#pragma unroll
    for (int mh = kMaxJ; mh >= 0; mh -= 2) {
      double2 hA, hB, lA, lB;
      if (mh <= jl) {
        elem(mh, hA, hB);
        if (mh > 0) elem(mh - 1, lA, lB);
        commit(mh, hA, hB);
        if (mh > 0) commit(mh - 1, lA, lB);
      } else if (mh > 0 && mh - 1 <= jl) {
        elem(mh - 1, lA, lB);
        commit(mh - 1, lA, lB);
      }
    }
  }
  if (C < col) {
    switch (C) {
      case 0: u_image<0>(uA); u_image<0>(uB); break;
      case 1: u_image<1>(uA); u_image<1>(uB); break;
      case 2: u_image<2>(uA); u_image<2>(uB); break;
      default: u_image<3>(uA); u_image<3>(uB); break;
    }
  }
}

__global__ void __launch_bounds__(32 * kMaxCol) snap_ui_kernel(const SnapTab *__restrict__ tab, const double *__restrict__ x,
                                                               const int *__restrict__ type, int n_local, const int *__restrict__ poff,
                                                               const int *__restrict__ pair_j, double2 *__restrict__ ulist, int ustride) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ double s_rootpq[kRootDim * kRootDim];
  const SnapTab &t = *tab;
  const int twojmax = t.twojmax;
  double2 *acc = reinterpret_cast<double2 *>(dyn);
  const int lane = threadIdx.x & 31, col = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < kRootDim * kRootDim; k += blockDim.x) s_rootpq[k] = t.rootpq[k];
  for (int j = 2 * col; j <= twojmax; j++)
    for (int ma = 0; ma <= j; ma++) acc[(size_t)(t.uh_block[j] + col * (j + 1) + ma) * 32 + lane] = make_double2(0.0, 0.0);
  __syncthreads();

  const int i = blockIdx.x * 32 + lane;
  const bool valid = i < n_local;
  const int ic = valid ? i : 0;
  const double x_i = x[3 * (size_t)ic], y_i = x[3 * (size_t)ic + 1], z_i = x[3 * (size_t)ic + 2];
  const double rad_i = t.radelem[t.elem_of_type[type[ic]]];
  const int pbeg = valid ? poff[i] : 0, pcnt = valid ? poff[i + 1] - pbeg : 0;
  int nmax = pcnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, o));

  // The geometry of a neighbor (square root, sincos, two divisions: ~350 instructions against the 1 600 of the recursion of
  // a neighbor pair) is the same for every column: the warps evaluate it in turns for kUiGeomIt iterations at a time and
  // hand it over through shared memory, instead of every warp evaluating all of it.
  double *geom = reinterpret_cast<double *>(acc + (size_t)t.nuh * 32) + lane; // [kUiGeomIt][2 neighbors][5][32]
  const int nit = (nmax + 1) >> 1, ncolw = blockDim.x >> 5;
  for (int it0 = 0; it0 < nit; it0 += kUiGeomIt) {
    const int it1 = min(nit, it0 + kUiGeomIt);
    if (it0 > 0) __syncthreads(); // the chunk before has been consumed
    for (int it = it0 + col; it < it1; it += ncolw) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = 2 * it + h;
        const bool act = k < pcnt;
        const UiGeom g = ui_geom(t, x, type, act ? pair_j[pbeg + k] : ic, act, x_i, y_i, z_i, rad_i);
        double *gp = geom + ((it - it0) * 2 + h) * 5 * 32;
        gp[0] = g.a_r; gp[32] = g.a_i; gp[64] = g.b_r; gp[96] = g.b_i; gp[128] = g.sfac;
      }
    }
    __syncthreads();
    for (int it = it0; it < it1; it++) {
      UiGeom gA, gB;
      {
        const double *gp = geom + (it - it0) * 2 * 5 * 32;
        gA.a_r = gp[0]; gA.a_i = gp[32]; gA.b_r = gp[64]; gA.b_i = gp[96]; gA.sfac = gp[128];
        gB.a_r = gp[160]; gB.a_i = gp[192]; gB.b_r = gp[224]; gB.b_i = gp[256]; gB.sfac = gp[288];
      }
      double2 uA[kMaxJ + 1], uB[kMaxJ + 1];
#pragma unroll
      for (int ma = 0; ma <= kMaxJ; ma++) uA[ma] = uB[ma] = make_double2(ma == 0 ? 1.0 : 0.0, 0.0); // all finite (u_elem)
      if (col == 0) { // level 0 belongs to column 0 (add_uarraytot :612-635)
        double2 &q = acc[(size_t)t.uh_block[0] * 32 + lane];
        q.x += gA.sfac + gB.sfac;
      }
      // column c: levels 2c, 2c+1 re-derived by every warp of a later column (first level = image of column c-1), all
      // remaining levels by its own warp
      for (int c = 0; c <= col; c++) ui_column(c, t, acc, s_rootpq, col, lane, uA, uB, gA, gB);
    }
  }
  // self term (addself_uarraytot :594-605) and write-out of this warp's column
  if (valid) {
    for (int j = 2 * col; j <= twojmax; j++)
      for (int ma = 0; ma <= j; ma++) {
        const int e = t.uh_block[j] + col * (j + 1) + ma;
        double2 v = acc[(size_t)e * 32 + lane];
        if (ma == col) v.x += t.wself;
        ulist[(size_t)e * ustride + i] = v;
      }
  }
}

// ------------------------------------------------------------------------------------ snap_yi
// block = 32 atoms (lanes) x W warps.  Dynamic shared memory: [front pad][sU [nuf][32] double2 = full U_tot of the batch][back
// pad][step table of Clebsch-Gordan factors].
//
// Register tiling of compute_zi's inner loop (sna_impl.hpp:248-262).  A warp takes an output STRIP (j, ma) = all
// NMB = j/2+1 values of mb at once.  For one row pair (ma1, ma2) of a block (j1,j2,j), j1 >= j2, the products are
//     y(mb) += [beta cg(ma1,ma2)] cg(mb1, mb2) u_j1(ma1, mb1) u_j2(ma2, mb2),   mb1 = mb + C - mb2,  C = (j1+j2-j)/2.
// The SHORT row (j2) is the one that is stepped through: mb2 = 0 .. min(j2, C+NMB-1); the NMB elements of the long row
// that the strip needs at one step form a WINDOW that slides down by one element per step (a ring of NMB registers, the
// step loop unrolled NMB-fold so that every slot index is a compile-time constant).  Per step: one new element of each
// row (2 LDS.128), the NMB coefficients of the step as one aligned row of the step table (ceil(NMB/2) broadcast LDS.128;
// a coefficient of an (mb1, mb2) outside the block is stored as 0, so the window is filled without range tests: what it
// reads outside its row is some other finite element of sU or the zeroed pad), 2 + 6 NMB FP64 instructions.  The factor
// beta*cg(ma1,ma2) of the row pair is folded into the stepped element, so the strip accumulates straight into its NMB
// outputs.  Everything that depends only on (strip, block) -- row offsets, strides, trip counts, table offset -- comes from
// a segment descriptor built at emd_snap_create (32 bytes, warp-uniform, fetched one segment ahead).
//
// History at 250 000 atoms (profiles/): plain loop nest 14.5 ms (shared-memory bound); fully unrolled per-block code (125
// instantiations, 380 KB of SASS) 13.5-24 ms (instruction-fetch bound); round 1's sliding window over the long row with
// the coefficients as indexed constant loads 10.7 ms: FP64 pipe 42 % busy, every term waited for its own LDC (one
// register pair for c at the 128-register cap) and a third of the executed terms were zero padding of the union range.
// Stepping the short row executes 45 656 instead of 51 476 terms per atom (38 358 are non-zero).
constexpr int kYiWarps = 16;
constexpr int kNumCg = 4098; // see snap_triples.inc
constexpr int kCgPad = 8;
__constant__ double c_cg[kNumCg + 2 * kCgPad]; // compact Clebsch-Gordan blocks for twojmax = 8 (a 2J = 6 run uses a subset)
constexpr int kYiFrontPad = 16, kYiBackPad = 8; // elements ([32] double2 each) before / after sU

struct __align__(16) YiSeg {
  int a_off;       // long row: element index (in sU) of the window's slot 0 at step 0, (j1, ma1lo, mb1 = C)
  int w_off;       // short row: element index of (j2, ma2 of ma1lo, mb2 = 0)
  int tab_off;     // step table: index (in doubles) of the block's row of step 0
  int cga_idx;     // c_cg index of cg(ma1lo, ma2)
  short a_stride;  // j1 + 1: next ma1
  short w_stride;  // -(j2 + 1): ma2 decreases as ma1 increases
  short cga_stride; // j2
  short tab_stride; // doubles per step row (NMB of the block's j rounded up to even)
  short nrows, nsteps, tr, pad;
};

// One segment: ROWS = 2 advances the output rows ma and ma+1 together (same ma1, hence the same window of the long row and
// the same step factors; the stepped elements are those of ma2 and ma2+1); the product c*a is shared, so a term costs
// (2 + 4 ROWS) / ROWS FP64 instructions.  Every operand of the next step is requested right after the last use of the
// registers it lands in (the window slot of output NMB-1 first), so the loads of a step fly during the step before it
// without a second set of registers.
template <int NMB, int ROWS>
__device__ __forceinline__ void z_segment(const double2 *__restrict__ U, const double *__restrict__ s_tab, const YiSeg &g, double bj,
                                          double (&y0r)[kMaxCol], double (&y0i)[kMaxCol], double (&y1r)[kMaxCol], double (&y1i)[kMaxCol]) {
  constexpr int NC2 = (NMB + 1) / 2;
  const double2 *arow = U + g.a_off * 32;
  const double2 *wrow = U + g.w_off * 32;
  const int row1 = -g.w_stride * 32; // the second output row reads the short row ma2 + 1
  int cga_idx = g.cga_idx;
  for (int row = 0; row < g.nrows; row++) {
    double sc0 = bj * c_cg[cga_idx], sc1 = ROWS == 2 ? bj * c_cg[cga_idx + 1] : 0.0;
    asm volatile("" : "+d"(sc0), "+d"(sc1)); // keep them in registers (the compiler would re-load and re-multiply them in every step)
    double2 w[NMB], c2[NC2];
#pragma unroll
    for (int k = 0; k < NMB; k++) w[k] = arow[k * 32];
    const double2 *an = arow - 32;   // element entering the window after the current step
    const double2 *wp = wrow;
    const double2 *tp = reinterpret_cast<const double2 *>(s_tab + g.tab_off);
    double2 r0 = wp[0], r1 = ROWS == 2 ? wp[row1] : make_double2(0.0, 0.0);
#pragma unroll
    for (int q = 0; q < NC2; q++) c2[q] = tp[q];
#pragma unroll 1
    for (int st = 0; st < g.nsteps; st++) { // not unrolled: the instruction cache decides this kernel (see the header)
      const double2 e0 = make_double2(sc0 * r0.x, sc0 * r0.y);
      double2 e1 = make_double2(0.0, 0.0);
      if (ROWS == 2) e1 = make_double2(sc1 * r1.x, sc1 * r1.y);
      wp += 32;
      r0 = wp[0];
      if (ROWS == 2) r1 = wp[row1];
      const double2 a_new = *an; // what output 0 needs at the next step
      an -= 32;
      tp = reinterpret_cast<const double2 *>(reinterpret_cast<const double *>(tp) + g.tab_stride);
      int done = 0;
#pragma unroll
      for (int kk = 0; kk < NMB; kk++) {
        const int k = NMB - 1 - kk; // downwards: the window moves up behind the terms
        const double2 a = w[k];
        const double c = (k & 1) ? c2[k >> 1].y : c2[k >> 1].x;
        const double cax = c * a.x, cay = c * a.y;
        y0r[k] = fma(e0.x, cax, y0r[k]); y0r[k] = fma(-e0.y, cay, y0r[k]);
        y0i[k] = fma(e0.x, cay, y0i[k]); y0i[k] = fma(e0.y, cax, y0i[k]);
        if (ROWS == 2) {
          y1r[k] = fma(e1.x, cax, y1r[k]); y1r[k] = fma(-e1.y, cay, y1r[k]);
          y1i[k] = fma(e1.x, cay, y1i[k]); y1i[k] = fma(e1.y, cax, y1i[k]);
        }
        w[k] = k > 0 ? w[k - 1] : a_new; // output k's element of the next step is output k-1's of this one
        done |= 1 << k;
        const int q = k >> 1, mate = k ^ 1;
        if (mate >= NMB || (done >> mate & 1)) c2[q] = tp[q]; // both factors of the pair are used: fetch the next step's
      }
    }
    arow += g.a_stride * 32;
    wrow += g.w_stride * 32;
    cga_idx += g.cga_stride;
  }
}

// storage form of a descriptor (16 bytes, kept in shared memory next to the step table)
__host__ __device__ inline int4 pack_seg(const YiSeg &g) {
  int4 p;
  p.x = g.a_off | (g.w_off << 16);
  p.y = g.tab_off | (g.cga_idx << 16);
  p.z = g.a_stride | (-g.w_stride << 4) | (g.cga_stride << 8) | (g.tab_stride << 12) | (g.nrows << 16) | (g.nsteps << 20) | (g.tr << 24);
  p.w = 0;
  return p;
}
__device__ __forceinline__ YiSeg load_seg(const int4 *__restrict__ segs, int q) {
  const int4 p = segs[q];
  YiSeg g;
  g.a_off = p.x & 0xffff; g.w_off = (unsigned)p.x >> 16;
  g.tab_off = p.y & 0xffff; g.cga_idx = (unsigned)p.y >> 16;
  g.a_stride = (short)(p.z & 15); g.w_stride = (short)-((p.z >> 4) & 15); g.cga_stride = (short)((p.z >> 8) & 15);
  g.tab_stride = (short)((p.z >> 12) & 15); g.nrows = (short)((p.z >> 16) & 15); g.nsteps = (short)((p.z >> 20) & 15);
  g.tr = (short)((unsigned)p.z >> 24); g.pad = 0;
  return g;
}

template <int NMB>
__device__ __forceinline__ void yi_item(const double2 *__restrict__ U, const double *__restrict__ s_tab, const int4 *__restrict__ segs,
                                        const double *__restrict__ beta_i, const int (&seg)[4], double (&y0r)[kMaxCol],
                                        double (&y0i)[kMaxCol], double (&y1r)[kMaxCol], double (&y1i)[kMaxCol]) {
  // the descriptor of the next segment travels while the current one runs (the list ends with a spare descriptor)
  YiSeg g = load_seg(segs, seg[0]);
  for (int q = seg[0]; q < seg[1]; q++) {
    const YiSeg gn = load_seg(segs, q + 1);
    z_segment<NMB, 2>(U, s_tab, g, beta_i[g.tr], y0r, y0i, y1r, y1i);
    g = gn;
  }
  for (int pass = 0; pass < 2; pass++) { // segments that reach only one of the two rows: one copy of the code, summed into a scratch row
    if (seg[1 + pass] == seg[2 + pass]) continue;
    double tr[kMaxCol], ti[kMaxCol];
#pragma unroll
    for (int k = 0; k < kMaxCol; k++) { tr[k] = 0.0; ti[k] = 0.0; }
    g = load_seg(segs, seg[1 + pass]);
    for (int q = seg[1 + pass]; q < seg[2 + pass]; q++) {
      const YiSeg gn = load_seg(segs, q + 1);
      z_segment<NMB, 1>(U, s_tab, g, beta_i[g.tr], tr, ti, tr, ti);
      g = gn;
    }
#pragma unroll
    for (int k = 0; k < NMB; k++) {
      if (pass == 0) { y0r[k] += tr[k]; y0i[k] += ti[k]; }
      else { y1r[k] += tr[k]; y1i[k] += ti[k]; }
    }
  }
}

__global__ void __launch_bounds__(32 * kYiWarps) snap_yi_kernel(const SnapTab *__restrict__ tab, const double *__restrict__ betaj,
                                                                const int4 *__restrict__ segs, int nsegs, const double *__restrict__ steptab,
                                                                int ntab, const int *__restrict__ type, int n_local, const double2 *__restrict__ ulist,
                                                                int ustride, double2 *__restrict__ ylist, size_t yhalf) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ int s_next;
  const SnapTab &t = *tab;
  double2 *sU = reinterpret_cast<double2 *>(dyn) + kYiFrontPad * 32;
  double *s_tab = reinterpret_cast<double *>(sU + ((size_t)t.nuf + kYiBackPad) * 32);
  int4 *s_segs = reinterpret_cast<int4 *>(s_tab + ntab);       // ntab is even
  double *s_beta = reinterpret_cast<double *>(s_segs + nsegs); // [nelements][kMaxTriples]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool valid = i < n_local;
  if (threadIdx.x == 0) s_next = 0;
  for (int k = threadIdx.x; k < kYiFrontPad * 32; k += blockDim.x) sU[k - kYiFrontPad * 32] = make_double2(0.0, 0.0);
  for (int k = threadIdx.x; k < kYiBackPad * 32; k += blockDim.x) sU[(size_t)t.nuf * 32 + k] = make_double2(0.0, 0.0);
  for (int k = threadIdx.x; k < ntab; k += blockDim.x) s_tab[k] = steptab[k];
  for (int k = threadIdx.x; k < nsegs; k += blockDim.x) s_segs[k] = segs[k];
  for (int k = threadIdx.x; k < t.nelements * kMaxTriples; k += blockDim.x) s_beta[k] = betaj[k];
  // expand the half range to the full (ma,mb) range with the inversion symmetry: a warp requests all of its elements
  // (at most kMaxHalf / kYiWarps + 1) before it stores the first one
  {
    constexpr int kPer = (kMaxHalf + kYiWarps - 1) / kYiWarps;
    double2 v[kPer];
#pragma unroll
    for (int n = 0; n < kPer; n++) {
      const int e = warp + n * nwarps;
      v[n] = (valid && e < t.nuh) ? ulist[(size_t)e * ustride + i] : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int n = 0; n < kPer; n++) {
      const int e = warp + n * nwarps;
      if (e < t.nuh) {
        sU[(size_t)t.exp_dst[e] * 32 + lane] = v[n];
        const int img = t.exp_img[e];
        if (img != 0) {
          const double sg = img < 0 ? -1.0 : 1.0;
          sU[(size_t)(abs(img) - 1) * 32 + lane] = make_double2(sg * v[n].x, -sg * v[n].y);
        }
      }
    }
  }
  __syncthreads();
  const int elem_i = valid ? t.elem_of_type[type[i]] : 0;
  const double *beta_i = s_beta + elem_i * kMaxTriples;
  const double2 *U = sU + lane;

  for (;;) { // work items from a queue sorted by cost, most expensive first
    int s = 0;
    if (lane == 0) s = atomicAdd(&s_next, 1);
    s = __shfl_sync(0xffffffffu, s, 0);
    if (s >= t.nitems) break;
    const int J = t.item[s].j, ma = t.item[s].ma, rows = t.item[s].rows;
    const int seg[4] = {t.item[s].seg[0], t.item[s].seg[1], t.item[s].seg[2], t.item[s].seg[3]};
    double y0r[kMaxCol], y0i[kMaxCol], y1r[kMaxCol], y1i[kMaxCol];
#pragma unroll
    for (int mb = 0; mb < kMaxCol; mb++) { y0r[mb] = 0.0; y0i[mb] = 0.0; y1r[mb] = 0.0; y1i[mb] = 0.0; }
    switch (t.item[s].nmb) {
      case 1: yi_item<1>(U, s_tab, s_segs, beta_i, seg, y0r, y0i, y1r, y1i); break;
      case 2: yi_item<2>(U, s_tab, s_segs, beta_i, seg, y0r, y0i, y1r, y1i); break;
      case 3: yi_item<3>(U, s_tab, s_segs, beta_i, seg, y0r, y0i, y1r, y1i); break;
      case 4: yi_item<4>(U, s_tab, s_segs, beta_i, seg, y0r, y0i, y1r, y1i); break;
      default: yi_item<5>(U, s_tab, s_segs, beta_i, seg, y0r, y0i, y1r, y1i); break;
    }
    if (valid) {
      double2 *Y = ylist + t.item[s].half * yhalf + (size_t)i * t.nuh + t.uh_block[J] + ma;
#pragma unroll
      for (int mb = 0; mb < kMaxCol; mb++)
        if (2 * mb <= J) {
          // half-column weights of compute_dbidrj (:393-424): 1, and on the middle column of even J: 1 above the
          // diagonal, 1/2 on it, 0 below (those outputs are not computed: nmb stops short of them)
          const double w0 = (2 * mb < J) ? 1.0 : (ma < mb ? 1.0 : (ma == mb ? 0.5 : 0.0));
          Y[mb * (J + 1)] = make_double2(w0 * y0r[mb], w0 * y0i[mb]);
          if (rows == 2) {
            const double w1 = (2 * mb < J) ? 1.0 : (ma + 1 < mb ? 1.0 : (ma + 1 == mb ? 0.5 : 0.0));
            Y[mb * (J + 1) + 1] = make_double2(w1 * y1r[mb], w1 * y1i[mb]);
          }
        }
    }
  }
}

// -------------------------------------------------------------------------------- snap_deidrj
constexpr int kDeThreads = 128;
constexpr int kDeStageAtoms = 12; // Y rows staged per CTA (155 double2 each at 2J = 8)

// one level of the dU recursion for column mb, in place, all three directions (compute_duarray :786-838)
__device__ __forceinline__ void du_level(double2 (&u)[kMaxJ + 1], double2 (&du)[3][kMaxJ + 1], int j, int mb, const double *__restrict__ s_rootpq,
                                         double a_r, double a_i, double b_r, double b_i, const double (&da_r)[3], const double (&da_i)[3],
                                         const double (&db_r)[3], const double (&db_i)[3]) {
#pragma unroll
  for (int ma = kMaxJ; ma >= 0; --ma) {
    if (ma <= j) {
      const double c1 = s_rootpq[(j - ma) * kRootDim + (j - mb)]; // rootpq(0, .) = 0 at the top element (see du_level_unit)
      double c2 = 0.0;
      const double2 uo = u[ma];
      double2 um = make_double2(0.0, 0.0);
      if (ma > 0) { c2 = s_rootpq[ma * kRootDim + (j - mb)]; um = u[ma - 1]; }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double2 dq = du[k][ma];
        double2 dm = make_double2(0.0, 0.0);
        if (ma > 0) dm = du[k][ma - 1];
        const double t1r = da_r[k] * uo.x + da_i[k] * uo.y + a_r * dq.x + a_i * dq.y;
        const double t1i = da_r[k] * uo.y - da_i[k] * uo.x + a_r * dq.y - a_i * dq.x;
        const double t2r = db_r[k] * um.x + db_i[k] * um.y + b_r * dm.x + b_i * dm.y;
        const double t2i = db_r[k] * um.y - db_i[k] * um.x + b_r * dm.y - b_i * dm.x;
        du[k][ma] = make_double2(c1 * t1r - c2 * t2r, c1 * t1i - c2 * t2i);
      }
      const double t1r = a_r * uo.x + a_i * uo.y, t1i = a_r * uo.y - a_i * uo.x;
      const double t2r = b_r * um.x + b_i * um.y, t2i = b_r * um.y - b_i * um.x;
      u[ma] = make_double2(c1 * t1r - c2 * t2r, c1 * t1i - c2 * t2i);
    }
  }
}

// The same level for the three UNIT tangents d/da_i, d/db_r, d/db_i (derivatives with respect to the Cayley-Klein
// parameters themselves): the inhomogeneous term of each is a single element of u, so a tangent costs 12 FP64
// instructions per element instead of the 20 of a direction x, y, z.  The fourth parameter derivative follows from
// homogeneity (every u_j is a homogeneous polynomial of degree j in (a_r, a_i, b_r, b_i): the recursion multiplies by
// conj(a) or conj(b) once per level, the inversion image conjugates), i.e. Euler:
//     a_r d/da_r u_j = j u_j - a_i d/da_i u_j - b_r d/db_r u_j - b_i d/db_i u_j,
// and the directional derivative is the chain rule over the four parameters (snap_deidrj_kernel).
__device__ __forceinline__ void du_level_unit(double2 (&u)[kMaxJ + 1], double2 (&d)[3][kMaxJ + 1], int j, int mb, const double *__restrict__ s_rootpq,
                                              double a_r, double a_i, double b_r, double b_i, const double2 *__restrict__ Yl, double &L,
                                              double (&T)[3]) {
#pragma unroll
  for (int ma = kMaxJ; ma >= 0; --ma) {
    if (ma <= j) {
      // the top element (ma == j) has no first term: its factor rootpq(0, .) is 0 and u[j], d[.][j] still hold finite values
      // (zero, or what an earlier column left there), so nothing is selected at run time
      const double c1 = s_rootpq[(j - ma) * kRootDim + (j - mb)];
      double c2 = 0.0;
      const double2 uo = u[ma];
      double2 um = make_double2(0.0, 0.0);
      if (ma > 0) { c2 = s_rootpq[ma * kRootDim + (j - mb)]; um = u[ma - 1]; }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double2 dq = d[k][ma];
        double2 dm = make_double2(0.0, 0.0);
        if (ma > 0) dm = d[k][ma - 1];
        // conj(da) uo with da = i (k = 0); conj(db) um with db = 1 (k = 1), db = i (k = 2)
        const double i1r = k == 0 ? uo.y : 0.0, i1i = k == 0 ? -uo.x : 0.0;
        const double i2r = k == 1 ? um.x : (k == 2 ? um.y : 0.0), i2i = k == 1 ? um.y : (k == 2 ? -um.x : 0.0);
        const double t1r = a_r * dq.x + (a_i * dq.y + i1r), t1i = a_r * dq.y - (a_i * dq.x - i1i);
        const double t2r = b_r * dm.x + (b_i * dm.y + i2r), t2i = b_r * dm.y - (b_i * dm.x - i2i);
        d[k][ma] = make_double2(c1 * t1r - c2 * t2r, c1 * t1i - c2 * t2i);
      }
      const double t1r = a_r * uo.x + a_i * uo.y, t1i = a_r * uo.y - a_i * uo.x;
      const double t2r = b_r * um.x + b_i * um.y, t2i = b_r * um.y - b_i * um.x;
      u[ma] = make_double2(c1 * t1r - c2 * t2r, c1 * t1i - c2 * t2i);
      // contraction with Y inside the same guarded block (the Y element travels during the update)
      const double2 y = Yl[ma];
      L += u[ma].x * y.x + u[ma].y * y.y;
#pragma unroll
      for (int k = 0; k < 3; k++) T[k] += d[k][ma].x * y.x + d[k][ma].y * y.y;
    }
  }
}

// S0 = sum w Re(conj(u) Y), Sj = the same weighted with the level j, T[k] = sum w Re(conj(d_k u) Y) over the half columns of one
// pair; d_k = the unit tangents d/da_i, d/db_r, d/db_i (UNIT) or the directions x, y, z themselves.
template <bool UNIT>
__device__ __forceinline__ void de_sums(const SnapTab &t, const double2 *__restrict__ Y, double2 *__restrict__ boot, int bs,
                                        const double *__restrict__ s_rootpq, double a_r, double a_i, double b_r, double b_i,
                                        const double (&da_r)[3], const double (&da_i)[3], const double (&db_r)[3], const double (&db_i)[3],
                                        double &S0, double &Sj, double (&T)[3]) {
  const int twojmax = t.twojmax, ncol = t.ncol;
  Sj = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) T[k] = 0.0;
  double2 u[kMaxJ + 1], du[3][kMaxJ + 1];
#pragma unroll
  for (int ma = 0; ma <= kMaxJ; ma++) { // every element finite from the start (du_level_unit)
    u[ma] = make_double2(ma == 0 ? 1.0 : 0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 3; k++) du[k][ma] = make_double2(0.0, 0.0);
  }
  S0 = Y[t.uh_block[0]].x; // level 0: u = 1, du = 0

  for (int c = 0; c < ncol; c++) {
    if (c > 0) {
#pragma unroll
      for (int ma = 0; ma <= kMaxJ; ma++)
        if (ma <= 2 * c - 1) {
          u[ma] = boot[(ma * 4 + 0) * bs];
#pragma unroll
          for (int k = 0; k < 3; k++) du[k][ma] = boot[(ma * 4 + 1 + k) * bs];
        }
    }
    for (int jl = max(1, 2 * c); jl <= twojmax; jl++) {
      const double2 *Yl = Y + t.uh_block[jl] + c * (jl + 1);
      double L = 0.0;
      if (UNIT) du_level_unit(u, du, jl, c, s_rootpq, a_r, a_i, b_r, b_i, Yl, L, T);
      else {
        du_level(u, du, jl, c, s_rootpq, a_r, a_i, b_r, b_i, da_r, da_i, db_r, db_i);
#pragma unroll
        for (int ma = 0; ma <= kMaxJ; ma++)
          if (ma <= jl) {
            const double2 y = Yl[ma];
            L += u[ma].x * y.x + u[ma].y * y.y;
#pragma unroll
            for (int k = 0; k < 3; k++) T[k] += du[k][ma].x * y.x + du[k][ma].y * y.y;
          }
      }
      S0 += L;
      if (UNIT) Sj += jl * L;
      if (jl == 2 * c + 1 && c + 1 < ncol) { // image that starts column c+1 (:840-864)
#pragma unroll
        for (int s = 0; s <= kMaxJ; s++)
          if (s <= jl) {
            const double sg = ((s + c) & 1) ? -1.0 : 1.0;
            boot[((jl - s) * 4 + 0) * bs] = make_double2(sg * u[s].x, -sg * u[s].y);
#pragma unroll
            for (int k = 0; k < 3; k++) boot[((jl - s) * 4 + 1 + k) * bs] = make_double2(sg * du[k][s].x, -sg * du[k][s].y);
          }
      }
    }
  }
}

// compute_duidrj :290-321, compute_duarray :740-771: Cayley-Klein parameters of a pair and their derivatives
struct DeGeom {
  double a_r, a_i, b_r, b_i, da_r[3], da_i[3], db_r[3], db_i[3], uhat[3], sfac, dsfac, rsq;
};
__device__ __forceinline__ DeGeom de_geom(const SnapTab &t, double dx, double dy, double dz, double rcut, double wj) {
  DeGeom g;
  const double rsq = dx * dx + dy * dy + dz * dz;
  const double r = sqrt(rsq);
  const double rscale0 = t.rfac0 * kPi / (rcut - t.rmin0);
  const double theta0 = (r - t.rmin0) * rscale0;
  double sn, cs;
  sincos(theta0, &sn, &cs);
  const double z0 = r * cs / sn;
  const double dz0dr = z0 / r - (r * rscale0) * (rsq + z0 * z0) / rsq;
  const double rinv = 1.0 / r;
  g.uhat[0] = dx * rinv; g.uhat[1] = dy * rinv; g.uhat[2] = dz * rinv;
  const double r0inv = 1.0 / sqrt(r * r + z0 * z0);
  g.a_r = z0 * r0inv; g.a_i = -dz * r0inv; g.b_r = dy * r0inv; g.b_i = -dx * r0inv;
  const double dr0invdr = -(r0inv * r0inv * r0inv) * (r + z0 * dz0dr);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double dr0inv = dr0invdr * g.uhat[k], dz0 = dz0dr * g.uhat[k];
    g.da_r[k] = dz0 * r0inv + z0 * dr0inv;
    g.da_i[k] = -dz * dr0inv;
    g.db_r[k] = dy * dr0inv;
    g.db_i[k] = -dx * dr0inv;
  }
  g.da_i[2] += -r0inv;
  g.db_i[0] += -r0inv;
  g.db_r[1] += r0inv;
  g.sfac = sfac_of(t, r, rcut) * wj; g.dsfac = dsfac_of(t, r, rcut) * wj;
  g.rsq = rsq;
  return g;
}

// lanes = in-cutoff pairs.  Dynamic shared memory: boot [kMaxJ][4][blockDim] double2.
// UNIT = true: the launch that does the work (pairs with |a_r| >= unit_min_ar); it raises *direct_flag if it met a pair
// below that bound.  UNIT = false: the second launch, which returns at once unless the flag is up and then handles
// exactly those pairs with the three directions in the recursion (two kernels, because one kernel holding both
// recursions spills its hot loop: 5.3 -> 7.1 ms).
template <bool UNIT>
__global__ void __launch_bounds__(kDeThreads) snap_deidrj_kernel(const SnapTab *__restrict__ tab, const double *__restrict__ x,
                                                                 const int *__restrict__ type, const int *__restrict__ pair_i,
                                                                 const int *__restrict__ pair_j, int npairs,
                                                                 const double2 *__restrict__ ylist, size_t yhalf, double *__restrict__ f,
                                                                 int *__restrict__ direct_flag) {
  extern __shared__ __align__(16) unsigned char dyn[];
  __shared__ double s_rootpq[kRootDim * kRootDim];
  if (!UNIT && *direct_flag == 0) return;
  const SnapTab &t = *tab;
  for (int k = threadIdx.x; k < kRootDim * kRootDim; k += blockDim.x) s_rootpq[k] = t.rootpq[k];
  // Pairs are sorted by their central atom, so the CTA's 128 pairs belong to a short run of consecutive atoms (7 at 18
  // in-cutoff neighbors): their Y rows (atom-major, 2.4 KB each) are staged in shared memory with one contiguous copy,
  // so the contraction reads Y at LDS latency.  (Reading Y from global at its point of use left the kernel
  // latency-bound at 8 warps/SM: 2.5 long-scoreboard stalls per issue, 42 % FP64 pipe.)  Y arrives as the two partial
  // sums of snap_yi's item halves and is added up on the way in.  A run longer than the staging area (very short rows)
  // is worked off in rounds.
  double2 *s_y = reinterpret_cast<double2 *>(dyn) + (size_t)kMaxJ * 4 * blockDim.x;
  const int p0 = blockIdx.x * blockDim.x;
  const int i_first = pair_i[p0], i_last = pair_i[min(p0 + (int)blockDim.x, npairs) - 1];
  const bool active = p0 + (int)threadIdx.x < npairs;
  const int p = active ? p0 + threadIdx.x : npairs - 1;
  double2 *boot = reinterpret_cast<double2 *>(dyn) + threadIdx.x;
  const int bs = blockDim.x;
  const int i = pair_i[p], j = pair_j[p];
  const double dx = x[3 * (size_t)j] - x[3 * (size_t)i], dy = x[3 * (size_t)j + 1] - x[3 * (size_t)i + 1],
               dz = x[3 * (size_t)j + 2] - x[3 * (size_t)i + 2];
  const int elem_i = t.elem_of_type[type[i]], elem_j = t.elem_of_type[type[j]];
  const double rcut = (t.radelem[elem_i] + t.radelem[elem_j]) * t.rcutfac;
  const double wj = t.wjelem[elem_j];
  for (int base = 0; base <= i_last - i_first; base += kDeStageAtoms) {
    if (base > 0) __syncthreads();
    {
      const int n = min(i_last - i_first + 1 - base, kDeStageAtoms) * t.nuh;
      const double2 *src = ylist + (size_t)(i_first + base) * t.nuh;
      for (int k0 = threadIdx.x; k0 < n; k0 += 4 * blockDim.x) { // eight loads in flight per thread (the copy was 8 % of the kernel's stall samples)
        double2 y0[4], y1[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = min(k0 + q * (int)blockDim.x, n - 1);
          y0[q] = src[k]; y1[q] = src[yhalf + k];
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int k = k0 + q * blockDim.x;
          if (k < n) s_y[k] = make_double2(y0[q].x + y1[q].x, y0[q].y + y1[q].y);
        }
      }
    }
    __syncthreads();
    if (!active || i - i_first < base || i - i_first >= base + kDeStageAtoms) continue;
    const double2 *Y = s_y + (size_t)(i - i_first - base) * t.nuh;
    double S0, Sj, T[3], S[3];
    double dxg = dx, dyg = dy, dzg = dz;
    DeGeom g = de_geom(t, dxg, dyg, dzg, rcut, wj);
    // |a_r| = |z0| / r0 passes through 0 where theta0 = pi/2 (r ~ half the cutoff: closer than any neighbor of the decks'
    // lattices, but legal); such pairs are left to the UNIT = false launch
    const bool unit_ok = fabs(g.a_r) >= t.unit_min_ar;
    if (UNIT && !unit_ok) atomicOr(direct_flag, 1);
    if (unit_ok != UNIT) continue;
    if (UNIT) {
      // The unit-tangent recursion needs only a and b: the rest of the geometry is evaluated again after it (a few hundred
      // instructions against 9 000) instead of being kept in 40 registers.
      de_sums<true>(t, Y, boot, bs, s_rootpq, g.a_r, g.a_i, g.b_r, g.b_i, g.da_r, g.da_i, g.db_r, g.db_i, S0, Sj, T);
      asm volatile("" : "+d"(dxg), "+d"(dyg), "+d"(dzg));
      g = de_geom(t, dxg, dyg, dzg, rcut, wj);
      const double T_ar = (Sj - g.a_i * T[0] - g.b_r * T[1] - g.b_i * T[2]) / g.a_r; // Euler (du_level_unit)
#pragma unroll
      for (int k = 0; k < 3; k++) S[k] = g.da_r[k] * T_ar + g.da_i[k] * T[0] + g.db_r[k] * T[1] + g.db_i[k] * T[2];
    } else {
      de_sums<false>(t, Y, boot, bs, s_rootpq, g.a_r, g.a_i, g.b_r, g.b_i, g.da_r, g.da_i, g.db_r, g.db_i, S0, Sj, T);
#pragma unroll
      for (int k = 0; k < 3; k++) S[k] = T[k];
    }
    // dU_full = dsfac u uhat + sfac dU (:873-892); F_ij = 2 sum w Re(conj(dU_full) Y) + rij * (-1.5e6 / r^14) (force_snap_neigh_impl.h:698-711)
    const double rsq7 = (g.rsq * g.rsq * g.rsq) * (g.rsq * g.rsq * g.rsq) * g.rsq;
    const double fdivr = -1.5e6 / rsq7;
    const double rij[3] = {dx, dy, dz};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const double fk = 2.0 * (g.dsfac * g.uhat[k] * S0 + g.sfac * S[k]) + rij[k] * fdivr;
      atomicAdd(&f[3 * (size_t)i + k], fk);
      atomicAdd(&f[3 * (size_t)j + k], -fk);
    }
  }
}

// ------------------------------------------------------------------------------- snap_energy
// E_i = e0(elem) + 2 sum_{half range} Re(conj(U_tot) Y_E) + sum_{j inside the cutoff} 1.25e5 / r_ij^12, with Y_E = the adjoint Y
// for the coefficients beta_k of the blocks (j1,j2,j) of the bispectrum list alone (no fold factors): the first sum is
// sum_k beta_k B_k of LAMMPS' SNA::compute_bi (B_k = 2 sum w Re(conj(U) Z_k), the same half-column weights), e0 = beta_0
// - sum_k beta_k bzero_k; the last term is the potential whose gradient the force kernel adds as rij * (-1.5e6 / r^14) from
// both ends of a pair (force_snap_neigh_impl.h:698-711).  Not in the reference (ForceSNAP inherits Force::compute_energy = 0).
__global__ void __launch_bounds__(256) snap_energy_kernel(const SnapTab *__restrict__ tab, const double *__restrict__ x, const int *__restrict__ type,
                                                          int n_local, const int *__restrict__ poff, const int *__restrict__ pair_j,
                                                          const double2 *__restrict__ ulist, int ustride, const double2 *__restrict__ ylist,
                                                          size_t yhalf, const double *__restrict__ e0, int which, double *__restrict__ partial) {
  __shared__ double sm[8];
  const SnapTab &t = *tab;
  double e = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += gridDim.x * blockDim.x) {
    double b = 0.0;
    const double2 *Y = ylist + (size_t)i * t.nuh;
    for (int k = 0; k < t.nuh; k++) {
      const double2 u = ulist[(size_t)k * ustride + i], y0 = Y[k], y1 = Y[yhalf + k];
      b += u.x * (y0.x + y1.x) + u.y * (y0.y + y1.y);
    }
    double rep = 0.0;
    const double x_i = x[3 * (size_t)i], y_i = x[3 * (size_t)i + 1], z_i = x[3 * (size_t)i + 2];
    for (int p = poff[i]; p < poff[i + 1]; p++) {
      const int j = pair_j[p];
      const double dx = x[3 * (size_t)j] - x_i, dy = x[3 * (size_t)j + 1] - y_i, dz = x[3 * (size_t)j + 2] - z_i;
      const double rsq = dx * dx + dy * dy + dz * dz, r6 = rsq * rsq * rsq;
      rep += 1.25e5 / (r6 * r6);
    }
    e += e0[2 * t.elem_of_type[type[i]] + which] + 2.0 * b + rep;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = e;
  __syncthreads();
  if (warp == 0) {
    double v = lane < 8 ? sm[lane] : 0.0;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) partial[blockIdx.x] = v;
  }
}

// ------------------------------------------------------------------ host: index tables and the snap_yi plan (no CUDA calls)
struct T3 { int j1, j2, j; };
const struct { int idx, j1, j2, j, cgoff; } kTri[] = {
#define X(IDX, TJ1, TJ2, TJ, CGOFF) {IDX, TJ1, TJ2, TJ, CGOFF},
#include "snap_triples.inc"
#undef X
};
static_assert(sizeof kTri / sizeof kTri[0] == kMaxTriples, "snap_triples.inc must list the 125 blocks of twojmax = 8");

// layouts of U (half / full range), the expansion table, the blocks (j1,j2,j) of twojmax = 8 and their Clebsch-Gordan factors
int build_index_tables(SnapTab &h, int J2, std::vector<T3> &full, std::vector<double> &cg) {
  int nuh = 0, nuf = 0;
  for (int j = 0; j <= J2; j++) { h.uh_block[j] = nuh; nuh += (j / 2 + 1) * (j + 1); h.uf_block[j] = nuf; nuf += (j + 1) * (j + 1); }
  h.nuh = nuh; h.nuf = nuf;
  for (int j = 0; j <= J2; j++)
    for (int mb = 0; 2 * mb <= j; mb++)
      for (int ma = 0; ma <= j; ma++) {
        const int e = h.uh_block[j] + mb * (j + 1) + ma;
        h.exp_dst[e] = (short)(h.uf_block[j] + ma * (j + 1) + mb);
        const int img = h.uf_block[j] + (j - ma) * (j + 1) + (j - mb) + 1;
        h.exp_img[e] = (short)(2 * mb == j ? 0 : (((ma + mb) & 1) ? -img : img));
      }
  h.ntriples = kMaxTriples;
  cg.assign(kNumCg, 0.0);
  full.clear();
  for (int tI = 0; tI < kMaxTriples; tI++) {
    const int j1 = kTri[tI].j1, j2 = kTri[tI].j2, j = kTri[tI].j;
    full.push_back({j1, j2, j});
    h.triple[tI].j1 = (short)j1; h.triple[tI].j2 = (short)j2; h.triple[tI].j = (short)j; h.triple[tI].cgoff = kTri[tI].cgoff;
    if (kTri[tI].idx != tI || kTri[tI].cgoff + (j1 + 1) * (j2 + 1) > kNumCg || (tI > 0 && full[tI - 1].j > j)) {
      set_error("emd_snap_create: snap_triples.inc is inconsistent"); return 1;
    }
    for (int m1 = 0; m1 <= j1; m1++)
      for (int m2 = 0; m2 <= j2; m2++) cg[kTri[tI].cgoff + m1 * (j2 + 1) + m2] = clebsch_gordan(j1, j2, j, m1, m2);
  }
  h.ncg = kNumCg;
  for (int j = 0, tI = 0; j <= kMaxJ + 1; j++) {
    while (tI < h.ntriples && full[tI].j < j) tI++;
    h.tri_begin[j] = tI;
  }
  return 0;
}

// the work items, segments and step table of snap_yi (see the kernel's header)
int build_yi_plan(SnapTab &h, int J2, const std::vector<T3> &full, const std::vector<double> &cg, std::vector<double> &steptab,
                  std::vector<YiSeg> &segs) {
  // snap_yi's step table: for block tI, step mb2 and output k the factor cg(mb1 = C + k - mb2, mb2), 0 outside the block
  steptab.clear();
  std::vector<int> tab_off(kMaxTriples, 0), tab_stride(kMaxTriples, 0);
  for (int tI = 0; tI < kMaxTriples; tI++) {
    const int j1 = full[tI].j1, j2 = full[tI].j2, j = full[tI].j, C = (j1 + j2 - j) / 2;
    const int nk = (j / 2 + 1 + 1) / 2 * 2;
    tab_off[tI] = (int)steptab.size(); tab_stride[tI] = nk;
    for (int mb2 = 0; mb2 <= j2; mb2++)
      for (int k = 0; k < nk; k++) {
        const int mb1 = C + k - mb2;
        steptab.push_back((k <= j / 2 && mb1 >= 0 && mb1 <= j1) ? cg[kTri[tI].cgoff + mb1 * (j2 + 1) + mb2] : 0.0);
      }
  }
  for (int k = 0; k < 8; k++) steptab.push_back(0.0); // a step row is fetched whole
  // snap_yi work items.  The rows ma = 0..j of level j are taken in pairs of equal output count (the rows below the diagonal
  // of an even level's middle column have one output less); per block of the level, the ma1 range that reaches both rows of a
  // pair becomes a two-row segment, what reaches only one of them a one-row segment.  Every group is split in two halves of
  // equal cost by blocks, so that 16 warps find ~50 items of similar size; the halves write two Y arrays that snap_deidrj adds.
  auto nmb_of = [](int j, int ma) { return j / 2 + 1 - ((j % 2 == 0 && ma > j / 2) ? 1 : 0); };
  auto make_seg = [&](int tI, int j, int ma, int nmb, int lo, int hi) {
    const int j1 = full[tI].j1, j2 = full[tI].j2, C = (j1 + j2 - j) / 2, ma2 = ma + C - lo;
    YiSeg g;
    g.a_off = h.uf_block[j1] + lo * (j1 + 1) + C;
    g.w_off = h.uf_block[j2] + ma2 * (j2 + 1);
    g.tab_off = tab_off[tI];
    g.cga_idx = kCgPad + kTri[tI].cgoff + lo * (j2 + 1) + ma2;
    g.a_stride = (short)(j1 + 1); g.w_stride = (short)-(j2 + 1); g.cga_stride = (short)j2; g.tab_stride = (short)tab_stride[tI];
    g.nrows = (short)(hi - lo + 1); g.nsteps = (short)(imin(j2, C + nmb - 1) + 1); g.tr = (short)tI; g.pad = 0;
    return g;
  };
  struct Item { int j, ma, rows, nmb, half; long cost; std::vector<YiSeg> seg[3]; };
  std::vector<Item> items;
  for (int j = 0; j <= J2; j++)
    for (int ma = 0; ma <= j;) {
      const int nmb = nmb_of(j, ma);
      const int rows = (ma + 1 <= j && nmb_of(j, ma + 1) == nmb) ? 2 : 1;
      if (nmb > 0) {
        struct Blk { long cost; YiSeg seg[3]; bool has[3]; };
        std::vector<Blk> blks;
        for (int tI = h.tri_begin[j]; tI < h.tri_begin[j + 1]; tI++) {
          const int j1 = full[tI].j1, j2 = full[tI].j2, C = (j1 + j2 - j) / 2;
          if (j1 > J2) continue;
          // 0 <= ma2 = ma + C - ma1 <= j2
          const int lo0 = imax(0, ma + C - j2), hi0 = imin(j1, ma + C);
          const int lo1 = rows == 2 ? imax(0, ma + 1 + C - j2) : 1, hi1 = rows == 2 ? imin(j1, ma + 1 + C) : 0;
          const int plo = imax(lo0, lo1), phi = imin(hi0, hi1);
          Blk bk{0, {}, {false, false, false}};
          const int nsteps = imin(j2, C + nmb - 1) + 1;
          if (plo <= phi) {
            bk.seg[0] = make_seg(tI, j, ma, nmb, plo, phi); bk.has[0] = true;
            bk.cost += (long)(phi - plo + 1) * (nsteps * (10 * nmb + 4) + 16) + 24;
            // what is left of either range is one ma1 at its outer end
            if (lo0 < plo) { bk.seg[1] = make_seg(tI, j, ma, nmb, lo0, plo - 1); bk.has[1] = true; }
            if (hi0 > phi) { set_error("emd_snap_create: internal range error"); return 1; }
            if (hi1 > phi) { bk.seg[2] = make_seg(tI, j, ma + 1, nmb, phi + 1, hi1); bk.has[2] = true; }
            if (lo1 < plo) { set_error("emd_snap_create: internal range error"); return 1; }
          } else {
            if (lo0 <= hi0) { bk.seg[1] = make_seg(tI, j, ma, nmb, lo0, hi0); bk.has[1] = true; }
            if (lo1 <= hi1) { bk.seg[2] = make_seg(tI, j, ma + 1, nmb, lo1, hi1); bk.has[2] = true; }
          }
          for (int k = 1; k < 3; k++)
            if (bk.has[k]) bk.cost += (long)bk.seg[k].nrows * (nsteps * (6 * nmb + 2) + 16) + 24;
          if (bk.has[0] || bk.has[1] || bk.has[2]) blks.push_back(bk);
        }
        std::stable_sort(blks.begin(), blks.end(), [](const Blk &x, const Blk &y) { return x.cost > y.cost; });
        Item half[2];
        for (int k = 0; k < 2; k++) half[k] = Item{j, ma, rows, nmb, k, 0, {}};
        for (const Blk &bk : blks) {
          Item &it = half[half[0].cost <= half[1].cost ? 0 : 1];
          it.cost += bk.cost;
          for (int k = 0; k < 3; k++)
            if (bk.has[k]) it.seg[k].push_back(bk.seg[k]);
        }
        // both halves exist even when one is empty: every Y element of both arrays is written
        items.push_back(half[0]); items.push_back(half[1]);
      }
      ma += rows;
    }
  std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) { return x.cost > y.cost; });
  if ((int)items.size() > kMaxItems) { set_error("emd_snap_create: too many snap_yi work items"); return 1; }
  h.nitems = (int)items.size();
  segs.clear();
  for (int k = 0; k < h.nitems; k++) {
    const Item &it = items[k];
    h.item[k].j = (short)it.j; h.item[k].ma = (short)it.ma; h.item[k].rows = (short)it.rows; h.item[k].nmb = (short)it.nmb;
    h.item[k].half = (short)it.half; h.item[k].pad = 0;
    for (int q = 0; q < 3; q++) {
      h.item[k].seg[q] = (int)segs.size();
      segs.insert(segs.end(), it.seg[q].begin(), it.seg[q].end());
    }
    h.item[k].seg[3] = (int)segs.size();
  }
  segs.push_back(YiSeg{}); // the kernel fetches one descriptor ahead
  return 0;
}

size_t ui_smem(const SnapTab &h) { return (size_t)h.nuh * 32 * sizeof(double2) + (size_t)kUiGeomIt * 2 * 5 * 32 * sizeof(double); }
size_t yi_smem(const SnapTab &h, int ntab, int nsegs) {
  return ((size_t)h.nuf + kYiFrontPad + kYiBackPad) * 32 * sizeof(double2) + sizeof(double) * (size_t)ntab + sizeof(int4) * (size_t)nsegs +
         sizeof(double) * (size_t)h.nelements * kMaxTriples;
}
size_t de_smem(const SnapTab &h) { return ((size_t)kMaxJ * 4 * kDeThreads + (size_t)kDeStageAtoms * h.nuh) * sizeof(double2); }

} // namespace

extern "C" {

// Host-only self-check of the snap_yi plan for a given twojmax (no device needed; tests/test_snap_plan.py): every (block, output row,
// ma1) of compute_zi (sna_impl.hpp:196-283) that has a non-empty mb range must be covered by exactly one segment row.  Returns the
// plan's sizes, the Clebsch-Gordan terms it executes and the non-zero ones among them.
int emd_snap_yi_plan_stats(int twojmax, int *nitems, int *nsegs, int *ntab, long long *terms_executed, long long *terms_nonzero) {
  if (twojmax < 0 || twojmax > kMaxJ) { set_error("emd_snap_yi_plan_stats: twojmax %d not in [0,%d]", twojmax, kMaxJ); return 1; }
  SnapTab *hp = new SnapTab();
  SnapTab &h = *hp;
  memset(&h, 0, sizeof h);
  h.twojmax = twojmax; h.ncol = twojmax / 2 + 1;
  std::vector<T3> full;
  std::vector<double> cg, steptab;
  std::vector<YiSeg> segs;
  if (build_index_tables(h, twojmax, full, cg) || build_yi_plan(h, twojmax, full, cg, steptab, segs)) { delete hp; return 1; }
  std::vector<int> cover((size_t)kMaxTriples * (kMaxJ + 1) * (kMaxJ + 1), 0);
  long long exec = 0, nonzero = 0;
  for (int it = 0; it < h.nitems; it++) {
    const int j = h.item[it].j, nmb = h.item[it].nmb;
    for (int q = 0; q < 3; q++)
      for (int sg = h.item[it].seg[q]; sg < h.item[it].seg[q + 1]; sg++) {
        const YiSeg &g = segs[sg];
        const int j1 = full[g.tr].j1, j2 = full[g.tr].j2, C = (j1 + j2 - j) / 2;
        if (full[g.tr].j != j) { set_error("emd_snap_yi_plan_stats: segment of another level"); delete hp; return 1; }
        const int lo = (g.a_off - h.uf_block[j1] - C) / (j1 + 1), ma2 = (g.w_off - h.uf_block[j2]) / (j2 + 1);
        const int rows = q == 0 ? 2 : 1;
        for (int r = 0; r < g.nrows; r++)
          for (int o = 0; o < rows; o++) {
            const int ma1 = lo + r, ma = ma2 - r + o - C + ma1; // ma2 of this row = ma2 - r (+ o for the second output row)
            const int want = h.item[it].ma + (q == 2 ? 1 : o);
            if (ma != want || ma1 < 0 || ma1 > j1) { set_error("emd_snap_yi_plan_stats: segment row does not belong to its item"); delete hp; return 1; }
            cover[((size_t)g.tr * (kMaxJ + 1) + ma) * (kMaxJ + 1) + ma1]++;
            exec += (long long)g.nsteps * nmb;
            for (int st = 0; st < g.nsteps; st++)
              for (int k = 0; k < nmb; k++) nonzero += steptab[g.tab_off + st * g.tab_stride + k] != 0.0;
          }
      }
  }
  for (int tI = 0; tI < kMaxTriples; tI++) {
    const int j1 = full[tI].j1, j2 = full[tI].j2, j = full[tI].j, C = (j1 + j2 - j) / 2;
    for (int ma = 0; ma <= j; ma++)
      for (int ma1 = 0; ma1 <= kMaxJ; ma1++) {
        const int nmb = j / 2 + 1 - ((j % 2 == 0 && ma > j / 2) ? 1 : 0);
        const bool needed = j1 <= twojmax && j <= twojmax && nmb > 0 && ma1 <= j1 && ma + C - ma1 >= 0 && ma + C - ma1 <= j2;
        if (cover[((size_t)tI * (kMaxJ + 1) + ma) * (kMaxJ + 1) + ma1] != (needed ? 1 : 0)) {
          set_error("emd_snap_yi_plan_stats: block %d row %d ma1 %d covered %d times", tI, ma, ma1, cover[((size_t)tI * (kMaxJ + 1) + ma) * (kMaxJ + 1) + ma1]);
          delete hp; return 1;
        }
      }
  }
  if (nitems) *nitems = h.nitems;
  if (nsegs) *nsegs = (int)segs.size();
  if (ntab) *ntab = (int)steptab.size();
  if (terms_executed) *terms_executed = exec;
  if (terms_nonzero) *terms_nonzero = nonzero;
  delete hp;
  return 0;
}

int emd_snap_create(emd_snap **out, const emd_snap_params *p) {
  if (!out || !p) { set_error("emd_snap_create: NULL argument"); return 1; }
  if (p->twojmax < 0 || p->twojmax > kMaxJ) { set_error("emd_snap_create: twojmax %d not in [0,%d]", p->twojmax, kMaxJ); return 1; }
  if (p->ntypes < 1 || p->ntypes > kMaxTypesConst || p->nelements < 1 || p->nelements > kMaxTypesConst) {
    set_error("emd_snap_create: ntypes/nelements out of range"); return 1;
  }
  emd_snap *s = new emd_snap();
  SnapTab &h = s->h;
  memset(&h, 0, sizeof h);
  const int J2 = p->twojmax;
  h.twojmax = J2; h.ncol = J2 / 2 + 1; h.ntypes = p->ntypes; h.nelements = p->nelements; h.switchflag = p->switchflag;
  h.rcutfac = p->rcutfac; h.rfac0 = p->rfac0; h.rmin0 = p->rmin0; h.wself = p->wself;
  { const char *e = getenv("EMD_SNAP_DEIDRJ_DIRECT"); h.unit_min_ar = (e && atoi(e)) ? 2.0 : 0.25; }
  std::vector<T3> full;
  std::vector<double> cg;
  if (build_index_tables(h, J2, full, cg)) { delete s; return 1; }
  for (int pp = 1; pp <= J2; pp++) // init_rootpqarray, sna_impl.hpp:1053-1061
    for (int q = 1; q <= J2; q++) h.rootpq[pp * kRootDim + q] = sqrt(static_cast<double>(pp) / q);
  double rcutmax = 0.0; // force_snap_neigh_impl.h:321-329
  for (int e = 0; e < p->nelements; e++) {
    h.radelem[e] = p->radelem[e]; h.wjelem[e] = p->wjelem[e];
    rcutmax = std::max(2.0 * p->radelem[e] * p->rcutfac, rcutmax);
  }
  h.cutsq = rcutmax * rcutmax;
  for (int ty = 0; ty < p->ntypes; ty++) h.elem_of_type[ty] = p->elem_of_type[ty];

  // index lists (build_indexlist, sna_impl.hpp:86-132, diagonalstyle 3): idxj = the ncoeff bispectrum components of THIS
  // twojmax; the Z blocks (idxj_full) are always the twojmax = 8 list of snap_triples.inc (sorted by j), of which a
  // smaller twojmax uses the blocks with j1 <= twojmax
  std::vector<T3> idxj;
  for (int j1 = 0; j1 <= J2; j1++)
    for (int j2 = 0; j2 <= j1; j2++)
      for (int j = abs(j1 - j2); j <= imin(J2, j1 + j2); j += 2)
        if (j >= j1) idxj.push_back({j1, j2, j});
  s->ncoeff = (int)idxj.size();
  if (s->ncoeff != p->ncoeffall - 1) { // :315-318
    set_error("emd_snap_create: coefficient count %d does not match twojmax %d (expected %d + 1)", p->ncoeffall, J2, s->ncoeff);
    delete s; return 1;
  }
  auto tri_index = [&](int a, int b, int c) {
    for (int tI = 0; tI < h.ntriples; tI++)
      if (full[tI].j1 == a && full[tI].j2 == b && full[tI].j == c) return tI;
    return -1;
  };
  // betaj: the coefficient of every Z block in Y (fold of compute_dbidrj's three sums, :393-527, with beta)
  std::vector<double> betaj((size_t)p->nelements * h.ntriples, 0.0), betaj_e((size_t)p->nelements * h.ntriples, 0.0), e0(2 * (size_t)p->nelements, 0.0);
  for (int e = 0; e < p->nelements; e++) {
    const double *coeff = p->coeffelem + (size_t)e * p->ncoeffall;
    double *bj = betaj.data() + (size_t)e * h.ntriples;
    for (int JJ = 0; JJ < s->ncoeff; JJ++) {
      const int j1 = idxj[JJ].j1, j2 = idxj[JJ].j2, j = idxj[JJ].j;
      const double b = coeff[JJ + 1];
      const int t1 = tri_index(imax(j1, j2), imin(j1, j2), j);
      const int t2 = tri_index(imax(j, j2), imin(j, j2), j1);
      const int t3 = tri_index(imax(j1, j), imin(j1, j), j2);
      if (t1 < 0 || t2 < 0 || t3 < 0) { set_error("emd_snap_create: internal index error"); delete s; return 1; }
      bj[t1] += b;
      bj[t2] += b * ((j + 1) / (j1 + 1.0));
      bj[t3] += b * ((j + 1) / (j2 + 1.0));
      betaj_e[(size_t)e * h.ntriples + t1] += b;
      e0[2 * e + 1] -= b * (p->wself * p->wself * p->wself) * (j + 1); // bzero[j] of SNA::init (LAMMPS sna.cpp)
    }
    e0[2 * e] = coeff[0];
    e0[2 * e + 1] += coeff[0];
  }
  std::vector<double> steptab;
  std::vector<YiSeg> segs;
  if (build_yi_plan(h, J2, full, cg, steptab, segs)) { delete s; return 1; }
  if (steptab.size() % 2) steptab.push_back(0.0);
  s->ntab = (int)steptab.size();
  if (steptab.size() > 0xffff) { set_error("emd_snap_create: step table too long for the packed descriptors"); delete s; return 1; }
  std::vector<int4> packed;
  for (const YiSeg &g : segs) packed.push_back(pack_seg(g));
  s->nsegs = (int)packed.size();
  EMD_CUDA(cudaMalloc((void **)&s->d_segs, sizeof(int4) * packed.size()));
  EMD_CUDA(cudaMemcpy(s->d_segs, packed.data(), sizeof(int4) * packed.size(), cudaMemcpyHostToDevice));
  EMD_CUDA(cudaMalloc((void **)&s->d_steptab, sizeof(double) * steptab.size()));
  EMD_CUDA(cudaMemcpy(s->d_steptab, steptab.data(), sizeof(double) * steptab.size(), cudaMemcpyHostToDevice));

  EMD_CUDA(cudaMalloc((void **)&s->d_tab, sizeof(SnapTab)));
  EMD_CUDA(cudaMemcpy(s->d_tab, &h, sizeof(SnapTab), cudaMemcpyHostToDevice));
  {
    std::vector<double> padded(kNumCg + 2 * kCgPad, 0.0);
    std::copy(cg.begin(), cg.end(), padded.begin() + kCgPad);
    EMD_CUDA(cudaMemcpyToSymbol(c_cg, padded.data(), sizeof(double) * padded.size())); // the same twojmax = 8 table for every emd_snap
  }
  EMD_CUDA(cudaMalloc((void **)&s->d_betaj, sizeof(double) * betaj.size()));
  EMD_CUDA(cudaMemcpy(s->d_betaj, betaj.data(), sizeof(double) * betaj.size(), cudaMemcpyHostToDevice));
  EMD_CUDA(cudaMalloc((void **)&s->d_betaj_e, sizeof(double) * betaj_e.size()));
  EMD_CUDA(cudaMemcpy(s->d_betaj_e, betaj_e.data(), sizeof(double) * betaj_e.size(), cudaMemcpyHostToDevice));
  EMD_CUDA(cudaMalloc((void **)&s->d_e0, sizeof(double) * e0.size()));
  EMD_CUDA(cudaMemcpy(s->d_e0, e0.data(), sizeof(double) * e0.size(), cudaMemcpyHostToDevice));
  int dev = 0;
  EMD_CUDA(cudaGetDevice(&dev));
  EMD_CUDA(cudaDeviceGetAttribute(&s->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (yi_smem(h, s->ntab, s->nsegs) > (size_t)s->max_smem_optin) { set_error("emd_snap_create: U_tot batch does not fit in shared memory"); delete s; return 1; }
  EMD_CUDA(cudaFuncSetAttribute(snap_ui_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ui_smem(h)));
  EMD_CUDA(cudaFuncSetAttribute(snap_yi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)yi_smem(h, s->ntab, s->nsegs)));
  EMD_CUDA(cudaFuncSetAttribute(snap_deidrj_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)de_smem(h)));
  EMD_CUDA(cudaFuncSetAttribute(snap_deidrj_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)de_smem(h)));
  EMD_CUDA(cudaMalloc((void **)&s->d_direct, sizeof(int)));
  *out = s;
  return 0;
}

void emd_snap_destroy(emd_snap *s) {
  if (!s) return;
  if (s->d_tab) cudaFree(s->d_tab);
  if (s->d_betaj) cudaFree(s->d_betaj);
  if (s->d_betaj_e) cudaFree(s->d_betaj_e);
  if (s->d_e0) cudaFree(s->d_e0);
  if (s->d_segs) cudaFree(s->d_segs);
  if (s->d_steptab) cudaFree(s->d_steptab);
  if (s->d_direct) cudaFree(s->d_direct);
  s->ulist.release(); s->ylist.release(); s->cnt.release(); s->pair_i.release(); s->pair_j.release(); s->queue.release();
  delete s;
}

int emd_snap_info(const emd_snap *s, int *ncoeff, int *nuh, int *ntriples, double *rcutmax, int *npairs, int *ustride) {
  if (!s) return 1;
  if (ncoeff) *ncoeff = s->ncoeff;
  if (nuh) *nuh = s->h.nuh;
  if (ntriples) *ntriples = s->h.ntriples;
  if (rcutmax) *rcutmax = sqrt(s->h.cutsq);
  if (npairs) *npairs = s->npairs;
  if (ustride) *ustride = s->ucap;
  return 0;
}

void *emd_snap_device_ptr(emd_snap *s, const char *what) {
  if (!s || !what) return nullptr;
  if (!strcmp(what, "ulist")) return s->ulist.p;
  if (!strcmp(what, "ylist")) return s->ylist.p;
  if (!strcmp(what, "pair_i")) return s->pair_i.p;
  if (!strcmp(what, "pair_j")) return s->pair_j.p;
  if (!strcmp(what, "pair_offsets")) return s->cnt.p;
  return nullptr;
}

// in-cutoff pair list + U_tot of every owned atom (the part ForceSNAP::compute and the energy share)
static int snap_pairs_and_ui(emd_ctx *ctx, emd_snap *s, const double *d_x, const int *d_type, int n_local, const emd_neigh_list *list) {
  const SnapTab &h = s->h;
  // in-cutoff pairs: count -> exclusive scan -> fill
  if (s->cnt.ensure(sizeof(int) * ((size_t)n_local + 1))) return 1;
  int *cnt = s->cnt.as<int>();
  EMD_CUDA(cudaMemsetAsync(cnt + n_local, 0, sizeof(int), ctx->stream));
  EMD_LAUNCH(ctx, (snap_pairs_kernel<false>), grid_for(n_local, 128), 128, 0, d_x, n_local, *list, h.cutsq, cnt, nullptr, nullptr);
  if (exclusive_scan_int(ctx, cnt, cnt, n_local + 1, nullptr)) return 1;
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, cnt + n_local, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream)); // like the reference's max_neighs read-back (:181)
  const int npairs = ctx->h_pinned[0];
  if (npairs < 0) { set_error("emd_force_snap_compute: pair count overflows 32 bits"); return 2; }
  s->npairs = npairs;
  if (s->pair_i.ensure(sizeof(int) * ((size_t)npairs + 1)) || s->pair_j.ensure(sizeof(int) * ((size_t)npairs + 1))) return 1;
  if (n_local > s->ucap) {
    const int want = (n_local + n_local / 8 + 31) / 32 * 32;
    if (s->ulist.ensure(sizeof(double2) * (size_t)h.nuh * want) || s->ylist.ensure(sizeof(double2) * (size_t)h.nuh * want * 2)) { s->ucap = 0; return 1; }
    s->ucap = want;
  }
  EMD_LAUNCH(ctx, (snap_pairs_kernel<true>), grid_for(n_local, 128), 128, 0, d_x, n_local, *list, h.cutsq, cnt, s->pair_i.as<int>(), s->pair_j.as<int>());
  EMD_LAUNCH(ctx, snap_ui_kernel, grid_for(n_local, 32), 32 * h.ncol, ui_smem(h), s->d_tab, d_x, d_type, n_local, cnt, s->pair_j.as<int>(),
             s->ulist.as<double2>(), s->ucap);
  return 0;
}

static int snap_yi(emd_ctx *ctx, emd_snap *s, const double *d_betaj, const int *d_type, int n_local) {
  const SnapTab &h = s->h;
  EMD_LAUNCH(ctx, snap_yi_kernel, grid_for(n_local, 32), 32 * kYiWarps, yi_smem(h, s->ntab, s->nsegs), s->d_tab, d_betaj, s->d_segs, s->nsegs, s->d_steptab,
             s->ntab, d_type, n_local, s->ulist.as<double2>(), s->ucap, s->ylist.as<double2>(), (size_t)h.nuh * s->ucap);
  return 0;
}

int emd_force_snap_compute(emd_ctx *ctx, emd_snap *s, const double *d_x, const int *d_type, double *d_f, int n_local, int n_all,
                           const emd_neigh_list *list) {
  if (!ctx || !s || !list) { set_error("emd_force_snap_compute: NULL argument"); return 1; }
  (void)n_all;
  if (n_local <= 0) return 0;
  const SnapTab &h = s->h;
  if (int rc = snap_pairs_and_ui(ctx, s, d_x, d_type, n_local, list)) return rc;
  if (int rc = snap_yi(ctx, s, s->d_betaj, d_type, n_local)) return rc;
  const int npairs = s->npairs;
  int *pair_i = s->pair_i.as<int>(), *pair_j = s->pair_j.as<int>();
  double2 *ylist = s->ylist.as<double2>();
  if (npairs > 0) {
    EMD_CUDA(cudaMemsetAsync(s->d_direct, 0, sizeof(int), ctx->stream));
    EMD_LAUNCH(ctx, snap_deidrj_kernel<true>, grid_for(npairs, kDeThreads), kDeThreads, de_smem(h), s->d_tab, d_x, d_type, pair_i, pair_j, npairs,
               ylist, (size_t)h.nuh * s->ucap, d_f, s->d_direct);
    EMD_LAUNCH(ctx, snap_deidrj_kernel<false>, grid_for(npairs, kDeThreads), kDeThreads, de_smem(h), s->d_tab, d_x, d_type, pair_i, pair_j, npairs,
               ylist, (size_t)h.nuh * s->ucap, d_f, s->d_direct);
  }
  return 0;
}

int emd_force_snap_energy(emd_ctx *ctx, emd_snap *s, const double *d_x, const int *d_type, int n_local, const emd_neigh_list *list,
                          int bzeroflag, double *h_energy) {
  if (!ctx || !s || !list || !h_energy) { set_error("emd_force_snap_energy: NULL argument"); return 1; }
  *h_energy = 0.0;
  if (n_local <= 0) return 0;
  const SnapTab &h = s->h;
  if (int rc = snap_pairs_and_ui(ctx, s, d_x, d_type, n_local, list)) return rc;
  if (int rc = snap_yi(ctx, s, s->d_betaj_e, d_type, n_local)) return rc; // overwrites the Y of the last force call
  const int grid = max(1, min(grid_for(n_local, 256), ctx->num_sms * 8));
  if (ctx->s_c.ensure(sizeof(double) * ((size_t)grid + 8))) return 1;
  double *partial = ctx->s_c.as<double>() + 8;
  EMD_LAUNCH(ctx, snap_energy_kernel, grid, 256, 0, s->d_tab, d_x, d_type, n_local, s->cnt.as<int>(), s->pair_j.as<int>(), s->ulist.as<double2>(),
             s->ucap, s->ylist.as<double2>(), (size_t)h.nuh * s->ucap, s->d_e0, bzeroflag ? 1 : 0, partial);
  return device_sum_partials(ctx, partial, grid, h_energy);
}

} // extern "C"
