// lattice.cu -- Input::create_lattice / create_velocities on the device (src/input.cpp:460-792, src/input.h:66-133).
// The host loops of the reference (three passes over the lattice + a five-draw RNG warm-up per atom) dominate the
// start-up of a 16 M-atom brick; everything here is bit-identical to them:
//   * sites are enumerated in the reference's loop order (z, y, x, basis) and kept when they lie in the rank's brick:
//     flag -> exclusive scan -> scatter = the same atom order;
//   * the velocity stream of an atom is a Jenkins one-at-a-time hash of (seed, position bytes) feeding Park-Miller:
//     integer arithmetic, exactly reproducible; v = (u - 0.5) / sqrt(m) uses IEEE division and square root;
//   * the momentum and temperature sums fix every velocity through the centre-of-mass shift and the rescale factor, so
//     they are accumulated IN ATOM ORDER by one thread (the other threads of the block stage the operands through
//     shared memory): ~10 cycles per atom, 0.1 s for 16 M atoms, instead of a parallel reduction with another rounding.
#include "common.cuh"

using namespace emd;

namespace {

struct LatticeArgs {
  long long ix0, iy0, iz0;
  int nx, ny, nz, nbasis;      // candidate ranges (inclusive ranges of the reference, as extents)
  double a;
  double basis[4][3];          // lattice basis + offset (fcc), or the offset alone in basis[0] (sc)
  int fcc;
  double lo[3], hi[3];         // the rank's brick
};

__device__ __forceinline__ bool site_of(const LatticeArgs &L, long long c, double p[3]) {
  const int k = (int)(c % L.nbasis);
  long long r = c / L.nbasis;
  const long long ix = L.ix0 + r % L.nx; r /= L.nx;
  const long long iy = L.iy0 + r % L.ny; r /= L.ny;
  const long long iz = L.iz0 + r;
  if (L.fcc) { // input.cpp:625-627
    p[0] = __dmul_rn(L.a, __dadd_rn(1.0 * ix, L.basis[k][0]));
    p[1] = __dmul_rn(L.a, __dadd_rn(1.0 * iy, L.basis[k][1]));
    p[2] = __dmul_rn(L.a, __dadd_rn(1.0 * iz, L.basis[k][2]));
  } else {     // input.cpp:506-508
    p[0] = __dmul_rn(L.a, __dadd_rn((double)ix, L.basis[0][0]));
    p[1] = __dmul_rn(L.a, __dadd_rn((double)iy, L.basis[0][1]));
    p[2] = __dmul_rn(L.a, __dadd_rn((double)iz, L.basis[0][2]));
  }
  return p[0] >= L.lo[0] && p[1] >= L.lo[1] && p[2] >= L.lo[2] && p[0] < L.hi[0] && p[1] < L.hi[1] && p[2] < L.hi[2];
}

__global__ void __launch_bounds__(256) lattice_flag_kernel(LatticeArgs L, long long ncand, int *__restrict__ flag) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  double p[3];
  flag[c] = site_of(L, c, p) ? 1 : 0;
}

// LAMMPS_RandomVelocityGeom (src/input.h:66-133): Park-Miller with Schrage's trick; Jenkins one-at-a-time over the bytes
// of (seed, x, y, z) read as plain (signed) char; 27-bit mask; five warm-up draws
__device__ __forceinline__ double pm_uniform(int &seed) {
  const int IA = 16807, IM = 2147483647, IQ = 127773, IR = 2836;
  const int k = seed / IQ;
  seed = IA * (seed - k * IQ) - IR * k;
  if (seed < 0) seed += IM;
  return __dmul_rn(1.0 / IM, (double)seed); // (no FMA contraction with the caller's - 0.5: the host rounds twice)
}
__device__ __forceinline__ void oaat(unsigned &hash, unsigned long long bits, int nbytes) {
  for (int i = 0; i < nbytes; i++) {
    hash += (unsigned)(int)(signed char)((bits >> (8 * i)) & 0xffu);
    hash += (hash << 10);
    hash ^= (hash >> 6);
  }
}

__global__ void __launch_bounds__(256) lattice_fill_kernel(LatticeArgs L, long long ncand, const int *__restrict__ offs, int seed_base, int id_offset,
                                                           const double *__restrict__ mass, double *__restrict__ x, double *__restrict__ v,
                                                           double *__restrict__ q, int *__restrict__ type, int *__restrict__ id) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  double p[3];
  if (!site_of(L, c, p)) return;
  const size_t k = (size_t)offs[c];
  x[3 * k] = p[0]; x[3 * k + 1] = p[1]; x[3 * k + 2] = p[2];
  type[k] = 0;                 // rand() % ntypes with one type (more types: host path, the libc stream is sequential)
  id[k] = (int)k + 1 + id_offset;
  q[k] = 0.0;
  unsigned hash = 0u;
  oaat(hash, (unsigned long long)(unsigned)seed_base, 4);
  for (int d = 0; d < 3; d++) oaat(hash, (unsigned long long)__double_as_longlong(p[d]), 8);
  hash += (hash << 3);
  hash ^= (hash >> 11);
  hash += (hash << 15);
  int seed = (int)(hash & 0x7ffffffu);
  if (!seed) seed = 1;
  for (int i = 0; i < 5; i++) pm_uniform(seed);
  const double sm = __dsqrt_rn(mass[0]);
  const double vx = __dadd_rn(pm_uniform(seed), -0.5), vy = __dadd_rn(pm_uniform(seed), -0.5), vz = __dadd_rn(pm_uniform(seed), -0.5); // input.cpp:744-746
  v[3 * k] = __ddiv_rn(vx, sm); v[3 * k + 1] = __ddiv_rn(vy, sm); v[3 * k + 2] = __ddiv_rn(vz, sm);
}

// sums in atom order: MODE 0: out = {sum m, sum m vx, sum m vy, sum m vz} (input.cpp:750-753); MODE 1: out[0] = sum m |v|^2
// (property_temperature.cpp:49 with one thread)
constexpr int kSeqTile = 512;
template <int MODE>
__global__ void __launch_bounds__(256) seq_sum_kernel(const double *__restrict__ v, const int *__restrict__ type, const double *__restrict__ mass,
                                                      long long n, double *__restrict__ out) {
  __shared__ double s_v[2][3 * kSeqTile];
  __shared__ double s_m[2][kSeqTile];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  auto stage = [&](int b, long long base) {
    const long long m = min((long long)kSeqTile, n - base);
    for (long long e = threadIdx.x; e < 3 * m; e += blockDim.x) s_v[b][e] = v[3 * base + e];
    for (long long e = threadIdx.x; e < m; e += blockDim.x) s_m[b][e] = mass[type[base + e]];
  };
  if (n > 0) stage(0, 0);
  __syncthreads();
  int b = 0;
  for (long long base = 0; base < n; base += kSeqTile, b ^= 1) {
    const long long m = min((long long)kSeqTile, n - base);
    if (threadIdx.x == 0) {
      for (int e = 0; e < (int)m; e++) {
        const double mi = s_m[b][e], vx = s_v[b][3 * e], vy = s_v[b][3 * e + 1], vz = s_v[b][3 * e + 2];
        if (MODE == 0) {
          a0 = __dadd_rn(a0, mi);
          a1 = __dadd_rn(a1, __dmul_rn(mi, vx)); a2 = __dadd_rn(a2, __dmul_rn(mi, vy)); a3 = __dadd_rn(a3, __dmul_rn(mi, vz));
        } else {
          a0 = __dadd_rn(a0, __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)), mi));
        }
      }
    } else if (base + kSeqTile < n) {
      // the other threads fetch the next tile meanwhile (thread 0's share of it is picked up by the strided loops of the rest)
      const long long nb = base + kSeqTile, mm = min((long long)kSeqTile, n - nb);
      for (long long e = threadIdx.x - 1; e < 3 * mm; e += blockDim.x - 1) s_v[b ^ 1][e] = v[3 * nb + e];
      for (long long e = threadIdx.x - 1; e < mm; e += blockDim.x - 1) s_m[b ^ 1][e] = mass[type[nb + e]];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = a0; if (MODE == 0) { out[1] = a1; out[2] = a2; out[3] = a3; } }
}

__global__ void __launch_bounds__(256) shift3_kernel(double *__restrict__ v, long long n, double sx, double sy, double sz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  v[3 * i] = __dadd_rn(v[3 * i], -sx); v[3 * i + 1] = __dadd_rn(v[3 * i + 1], -sy); v[3 * i + 2] = __dadd_rn(v[3 * i + 2], -sz);
}
__global__ void __launch_bounds__(256) scale_kernel(double *__restrict__ v, long long n3, double s) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n3) v[e] = __dmul_rn(v[e], s);
}

int fill_args(const emd_lattice *in, LatticeArgs &L, long long *ncand) {
  if (!in || in->n[0] <= 0 || in->n[1] <= 0 || in->n[2] <= 0) { set_error("emd_lattice: bad ranges"); return 1; }
  L.ix0 = in->i0[0]; L.iy0 = in->i0[1]; L.iz0 = in->i0[2];
  L.nx = in->n[0]; L.ny = in->n[1]; L.nz = in->n[2];
  L.fcc = in->fcc ? 1 : 0;
  L.nbasis = L.fcc ? 4 : 1;
  L.a = in->a;
  static const double fcc_basis[4][3] = {{0.0, 0.0, 0.0}, {0.5, 0.5, 0.0}, {0.5, 0.0, 0.5}, {0.0, 0.5, 0.5}};
  for (int k = 0; k < 4; k++)
    for (int d = 0; d < 3; d++) L.basis[k][d] = fcc_basis[k][d] + in->offset[d];
  if (!L.fcc) for (int d = 0; d < 3; d++) L.basis[0][d] = in->offset[d];
  for (int d = 0; d < 3; d++) { L.lo[d] = in->lo[d]; L.hi[d] = in->hi[d]; }
  const long long nc = (long long)L.nx * L.ny * L.nz * L.nbasis;
  if (nc > 0x7fffffffLL) { set_error("emd_lattice: more than 2^31 candidate sites in one brick"); return 1; }
  *ncand = nc;
  return 0;
}

} // namespace

extern "C" {

int emd_lattice_count(emd_ctx *ctx, const emd_lattice *lat, int *h_n) {
  LatticeArgs L;
  long long nc;
  if (fill_args(lat, L, &nc)) return 1;
  if (ctx->s_a.ensure(sizeof(int) * ((size_t)nc + 1))) return 1;
  int *flag = ctx->s_a.as<int>();
  EMD_LAUNCH(ctx, lattice_flag_kernel, grid_for(nc, 256), 256, 0, L, nc, flag);
  EMD_CUDA(cudaMemsetAsync(flag + nc, 0, sizeof(int), ctx->stream));
  if (exclusive_scan_int(ctx, flag, flag, (int)nc + 1, nullptr)) return 1;
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, flag + nc, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream));
  *h_n = ctx->h_pinned[0];
  return 0;
}

// must follow emd_lattice_count of the same lattice (it uses the scanned offsets left in the context's scratch)
int emd_lattice_fill(emd_ctx *ctx, const emd_lattice *lat, int seed, int id_offset, const double *d_mass, double *d_x, double *d_v, double *d_q,
                     int *d_type, int *d_id) {
  LatticeArgs L;
  long long nc;
  if (fill_args(lat, L, &nc)) return 1;
  EMD_LAUNCH(ctx, lattice_fill_kernel, grid_for(nc, 256), 256, 0, L, nc, ctx->s_a.as<int>(), seed, id_offset, d_mass, d_x, d_v, d_q, d_type, d_id);
  return 0;
}

int emd_velocity_sums(emd_ctx *ctx, const double *d_v, const int *d_type, const double *d_mass, int n, int mode, double *h_out4) {
  if (ctx->s_b.ensure(4 * sizeof(double))) return 1;
  double *out = ctx->s_b.as<double>();
  EMD_CUDA(cudaMemsetAsync(out, 0, 4 * sizeof(double), ctx->stream));
  if (mode == 0) EMD_LAUNCH(ctx, seq_sum_kernel<0>, 1, 256, 0, d_v, d_type, d_mass, (long long)n, out);
  else EMD_LAUNCH(ctx, seq_sum_kernel<1>, 1, 256, 0, d_v, d_type, d_mass, (long long)n, out);
  EMD_CUDA(cudaMemcpyAsync(ctx->h_pinned, out, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  EMD_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(h_out4, ctx->h_pinned, 4 * sizeof(double));
  return 0;
}

int emd_velocity_shift(emd_ctx *ctx, double *d_v, int n, double sx, double sy, double sz) {
  if (n > 0) EMD_LAUNCH(ctx, shift3_kernel, grid_for(n, 256), 256, 0, d_v, (long long)n, sx, sy, sz);
  return 0;
}
int emd_velocity_scale(emd_ctx *ctx, double *d_v, int n, double s) {
  if (n > 0) EMD_LAUNCH(ctx, scale_kernel, grid_for(3LL * n, 256), 256, 0, d_v, 3LL * n, s);
  return 0;
}

} // extern "C"
