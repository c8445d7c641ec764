// integrator.cu -- velocity-Verlet half steps.
// Replaces IntegratorNVE::initial_integrate / final_integrate (src/integrator_nve.cpp:77-83,
// 115-121; functors :47-74, :87-112).  Pure streaming kernels, HBM-bound:
//   initial: read x,v,f (72 B) + type (4 B), write x,v (48 B)  = 124 B/atom
//   final  : read v,f (48 B) + type (4 B), write v (24 B)      =  76 B/atom
// One thread per vector COMPONENT so that every warp access is a contiguous 256 B run; the
// arithmetic uses separate multiply and add (no FMA) so positions and velocities are
// bit-identical to the reference's x86-64 CPU build given identical forces.
#include "common.cuh"

using namespace emd;

namespace {

__global__ void __launch_bounds__(256) nve_initial_kernel(double *__restrict__ x, double *__restrict__ v,
                                                          const double *__restrict__ f, const int *__restrict__ type,
                                                          const double *__restrict__ mass, long long n3, double dtf,
                                                          double dtv) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int i = (int)(e / 3);
  const double dtfm = dtf / mass[type[i]];              // integrator_nve.cpp:67
  const double vn = __dadd_rn(v[e], __dmul_rn(dtfm, f[e])); // :68-70
  v[e] = vn;
  x[e] = __dadd_rn(x[e], __dmul_rn(dtv, vn));           // :71-73
}

__global__ void __launch_bounds__(256) nve_final_kernel(double *__restrict__ v, const double *__restrict__ f,
                                                        const int *__restrict__ type, const double *__restrict__ mass,
                                                        long long n3, double dtf) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int i = (int)(e / 3);
  const double dtfm = dtf / mass[type[i]];              // :106
  v[e] = __dadd_rn(v[e], __dmul_rn(dtfm, f[e]));        // :107-109
}

// final_integrate of step n and initial_integrate of step n+1 in one pass (124 B/atom instead of 76 + 124): the same
// operations in the same order on the same values, so x and v are bit-identical to the two separate kernels.
__global__ void __launch_bounds__(256) nve_final_initial_kernel(double *__restrict__ x, double *__restrict__ v,
                                                                const double *__restrict__ f, const int *__restrict__ type,
                                                                const double *__restrict__ mass, long long n3, double dtf,
                                                                double dtv) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n3) return;
  const int i = (int)(e / 3);
  const double dtfm = dtf / mass[type[i]];
  const double kick = __dmul_rn(dtfm, f[e]);
  const double v1 = __dadd_rn(v[e], kick);              // final_integrate, integrator_nve.cpp:107-109
  const double v2 = __dadd_rn(v1, kick);                // initial_integrate of the next step, :68-70 (f unchanged in between)
  v[e] = v2;
  x[e] = __dadd_rn(x[e], __dmul_rn(dtv, v2));           // :71-73
}

} // namespace

extern "C" {

int emd_nve_final_initial_integrate(emd_ctx *ctx, double *d_x, double *d_v, const double *d_f, const int *d_type,
                                    const double *d_mass, int n_local, double dtf, double dtv) {
  if (n_local <= 0) return 0;
  const long long n3 = 3LL * n_local;
  EMD_LAUNCH(ctx, nve_final_initial_kernel, grid_for(n3, 256), 256, 0, d_x, d_v, d_f, d_type, d_mass, n3, dtf, dtv);
  return 0;
}

int emd_nve_initial_integrate(emd_ctx *ctx, double *d_x, double *d_v, const double *d_f, const int *d_type,
                              const double *d_mass, int n_local, double dtf, double dtv) {
  if (n_local <= 0) return 0;
  const long long n3 = 3LL * n_local;
  EMD_LAUNCH(ctx, nve_initial_kernel, grid_for(n3, 256), 256, 0, d_x, d_v, d_f, d_type, d_mass, n3, dtf, dtv);
  return 0;
}

int emd_nve_final_integrate(emd_ctx *ctx, double *d_v, const double *d_f, const int *d_type, const double *d_mass,
                            int n_local, double dtf) {
  if (n_local <= 0) return 0;
  const long long n3 = 3LL * n_local;
  EMD_LAUNCH(ctx, nve_final_kernel, grid_for(n3, 256), 256, 0, d_v, d_f, d_type, d_mass, n3, dtf);
  return 0;
}

} // extern "C"
