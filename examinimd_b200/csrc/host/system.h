// system.h -- per-rank atom storage in HBM.
// Same public fields and layout contract as the reference's System (src/system.h:57-143):
// x,v,f double[N_max][3] row-major, type/id int[N_max], q double[N_max], mass double[ntypes];
// grow() is the only content-preserving resizer.  Differences that stay behind the API: the
// arrays live in device memory, a second set of buffers is kept as the destination of the
// cell sort (BinningKKSort swaps instead of copying back), and a host staging mirror is used
// by Input (lattice creation) and the binary dump.
#pragma once
#include <string>
#include <vector>
#include "types.h"

struct Particle { // src/system.h:43-55 (72 bytes = 18 ints; the multi-GPU wire format)
  T_X_FLOAT x, y, z;
  T_V_FLOAT vx, vy, vz, mass;
  T_FLOAT q;
  T_INT id;
  int type;
};

struct HostAtoms { // host staging of the first n atoms
  std::vector<double> x, v, f, q;
  std::vector<int> type, id;
  void resize(size_t n) { x.resize(3 * n); v.resize(3 * n); f.resize(3 * n); q.resize(n); type.resize(n); id.resize(n); }
};

class System {
public:
  T_INT N;       // global atoms
  T_INT N_max;   // capacity of the per-atom arrays
  T_INT N_local; // owned atoms
  T_INT N_ghost; // ghost atoms, stored behind the owned ones
  int ntypes;

  t_x x; t_v v; t_f f;
  t_type type; t_id id; t_q q;
  t_mass mass;

  T_X_FLOAT domain_x, domain_y, domain_z;
  T_X_FLOAT sub_domain_x, sub_domain_y, sub_domain_z;
  T_X_FLOAT sub_domain_lo_x, sub_domain_lo_y, sub_domain_lo_z;
  T_X_FLOAT sub_domain_hi_x, sub_domain_hi_y, sub_domain_hi_z;

  T_FLOAT boltz, mvv2e, dt;
  // sum m v^2 of the velocities thermo sees, left by a force launch that took the integrator kicks along on a thermo step
  // (Force::compute_with_nve after expect_energy(true)); valid until the next step begins
  bool mv2_cached = false;
  double mv2_cache = 0.0;
  bool do_print, print_lammps;

  emd_ctx *ctx; // device context all modules launch on
  std::string input_dir; // directory of the input deck (second place ForceSNAP looks for its coefficient files)

  System();
  ~System();
  void init();
  void destroy();
  void grow(T_INT new_N);

  // sort destination buffers (same capacity as the primary ones) and the swap that publishes them
  t_x x_alt; t_v v_alt; t_f f_alt; t_type type_alt; t_id id_alt; t_q q_alt;
  void swap_sorted();
  void swap_x() { t_x t = x; x = x_alt; x_alt = t; } // a kernel wrote the advanced positions of the owned atoms to x_alt

  // host <-> device staging of atoms [0,n)
  void upload(const HostAtoms &h, T_INT n, bool with_f);
  void download(HostAtoms &h, T_INT n) const;
  void set_mass(const std::vector<double> &m);
  std::vector<double> h_mass; // host copy of mass[] for host-side setup code

private:
  void release();
};
