#include "force_lj_neigh.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

static void fail(const char *what) {
  fprintf(stderr, "ForceLJNeigh: %s: %s\n", what, emd_last_error());
  emd_host_exit(1);
}

ForceLJNeigh::ForceLJNeigh(char **args, System *system, bool half_neigh_) : Force(args, system, half_neigh_), sys(system) {
  ntypes = system->ntypes;
  use_stackparams = (ntypes <= MAX_TYPES_STACKPARAMS);
  lj1.assign((size_t)ntypes * ntypes, 0.0);
  lj2.assign((size_t)ntypes * ntypes, 0.0);
  cutsq.assign((size_t)ntypes * ntypes, 0.0);
}

// src/force_types/force_lj_neigh_impl.h:57-98.  With <= 12 types every pair_coeff line fills
// the WHOLE table (reference quirk :66-74), otherwise the (t1,t2)/(t2,t1) entries.
void ForceLJNeigh::init_coeff(int nargs, char **args) {
  const int t1 = atoi(args[1]) - 1, t2 = atoi(args[2]) - 1;
  const double eps = atof(args[3]), sigma = atof(args[4]), cut = atof(args[5]);
  const double a = 48.0 * eps * pow(sigma, 12.0), b = 24.0 * eps * pow(sigma, 6.0), c = cut * cut;
  if (use_stackparams) {
    for (size_t k = 0; k < lj1.size(); k++) { lj1[k] = a; lj2[k] = b; cutsq[k] = c; }
  } else {
    lj1[t1 * ntypes + t2] = lj1[t2 * ntypes + t1] = a;
    lj2[t1 * ntypes + t2] = lj2[t2 * ntypes + t1] = b;
    cutsq[t1 * ntypes + t2] = cutsq[t2 * ntypes + t1] = c;
  }
  if (emd_force_lj_set_params(sys->ctx, ntypes, lj1.data(), lj2.data(), cutsq.data())) fail("set_params");
}

// src/force_types/force_lj_neigh_impl.h:100-126
void ForceLJNeigh::compute(System *system, Binning *, Neighbor *neighbor) {
  // fast path (kernels/tiles.cu): same forces on the owned atoms, every pair evaluated from both
  // sides out of shared memory.  A half list with newton on needs the ghost forces: generic kernel.
  emd_tiles *t = neighbor->tiles();
  if (t && !(half_neigh && comm_newton)) {
    if (comm_newton && system->N_ghost > 0) // ghost rows stay zero, as after the reference's deep_copy(f,0)
      emd_memset_zero(system->ctx, system->f + 3 * (size_t)system->N_local, sizeof(T_F_FLOAT) * 3 * (size_t)system->N_ghost);
    pe_cached = false;
    if (want_energy) { // a thermo step: forces and energy in one pass over the pairs
      if (emd_force_lj_compute_tiles_with_energy(system->ctx, t, system->x, system->type, system->f, &pe_cache)) fail("compute + energy (tiles)");
      pe_cached = true;
    } else if (emd_force_lj_compute_tiles(system->ctx, t, system->x, system->type, system->f, nullptr)) fail("compute (tiles)");
    return;
  }
  emd_ctx_halo_gate_wait(system->ctx);
  const emd_neigh_list l = neighbor->list_view();
  if (emd_force_lj_compute(system->ctx, system->x, system->type, system->f, system->N_local,
                           system->N_local + system->N_ghost, &l, half_neigh, /*zero_f=*/1))
    fail("compute");
}

// compute() + final_integrate + next initial_integrate in one launch (force.h): tile path only; the kernel reads the old
// positions of every staged atom, so the advanced positions go to the System's second position array, published here
bool ForceLJNeigh::compute_with_nve(System *system, Binning *, Neighbor *neighbor, T_V_FLOAT dtf, T_V_FLOAT dtv) {
  static const bool off = getenv("EMD_NO_FUSED_FORCE_NVE") && atoi(getenv("EMD_NO_FUSED_FORCE_NVE"));
  emd_tiles *t = neighbor->tiles();
  if (off || !t || comm_newton || !system->x_alt) return false;
  pe_cached = false;
  int rc;
  if (want_energy) { // a thermo step: the launch also sums the potential energy and m v^2 between the two kicks
    double mv2 = 0.0;
    rc = emd_force_lj_compute_tiles_nve_thermo(system->ctx, t, system->x, system->type, system->f, system->v, system->x_alt, system->mass,
                                               dtf, dtv, &pe_cache, &mv2);
    if (rc == 0) { pe_cached = true; system->mv2_cached = true; system->mv2_cache = mv2; }
  } else
    rc = emd_force_lj_compute_tiles_nve(system->ctx, t, system->x, system->type, system->f, system->v, system->x_alt, system->mass,
                                        dtf, dtv);
  if (rc == 3) return false; // an owned atom without a row: separate kernels
  if (rc) fail("compute_with_nve (tiles)");
  system->swap_x();
  return true;
}

// compute() in two launches for a decomposed run (force.h): the tile lists know which tiles read ghost atoms
bool ForceLJNeigh::can_split(System *, Neighbor *neighbor) {
  emd_tiles *t = neighbor->tiles();
  int n_free = 0, n_halo = 0;
  if (!t || comm_newton || emd_tiles_halo_split(t, &n_free, &n_halo)) return false;
  return n_free > 0 && n_halo > 0;
}

bool ForceLJNeigh::can_kick(System *system, Neighbor *neighbor) {
  static const bool off = getenv("EMD_NO_FUSED_FORCE_NVE") && atoi(getenv("EMD_NO_FUSED_FORCE_NVE"));
  emd_tiles *t = neighbor->tiles();
  int lists_complete = 0;
  return !off && t && !comm_newton && system->x_alt && !emd_tiles_complete(t, &lists_complete) && lists_complete;
}

// the tile kernel waits for the ghosts itself (tiles.cu: halo gate); the generic kernels below wait on the stream first
bool ForceLJNeigh::gates_halo(System *, Neighbor *neighbor) {
  // Measured on 2 and 8 B200s (profiles/r02_halo_transport.md): the split launch on the side stream still hides more (the
  // re-neighboring steps' exchanges dominate what is left), so the gate is opt-in.
  static const bool on = getenv("EMD_HALO_GATE") && atoi(getenv("EMD_HALO_GATE"));
  return on && neighbor->tiles() && !(half_neigh && comm_newton);
}

void ForceLJNeigh::compute_part(System *system, Binning *, Neighbor *neighbor, int part, const T_V_FLOAT *nve) {
  static const int reserve = getenv("EMD_OVERLAP_RESERVE") ? atoi(getenv("EMD_OVERLAP_RESERVE")) : 0;
  // part 1 runs on the side stream's SM partition (ctx.cu); EMD_OVERLAP_RESERVE additionally leaves CTA slots free
  pe_cached = false;
  if (nve) {
    if (emd_force_lj_compute_tiles_part_nve(system->ctx, neighbor->tiles(), system->x, system->type, system->f, part, part == 1 ? reserve : 0,
                                            system->v, system->x_alt, system->mass, nve[0], nve[1]))
      fail("compute_part + nve (tiles)");
    if (part == 2) system->swap_x(); // host pointers only: part 1 may still be running on the old array, which stays alive
    return;
  }
  if (emd_force_lj_compute_tiles_part(system->ctx, neighbor->tiles(), system->x, system->type, system->f, part, part == 1 ? reserve : 0))
    fail("compute_part (tiles)");
}

// src/force_types/force_lj_neigh_impl.h:128-156
T_F_FLOAT ForceLJNeigh::compute_energy(System *system, Binning *, Neighbor *neighbor) {
  double pe = 0.0;
  emd_tiles *t = neighbor->tiles();
  if (pe_cached) return pe_cache; // evaluated by the compute() of this step (expect_energy), positions unchanged since
  emd_ctx_halo_gate_wait(system->ctx);
  if (t && !(half_neigh && comm_newton)) {
    if (emd_force_lj_compute_tiles(system->ctx, t, system->x, system->type, system->f, &pe)) fail("energy (tiles)");
    return pe;
  }
  const emd_neigh_list l = neighbor->list_view();
  if (emd_force_lj_energy(system->ctx, system->x, system->type, system->N_local, &l, half_neigh, &pe)) fail("energy");
  return pe;
}

const char *ForceLJNeigh::name() { return half_neigh ? "ForceLJNeighHalf" : "ForceLJNeighFull"; }
