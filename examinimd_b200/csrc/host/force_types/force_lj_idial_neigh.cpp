#include "force_lj_idial_neigh.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

ForceLJIDialNeigh::ForceLJIDialNeigh(char **args, System *system, bool half_neigh_) : Force(args, system, half_neigh_), sys(system), step(0) {
  ntypes = system->ntypes;
  lj1.assign((size_t)ntypes * ntypes, 0.0);
  lj2.assign((size_t)ntypes * ntypes, 0.0);
  cutsq.assign((size_t)ntypes * ntypes, 0.0);
  intensity.assign((size_t)ntypes * ntypes, 0.0);
}

// src/force_types/force_lj_idial_neigh_impl.h:50-88: `pair_coeff t1 t2 eps sigma cut nrepeat`; unlike ForceLJNeigh there is
// no stack-parameter path, a line sets its own (t1,t2)/(t2,t1) entries only
void ForceLJIDialNeigh::init_coeff(int nargs, char **args) {
  if (nargs < 7) {
    fprintf(stderr, "ForceLJIDialNeigh: pair_coeff needs `t1 t2 eps sigma cut nrepeat`\n");
    emd_host_exit(1);
  }
  const int t1 = atoi(args[1]) - 1, t2 = atoi(args[2]) - 1;
  const double eps = atof(args[3]), sigma = atof(args[4]), cut = atof(args[5]);
  const int nrepeat = atoi(args[6]);
  if (t1 < 0 || t2 < 0 || t1 >= ntypes || t2 >= ntypes) {
    fprintf(stderr, "ForceLJIDialNeigh: pair_coeff type out of range\n");
    emd_host_exit(1);
  }
  lj1[t1 * ntypes + t2] = lj1[t2 * ntypes + t1] = 48.0 * eps * pow(sigma, 12.0);
  lj2[t1 * ntypes + t2] = lj2[t2 * ntypes + t1] = 24.0 * eps * pow(sigma, 6.0);
  cutsq[t1 * ntypes + t2] = cutsq[t2 * ntypes + t1] = cut * cut;
  intensity[t1 * ntypes + t2] = intensity[t2 * ntypes + t1] = nrepeat;
  if (emd_force_lj_set_params(sys->ctx, ntypes, lj1.data(), lj2.data(), cutsq.data())) {
    fprintf(stderr, "ForceLJIDialNeigh: set_params: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

// src/force_types/force_lj_idial_neigh_impl.h:90-111
void ForceLJIDialNeigh::compute(System *system, Binning *, Neighbor *neighbor) {
  const emd_neigh_list l = neighbor->list_view();
  if (emd_force_lj_idial_compute(system->ctx, system->x, system->type, system->f, system->N_local, system->N_local + system->N_ghost, &l,
                                 half_neigh, /*zero_f=*/1, intensity.data())) {
    fprintf(stderr, "ForceLJIDialNeigh: compute: %s\n", emd_last_error());
    emd_host_exit(1);
  }
  step++;
}

const char *ForceLJIDialNeigh::name() { return half_neigh ? "ForceLJIDialNeighHalf" : "ForceLJIDialNeighFull"; }
