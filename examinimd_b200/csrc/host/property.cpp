#include "property.h"
#include <cstdio>
#include <cstdlib>

static T_V_FLOAT sum_mv2(System *system) {
  if (system->mv2_cached) return system->mv2_cache; // summed by the fused force + integrator launch of this (thermo) step
  double s = 0.0;
  if (emd_reduce_mv2(system->ctx, system->v, system->type, system->mass, system->N_local, &s)) {
    fprintf(stderr, "thermo reduction failed: %s\n", emd_last_error());
    emd_host_exit(1);
  }
  return s;
}

// src/property_temperature.cpp:43-62
T_V_FLOAT Temperature::compute(System *system) {
  T_V_FLOAT T = sum_mv2(system);
  T_INT dof = 3 * system->N - 3;
  T_V_FLOAT factor = system->mvv2e / (1.0 * dof * system->boltz);
  comm->reduce_float(&T, 1);
  return T * factor;
}

// src/property_kine.cpp:43-61
T_V_FLOAT KinE::compute(System *system) {
  T_V_FLOAT KE = sum_mv2(system);
  T_V_FLOAT factor = 0.5 * system->mvv2e;
  comm->reduce_float(&KE, 1);
  return KE * factor;
}

// src/property_pote.cpp:44-49
T_F_FLOAT PotE::compute(System *system, Binning *binning, Neighbor *neighbor, Force *force) {
  T_F_FLOAT PE = force->compute_energy(system, binning, neighbor);
  comm->reduce_float(&PE, 1);
  return PE;
}
