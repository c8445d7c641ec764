#include "integrator_nve.h"
#include <cstdio>
#include <cstdlib>

// dt and mvv2e are captured at construction, i.e. after the deck is parsed (integrator_nve.cpp:41-44)
IntegratorNVE::IntegratorNVE(System *s) : Integrator(s) {
  dtv = system->dt;
  dtf = 0.5 * system->dt / system->mvv2e;
  timestep_size = system->dt;
}

void IntegratorNVE::initial_integrate() {
  if (emd_nve_initial_integrate(system->ctx, system->x, system->v, system->f, system->type, system->mass, system->N_local,
                                dtf, dtv)) {
    fprintf(stderr, "IntegratorNVE::initial_integrate: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

void IntegratorNVE::final_integrate() {
  if (emd_nve_final_integrate(system->ctx, system->v, system->f, system->type, system->mass, system->N_local, dtf)) {
    fprintf(stderr, "IntegratorNVE::final_integrate: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

void IntegratorNVE::final_initial_integrate() {
  if (emd_nve_final_initial_integrate(system->ctx, system->x, system->v, system->f, system->type, system->mass, system->N_local,
                                      dtf, dtv)) {
    fprintf(stderr, "IntegratorNVE::final_initial_integrate: %s\n", emd_last_error());
    emd_host_exit(1);
  }
}

const char *IntegratorNVE::name() { return "IntegratorNVE"; }
