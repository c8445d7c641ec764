// IntegratorNVE -- velocity Verlet (src/integrator_nve.h, src/integrator_nve.cpp:41-121).
#ifndef INTEGRATOR_NVE_H
#define INTEGRATOR_NVE_H
#include "integrator.h"

class IntegratorNVE : public Integrator {
  T_V_FLOAT dtv, dtf;

public:
  IntegratorNVE(System *s);
  void initial_integrate();
  void final_integrate();
  void final_initial_integrate();
  const char *name();
};
#endif
