// integrator.h -- abstract time integrator (plugin surface of src/integrator.h:45-55).
#pragma once
#include "types.h"
#include "system.h"

class Integrator {
public:
  System *system;
  T_V_FLOAT timestep_size;
  Integrator(System *s) : system(s), timestep_size(0.0) {}
  virtual ~Integrator() {}
  virtual void initial_integrate() {}
  virtual void final_integrate() {}
  // final_integrate() immediately followed by the next step's initial_integrate() (not in the reference; the driver calls
  // it when nothing observes the state between two steps).  Default: the two calls.
  virtual void final_initial_integrate() { final_integrate(); initial_integrate(); }
  // the two factors of a plain velocity-Verlet step (v += dtf/m f, x += dtv v), for Force::compute_with_nve; an integrator
  // that is not of that form returns false
  virtual bool step_factors(T_V_FLOAT *dtf, T_V_FLOAT *dtv) { return false; }
  virtual const char *name() { return "IntegratorNone"; }
};

#include "modules_integrator.h"
