// force.h -- abstract short-range force module (plugin surface of src/force.h:46-57).
#pragma once
#include "types.h"
#include "system.h"
#include "binning.h"
#include "neighbor.h"

class Force {
public:
  bool half_neigh, comm_newton;
  Force(char **args, System *system, bool half_neigh_) : half_neigh(half_neigh_), comm_newton(false) {}
  virtual ~Force() {}
  virtual void init_coeff(int nargs, char **args) {}
  virtual void compute(System *system, Binning *binning, Neighbor *neigh) {}
  virtual T_F_FLOAT compute_energy(System *system, Binning *binning, Neighbor *neigh) { return 0.0; } // thermo only
  // true when compute() zeroes/overwrites f itself, so the driver may skip deep_copy(f,0)
  // (src/examinimd.cpp:232); a foreign Force plugin simply inherits `false`
  virtual bool zeroes_forces() const { return false; }
  virtual const char *name() { return "ForceNone"; }
};

#include "modules_force.h"
