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
  // Optional (not in the reference): the driver announces that compute_energy() follows the next compute() on unchanged
  // positions (a thermo step, src/examinimd.cpp:252-267); a module may then evaluate both in one pass over the pairs and
  // answer compute_energy() from that pass.  The default ignores the hint.
  virtual void expect_energy(bool) {}
  // true when compute() zeroes/overwrites f itself, so the driver may skip deep_copy(f,0)
  // (src/examinimd.cpp:232); a foreign Force plugin simply inherits `false`
  virtual bool zeroes_forces() const { return false; }
  // Optional split of compute() for decomposed runs (not in the reference): part 1 = the share that reads no ghost atom
  // (the driver runs it on the context's side stream while Comm::update_halo is in flight), part 2 = the rest; the two
  // parts together equal compute().  A module that cannot split inherits `false` and the driver calls compute().
  // Optional: compute() and, in the same pass over the atoms, Integrator::final_integrate of this step + initial_integrate of
  // the next with the given factors (v += dtf/m f twice, x += dtv v); the module publishes the new positions itself.  The
  // driver offers it between two unobserved steps; a module that cannot do it returns false and nothing has happened.
  virtual bool compute_with_nve(System *system, Binning *binning, Neighbor *neigh, T_V_FLOAT dtf, T_V_FLOAT dtv) { return false; }
  virtual bool can_split(System *system, Neighbor *neigh) { return false; }
  // nve != nullptr: {dtf, dtv}; the part also carries the integrator kick as compute_with_nve does (only offered when
  // can_kick() said so); part 2 then publishes the new positions
  virtual void compute_part(System *system, Binning *binning, Neighbor *neigh, int part, const T_V_FLOAT *nve = nullptr) {}
  virtual bool can_kick(System *system, Neighbor *neigh) { return false; }
  // Optional: the next compute() / compute_with_nve() waits for the neighbours' ghost stores of this step itself (only before
  // the first pair that reads a ghost), so the driver may ask the Comm module for a refresh that does not block the stream
  // (Comm::update_halo_deferred).  A module that inherits `false` always sees a completed Comm::update_halo().
  virtual bool gates_halo(System *system, Neighbor *neigh) { return false; }
  virtual const char *name() { return "ForceNone"; }
};

#include "modules_force.h"
