// ForceSNAP -- SNAP bispectrum potential (src/force_types/force_snap_neigh.h).
#ifdef MODULES_OPTION_CHECK
#endif
#ifdef FORCE_MODULES_INSTANTIATION
    else if (input->force_type == FORCE_SNAP) {
      bool half_neigh = input->force_iteration_type == FORCE_ITER_NEIGH_HALF;
      force = new ForceSNAP(input->input_data.words[input->force_line], system, half_neigh);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(FORCE_MODULES_INSTANTIATION)
#ifndef FORCE_SNAP_NEIGH_H
#define FORCE_SNAP_NEIGH_H
#include "../force.h"

class ForceSNAP : public Force {
  System *sys;
  struct Impl;
  Impl *impl;

public:
  ForceSNAP(char **args, System *system, bool half_neigh_);
  ~ForceSNAP();
  void init_coeff(int nargs, char **args);
  void compute(System *system, Binning *binning, Neighbor *neighbor);
  bool zeroes_forces() const { return false; }
  const char *name();
};
#endif
#endif
