// ForceLJIDialNeigh -- pair_style lj/cut/idial: Lennard-Jones with a compute-intensity dial
// (src/force_types/force_lj_idial_neigh.h:41-140, force_lj_idial_neigh_impl.h).
#ifdef MODULES_OPTION_CHECK
      // the --force-iteration words are parsed by force_lj_neigh.h
#endif
#ifdef FORCE_MODULES_INSTANTIATION
    else if (input->force_type == FORCE_LJ_IDIAL) {
      bool half_neigh = input->force_iteration_type == FORCE_ITER_NEIGH_HALF;
      force = new ForceLJIDialNeigh(input->input_data.words[input->force_line], system, half_neigh);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(FORCE_MODULES_INSTANTIATION)
#ifndef FORCE_LJ_IDIAL_NEIGH_H
#define FORCE_LJ_IDIAL_NEIGH_H
#include "../force.h"
#include <vector>

class ForceLJIDialNeigh : public Force {
private:
  int ntypes;
  std::vector<T_F_FLOAT> lj1, lj2, cutsq, intensity; // [ntypes][ntypes] host tables
  System *sys;
  int step;

public:
  ForceLJIDialNeigh(char **args, System *system, bool half_neigh_);
  void init_coeff(int nargs, char **args);
  void compute(System *system, Binning *binning, Neighbor *neighbor);
  bool zeroes_forces() const { return true; }
  const char *name();
};
#endif
#endif
