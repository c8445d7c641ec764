// ForceLJNeigh -- Lennard-Jones over a neighbor list (src/force_types/force_lj_neigh.h).
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "NEIGH_FULL") == 0)) force_iteration_type = FORCE_ITER_NEIGH_FULL;
      if ((strcmp(argv[i + 1], "NEIGH_HALF") == 0)) force_iteration_type = FORCE_ITER_NEIGH_HALF;
      if ((strcmp(argv[i + 1], "CELL_FULL") == 0)) force_iteration_type = FORCE_ITER_CELL_FULL;
#endif
#ifdef FORCE_MODULES_INSTANTIATION
    else if (input->force_type == FORCE_LJ) {
      // CELL_FULL has no live implementation in the reference either (its factory fragment is
      // never compiled in, force_lj_cell.h:43-47) and falls through to the full neighbor list
      bool half_neigh = input->force_iteration_type == FORCE_ITER_NEIGH_HALF;
      force = new ForceLJNeigh(input->input_data.words[input->force_line], system, half_neigh);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(FORCE_MODULES_INSTANTIATION)
#ifndef FORCE_LJ_NEIGH_H
#define FORCE_LJ_NEIGH_H
#include "../force.h"
#include <vector>

class ForceLJNeigh : public Force {
private:
  int ntypes;
  bool use_stackparams;
  std::vector<T_F_FLOAT> lj1, lj2, cutsq; // [ntypes][ntypes] host tables, pushed into the context
  System *sys;
  bool want_energy = false, pe_cached = false; // expect_energy(): the next compute() also evaluates the energy
  T_F_FLOAT pe_cache = 0.0;

public:
  ForceLJNeigh(char **args, System *system, bool half_neigh_);
  void init_coeff(int nargs, char **args);
  void compute(System *system, Binning *binning, Neighbor *neighbor);
  T_F_FLOAT compute_energy(System *system, Binning *binning, Neighbor *neighbor);
  void expect_energy(bool on) { want_energy = on; if (!on) pe_cached = false; }
  bool zeroes_forces() const { return true; }
  bool compute_with_nve(System *system, Binning *binning, Neighbor *neighbor, T_V_FLOAT dtf, T_V_FLOAT dtv);
  bool can_split(System *system, Neighbor *neighbor);
  void compute_part(System *system, Binning *binning, Neighbor *neighbor, int part, const T_V_FLOAT *nve = nullptr);
  bool can_kick(System *system, Neighbor *neighbor);
  bool gates_halo(System *system, Neighbor *neighbor);
  const char *name();
};
#endif
#endif
