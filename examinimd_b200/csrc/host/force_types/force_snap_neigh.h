// ForceSNAP -- SNAP bispectrum potential (src/force_types/force_snap_neigh.h).
// Same three-way header as the reference module; the class parses the pair_coeff line and the two
// coefficient files exactly like ForceSNAP<>::init_coeff/read_files and hands the result to the
// sm_100a kernels of kernels/snap.cu through emd_snap_create / emd_force_snap_compute.
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
#include <string>
#include <vector>

class ForceSNAP : public Force {
  System *sys;
  emd_snap *snap;

  // what init_coeff / read_files leave behind (force_snap_neigh.h:95-130)
  int nelements, ncoeffall, ncoeff;
  std::vector<std::string> elements;
  std::vector<double> radelem, wjelem, coeffelem; // [nelements], [nelements][ncoeffall]
  std::vector<int> map;                          // [ntypes+1], filled from index 1 (:293-306)
  double rcutfac, rfac0, rmin0, rcutmax;
  int twojmax, diagonalstyle, switchflag, bzeroflag, quadraticflag;

  void read_files(const char *coefffilename, const char *paramfilename);

public:
  ForceSNAP(char **args, System *system, bool half_neigh_);
  ~ForceSNAP();
  void init_coeff(int nargs, char **args);
  void compute(System *system, Binning *binning, Neighbor *neighbor);
  // The reference does not override Force::compute_energy (thermo PE prints 0, src/force.h:54), and by default neither does
  // this class's behaviour differ: 0 is returned.  With EMD_SNAP_ENERGY=1 in the environment (an extension)
  // the SNAP energy of the owned atoms is evaluated (emd_force_snap_energy), which gives the deck a total-energy drift check.
  T_F_FLOAT compute_energy(System *system, Binning *binning, Neighbor *neighbor);
  bool zeroes_forces() const { return false; }
  const char *name();
  T_F_FLOAT cutoff() const { return rcutmax; }
  int num_coeff() const { return ncoeff; }
  emd_snap *handle() const { return snap; }
};
#endif
#endif
