// NeighborCSR -- CSR neighbor list built by the warp-ballot kernels of kernels/neighbor.cu.
// Same three-way header as the reference's src/neighbor_types/neighbor_csr.h: an option-check
// fragment for Input::read_command_line_args, an instantiation fragment for ExaMiniMD::init,
// and the class itself.
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "CSR") == 0)) neighbor_type = NEIGH_CSR;
#endif
#ifdef NEIGHBOR_MODULES_INSTANTIATION
    else if (input->neighbor_type == NEIGH_CSR) {
      neighbor = new NeighborCSR();
      neighbor->init(input->force_cutoff + input->neighbor_skin);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(NEIGHBOR_MODULES_INSTANTIATION)
#ifndef NEIGHBOR_CSR_H
#define NEIGHBOR_CSR_H
#include "../neighbor.h"

// the list concept of src/neighbor_types/neighbor_csr.h:81-124, as device pointers
struct NeighListCSR {
  const T_INT *row_map; // [N_local+1]
  const T_INT *entries; // [row_map[N_local]]
  T_INT N_local, total;
};

class NeighborCSR : public Neighbor {
protected:
  T_X_FLOAT neigh_cut;
  DeviceArray<T_INT> neigh_offsets, neighs; // grow-only, as the reference's views
  NeighListCSR neigh_list;

public:
  typedef NeighListCSR t_neigh_list;
  NeighborCSR() : neigh_cut(0.0) { neigh_type = NEIGH_CSR; neigh_list = {nullptr, nullptr, 0, 0}; }
  void init(T_X_FLOAT neigh_cut_) { neigh_cut = neigh_cut_; }
  void create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool);
  t_neigh_list get_neigh_list() { return neigh_list; }
  emd_neigh_list list_view() const { emd_neigh_list l = {neigh_list.row_map, nullptr, neigh_list.entries, 1}; return l; }
  T_INT total_neighs() const { return neigh_list.total; }
  const char *name() { return "NeighborCSR"; }
};

template <>
struct NeighborAdaptor<NEIGH_CSR> { typedef NeighborCSR type; };
#endif
#endif
