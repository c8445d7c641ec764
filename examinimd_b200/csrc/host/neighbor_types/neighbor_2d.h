// Neighbor2D -- dense [N_local+1][maxneighs] neighbor table (the reference's default
// --neigh-type, src/neighbor_types/neighbor_2d.h), filled by the same warp-ballot traversal.
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "2D") == 0)) neighbor_type = NEIGH_2D;
#endif
#ifdef NEIGHBOR_MODULES_INSTANTIATION
    else if (input->neighbor_type == NEIGH_2D) {
      neighbor = new Neighbor2D();
      neighbor->init(input->force_cutoff + input->neighbor_skin);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(NEIGHBOR_MODULES_INSTANTIATION)
#ifndef NEIGHBOR_2D_H
#define NEIGHBOR_2D_H
#include "../neighbor.h"

// the list concept of src/neighbor_types/neighbor_2d.h:80-120
struct NeighList2D {
  T_INT maxneighs;           // row stride
  const T_INT *num_neighs;   // [N_local+1]
  const T_INT *neighs;       // [N_local+1][maxneighs] row-major
  T_INT N_local;
  NeighList2D() : maxneighs(16), num_neighs(nullptr), neighs(nullptr), N_local(0) {} // :102-104
};

class Neighbor2D : public Neighbor {
protected:
  T_X_FLOAT neigh_cut;
  DeviceArray<T_INT> num_neighs_buf, neighs_buf;
  size_t rows_cap, cols_cap;
  NeighList2D neigh_list;
  long long last_total;

public:
  typedef NeighList2D t_neigh_list;
  int fill_passes; // how many fill passes the last build needed (the reference's do/while, :304-330)
  Neighbor2D() : neigh_cut(0.0), rows_cap(0), cols_cap(0), last_total(0), fill_passes(0) { neigh_type = NEIGH_2D; }
  void init(T_X_FLOAT neigh_cut_) { neigh_cut = neigh_cut_; neigh_list = NeighList2D(); }
  void create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool);
  t_neigh_list get_neigh_list() { return neigh_list; }
  emd_neigh_list list_view() const {
    emd_neigh_list l = {nullptr, neigh_list.num_neighs, neigh_list.neighs, neigh_list.maxneighs};
    return l;
  }
  T_INT total_neighs() const { return (T_INT)last_total; }
  const char *name() { return "Neighbor2D"; }
};

template <>
struct NeighborAdaptor<NEIGH_2D> { typedef Neighbor2D type; };
#endif
#endif
