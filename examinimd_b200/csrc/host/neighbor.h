// neighbor.h -- abstract neighbor-list module (plugin surface of src/neighbor.h:45-61).
// The reference's forces reach the concrete list through a static downcast and a duck-typed
// get_neigh_list(); here every list, CSR or 2D, is additionally exposed as the C-ABI's
// emd_neigh_list through one virtual, so a force module works with any Neighbor.
#pragma once
#include "types.h"
#include "system.h"
#include "binning.h"

class Neighbor {
public:
  int neigh_type;
  bool comm_newton;
  Neighbor() : neigh_type(NEIGH_NONE), comm_newton(false) {}
  virtual ~Neighbor() {}
  virtual void init(T_X_FLOAT neighcut) {}
  virtual void create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool ghost_neighs_) {}
  virtual emd_neigh_list list_view() const { emd_neigh_list l = {nullptr, nullptr, nullptr, 0}; return l; }
  virtual T_INT total_neighs() const { return 0; } // entries in the list (thermo / roofline accounting)
  virtual const char *name() { return "NeighborNone"; }
};

template <int Type>
struct NeighborAdaptor { typedef Neighbor type; };

#include "modules_neighbor.h"
