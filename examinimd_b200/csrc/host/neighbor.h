// neighbor.h -- abstract neighbor-list module (plugin surface of src/neighbor.h:45-61).
// The reference's forces reach the concrete list through a static downcast and a duck-typed
// get_neigh_list(); here every list, CSR or 2D, is additionally exposed as the C-ABI's
// emd_neigh_list through one virtual, so a force module works with any Neighbor.
#pragma once
#include "types.h"
#include "system.h"
#include "binning.h"

class Neighbor {
protected:
  emd_tiles *tile_lists; // B200 fast path: tile-local full adjacency built next to the API list (kernels/tiles.cu)
  // build the tile lists; false = not applicable here (or disabled with EMD_NO_TILES=1): use the generic kernels
  bool build_tiles(System *system, Binning *binning, T_X_FLOAT neigh_cut);

public:
  int neigh_type;
  bool comm_newton;
  Neighbor() : tile_lists(nullptr), neigh_type(NEIGH_NONE), comm_newton(false) {}
  virtual ~Neighbor() { if (tile_lists) emd_tiles_destroy(tile_lists); }
  // valid tile lists of the last create_neigh_list, or NULL
  emd_tiles *tiles() const { return (tile_lists && emd_tiles_valid(tile_lists)) ? tile_lists : nullptr; }
  virtual void init(T_X_FLOAT neighcut) {}
  virtual void create_neigh_list(System *system, Binning *binning, bool half_neigh_, bool ghost_neighs_) {}
  virtual emd_neigh_list list_view() const { emd_neigh_list l = {nullptr, nullptr, nullptr, 0}; return l; }
  virtual T_INT total_neighs() const { return 0; } // entries in the list (thermo / roofline accounting)
  virtual const char *name() { return "NeighborNone"; }
};

template <int Type>
struct NeighborAdaptor { typedef Neighbor type; };

#include "modules_neighbor.h"
