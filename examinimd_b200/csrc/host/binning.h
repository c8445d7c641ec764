// binning.h -- abstract cell-binning module (plugin surface of src/binning.h:43-67).
// Public data members keep the reference's names; bincount/binoffsets are the flattened
// [nbinx][nbiny][nbinz] device arrays (x slowest, z fastest).
#pragma once
#include "types.h"
#include "system.h"

class Binning {
protected:
  System *system;

public:
  T_INT nbinx, nbiny, nbinz, nhalo;
  T_X_FLOAT minx, maxx, miny, maxy, minz, maxz;

  typedef int *t_bincount;
  typedef T_INT *t_binoffsets;
  typedef T_INT *t_permute_vector;

  t_bincount bincount;
  t_binoffsets binoffsets;
  t_permute_vector permute_vector;

  bool is_sorted;

  Binning(System *s) : system(s), nbinx(0), nbiny(0), nbinz(0), nhalo(0), minx(0), maxx(0), miny(0), maxy(0), minz(0),
                       maxz(0), bincount(nullptr), binoffsets(nullptr), permute_vector(nullptr), is_sorted(false) {}
  virtual ~Binning() {}
  virtual void create_binning(T_X_FLOAT dx, T_X_FLOAT dy, T_X_FLOAT dz, int halo_depth, bool do_local, bool do_ghost,
                              bool sort) {}
  virtual const char *name() { return "BinningNone"; }

  emd_bin_geom geom() const {
    emd_bin_geom g;
    g.nbinx = nbinx; g.nbiny = nbiny; g.nbinz = nbinz; g.nhalo = nhalo;
    g.minx = minx; g.maxx = maxx; g.miny = miny; g.maxy = maxy; g.minz = minz; g.maxz = maxz;
    return g;
  }
};

#include "modules_binning.h"
