// BinningKKSort -- device counting sort into cells (keeps the reference class name so that the
// "Using:" line matches, src/examinimd.cpp:113).  Interface: src/binning_types/binning_kksort.h:42-52.
#ifndef BINNING_KKSORT_H
#define BINNING_KKSORT_H
#include "../binning.h"

class BinningKKSort : public Binning {
  DeviceArray<int> bincount_buf, binoffsets_buf, permute_buf;

public:
  BinningKKSort(System *s);
  void create_binning(T_X_FLOAT dx, T_X_FLOAT dy, T_X_FLOAT dz, int halo_depth, bool do_local, bool do_ghost, bool sort);
  const char *name();
};
#endif
