// CommSerial -- single-process periodic boundaries (src/comm_types/comm_serial.{h,cpp}).
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "SERIAL") == 0)) comm_type = COMM_SERIAL;
#endif
#ifdef COMM_MODULES_INSTANTIATION
      else if (input->comm_type == COMM_SERIAL) {
        comm = new CommSerial(system, input->force_cutoff + input->neighbor_skin);
      }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(COMM_MODULES_INSTANTIATION)
#ifndef COMM_SERIAL_H
#define COMM_SERIAL_H
#include "../comm.h"

class CommSerial : public Comm {
  T_INT num_ghost[6];
  T_INT ghost_offsets[6];
  DeviceArray<T_INT> pack_indicies[6]; // source index of every ghost, per phase (replayed by update_halo)
  DeviceArray<T_INT> ghost_root;       // owned root atom of every ghost   } resolved once per exchange_halo, so that
  DeviceArray<T_FLOAT> ghost_shift;    // its total periodic shift [.][3]   } update_halo is a single kernel

public:
  CommSerial(System *s, T_X_FLOAT comm_depth_);
  void exchange();
  void exchange_halo();
  void update_halo();
  void update_force();
  const char *name();
  const T_INT *ghost_counts() const { return num_ghost; }
};
#endif
#endif
