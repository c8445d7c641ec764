#include "comm_serial.h"
#include <cstdio>
#include <cstdlib>

static void fail(const char *what) {
  fprintf(stderr, "CommSerial: %s: %s\n", what, emd_last_error());
  emd_host_exit(1);
}

CommSerial::CommSerial(System *s, T_X_FLOAT comm_depth_) : Comm(s, comm_depth_) {
  printf("CommSerial\n"); // part of the reference's stdout (comm_serial.cpp:42)
  for (int p = 0; p < 6; p++) num_ghost[p] = ghost_offsets[p] = 0;
}

// src/comm_types/comm_serial.cpp:47-54
void CommSerial::exchange() {
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  if (emd_comm_wrap(system->ctx, system->x, system->N_local, L)) fail("wrap");
}

// src/comm_types/comm_serial.cpp:56-97: six dimension-ordered phases; phase p also scans the
// ghosts made by earlier dimensions, but an odd phase skips the ghosts of its own even twin.
void CommSerial::exchange_halo() {
  const T_INT N_local = system->N_local;
  T_INT N_ghost = 0;
  const double L[3] = {system->domain_x, system->domain_y, system->domain_z};
  const double lo[3] = {system->sub_domain_lo_x, system->sub_domain_lo_y, system->sub_domain_lo_z};
  const double hi[3] = {system->sub_domain_hi_x, system->sub_domain_hi_y, system->sub_domain_hi_z};
  for (int phase = 0; phase < 6; phase++) {
    const T_INT nparticles = N_local + N_ghost - ((phase % 2 == 1) ? num_ghost[phase - 1] : 0);
    int count = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
      if (emd_comm_halo_phase(system->ctx, phase, system->x, system->v, system->q, system->id, system->type, nparticles,
                              N_local + N_ghost, system->N_max, pack_indicies[phase].ptr,
                              (int)pack_indicies[phase].extent(), L, lo, hi, comm_depth, &count))
        fail("halo_phase");
      bool redo = false;
      if (N_local + N_ghost + count > system->N_max) { // :75-79
        system->grow(N_local + N_ghost + count + count / 4);
        redo = true;
      }
      if ((size_t)count > pack_indicies[phase].extent()) { // :80-84
        if (!pack_indicies[phase].alloc((size_t)(count * 1.5) + 1024)) fail("alloc pack_indicies");
        redo = true;
      }
      if (!redo) break;
    }
    num_ghost[phase] = count;
    N_ghost += count;
  }
  system->N_ghost = N_ghost;
  // resolve every ghost to its owned root atom + total shift for the single-kernel refresh
  if (ghost_root.extent() < (size_t)N_ghost) {
    if (!ghost_root.alloc((size_t)N_ghost + N_ghost / 2 + 1) || !ghost_shift.alloc(3 * ((size_t)N_ghost + N_ghost / 2 + 1))) fail("alloc ghost_root");
  }
  const int *lists[6];
  int counts[6];
  for (int p = 0; p < 6; p++) { lists[p] = pack_indicies[p].ptr; counts[p] = num_ghost[p]; }
  if (emd_comm_halo_resolve(system->ctx, lists, counts, N_local, L, ghost_root.ptr, ghost_shift.ptr)) fail("halo_resolve");
}

// src/comm_types/comm_serial.cpp:99-110: the six phase copies collapse into one kernel (see emd_comm_halo_resolve)
void CommSerial::update_halo() {
  if (emd_comm_halo_refresh(system->ctx, system->x, system->N_local, system->N_ghost, ghost_root.ptr, ghost_shift.ptr))
    fail("halo_refresh");
}

// src/comm_types/comm_serial.cpp:112-127
void CommSerial::update_force() {
  ghost_offsets[0] = system->N_local;
  for (int phase = 1; phase < 6; phase++) ghost_offsets[phase] = ghost_offsets[phase - 1] + num_ghost[phase - 1];
  for (int phase = 5; phase >= 0; phase--)
    if (emd_comm_force_fold_phase(system->ctx, system->f, pack_indicies[phase].ptr, num_ghost[phase], ghost_offsets[phase]))
      fail("force_fold_phase");
}

const char *CommSerial::name() { return "CommSerial"; }
