// types.h -- scalar types, enums and device-array handles of the host layer.
// Mirrors the role of the reference's src/types.h:45-113: the same enum names and the same
// per-atom element types (FP64 everywhere, 32-bit indices), but "views" are plain device
// pointers into HBM (row-major [N][3]) that are handed to the emd_* C ABI unchanged.
#pragma once
#include <cstddef>
#include "emd_b200.h"

enum { UNITS_REAL, UNITS_LJ, UNITS_METAL };
enum { LATTICE_SC, LATTICE_FCC };
enum { INTEGRATOR_NVE };
enum { BINNING_KKSORT };
enum { COMM_SERIAL, COMM_MPI };
enum { FORCE_LJ, FORCE_LJ_IDIAL, FORCE_SNAP };
enum { FORCE_ITER_CELL_FULL, FORCE_ITER_NEIGH_FULL, FORCE_ITER_NEIGH_HALF };
enum { NEIGH_NONE, NEIGH_CSR, NEIGH_CSR_MAPCONSTR, NEIGH_2D };
enum { INPUT_LAMMPS };

#define MAX_TYPES_STACKPARAMS 12

// Fatal errors of the host layer.  The reference prints and calls exit(1)/MPI_Abort (src/comm.cpp:65-68); the driver
// binary keeps that.  Inside the session C API (capi.cpp) the same call sites must not take the embedding process down:
// emd_host_exit then throws EmdFatal, which every emd_app_* entry point turns into a non-zero return code + emd_last_error().
struct EmdFatal { int code; };
void emd_host_throw_on_exit(bool on);
[[noreturn]] void emd_host_exit(int code);

typedef int T_INT;
typedef double T_FLOAT;
typedef double T_X_FLOAT;
typedef double T_V_FLOAT;
typedef double T_F_FLOAT;

typedef T_X_FLOAT *t_x;   // [N][3] device
typedef T_V_FLOAT *t_v;   // [N][3] device
typedef T_F_FLOAT *t_f;   // [N][3] device
typedef int *t_type;      // [N] device
typedef T_INT *t_id;      // [N] device
typedef T_FLOAT *t_q;     // [N] device
typedef T_V_FLOAT *t_mass; // [ntypes] device

// Owning device buffer (grow-only unless reset); the analogue of a managed Kokkos::View.
template <class T>
struct DeviceArray {
  T *ptr = nullptr;
  size_t count = 0;
  DeviceArray() {}
  DeviceArray(const DeviceArray &) = delete;
  DeviceArray &operator=(const DeviceArray &) = delete;
  ~DeviceArray() { reset(); }
  void reset() { if (ptr) emd_free(ptr); ptr = nullptr; count = 0; }
  // discard contents, new size (Kokkos::realloc)
  bool alloc(size_t n) {
    reset();
    void *p = nullptr;
    if (emd_malloc(&p, sizeof(T) * (n ? n : 1))) return false;
    ptr = static_cast<T *>(p); count = n;
    return true;
  }
  size_t extent() const { return count; }
};
