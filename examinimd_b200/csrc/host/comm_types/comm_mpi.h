// CommMPI -- 3-D brick domain decomposition across the GPUs of one NVSwitch box, one process per GPU
// (src/comm_types/comm_mpi.{h,cpp}).  Keeps the reference's class name, flag (`--comm-type MPI`; NCCL is
// accepted as an alias) and protocol; the messages travel as NCCL send/recv pairs over NVLink on the
// module stream instead of MPI (kernels/comm_mpi.cu).  Rank and world size come from the launcher's
// environment (RANK / WORLD_SIZE as set by torchrun, or OMPI_COMM_WORLD_* / PMI_*).
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "MPI") == 0) || (strcmp(argv[i + 1], "NCCL") == 0)) comm_type = COMM_MPI;
#endif
#ifdef COMM_MODULES_INSTANTIATION
      else if (input->comm_type == COMM_MPI) {
        comm = new CommMPI(system, input->force_cutoff + input->neighbor_skin);
      }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(COMM_MODULES_INSTANTIATION)
#ifndef COMM_MPI_H
#define COMM_MPI_H
#include "../comm.h"

class CommMPI : public Comm {
  emd_net *net;
  emd_peer *peer; // per-step halo refresh by peer stores over NVLink (kernels/comm_peer.cu); nullptr: NCCL send/recv
  emd_decomp dec;
  int proc_rank, proc_size;
  T_INT proc_num_send[6], proc_num_recv[6]; // atoms shipped / received per phase by the last exchange_halo
  T_INT num_ghost[6], ghost_offsets[6];
  DeviceArray<T_INT> pack_indicies[6];      // pack_indicies_all(phase, :)
  DeviceArray<char> pack_buffer, unpack_buffer;
  DeviceArray<char> pack_buffer2, unpack_buffer2; // the odd phase of a dimension (both phases of a dimension travel together)
  // the leading dimensions that are not decomposed: every ghost of theirs resolved to its owned root atom + total periodic
  // shift once per ghost build, so that their refresh is one kernel (as in CommSerial)
  DeviceArray<T_INT> local_root;
  DeviceArray<T_X_FLOAT> local_shift;
  int local_dims = 0;       // dims [0, local_dims) are handled by that kernel
  T_INT local_ghosts = 0;

  bool decomposed(int phase) const { return dec.grid[phase / 2] > 1; }
  void ensure_bytes(DeviceArray<char> &b, size_t bytes);
  void fail(const char *what);
  void refresh(bool defer);

public:
  CommMPI(System *s, T_X_FLOAT comm_depth_);
  ~CommMPI();
  void init();
  void create_domain_decomposition();
  void exchange();
  void exchange_halo();
  void update_halo();
  bool update_halo_deferred();
  void update_force();
  void reduce_float(T_FLOAT *values, T_INT N);
  void reduce_int(T_INT *values, T_INT N);
  void reduce_max_float(T_FLOAT *values, T_INT N);
  void reduce_max_int(T_INT *values, T_INT N);
  void reduce_min_float(T_FLOAT *values, T_INT N);
  void reduce_min_int(T_INT *values, T_INT N);
  void scan_int(T_INT *values, T_INT N);
  void weighted_reduce_float(T_FLOAT *values, T_INT *weight, T_INT N);
  int process_rank();
  int num_processes();
  void error(const char *msg);
  const char *name();
  const emd_decomp &decomposition() const { return dec; }
  const T_INT *ghost_counts() const { return num_ghost; }
};
#endif
#endif
