// CommNCCL -- 3-D brick domain decomposition across the GPUs of one NVSwitch box; takes the
// place of the reference's CommMPI (src/comm_types/comm_mpi.{h,cpp}) and answers to the same
// `--comm-type MPI` flag.  Implemented in comm_nccl.cpp (stub fragments until it lands).
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "MPI") == 0) || (strcmp(argv[i + 1], "NCCL") == 0)) comm_type = COMM_MPI;
#endif
#ifdef COMM_MODULES_INSTANTIATION
#endif
