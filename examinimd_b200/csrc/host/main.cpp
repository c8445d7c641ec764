// main.cpp -- thin driver (role of src/main.cpp:48-72): one process, one GPU.
// Multi-GPU runs start one process per GPU (torchrun or any launcher that sets
// RANK/WORLD_SIZE/LOCAL_RANK); see comm_types/comm_nccl.h.
#include "examinimd.h"
#include <cstdlib>

int main(int argc, char *argv[]) {
  setenv("CUDA_MODULE_LOADING", "EAGER", 0); // load every kernel with its module, not at its first launch inside the run loop
  int device = 0;
  if (const char *lr = getenv("LOCAL_RANK")) device = atoi(lr);
  ExaMiniMD examinimd(device);
  examinimd.init(argc, argv);
  examinimd.run(examinimd.input->nsteps);
  examinimd.print_performance();
  examinimd.shutdown();
  return 0;
}
