#include "force_snap_neigh.h"
#include <cstdio>
#include <cstdlib>

struct ForceSNAP::Impl {};

ForceSNAP::ForceSNAP(char **args, System *system, bool half_neigh_) : Force(args, system, half_neigh_), sys(system), impl(nullptr) {}
ForceSNAP::~ForceSNAP() { delete impl; }
void ForceSNAP::init_coeff(int, char **) {
  fprintf(stderr, "ForceSNAP: CUDA kernel not built yet\n");
  exit(1);
}
void ForceSNAP::compute(System *, Binning *, Neighbor *) {}
const char *ForceSNAP::name() { return "ForceSNAP"; }
