// NeighborCSRMapConstr -- the reference builds the identical CSR list through a hash-map pair
// set (src/neighbor_types/neighbor_csr_map_constr.h:128-321); the result does not depend on
// the construction, so the flag is accepted and served by the same kernels as NeighborCSR.
#ifdef MODULES_OPTION_CHECK
      if ((strcmp(argv[i + 1], "CSR_MAPCONSTR") == 0)) neighbor_type = NEIGH_CSR_MAPCONSTR;
#endif
#ifdef NEIGHBOR_MODULES_INSTANTIATION
    else if (input->neighbor_type == NEIGH_CSR_MAPCONSTR) {
      neighbor = new NeighborCSRMapConstr();
      neighbor->init(input->force_cutoff + input->neighbor_skin);
    }
#endif
#if !defined(MODULES_OPTION_CHECK) && !defined(NEIGHBOR_MODULES_INSTANTIATION)
#ifndef NEIGHBOR_CSR_MAP_CONSTR_H
#define NEIGHBOR_CSR_MAP_CONSTR_H
#include "neighbor_csr.h"

class NeighborCSRMapConstr : public NeighborCSR {
public:
  NeighborCSRMapConstr() { neigh_type = NEIGH_CSR_MAPCONSTR; }
  const char *name() { return "NeighborCSRMapConstr"; }
};

template <>
struct NeighborAdaptor<NEIGH_CSR_MAPCONSTR> { typedef NeighborCSRMapConstr type; };
#endif
#endif
