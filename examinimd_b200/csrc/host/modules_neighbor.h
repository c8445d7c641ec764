// Include Module header files for neighbor
#include "neighbor_types/neighbor_2d.h"
#include "neighbor_types/neighbor_csr.h"
#include "neighbor_types/neighbor_csr_map_constr.h"
