// Include Module header files for force
#include "force_types/force_lj_neigh.h"
#include "force_types/force_lj_idial_neigh.h"
#include "force_types/force_snap_neigh.h"
