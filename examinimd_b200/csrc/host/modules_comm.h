// Include Module header files for comm
#include "comm_types/comm_mpi.h"
#include "comm_types/comm_serial.h"
