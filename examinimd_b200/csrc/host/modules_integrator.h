// Include Module header files for integrator
#include "integrator_nve.h"
