// Include Module header files for binning
#include "binning_types/binning_kksort.h"
