// hoomd_shim/hoomd/Autotuner.h -- forwards to the single shim header (see ShimCore.h for what this is).
#include "ShimCore.h"
