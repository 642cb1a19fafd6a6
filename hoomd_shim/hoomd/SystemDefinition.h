// hoomd_shim/hoomd/SystemDefinition.h -- forwards to the single shim header (see ShimCore.h for what this is).
#include "ShimCore.h"
