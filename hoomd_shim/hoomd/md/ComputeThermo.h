// hoomd_shim/hoomd/md/ComputeThermo.h -- forwards to the single shim header (see ../ShimCore.h).
#include "../ShimCore.h"
