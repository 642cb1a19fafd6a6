// _cavitymd: same module name and class name as the reference (reference src/cavitymd/module.cc:27-34).
// This build exports the GPU class only; the CPU class CavityForceCompute stays the reference's own
// file when a CPU device is wanted (plugin/CMakeLists.txt option CAVB200_WITH_REFERENCE_CPU).
#include "CavityForceComputeGPU.h"
#include "TwoStepConstantVolumeCavity.h"

PYBIND11_MODULE(_cavitymd, m)
    {
    hoomd::cavitymd::detail::export_CavityForceComputeGPU(m);
    // the fused integration method: new functionality, not in the reference's module
    hoomd::md::detail::export_TwoStepConstantVolumeCavity(m);
    }
