// continuity.cpp — PerformContinuityPostProcess (Source/HP/Octree.cpp:1717-1762) on the device. Placeholder until the
// face-pair assembly and CG kernels land (task 5).
#include "octree.h"

namespace hpsdf
{
    hpsdf_status continuityPostProcess(hpsdf_octree&, const hpsdf_build_opts&, cudaStream_t)
    {
        setLastError("continuity post-process is not available in this build");
        return HPSDF_ERR_UNSUPPORTED;
    }
}
