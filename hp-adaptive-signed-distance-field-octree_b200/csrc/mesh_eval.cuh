// mesh_eval.cuh — float32 mesh signed distance on the device (Mesh::SignedDistanceAtPt, Source/Meshing/Mesh.cpp:54-63).
#pragma once
#include "hp_common.h"

namespace hpsdf
{
    struct DeviceMeshView
    {
        const float*    vertices;     // nVertices x 3
        const uint32_t* tris;         // nTris x 3
        uint32_t        nVertices, nTris;
    };

    // TODO(round 1, task 6): BVH closest-triangle + pseudonormal sign; until then MESH programs are rejected by the host
    // (HPSDF_ERR_UNSUPPORTED), so this is never reached.
    __device__ __noinline__ double meshSignedDistance(const DeviceMeshView*, double, double, double) { return 0.0; }
}
