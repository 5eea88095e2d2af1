// comm.h — one rank of a multi-GPU build: NCCL communicator loaded at run time (dlopen), so the library itself has no
// link-time dependency on NCCL and loads on a CPU-only box.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "hp_common.h"

namespace hpsdf
{
    struct CommSegment
    {
        double* ptr;      // device pointer (same address on every rank: pools are replicated)
        size_t  count;    // doubles
        int     root;     // rank that produced it
    };

    int          commRank(const hpsdf_comm* c);
    int          commWorld(const hpsdf_comm* c);
    // ncclGroupStart; one ncclBroadcast per segment; ncclGroupEnd — an all-gather with ragged, in-place segments.
    hpsdf_status commBroadcastSegments(hpsdf_comm* c, const std::vector<CommSegment>& segs, cudaStream_t stream);
}
