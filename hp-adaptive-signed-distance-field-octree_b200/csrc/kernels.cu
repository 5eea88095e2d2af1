// kernels.cu — the single device translation unit of libhpsdf: constant tables, all kernels, and their launchers.
// Built for sm_100a only (B200): nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>
#include "hp_common.h"
#include "device_ctx.h"

namespace hpsdf
{
    __constant__ double c_nl[kMaxDegree + 1][kMaxDepth + 1];   // NormalisedLengths (Utility.h:63-78)
    __constant__ double c_rec[kMaxDegree + 1][2];              // LegendreCoefficent (Utility.h:112-127)
}

#include "mesh_eval.cuh"
#include "query_eval.cuh"
#include "sdf_eval.cuh"
#include "fit_kernels.cuh"
#include "query_kernels.cuh"

namespace hpsdf
{
    void uploadConstants()
    {
        cudaMemcpyToSymbol(c_nl, tables().nl, sizeof(c_nl));
        cudaMemcpyToSymbol(c_rec, tables().rec, sizeof(c_rec));
    }

    cudaError_t launchQuery(const DeviceTreeView& view, const double* dXyz, size_t n, double* dOut, int smCount, cudaStream_t stream)
    {
        if (!n) return cudaSuccess;
        const size_t tiles = (n + kQueryThreads - 1) / kQueryThreads;
        // persistent grid: 8 CTAs of 256 threads per SM (22 KB shared memory each), fewer if the batch is small
        size_t grid = (size_t)(smCount > 0 ? smCount : 148) * 8;
        if (grid > tiles) grid = tiles;
        queryKernel<<<(unsigned)grid, kQueryThreads, 0, stream>>>(view, dXyz, n, dOut);
        return cudaGetLastError();
    }

    static const uint32_t* g_bidxDev[16] = { nullptr };
    void setBidxDev(int device, const uint32_t* p) { g_bidxDev[device & 15] = p; }

    cudaError_t launchQueryGradient(const DeviceTreeView& view, const double* dXyz, size_t n, double* dOut, double* dGrad, cudaStream_t stream)
    {
        if (!n) return cudaSuccess;
        int dev = 0;
        cudaGetDevice(&dev);
        queryGradientKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(view, dXyz, n, dOut, dGrad, g_bidxDev[dev & 15]);
        return cudaGetLastError();
    }

    cudaError_t launchGatherSegments(const double* src, double* dst, const uint32_t* srcOff, const uint32_t* dstOff,
                                     const uint32_t* count, uint32_t nSeg, cudaStream_t stream)
    {
        if (!nSeg) return cudaSuccess;
        const unsigned blocks = (unsigned)(((size_t)nSeg * 32 + 255) / 256);
        gatherSegmentsKernel<<<blocks, 256, 0, stream>>>(src, dst, srcOff, dstOff, count, nSeg);
        return cudaGetLastError();
    }
}
