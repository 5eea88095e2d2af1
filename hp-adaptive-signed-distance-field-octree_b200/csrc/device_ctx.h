// device_ctx.h — per-device tables and the host-callable launchers of kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "hp_common.h"

namespace hpsdf
{
    // Per-degree projection tables in device memory: q[d][c*n + k] = w_k * P_c(xi_k) for the (4d+1)-point rule,
    // roots[d][k] = xi_k; bidx[idx] = a | b << 8 | c << 16 (BasisIndexValues, Utility.h:133-160).
    struct FitTablesDev
    {
        const double*   q[kMaxDegree + 1];
        const double*   roots[kMaxDegree + 1];
        const uint32_t* bidx;
    };

    struct DeviceCtx
    {
        int          device = -1;
        FitTablesDev fitTab{};
        void*        tabMem = nullptr;
        int          smCount = 0;
    };

    // Lazily creates the context of `device` (uploads tables and constant memory). nullptr + err on failure.
    DeviceCtx* getDeviceCtx(int device, std::string& err);

    // kernels.cu
    void        uploadConstants();
    cudaError_t launchFitKernel(int degree, const FitTask* dTasks, int n, double* pool, FitRecord* recs,
                                const SdfProgramDev& prog, const RootMap& map, const FitTablesDev& tab, cudaStream_t stream);
    cudaError_t launchSdfEval(const SdfProgramDev& prog, const double* dXyz, size_t n, double* dOut, cudaStream_t stream);
    cudaError_t launchDfmaPeak(double* dOut, int blocks, cudaStream_t stream);
    cudaError_t launchQuery(const DeviceTreeView& view, const double* dXyz, size_t n, double* dOut, int smCount, cudaStream_t stream);
    cudaError_t launchQueryGradient(const DeviceTreeView& view, const double* dXyz, size_t n, double* dOut, double* dGrad, cudaStream_t stream);
    // dst[dstOff[s] + i] = src[srcOff[s] + i], i < count[s], for nSeg segments (ReallocCoeffs, Octree.cpp:474-555, on device)
    cudaError_t launchGatherSegments(const double* src, double* dst, const uint32_t* srcOff, const uint32_t* dstOff,
                                     const uint32_t* count, uint32_t nSeg, cudaStream_t stream);
}
