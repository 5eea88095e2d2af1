// fit_kernels.cuh — host launchers of fitKernel<D, EXT> (the kernel itself: fit_kernel_body.cuh).
#pragma once
#include <cuda_runtime.h>
#include "hp_common.h"
#include "device_ctx.h"
#include "fit_kernel_body.cuh"

namespace hpsdf
{
    template <int D, bool EXT>
    static cudaError_t launchOne(const FitTask* dTasks, int n, double* pool, FitRecord* recs, const SdfProgramDev& prog,
                                 const RootMap& map, const FitTablesDev& tab, cudaStream_t stream)
    {
        constexpr size_t smem = fitSmemDoubles(D) * sizeof(double);
        static bool attrSet[16] = { false };
        int dev = 0;
        cudaGetDevice(&dev);
        if (!attrSet[dev & 15])
        {
            cudaError_t e = cudaFuncSetAttribute(fitKernel<D, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            attrSet[dev & 15] = true;
        }
        fitKernel<D, EXT><<<n, fitThreads(D), smem, stream>>>(dTasks, pool, recs, prog, map, tab);
        return cudaGetLastError();
    }

    static bool programHasExt(const SdfProgramDev& prog)
    {
        for (uint32_t i = 0; i < prog.n; ++i)
            if (prog.instr[i].op == HPSDF_PRIM_MESH || prog.instr[i].op == HPSDF_PRIM_OCTREE) return true;
        return false;
    }

    template <bool EXT>
    static cudaError_t launchFitKernelT(int degree, const FitTask* dTasks, int n, double* pool, FitRecord* recs,
                                const SdfProgramDev& prog, const RootMap& map, const FitTablesDev& tab, cudaStream_t stream)
    {
        if (n <= 0) return cudaSuccess;
        switch (degree)
        {
            case 1:  return launchOne<1, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 2:  return launchOne<2, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 3:  return launchOne<3, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 4:  return launchOne<4, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 5:  return launchOne<5, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 6:  return launchOne<6, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 7:  return launchOne<7, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 8:  return launchOne<8, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 9:  return launchOne<9, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 10: return launchOne<10, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            case 11: return launchOne<11, EXT>(dTasks, n, pool, recs, prog, map, tab, stream);
            default: return cudaErrorInvalidValue;
        }
    }

    // Launch the fits of one degree. dTasks: n tasks, all with task.degree == degree.
    cudaError_t launchFitKernel(int degree, const FitTask* dTasks, int n, double* pool, FitRecord* recs,
                                const SdfProgramDev& prog, const RootMap& map, const FitTablesDev& tab, cudaStream_t stream)
    {
        return programHasExt(prog) ? launchFitKernelT<true>(degree, dTasks, n, pool, recs, prog, map, tab, stream)
                                   : launchFitKernelT<false>(degree, dTasks, n, pool, recs, prog, map, tab, stream);
    }

    // ---- program evaluation at arbitrary points (hpsdf_sdf_eval) and the FP64 peak probe ---------------------------
    __global__ void sdfEvalKernel(const SdfProgramDev prog, const double* __restrict__ xyz, size_t n, double* __restrict__ out)
    {
        __shared__ SdfProgramSmem sProg;
        stageProgram(sProg, prog);
        __syncthreads();
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) out[i] = sdfEval<true>(sProg, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }

    cudaError_t launchSdfEval(const SdfProgramDev& prog, const double* dXyz, size_t n, double* dOut, cudaStream_t stream)
    {
        if (!n) return cudaSuccess;
        sdfEvalKernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(prog, dXyz, n, dOut);
        return cudaGetLastError();
    }

    // 8 independent DFMA chains per thread, 4096 iterations: 2 * 8 * 4096 flops per thread.
    __global__ void __launch_bounds__(256) dfmaPeakKernel(double* out, double a, double b)
    {
        double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
        #pragma unroll 1
        for (int i = 0; i < 512; ++i)
        {
            #pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
                x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
            }
        }
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    }

    cudaError_t launchDfmaPeak(double* dOut, int blocks, cudaStream_t stream)
    {
        dfmaPeakKernel<<<blocks, 256, 0, stream>>>(dOut, 0.999999, 1e-9);
        return cudaGetLastError();
    }
}
