// fit_kernels.cuh — batched FitPolynomial (Source/HP/Octree.cpp:1007-1093) for every fit of a build round.
//
// Reference loop: for each of the n^3 Gauss-Legendre samples (n = 4d+1 per axis) and each coefficient index,
// coeffs[idx] += prod_axis(LpX(a_axis, xi_axis) * NL[a_axis][depth]) * (V * w_i w_j w_k * F(x))        (:1028-1056)
// i.e. n^3 * N_d * 3 Legendre recurrences. Here the same tensor contraction is sum-factorised:
//
//   stage 1 (fused with SDF sampling, registers only): one thread per (i, j) column walks k and accumulates the d+1
//           z-moments  T1[c][j][i] = sum_k F(x_i, y_j, z_k) * Q[c][k],     Q[c][k] = w_k * P_c(xi_k)
//   stage 2 (shared memory): T2[(b,c)][i] = sum_j T1[c][j][i] * Q[b][j]      for b + c <= d
//   stage 3: C[a][b][c] = V * NL[a] NL[b] NL[c] * sum_i T2[(b,c)][i] * Q[a][i]   for the wanted indices [start, end)
//
// followed by the top-shell energy (:1062-1069). One CTA per fit; the SDF program is evaluated in place (sdf_eval.cuh),
// so samples never touch HBM: per fit the kernel reads a 32-byte task and writes N_d coefficients + a 16-byte record.
// The nearness weight (:1071-1090) and the h-vs-p decision (:558-659) are host-side in the greedy replay (build.cpp),
// where they use the same libm as the CPU checker.
//
// FP64 on CUDA cores: the contraction has M = d+1 in 3..12 against K = n in 9..45 — far too thin for the DMMA shapes —
// and SDF evaluation (sqrt/div chains) dominates the instruction count, so there is no tensor-core path here; tcgen05
// has no FP64 kind at all.
#pragma once
#include <cuda_runtime.h>
#include "hp_common.h"
#include "device_ctx.h"
#include "sdf_eval.cuh"

namespace hpsdf
{
    HPSDF_HD constexpr int fitPasses(int d)  { return (fitRule(d) * fitRule(d) + 639) / 640; }
    HPSDF_HD constexpr int fitThreads(int d)
    {
        return (((fitRule(d) * fitRule(d) + fitPasses(d) - 1) / fitPasses(d)) + 31) / 32 * 32;
    }
    // shared memory (doubles): Q (d+1)*n | roots n | user-space z n | T1 (d+1)*n*n | T2 pairCount*n ; coefficients alias T1
    HPSDF_HD constexpr size_t fitSmemDoubles(int d)
    {
        return (size_t)(d + 1) * fitRule(d) + 2 * fitRule(d) + (size_t)(d + 1) * fitRule(d) * fitRule(d) + (size_t)pairCount(d) * fitRule(d);
    }

    template <int D, bool EXT>
    __global__ void __launch_bounds__(fitThreads(D))
    fitKernel(const FitTask* __restrict__ tasks, double* __restrict__ pool, FitRecord* __restrict__ recs,
              const SdfProgramDev prog, const RootMap map, const FitTablesDev tab)
    {
        constexpr int N  = fitRule(D);
        constexpr int N2 = N * N;
        constexpr int P2 = pairCount(D);
        extern __shared__ double smem[];
        __shared__ SdfProgramSmem sProg;
        stageProgram(sProg, prog);
        double* sQ  = smem;                    // Q[c][k]
        double* sR  = sQ + (D + 1) * N;        // roots
        double* sZ  = sR + N;                  // user-space z of sample k
        double* sT1 = sZ + N;                  // T1[c][j][i]
        double* sT2 = sT1 + (D + 1) * N2;      // T2[q][i]
        double* sC  = sT1;                     // final coefficients (T1 is dead after stage 2)

        const FitTask t = tasks[blockIdx.x];
        const int tid = threadIdx.x;
        const double half = (double)t.half;    // aabbScale = sizes * 0.5 (Octree.cpp:1020); cells are cubes

        for (int e = tid; e < (D + 1) * N; e += blockDim.x) sQ[e] = tab.q[D][e];
        for (int k = tid; k < N; k += blockDim.x)
        {
            const double r = tab.roots[D][k];
            sR[k] = r;
            sZ[k] = (r * half + (double)t.cz) * map.sizes[2] + map.centre[2];       // Octree.cpp:1039 then :327
        }
        __syncthreads();

        // ---- stage 1: sample F and contract z ---------------------------------------------------------------------
        for (int col = tid; col < N2; col += blockDim.x)
        {
            const int i = col % N, j = col / N;
            const double X = (sR[i] * half + (double)t.cx) * map.sizes[0] + map.centre[0];
            const double Y = (sR[j] * half + (double)t.cy) * map.sizes[1] + map.centre[1];
            double acc[D + 1];
            #pragma unroll
            for (int c = 0; c <= D; ++c) acc[c] = 0.0;
            #pragma unroll 1
            for (int k = 0; k < N; ++k)
            {
                const double f = sdfEval<EXT>(sProg, X, Y, sZ[k]);
                #pragma unroll
                for (int c = 0; c <= D; ++c) acc[c] = fma(f, sQ[c * N + k], acc[c]);
            }
            #pragma unroll
            for (int c = 0; c <= D; ++c) sT1[c * N2 + col] = acc[c];
        }
        __syncthreads();

        // ---- stage 2: contract y --------------------------------------------------------------------------------
        for (int o = tid; o < P2 * N; o += blockDim.x)
        {
            const int i = o % N, q = o / N;
            // q -> (b, c): pairs enumerated b = 0..D, c = 0..D-b
            int b = 0, rem = q;
            while (rem >= D + 1 - b) { rem -= D + 1 - b; ++b; }
            const int c = rem;
            const double* t1 = sT1 + c * N2 + i;
            const double* qb = sQ + b * N;
            double s = 0.0;
            #pragma unroll 4
            for (int j = 0; j < N; ++j) s = fma(t1[j * N], qb[j], s);
            sT2[q * N + i] = s;
        }
        __syncthreads();

        // ---- stage 3: contract x for the wanted indices, scale, write ----------------------------------------------
        const int start = t.degreeIn > 0 ? coeffCount(t.degreeIn) : 0;       // Octree.cpp:1012-1013
        const int end   = coeffCount(D);
        const double V  = half * half * half;                                // aabbScale.prod() (Octree.cpp:1022)
        for (int idx = tid; idx < end; idx += blockDim.x)
        {
            double v;
            if (idx >= start)
            {
                const uint32_t abc = tab.bidx[idx];
                const int a = abc & 0xFF, b = (abc >> 8) & 0xFF, c = (abc >> 16) & 0xFF;
                const int q = b * (D + 1) - (b * (b - 1)) / 2 + c;
                const double* t2 = sT2 + q * N;
                const double* qa = sQ + a * N;
                double s = 0.0;
                #pragma unroll 4
                for (int i = 0; i < N; ++i) s = fma(t2[i], qa[i], s);
                v = s * (V * (c_nl[a][t.depth] * c_nl[b][t.depth] * c_nl[c][t.depth]));
            }
            else v = pool[t.src + idx];                                      // kept lower shells (Octree.cpp:847)
            pool[t.out + idx] = v;
            sC[idx] = v;
        }
        __syncthreads();

        // ---- top-shell energy (Octree.cpp:1062-1069): sum of c^2 over idx < end with a+b+c == D --------------------
        if (tid < 32)
        {
            double e = 0.0;
            for (int idx = start + tid; idx < end; idx += 32)
            {
                const uint32_t abc = tab.bidx[idx];
                if ((int)((abc & 0xFF) + ((abc >> 8) & 0xFF) + ((abc >> 16) & 0xFF)) == D) e = fma(sC[idx], sC[idx], e);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
            if (tid == 0) { FitRecord r; r.rawErr = e; r.c0 = sC[0]; recs[t.rec] = r; }
        }
    }

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
