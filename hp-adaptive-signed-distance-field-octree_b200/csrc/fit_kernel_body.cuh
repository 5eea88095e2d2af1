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
// This header is also compiled at run time by NVRTC (jit.cpp) with HPSDF_JIT_PROGRAM defined: there the SDF program is not
// interpreted but generated as straight-line code (sdfEval defined by the generated translation unit before this
// include). It therefore includes nothing that needs system headers.
#include "hp_common.h"
#ifdef HPSDF_JIT_PROGRAM
namespace hpsdf
{
    struct SdfProgramSmem { int unused; };
    __device__ __forceinline__ void stageProgram(SdfProgramSmem&, const SdfProgramDev&) {}
}
#else
#include "sdf_eval.cuh"
#endif

namespace hpsdf
{
    template <int D, bool EXT>
    __global__ void __launch_bounds__(fitThreads(D))
    fitKernel(const FitTask* __restrict__ tasks, double* __restrict__ pool, FitRecord* __restrict__ recs,
              const SdfProgramDev prog, const RootMap map, const FitTablesDev tab, const double* __restrict__ samples)
    {
        constexpr int N  = fitRule(D);
        constexpr int N2 = N * N;
        constexpr int P2 = pairCount(D);
        extern __shared__ double smem[];
        __shared__ SdfProgramSmem sProg;
        if constexpr (!EXT) stageProgram(sProg, prog);
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
            sZ[k] = samplePos(r, half, (double)t.cz, map.sizes[2], map.centre[2]);
        }
        __syncthreads();

        // ---- stage 1: sample F and contract z ---------------------------------------------------------------------
        for (int col = tid; col < N2; col += blockDim.x)
        {
            double acc[D + 1];
            #pragma unroll
            for (int c = 0; c <= D; ++c) acc[c] = 0.0;
            if constexpr (EXT)
            {
                // mesh / octree programs: F was sampled by sampleKernel into samples[fit][k][j][i] (coalesced over col)
                const double* __restrict__ fs = samples + (size_t)blockIdx.x * (N * N2) + col;
                #pragma unroll 4
                for (int k = 0; k < N; ++k)
                {
                    const double f = fs[(size_t)k * N2];
                    #pragma unroll
                    for (int c = 0; c <= D; ++c) acc[c] = fma(f, sQ[c * N + k], acc[c]);
                }
            }
            else
            {
                const int i = col % N, j = col / N;
                const double X = samplePos(sR[i], half, (double)t.cx, map.sizes[0], map.centre[0]);
                const double Y = samplePos(sR[j], half, (double)t.cy, map.sizes[1], map.centre[1]);
                #pragma unroll 1
                for (int k = 0; k < N; ++k)
                {
                    const double f = sdfEval<false>(sProg, X, Y, sZ[k]);
                    #pragma unroll
                    for (int c = 0; c <= D; ++c) acc[c] = fma(f, sQ[c * N + k], acc[c]);
                }
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
}
