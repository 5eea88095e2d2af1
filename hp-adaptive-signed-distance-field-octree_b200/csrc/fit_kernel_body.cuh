// fit_kernel_body.cuh — batched FitPolynomial (Source/HP/Octree.cpp:1007-1093) for every fit of a build round.
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
// followed by the top-shell energy (:1062-1069). One CTA per fit (degrees 1-3: 5 / 3 / 3 fits per CTA, hp_common.h:
// fitGroup); closed-form SDF programs are evaluated in place (sdf_eval.cuh, or generated code under NVRTC), so samples
// never touch HBM: per fit the kernel reads a 32-byte task and writes N_d coefficients + a 16-byte record. Mesh / octree
// programs (EXT) read F from the scratch buffer a sample kernel filled (fit_kernels.cuh, mesh_sample_kernel.cuh).
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
              const SdfProgramDev prog, const RootMap map, const FitTablesDev tab, const double* __restrict__ samples, const int nFits)
    {
        constexpr int N  = fitRule(D);
        constexpr int N2 = N * N;
        constexpr int P2 = pairCount(D);
        constexpr int G  = fitGroup(D);                     // fits per CTA; fit g of the group is task blockIdx.x * G + g
        constexpr int PF = (int)fitSmemPerFit(D);
        extern __shared__ double smem[];
        __shared__ SdfProgramSmem sProg;
        if constexpr (!EXT) stageProgram(sProg, prog);
        double* sQ   = smem;                   // Q[c][k]
        double* sR   = sQ + (D + 1) * N;       // roots
        double* sFit = sR + N;                 // per fit: z of sample k (n) | T1[c][j][i] | T2[q][i]; coefficients alias T1
        auto sZ  = [&](int g) { return sFit + g * PF; };
        auto sT1 = [&](int g) { return sFit + g * PF + N; };
        auto sT2 = [&](int g) { return sFit + g * PF + N + (D + 1) * N2; };

        const int tid = threadIdx.x;
        const int first = blockIdx.x * G;
        const int nHere = nFits - first < G ? nFits - first : G;

        for (int e = tid; e < (D + 1) * N; e += blockDim.x) sQ[e] = tab.q[D][e];
        for (int k = tid; k < N; k += blockDim.x) sR[k] = tab.roots[D][k];
        for (int e = tid; e < nHere * N; e += blockDim.x)
        {
            const int g = e / N, k = e - g * N;
            const FitTask& t = tasks[first + g];
            sZ(g)[k] = samplePos(tab.roots[D][k], (double)t.half, (double)t.cz, map.sizes[2], map.centre[2]);
        }
        __syncthreads();

        // ---- stage 1: sample F and contract z ---------------------------------------------------------------------
        for (int item = tid; item < nHere * N2; item += blockDim.x)
        {
            const int g = item / N2, col = item - g * N2;
            double acc[D + 1];
            #pragma unroll
            for (int c = 0; c <= D; ++c) acc[c] = 0.0;
            if constexpr (EXT)
            {
                // mesh / octree programs: F was sampled by sampleKernel into samples[fit][k][j][i] (coalesced over col)
                const double* __restrict__ fs = samples + (size_t)(first + g) * (N * N2) + col;
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
                const FitTask& t = tasks[first + g];
                const double half = (double)t.half;    // aabbScale = sizes * 0.5 (Octree.cpp:1020); cells are cubes
                const int i = col % N, j = col / N;
                const double X = samplePos(sR[i], half, (double)t.cx, map.sizes[0], map.centre[0]);
                const double Y = samplePos(sR[j], half, (double)t.cy, map.sizes[1], map.centre[1]);
                const double* __restrict__ z = sZ(g);
                #pragma unroll 1
                for (int k = 0; k < N; ++k)
                {
                    const double f = sdfEval<false>(sProg, X, Y, z[k]);
                    #pragma unroll
                    for (int c = 0; c <= D; ++c) acc[c] = fma(f, sQ[c * N + k], acc[c]);
                }
            }
            double* t1 = sT1(g);
            #pragma unroll
            for (int c = 0; c <= D; ++c) t1[c * N2 + col] = acc[c];
        }
        __syncthreads();

        // ---- stage 2: contract y --------------------------------------------------------------------------------
        for (int item = tid; item < nHere * (P2 * N); item += blockDim.x)
        {
            const int g = item / (P2 * N), o = item - g * (P2 * N);
            const int i = o % N, q = o / N;
            // q -> (b, c): pairs enumerated b = 0..D, c = 0..D-b
            int b = 0, rem = q;
            while (rem >= D + 1 - b) { rem -= D + 1 - b; ++b; }
            const int c = rem;
            const double* t1 = sT1(g) + c * N2 + i;
            const double* qb = sQ + b * N;
            double s = 0.0;
            #pragma unroll 4
            for (int j = 0; j < N; ++j) s = fma(t1[j * N], qb[j], s);
            sT2(g)[q * N + i] = s;
        }
        __syncthreads();

        // ---- stage 3: contract x for the wanted indices, scale, write ----------------------------------------------
        constexpr int end = coeffCount(D);
        for (int item = tid; item < nHere * end; item += blockDim.x)
        {
            const int g = item / end, idx = item - g * end;
            const FitTask& t = tasks[first + g];
            const int start = t.degreeIn > 0 ? coeffCount(t.degreeIn) : 0;   // Octree.cpp:1012-1013
            double v;
            if (idx >= start)
            {
                const double half = (double)t.half;
                const double V = half * half * half;                         // aabbScale.prod() (Octree.cpp:1022)
                const uint32_t abc = tab.bidx[idx];
                const int a = abc & 0xFF, b = (abc >> 8) & 0xFF, c = (abc >> 16) & 0xFF;
                const int q = b * (D + 1) - (b * (b - 1)) / 2 + c;
                const double* t2 = sT2(g) + q * N;
                const double* qa = sQ + a * N;
                double s = 0.0;
                #pragma unroll 4
                for (int i = 0; i < N; ++i) s = fma(t2[i], qa[i], s);
                v = s * (V * (c_nl[a][t.depth] * c_nl[b][t.depth] * c_nl[c][t.depth]));
            }
            else v = pool[t.src + idx];                                      // kept lower shells (Octree.cpp:847)
            pool[t.out + idx] = v;
            sT1(g)[idx] = v;                                                 // T1 is dead after stage 2
        }
        __syncthreads();

        // ---- top-shell energy (Octree.cpp:1062-1069): sum of c^2 over idx < end with a+b+c == D; one warp per fit -------
        for (int g = tid >> 5; g < nHere; g += blockDim.x >> 5)
        {
            const FitTask& t = tasks[first + g];
            const int start = t.degreeIn > 0 ? coeffCount(t.degreeIn) : 0;
            const double* sC = sT1(g);
            double e = 0.0;
            for (int idx = start + (tid & 31); idx < end; idx += 32)
            {
                const uint32_t abc = tab.bidx[idx];
                if ((int)((abc & 0xFF) + ((abc >> 8) & 0xFF) + ((abc >> 16) & 0xFF)) == D) e = fma(sC[idx], sC[idx], e);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
            if ((tid & 31) == 0) { FitRecord r; r.rawErr = e; r.c0 = sC[0]; recs[t.rec] = r; }
        }
    }
}
