// continuity_kernels.cuh — PerformContinuityPostProcess (Source/HP/Octree.cpp:1717-1762) on the device.
//
// Reference: every leaf-leaf shared face contributes the Gram matrix of the jump (f_A - f_B) to a sparse matrix M
// (EvaluateSharedFaceIntegralAnalytically :1459-1546 for equal depths, ...Numerically :1250-1456 otherwise), lambda is
// added on the diagonal, and (M + lambda I) x = lambda c is solved with Eigen's ConjugateGradient (:1751-1755).
//
// Here: (1) one CTA per face emits that face's entries as (row << 32 | col, value) pairs at a precomputed offset —
// positions inside a face come from an ordered block-wide compaction, so the COO order (and hence the order in which
// duplicates are summed) is deterministic; (2) a radix sort + segmented reduction (CUB: plumbing) turns COO into CSR
// with duplicates summed, which is what setFromTriplets does (:1732-1735); (3) a hand-written persistent cooperative
// conjugate-gradient kernel (sparse matrix-vector product, fused dot products, grid-wide syncs) runs Eigen's CG loop
// with a diagonal preconditioner until |r|^2 < tol^2 |b|^2.
#pragma once
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include "hp_common.h"
#include "device_ctx.h"

namespace hpsdf
{
    namespace cg = cooperative_groups;

    __constant__ double c_lp1[kMaxDegree + 1];    // LpX(a, +1) by the reference's recurrence (Octree.cpp:988-1004)
    __constant__ double c_lm1[kMaxDegree + 1];    // LpX(a, -1)

    constexpr int kFaceThreads = 128;

    // Ordered compaction: every thread passes (keep, key, val) for candidate number `cand` of a chunk of blockDim
    // candidates; kept entries are written at base + (number of kept candidates before this one).
    __device__ __forceinline__ void emitOrdered(bool keep, uint64_t key, double val, uint64_t* __restrict__ keys,
                                                double* __restrict__ vals, unsigned long long& base, uint32_t* sWarp)
    {
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) sWarp[warp] = __popc(ballot);
        __syncthreads();
        uint32_t before = 0, total = 0;
        #pragma unroll
        for (int w = 0; w < kFaceThreads / 32; ++w) { const uint32_t c = sWarp[w]; if (w < (int)warp) before += c; total += c; }
        if (keep)
        {
            const unsigned long long pos = base + before + __popc(ballot & ((1u << lane) - 1u));
            keys[pos] = key; vals[pos] = val;
        }
        base += total;
        __syncthreads();
    }

    // One CTA per face. Analytic faces (equal depth) emit only the structurally non-zero entries (tangential indices
    // equal); numeric faces emit their three dense blocks with sub-threshold values written as explicit zeros, so the
    // per-face entry count is known on the host without evaluating anything.
    __global__ void __launch_bounds__(kFaceThreads)
    faceEmitKernel(const FaceJobDev* __restrict__ faces, uint32_t nFaces, const uint32_t* __restrict__ bidx,
                   const double* __restrict__ glRoots, const double* __restrict__ glWeights,
                   uint64_t* __restrict__ keys, double* __restrict__ vals)
    {
        __shared__ uint32_t sWarp[kFaceThreads / 32];
        __shared__ double sL[2][3][kMaxDegree + 1][kMaxDegree + 1];     // [side A/B][axis][degree][node]: LpX at the sample coordinates
        const FaceJobDev f = faces[blockIdx.x];
        const int dim = f.dim, t1 = (dim + 1) % 3, t2 = (dim + 2) % 3;
        const int nA = coeffCount(f.degA), nB = coeffCount(f.degB);
        unsigned long long base = f.cooOffset;

        if (f.analytic)
        {
            // blocks in the reference's order: AA, AB (+ its transpose), BB  (Octree.cpp:1473-1545)
            for (int blk = 0; blk < 3; ++blk)
            {
                const int nR = blk == 2 ? nB : nA, nC = blk == 0 ? nA : nB;
                const uint32_t rowStart = blk == 2 ? f.cstartB : f.cstartA, colStart = blk == 0 ? f.cstartA : f.cstartB;
                const int depthR = blk == 2 ? f.depthB : f.depthA, depthC = blk == 0 ? f.depthA : f.depthB;
                const double* lr = blk == 2 ? c_lm1 : c_lp1;      // rows of A sit at u_dim = +1, rows of B at -1
                const double* lc = blk == 0 ? c_lp1 : c_lm1;
                const int nCand = nR * nC;
                for (int c0 = 0; c0 < nCand; c0 += kFaceThreads)
                {
                    const int cand = c0 + threadIdx.x;
                    bool keep = false; uint64_t key = 0, keyT = 0; double v = 0.0;
                    if (cand < nCand)
                    {
                        const int i = cand / nC, j = cand % nC;
                        const uint32_t bi = bidx[i], bj = bidx[j];
                        const int it1 = (bi >> (8 * t1)) & 0xFF, it2 = (bi >> (8 * t2)) & 0xFF, id = (bi >> (8 * dim)) & 0xFF;
                        const int jt1 = (bj >> (8 * t1)) & 0xFF, jt2 = (bj >> (8 * t2)) & 0xFF, jd = (bj >> (8 * dim)) & 0xFF;
                        keep = it1 == jt1 && it2 == jt2;
                        if (keep)
                        {
                            v = blk == 1 ? -1.0 : 1.0;
                            v *= lr[id]; v *= c_nl[id][depthR]; v *= lc[jd]; v *= c_nl[jd][depthC];
                            key  = ((uint64_t)(rowStart + i) << 32) | (uint64_t)(colStart + j);
                            keyT = ((uint64_t)(colStart + j) << 32) | (uint64_t)(rowStart + i);
                        }
                    }
                    emitOrdered(keep, key, v, keys, vals, base, sWarp);
                    if (blk == 1) emitOrdered(keep, keyT, v, keys, vals, base, sWarp);     // -Pr*Pl, Octree.cpp:1518-1519
                }
            }
            return;
        }

        // ---- numeric face: GL rule n = max(degA, degB) + 1 on the smaller face (Octree.cpp:1269-1290) -----------------
        const int maxDeg = f.degA > f.degB ? f.degA : f.degB;
        const int n = maxDeg + 1;
        const double* roots = glRoots + (n * (n - 1)) / 2;
        const double* wts   = glWeights + (n * (n - 1)) / 2;
        // Legendre values at the sample coordinates of each side: axis dim is +-1, the tangential axes are the GL nodes,
        // mapped into the coarser cell by u * invDist + invTr (Octree.cpp:1309-1314, 1366-1376)
        for (int e = threadIdx.x; e < 2 * 3 * n; e += kFaceThreads)
        {
            const int side = e / (3 * n), axis = (e / n) % 3, node = e % n;
            const bool mapped = side == 0 ? (f.depthB > f.depthA) : (f.depthA > f.depthB);
            double u;
            if (axis == dim) u = side == 0 ? 1.0 : -1.0;
            else
            {
                u = roots[node];
                if (mapped) u = u * f.invDist + (axis == t1 ? f.invTr1 : f.invTr2);
            }
            double m2 = 0.0, m1 = 1.0, l = 1.0;
            sL[side][axis][0][node] = 1.0;
            for (int a = 1; a <= maxDeg; ++a)
            {
                l = c_rec[a][0] * u * m1 - c_rec[a][1] * m2; m2 = m1; m1 = l;
                sL[side][axis][a][node] = l;
            }
        }
        __syncthreads();
        for (int blk = 0; blk < 3; ++blk)
        {
            const int nR = blk == 2 ? nB : nA, nC = blk == 0 ? nA : nB;
            const int sideR = blk == 2 ? 1 : 0, sideC = blk == 0 ? 0 : 1;
            const uint32_t rowStart = blk == 2 ? f.cstartB : f.cstartA, colStart = blk == 0 ? f.cstartA : f.cstartB;
            const int depthR = blk == 2 ? f.depthB : f.depthA, depthC = blk == 0 ? f.depthA : f.depthB;
            const int nCand = nR * nC;
            for (int cand = threadIdx.x; cand < nCand; cand += kFaceThreads)
            {
                const int i = cand / nC, j = cand % nC;
                const uint32_t bi = bidx[i], bj = bidx[j];
                int ia[3], ja[3];
                #pragma unroll
                for (int k = 0; k < 3; ++k) { ia[k] = (bi >> (8 * k)) & 0xFF; ja[k] = (bj >> (8 * k)) & 0xFF; }
                double integral = 0.0;
                for (int x = 0; x < n; ++x)
                    for (int y = 0; y < n; ++y)
                    {
                        double area = wts[x] * wts[y];                                          // Octree.cpp:1317-1323
                        #pragma unroll
                        for (int k = 0; k < 3; ++k)
                        {
                            const int node = k == t1 ? x : (k == t2 ? y : 0);
                            area *= sL[sideR][k][ia[k]][node];
                            area *= sL[sideC][k][ja[k]][node];
                        }
                        integral += area;
                    }
                double bw = 1.0;
                #pragma unroll
                for (int k = 0; k < 3; ++k) { bw *= c_nl[ia[k]][depthR]; bw *= c_nl[ja[k]][depthC]; }   // Octree.cpp:1327-1332
                integral *= blk == 1 ? f.faceScale * bw * -1.0 : f.faceScale * bw;
                const double v = fabsf((float)integral) > 0.000001f ? integral : 0.0;                   // EPSILON_F32 drop (:1336)
                // dense layout per block: candidate order; the AB block is followed by its transpose
                const unsigned long long pos = base + (blk == 1 ? 2ull * cand : (unsigned long long)cand);
                keys[pos] = ((uint64_t)(rowStart + i) << 32) | (uint64_t)(colStart + j);
                vals[pos] = v;
                if (blk == 1)
                {
                    keys[pos + 1] = ((uint64_t)(colStart + j) << 32) | (uint64_t)(rowStart + i);
                    vals[pos + 1] = v;
                }
            }
            base += (blk == 1 ? 2ull : 1ull) * (unsigned long long)nCand;
        }
    }

    __global__ void diagEmitKernel(uint64_t* __restrict__ keys, double* __restrict__ vals, uint32_t n, double lambda)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) { keys[i] = ((uint64_t)i << 32) | i; vals[i] = lambda; }                   // Octree.cpp:1724-1729
    }

    __global__ void flagNonZeroKernel(const double* __restrict__ v, uint32_t n, uint8_t* __restrict__ flags)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) flags[i] = v[i] != 0.0;
    }

    // rowPtr from sorted unique keys: rowPtr[r] = first entry whose row >= r
    __global__ void rowPtrKernel(const uint64_t* __restrict__ keys, uint32_t nnz, uint32_t n, uint32_t* __restrict__ rowPtr,
                                 uint32_t* __restrict__ col)
    {
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k > nnz) return;
        const uint32_t rowHere = k < nnz ? (uint32_t)(keys[k] >> 32) : n;
        const uint32_t rowPrev = k > 0 ? (uint32_t)(keys[k - 1] >> 32) : 0xFFFFFFFFu;
        if (k < nnz) col[k] = (uint32_t)(keys[k] & 0xFFFFFFFFu);
        if (k == 0) { for (uint32_t r = 0; r <= rowHere && r <= n; ++r) rowPtr[r] = 0; }
        else if (rowHere != rowPrev) { for (uint32_t r = rowPrev + 1; r <= rowHere && r <= n; ++r) rowPtr[r] = k; }
    }

    // ---- conjugate gradient ------------------------------------------------------------------------------------
    struct CgParams
    {
        const uint32_t* rowPtr; const uint32_t* col; const double* val;
        uint32_t n; uint32_t maxIt;
        double   tol;
        const double* b;          // right-hand side (lambda * c)
        double* x;                // in: initial guess, out: solution
        double* r; double* p; double* ap; double* invDiag;
        double* partial;          // 3 * gridDim.x scratch for block partial sums
        double* result;           // [0] iterations, [1] relative residual
    };

    constexpr int kCgThreads = 256;

    __device__ __forceinline__ double blockSum(double v, double* sRed)
    {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0.0;
        #pragma unroll
        for (int w = 0; w < kCgThreads / 32; ++w) s += sRed[w];
        __syncthreads();
        return s;
    }

    // Every block sums all block partials in the same fixed order (warp 0: strided partial sums, then a shuffle tree),
    // so all blocks get bit-identical scalars and take the same branches; deterministic, no second kernel.
    __device__ __forceinline__ double gridSum(const double* partial, int nBlocks, double* sRed)
    {
        if (threadIdx.x < 32)
        {
            double s = 0.0;
            for (int i = threadIdx.x; i < nBlocks; i += 32) s += partial[i];
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            if (threadIdx.x == 0) sRed[0] = s;
        }
        __syncthreads();
        const double r = sRed[0];
        __syncthreads();
        return r;
    }

    // y[row] = sum_k val[k] * x[col[k]], 8 lanes per row. A row's columns come in contiguous runs (the coefficient ranges of
    // the leaf itself and of each face neighbour), so 8 lanes reading 8 consecutive entries also gather 8 mostly consecutive
    // x values: val / col are read as full 64 / 32-byte segments and the gathers coalesce into a few sectors. (One thread per
    // row scatters every lane of a warp over a different row and quadruples the L2 sector traffic.) The row -> lane
    // assignment is fixed, so the sums and the fused dot product are deterministic.
    constexpr int kCgLanesPerRow = 8;

    __device__ __forceinline__ void spmvRows(const CgParams& P, const double* __restrict__ x, double* __restrict__ y,
                                             const double* __restrict__ dotWith, double& dotAcc)
    {
        const uint32_t groupsPerGrid = (gridDim.x * kCgThreads) / kCgLanesPerRow;
        const uint32_t group = (blockIdx.x * kCgThreads + threadIdx.x) / kCgLanesPerRow;
        const uint32_t sub = threadIdx.x % kCgLanesPerRow;
        const uint32_t rows = (P.n + groupsPerGrid - 1) / groupsPerGrid * groupsPerGrid;     // keep whole warps converged for the shuffles
        for (uint32_t row = group; row < rows; row += groupsPerGrid)
        {
            double s = 0.0;
            if (row < P.n)
            {
                const uint32_t e = P.rowPtr[row + 1];
                for (uint32_t k = P.rowPtr[row] + sub; k < e; k += kCgLanesPerRow) s = fma(P.val[k], x[P.col[k]], s);
            }
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
            s += __shfl_xor_sync(0xFFFFFFFFu, s, 4);
            if (sub == 0 && row < P.n) { y[row] = s; dotAcc = fma(s, dotWith[row], dotAcc); }
        }
    }

    // Eigen's conjugate_gradient loop (the shim in the CPU checker has the same structure): diagonal preconditioner.
    __global__ void __launch_bounds__(kCgThreads) cgKernel(const CgParams P)
    {
        cg::grid_group grid = cg::this_grid();
        __shared__ double sRed[kCgThreads / 32];
        const int nb = gridDim.x;
        const uint32_t tid = blockIdx.x * kCgThreads + threadIdx.x, stride = gridDim.x * kCgThreads;
        double* part0 = P.partial; double* part1 = P.partial + nb; double* part2 = P.partial + 2 * nb;

        // diagonal, r = b - A x0, |b|^2, |r|^2
        for (uint32_t i = tid; i < P.n; i += stride)
        {
            double d = 1.0;
            for (uint32_t k = P.rowPtr[i]; k < P.rowPtr[i + 1]; ++k) if (P.col[k] == i && P.val[k] != 0.0) d = 1.0 / P.val[k];
            P.invDiag[i] = d;
        }
        double dummy = 0.0;
        spmvRows(P, P.x, P.ap, P.x, dummy);
        grid.sync();
        double bb = 0.0, rr = 0.0, rz = 0.0;
        for (uint32_t i = tid; i < P.n; i += stride)
        {
            const double bi = P.b[i], ri = bi - P.ap[i];
            P.r[i] = ri;
            const double zi = P.invDiag[i] * ri;
            P.p[i] = zi;
            bb = fma(bi, bi, bb); rr = fma(ri, ri, rr); rz = fma(ri, zi, rz);
        }
        bb = blockSum(bb, sRed); rr = blockSum(rr, sRed); rz = blockSum(rz, sRed);
        if (threadIdx.x == 0) { part0[blockIdx.x] = bb; part1[blockIdx.x] = rr; part2[blockIdx.x] = rz; }
        grid.sync();
        const double rhs2 = gridSum(part0, nb, sRed);
        double res2 = gridSum(part1, nb, sRed);
        double absNew = gridSum(part2, nb, sRed);
        const double threshold = fmax(P.tol * P.tol * rhs2, 2.2250738585072014e-308);
        uint32_t it = 0;
        if (rhs2 == 0.0)
        {
            for (uint32_t i = tid; i < P.n; i += stride) P.x[i] = 0.0;
            res2 = 0.0;
        }
        else if (res2 >= threshold)
        {
            while (it < P.maxIt)
            {
                grid.sync();                                   // p complete, partial buffers free
                double pAp = 0.0;
                spmvRows(P, P.p, P.ap, P.p, pAp);
                pAp = blockSum(pAp, sRed);
                if (threadIdx.x == 0) part0[blockIdx.x] = pAp;
                grid.sync();
                const double alpha = absNew / gridSum(part0, nb, sRed);
                double rr2 = 0.0, rz2 = 0.0;
                for (uint32_t i = tid; i < P.n; i += stride)
                {
                    P.x[i] = fma(alpha, P.p[i], P.x[i]);
                    const double ri = fma(-alpha, P.ap[i], P.r[i]);
                    P.r[i] = ri;
                    rr2 = fma(ri, ri, rr2); rz2 = fma(ri, P.invDiag[i] * ri, rz2);
                }
                rr2 = blockSum(rr2, sRed); rz2 = blockSum(rz2, sRed);
                if (threadIdx.x == 0) { part1[blockIdx.x] = rr2; part2[blockIdx.x] = rz2; }
                grid.sync();
                res2 = gridSum(part1, nb, sRed);
                if (res2 < threshold) break;
                const double absOld = absNew;
                absNew = gridSum(part2, nb, sRed);
                const double beta = absNew / absOld;
                for (uint32_t i = tid; i < P.n; i += stride) P.p[i] = fma(beta, P.p[i], P.invDiag[i] * P.r[i]);
                ++it;
            }
        }
        if (tid == 0) { P.result[0] = (double)it; P.result[1] = rhs2 > 0.0 ? sqrt(res2 / rhs2) : 0.0; }
    }

    __global__ void scaleKernel(const double* __restrict__ in, double* __restrict__ out, uint32_t n, double s)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) out[i] = in[i] * s;                                                        // oldCoeffs *= strength (:1741)
    }
}
