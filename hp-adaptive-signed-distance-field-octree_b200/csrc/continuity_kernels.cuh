// continuity_kernels.cuh — PerformContinuityPostProcess (Source/HP/Octree.cpp:1717-1762) on the device.
//
// Reference: every leaf-leaf shared face contributes the Gram matrix of the jump (f_A - f_B) to a sparse matrix M
// (EvaluateSharedFaceIntegralAnalytically :1459-1546 for equal depths, ...Numerically :1250-1456 otherwise), lambda is
// added on the diagonal, and (M + lambda I) x = lambda c is solved with Eigen's ConjugateGradient (:1751-1755).
//
// Here: (1) one CTA per face emits that face's entries as (row << bits | col, value) pairs (bits = ceil(log2 n): the radix sort then needs 2 bits / 8 passes) at a precomputed offset —
// positions inside a face come from an ordered block-wide compaction, so the COO order (and hence the order in which
// duplicates are summed) is deterministic; (2) a radix sort + segmented reduction (CUB: plumbing) turns COO into CSR
// with duplicates summed, which is what setFromTriplets does (:1732-1735); (3) a hand-written persistent cooperative
// conjugate-gradient kernel (sparse matrix-vector product, fused dot products, grid-wide syncs) runs Eigen's CG loop
// with a diagonal preconditioner until |r|^2 < tol^2 |b|^2.
#pragma once
#include <cub/cub.cuh>
#include "hp_common.h"
#include "device_ctx.h"

namespace hpsdf
{
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
                   uint64_t* __restrict__ keys, double* __restrict__ vals, const int keyShift)   // key = row << keyShift | col
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
                            key  = ((uint64_t)(rowStart + i) << keyShift) | (uint64_t)(colStart + j);
                            keyT = ((uint64_t)(colStart + j) << keyShift) | (uint64_t)(rowStart + i);
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
                keys[pos] = ((uint64_t)(rowStart + i) << keyShift) | (uint64_t)(colStart + j);
                vals[pos] = v;
                if (blk == 1)
                {
                    keys[pos + 1] = ((uint64_t)(colStart + j) << keyShift) | (uint64_t)(rowStart + i);
                    vals[pos + 1] = v;
                }
            }
            base += (blk == 1 ? 2ull : 1ull) * (unsigned long long)nCand;
        }
    }

    // ---- face-pair enumeration on the device -------------------------------------------------------------------------------
    // RunContinuityThreadPool enumerates every leaf-leaf shared face with the NodeProc / FaceProc recursion (Octree.cpp:
    // 1549-1612, 1663-1714). The same set, one thread per node of the 56-byte SDF::Node image (finish_kernels.cuh): a leaf looks
    // across each of its six faces for the node of its own depth (or the coarser leaf that contains that position);
    //   + direction: neighbour is a leaf of the same or a coarser depth  -> face (me = low side, neighbour = high side)
    //   - direction: neighbour is a strictly coarser leaf                -> face (neighbour = low side, me = high side)
    // A finer neighbourhood is left to its own leaves, so every shared face is produced exactly once, in node order.
    struct NodeRec { unsigned long long child; float mn[3], mx[3]; uint32_t cstart; uint32_t degree, depth; };
    __device__ __forceinline__ NodeRec readNodeRec(const unsigned char* __restrict__ image, uint32_t i)
    {
        const uint2* r = reinterpret_cast<const uint2*>(image + 56 * (size_t)i);
        NodeRec n;
        const uint2 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5], g = r[6];
        n.child = ((unsigned long long)a.y << 32) | a.x;
        n.mn[0] = __uint_as_float(b.x); n.mn[1] = __uint_as_float(b.y); n.mn[2] = __uint_as_float(c.x);
        n.mx[0] = __uint_as_float(c.y); n.mx[1] = __uint_as_float(d.x); n.mx[2] = __uint_as_float(d.y);
        n.cstart = e.x; n.degree = f.x & 0xFFu; n.depth = g.x & 0xFFu;
        return n;
    }

    // entries an analytic block emits: (i, j) with equal tangential indices (table built on the host, continuity.cpp)
    struct MatchTable { const uint32_t* count; };       // [dim][degR][degC], 13 x 13 per dim

    __device__ __forceinline__ bool faceAcross(const unsigned char* __restrict__ image, uint32_t me, const NodeRec& M, int axis, int sign,
                                               const MatchTable mt, FaceJobDev& f, uint32_t& entries)
    {
        const uint32_t d = M.depth;
        uint32_t ic[3];
        #pragma unroll
        for (int a = 0; a < 3; ++a) ic[a] = (uint32_t)((M.mn[a] + 0.5f) * (float)(1u << d));         // cell index at depth d (dyadic: exact)
        if (sign > 0) { if (ic[axis] + 1u >= (1u << d)) return false; ic[axis] += 1u; }
        else { if (ic[axis] == 0u) return false; ic[axis] -= 1u; }
        uint32_t cur = 0;
        NodeRec N = readNodeRec(image, 0);
        for (uint32_t l = 1; l <= d && N.child != kNoChild; ++l)
        {
            const uint32_t sh = d - l;
            cur = (uint32_t)N.child + ((ic[0] >> sh) & 1u) + 2u * ((ic[1] >> sh) & 1u) + 4u * ((ic[2] >> sh) & 1u);
            N = readNodeRec(image, cur);
        }
        if (N.child != kNoChild) return false;                       // finer leaves over there: they produce the face
        if (sign < 0 && N.depth >= d) return false;                   // same depth: produced from the other side
        const NodeRec& A = sign > 0 ? M : N;                           // low side (Octree.cpp:1593-1594)
        const NodeRec& B = sign > 0 ? N : M;
        f.cooOffset = 0;
        f.cstartA = A.cstart; f.cstartB = B.cstart;
        f.degA = (uint8_t)A.degree; f.degB = (uint8_t)B.degree; f.depthA = (uint8_t)A.depth; f.depthB = (uint8_t)B.depth;
        f.dim = (uint8_t)axis; f.analytic = A.depth == B.depth; f.pad0 = f.pad1 = 0;       // Octree.cpp:1651
        f.faceScale = 0.0; f.invDist = 0.0; f.invTr1 = 0.0; f.invTr2 = 0.0;
        const uint32_t nA = (uint32_t)coeffCount((int)A.degree), nB = (uint32_t)coeffCount((int)B.degree);
        if (f.analytic)
        {
            const uint32_t* mc = mt.count + axis * 169;
            entries = mc[A.degree * 13 + A.degree] + 2u * mc[A.degree * 13 + B.degree] + mc[B.degree * 13 + B.degree];
        }
        else
        {
            const int t1 = (axis + 1) % 3, t2 = (axis + 2) % 3;
            double fs[3];
            #pragma unroll
            for (int i = 0; i < 3; ++i)
            {
                const float lo = fmaxf(A.mn[i], B.mn[i]), hi = fminf(A.mx[i], B.mx[i]);              // shared face = A.aabb clamped to B.aabb (Octree.cpp:1265-1266)
                fs[i] = (double)__fsub_rn(hi, lo) * 0.5;
            }
            f.faceScale = fs[t1] * fs[t2];
            const uint32_t diff = A.depth > B.depth ? A.depth - B.depth : B.depth - A.depth;
            f.invDist = 1.0 / (double)(1u << diff);                                                    // Octree.cpp:1275-1276
            const NodeRec& fine = A.depth > B.depth ? A : B;
            const NodeRec& coarse = A.depth > B.depth ? B : A;
            const float cf1 = __fdiv_rn(__fadd_rn(fine.mn[t1], fine.mx[t1]), 2.0f), cc1 = __fdiv_rn(__fadd_rn(coarse.mn[t1], coarse.mx[t1]), 2.0f);
            const float cf2 = __fdiv_rn(__fadd_rn(fine.mn[t2], fine.mx[t2]), 2.0f), cc2 = __fdiv_rn(__fadd_rn(coarse.mn[t2], coarse.mx[t2]), 2.0f);
            f.invTr1 = (double)__fsub_rn(cf1, cc1) / ((double)__fsub_rn(fine.mx[t1], fine.mn[t1]) * 0.5);   // Octree.cpp:1280-1289
            f.invTr2 = (double)__fsub_rn(cf2, cc2) / ((double)__fsub_rn(fine.mx[t2], fine.mn[t2]) * 0.5);
            f.invTr1 *= f.invDist; f.invTr2 *= f.invDist;                                              // Octree.cpp:1290
            entries = nA * nA + 2u * nA * nB + nB * nB;
        }
        return true;
    }

    // pass 1: per node (faces << 32 | COO entries); pass 2 (after an exclusive scan): the face jobs at their offsets
    __global__ void __launch_bounds__(128) faceEnumKernel(const unsigned char* __restrict__ image, uint32_t nNodes, const MatchTable mt,
                                                          const unsigned long long* __restrict__ offsets, uint32_t cooBase,
                                                          unsigned long long* __restrict__ counts, FaceJobDev* __restrict__ faces)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= nNodes) return;
        const NodeRec M = readNodeRec(image, i);
        unsigned long long mine = 0;
        if (M.child == kNoChild)
        {
            unsigned long long off = offsets ? offsets[i] : 0ull;
            for (int axis = 0; axis < 3; ++axis)
                for (int sign = 1; sign >= -1; sign -= 2)
                {
                    FaceJobDev f;
                    uint32_t entries = 0;
                    if (!faceAcross(image, i, M, axis, sign, mt, f, entries)) continue;
                    if (faces)
                    {
                        f.cooOffset = (unsigned long long)cooBase + (off & 0xFFFFFFFFull);
                        faces[off >> 32] = f;
                        off += (1ull << 32) | entries;
                    }
                    mine += (1ull << 32) | entries;
                }
        }
        if (counts) counts[i] = mine;
    }

    __global__ void diagEmitKernel(uint64_t* __restrict__ keys, double* __restrict__ vals, uint32_t n, double lambda, int keyShift)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) { keys[i] = ((uint64_t)i << keyShift) | i; vals[i] = lambda; }                   // Octree.cpp:1724-1729
    }

    __global__ void flagNonZeroKernel(const double* __restrict__ v, uint32_t n, uint8_t* __restrict__ flags)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) flags[i] = v[i] != 0.0;
    }

    // rowPtr from sorted unique keys: rowPtr[r] = first entry whose row >= r
    __global__ void rowPtrKernel(const uint64_t* __restrict__ keys, uint32_t nnz, uint32_t n, uint32_t* __restrict__ rowPtr,
                                 uint32_t* __restrict__ col, int keyShift)
    {
        const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
        if (k > nnz) return;
        const uint32_t rowHere = k < nnz ? (uint32_t)(keys[k] >> keyShift) : n;
        const uint32_t rowPrev = k > 0 ? (uint32_t)(keys[k - 1] >> keyShift) : 0xFFFFFFFFu;
        if (k < nnz) col[k] = (uint32_t)(keys[k] & ((1ull << keyShift) - 1ull));
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
        double* r; double* p; double* ap; double* invDiag; double* u; double* s;
        double* partial;          // 4 * gridDim.x scratch for block partial sums
        unsigned* barrier;        // arrival counter of the grid-wide barrier (zero at launch)
        uint32_t eCap, slots;     // shared-memory staging of the matrix: entries per thread, rows per 8-lane group
        uint32_t rowsStaged;      // 1 = the (start, length) of every owned row is staged too
        double* result;           // [0] iterations, [1] relative residual
    };

    constexpr int kCgThreads = 1024;           // one CTA per SM: 148 arrivals per grid-wide barrier, and up to ~200 KB of shared memory for the matrix

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

    // three sums at once: one exchange through shared memory instead of three
    __device__ __forceinline__ void blockSum3(double& a, double& b, double& c, double* sRed3 /* 3 * warps */)
    {
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        }
        constexpr int W = kCgThreads / 32;
        if ((threadIdx.x & 31) == 0) { sRed3[threadIdx.x >> 5] = a; sRed3[W + (threadIdx.x >> 5)] = b; sRed3[2 * W + (threadIdx.x >> 5)] = c; }
        __syncthreads();
        double sa = 0.0, sb = 0.0, sc = 0.0;
        #pragma unroll
        for (int w = 0; w < W; ++w) { sa += sRed3[w]; sb += sRed3[W + w]; sc += sRed3[2 * W + w]; }
        __syncthreads();
        a = sa; b = sb; c = sc;
    }

    // every block sums the block partials of three quantities in the same fixed order (warps 0..2 take one each)
    __device__ __forceinline__ void gridSum3(const double* p0, const double* p1, const double* p2, int nBlocks, double* sOut3, double& a, double& b, double& c)
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (warp < 3)
        {
            const double* p = warp == 0 ? p0 : (warp == 1 ? p1 : p2);
            double s = 0.0;
            for (int i = lane; i < nBlocks; i += 32) s += p[i];
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            if (lane == 0) sOut3[warp] = s;
        }
        __syncthreads();
        a = sOut3[0]; b = sOut3[1]; c = sOut3[2];
        __syncthreads();
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

    // The matrix does not change during the solve and an 8-lane group owns the same rows in every iteration, so each CTA
    // keeps ITS part of the CSR arrays in shared memory for the whole kernel (95 k rows x 20 entries x 12 B = 23 MB over 148
    // SMs = 155 KB per SM): a matrix-vector product then reads nothing from L2 but the gathered vector entries, four
    // independent gathers in flight per lane. Entries beyond the staged capacity are read from global memory (any size works).
    struct CgStage
    {
        double*   val;        // [entry * blockDim + thread]
        uint32_t* col;
        uint32_t* rowStart;   // [slot * groupsPerBlock + group]: CSR start of the row the group owns in that slot
        uint32_t* rowLen;
        uint32_t  eCap;       // staged entries per thread
        uint32_t  slots;      // rows per group (uniform trip count)
    };

    __device__ __forceinline__ void cgStageMatrix(const CgParams& P, const CgStage& S)
    {
        const uint32_t groupsPerGrid = (gridDim.x * kCgThreads) / kCgLanesPerRow, gpb = kCgThreads / kCgLanesPerRow;
        const uint32_t group = (blockIdx.x * kCgThreads + threadIdx.x) / kCgLanesPerRow, gl = threadIdx.x / kCgLanesPerRow;
        const uint32_t sub = threadIdx.x % kCgLanesPerRow;
        uint32_t pos = 0;
        for (uint32_t q = 0; q < S.slots; ++q)
        {
            const uint32_t row = group + q * groupsPerGrid;
            uint32_t start = 0, len = 0;
            if (row < P.n) { start = P.rowPtr[row]; len = P.rowPtr[row + 1] - start; }
            if (sub == 0 && S.rowStart) { S.rowStart[q * gpb + gl] = start; S.rowLen[q * gpb + gl] = len; }
            for (uint32_t k = sub; k < len; k += kCgLanesPerRow, ++pos)
                if (pos < S.eCap) { S.val[pos * kCgThreads + threadIdx.x] = P.val[start + k]; S.col[pos * kCgThreads + threadIdx.x] = P.col[start + k]; }
        }
        __syncthreads();
    }

    __device__ __forceinline__ void spmvStaged(const CgParams& P, const CgStage& S, const double* __restrict__ x, double* __restrict__ y,
                                               const double* __restrict__ dotWith, double& dotAcc)
    {
        const uint32_t groupsPerGrid = (gridDim.x * kCgThreads) / kCgLanesPerRow, gpb = kCgThreads / kCgLanesPerRow;
        const uint32_t group = (blockIdx.x * kCgThreads + threadIdx.x) / kCgLanesPerRow, gl = threadIdx.x / kCgLanesPerRow;
        const uint32_t sub = threadIdx.x % kCgLanesPerRow;
        uint32_t pos = 0;
        for (uint32_t q = 0; q < S.slots; ++q)
        {
            const uint32_t row = group + q * groupsPerGrid;
            uint32_t start = 0, len = 0;
            if (S.rowStart) { start = S.rowStart[q * gpb + gl]; len = S.rowLen[q * gpb + gl]; }
            else if (row < P.n) { start = P.rowPtr[row]; len = P.rowPtr[row + 1] - start; }
            const uint32_t c = len > sub ? (len - sub + kCgLanesPerRow - 1) / kCgLanesPerRow : 0u;
            double acc = 0.0;
            for (uint32_t j = 0; j < c; j += 6)
            {
                // up to six gathers in flight per lane (a scalar loop would serialise the L2 round trips: issue is in order)
                double vv[6], xv[6];
                #pragma unroll
                for (uint32_t uu = 0; uu < 6; ++uu)
                {
                    const uint32_t jj = j + uu;
                    vv[uu] = 0.0; xv[uu] = 0.0;
                    if (jj < c)
                    {
                        uint32_t cc;
                        if (pos + jj < S.eCap) { const uint32_t e = (pos + jj) * kCgThreads + threadIdx.x; vv[uu] = S.val[e]; cc = S.col[e]; }
                        else { const uint32_t k = start + sub + jj * kCgLanesPerRow; vv[uu] = P.val[k]; cc = P.col[k]; }
                        xv[uu] = x[cc];
                    }
                }
                #pragma unroll
                for (uint32_t uu = 0; uu < 6; ++uu) acc = fma(vv[uu], xv[uu], acc);
            }
            pos += c;
            acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 1);
            acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 2);
            acc += __shfl_xor_sync(0xFFFFFFFFu, acc, 4);
            if (sub == 0 && row < P.n) { y[row] = acc; dotAcc = fma(acc, dotWith[row], dotAcc); }
        }
    }

    __device__ __forceinline__ unsigned long long cgTimerNs()
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        return t;
    }

    // Grid-wide barrier for a cooperatively launched (co-resident) grid: one arrival per block on a monotone counter, the
    // arriving thread spins until `phase * gridDim.x` arrivals are in. Cheaper than cg::grid.sync(), same guarantees here.
    __device__ __forceinline__ void gridBarrier(unsigned* counter, unsigned& phase)
    {
        __syncthreads();
        if (threadIdx.x == 0)
        {
            ++phase;
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned target = phase * gridDim.x;
            while (*((volatile unsigned*)counter) < target) { }
            __threadfence();
        }
        __syncthreads();
    }

    // Preconditioned conjugate gradients in the Chronopoulos-Gear form: both inner products of an iteration are taken after
    // its matrix-vector product, so an iteration needs TWO grid-wide barriers (vector update | product + reductions) where
    // the textbook loop (Eigen's conjugate_gradient, which the reference calls at Octree.cpp:1751-1755) needs three:
    //     p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;  u = D^-1 r          | barrier
    //     w = A u;  gamma' = (r, u);  delta = (w, u);  |r|^2                                    | barrier
    //     beta = gamma' / gamma;  alpha = gamma' / (delta - beta gamma' / alpha)
    // Same Krylov iterates in exact arithmetic; Eigen's stopping rule |r|^2 < tol^2 |b|^2 on the recursively updated residual,
    // diagonal preconditioner, x0 = what the caller left in x (c, or the reference's lambda c: Octree.cpp:1738-1755). Deterministic: fixed row -> lane assignment and
    // fixed-shape reductions.
    __global__ void __launch_bounds__(kCgThreads) cgKernel(const CgParams P)
    {
        __shared__ double sRed[kCgThreads / 32];
        __shared__ double sRed3[3 * (kCgThreads / 32)];
        __shared__ double sOut3[3];
        const int nb = gridDim.x;
        const uint32_t tid = blockIdx.x * kCgThreads + threadIdx.x, stride = gridDim.x * kCgThreads;
        double* part0 = P.partial; double* part1 = P.partial + nb; double* part2 = P.partial + 2 * nb; double* part3 = P.partial + 3 * nb;
        unsigned phase = 0;
        double* u = P.u; double* w = P.ap; double* s = P.s;
        unsigned long long tA = 0, tB1 = 0, tS = 0, tB2 = 0, t0 = 0, t1 = 0;      // diagnostics (block 0): ns in update | barrier | product | barrier + sums
        const bool timing = blockIdx.x == 0 && threadIdx.x == 0;
        extern __shared__ double cgSmem[];
        CgStage S;
        S.eCap = P.eCap; S.slots = P.slots;
        S.val = cgSmem; S.col = reinterpret_cast<uint32_t*>(cgSmem + (size_t)P.eCap * kCgThreads);
        S.rowStart = P.rowsStaged ? S.col + (size_t)P.eCap * kCgThreads : nullptr;
        S.rowLen = P.rowsStaged ? S.rowStart + (size_t)P.slots * (kCgThreads / kCgLanesPerRow) : nullptr;
        cgStageMatrix(P, S);

        // diagonal; w = A x0 (x0 was written by the launch before: no barrier needed)
        for (uint32_t i = tid; i < P.n; i += stride)
        {
            double d = 1.0;
            for (uint32_t k = P.rowPtr[i]; k < P.rowPtr[i + 1]; ++k) if (P.col[k] == i && P.val[k] != 0.0) d = 1.0 / P.val[k];
            P.invDiag[i] = d;
        }
        double dummy = 0.0;
        spmvStaged(P, S, P.x, w, P.x, dummy);
        gridBarrier(P.barrier, phase);
        // r = b - A x0, u = D^-1 r
        double bb = 0.0, rr = 0.0, ga = 0.0;
        for (uint32_t i = tid; i < P.n; i += stride)
        {
            const double bi = P.b[i], ri = bi - w[i];
            P.r[i] = ri;
            const double ui = P.invDiag[i] * ri;
            u[i] = ui;
            P.p[i] = 0.0; s[i] = 0.0;
            bb = fma(bi, bi, bb); rr = fma(ri, ri, rr); ga = fma(ri, ui, ga);
        }
        bb = blockSum(bb, sRed); rr = blockSum(rr, sRed); ga = blockSum(ga, sRed);
        gridBarrier(P.barrier, phase);                 // u complete
        double de = 0.0;
        spmvStaged(P, S, u, w, u, de);
        de = blockSum(de, sRed);
        if (threadIdx.x == 0) { part0[blockIdx.x] = bb; part1[blockIdx.x] = rr; part2[blockIdx.x] = ga; part3[blockIdx.x] = de; }
        gridBarrier(P.barrier, phase);
        const double rhs2 = gridSum(part0, nb, sRed);
        double res2 = gridSum(part1, nb, sRed);
        double gamma = gridSum(part2, nb, sRed);
        double delta = gridSum(part3, nb, sRed);
        const double threshold = fmax(P.tol * P.tol * rhs2, 2.2250738585072014e-308);
        uint32_t it = 0;
        if (rhs2 == 0.0)
        {
            for (uint32_t i = tid; i < P.n; i += stride) P.x[i] = 0.0;
            res2 = 0.0;
        }
        else if (res2 >= threshold)
        {
            double alpha = gamma / delta, beta = 0.0;
            while (it < P.maxIt)
            {
                if (timing) t0 = cgTimerNs();
                double rr2 = 0.0, ga2 = 0.0;
                for (uint32_t i = tid; i < P.n; i += stride)
                {
                    const double pi = fma(beta, P.p[i], u[i]);
                    const double si = fma(beta, s[i], w[i]);
                    P.p[i] = pi; s[i] = si;
                    P.x[i] = fma(alpha, pi, P.x[i]);
                    const double ri = fma(-alpha, si, P.r[i]);
                    P.r[i] = ri;
                    const double ui = P.invDiag[i] * ri;
                    u[i] = ui;
                    rr2 = fma(ri, ri, rr2); ga2 = fma(ri, ui, ga2);
                }
                if (timing) { t1 = cgTimerNs(); tA += t1 - t0; t0 = t1; }
                gridBarrier(P.barrier, phase);             // u complete (and the partial buffers of the last iteration are consumed)
                if (timing) { t1 = cgTimerNs(); tB1 += t1 - t0; t0 = t1; }
                double de2 = 0.0;
                spmvStaged(P, S, u, w, u, de2);
                blockSum3(rr2, ga2, de2, sRed3);
                if (threadIdx.x == 0) { part1[blockIdx.x] = rr2; part2[blockIdx.x] = ga2; part3[blockIdx.x] = de2; }
                if (timing) { t1 = cgTimerNs(); tS += t1 - t0; t0 = t1; }
                gridBarrier(P.barrier, phase);
                double gammaNew;
                gridSum3(part1, part2, part3, nb, sOut3, res2, gammaNew, delta);
                if (timing) { t1 = cgTimerNs(); tB2 += t1 - t0; }
                ++it;
                if (res2 < threshold) break;
                beta = gammaNew / gamma;
                alpha = gammaNew / (delta - beta * gammaNew / alpha);
                gamma = gammaNew;
            }
        }
        if (tid == 0)
        {
            P.result[0] = (double)it; P.result[1] = rhs2 > 0.0 ? sqrt(res2 / rhs2) : 0.0;
            P.result[2] = (double)tA; P.result[3] = (double)tB1; P.result[4] = (double)tS; P.result[5] = (double)tB2;
        }
    }

    __global__ void scaleKernel(const double* __restrict__ in, double* __restrict__ out, uint32_t n, double s)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) out[i] = in[i] * s;                                                        // oldCoeffs *= strength (:1741)
    }
}
