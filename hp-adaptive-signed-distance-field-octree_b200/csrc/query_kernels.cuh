// query_kernels.cuh — batched Octree::Query (Source/HP/Octree.cpp:662-702): the reference has no batch API
// (callers loop over Query, HPBenchmarks.cpp:107-110); this is the data-parallel form.
//
// Layout: points are AoS n x 3 f64 (the Eigen::Vector3d array a caller already has). A CTA stages a tile of 256 points
// (6 KB) into shared memory with coalesced 16-byte loads, then each thread evaluates one point. The 4096-entry top-of-tree
// table (16 KB) is staged once per CTA; the grid is persistent (a few CTAs per SM looping over tiles) so that staging
// amortises. Algorithmic HBM traffic: 24 B in + 8 B out per point; the tree (16 B nodes + padded coefficients) stays in L2.
#pragma once
#include "query_eval.cuh"

namespace hpsdf
{
    constexpr int kQueryThreads = 256;
#ifndef HPSDF_QUERY_MINBLOCKS
#define HPSDF_QUERY_MINBLOCKS 4
#endif
    constexpr int kQueryBlocksPerSm = HPSDF_QUERY_MINBLOCKS;

    __global__ void __launch_bounds__(kQueryThreads, kQueryBlocksPerSm)
    queryKernel(const DeviceTreeView view, const double* __restrict__ xyz, size_t n, double* __restrict__ out,
                const uint32_t* __restrict__ bidx, const int aligned /* xyz starts on a 16-byte boundary */)
    {
        __shared__ uint32_t sTop[4096];
        __shared__ __align__(16) double sPts[kQueryThreads / 32][96];          // per warp: 32 points x 3 doubles
        const bool useTop = view.top != nullptr;
        if (useTop)
        {
            for (int i = threadIdx.x; i < 1024; i += kQueryThreads)
                reinterpret_cast<uint4*>(sTop)[i] = __ldg(reinterpret_cast<const uint4*>(view.top) + i);
            __syncthreads();                        // the only block-wide barrier: the table is read-only afterwards
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const size_t nGroups = (n + 31) / 32;                                   // a warp handles 32 consecutive points
        const size_t warpsTotal = (size_t)gridDim.x * (kQueryThreads / 32);
        double* sp = sPts[warp];
        for (size_t g = (size_t)blockIdx.x * (kQueryThreads / 32) + warp; g < nGroups; g += warpsTotal)
        {
            const size_t base = g * 32;
            const size_t cnt = n - base < 32 ? n - base : 32;
            const double* src = xyz + 3 * base;                                 // 32 * 24 B = 768 B: 16-byte aligned
            if (cnt == 32 && aligned)
            {
                reinterpret_cast<double2*>(sp)[lane] = __ldcs(reinterpret_cast<const double2*>(src) + lane);
                if (lane < 16) reinterpret_cast<double2*>(sp)[32 + lane] = __ldcs(reinterpret_cast<const double2*>(src) + 32 + lane);
            }
            else
                for (size_t v = lane; v < 3 * cnt; v += 32) sp[v] = src[v];
            __syncwarp();
            LeafHit h;
            h.coeffs = view.coeffs; h.ux = h.uy = h.uz = 0.0; h.degree = -1; h.depth = 0;
            if ((size_t)lane < cnt) h = findLeaf(view.nodes, view.coeffs, useTop ? sTop : nullptr, view.map, sp[3 * lane], sp[3 * lane + 1], sp[3 * lane + 2]);
            __syncwarp();                           // sp is overwritten by the next group
            const double v = evalWarp(h, bidx);
            if ((size_t)lane < cnt) __stcs(out + base + lane, v);
        }
    }

    // Octree::QueryWithGradient (Octree.cpp:749-789) / FApproxWithGradient (:904-985): central differences with eps = 1e-4
    // in the leaf's local coordinate, gradient normalised. One thread per point, generic-degree loops (not a hot path).
    __global__ void __launch_bounds__(256)
    queryGradientKernel(const DeviceTreeView view, const double* __restrict__ xyz, size_t n, double* __restrict__ out,
                        double* __restrict__ grad, const uint32_t* __restrict__ bidx)
    {
        const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t >= n) return;
        const RootMap& map = view.map;
        const double px = (xyz[3 * t] - map.centre[0]) * map.invSizes[0];
        const double py = (xyz[3 * t + 1] - map.centre[1]) * map.invSizes[1];
        const double pz = (xyz[3 * t + 2] - map.centre[2]) * map.invSizes[2];
        const float fx = (float)px, fy = (float)py, fz = (float)pz;
        if (!(fx >= -0.5f && fx <= 0.5f && fy >= -0.5f && fy <= 0.5f && fz >= -0.5f && fz <= 0.5f)) { out[t] = DBL_MAX; return; }
        double c[3] = { 0.0, 0.0, 0.0 }, q = 0.25;
        const double p[3] = { px, py, pz };
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(view.nodes));
        while (raw.z == kInternalTag)
        {
            uint32_t b[3];
            for (int a = 0; a < 3; ++a) { b[a] = p[a] >= c[a]; c[a] += b[a] ? q : -q; }
            q *= 0.5;
            raw = __ldg(reinterpret_cast<const uint4*>(view.nodes) + raw.x + b[0] + (b[1] << 1) + (b[2] << 2));
        }
        const int depth = (int)raw.w, degree = (int)raw.z;
        const double* coeffs = view.coeffs + raw.y;
        const double scale = (double)(2u << depth), eps = 0.0001;
        double lut[kMaxDegree + 1][3][3];
        for (int a = 0; a < 3; ++a)
        {
            const double u = (p[a] - c[a]) * scale;
            const double us[3] = { u, u + eps, u - eps };
            for (int v = 0; v < 3; ++v)
            {
                lut[0][a][v] = c_nl[0][depth];
                double m2 = 0.0, m1 = 1.0, l = 1.0;
                for (int j = 1; j <= degree; ++j)
                {
                    l = c_rec[j][0] * us[v] * m1 - c_rec[j][1] * m2; m2 = m1; m1 = l;
                    lut[j][a][v] = l * c_nl[j][depth];
                }
            }
        }
        const int nc = coeffCount(degree);
        double g[3];
        for (int k = 0; k < 3; ++k)
        {
            double fp = 0.0, fm = 0.0;
            for (int i = 0; i < nc; ++i)
            {
                const int a = (bidx[i] >> (8 * k)) & 0xFF;
                fp += coeffs[i] * lut[a][k][1]; fm += coeffs[i] * lut[a][k][2];            // Octree.cpp:961-965
            }
            g[k] = (fp - fm) / (2.0 * eps);
        }
        const double nrm = sqrt(g[0] * g[0] + (g[1] * g[1] + g[2] * g[2]));
        if (nrm > 0.0) { g[0] /= nrm; g[1] /= nrm; g[2] /= nrm; }
        double f = 0.0;
        for (int i = 0; i < nc; ++i)
        {
            const uint32_t abc = bidx[i];
            f += coeffs[i] * ((lut[abc & 0xFF][0][0] * lut[(abc >> 8) & 0xFF][1][0]) * lut[(abc >> 16) & 0xFF][2][0]);
        }
        out[t] = f;
        grad[3 * t] = g[0]; grad[3 * t + 1] = g[1]; grad[3 * t + 2] = g[2];
    }

    __global__ void gatherSegmentsKernel(const double* __restrict__ src, double* __restrict__ dst, const uint32_t* __restrict__ srcOff,
                                         const uint32_t* __restrict__ dstOff, const uint32_t* __restrict__ count, uint32_t nSeg)
    {
        // one warp per segment
        const uint32_t seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (seg >= nSeg) return;
        const double* s = src + srcOff[seg];
        double* d = dst + dstOff[seg];
        for (uint32_t i = lane; i < count[seg]; i += 32) d[i] = s[i];
    }

    // Octree::QueryRay (Source/HP/Octree.cpp:705-746) + Ray::Ray / Ray::IntersectAABB (Source/HP/Ray.cpp:5-65), one ray per
    // thread, statement by statement — including what the reference does with its own intermediate results: the origin is
    // moved into the unit cube but the direction is not rescaled (:713); when the origin is outside the root, the vector the
    // march starts from is the `a_` output of IntersectAABB, whose x component holds the entry PARAMETER and whose y, z
    // components hold the slab parameters of those axes (:715-719, Ray.cpp:27-61), not a point; every sample goes through
    // Query, which applies the root map a second time (:729, :665); `t_` receives the last field value, not a distance
    // (:733). For the default root box and origins inside it these coincide with sphere tracing of the field.
    __global__ void __launch_bounds__(256)
    queryRayKernel(const DeviceTreeView view, const double* __restrict__ origins, const double* __restrict__ dirs, size_t n,
                   const double tMax, unsigned char* __restrict__ hit, double* __restrict__ tOut)
    {
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= n) return;
        constexpr unsigned MAX_STEPS = 200;
        constexpr double eps = 0.0001, minStep = 0.0001;
        const RootMap& map = view.map;
        double o[3], dir[3], inv[3];
        int sign[3];
        for (int a = 0; a < 3; ++a)
        {
            o[a] = (origins[3 * i + a] - map.centre[a]) * map.invSizes[a];                     // :713
            dir[a] = dirs[3 * i + a];
            inv[a] = 1.0 / dir[a];                                                              // Ray.cpp:10
            sign[a] = inv[a] < 0.0;                                                             // Ray.cpp:11-13
        }
        double intMin[3] = { o[0], o[1], o[2] };
        const float fx = (float)o[0], fy = (float)o[1], fz = (float)o[2];
        const bool inside = fx >= -0.5f && fx <= 0.5f && fy >= -0.5f && fy <= 0.5f && fz >= -0.5f && fz <= 0.5f;      // nodes[0].aabb.contains
        bool ok = true;
        if (!inside)
        {
            const double bounds[2] = { -0.5, 0.5 };
            double a_[3] = { o[0], o[1], o[2] }, b_[3] = { 0.0, 0.0, 0.0 };
            a_[0] = (bounds[sign[0]] - o[0]) * inv[0];     b_[0] = (bounds[1 - sign[0]] - o[0]) * inv[0];     // Ray.cpp:27-30
            a_[1] = (bounds[sign[1]] - o[1]) * inv[1];     b_[1] = (bounds[1 - sign[1]] - o[1]) * inv[1];
            if ((a_[0] > b_[1]) || (a_[1] > b_[0])) ok = false;                                                // :32-35
            if (ok)
            {
                if (a_[1] > a_[0]) a_[0] = a_[1];                                                              // :37-40
                if (b_[1] < b_[0]) b_[0] = b_[1];                                                              // :42-45
                a_[2] = (bounds[sign[2]] - o[2]) * inv[2]; b_[2] = (bounds[1 - sign[2]] - o[2]) * inv[2];     // :47-48
                if ((a_[0] > b_[2]) || (a_[2] > b_[0])) ok = false;                                            // :50-53
                if (ok)
                {
                    if (a_[2] > a_[0]) a_[0] = a_[2];                                                          // :55-58
                    intMin[0] = a_[0]; intMin[1] = a_[1]; intMin[2] = a_[2];
                }
            }
        }
        unsigned char h = 0;
        double t = 0.0;
        if (ok)
        {
            double d = 0.0;
            for (unsigned s = 0; s < MAX_STEPS; ++s)
            {
                const double v = queryPoint(view.nodes, view.coeffs, view.top, map, intMin[0] + d * dir[0], intMin[1] + d * dir[1], intMin[2] + d * dir[2]);
                if (v < eps) { t = v; h = 1; break; }                                           // :731-736
                d += v * 0.95 + minStep;                                                        // :739
                if (d > tMax) break;                                                            // :741-744
            }
        }
        hit[i] = h;
        tOut[i] = t;
    }
}
