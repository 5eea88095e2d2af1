// query_eval.cuh — Octree::Query (Source/HP/Octree.cpp:662-702) and FApprox (:859-901) for one point, on the device.
//
// Reference: map the point into the unit cube with the f32-ROUNDED inverse root sizes (:323, :420, :665); reject it if
// its f32 cast is outside [-0.5, 0.5]^3 (:668-671, DBL_MAX); descend comparing the f64 coordinate with the f32 cell
// midpoint (:677-685); evaluate sum_i coeffs[i] * Lx[a_i] Ly[b_i] Lz[c_i] with Legendre recurrences scaled by
// NormalisedLengths[.][depth] in BasisIndexValues order.
//
// Here: cells are dyadic, so midpoints are tracked exactly in f64 instead of being loaded; the complete 16^3 grid of
// UniformlyRefine (:112-191) is entered through a 4096-entry table (shared-memory staged by the batch kernel) after four
// compare-only levels; nodes are 16-byte records (one LDG.128); coefficients sit in a padded store where every leaf is
// 16-byte aligned so they stream in as LDG.128 pairs; the basis loop is unrolled per degree with all Legendre values in
// registers.
#pragma once
#include <cfloat>
#include "hp_common.h"

namespace hpsdf
{
    template <int DEG>
    __device__ __forceinline__ double evalLeaf(const double* __restrict__ coeffs, double ux, double uy, double uz, int depth)
    {
        double lx[DEG + 1], ly[DEG + 1], lz[DEG + 1];
        {
            const double nl0 = c_nl[0][depth];
            lx[0] = nl0; ly[0] = nl0; lz[0] = nl0;
            double ax2 = 0.0, ax1 = 1.0, ay2 = 0.0, ay1 = 1.0, az2 = 0.0, az1 = 1.0;
            #pragma unroll
            for (int j = 1; j <= DEG; ++j)
            {
                const double r0 = c_rec[j][0], r1 = c_rec[j][1], nl = c_nl[j][depth];
                const double ax = r0 * ux * ax1 - r1 * ax2; ax2 = ax1; ax1 = ax; lx[j] = ax * nl;      // Octree.cpp:879-883
                const double ay = r0 * uy * ay1 - r1 * ay2; ay2 = ay1; ay1 = ay; ly[j] = ay * nl;
                const double az = r0 * uz * az1 - r1 * az2; az2 = az1; az1 = az; lz[j] = az * nl;
            }
        }
        // coefficients are 16-byte aligned in the padded store: read them as double2
        const double2* c2 = reinterpret_cast<const double2*>(coeffs);
        double f = 0.0;
        int idx = 0;
        double2 cur = make_double2(0.0, 0.0);
        #pragma unroll
        for (int p = 0; p <= DEG; ++p)
        {
            #pragma unroll
            for (int i = 0; i <= p; ++i)
            {
                #pragma unroll
                for (int j = 0; j <= p - i; ++j)
                {
                    const int k = p - i - j;
                    if (idx < coeffCount(DEG))                       // drops (6,0,0) when DEG == 6: the reference stores 83 coefficients
                    {
                        if ((idx & 1) == 0) cur = __ldg(c2 + (idx >> 1));
                        const double cv = (idx & 1) ? cur.y : cur.x;
                        f = fma(cv, (lx[i] * ly[j]) * lz[k], f);     // Octree.cpp:891-897
                    }
                    ++idx;
                }
            }
        }
        return f;
    }

    __device__ __forceinline__ double evalLeafAnyDegree(const double* __restrict__ coeffs, int degree, double ux, double uy, double uz, int depth)
    {
        switch (degree)
        {
            case 0:  return evalLeaf<0>(coeffs, ux, uy, uz, depth);
            case 1:  return evalLeaf<1>(coeffs, ux, uy, uz, depth);
            case 2:  return evalLeaf<2>(coeffs, ux, uy, uz, depth);
            case 3:  return evalLeaf<3>(coeffs, ux, uy, uz, depth);
            case 4:  return evalLeaf<4>(coeffs, ux, uy, uz, depth);
            case 5:  return evalLeaf<5>(coeffs, ux, uy, uz, depth);
            case 6:  return evalLeaf<6>(coeffs, ux, uy, uz, depth);
            case 7:  return evalLeaf<7>(coeffs, ux, uy, uz, depth);
            case 8:  return evalLeaf<8>(coeffs, ux, uy, uz, depth);
            case 9:  return evalLeaf<9>(coeffs, ux, uy, uz, depth);
            case 10: return evalLeaf<10>(coeffs, ux, uy, uz, depth);
            case 11: return evalLeaf<11>(coeffs, ux, uy, uz, depth);
            default: return evalLeaf<12>(coeffs, ux, uy, uz, depth);
        }
    }

    // Descent + evaluation. `top` is the 4096-entry depth-4 table (global or shared), or nullptr to start at the root.
    __device__ __forceinline__ double queryPoint(const QNode* __restrict__ nodes, const double* __restrict__ coeffs,
                                                 const uint32_t* top, const RootMap& map, double x, double y, double z)
    {
        const double px = (x - map.centre[0]) * map.invSizes[0];          // Octree.cpp:665
        const double py = (y - map.centre[1]) * map.invSizes[1];
        const double pz = (z - map.centre[2]) * map.invSizes[2];
        const float fx = (float)px, fy = (float)py, fz = (float)pz;       // Octree.cpp:668: contains() on the f32 cast, inclusive
        if (!(fx >= -0.5f && fx <= 0.5f && fy >= -0.5f && fy <= 0.5f && fz >= -0.5f && fz <= 0.5f)) return DBL_MAX;

        double cx = 0.0, cy = 0.0, cz = 0.0, q = 0.25;                    // centre of the current node, quarter of its size
        uint32_t cur = 0;
        if (top)
        {
            uint32_t code = 0;
            #pragma unroll
            for (int l = 0; l < kCoarseDepth; ++l)
            {
                const uint32_t bx = px >= cx, by = py >= cy, bz = pz >= cz;         // Octree.cpp:681-683
                cx += bx ? q : -q; cy += by ? q : -q; cz += bz ? q : -q; q *= 0.5;
                code = (code << 1) | bx | (by << 4) | (bz << 8);                    // x bits in [0,4), y in [4,8), z in [8,12)
            }
            // code now holds the 4 x-bits, 4 y-bits (<<4), 4 z-bits (<<8), each MSB first
            cur = top[code & 0xFFF];
        }
        uint4 raw = __ldg(reinterpret_cast<const uint4*>(nodes) + cur);
        while (raw.z == kInternalTag)                                               // Octree.cpp:687
        {
            const uint32_t bx = px >= cx, by = py >= cy, bz = pz >= cz;
            cx += bx ? q : -q; cy += by ? q : -q; cz += bz ? q : -q; q *= 0.5;
            cur = raw.x + bx + (by << 1) + (bz << 2);                               // Octree.cpp:685
            raw = __ldg(reinterpret_cast<const uint4*>(nodes) + cur);
        }
        const int depth = (int)raw.w;
        const double scale = (double)(2u << depth);                                 // Octree.cpp:862
        return evalLeafAnyDegree(coeffs + raw.y, (int)raw.z, (px - cx) * scale, (py - cy) * scale, (pz - cz) * scale, depth);
    }

    __device__ __forceinline__ double treeQuery(const DeviceTreeView* tree, double x, double y, double z)
    {
        return queryPoint(tree->nodes, tree->coeffs, tree->top, tree->map, x, y, z);
    }
}
