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

    // Warp-uniform evaluation: all lanes run the instantiation of the warp's LARGEST leaf degree and switch whole shells
    // off with a predicate (p <= myDeg), so lanes whose leaves have different degrees do not serialise (ncu showed 17 of
    // 32 lanes active per instruction when each lane branched to its own degree). Lanes without a leaf pass myDeg = -1.
    template <int DEG>
    __device__ __forceinline__ double evalLeafMasked(const double* __restrict__ coeffs, int myDeg, double ux, double uy, double uz, int depth)
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
        const double2* c2 = reinterpret_cast<const double2*>(coeffs);
        double f = 0.0;
        int idx = 0;
        double2 cur = make_double2(0.0, 0.0);
        #pragma unroll
        for (int p = 0; p <= DEG; ++p)
        {
            const bool on = p <= myDeg;
            #pragma unroll
            for (int i = 0; i <= p; ++i)
            {
                #pragma unroll
                for (int j = 0; j <= p - i; ++j)
                {
                    const int k = p - i - j;
                    // the reference stores 83 coefficients for a degree-6 leaf: (6,0,0) exists only from degree 7 on
                    const bool term = on && !(idx == 83 && myDeg == 6);
                    if ((idx & 1) == 0) { if (on) cur = __ldg(c2 + (idx >> 1)); }
                    const double cv = (idx & 1) ? cur.y : cur.x;
                    if (term) f = fma(cv, (lx[i] * ly[j]) * lz[k], f);                                  // Octree.cpp:891-897
                    ++idx;
                }
            }
        }
        return f;
    }

    // Degrees above 6 are rare: a loop over the basis table with the Legendre values in local memory keeps the register
    // count of the kernel set by the common degrees.
    __device__ __noinline__ double evalLeafGeneric(const double* __restrict__ coeffs, int degree, double ux, double uy, double uz,
                                                   int depth, const uint32_t* __restrict__ bidx)
    {
        double l[3][kMaxDegree + 1];
        const double u[3] = { ux, uy, uz };
        for (int a = 0; a < 3; ++a)
        {
            l[a][0] = c_nl[0][depth];
            double m2 = 0.0, m1 = 1.0;
            for (int j = 1; j <= degree; ++j)
            {
                const double v = c_rec[j][0] * u[a] * m1 - c_rec[j][1] * m2; m2 = m1; m1 = v;
                l[a][j] = v * c_nl[j][depth];
            }
        }
        double f = 0.0;
        const int n = coeffCount(degree);
        for (int i = 0; i < n; ++i)
        {
            const uint32_t abc = __ldg(bidx + i);
            f = fma(coeffs[i], (l[0][abc & 0xFF] * l[1][(abc >> 8) & 0xFF]) * l[2][(abc >> 16) & 0xFF], f);
        }
        return f;
    }

    struct LeafHit
    {
        const double* coeffs;
        double ux, uy, uz;
        int    degree;        // -1: the point is outside the root
        int    depth;
    };

    // Octree::Query up to the leaf: root map, f32 containment test, descent (Octree.cpp:662-702).
    __device__ __forceinline__ LeafHit findLeaf(const QNode* __restrict__ nodes, const double* __restrict__ coeffs,
                                                const uint32_t* top, const RootMap& map, double x, double y, double z)
    {
        LeafHit h;
        h.coeffs = coeffs; h.ux = h.uy = h.uz = 0.0; h.degree = -1; h.depth = 0;
        const double px = (x - map.centre[0]) * map.invSizes[0];          // Octree.cpp:665
        const double py = (y - map.centre[1]) * map.invSizes[1];
        const double pz = (z - map.centre[2]) * map.invSizes[2];
        const float fx = (float)px, fy = (float)py, fz = (float)pz;       // Octree.cpp:668: contains() on the f32 cast, inclusive
        if (!(fx >= -0.5f && fx <= 0.5f && fy >= -0.5f && fy <= 0.5f && fz >= -0.5f && fz <= 0.5f)) return h;
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
                code = (code << 1) | bx | (by << 4) | (bz << 8);
            }
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
        const double scale = (double)(2u << raw.w);                                 // Octree.cpp:862
        h.coeffs = coeffs + raw.y; h.degree = (int)raw.z; h.depth = (int)raw.w;
        h.ux = (px - cx) * scale; h.uy = (py - cy) * scale; h.uz = (pz - cz) * scale;
        return h;
    }

    // Evaluate the hits of a whole warp (all 32 lanes must call this).
    __device__ __forceinline__ double evalWarp(const LeafHit& h, const uint32_t* __restrict__ bidx)
    {
        const int maxDeg = (int)__reduce_max_sync(0xFFFFFFFFu, (unsigned)(h.degree < 0 ? 0 : h.degree));
        double v;
        switch (maxDeg)
        {
            case 0:  v = evalLeafMasked<0>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 1:  v = evalLeafMasked<1>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 2:  v = evalLeafMasked<2>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 3:  v = evalLeafMasked<3>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 4:  v = evalLeafMasked<4>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 5:  v = evalLeafMasked<5>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            case 6:  v = evalLeafMasked<6>(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth); break;
            default: v = h.degree < 0 ? 0.0 : evalLeafGeneric(h.coeffs, h.degree, h.ux, h.uy, h.uz, h.depth, bidx); break;
        }
        return h.degree < 0 ? DBL_MAX : v;                                          // Octree.cpp:668-671
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
