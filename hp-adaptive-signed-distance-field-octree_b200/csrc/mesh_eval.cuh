// mesh_eval.cuh — float32 mesh signed distance on the device.
//
// Reference: Mesh::SignedDistanceAtPt (Source/Meshing/Mesh.cpp:54-63) = closest triangle through the BVH
// (BVH::ClosestTriangleToPt, Source/Meshing/BVH.cpp:263-342) + ClosestSimplexToPt (Source/Meshing/Utility.cpp:5-97, Ericson's
// region test with EPSILON_F32 thresholds) + angle-weighted pseudonormal of the hit face / edge / vertex
// (Mesh.cpp:162-242) for the sign; everything in float32.
//
// A 1-ulp difference in a float32 distance (6e-8 relative) would move fitted coefficients by far more than the 1e-10
// parity bar, so the arithmetic that produces the returned value is mirrored operation by operation with round-to-nearest
// intrinsics (__fadd_rn / __fmul_rn / __fdiv_rn / __fsqrt_rn are never contracted into FMAs), in the association order of
// the CPU checker (3-term sums as a0 + (a1 + a2)). The pseudonormals are precomputed on the host with the reference's
// formulas (mesh.cpp); the BVH itself is free to differ (any BVH yields the same closest triangle): here a host-built
// median-split binary tree with up to 4 triangles per leaf, every node bounded by an axis-aligned AND an oriented box,
// traversed with a small per-thread stack, nearer child first.
// Ties in squared distance go to the lower triangle index (the brute-force order of Mesh.cpp:134-159).
#pragma once
#include "hp_common.h"

namespace hpsdf
{
    struct F3 { float x, y, z; };
    __device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
    __device__ __forceinline__ F3 sub3(const F3& a, const F3& b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
    __device__ __forceinline__ F3 add3(const F3& a, const F3& b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
    __device__ __forceinline__ F3 scale3(const F3& a, float s) { return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
    __device__ __forceinline__ float dot3(const F3& a, const F3& b)
    {
        return __fadd_rn(__fmul_rn(a.x, b.x), __fadd_rn(__fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)));
    }
    __device__ __forceinline__ F3 cross3(const F3& a, const F3& b)
    {
        return f3(__fsub_rn(__fmul_rn(a.y, b.z), __fmul_rn(a.z, b.y)),
                  __fsub_rn(__fmul_rn(a.z, b.x), __fmul_rn(a.x, b.z)),
                  __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
    }

    // ClosestSimplexToPt (Utility.cpp:5-97). simplex: 0 = vertex, 1 = edge, 2 = face; id: vertex A/B/C or edge AB/BC/CA = 0/1/2.
    __device__ __forceinline__ F3 closestSimplex(const F3& pt, const F3& a, const F3& b, const F3& c, int& simplex, int& id)
    {
        constexpr float EPS = 0.000001f;                              // EPSILON_F32, Literals.h:13
        const F3 ab = sub3(b, a), ac = sub3(c, a), bc = sub3(c, b);
        const float snom   = dot3(sub3(pt, a), ab);
        const float sdenom = dot3(sub3(pt, b), sub3(a, b));
        const float tnom   = dot3(sub3(pt, a), ac);
        const float tdenom = dot3(sub3(pt, c), sub3(a, c));
        if (snom < EPS && tnom < EPS) { simplex = 0; id = 0; return a; }
        const float unom   = dot3(sub3(pt, b), bc);
        const float udenom = dot3(sub3(pt, c), sub3(b, c));
        if (sdenom < EPS && unom < EPS) { simplex = 0; id = 1; return b; }
        if (tdenom < EPS && udenom < EPS) { simplex = 0; id = 2; return c; }
        const F3 n = cross3(sub3(b, a), sub3(c, a));
        const float vc = dot3(n, cross3(sub3(a, pt), sub3(b, pt)));
        if (vc < EPS && snom > EPS && sdenom > EPS)
        {
            simplex = 1; id = 0;
            return add3(a, scale3(ab, __fdiv_rn(snom, __fadd_rn(snom, sdenom))));
        }
        const float va = dot3(n, cross3(sub3(b, pt), sub3(c, pt)));
        if (va < EPS && unom > EPS && udenom > EPS)
        {
            simplex = 1; id = 1;
            return add3(b, scale3(bc, __fdiv_rn(unom, __fadd_rn(unom, udenom))));
        }
        const float vb = dot3(n, cross3(sub3(c, pt), sub3(a, pt)));
        if (vb < EPS && tnom > EPS && tdenom > EPS)
        {
            simplex = 1; id = 2;
            return add3(a, scale3(ac, __fdiv_rn(tnom, __fadd_rn(tnom, tdenom))));
        }
        const float sum = __fadd_rn(__fadd_rn(va, vb), vc);
        const float u = __fdiv_rn(va, sum), v = __fdiv_rn(vb, sum);
        const float w = __fsub_rn(__fsub_rn(1.0f, u), v);
        simplex = 2; id = 0;
        return add3(add3(scale3(a, u), scale3(b, v)), scale3(c, w));
    }

    // squared distance from p to the box, a lower bound of the distance to anything inside (ClosestPtOnAABB, Utility.cpp:118-139)
    __device__ __forceinline__ float boxDist2(const BvhNode& n, const F3& p)
    {
        const float dx = fmaxf(fmaxf(n.mn[0] - p.x, p.x - n.mx[0]), 0.0f);
        const float dy = fmaxf(fmaxf(n.mn[1] - p.y, p.y - n.mx[1]), 0.0f);
        const float dz = fmaxf(fmaxf(n.mn[2] - p.z, p.z - n.mx[2]), 0.0f);
        return dx * dx + dy * dy + dz * dz;
    }

    // Closest-triangle state of one query point: winner = (smallest f32 d2, lowest triangle index).
    struct MeshHit
    {
        float    best = 3.402823466e+38f;                               // FLT_MAX, BVH.cpp:279
        uint32_t tri = 0xFFFFFFFFu;
        int      simplex = 2, id = 0;
        F3       pt;
    };

    __device__ __forceinline__ void testLeaf(const float4* __restrict__ tv, uint32_t first, uint32_t cnt, const F3& p, MeshHit& h)
    {
        for (uint32_t k = 0; k < cnt; ++k)
        {
            const float4 A = __ldg(tv + 3 * (first + k)), B = __ldg(tv + 3 * (first + k) + 1), C = __ldg(tv + 3 * (first + k) + 2);
            const uint32_t tri = __float_as_uint(A.w);
            int s, id;
            const F3 cp = closestSimplex(p, f3(A.x, A.y, A.z), f3(B.x, B.y, B.z), f3(C.x, C.y, C.z), s, id);
            const F3 d = sub3(p, cp);
            const float d2 = dot3(d, d);                               // (pt - closestPt).squaredNorm(), BVH.cpp:320
            if (d2 < h.best || (d2 == h.best && tri < h.tri)) { h.best = d2; h.tri = tri; h.simplex = s; h.id = id; h.pt = cp; }
        }
    }

    // sign from the pseudonormal of the hit simplex (Mesh.cpp:162-242, precomputed): face | edge AB, BC, CA | vertex A, B, C
    __device__ __forceinline__ float finishHit(const DeviceMeshView* __restrict__ mesh, const F3& p, const MeshHit& h)
    {
        if (h.tri == 0xFFFFFFFFu) return 3.402823466e+38f;
        const float* pn = mesh->pseudo + 21 * (size_t)h.tri + (h.simplex == 2 ? 0 : h.simplex == 1 ? 3 + 3 * h.id : 12 + 3 * h.id);
        const F3 d = sub3(p, h.pt);
        const float sgn = dot3(f3(__ldg(pn), __ldg(pn + 1), __ldg(pn + 2)), d) > 0.0f ? 1.0f : -1.0f;      // Mesh.cpp:61
        return __fmul_rn(sgn, __fsqrt_rn(dot3(d, d)));                                                      // Mesh.cpp:62
    }

    // Lower bound of the squared distance from p to anything inside node `i`: the larger of the axis-aligned bound and
    // the bound of the node's oriented box (mesh.cpp: frame of the mean normal, extents inflated against float32 rounding).
    // The oriented box is only fetched when the axis-aligned one does not already prune.
    __device__ __forceinline__ float nodeDist2(const BvhNode& n, const float4* __restrict__ obb, uint32_t i, const F3& p, float lim)
    {
        const float d = boxDist2(n, p);
        if (d > lim) return d;
        const float4 o0 = __ldg(obb + 4 * (size_t)i), o1 = __ldg(obb + 4 * (size_t)i + 1), o2 = __ldg(obb + 4 * (size_t)i + 2), o3 = __ldg(obb + 4 * (size_t)i + 3);
        const float qu = fmaxf(fabsf(p.x * o1.x + p.y * o1.y + p.z * o1.z - o0.x) - o0.w, 0.0f);
        const float qv = fmaxf(fabsf(p.x * o2.x + p.y * o2.y + p.z * o2.z - o0.y) - o1.w, 0.0f);
        const float qn = fmaxf(fabsf(p.x * o3.x + p.y * o3.y + p.z * o3.z - o0.z) - o2.w, 0.0f);
        return fmaxf(d, qu * qu + qv * qv + qn * qn);
    }

    // the same two bounds on data already in registers (mesh_sample_kernel.cuh requests everything of a step at once)
    __device__ __forceinline__ float aabbDist2(const float4& lo, const float4& hi, const F3& p)
    {
        const float dx = fmaxf(fmaxf(lo.x - p.x, p.x - hi.x), 0.0f);
        const float dy = fmaxf(fmaxf(lo.y - p.y, p.y - hi.y), 0.0f);
        const float dz = fmaxf(fmaxf(lo.z - p.z, p.z - hi.z), 0.0f);
        return dx * dx + dy * dy + dz * dz;
    }
    __device__ __forceinline__ float obbDist2(const float4& o0, const float4& o1, const float4& o2, const float4& o3, const F3& p)
    {
        const float qu = fmaxf(fabsf(p.x * o1.x + p.y * o1.y + p.z * o1.z - o0.x) - o0.w, 0.0f);
        const float qv = fmaxf(fabsf(p.x * o2.x + p.y * o2.y + p.z * o2.z - o0.y) - o1.w, 0.0f);
        const float qn = fmaxf(fabsf(p.x * o3.x + p.y * o3.y + p.z * o3.z - o0.z) - o2.w, 0.0f);
        return qu * qu + qv * qv + qn * qn;
    }

    // Per-thread traversal, nearer child first, "while-while": the lanes of a warp first all walk inner nodes until each
    // holds a leaf (or is done), then all test their leaf's triangles together — a lane in the 600-instruction triangle
    // test no longer stalls 31 lanes that only want a 30-instruction node step. Pruning is conservative (bounds are
    // evaluated with FMAs: a few ulps of slack). `h` may carry a candidate found earlier.
    __device__ __forceinline__ void traverseThread(const DeviceMeshView* __restrict__ mesh, const F3& p, MeshHit& h)
    {
        const BvhNode* __restrict__ nodes = mesh->nodes;
        const float4* __restrict__ obb = (const float4*)mesh->obb;
        const float4* __restrict__ tv = (const float4*)mesh->triVerts;
        uint32_t stack[48];
        float    stackD[48];
        int sp = 0;
        uint32_t cur = 0;
        bool alive = true;
        while (alive)
        {
            // phase 1: descend until `cur` is a leaf that can still hold something closer
            uint32_t leafFirst = 0, leafCnt = 0;
            for (;;)
            {
                const BvhNode n = nodes[cur];
                if (n.b & 0x80000000u) { leafFirst = n.a; leafCnt = n.b & 0x7FFFFFFFu; break; }
                const float lim = h.best * 1.000001f;
                const float dl = nodeDist2(nodes[n.a], obb, n.a, p, lim), dr = nodeDist2(nodes[n.b], obb, n.b, p, lim);
                const bool goL = dl <= lim, goR = dr <= lim;
                if (goL && goR)
                {
                    const bool leftFirst = dl <= dr;
                    if (sp < 48) { stack[sp] = leftFirst ? n.b : n.a; stackD[sp] = leftFirst ? dr : dl; ++sp; }
                    cur = leftFirst ? n.a : n.b;
                    continue;
                }
                if (goL) { cur = n.a; continue; }
                if (goR) { cur = n.b; continue; }
                bool found = false;
                while (sp > 0)
                {
                    --sp;
                    if (stackD[sp] <= h.best * 1.000001f) { cur = stack[sp]; found = true; break; }
                }
                if (!found) { alive = false; break; }
            }
            // phase 2: the leaf's triangles
            if (alive)
            {
                testLeaf(tv, leafFirst, leafCnt, p, h);
                bool found = false;
                while (sp > 0)
                {
                    --sp;
                    if (stackD[sp] <= h.best * 1.000001f) { cur = stack[sp]; found = true; break; }
                }
                alive = found;
            }
        }
    }

    __device__ __noinline__ float meshSignedDistanceF(const DeviceMeshView* __restrict__ mesh, float px, float py, float pz)
    {
        const F3 p = f3(px, py, pz);
        MeshHit h;
        h.pt = p;
        traverseThread(mesh, p, h);
        return finishHit(mesh, p, h);
    }

    // The double-precision SDF the fit samples: the point is cast to float32, the result widened (the user-side glue the
    // reference implies, Include/Meshing/Mesh.h:53-54).
    __device__ __forceinline__ double meshSignedDistance(const DeviceMeshView* mesh, double x, double y, double z)
    {
        return (double)meshSignedDistanceF(mesh, (float)x, (float)y, (float)z);
    }

    __global__ void meshDistanceKernel(const DeviceMeshView* __restrict__ mesh, const float* __restrict__ xyz, size_t n, float* __restrict__ out)
    {
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < n) out[i] = meshSignedDistanceF(mesh, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    }
}
