// mesh.cpp — device-resident triangle mesh as an SDF source: what replaces Meshing::Mesh + Meshing::BVH
// (Include/Meshing/Mesh.h, Include/Meshing/BVH.h) under Octree::Create.
//
// Host side, at hpsdf_mesh_create:
//   * half-edge twins            Mesh::CreateHalfEdges (Source/Meshing/Mesh.cpp:87-131); an unpaired edge -> HPSDF_ERR_MESH
//   * pseudonormals, float32     PseudoNormalFace / Edge / Vertex (Mesh.cpp:185-242), precomputed for every face, every
//                                half-edge and every triangle corner with the reference's formulas and operation order
//                                (this file is compiled with -ffp-contract=off; acosf is the host libm's, as in the reference)
//   * BVH                        top-down median split on the longest axis of the centroid bounds, <= 4 triangles per leaf.
//                                The reference's bottom-up pairing through an NNOctree (BVH.cpp:26-260) is not reproduced:
//                                any BVH gives the same closest triangle.
// Device side: mesh_eval.cuh.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <numeric>
#include <vector>
#include "octree.h"

struct hpsdf_mesh
{
    int      device = 0;
    void*    blob = nullptr;                  // nodes | triVerts | pseudo | view
    hpsdf::DeviceMeshView  view{};
    hpsdf::DeviceMeshView* dView = nullptr;
    float    mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
    uint32_t nTris = 0, nVerts = 0;
};

namespace hpsdf
{
    cudaError_t launchMeshDistance(const DeviceMeshView* dView, const float* dXyz, size_t n, float* dOut, cudaStream_t stream);   // kernels.cu

    const DeviceMeshView* meshDeviceView(const hpsdf_mesh* m) { return m ? m->dView : nullptr; }
    int meshDevice(const hpsdf_mesh* m) { return m ? m->device : -1; }

    namespace
    {
        struct V3 { float x, y, z; };
        inline V3 sub(const V3& a, const V3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
        inline V3 add(const V3& a, const V3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
        inline V3 mul(const V3& a, float s) { return { a.x * s, a.y * s, a.z * s }; }
        inline float dot(const V3& a, const V3& b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }           // a0 + (a1 + a2)
        inline V3 cross(const V3& a, const V3& b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
        inline V3 normalized(const V3& a)
        {
            const float z = dot(a, a);
            if (z > 0.0f) { const float n = std::sqrt(z); return { a.x / n, a.y / n, a.z / n }; }
            return a;
        }

        struct Builder
        {
            const std::vector<V3>& v;
            const std::vector<uint32_t>& tri;
            std::vector<uint32_t> he;

            V3 vert(uint32_t t, uint32_t k) const { return v[tri[3 * t + k]]; }
            // PseudoNormalFace, Mesh.cpp:185-193
            V3 faceNormal(uint32_t t) const
            {
                const V3 ab = sub(vert(t, 1), vert(t, 0)), ac = sub(vert(t, 2), vert(t, 0));
                return normalized(cross(ab, ac));
            }
            // PseudoNormalEdge, Mesh.cpp:196-213: n = nA * PI + nB * PI (PI converted to float), normalised
            V3 edgeNormal(uint32_t t, uint32_t s) const
            {
                const uint32_t adjEdge = he[3 * t + s];
                const uint32_t adjTri = (adjEdge - (adjEdge % 3)) / 3;
                const float PI = (float)3.14159265359;
                return normalized(add(mul(faceNormal(t), PI), mul(faceNormal(adjTri), PI)));
            }
            // PseudoNormalVertex, Mesh.cpp:216-242: walk the fan he -> next(twin(he)) until back at the start triangle
            V3 vertexNormal(uint32_t t, uint32_t s) const
            {
                V3 n = { 0.0f, 0.0f, 0.0f };
                uint32_t curHE = 3 * t + s, curTri = t;
                size_t guard = 0;
                do
                {
                    const V3 c0 = vert(curTri, curHE % 3), c1 = vert(curTri, (curHE + 1) % 3), c2 = vert(curTri, (curHE + 2) % 3);
                    const V3 ab = sub(c1, c0), ac = sub(c2, c0);
                    const float ang = acosf(dot(normalized(ab), normalized(ac)));
                    n = add(n, mul(faceNormal(curTri), ang));
                    curHE = he[curHE];
                    curHE = ((curHE % 3) == 2) ? (curHE - 2) : (curHE + 1);
                    curTri = (curHE - (curHE % 3)) / 3;
                } while (curTri != t && ++guard < 100000);
                return normalized(n);
            }
        };

        struct BuildNode { float mn[3], mx[3]; uint32_t a, b, begin, end; };

        // static partition of [0, n) over the host cores (mesh set-up is per-triangle / per-node independent work)
        template <typename F>
        void parallelFor(size_t n, F&& body)
        {
            const size_t nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
            if (nt <= 1 || n < 4096) { body((size_t)0, n); return; }
            std::vector<std::thread> th;
            for (size_t t = 0; t < nt; ++t)
            {
                const size_t b = n * t / nt, e = n * (t + 1) / nt;
                th.emplace_back([&body, b, e] { body(b, e); });
            }
            for (std::thread& x : th) x.join();
        }

        void triBounds(const std::vector<V3>& v, const std::vector<uint32_t>& tri, uint32_t t, float mn[3], float mx[3])
        {
            for (int k = 0; k < 3; ++k)
            {
                const V3& p = v[tri[3 * t + k]];
                const float c[3] = { p.x, p.y, p.z };
                for (int d = 0; d < 3; ++d) { mn[d] = k ? std::min(mn[d], c[d]) : c[d]; mx[d] = k ? std::max(mx[d], c[d]) : c[d]; }
            }
        }

        // nodes of the median-split tree over m triangles (leaf <= 4): known up front, so subtrees can be built in parallel
        // into their final index ranges (node, then its left subtree, then its right subtree — the order a serial build gives)
        uint32_t bvhNodeCount(uint32_t m) { return m <= 4 ? 1u : 1u + bvhNodeCount(m / 2) + bvhNodeCount(m - m / 2); }

        void buildBvh(BuildNode* nodes, uint32_t idx, uint32_t* order, const float* cen, const float* tmn, const float* tmx,
                      uint32_t begin, uint32_t end, int depth)
        {
            float mn[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, mx[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
            float cmn[3] = { 3.4e38f, 3.4e38f, 3.4e38f }, cmx[3] = { -3.4e38f, -3.4e38f, -3.4e38f };
            for (uint32_t i = begin; i < end; ++i)
            {
                const uint32_t t = order[i];
                for (int d = 0; d < 3; ++d)
                {
                    mn[d] = std::min(mn[d], tmn[3 * t + d]); mx[d] = std::max(mx[d], tmx[3 * t + d]);
                    cmn[d] = std::min(cmn[d], cen[3 * t + d]); cmx[d] = std::max(cmx[d], cen[3 * t + d]);
                }
            }
            memcpy(nodes[idx].mn, mn, 12); memcpy(nodes[idx].mx, mx, 12);
            nodes[idx].begin = begin; nodes[idx].end = end;
            if (end - begin <= 4)            // at most 7 fit the 3-bit count of the wide tree; 1 / 2 / 4 / 7 measured 98 / 90 / 83 / 80 ms on the binary walk, 4 best on the 4-wide one
            {
                nodes[idx].a = begin; nodes[idx].b = 0x80000000u | (end - begin);
                return;
            }
            int axis = 0;
            if (cmx[1] - cmn[1] > cmx[axis] - cmn[axis]) axis = 1;
            if (cmx[2] - cmn[2] > cmx[axis] - cmn[axis]) axis = 2;
            const uint32_t mid = (begin + end) / 2;
            std::nth_element(order + begin, order + mid, order + end,
                             [&](uint32_t x, uint32_t y) { return cen[3 * x + axis] < cen[3 * y + axis] || (cen[3 * x + axis] == cen[3 * y + axis] && x < y); });
            const uint32_t l = idx + 1, r = idx + 1 + bvhNodeCount(mid - begin);
            nodes[idx].a = l; nodes[idx].b = r;
            if (depth < 4 && end - begin > 65536)
            {
                // the two halves touch disjoint ranges of `order` and of `nodes`: the left one gets its own thread (16 at depth 4)
                std::thread left([=] { buildBvh(nodes, l, order, cen, tmn, tmx, begin, mid, depth + 1); });
                buildBvh(nodes, r, order, cen, tmn, tmx, mid, end, depth + 1);
                left.join();
            }
            else
            {
                buildBvh(nodes, l, order, cen, tmn, tmx, begin, mid, depth + 1);
                buildBvh(nodes, r, order, cen, tmn, tmx, mid, end, depth + 1);
            }
        }

        // Oriented boxes. The axis-aligned box of a slanted patch of surface overhangs it by about its own size L, so a
        // point at distance d sees ~2 pi d / L boxes of that size inside its search sphere at EVERY level of the tree
        // (hundreds of leaves for d = 0.1 on a 870 k-triangle mesh). In the frame of the patch's mean normal the overhang
        // is the patch's deviation from its plane (~L^2 / 8R), and the count per level drops to O(1). Per node: centre c,
        // orthonormal axes u, v, n (n = area-weighted mean normal) and half extents, inflated so that float32 rounding of
        // the frame and of the device-side test can never cut a triangle off. Nodes whose normals disagree (|sum n_i A_i|
        // < 0.5 sum A_i) or that hold more than 32 768 triangles keep only their axis-aligned box (half extents = FLT_MAX).
        void computeObbs(const std::vector<BuildNode>& nodes, const std::vector<uint32_t>& order, const std::vector<V3>& v,
                         const std::vector<uint32_t>& tri, std::vector<float>& obb)
        {
            obb.assign(16 * nodes.size(), 0.0f);
            double diag2 = 0.0;
            for (int d = 0; d < 3; ++d) diag2 += ((double)nodes[0].mx[d] - nodes[0].mn[d]) * ((double)nodes[0].mx[d] - nodes[0].mn[d]);
            double maxAbs = 0.0;
            for (int d = 0; d < 3; ++d) maxAbs = std::max(maxAbs, std::max(std::fabs((double)nodes[0].mn[d]), std::fabs((double)nodes[0].mx[d])));
            // float32 projection error on the device is ~1e-7 |p|; query points live in a root box of about the mesh's size
            const double inflate = 1e-5 * std::max(std::sqrt(diag2), maxAbs);
            parallelFor(nodes.size(), [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i)
            {
                float* o = &obb[16 * i];
                o[3] = o[7] = o[11] = 3.402823466e+38f;
                const BuildNode& nd = nodes[i];
                if (nd.end - nd.begin > 32768u) continue;
                double N[3] = { 0, 0, 0 }, total = 0.0;
                for (uint32_t k = nd.begin; k < nd.end; ++k)
                {
                    const uint32_t t = order[k];
                    const V3 &a = v[tri[3 * t]], &b = v[tri[3 * t + 1]], &c = v[tri[3 * t + 2]];
                    const double e1[3] = { (double)b.x - a.x, (double)b.y - a.y, (double)b.z - a.z }, e2[3] = { (double)c.x - a.x, (double)c.y - a.y, (double)c.z - a.z };
                    const double cr[3] = { e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0] };
                    N[0] += cr[0]; N[1] += cr[1]; N[2] += cr[2];
                    total += std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
                }
                const double len = std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
                if (!(len > 0.5 * total) || !(len > 0.0)) continue;
                const double n[3] = { N[0] / len, N[1] / len, N[2] / len };
                int ax = 0;
                if (std::fabs(n[1]) < std::fabs(n[ax])) ax = 1;
                if (std::fabs(n[2]) < std::fabs(n[ax])) ax = 2;
                double u[3] = { -n[ax] * n[0], -n[ax] * n[1], -n[ax] * n[2] };
                u[ax] += 1.0;
                const double ul = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
                for (int d = 0; d < 3; ++d) u[d] /= ul;
                const double w[3] = { n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0] };
                // the frame as the device will see it (float32), so the extents are measured in exactly that frame
                const float fu[3] = { (float)u[0], (float)u[1], (float)u[2] }, fw[3] = { (float)w[0], (float)w[1], (float)w[2] }, fn[3] = { (float)n[0], (float)n[1], (float)n[2] };
                double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
                for (uint32_t k = nd.begin; k < nd.end; ++k)
                {
                    const uint32_t t = order[k];
                    for (int c = 0; c < 3; ++c)
                    {
                        const V3& p = v[tri[3 * t + c]];
                        const double q[3] = { (double)p.x * fu[0] + (double)p.y * fu[1] + (double)p.z * fu[2],
                                              (double)p.x * fw[0] + (double)p.y * fw[1] + (double)p.z * fw[2],
                                              (double)p.x * fn[0] + (double)p.y * fn[1] + (double)p.z * fn[2] };
                        for (int d = 0; d < 3; ++d) { lo[d] = std::min(lo[d], q[d]); hi[d] = std::max(hi[d], q[d]); }
                    }
                }
                // centre in frame coordinates (the device projects p onto the axes and subtracts these)
                o[0] = (float)(0.5 * (lo[0] + hi[0])); o[1] = (float)(0.5 * (lo[1] + hi[1])); o[2] = (float)(0.5 * (lo[2] + hi[2]));
                for (int d = 0; d < 3; ++d)
                {
                    const double c = (double)o[d];
                    const double he = std::max(hi[d] - c, c - lo[d]) + inflate;
                    o[3 + 4 * d] = std::nextafter((float)he, 3.402823466e+38f);
                }
                o[4] = fu[0]; o[5] = fu[1]; o[6] = fu[2];
                o[8] = fw[0]; o[9] = fw[1]; o[10] = fw[2];
                o[12] = fn[0]; o[13] = fn[1]; o[14] = fn[2];
            }
            });
        }

        // 4-wide collapse for meshSampleKernel: a wide node holds the grandchildren of a binary node (or its children where
        // they are leaves), so a query needs half as many dependent memory round trips. 68 floats per node: the 4 child
        // references (kWideNone | leaf: 0x80000000 | count << 28 | first slot | index of a wide node), then per child one
        // oriented box (a 256-byte record with the references folded into the boxes' padding measured 5-12 % SLOWER: the
        // power-of-two stride maps the nodes onto half of the L1 sets); nodes without one (large or with incoherent normals) get their axis-aligned box in the same form.
        constexpr uint32_t kWideNone = 0xFFFFFFFFu;
        uint32_t collapseWide(const std::vector<BuildNode>& bn, const std::vector<float>& obb, std::vector<float>& wide, uint32_t b, double inflate)
        {
            const uint32_t idx = (uint32_t)(wide.size() / 68);
            wide.resize(wide.size() + 68, 0.0f);
            uint32_t kids[4], nk = 0;
            const uint32_t two[2] = { bn[b].a, bn[b].b };
            for (int s = 0; s < 2; ++s)
            {
                const BuildNode& c = bn[two[s]];
                if (c.b & 0x80000000u) kids[nk++] = two[s];
                else { kids[nk++] = c.a; kids[nk++] = c.b; }
            }
            for (uint32_t k = 0; k < 4; ++k)
            {
                uint32_t ref = kWideNone;
                float box[16] = { 0 };
                box[3] = box[7] = box[11] = -3.0e38f;                 // absent child: |q| - he overflows to +inf
                if (k < nk)
                {
                    const BuildNode& c = bn[kids[k]];
                    const float* o = &obb[16 * (size_t)kids[k]];
                    if (o[3] < 1e38f) memcpy(box, o, 64);
                    else
                    {
                        for (int d = 0; d < 3; ++d)
                        {
                            const double lo = c.mn[d], hi = c.mx[d];
                            box[d] = (float)(0.5 * (lo + hi));
                            const double ce = (double)box[d];
                            box[3 + 4 * d] = std::nextafter((float)(std::max(hi - ce, ce - lo) + inflate), 3.402823466e+38f);
                        }
                        box[4] = 1.0f; box[9] = 1.0f; box[14] = 1.0f;       // u = x, v = y, n = z
                    }
                    if (c.b & 0x80000000u) ref = 0x80000000u | ((c.b & 7u) << 28) | c.a;
                    else ref = collapseWide(bn, obb, wide, kids[k], inflate);
                }
                memcpy(&wide[68 * (size_t)idx + k], &ref, 4);
                memcpy(&wide[68 * (size_t)idx + 4 + 16 * k], box, 64);
            }
            return idx;
        }
    }
}

using namespace hpsdf;

extern "C"
{
    HPSDF_API hpsdf_status hpsdf_mesh_create(const float* vertices, size_t n_vertices, const uint32_t* tri_indices, size_t n_tris,
                                             int device, hpsdf_mesh** out)
    {
        if (!out) { setLastError("out is null"); return HPSDF_ERR_INVALID_ARG; }
        *out = nullptr;
        if (!vertices || !tri_indices || n_vertices == 0 || n_tris == 0 || n_tris >= 0x10000000ull)      // leaf slots are packed into 28 bits (mesh_sample_kernel.cuh)
        { setLastError("mesh arrays are null, empty or too large"); return HPSDF_ERR_INVALID_ARG; }
        std::string err;
        DeviceCtx* ctx = getDeviceCtx(device, err);
        if (!ctx) { setLastError(err); return HPSDF_ERR_NO_DEVICE; }

        const bool dbg = getenv("HPSDF_DEBUG_MESH") != nullptr;
        double tS = nowMs();
        auto stage = [&](const char* what) { if (dbg) { const double t = nowMs(); fprintf(stderr, "mesh_create: %-14s %.1f ms\n", what, t - tS); tS = t; } };
        std::vector<V3> v(n_vertices);
        memcpy(v.data(), vertices, n_vertices * 12);
        std::vector<uint32_t> tri(tri_indices, tri_indices + 3 * n_tris);
        for (uint32_t i : tri) if (i >= n_vertices) { setLastError("triangle index out of range"); return HPSDF_ERR_INVALID_ARG; }

        // CreateHalfEdges (Mesh.cpp:87-131): twin of the directed edge (a, b) is the edge (b, a); first occurrence wins
        Builder b{ v, tri, std::vector<uint32_t>(3 * n_tris, 0xFFFFFFFFu) };
        {
            // same find / insert sequence as the reference's std::map (Mesh.cpp:96-118): "is the reverse edge known? pair them :
            // remember this edge unless an equal one is already known", on a flat linear-probing table of packed (from, to) keys
            size_t cap = 1;
            while (cap < 6 * n_tris) cap <<= 1;                              // load factor <= 0.5
            const uint64_t kEmpty = ~0ull;                                   // (0xFFFFFFFF, 0xFFFFFFFF) is not an edge: indices are < n_vertices
            std::vector<uint64_t> keys(cap, kEmpty);
            std::vector<uint32_t> vals(cap);
            auto slotOf = [&](uint64_t key) -> size_t
            {
                size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
                while (keys[h] != kEmpty && keys[h] != key) h = (h + 1) & (cap - 1);
                return h;
            };
            for (uint32_t i = 0; i < 3 * n_tris; ++i)
            {
                const uint64_t from = tri[i], to = (i % 3 == 2) ? tri[i - 2] : tri[i + 1];
                const size_t f = slotOf(to << 32 | from);
                if (keys[f] != kEmpty) { b.he[vals[f]] = i; b.he[i] = vals[f]; }
                else
                {
                    const size_t g = slotOf(from << 32 | to);
                    if (keys[g] == kEmpty) { keys[g] = from << 32 | to; vals[g] = i; }
                }
            }
            for (uint32_t h : b.he)
                if (h == 0xFFFFFFFFu) { setLastError("mesh has an edge without a twin: not a closed manifold (Mesh.cpp:121-128)"); return HPSDF_ERR_MESH; }
        }

        stage("half-edges");
        // pseudonormals: 21 floats per triangle
        std::vector<float> pseudo(21 * n_tris);
        parallelFor(n_tris, [&](size_t t0, size_t t1)
        {
            for (uint32_t t = (uint32_t)t0; t < (uint32_t)t1; ++t)
            {
                float* p = pseudo.data() + 21 * (size_t)t;
                const V3 f = b.faceNormal(t);
                p[0] = f.x; p[1] = f.y; p[2] = f.z;
                for (uint32_t s = 0; s < 3; ++s)
                {
                    const V3 e = b.edgeNormal(t, s), w = b.vertexNormal(t, s);
                    p[3 + 3 * s] = e.x; p[4 + 3 * s] = e.y; p[5 + 3 * s] = e.z;
                    p[12 + 3 * s] = w.x; p[13 + 3 * s] = w.y; p[14 + 3 * s] = w.z;
                }
            }
        });

        stage("pseudonormals");
        // BVH
        std::vector<float> cen(3 * n_tris), tmn(3 * n_tris), tmx(3 * n_tris);
        for (uint32_t t = 0; t < n_tris; ++t)
        {
            triBounds(v, tri, t, &tmn[3 * t], &tmx[3 * t]);
            for (int d = 0; d < 3; ++d) cen[3 * t + d] = 0.5f * (tmn[3 * t + d] + tmx[3 * t + d]);
        }
        std::vector<uint32_t> order(n_tris);
        std::iota(order.begin(), order.end(), 0u);
        std::vector<BuildNode> bn(bvhNodeCount((uint32_t)n_tris));
        buildBvh(bn.data(), 0, order.data(), cen.data(), tmn.data(), tmx.data(), 0, (uint32_t)n_tris, 0);
        std::vector<BvhNode> nodes(bn.size());
        for (size_t i = 0; i < bn.size(); ++i)
        {
            memcpy(nodes[i].mn, bn[i].mn, 12); memcpy(nodes[i].mx, bn[i].mx, 12);
            nodes[i].a = bn[i].a; nodes[i].b = bn[i].b;
        }
        stage("bvh");
        std::vector<float> obb;
        computeObbs(bn, order, v, tri, obb);
        std::vector<float> wide;
        wide.reserve(68 * (bn.size() / 2 + 2));
        {
            double diag2 = 0.0, maxAbs = 0.0;
            for (int d = 0; d < 3; ++d)
            {
                diag2 += ((double)bn[0].mx[d] - bn[0].mn[d]) * ((double)bn[0].mx[d] - bn[0].mn[d]);
                maxAbs = std::max(maxAbs, std::max(std::fabs((double)bn[0].mn[d]), std::fabs((double)bn[0].mx[d])));
            }
            const double inflate = 1e-5 * std::max(std::sqrt(diag2), maxAbs);
            if (bn[0].b & 0x80000000u)
            {
                // the whole mesh is one leaf: a root with that single child
                wide.assign(68, 0.0f);
                const uint32_t none = kWideNone, ref = 0x80000000u | ((bn[0].b & 7u) << 28) | bn[0].a;
                for (int k = 0; k < 4; ++k) { memcpy(&wide[k], k ? &none : &ref, 4); wide[4 + 16 * k + 3] = wide[4 + 16 * k + 7] = wide[4 + 16 * k + 11] = k ? -3.0e38f : 3.0e38f; }
            }
            else collapseWide(bn, obb, wide, 0, inflate);
        }
        stage("oriented boxes");
        std::vector<float> tv(12 * n_tris);
        for (uint32_t slot = 0; slot < n_tris; ++slot)
        {
            const uint32_t t = order[slot];
            for (int k = 0; k < 3; ++k)
            {
                const V3 p = v[tri[3 * t + k]];
                float* o = &tv[12 * (size_t)slot + 4 * k];
                o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = 0.0f;
            }
            memcpy(&tv[12 * (size_t)slot + 3], &t, 4);          // original triangle index rides in a.w
        }

        hpsdf_mesh* m = new hpsdf_mesh();
        m->device = ctx->device; m->nTris = (uint32_t)n_tris; m->nVerts = (uint32_t)n_vertices;
        memcpy(m->mn, bn[0].mn, 12); memcpy(m->mx, bn[0].mx, 12);          // CalculateMeshAABB (Mesh.cpp:66-84)
        auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
        const size_t bNodes = align(nodes.size() * sizeof(BvhNode)), bTv = align(tv.size() * 4), bPs = align(pseudo.size() * 4), bObb = align(obb.size() * 4), bWide = align(wide.size() * 4);
        cudaError_t e = cudaMalloc(&m->blob, bNodes + bTv + bPs + bObb + bWide + 256);
        if (e != cudaSuccess) { delete m; return failCuda(e, "mesh allocation"); }
        char* p = (char*)m->blob;
        m->view.nodes = (const BvhNode*)p;
        m->view.triVerts = (const void*)(p + bNodes);
        m->view.pseudo = (const float*)(p + bNodes + bTv);
        m->view.nTris = (uint32_t)n_tris; m->view.nNodes = (uint32_t)nodes.size();
        m->view.obb = (const void*)(p + bNodes + bTv + bPs);
        m->view.wide = (const void*)(p + bNodes + bTv + bPs + bObb);
        m->dView = (DeviceMeshView*)(p + bNodes + bTv + bPs + bObb + bWide);
        e = cudaMemcpy(p, nodes.data(), nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p + bNodes, tv.data(), tv.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p + bNodes + bTv, pseudo.data(), pseudo.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p + bNodes + bTv + bPs, obb.data(), obb.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(p + bNodes + bTv + bPs + bObb, wide.data(), wide.size() * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(m->dView, &m->view, sizeof(DeviceMeshView), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(m->blob); delete m; return failCuda(e, "mesh upload"); }
        stage("upload");
        *out = m;
        return HPSDF_OK;
    }

    HPSDF_API hpsdf_status hpsdf_mesh_signed_distance(const hpsdf_mesh* mesh, const float* xyz, size_t n, float* out)
    {
        if (!mesh || (n && (!xyz || !out))) { setLastError("null pointer"); return HPSDF_ERR_INVALID_ARG; }
        if (!n) return HPSDF_OK;
        HPSDF_CUDA(cudaSetDevice(mesh->device));
        float *dIn = nullptr, *dOut = nullptr;
        HPSDF_CUDA(cudaMalloc((void**)&dIn, n * 12));
        cudaError_t e = cudaMalloc((void**)&dOut, n * 4);
        if (e == cudaSuccess) e = cudaMemcpy(dIn, xyz, n * 12, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = launchMeshDistance(mesh->dView, dIn, n, dOut, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(out, dOut, n * 4, cudaMemcpyDeviceToHost);
        cudaFree(dIn); cudaFree(dOut);
        if (e != cudaSuccess) return failCuda(e, "hpsdf_mesh_signed_distance");
        return HPSDF_OK;
    }

    HPSDF_API hpsdf_status hpsdf_mesh_aabb(const hpsdf_mesh* mesh, float mn[3], float mx[3])
    {
        if (!mesh || !mn || !mx) { setLastError("null pointer"); return HPSDF_ERR_INVALID_ARG; }
        memcpy(mn, mesh->mn, 12); memcpy(mx, mesh->mx, 12);
        return HPSDF_OK;
    }

    HPSDF_API void hpsdf_mesh_destroy(hpsdf_mesh* mesh)
    {
        if (!mesh) return;
        cudaSetDevice(mesh->device);
        cudaFree(mesh->blob);
        delete mesh;
    }
}
