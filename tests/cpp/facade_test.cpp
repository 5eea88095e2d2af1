// facade_test.cpp — the reference's own unit tests (Source/Tests/HPUnitTests.cpp:46-316) re-expressed against the C++
// facade include/hpsdf.hpp: same workloads, same acceptance tolerances (|Query - analytic| <= 0.01, <= 0.05 for the SDF
// operations). The lambdas of the reference become SDF::Program objects (the device SDF evaluator).
// Build: g++ -std=c++17 -I include tests/cpp/facade_test.cpp -L <pkg>/lib -lhpsdf ; run on a GPU box.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include "hpsdf.hpp"

using SDF::Vec3d;

static double sphere(const Vec3d& p, double cx, double cy, double cz, double r)
{
    return std::sqrt((p.x() - cx) * (p.x() - cx) + (p.y() - cy) * (p.y() - cy) + (p.z() - cz) * (p.z() - cz)) - r;
}

static std::vector<double> samples(size_t n, double lo, double hi, unsigned seed)
{
    std::mt19937_64 g(seed);
    std::uniform_real_distribution<double> u(lo, hi);
    std::vector<double> v(3 * n);
    for (double& x : v) x = u(g);
    return v;
}

template <class F>
static bool within(const SDF::Octree& t, const std::vector<double>& pts, F truth, double tol)
{
    std::vector<double> out(pts.size() / 3);
    t.Query(pts.data(), out.size(), out.data());
    for (size_t i = 0; i < out.size(); ++i)
    {
        const Vec3d p{ { pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] } };
        if (std::abs(out[i] - truth(p)) > tol) { std::printf("  point %zu: got %.6f want %.6f\n", i, out[i], truth(p)); return false; }
    }
    return true;
}

static SDF::Config baseConfig(bool continuity)
{
    SDF::Config c;
    c.targetErrorThreshold       = std::pow(10, -8);
    c.nearnessWeighting.type     = SDF::Config::NearnessWeighting::Polynomial;
    c.nearnessWeighting.strength = 3.0;
    c.continuity.enforce         = continuity;
    c.continuity.strength        = 8.0;
    c.threadCount                = std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() : 1;
    return c;
}

int main()
{
    const auto pts = samples(1000000, -0.5, 0.5, 1);
    auto sphereF = [](const Vec3d& p) { return sphere(p, 0.25, 0, 0, 0.5); };
    SDF::Program sphereProg;
    sphereProg.Sphere(0.25, 0, 0, 0.5);
    int failed = 0;
    auto report = [&](const char* name, bool ok) { std::printf("%s: %s\n", name, ok ? "Passed" : "Failed"); failed += !ok; };

    {   // TestOctreeCreation (HPUnitTests.cpp:46-77)
        SDF::Octree t;
        t.Create(baseConfig(false), sphereProg);
        report("Octree Creation", within(t, pts, sphereF, 0.01));
    }
    {   // TestOctreeContinuity (:80-112)
        SDF::Octree t;
        t.Create(baseConfig(true), sphereProg);
        report("Octree Continuity", within(t, pts, sphereF, 0.01));
    }
    {   // TestOctreeSerialisation (:115-154): Create -> ToMemoryBlock -> destroy -> FromMemoryBlock -> free block
        MemoryBlock b;
        {
            SDF::Octree t;
            t.Create(baseConfig(true), sphereProg);
            b = t.ToMemoryBlock();
        }
        SDF::Octree t2;
        t2.FromMemoryBlock(b);
        free(b.ptr);
        report("Octree Serialisation", within(t2, pts, sphereF, 0.01));
    }
    {   // TestOctreeCopying (:157-204): copy-construct, then move-assign
        SDF::Octree t;
        t.Create(baseConfig(false), sphereProg);
        SDF::Octree copy(t);
        t.Clear();
        bool ok = within(copy, pts, sphereF, 0.01);
        SDF::Octree moved;
        moved = std::move(copy);
        ok = ok && within(moved, pts, sphereF, 0.01);
        report("Octree Copying", ok);
    }
    {   // TestOctreeSDFOperations (:207-282): nearness None, second sphere, tolerance 0.05
        SDF::Config c = baseConfig(false);
        c.nearnessWeighting.type = SDF::Config::NearnessWeighting::None;
        SDF::Program other;
        other.Sphere(-0.25, 0, 0, 0.3);
        auto otherF = [](const Vec3d& p) { return sphere(p, -0.25, 0, 0, 0.3); };
        bool ok = true;
        { SDF::Octree t; t.Create(c, sphereProg); t.UnionSDF(other);     ok = ok && within(t, pts, [&](const Vec3d& p) { return std::min(sphereF(p), otherF(p)); }, 0.05); }
        { SDF::Octree t; t.Create(c, sphereProg); t.IntersectSDF(other); ok = ok && within(t, pts, [&](const Vec3d& p) { return std::max(sphereF(p), otherF(p)); }, 0.05); }
        { SDF::Octree t; t.Create(c, sphereProg); t.SubtractSDF(other);  ok = ok && within(t, pts, [&](const Vec3d& p) { return std::max(-sphereF(p), otherF(p)); }, 0.05); }
        report("SDF Operations", ok);
    }
    {   // TestOctreeCustomDomains (:285-316): root [-0.25,5]^3, |p - (0.25,0,0)| - 0.75, default nearness (None), continuity on
        SDF::Config c;
        c.targetErrorThreshold = std::pow(10, -8);
        c.continuity.enforce   = true;
        c.continuity.strength  = 8.0;
        for (int i = 0; i < 3; ++i) { c.root.lo[i] = -0.25f; c.root.hi[i] = 5.0f; }
        SDF::Program p;
        p.Sphere(0.25, 0.0, 0.0, 0.75);
        SDF::Octree t;
        t.Create(c, p);
        const auto big = samples(1000000, -0.25, 5.0, 2);
        report("Custom Domains", within(t, big, [](const Vec3d& q) { return sphere(q, 0.25, 0.0, 0.0, 0.75); }, 0.01));
        const SDF::Box3f root = t.GetRootAABB();
        report("GetRootAABB", root.lo[0] == -0.25f && root.hi[2] == 5.0f);
    }
    {   // error behaviour: out-of-domain Query is DBL_MAX (Octree.cpp:668-671); the std::function Create is refused
        SDF::Octree t;
        t.Create(baseConfig(false), sphereProg);
        bool ok = t.Query(Vec3d{ { 0.75, 0, 0 } }) == std::numeric_limits<double>::max();
        try { t.Create(baseConfig(false), [](const Vec3d&, unsigned long) { return 0.0; }); ok = false; } catch (const SDF::Error& e) { ok = ok && e.status == HPSDF_ERR_UNSUPPORTED; }
        report("Error behaviour", ok);
    }
    {   // QueryRay (Octree.cpp:705-746): a ray from inside the root towards the sphere stops on it, one pointing away misses
        SDF::Octree t;
        t.Create(baseConfig(false), sphereProg);
        double tHit = -1.0, tMiss = -1.0;
        const bool hit = t.QueryRay(SDF::Ray{ Vec3d{ { -0.45, 0.0, 0.0 } }, Vec3d{ { 1.0, 0.0, 0.0 } } }, 2.0, tHit);
        const bool miss = t.QueryRay(SDF::Ray{ Vec3d{ { -0.45, 0.4, 0.4 } }, Vec3d{ { -1.0, 0.0, 0.0 } } }, 2.0, tMiss);
        report("QueryRay", hit && tHit < 0.0001 && !miss && tMiss == -1.0);
    }
    {   // Meshing::Mesh as an SDF source (MeshingUnitTests-style): a regular octahedron |x|+|y|+|z| = 0.3, exact SDF known
        const float r = 0.3f;
        const float verts[18] = { r, 0, 0,  -r, 0, 0,  0, r, 0,  0, -r, 0,  0, 0, r,  0, 0, -r };
        const uint32_t tris[24] = { 0, 2, 4,  2, 1, 4,  1, 3, 4,  3, 0, 4,  2, 0, 5,  1, 2, 5,  3, 1, 5,  0, 3, 5 };
        Meshing::Mesh mesh;
        bool ok = mesh.Create(verts, 6, tris, 8);
        Meshing::Mesh open;
        ok = ok && !open.Create(verts, 6, tris, 7);                         // one face missing: not a closed manifold
        const SDF::Box3f box = mesh.CalculateMeshAABB();
        ok = ok && box.lo[0] == -r && box.hi[2] == r;
        ok = ok && mesh.SignedDistanceAtPt(0.0f, 0.0f, 0.0f) < 0.0f && std::fabs(mesh.SignedDistanceAtPt(0.0f, 0.0f, 0.0f) + r / std::sqrt(3.0f)) < 1e-6f;
        ok = ok && std::fabs(mesh.SignedDistanceAtPt(0.45f, 0.0f, 0.0f) - 0.15f) < 1e-6f;          // beyond a vertex
        SDF::Config c;
        c.targetErrorThreshold = std::pow(10, -7);
        c.continuity.enforce = false;
        SDF::Program p;
        p.Mesh(mesh.handle());
        SDF::Octree t;
        t.Create(c, p);
        // away from the edges and vertices (where the SDF has kinks the polynomial pieces only approach) the tree reproduces the mesh SDF
        double worst = 0.0;
        std::vector<float> q(3 * 20000), d(20000);
        const auto mp = samples(20000, -0.5, 0.5, 3);
        for (size_t i = 0; i < 60000; ++i) q[i] = (float)mp[i];
        mesh.SignedDistanceAtPts(q.data(), 20000, d.data());
        for (size_t i = 0; i < 20000; ++i)
            worst = std::max(worst, std::fabs(t.Query(Vec3d{ { (double)q[3 * i], (double)q[3 * i + 1], (double)q[3 * i + 2] } }) - (double)d[i]));
        report("Mesh SDF source", ok && worst < 0.01);
    }
    std::printf(failed ? "Some tests failed!\n" : "All tests passed!\n");
    return failed;
}
