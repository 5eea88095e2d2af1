// oracle/ref_driver.cpp — C entry points around the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY. Compiled by oracle/Makefile together with /root/reference/Source/{HP,Meshing}/*.cpp
// (where they lie; never copied) against oracle/eigen_shim into oracle/_ref/libhpref.so. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Two build modes:
//   mode 0  literal reference: SDF::Octree::Create as shipped (asynchronous thread pool, std::rand()
//           nearness) — non-deterministic run to run. Used as the CPU baseline.
//   mode 1  deterministic driver: strict greedy (window 1, termination checked after every applied job),
//           exact-mean nearness, reference totalCoeffError bookkeeping in pop order. It calls the
//           reference's OWN FitPolynomial / CornerAABB / Subdivide / ReallocCoeffs /
//           PerformContinuityPostProcess / nodeQueue (private members reached with `#define private public`)
//           and restates only the scheduler arithmetic of Octree.cpp:194-309, 558-659, 804-856.
#include <vector>
#include <queue>
#include <functional>
#include <map>
#include <limits>
#include <mutex>
#include <atomic>
#include <thread>
#include <cassert>
#include <chrono>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <tuple>
#include <malloc.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "Eigen/Core"
#include "Eigen/Geometry"
#include "Eigen/Sparse"

#define private public
#include "HP/Octree.h"
#include "Meshing/Mesh.h"
#include "Meshing/BVH.h"
#undef private

#include "sdf_cpu.h"

namespace
{
    struct RefMesh
    {
        Meshing::Mesh mesh;
        Meshing::BVH  bvh;
        bool          hasBvh = false;
    };

    struct ApplyRecord
    {
        uint64_t nodeIdx;
        uint32_t kind;        // 0 = P, 1 = H
        uint32_t degree;      // degree before the job
        double   initialErr;
        double   newErr;      // P: new error; H: max child error
        double   pImp, hImp;
        double   totalAfter;
    };

    struct RefTree
    {
        SDF::Octree                  oct;
        std::vector<hpsdf_sdf_instr> prog;
        std::vector<ApplyRecord>     log;
        double   buildSeconds = 0.0, continuitySeconds = 0.0;
        uint64_t fits = 0, jobs = 0, appliedP = 0, appliedH = 0;
        double   finalTotal = 0.0;
        long     cgIterations = 0;
        double   cgError = 0.0;
    };

    int currentThread()
    {
    #ifdef _OPENMP
        return omp_get_thread_num();
    #else
        return 0;
    #endif
    }

    // threadIdx_ the reference hands to F (Octree.h:50): its own pool threads are std::threads, not OpenMP threads, and
    // Mesh::SignedDistanceAtPt needs a distinct per-thread BVH queue (BVH.cpp:253-258)
    thread_local int g_callerThread = -1;

    double extEval(uint32_t op, const void* handle, const double x[3])
    {
        if (op == HPSDF_PRIM_MESH)
        {
            RefMesh* m = (RefMesh*)handle;
            const Eigen::Vector3f p((float)x[0], (float)x[1], (float)x[2]);
            return m->hasBvh ? (double)m->mesh.SignedDistanceAtPt(p, m->bvh, (u32)(g_callerThread >= 0 ? g_callerThread : currentThread()))
                             : (double)m->mesh.SignedDistanceAtPt(p);
        }
        if (op == HPSDF_PRIM_OCTREE)
        {
            const RefTree* t = (const RefTree*)handle;
            return t->oct.Query(Eigen::Vector3d(x[0], x[1], x[2]));
        }
        return NAN;
    }

    std::function<f64(const Eigen::Vector3d&, const u32)> makeF(const std::vector<hpsdf_sdf_instr>& prog)
    {
        const hpsdf_sdf_instr* instr = prog.data();
        const uint32_t n = (uint32_t)prog.size();
        return [instr, n](const Eigen::Vector3d& pt_, const u32 threadIdx_) -> f64
        {
            const double x[3] = { pt_.x(), pt_.y(), pt_.z() };
            g_callerThread = (int)threadIdx_;
            return hporacle_sdf_eval(instr, n, x, extEval);
        };
    }

    SDF::Config toConfig(const hpsdf_config* c)
    {
        SDF::Config cfg;
        static_assert(sizeof(SDF::Config) == sizeof(hpsdf_config), "Config image must be 80 bytes (SURVEY App. B)");
        memcpy((void*)&cfg, c, sizeof(cfg));
        return cfg;
    }

    // Deterministic statements of CalculatePolyWeighting / CalculateExpWeighting (Octree.cpp:1209-1247):
    //   default:          the mean of the approximant over its cell is coeffs[0] * NL[0][depth]^3 by orthonormality
    //                     (the limit of the 100-sample estimate);
    //   mc_counter(seed): the reference's own estimator — 100 calls of the reference's FApprox — with the sample points
    //                     from Philox4x32-10 (key = seed, counter = cell coordinates, depth, degree, sample index) instead
    //                     of aabb_.sample() / std::rand(). Same rule as oracle/hp_oracle.c: nearness_mean.
    int      g_nearnessMc = 0;
    uint64_t g_nearnessSeed = 0;

    void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
    {
        for (int r = 0; r < 10; ++r)
        {
            const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
            c[0] = n0; c[1] = (uint32_t)p1; c[2] = n2; c[3] = (uint32_t)p0;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
    }

    double nearnessWeight(const SDF::Octree& o, const SDF::Config& cfg, const SDF::Node::Basis& basis, const Eigen::AlignedBox3f& aabb, const u32 depth)
    {
        if (cfg.nearnessWeighting.type == SDF::Config::NearnessWeighting::None) return 1.0;
        double fIntegral = 0.0;
        if (!g_nearnessMc)
        {
            const double nl = SDF::NormalisedLengths[0][depth];
            fIntegral = basis.coeffs[0] * (nl * nl * nl);
        }
        else
        {
            uint32_t cell[3];
            for (int a = 0; a < 3; ++a) cell[a] = (uint32_t)((aabb.min()(a) + 0.5f) * (float)(1u << depth));
            for (u32 s = 0; s < 100; ++s)
            {
                uint32_t c[4] = { cell[0], cell[1], cell[2], depth | ((u32)basis.degree << 8) | (s << 16) };
                philox4x32_10(c, (uint32_t)g_nearnessSeed, (uint32_t)(g_nearnessSeed >> 32));
                Eigen::Vector3d pt;
                for (int a = 0; a < 3; ++a)
                {
                    const float u = (float)(c[a] >> 8) * 0x1p-24f;
                    const float ext = aabb.max()(a) - aabb.min()(a);
                    const float x = aabb.min()(a) + ext * u;
                    pt(a) = (double)x;
                }
                fIntegral += o.FApprox(basis, aabb, pt, depth);                  // the reference's own evaluator (Octree.cpp:859-901)
            }
            fIntegral /= 100;
        }
        fIntegral = std::abs<f64>(fIntegral);
        const double d = sqrt(3.0);
        if (cfg.nearnessWeighting.type == SDF::Config::NearnessWeighting::Polynomial)
        {
            const double k = std::pow(1.0 - fIntegral / d, cfg.nearnessWeighting.strength);
            return std::min<f64>(1.0, std::max<f64>(k, 0.0));
        }
        return std::exp(-1.0 * cfg.nearnessWeighting.strength * fIntegral / d);
    }

    void deterministicBuild(RefTree* t, const SDF::Config& userCfg, const uint32_t maxDegree, const uint32_t maxDepth,
                            const uint32_t totalMode, const int threads)
    {
        using namespace SDF;
        Octree& o = t->oct;
        o.Clear();
        o.config = userCfg;
        // FitPolynomial must return the raw top-shell energy; the weight is applied here with the exact mean.
        o.config.nearnessWeighting.type = Config::NearnessWeighting::None;

        // Octree.cpp:322-328
        o.configRootCentre               = o.config.root.center().cast<f64>();
        o.configRootInvSizes             = o.config.root.sizes().cwiseInverse().cast<f64>();
        const Eigen::Vector3d rootBounds = o.config.root.sizes().cast<f64>();
        auto userF = makeF(t->prog);
        o.F = [userF, centre = o.configRootCentre, rootBounds](const Eigen::Vector3d& pt_, const u32 threadIdx_) -> f64
        {
            return userF(pt_.cwiseProduct(rootBounds) + centre, threadIdx_);
        };

        o.CreateRoot();
        o.UniformlyRefine();

        const double INITIAL = Octree::INITIAL_NODE_ERR;
        f64 total = pow(8, 4) * INITIAL;          // Octree.cpp:212
        long double exactSum = 0.0L;              // HPSDF_TOTAL_EXACT_SUM: leaf errors without the sentinel
        long unfitted = (long)o.nodeQueue.size();

        while (true)
        {
            const double check = totalMode == HPSDF_TOTAL_EXACT_SUM
                               ? (unfitted > 0 ? std::numeric_limits<double>::infinity() : (double)exactSum) : total;
            if (check < userCfg.targetErrorThreshold || o.nodeQueue.empty()) break;       // Octree.cpp:216

            const std::pair<u32, f64> top = o.nodeQueue.top();                              // Octree.cpp:231-232
            o.nodeQueue.pop();
            const Node node = o.nodes[top.first];
            const f64 err = top.second;
            const bool isCoarse = std::abs<f64>(err - INITIAL) < std::numeric_limits<f64>::epsilon();   // Octree.cpp:806, 831
            const u32 p = node.basis.degree, depth = node.depth;

            Node::Basis hBases[8], pBasis;
            f64 hErrs[8] = { 0 }, pErr = 0.0, hImp = 0.0, pImp = 0.0;
            for (auto& b : hBases) { b.coeffs = nullptr; b.degree = 0; }
            pBasis.coeffs = nullptr; pBasis.degree = 0;

            const bool doH = !isCoarse && depth < maxDepth;
            const bool doP = isCoarse || p < maxDegree;
            if (doH) for (u32 i = 0; i < 8; ++i) hBases[i].coeffs = (f64*)malloc(sizeof(f64) * LegendreCoeffientCount[p]);   // Octree.cpp:817-818
            if (doP)
            {
                if (isCoarse) { pBasis.coeffs = (f64*)malloc(sizeof(f64) * LegendreCoeffientCount[2]); pBasis.degree = 0; }      // Octree.cpp:838-839
                else
                {
                    pBasis.coeffs = (f64*)malloc(sizeof(f64) * LegendreCoeffientCount[p + 1]);                                   // Octree.cpp:846-848
                    memcpy(pBasis.coeffs, node.basis.coeffs, sizeof(f64) * LegendreCoeffientCount[p]);
                    pBasis.degree = (u8)p;
                }
            }
            f64 rawH[8] = { 0 }, rawP = 0.0;
            #pragma omp parallel for schedule(dynamic, 1) num_threads(threads) if (threads > 1)
            for (int f = 0; f < 9; ++f)
            {
                if (f < 8) { if (doH) rawH[f] = o.FitPolynomial(hBases[f], o.CornerAABB(node.aabb, (u32)f), (u8)p, depth + 1, (u32)currentThread()); }   // Octree.cpp:820
                else if (doP) rawP = o.FitPolynomial(pBasis, node.aabb, isCoarse ? (u8)2 : (u8)(p + 1), depth, (u32)currentThread());                     // Octree.cpp:840, 851
            }
            t->jobs++;
            if (doH)
            {
                f64 maxNewErr = 0.0;
                for (u32 i = 0; i < 8; ++i)
                {
                    hErrs[i]  = rawH[i] * nearnessWeight(o, userCfg, hBases[i], o.CornerAABB(node.aabb, i), depth + 1);
                    maxNewErr = std::max<f64>(maxNewErr, hErrs[i]);                                                            // Octree.cpp:821
                }
                hImp = (1.0 / (7.0 * LegendreCoeffientCount[p])) * (err - 8.0 * maxNewErr);                                   // Octree.cpp:825
                t->fits += 8;
            }
            if (doP)
            {
                pErr = rawP * nearnessWeight(o, userCfg, pBasis, node.aabb, depth);
                pImp = isCoarse ? pErr                                                                                         // Octree.cpp:842
                                : (1.0 / (LegendreCoeffientCount[p + 1] - LegendreCoeffientCount[p])) * (err - 8.0 * pErr);   // Octree.cpp:854
                t->fits += 1;
            }

            // Octree.cpp:600-601 with BASIS_MAX_DEGREE-1 -> maxDegree, TREE_MAX_DEPTH -> maxDepth
            // A coarse cell always takes its degree-2 fit: the reference would fall into the H branch with null child
            // bases if that fit's error were exactly 0 (undefined behaviour) — avoided here.
            const bool refineP = isCoarse || (p < maxDegree && (depth == maxDepth || pImp > hImp));
            const bool refineH = depth < maxDepth && !refineP;

            ApplyRecord rec = { top.first, 0u, (uint32_t)p, err, 0.0, pImp, hImp, 0.0 };
            if (refineP)
            {
                for (auto& b : hBases) if (b.coeffs) free(b.coeffs);
                free(o.nodes[top.first].basis.coeffs);                                          // Octree.cpp:256
                total += (pErr - err);                                                          // Octree.cpp:257
                if (isCoarse) unfitted--; else exactSum -= (long double)err;
                exactSum += (long double)pErr;
                o.nodes[top.first].basis = pBasis;                                              // Octree.cpp:286
                o.nodeQueue.push({ top.first, pErr });                                          // Octree.cpp:289-290
                t->appliedP++;
                rec.kind = 0; rec.newErr = pErr;
            }
            else if (refineH)
            {
                if (pBasis.coeffs) free(pBasis.coeffs);
                free(o.nodes[top.first].basis.coeffs);                                          // Octree.cpp:267-272
                o.nodes[top.first].basis.degree = (BASIS_MAX_DEGREE + 1);
                o.Subdivide(top.first);
                total -= err;
                exactSum -= (long double)err;
                f64 mx = 0.0;
                for (u32 i = 0; i < 8; ++i)
                {
                    const u32 child = o.nodes[top.first].childIdx + i;                          // Octree.cpp:275-276
                    total += hErrs[i];
                    exactSum += (long double)hErrs[i];
                    o.nodes[child].basis = hBases[i];
                    o.nodeQueue.push({ child, hErrs[i] });
                    mx = std::max(mx, hErrs[i]);
                }
                t->appliedH++;
                rec.kind = 1; rec.newErr = mx;
            }
            else
            {
                for (auto& b : hBases) if (b.coeffs) free(b.coeffs);                            // Octree.cpp:643-655: node leaves the queue
                if (pBasis.coeffs) free(pBasis.coeffs);
                continue;
            }
            rec.totalAfter = totalMode == HPSDF_TOTAL_EXACT_SUM ? (unfitted > 0 ? std::numeric_limits<double>::infinity() : (double)exactSum) : total;
            t->log.push_back(rec);
        }
        t->finalTotal = totalMode == HPSDF_TOTAL_EXACT_SUM ? (double)exactSum : total;
        o.config = userCfg;    // the serialised Config is the caller's
    }
}

extern "C"
{
    // mode: 0 literal reference Create, 1 deterministic driver. Returns a handle.
    void* hpref_build(const hpsdf_config* cfg_, const hpsdf_sdf_instr* prog_, uint32_t nInstr_, int mode_,
                      uint32_t maxDegree_, uint32_t maxDepth_, uint32_t totalMode_, double cgTol_, int threads_)
    {
        RefTree* t = new RefTree();
        t->prog.assign(prog_, prog_ + nInstr_);
        SDF::Config cfg = toConfig(cfg_);
        Eigen::shim_cg_tolerance_override() = cgTol_;
        const auto t0 = std::chrono::high_resolution_clock::now();
        if (mode_ == 0)
        {
            t->oct.Create(cfg, makeF(t->prog));
            t->buildSeconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
        }
        else
        {
            deterministicBuild(t, cfg, maxDegree_, maxDepth_, totalMode_, threads_ > 0 ? threads_ : 1);
            const auto t1 = std::chrono::high_resolution_clock::now();
            const u32 nCoeffs = t->oct.ReallocCoeffs();                          // Octree.cpp:338
            if (cfg.continuity.enforce) t->oct.PerformContinuityPostProcess(nCoeffs);     // Octree.cpp:341-344
            t->oct.procMap.clear();
            while (!t->oct.nodeQueue.empty()) t->oct.nodeQueue.pop();
            const auto t2 = std::chrono::high_resolution_clock::now();
            t->buildSeconds      = std::chrono::duration<double>(t2 - t0).count();
            t->continuitySeconds = std::chrono::duration<double>(t2 - t1).count();
        }
        t->cgIterations = Eigen::shim_cg_last_iterations();
        t->cgError      = Eigen::shim_cg_last_error();
        Eigen::shim_cg_tolerance_override() = 0.0;
        return t;
    }

    void hpref_set_nearness_mc(int on_, uint64_t seed_) { g_nearnessMc = on_; g_nearnessSeed = seed_; }

    void* hpref_from_block(const void* ptr_, size_t size_)
    {
        RefTree* t = new RefTree();
        MemoryBlock b = { size_, const_cast<void*>(ptr_) };
        t->oct.FromMemoryBlock(b);
        return t;
    }

    void hpref_destroy(void* h_) { delete (RefTree*)h_; }

    size_t hpref_block_size(void* h_)
    {
        MemoryBlock b = ((RefTree*)h_)->oct.ToMemoryBlock();
        free(b.ptr);
        return b.size;
    }

    void hpref_block_copy(void* h_, void* dst_)
    {
        MemoryBlock b = ((RefTree*)h_)->oct.ToMemoryBlock();
        memcpy(dst_, b.ptr, b.size);
        free(b.ptr);
    }

    void hpref_query(void* h_, const double* xyz_, size_t n_, double* out_, int threads_)
    {
        const SDF::Octree& o = ((RefTree*)h_)->oct;
        #pragma omp parallel for schedule(static) num_threads(threads_) if (threads_ > 1)
        for (long i = 0; i < (long)n_; ++i)
            out_[i] = o.Query(Eigen::Vector3d(xyz_[3 * i], xyz_[3 * i + 1], xyz_[3 * i + 2]));
    }

    // Octree::QueryRay (Octree.cpp:705-746) on the reference's own code
    void hpref_query_ray(void* h_, const double* origins_, const double* dirs_, size_t n_, double tMax_, unsigned char* hit_, double* t_)
    {
        const SDF::Octree& o = ((RefTree*)h_)->oct;
        for (size_t i = 0; i < n_; ++i)
        {
            const SDF::Ray ray(Eigen::Vector3d(origins_[3 * i], origins_[3 * i + 1], origins_[3 * i + 2]), Eigen::Vector3d(dirs_[3 * i], dirs_[3 * i + 1], dirs_[3 * i + 2]));
            double t = 0.0;
            hit_[i] = o.QueryRay(ray, tMax_, t) ? 1 : 0;
            t_[i] = t;
        }
    }

    void hpref_query_gradient(void* h_, const double* xyz_, size_t n_, double* out_, double* grad_, int threads_)
    {
        const SDF::Octree& o = ((RefTree*)h_)->oct;
        #pragma omp parallel for schedule(static) num_threads(threads_) if (threads_ > 1)
        for (long i = 0; i < (long)n_; ++i)
        {
            Eigen::Vector3d g(0.0, 0.0, 0.0);
            out_[i] = o.QueryWithGradient(Eigen::Vector3d(xyz_[3 * i], xyz_[3 * i + 1], xyz_[3 * i + 2]), g);
            grad_[3 * i] = g.x(); grad_[3 * i + 1] = g.y(); grad_[3 * i + 2] = g.z();
        }
    }

    // out[0..9]: seconds, continuity seconds, fits, jobs, applied P, applied H, final total, cg iterations, cg error, log size
    void hpref_stats(void* h_, double* out_)
    {
        RefTree* t = (RefTree*)h_;
        out_[0] = t->buildSeconds; out_[1] = t->continuitySeconds; out_[2] = (double)t->fits; out_[3] = (double)t->jobs;
        out_[4] = (double)t->appliedP; out_[5] = (double)t->appliedH; out_[6] = t->finalTotal;
        out_[7] = (double)t->cgIterations; out_[8] = t->cgError; out_[9] = (double)t->log.size();
    }

    // rows of 8 doubles: nodeIdx, kind, degree, initialErr, newErr, pImp, hImp, totalAfter
    void hpref_apply_log(void* h_, double* out_)
    {
        RefTree* t = (RefTree*)h_;
        for (size_t i = 0; i < t->log.size(); ++i)
        {
            const ApplyRecord& r = t->log[i];
            double* o = out_ + 8 * i;
            o[0] = (double)r.nodeIdx; o[1] = r.kind; o[2] = r.degree; o[3] = r.initialErr; o[4] = r.newErr; o[5] = r.pImp; o[6] = r.hImp; o[7] = r.totalAfter;
        }
    }

    // One FitPolynomial call (Octree.cpp:1007-1093) with nearness None: the raw top-shell energy.
    double hpref_fit(const hpsdf_config* cfg_, const hpsdf_sdf_instr* prog_, uint32_t nInstr_,
                     const float* aabbMin_, const float* aabbMax_, uint32_t degreeIn_, const double* coeffsIn_,
                     uint32_t degree_, uint32_t depth_, double* coeffsOut_)
    {
        using namespace SDF;
        RefTree t;
        t.prog.assign(prog_, prog_ + nInstr_);
        Octree& o = t.oct;
        o.config = toConfig(cfg_);
        o.config.nearnessWeighting.type = Config::NearnessWeighting::None;
        const Eigen::Vector3d centre     = o.config.root.center().cast<f64>();
        const Eigen::Vector3d rootBounds = o.config.root.sizes().cast<f64>();
        auto userF = makeF(t.prog);
        o.F = [userF, centre, rootBounds](const Eigen::Vector3d& pt_, const u32 threadIdx_) -> f64
        {
            return userF(pt_.cwiseProduct(rootBounds) + centre, threadIdx_);
        };
        Node::Basis b;
        b.coeffs = coeffsOut_;
        b.degree = (u8)degreeIn_;
        if (degreeIn_ > 0) memcpy(coeffsOut_, coeffsIn_, sizeof(f64) * LegendreCoeffientCount[degreeIn_]);
        const Eigen::AlignedBox3f aabb(Eigen::Vector3f(aabbMin_[0], aabbMin_[1], aabbMin_[2]), Eigen::Vector3f(aabbMax_[0], aabbMax_[1], aabbMax_[2]));
        return o.FitPolynomial(b, aabb, (u8)degree_, depth_, 0);
    }

    // The history of n leaves replayed with the reference's own FitPolynomial, one leaf per OpenMP thread: a from-scratch fit at
    // degree d0_[i] (Octree.cpp:820 child fit / :840 coarse fit), then one kept-shell fit per degree up to d1_[i] (:846-851).
    // coeffsOut_: n x 455 doubles (row i holds N_{d1[i]} coefficients), errOut_: raw top-shell energy of the last fit.
    // keptDegree_ (may be null): where non-zero, the first fit of leaf i is itself a kept-shell fit on top of the degree
    // keptDegree_[i] coefficients already in row i (a p-fit on its own, as a timing sample of EstimatePImprovement, :829-856).
    void hpref_fit_chain_batch(const hpsdf_config* cfg_, const hpsdf_sdf_instr* prog_, uint32_t nInstr_, size_t n_,
                               const float* aabbMin_, const float* aabbMax_, const uint32_t* d0_, const uint32_t* d1_,
                               const uint32_t* depth_, const uint32_t* keptDegree_, double* coeffsOut_, double* errOut_, int threads_)
    {
        using namespace SDF;
        RefTree t;
        t.prog.assign(prog_, prog_ + nInstr_);
        Octree& o = t.oct;
        o.config = toConfig(cfg_);
        o.config.nearnessWeighting.type = Config::NearnessWeighting::None;
        const Eigen::Vector3d centre     = o.config.root.center().cast<f64>();
        const Eigen::Vector3d rootBounds = o.config.root.sizes().cast<f64>();
        auto userF = makeF(t.prog);
        o.F = [userF, centre, rootBounds](const Eigen::Vector3d& pt_, const u32 threadIdx_) -> f64
        {
            return userF(pt_.cwiseProduct(rootBounds) + centre, threadIdx_);
        };
        #pragma omp parallel for schedule(dynamic, 1) num_threads(threads_) if (threads_ > 1)
        for (long i = 0; i < (long)n_; ++i)
        {
            Node::Basis b;
            b.coeffs = coeffsOut_ + 455 * (size_t)i;
            b.degree = keptDegree_ ? (u8)keptDegree_[i] : (u8)0;
            const Eigen::AlignedBox3f aabb(Eigen::Vector3f(aabbMin_[3 * i], aabbMin_[3 * i + 1], aabbMin_[3 * i + 2]),
                                           Eigen::Vector3f(aabbMax_[3 * i], aabbMax_[3 * i + 1], aabbMax_[3 * i + 2]));
            double e = 0.0;
            for (uint32_t d = d0_[i]; d <= d1_[i]; ++d)
            {
                e = o.FitPolynomial(b, aabb, (u8)d, depth_[i], (u32)currentThread());
                b.degree = (u8)d;
            }
            errOut_[i] = e;
        }
    }

    // Reference constant tables (Include/HP/Utility.h, Include/HP/Legendre.h) for pinning the restatement's own tables.
    void hpref_tables(double* nl_ /*13*11*/, uint32_t* counts_ /*13*/, uint32_t* basisIdx_ /*455*3*/,
                      double* recur_ /*13*2*/, double* roots_ /*2080*/, double* weights_ /*2080*/, uint32_t* faceLookup_ /*24*/)
    {
        using namespace SDF;
        for (u32 i = 0; i <= BASIS_MAX_DEGREE; ++i)
        {
            for (u32 j = 0; j <= TREE_MAX_DEPTH; ++j) nl_[i * (TREE_MAX_DEPTH + 1) + j] = NormalisedLengths[i][j];
            counts_[i] = (uint32_t)LegendreCoeffientCount[i];
            recur_[2 * i] = LegendreCoefficent[i][0]; recur_[2 * i + 1] = LegendreCoefficent[i][1];
        }
        for (u32 i = 0; i < 455; ++i) for (u32 k = 0; k < 3; ++k) basisIdx_[3 * i + k] = (uint32_t)BasisIndexValues[i][k];
        for (u32 i = 0; i < 2080; ++i) { roots_[i] = LegendreRoots[i]; weights_[i] = LegendreWeights[i]; }
        for (u32 d = 0; d < 3; ++d) for (u32 j = 0; j < 4; ++j) for (u32 k = 0; k < 2; ++k) faceLookup_[(d * 4 + j) * 2 + k] = (uint32_t)SharedFaceLookup[d][j][k];
    }

    // Face-jump Gram matrix entries exactly as RunContinuityThreadPool emits them (Octree.cpp:1663-1714), before
    // the lambda diagonal. Two-call protocol: rows_ == NULL returns the count.
    size_t hpref_continuity_triplets(void* h_, int32_t* rows_, int32_t* cols_, double* vals_)
    {
        static thread_local std::vector<Eigen::Triplet<f64>> trips;
        RefTree* t = (RefTree*)h_;
        if (!rows_)
        {
            trips.clear();
            t->oct.procMap.clear();
            t->oct.RunContinuityThreadPool(trips);
            t->oct.procMap.clear();
            return trips.size();
        }
        for (size_t i = 0; i < trips.size(); ++i) { rows_[i] = trips[i].row(); cols_[i] = trips[i].col(); vals_[i] = trips[i].value(); }
        return trips.size();
    }

    // ---- Meshing (Source/Meshing/*.cpp) ------------------------------------------------------------------
    void* hpref_mesh_create(const float* v_, size_t nv_, const uint32_t* tri_, size_t nt_, int buildBvh_)
    {
        RefMesh* m = new RefMesh();
        m->mesh.vertices.resize(nv_);
        for (size_t i = 0; i < nv_; ++i) m->mesh.vertices[i] = Eigen::Vector3f(v_[3 * i], v_[3 * i + 1], v_[3 * i + 2]);
        m->mesh.triIndices.resize(3 * nt_);
        for (size_t i = 0; i < 3 * nt_; ++i) m->mesh.triIndices[i] = tri_[i];
        if (!m->mesh.CreateHalfEdges()) { delete m; return nullptr; }
        if (buildBvh_) { m->bvh.Create(m->mesh); m->hasBvh = true; }
        return m;
    }

    void* hpref_mesh_load_obj(const char* path_, int buildBvh_)
    {
        RefMesh* m = new RefMesh();
        if (!m->mesh.CreateFromObj(path_)) { delete m; return nullptr; }
        if (buildBvh_) { m->bvh.Create(m->mesh); m->hasBvh = true; }
        return m;
    }

    void hpref_mesh_counts(void* m_, size_t* nv_, size_t* nt_)
    {
        RefMesh* m = (RefMesh*)m_;
        *nv_ = m->mesh.vertices.size(); *nt_ = m->mesh.triIndices.size() / 3;
    }

    void hpref_mesh_arrays(void* m_, float* v_, uint32_t* tri_)
    {
        RefMesh* m = (RefMesh*)m_;
        for (size_t i = 0; i < m->mesh.vertices.size(); ++i) for (int k = 0; k < 3; ++k) v_[3 * i + k] = m->mesh.vertices[i](k);
        for (size_t i = 0; i < m->mesh.triIndices.size(); ++i) tri_[i] = (uint32_t)m->mesh.triIndices[i];
    }

    void hpref_mesh_sdf(void* m_, const float* xyz_, size_t n_, float* out_, int useBvh_, int threads_)
    {
        RefMesh* m = (RefMesh*)m_;
        const bool bvh = useBvh_ && m->hasBvh;
        #pragma omp parallel for schedule(dynamic, 64) num_threads(threads_) if (threads_ > 1)
        for (long i = 0; i < (long)n_; ++i)
        {
            const Eigen::Vector3f p(xyz_[3 * i], xyz_[3 * i + 1], xyz_[3 * i + 2]);
            out_[i] = bvh ? m->mesh.SignedDistanceAtPt(p, m->bvh, (u32)currentThread()) : m->mesh.SignedDistanceAtPt(p);
        }
    }

    void hpref_mesh_destroy(void* m_) { delete (RefMesh*)m_; }

    int hpref_hardware_threads() { const unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }
}
