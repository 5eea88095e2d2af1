// sched.h — device-resident greedy scheduler of Octree::Create: state shared by the host driver (build_device.cpp) and the
// kernels (sched_kernels.cuh).
//
// Reference: RunBuildThreadPool (Source/HP/Octree.cpp:194-309) pops the max-error leaf from a std::priority_queue, evaluates
// its refinement job on a worker (TickBuildThread :558-659), applies the h- or p-refinement, pushes the results back and stops
// when totalCoeffError < threshold. Here the queue, the decision, the error bookkeeping and the node allocation live on the
// device; the host only launches kernels and reads one 112-byte header per round.
//
//   nodes      structure of arrays, children in blocks of 8 (Subdivide :1115-1128), same numbering scheme as the reference for
//              the uniform depth-4 start (UniformlyRefine :112-191)
//   open list  node indices of every leaf that sits in the reference's queue (compacted once per round)
//   jobs       cached results of refinement jobs: 8 child errors + the p-fit error, and the pool slots of their coefficients
//   histograms errors of the open leaves by (binary exponent, top 4 mantissa bits): count and an integer upper bound of the sum,
//              updated with integer atomics only, so every rank of a multi-GPU build computes the same bits
#pragma once
#include "hp_common.h"

namespace hpsdf
{
    constexpr int      kSchedThreads  = 1024;
    constexpr uint32_t kSelItems      = 1;                   // open-list entries per thread of the split-mode selection kernels
    constexpr uint32_t kSelChunk      = kSchedThreads * kSelItems;      // entries per CTA
    constexpr int      kSubPerOctave  = 16;
    constexpr int      kOctaves       = 2200;
    constexpr int      kSubBuckets    = kOctaves * kSubPerOctave;
    constexpr int      kWindow        = 1024;                 // entries of the exactly ordered head of the queue handled per pass
    constexpr uint32_t kNone          = 0xFFFFFFFFu;

    // leaf states
    enum : uint8_t
    {
        kStInternal = 0,     // has children
        kStPending  = 1,     // in the queue, refinement job not evaluated yet
        kStEval     = 2,     // in the queue, job selected for the round in flight
        kStCached   = 3,     // in the queue, job result cached
        kStRetired  = 4      // left the queue: degree and depth both at their maximum (Octree.cpp:643-655)
    };

    struct SchedCounters
    {
        uint32_t nNodes, nOpen, nJobs, nCached;
        uint32_t poolUsed;                   // doubles
        uint32_t done;                       // 1 = terminated (Octree.cpp:216), 2 = capacity exceeded
        uint32_t round;
        uint32_t topSub;                     // highest sub-bucket ever used (walk start)
        uint32_t nLog, nDecision;
        uint32_t appliedP, appliedH, retired, nearTies;
        uint32_t passes, windowPasses, lastPassLogStart, pad0;
        uint64_t fitsEvaluated, jobsEvaluated;
        double   total;                      // totalCoeffError (Octree.cpp:212, 257, 272, 276)
        double   exactSum;                   // plain sum of the leaf errors
        double   totalBeforeLast, lastTotal; // around the last applied job (cut margin)
        uint32_t roundJob0;                  // first job of the round in flight
        uint32_t roundJobs;
        // guaranteed level in force between launches: every open entry with (error key, node index) at or above (levelKey,
        // levelNode) is certain to be refined; levelKey = ~0 means "sequential-greedy state, no level"
        unsigned long long levelKey;
        uint32_t levelNode;
        int32_t  aboveLevel;                 // open entries at or above the level that are not refined yet
        unsigned long long nsIngest, nsPasses, nsSelect;   // time inside the scheduler kernel by phase (diagnostics)
        // selection parameters handed from the pass kernel to the multi-block selection kernels (split mode)
        double   selLevel;
        int32_t  selCutSub;
        uint32_t selCutTau;                  // share of the cut sub-bucket taken by the top-up, in 1/65536
        uint32_t selOpen, selJob0, selPool0; // open-list length before compaction, first job index and first pool slot of the new round
    };

    // What the host reads after every scheduler launch (mapped pinned memory).
    struct RoundHeader
    {
        volatile uint32_t seq;               // written last: round number + 1
        uint32_t done;
        uint32_t nJobs;                      // jobs selected for the next round
        uint32_t nTasks;
        uint32_t cnt[kMaxDegree + 2];        // fits per degree of the next round
        uint32_t nNodes, nOpen, nCached, poolUsed;
        uint32_t pad[6];
    };

    static_assert(sizeof(RoundHeader) == 112, "RoundHeader: the size quoted in include/hpsdf.h and DESIGN.md");

    // Device templates of the uniform depth-4 start (built once per device).
    struct SchedTemplates
    {
        const float4*   cell; const uint32_t* child; const uint32_t* code; const uint8_t* depth; const uint8_t* degree; const uint8_t* state;
        const uint32_t* jobNode; const uint32_t* jobPSlot; const uint32_t* jobPPos; const uint8_t* jobFlags;
        uint32_t nNodes, nJobs;
    };

    // Everything the scheduler kernels touch. Passed by value.
    struct SchedDev
    {
        // nodes
        float4*   cell;          // centre, half size (internal unit cube; dyadic, exact in f32)
        uint32_t* child;
        uint32_t* slot;          // pool offset of the leaf's coefficients
        double*   err;
        uint32_t* code;          // child-slot path, 3 bits per level from the top (level 1 in bits 29..27): ascending = DFS order
        uint32_t* jobOf;
        uint8_t*  depth;
        uint8_t*  degree;
        uint8_t*  state;
        // jobs
        uint32_t* jobNode;
        uint32_t* jobHSlot;      // pool slot of child 0 (child c at + c * N_degree)
        uint32_t* jobPSlot;
        uint32_t* jobHPos;       // record index of child 0 / of the p-fit in the round that evaluated the job
        uint32_t* jobPPos;
        double*   jobErr;        // 9 per job: child errors, p-fit error (weighted)
        uint8_t*  jobFlags;      // 1 = child fits evaluated, 2 = p-fit evaluated, 128 = applied
        // lists
        uint32_t* open;
        uint32_t* openAlt;       // split mode: the compacted open list of the next round is written here (the host swaps the two per round)
        uint32_t* chunkCounts;   // split mode: 16 counters per 8192-entry chunk of the open list
        uint32_t* cached;
        uint32_t* scratch;       // candidates of a window selection (capacity = nodes)
        // histograms
        uint32_t* allCnt;
        unsigned long long* allSum;
        uint32_t* pendCnt;
        // outputs
        JobDesc*     jobsOut;
        RoundLayout* layout;
        hpsdf_apply_log_entry*    log;
        hpsdf_decision_log_entry* decisions;
        SchedCounters* ctr;
        RoundHeader*   hostHdr;  // device pointer of the mapped header
        const FitRecord* recs;   // records of the round in flight
        const double*    pool;   // coefficient pool (read by the mc_counter nearness estimate only)
        // capacities
        uint32_t capNodes, capJobs, capLog, capDecisions;
        // configuration
        double   threshold, nearnessStrength;
        unsigned long long nearnessSeed;
        uint32_t nearnessType, nearnessMode, maxDegree, maxDepth, totalMode, minRoundJobs, speculate;
        uint32_t split;          // 1 = ingest and selection run as multi-block kernels around the single-CTA pass kernel
        uint32_t dealJobs;       // 1 = deal the jobs of a round out along a stride permutation (multi-GPU shards get the same mix)
    };
}
