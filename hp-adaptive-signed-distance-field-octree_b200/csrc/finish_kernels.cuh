// finish_kernels.cuh — the end of Octree::Create on the device: ReallocCoeffs (Source/HP/Octree.cpp:474-555) and the Query /
// MemoryBlock structures, straight from the scheduler's node arrays (sched.h) without a round trip through the host.
//
// ReallocCoeffs packs the leaves' coefficients in DFS order (children 0..7). Every node carries its child-slot path as a
// left-aligned 30-bit code (3 bits per level), and leaves are disjoint cells, so ascending code order IS that DFS order:
//   leafKeysKernel        (code, node) per node, internal nodes keyed behind every leaf; leaves counted with warp-aggregated atomics
//   cub::DeviceRadixSort  31-bit keys (plumbing)
//   leafOffsetsKernel     exclusive sums of the coefficient counts in sorted order -> coeffsStart of every leaf, packed and
//                         16-byte-aligned (Query layout); totals to the host through mapped memory
//   emitTreeKernel        pool slots -> packed store; 16-byte Query nodes; 56-byte SDF::Node records (Include/HP/Node.h:10-33,
//                         the MemoryBlock image)
//   padCoefficientsKernel packed store -> padded Query store (after the optional continuity solve rewrote the packed one)
#pragma once
#include "hp_common.h"
#include "sched.h"

namespace hpsdf
{
    struct FinishHeader
    {
        volatile uint32_t seq;
        uint32_t nLeaves;
        uint32_t nCoeffs, nCoeffsPad;
        SchedCounters counters;              // the scheduler's final counters ride along (no separate copy)
    };

    // keys[i] = path code of leaf i, or a sentinel above every code for internal nodes (they sort to the end); the leaf count
    // comes from one warp-aggregated atomic per warp
    constexpr uint32_t kNotLeafKey = 0x40000000u;
    __global__ void __launch_bounds__(256) leafKeysKernel(const uint8_t* __restrict__ state, const uint32_t* __restrict__ code, uint32_t nNodes,
                                                          uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ counter)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
        const bool leaf = i < nNodes && state[i] != kStInternal;
        if (i < nNodes) { keys[i] = leaf ? code[i] : kNotLeafKey; vals[i] = i; }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, leaf);
        if (ballot && (threadIdx.x & 31u) == (uint32_t)(__ffs(ballot) - 1)) atomicAdd(counter, (uint32_t)__popc(ballot));
    }

    // Single CTA, 16 leaves per thread and chunk: sorted position -> coeffsStart (packed) and start in the padded store.
    __global__ void __launch_bounds__(1024, 1) leafOffsetsKernel(const uint32_t* __restrict__ sortedNodes, const uint32_t* __restrict__ counter,
                                                                  const uint8_t* __restrict__ degree, uint32_t* __restrict__ cstartOf,
                                                                  uint32_t* __restrict__ padOf, const SchedCounters* __restrict__ ctr,
                                                                  FinishHeader* __restrict__ hdr, uint32_t seq)
    {
        __shared__ unsigned long long sWarp[32];
        constexpr uint32_t IT = 16;
        const uint32_t n = *counter;
        const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
        unsigned long long carry = 0;                               // packed count in the low 32 bits, padded count in the high 32
        for (uint32_t base = 0; base < n; base += 1024u * IT)
        {
            uint32_t node[IT];
            unsigned long long mine = 0;
            #pragma unroll
            for (uint32_t r = 0; r < IT; ++r)
            {
                const uint32_t k = base + tid * IT + r;
                node[r] = k < n ? sortedNodes[k] : kNone;
            }
            uint32_t cnt[IT];
            #pragma unroll
            for (uint32_t r = 0; r < IT; ++r)
            {
                cnt[r] = node[r] != kNone ? (uint32_t)coeffCount((int)degree[node[r]]) : 0u;
                mine += (unsigned long long)cnt[r] | ((unsigned long long)((cnt[r] + 1u) & ~1u) << 32);
            }
            unsigned long long x = mine;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if ((int)lane >= o) x += y; }
            if (lane == 31u) sWarp[warp] = x;
            __syncthreads();
            if (warp == 0)
            {
                unsigned long long w = sWarp[lane];
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, w, o); if ((int)lane >= o) w += y; }
                sWarp[lane] = w;
            }
            __syncthreads();
            unsigned long long run = carry + (warp ? sWarp[warp - 1] : 0ull) + x - mine;
            const unsigned long long total = sWarp[31];
            #pragma unroll
            for (uint32_t r = 0; r < IT; ++r)
            {
                if (node[r] == kNone) continue;
                cstartOf[node[r]] = (uint32_t)(run & 0xFFFFFFFFull);
                padOf[node[r]] = (uint32_t)(run >> 32);
                run += (unsigned long long)cnt[r] | ((unsigned long long)((cnt[r] + 1u) & ~1u) << 32);
            }
            carry += total;
            __syncthreads();
        }
        if (tid == 0)
        {
            hdr->nLeaves = n; hdr->nCoeffs = (uint32_t)(carry & 0xFFFFFFFFull); hdr->nCoeffsPad = (uint32_t)(carry >> 32);
            hdr->counters = *ctr;
            __threadfence_system();
            hdr->seq = seq;
            __threadfence_system();
        }
    }

    // One thread per node writes its Query record and its SDF::Node record; one warp per 32 nodes then copies the leaves'
    // coefficients (a warp walks its 32 nodes, all lanes copy one leaf at a time: coalesced).
    __global__ void __launch_bounds__(256) emitTreeKernel(const SchedDev S, uint32_t nNodes, const uint32_t* __restrict__ cstartOf,
                                                          const uint32_t* __restrict__ padOf, const double* __restrict__ pool,
                                                          double* __restrict__ packed, QNode* __restrict__ qnodes, unsigned char* __restrict__ image)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
        uint32_t slot = 0, cstart = 0, count = 0;
        if (i < nNodes)
        {
            const bool leaf = S.state[i] != kStInternal;
            const uint32_t deg = S.degree[i], depth = S.depth[i];
            const float4 c = S.cell[i];
            QNode q;
            q.child = leaf ? 0xFFFFFFFFu : S.child[i];
            q.cstart = leaf ? padOf[i] : 0u;
            q.degree = leaf ? deg : (uint32_t)kInternalTag;
            q.depth = depth;
            qnodes[i] = q;
            // SDF::Node, LP64: childIdx u64 @0 | aabb.min 3 x f32 @8 | aabb.max @20 | coeffsStart u64 @32 | degree u8 @40 | depth u8 @48; 56 bytes
            uint2* rec = reinterpret_cast<uint2*>(image + 56 * (size_t)i);
            const unsigned long long child = leaf ? kNoChild : (unsigned long long)S.child[i];
            const float mn[3] = { c.x - c.w, c.y - c.w, c.z - c.w }, mx[3] = { c.x + c.w, c.y + c.w, c.z + c.w };       // dyadic: exact (= CornerAABB, Octree.cpp:1096-1112)
            rec[0] = make_uint2((uint32_t)child, (uint32_t)(child >> 32));
            rec[1] = make_uint2(__float_as_uint(mn[0]), __float_as_uint(mn[1]));
            rec[2] = make_uint2(__float_as_uint(mn[2]), __float_as_uint(mx[0]));
            rec[3] = make_uint2(__float_as_uint(mx[1]), __float_as_uint(mx[2]));
            rec[4] = make_uint2(leaf ? cstartOf[i] : 0u, 0u);
            rec[5] = make_uint2(leaf ? deg : (uint32_t)kInternalTag, 0u);
            rec[6] = make_uint2(depth, 0u);
            if (leaf) { slot = S.slot[i]; cstart = cstartOf[i]; count = (uint32_t)coeffCount((int)deg); }
        }
        for (uint32_t src = 0; src < 32u; ++src)
        {
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, count, src);
            if (!n) continue;
            const uint32_t s = __shfl_sync(0xFFFFFFFFu, slot, src), d = __shfl_sync(0xFFFFFFFFu, cstart, src);
            for (uint32_t k = lane; k < n; k += 32u) packed[d + k] = pool[s + k];
        }
    }

    __global__ void __launch_bounds__(256) padCoefficientsKernel(const uint8_t* __restrict__ state, const uint8_t* __restrict__ degree, uint32_t nNodes,
                                                                 const uint32_t* __restrict__ cstartOf, const uint32_t* __restrict__ padOf,
                                                                 const double* __restrict__ packed, double* __restrict__ padded)
    {
        const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
        uint32_t from = 0, to = 0, count = 0;
        if (i < nNodes && state[i] != kStInternal) { from = cstartOf[i]; to = padOf[i]; count = (uint32_t)coeffCount((int)degree[i]); }
        for (uint32_t src = 0; src < 32u; ++src)
        {
            const uint32_t n = __shfl_sync(0xFFFFFFFFu, count, src);
            if (!n) continue;
            const uint32_t s = __shfl_sync(0xFFFFFFFFu, from, src), d = __shfl_sync(0xFFFFFFFFu, to, src);
            for (uint32_t k = lane; k < n + (n & 1u); k += 32u) padded[d + k] = k < n ? packed[s + k] : 0.0;     // the slack element of an odd count is zero
        }
    }
}
