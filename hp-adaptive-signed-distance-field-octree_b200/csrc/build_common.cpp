// build_common.cpp — ReallocCoeffs and the cut-tie log, shared by the two schedulers of Octree::Create.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "build_common.h"

namespace hpsdf
{
    hpsdf_status packCoefficients(hpsdf_octree& t, const double* pool, cudaStream_t stream)
    {
        std::vector<HostNode>& nodes = t.nodes;
        BuildWorkspace& ws = t.ctx->ws;
        std::vector<uint32_t> srcOff, dstOff, count;
        size_t cur = 0;
        std::vector<uint64_t> stack;
        for (int i = 7; i >= 0; --i) stack.push_back(nodes[0].child + (uint64_t)i);
        while (!stack.empty())
        {
            const uint64_t idx = stack.back(); stack.pop_back();
            HostNode& n = nodes[idx];
            if (n.child == kNoChild)
            {
                const uint32_t c = (uint32_t)coeffCount(n.degree);
                srcOff.push_back(n.slot); dstOff.push_back((uint32_t)cur); count.push_back(c);
                n.cstart = cur; cur += c;
            }
            else for (int i = 7; i >= 0; --i) stack.push_back(n.child + (uint64_t)i);
        }
        t.nCoeffs = cur; t.nNodes = nodes.size(); t.nCoeffsPad = paddedCoeffCount(t);
        const uint32_t nSeg = (uint32_t)srcOff.size();
        hpsdf_status st = allocTreeBlob(t);
        if (st != HPSDF_OK) return st;
        HPSDF_CUDA(ws.segs.reserve(3 * (size_t)nSeg + 16));
        HPSDF_CUDA(ws.hSegs.reserve(3 * (size_t)nSeg + 16));
        memcpy(ws.hSegs.p, srcOff.data(), nSeg * 4);
        memcpy(ws.hSegs.p + nSeg, dstOff.data(), nSeg * 4);
        memcpy(ws.hSegs.p + 2 * (size_t)nSeg, count.data(), nSeg * 4);
        HPSDF_CUDA(cudaMemcpyAsync(ws.segs.p, ws.hSegs.p, 3 * (size_t)nSeg * 4, cudaMemcpyHostToDevice, stream));
        HPSDF_CUDA(launchGatherSegments(pool, t.dCoeffs, ws.segs.p, ws.segs.p + nSeg, ws.segs.p + 2 * (size_t)nSeg, nSeg, stream));
        t.stats.kernel_launches++;
        HPSDF_CUDA(cudaStreamSynchronize(stream));       // the pinned segment staging is reused by finalizeQueryStructures
        return HPSDF_OK;
    }

    void ensureApplyLog(hpsdf_octree& t)
    {
        if (!t.logOnDevice) return;
        t.logOnDevice = false;
        t.applyLog.resize(t.nLogDev);
        cudaSetDevice(t.device);
        if (t.nLogDev && cudaMemcpy(t.applyLog.data(), t.dApplyLog, t.nLogDev * sizeof(hpsdf_apply_log_entry), cudaMemcpyDeviceToHost) != cudaSuccess)
        { cudaGetLastError(); t.applyLog.clear(); }
    }

    void ensureDecisionLog(hpsdf_octree& t)
    {
        if (!t.cutLogPending) return;
        t.cutLogPending = false;
        ensureApplyLog(t);
        if (ensureHostNodes(t) != HPSDF_OK) return;
        std::vector<double> errOf(t.nNodes);
        if (cudaMemcpy(errOf.data(), t.dLeafErr, t.nNodes * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return; }
        const double thr = t.cfg.target_error_threshold;
        if (t.applyLog.size() > 4096)
        {
            // the last job applied before the termination cut, with how far the total was from the threshold around it
            const hpsdf_apply_log_entry& a = t.applyLog.back();
            hpsdf_decision_log_entry e{};
            e.node_idx = a.node_idx; e.depth = t.nodes[a.node_idx].depth; e.degree = a.degree; e.chose_p = a.kind == 0; e.kind = 1;
            e.p_improvement = a.p_improvement; e.h_improvement = a.h_improvement;
            for (int k = 0; k < 3; ++k) e.centre[k] = (t.nodes[a.node_idx].mn[k] + t.nodes[a.node_idx].mx[k]) / 2.0f;
            e.relative_margin = std::min(std::fabs(t.cutTotalBeforeLast - thr), std::fabs(thr - t.cutCheck)) / thr;
            t.decisionLog.push_back(e);
        }
        logCutTieGroup(t, errOf, t.cutLogStart, t.cutQueueEmpty);
    }

    void logCutTieGroup(hpsdf_octree& t, const std::vector<double>& errOf, size_t levelLogStart, bool queueEmpty)
    {
        const std::vector<HostNode>& nodes = t.nodes;
        if (t.applyLog.empty() || queueEmpty) return;
        // the sequential loop's last pop is the smallest-error entry applied since the last sequential state
        double eLast = t.applyLog.back().initial_err;
        for (size_t k = std::min(levelLogStart, t.applyLog.size() - 1); k < t.applyLog.size(); ++k) eLast = std::min(eLast, t.applyLog[k].initial_err);
        if (!(eLast > 0.0) || std::abs(eLast - kInitialErr) < 1e-9) return;
        // errors of the members of a symmetric group agree to ~1e-10 relative (they are sums of squares of top-shell
        // coefficients that carry ~1e-16 |c000| of rounding each); the same noise separates two implementations
        const double band = 3e-9 * eLast;
        auto entry = [&](uint64_t idx, uint32_t degree, uint32_t kind, double err)
        {
            hpsdf_decision_log_entry e{};
            e.node_idx = idx; e.depth = nodes[idx].depth; e.degree = degree; e.kind = kind; e.chose_p = 0;
            for (int a = 0; a < 3; ++a) e.centre[a] = (nodes[idx].mn[a] + nodes[idx].mx[a]) / 2.0f;
            e.relative_margin = std::fabs(err - eLast) / eLast;
            t.decisionLog.push_back(e);
        };
        size_t refined = 0, unrefined = 0;
        for (size_t k = t.applyLog.size(); k-- > 0;)
        {
            const hpsdf_apply_log_entry& a = t.applyLog[k];
            if (std::fabs(a.initial_err - eLast) > band) continue;
            entry(a.node_idx, a.degree, 2u, a.initial_err);
            t.decisionLog.back().chose_p = a.kind == 0;
            ++refined;
        }
        for (uint64_t idx = 0; idx < nodes.size(); ++idx)
            if (nodes[idx].child == kNoChild && std::fabs(errOf[idx] - eLast) <= band) { entry(idx, nodes[idx].degree, 3u, errOf[idx]); ++unrefined; }
        if (unrefined == 0)
        {
            // nothing was left behind: the cut does not fall inside a tie group, drop the kind-2 entries again
            t.decisionLog.resize(t.decisionLog.size() - refined);
        }
    }
}
