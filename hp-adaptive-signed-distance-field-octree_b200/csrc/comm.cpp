// comm.cpp — NCCL plumbing for the sharded build: the frontier's fits are split over the ranks and each
// round ends with one grouped broadcast of every rank's results, so all ranks hold the same coefficient pool and replay
// the same greedy order. NCCL is resolved with dlopen("libnccl.so.2") at communicator creation.
#include <dlfcn.h>
#include <cstring>
#include <string>
#include "octree.h"
#include "comm.h"

struct hpsdf_comm
{
    void* nccl = nullptr;     // ncclComm_t
    int   rank = 0, world = 1, device = 0;
};

namespace hpsdf
{
    namespace
    {
        typedef int (*GetUniqueIdFn)(void*);
        struct NcclId { char internal[128]; };
        typedef int (*InitRankFn)(void**, int, NcclId, int);
        typedef int (*DestroyFn)(void*);
        typedef int (*BroadcastFn)(const void*, void*, size_t, int /*dtype*/, int /*root*/, void*, cudaStream_t);
        typedef int (*VoidFn)(void);
        typedef const char* (*ErrStrFn)(int);

        struct Nccl
        {
            void* lib = nullptr;
            GetUniqueIdFn getUniqueId = nullptr;
            InitRankFn    initRank = nullptr;
            DestroyFn     destroy = nullptr;
            BroadcastFn   broadcast = nullptr;
            VoidFn        groupStart = nullptr, groupEnd = nullptr;
            ErrStrFn      errStr = nullptr;
        };

        Nccl* loadNccl(std::string& err)
        {
            static Nccl n;
            if (n.lib) return &n;
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
            if (!h) { err = std::string("cannot load NCCL: ") + dlerror(); return nullptr; }
            n.getUniqueId = (GetUniqueIdFn)dlsym(h, "ncclGetUniqueId");
            n.initRank    = (InitRankFn)dlsym(h, "ncclCommInitRank");
            n.destroy     = (DestroyFn)dlsym(h, "ncclCommDestroy");
            n.broadcast   = (BroadcastFn)dlsym(h, "ncclBroadcast");
            n.groupStart  = (VoidFn)dlsym(h, "ncclGroupStart");
            n.groupEnd    = (VoidFn)dlsym(h, "ncclGroupEnd");
            n.errStr      = (ErrStrFn)dlsym(h, "ncclGetErrorString");
            if (!n.getUniqueId || !n.initRank || !n.destroy || !n.broadcast || !n.groupStart || !n.groupEnd)
            {
                err = "NCCL library lacks required symbols";
                return nullptr;
            }
            n.lib = h;
            return &n;
        }

        hpsdf_status failNccl(Nccl* n, int code, const char* what)
        {
            setLastError(std::string("NCCL error in ") + what + ": " + (n && n->errStr ? n->errStr(code) : "?"));
            return HPSDF_ERR_COMM;
        }
    }

    int commRank(const hpsdf_comm* c)  { return c ? c->rank : 0; }
    int commWorld(const hpsdf_comm* c) { return c ? c->world : 1; }

    hpsdf_status commBroadcastSegments(hpsdf_comm* c, const std::vector<CommSegment>& segs, cudaStream_t stream)
    {
        std::string err;
        Nccl* n = loadNccl(err);
        if (!n || !c || !c->nccl) { setLastError(err.empty() ? "communicator not initialised" : err); return HPSDF_ERR_COMM; }
        int rc = n->groupStart();
        if (rc) return failNccl(n, rc, "ncclGroupStart");
        for (const CommSegment& s : segs)
        {
            rc = n->broadcast(s.ptr, s.ptr, s.count, 8 /* ncclFloat64 */, s.root, c->nccl, stream);
            if (rc) { n->groupEnd(); return failNccl(n, rc, "ncclBroadcast"); }
        }
        rc = n->groupEnd();
        if (rc) return failNccl(n, rc, "ncclGroupEnd");
        return HPSDF_OK;
    }
}

using namespace hpsdf;

extern "C"
{
    HPSDF_API hpsdf_status hpsdf_comm_get_unique_id(void* id_out)
    {
        if (!id_out) { setLastError("id_out is null"); return HPSDF_ERR_INVALID_ARG; }
        std::string err;
        Nccl* n = loadNccl(err);
        if (!n) { setLastError(err); return HPSDF_ERR_COMM; }
        NcclId id;
        memset(&id, 0, sizeof(id));
        const int rc = n->getUniqueId(&id);
        if (rc) return failNccl(n, rc, "ncclGetUniqueId");
        memcpy(id_out, &id, HPSDF_COMM_ID_BYTES);
        return HPSDF_OK;
    }

    HPSDF_API hpsdf_status hpsdf_comm_init(const void* id, int rank, int world_size, int device, hpsdf_comm** out)
    {
        if (!id || !out || world_size < 1 || rank < 0 || rank >= world_size) { setLastError("bad communicator arguments"); return HPSDF_ERR_INVALID_ARG; }
        std::string err;
        if (!getDeviceCtx(device, err)) { setLastError(err); return HPSDF_ERR_NO_DEVICE; }
        Nccl* n = loadNccl(err);
        if (!n) { setLastError(err); return HPSDF_ERR_COMM; }
        NcclId nid;
        memcpy(&nid, id, HPSDF_COMM_ID_BYTES);
        hpsdf_comm* c = new hpsdf_comm();
        c->rank = rank; c->world = world_size;
        cudaGetDevice(&c->device);
        const int rc = n->initRank(&c->nccl, world_size, nid, rank);
        if (rc) { delete c; return failNccl(n, rc, "ncclCommInitRank"); }
        *out = c;
        return HPSDF_OK;
    }

    HPSDF_API void hpsdf_comm_destroy(hpsdf_comm* comm)
    {
        if (!comm) return;
        std::string err;
        Nccl* n = loadNccl(err);
        if (n && comm->nccl) n->destroy(comm->nccl);
        delete comm;
    }

    HPSDF_API void hpsdf_shard_range(size_t n, int rank, int world_size, size_t* begin, size_t* end)
    {
        if (world_size < 1) world_size = 1;
        const size_t base = n / (size_t)world_size, rem = n % (size_t)world_size;
        const size_t r = (size_t)(rank < 0 ? 0 : rank);
        const size_t b = r * base + (r < rem ? r : rem);
        if (begin) *begin = b;
        if (end) *end = b + base + (r < rem ? 1 : 0);
    }
}
