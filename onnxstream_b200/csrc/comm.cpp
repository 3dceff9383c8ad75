// comm.cpp -- NCCL plumbing for the multi-GPU weight broadcast.  libnccl is resolved at run time (dlopen) so the engine
// library loads on boxes without it; with torch imported the already-loaded torch-bundled libnccl.so.2 is reused.
#include "engine_impl.h"
#include <dlfcn.h>
#include <cstring>
#include <stdexcept>

namespace {

struct NcclId { char b[128]; };   // ncclUniqueId is passed by value: 128 opaque bytes

struct NcclApi {
    using Id = NcclId;
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& api()
{
    static NcclApi a;
    if (a.handle) return a;
    const char* names[] = { "libnccl.so.2", "libnccl.so" };
    for (auto n : names) { a.handle = dlopen(n, RTLD_LAZY | RTLD_GLOBAL); if (a.handle) break; }
    if (!a.handle) throw std::runtime_error("onnxstream_b200: libnccl.so.2 not found (multi-GPU weight broadcast needs NCCL)");
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.Broadcast = (decltype(a.Broadcast))dlsym(a.handle, "ncclBroadcast");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.Broadcast) throw std::runtime_error("onnxstream_b200: incomplete NCCL library");
    return a;
}

}  // namespace

namespace osb {

void WeightStreamer::nccl_broadcast(void* dev, size_t bytes, cudaStream_t st)
{
    // ncclChar = 0 in ncclDataType_t; root 0 is the rank that did the H2D
    int r = api().Broadcast(dev, dev, bytes, 0, 0, (void*)m_comm, st);
    if (r != 0) throw std::runtime_error(std::string("ncclBroadcast failed: ") + (api().GetErrorString ? api().GetErrorString(r) : "?"));
}

// in-place all-gather over a ring slot laid out as nranks chunks of `chunk_bytes`: rank r contributes [r * chunk, (r + 1) * chunk)
void WeightStreamer::nccl_allgather_inplace(void* dev, size_t chunk_bytes, cudaStream_t st)
{
    if (!api().AllGather) throw std::runtime_error("onnxstream_b200: ncclAllGather not found in the NCCL library");
    int r = api().AllGather((const char*)dev + (size_t)m_rank * chunk_bytes, dev, chunk_bytes, 0 /* ncclChar */, (void*)m_comm, st);
    if (r != 0) throw std::runtime_error(std::string("ncclAllGather failed: ") + (api().GetErrorString ? api().GetErrorString(r) : "?"));
}

}  // namespace osb

extern "C" {

int osb_comm_unique_id(char* out128)
{
    try { return api().GetUniqueId(out128); } catch (...) { return -1; }
}

void* osb_comm_init(int nranks, int rank, const char* id128)
{
    try {
        NcclApi::Id id;
        memcpy(id.b, id128, 128);
        void* comm = nullptr;
        int r = api().CommInitRank(&comm, nranks, id, rank);
        return r == 0 ? comm : nullptr;
    } catch (...) { return nullptr; }
}

void osb_comm_destroy(void* comm)
{
    try { if (comm) api().CommDestroy(comm); } catch (...) {}
}

}
