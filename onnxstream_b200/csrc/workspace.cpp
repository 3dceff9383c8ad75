// workspace.cpp -- per-stream, fixed-capacity kernel scratch (see workspace.h).
#include "workspace.h"

#include <map>
#include <mutex>

namespace {
std::mutex g_mu;
std::map<cudaStream_t, OsbWorkspace> g_ws;

bool capturing(cudaStream_t st)
{
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

template <typename T>
bool alloc_zeroed(T** p, size_t count, cudaStream_t st)
{
    if (cudaMalloc((void**)p, count * sizeof(T)) != cudaSuccess) { *p = nullptr; cudaGetLastError(); return false; }
    // ordered before the first kernel that uses the piece: same stream
    return cudaMemsetAsync(*p, 0, count * sizeof(T), st) == cudaSuccess;
}
}  // namespace

OsbWorkspace* osb_workspace(cudaStream_t st, int pieces)
{
    std::lock_guard<std::mutex> lock(g_mu);
    OsbWorkspace& w = g_ws[st];
    if (w.device < 0) cudaGetDevice(&w.device);
    bool need = ((pieces & OSB_WS_SPLITK) && !w.splitk) || ((pieces & OSB_WS_GEMV) && !w.gemv) || ((pieces & OSB_WS_INORM) && !w.inorm);
    if (!need) return &w;
    if (capturing(st)) return nullptr;     // never allocate inside a capture: the eager warm-up runs size everything
    if ((pieces & OSB_WS_SPLITK) && !w.splitk) {
        if (!alloc_zeroed(&w.splitk_counters, 4096, st)) return nullptr;
        if (cudaMalloc((void**)&w.splitk, OSB_WS_SPLITK_BYTES) != cudaSuccess) { w.splitk = nullptr; cudaGetLastError(); return nullptr; }
    }
    if ((pieces & OSB_WS_GEMV) && !w.gemv) {
        if (!alloc_zeroed(&w.gemv_counters, 4096, st)) return nullptr;
        if (!alloc_zeroed(&w.gemv, OSB_WS_GEMV_FLOATS, st)) return nullptr;
    }
    if ((pieces & OSB_WS_INORM) && !w.inorm) {
        if (cudaMalloc((void**)&w.inorm, OSB_WS_INORM_DOUBLES * sizeof(double)) != cudaSuccess) { w.inorm = nullptr; cudaGetLastError(); return nullptr; }
    }
    return &w;
}

extern "C" void osb_workspace_release(void* stream)
{
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_ws.find((cudaStream_t)stream);
    if (it == g_ws.end()) return;
    OsbWorkspace& w = it->second;
    cudaFree(w.splitk); cudaFree(w.splitk_counters); cudaFree(w.gemv); cudaFree(w.gemv_counters); cudaFree(w.inorm);
    g_ws.erase(it);
}
