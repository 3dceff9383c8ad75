// workspace.h -- per-stream kernel scratch (split-K partials, GEMV sums, instance-norm partials).
//
// Every Engine owns one compute stream, so "per stream" = per Engine and per device.  Each piece has a FIXED capacity and is
// allocated once, lazily, outside stream capture: an address baked into a captured CUDA graph stays valid for the life of the
// stream's workspace, and two Engines (or two devices) never share scratch.  Callers that need more than the fixed capacity
// take their non-scratch path instead of growing the buffer.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

struct OsbWorkspace {
    float* splitk = nullptr;          // fp32 split-K partial planes
    int* splitk_counters = nullptr;   // 4096 self-resetting ints
    float* gemv = nullptr;            // fp32 column sums of the single-launch GEMV (self-re-arming, zeroed at allocation)
    int* gemv_counters = nullptr;     // 4096 arrival counters
    double* inorm = nullptr;          // instance-norm partial sums
    int device = -1;
};

constexpr size_t OSB_WS_SPLITK_BYTES = (size_t)96 << 20;
constexpr size_t OSB_WS_GEMV_FLOATS = (size_t)2 << 20;      // M (<= 8) x N sums: N up to 262144 at M = 8
constexpr size_t OSB_WS_INORM_DOUBLES = (size_t)1 << 16;

enum { OSB_WS_SPLITK = 1, OSB_WS_GEMV = 2, OSB_WS_INORM = 4 };

// Returns the workspace of `st` with the requested pieces allocated, or nullptr when a piece is missing and cannot be allocated
// now (the stream is capturing, or cudaMalloc failed).  Thread-safe.
OsbWorkspace* osb_workspace(cudaStream_t st, int pieces);
extern "C" void osb_workspace_release(void* stream);
