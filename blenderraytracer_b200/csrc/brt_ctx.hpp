// Private host-side state of libbrt shared by brt_api.cu (single-device ABI) and brt_multi.cu (peer groups: the fused
// cross-GPU exchange, one process per GPU or n GPUs in one process).  Not part of the ABI.
#pragma once
#include <atomic>
#include <string>
#include <vector>
#include "brt_host.hpp"
#include "brt_kernels.h"

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinnedBuf {                       // grow-only page-locked staging memory (scene upload)
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = bytes ? bytes : 16;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// One rank's share of a peer group (brt_multi.cu).  Every rank owns ONE device block
//   [accum 0 | accum 1 | rgba8 | floatData | linear | flags]
// that its peers map (CUDA IPC across processes, plain peer access inside one process).  accum is double-buffered by the
// parity of the render epoch, so one "ready" flag per rank orders the whole exchange (see k_peer_reduce_resolve).
struct PeerGroup {
    int rank = 0, world = 0;
    bool ipc = false, connected = false;
    int W = 0, H = 0;
    void* block = nullptr; size_t blockBytes = 0;
    size_t offAccum[2] = { 0, 0 }, offRgba = 0, offFloat = 0, offLinear = 0, offFlags = 0;
    void* peerBlock[brt::MAX_PEERS] = {};
    unsigned epoch = 0;
};

struct brt_ctx {
    int device = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    std::string err;
    brt::HostScene scene; bool haveScene = false; unsigned long long sceneVersion = 0, syncedVersion = ~0ull;
    const brt::HostScene* sceneRef = nullptr;     // a follower of a multi-device ctx reads the leader's scene (never copied)
    brt::HostBackground bg;
    brt_camera cam{}; bool haveCam = false;
    brt_render_params rp{};
    // device scene: one arena holding every SoA array (filled through one pinned staging buffer, one copy)
    DevBuf dArena, dPerm, dPrim64;
    PinnedBuf hStage;
    cudaEvent_t evStage = nullptr; bool stagePending = false;   // marks the last asynchronous copy out of hStage
    brt::BvhWorkspace bvhWs;
    brt::BvhBuildResult bin{};                    // the binary hierarchy of the current scene
    brt::WideBuildResult wide{};                  // its wide collapse (built when a launch wants it)
    brt::DevScene dev{};
    bool sceneDirty = true, permDirty = true, bvhDirty = true, prim64Dirty = true;
    int nBounded = 0;
    brt_scene_info info{};
    // frame buffers
    DevBuf dAccum, dRgba, dFloat, dFloat2, dLinear, dCounters, dScratch, dPlanes;
    // fp64 parity data (lazy)
    DevBuf dObj64, dTris64; bool obj64Dirty = true;
    std::atomic<int> cancel{ 0 };
    brt_stats stats{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    // multi-GPU
    PeerGroup pg;
    std::vector<brt_ctx*> followers;              // brt_create_multi: contexts of devices 1..n-1, owned by this (leader) ctx
    brt_ctx* leader = nullptr;
    const brt::HostScene& hostScene() const { return sceneRef ? *sceneRef : scene; }
};

namespace brt {
int fail(brt_ctx* c, int code, const std::string& msg);
int cuda_fail(brt_ctx* c, cudaError_t e, const char* where);
int prepare(brt_ctx* ctx, PTParams& p);                                 // upload scene / BVH / perm as needed, fill the launch params
int launch_samples(brt_ctx* ctx, PTParams& p, float* dAccum, int sBegin, int sCount);
int reserve_launch_buffers(brt_ctx* ctx, int maxSamplesPerLaunch);      // everything launch_samples may allocate, ahead of time
PostParams post_params(const brt_ctx* ctx);
int effective_spp(const brt_render_params& rp);
int spp_batch(const brt_ctx* ctx, int spp, bool haveCallback);          // samples per launch between progress / cancel polls
int render_multi(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean, brt_progress_cb cb, void* user);
void peer_release(brt_ctx* ctx);
}
#define NEED_GPU() do { if (ctx->device < 0) return brt::fail(ctx, BRT_E_CUDA, "host-only context (device_id = -1): no GPU, and libbrt has no CPU fallback"); } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return brt::cuda_fail(ctx, e_, #call); } while (0)
