// Host-visible declarations of the kernel launchers (one .cu per subsystem, linked into libbrt.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "brt_device.cuh"

namespace brt {

constexpr int PT_BLOCK = 128;
constexpr int PT_MIN_BLOCKS = 6;      // __launch_bounds__ min blocks/SM of the wavefront kernel (<= 80 registers)
#ifndef PT_MIN_BLOCKS_MEGA
#define PT_MIN_BLOCKS_MEGA 8          // megakernel at 64 registers: 8 blocks = 32 warps / SM (measured 5..8: 8 best; 32 B of spills in the fast-sampler kernels)
#endif
constexpr int PT_SLOT_WORDS = 13;    // shared-memory words per path slot (pathtrace.cu: SlotField; +1 for the sequential sampler)

struct PTParams {
    DevScene sc;
    DevCamera cam;
    DevCamera32 cam32;
    int W, H;
    int rowBegin, rowEnd;        // only pixels of rows [rowBegin, rowEnd) are traced (the whole frame unless BRT_DEBUG_ROW_WINDOW narrows it: tools/rare_event_check.py)
    int sBegin, sCount;          // global sample indices [sBegin, sBegin + sCount) for every pixel
    int maxDepth;
    int aaMode;
    uint32_t seedLo, seedHi;
    int directLighting;
    int wavefront;               // 0 = megakernel (one blocking traversal per thread), 1 = warp-local wavefront
    int inflight;                // path slots per lane (samples of a pixel in flight): 1, 2, 3 or 4
    int refill;                  // extend phase: idle lanes fetch the next queued ray once at least this many are idle
    float4* accum;               // W*H fp32 RGBA sums (alpha = number of samples); with z chunks: planes of W*H each
    size_t planeStride;          // float4 elements between the planes of consecutive z chunks (0 = single plane)
    unsigned long long* counters;// 8 x u64 (Counters), counting build only
};

// pathtrace.cu
cudaError_t launch_pathtrace(const PTParams& p, int sampler, bool useBvh, bool count, int zSplit, cudaStream_t st);
// pathtrace_wave.cu
cudaError_t launch_pathtrace_wave(const PTParams& p, int sampler, bool useBvh, int zSplit, cudaStream_t st);
cudaError_t launch_sum_planes(float4* accum, const float4* planes, int nPlanes, size_t px, cudaStream_t st);
cudaError_t launch_primary_aov(const PTParams& p, bool useBvh, int* objId, int* triId, float* t, float* nrm, unsigned char* front,
                               cudaStream_t st);
cudaError_t launch_eval_background(const DevScene& sc, const float* dirs, int n, float* out, cudaStream_t st);
cudaError_t launch_eval_texture(const DevScene& sc, int texIndex, const float* points, int n, float* out, cudaStream_t st);
cudaError_t launch_rng_stream(uint32_t lo, uint32_t hi, uint32_t pixel, uint32_t sample, int n, float* out, cudaStream_t st);
cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st);

// post.cu  (ray-tracer.js:208-233, 266-276; post-processor.js)
struct PostParams {
    int W, H;
    int tonemap;
    double exposure, invGamma;
    double w1, w2;               // exp(-1/(2σ²)), exp(-2/(2σ²)) computed on the host in float64 (post-processor.js:59)
};
// accum (sum, alpha = count) -> rgba8 [+ floatData] [+ linear mean]; rows [rowBegin,rowEnd)
cudaError_t launch_resolve(const PostParams& pp, const float4* accum, uchar4* rgba, float4* floatData, float4* linear,
                           int rowBegin, int rowEnd, cudaStream_t st);
// sums rows of nPeers accumulation buffers in rank order, then resolves (fused NVLink reduce + resolve)
cudaError_t launch_reduce_resolve(const PostParams& pp, const float4* const* peers, int nPeers, uchar4* rgba, float4* floatData,
                                  int rowBegin, int rowEnd, cudaStream_t st);
cudaError_t launch_denoise(const PostParams& pp, const float4* floatData, uchar4* rgba, float4* outFloat, cudaStream_t st);

// The synchronised form of the fused exchange (brt_multi.cu): every rank owns three flag words in its peer-mapped block —
//   FLAG_READY: the epoch whose sums this rank has finished tracing;  FLAG_DONE: the epoch whose row stripe this rank has
//   resolved into the root's image;  FLAG_BLOCKS: completion counter of the stripe kernel;  FLAG_ERR: a wait timed out.
constexpr int MAX_PEERS = 16;
enum PeerFlag : int { FLAG_READY = 0, FLAG_DONE = 1, FLAG_BLOCKS = 2, FLAG_ERR = 3, FLAG_WORDS = 64 };
struct PeerSync {
    const float4* accum[MAX_PEERS];   // every rank's sums of this epoch (peer-mapped), rank order
    unsigned* flags[MAX_PEERS];       // every rank's flag words (peer-mapped)
    int nPeers, self;
    unsigned epoch;
};
// k_peer_reduce_resolve: publishes READY(self) = epoch, waits until READY(r) >= epoch for every rank (acquire at system scope),
// sums rows [rowBegin, rowEnd) of all ranks' buffers in rank order, resolves, stores RGBA8 (+ floatData / linear) into the
// ROOT's image, and the last block to finish publishes DONE(self) = epoch.
cudaError_t launch_peer_reduce_resolve(const PostParams& pp, const PeerSync& ps, uchar4* rgbaRoot, float4* floatRoot, float4* linearRoot,
                                       int rowBegin, int rowEnd, cudaStream_t st);
// k_peer_wait: one block that spins until flag `which` of every rank reached ps.epoch (the root runs it with FLAG_DONE before it reads its image)
cudaError_t launch_peer_wait(const PeerSync& ps, int which, cudaStream_t st);

// aov64.cu — float64, FMA-free, brute-force primary visibility with the reference's exact operation order
struct Obj64 {                   // one per world.objects entry, in order
    int type, material;
    double a[3], b[3], c[3];     // as brt_object (plane normal already normalised, geometry.js:52)
    long long firstTri, triCount;
};
struct Cam64 {
    double origin[3], llc[3], horizontal[3], vertical[3], w[3];
    int type;
};
cudaError_t launch_primary_aov64(const Obj64* objs, int nObjs, const double* meshTris /* 9 per tri */, const Cam64& cam, int W, int H,
                                 int* objId, int* triId, double* t, double* nrm, unsigned char* front, cudaStream_t st);

// bvh.cu — GPU LBVH: Morton codes -> radix sort -> Karras hierarchy -> bottom-up refit
struct BvhWorkspace {            // grow-only build memory, owned by the ctx (free_bvh_workspace at destroy)
    void* arena = nullptr; size_t arenaCap = 0;
    float4* nodes[2] = { nullptr, nullptr }; size_t nodeCap[2] = { 0, 0 };
    float4* wide = nullptr; size_t wideCap = 0;         // wide hierarchy (build_wide)
    int* level = nullptr; size_t levelCap = 0;
    float4* cnodes = nullptr; size_t cnodeCap = 0;      // centre / half-extent copy of the chosen binary tree
    void* host = nullptr; size_t hostCap = 0;           // pinned block the build's small results are copied into (asynchronously)
};
void free_bvh_workspace(BvhWorkspace* ws);
struct BvhBuildResult {
    float4* nodes;               // device, 4 x float4 per node, root = 0 (points into the workspace: do not free)
    float4* cnodes;              // the same nodes with child boxes as centre / half-extent (node_visit_ch)
    long long nNodes;
    int depth;
    float buildMs;
    float extent[3];             // extent of the primitive centroids' bounds per axis (the N = 4 wide collapse orders slots along the two widest)
};
struct WideBuildResult {
    float4* wnodes = nullptr;    // device, 2 x float4 per child slot, `width` slots per node, node of binary node b at index b (root = 0)
    int width = 0, depth = 0, axes = 0;
    float buildMs = 0.f;
};
// collapses the binary hierarchy `bin` into a `width`-wide one (4 or 8); memory lives in the workspace
cudaError_t build_wide(const BvhBuildResult& bin, int width, BvhWorkspace* ws, WideBuildResult* out, cudaStream_t st);
// prim AABBs are computed on the device from the SoA arrays; pids = primitive ids of the bounded primitives
cudaError_t build_lbvh(const DevScene& sc, BvhWorkspace* ws, BvhBuildResult* out, cudaStream_t st);

}  // namespace brt
