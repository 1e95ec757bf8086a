// Multi-GPU behind the C ABI: the samples-per-pixel split with the exchange fused into its consumer.
//
// A *peer group* is N contexts, one per GPU, each owning one device block [accum 0 | accum 1 | rgba8 | floatData | linear |
// flags] that all the others map — through CUDA IPC handles when every GPU has its own process (brt_peer_alloc /
// brt_peer_connect: torchrun, MPI, one Node worker per GPU), or through plain peer access when one process drives all GPUs
// (brt_create_multi: the form a Node host binds — RayTracer.render() stays ONE call, ray-tracer.js:166-281 / ui-controller.js:189).
// brt_peer_render(rank) = zero this epoch's sums, trace this rank's sample range, then ONE kernel (k_peer_reduce_resolve,
// post.cu) that publishes "my sums are ready", waits for every peer's flag with acquire loads over NVLink, pulls its row
// stripe of all N buffers with 128-bit loads, sums in rank order, tone-maps and stores RGBA8 straight into the root's
// image.  No host barrier, no NCCL call, no separate reduce: the only cross-GPU traffic is (N-1)/N of the fp32 sums inbound
// per GPU plus the RGBA8 stripes outbound, and it is ordered by three flag words per rank.
#include <chrono>
#include <cstring>
#include "brt_ctx.hpp"

using namespace brt;

namespace {

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

void layout(PeerGroup& g, int W, int H) {
    const size_t px = (size_t)W * H;
    size_t off = 0;
    g.offAccum[0] = off; off += align256(px * 16);
    g.offAccum[1] = off; off += align256(px * 16);
    g.offRgba = off; off += align256(px * 4);
    g.offFloat = off; off += align256(px * 16);
    g.offLinear = off; off += align256(px * 16);
    g.offFlags = off; off += align256(FLAG_WORDS * sizeof(unsigned));
    g.blockBytes = off; g.W = W; g.H = H;
}

int group_alloc(brt_ctx* ctx, int rank, int world, bool ipc) {
    NEED_GPU();
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return fail(ctx, BRT_E_INVALID, "peer group: bad rank / world (1..16 ranks)");
    CK(cudaSetDevice(ctx->device));
    peer_release(ctx);
    PeerGroup& g = ctx->pg;
    g = PeerGroup{};
    g.rank = rank; g.world = world; g.ipc = ipc;
    layout(g, ctx->rp.width, ctx->rp.height);
    CK(cudaMalloc(&g.block, g.blockBytes));
    CK(cudaMemsetAsync((char*)g.block + g.offFlags, 0, FLAG_WORDS * sizeof(unsigned), ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    g.peerBlock[rank] = g.block;
    return BRT_OK;
}

void row_stripe(int H, int rank, int world, int& r0, int& r1) {
    const int base = H / world, rem = H % world;
    r0 = rank * base + (rank < rem ? rank : rem);
    r1 = r0 + base + (rank < rem ? 1 : 0);
}

PeerSync sync_of(const PeerGroup& g) {
    PeerSync ps{};
    ps.nPeers = g.world; ps.self = g.rank; ps.epoch = g.epoch;
    for (int r = 0; r < g.world; r++) {
        ps.accum[r] = (const float4*)((const char*)g.peerBlock[r] + g.offAccum[g.epoch & 1]);
        ps.flags[r] = (unsigned*)((char*)g.peerBlock[r] + g.offFlags);
    }
    return ps;
}

int check_group(brt_ctx* ctx) {
    const PeerGroup& g = ctx->pg;
    if (!g.block || !g.connected) return fail(ctx, BRT_E_STATE, "peer group not allocated / connected");
    if (g.W != ctx->rp.width || g.H != ctx->rp.height) return fail(ctx, BRT_E_STATE, "image size changed since brt_peer_alloc: allocate the peer group again");
    return BRT_OK;
}

// Exchange step of one rank (asynchronous, launches only — nothing here allocates or synchronises): advance the epoch, put
// this rank's sums into the epoch's buffer — traced straight into it (`running` == nullptr: zero, trace [sBegin, sBegin +
// sCount)) or copied from the rank's running sums of a multi-batch render — then the fused exchange + resolve kernel.
int peer_exchange(brt_ctx* ctx, PTParams& p, const float* running, int sBegin, int sCount, bool wantFloat, bool wantLinear) {
    PeerGroup& g = ctx->pg;
    g.epoch++;
    const size_t px = (size_t)g.W * g.H;
    float* acc = (float*)((char*)g.block + g.offAccum[g.epoch & 1]);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    if (running) CK(cudaMemcpyAsync(acc, running, px * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    else {
        CK(cudaMemsetAsync(acc, 0, px * 16, ctx->stream));
        int rc;
        if (sCount > 0 && (rc = launch_samples(ctx, p, acc, sBegin, sCount)) != BRT_OK) return rc;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    int r0, r1;
    row_stripe(g.H, g.rank, g.world, r0, r1);
    char* root = (char*)g.peerBlock[0];
    const bool dn = ctx->rp.denoise != 0;
    CK(launch_peer_reduce_resolve(post_params(ctx), sync_of(g), dn ? nullptr : (uchar4*)(root + g.offRgba),
                                  (wantFloat || dn) ? (float4*)(root + g.offFloat) : nullptr, wantLinear ? (float4*)(root + g.offLinear) : nullptr,
                                  r0, r1, ctx->stream));
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    ctx->stats.launches += 1;
    return BRT_OK;
}

// root: wait until every rank's stripe of this epoch is in the root's image (+ denoise), leave the image in the block
int peer_finish_root(brt_ctx* ctx) {
    PeerGroup& g = ctx->pg;
    CK(launch_peer_wait(sync_of(g), FLAG_DONE, ctx->stream));
    if (ctx->rp.denoise) {
        // post-processor.js:45-77 runs on the WHOLE tone-mapped image: after all stripes have arrived, on the root
        CK(launch_denoise(post_params(ctx), (const float4*)((char*)g.block + g.offFloat), (uchar4*)((char*)g.block + g.offRgba), nullptr, ctx->stream));
        ctx->stats.launches += 1;
    }
    ctx->stats.launches += 1;
    return BRT_OK;
}

int check_err_flag(brt_ctx* ctx) {
    unsigned e = 0;
    CK(cudaMemcpy(&e, (char*)ctx->pg.block + ctx->pg.offFlags + FLAG_ERR * sizeof(unsigned), 4, cudaMemcpyDeviceToHost));
    if (e) {
        cudaMemset((char*)ctx->pg.block + ctx->pg.offFlags + FLAG_ERR * sizeof(unsigned), 0, 4);
        return fail(ctx, BRT_E_STATE, "peer exchange timed out: a rank of the group did not reach this epoch");
    }
    return BRT_OK;
}

void sample_range(int n, int rank, int world, int& begin, int& count) {      // contiguous ranges, remainder to the low ranks
    const int base = n / world, rem = n % world;
    begin = rank * base + (rank < rem ? rank : rem);
    count = base + (rank < rem ? 1 : 0);
}

}  // namespace

namespace brt {

void peer_release(brt_ctx* ctx) {
    PeerGroup& g = ctx->pg;
    if (!g.block) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (g.ipc) for (int r = 0; r < g.world; r++) if (r != g.rank && g.peerBlock[r]) cudaIpcCloseMemHandle(g.peerBlock[r]);
    cudaFree(g.block);
    g = PeerGroup{};
}

// brt_render of a multi-device context: the same contract as the single-device call (ray-tracer.js:166-281) — batches,
// progress callback, preview, cancel — with every batch's samples split over the devices.
int render_multi(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean, brt_progress_cb cb, void* user) {
    auto w0 = std::chrono::steady_clock::now();
    ctx->cancel.store(0);
    std::vector<brt_ctx*> all;
    all.push_back(ctx);
    for (brt_ctx* f : ctx->followers) all.push_back(f);
    const int N = (int)all.size();
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    if (!ctx->haveCam) return fail(ctx, BRT_E_NOSCENE, "no camera set");
    auto on_device = [&](int r, int rc) { return r == 0 ? rc : fail(ctx, rc, "device " + std::to_string(all[r]->device) + ": " + all[r]->err); };
    auto sync_all = [&]() -> cudaError_t {
        for (int r = N - 1; r >= 0; r--) { cudaSetDevice(all[r]->device); cudaError_t e = cudaStreamSynchronize(all[r]->stream); if (e != cudaSuccess) return e; }
        return cudaSetDevice(ctx->device);
    };
    // mirror the leader's host state into the followers (the scene itself is shared by reference, never copied)
    for (int r = 1; r < N; r++) {
        brt_ctx* f = all[r];
        f->sceneRef = &ctx->scene; f->haveScene = true;
        if (f->syncedVersion != ctx->sceneVersion) { f->sceneDirty = true; f->obj64Dirty = true; f->syncedVersion = ctx->sceneVersion; }
        if (memcmp(f->bg.perm, ctx->bg.perm, 512) != 0) f->permDirty = true;
        f->bg = ctx->bg; f->cam = ctx->cam; f->haveCam = true; f->rp = ctx->rp;
    }
    const size_t px = (size_t)ctx->rp.width * ctx->rp.height;
    const int spp = effective_spp(ctx->rp);
    const int batch = spp_batch(ctx, spp, cb != nullptr);
    const bool oneBatch = batch >= spp;
    // ---- phase 1: everything that may allocate or synchronise (scene upload, LBVH build, buffers) happens before any kernel
    // that waits on a peer is in flight
    if (!ctx->pg.block || ctx->pg.W != ctx->rp.width || ctx->pg.H != ctx->rp.height || ctx->pg.world != N) {
        for (int r = 0; r < N; r++) { int rc = group_alloc(all[r], r, N, false); if (rc != BRT_OK) return on_device(r, rc); }
        for (int r = 0; r < N; r++) {
            for (int q = 0; q < N; q++) all[r]->pg.peerBlock[q] = all[q]->pg.block;
            all[r]->pg.connected = true;
        }
    }
    std::vector<PTParams> P(N);
    for (int r = 0; r < N; r++) {
        brt_ctx* c = all[r];
        int rc = prepare(c, P[r]);
        if (rc == BRT_OK) rc = reserve_launch_buffers(c, (batch + N - 1) / N);
        if (rc == BRT_OK && !oneBatch) {
            cudaError_t e = c->dAccum.ensure(px * 16);
            if (e == cudaSuccess) e = cudaMemsetAsync(c->dAccum.p, 0, px * 16, c->stream);
            if (e != cudaSuccess) rc = cuda_fail(c, e, "running sums");
        }
        if (rc == BRT_OK && c->rp.count_tests) {
            cudaError_t e = c->dCounters.ensure(N_COUNTERS * 8);
            if (e == cudaSuccess) e = cudaMemsetAsync(c->dCounters.p, 0, N_COUNTERS * 8, c->stream);
            if (e != cudaSuccess) rc = cuda_fail(c, e, "counters");
        }
        if (rc != BRT_OK) return on_device(r, rc);
        c->stats = brt_stats{};
    }
    CK(sync_all());
    // ---- phase 2: launches
    for (int traced = 0; traced < spp;) {
        const int n = spp - traced < batch ? spp - traced : batch;
        const bool last = traced + n >= spp;
        const bool exchange = last || (cb && ctx->rp.preview);        // an intermediate batch is exchanged only for a preview
        for (int r = 0; r < N; r++) {
            brt_ctx* c = all[r];
            cudaSetDevice(c->device);
            int b, k, rc = BRT_OK;
            sample_range(n, r, N, b, k);
            if (oneBatch) rc = peer_exchange(c, P[r], nullptr, traced + b, k, float_data != nullptr, linear_mean != nullptr);
            else {
                // several batches: each device keeps running sums of its share and exchanges a copy of them when an image is due
                if (k > 0) rc = launch_samples(c, P[r], (float*)c->dAccum.p, traced + b, k);
                if (rc == BRT_OK && exchange) rc = peer_exchange(c, P[r], (const float*)c->dAccum.p, 0, 0, float_data != nullptr, linear_mean != nullptr);
            }
            if (rc != BRT_OK) return on_device(r, rc);
        }
        traced += n;
        CK(cudaSetDevice(ctx->device));
        if (exchange) { int rc = peer_finish_root(ctx); if (rc != BRT_OK) return rc; }
        if (cb || !last) {
            // progress + cooperative cancel between batches (ray-tracer.js:190,256-261)
            CK(sync_all());
            if (ctx->cancel.load()) return fail(ctx, BRT_E_CANCELLED, "render cancelled");
            if (cb && !last) {
                if (ctx->rp.preview) {
                    CK(cudaMemcpyAsync(rgba8, (char*)ctx->pg.block + ctx->pg.offRgba, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream));
                }
                cb((double)traced / spp, user);
            }
        }
    }
    const PeerGroup& g = ctx->pg;
    CK(cudaMemcpyAsync(rgba8, (char*)g.block + g.offRgba, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (float_data) CK(cudaMemcpyAsync(float_data, (char*)g.block + g.offFloat, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (linear_mean) CK(cudaMemcpyAsync(linear_mean, (char*)g.block + g.offLinear, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(sync_all());
    int rc = check_err_flag(ctx);
    if (rc != BRT_OK) return rc;
    if (ctx->cancel.load()) return fail(ctx, BRT_E_CANCELLED, "render cancelled");
    double kernelMs = 0, postMs = 0;
    uint64_t launches = 0;
    for (int r = 0; r < N; r++) {
        float k = 0;
        cudaSetDevice(all[r]->device);
        if (cudaEventElapsedTime(&k, all[r]->ev0, all[r]->ev1) == cudaSuccess && k > kernelMs) kernelMs = k;
        if (cudaEventElapsedTime(&k, all[r]->ev1, all[r]->ev2) == cudaSuccess && k > postMs) postMs = k;
        launches += all[r]->stats.launches;
    }
    cudaGetLastError();
    cudaSetDevice(ctx->device);
    ctx->stats.kernel_ms = kernelMs; ctx->stats.post_ms = postMs; ctx->stats.launches = launches;
    ctx->stats.samples = (uint64_t)px * (uint64_t)spp;
    ctx->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    if (cb) cb(1.0, user);                                          // ray-tracer.js:279
    return BRT_OK;
}

}  // namespace brt

extern "C" {

int brt_create_multi(brt_ctx** out, const int* device_ids, int n_devices) {
    if (!out || !device_ids || n_devices < 1 || n_devices > MAX_PEERS) return BRT_E_INVALID;
    *out = nullptr;
    for (int i = 0; i < n_devices; i++) for (int j = 0; j < i; j++) if (device_ids[i] == device_ids[j]) return BRT_E_INVALID;
    brt_ctx* leader = nullptr;
    int rc = brt_create(&leader, device_ids[0]);
    if (rc != BRT_OK) return rc;
    for (int i = 1; i < n_devices; i++) {
        brt_ctx* f = nullptr;
        rc = brt_create(&f, device_ids[i]);
        if (rc != BRT_OK) { brt_destroy(leader); return rc; }
        f->leader = leader;
        leader->followers.push_back(f);
    }
    // every device maps every other device's block with plain loads / stores over NVLink
    for (int i = 0; i < n_devices; i++) {
        for (int j = 0; j < n_devices; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_ids[i], device_ids[j]) != cudaSuccess || !can) { brt_destroy(leader); return BRT_E_CUDA; }
            cudaSetDevice(device_ids[i]);
            cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { brt_destroy(leader); return BRT_E_CUDA; }
        }
    }
    cudaSetDevice(device_ids[0]);
    *out = leader;
    return BRT_OK;
}

int brt_device_count(const brt_ctx* ctx) { return ctx ? (ctx->device < 0 ? 0 : 1 + (int)ctx->followers.size()) : 0; }

int brt_peer_alloc(brt_ctx* ctx, int rank, int world, uint8_t handle[64]) {
    if (!ctx || !handle) return BRT_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    int rc = group_alloc(ctx, rank, world, true);
    if (rc != BRT_OK) return rc;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ctx->pg.block);
    if (e != cudaSuccess) { peer_release(ctx); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    return BRT_OK;
}

int brt_peer_connect(brt_ctx* ctx, const uint8_t* handles) {
    if (!ctx || !handles) return BRT_E_INVALID;
    NEED_GPU();
    PeerGroup& g = ctx->pg;
    if (!g.block || !g.ipc) return fail(ctx, BRT_E_STATE, "brt_peer_alloc first");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < g.world; r++) {
        if (r == g.rank || g.peerBlock[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        CK(cudaIpcOpenMemHandle(&g.peerBlock[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    g.connected = true;
    return BRT_OK;
}

int brt_peer_render(brt_ctx* ctx, int sample_begin, int sample_count, int want_float_data, int want_linear_mean) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    if (sample_begin < 0 || sample_count < 0 || (long long)sample_begin + sample_count > (1LL << 24)) return fail(ctx, BRT_E_INVALID, "bad sample range");
    CK(cudaSetDevice(ctx->device));
    int rc = check_group(ctx);
    if (rc != BRT_OK) return rc;
    PTParams p;
    if ((rc = prepare(ctx, p)) != BRT_OK) return rc;
    if ((rc = reserve_launch_buffers(ctx, sample_count)) != BRT_OK) return rc;
    if (ctx->rp.count_tests) { CK(ctx->dCounters.ensure(N_COUNTERS * 8)); CK(cudaMemsetAsync(ctx->dCounters.p, 0, N_COUNTERS * 8, ctx->stream)); }
    ctx->stats.launches = 0;
    if ((rc = peer_exchange(ctx, p, nullptr, sample_begin, sample_count, want_float_data != 0, want_linear_mean != 0)) != BRT_OK) return rc;
    ctx->stats.samples = (uint64_t)ctx->pg.W * ctx->pg.H * (uint64_t)sample_count;
    if (ctx->pg.rank == 0) return peer_finish_root(ctx);
    return BRT_OK;
}

int brt_peer_fetch(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    int rc = check_group(ctx);
    if (rc != BRT_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const PeerGroup& g = ctx->pg;
    const size_t px = (size_t)g.W * g.H;
    if (g.rank == 0) {
        if (rgba8) CK(cudaMemcpyAsync(rgba8, (char*)g.block + g.offRgba, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (float_data) CK(cudaMemcpyAsync(float_data, (char*)g.block + g.offFloat, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (linear_mean) CK(cudaMemcpyAsync(linear_mean, (char*)g.block + g.offLinear, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    float k = 0, q = 0;
    if (cudaEventElapsedTime(&k, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->stats.kernel_ms = k;
    if (cudaEventElapsedTime(&q, ctx->ev1, ctx->ev2) == cudaSuccess) ctx->stats.post_ms = q;
    cudaGetLastError();
    return check_err_flag(ctx);
}

int brt_peer_image_ptr(brt_ctx* ctx, void** d_rgba8) {
    if (!ctx || !d_rgba8) return BRT_E_INVALID;
    if (!ctx->pg.block) return fail(ctx, BRT_E_STATE, "brt_peer_alloc first");
    *d_rgba8 = (char*)ctx->pg.block + ctx->pg.offRgba;
    return BRT_OK;
}

int brt_peer_free(brt_ctx* ctx) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    peer_release(ctx);
    return BRT_OK;
}

}  // extern "C"
