// C ABI of libbrt (include/brt.h): context, scene upload (SoA flattening), LBVH build, render orchestration.
// Replaces RayTracer.render() (reference js/ray-tracer.js:166-281).  No CPU fallback exists: every compute entry
// point needs a CUDA device.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "brt_ctx.hpp"

using namespace brt;

namespace brt {
int fail(brt_ctx* c, int code, const std::string& msg) { if (c) c->err = msg; return code; }
int cuda_fail(brt_ctx* c, cudaError_t e, const char* where) {
    return fail(c, BRT_E_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}
}

// Host-side work over large meshes (float64 -> fp32 flattening, the defensive copy of a borrowed mesh) is split over a few
// threads: at 1 M triangles it is otherwise the largest part of a scene (re)load (25 ms single-threaded vs 3 ms of DMA).
template <class F>
static void parallel_chunks(size_t n, size_t minPerThread, F fn) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = hw ? hw : 1;
    if (nt > 8) nt = 8;
    if (n / (minPerThread ? minPerThread : 1) < nt) nt = n / (minPerThread ? minPerThread : 1);
    if (nt <= 1) { fn((size_t)0, n); return; }
    std::vector<std::thread> th;
    const size_t per = (n + nt - 1) / nt;
    for (size_t k = 0; k < nt; k++) {
        const size_t lo = k * per, hi = lo + per < n ? lo + per : n;
        if (lo < hi) th.emplace_back([=] { fn(lo, hi); });
    }
    for (std::thread& t : th) t.join();
}

static void default_params(brt_render_params& p) {                 // ray-tracer.js:19-30
    memset(&p, 0, sizeof(p));
    p.width = 600; p.height = 400; p.spp = 4; p.max_depth = 5; p.aa_mode = BRT_AA_SUPERSAMPLING; p.tonemap = BRT_TONEMAP_REINHARD;
    p.exposure = 1.0; p.gamma = 2.2; p.denoise = 0; p.denoise_strength = 0.5; p.seed = 1;
}
static void identity_perm(uint8_t* perm) { for (int i = 0; i < 512; i++) perm[i] = (uint8_t)(i & 255); }

extern "C" {

int brt_abi_version(void) { return BRT_ABI_VERSION; }
const char* brt_version(void) { return "libbrt 0.1 (sm_100a)"; }

int brt_create(brt_ctx** out, int device_id) {
    if (!out) return BRT_E_INVALID;
    *out = nullptr;
    if (device_id == -1) {
        // host-only context: scene ingest / camera / parameter logic without a device (used by CPU-side tests and
        // tooling).  Every compute entry point on such a context fails with BRT_E_CUDA — there is no CPU renderer.
        brt_ctx* h = new (std::nothrow) brt_ctx();
        if (!h) return BRT_E_NOMEM;
        h->device = -1;
        default_params(h->rp);
        identity_perm(h->bg.perm);
        *out = h;
        return BRT_OK;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return BRT_E_CUDA;             // no CPU fallback, by design
    if (device_id < 0 || device_id >= n) return BRT_E_INVALID;
    if (cudaSetDevice(device_id) != cudaSuccess) return BRT_E_CUDA;
    brt_ctx* ctx = new (std::nothrow) brt_ctx();
    if (!ctx) return BRT_E_NOMEM;
    ctx->device = device_id;
    default_params(ctx->rp);
    identity_perm(ctx->bg.perm);
    if (cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BRT_E_CUDA; }
    ctx->stream = ctx->ownStream;
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->ev2);
    *out = ctx;
    return BRT_OK;
}

void brt_destroy(brt_ctx* ctx) {
    if (!ctx) return;
    if (ctx->device < 0) { delete ctx; return; }
    for (brt_ctx* f : ctx->followers) { if (f) { cudaSetDevice(f->device); cudaStreamSynchronize(f->stream); } }
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (brt_ctx* f : ctx->followers) brt_destroy(f);
    ctx->followers.clear();
    cudaSetDevice(ctx->device);
    peer_release(ctx);
    DevBuf* bufs[] = { &ctx->dArena, &ctx->dPerm, &ctx->dPrim64,
                       &ctx->dAccum, &ctx->dRgba, &ctx->dFloat, &ctx->dFloat2, &ctx->dLinear, &ctx->dCounters, &ctx->dScratch, &ctx->dPlanes,
                       &ctx->dObj64, &ctx->dTris64 };
    for (DevBuf* b : bufs) b->release();
    if (ctx->stagePending) { cudaEventSynchronize(ctx->evStage); ctx->stagePending = false; }
    if (ctx->evStage) { cudaEventDestroy(ctx->evStage); ctx->evStage = nullptr; }
    ctx->hStage.release();
    free_bvh_workspace(&ctx->bvhWs);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    delete ctx;
}

const char* brt_last_error(const brt_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int brt_set_stream(brt_ctx* ctx, void* s) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    ctx->stream = s ? (cudaStream_t)s : ctx->ownStream;
    return BRT_OK;
}
int brt_stream_synchronize(brt_ctx* ctx) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

// ------------------------------------------------------------------------------------------- scene
static int validate_scene(brt_ctx* ctx, const HostScene& s, uint64_t nMeshTris = ~0ull) {
    for (size_t i = 0; i < s.objects.size(); i++) {
        const brt_object& o = s.objects[i];
        if (o.type < BRT_OBJ_SPHERE || o.type > BRT_OBJ_MESH) return fail(ctx, BRT_E_INVALID, "object " + std::to_string(i) + ": bad type");
        if (o.material < 0 || (size_t)o.material >= s.materials.size()) return fail(ctx, BRT_E_INVALID, "object " + std::to_string(i) + ": bad material index");
        if (o.type == BRT_OBJ_MESH) {
            const uint64_t nT = nMeshTris != ~0ull ? nMeshTris : s.meshTris.size() / 9;   // no arithmetic on the untrusted values: nothing can wrap
            if (o.first_tri < 0 || o.tri_count < 0 || (uint64_t)o.first_tri > nT || (uint64_t)o.tri_count > nT - (uint64_t)o.first_tri)
                return fail(ctx, BRT_E_INVALID, "object " + std::to_string(i) + ": mesh triangle range out of bounds");
        }
    }
    for (const brt_material& m : s.materials) {
        if (m.type < 0 || m.type > 3) return fail(ctx, BRT_E_INVALID, "bad material type");
        if (m.texture < 0 || (size_t)m.texture > s.textures.size()) return fail(ctx, BRT_E_INVALID, "bad texture index");
    }
    for (const brt_texture& t : s.textures) if (t.kind < BRT_TEX_SOLID || t.kind > BRT_TEX_WOOD) return fail(ctx, BRT_E_INVALID, "bad texture kind");
    for (const brt_light& l : s.lights) if (l.type < 0 || l.type > 1) return fail(ctx, BRT_E_INVALID, "bad light type");
    return BRT_OK;
}

static int scene_load_common(brt_ctx* ctx, const void* data, size_t len, bool binary, int fw, int fh, int* out_has_camera, int* out_w, int* out_h) {
    if (!ctx || !data) return BRT_E_INVALID;
    if (fw <= 0 || fh <= 0) return fail(ctx, BRT_E_INVALID, "fallback width/height must be positive");
    HostScene sc; bool hasCam = false; int w = 0, h = 0; std::string err;
    brt_camera cam = ctx->cam; HostBackground bg = ctx->bg;
    int rc = binary ? load_scene_binary((const unsigned char*)data, len, fw, fh, sc, bg, cam, hasCam, w, h, err)
                    : load_scene_json((const char*)data, len, fw, fh, sc, bg, cam, hasCam, w, h, err);
    if (rc != BRT_OK) return fail(ctx, rc, err);
    if ((rc = validate_scene(ctx, sc)) != BRT_OK) return rc;
    ctx->scene = std::move(sc); ctx->haveScene = true; ctx->sceneDirty = true; ctx->obj64Dirty = true; ctx->sceneVersion++;
    ctx->bg = bg;
    if (hasCam) { ctx->cam = cam; ctx->haveCam = true; }           // ray-tracer.js:315-317
    if (out_has_camera) *out_has_camera = hasCam ? 1 : 0;
    if (out_w) *out_w = w;
    if (out_h) *out_h = h;
    return BRT_OK;
}
int brt_scene_load_json(brt_ctx* ctx, const char* utf8, size_t len, int fw, int fh, int* out_has_camera, int* out_w, int* out_h) {
    return scene_load_common(ctx, utf8, len, false, fw, fh, out_has_camera, out_w, out_h);
}
int brt_scene_load_binary(brt_ctx* ctx, const void* bytes, size_t len, int fw, int fh, int* out_has_camera, int* out_w, int* out_h) {
    return scene_load_common(ctx, bytes, len, true, fw, fh, out_has_camera, out_w, out_h);
}

int brt_scene_set_flat(brt_ctx* ctx, const brt_scene_desc* d) {
    if (!ctx || !d) return BRT_E_INVALID;
    if (d->n_objects < 0 || d->n_materials < 0 || d->n_lights < 0 || d->n_mesh_triangles < 0 || d->n_textures < 0) return fail(ctx, BRT_E_INVALID, "negative count");
    if (d->n_textures && !d->textures) return fail(ctx, BRT_E_INVALID, "null array with non-zero count");
    if ((d->n_objects && !d->objects) || (d->n_materials && !d->materials) || (d->n_lights && !d->lights) || (d->n_mesh_triangles && !d->mesh_triangles))
        return fail(ctx, BRT_E_INVALID, "null array with non-zero count");
    HostScene sc;
    sc.objects.assign(d->objects, d->objects + d->n_objects);
    sc.materials.assign(d->materials, d->materials + d->n_materials);
    sc.lights.assign(d->lights, d->lights + d->n_lights);
    if (d->n_textures) sc.textures.assign(d->textures, d->textures + d->n_textures);
    const bool raw = !(d->flags & BRT_SCENE_CONSTRUCTED);             // constructor arguments: apply what the reference's constructors do
    if (raw) for (brt_object& o : sc.objects) if (o.type == BRT_OBJ_PLANE) {      // geometry.js:52
        double l = std::sqrt(o.b[0] * o.b[0] + o.b[1] * o.b[1] + o.b[2] * o.b[2]);
        if (l > 0) { o.b[0] /= l; o.b[1] /= l; o.b[2] /= l; } else o.b[0] = o.b[1] = o.b[2] = 0;
    }
    if (raw) for (brt_material& m : sc.materials) if (m.type == BRT_MAT_METAL && !(m.param != m.param)) m.param = std::fmin(m.param, 1.0);   // materials.js:33
    if (raw) for (brt_light& l : sc.lights) if (l.type == BRT_LIGHT_DIRECTIONAL) {   // lights.js:38
        double n = std::sqrt(l.v[0] * l.v[0] + l.v[1] * l.v[1] + l.v[2] * l.v[2]);
        if (n > 0) { l.v[0] /= n; l.v[1] /= n; l.v[2] /= n; } else l.v[0] = l.v[1] = l.v[2] = 0;
    }
    int rc = validate_scene(ctx, sc, (uint64_t)d->n_mesh_triangles);
    if (rc != BRT_OK) return rc;                                    // the previous scene stays in place
    // the mesh triangles (72 B each) are copied last, in parallel, into the ctx's own buffer: re-sending a scene of the same
    // size touches no allocator (a fresh 72 MB vector costs more in page faults than the copy itself)
    std::vector<double> keep;
    keep.swap(ctx->scene.meshTris);
    keep.resize(9 * (size_t)d->n_mesh_triangles);
    {
        const double* src = d->mesh_triangles; double* dst = keep.data();
        parallel_chunks(9 * (size_t)d->n_mesh_triangles, 1u << 20, [=](size_t lo, size_t hi) { memcpy(dst + lo, src + lo, (hi - lo) * sizeof(double)); });
    }
    sc.meshTris.swap(keep);
    ctx->scene = std::move(sc); ctx->haveScene = true; ctx->sceneDirty = true; ctx->obj64Dirty = true; ctx->sceneVersion++;
    return BRT_OK;
}

// The pinned staging buffer is written by the host and read by an asynchronous copy: an event marks the copy, and the next writer
// (or whoever frees the buffer) waits for it — instead of a stream synchronisation inside every scene upload.
static cudaError_t stage_mark(brt_ctx* ctx) {
    if (!ctx->evStage) { cudaError_t e = cudaEventCreateWithFlags(&ctx->evStage, cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    ctx->stagePending = true;
    return cudaEventRecord(ctx->evStage, ctx->stream);
}
static cudaError_t stage_wait(brt_ctx* ctx) {
    if (!ctx->stagePending) return cudaSuccess;
    ctx->stagePending = false;
    return cudaEventSynchronize(ctx->evStage);
}

static float4 f4(double x, double y, double z, double w) { return make_float4((float)x, (float)y, (float)z, (float)w); }

// Flatten world.objects into per-type SoA float4 arrays + unified meta.  The arrays are written straight into ONE pinned
// staging buffer laid out like the device arena, so the whole scene goes up in a single host-to-device copy (the 1.0 M
// triangle scene: 64 MB, one DMA from page-locked memory).  The float64 primitive copy (prim64, 72 B / primitive) is only
// needed where primary hits are re-evaluated in float64 and is uploaded on first use (ensure_prim64).
static int upload_scene(brt_ctx* ctx) {
    NEED_GPU();
    if (!ctx->sceneDirty) return BRT_OK;
    CK(cudaSetDevice(ctx->device));
    auto t0 = std::chrono::steady_clock::now();
    const HostScene& s = ctx->hostScene();
    size_t nSph = 0, nPln = 0, nBox = 0, nTri = 0;
    for (const brt_object& o : s.objects) {
        if (o.type == BRT_OBJ_SPHERE) nSph++; else if (o.type == BRT_OBJ_PLANE) nPln++; else if (o.type == BRT_OBJ_BOX) nBox++;
        else nTri += o.type == BRT_OBJ_TRIANGLE ? 1 : (size_t)o.tri_count;
    }
    if (nTri >= (1u << 28) || s.objects.size() >= (1u << 28)) return fail(ctx, BRT_E_INVALID, "too many primitives (limit 2^28 per type)");
    const size_t nMat = s.materials.size(), nLights = s.lights.size(), nTex = s.textures.size(), nPrim = nSph + nPln + nBox + nTri;
    // arena layout (256-byte aligned sub-arrays)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    const size_t oSph = take(nSph * 16), oPln = take(nPln * 32), oBox = take(nBox * 32), oTri = take(nTri * 48), oMeta = take(nPrim * 16),
                 oMat = take(nMat * 16), oMatType = take(nMat * 4), oLights = take(nLights * 32), oTex = take(nTex * 32), oTexPerm = take(nTex * 512);
    const size_t arenaBytes = off;
    CK(stage_wait(ctx));                                             // the previous upload may still be reading the staging buffer
    CK(ctx->hStage.ensure(arenaBytes));
    CK(ctx->dArena.ensure(arenaBytes));
    char* H = (char*)ctx->hStage.p;
    float4 *sph = (float4*)(H + oSph), *pln = (float4*)(H + oPln), *box = (float4*)(H + oBox), *tri = (float4*)(H + oTri);
    float4 *mat = (float4*)(H + oMat), *lights = (float4*)(H + oLights), *tex = (float4*)(H + oTex);
    int4* meta = (int4*)(H + oMeta);
    int* matType = (int*)(H + oMatType);
    unsigned char* texPerm = (unsigned char*)(H + oTexPerm);
    DevScene& d = ctx->dev;
    d.baseSph = 0; d.basePln = (int)nSph; d.baseBox = (int)(nSph + nPln); d.baseTri = (int)(nSph + nPln + nBox);
    size_t iS = 0, iP = 0, iB = 0, iT = 0;
    auto put_tri = [&](const double* v0, const double* v1, const double* v2, int obj, int m, int triId) {
        float4* t = tri + 3 * iT;
        t[0] = f4(v0[0], v0[1], v0[2], 0);
        // edges are formed in float64 and rounded once (geometry.js:150-151 recomputes them per hit in float64)
        t[1] = f4(v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2], 0);
        t[2] = f4(v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2], 0);
        meta[d.baseTri + iT] = make_int4(obj, m, triId, 0);
        iT++;
    };
    for (size_t i = 0; i < s.objects.size(); i++) {
        const brt_object& o = s.objects[i];
        const int obj = (int)i;
        switch (o.type) {
        case BRT_OBJ_SPHERE: sph[iS] = f4(o.a[0], o.a[1], o.a[2], o.b[0]); meta[d.baseSph + iS] = make_int4(obj, o.material, -1, 0); iS++; break;
        case BRT_OBJ_PLANE: pln[2 * iP] = f4(o.b[0], o.b[1], o.b[2], 0); pln[2 * iP + 1] = f4(o.a[0], o.a[1], o.a[2], 0); meta[d.basePln + iP] = make_int4(obj, o.material, -1, 0); iP++; break;
        case BRT_OBJ_BOX: box[2 * iB] = f4(o.a[0], o.a[1], o.a[2], 0); box[2 * iB + 1] = f4(o.b[0], o.b[1], o.b[2], 0); meta[d.baseBox + iB] = make_int4(obj, o.material, -1, 0); iB++; break;
        case BRT_OBJ_TRIANGLE: put_tri(o.a, o.b, o.c, obj, o.material, -1); break;
        default: {
            // a mesh: its triangles are independent rows of the SoA arrays, filled in parallel
            const size_t base = iT, first = (size_t)o.first_tri;
            const int mat_ = o.material;
            const int baseTri = d.baseTri;
            const double* src = s.meshTris.data();
            parallel_chunks((size_t)o.tri_count, 65536, [=](size_t lo, size_t hi) {
                for (size_t t = lo; t < hi; t++) {
                    const double* p = src + 9 * (first + t);
                    float4* q = tri + 3 * (base + t);
                    q[0] = f4(p[0], p[1], p[2], 0);
                    q[1] = f4(p[3] - p[0], p[4] - p[1], p[5] - p[2], 0);
                    q[2] = f4(p[6] - p[0], p[7] - p[1], p[8] - p[2], 0);
                    meta[baseTri + base + t] = make_int4(obj, mat_, (int)t, 0);
                }
            });
            iT += (size_t)o.tri_count;
        }
        }
    }
    // material type in the low byte, 1-based texture index above it (TexturedLambertian / TexturedMetal)
    for (size_t i = 0; i < nMat; i++) { const brt_material& m = s.materials[i]; mat[i] = f4(m.color[0], m.color[1], m.color[2], m.param); matType[i] = m.type | (m.texture << 8); }
    for (size_t i = 0; i < nTex; i++) {
        const brt_texture& t = s.textures[i];
        tex[2 * i] = f4(t.odd[0], t.odd[1], t.odd[2], (double)t.kind);
        tex[2 * i + 1] = f4(t.even[0], t.even[1], t.even[2], t.scale);
        for (int k = 0; k < 512; k++) texPerm[512 * i + k] = t.perm[k & 255];      // doubled table (noise.js:16-17)
    }
    for (size_t i = 0; i < nLights; i++) {
        const brt_light& l = s.lights[i];
        lights[2 * i] = f4(l.v[0], l.v[1], l.v[2], l.type == BRT_LIGHT_DIRECTIONAL ? 1.0 : 0.0);
        lights[2 * i + 1] = f4(l.color[0] * l.intensity, l.color[1] * l.intensity, l.color[2] * l.intensity, 0);
    }
    d.nSph = (int)nSph; d.nPln = (int)nPln; d.nBox = (int)nBox; d.nTri = (int)nTri;
    d.nLights = (int)nLights; d.nTex = (int)nTex;
    if (arenaBytes) CK(cudaMemcpyAsync(ctx->dArena.p, H, arenaBytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(stage_mark(ctx));                                             // no host synchronisation here: whoever writes the staging buffer next waits (stage_wait)
    const char* D = (const char*)ctx->dArena.p;
    d.sph = (const float4*)(D + oSph); d.pln = (const float4*)(D + oPln); d.box = (const float4*)(D + oBox); d.tri = (const float4*)(D + oTri);
    d.meta = (const int4*)(D + oMeta); d.mat = (const float4*)(D + oMat); d.matType = (const int*)(D + oMatType);
    d.lights = (const float4*)(D + oLights); d.tex = (const float4*)(D + oTex); d.texPerm = (const unsigned char*)(D + oTexPerm);
    d.prim64 = nullptr; ctx->prim64Dirty = true;
    auto t1 = std::chrono::steady_clock::now();
    // the LBVH over the bounded primitives is built on first use (ensure_bvh): tiny scenes render with the linear loop
    d.nodes = nullptr; d.cnodes = nullptr; d.nNodes = 0; d.bvhStackDepth = 0;
    d.wnodes = nullptr; d.wideN = 0; d.wideDepth = 0; d.wideAxes = 0;
    ctx->bin = BvhBuildResult{}; ctx->wide = WideBuildResult{};
    ctx->nBounded = d.nSph + d.nBox + d.nTri;
    ctx->bvhDirty = true;
    brt_scene_info& inf = ctx->info;
    inf.n_objects = (int)s.objects.size(); inf.n_materials = (int)nMat; inf.n_lights = (int)nLights;
    inf.n_spheres = d.nSph; inf.n_planes = d.nPln; inf.n_boxes = d.nBox; inf.n_triangles = d.nTri;
    inf.n_bvh_nodes = 0; inf.bvh_depth = 0; inf.bvh_build_ms = 0; inf.bvh_width = 0; inf.bvh_wide_depth = 0; inf.bvh_wide_build_ms = 0;
    inf.upload_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    inf.upload_bytes = (int64_t)arenaBytes;
    ctx->sceneDirty = false;
    return BRT_OK;
}

// float64 copy of every primitive (9 doubles, unified index as meta): what refine_primary evaluates.  Uploaded only when a
// launch needs it (sampler = reference, the fp32 AOV kernel).
static int ensure_prim64(brt_ctx* ctx) {
    if (!ctx->prim64Dirty && ctx->dev.prim64) return BRT_OK;
    const HostScene& s = ctx->hostScene();
    const DevScene& d = ctx->dev;
    const size_t nPrim = (size_t)d.nSph + d.nPln + d.nBox + d.nTri;
    CK(stage_wait(ctx));
    CK(ctx->hStage.ensure(nPrim * 72));
    CK(ctx->dPrim64.ensure(nPrim * 72));
    double* q = (double*)ctx->hStage.p;
    memset(q, 0, nPrim * 72);
    size_t iS = 0, iP = 0, iB = 0, iT = 0;
    for (const brt_object& o : s.objects) {
        switch (o.type) {
        case BRT_OBJ_SPHERE: { double* r = q + 9 * (d.baseSph + iS++); memcpy(r, o.a, 24); memcpy(r + 3, o.b, 24); break; }
        case BRT_OBJ_PLANE: { double* r = q + 9 * (d.basePln + iP++); memcpy(r, o.b, 24); memcpy(r + 3, o.a, 24); break; }
        case BRT_OBJ_BOX: { double* r = q + 9 * (d.baseBox + iB++); memcpy(r, o.a, 24); memcpy(r + 3, o.b, 24); break; }
        case BRT_OBJ_TRIANGLE: { double* r = q + 9 * (d.baseTri + iT++); memcpy(r, o.a, 24); memcpy(r + 3, o.b, 24); memcpy(r + 6, o.c, 24); break; }
        default:
            if (o.tri_count > 0) memcpy(q + 9 * (d.baseTri + iT), &s.meshTris[9 * (size_t)o.first_tri], 72 * (size_t)o.tri_count);
            iT += (size_t)o.tri_count;
        }
    }
    if (nPrim) CK(cudaMemcpyAsync(ctx->dPrim64.p, q, nPrim * 72, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dev.prim64 = (const double*)ctx->dPrim64.p;
    ctx->prim64Dirty = false;
    return BRT_OK;
}

static int ensure_bvh(brt_ctx* ctx) {
    if (!ctx->bvhDirty) return BRT_OK;
    BvhBuildResult br{};
    CK(build_lbvh(ctx->dev, &ctx->bvhWs, &br, ctx->stream));
    if (br.depth + 1 > SMEM_STACK + LOCAL_STACK) return fail(ctx, BRT_E_STATE, "LBVH deeper than the traversal stack (" + std::to_string(br.depth) + ")");
    ctx->dev.nodes = br.nodes; ctx->dev.cnodes = br.cnodes; ctx->dev.nNodes = (int)br.nNodes; ctx->dev.bvhStackDepth = br.depth;
    ctx->info.n_bvh_nodes = br.nNodes; ctx->info.bvh_depth = br.depth; ctx->info.bvh_build_ms = br.buildMs;
    ctx->bin = br; ctx->wide = WideBuildResult{};
    ctx->dev.wnodes = nullptr; ctx->dev.wideN = 0; ctx->dev.wideDepth = 0; ctx->dev.wideAxes = 0;
    ctx->info.bvh_width = 2; ctx->info.bvh_wide_depth = 0; ctx->info.bvh_wide_build_ms = 0;
    ctx->bvhDirty = false;
    return BRT_OK;
}
// Width of the hierarchy the megakernel walks with the fast sampler: 2 = the binary LBVH itself, 4 / 8 = the wide collapse of it.
// BRT_BVH_WIDTH (environment) overrides the render parameter — kernel A/B runs; 0 / unset = automatic.
constexpr int WIDE_MAX_DEPTH = 32;                   // one shared-memory stack word per level and thread (16 KB / block at 32)
static int wanted_width(const brt_ctx* ctx) {
    int w = ctx->rp.bvh_width;
    if (const char* e = getenv("BRT_BVH_WIDTH")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) w = v; }
    if (w == 0) w = BRT_BVH_WIDTH_AUTO;
    if (ctx->rp.sampler != BRT_SAMPLER_FAST || ctx->rp.integrator == BRT_INTEGRATOR_WAVEFRONT) return 2;
    if (ctx->bin.nNodes >= (1LL << 24)) return 2;     // node ids share a stack word with the 8-bit pending mask
    return w;
}
static int ensure_wide(brt_ctx* ctx) {
    const int w = ctx->bvhDirty || ctx->dev.nNodes == 0 ? 2 : wanted_width(ctx);
    if (w != 4 && w != 8) { ctx->dev.wnodes = nullptr; ctx->dev.wideN = 0; ctx->info.bvh_width = 2; return BRT_OK; }
    if (ctx->wide.width != w) {
        CK(build_wide(ctx->bin, w, &ctx->bvhWs, &ctx->wide, ctx->stream));
        ctx->info.bvh_wide_depth = ctx->wide.depth; ctx->info.bvh_wide_build_ms = ctx->wide.buildMs;
    }
    if (!ctx->wide.wnodes || ctx->wide.depth > WIDE_MAX_DEPTH) {   // a degenerate chain of a tree: stay with the binary traversal (hybrid stack)
        ctx->dev.wnodes = nullptr; ctx->dev.wideN = 0; ctx->info.bvh_width = 2;
        return BRT_OK;
    }
    ctx->dev.wnodes = ctx->wide.wnodes; ctx->dev.wideN = w; ctx->dev.wideDepth = ctx->wide.depth; ctx->dev.wideAxes = ctx->wide.axes;
    ctx->info.bvh_width = w;
    return BRT_OK;
}
// BRUTE reproduces the reference's loops; AUTO takes the hierarchy from 6 bounded primitives up.  Measured on a B200
// (tools/accel_threshold.py, profiles/r02c_accel_threshold.json; 1280x720x32 spp): with 4-5 primitives — the reference's presets — the
// linear loop wins by 1-17 %, at 6 the hierarchy is 9 % ahead, at 8 22 %, at 16 73 %, at 64 4.5x.
static bool wants_bvh(const brt_ctx* ctx) {
    if (ctx->rp.accel == BRT_ACCEL_BRUTE) return false;
    if (ctx->rp.accel == BRT_ACCEL_BVH) return ctx->nBounded >= 2;
    return ctx->nBounded >= 6;
}

static int upload_perm(brt_ctx* ctx) {
    NEED_GPU();
    if (!ctx->permDirty && ctx->dPerm.p) { ctx->dev.perm = (const unsigned char*)ctx->dPerm.p; return BRT_OK; }
    CK(cudaSetDevice(ctx->device));
    CK(ctx->dPerm.ensure(512));
    CK(cudaMemcpyAsync(ctx->dPerm.p, ctx->bg.perm, 512, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->dev.perm = (const unsigned char*)ctx->dPerm.p;
    ctx->permDirty = false;
    return BRT_OK;
}

int brt_scene_info_get(brt_ctx* ctx, brt_scene_info* out) {
    if (!ctx || !out) return BRT_E_INVALID;
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    if (ctx->device < 0) {
        memset(out, 0, sizeof(*out));
        const HostScene& s = ctx->hostScene();
        out->n_objects = (int)s.objects.size(); out->n_materials = (int)s.materials.size(); out->n_lights = (int)s.lights.size();
        for (const brt_object& o : s.objects) {
            if (o.type == BRT_OBJ_SPHERE) out->n_spheres++; else if (o.type == BRT_OBJ_PLANE) out->n_planes++;
            else if (o.type == BRT_OBJ_BOX) out->n_boxes++; else if (o.type == BRT_OBJ_TRIANGLE) out->n_triangles++;
            else out->n_triangles += o.tri_count;
        }
        return BRT_OK;
    }
    int rc = upload_scene(ctx);
    if (rc != BRT_OK) return rc;
    if ((rc = ensure_bvh(ctx)) != BRT_OK) return rc;               // the query reports the hierarchy, so it builds it
    if ((rc = ensure_wide(ctx)) != BRT_OK) return rc;
    *out = ctx->info;
    return BRT_OK;
}
int brt_scene_get_flat(brt_ctx* ctx, brt_scene_desc* out) {
    if (!ctx || !out) return BRT_E_INVALID;
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    const HostScene& s = ctx->scene;
    memset(out, 0, sizeof(*out));
    out->objects = s.objects.data(); out->n_objects = (int)s.objects.size();
    out->materials = s.materials.data(); out->n_materials = (int)s.materials.size();
    out->mesh_triangles = s.meshTris.data(); out->n_mesh_triangles = (int64_t)(s.meshTris.size() / 9);
    out->lights = s.lights.data(); out->n_lights = (int)s.lights.size();
    out->textures = s.textures.data(); out->n_textures = (int)s.textures.size();
    out->flags = BRT_SCENE_CONSTRUCTED;                             // what the ctx holds is post-constructor: get -> set round-trips exactly
    return BRT_OK;
}

int brt_set_camera(brt_ctx* ctx, const brt_camera* cam) {
    if (!ctx || !cam) return BRT_E_INVALID;
    if (cam->type < BRT_CAM_PERSPECTIVE || cam->type > BRT_CAM_OTHER) return fail(ctx, BRT_E_INVALID, "bad camera type");
    ctx->cam = *cam;
    if (!cam->use_derived) derive_camera(ctx->cam);
    ctx->haveCam = true;
    return BRT_OK;
}
int brt_get_camera(brt_ctx* ctx, brt_camera* out) {
    if (!ctx || !out) return BRT_E_INVALID;
    if (!ctx->haveCam) return fail(ctx, BRT_E_NOSCENE, "no camera set");
    *out = ctx->cam;
    return BRT_OK;
}
int brt_set_background(brt_ctx* ctx, int kind, const double color[3], double intensity, const uint8_t* perm256) {
    if (!ctx) return BRT_E_INVALID;
    if (kind < BRT_BG_GRADIENT || kind > BRT_BG_PROCEDURAL_SKY) return fail(ctx, BRT_E_INVALID, "bad background kind");
    ctx->bg.kind = kind; ctx->bg.intensity = intensity;
    if (color) for (int k = 0; k < 3; k++) ctx->bg.color[k] = color[k];
    if (perm256) { for (int i = 0; i < 512; i++) ctx->bg.perm[i] = perm256[i & 255]; ctx->permDirty = true; }   // noise.js:16-17
    return BRT_OK;
}
int brt_get_background(brt_ctx* ctx, int* kind, double color[3], double* intensity) {
    if (!ctx) return BRT_E_INVALID;
    if (kind) *kind = ctx->bg.kind;
    if (color) for (int k = 0; k < 3; k++) color[k] = ctx->bg.color[k];
    if (intensity) *intensity = ctx->bg.intensity;
    return BRT_OK;
}
int brt_set_render_params(brt_ctx* ctx, const brt_render_params* p) {
    if (!ctx || !p) return BRT_E_INVALID;
    if (p->width < 1 || p->height < 1 || p->width > 65536 || p->height > 65536) return fail(ctx, BRT_E_INVALID, "bad image size");
    if (p->spp < 1) return fail(ctx, BRT_E_INVALID, "spp must be >= 1");
    if (p->max_depth < 0 || p->max_depth > 250) return fail(ctx, BRT_E_INVALID, "max_depth must be in [0, 250]");
    if (p->spp > (1 << 24)) return fail(ctx, BRT_E_INVALID, "spp must be <= 2^24");
    if (p->aa_mode < 0 || p->aa_mode > 3 || p->tonemap < 0 || p->tonemap > 2) return fail(ctx, BRT_E_INVALID, "bad aa_mode / tonemap");
    if (p->sampler < 0 || p->sampler > 1 || p->integrator < 0 || p->integrator > 2 || p->accel < 0 || p->accel > 2)
        return fail(ctx, BRT_E_INVALID, "bad sampler / integrator / accel");
    if (p->bvh_width != 0 && p->bvh_width != 2 && p->bvh_width != 4 && p->bvh_width != 8) return fail(ctx, BRT_E_INVALID, "bvh_width must be 0 (auto), 2, 4 or 8");
    ctx->rp = *p;
    return BRT_OK;
}
int brt_get_render_params(brt_ctx* ctx, brt_render_params* out) {
    if (!ctx || !out) return BRT_E_INVALID;
    *out = ctx->rp;
    return BRT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- render
static bool use_bvh(const brt_ctx* ctx) { return wants_bvh(ctx) && !ctx->bvhDirty && ctx->dev.nNodes > 0; }
namespace brt {
int effective_spp(const brt_render_params& rp) { return rp.aa_mode == BRT_AA_NONE ? 1 : rp.spp; }    // ray-tracer.js:201
int spp_batch(const brt_ctx* ctx, int spp, bool haveCallback) {
    int batch = ctx->rp.spp_batch > 0 ? ctx->rp.spp_batch : (haveCallback ? (spp + 15) / 16 : spp);
    if (ctx->rp.spp_batch <= 0) {
        // keep one launch to a few seconds at most so brt_cancel is honoured promptly (it is polled between launches):
        // about 4e9 path samples per launch
        const long long px = (long long)ctx->rp.width * ctx->rp.height;
        long long cap = 4000000000LL / (px ? px : 1);
        if (cap < 1) cap = 1;
        if (batch > cap) batch = (int)cap;
    }
    return batch < 1 ? 1 : batch;
}

int prepare(brt_ctx* ctx, PTParams& p) {
    NEED_GPU();
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    if (!ctx->haveCam) return fail(ctx, BRT_E_NOSCENE, "no camera set");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_scene(ctx);
    if (rc != BRT_OK) return rc;
    if (wants_bvh(ctx) && (rc = ensure_bvh(ctx)) != BRT_OK) return rc;
    if ((rc = ensure_wide(ctx)) != BRT_OK) return rc;
    if ((rc = upload_perm(ctx)) != BRT_OK) return rc;
    const brt_render_params& rp = ctx->rp;
    if (rp.sampler == BRT_SAMPLER_REFERENCE && (rc = ensure_prim64(ctx)) != BRT_OK) return rc;   // float64 primary-hit evaluation
    memset(&p, 0, sizeof(p));
    ctx->dev.bgKind = ctx->bg.kind; ctx->dev.bgR = (float)ctx->bg.color[0]; ctx->dev.bgG = (float)ctx->bg.color[1];
    ctx->dev.bgB = (float)ctx->bg.color[2]; ctx->dev.skyIntensity = (float)ctx->bg.intensity;
    p.sc = ctx->dev;
    const brt_camera& c = ctx->cam;
    DevCamera& dc = p.cam;
    for (int k = 0; k < 3; k++) {
        dc.o[k] = c.origin[k]; dc.ll[k] = c.lower_left_corner[k]; dc.h[k] = c.horizontal[k]; dc.v[k] = c.vertical[k];
        dc.cu[k] = c.u[k]; dc.cv[k] = c.v[k]; dc.cw[k] = c.w[k];
    }
    dc.lensRadius = c.lens_radius; dc.type = c.type;
    DevCamera32& d32 = p.cam32;
    for (int k = 0; k < 3; k++) {
        d32.o[k] = (float)c.origin[k]; d32.llo[k] = (float)(c.lower_left_corner[k] - c.origin[k]);
        d32.h[k] = (float)c.horizontal[k]; d32.v[k] = (float)c.vertical[k];
        d32.cu[k] = (float)c.u[k]; d32.cv[k] = (float)c.v[k]; d32.cw[k] = (float)c.w[k];
    }
    d32.lensRadius = (float)c.lens_radius; d32.type = c.type;
    p.W = rp.width; p.H = rp.height; p.maxDepth = rp.max_depth; p.aaMode = rp.aa_mode;
    p.rowBegin = 0; p.rowEnd = rp.height;
    if (const char* e = getenv("BRT_DEBUG_ROW_WINDOW")) {          // "a,b": trace rows [a, b) only (debug: brute force on a few rows of a big frame)
        int a = 0, b = 0;
        if (sscanf(e, "%d,%d", &a, &b) == 2 && a >= 0 && b > a) { p.rowBegin = a; p.rowEnd = b < rp.height ? b : rp.height; }
    }
    p.seedLo = (uint32_t)rp.seed; p.seedHi = (uint32_t)(rp.seed >> 32);
    p.directLighting = rp.direct_lighting ? 1 : 0;
    p.refill = rp.refill_threshold > 0 ? (rp.refill_threshold > 32 ? 32 : rp.refill_threshold) : 8;
    p.wavefront = rp.integrator == BRT_INTEGRATOR_WAVEFRONT ? 1 : 0;   // AUTO = megakernel: faster on every measured config
    p.inflight = (rp.paths_in_flight >= 1 && rp.paths_in_flight <= 4) ? rp.paths_in_flight : 2;
    return BRT_OK;
}

static int z_split(const brt_ctx* ctx, int samplesInLaunch) {
    // keep >= ~4 resident waves of threads on 148 SMs when the image is small (the chunks' planes are folded in fixed order)
    const long long want = 148LL * 2048 * 2;
    long long px = (long long)ctx->rp.width * ctx->rp.height;
    if (px >= want || samplesInLaunch < 2) return 1;
    long long z = (want + px - 1) / px;
    if (z > samplesInLaunch) z = samplesInLaunch;
    if (z > 64) z = 64;
    return (int)z;
}

int reserve_launch_buffers(brt_ctx* ctx, int maxSamplesPerLaunch) {
    const int z = z_split(ctx, maxSamplesPerLaunch);
    if (z > 1) CK(ctx->dPlanes.ensure((size_t)ctx->rp.width * ctx->rp.height * 16 * (size_t)z));
    if (ctx->rp.count_tests) CK(ctx->dCounters.ensure(N_COUNTERS * sizeof(unsigned long long)));
    return BRT_OK;
}

int launch_samples(brt_ctx* ctx, PTParams& p, float* dAccum, int sBegin, int sCount) {
    p.accum = (float4*)dAccum; p.sBegin = sBegin; p.sCount = sCount;
    const bool count = ctx->rp.count_tests != 0;
    if (count) {
        CK(ctx->dCounters.ensure(N_COUNTERS * sizeof(unsigned long long)));
        p.counters = (unsigned long long*)ctx->dCounters.p;
    }
    const int z = z_split(ctx, sCount);
    if (z > 1) {
        // small image: split the samples over z chunks to fill the GPU; every chunk accumulates into its own zeroed plane
        // and the planes are folded in fixed order (no atomics: results are bit-reproducible)
        const size_t px = (size_t)p.W * p.H;
        CK(ctx->dPlanes.ensure(px * 16 * (size_t)z));
        CK(cudaMemsetAsync(ctx->dPlanes.p, 0, px * 16 * (size_t)z, ctx->stream));
        p.accum = (float4*)ctx->dPlanes.p; p.planeStride = px;
        CK(launch_pathtrace(p, ctx->rp.sampler, use_bvh(ctx), count, z, ctx->stream));
        CK(launch_sum_planes((float4*)dAccum, (const float4*)ctx->dPlanes.p, z, px, ctx->stream));
        ctx->stats.launches += 2;
        return BRT_OK;
    }
    p.planeStride = 0;
    CK(launch_pathtrace(p, ctx->rp.sampler, use_bvh(ctx), count, 1, ctx->stream));
    ctx->stats.launches++;
    return BRT_OK;
}

PostParams post_params(const brt_ctx* ctx) {
    PostParams pp{};
    pp.W = ctx->rp.width; pp.H = ctx->rp.height; pp.tonemap = ctx->rp.tonemap; pp.exposure = ctx->rp.exposure;
    pp.invGamma = 1.0 / ctx->rp.gamma;                              // post-processor.js:36
    double s = ctx->rp.denoise_strength;
    pp.w1 = std::exp(-1.0 / (2 * s * s)); pp.w2 = std::exp(-2.0 / (2 * s * s));   // post-processor.js:59
    return pp;
}

}  // namespace brt

extern "C" {

int brt_render_accumulate(brt_ctx* ctx, float* d_accum, int sample_begin, int sample_count) {
    if (!ctx) return BRT_E_INVALID;
    if (sample_begin < 0 || sample_count < 0) return fail(ctx, BRT_E_INVALID, "negative sample range");
    // sample indices are Philox counter words and (wavefront) packed into 24 bits of the slot state
    if ((long long)sample_begin + sample_count > (1LL << 24)) return fail(ctx, BRT_E_INVALID, "sample_begin + sample_count must be <= 2^24");
    PTParams p;
    int rc = prepare(ctx, p);
    if (rc != BRT_OK) return rc;
    size_t px = (size_t)ctx->rp.width * ctx->rp.height;
    if (!d_accum) {
        bool fresh = ctx->dAccum.cap < px * 16;
        CK(ctx->dAccum.ensure(px * 16));
        if (fresh) CK(cudaMemsetAsync(ctx->dAccum.p, 0, px * 16, ctx->stream));
        d_accum = (float*)ctx->dAccum.p;
    }
    if (ctx->rp.count_tests) { CK(ctx->dCounters.ensure(N_COUNTERS * 8)); CK(cudaMemsetAsync(ctx->dCounters.p, 0, N_COUNTERS * 8, ctx->stream)); }
    ctx->stats.launches = 0;
    if (sample_count == 0) return BRT_OK;
    if ((rc = launch_samples(ctx, p, d_accum, sample_begin, sample_count)) != BRT_OK) return rc;
    ctx->stats.samples = (uint64_t)px * (uint64_t)sample_count;
    return BRT_OK;
}

int brt_resolve_device(brt_ctx* ctx, const float* d_accum, uint8_t* d_rgba8, float* d_float_data, float* d_linear_mean) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    size_t px = (size_t)ctx->rp.width * ctx->rp.height;
    if (!d_accum) d_accum = (const float*)ctx->dAccum.p;
    if (!d_accum) return fail(ctx, BRT_E_STATE, "no accumulation buffer");
    PostParams pp = post_params(ctx);
    if (ctx->rp.denoise) {
        float4* fd = (float4*)d_float_data;
        CK(ctx->dFloat2.ensure(px * 16));
        // denoise reads the un-filtered floatData and (ray-tracer.js:267-275) only the 8-bit image is replaced
        CK(launch_resolve(pp, (const float4*)d_accum, nullptr, (float4*)ctx->dFloat2.p, (float4*)d_linear_mean, 0, pp.H, ctx->stream));
        CK(launch_denoise(pp, (const float4*)ctx->dFloat2.p, (uchar4*)d_rgba8, nullptr, ctx->stream));
        if (fd) CK(cudaMemcpyAsync(fd, ctx->dFloat2.p, px * 16, cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->stats.launches += 2;
    } else {
        CK(launch_resolve(pp, (const float4*)d_accum, (uchar4*)d_rgba8, (float4*)d_float_data, (float4*)d_linear_mean, 0, pp.H, ctx->stream));
        ctx->stats.launches += 1;
    }
    return BRT_OK;
}

int brt_reduce_resolve_peers(brt_ctx* ctx, const float* const* d_peer_accum, int n_peers, int row_begin, int row_end,
                             uint8_t* d_rgba8_root, float* d_float_data_root) {
    if (!ctx || !d_peer_accum) return BRT_E_INVALID;
    if (n_peers < 1 || n_peers > 16) return fail(ctx, BRT_E_INVALID, "n_peers must be in [1,16]");
    NEED_GPU();
    if (row_begin < 0 || row_end > ctx->rp.height || row_begin > row_end) return fail(ctx, BRT_E_INVALID, "bad row range");
    CK(cudaSetDevice(ctx->device));
    PostParams pp = post_params(ctx);
    CK(launch_reduce_resolve(pp, (const float4* const*)d_peer_accum, n_peers, (uchar4*)d_rgba8_root, (float4*)d_float_data_root,
                             row_begin, row_end, ctx->stream));
    ctx->stats.launches += 1;
    return BRT_OK;
}

int brt_shared_alloc(brt_ctx* ctx, size_t bytes, void** d_ptr, uint8_t handle[64]) {
    if (!ctx || !d_ptr || !handle || bytes == 0) return BRT_E_INVALID;
    NEED_GPU();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    CK(cudaSetDevice(ctx->device));
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    *d_ptr = p;
    return BRT_OK;
}
int brt_shared_free(brt_ctx* ctx, void* d_ptr) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(d_ptr));
    return BRT_OK;
}
int brt_shared_open(brt_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
    if (!ctx || !handle || !d_ptr) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BRT_OK;
}
int brt_shared_close(brt_ctx* ctx, void* d_ptr) {
    if (!ctx) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return BRT_OK;
}
int brt_device_memset(brt_ctx* ctx, void* d_ptr, int value, size_t bytes) {
    if (!ctx || !d_ptr) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(d_ptr, value, bytes, ctx->stream));
    return BRT_OK;
}

int brt_copy_to_host(brt_ctx* ctx, void* host_dst, const void* d_src, size_t bytes) {
    if (!ctx || !host_dst || !d_src) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_render(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean, brt_progress_cb cb, void* user) {
    if (!ctx || !rgba8) return BRT_E_INVALID;
    if (!ctx->followers.empty()) return render_multi(ctx, rgba8, float_data, linear_mean, cb, user);   // brt_create_multi: n GPUs behind the same call
    auto w0 = std::chrono::steady_clock::now();
    ctx->cancel.store(0);
    PTParams p;
    int rc = prepare(ctx, p);
    if (rc != BRT_OK) return rc;
    const size_t px = (size_t)ctx->rp.width * ctx->rp.height;
    CK(ctx->dAccum.ensure(px * 16)); CK(ctx->dRgba.ensure(px * 4));
    if (float_data || ctx->rp.denoise) CK(ctx->dFloat.ensure(px * 16));
    if (linear_mean) CK(ctx->dLinear.ensure(px * 16));
    CK(cudaMemsetAsync(ctx->dAccum.p, 0, px * 16, ctx->stream));
    if (ctx->rp.count_tests) { CK(ctx->dCounters.ensure(N_COUNTERS * 8)); CK(cudaMemsetAsync(ctx->dCounters.p, 0, N_COUNTERS * 8, ctx->stream)); }
    ctx->stats = brt_stats{};
    const int spp = effective_spp(ctx->rp);
    const int batch = spp_batch(ctx, spp, cb != nullptr);
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int s = 0; s < spp; s += batch) {
        int n = spp - s < batch ? spp - s : batch;
        if ((rc = launch_samples(ctx, p, (float*)ctx->dAccum.p, s, n)) != BRT_OK) return rc;
        if (cb || s + n < spp) {
            // progress + cooperative cancel between batches (ray-tracer.js:190,256-261)
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->cancel.load()) return fail(ctx, BRT_E_CANCELLED, "render cancelled");
            if (cb && s + n < spp) {
                if (ctx->rp.preview) {
                    // progressive preview: the sums so far divided by their own sample count (alpha) are a complete image
                    bool dn = ctx->rp.denoise != 0;
                    if (dn) CK(ctx->dFloat.ensure(px * 16));
                    rc = brt_resolve_device(ctx, (const float*)ctx->dAccum.p, (uint8_t*)ctx->dRgba.p, dn ? (float*)ctx->dFloat.p : nullptr, nullptr);
                    if (rc != BRT_OK) return rc;
                    CK(cudaMemcpyAsync(rgba8, ctx->dRgba.p, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
                    CK(cudaStreamSynchronize(ctx->stream));
                }
                cb((double)(s + n) / spp, user);
            }
        }
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    rc = brt_resolve_device(ctx, (const float*)ctx->dAccum.p, (uint8_t*)ctx->dRgba.p, (float_data || ctx->rp.denoise) ? (float*)ctx->dFloat.p : nullptr,
                            linear_mean ? (float*)ctx->dLinear.p : nullptr);
    if (rc != BRT_OK) return rc;
    CK(cudaEventRecord(ctx->ev2, ctx->stream));
    CK(cudaMemcpyAsync(rgba8, ctx->dRgba.p, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (float_data) CK(cudaMemcpyAsync(float_data, ctx->dFloat.p, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    if (linear_mean) CK(cudaMemcpyAsync(linear_mean, ctx->dLinear.p, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->cancel.load()) return fail(ctx, BRT_E_CANCELLED, "render cancelled");
    float k = 0, q = 0;
    cudaEventElapsedTime(&k, ctx->ev0, ctx->ev1); cudaEventElapsedTime(&q, ctx->ev1, ctx->ev2);
    ctx->stats.kernel_ms = k; ctx->stats.post_ms = q;
    ctx->stats.samples = (uint64_t)px * (uint64_t)spp;
    ctx->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
    if (cb) cb(1.0, user);                                          // ray-tracer.js:279
    return BRT_OK;
}

void brt_cancel(brt_ctx* ctx) { if (ctx) ctx->cancel.store(1); }

int brt_get_stats(brt_ctx* ctx, brt_stats* out) {
    if (!ctx || !out) return BRT_E_INVALID;
    if (ctx->device >= 0 && ctx->rp.count_tests && ctx->dCounters.p) {
        CK(cudaSetDevice(ctx->device));
        unsigned long long c[N_COUNTERS];
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaMemcpy(c, ctx->dCounters.p, sizeof(c), cudaMemcpyDeviceToHost));
        ctx->stats.rays = c[0]; ctx->stats.tests_sphere = c[1]; ctx->stats.tests_plane = c[2]; ctx->stats.tests_box = c[3];
        ctx->stats.tests_tri_a = c[4]; ctx->stats.tests_tri_b = c[5]; ctx->stats.tests_tri_c = c[6]; ctx->stats.tests_aabb = c[7];
        ctx->stats.trav_warp_iters = c[8]; ctx->stats.trav_lane_iters = c[9]; ctx->stats.trav_alive_lanes = c[10];
        ctx->stats.trav_node_issues = c[11]; ctx->stats.trav_leaf_issues = c[12]; ctx->stats.trav_leaf_lanes = c[13];
        ctx->stats.path_warp_iters = c[14]; ctx->stats.path_lane_iters = c[15]; ctx->stats.node_visits = c[16];
    }
    *out = ctx->stats;
    return BRT_OK;
}

// ------------------------------------------------------------------------------------------- parity AOVs
int brt_primary_aov_f32(brt_ctx* ctx, int32_t* obj_id, int32_t* tri_id, float* t, float* normal3, uint8_t* front_face) {
    if (!ctx || !obj_id || !tri_id || !t || !normal3 || !front_face) return BRT_E_INVALID;
    PTParams p;
    int rc = prepare(ctx, p);
    if (rc != BRT_OK) return rc;
    if ((rc = ensure_prim64(ctx)) != BRT_OK) return rc;
    p.sc = ctx->dev;
    size_t px = (size_t)p.W * p.H;
    CK(ctx->dScratch.ensure(px * (4 + 4 + 4 + 12 + 1) + 64));
    char* base = (char*)ctx->dScratch.p;
    int* dObj = (int*)base; int* dTri = (int*)(base + px * 4); float* dT = (float*)(base + px * 8); float* dN = (float*)(base + px * 12);
    unsigned char* dF = (unsigned char*)(base + px * 24);
    CK(launch_primary_aov(p, use_bvh(ctx), dObj, dTri, dT, dN, dF, ctx->stream));
    CK(cudaMemcpyAsync(obj_id, dObj, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(tri_id, dTri, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(t, dT, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(normal3, dN, px * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(front_face, dF, px, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_primary_aov_f64(brt_ctx* ctx, int32_t* obj_id, int32_t* tri_id, double* t, double* normal3, uint8_t* front_face) {
    if (!ctx || !obj_id || !tri_id || !t || !normal3 || !front_face) return BRT_E_INVALID;
    NEED_GPU();
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    if (!ctx->haveCam) return fail(ctx, BRT_E_NOSCENE, "no camera set");
    CK(cudaSetDevice(ctx->device));
    const HostScene& s = ctx->scene;
    if (ctx->obj64Dirty) {
        std::vector<Obj64> objs(s.objects.size());
        for (size_t i = 0; i < s.objects.size(); i++) {
            const brt_object& o = s.objects[i]; Obj64& d = objs[i];
            d.type = o.type; d.material = o.material; d.firstTri = o.first_tri; d.triCount = o.tri_count;
            for (int k = 0; k < 3; k++) { d.a[k] = o.a[k]; d.b[k] = o.b[k]; d.c[k] = o.c[k]; }
        }
        CK(ctx->dObj64.ensure(objs.size() * sizeof(Obj64)));
        if (!objs.empty()) CK(cudaMemcpy(ctx->dObj64.p, objs.data(), objs.size() * sizeof(Obj64), cudaMemcpyHostToDevice));
        CK(ctx->dTris64.ensure(s.meshTris.size() * sizeof(double)));
        if (!s.meshTris.empty()) CK(cudaMemcpy(ctx->dTris64.p, s.meshTris.data(), s.meshTris.size() * sizeof(double), cudaMemcpyHostToDevice));
        ctx->obj64Dirty = false;
    }
    Cam64 c{};
    for (int k = 0; k < 3; k++) {
        c.origin[k] = ctx->cam.origin[k]; c.llc[k] = ctx->cam.lower_left_corner[k]; c.horizontal[k] = ctx->cam.horizontal[k];
        c.vertical[k] = ctx->cam.vertical[k]; c.w[k] = ctx->cam.w[k];
    }
    c.type = ctx->cam.type;
    const int W = ctx->rp.width, H = ctx->rp.height;
    size_t px = (size_t)W * H;
    CK(ctx->dScratch.ensure(px * (4 + 4 + 8 + 24 + 1) + 64));
    char* base = (char*)ctx->dScratch.p;
    double* dT = (double*)base; double* dN = (double*)(base + px * 8); int* dObj = (int*)(base + px * 32); int* dTri = (int*)(base + px * 36);
    unsigned char* dF = (unsigned char*)(base + px * 40);
    CK(launch_primary_aov64((const Obj64*)ctx->dObj64.p, (int)s.objects.size(), (const double*)ctx->dTris64.p, c, W, H, dObj, dTri, dT, dN, dF, ctx->stream));
    CK(cudaMemcpyAsync(obj_id, dObj, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(tri_id, dTri, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(t, dT, px * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(normal3, dN, px * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(front_face, dF, px, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

// ------------------------------------------------------------------------------------------- unit hooks
int brt_eval_background(brt_ctx* ctx, const double* dirs, int n, float* out_rgb) {
    if (!ctx || !dirs || !out_rgb || n < 0) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    int rc = upload_perm(ctx);
    if (rc != BRT_OK) return rc;
    if (n == 0) return BRT_OK;
    std::vector<float> f(3 * (size_t)n);
    for (size_t i = 0; i < f.size(); i++) f[i] = (float)dirs[i];
    CK(ctx->dScratch.ensure(f.size() * 8));
    float* dIn = (float*)ctx->dScratch.p; float* dOut = dIn + f.size();
    CK(cudaMemcpyAsync(dIn, f.data(), f.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    DevScene sc = ctx->dev;
    sc.bgKind = ctx->bg.kind; sc.bgR = (float)ctx->bg.color[0]; sc.bgG = (float)ctx->bg.color[1]; sc.bgB = (float)ctx->bg.color[2];
    sc.skyIntensity = (float)ctx->bg.intensity;
    CK(launch_eval_background(sc, dIn, n, dOut, ctx->stream));
    CK(cudaMemcpyAsync(out_rgb, dOut, f.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_eval_texture(brt_ctx* ctx, int tex_index, const double* points, int n, float* out_rgb) {
    if (!ctx || !points || !out_rgb || n < 0) return BRT_E_INVALID;
    NEED_GPU();
    if (!ctx->haveScene) return fail(ctx, BRT_E_NOSCENE, "no scene loaded");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_scene(ctx);
    if (rc != BRT_OK) return rc;
    if (tex_index < 0 || tex_index >= ctx->dev.nTex) return fail(ctx, BRT_E_INVALID, "texture index out of range");
    if (n == 0) return BRT_OK;
    std::vector<float> f(3 * (size_t)n);
    for (size_t i = 0; i < f.size(); i++) f[i] = (float)points[i];
    CK(ctx->dScratch.ensure(f.size() * 8));
    float* dIn = (float*)ctx->dScratch.p; float* dOut = dIn + f.size();
    CK(cudaMemcpyAsync(dIn, f.data(), f.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_eval_texture(ctx->dev, tex_index, dIn, n, dOut, ctx->stream));
    CK(cudaMemcpyAsync(out_rgb, dOut, f.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_debug_rng_stream(brt_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, int n, float* out) {
    if (!ctx || !out || n < 0) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return BRT_OK;
    CK(ctx->dScratch.ensure((size_t)n * 4));
    CK(launch_rng_stream((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, n, (float*)ctx->dScratch.p, ctx->stream));
    CK(cudaMemcpyAsync(out, ctx->dScratch.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_postprocess_host(brt_ctx* ctx, const float* linear_mean, uint8_t* rgba8, float* float_data) {
    if (!ctx || !linear_mean || !rgba8) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    const size_t px = (size_t)ctx->rp.width * ctx->rp.height;
    // a linear-mean image is an accumulation buffer with one sample per pixel
    std::vector<float> acc(linear_mean, linear_mean + px * 4);
    for (size_t i = 0; i < px; i++) acc[4 * i + 3] = 1.0f;
    CK(ctx->dAccum.ensure(px * 16)); CK(ctx->dRgba.ensure(px * 4)); CK(ctx->dFloat.ensure(px * 16));
    CK(cudaMemcpyAsync(ctx->dAccum.p, acc.data(), px * 16, cudaMemcpyHostToDevice, ctx->stream));
    int rc = brt_resolve_device(ctx, (const float*)ctx->dAccum.p, (uint8_t*)ctx->dRgba.p, (float*)ctx->dFloat.p, nullptr);
    if (rc != BRT_OK) return rc;
    CK(cudaMemcpyAsync(rgba8, ctx->dRgba.p, px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (float_data) CK(cudaMemcpyAsync(float_data, ctx->dFloat.p, px * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return BRT_OK;
}

int brt_measure_fp32_peak(brt_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return BRT_E_INVALID;
    NEED_GPU();
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, ctx->device));
    CK(ctx->dScratch.ensure(64));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    CK(launch_fp32_peak((float*)ctx->dScratch.p, blocks, 64, ctx->stream));           // warm-up
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(ctx->ev0, ctx->stream));
        CK(launch_fp32_peak((float*)ctx->dScratch.p, blocks, iters, ctx->stream));
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        double flops = (double)blocks * 256 * (double)iters * 16 * 8 * 2;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    *tflops = best;
    return BRT_OK;
}

}  // extern "C"
