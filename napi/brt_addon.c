/* brt_addon.c — thin Node-API (N-API) binding of libbrt (include/brt.h).
 *
 * Replaces the per-pixel loop of the reference's RayTracer.render() (js/ray-tracer.js:166-281): the JS shim
 * (napi/raytracer_gpu.mjs) flattens the live World / Camera / settings into typed arrays, calls into this addon, and the
 * GPU writes the RGBA8 image straight into the Uint8ClampedArray behind `imageData.data` (zero copy on the JS side).
 * render() runs brt_render on a libuv worker thread (napi_async_work) so the event loop stays responsive — the role of
 * the reference's `await new Promise(r => setTimeout(r, 0))` yields (:241,:261) — and reports progress through a
 * thread-safe function; cancel() maps window.renderCancelled (:190,:256,:264) onto brt_cancel.
 *
 * Build (with Node headers; this repo's image has none, so __graft_entry__.build() compiles against node_api_min.h — the same
 * C ABI — and tests/test_napi_mock.py loads the result into a mock N-API host and runs it end to end):
 *   cc -shared -fPIC -I../include -I$(node -p "process.execPath+'/../../include/node'") brt_addon.c \
 *      -L../blenderraytracer_b200 -lbrt -Wl,-rpath,'$ORIGIN/../blenderraytracer_b200' -o brt_addon.node
 */
#if defined(__has_include)
#  if __has_include(<node_api.h>)
#    include <node_api.h>
#  else
#    include "node_api_min.h"
#  endif
#else
#  include "node_api_min.h"
#endif
#include <stdlib.h>
#include <string.h>
#include "brt.h"

#define ARGS(n) size_t argc = (n); napi_value argv[(n)]; napi_get_cb_info(env, info, &argc, argv, NULL, NULL)

static napi_value throw_brt(napi_env env, brt_ctx* ctx, int rc) {
    char code[16];
    const char* names[] = { "BRT_OK", "BRT_E_INVALID", "BRT_E_CUDA", "BRT_E_PARSE", "BRT_E_NOSCENE", "BRT_E_CANCELLED", "BRT_E_NOMEM", "BRT_E_STATE" };
    strncpy(code, (-rc >= 0 && -rc < 8) ? names[-rc] : "BRT_E_?", sizeof(code) - 1); code[sizeof(code) - 1] = 0;
    napi_throw_error(env, code, ctx ? brt_last_error(ctx) : "libbrt call failed");
    return NULL;
}
static brt_ctx* get_ctx(napi_env env, napi_value v) { void* p = NULL; napi_get_value_external(env, v, &p); return (brt_ctx*)p; }
static void finalize_ctx(napi_env env, void* data, void* hint) { (void)env; (void)hint; brt_destroy((brt_ctx*)data); }
static double* f64_array(napi_env env, napi_value v, size_t* n) {
    napi_typedarray_type ty; void* data = NULL; size_t len = 0;
    if (napi_get_typedarray_info(env, v, &ty, &len, &data, NULL, NULL) != napi_ok || ty != napi_float64_array) { *n = 0; return NULL; }
    *n = len; return (double*)data;
}
static int32_t get_i32(napi_env env, napi_value obj, const char* key, int32_t dflt) {
    bool has = false; napi_value v; int32_t r = dflt;
    if (napi_has_named_property(env, obj, key, &has) == napi_ok && has && napi_get_named_property(env, obj, key, &v) == napi_ok) {
        napi_valuetype t; napi_typeof(env, v, &t);
        if (t == napi_number) napi_get_value_int32(env, v, &r);
        else if (t == napi_boolean) { bool b; napi_get_value_bool(env, v, &b); r = b ? 1 : 0; }
    }
    return r;
}
static double get_f64(napi_env env, napi_value obj, const char* key, double dflt) {
    bool has = false; napi_value v; double r = dflt;
    if (napi_has_named_property(env, obj, key, &has) == napi_ok && has && napi_get_named_property(env, obj, key, &v) == napi_ok) {
        napi_valuetype t; napi_typeof(env, v, &t);
        if (t == napi_number) napi_get_value_double(env, v, &r);
    }
    return r;
}

/* create(deviceId | [deviceId, ...]) -> external.  An array makes ONE context that spans several GPUs (brt_create_multi):
 * render() stays one call, the samples of every batch are split over the devices and exchanged by the fused peer kernel. */
static napi_value js_create(napi_env env, napi_callback_info info) {
    ARGS(1);
    brt_ctx* ctx = NULL;
    int rc;
    bool is_arr = false;
    if (argc >= 1 && napi_is_array(env, argv[0], &is_arr) == napi_ok && is_arr) {
        uint32_t n = 0; napi_get_array_length(env, argv[0], &n);
        int ids[16];
        if (n < 1 || n > 16) { napi_throw_error(env, "BRT_E_INVALID", "devices: 1..16 CUDA device ids"); return NULL; }
        for (uint32_t i = 0; i < n; i++) { napi_value e; int32_t d = 0; napi_get_element(env, argv[0], i, &e); napi_get_value_int32(env, e, &d); ids[i] = d; }
        rc = brt_create_multi(&ctx, ids, (int)n);
    } else {
        int32_t dev = 0; if (argc >= 1) napi_get_value_int32(env, argv[0], &dev);
        rc = brt_create(&ctx, dev);
    }
    if (rc != BRT_OK) { napi_throw_error(env, rc == BRT_E_INVALID ? "BRT_E_INVALID" : "BRT_E_CUDA", "brt_create failed: libbrt needs CUDA device(s) it can use (no CPU fallback)"); return NULL; }
    napi_value ext; napi_create_external(env, ctx, finalize_ctx, NULL, &ext);
    return ext;
}

/* loadSceneJSON(ctx, text, fallbackW, fallbackH) -> { hasCamera, width, height }   (RayTracer.loadFromJSON, ray-tracer.js:305-334) */
static napi_value js_load_scene_json(napi_env env, napi_callback_info info) {
    ARGS(4);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    size_t len = 0; napi_get_value_string_utf8(env, argv[1], NULL, 0, &len);
    char* buf = (char*)malloc(len + 1);
    if (!buf) return throw_brt(env, ctx, BRT_E_NOMEM);
    napi_get_value_string_utf8(env, argv[1], buf, len + 1, &len);
    int32_t fw = 600, fh = 400; napi_get_value_int32(env, argv[2], &fw); napi_get_value_int32(env, argv[3], &fh);
    int has = 0, w = 0, h = 0;
    int rc = brt_scene_load_json(ctx, buf, len, fw, fh, &has, &w, &h);
    free(buf);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value out, v; napi_create_object(env, &out);
    napi_get_boolean(env, has != 0, &v); napi_set_named_property(env, out, "hasCamera", v);
    napi_create_int32(env, w, &v); napi_set_named_property(env, out, "width", v);
    napi_create_int32(env, h, &v); napi_set_named_property(env, out, "height", v);
    return out;
}

/* setSceneFlat(ctx, objects Float64Array[13 n], materials Float64Array[6 m], meshTris Float64Array[9 t], lights Float64Array[8 l])
 * objects: type, material, a.xyz, b.xyz, c.xyz, firstTri, triCount — one row per world.objects entry, in order (world.js:24-30)
 * materials: type, r, g, b, param, texture(1-based, 0 = none) — 6 per material
 * optional: textures Float64Array[8 k] (kind, odd.rgb, even.rgb, scale) + texturePerms Uint8Array[256 k]   (js/textures.js);
 * flags (number): BRT_SCENE_CONSTRUCTED = 1 when the rows were read from constructed objects (normals already normalised) */
static napi_value js_set_scene_flat(napi_env env, napi_callback_info info) {
    ARGS(8);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    size_t no, nm, nt, nl;
    double* o = f64_array(env, argv[1], &no); double* m = f64_array(env, argv[2], &nm);
    double* t = f64_array(env, argv[3], &nt); double* l = f64_array(env, argv[4], &nl);
    no /= 13; nm /= 6; nt /= 9; nl /= 8;
    size_t nx = 0; double* x = argc >= 6 ? f64_array(env, argv[5], &nx) : NULL; nx /= 8;
    const uint8_t* xp = NULL;
    if (argc >= 7) { napi_typedarray_type ty; void* data = NULL; size_t len = 0;
        if (napi_get_typedarray_info(env, argv[6], &ty, &len, &data, NULL, NULL) == napi_ok && ty == napi_uint8_array && len >= 256 * nx) xp = (const uint8_t*)data; }
    brt_object* objs = (brt_object*)calloc(no ? no : 1, sizeof(brt_object));
    brt_material* mats = (brt_material*)calloc(nm ? nm : 1, sizeof(brt_material));
    brt_light* lights = (brt_light*)calloc(nl ? nl : 1, sizeof(brt_light));
    if (!objs || !mats || !lights) { free(objs); free(mats); free(lights); return throw_brt(env, ctx, BRT_E_NOMEM); }
    for (size_t i = 0; i < no; i++) {
        const double* r = o + 13 * i;
        objs[i].type = (int32_t)r[0]; objs[i].material = (int32_t)r[1];
        memcpy(objs[i].a, r + 2, 24); memcpy(objs[i].b, r + 5, 24); memcpy(objs[i].c, r + 8, 24);
        objs[i].first_tri = (int64_t)r[11]; objs[i].tri_count = (int64_t)r[12];
    }
    for (size_t i = 0; i < nm; i++) { const double* r = m + 6 * i; mats[i].type = (int32_t)r[0]; memcpy(mats[i].color, r + 1, 24); mats[i].param = r[4]; mats[i].texture = (int32_t)r[5]; }
    brt_texture* texs = (brt_texture*)calloc(nx ? nx : 1, sizeof(brt_texture));
    if (!texs) { free(objs); free(mats); free(lights); return throw_brt(env, ctx, BRT_E_NOMEM); }
    for (size_t i = 0; i < nx; i++) {
        const double* r = x + 8 * i; texs[i].kind = (int32_t)r[0]; memcpy(texs[i].odd, r + 1, 24); memcpy(texs[i].even, r + 4, 24); texs[i].scale = r[7];
        for (int k = 0; k < 256; k++) texs[i].perm[k] = xp ? xp[256 * i + k] : (uint8_t)k;
    }
    for (size_t i = 0; i < nl; i++) { const double* r = l + 8 * i; lights[i].type = (int32_t)r[0]; memcpy(lights[i].v, r + 1, 24); memcpy(lights[i].color, r + 4, 24); lights[i].intensity = r[7]; }
    brt_scene_desc d; memset(&d, 0, sizeof(d));
    d.objects = objs; d.n_objects = (int32_t)no; d.materials = mats; d.n_materials = (int32_t)nm;
    d.mesh_triangles = t; d.n_mesh_triangles = (int64_t)nt; d.lights = lights; d.n_lights = (int32_t)nl;
    d.textures = texs; d.n_textures = (int32_t)nx;
    if (argc >= 8) { int32_t fl = 0; if (napi_get_value_int32(env, argv[7], &fl) == napi_ok) d.flags = fl; }   /* 1 = rows read from constructed objects */
    int rc = brt_scene_set_flat(ctx, &d);                       /* borrowed, copied before return */
    free(objs); free(mats); free(lights); free(texs);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value u; napi_get_undefined(env, &u); return u;
}

/* setCameraDerived(ctx, Float64Array[23]): origin, lowerLeftCorner, horizontal, vertical, u, v, w, lensRadius, type —
 * the reference Camera object's own derived members (camera.js:14-35), so not even Math.tan can differ. */
static napi_value js_set_camera_derived(napi_env env, napi_callback_info info) {
    ARGS(2);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    size_t n; double* c = f64_array(env, argv[1], &n);
    if (!c || n < 23) return throw_brt(env, ctx, BRT_E_INVALID);
    brt_camera cam; memset(&cam, 0, sizeof(cam));
    memcpy(cam.origin, c, 24); memcpy(cam.lower_left_corner, c + 3, 24); memcpy(cam.horizontal, c + 6, 24); memcpy(cam.vertical, c + 9, 24);
    memcpy(cam.u, c + 12, 24); memcpy(cam.v, c + 15, 24); memcpy(cam.w, c + 18, 24);
    cam.lens_radius = c[21]; cam.type = (int32_t)c[22]; cam.use_derived = 1;
    int rc = brt_set_camera(ctx, &cam);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value u; napi_get_undefined(env, &u); return u;
}

/* setBackground(ctx, kind, r, g, b, intensity, perm Uint8Array[256] | undefined)   (world.js:12-14, ray-tracer.js:568-585) */
static napi_value js_set_background(napi_env env, napi_callback_info info) {
    ARGS(7);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    int32_t kind = 0; double col[3] = { 0.1, 0.1, 0.1 }, inten = 1.0;
    napi_get_value_int32(env, argv[1], &kind);
    napi_get_value_double(env, argv[2], &col[0]); napi_get_value_double(env, argv[3], &col[1]); napi_get_value_double(env, argv[4], &col[2]);
    napi_get_value_double(env, argv[5], &inten);
    const uint8_t* perm = NULL;
    if (argc >= 7) {
        napi_typedarray_type ty; void* data = NULL; size_t len = 0;
        if (napi_get_typedarray_info(env, argv[6], &ty, &len, &data, NULL, NULL) == napi_ok && ty == napi_uint8_array && len >= 256) perm = (const uint8_t*)data;
    }
    int rc = brt_set_background(ctx, kind, col, inten, perm);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value u; napi_get_undefined(env, &u); return u;
}

/* setRenderParams(ctx, { width, height, samples, maxBounces, aaMode, toneMapping, exposure, gamma, denoising, denoiseStrength, seed, ... }) */
static napi_value js_set_render_params(napi_env env, napi_callback_info info) {
    ARGS(2);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    brt_render_params p; brt_get_render_params(ctx, &p);
    napi_value o = argv[1];
    p.width = get_i32(env, o, "width", p.width); p.height = get_i32(env, o, "height", p.height);
    p.spp = get_i32(env, o, "samples", p.spp); p.max_depth = get_i32(env, o, "maxBounces", p.max_depth);
    p.aa_mode = get_i32(env, o, "aaMode", p.aa_mode); p.tonemap = get_i32(env, o, "toneMapping", p.tonemap);
    p.exposure = get_f64(env, o, "exposure", p.exposure); p.gamma = get_f64(env, o, "gamma", p.gamma);
    p.denoise = get_i32(env, o, "denoising", p.denoise); p.denoise_strength = get_f64(env, o, "denoiseStrength", p.denoise_strength);
    p.seed = (uint64_t)get_f64(env, o, "seed", (double)p.seed);
    p.direct_lighting = get_i32(env, o, "directLighting", p.direct_lighting);
    p.sampler = get_i32(env, o, "sampler", p.sampler); p.accel = get_i32(env, o, "accel", p.accel);
    p.integrator = get_i32(env, o, "integrator", p.integrator); p.spp_batch = get_i32(env, o, "sppBatch", p.spp_batch);
    p.preview = get_i32(env, o, "preview", p.preview); p.bvh_width = get_i32(env, o, "bvhWidth", p.bvh_width);
    int rc = brt_set_render_params(ctx, &p);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value u; napi_get_undefined(env, &u); return u;
}

/* render(ctx, data Uint8ClampedArray[W*H*4], onProgress?) -> Promise<void>   (RayTracer.render, ray-tracer.js:166-281) */
typedef struct {
    brt_ctx* ctx; uint8_t* rgba; int rc; char err[256];
    napi_deferred deferred; napi_async_work work; napi_ref data_ref, ctx_ref; napi_threadsafe_function tsfn;
} render_job;
static void progress_from_worker(double fraction, void* user) {
    render_job* j = (render_job*)user;
    if (!j->tsfn) return;
    double* f = (double*)malloc(sizeof(double));
    if (!f) return;
    *f = fraction;
    if (napi_call_threadsafe_function(j->tsfn, f, napi_tsfn_nonblocking) != napi_ok) free(f);
}
static void progress_call_js(napi_env env, napi_value cb, void* context, void* data) {
    (void)context;
    if (env && cb && data) {
        napi_value arg, undef; napi_create_double(env, *(double*)data, &arg); napi_get_undefined(env, &undef);
        napi_call_function(env, undef, cb, 1, &arg, NULL);          /* onProgress(fraction) (ray-tracer.js:258-259,279) */
    }
    free(data);
}
static void render_execute(napi_env env, void* data) {
    (void)env;
    render_job* j = (render_job*)data;
    j->rc = brt_render(j->ctx, j->rgba, NULL, NULL, progress_from_worker, j);
    if (j->rc != BRT_OK) { strncpy(j->err, brt_last_error(j->ctx), sizeof(j->err) - 1); j->err[sizeof(j->err) - 1] = 0; }
}
static void render_complete(napi_env env, napi_status status, void* data) {
    render_job* j = (render_job*)data;
    napi_value v;
    if (status == napi_ok && j->rc == BRT_OK) { napi_get_undefined(env, &v); napi_resolve_deferred(env, j->deferred, v); }
    else {
        napi_value msg, code;
        napi_create_string_utf8(env, j->rc == BRT_E_CANCELLED ? "render cancelled" : j->err, NAPI_AUTO_LENGTH, &msg);
        napi_create_string_utf8(env, j->rc == BRT_E_CANCELLED ? "BRT_E_CANCELLED" : "BRT_E", NAPI_AUTO_LENGTH, &code);
        napi_create_error(env, code, msg, &v);
        napi_reject_deferred(env, j->deferred, v);
    }
    if (j->tsfn) napi_release_threadsafe_function(j->tsfn, napi_tsfn_release);
    napi_delete_reference(env, j->data_ref);
    napi_delete_reference(env, j->ctx_ref);
    napi_delete_async_work(env, j->work);
    free(j);
}
static napi_value js_render(napi_env env, napi_callback_info info) {
    ARGS(3);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    napi_typedarray_type ty; void* data = NULL; size_t len = 0;
    if (napi_get_typedarray_info(env, argv[1], &ty, &len, &data, NULL, NULL) != napi_ok || (ty != napi_uint8_clamped_array && ty != napi_uint8_array))
        return throw_brt(env, ctx, BRT_E_INVALID);
    brt_render_params p; brt_get_render_params(ctx, &p);
    if (len < (size_t)p.width * (size_t)p.height * 4) { napi_throw_error(env, "BRT_E_INVALID", "imageData.data is smaller than width*height*4"); return NULL; }
    render_job* j = (render_job*)calloc(1, sizeof(render_job));
    if (!j) return throw_brt(env, ctx, BRT_E_NOMEM);
    j->ctx = ctx; j->rgba = (uint8_t*)data;
    napi_value promise, name;
    napi_create_promise(env, &j->deferred, &promise);
    napi_create_reference(env, argv[1], 1, &j->data_ref);          /* keep the pixel buffer alive while the worker writes it */
    napi_create_reference(env, argv[0], 1, &j->ctx_ref);           /* and the ctx: dropping the handle mid-render must not run finalize_ctx -> brt_destroy */
    napi_create_string_utf8(env, "brt_render", NAPI_AUTO_LENGTH, &name);
    napi_valuetype t = napi_undefined;
    if (argc >= 3) napi_typeof(env, argv[2], &t);
    if (t == napi_function) napi_create_threadsafe_function(env, argv[2], NULL, name, 0, 1, NULL, NULL, NULL, progress_call_js, &j->tsfn);
    napi_create_async_work(env, NULL, name, render_execute, render_complete, j, &j->work);
    napi_queue_async_work(env, j->work);
    return promise;
}

/* cancel(ctx): window.renderCancelled = true (ui-controller.js:134-137); the only call allowed while render() is in flight */
static napi_value js_cancel(napi_env env, napi_callback_info info) {
    ARGS(1);
    brt_cancel(get_ctx(env, argv[0]));
    napi_value u; napi_get_undefined(env, &u); return u;
}

/* stats(ctx) -> { samples, kernelMs, postMs, totalMs, launches } */
static napi_value js_stats(napi_env env, napi_callback_info info) {
    ARGS(1);
    brt_ctx* ctx = get_ctx(env, argv[0]);
    brt_stats s; int rc = brt_get_stats(ctx, &s);
    if (rc != BRT_OK) return throw_brt(env, ctx, rc);
    napi_value out, v; napi_create_object(env, &out);
    napi_create_double(env, (double)s.samples, &v); napi_set_named_property(env, out, "samples", v);
    napi_create_double(env, s.kernel_ms, &v); napi_set_named_property(env, out, "kernelMs", v);
    napi_create_double(env, s.post_ms, &v); napi_set_named_property(env, out, "postMs", v);
    napi_create_double(env, s.total_ms, &v); napi_set_named_property(env, out, "totalMs", v);
    napi_create_double(env, (double)s.launches, &v); napi_set_named_property(env, out, "launches", v);
    napi_create_int32(env, brt_device_count(ctx), &v); napi_set_named_property(env, out, "devices", v);
    return out;
}

napi_value napi_register_module_v1(napi_env env, napi_value exports) {
    struct { const char* name; napi_callback fn; } fns[] = {
        { "create", js_create }, { "loadSceneJSON", js_load_scene_json }, { "setSceneFlat", js_set_scene_flat },
        { "setCameraDerived", js_set_camera_derived }, { "setBackground", js_set_background }, { "setRenderParams", js_set_render_params },
        { "render", js_render }, { "cancel", js_cancel }, { "stats", js_stats },
    };
    for (size_t i = 0; i < sizeof(fns) / sizeof(fns[0]); i++) {
        napi_value f;
        napi_create_function(env, fns[i].name, NAPI_AUTO_LENGTH, fns[i].fn, NULL, &f);
        napi_set_named_property(env, exports, fns[i].name, f);
    }
    return exports;
}
