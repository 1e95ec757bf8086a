/* brt.h — C ABI of libbrt, the B200-native (sm_100a) path-tracing engine that replaces the per-pixel
 * render loop of Shinzef/BlenderRayTracer ("RayCast").
 *
 * The reference has no FFI/plugin interface of its own (it is browser JavaScript); the seam this library
 * replaces is the body of `RayTracer.render(onProgress)` — reference js/ray-tracer.js:166-281 — and the
 * scene objects it reads (js/world.js, js/camera.js, js/geometry.js, js/materials.js, js/lights.js,
 * js/noise.js, js/post-processor.js, js/scene-loader.js).  Each entry point below cites the reference
 * code whose behaviour it reproduces.  INTEGRATION.md shows the Node N-API / JS binding (napi/) and the
 * Python ctypes binding (blenderraytracer_b200/) that sit on top of it.
 *
 * Conventions
 *   - C linkage, plain pointers and sizes only; no exceptions cross the boundary.
 *   - Every call returns BRT_OK (0) or a negative BRT_E_* code; brt_last_error(ctx) returns a ctx-owned
 *     NUL-terminated message that stays valid until the next call on that ctx.
 *   - brt_create binds a ctx to ONE CUDA device; brt_create_multi spans several devices of one process behind the same
 *     calls; with one process per GPU every rank owns a ctx and joins a peer group (brt_peer_*).  A ctx is not
 *     re-entrant: one call at a time, except brt_cancel which may be called from any thread.
 *   - Host descriptors are double precision (JavaScript Numbers are doubles); the device path computes in
 *     fp32 except where stated.
 *   - Images are row-major with row 0 = TOP (pixelIndex = ((H-1-j)*W + i)*4, ray-tracer.js:215).
 *   - There is no CPU fallback: every compute entry point fails with BRT_E_CUDA when no sm_100 device is usable.
 */
#ifndef BRT_H
#define BRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRT_ABI_VERSION 3

typedef struct brt_ctx brt_ctx;

/* ---- error codes --------------------------------------------------------------------------------- */
enum {
    BRT_OK = 0,
    BRT_E_INVALID = -1,    /* bad argument / descriptor */
    BRT_E_CUDA = -2,       /* CUDA runtime failure or no usable device */
    BRT_E_PARSE = -3,      /* scene JSON rejected (the reference would throw; ray-tracer.js:330-333) */
    BRT_E_NOSCENE = -4,    /* render requested before a scene / camera was set */
    BRT_E_CANCELLED = -5,  /* brt_cancel() observed (window.renderCancelled, ray-tracer.js:190,256,264) */
    BRT_E_NOMEM = -6,
    BRT_E_STATE = -7
};

/* ---- enumerations (values are ABI) ---------------------------------------------------------------- */
enum { BRT_OBJ_SPHERE = 0, BRT_OBJ_PLANE = 1, BRT_OBJ_BOX = 2, BRT_OBJ_TRIANGLE = 3, BRT_OBJ_MESH = 4 };     /* js/geometry.js */
enum { BRT_MAT_LAMBERTIAN = 0, BRT_MAT_METAL = 1, BRT_MAT_DIELECTRIC = 2, BRT_MAT_EMISSIVE = 3 };            /* js/materials.js */
enum { BRT_TEX_SOLID = 0, BRT_TEX_CHECKER = 1, BRT_TEX_NOISE = 2, BRT_TEX_MARBLE = 3, BRT_TEX_WOOD = 4 };       /* js/textures.js */
enum { BRT_LIGHT_POINT = 0, BRT_LIGHT_DIRECTIONAL = 1 };                                                    /* js/lights.js */
enum { BRT_BG_GRADIENT = 0, BRT_BG_SOLID = 1, BRT_BG_HDRI = 2, BRT_BG_PROCEDURAL_SKY = 3 };                 /* js/world.js:35-110 */
/* CENTER = any other antiAliasing string: pixel-centre samples, but still `samples` of them (ray-tracer.js:142-148 with :201) */
enum { BRT_AA_NONE = 0, BRT_AA_SUPERSAMPLING = 1, BRT_AA_STOCHASTIC = 2, BRT_AA_CENTER = 3 };               /* ray-tracer.js:125-149 */
enum { BRT_TONEMAP_REINHARD = 0, BRT_TONEMAP_ACES = 1, BRT_TONEMAP_LINEAR = 2 };                            /* ray-tracer.js:151-161 */
/* camera.js:25 tests `type === 'perspective'`, camera.js:39 tests `type === 'orthographic'`; any other
 * string gets the un-scaled viewport with the perspective ray formula = BRT_CAM_OTHER. */
enum { BRT_CAM_PERSPECTIVE = 0, BRT_CAM_ORTHOGRAPHIC = 1, BRT_CAM_OTHER = 2 };
/* FAST: direct-inversion sampling, one Philox block per bounce (same distributions as math.js:22-31); fp32 throughout.
 * REFERENCE: rejection sampling, draws consumed in the reference's exact order from the sequential
 * (seed, pixel, sample) stream — sample-for-sample comparable with the float64 oracle; primary rays are generated and
 * the primary hit is evaluated in float64 (as in the AOV kernel). */
enum { BRT_SAMPLER_FAST = 0, BRT_SAMPLER_REFERENCE = 1 };
/* MEGAKERNEL: one thread per pixel, blocking traversal.  WAVEFRONT: persistent warp-local wavefront (path slots and ray queue
 * in shared memory).  AUTO = MEGAKERNEL, the faster of the two on every measured configuration.  Both trace
 * the same paths; images agree up to fp32 summation order.  count_tests renders always run the megakernel. */
enum { BRT_INTEGRATOR_AUTO = 0, BRT_INTEGRATOR_MEGAKERNEL = 1, BRT_INTEGRATOR_WAVEFRONT = 2 };
/* BRUTE reproduces the reference's linear loops (world.js:24-30, geometry.js:253-259); BVH must give identical results. */
enum { BRT_ACCEL_AUTO = 0, BRT_ACCEL_BRUTE = 1, BRT_ACCEL_BVH = 2 };
#define BRT_BVH_WIDTH_AUTO 2   /* what bvh_width = 0 selects: the binary tree is the fastest on every measured config */

/* ---- scene descriptors ----------------------------------------------------------------------------- */
typedef struct brt_material {
    int32_t type;        /* BRT_MAT_* */
    int32_t texture;     /* 0 = none; k > 0 = brt_scene_desc.textures[k-1]: TexturedLambertian / TexturedMetal (materials.js:99-126),
                            attenuation = texture.value(u, v, point) instead of `color` (lambertian / metal only) */
    double color[3];     /* albedo (lambertian/metal) or emissive colour; ignored for dielectric */
    double param;        /* metal: roughness (clamped min(r,1), materials.js:33) | dielectric: ior | emissive: intensity */
} brt_material;

/* js/textures.js:9-82.  Every value(u, v, p) there uses only the hit point p.  The reference never instantiates these
 * (createTexture has no caller, ray-tracer.js:79-100) and its scene JSON has no texture fields, so textures reach the
 * engine only through brt_scene_set_flat.  Each noise-based texture owns a PerlinNoise with a random permutation
 * (textures.js:44,58,74): `perm` is that table (first 256 entries of PerlinNoise.p). */
typedef struct brt_texture {
    int32_t kind;        /* BRT_TEX_* */
    int32_t _pad;
    double odd[3];       /* checker: odd colour | solid: the colour */
    double even[3];      /* checker: even colour */
    double scale;        /* checker / noise / marble / wood scale */
    uint8_t perm[256];
} brt_texture;

/* One entry per element of world.objects, IN ORDER — the index is the "object ID" (world.js:24-30). */
typedef struct brt_object {
    int32_t type;        /* BRT_OBJ_* */
    int32_t material;    /* index into brt_scene_desc.materials */
    double a[3];         /* sphere: center | plane: point  | box: min | triangle: v0 */
    double b[3];         /* sphere: b[0] = radius (sign kept, geometry.js:34) | plane: normal (normalised by the library, geometry.js:52) | box: max | triangle: v1 */
    double c[3];         /* triangle: v2 */
    int64_t first_tri;   /* mesh: first triangle in brt_scene_desc.mesh_triangles */
    int64_t tri_count;   /* mesh: number of triangles (post-filter ordinals = triangle IDs, geometry.js:206-231) */
} brt_object;

typedef struct brt_light {
    int32_t type;        /* BRT_LIGHT_* */
    int32_t _pad;
    double v[3];         /* point: position | directional: direction (normalised by the library, lights.js:38) */
    double color[3];
    double intensity;
} brt_light;

/* flags: 0 = the rows are CONSTRUCTOR ARGUMENTS (new Plane(point, normal), new DirectionalLight(direction, ...), new Metal(albedo,
 * roughness)): the library applies what those constructors do — plane normals and directional-light directions are normalised
 * (geometry.js:52, lights.js:38), metal roughness is clamped to 1 (materials.js:33).  BRT_SCENE_CONSTRUCTED = the rows were read
 * from CONSTRUCTED objects (the JS shim flattening a live World; brt_scene_get_flat): they are stored as they are — normalising a
 * normalised vector a second time moves its last bit. */
#define BRT_SCENE_CONSTRUCTED 1
typedef struct brt_scene_desc {
    const brt_object* objects;      int32_t n_objects;     int32_t flags;
    const brt_material* materials;  int32_t n_materials;   int32_t _pad1;
    const double* mesh_triangles;   /* 9 doubles per triangle: v0.xyz v1.xyz v2.xyz */
    int64_t n_mesh_triangles;
    const brt_light* lights;        int32_t n_lights;      int32_t _pad2;
    const brt_texture* textures;    int32_t n_textures;    int32_t _pad3;   /* may be NULL / 0 */
} brt_scene_desc;

/* js/camera.js.  Either the constructor arguments (camera.js:8) or, with use_derived = 1, the derived
 * vectors exactly as the reference's own Camera object holds them (camera.js:14-35) — the JS shim passes
 * those so that not even Math.tan can differ. */
typedef struct brt_camera {
    double look_from[3], look_at[3], vup[3];
    double vfov, aspect, aperture, focus_dist;
    int32_t type;          /* BRT_CAM_* */
    int32_t use_derived;
    double origin[3], lower_left_corner[3], horizontal[3], vertical[3], u[3], v[3], w[3];
    double lens_radius;
} brt_camera;

/* RayTracer settings (ray-tracer.js:23-30, 554-566) plus engine knobs. */
typedef struct brt_render_params {
    int32_t width, height;
    int32_t spp;                 /* this.samples; AA 'none' forces 1 sample (ray-tracer.js:201) */
    int32_t max_depth;           /* this.maxBounces: at most this many intersections per path (ray-tracer.js:103); 0..250 */
    int32_t aa_mode;             /* BRT_AA_* */
    int32_t tonemap;             /* BRT_TONEMAP_* */
    double exposure, gamma;
    int32_t denoise;             /* 3x3 gaussian on the tone-mapped image (post-processor.js:45-77) */
    int32_t _pad0;
    double denoise_strength;
    uint64_t seed;               /* Philox key; Math.random (math.js:21) is unseeded in the reference */
    int32_t direct_lighting;     /* EXTENSION, default 0: point/directional shadow rays (lights.js:22-47 are never called by the reference) */
    int32_t sampler;             /* BRT_SAMPLER_* */
    int32_t integrator;          /* BRT_INTEGRATOR_* */
    int32_t accel;               /* BRT_ACCEL_* */
    int32_t spp_batch;           /* samples per launch between progress callbacks / cancel polls; 0 = auto */
    int32_t count_tests;         /* 1 = counting build of the same traversal (fills brt_stats.tests_*) */
    int32_t refill_threshold;    /* WAVEFRONT tuning: idle lanes of a warp fetch the next queued ray once this many are idle; 0 = default (8) */
    int32_t paths_in_flight;     /* WAVEFRONT tuning: samples of a pixel in flight per lane (1..4); 0 = default (2) */
    int32_t preview;             /* 1 = progressive preview: before every progress callback the image of the samples traced so far
                                    is resolved into the caller's rgba8 buffer (the reference blits finished rows, ray-tracer.js:236-238) */
    int32_t bvh_width;           /* children per hierarchy node walked by the megakernel with the FAST sampler: 0 = auto, 2 = the binary
                                    LBVH, 4 / 8 = its wide collapse (results are identical; the reference has no hierarchy at all) */
} brt_render_params;

typedef struct brt_scene_info {
    int32_t n_objects, n_materials, n_lights;
    int32_t n_spheres, n_planes, n_boxes;
    int64_t n_triangles;         /* standalone + mesh triangles */
    int64_t n_bvh_nodes;
    int32_t bvh_depth;
    int32_t bvh_width;           /* width of the hierarchy the current render parameters select: 2, 4 or 8 (0 = none built) */
    double bvh_build_ms;         /* device time of the LBVH build (Morton, radix sort, Karras, refit) */
    double upload_ms;
    int64_t upload_bytes;        /* host -> device bytes of the flattened SoA scene (primitives, meta, materials, lights) */
    int32_t bvh_wide_depth, _pad;/* levels of the wide hierarchy (0 when the binary one is walked) */
    double bvh_wide_build_ms;    /* device time of the wide collapse */
} brt_scene_info;

typedef struct brt_stats {
    uint64_t samples;            /* path samples traced by the last render / accumulate call */
    uint64_t rays;               /* rays traced (count_tests only) */
    uint64_t tests_sphere, tests_plane, tests_box;
    uint64_t tests_tri_a, tests_tri_b, tests_tri_c;   /* Möller–Trumbore stages reached */
    uint64_t tests_aabb;         /* BVH child-slab tests */
    double kernel_ms;            /* device time of the path-tracing launches */
    double post_ms;              /* resolve (+ denoise) */
    double total_ms;             /* brt_render wall time incl. copies */
    uint64_t launches;           /* kernels launched by the last call */
    /* SIMD-lane attribution of the megakernel (count_tests only; one count per WARP iteration of the loop named):
     * 32*trav_warp_iters lane slots = trav_lane_iters working + (trav_alive_lanes - trav_lane_iters) waiting for the slowest
     * ray of the warp + (32*trav_warp_iters - trav_alive_lanes) drained (no samples left for that lane). */
    uint64_t trav_warp_iters, trav_lane_iters, trav_alive_lanes;
    uint64_t trav_node_issues, trav_leaf_issues, trav_leaf_lanes;   /* warp iterations with >= 1 lane at a node / at a leaf; lanes at a leaf */
    uint64_t path_warp_iters, path_lane_iters;                       /* the path loop (one trace call per iteration) */
    uint64_t node_visits;        /* internal-node visits (count_tests only); tests_aabb / node_visits = child boxes tested per visit */
} brt_stats;

typedef void (*brt_progress_cb)(double fraction, void* user);   /* onProgress (ray-tracer.js:258-259,279) */

/* ---- lifecycle -------------------------------------------------------------------------------------- */
int brt_abi_version(void);
const char* brt_version(void);
/* device_id >= 0: a CUDA device.  device_id == -1: a HOST-ONLY context for scene ingest / camera / parameter logic
 * (CPU-side tests, tooling); every compute entry point on it returns BRT_E_CUDA — libbrt has no CPU renderer. */
int brt_create(brt_ctx** out, int device_id);
void brt_destroy(brt_ctx* ctx);
const char* brt_last_error(const brt_ctx* ctx);
/* Launch on this cudaStream_t (e.g. torch's current stream) instead of the ctx-owned one; NULL restores it.
 * The legacy default stream is named by its CUDA handle cudaStreamLegacy = (cudaStream_t)0x1, not by 0. */
int brt_set_stream(brt_ctx* ctx, void* cuda_stream);

/* ---- scene ------------------------------------------------------------------------------------------- */
/* SceneLoader.loadFromJSON (scene-loader.js:20-84) + RayTracer.loadFromJSON (ray-tracer.js:305-334):
 * parses the Blender-exported JSON, applies every default / skip rule, replaces the scene, replaces the
 * camera iff the JSON has one, and reports camera.resolution when present (out_w/out_h = 0 otherwise).
 * The caller then applies resizeCanvas (ray-tracer.js:598-614) through brt_set_render_params/brt_set_camera. */
int brt_scene_load_json(brt_ctx* ctx, const char* utf8, size_t len, int fallback_w, int fallback_h,
                        int* out_has_camera, int* out_w, int* out_h);
/* The same ingest from the compact container tools/scene_binary.py writes — "BRTSCN01" | u64 json_len | scene JSON |
 * pad to 8 | blob — where a mesh may carry `vertices_bin` / `indices_bin` = {offset, count} (float64 vertex triples /
 * uint32 indices in the blob) instead of `vertices` / `indices`.  Same defaults, skip rules and triangle filtering; it only
 * spares the text round trip of large meshes (the 1.0 M-triangle scene: 38 MB of JSON text vs 24.8 MB binary). */
int brt_scene_load_binary(brt_ctx* ctx, const void* bytes, size_t len, int fallback_w, int fallback_h,
                          int* out_has_camera, int* out_w, int* out_h);
/* World.add / addLight with live objects flattened by the caller (world.js:16-18).  Borrowed; copied before return. */
int brt_scene_set_flat(brt_ctx* ctx, const brt_scene_desc* desc);
int brt_scene_info_get(brt_ctx* ctx, brt_scene_info* out);
/* The scene as held after ingest: world.objects / world.lights in order (scene-loader.js:59-76).  Pointers are
 * borrowed from the ctx and stay valid until the next scene call. */
int brt_scene_get_flat(brt_ctx* ctx, brt_scene_desc* out);

int brt_set_camera(brt_ctx* ctx, const brt_camera* cam);            /* new Camera(...) (camera.js:8-36) */
int brt_get_camera(brt_ctx* ctx, brt_camera* out);                  /* Camera.debugReport (camera.js:56-78) */
/* world.background / world.skyIntensity / world.cloudNoise.p (world.js:12-14; ray-tracer.js:568-585).
 * perm256 = the first 256 entries of PerlinNoise.p (noise.js:7-13), NULL keeps the current table. */
int brt_set_background(brt_ctx* ctx, int kind, const double color[3], double intensity, const uint8_t* perm256);
int brt_get_background(brt_ctx* ctx, int* kind, double color[3], double* intensity);
int brt_set_render_params(brt_ctx* ctx, const brt_render_params* p);
int brt_get_render_params(brt_ctx* ctx, brt_render_params* out);

/* ---- render (the replaced seam: ray-tracer.js:166-281) ------------------------------------------------ */
/* Blocking.  rgba8: W*H*4 bytes (the Uint8ClampedArray behind imageData.data).  float_data (nullable):
 * W*H*4 fp32 = the reference's `floatData` (tone-mapped + gamma'd, alpha 1).  linear_mean (nullable):
 * W*H*4 fp32 per-pixel mean radiance before tone mapping.  All three are HOST buffers. */
int brt_render(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean, brt_progress_cb cb, void* user);
void brt_cancel(brt_ctx* ctx);
int brt_get_stats(brt_ctx* ctx, brt_stats* out);

/* ---- device-resident pieces (multi-GPU spp split, benchmarking) ---------------------------------------- */
/* Adds the radiance of samples [sample_begin, sample_begin+sample_count) of EVERY pixel into d_accum
 * (DEVICE fp32 RGBA sums, W*H*4; alpha accumulates the sample count).  Asynchronous on the ctx stream.
 * The RNG is keyed by (seed, pixel, global sample index), so any partition of the samples over ranks
 * yields the same sample set. */
int brt_render_accumulate(brt_ctx* ctx, float* d_accum, int sample_begin, int sample_count);
/* color/spp -> tone map -> gamma -> floor(c*255) (ray-tracer.js:208-233) [-> denoise :266-276] on DEVICE buffers.
 * d_float_data / d_linear_mean nullable.  Asynchronous on the ctx stream. */
int brt_resolve_device(brt_ctx* ctx, const float* d_accum, uint8_t* d_rgba8, float* d_float_data, float* d_linear_mean);
/* Fused cross-GPU reduce + resolve over peer-mapped accumulation buffers (NVLink P2P): this rank sums rows
 * [row_begin,row_end) of all n_peers buffers in fixed rank order, resolves them and writes RGBA8 into
 * d_rgba8_root (a peer-mapped pointer on the root GPU, or local).  See INTEGRATION.md. */
int brt_reduce_resolve_peers(brt_ctx* ctx, const float* const* d_peer_accum, int n_peers, int row_begin, int row_end,
                             uint8_t* d_rgba8_root, float* d_float_data_root);
int brt_stream_synchronize(brt_ctx* ctx);

/* ---- multi-GPU: the samples-per-pixel split with the exchange fused into its consumer ------------------------
 * (a) ONE process, n GPUs — what a Node host binds: a context that spans the listed devices.  Every other call is unchanged:
 *     brt_render(ctx, host_rgba8, ...) splits each batch's samples over the devices, every device traces its share, one fused
 *     kernel per device pulls its row stripe of all devices' fp32 sums over NVLink (peer access), sums in device order,
 *     tone-maps and stores RGBA8 into device 0's image, which is copied into the caller's buffer.  The reference's caller
 *     still makes ONE call (js/ui-controller.js:189 `await raytracer.render(cb)`).  Scene / camera / parameter setters on
 *     the returned context apply to all devices.  n = 1 is the same as brt_create. */
int brt_create_multi(brt_ctx** out, const int* device_ids, int n_devices);
int brt_device_count(const brt_ctx* ctx);
/* (b) ONE process PER GPU (torchrun / MPI): rank r of `world` allocates its exchange block (sized from the current render
 *     params; double-buffered fp32 sums, RGBA8 / floatData / linear images, three epoch flag words) and gets a 64-byte CUDA
 *     IPC handle for it; the caller gathers the handles of all ranks (world x 64 bytes, rank order) and every rank connects.
 *     brt_peer_render then runs one exchange epoch asynchronously on the ctx stream: zero, trace this rank's samples
 *     [sample_begin, sample_begin + sample_count), publish "ready", wait for every peer's flag (acquire loads on peer-mapped
 *     memory: no host barrier, no NCCL), reduce this rank's row stripe in rank order, resolve, store into RANK 0's image.
 *     Every rank must call it once per epoch with its own sample range.  brt_peer_fetch on rank 0 waits until all stripes
 *     have arrived and copies the image(s) into HOST buffers (any may be NULL); on other ranks it only synchronises.
 *     A peer that never arrives makes the wait time out after 20 s (BRT_E_STATE), it cannot hang the GPU. */
int brt_peer_alloc(brt_ctx* ctx, int rank, int world, uint8_t handle[64]);
int brt_peer_connect(brt_ctx* ctx, const uint8_t* handles /* world x 64 bytes */);
int brt_peer_render(brt_ctx* ctx, int sample_begin, int sample_count, int want_float_data, int want_linear_mean);
int brt_peer_fetch(brt_ctx* ctx, uint8_t* rgba8, float* float_data, float* linear_mean);
int brt_peer_image_ptr(brt_ctx* ctx, void** d_rgba8);      /* this rank's RGBA8 image region (the group's image on rank 0), DEVICE pointer */
int brt_peer_free(brt_ctx* ctx);

/* ---- unsynchronised building blocks of the same exchange (the caller orders the ranks itself, e.g. with NCCL) --------------
 * Buffers for the cross-process (one process per GPU) form of the fused reduce: cudaMalloc'd by the library so that a
 * CUDA IPC handle (64 opaque bytes, sent to the peers by the caller, e.g. torch.distributed.all_gather_object) names
 * them.  brt_shared_open maps a peer's buffer into this process (NVLink peer access); close before the owner frees. */
int brt_shared_alloc(brt_ctx* ctx, size_t bytes, void** d_ptr, uint8_t handle[64]);
int brt_shared_free(brt_ctx* ctx, void* d_ptr);
int brt_shared_open(brt_ctx* ctx, const uint8_t handle[64], void** d_ptr);
int brt_shared_close(brt_ctx* ctx, void* d_ptr);
/* cudaMemsetAsync on the ctx stream (zeroing an accumulation buffer between renders). */
int brt_device_memset(brt_ctx* ctx, void* d_ptr, int value, size_t bytes);
/* Blocking device -> host copy on the ctx stream (reading back a brt_shared_alloc buffer). */
int brt_copy_to_host(brt_ctx* ctx, void* host_dst, const void* d_src, size_t bytes);

/* ---- parity AOVs (north star: primary-hit IDs bit-exact, t / normal within 1e-5) ------------------------ */
/* Primary visibility at pixel centres with lens offset 0 (ray-tracer.js:144-147, camera.js:45-49, world.js:20-33).
 * HOST outputs, W*H each (normal3: W*H*3).  Miss: obj_id = tri_id = -1, t = +inf.
 * _f32 runs the render path's own fp32 intersection code (BVH or brute per render params);
 * _f64 runs a float64, FMA-free, brute-force kernel with the reference's exact operation order. */
int brt_primary_aov_f32(brt_ctx* ctx, int32_t* obj_id, int32_t* tri_id, float* t, float* normal3, uint8_t* front_face);
int brt_primary_aov_f64(brt_ctx* ctx, int32_t* obj_id, int32_t* tri_id, double* t, double* normal3, uint8_t* front_face);

/* ---- unit-level hooks used by the parity tests and the benchmark ----------------------------------------- */
/* world.background(ray) for n directions (HOST in: n*3 doubles; HOST out: n*3 floats), evaluated by the device code. */
int brt_eval_background(brt_ctx* ctx, const double* dirs, int n, float* out_rgb);
/* First n uniforms of the (seed, pixel, sample) Philox stream as the device produces them. */
int brt_debug_rng_stream(brt_ctx* ctx, uint64_t seed, uint32_t pixel, uint32_t sample, int n, float* out);
/* Run only the post-processing kernels on a HOST linear-mean image (W*H*4 fp32) with the ctx render params. */
int brt_postprocess_host(brt_ctx* ctx, const float* linear_mean, uint8_t* rgba8, float* float_data);
/* textures[tex_index].value(u, v, p) for n points (HOST in: n*3 doubles; HOST out: n*3 floats), evaluated by the device code. */
int brt_eval_texture(brt_ctx* ctx, int tex_index, const double* points, int n, float* out_rgb);
/* Dense FFMA micro-benchmark on this device: the measured FP32 roofline denominator (TFLOP/s). */
int brt_measure_fp32_peak(brt_ctx* ctx, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* BRT_H */
