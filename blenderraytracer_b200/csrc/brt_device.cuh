// Device-side core of libbrt: scene views, Philox RNG, fp32 intersection, BVH traversal, materials and
// backgrounds.  Every routine cites the reference JavaScript it reproduces (Shinzef/BlenderRayTracer).
// This is new sm_100a code, not a translation: SoA float4 primitive arrays read with 128-bit loads through
// the read-only path, an Aila–Laine style 64-byte two-child BVH node, a shared-memory traversal stack, and
// counter-based RNG so that every (pixel, sample) is independent of how work is partitioned.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math_constants.h>
#include <cstdio>

namespace brt {

// ------------------------------------------------------------------------------------------- ids
// A primitive id packs the primitive type (bits 28..29) and its index within the type's array.
// Bit 31 marks a BVH leaf reference on the traversal stack.
constexpr uint32_t PID_NONE = 0x7FFFFFFFu;
constexpr uint32_t LEAF_BIT = 0x80000000u;
enum PrimType : uint32_t { PT_SPHERE = 0, PT_PLANE = 1, PT_BOX = 2, PT_TRI = 3 };
__host__ __device__ inline uint32_t make_pid(uint32_t type, uint32_t idx) { return (type << 28) | idx; }
__host__ __device__ inline uint32_t pid_type(uint32_t pid) { return (pid >> 28) & 3u; }
__host__ __device__ inline uint32_t pid_index(uint32_t pid) { return pid & 0x0FFFFFFFu; }

// ------------------------------------------------------------------------------------------- scene views
struct DevScene {
    const float4* sph;   // (cx, cy, cz, r)                                  geometry.js:9-13
    const float4* pln;   // 2 per plane: (n.xyz, 0), (p.xyz, 0)              geometry.js:50-54
    const float4* box;   // 2 per box: min, max                              geometry.js:79-83
    const float4* tri;   // 3 per triangle: v0, e1 = v1-v0, e2 = v2-v0       geometry.js:137-146
    const int4* meta;    // per primitive (unified index): {objId, matId, triId, 0}
    const float4* mat;   // (r, g, b, param)                                 materials.js
    const int* matType;  // BRT_MAT_*
    const float4* nodes; // 4 x float4 per BVH node (see bvh.cu)
    const float4* cnodes;// the same binary hierarchy with child boxes as centre / half-extent (bvh.cu: k_centre_half; node_visit_ch)
    const float4* wnodes;// wide hierarchy collapsed from `nodes` (bvh.cu: k_wide_level): 2 x float4 per child slot, wideN slots per node
    const float4* lights;// 2 per light: (v.xyz, type), (color*intensity .xyz, 0)
    const unsigned char* perm;  // 512-entry doubled Perlin permutation      noise.js:7-17
    const float4* tex;    // 2 per texture: (odd.rgb, kind), (even.rgb, scale)        textures.js
    const unsigned char* texPerm;   // 512 per texture: its own doubled Perlin table     textures.js:44,58,74
    const double* prim64; // 9 doubles per primitive (unified index, as meta): float64 copy used ONLY to evaluate the
                          // primitive a PRIMARY ray's fp32 traversal selected (t / P / N within 1e-5 of the float64 reference)
    int nSph, nPln, nBox, nTri;
    int baseSph, basePln, baseBox, baseTri;   // offsets into meta
    int nNodes, nLights, nTex;
    int bgKind;
    float bgR, bgG, bgB, skyIntensity;
    int bvhStackDepth;
    int wideN;           // 0 = no wide hierarchy; 4 / 8 = children per wide node
    int wideDepth;       // levels of the wide hierarchy (= traversal stack entries per ray)
    int wideAxes;        // wideN == 4: the two axes (a0 | a1 << 2) whose ray signs order the four child slots
};

struct DevCamera {         // camera.js:14-35, derived on the host, kept in float64: primary rays are generated in double
    double o[3];           // (12 DFMA per camera sample) and rounded once to fp32 for traversal
    double ll[3];          // lowerLeftCorner
    double h[3], v[3];     // horizontal, vertical
    double cu[3], cv[3], cw[3];   // camera.u / .v / .w
    double lensRadius;
    int type;
};

struct DevCamera32 {       // the same camera rounded once on the host for the fp32 (fast-sampler) ray generator;
    float o[3], llo[3];    // llo = lowerLeftCorner - origin, formed in float64 before rounding
    float h[3], v[3], cu[3], cv[3], cw[3];
    float lensRadius;
    int type;
};

struct Counters {          // counting build
    unsigned long long rays, sph, pln, box, triA, triB, triC, aabb;
    // SIMD-lane attribution of the BVH loop (one count per WARP iteration, taken by the lowest active lane):
    unsigned long long trIter;       // warp-level iterations of the traversal loop
    unsigned long long trLanes;      // lanes that executed them (sum of popc(active mask))
    unsigned long long trAlive;      // lanes of the warp that still had samples to trace at that time (32 - drained)
    unsigned long long trNodeIssue;  // warp iterations in which >= 1 lane visited an internal node
    unsigned long long trLeafIssue;  // warp iterations in which >= 1 lane tested a leaf primitive
    unsigned long long trLeafLanes;  // lanes that tested a leaf primitive
    unsigned long long mainIter;     // warp-level iterations of the path loop (one trace call each)
    unsigned long long mainLanes;    // lanes that executed them
    unsigned long long nodeVisits;   // internal-node visits (binary: aabb / 2; wide: one per node, aabb counts its occupied slots)
};
constexpr int N_COUNTERS = sizeof(Counters) / sizeof(unsigned long long);

// ------------------------------------------------------------------------------------------- float3 helpers
// Every operation is spelled with an explicit rounding intrinsic (or an explicit fmaf), so nvcc's context-dependent
// mul+add contraction cannot make the same expression round differently in two instantiations: the brute-force and BVH
// kernels then produce bit-identical hits, hit points and scattered rays (tests: BVH must not change results).
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z)); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(float3 a, float3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, __fmul_rn(a.x, b.x))); }
__device__ __forceinline__ float3 cross(float3 a, float3 b) {
    return f3(fmaf(a.y, b.z, -__fmul_rn(a.z, b.y)), fmaf(a.z, b.x, -__fmul_rn(a.x, b.z)), fmaf(a.x, b.y, -__fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float3 madd(float3 a, float s, float3 b) { return f3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z)); }   // a*s + b
__device__ __forceinline__ float3 normalize0(float3 a) {                  // math.js:18 (zero vector stays zero)
    float l2 = dot(a, a);
    return l2 > 0.f ? a * rsqrtf(l2) : f3(0.f, 0.f, 0.f);
}
__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 reflect(float3 v, float3 n) { return madd(n, -2.f * dot(v, n), v); }   // math.js:19

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }   // FMNMX3 (sm_100)
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

// ------------------------------------------------------------------------------------------- Philox4x32-10
// Stands in for Math.random (math.js:21-31).  counter = (pixel, sample, block, tag), key = seed.
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}
// The fast sampler's generator.  Philox4x32-7 is the fewest rounds that pass BigCrush (Salmon et al., SC'11, table 2);
// -10 is Random123's safety-margin default and what the sequential (reference) sampler and the oracle use.
#ifndef BRT_PHILOX_FAST_ROUNDS
#define BRT_PHILOX_FAST_ROUNDS 10
#endif
__device__ __forceinline__ uint4 philox_fast(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    return philox4x32<BRT_PHILOX_FAST_ROUNDS>(c0, c1, c2, c3, k0, k1);
}
// Reciprocal used by the fp32 intersection code: the 1-ulp MUFU approximation (one instruction instead of the ~8 of the
// IEEE sequence; +6 % on C3).  Deterministic, so brute force and BVH still agree bit for bit; primary-hit t / normals that
// are compared with the float64 reference come from refine_primary, not from here.  -DBRT_IEEE_RCP restores __frcp_rn.
#ifdef BRT_IEEE_RCP
__device__ __forceinline__ float rcpf(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float sqrtfa(float x) { return sqrtf(x); }
#else
__device__ __forceinline__ float rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sqrtfa(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }   // same reasoning for the sphere root
#endif
constexpr uint32_t PHILOX_TAG = 0x42525431u;   // "BRT1"
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// Sequential stream for BRT_SAMPLER_REFERENCE: the reference's draw order, one uniform at a time.  The whole state is
// one counter (`pos` = uniforms consumed so far), so a path can be suspended to shared memory and resumed.
struct RngSeq {
    uint32_t pix, samp, pos, k0, k1;
    uint4 buf;
    __device__ __forceinline__ void init(uint32_t pixel, uint32_t sample, uint32_t seedLo, uint32_t seedHi) { resume(pixel, sample, 0u, seedLo, seedHi); }
    __device__ __forceinline__ void resume(uint32_t pixel, uint32_t sample, uint32_t position, uint32_t seedLo, uint32_t seedHi) {
        pix = pixel; samp = sample; pos = position; k0 = seedLo; k1 = seedHi;
        if (pos & 3u) buf = philox4x32_10(pix, samp, pos >> 2, PHILOX_TAG, k0, k1);
    }
    __device__ __forceinline__ float next() {
        uint32_t k = pos & 3u;
        if (k == 0u) buf = philox4x32_10(pix, samp, pos >> 2, PHILOX_TAG, k0, k1);
        uint32_t v = k == 0u ? buf.x : k == 1u ? buf.y : k == 2u ? buf.z : buf.w;
        pos++;
        return u01(v);
    }
};

// ------------------------------------------------------------------------------------------- primitive tests
// All tests return the candidate parametric t chosen by the reference's root-selection logic for t >= tMin
// (tMin = 0.001, ray-tracer.js:105); the caller applies the `t > closestT` rejection and the tie rule.
// `self` = the ray starts ON this primitive.  In exact arithmetic one root is then exactly 0 (< tMin, always
// rejected by the reference); in fp32 it is noise of either sign, so that root is dropped analytically.

// geometry.js:15-29.  Discriminant evaluated as a*(r^2 - |oc - (hb/a) D|^2) (cancellation-robust form of hb^2 - a c).
__device__ __forceinline__ bool hit_sphere(float4 s, float3 O, float3 D, float tMin, bool self, float& t) {
    float a = dot(D, D), inva = rcpf(a);
    float3 oc = O - f3(s.x, s.y, s.z);
    float hb = dot(oc, D);
    if (self) { t = __fmul_rn(__fmul_rn(-2.f, hb), inva); return t >= tMin; }
    float k = __fmul_rn(hb, inva);
    float3 l = madd(D, -k, oc);
    float disc = fmaf(s.w, s.w, -dot(l, l));
    if (disc < 0.f) return false;
    float sq = sqrtfa(__fmul_rn(a, disc));
    float root = __fmul_rn(-hb - sq, inva);
    if (root < tMin) {
        root = __fmul_rn(sq - hb, inva);
        if (!(root >= tMin)) return false;
    }
    t = root;
    return true;
}
// geometry.js:56-61
__device__ __forceinline__ bool hit_plane(float4 n, float4 p, float3 O, float3 D, float tMin, float& t) {
    float3 N = xyz(n);
    float denom = dot(N, D);
    if (fabsf(denom) < 1e-6f) return false;
    float tt = __fmul_rn(dot(xyz(p) - O, N), rcpf(denom));
    if (!(tt >= tMin)) return false;
    t = tt;
    return true;
}
// geometry.js:85-112.  `face` returns 0..5 = x-,x+,y-,y+,z-,z+ : the slab plane that produced t (the reference picks
// the face by |p - face| < 1e-6, :119-126, which is the same face away from edges; ties resolve x, y, z as there).
__device__ __forceinline__ bool hit_box(float4 mn, float4 mx, float3 O, float3 D, float tMin, bool self, float& t, int& face) {
    float3 inv = f3(rcpf(D.x), rcpf(D.y), rcpf(D.z));
    float t0 = __fmul_rn(mn.x - O.x, inv.x), t1 = __fmul_rn(mx.x - O.x, inv.x);
    int fe = 0, fx = 1;                         // entering / exiting face ids
    if (t0 > t1) { float tmp = t0; t0 = t1; t1 = tmp; fe = 1; fx = 0; }
    float y0 = __fmul_rn(mn.y - O.y, inv.y), y1 = __fmul_rn(mx.y - O.y, inv.y);
    int fye = 2, fyx = 3;
    if (y0 > y1) { float tmp = y0; y0 = y1; y1 = tmp; fye = 3; fyx = 2; }
    if (t0 > y1 || y0 > t1) return false;
    if (y0 > t0) { t0 = y0; fe = fye; }         // Math.max(tMinBox, tMinY): x wins ties
    if (y1 < t1) { t1 = y1; fx = fyx; }
    float z0 = __fmul_rn(mn.z - O.z, inv.z), z1 = __fmul_rn(mx.z - O.z, inv.z);
    int fze = 4, fzx = 5;
    if (z0 > z1) { float tmp = z0; z0 = z1; z1 = tmp; fze = 5; fzx = 4; }
    if (t0 > z1 || z0 > t1) return false;
    if (z0 > t0) { t0 = z0; fe = fze; }
    if (z1 < t1) { t1 = z1; fx = fzx; }
    float tt; int ff;
    if (self) {
        // origin on the box surface: drop the root nearest to zero, keep the other (see header comment)
        if (fabsf(t0) < fabsf(t1)) { tt = t1; ff = fx; } else { tt = t0; ff = fe; }
        if (!(tt >= tMin)) return false;
    } else {
        if (t0 > tMin) { tt = t0; ff = fe; } else { tt = t1; ff = fx; }
        if (!(tt >= tMin)) return false;
    }
    t = tt; face = ff;
    return true;
}
// geometry.js:148-175 (Möller–Trumbore with the reference's absolute 1e-4 parallel threshold).
template <bool COUNT>
__device__ __forceinline__ bool hit_tri(float4 v0, float4 e1, float4 e2, float3 O, float3 D, float tMin, float& t, Counters& cnt) {
    float3 E1 = xyz(e1), E2 = xyz(e2);
    float3 h = cross(D, E2);
    float a = dot(E1, h);
    if (fabsf(a) < 0.0001f) return false;
    float f = rcpf(a);
    float3 s = O - xyz(v0);
    float u = __fmul_rn(f, dot(s, h));
    if (u < 0.f || u > 1.f) return false;
    if (COUNT) cnt.triB++;
    float3 q = cross(s, E1);
    float v = __fmul_rn(f, dot(D, q));
    if (v < 0.f || __fadd_rn(u, v) > 1.f) return false;
    if (COUNT) cnt.triC++;
    float tt = __fmul_rn(f, dot(E2, q));
    if (!(tt >= tMin)) return false;
    t = tt;
    return true;
}

struct Hit {
    float t;
    uint32_t pid;
};

// Tie rule: across objects the FIRST object wins an exact tie (world.js:26 `hit.t < closestT`);
// inside one mesh the LAST triangle wins (geometry.js:175 accepts t == tMax, :255-258 replaces).
__device__ __forceinline__ int meta_index(const DevScene& sc, uint32_t pid) {
    uint32_t ty = pid_type(pid), ix = pid_index(pid);
    int base = ty == PT_SPHERE ? sc.baseSph : ty == PT_PLANE ? sc.basePln : ty == PT_BOX ? sc.baseBox : sc.baseTri;
    return base + (int)ix;
}
static __device__ __forceinline__ bool tie_wins(const DevScene& sc, uint32_t cand, uint32_t cur) {
    int4 mc = __ldg(&sc.meta[meta_index(sc, cand)]);
    int4 mb = __ldg(&sc.meta[meta_index(sc, cur)]);
    if (mc.x != mb.x) return mc.x < mb.x;
    return mc.z > mb.z;
}
template <bool SHADOW>
__device__ __forceinline__ void consider(const DevScene& sc, Hit& best, float t, uint32_t pid) {
    if (SHADOW) {
        // any-hit for the direct-lighting extension: strictly closer than the light, emissive primitives ignored
        if (t < best.t) {
            int4 m = __ldg(&sc.meta[meta_index(sc, pid)]);
            if ((__ldg(&sc.matType[m.y]) & 255) != 3) { best.t = t; best.pid = pid; }
        }
        return;
    }
    if (t < best.t || (t == best.t && best.pid != PID_NONE && tie_wins(sc, pid, best.pid))) {
        best.t = t; best.pid = pid;
    }
}

// PRIMS: compile-time mask of the bounded primitive types the BVH can hold (1 sphere | 2 box | 4 triangle; 7 = any).  A scene
// with a single bounded type (random spheres; one big mesh) runs a kernel whose leaf code has no type dispatch at all.
constexpr int PRIMS_ANY = 7, PRIMS_SPHERE = 1, PRIMS_BOX = 2, PRIMS_TRI = 4;
template <bool COUNT, bool SHADOW, int PRIMS = PRIMS_ANY>
__device__ __forceinline__ void test_prim(const DevScene& sc, uint32_t pid, float3 O, float3 D, float tMin, uint32_t self, Hit& best, Counters& cnt) {
    uint32_t ty = pid_type(pid), ix = pid_index(pid);
    if (PRIMS == PRIMS_SPHERE) ty = PT_SPHERE; else if (PRIMS == PRIMS_TRI) ty = PT_TRI; else if (PRIMS == PRIMS_BOX) ty = PT_BOX;
    float t; int face = 0;
    bool h;
    if (ty == PT_TRI) {
        if (pid == self) return;                  // a ray leaving a planar primitive cannot meet it again
        if (COUNT) cnt.triA++;
        h = hit_tri<COUNT>(ldg4(sc.tri + 3 * ix), ldg4(sc.tri + 3 * ix + 1), ldg4(sc.tri + 3 * ix + 2), O, D, tMin, t, cnt);
    } else if (ty == PT_SPHERE) {
        if (COUNT) cnt.sph++;
        h = hit_sphere(ldg4(sc.sph + ix), O, D, tMin, pid == self, t);
    } else {
        if (COUNT) cnt.box++;
        h = hit_box(ldg4(sc.box + 2 * ix), ldg4(sc.box + 2 * ix + 1), O, D, tMin, pid == self, t, face);
    }
    if (h && t <= best.t) consider<SHADOW>(sc, best, t, pid);
}

// Unbounded planes live outside the BVH in a linear list (world.js:24-30 order is irrelevant given the tie rule).
template <bool COUNT, bool SHADOW>
__device__ __forceinline__ void test_planes(const DevScene& sc, float3 O, float3 D, float tMin, uint32_t self, Hit& best, Counters& cnt) {
    for (int i = 0; i < sc.nPln; i++) {
        uint32_t pid = make_pid(PT_PLANE, i);
        if (pid == self) continue;
        if (COUNT) cnt.pln++;
        float t;
        if (hit_plane(ldg4(sc.pln + 2 * i), ldg4(sc.pln + 2 * i + 1), O, D, tMin, t) && t <= best.t) consider<SHADOW>(sc, best, t, pid);
    }
}

// ------------------------------------------------------------------------------------------- closest hit
// Brute force: the reference's own O(N) loops (world.js:24-30, geometry.js:253-259) over the SoA arrays.
template <bool COUNT, bool SHADOW>
__device__ __forceinline__ Hit trace_brute(const DevScene& sc, float3 O, float3 D, float tMin, float tMax, uint32_t self, Counters& cnt) {
    Hit best; best.t = tMax; best.pid = PID_NONE;
    test_planes<COUNT, SHADOW>(sc, O, D, tMin, self, best, cnt);
    for (int i = 0; i < sc.nSph; i++) test_prim<COUNT, SHADOW>(sc, make_pid(PT_SPHERE, i), O, D, tMin, self, best, cnt);
    for (int i = 0; i < sc.nBox; i++) test_prim<COUNT, SHADOW>(sc, make_pid(PT_BOX, i), O, D, tMin, self, best, cnt);
    for (int i = 0; i < sc.nTri; i++) test_prim<COUNT, SHADOW>(sc, make_pid(PT_TRI, i), O, D, tMin, self, best, cnt);
    return best;
}

// BVH node (64 B, four 128-bit loads):
//   n0 = (c0.min.x, c0.max.x, c0.min.y, c0.max.y)   n1 = same for child 1
//   n2 = (c0.min.z, c0.max.z, c1.min.z, c1.max.z)   n3 = (child0, child1, -, -) as bit patterns;
//   a child with LEAF_BIT set is a primitive id, otherwise an internal node index.
// Traversal stack: the first SMEM_STACK entries of every thread live in shared memory ([depth][thread], conflict
// free); deeper entries (rare) spill to a per-thread local array.
constexpr int SMEM_STACK = 12;
constexpr int LOCAL_STACK = 52;
constexpr int SMEM_ONLY_MAX_DEPTH = 32;   // trees up to this depth run with the whole stack in shared memory (16.5 KB / block at 32)
constexpr uint32_t TRAV_DONE = 0xFFFFFFFFu;

__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
struct RayInv { float3 inv, ood; };
// A direction component that is exactly 0 (axis-parallel camera rays through the centre row / column, mirror bounces off
// axis-aligned faces) or denormal would make inv = +-inf and `bound*inv - ood` = inf - inf = NaN on a slab that straddles the
// origin's coordinate, which fminf / fmaxf then silently drop: the child would be culled although the origin lies inside the
// slab.  Clamping |d| to 2^-80 keeps every slab value finite with the right sign (Aila-Laine); the primitive tests do not use
// this reciprocal, so brute force and BVH still see the same hits.
__device__ __forceinline__ float slab_dir(float d) { return fabsf(d) > 8.271806125530277e-25f ? d : copysignf(8.271806125530277e-25f, d); }
__device__ __forceinline__ RayInv ray_inv(float3 O, float3 D) {
    RayInv r;
    r.inv = f3(rcpf(slab_dir(D.x)), rcpf(slab_dir(D.y)), rcpf(slab_dir(D.z)));
    r.ood = O * r.inv;
    return r;
}

// One internal-node visit: slab-tests both children against [0, tBest] and returns the next node to visit (TRAV_DONE-free:
// `hitAny` false = neither child is hit); `farc` is valid when both were hit.  12 FFMA + 12 FMNMX + 4 FMNMX3.
__device__ __forceinline__ bool node_visit(const float4* __restrict__ nodes, uint32_t cur, const RayInv& r, float tBest,
                                           uint32_t& nearc, uint32_t& farc, bool& both) {
    const float4* np = nodes + 4 * (size_t)cur;
    float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2);
    float4 n3f = ldg4(np + 3);
    uint32_t c0 = __float_as_uint(n3f.x), c1 = __float_as_uint(n3f.y);
    float ax0 = fmaf(n0.x, r.inv.x, -r.ood.x), ax1 = fmaf(n0.y, r.inv.x, -r.ood.x);
    float ay0 = fmaf(n0.z, r.inv.y, -r.ood.y), ay1 = fmaf(n0.w, r.inv.y, -r.ood.y);
    float az0 = fmaf(n2.x, r.inv.z, -r.ood.z), az1 = fmaf(n2.y, r.inv.z, -r.ood.z);
    float bx0 = fmaf(n1.x, r.inv.x, -r.ood.x), bx1 = fmaf(n1.y, r.inv.x, -r.ood.x);
    float by0 = fmaf(n1.z, r.inv.y, -r.ood.y), by1 = fmaf(n1.w, r.inv.y, -r.ood.y);
    float bz0 = fmaf(n2.z, r.inv.z, -r.ood.z), bz1 = fmaf(n2.w, r.inv.z, -r.ood.z);
    float tn0 = fmaxf(fmax3(fminf(ax0, ax1), fminf(ay0, ay1), fminf(az0, az1)), 0.f);
    float tf0 = fminf(fmin3(fmaxf(ax0, ax1), fmaxf(ay0, ay1), fmaxf(az0, az1)), tBest);
    float tn1 = fmaxf(fmax3(fminf(bx0, bx1), fminf(by0, by1), fminf(bz0, bz1)), 0.f);
    float tf1 = fminf(fmin3(fmaxf(bx0, bx1), fmaxf(by0, by1), fmaxf(bz0, bz1)), tBest);
    // conservative: boxes are inflated at build time and the far bound is widened by a few ulps (the slab arithmetic's
    // own error grows with the distance to the ray origin), so the BVH can only add candidates, never lose one the
    // brute-force loop would have found.
    // (without the widening the 1 M-triangle terrain loses hits the linear loop finds — measured, tools/ab.py — so it stays)
    bool h0 = tn0 <= __fmul_rn(tf0, 1.0000005f), h1 = tn1 <= __fmul_rn(tf1, 1.0000005f);
    bool swap = tn1 < tn0;
    both = h0 & h1;
    nearc = both ? (swap ? c1 : c0) : (h0 ? c0 : c1);
    farc = swap ? c0 : c1;
    return h0 | h1;
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2: `fma.rn.f32x2`, `mul.rn.f32x2`): ONE instruction performs the IEEE operation on both
// halves of a 64-bit register pair — same results as two scalar fmaf / __fmul_rn, half the issue slots.  ptxas folds a pair built
// from one register into a broadcast operand (`R.F32`) and a negated pair into a negate modifier, so pk2(x, x) costs nothing.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// The same visit over the centre / half-extent copy of the node (fast-sampler megakernel), children interleaved per axis:
//   n0 = (c0.x, c1.x, c0.y, c1.y)  n1 = (c0.z, c1.z, h0.x, h1.x)  n2 = (h0.y, h1.y, h0.z, h1.z)  n3 = (child0, child1, -, -)
//   [c - h, c + h] contains the lo / hi box; 56 of the 64 bytes are read (three 128-bit loads + one 64-bit), as for the lo / hi node
// With tm = c*inv - ood the two slab planes of an axis are tm -+ h*|inv|: near and far are known WITHOUT the per-axis min / max
// pair, and each of the nine fused multiply-adds serves BOTH children (FFMA2 on the register pair a 128-bit load delivers).  A visit
// is 9 FFMA2 + 4 FMNMX3 + 4 FMNMX + 1 FMUL2 where the lo / hi form needs 12 FFMA + 16 FMNMX + 4 FMNMX3 + 2 FMUL: 12 fewer
// instructions on the ALU pipe (the busiest pipe of this kernel, profiles/) and 4 fewer on the FMA pipe.
// Same widening, same ordering rule; the boxes are supersets of the lo / hi boxes.
#ifndef BRT_NODE_CH
#define BRT_NODE_CH 1
#endif
#ifndef BRT_NODE_LDG256
#define BRT_NODE_LDG256 0      // two LDG.E.256 per node instead of 3 x 128 + 64 bit: measured C5 -1.3 %, C3 / C4 +0.4 %, C2 +1.1 % (B200) — off
#endif
// Slab test of the TWO boxes of one 64-byte centre / half-extent record (see node_visit_ch for the layout): 9 FFMA2 + 1 FMUL2.
// -> entry distances tn and widened exit distances tw of both boxes; box k is hit iff tn_k <= tw_k (the comparisons are left to
// the caller: handed back as bool references they are materialised as 0 / 1 integers instead of predicates)
__device__ __forceinline__ void pair_slabs(float4 n0, float4 n1, float4 n2, const RayInv& r, float3 ainv, float tBest,
                                           float& tn0, float& tn1, float& tw0, float& tw1) {
    const unsigned long long tmx = fma2(pk2(n0.x, n0.y), pk2(r.inv.x, r.inv.x), pk2(-r.ood.x, -r.ood.x));
    const unsigned long long tmy = fma2(pk2(n0.z, n0.w), pk2(r.inv.y, r.inv.y), pk2(-r.ood.y, -r.ood.y));
    const unsigned long long tmz = fma2(pk2(n1.x, n1.y), pk2(r.inv.z, r.inv.z), pk2(-r.ood.z, -r.ood.z));
    float nx0, nx1, ny0, ny1, nz0, nz1, fx0, fx1, fy0, fy1, fz0, fz1;
    upk2(fma2(pk2(-n1.z, -n1.w), pk2(ainv.x, ainv.x), tmx), nx0, nx1); upk2(fma2(pk2(n1.z, n1.w), pk2(ainv.x, ainv.x), tmx), fx0, fx1);
    upk2(fma2(pk2(-n2.x, -n2.y), pk2(ainv.y, ainv.y), tmy), ny0, ny1); upk2(fma2(pk2(n2.x, n2.y), pk2(ainv.y, ainv.y), tmy), fy0, fy1);
    upk2(fma2(pk2(-n2.z, -n2.w), pk2(ainv.z, ainv.z), tmz), nz0, nz1); upk2(fma2(pk2(n2.z, n2.w), pk2(ainv.z, ainv.z), tmz), fz0, fz1);
    tn0 = fmaxf(fmax3(nx0, ny0, nz0), 0.f); tn1 = fmaxf(fmax3(nx1, ny1, nz1), 0.f);
    const float tf0 = fminf(fmin3(fx0, fy0, fz0), tBest), tf1 = fminf(fmin3(fx1, fy1, fz1), tBest);
    // conservative: boxes are inflated at build time and the far bound is widened by a few ulps (see node_visit)
    upk2(mul2(pk2(tf0, tf1), pk2(1.0000005f, 1.0000005f)), tw0, tw1);
}
__device__ __forceinline__ bool node_visit_ch(const float4* __restrict__ nodes, uint32_t cur, const RayInv& r, float3 ainv, float tBest,
                                              uint32_t& nearc, uint32_t& farc, bool& both) {
    const float4* np = nodes + 4 * (size_t)cur;
#if BRT_NODE_LDG256
    // the 64-byte node in TWO 256-bit read-only loads (sm_100 LDG.E.256) instead of three 128-bit + one 64-bit: half the
    // load instructions and half the L1 requests of a visit whose lanes all sit at different nodes
    float4 n0, n1, n2; float2 n3; float pad0, pad1;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w) : "l"(np));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(pad0), "=f"(pad1) : "l"(np + 2));
#else
    const float4 n0 = ldg4(np), n1 = ldg4(np + 1), n2 = ldg4(np + 2);
    const float2 n3 = __ldg(reinterpret_cast<const float2*>(np + 3));
#endif
    const uint32_t c0 = __float_as_uint(n3.x), c1 = __float_as_uint(n3.y);
    float tn0, tn1, tw0, tw1;
    pair_slabs(n0, n1, n2, r, ainv, tBest, tn0, tn1, tw0, tw1);
    const bool h0 = tn0 <= tw0, h1 = tn1 <= tw1;
    const bool swap = tn1 < tn0;
    both = h0 & h1;
    nearc = both ? (swap ? c1 : c0) : (h0 ? c0 : c1);
    farc = swap ? c0 : c1;
    return h0 | h1;
}

#ifndef BRT_LEAF_WAIT_K
#define BRT_LEAF_WAIT_K 0
#endif
// Blocking traversal: runs one ray to completion.  HYBRID = false: the whole stack lives in shared memory (the launcher
// sized it from the tree depth: depth + 1 entries per thread) and the loop carries no local-memory path at all;
// HYBRID = true (trees deeper than SMEM_ONLY_MAX_DEPTH): SMEM_STACK entries in shared memory, the rest in a local array.
template <bool COUNT, bool SHADOW, bool HYBRID = true, int PRIMS = PRIMS_ANY, bool CH = false>
__device__ __forceinline__ Hit trace_bvh(const DevScene& sc, float3 O, float3 D, float tMin, float tMax, uint32_t self, Counters& cnt,
                                         uint32_t* sstack /* &smem[threadIdx.x] */, int sstride, unsigned aliveMask = 0xffffffffu) {
    Hit best; best.t = tMax; best.pid = PID_NONE;
    test_planes<COUNT, SHADOW>(sc, O, D, tMin, self, best, cnt);
    if (sc.nNodes == 0) return best;
    RayInv r = ray_inv(O, D);
    const float3 ainv = f3(fabsf(r.inv.x), fabsf(r.inv.y), fabsf(r.inv.z));
    uint32_t lstack[HYBRID ? LOCAL_STACK : 1];
    // (A variant with a uniform tail — far child stored unconditionally, "pop" as a select on a stack top that every lane
    // loads each iteration — removed the divergent pop block, 6-8 % of issued instructions at 6-7 lanes, but put a shared-memory
    // load on every iteration's critical path: C3 -2.8 %, C5 -0.9 %, C4 +3.3 % on a B200.  The branchy form stays.)
    // The bottom stack entry is a sentinel: popping it ends the loop, so a pop needs no emptiness test and the loop has ONE
    // back edge and one exit (the multi-exit form made ptxas copy best.t / best.pid between two register sets every iteration).
    // shared-memory stack through its 32-bit shared address, kept opaque so that ptxas holds it in a register instead of
    // re-deriving it from %tid at every push / pop (S2R + 4 instructions when it rematerialises)
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sstack);
    asm volatile("" : "+r"(sbase));
    const uint32_t sstep = (uint32_t)sstride * 4u;
    sts32(sbase, TRAV_DONE);
    int sp = 1;
    uint32_t cur = 0;
    do {
        if (COUNT && !SHADOW) {
            // lane attribution: which lanes of the warp execute this iteration, and doing what (profiles/*lane_attribution*)
            const unsigned act = __activemask();
            const unsigned leafM = __ballot_sync(act, (cur & LEAF_BIT) != 0u);
            if ((int)(threadIdx.x & 31u) == __ffs(act) - 1) {
                cnt.trIter++; cnt.trLanes += __popc(act); cnt.trAlive += __popc(aliveMask);
                cnt.trLeafIssue += leafM ? 1 : 0; cnt.trNodeIssue += (leafM != act) ? 1 : 0; cnt.trLeafLanes += __popc(leafM);
            }
        }
        bool pop = true;
#if BRT_LEAF_WAIT_K > 0
        // EXPERIMENT (rejected by measurement, profiles/r02c_leaf_wait_rejected.md; off by default): a lane that holds a leaf waits
        // until at least BRT_LEAF_WAIT_K lanes of the warp hold one (or no lane has node work left), so that the leaf block is
        // issued for fuller warps.  No speculation: the waiting lane does nothing, its ray's state is untouched.
        bool hold = false;
        {
            const unsigned actW = __activemask();
            const unsigned leafW = __ballot_sync(actW, (cur & LEAF_BIT) != 0u);
            hold = (cur & LEAF_BIT) != 0u && __popc(leafW) < BRT_LEAF_WAIT_K && leafW != actW;
        }
        if (hold) pop = false;
        else
#endif
        if (cur & LEAF_BIT) {
            test_prim<COUNT, SHADOW, PRIMS>(sc, cur & ~LEAF_BIT, O, D, tMin, self, best, cnt);
            if (SHADOW && best.pid != PID_NONE) break;
        } else {
            if (COUNT) { cnt.aabb += 2; cnt.nodeVisits++; }
            uint32_t nearc, farc; bool both;
            if (CH ? node_visit_ch(sc.cnodes, cur, r, ainv, best.t, nearc, farc, both) : node_visit(sc.nodes, cur, r, best.t, nearc, farc, both)) {
                if (both) {
                    if (!HYBRID || sp < SMEM_STACK) sts32(sbase + (uint32_t)sp * sstep, farc); else lstack[sp - SMEM_STACK] = farc;
                    sp++;
                }
                cur = nearc;
                pop = false;
            }
        }
        if (pop) {
            sp--;
            cur = (!HYBRID || sp < SMEM_STACK) ? lds32(sbase + (uint32_t)sp * sstep) : lstack[sp - SMEM_STACK];
        }
    } while (cur != TRAV_DONE);
    return best;
}

// ------------------------------------------------------------------------------------------- wide hierarchy
// N-wide (4 or 8) hierarchy collapsed from the binary tree (bvh.cu: k_wide_level).  A wide node (32 N bytes) is N / 2 PAIR RECORDS of
// 48 bytes, each holding two child slots in the box layout of a centre / half-extent binary node (node_visit_ch):
//   (cA.x, cB.x, cA.y, cB.y) (cA.z, cB.z, hA.x, hB.x) (hA.y, hB.y, hA.z, hB.z)
// followed by the N child refs (LEAF_BIT | pid, or the id of the child's wide node) as one contiguous uint32 array.  An empty slot
// has h.x = -inf (near = +inf, far = -inf: never hit).
// The boxes contain the binary tree's (padded) boxes and the slab arithmetic is pair_slabs — the same nine FFMA2 per TWO children
// as the binary visit — so the wide traversal is as conservative as the binary one: same hits, same images.
// Order: the builder puts a child into the slot whose index bits say on which side of the node's centre it lies (bit a set = the
// + side of axis a; N = 8: x, y, z; N = 4: the tree's two widest axes).  XOR-ing the slot index with the ray's direction-sign bits
// gives a front-to-back visiting order (Ylitie et al. 2017): the pair records are loaded in slot order (compile-time offsets from
// one base register) and the hit mask is permuted into traversal order afterwards (wide_order), so mask bit j is "j-th nearest
// slot" and the next child is simply the lowest set bit.  The stack holds one word per node that still has pending children (node id << 8 | pending
// mask) — one entry per LEVEL, not per far child.
// Measured on a B200 (profiles/r02_wide_hierarchy.md): the wide trees test the SAME number of boxes per ray as the binary tree at
// N = 4 (C5 55.4 vs 55.7, C3 19.6 vs 18.8) in half the node visits (C5 14.2 vs 27.9, N = 8: 9.8), but the walk is slower on every
// config but the Cornell box (N = 4: C3 -12 %, C5 -6 %; N = 8: C3 -20 %, C5 -23 %, C4 +0.6 %): only 4 % fewer warp-instructions
// than the binary kernel on C5, two dependent loads per level (node, then the picked slot's ref), sparser nodes (L2 hit 91 % vs 99 %).
// bvh_width = 0 (auto) therefore selects the binary tree; 4 / 8 stay selectable, with identical images.
// Hit mask in slot order -> hit mask in traversal order (bit j = slot j ^ c).  N = 4: one 64-bit table per octant code, looked up
// with a funnel shift (the 16 possible masks x 4 bits); N = 8: three conditional swaps of bit groups.
__device__ __forceinline__ unsigned long long wide4_table(uint32_t c) {
    // T_c = sum over the 16 masks m of perm_c(m) << 4m, perm_c(m) bit j = m bit (j ^ c); folded to a constant for a constant c
    unsigned long long t = 0;
#pragma unroll
    for (uint32_t m = 0; m < 16; m++) {
        uint32_t q = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; j++) q |= ((m >> (j ^ c)) & 1u) << j;
        t |= (unsigned long long)q << (4 * m);
    }
    return t;
}
template <int N>
__device__ __forceinline__ uint32_t wide_order(uint32_t m, uint32_t c, unsigned long long table4) {
    if (N == 4) return (uint32_t)(table4 >> (4u * m)) & 15u;
    if (c & 1u) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
    if (c & 2u) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
    if (c & 4u) m = ((m & 0x0Fu) << 4) | ((m >> 4) & 0x0Fu);
    return m;
}

template <int N, bool COUNT, bool SHADOW, int PRIMS = PRIMS_ANY>
__device__ __forceinline__ Hit trace_wide(const DevScene& sc, float3 O, float3 D, float tMin, float tMax, uint32_t self, Counters& cnt,
                                          uint32_t* sstack /* &smem[threadIdx.x] */, int sstride, unsigned aliveMask = 0xffffffffu) {
    static_assert(N == 4 || N == 8, "wide hierarchy: 4 or 8 children per node");
    constexpr uint32_t STRIDE = 32u * N, REFS = 24u * N;             // node bytes; byte offset of the N child refs (after N / 2 pair records of 48 B)
    Hit best; best.t = tMax; best.pid = PID_NONE;
    test_planes<COUNT, SHADOW>(sc, O, D, tMin, self, best, cnt);
    if (sc.nNodes == 0) return best;
    const RayInv r = ray_inv(O, D);
    const float3 ainv = f3(fabsf(r.inv.x), fabsf(r.inv.y), fabsf(r.inv.z));
    uint32_t c;
    if (N == 8) c = (D.x < 0.f ? 1u : 0u) | (D.y < 0.f ? 2u : 0u) | (D.z < 0.f ? 4u : 0u);
    else {
        const int a0 = sc.wideAxes & 3, a1 = (sc.wideAxes >> 2) & 3;
        const float d0 = a0 == 0 ? D.x : a0 == 1 ? D.y : D.z, d1 = a1 == 0 ? D.x : a1 == 1 ? D.y : D.z;
        c = (d0 < 0.f ? 1u : 0u) | (d1 < 0.f ? 2u : 0u);
    }
    unsigned long long table4 = 0;
    if (N == 4) {
        // the four tables are compile-time constants (wide4_table is evaluated per constant code); the ray keeps one of them
        const unsigned long long t0 = wide4_table(0), t1 = wide4_table(1), t2 = wide4_table(2), t3 = wide4_table(3);
        table4 = c == 0 ? t0 : c == 1 ? t1 : c == 2 ? t2 : t3;
    }
    const char* const base = reinterpret_cast<const char*>(sc.wnodes);
    uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sstack);
    asm volatile("" : "+r"(sbase));
    const uint32_t sstep = (uint32_t)sstride * 4u;
    int sp = 0;
    uint32_t g = 0u, node = 0u;
    bool visit = true;
    for (;;) {
        if (COUNT && !SHADOW) {
            const unsigned act = __activemask();
            const unsigned nodeM = __ballot_sync(act, visit);
            if ((int)(threadIdx.x & 31u) == __ffs(act) - 1) {
                cnt.trIter++; cnt.trLanes += __popc(act); cnt.trAlive += __popc(aliveMask); cnt.trNodeIssue += nodeM ? 1 : 0;
            }
        }
        if (visit) {
            const float4* np = reinterpret_cast<const float4*>(base + (size_t)node * STRIDE);
            uint32_t m = 0u;
            if (COUNT) cnt.nodeVisits++;
#pragma unroll
            for (int k = 0; k < N / 2; k++) {                          // pair records in SLOT order: compile-time offsets from one base
                const float4 n0 = ldg4(np + 3 * k), n1 = ldg4(np + 3 * k + 1), n2 = ldg4(np + 3 * k + 2);
                if (COUNT) cnt.aabb += (n1.z >= 0.f ? 1 : 0) + (n1.w >= 0.f ? 1 : 0);      // occupied slots
                float tn0, tn1, tw0, tw1;
                pair_slabs(n0, n1, n2, r, ainv, best.t, tn0, tn1, tw0, tw1);
                if (tn0 <= tw0) m |= 1u << (2 * k);
                if (tn1 <= tw1) m |= 2u << (2 * k);
            }
            g = (node << 8) | wide_order<N>(m, c, table4);
        }
        if ((g & 0xFFu) == 0u) {
            if (sp == 0) break;
            sp--;
            g = lds32(sbase + (uint32_t)sp * sstep);
        }
        const uint32_t slot = ((uint32_t)__ffs((int)g) - 1u) ^ c;          // the nearest pending slot (traversal order -> slot)
        g &= g - 1u;
        const uint32_t ref = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(g >> 8) * STRIDE + REFS) + slot);
        const bool leaf = (ref & LEAF_BIT) != 0u;
        if (COUNT && !SHADOW) {
            const unsigned act = __activemask();
            const unsigned leafM = __ballot_sync(act, leaf);
            if ((int)(threadIdx.x & 31u) == __ffs(act) - 1) { cnt.trLeafIssue += leafM ? 1 : 0; cnt.trLeafLanes += __popc(leafM); }
        }
        if (leaf) {
            test_prim<COUNT, SHADOW, PRIMS>(sc, ref & ~LEAF_BIT, O, D, tMin, self, best, cnt);
            if (SHADOW && best.pid != PID_NONE) break;
            visit = false;
        } else {
            if (g & 0xFFu) { sts32(sbase + (uint32_t)sp * sstep, g); sp++; }
            node = ref;
            visit = true;
        }
    }
    return best;
}

// ------------------------------------------------------------------------------------------- float64 primary hits
// Mixed precision by design: fp32 traversal SELECTS the primitive; for primary rays the selected primitive is then
// re-evaluated in float64 from the float64 camera ray (one primitive, ~60 DFMA per camera sample), which puts t, P and N
// within rounding of the float64 reference even at silhouettes and grazing angles where fp32 is ill-conditioned.
struct D3 { double x, y, z; };
// Explicitly rounded, never contracted, in the reference's operation order (math.js:11-19): given the same inputs these
// produce the same bits as JavaScript, so a primary hit's t equals the float64 reference's t.
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(__dmul_rn(a.x, s), __dmul_rn(a.y, s), __dmul_rn(a.z, s)); }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return d3(__ddiv_rn(a.x, s), __ddiv_rn(a.y, s), __ddiv_rn(a.z, s)); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z)); }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3(__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
              __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__device__ __forceinline__ double js_max(double a, double b) { return (a != a || b != b) ? __dadd_rn(a, b) : (a > b ? a : b); }   // Math.max: NaN if either is
__device__ __forceinline__ double js_min(double a, double b) { return (a != a || b != b) ? __dadd_rn(a, b) : (a < b ? a : b); }
__device__ __forceinline__ D3 normalize0(D3 a) { double l = sqrt(dot(a, a)); return l > 0 ? a / l : d3(0, 0, 0); }
__device__ __forceinline__ D3 ldd3(const double* p) { return d3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ float3 tof3(D3 a) { return f3((float)a.x, (float)a.y, (float)a.z); }

struct Surface {
    float3 P, N;        // hit point, shading normal after setFaceNormal (math.js:55-58)
    bool front;
    int objId, matId, triId;
};

// Returns false when the float64 evaluation rejects the primitive fp32 selected (a silhouette flip): the caller keeps fp32.
static __device__ __noinline__ bool refine_primary(const DevScene& sc, uint32_t pid, D3 O, D3 D, double& tOut, Surface& s) {
    const uint32_t ty = pid_type(pid);
    const double* q = sc.prim64 + 9 * (size_t)meta_index(sc, pid);
    const double tMin = 0.001;
    double t; D3 n;
    if (ty == PT_SPHERE) {                                                     // geometry.js:15-34
        D3 c = ldd3(q); double r = __ldg(q + 3);
        D3 oc = O - c;
        double a = dot(D, D), hb = dot(oc, D), cc = __dsub_rn(dot(oc, oc), __dmul_rn(r, r));
        double disc = __dsub_rn(__dmul_rn(hb, hb), __dmul_rn(a, cc));
        if (disc < 0) return false;
        double sq = sqrt(disc);
        t = __ddiv_rn(__dsub_rn(-hb, sq), a);
        if (t < tMin) { t = __ddiv_rn(__dadd_rn(-hb, sq), a); if (!(t >= tMin)) return false; }
        D3 P = O + D * t;
        n = (P - c) / r;
    } else if (ty == PT_PLANE) {                                               // geometry.js:56-61
        n = ldd3(q);
        double den = dot(n, D);
        if (fabs(den) < 1e-6) return false;
        t = __ddiv_rn(dot(ldd3(q + 3) - O, n), den);
        if (!(t >= tMin)) return false;
    } else if (ty == PT_BOX) {                                                 // geometry.js:85-126 incl. the |p - face| < 1e-6 face rule
        D3 mn = ldd3(q), mx = ldd3(q + 3);
        double t0 = __ddiv_rn(__dsub_rn(mn.x, O.x), D.x), t1 = __ddiv_rn(__dsub_rn(mx.x, O.x), D.x);
        if (t0 > t1) { double w = t0; t0 = t1; t1 = w; }
        double y0 = __ddiv_rn(__dsub_rn(mn.y, O.y), D.y), y1 = __ddiv_rn(__dsub_rn(mx.y, O.y), D.y);
        if (y0 > y1) { double w = y0; y0 = y1; y1 = w; }
        if (t0 > y1 || y0 > t1) return false;
        // Math.max / Math.min PROPAGATE NaN (fmax / fmin drop it): a ray parallel to a slab whose face passes through the ray
        // origin gives 0 / 0 = NaN, and the reference then leaves through `t0 > tMin ? t0 : t1` with t1 (pinned against the
        // reference's own Box.hit: tests/golden/reference_aov_vectors.json, case axis_parallel_rays_and_zero_over_zero)
        t0 = js_max(t0, y0); t1 = js_min(t1, y1);
        double z0 = __ddiv_rn(__dsub_rn(mn.z, O.z), D.z), z1 = __ddiv_rn(__dsub_rn(mx.z, O.z), D.z);
        if (z0 > z1) { double w = z0; z0 = z1; z1 = w; }
        if (t0 > z1 || z0 > t1) return false;
        t0 = js_max(t0, z0); t1 = js_min(t1, z1);
        t = t0 > tMin ? t0 : t1;
        if (!(t >= tMin)) return false;
        D3 P = O + D * t;
        const double eps = 1e-6;
        if (fabs(P.x - mn.x) < eps) n = d3(-1, 0, 0);
        else if (fabs(P.x - mx.x) < eps) n = d3(1, 0, 0);
        else if (fabs(P.y - mn.y) < eps) n = d3(0, -1, 0);
        else if (fabs(P.y - mx.y) < eps) n = d3(0, 1, 0);
        else if (fabs(P.z - mn.z) < eps) n = d3(0, 0, -1);
        else n = d3(0, 0, 1);
    } else {                                                                   // geometry.js:148-175
        D3 v0 = ldd3(q), e1 = ldd3(q + 3) - v0, e2 = ldd3(q + 6) - v0;
        D3 h = cross(D, e2);
        double a = dot(e1, h);
        if (fabs(a) < 0.0001) return false;
        double f = __ddiv_rn(1.0, a);
        D3 sv = O - v0;
        double u = __dmul_rn(f, dot(sv, h));
        if (u < 0.0 || u > 1.0) return false;
        D3 qq = cross(sv, e1);
        double v = __dmul_rn(f, dot(D, qq));
        if (v < 0.0 || __dadd_rn(u, v) > 1.0) return false;
        t = __dmul_rn(f, dot(e2, qq));
        if (!(t >= tMin)) return false;
        n = normalize0(cross(e1, e2));
    }
    D3 P = O + D * t;
    bool front = dot(D, n) < 0;
    s.P = tof3(P);
    s.N = front ? tof3(n) : tof3(d3(-n.x, -n.y, -n.z));
    s.front = front;
    tOut = t;
    return true;
}

// Primary visibility decided in float64 (the PRECISE instantiations: sampler = reference, the fp32 AOV kernel).  fp32 only
// PROPOSES: the hierarchy is walked with fp32 slab tests on padded boxes (conservative: a box is ~100x wider than the rounding
// of the ray and of the slab arithmetic), and EVERY primitive whose leaf is reached is evaluated by refine_primary — float64,
// the reference's operation order, its root selection and thresholds — from the float64 camera ray.  The closest float64 hit
// wins, exact ties go by the reference's loop order (tie_wins), and the fp32 cull distance is the float64 best widened by a
// few ulps so a candidate at an equal distance is still visited.  The chosen primitive, its t, point and normal are therefore
// the float64 reference's own (IDs bit-exact; tests/test_gpu_parity.py asserts zero mismatches against the brute-force
// float64 kernel and the oracle), at the price of ~60 double-precision operations per candidate leaf of a PRIMARY ray.
template <bool USE_BVH, bool HYBRID>
static __device__ __noinline__ bool trace_primary64(const DevScene& sc, D3 O64, D3 D64, uint32_t* sstack, int sstride, uint32_t& pidOut,
                                                    double& tOut, Surface& sfOut) {
    double bestT = CUDART_INF; uint32_t bestPid = PID_NONE;
    auto candidate = [&](uint32_t pid) {
        Surface s; double t;
        if (!refine_primary(sc, pid, O64, D64, t, s)) return;
        if (t < bestT || (t == bestT && bestPid != PID_NONE && tie_wins(sc, pid, bestPid))) { bestT = t; bestPid = pid; sfOut = s; }
    };
    for (int i = 0; i < sc.nPln; i++) candidate(make_pid(PT_PLANE, i));
    if (!USE_BVH || sc.nNodes == 0) {                                            // the reference's own linear loops, in float64
        for (int i = 0; i < sc.nSph; i++) candidate(make_pid(PT_SPHERE, i));
        for (int i = 0; i < sc.nBox; i++) candidate(make_pid(PT_BOX, i));
        for (int i = 0; i < sc.nTri; i++) candidate(make_pid(PT_TRI, i));
    } else {
        const float3 O = tof3(O64), D = tof3(D64);
        RayInv r = ray_inv(O, D);
        uint32_t lstack[HYBRID ? LOCAL_STACK : 1];
        int sp = 0;
        uint32_t cur = 0;
        for (;;) {
            if (cur & LEAF_BIT) candidate(cur & ~LEAF_BIT);
            else {
                // cull distance: the float64 best, rounded up and widened (equal-distance candidates must still be reached)
                const float tCull = bestT < 1e30 ? __fmul_rn(__double2float_ru(bestT), 1.000002f) + 1e-30f : CUDART_INF_F;
                uint32_t nearc, farc; bool both;
                if (node_visit(sc.nodes, cur, r, tCull, nearc, farc, both)) {
                    if (both) {
                        if (!HYBRID || sp < SMEM_STACK) sstack[sp * sstride] = farc; else lstack[sp - SMEM_STACK] = farc;
                        sp++;
                    }
                    cur = nearc;
                    continue;
                }
            }
            if (sp == 0) break;
            sp--;
            cur = (!HYBRID || sp < SMEM_STACK) ? sstack[sp * sstride] : lstack[sp - SMEM_STACK];
        }
    }
    if (bestPid == PID_NONE) return false;
    int4 m = __ldg(&sc.meta[meta_index(sc, bestPid)]);
    sfOut.objId = m.x; sfOut.matId = m.y; sfOut.triId = m.z;
    pidOut = bestPid; tOut = bestT;
    return true;
}

// ------------------------------------------------------------------------------------------- surface frame
__device__ __forceinline__ Surface make_surface(const DevScene& sc, const Hit& h, float3 O, float3 D, uint32_t self) {
    Surface s;
    s.P = madd(D, h.t, O);                                                     // Ray.at (math.js:41)
    uint32_t ty = pid_type(h.pid), ix = pid_index(h.pid);
    int4 m = __ldg(&sc.meta[meta_index(sc, h.pid)]);
    s.objId = m.x; s.matId = m.y; s.triId = m.z;
    float3 n;
    if (ty == PT_SPHERE) {
        float4 sp = ldg4(sc.sph + ix);
        n = (s.P - f3(sp.x, sp.y, sp.z)) * rcpf(sp.w);                    // geometry.js:34 (negative radius flips)
    } else if (ty == PT_PLANE) {
        n = xyz(ldg4(sc.pln + 2 * ix));
    } else if (ty == PT_BOX) {
        float t; int f = 0;                                                    // the slab plane that produced t (re-derived: keeps `face` out of the traversal state)
        hit_box(ldg4(sc.box + 2 * ix), ldg4(sc.box + 2 * ix + 1), O, D, 0.001f, h.pid == self, t, f);
        float sgn = (f & 1) ? 1.f : -1.f;
        n = f3((f >> 1) == 0 ? sgn : 0.f, (f >> 1) == 1 ? sgn : 0.f, (f >> 1) == 2 ? sgn : 0.f);
    } else {
        n = normalize0(cross(xyz(ldg4(sc.tri + 3 * ix + 1)), xyz(ldg4(sc.tri + 3 * ix + 2))));   // geometry.js:143-145
    }
    s.front = dot(D, n) < 0.f;
    s.N = s.front ? n : -n;
    return s;
}

// ------------------------------------------------------------------------------------------- backgrounds (world.js:35-110)
// (explicit fmaf / __fmul_rn throughout: see the note on pinned numerics at the top of this file)
__device__ __forceinline__ float perlin_fade(float t) { return __fmul_rn(__fmul_rn(__fmul_rn(t, t), t), fmaf(t, fmaf(t, 6.f, -15.f), 10.f)); }      // noise.js:20
__device__ __forceinline__ float perlin_grad(int hash, float x, float y, float z) {                               // noise.js:22-27
    int h = hash & 15;
    float u = h < 8 ? x : y;
    float v = h < 4 ? y : (h == 12 || h == 14) ? x : z;
    return __fadd_rn((h & 1) == 0 ? u : -u, (h & 2) == 0 ? v : -v);
}
__device__ __forceinline__ float lerpf(float t, float a, float b) { return fmaf(t, __fsub_rn(b, a), a); }          // noise.js:21
__device__ inline float perlin_noise(const unsigned char* __restrict__ p, float x, float y, float z) {             // noise.js:29-61
    float flx = floorf(x), fly = floorf(y), flz = floorf(z);
    int X = ((int)flx) & 255, Y = ((int)fly) & 255, Z = ((int)flz) & 255;
    float fx = __fsub_rn(x, flx), fy = __fsub_rn(y, fly), fz = __fsub_rn(z, flz);
    float u = perlin_fade(fx), v = perlin_fade(fy), w = perlin_fade(fz);
    int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
    int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
    return lerpf(w,
        lerpf(v, lerpf(u, perlin_grad(p[AA], fx, fy, fz), perlin_grad(p[BA], fx - 1, fy, fz)),
                 lerpf(u, perlin_grad(p[AB], fx, fy - 1, fz), perlin_grad(p[BB], fx - 1, fy - 1, fz))),
        lerpf(v, lerpf(u, perlin_grad(p[AA + 1], fx, fy, fz - 1), perlin_grad(p[BA + 1], fx - 1, fy, fz - 1)),
                 lerpf(u, perlin_grad(p[AB + 1], fx, fy - 1, fz - 1), perlin_grad(p[BB + 1], fx - 1, fy - 1, fz - 1))));
}
// a*x + b*y + c*z + d*w + e*v with explicit fused steps
__device__ __forceinline__ float mix5(float a, float x, float b, float y, float c, float z, float d, float w, float e, float v) {
    return fmaf(e, v, fmaf(d, w, fmaf(c, z, fmaf(b, y, __fmul_rn(a, x)))));
}
__device__ inline float3 background(const DevScene& sc, float3 D) {
    float3 dir = normalize0(D);
    float I = sc.skyIntensity;
    switch (sc.bgKind) {
    case 1:                                                                                                         // world.js:42-44
        return f3(__fmul_rn(sc.bgR, I), __fmul_rn(sc.bgG, I), __fmul_rn(sc.bgB, I));
    case 2: {                                                                                                       // world.js:74-110
        const float il = 1.1952286093343936f;        // 1/|(-0.3,0.6,-0.5)|
        float sunDot = fmaxf(0.f, __fmul_rn(fmaf(dir.z, -0.5f, fmaf(dir.y, 0.6f, __fmul_rn(dir.x, -0.3f))), il));
        float sunMask = sunDot > 0.96f ? 20.f : 0.f;
        float corona = fmaxf(0.f, __fmul_rn(__fsub_rn(sunDot, 0.8f), 5.f));
        float c3 = __fmul_rn(__fmul_rn(corona, corona), 3.f);
        float y = dir.y;
        float sky = __fmul_rn(fmaxf(0.f, fmaf(y, 0.5f, 0.5f)), 2.f);
        float ground = fmaxf(0.f, __fmul_rn(y, -0.3f));
        float sc1 = fmaxf(0.f, __fsub_rn(1.f, fabsf(y)));
        float scat = __fmul_rn(__fmul_rn(sc1, sc1), 0.3f);
        return f3(__fmul_rn(mix5(0.3f, sky, 0.2f, ground, 0.8f, scat, 1.0f, sunMask, 1.0f, c3), I),
                  __fmul_rn(mix5(0.5f, sky, 0.15f, ground, 0.9f, scat, 0.95f, sunMask, 0.8f, c3), I),
                  __fmul_rn(mix5(0.8f, sky, 0.1f, ground, 1.0f, scat, 0.8f, sunMask, 0.6f, c3), I));
    }
    case 3: {                                                                                                       // world.js:46-72
        const float il = 0.95782628522115137f;       // 1/|(0.3,0.6,0.8)|
        float sunDot = fmaxf(0.f, __fmul_rn(fmaf(dir.z, 0.8f, fmaf(dir.y, 0.6f, __fmul_rn(dir.x, 0.3f))), il));
        float s = sunDot;                             // pow(x, 512) by nine exact squarings
#pragma unroll
        for (int k = 0; k < 9; k++) s = __fmul_rn(s, s);
        float sun = __fmul_rn(s, 10.f);
        float hb = __fmul_rn(fmaxf(0.f, dir.y), 0.8f);
        float glow = __fmul_rn(expf(__fmul_rn(fabsf(dir.y), -4.f)), 0.3f);
        float ground = fmaxf(0.f, __fmul_rn(dir.y, -0.5f));
        float cloud = fmaxf(0.f, fmaf(perlin_noise(sc.perm, __fmul_rn(dir.x, 10.f), fmaf(dir.y, 3.f, 2.f), __fmul_rn(dir.z, 10.f)), 0.8f, 0.2f));
        float cl = __fmul_rn(__fmul_rn(cloud, fmaxf(0.f, dir.y)), 0.5f);
        return f3(__fmul_rn(mix5(0.4f, hb, 1.0f, glow, 0.1f, ground, 1.0f, sun, 0.9f, cl), I),
                  __fmul_rn(mix5(0.7f, hb, 0.8f, glow, 0.15f, ground, 0.95f, sun, 0.9f, cl), I),
                  __fmul_rn(mix5(1.0f, hb, 0.6f, glow, 0.1f, ground, 0.8f, sun, 1.0f, cl), I));
    }
    default: {                                                                                                      // world.js:35-40
        float t = __fmul_rn(0.5f, __fadd_rn(dir.y, 1.0f));
        // (1-t)*(1,1,1) + t*(0.5,0.7,1.0)
        return f3(__fmul_rn(fmaf(t, 0.5f, __fsub_rn(1.0f, t)), I), __fmul_rn(fmaf(t, 0.7f, __fsub_rn(1.0f, t)), I), __fmul_rn(fmaf(t, 1.0f, __fsub_rn(1.0f, t)), I));
    }
    }
}

// ------------------------------------------------------------------------------------------- surface textures (textures.js)
// value(u, v, p) of every reference texture depends on the hit point only.
__device__ inline float3 texture_value(const DevScene& sc, int ti, float3 P) {
    float4 a = ldg4(sc.tex + 2 * ti), b = ldg4(sc.tex + 2 * ti + 1);
    const int kind = (int)a.w;
    const float scale = b.w;
    const unsigned char* perm = sc.texPerm + 512 * ti;
    switch (kind) {
    case 1: {                                                                                                        // :33-36
        float sines = __fmul_rn(__fmul_rn(sinf(__fmul_rn(scale, P.x)), sinf(__fmul_rn(scale, P.y))), sinf(__fmul_rn(scale, P.z)));
        return sines < 0.f ? f3(a.x, a.y, a.z) : f3(b.x, b.y, b.z);
    }
    case 2: {                                                                                                        // :47-50
        float n = perlin_noise(perm, __fmul_rn(P.x, scale), __fmul_rn(P.y, scale), __fmul_rn(P.z, scale));
        float v = __fmul_rn(0.5f, __fadd_rn(1.f, n));
        return f3(v, v, v);
    }
    case 3: {                                                                                                        // :61-65, noise.js:63-75
        float accum = 0.f, weight = 1.f;
        float3 q = P * scale;
        for (int i = 0; i < 7; i++) {
            accum = fmaf(weight, perlin_noise(perm, q.x, q.y, q.z), accum);
            weight = __fmul_rn(weight, 0.5f);
            q = q * 2.f;
        }
        float m = __fmul_rn(0.5f, __fadd_rn(1.f, sinf(fmaf(10.f, fabsf(accum), __fmul_rn(scale, P.z)))));
        float w = __fsub_rn(1.f, m);
        return f3(fmaf(0.6f, w, __fmul_rn(0.9f, m)), fmaf(0.4f, w, __fmul_rn(0.8f, m)), fmaf(0.3f, w, __fmul_rn(0.7f, m)));
    }
    case 4: {                                                                                                        // :77-82
        float s20 = __fmul_rn(scale, 20.f);
        float grain = perlin_noise(perm, __fmul_rn(P.x, s20), __fmul_rn(P.y, s20), __fmul_rn(P.z, s20));
        float rings = sinf(fmaf(grain, 10.f, __fmul_rn(scale, sqrtf(fmaf(P.x, P.x, __fmul_rn(P.z, P.z))))));
        float m = __fmul_rn(0.5f, __fadd_rn(1.f, rings));
        float w = __fsub_rn(1.f, m);
        return f3(fmaf(0.4f, w, __fmul_rn(0.8f, m)), fmaf(0.2f, w, __fmul_rn(0.5f, m)), fmaf(0.1f, w, __fmul_rn(0.2f, m)));
    }
    default: return f3(a.x, a.y, a.z);                                                                               // :21
    }
}

// ------------------------------------------------------------------------------------------- sampling
// math.js:22-31 by direct inversion (identical distributions; used by BRT_SAMPLER_FAST).
// sin / cos of 2*pi*u for the FAST sampler's random angles: the MUFU approximations (absolute error ~2^-21 on [0, 2 pi)) in
// place of the ~40-instruction exact sincospif — these angles are uniform random numbers, so the error is far below the
// Monte-Carlo noise floor and unbiased; the reference sampler keeps sincospif (compared sample for sample with the oracle).
__device__ __forceinline__ void fast_sincos2pi(float u, float* sn, float* cs) {
#ifndef BRT_EXACT_SINCOS
    const float x = __fmul_rn(6.283185307179586f, u);
    *sn = __sinf(x); *cs = __cosf(x);                      // +3 % on C3 / C4, +4 % on C2 (B200, tools/ab.py)
#else
    sincospif(__fmul_rn(2.f, u), sn, cs);
#endif
}
__device__ __forceinline__ float3 uniform_sphere(float u0, float u1) {
    float z = fmaf(-2.f, u0, 1.f);
    float r = sqrtf(fmaxf(0.f, fmaf(-z, z, 1.f)));
    float sn, cs;
    fast_sincos2pi(u1, &sn, &cs);
    return f3(__fmul_rn(r, cs), __fmul_rn(r, sn), z);
}

}  // namespace brt
