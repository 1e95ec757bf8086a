// Path tracer: a persistent, warp-local WAVEFRONT inside one kernel.  Every lane owns one pixel and keeps PT_K samples
// of it in flight; the 32*PT_K path slots of a warp live in shared memory.  The warp alternates between a shade phase
// (every lane shades / scatters / regenerates its own slots — all lanes busy) and an extend phase (the warp's rays are a
// queue in shared memory; a lane that finishes a ray fetches the next one, so the BVH loop stays full although rays need
// very different numbers of steps).  No global-memory queues, no launches per bounce, no block-wide barriers.
// Replaces the three nested loops of RayTracer.render (ray-tracer.js:189-206) and the recursion of
// rayColor (ray-tracer.js:102-123) — written iteratively: sum += beta ⊙ E; beta ⊙= attenuation.
#include "brt_device.cuh"
#include "brt_kernels.h"

namespace brt {

// ------------------------------------------------------------------------------------------- camera (camera.js:38-51)
// The lens / pixel sample of one camera ray: s, t (getAntiAliasSample, ray-tracer.js:125-149) and the unit-disk point.
struct CamSample { float s, t, dx, dy; };

// float64 ray exactly as camera.js:38-51 forms it (operation order kept), from fp32-exact sample values.
__device__ __forceinline__ void camera_ray64(const DevCamera& c, int W, int H, int aaMode, int col, int jUp, CamSample cs, D3& O, D3& D) {
    double u, v;
    if (aaMode == 1) { u = __ddiv_rn(__dadd_rn((double)col, (double)cs.s), (double)W); v = __ddiv_rn(__dadd_rn((double)jUp, (double)cs.t), (double)H); }
    else if (aaMode == 2) {
        u = __ddiv_rn(__dadd_rn((double)col + 0.5, __dmul_rn((double)cs.s, 0.5)), (double)W);
        v = __ddiv_rn(__dadd_rn((double)jUp + 0.5, __dmul_rn((double)cs.t, 0.5)), (double)H);
    } else { u = __ddiv_rn((double)col + 0.5, (double)W); v = __ddiv_rn((double)jUp + 0.5, (double)H); }
    double rx = __dmul_rn((double)cs.dx, c.lensRadius), ry = __dmul_rn((double)cs.dy, c.lensRadius);
    D3 cu = d3(c.cu[0], c.cu[1], c.cu[2]), cv = d3(c.cv[0], c.cv[1], c.cv[2]);
    D3 ll = d3(c.ll[0], c.ll[1], c.ll[2]), hh = d3(c.h[0], c.h[1], c.h[2]), vv = d3(c.v[0], c.v[1], c.v[2]);
    if (c.type == 1) {                                          // camera.js:39-43
        O = d3(c.o[0], c.o[1], c.o[2]) + cu * rx + cv * ry;
        D = normalize0(ll + hh * u + vv * v - O + d3(c.cw[0], c.cw[1], c.cw[2]) * -1.0);
    } else {                                                    // camera.js:44-49
        O = d3(c.o[0], c.o[1], c.o[2]) + (cu * rx + cv * ry);
        D = ll + hh * u + vv * v - O;
    }
}

// fp32 form of the same ray (BRT_SAMPLER_FAST render path: jittered, lens-offset camera samples have no float64 reference
// to match bit for bit; LLC - origin is formed in float64 on the host side of this call and rounded once).
__device__ __forceinline__ void camera_ray32(const DevCamera& c, int W, int H, int aaMode, int col, int jUp, CamSample cs, float3& O, float3& D) {
    float u, v;
    if (aaMode == 1) { u = __fdiv_rn((float)col + cs.s, (float)W); v = __fdiv_rn((float)jUp + cs.t, (float)H); }
    else if (aaMode == 2) { u = __fdiv_rn(fmaf(cs.s, 0.5f, (float)col + 0.5f), (float)W); v = __fdiv_rn(fmaf(cs.t, 0.5f, (float)jUp + 0.5f), (float)H); }
    else { u = __fdiv_rn((float)col + 0.5f, (float)W); v = __fdiv_rn((float)jUp + 0.5f, (float)H); }
    float rx = __fmul_rn(cs.dx, (float)c.lensRadius), ry = __fmul_rn(cs.dy, (float)c.lensRadius);
    float3 off = madd(f3((float)c.cv[0], (float)c.cv[1], (float)c.cv[2]), ry, f3((float)c.cu[0], (float)c.cu[1], (float)c.cu[2]) * rx);
    O = f3((float)c.o[0], (float)c.o[1], (float)c.o[2]) + off;
    // D = (LLC - origin) + u*H + v*V - off
    float3 llo = f3((float)(c.ll[0] - c.o[0]), (float)(c.ll[1] - c.o[1]), (float)(c.ll[2] - c.o[2]));
    D = madd(f3((float)c.v[0], (float)c.v[1], (float)c.v[2]), v, madd(f3((float)c.h[0], (float)c.h[1], (float)c.h[2]), u, llo)) - off;
    if (c.type == 1) D = normalize0(D - f3((float)c.cw[0], (float)c.cw[1], (float)c.cw[2]));
}

template <int SAMPLER>
__device__ __forceinline__ CamSample camera_sample(const PTParams& p, uint32_t pix, uint32_t s, RngSeq& rng) {
    CamSample cs; cs.s = 0.f; cs.t = 0.f;
    float a0 = 0.f, a1 = 0.f;
    if (SAMPLER == 0) {
        uint4 r = philox_fast(pix, s, 0u, PHILOX_TAG, p.seedLo, p.seedHi);
        a0 = u01(r.x); a1 = u01(r.y);
        float rr = sqrtf(u01(r.z)), sn, cs_;                   // unit disk by inversion (math.js:27-31 distribution)
        sincospif(__fmul_rn(2.f, u01(r.w)), &sn, &cs_);
        cs.dx = __fmul_rn(rr, cs_); cs.dy = __fmul_rn(rr, sn);
    } else {
        rng.init(pix, s, p.seedLo, p.seedHi);
        if (p.aaMode == 1 || p.aaMode == 2) { a0 = rng.next(); a1 = rng.next(); }
        do { cs.dx = rng.next() * 2.f - 1.f; cs.dy = rng.next() * 2.f - 1.f; } while (fmaf(cs.dx, cs.dx, __fmul_rn(cs.dy, cs.dy)) >= 1.0f);   // math.js:29
    }
    if (p.aaMode == 1) { cs.s = a0; cs.t = a1; }
    else if (p.aaMode == 2) {                                    // stochastic: disk of radius 0.5 about the pixel centre
        float sr = sqrtf(a0), sn, c2;
        sincospif(__fmul_rn(2.f, a1), &sn, &c2);
        cs.s = __fmul_rn(sr, c2); cs.t = __fmul_rn(sr, sn);
    }
    return cs;
}

// ------------------------------------------------------------------------------------------- materials (materials.js)
// Returns false when the path ends here (emissive, absorbed metal).  `att` multiplies the throughput.
template <int SAMPLER>
__device__ __forceinline__ bool scatter(const PTParams& p, int matWord, float4 m, const Surface& sf, float3 Din, uint32_t pix,
                                        uint32_t s, int bounce, RngSeq& rng, float3& Dout, float3& att) {
    const int matType = matWord & 255, tex = matWord >> 8;           // 1-based texture index above the type (materials.js:99-126)
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (SAMPLER == 0) {
        uint4 r = philox_fast(pix, s, (uint32_t)(bounce + 1), PHILOX_TAG, p.seedLo, p.seedHi);
        u0 = u01(r.x); u1 = u01(r.y); u2 = u01(r.z);
    }
    if (matType == 0) {                                                       // Lambertian (materials.js:20-25)
        float3 unit;
        if (SAMPLER == 0) unit = uniform_sphere(u0, u1);
        else {
            float3 q;
            do { q = f3(rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f); } while (dot(q, q) >= 1.0f);
            unit = normalize0(q);
        }
        Dout = sf.N + unit;
        att = tex ? texture_value(p.sc, tex - 1, sf.P) : f3(m.x, m.y, m.z);
        return true;
    }
    if (matType == 1) {                                                       // Metal (materials.js:36-41)
        float3 refl = reflect(normalize0(Din), sf.N);
        float3 ball;
        if (SAMPLER == 0) ball = uniform_sphere(u0, u1) * cbrtf(u2);
        else { do { ball = f3(rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f); } while (dot(ball, ball) >= 1.0f); }
        Dout = madd(ball, m.w, refl);
        att = tex ? texture_value(p.sc, tex - 1, sf.P) : f3(m.x, m.y, m.z);
        return dot(Dout, sf.N) > 0.f;
    }
    if (matType == 2) {                                                       // Dielectric (materials.js:51-83)
        float ratio = sf.front ? (1.0f / m.w) : m.w;
        float3 ud = normalize0(Din);
        float cosT = fminf(-dot(ud, sf.N), 1.0f);
        float sinT = sqrtf(fmaxf(0.f, fmaf(-cosT, cosT, 1.0f)));
        bool cannot = __fmul_rn(ratio, sinT) > 1.0f;
        bool refl = cannot;
        if (!cannot) {                                                        // the uniform is drawn only here (:62)
            float r0 = __fdiv_rn(1.f - ratio, 1.f + ratio); r0 = __fmul_rn(r0, r0);
            float c1 = 1.f - cosT, c2 = __fmul_rn(c1, c1);
            float R = fmaf(1.f - r0, __fmul_rn(__fmul_rn(c2, c2), c1), r0);
            float xi = SAMPLER == 0 ? u0 : rng.next();
            refl = R > xi;
        }
        if (refl) Dout = reflect(ud, sf.N);
        else {
            float3 perp = madd(sf.N, cosT, ud) * ratio;
            Dout = madd(sf.N, -sqrtf(fabsf(1.0f - dot(perp, perp))), perp);
        }
        att = f3(1.f, 1.f, 1.f);
        return true;
    }
    return false;                                                             // Emissive (materials.js:94)
}

template <bool USE_BVH, bool COUNT, bool SHADOW, bool HYBRID = true>
__device__ __forceinline__ Hit trace(const DevScene& sc, float3 O, float3 D, float tMax, uint32_t self, Counters& cnt,
                                     uint32_t* sstack, int sstride) {
    if (COUNT && !SHADOW) cnt.rays++;
    if (USE_BVH) return trace_bvh<COUNT, SHADOW, HYBRID>(sc, O, D, 0.001f, tMax, self, cnt, sstack, sstride);
    return trace_brute<COUNT, SHADOW>(sc, O, D, 0.001f, tMax, self, cnt);
}

// ------------------------------------------------------------------------------------------- the warp-local wavefront kernel
// Block = 128 threads = a 16x8 pixel tile; a warp = an 8x4 sub-tile (coherent primary rays).
// Slot fields (shared memory, [field][slot] per warp, slot = k*32 + owner lane):
// F_STATE packs (camera sample index << 8) | (depth + 2): 0 = dead (no samples left), 1 = needs a camera ray, >= 2 = a path
// at `depth` waiting for / holding a hit.  F_RNGPOS exists only for the sequential (reference) sampler.
enum SlotField : int { F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_SELF, F_T, F_PID, F_BX, F_BY, F_BZ, F_STATE, F_RNGPOS };
static_assert(F_RNGPOS == PT_SLOT_WORDS, "slot layout");
constexpr int DEPTH_NEED_RAY = -1, DEPTH_DEAD = -2;
__host__ __device__ constexpr int slot_words(int sampler) { return PT_SLOT_WORDS + (sampler == 1 ? 1 : 0); }

template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT, int K>
__global__ void __launch_bounds__(PT_BLOCK, PT_MIN_BLOCKS) k_pathtrace_wave(const __grid_constant__ PTParams p) {
    extern __shared__ uint32_t smem[];
    constexpr int NS = 32 * K;                                        // path slots per warp
    constexpr int NW = slot_words(SAMPLER);
    constexpr bool PRECISE = SAMPLER == 1;                            // see k_pathtrace_mega
    constexpr int WARP_WORDS = NW * NS + 32 * SMEM_STACK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = col < p.W && row < p.H;
    const DevScene& sc = p.sc;
    uint32_t* slots = smem + warp * WARP_WORDS;
    float* slotsF = reinterpret_cast<float*>(slots);
    uint32_t* sstack = slots + NW * NS + lane;                        // [depth][lane]
#define SLOT_U(f, i) slots[(f) * NS + (i)]
#define SLOT_F(f, i) slotsF[(f) * NS + (i)]
    Counters cnt = {};
    const int jUp = p.H - 1 - row;
    const uint32_t pix = (uint32_t)(row * p.W + col);
    // sample range of this thread: gridDim.z chunks split the launch's samples
    const int per = (p.sCount + gridDim.z - 1) / gridDim.z;
    int s = p.sBegin + blockIdx.z * per;
    const int sEnd = inside ? min(p.sBegin + p.sCount, s + per) : s;
    const int nMine = max(0, sEnd - s);
    float3 sum = f3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < K; k++) SLOT_U(F_STATE, k * 32 + lane) = (uint32_t)(DEPTH_NEED_RAY + 2);
    uint32_t lstack[LOCAL_STACK];

    for (;;) {
        // ================================================================ shade phase: each lane works on its own K slots
        bool alive = false;
#pragma unroll 1
        for (int k = 0; k < K; k++) {
            const int slot = k * 32 + lane;
            const uint32_t st = SLOT_U(F_STATE, slot);
            int depth = (int)(st & 255u) - 2;
            uint32_t cs = st >> 8;
            RngSeq rng;
            if (depth >= 0) {
                float3 O = f3(SLOT_F(F_OX, slot), SLOT_F(F_OY, slot), SLOT_F(F_OZ, slot));
                float3 D = f3(SLOT_F(F_DX, slot), SLOT_F(F_DY, slot), SLOT_F(F_DZ, slot));
                float3 beta = f3(SLOT_F(F_BX, slot), SLOT_F(F_BY, slot), SLOT_F(F_BZ, slot));
                Hit best; best.t = SLOT_F(F_T, slot); best.pid = SLOT_U(F_PID, slot);
                const uint32_t self = SLOT_U(F_SELF, slot);
                if (SAMPLER == 1) rng.resume(pix, cs, SLOT_U(F_RNGPOS, slot), p.seedLo, p.seedHi);
                bool cont = false;
                if (best.pid == PID_NONE) {                                   // ray-tracer.js:122
                    sum = sum + beta * background(sc, D);
                } else {
                    Surface sf = make_surface(sc, best, O, D, self);
                    if (PRECISE && depth == 0) {                              // primary hit: float64 evaluation of the selected primitive
                        D3 O64, D64; double t64;
                        RngSeq again;                                         // re-derive this path's camera sample (cheaper than 4 words per slot)
                        CamSample cam = camera_sample<SAMPLER>(p, pix, cs, again);
                        camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                        refine_primary(sc, best.pid, O64, D64, t64, sf);
                    }
                    float4 m = ldg4(sc.mat + sf.matId);
                    int mt = __ldg(sc.matType + sf.matId);
                    if ((mt & 255) == 3) sum = sum + beta * (f3(m.x, m.y, m.z) * m.w);    // emitted (materials.js:95)
                    if (DIRECT && (mt & 255) == 0) {
                        // EXTENSION (off by default): lights.js:22-47 give direction / colour / distance.
                        for (int li = 0; li < sc.nLights; li++) {
                            float4 l0 = ldg4(sc.lights + 2 * li), l1 = ldg4(sc.lights + 2 * li + 1);
                            float3 ldir, lcol = f3(l1.x, l1.y, l1.z); float ldist;
                            if (l0.w == 0.f) {
                                float3 d = f3(l0.x, l0.y, l0.z) - sf.P;
                                ldist = sqrtf(dot(d, d));
                                ldir = normalize0(d);
                                lcol = lcol * (1.0f / (1.0f + 0.1f * ldist + 0.01f * ldist * ldist));
                            } else { ldir = f3(-l0.x, -l0.y, -l0.z); ldist = CUDART_INF_F; }
                            float cosN = dot(sf.N, ldir);
                            if (!(cosN > 0.f)) continue;
                            // blocking any-hit traversal (the extend-phase stacks are idle during the shade phase)
                            Hit sh = USE_BVH ? trace_bvh<COUNT, true>(sc, sf.P, ldir, 0.001f, ldist, best.pid, cnt, sstack, 32)
                                             : trace_brute<COUNT, true>(sc, sf.P, ldir, 0.001f, ldist, best.pid, cnt);
                            if (sh.pid != PID_NONE) continue;
                            sum = sum + beta * (f3(m.x, m.y, m.z) * lcol) * cosN;
                        }
                    }
                    float3 Dn, att;
                    cont = scatter<SAMPLER>(p, mt, m, sf, D, pix, cs, depth, rng, Dn, att);
                    depth++;
                    cont = cont && depth < p.maxDepth;                        // depth <= 0 returns black (ray-tracer.js:103)
                    if (cont) {
                        beta = beta * att;
                        SLOT_F(F_BX, slot) = beta.x; SLOT_F(F_BY, slot) = beta.y; SLOT_F(F_BZ, slot) = beta.z;
                        SLOT_F(F_OX, slot) = sf.P.x; SLOT_F(F_OY, slot) = sf.P.y; SLOT_F(F_OZ, slot) = sf.P.z;
                        SLOT_F(F_DX, slot) = Dn.x; SLOT_F(F_DY, slot) = Dn.y; SLOT_F(F_DZ, slot) = Dn.z;
                        SLOT_U(F_SELF, slot) = best.pid;
                        if (SAMPLER == 1) SLOT_U(F_RNGPOS, slot) = rng.pos;
                    }
                }
                if (!cont) depth = DEPTH_NEED_RAY;
            }
            if (depth == DEPTH_NEED_RAY) {
                if (s < sEnd) {
                    cs = (uint32_t)s++;
                    CamSample cam = camera_sample<SAMPLER>(p, pix, cs, rng);
                    float3 O, D;
                    if (PRECISE) {
                        D3 O64, D64;
                        camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                        O = tof3(O64); D = tof3(D64);
                    } else camera_ray32(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O, D);
                    SLOT_F(F_OX, slot) = O.x; SLOT_F(F_OY, slot) = O.y; SLOT_F(F_OZ, slot) = O.z;
                    SLOT_F(F_DX, slot) = D.x; SLOT_F(F_DY, slot) = D.y; SLOT_F(F_DZ, slot) = D.z;
                    SLOT_F(F_BX, slot) = 1.f; SLOT_F(F_BY, slot) = 1.f; SLOT_F(F_BZ, slot) = 1.f;
                    SLOT_U(F_SELF, slot) = PID_NONE;
                    if (SAMPLER == 1) SLOT_U(F_RNGPOS, slot) = rng.pos;
                    depth = 0;
                } else depth = DEPTH_DEAD;
            }
            SLOT_U(F_STATE, slot) = (cs << 8) | (uint32_t)(depth + 2);
            if (depth >= 0) {
                alive = true;
                // arm the ray: unbounded planes (outside the BVH) are tested here, by the owner lane
                if (COUNT) cnt.rays++;
                float3 O = f3(SLOT_F(F_OX, slot), SLOT_F(F_OY, slot), SLOT_F(F_OZ, slot));
                float3 D = f3(SLOT_F(F_DX, slot), SLOT_F(F_DY, slot), SLOT_F(F_DZ, slot));
                const uint32_t self = SLOT_U(F_SELF, slot);
                Hit best;
                if (USE_BVH) {
                    best.t = CUDART_INF_F; best.pid = PID_NONE;
                    test_planes<COUNT, false>(sc, O, D, 0.001f, self, best, cnt);
                } else {
                    best = trace_brute<COUNT, false>(sc, O, D, 0.001f, CUDART_INF_F, self, cnt);
                }
                SLOT_F(F_T, slot) = best.t; SLOT_U(F_PID, slot) = best.pid;
            }
        }
        if (!__any_sync(0xffffffffu, alive)) break;
        if (!USE_BVH) continue;
        __syncwarp();
        // ================================================================ extend phase: the warp's NS rays are a queue
        {
            int next = 0;                                             // warp-uniform queue head
            int idx = -1;                                             // slot this lane is traversing (-1: idle)
            float3 O = f3(0, 0, 0), D = f3(0, 0, 1);
            RayInv ri; ri.inv = f3(1, 1, 1); ri.ood = f3(0, 0, 0);
            Hit best; best.t = 0.f; best.pid = PID_NONE;
            uint32_t self = PID_NONE, cur = TRAV_DONE;
            int sp = 0;
            for (;;) {
                // refill idle lanes from the queue once enough of them are idle (or nothing is running)
                const unsigned idleMask = __ballot_sync(0xffffffffu, idx < 0);
                if (next < NS && (__popc(idleMask) >= p.refill || idleMask == 0xffffffffu)) {
                    if (idx < 0) {
                        const int cand = next + __popc(idleMask & ((1u << lane) - 1u));
                        if (cand < NS && (SLOT_U(F_STATE, cand) & 255u) >= 2u) {
                            idx = cand;
                            O = f3(SLOT_F(F_OX, idx), SLOT_F(F_OY, idx), SLOT_F(F_OZ, idx));
                            D = f3(SLOT_F(F_DX, idx), SLOT_F(F_DY, idx), SLOT_F(F_DZ, idx));
                            self = SLOT_U(F_SELF, idx);
                            best.t = SLOT_F(F_T, idx); best.pid = SLOT_U(F_PID, idx);
                            ri = ray_inv(O, D);
                            cur = 0; sp = 0;
                        }
                    }
                    next += __popc(idleMask);
                    continue;                                         // re-evaluate: a fetched slot may have been dead
                }
                if (idleMask == 0xffffffffu) break;                   // queue drained and nobody is traversing
                if (idx >= 0) {
                    if (!(cur & LEAF_BIT)) {
                        if (COUNT) cnt.aabb += 2;
                        uint32_t nearc, farc; bool both;
                        if (node_visit(sc.nodes, cur, ri, best.t, nearc, farc, both)) {
                            if (both) {
                                if (sp < SMEM_STACK) sstack[sp * 32] = farc; else lstack[sp - SMEM_STACK] = farc;
                                sp++;
                            }
                            cur = nearc;
                        } else cur = TRAV_DONE - 1u;                   // "pop" marker (a leaf-bit value that is never a pid)
                    }
                    if ((cur & LEAF_BIT) && cur < TRAV_DONE - 1u) {
                        test_prim<COUNT, false>(sc, cur & ~LEAF_BIT, O, D, 0.001f, self, best, cnt);
                        cur = TRAV_DONE - 1u;
                    }
                    if (cur == TRAV_DONE - 1u) {
                        if (sp == 0) {
                            SLOT_F(F_T, idx) = best.t; SLOT_U(F_PID, idx) = best.pid;
                            idx = -1;
                        } else { sp--; cur = sp < SMEM_STACK ? sstack[sp * 32] : lstack[sp - SMEM_STACK]; }
                    }
                }
            }
        }
        __syncwarp();
    }
    if (inside) {
        // each z chunk owns its own plane of the accumulation target (planeStride = 0 when there is one chunk):
        // no atomics, so the sum is deterministic; k_sum_planes folds the planes in fixed order afterwards
        float4* dst = p.accum + (size_t)blockIdx.z * p.planeStride + pix;
        float4 a = *dst;
        a.x += sum.x; a.y += sum.y; a.z += sum.z; a.w += (float)nMine;
        *dst = a;
    }
    if (COUNT) {
        unsigned long long* v = reinterpret_cast<unsigned long long*>(&cnt);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(p.counters + k, x);
        }
    }
#undef SLOT_U
#undef SLOT_F
}

// ------------------------------------------------------------------------------------------- the megakernel
// One thread = one pixel, looping over its samples and regenerating a camera ray as soon as the current path ends, with a
// blocking per-ray BVH traversal.  Measured against the warp-local wavefront above on the 1920x1080 random-spheres scene
// (profiles/): the wavefront raises SIMD efficiency of the traversal loop (15 -> 22 active lanes per instruction) but pays
// for it in queue traffic, refill code and L1 capacity lost to shared memory; the megakernel is faster there and is what
// BRT_INTEGRATOR_AUTO selects.  Both produce the same image up to fp32 summation order.
template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT, bool HYBRID>
__global__ void __launch_bounds__(PT_BLOCK, PT_MIN_BLOCKS_MEGA) k_pathtrace_mega(const __grid_constant__ PTParams p) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = col < p.W && row < p.H;
    const DevScene& sc = p.sc;
    // float64 primary rays + float64 evaluation of the primary hit: always with the sequential (reference) sampler — the mode
    // that is compared sample for sample with the float64 oracle — and in the AOV kernel; the fast sampler is fp32 throughout
#ifdef BRT_PRECISE_ALWAYS
    constexpr bool PRECISE = true;
#else
    constexpr bool PRECISE = SAMPLER == 1;
#endif
    Counters cnt = {};
    if (inside) {
        const int jUp = p.H - 1 - row;
        const uint32_t pix = (uint32_t)(row * p.W + col);
        const int per = (p.sCount + gridDim.z - 1) / gridDim.z;
        int s = p.sBegin + blockIdx.z * per;
        const int sEnd = min(p.sBegin + p.sCount, s + per);
        const int nMine = max(0, sEnd - s);
        float3 sum = f3(0.f, 0.f, 0.f), beta = f3(1.f, 1.f, 1.f), O = f3(0, 0, 0), D = f3(0, 0, 1);
        uint32_t self = PID_NONE, cs = 0;
        int depth = 0;
        bool alive = false;
        CamSample cam = {};
        RngSeq rng;
        uint32_t* sstack = smem + threadIdx.x;
        for (;;) {
            if (!alive) {
                if (s >= sEnd) break;
                cs = (uint32_t)s++;
                cam = camera_sample<SAMPLER>(p, pix, cs, rng);
                if (PRECISE) {
                    D3 O64, D64;
                    camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                    O = tof3(O64); D = tof3(D64);
                } else camera_ray32(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O, D);
                beta = f3(1.f, 1.f, 1.f); self = PID_NONE; depth = 0; alive = true;
            }
            Hit h = trace<USE_BVH, COUNT, false, HYBRID>(sc, O, D, CUDART_INF_F, self, cnt, sstack, PT_BLOCK);
            if (h.pid == PID_NONE) {                                          // ray-tracer.js:122
                sum = sum + beta * background(sc, D);
                alive = false;
                continue;
            }
            Surface sf = make_surface(sc, h, O, D, self);
            if (PRECISE && depth == 0) {                                      // primary hit: float64 evaluation of the selected primitive
                D3 O64, D64; double t64;
                camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                refine_primary(sc, h.pid, O64, D64, t64, sf);
            }
            float4 m = ldg4(sc.mat + sf.matId);
            int mt = __ldg(sc.matType + sf.matId);
            if ((mt & 255) == 3) sum = sum + beta * (f3(m.x, m.y, m.z) * m.w);        // emitted (materials.js:95)
            if (DIRECT && (mt & 255) == 0) {
                // EXTENSION (off by default): lights.js:22-47 give direction / colour / distance.
                for (int li = 0; li < sc.nLights; li++) {
                    float4 l0 = ldg4(sc.lights + 2 * li), l1 = ldg4(sc.lights + 2 * li + 1);
                    float3 ldir, lcol = f3(l1.x, l1.y, l1.z); float ldist;
                    if (l0.w == 0.f) {
                        float3 d = f3(l0.x, l0.y, l0.z) - sf.P;
                        ldist = sqrtf(dot(d, d));
                        ldir = normalize0(d);
                        lcol = lcol * (1.0f / (1.0f + 0.1f * ldist + 0.01f * ldist * ldist));
                    } else { ldir = f3(-l0.x, -l0.y, -l0.z); ldist = CUDART_INF_F; }
                    float cosN = dot(sf.N, ldir);
                    if (!(cosN > 0.f)) continue;
                    Hit sh = trace<USE_BVH, COUNT, true, HYBRID>(sc, sf.P, ldir, ldist, h.pid, cnt, sstack, PT_BLOCK);
                    if (sh.pid != PID_NONE) continue;
                    sum = sum + beta * (f3(m.x, m.y, m.z) * lcol) * cosN;
                }
            }
            float3 Dn, att;
            bool cont = scatter<SAMPLER>(p, mt, m, sf, D, pix, cs, depth, rng, Dn, att);
            depth++;
            if (!cont || depth >= p.maxDepth) { alive = false; continue; }   // depth <= 0 returns black (ray-tracer.js:103)
            beta = beta * att;
            O = sf.P; D = Dn; self = h.pid;
        }
        // each z chunk owns its own plane of the accumulation target (planeStride = 0 when there is one chunk):
        // no atomics, so the sum is deterministic; k_sum_planes folds the planes in fixed order afterwards
        float4* dst = p.accum + (size_t)blockIdx.z * p.planeStride + pix;
        float4 a = *dst;
        a.x += sum.x; a.y += sum.y; a.z += sum.z; a.w += (float)nMine;
        *dst = a;
    }
    if (COUNT) {
        unsigned long long* v = reinterpret_cast<unsigned long long*>(&cnt);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned long long x = v[k];
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0 && x) atomicAdd(p.counters + k, x);
        }
    }
}

// ------------------------------------------------------------------------------------------- primary AOVs (fp32 render-path code)
template <bool USE_BVH>
__global__ void __launch_bounds__(PT_BLOCK) k_primary_aov(const __grid_constant__ PTParams p, int* objId, int* triId, float* tOut,
                                                          float* nrm, unsigned char* front) {
    extern __shared__ uint32_t smem_stack[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (col >= p.W || row >= p.H) return;
    const int jUp = p.H - 1 - row;
    CamSample cam = {};                                        // pixel centre, lens offset 0 (ray-tracer.js:144-147)
    D3 O64, D64;
    camera_ray64(p.cam, p.W, p.H, 0, col, jUp, cam, O64, D64);
    float3 O = tof3(O64), D = tof3(D64);
    Counters cnt;
    Hit h = trace<USE_BVH, false, false>(p.sc, O, D, CUDART_INF_F, PID_NONE, cnt, smem_stack + threadIdx.x, PT_BLOCK);
    size_t k = (size_t)row * p.W + col;
    if (h.pid == PID_NONE) {
        objId[k] = -1; triId[k] = -1; tOut[k] = CUDART_INF_F; nrm[3 * k] = nrm[3 * k + 1] = nrm[3 * k + 2] = 0.f; front[k] = 0;
    } else {
        Surface sf = make_surface(p.sc, h, O, D, PID_NONE);
        double t64 = (double)h.t;
        refine_primary(p.sc, h.pid, O64, D64, t64, sf);           // the render path's own primary-hit code
        objId[k] = sf.objId; triId[k] = sf.triId; tOut[k] = (float)t64;
        nrm[3 * k] = sf.N.x; nrm[3 * k + 1] = sf.N.y; nrm[3 * k + 2] = sf.N.z; front[k] = sf.front ? 1 : 0;
    }
}

__global__ void k_eval_background(DevScene sc, const float* dirs, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = background(sc, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

__global__ void k_eval_texture(DevScene sc, int ti, const float* pts, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float3 c = texture_value(sc, ti, f3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
}

__global__ void k_rng_stream(uint32_t seedLo, uint32_t seedHi, uint32_t pixel, uint32_t sample, int n, float* out) {
    if (blockIdx.x || threadIdx.x) return;
    RngSeq r; r.init(pixel, sample, seedLo, seedHi);
    for (int i = 0; i < n; i++) out[i] = r.next();
}

// Dense FFMA throughput probe (the measured FP32 roofline denominator): 8 independent chains per thread.
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;
}

// ------------------------------------------------------------------------------------------- host launchers
template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT>
static cudaError_t launch_pt3(const PTParams& p, dim3 grid, cudaStream_t st) {
    auto go = [&](auto kern, int K) -> cudaError_t {
        size_t smem = (size_t)(PT_BLOCK / 32) * (slot_words(SAMPLER) * 32 * K + 32 * SMEM_STACK) * sizeof(uint32_t);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, PT_BLOCK, smem, st>>>(p);
        return cudaGetLastError();
    };
    if (!p.wavefront) {
        // the stack: depth + 1 entries per thread, all in shared memory, unless the tree is unusually deep
        const bool hybrid = USE_BVH && p.sc.bvhStackDepth > SMEM_ONLY_MAX_DEPTH;
        size_t smem = USE_BVH ? (size_t)(hybrid ? SMEM_STACK : p.sc.bvhStackDepth + 1) * PT_BLOCK * sizeof(uint32_t) : 0;
        if (hybrid) k_pathtrace_mega<SAMPLER, USE_BVH, COUNT, DIRECT, true><<<grid, PT_BLOCK, smem, st>>>(p);
        else k_pathtrace_mega<SAMPLER, USE_BVH, COUNT, DIRECT, false><<<grid, PT_BLOCK, smem, st>>>(p);
        return cudaGetLastError();
    }
    switch (p.inflight) {
    case 1: return go(k_pathtrace_wave<SAMPLER, USE_BVH, COUNT, DIRECT, 1>, 1);
    case 3: return go(k_pathtrace_wave<SAMPLER, USE_BVH, COUNT, DIRECT, 3>, 3);
    case 4: return go(k_pathtrace_wave<SAMPLER, USE_BVH, COUNT, DIRECT, 4>, 4);
    default: return go(k_pathtrace_wave<SAMPLER, USE_BVH, COUNT, DIRECT, 2>, 2);
    }
}
template <int SAMPLER, bool USE_BVH, bool COUNT>
static cudaError_t launch_pt2(const PTParams& p, dim3 grid, cudaStream_t st) {
    return p.directLighting ? launch_pt3<SAMPLER, USE_BVH, COUNT, true>(p, grid, st) : launch_pt3<SAMPLER, USE_BVH, COUNT, false>(p, grid, st);
}
template <int SAMPLER>
static cudaError_t launch_pt1(const PTParams& p, bool bvh, bool count, dim3 grid, cudaStream_t st) {
    if (bvh) return count ? launch_pt2<SAMPLER, true, true>(p, grid, st) : launch_pt2<SAMPLER, true, false>(p, grid, st);
    return count ? launch_pt2<SAMPLER, false, true>(p, grid, st) : launch_pt2<SAMPLER, false, false>(p, grid, st);
}

cudaError_t launch_pathtrace(const PTParams& p, int sampler, bool useBvh, bool count, int zSplit, cudaStream_t st) {
    dim3 grid((p.W + 15) / 16, (p.H + 7) / 8, zSplit < 1 ? 1 : zSplit);
    return sampler == 1 ? launch_pt1<1>(p, useBvh, count, grid, st) : launch_pt1<0>(p, useBvh, count, grid, st);
}

__global__ void __launch_bounds__(256) k_sum_planes(float4* __restrict__ accum, const float4* __restrict__ planes, int nPlanes, size_t px) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= px) return;
    float4 a = accum[i];
    for (int z = 0; z < nPlanes; z++) {                     // fixed order: bit-reproducible
        float4 v = __ldg(planes + (size_t)z * px + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    accum[i] = a;
}
cudaError_t launch_sum_planes(float4* accum, const float4* planes, int nPlanes, size_t px, cudaStream_t st) {
    k_sum_planes<<<(unsigned)((px + 255) / 256), 256, 0, st>>>(accum, planes, nPlanes, px);
    return cudaGetLastError();
}

cudaError_t launch_primary_aov(const PTParams& p, bool useBvh, int* objId, int* triId, float* t, float* nrm, unsigned char* front,
                               cudaStream_t st) {
    dim3 grid((p.W + 15) / 16, (p.H + 7) / 8, 1);
    size_t smem = useBvh ? (size_t)SMEM_STACK * PT_BLOCK * sizeof(uint32_t) : 0;
    if (useBvh) k_primary_aov<true><<<grid, PT_BLOCK, smem, st>>>(p, objId, triId, t, nrm, front);
    else k_primary_aov<false><<<grid, PT_BLOCK, smem, st>>>(p, objId, triId, t, nrm, front);
    return cudaGetLastError();
}

cudaError_t launch_eval_background(const DevScene& sc, const float* dirs, int n, float* out, cudaStream_t st) {
    k_eval_background<<<(n + 127) / 128, 128, 0, st>>>(sc, dirs, n, out);
    return cudaGetLastError();
}
cudaError_t launch_eval_texture(const DevScene& sc, int texIndex, const float* points, int n, float* out, cudaStream_t st) {
    k_eval_texture<<<(n + 127) / 128, 128, 0, st>>>(sc, texIndex, points, n, out);
    return cudaGetLastError();
}
cudaError_t launch_rng_stream(uint32_t lo, uint32_t hi, uint32_t pixel, uint32_t sample, int n, float* out, cudaStream_t st) {
    k_rng_stream<<<1, 32, 0, st>>>(lo, hi, pixel, sample, n, out);
    return cudaGetLastError();
}
cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st) {
    k_fp32_peak<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
    return cudaGetLastError();
}

}  // namespace brt
