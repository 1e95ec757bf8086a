// Megakernel path tracer: one thread owns one pixel and regenerates a new camera sample as soon as its
// current path terminates ("persistent per pixel over samples"), so warps stay full while paths of very
// different length (1 bounce into the sky … max_depth bounces between glass spheres) share a warp.
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
    if (aaMode == 1) { u = ((double)col + (double)cs.s) / W; v = ((double)jUp + (double)cs.t) / H; }
    else if (aaMode == 2) { u = ((double)col + 0.5 + (double)cs.s * 0.5) / W; v = ((double)jUp + 0.5 + (double)cs.t * 0.5) / H; }
    else { u = ((double)col + 0.5) / W; v = ((double)jUp + 0.5) / H; }
    double rx = (double)cs.dx * c.lensRadius, ry = (double)cs.dy * c.lensRadius;
    D3 cu = d3(c.cu[0], c.cu[1], c.cu[2]), cv = d3(c.cv[0], c.cv[1], c.cv[2]);
    O = d3(c.o[0], c.o[1], c.o[2]) + cu * rx + cv * ry;
    D = d3(c.ll[0], c.ll[1], c.ll[2]) + d3(c.h[0], c.h[1], c.h[2]) * u + d3(c.v[0], c.v[1], c.v[2]) * v - O;
    if (c.type == 1) D = normalize0(D + d3(c.cw[0], c.cw[1], c.cw[2]) * -1.0);   // camera.js:42-43
}

template <int SAMPLER>
__device__ __forceinline__ CamSample camera_sample(const PTParams& p, uint32_t pix, uint32_t s, RngSeq& rng) {
    CamSample cs; cs.s = 0.f; cs.t = 0.f;
    float a0 = 0.f, a1 = 0.f;
    if (SAMPLER == 0) {
        uint4 r = philox4x32_10(pix, s, 0u, PHILOX_TAG, p.seedLo, p.seedHi);
        a0 = u01(r.x); a1 = u01(r.y);
        float rr = sqrtf(u01(r.z)), sn, cs_;                   // unit disk by inversion (math.js:27-31 distribution)
        sincospif(2.f * u01(r.w), &sn, &cs_);
        cs.dx = rr * cs_; cs.dy = rr * sn;
    } else {
        rng.init(pix, s, p.seedLo, p.seedHi);
        if (p.aaMode == 1 || p.aaMode == 2) { a0 = rng.next(); a1 = rng.next(); }
        do { cs.dx = rng.next() * 2.f - 1.f; cs.dy = rng.next() * 2.f - 1.f; } while (cs.dx * cs.dx + cs.dy * cs.dy >= 1.0f);   // math.js:29
    }
    if (p.aaMode == 1) { cs.s = a0; cs.t = a1; }
    else if (p.aaMode == 2) {                                    // stochastic: disk of radius 0.5 about the pixel centre
        float sr = sqrtf(a0), sn, c2;
        sincospif(2.f * a1, &sn, &c2);
        cs.s = sr * c2; cs.t = sr * sn;
    }
    return cs;
}

// ------------------------------------------------------------------------------------------- materials (materials.js)
// Returns false when the path ends here (emissive, absorbed metal).  `att` multiplies the throughput.
template <int SAMPLER>
__device__ __forceinline__ bool scatter(const PTParams& p, int matType, float4 m, const Surface& sf, float3 Din, uint32_t pix,
                                        uint32_t s, int bounce, RngSeq& rng, float3& Dout, float3& att) {
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (SAMPLER == 0) {
        uint4 r = philox4x32_10(pix, s, (uint32_t)(bounce + 1), PHILOX_TAG, p.seedLo, p.seedHi);
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
        att = f3(m.x, m.y, m.z);
        return true;
    }
    if (matType == 1) {                                                       // Metal (materials.js:36-41)
        float3 refl = reflect(normalize0(Din), sf.N);
        float3 ball;
        if (SAMPLER == 0) ball = uniform_sphere(u0, u1) * cbrtf(u2);
        else { do { ball = f3(rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f, rng.next() * 2.f - 1.f); } while (dot(ball, ball) >= 1.0f); }
        Dout = madd(ball, m.w, refl);
        att = f3(m.x, m.y, m.z);
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

template <bool USE_BVH, bool COUNT, bool SHADOW>
__device__ __forceinline__ Hit trace(const DevScene& sc, float3 O, float3 D, float tMax, uint32_t self, Counters& cnt,
                                     uint32_t* sstack, int sstride) {
    if (COUNT && !SHADOW) cnt.rays++;
    if (USE_BVH) return trace_bvh<COUNT, SHADOW>(sc, O, D, 0.001f, tMax, self, cnt, sstack, sstride);
    return trace_brute<COUNT, SHADOW>(sc, O, D, 0.001f, tMax, self, cnt);
}

// ------------------------------------------------------------------------------------------- the megakernel
// Block = 128 threads = a 16x8 pixel tile; a warp = an 8x4 sub-tile (coherent primary rays).  One thread owns one pixel
// and all of its samples; its state machine is
//     NEED_RAY -> [camera sample] -> TRAV -> [BVH steps ...] -> SHADE -> [scatter] -> TRAV ... -> NEED_RAY -> ... -> EXIT
// A warp alternates between two phases:
//   S  every lane whose traversal has finished shades its hit, scatters or regenerates a camera ray and re-arms traversal;
//   T  all traversing lanes take BVH steps in lock step UNTIL fewer than `refill` lanes are still traversing — then the
//      warp goes back to S to refill the idle lanes while the unfinished lanes keep their traversal state (cur, sp, stack).
// So the traversal loop never runs with fewer than `refill` of 32 lanes busy (a plain per-ray while loop ran at ~40 % SIMD
// efficiency on the random-spheres scene because every warp waited for its longest ray), and shading never runs with
// fewer than 32 - refill lanes unless the warp is draining.
enum LaneState : int { ST_NEED_RAY = 0, ST_TRAV = 1, ST_SHADE = 2, ST_EXIT = 3 };

template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT>
__global__ void __launch_bounds__(PT_BLOCK, PT_MIN_BLOCKS) k_pathtrace(const __grid_constant__ PTParams p) {
    extern __shared__ uint32_t smem_stack[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = col < p.W && row < p.H;
    const DevScene& sc = p.sc;
    Counters cnt = {};
    const int jUp = p.H - 1 - row;
    const uint32_t pix = (uint32_t)(row * p.W + col);
    // sample range of this thread: gridDim.z chunks split the launch's samples
    const int per = (p.sCount + gridDim.z - 1) / gridDim.z;
    int s = p.sBegin + blockIdx.z * per;
    const int sEnd = inside ? min(p.sBegin + p.sCount, s + per) : s;
    const int nMine = max(0, sEnd - s);

    // cold per-path state (touched once per bounce / per sample) lives in shared memory, [slot][thread]: keeps the
    // traversal loop's register set small.  slots: 0-2 radiance sum, 3-5 throughput beta, 6-9 camera sample
    float* cold = reinterpret_cast<float*>(smem_stack + (size_t)p.stackSlots * PT_BLOCK) + threadIdx.x;
#define COLD(k) cold[(k) * PT_BLOCK]
    COLD(0) = 0.f; COLD(1) = 0.f; COLD(2) = 0.f;
    float3 O = f3(0, 0, 0), D = f3(0, 0, 1);
    RayInv ri; ri.inv = f3(1, 1, 1); ri.ood = f3(0, 0, 0);
    Hit best; best.t = CUDART_INF_F; best.pid = PID_NONE;
    uint32_t self = PID_NONE, cur = TRAV_DONE;
    int depth = 0, sp = 0, state = ST_NEED_RAY;
    RngSeq rng;
    uint32_t* sstack = smem_stack + threadIdx.x;
    uint32_t lstack[LOCAL_STACK];
    const int refill = p.refill;

    for (;;) {
        // ---------------------------------------------------------------- phase S: shade / scatter / regenerate
        if (state == ST_SHADE) {
            state = ST_NEED_RAY;
            const uint32_t cs = (uint32_t)(s - 1);                            // the camera sample this path belongs to
            float3 beta = f3(COLD(3), COLD(4), COLD(5));
            float3 add = f3(0.f, 0.f, 0.f);
            if (best.pid == PID_NONE) {                                       // ray-tracer.js:122
                add = beta * background(sc, D);
            } else {
                Surface sf = make_surface(sc, best, O, D, self);
                if (depth == 0) {                                             // primary hit: float64 evaluation of the selected primitive
                    D3 O64, D64; double t64;
                    CamSample cam; cam.s = COLD(6); cam.t = COLD(7); cam.dx = COLD(8); cam.dy = COLD(9);
                    camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                    refine_primary(sc, best.pid, O64, D64, t64, sf);
                }
                float4 m = ldg4(sc.mat + sf.matId);
                int mt = __ldg(sc.matType + sf.matId);
                if (mt == 3) add = beta * (f3(m.x, m.y, m.z) * m.w);          // emitted (materials.js:95)
                if (DIRECT && mt == 0) {
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
                        // blocking any-hit traversal; it uses stack slots above this lane's live entries
                        Hit sh = USE_BVH ? trace_bvh<COUNT, true>(sc, sf.P, ldir, 0.001f, ldist, best.pid, cnt, sstack + (SMEM_STACK + PT_COLD_SLOTS) * PT_BLOCK, PT_BLOCK)
                                         : trace_brute<COUNT, true>(sc, sf.P, ldir, 0.001f, ldist, best.pid, cnt);
                        if (sh.pid != PID_NONE) continue;
                        add = add + beta * (f3(m.x, m.y, m.z) * lcol) * cosN;
                    }
                }
                float3 Dn, att;
                bool cont = scatter<SAMPLER>(p, mt, m, sf, D, pix, cs, depth, rng, Dn, att);
                depth++;
                if (cont && depth < p.maxDepth) {                             // depth <= 0 returns black (ray-tracer.js:103)
                    beta = beta * att;
                    COLD(3) = beta.x; COLD(4) = beta.y; COLD(5) = beta.z;
                    O = sf.P; D = Dn; self = best.pid;
                    state = ST_TRAV;
                }
            }
            if (add.x != 0.f || add.y != 0.f || add.z != 0.f) { COLD(0) += add.x; COLD(1) += add.y; COLD(2) += add.z; }
        }
        if (state == ST_NEED_RAY) {
            if (s < sEnd) {
                CamSample cam = camera_sample<SAMPLER>(p, pix, (uint32_t)s++, rng);
                D3 O64, D64;
                camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                O = tof3(O64); D = tof3(D64);
                COLD(3) = 1.f; COLD(4) = 1.f; COLD(5) = 1.f;
                COLD(6) = cam.s; COLD(7) = cam.t; COLD(8) = cam.dx; COLD(9) = cam.dy;
                self = PID_NONE; depth = 0;
                state = ST_TRAV;
            } else state = ST_EXIT;
        }
        if (state == ST_TRAV && cur == TRAV_DONE) {                           // arm traversal for a fresh ray
            if (COUNT) cnt.rays++;
            if (USE_BVH) {
                best.t = CUDART_INF_F; best.pid = PID_NONE;
                test_planes<COUNT, false>(sc, O, D, 0.001f, self, best, cnt);
                ri = ray_inv(O, D);
                cur = 0; sp = 0;
            } else {
                best = trace_brute<COUNT, false>(sc, O, D, 0.001f, CUDART_INF_F, self, cnt);
                state = ST_SHADE;
            }
        }
        const unsigned live = __ballot_sync(0xffffffffu, state != ST_EXIT);
        if (live == 0u) break;
        if (!USE_BVH) continue;
        // ---------------------------------------------------------------- phase T: lock-step BVH steps
        const int need = min(refill, __popc(live));
        for (;;) {
            if (cur != TRAV_DONE) {
                if (!(cur & LEAF_BIT)) {
                    if (COUNT) cnt.aabb += 2;
                    uint32_t nearc, farc; bool both;
                    if (node_visit(sc.nodes, cur, ri, best.t, nearc, farc, both)) {
                        if (both) {
                            if (sp < SMEM_STACK) sstack[sp * PT_BLOCK] = farc; else lstack[sp - SMEM_STACK] = farc;
                            sp++;
                        }
                        cur = nearc;
                    } else cur = TRAV_DONE - 1u;                               // "pop" marker (a leaf-bit value that is never a pid)
                }
                if ((cur & LEAF_BIT) && cur < TRAV_DONE - 1u) {
                    test_prim<COUNT, false>(sc, cur & ~LEAF_BIT, O, D, 0.001f, self, best, cnt);
                    cur = TRAV_DONE - 1u;
                }
                if (cur == TRAV_DONE - 1u) {
                    if (sp == 0) { cur = TRAV_DONE; state = ST_SHADE; }
                    else { sp--; cur = sp < SMEM_STACK ? sstack[sp * PT_BLOCK] : lstack[sp - SMEM_STACK]; }
                }
            }
            if (__popc(__ballot_sync(0xffffffffu, cur != TRAV_DONE)) < need) break;
        }
    }
    if (inside) {
        const float3 sum = f3(COLD(0), COLD(1), COLD(2));
        float4* dst = p.accum + pix;
        if (gridDim.z == 1) {
            float4 a = *dst;
            a.x += sum.x; a.y += sum.y; a.z += sum.z; a.w += (float)nMine;
            *dst = a;
        } else {
            atomicAdd(&dst->x, sum.x); atomicAdd(&dst->y, sum.y); atomicAdd(&dst->z, sum.z); atomicAdd(&dst->w, (float)nMine);
        }
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
template <int SAMPLER, bool USE_BVH, bool COUNT>
static cudaError_t launch_pt2(const PTParams& p, dim3 grid, size_t smem, cudaStream_t st) {
    if (p.directLighting) k_pathtrace<SAMPLER, USE_BVH, COUNT, true><<<grid, PT_BLOCK, smem, st>>>(p);
    else k_pathtrace<SAMPLER, USE_BVH, COUNT, false><<<grid, PT_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SAMPLER>
static cudaError_t launch_pt1(const PTParams& p, bool bvh, bool count, dim3 grid, size_t smem, cudaStream_t st) {
    if (bvh) return count ? launch_pt2<SAMPLER, true, true>(p, grid, smem, st) : launch_pt2<SAMPLER, true, false>(p, grid, smem, st);
    return count ? launch_pt2<SAMPLER, false, true>(p, grid, smem, st) : launch_pt2<SAMPLER, false, false>(p, grid, smem, st);
}

cudaError_t launch_pathtrace(const PTParams& p, int sampler, bool useBvh, bool count, int zSplit, cudaStream_t st) {
    dim3 grid((p.W + 15) / 16, (p.H + 7) / 8, zSplit < 1 ? 1 : zSplit);
    // [traversal stack (BVH only)] [cold path state] [shadow-ray stack (BVH + direct lighting)]
    PTParams q = p;
    q.stackSlots = useBvh ? SMEM_STACK : 0;
    size_t smem = (size_t)(q.stackSlots + PT_COLD_SLOTS + (useBvh && p.directLighting ? SMEM_STACK : 0)) * PT_BLOCK * sizeof(uint32_t);
    const PTParams& pq = q;
    return sampler == 1 ? launch_pt1<1>(pq, useBvh, count, grid, smem, st) : launch_pt1<0>(pq, useBvh, count, grid, smem, st);
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
cudaError_t launch_rng_stream(uint32_t lo, uint32_t hi, uint32_t pixel, uint32_t sample, int n, float* out, cudaStream_t st) {
    k_rng_stream<<<1, 32, 0, st>>>(lo, hi, pixel, sample, n, out);
    return cudaGetLastError();
}
cudaError_t launch_fp32_peak(float* out, int blocks, int iters, cudaStream_t st) {
    k_fp32_peak<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f);
    return cudaGetLastError();
}

}  // namespace brt
