// Megakernel path tracer (the default integrator) + the primary-AOV and utility kernels.
// One thread owns one pixel and regenerates a new camera sample as soon as its current path terminates, so warps stay full
// while paths of very different length (1 bounce into the sky … max_depth bounces between glass spheres) share a warp.
// Replaces the three nested loops of RayTracer.render (ray-tracer.js:189-206) and the recursion of
// rayColor (ray-tracer.js:102-123) — written iteratively: sum += beta ⊙ E; beta ⊙= attenuation.
#include "pathtrace_common.cuh"
#ifndef BRT_STEAL
#define BRT_STEAL 0
#endif

namespace brt {

#ifndef MEGA_BLOCK
#define MEGA_BLOCK PT_BLOCK            // threads per megakernel block: 16 x (MEGA_BLOCK / 16) pixels, one 8 x 4 sub-tile per warp
                                       // (128 measured best on a B200: 64 threads x 16 blocks / SM -1.0 % C3, -3.0 % C5; 256 x 4 -2.5 %, -0.8 %)
#endif

// ------------------------------------------------------------------------------------------- the megakernel
// One thread = one pixel, looping over its samples and regenerating a camera ray as soon as the current path ends, with a
// blocking per-ray BVH traversal.  Measured against the warp-local wavefront (pathtrace_wave.cu) on the 1920x1080 random-spheres scene
// (profiles/): the wavefront raises SIMD efficiency of the traversal loop (15 -> 22 active lanes per instruction) but pays
// for it in queue traffic, refill code and L1 capacity lost to shared memory; the megakernel is faster there and is what
// BRT_INTEGRATOR_AUTO selects.  Both produce the same image up to fp32 summation order.
// Lane attribution (COUNT build, profiles/r02_lane_attribution_*.md): of the 32 lane slots of a BVH-loop iteration on C3 at
// 256 spp, 58 % work, 35 % wait for the slowest ray of the warp and 7 % are drained (their pixel has no samples left).  A
// variant that dealt each tile's samples to the 32 lanes in balanced runs (lane L traces run r of pixel (L + r) mod 32,
// order-independent fixed-point tile sums in shared memory) cut the drained share only to 6 % — what remains is the random
// spread of 256-sample sums, not a systematic difference between pixels — and lost 2 % to its bookkeeping; it was removed.
template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT, bool HYBRID, int PRIMS = PRIMS_ANY, int WIDE = 0>
__global__ void __launch_bounds__(MEGA_BLOCK, PT_MIN_BLOCKS_MEGA) k_pathtrace_mega(const __grid_constant__ PTParams p) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col0 = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row0 = blockIdx.y * (MEGA_BLOCK / 16) + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = col0 < p.W && row0 < p.rowEnd && row0 >= p.rowBegin;
    const DevScene& sc = p.sc;
    // float64 primary rays + float64 evaluation of the primary hit: always with the sequential (reference) sampler — the mode
    // that is compared sample for sample with the float64 oracle — and in the AOV kernel; the fast sampler is fp32 throughout
#ifdef BRT_PRECISE_ALWAYS
    constexpr bool PRECISE = true;
#else
    constexpr bool PRECISE = SAMPLER == 1;
#endif
    constexpr bool CH = BRT_NODE_CH && SAMPLER == 0;                      // fast sampler: centre / half-extent node copy (node_visit_ch)
    // STEAL (fast sampler): the samples of a warp's 8x4 tile are ONE pool of work items (pixel slot, sample index) that the 32 lanes
    // draw from through a shared-memory counter, instead of lane L owning pixel L's samples.  Path lengths are random, so with the
    // static binding the lanes of a warp finish their pixels at different times and wait drained (6.7 % of the BVH-loop lane
    // slots at 256 spp per launch, 22 % at 16: profiles/r02_lane_attribution_balanced_vs_static.md); drawing from a pool, they all
    // finish within one path of each other.  A sample's value depends only on (pixel, sample index) — the Philox counter — never on
    // the lane that traced it, and the per-pixel sums are kept in fixed point (2^-20, 16-bit limbs) in shared memory, so the result does
    // not depend on which lane added what when: bit-reproducible, and identical for every hierarchy / integrator variant.
    constexpr bool STEAL = BRT_STEAL && SAMPLER == 0;
    __shared__ unsigned int tileSum[STEAL ? MEGA_BLOCK / 32 : 1][STEAL ? 32 : 1][6];
    __shared__ unsigned int tileNext[STEAL ? MEGA_BLOCK / 32 : 1], tileNaN[STEAL ? MEGA_BLOCK / 32 : 1], tileInf[STEAL ? MEGA_BLOCK / 32 : 1];
    const unsigned insideMask = STEAL ? __ballot_sync(0xffffffffu, inside) : 0u;
    if (STEAL) {
#pragma unroll
        for (int k = 0; k < 6; k++) tileSum[warp][lane][k] = 0u;
        if (lane == 0) { tileNext[warp] = 0u; tileNaN[warp] = 0u; tileInf[warp] = 0u; }
        __syncwarp();
    }
    Counters cnt = {};
    if (STEAL ? insideMask != 0u : inside) {
        int col = col0, row = row0;                                           // STEAL: the pixel of the path in flight, not the lane's own
        int jUp = p.H - 1 - row;
        uint32_t pix = (uint32_t)(row * p.W + col);
        const int per = (p.sCount + gridDim.z - 1) / gridDim.z;
        int s = p.sBegin + blockIdx.z * per;
        const int sEnd = min(p.sBegin + p.sCount, s + per);
        const int nMine = max(0, sEnd - s);
        const int tileCnt = __popc(insideMask);
        uint32_t tileTotal = (uint32_t)tileCnt * (uint32_t)nMine;            // work items of this warp's tile
        uint32_t slot = (uint32_t)lane;                                       // pixel slot (0..31 within the tile) of the path in flight
        bool have = false;
        float3 sum = f3(0.f, 0.f, 0.f), beta = f3(1.f, 1.f, 1.f), O = f3(0, 0, 0), D = f3(0, 0, 1);
        uint32_t self = PID_NONE, cs = 0;
        int depth = 0;
        bool alive = false;                                                   // true: the lane holds a hit to scatter from
        Surface sf = {};                                                      // that hit (kept across the back edge only, not across trace)
        uint32_t hitPid = PID_NONE;
        RngSeq rng;
        uint32_t* sstack = smem + threadIdx.x;
        if (p.maxDepth <= 0) { s = sEnd; tileTotal = 0u; }                    // rayColor(depth <= 0) is black (ray-tracer.js:103): only alpha moves
        // One iteration = [draw] -> [start a path: camera ray | continue one: scatter at the previous hit] -> trace -> [miss:
        // background | hit: surface + emission].  The lanes that start a path and the lanes that continue one draw their
        // Philox block in the SAME call (counter block 0 / block depth): one full-warp instance of the generator instead of
        // two half-empty ones.
        for (;;) {
            if (!alive) {
                if (STEAL) {
                    if (have) {                                               // the path that just ended: its radiance into its pixel's tile sum
                        const float m = fmaxf(fmaxf(sum.x, sum.y), sum.z), lo = fminf(fminf(sum.x, sum.y), sum.z);
                        if (m < 2048.f && lo >= 0.f) {                        // (false for NaN too)
                            // 2^-20 fixed point in two 16-bit limbs per channel, each in its own 32-bit word: native 32-bit
                            // shared-memory atomics (the 64-bit add is a compare-and-swap loop), 65 536 samples before a limb can wrap
                            const uint32_t vx = __float2uint_rn(sum.x * 1048576.f), vy = __float2uint_rn(sum.y * 1048576.f), vz = __float2uint_rn(sum.z * 1048576.f);
                            atomicAdd(&tileSum[warp][slot][0], vx & 0xFFFFu); atomicAdd(&tileSum[warp][slot][1], vx >> 16);
                            atomicAdd(&tileSum[warp][slot][2], vy & 0xFFFFu); atomicAdd(&tileSum[warp][slot][3], vy >> 16);
                            atomicAdd(&tileSum[warp][slot][4], vz & 0xFFFFu); atomicAdd(&tileSum[warp][slot][5], vz >> 16);
                        } else if (m != m || lo != lo) atomicOr(&tileNaN[warp], 1u << slot);
                        else atomicOr(&tileInf[warp], 1u << slot);            // beyond the fixed-point range: the pixel saturates
                        sum = f3(0.f, 0.f, 0.f);
                    }
                    const uint32_t idx = atomicAdd(&tileNext[warp], 1u);     // (one warp-aggregated add — ballot, leader, shuffle — measured slower: C3 +0.3 % instead of +1.6 %)
                    if (idx >= tileTotal) break;
                    uint32_t sm;
                    if (tileCnt == 32) { slot = idx & 31u; sm = idx >> 5; }
                    else { sm = idx / (uint32_t)tileCnt; slot = __fns(insideMask, 0u, (int)(idx - sm * (uint32_t)tileCnt) + 1); }
                    cs = (uint32_t)s + sm;
                    col = blockIdx.x * 16 + (warp & 1) * 8 + (int)(slot & 7u);
                    row = blockIdx.y * (MEGA_BLOCK / 16) + (warp >> 1) * 4 + (int)(slot >> 3);
                    jUp = p.H - 1 - row;
                    pix = (uint32_t)(row * p.W + col);
                    have = true;
                } else {
                    if (s >= sEnd) break;
                    cs = (uint32_t)s++;
                }
                depth = 0;
            }
            uint4 r = make_uint4(0u, 0u, 0u, 0u);
            if (SAMPLER == 0) r = philox_fast(pix, cs, (uint32_t)depth, PHILOX_TAG, p.seedLo, p.seedHi);
            Hit h;
            bool primaryDone = false;
            if (!alive) {
                CamSample cam = camera_sample_drawn<SAMPLER>(p, r, pix, cs, rng);
                if (PRECISE) {                                                // primary visibility decided in float64 (trace_primary64)
                    D3 O64, D64; double t64 = 0.0;
                    camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                    O = tof3(O64); D = tof3(D64);
                    if (COUNT) cnt.rays++;
                    h.pid = PID_NONE; h.t = CUDART_INF_F;
                    if (trace_primary64<USE_BVH, HYBRID>(sc, O64, D64, sstack, MEGA_BLOCK, h.pid, t64, sf)) h.t = (float)t64;
                    primaryDone = true;
                } else camera_ray32(p.cam32, p.W, p.H, p.aaMode, col, jUp, cam, O, D);
                beta = f3(1.f, 1.f, 1.f); self = PID_NONE;
            } else {
                float4 m = ldg4(sc.mat + sf.matId);
                int mt = __ldg(sc.matType + sf.matId);
                float3 Dn, att;
                if (!scatter_drawn<SAMPLER>(p, mt, m, sf, D, r, rng, Dn, att)) { alive = false; continue; }   // absorbed (metal below the horizon)
                beta = beta * att;
                O = sf.P; D = Dn; self = hitPid;
            }
            alive = false;
            unsigned aliveMask = 0xffffffffu;
            if (COUNT) {
                aliveMask = __activemask();
                if (lane == __ffs(aliveMask) - 1) { cnt.mainIter++; cnt.mainLanes += __popc(aliveMask); }
            }
            if (!(PRECISE && primaryDone)) h = trace<USE_BVH, COUNT, false, HYBRID, PRIMS, WIDE, CH>(sc, O, D, CUDART_INF_F, self, cnt, sstack, MEGA_BLOCK, aliveMask);
            if (h.pid == PID_NONE) {                                          // ray-tracer.js:122
                sum = sum + beta * background(sc, D);
                continue;
            }
            if (!(PRECISE && primaryDone)) sf = make_surface(sc, h, O, D, self);
            float4 m = ldg4(sc.mat + sf.matId);
            int mt = __ldg(sc.matType + sf.matId);
            if ((mt & 255) == 3) { sum = sum + beta * (f3(m.x, m.y, m.z) * m.w); continue; }   // emitted, no scatter (materials.js:94-95)
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
                    Hit sh = trace<USE_BVH, COUNT, true, HYBRID, PRIMS_ANY, WIDE, CH>(sc, sf.P, ldir, ldist, h.pid, cnt, sstack, MEGA_BLOCK);
                    if (sh.pid != PID_NONE) continue;
                    sum = sum + beta * (f3(m.x, m.y, m.z) * lcol) * cosN;
                }
            }
            depth++;
            if (depth >= p.maxDepth) continue;                                // the next ray would return black (ray-tracer.js:103): at most maxDepth intersections
            hitPid = h.pid;
            alive = true;                                                     // scatter at the top of the next iteration (sf, D = incoming direction)
        }
        if (STEAL) {
            __syncwarp();
            if (inside) {                                                     // lane L writes pixel slot L of the tile
                const float k = 1.f / 1048576.f;
                const unsigned int* t = tileSum[warp][lane];
                sum = f3((float)((unsigned long long)t[0] + ((unsigned long long)t[1] << 16)) * k, (float)((unsigned long long)t[2] + ((unsigned long long)t[3] << 16)) * k,
                         (float)((unsigned long long)t[4] + ((unsigned long long)t[5] << 16)) * k);
                if ((tileInf[warp] >> lane) & 1u) sum = f3(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
                if ((tileNaN[warp] >> lane) & 1u) sum = f3(CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F);
                pix = (uint32_t)(row0 * p.W + col0);
            }
        }
        if (!STEAL || inside) {
            // each z chunk owns its own plane of the accumulation target (planeStride = 0 when there is one chunk):
            // no atomics, so the sum is deterministic; k_sum_planes folds the planes in fixed order afterwards
            float4* dst = p.accum + (size_t)blockIdx.z * p.planeStride + pix;
            float4 a = *dst;
            a.x += sum.x; a.y += sum.y; a.z += sum.z; a.w += (float)nMine;
            *dst = a;
        }
    }
    if (COUNT) {
        unsigned long long* v = reinterpret_cast<unsigned long long*>(&cnt);
#pragma unroll
        for (int k = 0; k < N_COUNTERS; k++) {
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
    uint32_t pid = PID_NONE; double t64 = 0.0; Surface sf;
    const bool hit = trace_primary64<USE_BVH, true>(p.sc, O64, D64, smem_stack + threadIdx.x, PT_BLOCK, pid, t64, sf);   // the render path's own primary-hit code
    size_t k = (size_t)row * p.W + col;
    if (!hit) {
        objId[k] = -1; triId[k] = -1; tOut[k] = CUDART_INF_F; nrm[3 * k] = nrm[3 * k + 1] = nrm[3 * k + 2] = 0.f; front[k] = 0;
    } else {
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
// the wide hierarchy: fast sampler only (the reference sampler decides primary visibility in float64 over the binary tree)
template <int WIDE, bool COUNT, bool DIRECT>
static cudaError_t launch_wide(const PTParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = (size_t)(p.sc.wideDepth < 1 ? 1 : p.sc.wideDepth) * MEGA_BLOCK * sizeof(uint32_t);   // one entry per level
    if constexpr (!COUNT && !DIRECT) {
        if (p.sc.nBox == 0 && p.sc.nTri == 0) { k_pathtrace_mega<0, true, false, false, false, PRIMS_SPHERE, WIDE><<<grid, MEGA_BLOCK, smem, st>>>(p); return cudaGetLastError(); }
        if (p.sc.nBox == 0 && p.sc.nSph == 0) { k_pathtrace_mega<0, true, false, false, false, PRIMS_TRI, WIDE><<<grid, MEGA_BLOCK, smem, st>>>(p); return cudaGetLastError(); }
    }
    k_pathtrace_mega<0, true, COUNT, DIRECT, false, PRIMS_ANY, WIDE><<<grid, MEGA_BLOCK, smem, st>>>(p);
    return cudaGetLastError();
}
template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT>
static cudaError_t launch_pt3(const PTParams& p, dim3 grid, cudaStream_t st) {
    if constexpr (!USE_BVH) {                              // brute force: no traversal stack at all
        k_pathtrace_mega<SAMPLER, false, COUNT, DIRECT, false><<<grid, MEGA_BLOCK, 0, st>>>(p);
        return cudaGetLastError();
    } else {
        if constexpr (SAMPLER == 0) {
            if (p.sc.wideN == 8 && p.sc.wnodes) return launch_wide<8, COUNT, DIRECT>(p, grid, st);
            if (p.sc.wideN == 4 && p.sc.wnodes) return launch_wide<4, COUNT, DIRECT>(p, grid, st);
        }
        // the stack: depth + 1 entries per thread, all in shared memory, unless the tree is unusually deep
        const bool hybrid = p.sc.bvhStackDepth > SMEM_ONLY_MAX_DEPTH;
        size_t smem = (size_t)(hybrid ? SMEM_STACK : p.sc.bvhStackDepth + 2) * MEGA_BLOCK * sizeof(uint32_t);   // + the sentinel entry
        if (hybrid) { k_pathtrace_mega<SAMPLER, true, COUNT, DIRECT, true><<<grid, MEGA_BLOCK, smem, st>>>(p); return cudaGetLastError(); }
        if constexpr (SAMPLER == 0 && !COUNT && !DIRECT) {  // the common hot configurations get a leaf test without type dispatch
            if (p.sc.nBox == 0 && p.sc.nTri == 0) { k_pathtrace_mega<0, true, false, false, false, PRIMS_SPHERE><<<grid, MEGA_BLOCK, smem, st>>>(p); return cudaGetLastError(); }
            if (p.sc.nBox == 0 && p.sc.nSph == 0) { k_pathtrace_mega<0, true, false, false, false, PRIMS_TRI><<<grid, MEGA_BLOCK, smem, st>>>(p); return cudaGetLastError(); }
        }
        k_pathtrace_mega<SAMPLER, true, COUNT, DIRECT, false><<<grid, MEGA_BLOCK, smem, st>>>(p);
        return cudaGetLastError();
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
    if (p.wavefront && !count && p.maxDepth > 0) return launch_pathtrace_wave(p, sampler, useBvh, zSplit, st);
    dim3 grid((p.W + 15) / 16, (p.H + MEGA_BLOCK / 16 - 1) / (MEGA_BLOCK / 16), zSplit < 1 ? 1 : zSplit);
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
