// Warp-local wavefront path tracer (the alternative integrator; pathtrace.cu holds the default megakernel).
// Every lane owns one pixel and keeps PT_K samples of it in flight; the 32*PT_K path slots of a warp live in shared memory.
// The warp alternates between a shade phase (every lane shades / scatters / regenerates its own slots — all lanes busy) and
// an extend phase (the warp's rays are a queue in shared memory; a lane that finishes a ray fetches the next one, so the
// BVH loop stays full although rays need very different numbers of steps).  No global-memory queues, no launches per
// bounce, no block-wide barriers.  Replaces the same reference loops as the megakernel (ray-tracer.js:102-123, 189-206).
#include "pathtrace_common.cuh"

namespace brt {

// ------------------------------------------------------------------------------------------- the warp-local wavefront kernel
// Block = 128 threads = a 16x8 pixel tile; a warp = an 8x4 sub-tile (coherent primary rays).
// Slot fields (shared memory, [field][slot] per warp, slot = k*32 + owner lane):
// F_STATE packs (camera sample index << 8) | (depth + 2): 0 = dead (no samples left), 1 = needs a camera ray, >= 2 = a path
// at `depth` waiting for / holding a hit.  F_RNGPOS exists only for the sequential (reference) sampler.
enum SlotField : int { F_OX = 0, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_SELF, F_T, F_PID, F_BX, F_BY, F_BZ, F_STATE, F_RNGPOS };
static_assert(F_RNGPOS == PT_SLOT_WORDS, "slot layout");
constexpr int DEPTH_NEED_RAY = -1, DEPTH_DEAD = -2;
constexpr uint32_t SELF_TRACED = 0x7FFFFFFEu;     // F_SELF marker: the slot's hit is already final (PRECISE primary rays are traced in float64 when armed)
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
                uint32_t self = SLOT_U(F_SELF, slot);
                if (self == SELF_TRACED) self = PID_NONE;
                if (SAMPLER == 1) rng.resume(pix, cs, SLOT_U(F_RNGPOS, slot), p.seedLo, p.seedHi);
                bool cont = false;
                if (best.pid == PID_NONE) {                                   // ray-tracer.js:122
                    sum = sum + beta * background(sc, D);
                } else {
                    Surface sf = make_surface(sc, best, O, D, self);
                    if (PRECISE && depth == 0) {                              // primary hit: the float64 evaluation of the primitive trace_primary64 chose
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
                    uint32_t selfMark = PID_NONE;
                    if (PRECISE) {
                        // primary visibility decided in float64, here and now (blocking; the extend-phase stacks are idle during
                        // the shade phase): the slot is marked as traced and the extend phase passes it by
                        D3 O64, D64;
                        camera_ray64(p.cam, p.W, p.H, p.aaMode, col, jUp, cam, O64, D64);
                        O = tof3(O64); D = tof3(D64);
                        uint32_t hp = PID_NONE; double t64 = 0.0; Surface tmp;
                        if (COUNT) cnt.rays++;
                        const bool hit = trace_primary64<USE_BVH, true>(sc, O64, D64, sstack, 32, hp, t64, tmp);
                        SLOT_F(F_T, slot) = hit ? (float)t64 : CUDART_INF_F; SLOT_U(F_PID, slot) = hit ? hp : PID_NONE;
                        selfMark = SELF_TRACED;
                    } else camera_ray32(p.cam32, p.W, p.H, p.aaMode, col, jUp, cam, O, D);
                    SLOT_F(F_OX, slot) = O.x; SLOT_F(F_OY, slot) = O.y; SLOT_F(F_OZ, slot) = O.z;
                    SLOT_F(F_DX, slot) = D.x; SLOT_F(F_DY, slot) = D.y; SLOT_F(F_DZ, slot) = D.z;
                    SLOT_F(F_BX, slot) = 1.f; SLOT_F(F_BY, slot) = 1.f; SLOT_F(F_BZ, slot) = 1.f;
                    SLOT_U(F_SELF, slot) = selfMark;
                    if (SAMPLER == 1) SLOT_U(F_RNGPOS, slot) = rng.pos;
                    depth = 0;
                } else depth = DEPTH_DEAD;
            }
            SLOT_U(F_STATE, slot) = (cs << 8) | (uint32_t)(depth + 2);
            if (depth >= 0) alive = true;
            if (depth >= 0 && SLOT_U(F_SELF, slot) != SELF_TRACED) {
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
                        if (cand < NS && (SLOT_U(F_STATE, cand) & 255u) >= 2u && SLOT_U(F_SELF, cand) != SELF_TRACED) {
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


// ------------------------------------------------------------------------------------------- host launcher
// (the counting build exists only for the megakernel: brt_api.cu routes count_tests renders there)
template <int SAMPLER, bool USE_BVH, bool DIRECT>
static cudaError_t launch_w3(const PTParams& p, dim3 grid, cudaStream_t st) {
    auto go = [&](auto kern, int K) -> cudaError_t {
        size_t smem = (size_t)(PT_BLOCK / 32) * (slot_words(SAMPLER) * 32 * K + 32 * SMEM_STACK) * sizeof(uint32_t);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, PT_BLOCK, smem, st>>>(p);
        return cudaGetLastError();
    };
    switch (p.inflight) {
    case 1: return go(k_pathtrace_wave<SAMPLER, USE_BVH, false, DIRECT, 1>, 1);
    case 3: return go(k_pathtrace_wave<SAMPLER, USE_BVH, false, DIRECT, 3>, 3);
    case 4: return go(k_pathtrace_wave<SAMPLER, USE_BVH, false, DIRECT, 4>, 4);
    default: return go(k_pathtrace_wave<SAMPLER, USE_BVH, false, DIRECT, 2>, 2);
    }
}
template <int SAMPLER>
static cudaError_t launch_w1(const PTParams& p, bool bvh, dim3 grid, cudaStream_t st) {
    if (bvh) return p.directLighting ? launch_w3<SAMPLER, true, true>(p, grid, st) : launch_w3<SAMPLER, true, false>(p, grid, st);
    return p.directLighting ? launch_w3<SAMPLER, false, true>(p, grid, st) : launch_w3<SAMPLER, false, false>(p, grid, st);
}
cudaError_t launch_pathtrace_wave(const PTParams& p, int sampler, bool useBvh, int zSplit, cudaStream_t st) {
    dim3 grid((p.W + 15) / 16, (p.H + 7) / 8, zSplit < 1 ? 1 : zSplit);
    return sampler == 1 ? launch_w1<1>(p, useBvh, grid, st) : launch_w1<0>(p, useBvh, grid, st);
}

}  // namespace brt
