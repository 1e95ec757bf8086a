// Megakernel path tracer: one thread owns one pixel and regenerates a new camera sample as soon as its
// current path terminates ("persistent per pixel over samples"), so warps stay full while paths of very
// different length (1 bounce into the sky … max_depth bounces between glass spheres) share a warp.
// Replaces the three nested loops of RayTracer.render (ray-tracer.js:189-206) and the recursion of
// rayColor (ray-tracer.js:102-123) — written iteratively: sum += beta ⊙ E; beta ⊙= attenuation.
#include "brt_device.cuh"
#include "brt_kernels.h"

namespace brt {

// ------------------------------------------------------------------------------------------- camera (camera.js:38-51)
template <int SAMPLER>
__device__ __forceinline__ void camera_ray(const PTParams& p, int col, int jUp, uint32_t pix, uint32_t s, RngSeq& rng,
                                           float3& O, float3& D) {
    const DevCamera& c = p.cam;
    float a0 = 0.f, a1 = 0.f, dx, dy;
    if (SAMPLER == 0) {
        uint4 r = philox4x32_10(pix, s, 0u, PHILOX_TAG, p.seedLo, p.seedHi);
        a0 = u01(r.x); a1 = u01(r.y);
        float rr = sqrtf(u01(r.z)), sn, cs;                    // unit disk by inversion (math.js:27-31 distribution)
        sincospif(2.f * u01(r.w), &sn, &cs);
        dx = rr * cs; dy = rr * sn;
    } else {
        rng.init(pix, s, p.seedLo, p.seedHi);
        if (p.aaMode == 1 || p.aaMode == 2) { a0 = rng.next(); a1 = rng.next(); }
        do { dx = rng.next() * 2.f - 1.f; dy = rng.next() * 2.f - 1.f; } while (dx * dx + dy * dy >= 1.0f);   // math.js:29
    }
    float u, v;                                                  // getAntiAliasSample (ray-tracer.js:125-149)
    if (p.aaMode == 1) {
        u = ((float)col + a0) / (float)p.W;
        v = ((float)jUp + a1) / (float)p.H;
    } else if (p.aaMode == 2) {
        float sr = sqrtf(a0), sn, cs;
        sincospif(2.f * a1, &sn, &cs);
        u = ((float)col + 0.5f + sr * cs * 0.5f) / (float)p.W;
        v = ((float)jUp + 0.5f + sr * sn * 0.5f) / (float)p.H;
    } else {
        u = ((float)col + 0.5f) / (float)p.W;
        v = ((float)jUp + 0.5f) / (float)p.H;
    }
    dx *= c.lensRadius; dy *= c.lensRadius;
    float3 off = f3(c.ux * dx + c.vvx * dy, c.uy * dx + c.vvy * dy, c.uz * dx + c.vvz * dy);
    O = f3(c.ox + off.x, c.oy + off.y, c.oz + off.z);
    D = f3(c.llx + u * c.hx + v * c.vx - O.x, c.lly + u * c.hy + v * c.vy - O.y, c.llz + u * c.hz + v * c.vz - O.z);
    if (c.type == 1) D = normalize0(f3(D.x - c.wx, D.y - c.wy, D.z - c.wz));   // camera.js:42-43
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
        Dout = refl + ball * m.w;
        att = f3(m.x, m.y, m.z);
        return dot(Dout, sf.N) > 0.f;
    }
    if (matType == 2) {                                                       // Dielectric (materials.js:51-83)
        float ratio = sf.front ? (1.0f / m.w) : m.w;
        float3 ud = normalize0(Din);
        float cosT = fminf(-dot(ud, sf.N), 1.0f);
        float sinT = sqrtf(fmaxf(0.f, 1.0f - cosT * cosT));
        bool cannot = ratio * sinT > 1.0f;
        bool refl = cannot;
        if (!cannot) {                                                        // the uniform is drawn only here (:62)
            float r0 = (1.f - ratio) / (1.f + ratio); r0 = r0 * r0;
            float c1 = 1.f - cosT, c2 = c1 * c1;
            float R = r0 + (1.f - r0) * (c2 * c2 * c1);
            float xi = SAMPLER == 0 ? u0 : rng.next();
            refl = R > xi;
        }
        if (refl) Dout = reflect(ud, sf.N);
        else {
            float3 perp = (ud + sf.N * cosT) * ratio;
            float3 par = sf.N * (-sqrtf(fabsf(1.0f - dot(perp, perp))));
            Dout = perp + par;
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
// Block = 128 threads = a 16x8 pixel tile; a warp = an 8x4 sub-tile (coherent primary rays).
template <int SAMPLER, bool USE_BVH, bool COUNT, bool DIRECT>
__global__ void __launch_bounds__(PT_BLOCK) k_pathtrace(const __grid_constant__ PTParams p) {
    extern __shared__ uint32_t smem_stack[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    const bool inside = col < p.W && row < p.H;
    Counters cnt = {};
    if (inside) {
        const int jUp = p.H - 1 - row;
        const uint32_t pix = (uint32_t)(row * p.W + col);
        // sample range of this thread: gridDim.z chunks split the launch's samples
        const int per = (p.sCount + gridDim.z - 1) / gridDim.z;
        int s = p.sBegin + blockIdx.z * per;
        const int sEnd = min(p.sBegin + p.sCount, s + per);
        const int nMine = max(0, sEnd - s);

        float3 sum = f3(0.f, 0.f, 0.f), beta = f3(1.f, 1.f, 1.f), O = f3(0, 0, 0), D = f3(0, 0, 1);
        uint32_t self = PID_NONE, cs = 0;
        int depth = 0;
        bool alive = false;
        RngSeq rng;
        uint32_t* sstack = smem_stack + threadIdx.x;
        for (;;) {
            if (!alive) {
                if (s >= sEnd) break;
                cs = (uint32_t)s++;
                camera_ray<SAMPLER>(p, col, jUp, pix, cs, rng, O, D);
                beta = f3(1.f, 1.f, 1.f); self = PID_NONE; depth = 0; alive = true;
            }
            Hit h = trace<USE_BVH, COUNT, false>(p.sc, O, D, CUDART_INF_F, self, cnt, sstack, PT_BLOCK);
            if (h.pid == PID_NONE) {                                          // ray-tracer.js:122
                sum = sum + beta * background(p.sc, D);
                alive = false;
                continue;
            }
            Surface sf = make_surface(p.sc, h, O, D);
            float4 m = ldg4(p.sc.mat + sf.matId);
            int mt = __ldg(p.sc.matType + sf.matId);
            if (mt == 3) sum = sum + beta * (f3(m.x, m.y, m.z) * m.w);        // emitted (materials.js:95)
            if (DIRECT && mt == 0) {
                // EXTENSION (off by default): lights.js:22-47 give direction / colour / distance.
                for (int li = 0; li < p.sc.nLights; li++) {
                    float4 l0 = ldg4(p.sc.lights + 2 * li), l1 = ldg4(p.sc.lights + 2 * li + 1);
                    float3 ldir, lcol = f3(l1.x, l1.y, l1.z); float ldist;
                    if (l0.w == 0.f) {
                        float3 d = f3(l0.x - sf.P.x, l0.y - sf.P.y, l0.z - sf.P.z);
                        ldist = sqrtf(dot(d, d));
                        ldir = normalize0(d);
                        lcol = lcol * (1.0f / (1.0f + 0.1f * ldist + 0.01f * ldist * ldist));
                    } else { ldir = f3(-l0.x, -l0.y, -l0.z); ldist = CUDART_INF_F; }
                    float cosN = dot(sf.N, ldir);
                    if (!(cosN > 0.f)) continue;
                    Hit sh = trace<USE_BVH, COUNT, true>(p.sc, sf.P, ldir, ldist, h.pid, cnt, sstack, PT_BLOCK);
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
    const DevCamera& c = p.cam;
    const int jUp = p.H - 1 - row;
    float u = ((float)col + 0.5f) / (float)p.W, v = ((float)jUp + 0.5f) / (float)p.H;
    float3 O = f3(c.ox, c.oy, c.oz);
    float3 D = f3(c.llx + u * c.hx + v * c.vx - O.x, c.lly + u * c.hy + v * c.vy - O.y, c.llz + u * c.hz + v * c.vz - O.z);
    if (c.type == 1) D = normalize0(f3(D.x - c.wx, D.y - c.wy, D.z - c.wz));
    Counters cnt;
    Hit h = trace<USE_BVH, false, false>(p.sc, O, D, CUDART_INF_F, PID_NONE, cnt, smem_stack + threadIdx.x, PT_BLOCK);
    size_t k = (size_t)row * p.W + col;
    if (h.pid == PID_NONE) {
        objId[k] = -1; triId[k] = -1; tOut[k] = CUDART_INF_F; nrm[3 * k] = nrm[3 * k + 1] = nrm[3 * k + 2] = 0.f; front[k] = 0;
    } else {
        Surface sf = make_surface(p.sc, h, O, D);
        objId[k] = sf.objId; triId[k] = sf.triId; tOut[k] = h.t;
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
    size_t smem = useBvh ? (size_t)SMEM_STACK * PT_BLOCK * sizeof(uint32_t) : 0;
    return sampler == 1 ? launch_pt1<1>(p, useBvh, count, grid, smem, st) : launch_pt1<0>(p, useBvh, count, grid, smem, st);
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
