// Device code shared by the two integrators (pathtrace.cu: megakernel + AOV / utility kernels; pathtrace_wave.cu: the
// warp-local wavefront): camera sampling (camera.js), materials (materials.js) and the blocking trace wrapper.
#pragma once
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
__device__ __forceinline__ void camera_ray32(const DevCamera32& c, int W, int H, int aaMode, int col, int jUp, CamSample cs, float3& O, float3& D) {
    const float iw = rcpf((float)W), ih = rcpf((float)H);
    float u, v;
    if (aaMode == 1) { u = __fmul_rn((float)col + cs.s, iw); v = __fmul_rn((float)jUp + cs.t, ih); }
    else if (aaMode == 2) { u = __fmul_rn(fmaf(cs.s, 0.5f, (float)col + 0.5f), iw); v = __fmul_rn(fmaf(cs.t, 0.5f, (float)jUp + 0.5f), ih); }
    else { u = __fmul_rn((float)col + 0.5f, iw); v = __fmul_rn((float)jUp + 0.5f, ih); }
    float rx = __fmul_rn(cs.dx, c.lensRadius), ry = __fmul_rn(cs.dy, c.lensRadius);
    float3 off = madd(f3(c.cv[0], c.cv[1], c.cv[2]), ry, f3(c.cu[0], c.cu[1], c.cu[2]) * rx);
    O = f3(c.o[0], c.o[1], c.o[2]) + off;
    // D = (LLC - origin) + u*H + v*V - off
    D = madd(f3(c.v[0], c.v[1], c.v[2]), v, madd(f3(c.h[0], c.h[1], c.h[2]), u, f3(c.llo[0], c.llo[1], c.llo[2]))) - off;
    if (c.type == 1) D = normalize0(D - f3(c.cw[0], c.cw[1], c.cw[2]));
}

// `r` = the Philox block (pixel, sample, 0) of the fast sampler, drawn by the caller (the megakernel draws the camera block of
// the lanes that start a path and the scatter block of the lanes that continue one in ONE call: both are the same code).
template <int SAMPLER>
__device__ __forceinline__ CamSample camera_sample_drawn(const PTParams& p, uint4 r, uint32_t pix, uint32_t s, RngSeq& rng) {
    CamSample cs; cs.s = 0.f; cs.t = 0.f;
    float a0 = 0.f, a1 = 0.f;
    if (SAMPLER == 0) {
        a0 = u01(r.x); a1 = u01(r.y);
        float rr = sqrtf(u01(r.z)), sn, cs_;                   // unit disk by inversion (math.js:27-31 distribution)
        fast_sincos2pi(u01(r.w), &sn, &cs_);
        cs.dx = __fmul_rn(rr, cs_); cs.dy = __fmul_rn(rr, sn);
    } else {
        rng.init(pix, s, p.seedLo, p.seedHi);
        if (p.aaMode == 1 || p.aaMode == 2) { a0 = rng.next(); a1 = rng.next(); }
        do { cs.dx = rng.next() * 2.f - 1.f; cs.dy = rng.next() * 2.f - 1.f; } while (fmaf(cs.dx, cs.dx, __fmul_rn(cs.dy, cs.dy)) >= 1.0f);   // math.js:29
    }
    if (p.aaMode == 1) { cs.s = a0; cs.t = a1; }
    else if (p.aaMode == 2) {                                    // stochastic: disk of radius 0.5 about the pixel centre
        float sr = sqrtf(a0), sn, c2;
        if (SAMPLER == 0) fast_sincos2pi(a1, &sn, &c2); else sincospif(__fmul_rn(2.f, a1), &sn, &c2);
        cs.s = __fmul_rn(sr, c2); cs.t = __fmul_rn(sr, sn);
    }
    return cs;
}
template <int SAMPLER>
__device__ __forceinline__ CamSample camera_sample(const PTParams& p, uint32_t pix, uint32_t s, RngSeq& rng) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (SAMPLER == 0) r = philox_fast(pix, s, 0u, PHILOX_TAG, p.seedLo, p.seedHi);
    return camera_sample_drawn<SAMPLER>(p, r, pix, s, rng);
}

// ------------------------------------------------------------------------------------------- materials (materials.js)
// Returns false when the path ends here (emissive, absorbed metal).  `att` multiplies the throughput.
// `r` = the Philox block (pixel, sample, bounce + 1) of the fast sampler, drawn by the caller (see camera_sample_drawn).
template <int SAMPLER>
__device__ __forceinline__ bool scatter_drawn(const PTParams& p, int matWord, float4 m, const Surface& sf, float3 Din, uint4 r,
                                              RngSeq& rng, float3& Dout, float3& att) {
    const int matType = matWord & 255, tex = matWord >> 8;           // 1-based texture index above the type (materials.js:99-126)
    float u0 = 0.f, u1 = 0.f, u2 = 0.f;
    if (SAMPLER == 0) { u0 = u01(r.x); u1 = u01(r.y); u2 = u01(r.z); }
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
template <int SAMPLER>
__device__ __forceinline__ bool scatter(const PTParams& p, int matWord, float4 m, const Surface& sf, float3 Din, uint32_t pix,
                                        uint32_t s, int bounce, RngSeq& rng, float3& Dout, float3& att) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (SAMPLER == 0) r = philox_fast(pix, s, (uint32_t)(bounce + 1), PHILOX_TAG, p.seedLo, p.seedHi);
    return scatter_drawn<SAMPLER>(p, matWord, m, sf, Din, r, rng, Dout, att);
}

// WIDE = 0: the binary hierarchy (or the linear loops); 4 / 8: the wide hierarchy collapsed from it (megakernel, fast sampler)
// CH: walk the centre / half-extent copy of the binary nodes (node_visit_ch) — the megakernel with the fast sampler
template <bool USE_BVH, bool COUNT, bool SHADOW, bool HYBRID = true, int PRIMS = PRIMS_ANY, int WIDE = 0, bool CH = false>
__device__ __forceinline__ Hit trace(const DevScene& sc, float3 O, float3 D, float tMax, uint32_t self, Counters& cnt,
                                     uint32_t* sstack, int sstride, unsigned aliveMask = 0xffffffffu) {
    if (COUNT && !SHADOW) cnt.rays++;
    if (USE_BVH && WIDE == 8) return trace_wide<8, COUNT, SHADOW, PRIMS>(sc, O, D, 0.001f, tMax, self, cnt, sstack, sstride, aliveMask);
    if (USE_BVH && WIDE == 4) return trace_wide<4, COUNT, SHADOW, PRIMS>(sc, O, D, 0.001f, tMax, self, cnt, sstack, sstride, aliveMask);
    if (USE_BVH) return trace_bvh<COUNT, SHADOW, HYBRID, PRIMS, CH>(sc, O, D, 0.001f, tMax, self, cnt, sstack, sstride, aliveMask);
    return trace_brute<COUNT, SHADOW>(sc, O, D, 0.001f, tMax, self, cnt);
}


}  // namespace brt
