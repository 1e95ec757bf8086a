// Post-processing kernels: resolve (÷spp → tone map → gamma → floor·255 → RGBA8), the fused cross-GPU
// reduce+resolve over peer-mapped accumulation buffers, and the 3x3 gaussian "denoise".
// Reference: ray-tracer.js:208-233 and :266-276, post-processor.js:9-77.  Arithmetic is float64 and this file
// is compiled with -fmad=false so that, given identical inputs, the denoise pass is bit-identical to the
// reference's double arithmetic and the resolve pass differs only by pow()'s last ulp.
// All three kernels are HBM-bound streaming passes: 16 B (float4) loads, 4 B (uchar4) stores, fully coalesced.
#include "brt_kernels.h"

namespace brt {

__device__ __forceinline__ double js_max0(double c) { return c > 0.0 ? c : (c != c ? c : 0.0); }   // Math.max(0, c) incl. NaN
__device__ __forceinline__ unsigned char quantize(double c) {                                       // ray-tracer.js:226-228
    double q = floor(c * 255.0);
    if (q != q) return 0;                    // NaN → Uint8ClampedArray stores 0
    q = q < 0.0 ? 0.0 : q;
    q = q > 255.0 ? 255.0 : q;
    return (unsigned char)q;
}
__device__ __forceinline__ double tonemap1(int kind, double x, double e) {
    if (kind == 1) {                                                                                 // post-processor.js:19-32
        double c = x * e;
        const double a = 2.51, b = 0.03, cc = 2.43, d = 0.59, ee = 0.14;
        double v = (c * (a * c + b)) / (c * (cc * c + d) + ee);
        return js_max0(v);
    }
    if (kind == 2) return x * e;                                                                     // ray-tracer.js:156
    double m = x * e;                                                                                // post-processor.js:9-16
    return m / (1.0 + m);
}

__device__ __forceinline__ void resolve_pixel(const PostParams& pp, float4 a, size_t k, uchar4* rgba, float4* floatData, float4* linear) {
    double n = (double)a.w;
    double r = (double)a.x / n, g = (double)a.y / n, b = (double)a.z / n;                            // ray-tracer.js:208
    if (linear) linear[k] = make_float4((float)r, (float)g, (float)b, 1.0f);
    r = tonemap1(pp.tonemap, r, pp.exposure); g = tonemap1(pp.tonemap, g, pp.exposure); b = tonemap1(pp.tonemap, b, pp.exposure);
    r = pow(js_max0(r), pp.invGamma); g = pow(js_max0(g), pp.invGamma); b = pow(js_max0(b), pp.invGamma);   // post-processor.js:35-42
    if (floatData) floatData[k] = make_float4((float)r, (float)g, (float)b, 1.0f);                   // ray-tracer.js:215-219
    if (rgba) rgba[k] = make_uchar4(quantize(r), quantize(g), quantize(b), 255);
}

__global__ void __launch_bounds__(256) k_resolve(PostParams pp, const float4* __restrict__ accum, uchar4* rgba, float4* floatData,
                                                 float4* linear, size_t begin, size_t end) {
    size_t k = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end) return;
    resolve_pixel(pp, __ldg(accum + k), k, rgba, floatData, linear);
}

struct PeerPtrs { const float4* p[16]; };

// Fused collective + consumer: each rank pulls its row slice from every peer's accumulation buffer over NVLink
// (plain 128-bit loads on peer-mapped pointers), sums in FIXED rank order (deterministic), resolves and writes RGBA8
// (4x smaller than the fp32 sums) straight into the root's output buffer.
__global__ void __launch_bounds__(256) k_reduce_resolve(PostParams pp, PeerPtrs peers, int nPeers, uchar4* rgba, float4* floatData,
                                                        size_t begin, size_t end) {
    size_t k = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= end) return;
    float4 s = peers.p[0][k];
    for (int r = 1; r < nPeers; r++) {
        float4 v = peers.p[r][k];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    resolve_pixel(pp, s, k, rgba, floatData, nullptr);
}

// ---- synchronised cross-GPU exchange: device-side epoch flags in peer-mapped memory (no host or NCCL barrier) -------------
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;   // a peer that never arrives must not hang the GPU

// thread r < nPeers waits for rank r; epochs only grow, so >= (as a signed difference: wrap-safe)
__device__ __forceinline__ void wait_all(const PeerSync& ps, int which) {
    if ((int)threadIdx.x < ps.nPeers) {
        const unsigned* f = ps.flags[threadIdx.x] + which;
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_acquire_sys(f) - ps.epoch) < 0) {
            __nanosleep(200);
            if (globaltimer_ns() - t0 > PEER_TIMEOUT_NS) { atomicExch(ps.flags[ps.self] + FLAG_ERR, 1u); break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_peer_reduce_resolve(PostParams pp, PeerSync ps, uchar4* rgba, float4* floatData, float4* linear,
                                                             size_t begin, size_t end) {
    unsigned* mine = ps.flags[ps.self];
    // the stream ordered this kernel after this rank's path-tracing launch: its sums are complete and, after the fence, visible to peers
    if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); st_release_sys(mine + FLAG_READY, ps.epoch); }
    wait_all(ps, FLAG_READY);
    size_t k = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < end) {
        float4 s = __ldcg(ps.accum[0] + k);
        for (int r = 1; r < ps.nPeers; r++) {                         // fixed rank order: the sum is deterministic
            float4 v = __ldcg(ps.accum[r] + k);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        resolve_pixel(pp, s, k, rgba, floatData, linear);
    }
    __threadfence_system();                                           // this thread's stores into the root's image are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {                                           // ... before the last block publishes DONE
        if (atomicAdd(mine + FLAG_BLOCKS, 1u) == gridDim.x - 1) {
            mine[FLAG_BLOCKS] = 0u;
            __threadfence_system();
            st_release_sys(mine + FLAG_DONE, ps.epoch);
        }
    }
}
__global__ void __launch_bounds__(32) k_peer_wait(PeerSync ps, int which) { wait_all(ps, which); }

// post-processor.js:45-77 — clamp-to-edge 3x3, weights exp(-(kx²+ky²)/(2σ²)), accumulation order ky outer / kx inner.
__global__ void __launch_bounds__(256) k_denoise(PostParams pp, const float4* __restrict__ in, uchar4* rgba, float4* outFloat) {
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= pp.W || y >= pp.H) return;
    double r = 0, g = 0, b = 0, weight = 0;
#pragma unroll
    for (int ky = -1; ky <= 1; ky++) {
#pragma unroll
        for (int kx = -1; kx <= 1; kx++) {
            int nx = max(0, min(pp.W - 1, x + kx)), ny = max(0, min(pp.H - 1, y + ky));
            float4 v = __ldg(in + (size_t)ny * pp.W + nx);
            int d2 = kx * kx + ky * ky;
            double w = d2 == 0 ? 1.0 : d2 == 1 ? pp.w1 : pp.w2;
            r += (double)v.x * w; g += (double)v.y * w; b += (double)v.z * w; weight += w;
        }
    }
    size_t k = (size_t)y * pp.W + x;
    float fr = (float)(r / weight), fg = (float)(g / weight), fb = (float)(b / weight);
    float fa = __ldg(in + k).w;
    if (outFloat) outFloat[k] = make_float4(fr, fg, fb, fa);
    rgba[k] = make_uchar4(quantize((double)fr), quantize((double)fg), quantize((double)fb), 255);   // ray-tracer.js:270-275
}

cudaError_t launch_resolve(const PostParams& pp, const float4* accum, uchar4* rgba, float4* floatData, float4* linear, int rowBegin,
                           int rowEnd, cudaStream_t st) {
    size_t begin = (size_t)rowBegin * pp.W, end = (size_t)rowEnd * pp.W;
    if (end <= begin) return cudaSuccess;
    unsigned blocks = (unsigned)((end - begin + 255) / 256);
    k_resolve<<<blocks, 256, 0, st>>>(pp, accum, rgba, floatData, linear, begin, end);
    return cudaGetLastError();
}
cudaError_t launch_reduce_resolve(const PostParams& pp, const float4* const* peers, int nPeers, uchar4* rgba, float4* floatData,
                                  int rowBegin, int rowEnd, cudaStream_t st) {
    if (nPeers < 1 || nPeers > 16) return cudaErrorInvalidValue;
    PeerPtrs pr;
    for (int i = 0; i < 16; i++) pr.p[i] = i < nPeers ? peers[i] : nullptr;
    size_t begin = (size_t)rowBegin * pp.W, end = (size_t)rowEnd * pp.W;
    if (end <= begin) return cudaSuccess;
    unsigned blocks = (unsigned)((end - begin + 255) / 256);
    k_reduce_resolve<<<blocks, 256, 0, st>>>(pp, pr, nPeers, rgba, floatData, begin, end);
    return cudaGetLastError();
}
cudaError_t launch_peer_reduce_resolve(const PostParams& pp, const PeerSync& ps, uchar4* rgbaRoot, float4* floatRoot, float4* linearRoot,
                                       int rowBegin, int rowEnd, cudaStream_t st) {
    if (ps.nPeers < 1 || ps.nPeers > MAX_PEERS) return cudaErrorInvalidValue;
    size_t begin = (size_t)rowBegin * pp.W, end = (size_t)rowEnd * pp.W;
    // an empty stripe still takes part in the flag protocol (READY / DONE), with one block
    unsigned blocks = end > begin ? (unsigned)((end - begin + 255) / 256) : 1u;
    k_peer_reduce_resolve<<<blocks, 256, 0, st>>>(pp, ps, rgbaRoot, floatRoot, linearRoot, begin, end);
    return cudaGetLastError();
}
cudaError_t launch_peer_wait(const PeerSync& ps, int which, cudaStream_t st) {
    k_peer_wait<<<1, 32, 0, st>>>(ps, which);
    return cudaGetLastError();
}
cudaError_t launch_denoise(const PostParams& pp, const float4* floatData, uchar4* rgba, float4* outFloat, cudaStream_t st) {
    dim3 grid((pp.W + 31) / 32, (pp.H + 7) / 8);
    k_denoise<<<grid, 256, 0, st>>>(pp, floatData, rgba, outFloat);
    return cudaGetLastError();
}

}  // namespace brt
