// GPU LBVH build over the bounded primitives (spheres, boxes, triangles):
//   k_prim_bounds → k_scene_bounds → k_morton → 4-pass LSD radix sort (k_sort_hist / k_sort_scan / k_sort_scatter)
//   → k_karras (Karras 2012 hierarchy from sorted 30-bit Morton codes, index tie-break) → k_refit (bottom-up AABBs).
// The reference has no acceleration structure (README.md:102 lists it as future work; world.js:24-30 and
// geometry.js:253-259 are linear loops); the BVH must therefore be invisible in the results: boxes are padded
// and traversal (brt_device.cuh) resolves exact ties by the reference's loop-order rule.
// Node layout (64 B): n0 = child0 (min.x,max.x,min.y,max.y); n1 = child1 (same); n2 = (c0.min.z,c0.max.z,c1.min.z,c1.max.z);
// n3 = (child0, child1, 0, 0) bit patterns, LEAF_BIT | pid for leaves.
#include "brt_kernels.h"
#include <algorithm>
#include <cfloat>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace brt {

struct Aabb { float mn[3], mx[3]; };
static inline float __uint_as_float_host(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

__device__ __forceinline__ Aabb prim_aabb(const DevScene& sc, uint32_t pid) {
    uint32_t ty = pid_type(pid), ix = pid_index(pid);
    Aabb b;
    if (ty == PT_SPHERE) {
        float4 s = sc.sph[ix]; float r = fabsf(s.w);
        b.mn[0] = s.x - r; b.mn[1] = s.y - r; b.mn[2] = s.z - r; b.mx[0] = s.x + r; b.mx[1] = s.y + r; b.mx[2] = s.z + r;
    } else if (ty == PT_BOX) {
        float4 a = sc.box[2 * ix], c = sc.box[2 * ix + 1];
        b.mn[0] = fminf(a.x, c.x); b.mn[1] = fminf(a.y, c.y); b.mn[2] = fminf(a.z, c.z);
        b.mx[0] = fmaxf(a.x, c.x); b.mx[1] = fmaxf(a.y, c.y); b.mx[2] = fmaxf(a.z, c.z);
    } else {
        float4 v0 = sc.tri[3 * ix], e1 = sc.tri[3 * ix + 1], e2 = sc.tri[3 * ix + 2];
        float p1[3] = { v0.x + e1.x, v0.y + e1.y, v0.z + e1.z }, p2[3] = { v0.x + e2.x, v0.y + e2.y, v0.z + e2.z };
        float p0[3] = { v0.x, v0.y, v0.z };
        for (int k = 0; k < 3; k++) { b.mn[k] = fminf(p0[k], fminf(p1[k], p2[k])); b.mx[k] = fmaxf(p0[k], fmaxf(p1[k], p2[k])); }
    }
    // conservative padding: fp32 slab / primitive-test rounding must never cull a primitive the brute-force loop would hit
    float ext = fmaxf(b.mx[0] - b.mn[0], fmaxf(b.mx[1] - b.mn[1], b.mx[2] - b.mn[2]));
    for (int k = 0; k < 3; k++) {
        float pad = 1e-5f * fmaxf(fmaxf(fabsf(b.mn[k]), fabsf(b.mx[k])), ext) + 1e-7f;
        b.mn[k] -= pad; b.mx[k] += pad;
    }
    return b;
}
__device__ __forceinline__ uint32_t bounded_pid(const DevScene& sc, int i) {
    if (i < sc.nSph) return make_pid(PT_SPHERE, i);
    i -= sc.nSph;
    if (i < sc.nBox) return make_pid(PT_BOX, i);
    return make_pid(PT_TRI, i - sc.nBox);
}

__global__ void k_prim_bounds(DevScene sc, int n, Aabb* boxes, float* blockBounds /* 6 per block */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float c[6] = { FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX };
    if (i < n) {
        Aabb b = prim_aabb(sc, bounded_pid(sc, i));
        boxes[i] = b;
        for (int k = 0; k < 3; k++) { float m = 0.5f * (b.mn[k] + b.mx[k]); c[k] = m; c[3 + k] = m; }
    }
    __shared__ float red[6][8];
    for (int k = 0; k < 6; k++) {
        float v = c[k];
        for (int o = 16; o > 0; o >>= 1) { float w = __shfl_down_sync(0xffffffffu, v, o); v = k < 3 ? fminf(v, w) : fmaxf(v, w); }
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int k = threadIdx.x; float v = red[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) v = k < 3 ? fminf(v, red[k][w]) : fmaxf(v, red[k][w]);
        blockBounds[6 * blockIdx.x + k] = v;
    }
}
__global__ void k_scene_bounds(const float* blockBounds, int nBlocks, float* out6) {
    __shared__ float red[6][256];
    for (int k = 0; k < 6; k++) {
        float v = k < 3 ? FLT_MAX : -FLT_MAX;
        for (int b = threadIdx.x; b < nBlocks; b += blockDim.x) { float w = blockBounds[6 * b + k]; v = k < 3 ? fminf(v, w) : fmaxf(v, w); }
        red[k][threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int k = threadIdx.x; float v = red[k][0];
        for (int i = 1; i < (int)blockDim.x; i++) v = k < 3 ? fminf(v, red[k][i]) : fmaxf(v, red[k][i]);
        out6[k] = v;
    }
}
__device__ __forceinline__ uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_morton(const Aabb* boxes, int n, const float* sb, int uniformScale, uint32_t* keys, uint32_t* vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Aabb b = boxes[i];
    uint32_t q[3];
    // uniformScale = 1: one scale for all three axes (the largest scene extent) — a thin axis, a terrain's height, then
    // occupies few Morton cells and only splits at fine levels instead of cutting every third level through geometry that
    // is flat there.  uniformScale = 0: every axis normalised to its own extent.  build_lbvh builds both, keeps the cheaper.
    float extMax = fmaxf(sb[3] - sb[0], fmaxf(sb[4] - sb[1], sb[5] - sb[2]));
    for (int k = 0; k < 3; k++) {
        float ext = uniformScale ? extMax : sb[3 + k] - sb[k];
        float x = ext > 0.f ? (0.5f * (b.mn[k] + b.mx[k]) - sb[k]) / ext : 0.f;
        q[k] = (uint32_t)fminf(fmaxf(x * 1024.f, 0.f), 1023.f);
    }
    keys[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    vals[i] = (uint32_t)i;
}

// ---- LSD radix sort, 8 bits per pass, stable ------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                       // keys per thread per tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint32_t* keys, int n, int shift, uint32_t* hist, int nBlocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_TILE;
    for (int k = 0; k < SORT_ITEMS; k++) {
        int i = base + k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * nBlocks + blockIdx.x] = h[threadIdx.x];     // digit-major so one scan orders (digit, block)
}
__global__ void __launch_bounds__(1024) k_sort_scan(uint32_t* hist, int total) {
    // single-block exclusive scan (total = 256 * nBlocks, tens of thousands of entries at most)
    __shared__ uint32_t warpSums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = i < total ? hist[i] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warpSums[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = warpSums[threadIdx.x], s = w;
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
            warpSums[threadIdx.x] = s - w;
        }
        __syncthreads();
        uint32_t excl = x - v + warpSums[threadIdx.x >> 5] + carry;
        if (i < total) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t* keysIn, const uint32_t* valsIn, uint32_t* keysOut,
                                                               uint32_t* valsOut, int n, int shift, const uint32_t* hist, int nBlocks) {
    __shared__ uint32_t offs[256];                   // running global offset of each digit for this block
    __shared__ uint32_t wcount[SORT_THREADS / 32][256];
    offs[threadIdx.x] = hist[threadIdx.x * nBlocks + blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = blockIdx.x * SORT_TILE;
    for (int k = 0; k < SORT_ITEMS; k++) {
        for (int w = 0; w < SORT_THREADS / 32; w++) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        int i = base + k * SORT_THREADS + threadIdx.x;
        bool valid = i < n;
        uint32_t key = valid ? keysIn[i] : 0xFFFFFFFFu, val = valid ? valsIn[i] : 0u;
        uint32_t d = (key >> shift) & 255u;
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFu);
        uint32_t rankInWarp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rankInWarp == 0) wcount[warp][d] = __popc(peers);
        __syncthreads();
        if (valid) {
            uint32_t before = 0;
            for (int w = 0; w < warp; w++) before += wcount[w][d];
            uint32_t dst = offs[d] + before + rankInWarp;
            keysOut[dst] = key; valsOut[dst] = val;
        }
        __syncthreads();
        uint32_t tot = 0;
        for (int w = 0; w < SORT_THREADS / 32; w++) tot += wcount[w][threadIdx.x];
        offs[threadIdx.x] += tot;
        __syncthreads();
    }
}

// ---- Karras hierarchy -------------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint32_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = keys[i], b = keys[j];
    if (a == b) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(a ^ b);
}
__global__ void k_karras(const uint32_t* keys, int n, int2* children /* n-1 */, int* parent /* 2n-1: internal [0,n-1), leaves [n-1,2n-1) */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? (n - 1 + gamma) : gamma;            // leaf ids are offset by n-1
    int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left] = i; parent[right] = i;
    if (i == 0) parent[0] = -1;
}
__global__ void k_refit(const DevScene sc, int n, const uint32_t* sortedPrim, const Aabb* primBoxes, const int2* children, const int* parent,
                        Aabb* nodeBox /* n-1 */, int* nodeDepth /* n-1 */, unsigned int* flags /* n-1, zeroed */, float4* nodes) {
    int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int cur = parent[n - 1 + leaf];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1u) == 0u) return;               // first child to arrive stops; the second continues
        __threadfence();
        int2 ch = children[cur];
        Aabb b[2]; int dep[2]; uint32_t ref[2];
        int cc[2] = { ch.x, ch.y };
        for (int k = 0; k < 2; k++) {
            if (cc[k] >= n - 1) {
                uint32_t pi = sortedPrim[cc[k] - (n - 1)];
                b[k] = primBoxes[pi]; dep[k] = 0; ref[k] = LEAF_BIT | bounded_pid(sc, (int)pi);
            } else {
                // written by another thread before its atomicAdd + fence: read through L2
                const volatile Aabb* vb = nodeBox + cc[k];
                for (int a = 0; a < 3; a++) { b[k].mn[a] = vb->mn[a]; b[k].mx[a] = vb->mx[a]; }
                dep[k] = ((const volatile int*)nodeDepth)[cc[k]]; ref[k] = (uint32_t)cc[k];
            }
        }
        float4* np = nodes + 4 * (size_t)cur;
        np[0] = make_float4(b[0].mn[0], b[0].mx[0], b[0].mn[1], b[0].mx[1]);
        np[1] = make_float4(b[1].mn[0], b[1].mx[0], b[1].mn[1], b[1].mx[1]);
        np[2] = make_float4(b[0].mn[2], b[0].mx[2], b[1].mn[2], b[1].mx[2]);
        np[3] = make_float4(__uint_as_float(ref[0]), __uint_as_float(ref[1]), 0.f, 0.f);
        Aabb u;
        for (int a = 0; a < 3; a++) { u.mn[a] = fminf(b[0].mn[a], b[1].mn[a]); u.mx[a] = fmaxf(b[0].mx[a], b[1].mx[a]); }
        nodeBox[cur] = u;
        nodeDepth[cur] = 1 + max(dep[0], dep[1]);
        __threadfence();
        cur = parent[cur];
    }
}

// Surface-area cost of a hierarchy: sum of the internal nodes' box areas (leaves hold one primitive each, so the leaf term
// is the same for every candidate tree).  Per-block partial sums, folded on the host in fixed order (deterministic choice).
__global__ void __launch_bounds__(256) k_sah_cost(const Aabb* nodeBox, int nInternal, double* blockSums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0;
    if (i < nInternal) {
        Aabb b = nodeBox[i];
        double dx = (double)b.mx[0] - b.mn[0], dy = (double)b.mx[1] - b.mn[1], dz = (double)b.mx[2] - b.mn[2];
        a = dx * dy + dy * dz + dz * dx;
    }
    __shared__ double red[256];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) blockSums[blockIdx.x] = red[0];
}

// ---- small scenes: a full-sweep SAH hierarchy built on the host -------------------------------------------------------
// For up to SAH_HOST_MAX bounded primitives (a few hundred objects placed by hand or by the exporter — the reference's own
// scale) a top-down surface-area-heuristic build over the SAME padded device AABBs costs well under a millisecond on one
// host core and gives a markedly better tree than any Morton-order LBVH.  It is offered as a third candidate and ranked by
// the same surface-area cost; large meshes stay with the GPU LBVH.  Same node format, same single-primitive leaves, so the
// traversal (and therefore every result) is unchanged.
constexpr int SAH_HOST_MAX = 16384;
struct HostSah {
    const Aabb* box; const uint32_t* pid; int n;
    std::vector<float4> nodes; double cost = 0.0;
    // the primitives of the current range, sorted by centroid along each axis (presorted once, then stably partitioned at every
    // split: O(n log n) for the whole build instead of four sorts per node)
    std::vector<int> ord[3], tmp; std::vector<unsigned char> side; std::vector<double> right;
    HostSah(const Aabb* b, const uint32_t* p, int count) : box(b), pid(p), n(count), tmp(count), side(count), right(count) {
        for (int ax = 0; ax < 3; ax++) {
            ord[ax].resize(n);
            for (int i = 0; i < n; i++) ord[ax][i] = i;
            std::stable_sort(ord[ax].begin(), ord[ax].end(), [&](int x, int y) { return box[x].mn[ax] + box[x].mx[ax] < box[y].mn[ax] + box[y].mx[ax]; });
        }
        nodes.reserve(4 * (size_t)(n > 1 ? n - 1 : 1));
    }
    static double area(const Aabb& b) {
        double dx = (double)b.mx[0] - b.mn[0], dy = (double)b.mx[1] - b.mn[1], dz = (double)b.mx[2] - b.mn[2];
        return dx * dy + dy * dz + dz * dx;
    }
    static void grow(Aabb& a, const Aabb& b) { for (int k = 0; k < 3; k++) { a.mn[k] = std::min(a.mn[k], b.mn[k]); a.mx[k] = std::max(a.mx[k], b.mx[k]); } }
    static Aabb empty() { Aabb e; for (int k = 0; k < 3; k++) { e.mn[k] = FLT_MAX; e.mx[k] = -FLT_MAX; } return e; }
    // builds the subtree over positions [lo, hi) of the three sorted lists (hi - lo >= 2; the same SET of primitives in each);
    // returns its node index, its box in `out`, its height in `h`
    int build(int lo, int hi, Aabb& out, int& h) {
        const int m = hi - lo, me = (int)(nodes.size() / 4);
        nodes.resize(nodes.size() + 4);
        int bestAxis = 0, bestK = m / 2; double bestC = DBL_MAX;
        for (int ax = 0; ax < 3; ax++) {
            const int* idx = ord[ax].data() + lo;
            Aabb acc = empty();
            for (int k = m - 1; k >= 1; k--) { grow(acc, box[idx[k]]); right[k] = area(acc); }
            acc = empty();
            for (int k = 1; k < m; k++) {
                grow(acc, box[idx[k - 1]]);
                double c = area(acc) * k + right[k] * (m - k);
                // ties (identical boxes) go to the most balanced split so duplicates cannot degenerate into a chain
                if (c < bestC || (c == bestC && std::abs(k - m / 2) < std::abs(bestK - m / 2))) { bestC = c; bestAxis = ax; bestK = k; }
            }
        }
        for (int k = 0; k < m; k++) side[ord[bestAxis][lo + k]] = k < bestK ? 0 : 1;
        for (int ax = 0; ax < 3; ax++) {
            if (ax == bestAxis) continue;
            int* idx = ord[ax].data() + lo;
            int l = 0, r = 0;
            for (int k = 0; k < m; k++) { if (side[idx[k]]) tmp[r++] = idx[k]; else idx[l++] = idx[k]; }
            for (int k = 0; k < r; k++) idx[l + k] = tmp[k];
        }
        Aabb cb[2]; uint32_t ref[2]; int ch[2] = { 0, 0 };
        const int range[2][2] = { { lo, lo + bestK }, { lo + bestK, hi } };
        for (int c = 0; c < 2; c++) {
            if (range[c][1] - range[c][0] == 1) { int pi = ord[0][range[c][0]]; cb[c] = box[pi]; ref[c] = LEAF_BIT | pid[pi]; }
            else { ref[c] = (uint32_t)build(range[c][0], range[c][1], cb[c], ch[c]); }
        }
        float4* np = &nodes[4 * (size_t)me];
        np[0] = make_float4(cb[0].mn[0], cb[0].mx[0], cb[0].mn[1], cb[0].mx[1]);
        np[1] = make_float4(cb[1].mn[0], cb[1].mx[0], cb[1].mn[1], cb[1].mx[1]);
        np[2] = make_float4(cb[0].mn[2], cb[0].mx[2], cb[1].mn[2], cb[1].mx[2]);
        np[3] = make_float4(__uint_as_float_host(ref[0]), __uint_as_float_host(ref[1]), 0.f, 0.f);
        out = cb[0]; grow(out, cb[1]);
        cost += area(out);
        h = 1 + std::max(ch[0], ch[1]);
        return me;
    }
};

// ---- centre / half-extent copy of the binary nodes (node_visit_ch) -------------------------------------------------------
// c = (lo + hi) / 2 rounded to nearest, h = max(hi - c, c - lo) widened by 4 ulps: [c - h, c + h] contains [lo, hi] whatever
// the rounding of the three operations, so the copy is at least as conservative as the original.
__global__ void __launch_bounds__(256) k_centre_half(const float4* __restrict__ nodes, int nInternal, float4* __restrict__ cnodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nInternal) return;
    const float4 n0 = nodes[4 * (size_t)i], n1 = nodes[4 * (size_t)i + 1], n2 = nodes[4 * (size_t)i + 2], n3 = nodes[4 * (size_t)i + 3];
    const float lo[2][3] = { { n0.x, n0.z, n2.x }, { n1.x, n1.z, n2.z } }, hi[2][3] = { { n0.y, n0.w, n2.y }, { n1.y, n1.w, n2.w } };
    float c[2][3], h[2][3];
    for (int k = 0; k < 2; k++)
        for (int a = 0; a < 3; a++) {
            c[k][a] = 0.5f * lo[k][a] + 0.5f * hi[k][a];
            h[k][a] = fmaxf(hi[k][a] - c[k][a], c[k][a] - lo[k][a]) * 1.0000005f + 1e-37f;
        }
    float4* cp = cnodes + 4 * (size_t)i;
    cp[0] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);          // children interleaved per axis: one 64-bit register
    cp[1] = make_float4(c[0][2], c[1][2], h[0][0], h[1][0]);          // pair = the same quantity of both children (FFMA2)
    cp[2] = make_float4(h[0][1], h[1][1], h[0][2], h[1][2]);
    cp[3] = make_float4(n3.x, n3.y, 0.f, 0.f);
}

// ---- wide hierarchy: N-wide collapse of the binary tree --------------------------------------------------------------
// One launch per level, top-down (launch L builds the wide nodes whose level is L and marks their internal children L + 1).
// The wide node of binary node b lives at index b, so there is no allocation, no atomics and the result is deterministic;
// only nodes reachable from the root are written.  Collapse rule: start from b's two children and keep replacing the internal
// child with the largest surface area by its own two children until there are N children or only leaves are left (the
// surface-area-guided collapse of Wald et al. 2008 / Ylitie et al. 2017).  Child boxes are copied from the binary nodes
// unchanged.  Slot assignment: greedy maximisation of sum_i dot(centroid_i - centre, sign vector of slot_i) — see trace_wide.
template <int N>
__global__ void __launch_bounds__(128) k_wide_level(const float4* __restrict__ nodes, int nInternal, int level, int* __restrict__ levelOf,
                                                    float4* __restrict__ wnodes, int axes, int* __restrict__ maxLevel) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nInternal) return;
    if (level == 1 ? b != 0 : levelOf[b] != level) return;
    uint32_t ref[N]; float lo[N][3], hi[N][3];
    auto expand = [&](uint32_t id, int i0, int i1) {
        const float4* np = nodes + 4 * (size_t)id;
        const float4 n0 = np[0], n1 = np[1], n2 = np[2], n3 = np[3];
        lo[i0][0] = n0.x; hi[i0][0] = n0.y; lo[i0][1] = n0.z; hi[i0][1] = n0.w; lo[i0][2] = n2.x; hi[i0][2] = n2.y; ref[i0] = __float_as_uint(n3.x);
        lo[i1][0] = n1.x; hi[i1][0] = n1.y; lo[i1][1] = n1.z; hi[i1][1] = n1.w; lo[i1][2] = n2.z; hi[i1][2] = n2.w; ref[i1] = __float_as_uint(n3.y);
    };
    expand((uint32_t)b, 0, 1);
    int cnt = 2;
    while (cnt < N) {
        int pick = -1; float bestA = -1.f;
        for (int i = 0; i < cnt; i++) {
            if (ref[i] & LEAF_BIT) continue;
            const float dx = hi[i][0] - lo[i][0], dy = hi[i][1] - lo[i][1], dz = hi[i][2] - lo[i][2];
            const float a = dx * dy + dy * dz + dz * dx;
            if (a > bestA) { bestA = a; pick = i; }
        }
        if (pick < 0) break;
        expand(ref[pick], pick, cnt);
        cnt++;
    }
    // slot assignment
    float ctr[3];
    for (int a = 0; a < 3; a++) {
        float mn = lo[0][a], mx = hi[0][a];
        for (int i = 1; i < cnt; i++) { mn = fminf(mn, lo[i][a]); mx = fmaxf(mx, hi[i][a]); }
        ctr[a] = 0.5f * (mn + mx);
    }
    constexpr int NB = N == 8 ? 3 : 2;
    int ax[3] = { 0, 1, 2 };
    if (N == 4) { ax[0] = axes & 3; ax[1] = (axes >> 2) & 3; }
    int slotOf[N], childAt[N];
    for (int i = 0; i < N; i++) { slotOf[i] = -1; childAt[i] = -1; }
    for (int round = 0; round < cnt; round++) {
        int bi = -1, bs = -1; float bc = -CUDART_INF_F;
        for (int i = 0; i < cnt; i++) {
            if (slotOf[i] >= 0) continue;
            for (int s2 = 0; s2 < N; s2++) {
                if (childAt[s2] >= 0) continue;
                float c = 0.f;
                for (int k = 0; k < NB; k++) {
                    const float d = 0.5f * (lo[i][ax[k]] + hi[i][ax[k]]) - ctr[ax[k]];
                    c += ((s2 >> k) & 1) ? d : -d;
                }
                if (c > bc) { bc = c; bi = i; bs = s2; }
            }
        }
        slotOf[bi] = bs; childAt[bs] = bi;
    }
    // node = N / 2 pair records of 48 bytes (slots 2p, 2p + 1 interleaved per axis: node_visit_ch's box layout) + the N refs
    float4* wp = wnodes + (size_t)b * (2 * N);
    uint32_t* refs = reinterpret_cast<uint32_t*>(wp) + 6 * N;           // byte offset 24 N
    bool anyInternal = false;
    for (int pr = 0; pr < N / 2; pr++) {
        float c[2][3], h[2][3];
        for (int j = 0; j < 2; j++) {
            const int i = childAt[2 * pr + j];
            if (i < 0) {
                // an empty slot: h.x = -inf makes the near plane +inf and the far plane -inf whatever the ray, so it is never hit
                c[j][0] = c[j][1] = c[j][2] = 0.f; h[j][0] = -CUDART_INF_F; h[j][1] = h[j][2] = 0.f; refs[2 * pr + j] = 0xFFFFFFFFu;
                continue;
            }
            for (int a = 0; a < 3; a++) {
                c[j][a] = 0.5f * lo[i][a] + 0.5f * hi[i][a];
                h[j][a] = fmaxf(hi[i][a] - c[j][a], c[j][a] - lo[i][a]) * 1.0000005f + 1e-37f;     // as k_centre_half: [c - h, c + h] contains [lo, hi]
            }
            refs[2 * pr + j] = ref[i];
            if (!(ref[i] & LEAF_BIT)) { levelOf[ref[i]] = level + 1; anyInternal = true; }
        }
        wp[3 * pr + 0] = make_float4(c[0][0], c[1][0], c[0][1], c[1][1]);
        wp[3 * pr + 1] = make_float4(c[0][2], c[1][2], h[0][0], h[1][0]);
        wp[3 * pr + 2] = make_float4(h[0][1], h[1][1], h[0][2], h[1][2]);
    }
    if (anyInternal) atomicMax(maxLevel, level + 1);
}

cudaError_t build_wide(const BvhBuildResult& bin, int width, BvhWorkspace* ws, WideBuildResult* out, cudaStream_t st) {
    *out = WideBuildResult{};
    if (!bin.nodes || bin.nNodes < 1 || (width != 4 && width != 8)) return cudaSuccess;
    const int nInternal = (int)bin.nNodes;
    const size_t bytes = (size_t)nInternal * 32 * (size_t)width, lvlBytes = 4 * ((size_t)nInternal + 1);
    cudaError_t e;
    if (bytes > ws->wideCap) {
        cudaFree(ws->wide); ws->wide = nullptr; ws->wideCap = 0;
        if ((e = cudaMalloc(&ws->wide, bytes)) != cudaSuccess) return e;
        ws->wideCap = bytes;
    }
    if (lvlBytes > ws->levelCap) {
        cudaFree(ws->level); ws->level = nullptr; ws->levelCap = 0;
        if ((e = cudaMalloc(&ws->level, lvlBytes)) != cudaSuccess) return e;
        ws->levelCap = lvlBytes;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    cudaEventRecord(e0, st);
    int* levelOf = ws->level;                         // [nInternal] level of every wide node (0 = not part of the wide tree), then maxLevel
    int* maxLevel = levelOf + nInternal;
    cudaMemsetAsync(levelOf, 0, lvlBytes, st);
    // the N = 4 slot order uses the two axes along which the scene is widest
    int axes = 0 | (2 << 2);
    {
        int order[3] = { 0, 1, 2 };
        std::sort(order, order + 3, [&](int a, int b2) { return bin.extent[a] > bin.extent[b2] || (bin.extent[a] == bin.extent[b2] && a < b2); });
        const int a0 = std::min(order[0], order[1]), a1 = std::max(order[0], order[1]);
        axes = a0 | (a1 << 2);
    }
    const int blocks = (nInternal + 127) / 128;
    // a wide level consumes at least one binary level, so `depth` launches reach every node (extra launches find nothing to do)
    for (int level = 1; level <= bin.depth; level++) {
        if (width == 8) k_wide_level<8><<<blocks, 128, 0, st>>>(bin.nodes, nInternal, level, levelOf, ws->wide, axes, maxLevel);
        else k_wide_level<4><<<blocks, 128, 0, st>>>(bin.nodes, nInternal, level, levelOf, ws->wide, axes, maxLevel);
    }
    int maxL = 0;
    cudaMemcpyAsync(&maxL, maxLevel, 4, cudaMemcpyDeviceToHost, st);
    cudaEventRecord(e1, st);
    e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    out->wnodes = ws->wide; out->width = width; out->depth = maxL < 1 ? 1 : maxL; out->axes = axes; out->buildMs = ms;
    if (getenv("BRT_DEBUG")) {
        // walk the wide tree on the host: structure check + shape statistics
        std::vector<float4> h((size_t)nInternal * 2 * width);
        std::vector<int> lv(nInternal);
        cudaMemcpy(h.data(), ws->wide, bytes, cudaMemcpyDeviceToHost);
        cudaMemcpy(lv.data(), levelOf, 4 * (size_t)nInternal, cudaMemcpyDeviceToHost);
        long long nodesSeen = 0, kids = 0, leaves = 0, bad = 0; int deepest = 0;
        std::vector<std::pair<uint32_t, int>> stack{ { 0u, 1 } };
        while (!stack.empty()) {
            auto [id, d] = stack.back(); stack.pop_back();
            nodesSeen++; deepest = std::max(deepest, d);
            if (id != 0 && lv[id] != d) bad++;
            for (int s2 = 0; s2 < width; s2++) {
                const float4* nodep = &h[(size_t)id * 2 * width];
                const float4* rec = nodep + 3 * (s2 >> 1);
                const float hx = (s2 & 1) ? rec[1].w : rec[1].z;
                if (!(hx >= 0.f)) continue;
                uint32_t r; memcpy(&r, reinterpret_cast<const char*>(nodep) + 24 * width + 4 * s2, 4);
                kids++;
                if (r & LEAF_BIT) leaves++;
                else if (r >= (uint32_t)nInternal) bad++;
                else stack.push_back({ r, d + 1 });
            }
        }
        fprintf(stderr, "[brt] wide-%d: %lld nodes (of %d binary), %.2f children / node, %lld leaves, depth %d (reported %d), %lld inconsistencies, %.3f ms\n",
                width, nodesSeen, nInternal, nodesSeen ? (double)kids / nodesSeen : 0.0, leaves, deepest, out->depth, bad, ms);
    }
    return cudaSuccess;
}

#define BVH_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { cleanup(); return e_; } } while (0)

void free_bvh_workspace(BvhWorkspace* ws) {
    if (!ws) return;
    cudaFree(ws->arena); cudaFree(ws->nodes[0]); cudaFree(ws->nodes[1]); cudaFree(ws->wide); cudaFree(ws->level); cudaFree(ws->cnodes);
    if (ws->host) cudaFreeHost(ws->host);
    *ws = BvhWorkspace{};
}

// All build memory lives in `ws` (grow-only, owned by the ctx): one scratch arena carved into the temporaries plus two
// node buffers for the two candidate trees, so re-building for a new scene of similar size allocates nothing
// (cudaMalloc / cudaFree of ~100 MB blocks cost far more than the 1 ms build itself).
cudaError_t build_lbvh(const DevScene& sc, BvhWorkspace* ws, BvhBuildResult* out, cudaStream_t st) {
    out->nodes = nullptr; out->cnodes = nullptr; out->nNodes = 0; out->depth = 0; out->buildMs = 0.f;
    out->extent[0] = out->extent[1] = out->extent[2] = 0.f;
    const int n = sc.nSph + sc.nBox + sc.nTri;
    if (n < 2) return cudaSuccess;                    // 0 or 1 bounded primitive: traversal falls back to the linear loop
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
    const int nb = (n + 255) / 256, sortBlocks = (n + SORT_TILE - 1) / SORT_TILE;
    const int costBlocks = (n - 1 + 255) / 256;
    // carve the arena
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off += (bytes + 255) & ~(size_t)255; return at; };
    const size_t oBoxes = take(sizeof(Aabb) * n), oNodeBox = take(sizeof(Aabb) * (n - 1)), oBlockBounds = take(sizeof(float) * 6 * nb),
                 oSb = take(sizeof(float) * 6), oK0 = take(4 * (size_t)n), oK1 = take(4 * (size_t)n), oV0 = take(4 * (size_t)n), oV1 = take(4 * (size_t)n),
                 oHist = take(4 * 256 * (size_t)sortBlocks), oFlags = take(4 * (size_t)(n - 1)), oChildren = take(sizeof(int2) * (n - 1)),
                 oParent = take(4 * (size_t)(2 * n - 1)), oDepth = take(4 * (size_t)(n - 1)), oSums = take(sizeof(double) * costBlocks);
    if (off > ws->arenaCap) {
        cudaFree(ws->arena); ws->arena = nullptr; ws->arenaCap = 0;
        BVH_CK(cudaMalloc(&ws->arena, off));
        ws->arenaCap = off;
    }
    const size_t nodeBytes = sizeof(float4) * 4 * (size_t)(n - 1);
    for (int k = 0; k < 2; k++) if (nodeBytes > ws->nodeCap[k]) {
        cudaFree(ws->nodes[k]); ws->nodes[k] = nullptr; ws->nodeCap[k] = 0;
        BVH_CK(cudaMalloc(&ws->nodes[k], nodeBytes));
        ws->nodeCap[k] = nodeBytes;
    }
    char* A = (char*)ws->arena;
    Aabb *boxes = (Aabb*)(A + oBoxes), *nodeBox = (Aabb*)(A + oNodeBox);
    float *blockBounds = (float*)(A + oBlockBounds), *sb = (float*)(A + oSb);
    uint32_t *k0 = (uint32_t*)(A + oK0), *k1 = (uint32_t*)(A + oK1), *v0 = (uint32_t*)(A + oV0), *v1 = (uint32_t*)(A + oV1), *hist = (uint32_t*)(A + oHist);
    unsigned int* flags = (unsigned int*)(A + oFlags);
    int2* children = (int2*)(A + oChildren);
    int *parent = (int*)(A + oParent), *nodeDepth = (int*)(A + oDepth);
    double* blockSums = (double*)(A + oSums);
    float4 *nodes = ws->nodes[0], *nodesAlt = ws->nodes[1];
    // results come back through ONE small pinned block (grow-only, owned by the workspace), so the copies are truly asynchronous:
    //   [cost sums of candidate 0 | of candidate 1 | depth 0, depth 1 | scene bounds | padded boxes (small scenes only)]
    const bool small = n <= SAH_HOST_MAX;
    const size_t hSums = sizeof(double) * (size_t)costBlocks, hMisc = 2 * hSums, hBoxes = hMisc + 64;
    const size_t hostBytes = hBoxes + (small ? sizeof(Aabb) * (size_t)n : 0);
    if (hostBytes > ws->hostCap) {
        if (ws->host) cudaFreeHost(ws->host);
        ws->host = nullptr; ws->hostCap = 0;
        BVH_CK(cudaHostAlloc(&ws->host, hostBytes, cudaHostAllocDefault));
        ws->hostCap = hostBytes;
    }
    char* HB = (char*)ws->host;
    double* hostSums[2] = { (double*)HB, (double*)(HB + hSums) };
    int* hostDepth = (int*)(HB + hMisc);
    float* hostSb = (float*)(HB + hMisc + 8);
    const Aabb* hostBoxes = (const Aabb*)(HB + hBoxes);
    cudaEvent_t eBoxes = nullptr;
    BVH_CK(cudaEventCreate(&e0)); BVH_CK(cudaEventCreate(&e1));
    BVH_CK(cudaEventRecord(e0, st));
    k_prim_bounds<<<nb, 256, 0, st>>>(sc, n, boxes, blockBounds);
    k_scene_bounds<<<1, 256, 0, st>>>(blockBounds, nb, sb);
    BVH_CK(cudaMemcpyAsync(hostSb, sb, 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (small) {
        BVH_CK(cudaMemcpyAsync((void*)hostBoxes, boxes, sizeof(Aabb) * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (cudaEventCreateWithFlags(&eBoxes, cudaEventDisableTiming) != cudaSuccess) eBoxes = nullptr;
        if (eBoxes) cudaEventRecord(eBoxes, st);
    }
    // two candidate hierarchies (Morton quantisation per axis / uniform), enqueued back to back without a host synchronisation:
    // candidate 1 reuses the scratch arrays of candidate 0, whose results have been copied out in stream order by then.  Small
    // scenes get the host SAH tree as their alternative instead of the uniform-scale LBVH (17 launches less on the e2e path).
    const int nModes = small ? 1 : 2;
    for (int mode = 0; mode < nModes; mode++) {
        float4* target = mode == 0 ? nodes : nodesAlt;
        k_morton<<<nb, 256, 0, st>>>(boxes, n, sb, mode, k0, v0);
        uint32_t *ki = k0, *ko = k1, *vi = v0, *vo = v1;
        for (int shift = 0; shift < 32; shift += 8) {
            k_sort_hist<<<sortBlocks, SORT_THREADS, 0, st>>>(ki, n, shift, hist, sortBlocks);
            k_sort_scan<<<1, 1024, 0, st>>>(hist, 256 * sortBlocks);
            k_sort_scatter<<<sortBlocks, SORT_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, hist, sortBlocks);
            uint32_t* t = ki; ki = ko; ko = t; t = vi; vi = vo; vo = t;
        }
        BVH_CK(cudaMemsetAsync(flags, 0, 4 * (size_t)(n - 1), st));
        k_karras<<<(n - 1 + 255) / 256, 256, 0, st>>>(ki, n, children, parent);
        k_refit<<<nb, 256, 0, st>>>(sc, n, vi, boxes, children, parent, nodeBox, nodeDepth, flags, target);
        k_sah_cost<<<costBlocks, 256, 0, st>>>(nodeBox, n - 1, blockSums);
        BVH_CK(cudaMemcpyAsync(hostSums[mode], blockSums, hSums, cudaMemcpyDeviceToHost, st));
        BVH_CK(cudaMemcpyAsync(hostDepth + mode, nodeDepth, 4, cudaMemcpyDeviceToHost, st));
    }
    // third candidate for small scenes, built on the host over the padded device boxes WHILE the GPU builds the other two
    HostSah* sah = nullptr; int sahHeight = 0;
    std::vector<uint32_t> pids;
    if (small) {
        cudaError_t e = eBoxes ? cudaEventSynchronize(eBoxes) : cudaStreamSynchronize(st);
        if (eBoxes) cudaEventDestroy(eBoxes);
        if (e != cudaSuccess) { cleanup(); return e; }
        pids.resize(n);
        for (int i = 0; i < n; i++) pids[i] = i < sc.nSph ? make_pid(PT_SPHERE, i) : i < sc.nSph + sc.nBox ? make_pid(PT_BOX, i - sc.nSph) : make_pid(PT_TRI, i - sc.nSph - sc.nBox);
        sah = new HostSah(hostBoxes, pids.data(), n);
        Aabb rootBox;
        sah->build(0, n, rootBox, sahHeight);
    }
    {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { delete sah; cleanup(); return e; }
    }
    double bestCost = 0.0; int bestMode = -1, bestDepth = 0;
    for (int mode = 0; mode < nModes; mode++) {
        double cost = 0.0;
        for (int k = 0; k < costBlocks; k++) cost += hostSums[mode][k];                  // fixed order: deterministic choice
        if (getenv("BRT_DEBUG")) fprintf(stderr, "[brt] lbvh candidate %d: sah cost %.6g depth %d\n", mode, cost, hostDepth[mode]);
        // the surface-area estimate assumes uniformly distributed rays; differences of a few percent are not predictive
        // (measured: C3 2 % apart, slower tree estimated cheaper), so the uniform-scale tree must win by > 10 % to be taken
        if (bestMode < 0 || cost < 0.9 * bestCost) { bestCost = cost; bestMode = mode; bestDepth = hostDepth[mode]; }
    }
    if (bestMode == 1) nodes = nodesAlt;
    if (sah) {
        if (getenv("BRT_DEBUG")) fprintf(stderr, "[brt] host sah candidate: sah cost %.6g depth %d\n", sah->cost, sahHeight);
        const char* mg = getenv("BRT_SAH_MARGIN");
        // taken only when clearly cheaper: at 0.89 of the LBVH's cost (C3) the two trees render equally fast (6 355 vs 6 388
        // Msamples/s), at 0.85 (the Cornell box) the SAH tree is 4 % faster
        const double margin = mg ? atof(mg) : 0.88;
        if (sah->cost < margin * bestCost && sahHeight <= SMEM_ONLY_MAX_DEPTH) {
            // taken: written into the node buffer the LBVH choice did not take (pageable source: the copy is staged before the call returns)
            float4* target = nodes == ws->nodes[0] ? ws->nodes[1] : ws->nodes[0];
            cudaError_t e = cudaMemcpyAsync(target, sah->nodes.data(), sizeof(float4) * sah->nodes.size(), cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { delete sah; cleanup(); return e; }
            nodes = target; bestCost = sah->cost; bestDepth = sahHeight; bestMode = 2;
        }
        delete sah; sah = nullptr;
    }
    if (nodeBytes > ws->cnodeCap) {
        cudaFree(ws->cnodes); ws->cnodes = nullptr; ws->cnodeCap = 0;
        BVH_CK(cudaMalloc(&ws->cnodes, nodeBytes));
        ws->cnodeCap = nodeBytes;
    }
    k_centre_half<<<(n - 1 + 255) / 256, 256, 0, st>>>(nodes, n - 1, ws->cnodes);
    BVH_CK(cudaEventRecord(e1, st));
    BVH_CK(cudaGetLastError());
    BVH_CK(cudaEventSynchronize(e1));
    float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
    int depth = bestDepth;
    cleanup();
    out->nodes = nodes; out->cnodes = ws->cnodes; out->nNodes = n - 1; out->depth = depth; out->buildMs = ms;
    for (int k = 0; k < 3; k++) out->extent[k] = hostSb[3 + k] - hostSb[k];
    return cudaSuccess;
}

}  // namespace brt
