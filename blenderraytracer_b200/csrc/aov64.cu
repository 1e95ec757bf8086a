// float64 primary-visibility kernel for the bit-exact parity contract ("primary-hit object IDs are bit-exact
// against the reference").  Compiled with -fmad=false: IEEE double +,-,*,/,sqrt are correctly rounded on the
// device exactly as in JavaScript, and every expression below keeps the reference's operation ORDER
// (math.js:11-19, geometry.js, world.js:20-33, camera.js:45-49), so object / triangle IDs, t and normals are
// the same bits a JS engine would produce.  Brute force on purpose: the same O(N) loops as the reference.
#include "brt_kernels.h"

namespace brt {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 mul(V3 a, double s) { return v3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 divs(V3 a, double s) { return v3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double length(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 normalize(V3 a) { double l = length(a); return l > 0 ? divs(a, l) : v3(0, 0, 0); }
__device__ __forceinline__ V3 ld3(const double* p) { return v3(p[0], p[1], p[2]); }

// geometry.js:15-29
__device__ __forceinline__ bool sphere64(const Obj64& o, V3 O, V3 D, double tMin, double tMax, double& t) {
    V3 oc = sub(O, ld3(o.a));
    double a = dot(D, D);
    double halfB = dot(oc, D);
    double c = dot(oc, oc) - o.b[0] * o.b[0];
    double disc = halfB * halfB - a * c;
    if (disc < 0) return false;
    double sqrtd = sqrt(disc);
    double root = (-halfB - sqrtd) / a;
    if (root < tMin || tMax < root) {
        root = (-halfB + sqrtd) / a;
        if (root < tMin || tMax < root) return false;
    }
    t = root;
    return true;
}
// geometry.js:56-61
__device__ __forceinline__ bool plane64(const Obj64& o, V3 O, V3 D, double tMin, double tMax, double& t) {
    V3 n = ld3(o.b);
    double denom = dot(n, D);
    if (fabs(denom) < 1e-6) return false;
    double tt = dot(sub(ld3(o.a), O), n) / denom;
    if (tt < tMin || tt > tMax) return false;
    t = tt;
    return true;
}
// geometry.js:85-112
__device__ __forceinline__ bool box64(const Obj64& o, V3 O, V3 D, double tMin, double tMax, double& t) {
    double t0 = (o.a[0] - O.x) / D.x, t1 = (o.b[0] - O.x) / D.x;
    if (t0 > t1) { double s = t0; t0 = t1; t1 = s; }
    double y0 = (o.a[1] - O.y) / D.y, y1 = (o.b[1] - O.y) / D.y;
    if (y0 > y1) { double s = y0; y0 = y1; y1 = s; }
    if (t0 > y1 || y0 > t1) return false;
    // Math.max / Math.min propagate NaN
    t0 = (t0 != t0 || y0 != y0) ? t0 + y0 : (t0 > y0 ? t0 : y0);
    t1 = (t1 != t1 || y1 != y1) ? t1 + y1 : (t1 < y1 ? t1 : y1);
    double z0 = (o.a[2] - O.z) / D.z, z1 = (o.b[2] - O.z) / D.z;
    if (z0 > z1) { double s = z0; z0 = z1; z1 = s; }
    if (t0 > z1 || z0 > t1) return false;
    t0 = (t0 != t0 || z0 != z0) ? t0 + z0 : (t0 > z0 ? t0 : z0);
    t1 = (t1 != t1 || z1 != z1) ? t1 + z1 : (t1 < z1 ? t1 : z1);
    double tt = t0 > tMin ? t0 : t1;
    if (tt < tMin || tt > tMax) return false;
    t = tt;
    return true;
}
// geometry.js:148-175
__device__ __forceinline__ bool tri64(V3 v0, V3 v1, V3 v2, V3 O, V3 D, double tMin, double tMax, double& t) {
    V3 e1 = sub(v1, v0), e2 = sub(v2, v0);
    V3 h = cross(D, e2);
    double a = dot(e1, h);
    if (fabs(a) < 0.0001) return false;
    double f = 1.0 / a;
    V3 s = sub(O, v0);
    double u = f * dot(s, h);
    if (u < 0 || u > 1) return false;
    V3 q = cross(s, e1);
    double v = f * dot(D, q);
    if (v < 0 || u + v > 1) return false;
    double tt = f * dot(e2, q);
    if (tt < tMin || tt > tMax) return false;
    t = tt;
    return true;
}

__global__ void __launch_bounds__(128) k_primary_aov64(const Obj64* __restrict__ objs, int nObjs, const double* __restrict__ meshTris,
                                                       Cam64 cam, int W, int H, int* objId, int* triId, double* tOut, double* nrm,
                                                       unsigned char* front) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const int row = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    if (col >= W || row >= H) return;
    const int j = H - 1 - row;
    double u = (col + 0.5) / W, v = (j + 0.5) / H;                                    // ray-tracer.js:144-147
    V3 O = ld3(cam.origin);
    V3 D = sub(add(add(ld3(cam.llc), mul(ld3(cam.horizontal), u)), mul(ld3(cam.vertical), v)), O);   // camera.js:48
    if (cam.type == 1) D = normalize(add(D, mul(ld3(cam.w), -1.0)));                  // camera.js:42-43
    const double tMin = 0.001;
    double closestT = CUDART_INF;
    int bestObj = -1, bestTri = -1;
    for (int k = 0; k < nObjs; k++) {                                                 // world.js:24-30
        const Obj64& o = objs[k];
        double t; bool h = false; int tr = -1;
        if (o.type == 0) h = sphere64(o, O, D, tMin, closestT, t);
        else if (o.type == 1) h = plane64(o, O, D, tMin, closestT, t);
        else if (o.type == 2) h = box64(o, O, D, tMin, closestT, t);
        else if (o.type == 3) h = tri64(ld3(o.a), ld3(o.b), ld3(o.c), O, D, tMin, closestT, t);
        else {                                                                        // geometry.js:248-262
            double ct = closestT;
            for (long long i = 0; i < o.triCount; i++) {
                const double* p = meshTris + 9 * (o.firstTri + i);
                double tt;
                if (tri64(ld3(p), ld3(p + 3), ld3(p + 6), O, D, tMin, ct, tt)) { ct = tt; t = tt; tr = (int)i; h = true; }
            }
        }
        if (h && t < closestT) { closestT = t; bestObj = k; bestTri = tr; }
    }
    size_t k = (size_t)row * W + col;
    if (bestObj < 0) {
        objId[k] = -1; triId[k] = -1; tOut[k] = CUDART_INF; nrm[3 * k] = nrm[3 * k + 1] = nrm[3 * k + 2] = 0; front[k] = 0;
        return;
    }
    const Obj64& o = objs[bestObj];
    V3 P = add(O, mul(D, closestT));                                                  // math.js:41
    V3 n;
    if (o.type == 0) n = divs(sub(P, ld3(o.a)), o.b[0]);                              // geometry.js:34
    else if (o.type == 1) n = ld3(o.b);
    else if (o.type == 2) {                                                           // geometry.js:119-126
        const double eps = 1e-6;
        if (fabs(P.x - o.a[0]) < eps) n = v3(-1, 0, 0);
        else if (fabs(P.x - o.b[0]) < eps) n = v3(1, 0, 0);
        else if (fabs(P.y - o.a[1]) < eps) n = v3(0, -1, 0);
        else if (fabs(P.y - o.b[1]) < eps) n = v3(0, 1, 0);
        else if (fabs(P.z - o.a[2]) < eps) n = v3(0, 0, -1);
        else n = v3(0, 0, 1);
    } else {
        V3 v0, v1, v2;
        if (o.type == 3) { v0 = ld3(o.a); v1 = ld3(o.b); v2 = ld3(o.c); }
        else { const double* p = meshTris + 9 * (o.firstTri + bestTri); v0 = ld3(p); v1 = ld3(p + 3); v2 = ld3(p + 6); }
        n = normalize(cross(sub(v1, v0), sub(v2, v0)));                               // geometry.js:143-145
    }
    bool ff = dot(D, n) < 0;                                                          // math.js:56-57
    if (!ff) n = mul(n, -1.0);
    objId[k] = bestObj; triId[k] = bestTri; tOut[k] = closestT;
    nrm[3 * k] = n.x; nrm[3 * k + 1] = n.y; nrm[3 * k + 2] = n.z; front[k] = ff ? 1 : 0;
}

cudaError_t launch_primary_aov64(const Obj64* objs, int nObjs, const double* meshTris, const Cam64& cam, int W, int H, int* objId,
                                 int* triId, double* t, double* nrm, unsigned char* front, cudaStream_t st) {
    dim3 grid((W + 15) / 16, (H + 7) / 8);
    k_primary_aov64<<<grid, 128, 0, st>>>(objs, nObjs, meshTris, cam, W, H, objId, triId, t, nrm, front);
    return cudaGetLastError();
}

}  // namespace brt
