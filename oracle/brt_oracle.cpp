// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// float64, brute-force, CPU restatement of the reference renderer's hot path
// (Shinzef/BlenderRayTracer, js/*.js).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library;
// the product (libbrt.so) never links, loads or calls it.
//
// PARITY STATUS: PINNED to the reference's own source.  The reference ships no tests or golden vectors and the image has
// no JavaScript engine, so baseline/minijs.py (a small interpreter for the JavaScript subset the reference uses) executes the
// UNMODIFIED js/*.js with Math.random replaced by this file's Philox stream; tests/test_reference_pin.py requires this oracle
// to reproduce the resulting tests/golden/reference_vectors.json bit for bit (20 cases: radiance, floatData, RGBA8).  It is
// also held by hand-derived float64 known-answer vectors (tests/golden/kat.json) and a second, independently written
// restatement (tests/golden/independent_port.py -> independent_vectors.npz, bit for bit on 13 images).
//
// Every function cites the reference file:line it follows.  Arithmetic is
// IEEE double with no FMA contraction (compile with -ffp-contract=off) because
// JavaScript Numbers are doubles and V8 never fuses multiply-add.
//
// The one deliberate substitution: Math.random() (unseeded xorshift128+ in V8,
// js/math.js:21-31) is replaced by a counter-based Philox4x32-10 stream keyed
// by (seed, pixel, sample), consumed in exactly the reference's draw order, so
// renders are reproducible and the GPU engine's "reference sampler" mode can
// consume the identical stream.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <memory>
#include <thread>
#include <vector>
#include <algorithm>
#include <atomic>

namespace {

// ---------------------------------------------------------------- JS Math.* helpers
// Math.max / Math.min propagate NaN (ECMA-262 21.3.2.24/25); std::max does not.
inline double js_max(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    return a > b ? a : b;
}
inline double js_min(double a, double b) {
    if (std::isnan(a) || std::isnan(b)) return std::numeric_limits<double>::quiet_NaN();
    return a < b ? a : b;
}
const double JS_PI = 3.141592653589793;
const double JS_INF = std::numeric_limits<double>::infinity();

// ---------------------------------------------------------------- RNG (stands in for Math.random)
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have;
    Philox(uint64_t seed, uint32_t pixel, uint32_t sample) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = pixel; ctr[1] = sample; ctr[2] = 0; ctr[3] = 0x42525431u; // "BRT1"
        have = 0;
    }
    static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b; hi = (uint32_t)(p >> 32); lo = (uint32_t)p;
    }
    void refill() {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; r++) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        buf[0] = c0; buf[1] = c1; buf[2] = c2; buf[3] = c3;
        ctr[2]++;              // next block of this (pixel, sample) stream
        have = 4;
    }
    // uniform in [0,1) with 24 bits — exactly representable in fp32 and fp64
    double next() {
        if (!have) refill();
        uint32_t u = buf[4 - have]; have--;
        return (double)(u >> 8) * (1.0 / 16777216.0);
    }
};

// ---------------------------------------------------------------- js/math.js:6-32
struct Vec3 {
    double x, y, z;
    Vec3(double x_ = 0, double y_ = 0, double z_ = 0) : x(x_), y(y_), z(z_) {}
    Vec3 add(const Vec3& v) const { return Vec3(x + v.x, y + v.y, z + v.z); }          // :11
    Vec3 sub(const Vec3& v) const { return Vec3(x - v.x, y - v.y, z - v.z); }          // :12
    Vec3 mul(double s) const { return Vec3(x * s, y * s, z * s); }                     // :13
    Vec3 div(double s) const { return Vec3(x / s, y / s, z / s); }                     // :14
    double dot(const Vec3& v) const { return x * v.x + y * v.y + z * v.z; }            // :15
    Vec3 cross(const Vec3& v) const {                                                  // :16
        return Vec3(y * v.z - z * v.y, z * v.x - x * v.z, x * v.y - y * v.x);
    }
    double length() const { return std::sqrt(x * x + y * y + z * z); }                 // :17
    Vec3 normalize() const { double l = length(); return l > 0 ? div(l) : Vec3(); }    // :18
    Vec3 reflect(const Vec3& n) const { return sub(n.mul(2 * dot(n))); }               // :19

    static Vec3 random(Philox& g) {                                                    // :21
        double a = g.next(), b = g.next(), c = g.next();
        return Vec3(a, b, c);
    }
    static Vec3 randomInUnitSphere(Philox& g) {                                        // :22-26
        Vec3 p;
        do { p = Vec3::random(g).mul(2).sub(Vec3(1, 1, 1)); } while (p.dot(p) >= 1.0);
        return p;
    }
    static Vec3 randomInUnitDisk(Philox& g) {                                          // :27-31
        Vec3 p;
        do { double a = g.next(); double b = g.next(); p = Vec3(a * 2 - 1, b * 2 - 1, 0); } while (p.dot(p) >= 1.0);
        return p;
    }
};

struct Ray {                                                                           // math.js:35-42
    Vec3 origin, direction;
    Ray() {}
    Ray(const Vec3& o, const Vec3& d) : origin(o), direction(d) {}
    Vec3 at(double t) const { return origin.add(direction.mul(t)); }
};

struct Material;
struct HitRecord {                                                                     // math.js:45-59
    Vec3 point, normal;
    double t = 0;
    bool frontFace = true;
    const Material* material = nullptr;
    double u = 0, v = 0;
    int objId = -1, triId = -1;     // oracle bookkeeping for the AOV contract (not in the reference)
    void setFaceNormal(const Ray& ray, const Vec3& outwardNormal) {                    // :55-58
        frontFace = ray.direction.dot(outwardNormal) < 0;
        normal = frontFace ? outwardNormal : outwardNormal.mul(-1);
    }
};

// ---------------------------------------------------------------- js/materials.js
struct Scatter { Ray scattered; Vec3 attenuation; };
enum MatType { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_EMISSIVE = 3 };

struct Texture;                                                                        // js/textures.js (defined after PerlinNoise)
Vec3 texture_value(const Texture* t, const Vec3& p);

struct Material {
    int type;
    Vec3 color;        // albedo / emissive colour
    double p;          // roughness | refractionIndex | intensity
    const Texture* tex = nullptr;   // TexturedLambertian / TexturedMetal (materials.js:99-126): attenuation = texture.value(u, v, point)
    Material(int t, const Vec3& c, double p_) : type(t), color(c), p(p_) {
        if (type == MAT_METAL) p = js_min(p_, 1);                                      // materials.js:33
    }
    // returns false for "null" (absorbed / no scatter)
    bool scatter(const Ray& ray, const HitRecord& rec, Philox& g, Scatter& out) const {
        switch (type) {
        case MAT_LAMBERTIAN: {                                                         // materials.js:20-25
            Vec3 scatterDirection = rec.normal.add(Vec3::randomInUnitSphere(g).normalize());
            out.scattered = Ray(rec.point, scatterDirection);
            out.attenuation = tex ? texture_value(tex, rec.point) : color;                 // materials.js:109
            return true;
        }
        case MAT_METAL: {                                                              // materials.js:36-41
            Vec3 reflected = ray.direction.normalize().reflect(rec.normal);
            out.scattered = Ray(rec.point, reflected.add(Vec3::randomInUnitSphere(g).mul(p)));
            out.attenuation = tex ? texture_value(tex, rec.point) : color;                 // materials.js:124
            return out.scattered.direction.dot(rec.normal) > 0;
        }
        case MAT_DIELECTRIC: {                                                         // materials.js:51-70
            out.attenuation = Vec3(1, 1, 1);
            double refractionRatio = rec.frontFace ? (1.0 / p) : p;
            Vec3 unitDirection = ray.direction.normalize();
            double cosTheta = js_min(unitDirection.mul(-1).dot(rec.normal), 1.0);
            double sinTheta = std::sqrt(1.0 - cosTheta * cosTheta);
            bool cannotRefract = refractionRatio * sinTheta > 1.0;
            Vec3 direction;
            // short-circuit ||: the uniform is drawn only when refraction is possible (:62)
            if (cannotRefract || reflectance(cosTheta, refractionRatio) > g.next()) {
                direction = unitDirection.reflect(rec.normal);
            } else {
                direction = refract(unitDirection, rec.normal, refractionRatio);
            }
            out.scattered = Ray(rec.point, direction);
            return true;
        }
        default:                                                                       // materials.js:9,94
            return false;
        }
    }
    Vec3 emitted() const {                                                             // materials.js:10,95
        if (type == MAT_EMISSIVE) return color.mul(p);
        return Vec3(0, 0, 0);
    }
    static Vec3 refract(const Vec3& uv, const Vec3& n, double etaiOverEtat) {          // materials.js:72-77
        double cosTheta = js_min(uv.mul(-1).dot(n), 1.0);
        Vec3 rOutPerpendicular = uv.add(n.mul(cosTheta)).mul(etaiOverEtat);
        Vec3 rOutParallel = n.mul(-std::sqrt(std::fabs(1.0 - rOutPerpendicular.dot(rOutPerpendicular))));
        return rOutPerpendicular.add(rOutParallel);
    }
    static double reflectance(double cosine, double refIdx) {                          // materials.js:79-83
        double r0 = (1 - refIdx) / (1 + refIdx);
        r0 = r0 * r0;
        return r0 + (1 - r0) * std::pow((1 - cosine), 5);
    }
};

// ---------------------------------------------------------------- js/geometry.js
struct Hittable {
    const Material* material = nullptr;
    virtual ~Hittable() {}
    virtual bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const = 0;
};

struct Sphere : Hittable {                                                             // geometry.js:8-46
    Vec3 center; double radius;
    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const override {
        Vec3 oc = ray.origin.sub(center);
        double a = ray.direction.dot(ray.direction);
        double halfB = oc.dot(ray.direction);
        double c = oc.dot(oc) - radius * radius;
        double discriminant = halfB * halfB - a * c;
        if (discriminant < 0) return false;
        double sqrtd = std::sqrt(discriminant);
        double root = (-halfB - sqrtd) / a;
        if (root < tMin || tMax < root) {
            root = (-halfB + sqrtd) / a;
            if (root < tMin || tMax < root) return false;
        }
        rec = HitRecord();
        rec.t = root;
        rec.point = ray.at(rec.t);
        Vec3 outwardNormal = rec.point.sub(center).div(radius);
        rec.setFaceNormal(ray, outwardNormal);
        rec.material = material;
        // UV mapping (:39-42) — consumed by no live material, but the reference pays for it
        double theta = std::acos(-outwardNormal.y);
        double phi = std::atan2(-outwardNormal.z, outwardNormal.x) + JS_PI;
        rec.u = phi / (2 * JS_PI);
        rec.v = theta / JS_PI;
        return true;
    }
};

struct Plane : Hittable {                                                              // geometry.js:49-75
    Vec3 point, normal;   // normal already normalised by the ctor (:52)
    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const override {
        double denom = normal.dot(ray.direction);
        if (std::fabs(denom) < 1e-6) return false;
        double t = point.sub(ray.origin).dot(normal) / denom;
        if (t < tMin || t > tMax) return false;
        rec = HitRecord();
        rec.t = t;
        rec.point = ray.at(t);
        rec.setFaceNormal(ray, normal);
        rec.material = material;
        rec.u = (rec.point.x + 10) / 20;
        rec.v = (rec.point.z + 10) / 20;
        return true;
    }
};

struct Box : Hittable {                                                                // geometry.js:78-133
    Vec3 min, max;
    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const override {
        double tMinBox = (min.x - ray.origin.x) / ray.direction.x;
        double tMaxBox = (max.x - ray.origin.x) / ray.direction.x;
        if (tMinBox > tMaxBox) std::swap(tMinBox, tMaxBox);
        double tMinY = (min.y - ray.origin.y) / ray.direction.y;
        double tMaxY = (max.y - ray.origin.y) / ray.direction.y;
        if (tMinY > tMaxY) std::swap(tMinY, tMaxY);
        if (tMinBox > tMaxY || tMinY > tMaxBox) return false;
        tMinBox = js_max(tMinBox, tMinY);
        tMaxBox = js_min(tMaxBox, tMaxY);
        double tMinZ = (min.z - ray.origin.z) / ray.direction.z;
        double tMaxZ = (max.z - ray.origin.z) / ray.direction.z;
        if (tMinZ > tMaxZ) std::swap(tMinZ, tMaxZ);
        if (tMinBox > tMaxZ || tMinZ > tMaxBox) return false;
        tMinBox = js_max(tMinBox, tMinZ);
        tMaxBox = js_min(tMaxBox, tMaxZ);
        double t = tMinBox > tMin ? tMinBox : tMaxBox;
        if (t < tMin || t > tMax) return false;
        rec = HitRecord();
        rec.t = t;
        rec.point = ray.at(t);
        const Vec3& p = rec.point;
        const double eps = 1e-6;
        if (std::fabs(p.x - min.x) < eps) rec.normal = Vec3(-1, 0, 0);
        else if (std::fabs(p.x - max.x) < eps) rec.normal = Vec3(1, 0, 0);
        else if (std::fabs(p.y - min.y) < eps) rec.normal = Vec3(0, -1, 0);
        else if (std::fabs(p.y - max.y) < eps) rec.normal = Vec3(0, 1, 0);
        else if (std::fabs(p.z - min.z) < eps) rec.normal = Vec3(0, 0, -1);
        else rec.normal = Vec3(0, 0, 1);
        Vec3 n = rec.normal;
        rec.setFaceNormal(ray, n);
        rec.material = material;
        return true;
    }
};

struct Triangle : Hittable {                                                           // geometry.js:136-189
    Vec3 v0, v1, v2, normal;
    Triangle(const Vec3& a, const Vec3& b, const Vec3& c, const Material* m) : v0(a), v1(b), v2(c) {
        material = m;
        Vec3 edge1 = v1.sub(v0);
        Vec3 edge2 = v2.sub(v0);
        normal = edge1.cross(edge2).normalize();                                       // :143-145
    }
    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const override {
        Vec3 edge1 = v1.sub(v0);
        Vec3 edge2 = v2.sub(v0);
        Vec3 h = ray.direction.cross(edge2);
        double a = edge1.dot(h);
        if (std::fabs(a) < 0.0001) return false;
        double f = 1.0 / a;
        Vec3 s = ray.origin.sub(v0);
        double u = f * s.dot(h);
        if (u < 0 || u > 1) return false;
        Vec3 q = s.cross(edge1);
        double v = f * ray.direction.dot(q);
        if (v < 0 || u + v > 1) return false;
        double t = f * edge2.dot(q);
        if (t < tMin || t > tMax) return false;
        rec = HitRecord();
        rec.t = t;
        rec.point = ray.at(t);
        rec.setFaceNormal(ray, normal);
        rec.material = material;
        rec.u = u;
        rec.v = v;
        return true;
    }
};

struct TriangleMesh : Hittable {                                                       // geometry.js:192-263
    std::vector<Triangle> triangles;
    // ctor (:193-237): `indices` arrive as doubles so the JS index semantics survive:
    // `idx >= vertices.length` skips the triangle; any other index that does not name an
    // element (negative, fractional, NaN) reads `undefined` → _ensureVec3 → (0,0,0) (:240-246).
    TriangleMesh(const double* verts, int nverts, const double* idx, int nidx, const Material* m) {
        material = m;
        for (int i = 0; i < nidx; i += 3) {
            if (i + 2 >= nidx) continue;                                               // :207-210
            double i0 = idx[i], i1 = idx[i + 1], i2 = idx[i + 2];
            if (i0 >= nverts || i1 >= nverts || i2 >= nverts) continue;                // :216-219
            triangles.emplace_back(fetch(verts, nverts, i0), fetch(verts, nverts, i1), fetch(verts, nverts, i2), m);
        }
    }
    static Vec3 fetch(const double* verts, int nverts, double i) {
        if (!(i >= 0) || i != std::floor(i) || i >= nverts) return Vec3(0, 0, 0);
        int k = (int)i;
        return Vec3(verts[3 * k], verts[3 * k + 1], verts[3 * k + 2]);
    }
    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& rec) const override { // :248-262
        bool any = false;
        double closestT = tMax;
        HitRecord tmp;
        for (size_t k = 0; k < triangles.size(); k++) {
            if (triangles[k].hit(ray, tMin, closestT, tmp)) {
                rec = tmp;
                rec.triId = (int)k;
                closestT = tmp.t;
                any = true;
            }
        }
        return any;
    }
};

// ---------------------------------------------------------------- js/lights.js
struct Light {
    int type;          // 0 point, 1 directional
    Vec3 position;     // point
    Vec3 direction;    // directional: normalised by the ctor (lights.js:38)
    Vec3 color;
    double intensity;
};
struct Illum { Vec3 direction, color; double distance; };
Illum illuminate(const Light& L, const Vec3& point) {
    Illum r;
    if (L.type == 0) {                                                                 // lights.js:22-31
        Vec3 direction = L.position.sub(point);
        double distance = direction.length();
        double attenuation = 1.0 / (1.0 + 0.1 * distance + 0.01 * distance * distance);
        r.direction = direction.normalize();
        r.color = L.color.mul(L.intensity * attenuation);
        r.distance = distance;
    } else {                                                                           // lights.js:41-47
        r.direction = L.direction.mul(-1);
        r.color = L.color.mul(L.intensity);
        r.distance = JS_INF;
    }
    return r;
}

// ---------------------------------------------------------------- js/noise.js:6-61
struct PerlinNoise {
    int p[512];
    PerlinNoise() { for (int i = 0; i < 512; i++) p[i] = i & 255; }   // identity until set_perm
    static double fade(double t) { return t * t * t * (t * (t * 6 - 15) + 10); }       // :20
    static double lerp(double t, double a, double b) { return a + t * (b - a); }       // :21
    static double grad(int hash, double x, double y, double z) {                       // :22-27
        int h = hash & 15;
        double u = h < 8 ? x : y;
        double v = h < 4 ? y : (h == 12 || h == 14) ? x : z;
        return ((h & 1) == 0 ? u : -u) + ((h & 2) == 0 ? v : -v);
    }
    static int floor_and_255(double x) {   // Math.floor(x) & 255 (ToInt32 then mask)
        double f = std::floor(x);
        if (!std::isfinite(f)) return 0;
        double m = std::fmod(f, 4294967296.0);
        int64_t i = (int64_t)m;
        return (int)(i & 255);
    }
    double noise(const Vec3& point) const {                                            // :29-61
        double x = point.x, y = point.y, z = point.z;
        int X = floor_and_255(x), Y = floor_and_255(y), Z = floor_and_255(z);
        double fx = x - std::floor(x), fy = y - std::floor(y), fz = z - std::floor(z);
        double u = fade(fx), v = fade(fy), w = fade(fz);
        int A = p[X] + Y, AA = p[A] + Z, AB = p[A + 1] + Z;
        int B = p[X + 1] + Y, BA = p[B] + Z, BB = p[B + 1] + Z;
        return lerp(w,
            lerp(v,
                lerp(u, grad(p[AA], fx, fy, fz), grad(p[BA], fx - 1, fy, fz)),
                lerp(u, grad(p[AB], fx, fy - 1, fz), grad(p[BB], fx - 1, fy - 1, fz))),
            lerp(v,
                lerp(u, grad(p[AA + 1], fx, fy, fz - 1), grad(p[BA + 1], fx - 1, fy, fz - 1)),
                lerp(u, grad(p[AB + 1], fx, fy - 1, fz - 1), grad(p[BB + 1], fx - 1, fy - 1, fz - 1))));
    }
    double turbulence(Vec3 point, int depth) const {                                   // :63-75
        double accum = 0, weight = 1.0;
        Vec3 tempP = point;
        for (int i = 0; i < depth; i++) {
            accum += weight * noise(tempP);
            weight *= 0.5;
            tempP = tempP.mul(2);
        }
        return std::fabs(accum);
    }
};

// ---------------------------------------------------------------- js/textures.js
// Every texture's value(u, v, p) ignores u and v and uses only the hit point p (textures.js:21,33-36,47-50,61-65,77-82).
enum TexKind { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_MARBLE = 3, TEX_WOOD = 4 };
struct Texture {
    int kind = TEX_SOLID;
    Vec3 odd, even;      // SolidColor.color lives in `odd`
    double scale = 1;
    PerlinNoise noise;   // each noise-based texture owns a PerlinNoise (random permutation in the reference: an input here)
};
Vec3 texture_value(const Texture* t, const Vec3& p) {
    switch (t->kind) {
    case TEX_CHECKER: {                                                                // textures.js:33-36
        double sines = std::sin(t->scale * p.x) * std::sin(t->scale * p.y) * std::sin(t->scale * p.z);
        return sines < 0 ? t->odd : t->even;
    }
    case TEX_NOISE: {                                                                  // textures.js:47-50
        double n = t->noise.noise(p.mul(t->scale));
        return Vec3(1, 1, 1).mul(0.5 * (1 + n));
    }
    case TEX_MARBLE: {                                                                 // textures.js:61-65
        double n = t->noise.turbulence(p.mul(t->scale), 7);
        double marble = 0.5 * (1 + std::sin(t->scale * p.z + 10 * n));
        return Vec3(0.9, 0.8, 0.7).mul(marble).add(Vec3(0.6, 0.4, 0.3).mul(1 - marble));
    }
    case TEX_WOOD: {                                                                   // textures.js:77-82
        double grain = t->noise.noise(p.mul(t->scale * 20));
        double rings = std::sin(t->scale * std::sqrt(p.x * p.x + p.z * p.z) + grain * 10);
        double wood = 0.5 * (1 + rings);
        return Vec3(0.8, 0.5, 0.2).mul(wood).add(Vec3(0.4, 0.2, 0.1).mul(1 - wood));
    }
    default: return t->odd;                                                            // textures.js:21
    }
}

// ---------------------------------------------------------------- js/camera.js
enum CamType { CAM_PERSPECTIVE = 0, CAM_ORTHOGRAPHIC = 1, CAM_OTHER = 2 };
struct Camera {
    int type = CAM_PERSPECTIVE;
    double aperture = 0, focusDist = 1, fov = 45, lensRadius = 0;
    Vec3 w, u, v, origin, horizontal, vertical, lowerLeftCorner;
    Camera() {}
    Camera(const Vec3& lookFrom, const Vec3& lookAt, const Vec3& vup, double vfov, double aspect,
           double aperture_, double focusDist_, int type_) {                           // camera.js:8-36
        type = type_; aperture = aperture_; focusDist = focusDist_; fov = vfov;
        double theta = vfov * JS_PI / 180;
        double h = std::tan(theta / 2);
        double viewportHeight = 2.0 * h;
        double viewportWidth = aspect * viewportHeight;
        w = lookFrom.sub(lookAt).normalize();
        u = vup.cross(w).normalize();
        v = w.cross(u);
        origin = lookFrom;
        if (type == CAM_PERSPECTIVE) {                                                 // :25
            horizontal = u.mul(viewportWidth * focusDist);
            vertical = v.mul(viewportHeight * focusDist);
            lowerLeftCorner = origin.sub(horizontal.div(2)).sub(vertical.div(2)).sub(w.mul(focusDist));
        } else {
            horizontal = u.mul(viewportWidth);
            vertical = v.mul(viewportHeight);
            lowerLeftCorner = origin.sub(horizontal.div(2)).sub(vertical.div(2));
        }
        lensRadius = aperture / 2;
    }
    Ray getRay(double s, double t, Philox& g) const {                                  // camera.js:38-51
        if (type == CAM_ORTHOGRAPHIC) {                                                // :39
            Vec3 offset = Vec3::randomInUnitDisk(g).mul(lensRadius);
            Vec3 rayOrigin = origin.add(u.mul(offset.x)).add(v.mul(offset.y));
            Vec3 rayDirection = lowerLeftCorner.add(horizontal.mul(s)).add(vertical.mul(t)).sub(rayOrigin).add(w.mul(-1));
            return Ray(rayOrigin, rayDirection.normalize());
        } else {
            Vec3 rd = Vec3::randomInUnitDisk(g).mul(lensRadius);
            Vec3 offset = u.mul(rd.x).add(v.mul(rd.y));
            Vec3 rayOrigin = origin.add(offset);
            Vec3 rayDirection = lowerLeftCorner.add(horizontal.mul(s)).add(vertical.mul(t)).sub(rayOrigin);
            return Ray(rayOrigin, rayDirection);
        }
    }
    // pixel-centre ray with lens offset exactly 0 (AOV contract; the disk draw scaled by lensRadius=0)
    Ray getRayNoLens(double s, double t) const {
        if (type == CAM_ORTHOGRAPHIC) {
            Vec3 rayOrigin = origin;
            Vec3 rayDirection = lowerLeftCorner.add(horizontal.mul(s)).add(vertical.mul(t)).sub(rayOrigin).add(w.mul(-1));
            return Ray(rayOrigin, rayDirection.normalize());
        }
        Vec3 rayDirection = lowerLeftCorner.add(horizontal.mul(s)).add(vertical.mul(t)).sub(origin);
        return Ray(origin, rayDirection);
    }
};

// ---------------------------------------------------------------- js/world.js
enum BgKind { BG_GRADIENT = 0, BG_SOLID = 1, BG_HDRI = 2, BG_PROCEDURAL_SKY = 3 };

struct World {
    std::vector<std::unique_ptr<Hittable>> objects;
    std::vector<std::unique_ptr<Material>> materials;
    std::vector<std::unique_ptr<Texture>> textures;
    std::vector<Light> lights;
    int background = BG_GRADIENT;                                                      // world.js:12
    Vec3 solidColor = Vec3(0.1, 0.1, 0.1);
    double skyIntensity = 1.0;                                                         // world.js:13
    PerlinNoise cloudNoise;                                                            // world.js:14

    bool hit(const Ray& ray, double tMin, double tMax, HitRecord& out) const {         // world.js:20-33
        bool any = false;
        double closestT = tMax;
        HitRecord h;
        for (size_t k = 0; k < objects.size(); k++) {
            if (objects[k]->hit(ray, tMin, closestT, h) && h.t < closestT) {
                closestT = h.t;
                out = h;
                out.objId = (int)k;
                any = true;
            }
        }
        return any;
    }
    // shadow-ray any-hit used ONLY by the direct-lighting extension (no reference analogue)
    bool occluded(const Ray& ray, double tMin, double tMax) const {
        HitRecord h;
        for (size_t k = 0; k < objects.size(); k++) {
            if (objects[k]->material->type == MAT_EMISSIVE) continue;
            if (objects[k]->hit(ray, tMin, tMax, h) && h.t < tMax) return true;
        }
        return false;
    }
    Vec3 skyGradient(const Ray& ray) const {                                           // world.js:35-40
        double t = 0.5 * (ray.direction.normalize().y + 1.0);
        Vec3 white(1.0, 1.0, 1.0), blue(0.5, 0.7, 1.0);
        return white.mul(1.0 - t).add(blue.mul(t)).mul(skyIntensity);
    }
    Vec3 solidBackground() const { return solidColor.mul(skyIntensity); }              // world.js:42-44 (intended use, ray-tracer.js:573)
    Vec3 proceduralSky(const Ray& ray) const {                                         // world.js:46-72
        Vec3 dir = ray.direction.normalize();
        Vec3 sunDir = Vec3(0.3, 0.6, 0.8).normalize();
        double sunDot = js_max(0, dir.dot(sunDir));
        double sunIntensity = std::pow(sunDot, 512);
        Vec3 sunColor = Vec3(1.0, 0.95, 0.8).mul(sunIntensity * 10);
        double horizonBlend = js_max(0, dir.y);
        Vec3 skyColor = Vec3(0.4, 0.7, 1.0).mul(horizonBlend * 0.8);
        double horizonGlow = std::exp(-std::fabs(dir.y) * 4) * 0.3;
        Vec3 glowColor = Vec3(1.0, 0.8, 0.6).mul(horizonGlow);
        Vec3 groundColor = Vec3(0.1, 0.15, 0.1).mul(js_max(0, -dir.y * 0.5));
        Vec3 cloudPos(dir.x * 10, dir.y * 3 + 2, dir.z * 10);
        double cloud = js_max(0, cloudNoise.noise(cloudPos) * 0.8 + 0.2);
        Vec3 cloudColor = Vec3(0.9, 0.9, 1.0).mul(cloud * js_max(0, dir.y) * 0.5);
        return skyColor.add(glowColor).add(groundColor).add(sunColor).add(cloudColor).mul(skyIntensity);
    }
    Vec3 hdriBackground(const Ray& ray) const {                                        // world.js:74-110
        Vec3 dir = ray.direction.normalize();
        Vec3 sunDir = Vec3(-0.3, 0.6, -0.5).normalize();
        double sunDot = js_max(0, dir.dot(sunDir));
        double sunSize = 0.04;
        double sunMask = sunDot > (1.0 - sunSize) ? 1.0 : 0.0;
        Vec3 sunColor = Vec3(1.0, 0.95, 0.8).mul(sunMask * 20);
        double coronaSize = 0.2;
        double coronaIntensity = js_max(0, (sunDot - (1.0 - coronaSize)) / coronaSize);
        Vec3 coronaColor = Vec3(1.0, 0.8, 0.6).mul(std::pow(coronaIntensity, 2) * 3);
        double y = dir.y;
        double skyI = js_max(0, y * 0.5 + 0.5);
        Vec3 skyColor = Vec3(0.3, 0.5, 0.8).mul(skyI * 2);
        double groundBounce = js_max(0, -y * 0.3);
        Vec3 groundColor = Vec3(0.2, 0.15, 0.1).mul(groundBounce);
        double scatter = std::pow(js_max(0, 1.0 - std::fabs(y)), 2) * 0.3;
        Vec3 scatterColor = Vec3(0.8, 0.9, 1.0).mul(scatter);
        return skyColor.add(groundColor).add(scatterColor).add(sunColor).add(coronaColor).mul(skyIntensity);
    }
    Vec3 backgroundColor(const Ray& ray) const {
        switch (background) {
        case BG_SOLID: return solidBackground();
        case BG_HDRI: return hdriBackground(ray);
        case BG_PROCEDURAL_SKY: return proceduralSky(ray);
        default: return skyGradient(ray);
        }
    }
};

// ---------------------------------------------------------------- js/post-processor.js
Vec3 reinhardToneMap(const Vec3& color, double exposure) {                             // :9-16
    Vec3 m = color.mul(exposure);
    return Vec3(m.x / (1.0 + m.x), m.y / (1.0 + m.y), m.z / (1.0 + m.z));
}
double aces1(double x) {                                                               // :19-32
    const double a = 2.51, b = 0.03, c = 2.43, d = 0.59, e = 0.14;
    return js_max(0, (x * (a * x + b)) / (x * (c * x + d) + e));
}
Vec3 acesToneMap(const Vec3& color, double exposure) {
    Vec3 ec = color.mul(exposure);
    return Vec3(aces1(ec.x), aces1(ec.y), aces1(ec.z));
}
Vec3 gammaCorrect(const Vec3& color, double gamma) {                                   // :35-42
    double invGamma = 1.0 / gamma;
    return Vec3(std::pow(js_max(0, color.x), invGamma), std::pow(js_max(0, color.y), invGamma),
                std::pow(js_max(0, color.z), invGamma));
}
void denoise(const float* imageData, int width, int height, double strength, float* result) {   // :45-77
    const int halfKernel = 1;
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width; x++) {
            double r = 0, g = 0, b = 0, weight = 0;
            for (int ky = -halfKernel; ky <= halfKernel; ky++) {
                for (int kx = -halfKernel; kx <= halfKernel; kx++) {
                    int nx = std::max(0, std::min(width - 1, x + kx));
                    int ny = std::max(0, std::min(height - 1, y + ky));
                    size_t idx = ((size_t)ny * width + nx) * 4;
                    double w = std::exp(-(double)(kx * kx + ky * ky) / (2 * strength * strength));
                    r += imageData[idx] * w;
                    g += imageData[idx + 1] * w;
                    b += imageData[idx + 2] * w;
                    weight += w;
                }
            }
            size_t idx = ((size_t)y * width + x) * 4;
            result[idx] = (float)(r / weight);
            result[idx + 1] = (float)(g / weight);
            result[idx + 2] = (float)(b / weight);
            result[idx + 3] = imageData[idx + 3];
        }
    }
}
// Math.min(255, Math.max(0, Math.floor(c * 255))) stored into a Uint8ClampedArray (ray-tracer.js:226-233):
// NaN survives min/max and the clamped store turns it into 0.
uint8_t quantize(double c) {
    double q = js_min(255, js_max(0, std::floor(c * 255)));
    if (std::isnan(q)) return 0;
    return (uint8_t)q;
}

// ---------------------------------------------------------------- js/ray-tracer.js
struct RenderParams {
    int32_t width, height;
    int32_t samples, maxBounces;
    int32_t antiAliasing;     // 0 none, 1 supersampling, 2 stochastic
    int32_t toneMapping;      // 0 reinhard, 1 aces, 2 linear
    double exposure, gamma;
    int32_t denoising;
    double denoiseStrength;
    uint64_t seed;
    int32_t directLighting;   // extension, default 0 = reference behaviour
    int32_t sampleBegin;      // first global sample index (0 = reference loop); lets tests restate a multi-GPU spp split
};

struct Scene {
    World world;
    Camera camera;
    std::atomic<long long> rays{0};
};

Vec3 rayColor(const Scene& sc, const RenderParams& rp, const Ray& ray, int depth, Philox& g, long long& nrays) {   // ray-tracer.js:102-123
    if (depth <= 0) return Vec3(0, 0, 0);
    HitRecord hit;
    nrays++;
    if (sc.world.hit(ray, 0.001, JS_INF, hit)) {
        Vec3 emitted = hit.material->emitted();
        Vec3 direct(0, 0, 0);
        if (rp.directLighting && hit.material->type == MAT_LAMBERTIAN) {
            // EXTENSION (off by default): SURVEY §8a-18.  lights.js:22-47 supply direction/colour/distance.
            for (const Light& L : sc.world.lights) {
                Illum il = illuminate(L, hit.point);
                double cosN = hit.normal.dot(il.direction);
                if (!(cosN > 0)) continue;
                if (sc.world.occluded(Ray(hit.point, il.direction), 0.001, il.distance)) continue;
                const Vec3& al = hit.material->color;
                direct = direct.add(Vec3(al.x * il.color.x, al.y * il.color.y, al.z * il.color.z).mul(cosN));
            }
        }
        Scatter s;
        if (hit.material->scatter(ray, hit, g, s)) {
            Vec3 scattered = rayColor(sc, rp, s.scattered, depth - 1, g, nrays);
            return emitted.add(direct).add(Vec3(s.attenuation.x * scattered.x, s.attenuation.y * scattered.y,
                                    s.attenuation.z * scattered.z));
        }
        return emitted.add(direct);
    }
    return sc.world.backgroundColor(ray);
}

void getAntiAliasSample(const RenderParams& rp, int i, int j, Philox& g, double& u, double& v) {   // ray-tracer.js:125-149
    if (rp.antiAliasing == 2) {
        double r1 = g.next();
        double r2 = g.next();
        double offsetX = std::sqrt(r1) * std::cos(2 * JS_PI * r2);
        double offsetY = std::sqrt(r1) * std::sin(2 * JS_PI * r2);
        u = (i + 0.5 + offsetX * 0.5) / rp.width;
        v = (j + 0.5 + offsetY * 0.5) / rp.height;
    } else if (rp.antiAliasing == 1) {
        double a = g.next();
        double b = g.next();
        u = (i + a) / rp.width;
        v = (j + b) / rp.height;
    } else {
        u = (i + 0.5) / rp.width;
        v = (j + 0.5) / rp.height;
    }
}

Vec3 toneMap(const RenderParams& rp, const Vec3& color) {                              // ray-tracer.js:151-161
    switch (rp.toneMapping) {
    case 1: return acesToneMap(color, rp.exposure);
    case 2: return color.mul(rp.exposure);
    default: return reinhardToneMap(color, rp.exposure);
    }
}

// One pixel of render()'s double loop (ray-tracer.js:195-253).  (i, j): column, UP-row.
void renderPixel(const Scene& sc, const RenderParams& rp, int i, int j,
                 uint8_t* rgba, float* floatData, double* linear, long long& nrays) {
    Vec3 color(0, 0, 0);
    int sampleCount = rp.antiAliasing == 0 ? 1 : rp.samples;                           // :201
    uint32_t pix = (uint32_t)((rp.height - 1 - j) * rp.width + i);
    for (int s = 0; s < sampleCount; s++) {
        Philox g(rp.seed, pix, (uint32_t)(s + rp.sampleBegin));
        double u, v;
        getAntiAliasSample(rp, i, j, g, u, v);
        Ray ray = sc.camera.getRay(u, v, g);
        color = color.add(rayColor(sc, rp, ray, rp.maxBounces, g, nrays));
    }
    color = color.div(sampleCount);                                                    // :208
    size_t pixelIndex = ((size_t)(rp.height - 1 - j) * rp.width + i) * 4;              // :215
    if (linear) { linear[pixelIndex] = color.x; linear[pixelIndex + 1] = color.y; linear[pixelIndex + 2] = color.z; linear[pixelIndex + 3] = 1.0; }
    color = toneMap(rp, color);
    color = gammaCorrect(color, rp.gamma);
    if (floatData) {
        floatData[pixelIndex] = (float)color.x; floatData[pixelIndex + 1] = (float)color.y;
        floatData[pixelIndex + 2] = (float)color.z; floatData[pixelIndex + 3] = 1.0f;
    }
    if (rgba) {
        rgba[pixelIndex] = quantize(color.x); rgba[pixelIndex + 1] = quantize(color.y);
        rgba[pixelIndex + 2] = quantize(color.z); rgba[pixelIndex + 3] = 255;
    }
}

Material* makeMaterial(Scene* s, int mtype, const double* col, double mp) {
    s->world.materials.emplace_back(new Material(mtype, col ? Vec3(col[0], col[1], col[2]) : Vec3(), mp));
    return s->world.materials.back().get();
}

} // namespace

// ================================================================= C API (ctypes)
extern "C" {

void* orc_scene_new() { return new Scene(); }
void orc_scene_free(void* s) { delete (Scene*)s; }

int orc_add_sphere(void* s_, const double* c, double r, int mtype, const double* col, double mp) {
    Scene* s = (Scene*)s_;
    auto* o = new Sphere(); o->center = Vec3(c[0], c[1], c[2]); o->radius = r; o->material = makeMaterial(s, mtype, col, mp);
    s->world.objects.emplace_back(o); return (int)s->world.objects.size() - 1;
}
int orc_add_plane(void* s_, const double* pt, const double* n, int mtype, const double* col, double mp) {
    Scene* s = (Scene*)s_;
    auto* o = new Plane(); o->point = Vec3(pt[0], pt[1], pt[2]); o->normal = Vec3(n[0], n[1], n[2]).normalize();   // geometry.js:52
    o->material = makeMaterial(s, mtype, col, mp);
    s->world.objects.emplace_back(o); return (int)s->world.objects.size() - 1;
}
int orc_add_box(void* s_, const double* mn, const double* mx, int mtype, const double* col, double mp) {
    Scene* s = (Scene*)s_;
    auto* o = new Box(); o->min = Vec3(mn[0], mn[1], mn[2]); o->max = Vec3(mx[0], mx[1], mx[2]); o->material = makeMaterial(s, mtype, col, mp);
    s->world.objects.emplace_back(o); return (int)s->world.objects.size() - 1;
}
int orc_add_triangle(void* s_, const double* a, const double* b, const double* c, int mtype, const double* col, double mp) {
    Scene* s = (Scene*)s_;
    auto* o = new Triangle(Vec3(a[0], a[1], a[2]), Vec3(b[0], b[1], b[2]), Vec3(c[0], c[1], c[2]), makeMaterial(s, mtype, col, mp));
    s->world.objects.emplace_back(o); return (int)s->world.objects.size() - 1;
}
int orc_add_mesh(void* s_, const double* verts, int nverts, const double* idx, int nidx, int mtype, const double* col, double mp) {
    Scene* s = (Scene*)s_;
    auto* o = new TriangleMesh(verts, nverts, idx, nidx, makeMaterial(s, mtype, col, mp));
    s->world.objects.emplace_back(o); return (int)s->world.objects.size() - 1;
}
int orc_mesh_triangle_count(void* s_, int obj) {
    Scene* s = (Scene*)s_;
    auto* m = dynamic_cast<TriangleMesh*>(s->world.objects[obj].get());
    return m ? (int)m->triangles.size() : -1;
}
int orc_object_count(void* s_) { return (int)((Scene*)s_)->world.objects.size(); }

void orc_add_point_light(void* s_, const double* pos, const double* col, double intensity) {
    Light L; L.type = 0; L.position = Vec3(pos[0], pos[1], pos[2]); L.color = Vec3(col[0], col[1], col[2]); L.intensity = intensity;
    ((Scene*)s_)->world.lights.push_back(L);
}
void orc_add_directional_light(void* s_, const double* dir, const double* col, double intensity) {
    Light L; L.type = 1; L.direction = Vec3(dir[0], dir[1], dir[2]).normalize();      // lights.js:38
    L.color = Vec3(col[0], col[1], col[2]); L.intensity = intensity;
    ((Scene*)s_)->world.lights.push_back(L);
}
// lights.js illuminate(): out = direction[3], color[3], distance
void orc_illuminate(void* s_, int light, const double* p, double* out7) {
    Illum il = illuminate(((Scene*)s_)->world.lights[light], Vec3(p[0], p[1], p[2]));
    out7[0] = il.direction.x; out7[1] = il.direction.y; out7[2] = il.direction.z;
    out7[3] = il.color.x; out7[4] = il.color.y; out7[5] = il.color.z; out7[6] = il.distance;
}

void orc_set_camera(void* s_, const double* from, const double* at, const double* up, double vfov, double aspect,
                    double aperture, double focusDist, int type) {
    ((Scene*)s_)->camera = Camera(Vec3(from[0], from[1], from[2]), Vec3(at[0], at[1], at[2]), Vec3(up[0], up[1], up[2]),
                                  vfov, aspect, aperture, focusDist, type);
}
// out[0..20] = origin, lowerLeftCorner, horizontal, vertical, u, v, w ; out[21..24] = lensRadius, fov, aperture, focusDist ; out[25] = type
void orc_get_camera(void* s_, double* out) {
    const Camera& c = ((Scene*)s_)->camera;
    const Vec3* vs[7] = { &c.origin, &c.lowerLeftCorner, &c.horizontal, &c.vertical, &c.u, &c.v, &c.w };
    for (int k = 0; k < 7; k++) { out[3 * k] = vs[k]->x; out[3 * k + 1] = vs[k]->y; out[3 * k + 2] = vs[k]->z; }
    out[21] = c.lensRadius; out[22] = c.fov; out[23] = c.aperture; out[24] = c.focusDist; out[25] = c.type;
}
void orc_copy_camera(void* dst, const void* src) { ((Scene*)dst)->camera = ((const Scene*)src)->camera; }
void orc_set_background(void* s_, int kind, const double* color, double intensity) {
    Scene* s = (Scene*)s_;
    s->world.background = kind;
    if (color) s->world.solidColor = Vec3(color[0], color[1], color[2]);
    s->world.skyIntensity = intensity;
}
// new TexturedLambertian(texture) / new TexturedMetal(texture, roughness) on object `obj` (materials.js:99-126)
int orc_set_object_texture(void* s_, int obj, int kind, const double* odd, const double* even, double scale, const int* perm256) {
    Scene* s = (Scene*)s_;
    if (obj < 0 || obj >= (int)s->world.objects.size() || kind < 0 || kind > 4) return -1;
    s->world.textures.emplace_back(new Texture());
    Texture* t = s->world.textures.back().get();
    t->kind = kind; t->odd = Vec3(odd[0], odd[1], odd[2]); t->even = Vec3(even[0], even[1], even[2]); t->scale = scale;
    if (perm256) for (int i = 0; i < 256; i++) { t->noise.p[i] = perm256[i] & 255; t->noise.p[256 + i] = perm256[i] & 255; }
    const_cast<Material*>(s->world.objects[obj]->material)->tex = t;
    return 0;
}
void orc_texture_value(void* s_, int obj, const double* p, double* out3) {
    Scene* s = (Scene*)s_;
    Vec3 c = texture_value(s->world.objects[obj]->material->tex, Vec3(p[0], p[1], p[2]));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
void orc_set_perm(void* s_, const int* perm256) {                                      // noise.js:6-18 (shuffle result supplied)
    Scene* s = (Scene*)s_;
    for (int i = 0; i < 256; i++) { s->world.cloudNoise.p[i] = perm256[i] & 255; s->world.cloudNoise.p[256 + i] = perm256[i] & 255; }
}
void orc_background(void* s_, const double* dir, double* out3) {
    Scene* s = (Scene*)s_;
    Vec3 c = s->world.backgroundColor(Ray(Vec3(), Vec3(dir[0], dir[1], dir[2])));
    out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
double orc_perlin(void* s_, const double* p) { return ((Scene*)s_)->world.cloudNoise.noise(Vec3(p[0], p[1], p[2])); }
double orc_turbulence(void* s_, const double* p, int depth) { return ((Scene*)s_)->world.cloudNoise.turbulence(Vec3(p[0], p[1], p[2]), depth); }

// scalar known-answer helpers
void orc_tonemap(int kind, double exposure, const double* in3, double* out3) {
    RenderParams rp{}; rp.toneMapping = kind; rp.exposure = exposure;
    Vec3 c = toneMap(rp, Vec3(in3[0], in3[1], in3[2])); out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
void orc_gamma(double gamma, const double* in3, double* out3) {
    Vec3 c = gammaCorrect(Vec3(in3[0], in3[1], in3[2]), gamma); out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
}
int orc_quantize(double c) { return quantize(c); }
double orc_schlick(double cosine, double refIdx) { return Material::reflectance(cosine, refIdx); }
void orc_refract(const double* uv, const double* n, double eta, double* out3) {
    Vec3 r = Material::refract(Vec3(uv[0], uv[1], uv[2]), Vec3(n[0], n[1], n[2]), eta); out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}
void orc_denoise(const float* in, int w, int h, double strength, float* out) { denoise(in, w, h, strength, out); }
void orc_quantize_image(const float* in, int w, int h, uint8_t* rgba) {               // ray-tracer.js:270-275
    size_t n = (size_t)w * h * 4;
    for (size_t i = 0; i < n; i += 4) {
        rgba[i] = quantize(in[i]); rgba[i + 1] = quantize(in[i + 1]); rgba[i + 2] = quantize(in[i + 2]); rgba[i + 3] = 255;
    }
}
// raw Philox4x32-10 block (Random123 known-answer vectors pin the generator itself)
void orc_philox_raw(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4) {
    Philox g(0, 0, 0);
    g.key[0] = key2[0]; g.key[1] = key2[1];
    for (int i = 0; i < 4; i++) g.ctr[i] = ctr4[i];
    g.refill();
    for (int i = 0; i < 4; i++) out4[i] = g.buf[i];
}
// first n uniforms of the (seed, pixel, sample) stream — lets tests pin the GPU's Philox against this one
void orc_rng_stream(uint64_t seed, uint32_t pixel, uint32_t sample, int n, double* out) {
    Philox g(seed, pixel, sample);
    for (int i = 0; i < n; i++) out[i] = g.next();
}

// Render the image rectangle [x0,x1) × [y0,y1) (image coordinates, row 0 = top) into full-size W×H buffers
// (any of rgba / floatData / linear may be NULL).  Rows are split over `nthreads` host threads; the result
// is independent of the thread count because the RNG is keyed by (pixel, sample).  Returns rays traced.
long long orc_render_rect(void* s_, const RenderParams* rp, int x0, int y0, int x1, int y1, int nthreads,
                          uint8_t* rgba, float* floatData, double* linear) {
    Scene* s = (Scene*)s_;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> nextRow(y0);
    std::atomic<long long> total(0);
    auto work = [&]() {
        long long nrays = 0;
        for (;;) {
            int row = nextRow.fetch_add(1);
            if (row >= y1) break;
            int j = rp->height - 1 - row;
            for (int i = x0; i < x1; i++) renderPixel(*s, *rp, i, j, rgba, floatData, linear, nrays);
        }
        total += nrays;
    };
    std::vector<std::thread> th;
    for (int k = 1; k < nthreads; k++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return total.load();
}

// Primary-visibility AOVs at pixel centres, lens offset 0 (ray-tracer.js:144-147 + camera.js:45-49 + world.js:20-33).
void orc_primary_aov(void* s_, int width, int height, int32_t* objId, int32_t* triId, double* t, double* normal3, uint8_t* frontFace) {
    Scene* s = (Scene*)s_;
    for (int row = 0; row < height; row++) {
        int j = height - 1 - row;
        for (int i = 0; i < width; i++) {
            double u = (i + 0.5) / width, v = (j + 0.5) / height;
            Ray ray = s->camera.getRayNoLens(u, v);
            HitRecord h;
            size_t k = (size_t)row * width + i;
            if (s->world.hit(ray, 0.001, JS_INF, h)) {
                objId[k] = h.objId; triId[k] = h.triId; t[k] = h.t;
                normal3[3 * k] = h.normal.x; normal3[3 * k + 1] = h.normal.y; normal3[3 * k + 2] = h.normal.z;
                frontFace[k] = h.frontFace ? 1 : 0;
            } else {
                objId[k] = -1; triId[k] = -1; t[k] = JS_INF;
                normal3[3 * k] = normal3[3 * k + 1] = normal3[3 * k + 2] = 0; frontFace[k] = 0;
            }
        }
    }
}

} // extern "C"
