"""A SECOND, independent restatement of the reference renderer's hot path, in pure Python floats (IEEE double, no FMA —
the arithmetic JavaScript performs), written from the reference JavaScript without looking at oracle/brt_oracle.cpp.

Purpose: cross-pin the C++ oracle.  The reference cannot be executed in this image (browser JS, no engine), so instead of
one hand transcription trusted on its own, two independent transcriptions — this one (functional style, tuples) and the
oracle (C++ classes) — must agree BIT FOR BIT on whole stochastic images when fed the same Philox stream in place of
Math.random().  `python tests/golden/independent_port.py` regenerates tests/golden/independent_vectors.npz;
tests/test_oracle_kat.py::test_oracle_matches_independent_port compares the oracle against those vectors.
This is NOT output of the reference; it is test infrastructure, like the oracle.

Followed sources (Shinzef/BlenderRayTracer): js/math.js, js/camera.js, js/geometry.js, js/materials.js, js/world.js,
js/noise.js, js/post-processor.js, js/ray-tracer.js:102-165,183-281, js/scene-loader.js (well-formed input only).
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INF = float("inf")


# ------------------------------------------------------------------------------------------------ RNG (stands in for Math.random)
class Philox:
    """Philox4x32-10, counter (pixel, sample, block, 'BRT1'), key = seed; uniforms = top 24 bits, consumed x, y, z, w."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

    def __init__(self, seed, pixel, sample):
        self.key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        self.pixel, self.sample, self.block, self.buf = pixel & 0xFFFFFFFF, sample & 0xFFFFFFFF, 0, []

    def _block(self, blk):
        c = [self.pixel, self.sample, blk, 0x42525431]
        k0, k1 = self.key
        for _ in range(10):
            p0, p1 = self.M0 * c[0], self.M1 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
            k0, k1 = (k0 + self.W0) & 0xFFFFFFFF, (k1 + self.W1) & 0xFFFFFFFF
        return c

    def random(self):
        if not self.buf:
            self.buf = self._block(self.block)
            self.block += 1
        return (self.buf.pop(0) >> 8) / 16777216.0


# ------------------------------------------------------------------------------------------------ vectors (math.js)
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def mul(a, s): return (a[0] * s, a[1] * s, a[2] * s)
def div(a, s): return (a[0] / s, a[1] / s, a[2] / s)
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def length(a): return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def normalize(a):
    n = length(a)
    return div(a, n) if n > 0 else (0.0, 0.0, 0.0)


def reflect(v, n): return sub(v, mul(n, 2 * dot(v, n)))
def at(o, d, t): return add(o, mul(d, t))


def in_unit_sphere(rng):
    while True:
        p = sub(mul((rng.random(), rng.random(), rng.random()), 2), (1, 1, 1))
        if not dot(p, p) >= 1.0:
            return p


def in_unit_disk(rng):
    while True:
        p = (rng.random() * 2 - 1, rng.random() * 2 - 1, 0.0)
        if not dot(p, p) >= 1.0:
            return p


def face(d, n):
    front = dot(d, n) < 0
    return front, (n if front else mul(n, -1))


# ------------------------------------------------------------------------------------------------ geometry (geometry.js)
def hit_sphere(ob, o, d, tmin, tmax):
    c, r = ob["center"], ob["radius"]
    oc = sub(o, c)
    a = dot(d, d)
    hb = dot(oc, d)
    cc = dot(oc, oc) - r * r
    disc = hb * hb - a * cc
    if disc < 0:
        return None
    s = math.sqrt(disc)
    root = (-hb - s) / a
    if root < tmin or tmax < root:
        root = (-hb + s) / a
        if root < tmin or tmax < root:
            return None
    p = at(o, d, root)
    front, n = face(d, div(sub(p, c), r))
    return dict(t=root, p=p, n=n, front=front)


def hit_plane(ob, o, d, tmin, tmax):
    n0 = ob["normal"]
    den = dot(n0, d)
    if abs(den) < 1e-6:
        return None
    t = dot(sub(ob["point"], o), n0) / den
    if t < tmin or t > tmax:
        return None
    front, n = face(d, n0)
    return dict(t=t, p=at(o, d, t), n=n, front=front)


def _div(a, b):
    """JavaScript division: x/0 = +-Infinity, 0/0 = NaN."""
    if b == 0:
        if a == 0 or a != a:
            return float("nan")
        return math.copysign(INF, a) * math.copysign(1.0, b)
    return a / b


def _jsmax(a, b): return float("nan") if (a != a or b != b) else max(a, b)
def _jsmin(a, b): return float("nan") if (a != a or b != b) else min(a, b)


def hit_box(ob, o, d, tmin, tmax):
    mn, mx = ob["min"], ob["max"]
    t0, t1 = _div(mn[0] - o[0], d[0]), _div(mx[0] - o[0], d[0])
    if t0 > t1:
        t0, t1 = t1, t0
    y0, y1 = _div(mn[1] - o[1], d[1]), _div(mx[1] - o[1], d[1])
    if y0 > y1:
        y0, y1 = y1, y0
    if t0 > y1 or y0 > t1:
        return None
    t0, t1 = _jsmax(t0, y0), _jsmin(t1, y1)
    z0, z1 = _div(mn[2] - o[2], d[2]), _div(mx[2] - o[2], d[2])
    if z0 > z1:
        z0, z1 = z1, z0
    if t0 > z1 or z0 > t1:
        return None
    t0, t1 = _jsmax(t0, z0), _jsmin(t1, z1)
    t = t0 if t0 > tmin else t1
    if t < tmin or t > tmax:
        return None
    p = at(o, d, t)
    eps = 1e-6
    if abs(p[0] - mn[0]) < eps: n = (-1.0, 0.0, 0.0)
    elif abs(p[0] - mx[0]) < eps: n = (1.0, 0.0, 0.0)
    elif abs(p[1] - mn[1]) < eps: n = (0.0, -1.0, 0.0)
    elif abs(p[1] - mx[1]) < eps: n = (0.0, 1.0, 0.0)
    elif abs(p[2] - mn[2]) < eps: n = (0.0, 0.0, -1.0)
    else: n = (0.0, 0.0, 1.0)
    front, n = face(d, n)
    return dict(t=t, p=p, n=n, front=front)


def hit_triangle(tri, o, d, tmin, tmax):
    v0, v1, v2, nrm = tri
    e1, e2 = sub(v1, v0), sub(v2, v0)
    h = cross(d, e2)
    a = dot(e1, h)
    if abs(a) < 0.0001:
        return None
    f = 1.0 / a
    s = sub(o, v0)
    u = f * dot(s, h)
    if u < 0 or u > 1:
        return None
    q = cross(s, e1)
    v = f * dot(d, q)
    if v < 0 or u + v > 1:
        return None
    t = f * dot(e2, q)
    if t < tmin or t > tmax:
        return None
    front, n = face(d, nrm)
    return dict(t=t, p=at(o, d, t), n=n, front=front)


def make_triangle(v0, v1, v2):
    return (v0, v1, v2, normalize(cross(sub(v1, v0), sub(v2, v0))))


def hit_mesh(ob, o, d, tmin, tmax):
    closest, ct = None, tmax
    for tri in ob["tris"]:
        h = hit_triangle(tri, o, d, tmin, ct)
        if h:
            closest, ct = h, h["t"]
    return closest


HIT = dict(sphere=hit_sphere, plane=hit_plane, box=hit_box, triangle=lambda ob, o, d, a, b: hit_triangle(ob["tri"], o, d, a, b), mesh=hit_mesh)


def world_hit(world, o, d, tmin, tmax):
    best, ct = None, tmax
    for ob in world["objects"]:
        h = HIT[ob["kind"]](ob, o, d, tmin, ct)
        if h and h["t"] < ct:
            ct, best = h["t"], dict(h, mat=ob["mat"])
    return best


# ------------------------------------------------------------------------------------------------ noise.js + backgrounds (world.js)
def perlin(p512, pt):
    x, y, z = pt
    X, Y, Z = math.floor(x) & 255, math.floor(y) & 255, math.floor(z) & 255
    fx, fy, fz = x - math.floor(x), y - math.floor(y), z - math.floor(z)
    fade = lambda t: t * t * t * (t * (t * 6 - 15) + 10)
    lerp = lambda t, a, b: a + t * (b - a)

    def grad(hh, x, y, z):
        h = hh & 15
        u = x if h < 8 else y
        v = y if h < 4 else (x if (h == 12 or h == 14) else z)
        return (u if (h & 1) == 0 else -u) + (v if (h & 2) == 0 else -v)
    u, v, w = fade(fx), fade(fy), fade(fz)
    p = p512
    A = p[X] + Y; AA = p[A] + Z; AB = p[A + 1] + Z
    B = p[X + 1] + Y; BA = p[B] + Z; BB = p[B + 1] + Z
    return lerp(w,
                lerp(v, lerp(u, grad(p[AA], fx, fy, fz), grad(p[BA], fx - 1, fy, fz)),
                     lerp(u, grad(p[AB], fx, fy - 1, fz), grad(p[BB], fx - 1, fy - 1, fz))),
                lerp(v, lerp(u, grad(p[AA + 1], fx, fy, fz - 1), grad(p[BA + 1], fx - 1, fy, fz - 1)),
                     lerp(u, grad(p[AB + 1], fx, fy - 1, fz - 1), grad(p[BB + 1], fx - 1, fy - 1, fz - 1))))


def background(world, d):
    kind, I = world["bg"], world["sky"]
    if kind == "solid":
        return mul(world["bg_color"], I)
    dr = normalize(d)
    if kind == "procedural_sky":
        sun = max(0, dot(dr, normalize((0.3, 0.6, 0.8))))
        sun_col = mul((1.0, 0.95, 0.8), math.pow(sun, 512) * 10)
        sky = mul((0.4, 0.7, 1.0), max(0, dr[1]) * 0.8)
        glow = mul((1.0, 0.8, 0.6), math.exp(-abs(dr[1]) * 4) * 0.3)
        ground = mul((0.1, 0.15, 0.1), max(0, -dr[1] * 0.5))
        cloud = max(0, perlin(world["perm"], (dr[0] * 10, dr[1] * 3 + 2, dr[2] * 10)) * 0.8 + 0.2)
        cloud_col = mul((0.9, 0.9, 1.0), cloud * max(0, dr[1]) * 0.5)
        return mul(add(add(add(add(sky, glow), ground), sun_col), cloud_col), I)
    if kind == "hdri":
        sun = max(0, dot(dr, normalize((-0.3, 0.6, -0.5))))
        mask = 1.0 if sun > (1.0 - 0.04) else 0.0
        sun_col = mul((1.0, 0.95, 0.8), mask * 20)
        corona = max(0, (sun - (1.0 - 0.2)) / 0.2)
        corona_col = mul((1.0, 0.8, 0.6), math.pow(corona, 2) * 3)
        y = dr[1]
        sky = mul((0.3, 0.5, 0.8), max(0, y * 0.5 + 0.5) * 2)
        ground = mul((0.2, 0.15, 0.1), max(0, -y * 0.3))
        scat = mul((0.8, 0.9, 1.0), math.pow(max(0, 1.0 - abs(y)), 2) * 0.3)
        return mul(add(add(add(add(sky, ground), scat), sun_col), corona_col), I)
    t = 0.5 * (dr[1] + 1.0)
    return mul(add(mul((1.0, 1.0, 1.0), 1.0 - t), mul((0.5, 0.7, 1.0), t)), I)


# ------------------------------------------------------------------------------------------------ materials.js
def scatter(mat, d, h, rng):
    """-> (direction, attenuation) or None"""
    kind = mat["type"]
    if kind == "lambertian":
        return add(h["n"], normalize(in_unit_sphere(rng))), mat["color"]
    if kind == "metal":
        refl = reflect(normalize(d), h["n"])
        out = add(refl, mul(in_unit_sphere(rng), mat["roughness"]))
        return (out, mat["color"]) if dot(out, h["n"]) > 0 else None
    if kind == "dielectric":
        ratio = (1.0 / mat["ior"]) if h["front"] else mat["ior"]
        ud = normalize(d)
        cos_t = min(dot(mul(ud, -1), h["n"]), 1.0)
        sin_t = math.sqrt(1.0 - cos_t * cos_t)
        cannot = ratio * sin_t > 1.0

        def reflectance(c, ri):
            r0 = (1 - ri) / (1 + ri)
            r0 = r0 * r0
            return r0 + (1 - r0) * math.pow((1 - c), 5)
        if cannot or reflectance(cos_t, ratio) > rng.random():
            out = reflect(ud, h["n"])
        else:
            c2 = min(dot(mul(ud, -1), h["n"]), 1.0)
            perp = mul(add(ud, mul(h["n"], c2)), ratio)
            par = mul(h["n"], -math.sqrt(abs(1.0 - dot(perp, perp))))
            out = add(perp, par)
        return out, (1.0, 1.0, 1.0)
    return None


def emitted(mat):
    return mul(mat["color"], mat["intensity"]) if mat["type"] == "emissive" else (0.0, 0.0, 0.0)


# ------------------------------------------------------------------------------------------------ camera.js
def make_camera(look_from, look_at, vup, vfov, aspect, aperture, focus, kind="perspective"):
    theta = vfov * math.pi / 180
    hh = math.tan(theta / 2)
    vh = 2.0 * hh
    vw = aspect * vh
    w = normalize(sub(look_from, look_at))
    u = normalize(cross(vup, w))
    v = cross(w, u)
    if kind == "perspective":
        H, V = mul(u, vw * focus), mul(v, vh * focus)
        llc = sub(sub(sub(look_from, div(H, 2)), div(V, 2)), mul(w, focus))
    else:
        H, V = mul(u, vw), mul(v, vh)
        llc = sub(sub(look_from, div(H, 2)), div(V, 2))
    return dict(o=look_from, H=H, V=V, llc=llc, u=u, v=v, w=w, lens=aperture / 2, kind=kind)


def get_ray(cam, s, t, rng):
    if cam["kind"] == "orthographic":
        off = mul(in_unit_disk(rng), cam["lens"])
        o = add(add(cam["o"], mul(cam["u"], off[0])), mul(cam["v"], off[1]))
        d = add(sub(add(add(cam["llc"], mul(cam["H"], s)), mul(cam["V"], t)), o), mul(cam["w"], -1))
        return o, normalize(d)
    rd = mul(in_unit_disk(rng), cam["lens"])
    off = add(mul(cam["u"], rd[0]), mul(cam["v"], rd[1]))
    o = add(cam["o"], off)
    return o, sub(add(add(cam["llc"], mul(cam["H"], s)), mul(cam["V"], t)), o)


# ------------------------------------------------------------------------------------------------ ray-tracer.js
def ray_color(world, o, d, depth, rng):
    if depth <= 0:
        return (0.0, 0.0, 0.0)
    h = world_hit(world, o, d, 0.001, INF)
    if h:
        e = emitted(h["mat"])
        sc = scatter(h["mat"], d, h, rng)
        if sc:
            c = ray_color(world, h["p"], sc[0], depth - 1, rng)
            return add(e, (sc[1][0] * c[0], sc[1][1] * c[1], sc[1][2] * c[2]))
        return e
    return background(world, d)


def tone_map(c, mode, exposure):
    if mode == "aces":
        def aces(x):
            a, b, cc, dd, e = 2.51, 0.03, 2.43, 0.59, 0.14
            return max(0, (x * (a * x + b)) / (x * (cc * x + dd) + e))
        m = mul(c, exposure)
        return (aces(m[0]), aces(m[1]), aces(m[2]))
    if mode == "linear":
        return mul(c, exposure)
    m = mul(c, exposure)
    return (m[0] / (1.0 + m[0]), m[1] / (1.0 + m[1]), m[2] / (1.0 + m[2]))


def render(world, cam, W, H, spp, depth, seed, aa="supersampling", tonemap="reinhard", exposure=1.0, gamma=2.2,
           denoise=False, strength=0.5):
    lin = np.zeros((H, W, 3), np.float64)
    fdat = np.zeros((H, W, 4), np.float32)
    rgba = np.zeros((H, W, 4), np.uint8)
    q = lambda c: int(min(255, max(0, math.floor(c * 255))))
    for j in range(H - 1, -1, -1):
        for i in range(W):
            color = (0.0, 0.0, 0.0)
            n = 1 if aa == "none" else spp
            row = H - 1 - j
            for s in range(n):
                rng = Philox(seed, row * W + i, s)
                if aa == "stochastic":
                    r1, r2 = rng.random(), rng.random()
                    ox = math.sqrt(r1) * math.cos(2 * math.pi * r2)
                    oy = math.sqrt(r1) * math.sin(2 * math.pi * r2)
                    u, v = (i + 0.5 + ox * 0.5) / W, (j + 0.5 + oy * 0.5) / H
                elif aa == "supersampling":
                    u = (i + rng.random()) / W
                    v = (j + rng.random()) / H
                else:
                    u, v = (i + 0.5) / W, (j + 0.5) / H
                o, d = get_ray(cam, u, v, rng)
                color = add(color, ray_color(world, o, d, depth, rng))
            color = div(color, n)
            lin[row, i] = color
            c = tone_map(color, tonemap, exposure)
            ig = 1.0 / gamma
            c = (math.pow(max(0, c[0]), ig), math.pow(max(0, c[1]), ig), math.pow(max(0, c[2]), ig))
            fdat[row, i] = (c[0], c[1], c[2], 1.0)
            rgba[row, i] = (q(c[0]), q(c[1]), q(c[2]), 255)
    if denoise:
        out = np.zeros_like(fdat)
        for y in range(H):
            for x in range(W):
                acc, wsum = [0.0, 0.0, 0.0], 0.0
                for ky in (-1, 0, 1):
                    for kx in (-1, 0, 1):
                        nx, ny = max(0, min(W - 1, x + kx)), max(0, min(H - 1, y + ky))
                        w = math.exp(-(kx * kx + ky * ky) / (2 * strength * strength))
                        for k in range(3):
                            acc[k] += float(fdat[ny, nx, k]) * w
                        wsum += w
                out[y, x] = (acc[0] / wsum, acc[1] / wsum, acc[2] / wsum, fdat[y, x, 3])
                rgba[y, x] = (q(float(out[y, x, 0])), q(float(out[y, x, 1])), q(float(out[y, x, 2])), 255)
    return lin, fdat, rgba


# ------------------------------------------------------------------------------------------------ scenes (scene-loader.js, well-formed input)
def _mat(m):
    if not m or not m.get("type"):
        return dict(type="lambertian", color=(0.8, 0.8, 0.8))
    t = m["type"].lower()
    v3 = lambda a: (float(a[0]), float(a[1]), float(a[2]))
    if t == "lambertian": return dict(type=t, color=v3(m["color"]))
    if t == "metal": return dict(type=t, color=v3(m["color"]), roughness=min(m.get("roughness", 0), 1))
    if t == "dielectric": return dict(type=t, ior=m.get("ior", 1.5))
    if t == "emissive": return dict(type=t, color=v3(m["color"]), intensity=m.get("intensity", 1))
    return dict(type="lambertian", color=(0.8, 0.8, 0.8))


def load_scene(js, W, H, perm=None):
    v3 = lambda a: (float(a[0]), float(a[1]), float(a[2]))
    objs = []
    for o in js.get("objects", []):
        t, m = o["type"].lower(), _mat(o.get("material"))
        if t == "sphere": objs.append(dict(kind=t, center=v3(o["center"]), radius=o.get("radius") or 1.0, mat=m))
        elif t == "plane": objs.append(dict(kind=t, point=v3(o["point"]), normal=normalize(v3(o["normal"])), mat=m))
        elif t == "box": objs.append(dict(kind=t, min=v3(o["min"]), max=v3(o["max"]), mat=m))
        elif t == "triangle": objs.append(dict(kind=t, tri=make_triangle(v3(o["v0"]), v3(o["v1"]), v3(o["v2"])), mat=m))
        elif t == "mesh":
            vs, idx, tris = [v3(v) for v in o["vertices"]], o["indices"], []
            for k in range(0, len(idx), 3):
                if k + 2 >= len(idx):
                    continue
                a, b, c = idx[k], idx[k + 1], idx[k + 2]
                if a >= len(vs) or b >= len(vs) or c >= len(vs):
                    continue
                tris.append(make_triangle(vs[a], vs[b], vs[c]))
            objs.append(dict(kind=t, tris=tris, mat=m))
    bg = js.get("background") or {}
    world = dict(objects=objs, bg=bg.get("type", "gradient"), bg_color=v3(bg.get("color", [0.1, 0.1, 0.1])),
                 sky=bg.get("intensity", 1.0), perm=list(perm) * 2 if perm is not None else list(range(256)) * 2)
    c = js["camera"]
    pos, at_, up = v3(c.get("position", [0, 0, 5])), v3(c.get("lookAt", [0, 0, 0])), v3(c.get("up", [0, 1, 0]))
    fd = c["focusDist"] if "focusDist" in c else length(sub(pos, at_))
    cam = make_camera(pos, at_, up, c.get("fov", 45), c.get("aspect") or W / H, c.get("aperture", 0.0), fd, c.get("type") or "perspective")
    return world, cam


def preset(name, W, H):
    """ray-tracer.js:42-77 (default), :336-364 (glass), :366-398 (metal), :400-435 (cornell) — as scene dicts."""
    S = lambda c, r, m: dict(type="sphere", center=c, radius=r, material=m)
    P = lambda p, n, m: dict(type="plane", point=p, normal=n, material=m)
    B = lambda a, b, m: dict(type="box", min=a, max=b, material=m)
    lam = lambda c: dict(type="lambertian", color=c)
    met = lambda c, r: dict(type="metal", color=c, roughness=r)
    gl = lambda i: dict(type="dielectric", ior=i)
    em = lambda c, i: dict(type="emissive", color=c, intensity=i)
    cam = dict(position=[3, 2, 2], lookAt=[0, 0, -1], up=[0, 1, 0], fov=45, aspect=W / H, aperture=0.0, focusDist=10.0)
    bg = dict(type="gradient", intensity=1.0)
    if name == "glass":
        objs = [P([0, -0.5, 0], [0, 1, 0], lam([0.8, 0.8, 0.0])), S([0, 0, -1], 0.5, gl(1.5)), S([0, 0, -1], -0.45, gl(1.5)),
                S([-1, 0, -1], 0.5, gl(2.4)), S([1, 0, -1], 0.5, gl(1.5)), S([0, 4, -1], 1, em([1, 1, 1], 8))]
        cam.update(aperture=0.02, focusDist=math.sqrt(3 * 3 + 2 * 2 + 3 * 3))
    elif name == "metal":
        m2 = met([0.8, 0.6, 0.2], 0.1)
        objs = [P([0, -0.5, 0], [0, 1, 0], lam([0.5, 0.5, 0.5])), S([0, 0, -1], 0.5, met([0.8, 0.8, 0.9], 0.0)), S([-1, 0, -1], 0.5, m2),
                S([1, 0, -1], 0.5, met([0.7, 0.6, 0.5], 0.3)), B([-0.3, -0.5, -2], [0.3, 0.3, -1.4], m2), S([2, 3, 0], 0.5, em([1, 0.8, 0.6], 10))]
        cam.update(position=[4, 2, 3])
    elif name == "cornell":
        red, white, green = lam([0.65, 0.05, 0.05]), lam([0.73, 0.73, 0.73]), lam([0.12, 0.45, 0.15])
        objs = [P([0, 0, -5], [0, 0, 1], white), P([0, -2.5, 0], [0, 1, 0], white), P([0, 2.5, 0], [0, -1, 0], white),
                P([-2.5, 0, 0], [1, 0, 0], red), P([2.5, 0, 0], [-1, 0, 0], green),
                B([-1, -2.5, -3.5], [-0.2, -1, -2.7], white), B([0.2, -2.5, -4], [1.2, -0.5, -3], white),
                S([-0.6, -1.8, -2.2], 0.7, gl(1.5)), S([0.7, -1.8, -3.5], 0.7, met([0.8, 0.85, 0.88], 0.0)),
                B([-0.5, 2.45, -3.5], [0.5, 2.49, -2.5], em([1, 1, 1], 15))]
        cam.update(position=[0, 0, 2], fov=40)
        bg = dict(type="solid", color=[0, 0, 0], intensity=1.0)
    else:
        objs = [P([0, -0.5, 0], [0, 1, 0], lam([0.5, 0.5, 0.5])), S([0, 0, -1], 0.5, lam([0.7, 0.3, 0.3])), S([-1, 0, -1], 0.5, gl(1.5)),
                S([1, 0, -1], 0.5, met([0.8, 0.8, 0.9], 0.1)), S([0, 1.5, -1], 0.3, em([1, 1, 1], 5))]
    return dict(objects=objs, camera=cam, background=bg)


def shuffled_perm(seed):
    """PerlinNoise ctor shuffle (noise.js:7-13) with numpy's default_rng in place of Math.random (same as the oracle's helper)."""
    rng = np.random.default_rng(seed)
    p = list(range(256))
    for i in range(255, -1, -1):
        j = int(math.floor(rng.random() * (i + 1)))
        p[i], p[j] = p[j], p[i]
    return p


def cases():
    fx = lambda n: json.load(open(os.path.join(HERE, n)))
    ss, sm = fx("sample_scene.json"), fx("sample_mesh.json")
    mirror = json.loads(json.dumps(ss))
    mirror["objects"][0]["material"] = dict(type="metal", color=[0.9, 0.9, 0.9], roughness=0.05)
    ortho = json.loads(json.dumps(ss)); ortho["camera"]["type"] = "orthographic"
    out = []
    add_case = lambda name, scene, W, H, **kw: out.append((name, scene, W, H, kw))
    add_case("sample_scene", ss, 18, 12, spp=3, depth=6, seed=11)
    add_case("sample_mesh", sm, 18, 12, spp=3, depth=6, seed=12)
    for k, p in enumerate(("default", "glass", "metal", "cornell")):
        add_case("preset_" + p, preset(p, 15, 10), 15, 10, spp=2, depth=7, seed=20 + k)
    for k, bgk in enumerate(("hdri", "procedural_sky", "solid")):
        sc = json.loads(json.dumps(mirror)); sc["background"] = dict(type=bgk, intensity=0.9, color=[0.2, 0.3, 0.5])
        add_case("bg_" + bgk, sc, 12, 8, spp=2, depth=5, seed=30 + k, perm_seed=5)
    add_case("aa_stochastic_aces", ss, 12, 8, spp=3, depth=4, seed=40, aa="stochastic", tonemap="aces", exposure=1.5, gamma=1.8)
    add_case("aa_none_linear", ss, 12, 8, spp=3, depth=4, seed=41, aa="none", tonemap="linear", exposure=0.7, gamma=2.2)
    add_case("denoise", ss, 12, 8, spp=2, depth=4, seed=42, denoise=True, strength=0.5)
    add_case("orthographic", ortho, 12, 8, spp=2, depth=4, seed=43)
    return out


def run_case(scene, W, H, spp, depth, seed, perm_seed=None, **kw):
    perm = shuffled_perm(perm_seed) if perm_seed is not None else None
    world, cam = load_scene(scene, W, H, perm)
    return render(world, cam, W, H, spp, depth, seed, **kw)


if __name__ == "__main__":
    arrays, meta = {}, []
    for name, scene, W, H, kw in cases():
        lin, fdat, rgba = run_case(scene, W, H, **kw)
        arrays[name + "_linear"], arrays[name + "_float"], arrays[name + "_rgba"] = lin, fdat, rgba
        meta.append(dict(name=name, W=W, H=H, scene=scene, **kw))
        print(name, lin.mean())
    np.savez_compressed(os.path.join(HERE, "independent_vectors.npz"), meta=json.dumps(meta), **arrays)
