"""Derives tests/golden/kat.json — the known-answer vectors that pin the oracle.

The reference ships no tests or golden vectors and cannot be executed here (no JavaScript engine), so these vectors
are derived by following the cited reference formulas by hand in plain Python float64 (identical to JS Numbers for
+ - * / sqrt).  This script is deliberately independent of oracle/ and of blenderraytracer_b200/: it is a third,
tiny, closed-form evaluation for a handful of pixels, so agreement between it and the C++ oracle is meaningful.

    python tests/golden/derive_kat.py        # rewrites kat.json (committed)
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
def add(a, b): return [a[0] + b[0], a[1] + b[1], a[2] + b[2]]
def mul(a, s): return [a[0] * s, a[1] * s, a[2] * s]
def div(a, s): return [a[0] / s, a[1] / s, a[2] / s]
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
def length(a): return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
def norm(a):
    l = length(a)
    return div(a, l) if l > 0 else [0.0, 0.0, 0.0]
def reflect(v, n): return sub(v, mul(n, 2 * dot(v, n)))


def camera(look_from, look_at, vup, vfov, aspect, aperture, focus):          # camera.js:8-36 (perspective)
    theta = vfov * math.pi / 180
    h = math.tan(theta / 2)
    vh = 2.0 * h
    vw = aspect * vh
    w = norm(sub(look_from, look_at))
    u = norm(cross(vup, w))
    v = cross(w, u)
    H = mul(u, vw * focus)
    V = mul(v, vh * focus)
    llc = sub(sub(sub(look_from, div(H, 2)), div(V, 2)), mul(w, focus))
    return dict(origin=look_from, w=w, u=u, v=v, horizontal=H, vertical=V, lowerLeftCorner=llc, lensRadius=aperture / 2)


def primary_ray(cam, i, j, W, H):                                             # ray-tracer.js:144-147, camera.js:45-49 (lens offset 0)
    s, t = (i + 0.5) / W, (j + 0.5) / H
    d = sub(add(add(cam["lowerLeftCorner"], mul(cam["horizontal"], s)), mul(cam["vertical"], t)), cam["origin"])
    return cam["origin"], d


def hit_sphere(o, d, c, r, tmin, tmax):                                       # geometry.js:15-35
    oc = sub(o, c)
    a = dot(d, d); hb = dot(oc, d); cc = dot(oc, oc) - r * r
    disc = hb * hb - a * cc
    if disc < 0: return None
    sq = math.sqrt(disc)
    root = (-hb - sq) / a
    if root < tmin or tmax < root:
        root = (-hb + sq) / a
        if root < tmin or tmax < root: return None
    p = add(o, mul(d, root))
    n = div(sub(p, c), r)
    front = dot(d, n) < 0
    return root, (n if front else mul(n, -1)), front, p


def hit_plane(o, d, p0, n, tmin, tmax):                                       # geometry.js:56-66
    n = norm(n)
    den = dot(n, d)
    if abs(den) < 1e-6: return None
    t = dot(sub(p0, o), n) / den
    if t < tmin or t > tmax: return None
    front = dot(d, n) < 0
    return t, (n if front else mul(n, -1)), front, add(o, mul(d, t))


def hit_tri(o, d, v0, v1, v2, tmin, tmax):                                    # geometry.js:148-180
    e1, e2 = sub(v1, v0), sub(v2, v0)
    h = cross(d, e2); a = dot(e1, h)
    if abs(a) < 0.0001: return None
    f = 1.0 / a; s = sub(o, v0); u = f * dot(s, h)
    if u < 0 or u > 1: return None
    q = cross(s, e1); v = f * dot(d, q)
    if v < 0 or u + v > 1: return None
    t = f * dot(e2, q)
    if t < tmin or t > tmax: return None
    n = norm(cross(e1, e2))
    front = dot(d, n) < 0
    return t, (n if front else mul(n, -1)), front, add(o, mul(d, t))


def world_hit(objects, o, d):                                                 # world.js:20-33 (+ geometry.js:248-262 for meshes)
    best, closest = None, math.inf
    for k, ob in enumerate(objects):
        h, tri = None, -1
        if ob["type"] == "sphere": h = hit_sphere(o, d, ob["center"], ob["radius"], 0.001, closest)
        elif ob["type"] == "plane": h = hit_plane(o, d, ob["point"], ob["normal"], 0.001, closest)
        elif ob["type"] == "mesh":
            ct = closest
            idx, vs = ob["indices"], ob["vertices"]
            for ti in range(len(idx) // 3):
                hh = hit_tri(o, d, vs[idx[3 * ti]], vs[idx[3 * ti + 1]], vs[idx[3 * ti + 2]], 0.001, ct)
                if hh: h, tri, ct = hh, ti, hh[0]
        if h and h[0] < closest:
            closest = h[0]
            best = dict(obj=k, tri=tri, t=h[0], normal=h[1], front=h[2], point=h[3])
    return best


def sky(d, intensity=1.0):                                                    # world.js:35-40
    t = 0.5 * (norm(d)[1] + 1.0)
    return mul(add(mul([1.0, 1.0, 1.0], 1.0 - t), mul([0.5, 0.7, 1.0], t)), intensity)


def to_u8(c):                                                                 # reinhard e=1, gamma 2.2, floor (ray-tracer.js:211-228)
    out = []
    for x in c:
        m = x / (1.0 + x)
        g = math.pow(max(0.0, m), 1.0 / 2.2)
        out.append(int(min(255, max(0, math.floor(g * 255)))))
    return out


def main():
    scene = json.load(open(os.path.join(HERE, "sample_scene.json")))
    mesh = json.load(open(os.path.join(HERE, "sample_mesh.json")))
    kat = {"_doc": "hand-derived float64 known-answer vectors (tests/golden/derive_kat.py); pins oracle/ — see SURVEY.md §8c"}

    # ---- KAT-A cameras
    c1 = scene["camera"]; c2 = mesh["camera"]
    camA = camera(c1["position"], c1["lookAt"], c1["up"], c1["fov"], c1["aspect"], c1["aperture"], c1["focusDist"])
    camB = camera(c2["position"], c2["lookAt"], c2["up"], c2["fov"], c2["aspect"], c2["aperture"], c2["focusDist"])
    camB169 = camera(c2["position"], c2["lookAt"], c2["up"], c2["fov"], 16 / 9, c2["aperture"], c2["focusDist"])
    kat["camera"] = {"sample_scene": camA, "sample_mesh": camB, "sample_mesh_16_9": camB169}

    # ---- KAT-B primary hits
    def prim(objs, cam, W, H, pixels):
        rows = []
        for (i, j) in pixels:
            o, d = primary_ray(cam, i, j, W, H)
            h = world_hit(objs, o, d)
            rows.append(dict(i=i, j=j, obj=h["obj"], tri=h["tri"], t=h["t"], normal=h["normal"], front=h["front"]))
        return rows
    kat["primary"] = {
        "sample_scene_600x400": prim(scene["objects"], camA, 600, 400, [(300, 200), (0, 0), (599, 399), (150, 100), (450, 133)]),
        "sample_mesh_1280x720": prim(mesh["objects"], camB, 1280, 720, [(640, 360), (0, 0), (1279, 719)]),
    }

    # ---- KAT-C scalars
    e1, e2 = math.exp(-1 / (2 * 0.25)), math.exp(-2 / (2 * 0.25))
    sdir = norm([-0.3, 0.6, -0.5])

    def hdri(d):                                                              # world.js:74-110
        d = norm(d)
        sd = max(0.0, dot(d, norm([-0.3, 0.6, -0.5])))
        mask = 1.0 if sd > 0.96 else 0.0
        cor = max(0.0, (sd - 0.8) / 0.2)
        y = d[1]
        skyv = max(0.0, y * 0.5 + 0.5); gb = max(0.0, -y * 0.3); sc = math.pow(max(0.0, 1.0 - abs(y)), 2) * 0.3
        col = add(add(add(add(mul([0.3, 0.5, 0.8], skyv * 2), mul([0.2, 0.15, 0.1], gb)), mul([0.8, 0.9, 1.0], sc)),
                      mul([1.0, 0.95, 0.8], mask * 20)), mul([1.0, 0.8, 0.6], math.pow(cor, 2) * 3))
        return col
    r0 = ((1 - 1 / 1.5) / (1 + 1 / 1.5)) ** 2
    kat["scalars"] = dict(
        reinhard_1=0.5, reinhard_1_gamma=math.pow(0.5, 1 / 2.2), reinhard_1_u8=int(math.floor(math.pow(0.5, 1 / 2.2) * 255)),
        aces_1=(1.0 * (2.51 * 1.0 + 0.03)) / (1.0 * (2.43 * 1.0 + 0.59) + 0.14),
        point_att_d10=1.0 / (1.0 + 0.1 * 10 + 0.01 * 10 * 10),
        schlick_r0=r0, schlick_cos05=r0 + (1 - r0) * math.pow(0.5, 5),
        denoise_sigma05=dict(edge=e1, corner=e2, total=1 + 4 * e1 + 4 * e2),
        denoise_sigma1_total=1 + 4 * math.exp(-0.5) + 4 * math.exp(-1.0),
        hdri={"0,1,0": hdri([0, 1, 0]), "1,0,0": hdri([1, 0, 0]), "0,-1,0": hdri([0, -1, 0]), "sun": hdri(sdir)},
        sky_up=sky([0, 1, 0]),
    )

    # ---- KAT-D deterministic multi-bounce: mirror sphere + emissive ground, gradient sky, depth 5, AA none, aperture 0
    objs = [dict(type="sphere", center=[0, 0, -1], radius=0.5, mat=("metal", [0.8, 0.8, 0.8], 0.0)),
            dict(type="plane", point=[0, -0.5, 0], normal=[0, 1, 0], mat=("emissive", [1, 0.5, 0.25], 2.0))]
    camD = camera(c1["position"], c1["lookAt"], c1["up"], c1["fov"], 1.5, 0.0, c1["focusDist"])

    def ray_color(o, d, depth):                                               # ray-tracer.js:102-123 without RNG influence
        if depth <= 0: return [0.0, 0.0, 0.0], []
        h = world_hit(objs, o, d)
        if not h: return sky(d), []
        kind, col, p = objs[h["obj"]]["mat"]
        if kind == "emissive": return mul(col, p), [h["t"]]
        refl = reflect(norm(d), h["normal"])                                  # roughness 0: ball sample * 0
        if dot(refl, h["normal"]) <= 0: return [0.0, 0.0, 0.0], [h["t"]]
        c, ts = ray_color(h["point"], refl, depth - 1)
        return [col[0] * c[0], col[1] * c[1], col[2] * c[2]], [h["t"]] + ts
    rows = []
    for (i, j) in [(300, 200), (300, 230), (280, 190), (300, 170)]:
        o, d = primary_ray(camD, i, j, 600, 400)
        c, ts = ray_color(o, d, 5)
        rows.append(dict(i=i, j=j, linear=c, ts=ts, rgba8=to_u8(c) + [255]))
    kat["deterministic"] = dict(
        scene=dict(objects=[dict(type="sphere", center=[0, 0, -1], radius=0.5, material=dict(type="metal", color=[0.8, 0.8, 0.8], roughness=0.0)),
                            dict(type="plane", point=[0, -0.5, 0], normal=[0, 1, 0], material=dict(type="emissive", color=[1, 0.5, 0.25], intensity=2.0))],
                   camera=dict(position=c1["position"], lookAt=c1["lookAt"], up=c1["up"], fov=c1["fov"], aspect=1.5, aperture=0.0,
                               focusDist=c1["focusDist"], type="perspective"),
                   background=dict(type="gradient", intensity=1.0)),
        width=600, height=400, depth=5, pixels=rows)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("wrote kat.json")


if __name__ == "__main__":
    main()
