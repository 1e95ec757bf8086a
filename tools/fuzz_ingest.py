"""Differential fuzz of scene ingest: random, type-correct but value-weird scene JSON (missing fields, zeros, negatives, short and
long arrays, odd capitalisation, unknown types, out-of-range mesh indices, cameras on top of their target, resolution overrides)
goes through the reference's OWN loader (js/scene-loader.js executed by baseline/minijs.py) and through libbrt's native loader
(csrc/scene_loader.cpp, host-only context); the resulting objects, materials, lights, mesh triangles, camera vectors, background
and canvas size must be the same doubles.  Needs a checkout of the reference (never copied).

    python tools/fuzz_ingest.py [--seed 1] [--n 300] [--ref /root/reference]"""
import argparse
import ctypes as C
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

NUMS = [0, 1, -1, 0.5, 2.5, 1e-7, 1e6, -0.25, 3, 10, 0.001, 45]


def gen_scene(r):
    def num(): return r.choice(NUMS) if r.random() < 0.7 else round(r.uniform(-5, 5), 3)
    def maybe(d, key, make, p=0.75):
        if r.random() < p: d[key] = make()
    def vec():
        k = r.random()
        n = 3 if k < 0.8 else 2 if k < 0.87 else 4 if k < 0.94 else 0
        return [num() for _ in range(n)]
    def color(): return [round(r.random(), 3) for _ in range(3)] if r.random() < 0.85 else vec()
    def case(s): return r.choice([s, s, s.capitalize(), s.upper()])
    def material():
        m = {}
        maybe(m, "type", lambda: r.choice([case("lambertian"), case("metal"), case("dielectric"), case("emissive"), "plastic", ""]), 0.9)
        maybe(m, "color", color)
        maybe(m, "roughness", lambda: r.choice([0, 0.25, 0.5, 1, 7, -1]), 0.5)
        maybe(m, "ior", lambda: r.choice([0, 1, 1.33, 1.5, 2.4]), 0.5)
        maybe(m, "intensity", lambda: r.choice([0, 1, 4.5, 15]), 0.5)
        return m
    def obj():
        o = {}
        maybe(o, "type", lambda: r.choice([case("sphere"), case("plane"), case("box"), case("triangle"), case("mesh"), "torus", ""]), 0.95)
        maybe(o, "material", material, 0.8)
        t = str(o.get("type", "")).lower()
        if t == "sphere": maybe(o, "center", vec, 0.9); maybe(o, "radius", lambda: r.choice([0, 1, 0.5, -0.45, 100, 2.5]), 0.85)
        elif t == "plane": maybe(o, "point", vec, 0.9); maybe(o, "normal", lambda: r.choice([[0, 1, 0], [0, 0, 0], [0, 5, 0], vec()]), 0.9)
        elif t == "box": maybe(o, "min", vec, 0.9); maybe(o, "max", vec, 0.9)
        elif t == "triangle":
            for k in ("v0", "v1", "v2"): maybe(o, k, vec, 0.92)
        elif t == "mesh":
            nv = r.choice([0, 1, 3, 4, 8])
            maybe(o, "vertices", lambda: [vec() for _ in range(nv)], 0.92)
            def idx():
                n = r.choice([0, 3, 6, 7, 12, 14])
                return [r.choice([r.randrange(0, max(1, nv)), r.randrange(0, max(1, nv)), nv, nv + 3, -1]) if r.random() < 0.25 else r.randrange(0, max(1, nv)) for _ in range(n)]
            maybe(o, "indices", idx, 0.92)
        return o
    def light():
        l = {}
        maybe(l, "type", lambda: r.choice([case("point"), case("directional"), "spot", ""]), 0.9)
        maybe(l, "position", vec, 0.7); maybe(l, "direction", vec, 0.7); maybe(l, "color", color, 0.6)
        maybe(l, "intensity", lambda: r.choice([0, 1, 3, 12.5]), 0.6)
        return l
    s = {}
    maybe(s, "objects", lambda: [obj() for _ in range(r.randrange(0, 7))], 0.95)
    maybe(s, "lights", lambda: [light() for _ in range(r.randrange(0, 4))], 0.6)
    def camera():
        c = {}
        pos = vec()
        maybe(c, "position", lambda: pos, 0.85)
        maybe(c, "lookAt", lambda: r.choice([vec(), pos, [v + 0.1 for v in pos] if len(pos) >= 3 else vec()]), 0.85)
        maybe(c, "up", lambda: r.choice([[0, 1, 0], [0, 0, 1], [0, 0, 0], vec()]), 0.6)
        maybe(c, "fov", lambda: r.choice([0, 20, 45, 90, 179]), 0.8)
        maybe(c, "aspect", lambda: r.choice([0, 1, 1.5, 16 / 9]), 0.5)
        maybe(c, "aperture", lambda: r.choice([0, 0.1, 2]), 0.6)
        maybe(c, "focusDist", lambda: r.choice([0, 1, 10, 4.5]), 0.5)
        maybe(c, "type", lambda: r.choice(["perspective", "orthographic", "fisheye", ""]), 0.6)
        maybe(c, "resolution", lambda: r.choice([[320, 200], [64, 64], [1920, 1080]]), 0.25)
        return c
    maybe(s, "camera", camera, 0.8)
    def background():
        b = {}
        maybe(b, "type", lambda: r.choice(["gradient", "solid", "hdri", "procedural_sky", "weird", ""]), 0.9)
        maybe(b, "color", color, 0.5)
        maybe(b, "intensity", lambda: r.choice([0, 0.5, 1, 2]), 0.6)
        return b
    maybe(s, "background", background, 0.7)
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    bad = run(args.seed, args.n, args.ref, verbose=True)
    print(f"{args.n} scenes, {len(bad)} disagreements")
    sys.exit(1 if bad else 0)


def run(seed, n, ref="/root/reference", verbose=False):
    import minijs as J
    import make_fixtures_minijs as M
    from make_host_fixtures_minijs import dump_state
    import blenderraytracer_b200 as brt
    from blenderraytracer_b200 import _lib as L
    from test_reference_host_pin import check_state, resize_canvas_camera
    sys.setrecursionlimit(20000)
    interp, RayTracer, Vec3 = M.load_reference(os.path.join(ref, "js"))
    lib = brt.load()
    r = random.Random(seed)
    bad = []
    for k in range(n):
        scene = gen_scene(r)
        rt = interp.construct(RayTracer, [M.fake_canvas(interp, 600, 400)])
        try:
            ok = J.truthy(M.method(interp, rt, "loadFromJSON", J.py_to_js(json.loads(json.dumps(scene)))))
        except J.JSThrow as e:                                         # the reference catches its own errors: this would be a harness problem
            bad.append((k, "reference threw: " + J.to_str(e.value), scene)); continue
        h = C.c_void_p()
        assert lib.brt_create(C.byref(h), -1) == L.BRT_OK
        try:
            text = json.dumps(scene).encode()
            hc, w, hh = C.c_int(), C.c_int(), C.c_int()
            rc = lib.brt_scene_load_json(h, text, len(text), 600, 400, C.byref(hc), C.byref(w), C.byref(hh))
            if (rc == L.BRT_OK) != ok:
                bad.append((k, f"reference ok={ok}, libbrt rc={rc} ({lib.brt_last_error(h)})", scene)); continue
            if not ok: continue
            W, H = (w.value, hh.value) if w.value and hh.value else (600, 400)
            if w.value and hh.value and hc.value: resize_canvas_camera(lib, h, W, H)
            st = dump_state(rt)
            if (W, H) != (st["width"], st["height"]):
                bad.append((k, f"canvas {W}x{H} vs reference {st['width']}x{st['height']}", scene)); continue
            if not hc.value: st = dict(st, camera=None)                # no camera in the JSON: the reference keeps its constructor's camera
            try:
                check_state(lib, h, st, f"fuzz {seed}/{k}")
            except AssertionError as e:
                bad.append((k, str(e)[:400], scene))
        finally:
            lib.brt_destroy(h)
    if verbose:
        for k, why, scene in bad[:10]:
            print(f"--- scene {k}: {why}\n{json.dumps(scene)}")
    return bad


if __name__ == "__main__":
    main()
