#!/usr/bin/env python
"""make_aov_fixtures_minijs.py — primary-visibility AOVs from the reference's OWN World.hit (js/world.js:20-33,
js/geometry.js), executed from the unmodified source by baseline/minijs.py.

For every pixel centre (lens offset 0: Math.random returns 0.5, so Vec3.randomInUnitDisk() is (0, 0)) the harness calls
camera.getRay((i + 0.5) / W, (j + 0.5) / H) and world.hit(ray, 0.001, Infinity) — exactly what rayColor does first
(js/ray-tracer.js:105-106) — and records the hit's t, normal and frontFace.  WHICH object (and which triangle of a mesh) produced
the hit is learnt without re-implementing the loops: every object's and every triangle's `hit` method is wrapped by a recorder
that calls the original and remembers the HitRecord it returned; the winner is the object / triangle whose record World.hit hands
back (identity).  North-star gate: primary-hit object IDs bit-exact.

Writes tests/golden/reference_aov_vectors.json (object id = index in world.objects, triangle id = index in mesh.triangles, -1 = miss /
not a mesh; row 0 = top).

    python baseline/make_aov_fixtures_minijs.py [--ref /root/reference]
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import minijs as J  # noqa: E402
import make_fixtures_minijs as M  # noqa: E402


def cases():
    from tools import gen_scenes
    import make_reference_cases_extra as X                    # tie_scene()
    g = lambda n: json.load(open(os.path.join(ROOT, "tests", "golden", n)))
    return [
        dict(name="sample_scene", W=60, H=40, scene=g("sample_scene.json")),
        dict(name="sample_mesh", W=64, H=36, scene=g("sample_mesh.json")),
        dict(name="ties_duplicates_coplanar", W=54, H=36, scene=X.tie_scene()),
        dict(name="c3_random_spheres", W=48, H=27, scene=gen_scenes.random_spheres(grid=3)),
        dict(name="c4_cornell", W=48, H=27, scene=gen_scenes.cornell("hdri")),
        dict(name="c5_terrain", W=48, H=27, scene=gen_scenes.terrain(quads=10, extent=200.0)),
        dict(name="axis_parallel_rays_and_zero_over_zero", W=45, H=31, scene=axis_parallel_scene()),
        dict(name="orthographic_from_inside", W=40, H=30, scene=inside_scene()),
    ]


def axis_parallel_scene():
    """An axis-aligned camera with odd W and H: the centre column has D.x = 0 and the centre row D.y = 0 EXACTLY, so Box.hit
    (geometry.js:85-112) divides by zero — (bound - O) / 0 = +-Infinity, and 0 / 0 = NaN where a box face lies exactly at the ray
    origin's coordinate (boxes 0 and 1 below) — and Math.max / Math.min propagate the NaN.  Spheres, triangles, a mesh and a plane
    straddle the same axes."""
    lam = lambda c: dict(type="lambertian", color=c)
    objs = [
        dict(type="box", min=[0.5, 0.2, -1.0], max=[1.5, 1.4, 0.0], material=lam([0.8, 0.2, 0.2])),      # min.x = O.x: 0/0 on the centre column
        dict(type="box", min=[-1.5, 1.0, -2.0], max=[-0.2, 1.8, -1.0], material=lam([0.2, 0.8, 0.2])),   # min.y = O.y: 0/0 on the centre row
        dict(type="box", min=[-0.4, -0.5, -3.0], max=[0.4, 0.6, -2.5], material=lam([0.2, 0.2, 0.8])),   # straddles x = 0.5? no: left of it
        dict(type="box", min=[0.1, 0.6, -4.0], max=[0.9, 1.6, -3.5], material=lam([0.8, 0.8, 0.2])),     # straddles both centre lines
        dict(type="sphere", center=[0.5, 1.0, -6.0], radius=1.2, material=lam([0.5, 0.5, 0.5])),
        dict(type="sphere", center=[-1.0, 0.2, 1.0], radius=0.6, material=dict(type="metal", color=[0.9, 0.9, 0.9], roughness=0.1)),
        dict(type="sphere", center=[2.0, 1.9, 0.5], radius=0.7, material=dict(type="dielectric", ior=1.5)),
        dict(type="triangle", v0=[0.5, 1.0, -4.5], v1=[2.6, 1.0, -4.5], v2=[0.5, 3.2, -4.5], material=lam([0.9, 0.4, 0.1])),   # edges ON the centre lines
        dict(type="mesh", vertices=[[-3, -0.5, -4.6], [0.5, -0.5, -4.6], [0.5, 1.0, -4.6], [-3, 1.0, -4.6]], indices=[0, 1, 2, 0, 2, 3], material=lam([0.3, 0.6, 0.9])),
        dict(type="plane", point=[0, -0.5, 0], normal=[0, 1, 0], material=lam([0.4, 0.4, 0.4])),
        dict(type="plane", point=[0, 0, -9], normal=[0, 0, 1], material=lam([0.6, 0.5, 0.4])),
    ]
    return dict(objects=objs, lights=[], camera=dict(position=[0.5, 1.0, 5.0], lookAt=[0.5, 1.0, 0.0], up=[0, 1, 0], fov=50, aspect=45 / 31, aperture=0.0, focusDist=5.0),
                background=dict(type="gradient", intensity=1.0))


def inside_scene():
    """orthographic camera placed INSIDE a large glass sphere and a box: back-face hits (frontFace false, flipped normals), the second
    sphere root (geometry.js:22-26) and Box.hit's `t0 > tMin ? t0 : t1` exit-face choice (:109)"""
    objs = [
        dict(type="sphere", center=[0, 0, 0], radius=3.0, material=dict(type="dielectric", ior=1.5)),
        dict(type="box", min=[-2.5, -2.0, -6.0], max=[2.5, 2.0, 2.0], material=dict(type="lambertian", color=[0.7, 0.7, 0.7])),
        dict(type="sphere", center=[0.8, 0.3, -1.5], radius=-0.5, material=dict(type="dielectric", ior=1.5)),                      # negative radius
        dict(type="triangle", v0=[-1, -1, -2.2], v1=[1, -1, -2.2], v2=[0, 1, -2.2], material=dict(type="metal", color=[0.8, 0.8, 0.8], roughness=0.0)),
    ]
    return dict(objects=objs, lights=[], camera=dict(position=[0.2, 0.1, 1.0], lookAt=[0.2, 0.1, -1.0], up=[0, 1, 0], fov=60, aspect=40 / 30, aperture=0.0,
                                                      focusDist=2.0, type="orthographic"),
                background=dict(type="solid", color=[0.2, 0.2, 0.2], intensity=1.0))


def aov_of_case(js_dir, c):
    """camera.getRay((i + .5) / W, (j + .5) / H) + world.hit(ray, 0.001, Infinity) of the reference for every pixel of case c;
    which object / triangle won is learnt by wrapping their hit methods.  Raises RuntimeError when the reference refuses the scene."""
    interp, RayTracer, Vec3 = M.load_reference(js_dir)
    W, H = c["W"], c["H"]
    rt = interp.construct(RayTracer, [M.fake_canvas(interp, W, H)])
    if not J.truthy(M.method(interp, rt, "loadFromJSON", J.py_to_js(json.loads(json.dumps(c["scene"]))))):
        raise RuntimeError("reference loadFromJSON returned false")
    interp.globals.vars["Math"].set("random", J.native(lambda t, a: 0.5))     # lens offset 0 (camera.js: the disk sample is (0, 0))
    world, cam = rt.get("world"), rt.get("camera")
    last = {}                                              # id(HitRecord) -> (object index, triangle index)
    def wrap(owner, oi, ti):
        orig = owner.get("hit")
        def rec(this, a):
            r = interp.call(orig, this, a)
            if isinstance(r, J.JSObject) and id(r) not in last: last[id(r)] = (oi, ti, r)
            return r
        owner.set("hit", J.native(rec))
    for oi, o in enumerate(world.get("objects").items):
        tris = o.get("triangles")
        if isinstance(tris, J.JSArray):
            for ti, t in enumerate(tris.items): wrap(t, oi, ti)
        else:
            wrap(o, oi, -1)
    obj, tri, ts, nrm, ff = [], [], [], [], []
    for row in range(H):
        j = H - 1 - row
        for i in range(W):
            last.clear()
            ray = M.method(interp, cam, "getRay", (i + 0.5) / W, (j + 0.5) / H)
            hit = M.method(interp, world, "hit", ray, 0.001, float("inf"))
            if not isinstance(hit, J.JSObject):
                obj.append(-1); tri.append(-1); ts.append(None); nrm.append([0.0, 0.0, 0.0]); ff.append(0); continue
            oi, ti, _ = last[id(hit)]
            n = hit.get("normal")
            obj.append(oi); tri.append(ti); ts.append(hit.get("t")); nrm.append([n.get("x"), n.get("y"), n.get("z")]); ff.append(1 if J.truthy(hit.get("frontFace")) else 0)
    return dict(name=c["name"], W=W, H=H, scene=c["scene"], obj_id=obj, tri_id=tri, t=ts, normal=nrm, front_face=ff)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_aov_vectors.json"))
    args = ap.parse_args()
    js_dir = os.path.join(args.ref, "js")
    sys.setrecursionlimit(20000)
    out = {"generator": "baseline/make_aov_fixtures_minijs.py: camera.getRay + world.hit of the unmodified reference executed by baseline/minijs.py", "cases": []}
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        all_cases = cases()
    for c in all_cases:
        rec = aov_of_case(js_dir, c)
        out["cases"].append(rec)
        print(c["name"], f"{c['W']}x{c['H']}", "hits", sum(1 for x in rec["obj_id"] if x >= 0), flush=True)
    json.dump(out, open(args.out, "w"))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
