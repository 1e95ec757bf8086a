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
    ]


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
        interp, RayTracer, Vec3 = M.load_reference(js_dir)
        W, H = c["W"], c["H"]
        rt = interp.construct(RayTracer, [M.fake_canvas(interp, W, H)])
        assert J.truthy(M.method(interp, rt, "loadFromJSON", J.py_to_js(json.loads(json.dumps(c["scene"])))))
        if not (c["scene"].get("camera") or {}).get("resolution"):
            pass
        interp.globals.vars["Math"].set("random", J.native(lambda t, a: 0.5))
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
        out["cases"].append(dict(name=c["name"], W=W, H=H, scene=c["scene"], obj_id=obj, tri_id=tri, t=ts, normal=nrm, front_face=ff))
        print(c["name"], f"{W}x{H}", "hits", sum(1 for x in obj if x >= 0), "objects", len(world.get("objects").items), flush=True)
    json.dump(out, open(args.out, "w"))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
