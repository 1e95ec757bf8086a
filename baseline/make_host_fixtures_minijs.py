#!/usr/bin/env python
"""make_host_fixtures_minijs.py — what the reference's OWN scene loader and camera make of a JSON scene.

Executes the unmodified reference (js/ray-tracer.js -> RayTracer.loadFromJSON -> js/scene-loader.js, js/camera.js, js/geometry.js,
js/materials.js, js/lights.js) through baseline/minijs.py for a list of ingest cases — the two shipped scenes and the scenes of
tests/test_host_abi.py that probe every default, `||` fallback, skip rule, clamp, mesh-index filter, camera push and resolution
override — and dumps the resulting World / Camera objects field by field into tests/golden/reference_host_vectors.json.
tests/test_reference_host_pin.py compares libbrt's native ingest (scene_loader.cpp, host-only context) with that file: object
order and kinds, every coordinate, materials, lights, mesh triangles, the derived camera vectors, canvas size — exact doubles.

    python baseline/make_host_fixtures_minijs.py [--ref /root/reference]
"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import minijs as J  # noqa: E402
import make_fixtures_minijs as M  # noqa: E402


def cls_name(o):
    c = o.proto.get("constructor") if isinstance(o, J.JSObject) and o.proto is not None else J.UNDEF
    return c.name if isinstance(c, J.JSFunction) else ""


def vec(v):
    return [v.get("x"), v.get("y"), v.get("z")] if isinstance(v, J.JSObject) else None


def num(v):
    return v if isinstance(v, float) else None


def dump_material(m):
    if not isinstance(m, J.JSObject):
        return None
    d = {"cls": cls_name(m)}
    for k in ("albedo", "color"):
        if isinstance(m.get(k), J.JSObject): d[k] = vec(m.get(k))
    for k in ("roughness", "refractionIndex", "intensity"):
        if isinstance(m.get(k), float): d[k] = m.get(k)
    return d


def dump_object(o):
    d = {"cls": cls_name(o), "material": dump_material(o.get("material"))}
    for k in ("center", "point", "normal", "min", "max", "v0", "v1", "v2"):
        if isinstance(o.get(k), J.JSObject): d[k] = vec(o.get(k))
    if isinstance(o.get("radius"), float): d["radius"] = o.get("radius")
    tris = o.get("triangles")
    if isinstance(tris, J.JSArray):
        d["triangles"] = [[vec(t.get("v0")), vec(t.get("v1")), vec(t.get("v2"))] for t in tris.items]
        if d["material"] is None and tris.items:            # a mesh keeps its material in its triangles (geometry.js:231)
            d["material"] = dump_material(tris.items[0].get("material"))
    return d


def dump_light(l):
    d = {"cls": cls_name(l), "color": vec(l.get("color")), "intensity": num(l.get("intensity"))}
    for k in ("position", "direction"):
        if isinstance(l.get(k), J.JSObject): d[k] = vec(l.get(k))
    return d


def dump_camera(c):
    if not isinstance(c, J.JSObject):
        return None
    d = {k: vec(c.get(k)) for k in ("origin", "lowerLeftCorner", "horizontal", "vertical", "u", "v", "w")}
    d.update(type=c.get("type"), lensRadius=num(c.get("lensRadius")), fov=num(c.get("fov")), aperture=num(c.get("aperture")), focusDist=num(c.get("focusDist")))
    return d


def dump_state(rt):
    world = rt.get("world")
    bg = world.get("background")
    return dict(objects=[dump_object(o) for o in world.get("objects").items], lights=[dump_light(l) for l in world.get("lights").items],
                camera=dump_camera(rt.get("camera")), background=(bg.name if isinstance(bg, J.JSFunction) else None),
                skyIntensity=num(world.get("skyIntensity")), width=rt.get("width"), height=rt.get("height"))


def tricky_scene():
    """tests/test_host_abi.py::test_ingest_defaults_and_skip_rules"""
    return dict(
        objects=[
            dict(type="Sphere", center=[1, 2, 3]),
            dict(type="sphere", center=[0, 0, 0], radius=0, material=dict(type="METAL", color=[1, 1, 1], roughness=7)),
            dict(type="torus", center=[0, 0, 0]),
            dict(center=[9, 9, 9]),
            dict(type="plane", point=[0, -1, 0], normal=[0, 5, 0], material=dict(type="dielectric")),
            dict(type="box", min=[0, 0], max=[1, 1, 1], material=dict(type="emissive", color=[1, 0.5, 0.25])),
            dict(type="mesh", vertices=[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]],
                 indices=[0, 1, 2, 1, 3, 2, 0, 1, 9, 0, 1, -1, 0, 1, 2.5, 3, 2], material=dict(type="plastic")),
            dict(type="mesh", vertices=[[0, 0, 0]]),
            dict(type="triangle", v0=[0, 0, 0], v1=[1, 0, 0], v2=[0, 1, 0], material=dict(type="lambertian", color=[0.1, 0.2, 0.3])),
        ],
        lights=[dict(type="point", position=[1, 2, 3]), dict(type="DIRECTIONAL", direction=[0, -2, 0], color=[1, 0, 0], intensity=3),
                dict(type="spot"), dict(position=[0, 0, 0])],
        camera=dict(position=[0, 0, 0.5], lookAt=[0, 0, 0], fov=30),
        background=dict(type="procedural_sky", intensity=0.5),
    )


def cases():
    g = lambda n: json.load(open(os.path.join(ROOT, "tests", "golden", n)))
    sample_scene, sample_mesh = g("sample_scene.json"), g("sample_mesh.json")
    res = json.loads(json.dumps(sample_scene)); res["camera"]["resolution"] = [320, 200]; res["camera"].pop("aspect", None)
    ortho = json.loads(json.dumps(sample_mesh)); ortho["camera"] = dict(ortho["camera"], type="orthographic", aperture=0.3, focusDist=4.5)
    out = [
        dict(name="sample_scene", W=600, H=400, scenes=[sample_scene]),
        dict(name="sample_mesh", W=1280, H=720, scenes=[sample_mesh]),
        dict(name="defaults_and_skip_rules", W=600, H=400, scenes=[tricky_scene()]),
        dict(name="resolution_override", W=600, H=400, scenes=[res]),
        dict(name="orthographic_thin_lens", W=800, H=450, scenes=[ortho]),
        dict(name="no_camera_keeps_camera", W=600, H=400, scenes=[sample_scene, dict(objects=[dict(type="sphere", center=[0, 0, 0], radius=2)])]),
        dict(name="failed_load_keeps_scene", W=600, H=400, scenes=[sample_scene, dict(objects=[None]), dict(objects=[dict(type=5)])]),
        dict(name="camera_defaults", W=640, H=360, scenes=[dict(objects=[dict(type="sphere", center=[0, 0, -3], radius=1)], camera=dict())]),
        dict(name="material_variants", W=600, H=400, scenes=[dict(objects=[
            dict(type="sphere", center=[0, 0, 0], radius=-1.5, material=dict(type="metal", color=[0.2, 0.4, 0.6])),
            dict(type="sphere", center=[1, 0, 0], radius=1, material=dict(type="metal", color=[0.2, 0.4, 0.6], roughness=0.25)),
            dict(type="sphere", center=[2, 0, 0], radius=1, material=dict(type="dielectric", ior=1.33)),
            dict(type="sphere", center=[3, 0, 0], radius=1, material=dict(type="emissive", color=[1, 2, 3], intensity=0)),
            dict(type="sphere", center=[4, 0, 0], radius=1, material=dict(type="emissive", intensity=4.5)),
            dict(type="sphere", center=[5, 0, 0], radius=1, material=dict(color=[0.3, 0.3, 0.3])),
            dict(type="plane", point=[0, 0, 0], normal=[0, 0, 0]),
            dict(type="box", min=[2, 2, 2], max=[1, 1, 1]),
        ], background=dict(type="weird", intensity=0))]),
    ]
    for bgname in ("gradient", "solid", "hdri", "procedural_sky"):
        out.append(dict(name="background_" + bgname, W=600, H=400, scenes=[dict(objects=[], background=dict(type=bgname, color=[0.2, 0.3, 0.4], intensity=2))]))
    return out


def control_calls():
    """A session at the reference's control surface (ray-tracer.js:439-614, 627-680): every camera preset and an unknown one, partial
    updateCamera calls that exercise each `||` fallback (0 and absent mean "keep"), setCameraPosition with and without arguments,
    resizes (setupCamera rebuilds the camera from its own derived vectors), every background kind."""
    calls = [["loadCameraPreset", n] for n in ("close-up", "wide-angle", "top-down", "side-view", "default", "no-such-preset")]
    calls += [["updateCamera", dict(fov=30)], ["updateCamera", dict(position=[1, 2, 3])], ["updateCamera", dict(aperture=0.1, focusDist=3)],
              ["updateCamera", dict(type="orthographic")], ["updateCamera", dict(fov=0, aperture=0, focusDist=0)],
              ["updateCamera", dict(up=[0, 0, 1], lookAt=[1, 1, 1], type="perspective")], ["updateCamera", dict()],
              ["resizeCanvas", 800, 450], ["resizeCanvas", 333, 777], ["setCameraPosition", [4, 1, -2], None, None],
              ["setCameraPosition", None, [0, 0.5, 0], [0, 1, 0]], ["setCameraPosition", None, None, None], ["resizeCanvas", 600, 400],
              ["updateCamera", dict(position=[0, 0, 5], lookAt=[0, 0, 0], up=[0, 1, 0], fov=90, aperture=2, focusDist=5)], ["getCameraPosition"]]
    calls += [["updateBackground", k, i] for k, i in (("solid", 0.5), ("hdri", 2.0), ("procedural_sky", 0.25), ("gradient", 3.0), ("weird", 1.5))]
    calls += [["updateBackground", "solid"], ["loadPreset", "cornell"], ["loadCameraPreset", "close-up"], ["resizeCanvas", 512, 512]]
    return calls


PROBE_PERM = [(i * 167 + 13) % 256 for i in range(256)]          # a fixed permutation of 0..255 (167 is odd)
PROBE_DIRS = [[0.0, 1.0, 0.0], [0.0, -1.0, 0.0], [0.3, 0.6, 0.8], [-0.3, 0.6, -0.5], [1.0, 0.05, -0.2], [-0.7, 0.2, 0.4]]


def run_controls(interp, RayTracer, Vec3, Ray):
    rt = interp.construct(RayTracer, [M.fake_canvas(interp, 600, 400)])
    v = lambda c: J.UNDEF if c is None else interp.construct(Vec3, [float(c[0]), float(c[1]), float(c[2])])
    def view(ret=None):
        st = dump_state(rt)
        world = rt.get("world")
        # world.background evaluated on fixed rays: says which background is installed (solid / hdri are anonymous closures,
        # world.js:42-44, 74-111) and with which intensity; the cloud table (random per World, noise.js:7-17) is an input here as in
        # the render cases
        p = world.get("cloudNoise").get("p")
        for i in range(256):
            J.set_member(p, float(i), float(PROBE_PERM[i])); J.set_member(p, float(256 + i), float(PROBE_PERM[i]))
        probe = []
        for d in PROBE_DIRS:
            probe.append(vec(interp.call(world.get("background"), world, [interp.construct(Ray, [v([0, 0, 0]), v(d)])])))
        kind = {"bound skyGradient": "gradient", "bound proceduralSky": "procedural_sky"}.get(st["background"])
        if kind is None:
            kind = "solid" if probe[0] == probe[1] == probe[4] else "hdri"
        return dict(camera=st["camera"], background=kind, bgProbe=probe, skyIntensity=st["skyIntensity"], width=st["width"], height=st["height"],
                    n_objects=len(st["objects"]), ret=ret)
    steps = [dict(call=["constructor"], state=view())]
    for call in control_calls():
        name, args = call[0], call[1:]
        if name == "setCameraPosition":
            ret = M.method(interp, rt, name, *[v(a) for a in args])
        else:
            ret = M.method(interp, rt, name, *[J.py_to_js(a) if isinstance(a, (dict, list)) else (float(a) if isinstance(a, (int, float)) else a) for a in args])
        if name == "getCameraPosition":
            ret = dict(position=vec(ret.get("position")), lookAt=vec(ret.get("lookAt")), up=vec(ret.get("up")), fov=ret.get("fov"),
                       aperture=ret.get("aperture"), focusDist=ret.get("focusDist"), type=ret.get("type"))
        else:
            ret = ret if isinstance(ret, bool) else None
        steps.append(dict(call=call, state=view(ret)))
    return steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_host_vectors.json"))
    args = ap.parse_args()
    js_dir = os.path.join(args.ref, "js")
    sys.setrecursionlimit(20000)
    out = {"generator": "baseline/make_host_fixtures_minijs.py: RayTracer.loadFromJSON of the unmodified reference executed by baseline/minijs.py", "cases": []}
    for c in cases():
        interp, RayTracer, Vec3 = M.load_reference(js_dir)
        rt = interp.construct(RayTracer, [M.fake_canvas(interp, c["W"], c["H"])])
        steps = []
        for scene in c["scenes"]:
            ok = J.truthy(M.method(interp, rt, "loadFromJSON", J.py_to_js(json.loads(json.dumps(scene)))))
            steps.append(dict(ok=ok, state=dump_state(rt)))
        out["cases"].append(dict(name=c["name"], W=c["W"], H=c["H"], scenes=c["scenes"], steps=steps))
        print(c["name"], [s["ok"] for s in steps], len(steps[-1]["state"]["objects"]), "objects")
    # lights.js illuminate() — never called by the reference's render loop (the direct-lighting EXTENSION uses its formulas)
    interp, RayTracer, Vec3 = M.load_reference(js_dir)
    lights_ex = interp.load_module(os.path.join(js_dir, "lights.js"))
    v = lambda c: interp.construct(Vec3, [float(c[0]), float(c[1]), float(c[2])])
    specs = [dict(kind="point", v=[1.0, 4.0, -2.5], color=[1.0, 0.9, 0.8], intensity=12.0), dict(kind="point", v=[0.0, 0.0, 0.0], color=[0.2, 0.4, 0.6], intensity=1.0),
             dict(kind="directional", v=[0.3, -1.0, 0.2], color=[1.0, 1.0, 0.9], intensity=2.5)]
    pts = [[0.0, 0.0, 0.0], [1.0, 4.0, -2.5], [0.5, -0.5, 3.0], [-7.25, 2.125, 0.001], [100.0, -50.0, 25.0]]
    out["lights"] = []
    for sp in specs:
        cls = lights_ex["PointLight" if sp["kind"] == "point" else "DirectionalLight"]
        light = interp.construct(cls, [v(sp["v"]), v(sp["color"]), float(sp["intensity"])])
        rows = []
        for p in pts:
            r = interp.call(light.get("illuminate"), light, [v(p)])
            rows.append(dict(direction=vec(r.get("direction")), color=vec(r.get("color")), distance=(None if r.get("distance") == float("inf") else r.get("distance"))))
        out["lights"].append(dict(sp, points=pts, illuminate=rows))
    print("lights:", len(out["lights"]))
    interp, RayTracer, Vec3 = M.load_reference(js_dir)
    out["controls"] = run_controls(interp, RayTracer, Vec3, interp.load_module(os.path.join(js_dir, "math.js"))["Ray"])
    out["probe_dirs"], out["probe_perm"] = PROBE_DIRS, PROBE_PERM
    print("controls:", len(out["controls"]), "steps")
    json.dump(out, open(args.out, "w"))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
