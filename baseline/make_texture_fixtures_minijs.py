#!/usr/bin/env python
"""make_texture_fixtures_minijs.py — the reference's procedural textures (js/textures.js, js/noise.js) and textured materials
(js/materials.js:99-126), executed from their unmodified source by baseline/minijs.py.

  * values: SolidColor / CheckerTexture / NoiseTexture / MarbleTexture / WoodTexture .value(u, v, p) at a set of points, each
    noise-based texture with an explicit Perlin permutation (the reference shuffles it with Math.random, noise.js:7-17);
  * one seeded render of tests/golden/sample_scene.json whose Lambertian / Metal objects were replaced, in the reference's own
    World, by TexturedLambertian / TexturedMetal over those textures (the reference's JSON has no texture fields).

Writes tests/golden/reference_texture_vectors.json; tests/test_reference_pin.py::test_oracle_textures_match_the_reference compares
the oracle's restatement with it bit for bit.

    python baseline/make_texture_fixtures_minijs.py [--ref /root/reference]
"""
import argparse
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import minijs as J  # noqa: E402
import make_fixtures_minijs as M  # noqa: E402

TEXTURES = [                                                  # (name, kind, ctor args as the oracle takes them, perm seed)
    dict(name="solid", kind="solid", odd=[0.2, 0.5, 0.9], even=[1, 1, 1], scale=1.0, perm_seed=None),
    dict(name="checker10", kind="checker", odd=[0.1, 0.1, 0.1], even=[0.9, 0.8, 0.7], scale=10.0, perm_seed=None),
    dict(name="checker_small", kind="checker", odd=[0.3, 0.6, 0.2], even=[0.05, 0.05, 0.4], scale=2.5, perm_seed=None),
    dict(name="noise1", kind="noise", odd=[1, 1, 1], even=[1, 1, 1], scale=1.0, perm_seed=3),
    dict(name="noise4", kind="noise", odd=[1, 1, 1], even=[1, 1, 1], scale=4.0, perm_seed=4),
    dict(name="marble", kind="marble", odd=[1, 1, 1], even=[1, 1, 1], scale=3.0, perm_seed=5),
    dict(name="wood", kind="wood", odd=[1, 1, 1], even=[1, 1, 1], scale=1.5, perm_seed=6),
]


def perm_of(seed):
    p = list(range(256))
    random.Random(seed).shuffle(p)
    return p


def points():
    r = random.Random(77)
    pts = [[0.0, 0.0, 0.0], [1.0, 2.0, 3.0], [-0.5, 0.25, -7.75], [255.5, -256.25, 1e-3], [-1e-9, 1e3, -1e3]]
    pts += [[round(r.uniform(-6, 6), 6), round(r.uniform(-6, 6), 6), round(r.uniform(-6, 6), 6)] for _ in range(43)]
    return pts


def make_texture(interp, tex_ex, Vec3, t):
    v = lambda c: interp.construct(Vec3, [float(c[0]), float(c[1]), float(c[2])])
    if t["kind"] == "solid": tex = interp.construct(tex_ex["SolidColor"], [v(t["odd"])])
    elif t["kind"] == "checker": tex = interp.construct(tex_ex["CheckerTexture"], [v(t["odd"]), v(t["even"]), float(t["scale"])])
    else: tex = interp.construct(tex_ex[{"noise": "NoiseTexture", "marble": "MarbleTexture", "wood": "WoodTexture"}[t["kind"]]], [float(t["scale"])])
    if t["perm_seed"] is not None:
        p = tex.get("noise").get("p"); perm = perm_of(t["perm_seed"])
        for i in range(256):
            J.set_member(p, float(i), float(perm[i])); J.set_member(p, float(256 + i), float(perm[i]))
    return tex


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_texture_vectors.json"))
    args = ap.parse_args()
    js_dir = os.path.join(args.ref, "js")
    sys.setrecursionlimit(20000)
    interp, RayTracer, Vec3 = M.load_reference(js_dir)
    tex_ex = interp.load_module(os.path.join(js_dir, "textures.js"))
    mat_ex = interp.load_module(os.path.join(js_dir, "materials.js"))
    pts = points()
    out = {"generator": "baseline/make_texture_fixtures_minijs.py: js/textures.js + js/noise.js + js/materials.js of the unmodified reference executed by baseline/minijs.py",
           "points": pts, "textures": []}
    for t in TEXTURES:
        tex = make_texture(interp, tex_ex, Vec3, t)
        vals = []
        for p in pts:
            r = interp.call(tex.get("value"), tex, [0.0, 0.0, interp.construct(Vec3, [float(p[0]), float(p[1]), float(p[2])])])
            vals.append([r.get("x"), r.get("y"), r.get("z")])
        out["textures"].append(dict(t, perm=(perm_of(t["perm_seed"]) if t["perm_seed"] is not None else None), values=vals))
        print(t["name"], vals[1])
    # a render through TexturedLambertian / TexturedMetal
    scene = json.load(open(os.path.join(ROOT, "tests", "golden", "sample_scene.json")))
    case = dict(name="textured_sample_scene", W=20, H=14, spp=3, depth=6, seed=210, aa="supersampling", tonemap="reinhard", exposure=1.0, gamma=2.2,
                denoise=False, strength=0.5, scene=scene, perm=list(range(256)))
    assign = {0: "checker10", 1: "marble", 2: "wood", 3: "noise4"}            # object index -> texture
    by_name = {t["name"]: t for t in TEXTURES}
    def retexture(rt):
        objs = rt.get("world").get("objects").items
        used = {}
        for i, name in assign.items():
            if i >= len(objs): continue
            m = objs[i].get("material"); cls = m.proto.get("constructor").name
            tex = make_texture(interp, tex_ex, Vec3, by_name[name])
            if cls == "Lambertian": objs[i].set("material", interp.construct(mat_ex["TexturedLambertian"], [tex])); used[i] = name
            elif cls == "Metal": objs[i].set("material", interp.construct(mat_ex["TexturedMetal"], [tex, m.get("roughness")])); used[i] = name
        return used
    orig_build = M.build_case
    used = {}
    def build_and_retexture(interp_, RayTracer_, Vec3_, c):
        rt = orig_build(interp_, RayTracer_, Vec3_, c); used.update(retexture(rt)); return rt
    M.build_case = build_and_retexture
    try:
        r = M.render_seeded(interp, RayTracer, Vec3, case)
    finally:
        M.build_case = orig_build
    out["render"] = dict(case=case, textured_objects={str(k): v for k, v in used.items()}, **r)
    print("render:", used, "mean linear", sum(r["linear"]) / len(r["linear"]))
    json.dump(out, open(args.out, "w"))
    print("wrote", args.out)


if __name__ == "__main__":
    main()
