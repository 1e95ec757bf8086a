"""Writes tests/golden/reference_cases_extra.json: further cases for the reference pin (baseline/make_fixtures_minijs.py /
baseline/make_fixtures.mjs), shaped like the BASELINE configs at sizes an interpreter finishes in seconds — the two fixtures at
depth 10, the Cornell-style scene at depth 16 with ACES + denoise, a random-spheres scene with a thin lens, a terrain mesh under
the procedural sky, and the duplicate / coplanar "tie" scene (first object wins, last triangle of a mesh wins).  Run once; the
JSON is committed."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from tools import gen_scenes  # noqa: E402
import independent_port as ip  # noqa: E402


def load(name):
    return json.load(open(os.path.join(HERE, name)))


def tie_scene(seed=7):
    """the scene of tests/test_gpu_parity.py::_tie_scene: exact duplicates, coplanar overlapping triangles, a mesh listed twice"""
    rng = np.random.default_rng(seed)
    objs = []
    lam = lambda c: dict(type="lambertian", color=c)
    for k in range(40):
        c = rng.uniform(-3, 3, 3).round(3).tolist()
        r = round(float(rng.uniform(0.2, 0.7)), 3)
        objs.append(dict(type="sphere", center=c, radius=r, material=lam([0.5, 0.5, 0.5])))
        if k % 4 == 0:
            objs.append(dict(type="sphere", center=c, radius=r, material=dict(type="metal", color=[0.9, 0.9, 0.9], roughness=0.1)))
    for k in range(10):
        mn = rng.uniform(-3, 2, 3).round(2)
        objs.append(dict(type="box", min=mn.tolist(), max=(mn + rng.uniform(0.3, 1.0, 3).round(2)).tolist(), material=lam([0.2, 0.6, 0.3])))
    objs.append(dict(type="box", min=[-1, -1, -1], max=[1, 1, 1], material=lam([0.7, 0.2, 0.2])))
    objs.append(dict(type="box", min=[-1, -1, -1], max=[1, 1, 1], material=lam([0.2, 0.2, 0.7])))
    verts = [[-2, -2, 2.5], [2, -2, 2.5], [2, 2, 2.5], [-2, 2, 2.5], [-1, -1, 2.5], [3, -1, 2.5], [3, 3, 2.5]]
    idx = [0, 1, 2, 0, 2, 3, 0, 1, 2, 0, 2, 3, 4, 5, 6]
    objs.append(dict(type="mesh", vertices=verts, indices=idx, material=lam([0.8, 0.8, 0.1])))
    objs.append(dict(type="triangle", v0=[-2, -2, 2.5], v1=[2, -2, 2.5], v2=[2, 2, 2.5], material=lam([0.1, 0.8, 0.8])))
    objs.append(dict(type="plane", point=[0, -3, 0], normal=[0, 1, 0], material=lam([0.5, 0.5, 0.5])))
    return dict(objects=objs, camera=dict(position=[0.5, 1.0, 9.0], lookAt=[0, 0, 0], fov=50, aspect=1.5, aperture=0.0, focusDist=9.0),
                background=dict(type="gradient"))


def case(name, scene, W, H, spp, depth, seed, perm_seed=None, **kw):
    c = dict(name=name, W=W, H=H, spp=spp, depth=depth, seed=seed, aa=kw.get("aa", "supersampling"), tonemap=kw.get("tonemap", "reinhard"),
             exposure=kw.get("exposure", 1.0), gamma=kw.get("gamma", 2.2), denoise=bool(kw.get("denoise", False)), strength=kw.get("strength", 0.5),
             scene=scene)
    c["perm"] = ip.shuffled_perm(perm_seed) if perm_seed is not None else list(range(256))
    return c


mesh_ortho = load("sample_mesh.json")
mesh_ortho["camera"] = dict(mesh_ortho.get("camera") or {}, type="orthographic")
mesh_ortho["background"] = dict(type="hdri")
out = [
    case("c1_sample_scene_d10", load("sample_scene.json"), 36, 24, 4, 10, 101),
    case("c2_sample_mesh_d10", load("sample_mesh.json"), 32, 18, 4, 10, 102),
    case("c3_random_spheres_thin_lens", gen_scenes.random_spheres(grid=3), 32, 18, 3, 10, 103),
    case("c4_cornell_d16_aces_denoise", gen_scenes.cornell("procedural_sky"), 24, 14, 3, 16, 104, perm_seed=5, tonemap="aces", exposure=1.2, denoise=True, strength=0.8),
    case("c5_terrain_procedural_sky", dict(gen_scenes.terrain(quads=6, extent=200.0), background=dict(type="procedural_sky")), 24, 14, 2, 6, 105, perm_seed=9),
    case("ties_duplicates_coplanar", tie_scene(), 30, 20, 2, 5, 106),
    case("mesh_orthographic_hdri_stochastic_linear", mesh_ortho, 24, 14, 3, 6, 107, aa="stochastic", tonemap="linear", exposure=0.8, gamma=2.0),
]
json.dump(out, open(os.path.join(HERE, "reference_cases_extra.json"), "w"))
print(len(out), "cases:", [c["name"] for c in out], [len(c["scene"]["objects"]) for c in out])
