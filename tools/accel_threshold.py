"""Where does the hierarchy start to pay?  Msamples/s of the megakernel with the linear loops (accel = brute) and with the LBVH
(accel = bvh) on small scenes: the reference's four presets, the Cornell bench scene (C4) and random-sphere fields of 4 .. 64 objects.
    python tools/accel_threshold.py > gpurun_out/accel_threshold.json"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import blenderraytracer_b200 as brt  # noqa: E402
import gen_scenes  # noqa: E402

W, H, SPP = 1280, 720, 32


def time_scene(load, depth):
    out = {}
    for accel in ("brute", "bvh"):
        rt = brt.RayTracer(W, H, seed=3)
        load(rt)
        rt.resizeCanvas(W, H)
        rt.updateRenderSettings(dict(samples=SPP, maxBounces=depth))
        rt.accel = accel
        rt.render(want_float=False)
        best = min((rt.render(want_float=False), rt.stats()["kernel_ms"])[1] for _ in range(4))
        out[accel] = round(W * H * SPP / best / 1e3, 1)
        info = rt.sceneInfo()
        out["bounded"] = info["n_spheres"] + info["n_boxes"] + info["n_triangles"]
        out["planes"] = info["n_planes"]
        rt.close()
    out["bvh_over_brute"] = round(out["bvh"] / out["brute"], 3)
    return out


res = {}
for preset in ("default", "glass", "metals", "cornell"):
    res["preset_" + preset] = time_scene(lambda rt, p=preset: rt.loadPreset(p), 10)
res["c4_cornell_depth16"] = time_scene(lambda rt: rt.loadFromJSON(gen_scenes.cornell("procedural_sky")), 16)
full = gen_scenes.random_spheres()
for n in (4, 6, 8, 12, 16, 24, 32, 64):
    sc = dict(full, objects=full["objects"][:1] + full["objects"][-3:] + full["objects"][1:1 + max(0, n - 3)])
    res[f"spheres_{n}"] = time_scene(lambda rt, s=sc: rt.loadFromJSON(s), 10)
for k, v in res.items():
    print(k, v, file=sys.stderr, flush=True)
print(json.dumps(res, indent=1))
