"""C3 with the plane ground vs the r = 1000 "ground sphere" idiom: Msamples/s and what the BVH builder made of each.
    python tools/ground_variant.py [spp]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import blenderraytracer_b200 as brt  # noqa: E402
import gen_scenes  # noqa: E402

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
out = {}
for ground in ("plane", "sphere"):
    rt = brt.RayTracer(1920, 1080, seed=5)
    assert rt.loadFromJSON(gen_scenes.random_spheres(ground=ground))
    rt.updateRenderSettings(dict(samples=spp, maxBounces=10))
    rt.render(want_float=False)
    best = 1e9
    for _ in range(3):
        rt.render(want_float=False)
        best = min(best, rt.stats()["kernel_ms"])
    rt.countTests = True
    rt.updateRenderSettings(dict(samples=4, maxBounces=10))
    rt.render(want_float=False)
    st, info = rt.stats(), rt.sceneInfo()
    out[ground] = dict(msamples_per_s=1920 * 1080 * spp / best / 1e3, kernel_ms=best, info={k: info[k] for k in info if "bvh" in k or k.startswith("n_")},
                       node_visits_per_ray=st["node_visits"] / max(1, st["rays"]) if "rays" in st else None)
    print(ground, json.dumps(out[ground]), flush=True)
    rt.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ground_variant.json", "w"), indent=1)
