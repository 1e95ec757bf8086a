import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blenderraytracer_b200 as brt
from tools import gen_scenes
scene = gen_scenes.terrain(quads=24, extent=200.0)
W, H = 320, 180
rt = brt.RayTracer(W, H, seed=5); assert rt.loadFromJSON(scene)
for sampler in ("reference", "fast"):
    for depth in (1, 2, 3, 6):
        rt.updateRenderSettings(dict(samples=1, maxBounces=depth, toneMapping="linear", gamma=1.0)); rt.sampler = sampler
        out = {}
        for accel in ("brute", "bvh"):
            rt.accel = accel
            rt.render(want_linear=True); out[accel] = rt.linearMean.copy()
        d = np.abs(out["brute"] - out["bvh"]).max(axis=-1)
        ys, xs = np.nonzero(d > 0)
        print(sampler, "depth", depth, "differ", len(ys), "max", d.max())
        for y, x in list(zip(ys, xs))[:4]:
            print(f"   px ({x},{y}) brute {out['brute'][y,x][:3]} bvh {out['bvh'][y,x][:3]}")
