"""Writes tests/golden/reference_cases.json: the 13 cases of independent_port.cases() in the form baseline/make_fixtures.mjs
feeds to the UNMODIFIED reference under Node (explicit Philox seeds; the Perlin table of World.cloudNoise spelled out, because
the reference shuffles it with Math.random, noise.js:7-14).  Run once; the JSON is committed."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import independent_port as ip  # noqa: E402

out = []
for name, scene, W, H, kw in ip.cases():
    c = dict(name=name, W=W, H=H, spp=kw["spp"], depth=kw["depth"], seed=kw["seed"], aa=kw.get("aa", "supersampling"),
             tonemap=kw.get("tonemap", "reinhard"), exposure=kw.get("exposure", 1.0), gamma=kw.get("gamma", 2.2),
             denoise=bool(kw.get("denoise", False)), strength=kw.get("strength", 0.5))
    if name.startswith("preset_"):
        c["preset"] = name[len("preset_"):]                       # RayTracer.loadPreset (ray-tracer.js:282-299), not a JSON scene
    else:
        c["scene"] = scene
    c["perm"] = ip.shuffled_perm(kw["perm_seed"]) if kw.get("perm_seed") is not None else list(range(256))
    out.append(c)
json.dump(out, open(os.path.join(HERE, "reference_cases.json"), "w"))
print(len(out), "cases")
