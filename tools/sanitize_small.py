"""Small invocations of every kernel variant (for compute-sanitizer --tool memcheck / racecheck; seconds under the tool)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blenderraytracer_b200 as brt
from tools import gen_scenes

scenes = {
    "mesh": json.load(open(os.path.join(ROOT, "tests/golden/sample_mesh.json"))),
    "c3": gen_scenes.random_spheres(grid=3),
    "c4": gen_scenes.cornell("procedural_sky"),
    "c5": gen_scenes.terrain(quads=12),
    "empty": dict(objects=[], camera=dict(position=[0, 0, 3], lookAt=[0, 0, 0], fov=40, aspect=1.5)),
}
for name, sc in scenes.items():
    rt = brt.RayTracer(45, 27, seed=2)          # ragged size: partial tiles
    assert rt.loadFromJSON(sc)
    rt.updateRenderSettings(dict(samples=3, maxBounces=4, denoising=(name == "mesh")))
    for integ in ("megakernel", "wavefront"):
        for accel in ("brute", "bvh"):
            for sampler in ("fast", "reference"):
                rt.integrator, rt.accel, rt.sampler = integ, accel, sampler
                rt.directLighting = (name == "mesh" and sampler == "reference")
                img = rt.render(want_linear=True)
                assert img.shape == (27, 45, 4)
    rt.integrator, rt.accel, rt.sampler, rt.directLighting = "megakernel", "bvh", "fast", False
    for width in (4, 8, 2, 0):                   # the wide collapses of the hierarchy and back
        rt.bvhWidth = width
        rt.render()
    rt.countTests = True
    rt.render()
    rt.countTests = False
    rt.primaryAOV(32); rt.primaryAOV(64)
    rt.evalBackground(np.random.default_rng(0).normal(size=(100, 3)))
    print(name, "ok", flush=True)
# a mesh large enough for the GPU LBVH + the host SAH candidate + the wide level builder to run with many levels
rt = brt.RayTracer(64, 36, seed=3)
assert rt.loadFromJSON(gen_scenes.terrain(quads=40))
rt.updateRenderSettings(dict(samples=2, maxBounces=3))
for width in (2, 4, 8):
    rt.bvhWidth = width
    rt.render()
print("terrain40 ok", flush=True)
print("SANITIZE_SMALL_DONE")
