"""Every kernel variant (-m gpu): integrator {megakernel, wavefront} x accel {brute, bvh} x sampler {fast, reference}
(+ direct lighting, counting build, denoise) on ragged image sizes.  Within one sampler all variants trace the same paths:
brute == bvh bit for bit, wavefront == megakernel up to fp32 summation order."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scenes():
    from tools import gen_scenes
    return {
        "mesh": json.load(open(os.path.join(ROOT, "tests/golden/sample_mesh.json"))),
        "c3": gen_scenes.random_spheres(grid=3),
        "c4": gen_scenes.cornell("procedural_sky"),
        "c5": gen_scenes.terrain(quads=12),
        "empty": dict(objects=[], camera=dict(position=[0, 0, 3], lookAt=[0, 0, 0], fov=40, aspect=1.5)),
    }


@pytest.mark.parametrize("name", ["mesh", "c3", "c4", "c5", "empty"])
@pytest.mark.parametrize("sampler", ["fast", "reference"])
def test_all_variants_agree(name, sampler):
    import blenderraytracer_b200 as brt
    rt = brt.RayTracer(45, 27, seed=2)                    # partial tiles on both axes
    assert rt.loadFromJSON(_scenes()[name])
    rt.updateRenderSettings(dict(samples=5, maxBounces=4, toneMapping="linear", gamma=1.0))
    rt.sampler = sampler
    rt.directLighting = name == "mesh"
    out = {}
    for integ in ("megakernel", "wavefront"):
        for accel in ("brute", "bvh"):
            rt.integrator, rt.accel = integ, accel
            rt.render(want_linear=True)
            out[integ, accel] = rt.linearMean.copy()
    for integ in ("megakernel", "wavefront"):
        assert np.array_equal(out[integ, "brute"], out[integ, "bvh"]), integ
    np.testing.assert_allclose(out["wavefront", "bvh"], out["megakernel", "bvh"], rtol=2e-6, atol=1e-7)
    for k in (1, 3, 4):                                   # paths in flight per lane only reorders the sums
        rt.integrator, rt.accel, rt.pathsInFlight = "wavefront", "bvh", k
        rt.render(want_linear=True)
        np.testing.assert_allclose(rt.linearMean, out["megakernel", "bvh"], rtol=2e-6, atol=1e-7)


def test_counting_build_counts_the_same_traversal():
    """count_tests = 1 must not change the image, and its counters obey simple identities."""
    import blenderraytracer_b200 as brt
    from tools import gen_scenes
    rt = brt.RayTracer(64, 40, seed=4)
    assert rt.loadFromJSON(gen_scenes.random_spheres(grid=4))
    rt.updateRenderSettings(dict(samples=4, maxBounces=6))
    a = rt.render()
    rt.countTests = True
    b = rt.render()
    st = rt.stats()
    assert np.array_equal(a, b)
    assert st["samples"] == 64 * 40 * 4
    assert st["samples"] <= st["rays"] <= st["samples"] * 6
    assert st["tests_plane"] <= st["rays"] and st["tests_aabb"] % 2 == 0 and st["tests_sphere"] > 0
    rt.accel = "brute"
    rt.render()
    sb = rt.stats()
    assert sb["rays"] == st["rays"] and sb["tests_aabb"] == 0
    assert sb["tests_sphere"] == sb["rays"] * rt.sceneInfo()["n_spheres"]          # the reference's O(N) loop
