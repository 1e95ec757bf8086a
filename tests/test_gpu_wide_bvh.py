"""The wide (4- / 8-child) hierarchy collapsed from the binary LBVH (bvh.cu: k_wide_level, brt_device.cuh: trace_wide) must be
INVISIBLE in the results, exactly like the binary one: the reference has no acceleration structure at all (world.js:24-30 and
geometry.js:253-259 are linear loops), so for one Philox stream the brute-force loops, the binary tree and both wide trees give
bit-identical sums — every hit, tie and scattered ray the same.  Also: structure sanity, the counting build, the deep-chain
fallback, shadow rays (direct-lighting extension) through the wide tree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LAM = dict(type="lambertian", color=[0.6, 0.6, 0.6])


@pytest.fixture(scope="module")
def brt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import blenderraytracer_b200 as m
    m.load()
    return m


def _scenes():
    from test_gpu_parity import _scenes as base, _tie_scene          # pytest puts tests/ on sys.path (rootdir conftest)
    d = dict(base())
    d["ties"] = (_tie_scene(), 360, 240)
    return d


def _sums(rt, spp, depth):
    """fp32 sums of `spp` samples per pixel through brt_render_accumulate (library-owned buffer) -> (H, W, 4) float32"""
    import torch
    rt.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    rt._push_params()
    acc = torch.zeros((rt.height, rt.width, 4), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rt.renderAccumulate(acc.data_ptr(), 0, spp)
    rt.synchronize()
    torch.cuda.synchronize()
    return acc.cpu().numpy()


@pytest.mark.parametrize("name", ["ties", "sample_scene", "sample_mesh", "c3_spheres", "c3_ground_sphere", "c4_cornell", "c5_terrain_small"])
def test_wide_hierarchy_is_invisible(brt, name):
    scene, W, H = _scenes()[name]
    rt = brt.RayTracer(W, H, seed=11)
    assert rt.loadFromJSON(scene)
    rt.sampler = "fast"
    out = {}
    for accel, width in (("brute", 0), ("bvh", 2), ("bvh", 4), ("bvh", 8)):
        rt.accel, rt.bvhWidth = accel, width
        out[(accel, width)] = _sums(rt, 4, 8)
        info = rt.sceneInfo()
        if accel == "bvh" and info["n_bvh_nodes"] > 0:
            assert info["bvh_width"] == width, info
            if width > 2:
                assert 1 <= info["bvh_wide_depth"] <= info["bvh_depth"]
    ref = out[("brute", 0)]
    assert ref[..., 3].min() == 4 and ref[..., :3].max() > 0
    for k, v in out.items():
        assert np.array_equal(ref, v), (name, k, int((ref != v).sum()))
    rt.close()


def test_wide_counting_build_and_depth(brt):
    """Counting build over the wide trees: same rays, same primitive hits as over the binary tree; fewer node visits, and the
    visit counter times the width bounds the slab tests."""
    scene, W, H = _scenes()["c5_terrain_small"]
    rt = brt.RayTracer(W, H, seed=3)
    assert rt.loadFromJSON(scene)
    rt.sampler, rt.accel, rt.countTests = "fast", "bvh", True
    st, img = {}, {}
    for width in (2, 4, 8):
        rt.bvhWidth = width
        img[width] = _sums(rt, 2, 6)
        st[width] = rt.stats()
    rt.countTests = False
    for width in (4, 8):
        assert np.array_equal(img[2], img[width])
        assert st[width]["rays"] == st[2]["rays"] > 0
        assert 0 < st[width]["node_visits"] < st[2]["node_visits"]
        assert st[width]["node_visits"] * 2 <= st[width]["tests_aabb"] <= st[width]["node_visits"] * width
        assert st[width]["trav_warp_iters"] > 0 and st[width]["trav_lane_iters"] <= st[width]["trav_alive_lanes"] <= 32 * st[width]["trav_warp_iters"]
    assert st[2]["tests_aabb"] == 2 * st[2]["node_visits"]
    assert st[8]["node_visits"] < st[4]["node_visits"]
    rt.close()


def test_wide_direct_lighting_shadow_rays(brt, ):
    """The direct-lighting extension traces its shadow rays (any-hit) through the same wide tree."""
    from conftest import load_scene
    rt = brt.RayTracer(320, 180, seed=9)
    assert rt.loadFromJSON(load_scene("sample_mesh.json"))
    rt.resizeCanvas(320, 180)
    rt.sampler, rt.directLighting = "fast", True
    out = {}
    for accel, width in (("brute", 0), ("bvh", 2), ("bvh", 4), ("bvh", 8)):
        rt.accel, rt.bvhWidth = accel, width
        out[(accel, width)] = _sums(rt, 4, 5)
    for k, v in out.items():
        assert np.array_equal(out[("brute", 0)], v), k
    rt.close()


def test_wide_falls_back_on_a_degenerate_chain(brt):
    """A chain-shaped LBVH (depth > 32) collapses into a wide tree that is still deeper than the shared-memory stack allows for
    (or not — then it is used); either way the image equals the brute-force image."""
    E = 8.0
    objs = [dict(type="sphere", center=[E, E, E], radius=0.4, material=LAM)]
    for j in range(1, 11):
        for ax in range(3):
            c = [0.0, 0.0, 0.0]
            c[ax] = E * 2.0 ** -j
            objs.append(dict(type="sphere", center=c, radius=0.02 + 0.01 * j, material=LAM))
    objs += [dict(type="sphere", center=[0, 0, 0], radius=0.05 + 0.0004 * k, material=dict(type="metal", color=[0.9, 0.8, 0.7], roughness=0.1)) for k in range(120)]
    scene = dict(objects=objs, camera=dict(position=[3, 2.5, 9], lookAt=[1.5, 1, 0], fov=55, aspect=1.5, aperture=0.0, focusDist=9.0),
                 background=dict(type="gradient"))
    rt = brt.RayTracer(150, 100, seed=2)
    assert rt.loadFromJSON(scene)
    rt.sampler = "fast"
    out = {}
    for accel, width in (("brute", 0), ("bvh", 2), ("bvh", 4), ("bvh", 8)):
        rt.accel, rt.bvhWidth = accel, width
        out[(accel, width)] = _sums(rt, 3, 6)
        if accel == "bvh":
            info = rt.sceneInfo()
            assert info["bvh_depth"] > 32
            assert info["bvh_width"] in (2, width)
            if info["bvh_width"] > 2:
                assert info["bvh_wide_depth"] <= 32
    for k, v in out.items():
        assert np.array_equal(out[("brute", 0)], v), k
    rt.close()


def test_bad_width_is_rejected(brt):
    rt = brt.RayTracer(16, 16)
    rt.bvhWidth = 3
    with pytest.raises(brt.BrtError):
        rt._push_params()
    rt.bvhWidth = 0
    rt._push_params()
    rt.close()
