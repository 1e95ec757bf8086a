"""Edge cases of the hot path on the GPU (-m gpu), each checked against the float64 oracle: empty and degenerate scenes,
scenes with no / one / two bounded primitives (no BVH, trivial BVH), ragged image sizes, axis-parallel rays, cameras inside
objects, negative radii, the reference's presets built object by object, ragged meshes."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def brt():
    import blenderraytracer_b200 as b
    return b


CAM = dict(position=[0, 1, 4], lookAt=[0, 0, 0], fov=45, aspect=1.5, aperture=0.0, focusDist=4.0)
LAM = dict(type="lambertian", color=[0.6, 0.5, 0.4])


def _check(brt, scene, W=150, H=100, spp=4, depth=5, max_id_mismatch=0, accels=("brute", "bvh")):
    """AOV f64 bit-exact, AOV f32 IDs / t / normal, and a same-stream render within 2 LSB, for every accel mode."""
    from oracle.oracle import OracleRayTracer
    rt = brt.RayTracer(W, H, seed=31)
    orc = OracleRayTracer(W, H, seed=31, threads=4)
    assert rt.loadFromJSON(scene) and orc.loadFromJSON(scene)
    for r in (rt, orc):
        r.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    o = orc.primary_aov()
    a64 = rt.primaryAOV(64)
    for key in ("obj_id", "tri_id", "t", "normal", "front_face"):
        assert np.array_equal(a64[key], o[key]), key
    ref = orc.render()
    rt.sampler = "reference"
    for accel in accels:
        rt.accel = accel
        a = rt.primaryAOV(32)
        mism = (a["obj_id"] != o["obj_id"]) | (a["tri_id"] != o["tri_id"])
        assert mism.sum() <= max_id_mismatch, (accel, int(mism.sum()))
        ok = ~mism & (o["obj_id"] >= 0)
        if ok.any():
            assert np.array_equal(a["t"][ok], o["t"][ok].astype(np.float32)), accel
            assert np.abs(a["normal"][ok] - o["normal"][ok]).max() <= 1e-5, accel
        img = rt.render()
        d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=-1)
        assert (d <= 2).mean() >= 0.97, (accel, (d <= 2).mean())
    return rt, orc


def test_empty_world_is_background_only(brt):
    for bg in ("gradient", "hdri", "procedural_sky", "solid"):
        rt, orc = _check(brt, dict(objects=[], camera=CAM, background=dict(type=bg, intensity=0.8)))
        assert np.all(rt.primaryAOV(32)["obj_id"] == -1)
        assert rt.sceneInfo()["n_bvh_nodes"] == 0


def test_only_unbounded_planes(brt):
    scene = dict(objects=[dict(type="plane", point=[0, -1, 0], normal=[0, 1, 0], material=LAM),
                          dict(type="plane", point=[0, 0, -6], normal=[0, 0, 1], material=dict(type="metal", color=[0.9, 0.9, 0.9], roughness=0.0))],
                 camera=CAM, background=dict(type="gradient"))
    rt, _ = _check(brt, scene)
    assert rt.sceneInfo()["n_bvh_nodes"] == 0


def test_one_and_two_bounded_primitives(brt):
    one = dict(objects=[dict(type="sphere", center=[0, 0, 0], radius=1.0, material=LAM)], camera=CAM, background=dict(type="gradient"))
    rt, _ = _check(brt, one)
    assert rt.sceneInfo()["n_bvh_nodes"] == 0                       # a single primitive needs no hierarchy
    two = dict(objects=one["objects"] + [dict(type="box", min=[1.2, -1, -1], max=[2.2, 0.5, 0.3], material=LAM)], camera=CAM)
    rt, _ = _check(brt, two)
    assert rt.sceneInfo()["n_bvh_nodes"] == 1


@pytest.mark.parametrize("W,H", [(1, 1), (37, 19), (16, 8), (17, 9), (255, 3)])
def test_ragged_image_sizes(brt, sample_scene, W, H):
    _check(brt, sample_scene, W, H)


def test_axis_parallel_rays_and_box_faces(brt):
    """Odd image size + symmetric camera: the centre column / row have direction components that are exactly 0
    (division by zero inside the slab test: +-inf / NaN semantics of geometry.js:86-103)."""
    scene = dict(objects=[dict(type="box", min=[-1, -1, -1], max=[1, 1, 1], material=LAM),
                          dict(type="box", min=[-3, -0.5, -0.5], max=[-2, 0.5, 0.5], material=LAM),
                          dict(type="plane", point=[0, -1, 0], normal=[0, 1, 0], material=LAM)],
                 camera=dict(position=[0, 0, 5], lookAt=[0, 0, 0], fov=50, aspect=1.0, aperture=0.0, focusDist=5.0),
                 background=dict(type="gradient"))
    _check(brt, scene, 101, 101)


def test_axis_parallel_rays_off_origin_bvh_equals_brute(brt):
    """Camera (0.3, 1, 5) -> (0.3, 1, 0): the centre column has D.x == 0 and the centre row D.y == 0 with an origin that is
    NOT on a slab plane, so `bound*inv - O*inv` would be inf - inf = NaN for every BVH box straddling x = 0.3 / y = 1 and
    the hierarchy would lose primitives the linear loop hits.  The slab reciprocal clamps |d| >= 2^-80 instead."""
    objs = []
    for k in range(6):                                                      # 12 bounded primitives straddling x = 0.3 and / or y = 1
        objs.append(dict(type="sphere", center=[0.3 + 0.05 * (k - 2.5), 1.0 + 0.45 * (k - 2.5), -1.0 - 0.3 * k], radius=0.4, material=LAM))
        objs.append(dict(type="box", min=[-2.5 + 0.9 * k, 0.8, -3.0], max=[-1.9 + 0.9 * k, 1.2, -2.6], material=LAM))
    objs.append(dict(type="plane", point=[0, -1, 0], normal=[0, 1, 0], material=LAM))
    scene = dict(objects=objs, camera=dict(position=[0.3, 1, 5], lookAt=[0.3, 1, 0], fov=40, aspect=1.0, aperture=0.0, focusDist=5.0),
                 background=dict(type="gradient"))
    W = H = 101
    rt = brt.RayTracer(W, H, seed=5)
    assert rt.loadFromJSON(scene)
    rt.updateRenderSettings(dict(samples=1, maxBounces=4, antiAliasing="none"))
    out = {}
    for accel in ("brute", "bvh"):
        rt.accel = accel
        out[accel] = (rt.primaryAOV(32), rt.render(want_linear=True).copy(), rt.linearMean.copy())
    assert rt.sceneInfo()["n_bvh_nodes"] == 11
    for key in ("obj_id", "tri_id", "t", "normal", "front_face"):
        assert np.array_equal(out["brute"][0][key], out["bvh"][0][key]), key
    assert np.array_equal(out["brute"][1], out["bvh"][1])
    assert np.array_equal(out["brute"][2], out["bvh"][2])
    col, row = out["bvh"][0]["obj_id"][:, W // 2], out["bvh"][0]["obj_id"][H // 2, :]
    assert (col >= 0).sum() > 40 and len(set(row.tolist()) - {-1, 12}) >= 3      # the centre column / row do see the primitives
    _check(brt, scene, W, H)


def test_camera_inside_objects(brt):
    inside_sphere = dict(objects=[dict(type="sphere", center=[0, 0, 0], radius=10.0, material=dict(type="emissive", color=[1, 0.9, 0.8], intensity=0.5)),
                                  dict(type="sphere", center=[0, 0, -2], radius=0.7, material=dict(type="dielectric", ior=1.5))],
                         camera=dict(position=[0, 0, 2], lookAt=[0, 0, -2], fov=50, aspect=1.5, aperture=0.0, focusDist=4.0))
    rt, orc = _check(brt, inside_sphere, depth=8)
    assert rt.primaryAOV(64)["front_face"].min() == 0               # the big sphere is seen from inside
    inside_box = dict(objects=[dict(type="box", min=[-3, -2, -6], max=[3, 2, 3], material=LAM),
                               dict(type="sphere", center=[0, 1.5, -2], radius=0.3, material=dict(type="emissive", color=[1, 1, 1], intensity=20))],
                      camera=dict(position=[0, 0, 2], lookAt=[0, 0, -2], fov=60, aspect=1.5, aperture=0.0, focusDist=4.0))
    _check(brt, inside_box, depth=6)


def test_negative_radius_hollow_glass(brt):
    """Sphere(center, -0.45, glass) inside Sphere(center, 0.5, glass): the hollow-glass idiom (ray-tracer.js:347)."""
    g = dict(type="dielectric", ior=1.5)
    scene = dict(objects=[dict(type="sphere", center=[0, 0, -1], radius=0.5, material=g), dict(type="sphere", center=[0, 0, -1], radius=-0.45, material=g),
                          dict(type="plane", point=[0, -0.5, 0], normal=[0, 1, 0], material=LAM)],
                 camera=dict(position=[0, 0.3, 1.2], lookAt=[0, 0, -1], fov=40, aspect=1.5, aperture=0.0, focusDist=2.2), background=dict(type="gradient"))
    _check(brt, scene, depth=10)


def test_degenerate_and_ragged_meshes(brt):
    verts = [[-1, 0, 0], [1, 0, 0], [0, 1.5, 0], [0, 0.5, 0], [2, 2, -1]]
    idx = [0, 1, 2,   0, 0, 0,   0, 1, 1,   3, 3, 4,   0, 1]            # one real triangle, three zero-area ones, an incomplete tail
    scene = dict(objects=[dict(type="mesh", vertices=verts, indices=idx, material=LAM),
                          dict(type="mesh", vertices=verts, indices=[], material=LAM),          # a mesh with no triangles is still object 1
                          dict(type="triangle", v0=[2, 0, 0], v1=[2, 0, 0], v2=[2, 0, 0], material=LAM),
                          dict(type="sphere", center=[-2, 0.5, 0], radius=0.5, material=LAM)],
                 camera=CAM, background=dict(type="gradient"))
    rt, orc = _check(brt, scene)
    a = rt.primaryAOV(64)
    assert set(np.unique(a["obj_id"])) == {-1, 0, 3} and set(np.unique(a["tri_id"][a["obj_id"] == 0])) == {0}
    assert rt.sceneInfo()["n_triangles"] == 5


@pytest.mark.parametrize("preset", ["default", "glass", "metal", "cornell"])
def test_presets_built_object_by_object(brt, preset):
    """loadPreset (ray-tracer.js:282-299, 42-77, 336-435): scenes built through World.add / brt_scene_set_flat rather than JSON."""
    from oracle.oracle import OracleRayTracer
    W, H = 180, 120
    rt = brt.RayTracer(W, H, seed=8)
    orc = OracleRayTracer(W, H, seed=8, threads=4)
    rt.loadPreset(preset)
    orc.loadPreset(preset)
    o, a64 = orc.primary_aov(), rt.primaryAOV(64)
    for key in ("obj_id", "t", "normal", "front_face"):
        assert np.array_equal(a64[key], o[key]), key
    for r in (rt, orc):
        r.updateRenderSettings(dict(samples=4, maxBounces=6))
    rt.sampler = "reference"
    img, ref = rt.render(), orc.render()
    d = np.abs(img[..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=-1)
    assert (d <= 2).mean() >= 0.96, (d <= 2).mean()


def test_camera_presets_and_ui_setters(brt, sample_scene):
    """updateCamera / loadCameraPreset / resizeCanvas (ray-tracer.js:475-510, 598-680) keep the two mirrors in step."""
    from oracle.oracle import OracleRayTracer
    rt, orc = brt.RayTracer(120, 80), OracleRayTracer(120, 80)
    for r in (rt, orc):
        assert r.loadFromJSON(sample_scene)
        assert r.loadCameraPreset("close-up") and not r.loadCameraPreset("nope")
        r.updateCamera(dict(fov=55, aperture=0))          # aperture 0 is falsy: keeps the preset's 0.02 (`||`)
        r.resizeCanvas(200, 100)
    cg, co = rt.camera, orc.scene.camera()
    for key in ("origin", "lowerLeftCorner", "horizontal", "vertical", "u", "v", "w"):
        np.testing.assert_allclose(cg[key], co[key], rtol=0, atol=1e-14, err_msg=key)
    assert cg["lensRadius"] == co["lensRadius"] == 0.01
    assert np.array_equal(rt.primaryAOV(64)["t"], orc.primary_aov()["t"])


def test_depth_of_field_statistics(brt, sample_scene):
    """Thin lens (aperture 0.05 in the fixture, camera.js:44-48): blurred edges agree statistically with the oracle."""
    from oracle.oracle import OracleRayTracer
    sc = json.loads(json.dumps(sample_scene))
    sc["camera"]["aperture"] = 0.4
    W, H, spp = 150, 100, 64
    rt = brt.RayTracer(W, H, seed=3); orc = OracleRayTracer(W, H, seed=4, threads=8); orc2 = OracleRayTracer(W, H, seed=5, threads=8)
    for r in (rt, orc, orc2):
        assert r.loadFromJSON(sc)
        r.updateRenderSettings(dict(samples=spp, maxBounces=4))
    rt.render(); orc.render(); orc2.render()
    g, a, b = (x.floatData[..., :3].astype(np.float64) for x in (rt, orc, orc2))
    assert np.sqrt(np.mean((g - a) ** 2)) <= 1.5 * np.sqrt(np.mean((b - a) ** 2)) + 1e-4
    assert np.abs((g - a).mean(axis=(0, 1))).max() <= 3e-3


# ---------------------------------------------------------------------------------------------- surface textures (§8f)
def _textured_pair(brt, W, H, seed=5):
    """The same textured scene built object by object in both engines (textures.js + materials.js:99-126; the reference has
    the classes but never instantiates them, so they exist only on the flat-scene path)."""
    from oracle.oracle import OracleRayTracer, make_perm
    pa, pb, pc = make_perm(11), make_perm(12), make_perm(13)
    rt = brt.RayTracer(W, H, seed=seed)
    w = brt.World()
    w.add(brt.Plane((0, -0.5, 0), (0, 1, 0), brt.TexturedLambertian(brt.CheckerTexture((0.1, 0.1, 0.1), (0.9, 0.9, 0.9), 3.0))))
    w.add(brt.Sphere((-1.1, 0, -1), 0.5, brt.TexturedLambertian(brt.MarbleTexture(4.0, pa))))
    w.add(brt.Sphere((0, 0, -1), 0.5, brt.TexturedMetal(brt.WoodTexture(2.0, pb), 0.2)))
    w.add(brt.Sphere((1.1, 0, -1), 0.5, brt.TexturedLambertian(brt.NoiseTexture(5.0, pc))))
    w.add(brt.Box((-0.3, -0.5, 0.2), (0.3, 0.1, 0.8), brt.TexturedLambertian(brt.SolidColor((0.2, 0.5, 0.8)))))
    w.add(brt.Sphere((0, 2.5, -1), 0.6, brt.Emissive((1, 1, 1), 4)))
    rt._set_world(w)
    rt._set_camera_raw((0, 1.0, 3.0), (0, 0, -1), (0, 1, 0), 40, W / H, 0.0, 4.0)
    orc = OracleRayTracer(W, H, seed=seed, threads=4)
    orc.scene = type(orc.scene)()
    sc = orc.scene
    sc.add_plane((0, -0.5, 0), (0, 1, 0), ("lambertian", [1, 1, 1], 0.0)); sc.set_object_texture(0, "checker", (0.1, 0.1, 0.1), (0.9, 0.9, 0.9), 3.0)
    sc.add_sphere((-1.1, 0, -1), 0.5, ("lambertian", [1, 1, 1], 0.0)); sc.set_object_texture(1, "marble", scale=4.0, perm256=pa)
    sc.add_sphere((0, 0, -1), 0.5, ("metal", [1, 1, 1], 0.2)); sc.set_object_texture(2, "wood", scale=2.0, perm256=pb)
    sc.add_sphere((1.1, 0, -1), 0.5, ("lambertian", [1, 1, 1], 0.0)); sc.set_object_texture(3, "noise", scale=5.0, perm256=pc)
    sc.add_box((-0.3, -0.5, 0.2), (0.3, 0.1, 0.8), ("lambertian", [1, 1, 1], 0.0)); sc.set_object_texture(4, "solid", (0.2, 0.5, 0.8))
    sc.add_sphere((0, 2.5, -1), 0.6, ("emissive", [1, 1, 1], 4.0))
    sc.set_camera((0, 1.0, 3.0), (0, 0, -1), (0, 1, 0), 40, W / H, 0.0, 4.0)
    return rt, orc


def test_texture_values_match_oracle(brt):
    rt, orc = _textured_pair(brt, 32, 32)
    rng = np.random.default_rng(4)
    pts = rng.uniform(-3, 3, size=(4000, 3))
    for tex_index, obj in enumerate(range(5)):
        got = rt.evalTexture(tex_index, pts)
        want = np.array([orc.scene.texture_value(obj, p) for p in pts])
        close = np.abs(got - want).max(axis=-1) <= 2e-4
        # checker is a sign test of a product of sines: points within fp32 rounding of a zero crossing may flip
        assert close.mean() >= (0.995 if obj == 0 else 0.9995), (obj, close.mean())
    assert np.array_equal(rt.evalTexture(4, pts[:3]), np.tile(np.float32([0.2, 0.5, 0.8]), (3, 1)))


def test_textured_materials_render_like_the_oracle(brt):
    W, H = 180, 120
    rt, orc = _textured_pair(brt, W, H)
    for r in (rt, orc):
        r.updateRenderSettings(dict(samples=6, maxBounces=6))
    a64, o = rt.primaryAOV(64), orc.primary_aov()
    assert np.array_equal(a64["obj_id"], o["obj_id"]) and np.array_equal(a64["t"], o["t"])
    rt.sampler = "reference"
    out = {}
    for accel in ("brute", "bvh"):
        rt.accel = accel
        out[accel] = rt.render()
    assert np.array_equal(out["brute"], out["bvh"])
    ref = orc.render()
    d = np.abs(out["bvh"][..., :3].astype(int) - ref[..., :3].astype(int)).max(axis=-1)
    assert (d <= 2).mean() >= 0.95, (d <= 2).mean()
    # the textures are really there: the checkered floor is not a flat colour
    floor = out["bvh"][H - 10:, :, 0].astype(int)
    assert floor.max() - floor.min() > 60


def test_pathologically_deep_lbvh_uses_the_hybrid_stack(brt):
    """Centroids at E*2^-j along each axis give Morton codes with a single set bit each: the LBVH degenerates into a chain
    deeper than SMEM_ONLY_MAX_DEPTH (32), which selects the kernel variant whose stack spills from shared to local memory.
    Results must not care."""
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
    rt, _ = _check(brt, scene, 150, 100, spp=4, depth=6)
    assert rt.sceneInfo()["bvh_depth"] > 32, rt.sceneInfo()["bvh_depth"]


def test_every_kernel_variant_survives_the_random_scenes(brt):
    """The 36 random degenerate scenes of tests/golden/reference_fuzz_vectors.json (zero radii, zero normals, fov 0, coordinates of
    1e6, filtered-out meshes, unknown types ...) through the FAST sampler and every traversal variant: no error, and the linear
    loops, the binary hierarchy and its 4- / 8-wide collapses give bit-identical images (the hierarchy is invisible), also with
    the wavefront integrator up to fp32 summation order."""
    import json, os
    doc = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fuzz_vectors.json")))
    n_bvh = 0
    for e in doc["cases"]:
        c = e["case"]
        W, H = c["W"], c["H"]
        rt = brt.RayTracer(W, H, seed=c["seed"])
        assert rt.loadFromJSON(c["scene"]), c["name"]
        rt.setCloudPermutation(np.asarray(c["perm"], np.uint8))
        rt.updateRenderSettings(dict(samples=4, maxBounces=c["depth"], antiAliasing=c["aa"], toneMapping=c["tonemap"], exposure=c["exposure"], gamma=c["gamma"]))
        rt.sampler, rt.accel = "fast", "brute"
        base = rt.render(want_linear=True).copy(); lin = rt.linearMean.copy()
        info = rt.sceneInfo()
        if info["n_spheres"] + info["n_boxes"] + info["n_triangles"] >= 2:
            n_bvh += 1
            rt.accel = "bvh"
            for width in (2, 4, 8):
                rt.bvhWidth = width
                img = rt.render(want_linear=True)
                assert np.array_equal(img, base) and np.array_equal(rt.linearMean, lin, equal_nan=True), (c["name"], width)
            rt.bvhWidth = 0
            rt.integrator = "wavefront"
            img = rt.render()
            assert (np.abs(img.astype(int) - base.astype(int)) <= 1).mean() >= 0.99, c["name"]
        rt.close()
    assert n_bvh >= 10
