"""G1: the float64 oracle against the hand-derived known-answer vectors (tests/golden/kat.json).
The reference ships no tests of its own, so these vectors — derived independently by
tests/golden/derive_kat.py following the cited reference formulas — are what pins the oracle."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle.oracle import OracleRayTracer, OracleScene, lib, make_perm


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def test_kat_a_camera(kat, sample_scene, sample_mesh):
    for name, scene, W, H in (("sample_scene", sample_scene, 600, 400), ("sample_mesh", sample_mesh, 1280, 720)):
        rt = OracleRayTracer(W, H)
        assert rt.loadFromJSON(scene)
        cam = rt.scene.camera()
        for key in ("w", "u", "v", "horizontal", "vertical", "lowerLeftCorner", "origin"):
            np.testing.assert_allclose(cam[key], kat["camera"][name][key], rtol=0, atol=1e-15, err_msg=f"{name}.{key}")
        assert cam["lensRadius"] == kat["camera"][name]["lensRadius"]


def test_kat_a_camera_ui_path_aspect(kat, sample_mesh):
    """updateCamera (ray-tracer.js:505) rebuilds the camera with width/height — the 16:9 variant of KAT-A."""
    rt = OracleRayTracer(1280, 720)
    assert rt.loadFromJSON(sample_mesh)
    rt.updateCamera({})
    cam = rt.scene.camera()
    np.testing.assert_allclose(cam["horizontal"], kat["camera"]["sample_mesh_16_9"]["horizontal"], rtol=1e-14)
    np.testing.assert_allclose(cam["lowerLeftCorner"], kat["camera"]["sample_mesh_16_9"]["lowerLeftCorner"], rtol=1e-14)


@pytest.mark.parametrize("name,scene_fx,W,H", [("sample_scene_600x400", "sample_scene", 600, 400),
                                                ("sample_mesh_1280x720", "sample_mesh", 1280, 720)])
def test_kat_b_primary_hits(kat, request, name, scene_fx, W, H):
    rt = OracleRayTracer(W, H)
    assert rt.loadFromJSON(request.getfixturevalue(scene_fx))
    a = rt.primary_aov()
    for row in kat["primary"][name]:
        r, i = H - 1 - row["j"], row["i"]
        assert a["obj_id"][r, i] == row["obj"]
        assert a["tri_id"][r, i] == row["tri"]
        assert a["t"][r, i] == row["t"], (row, a["t"][r, i])            # float64 bit-exact
        np.testing.assert_array_equal(a["normal"][r, i], np.asarray(row["normal"]) + 0.0)
        assert bool(a["front_face"][r, i]) == row["front"]


def test_kat_b_survey_values(sample_scene, sample_mesh):
    """The literal numbers quoted in SURVEY.md §8c KAT-B."""
    a = OracleRayTracer(600, 400); a.loadFromJSON(sample_scene); A = a.primary_aov()
    assert A["t"][400 - 1 - 200, 300] == 0.9930405207627144 and A["obj_id"][199, 300] == 0
    assert A["t"][399, 0] == 0.8096435856689774 and A["obj_id"][399, 0] == 3
    assert A["t"][0, 599] == 5.102766307093086
    b = OracleRayTracer(1280, 720); b.loadFromJSON(sample_mesh); B = b.primary_aov()
    assert B["t"][720 - 1 - 360, 640] == 0.800674890604618 and B["tri_id"][359, 640] == 9 and B["front_face"][359, 640] == 0
    assert B["t"][719, 0] == 0.7571329723835393 and B["obj_id"][719, 0] == 2
    assert B["t"][0, 1279] == 2.8914618913740564


def test_kat_c_scalars(kat):
    L = lib()
    s = kat["scalars"]
    out = (C.c_double * 3)()
    L.orc_tonemap(0, 1.0, _d3([1, 1, 1]), out)
    assert out[0] == s["reinhard_1"] == 0.5
    L.orc_gamma(2.2, _d3([0.5, 0.5, 0.5]), out)
    assert out[0] == pytest.approx(s["reinhard_1_gamma"], rel=1e-15) and out[0] == pytest.approx(0.7297400528407231, rel=1e-15)
    assert L.orc_quantize(out[0]) == s["reinhard_1_u8"] == 186
    L.orc_tonemap(1, 1.0, _d3([1, 1, 1]), out)
    assert out[0] == pytest.approx(s["aces_1"], rel=1e-15) and out[0] == pytest.approx(0.8037974683544302, rel=1e-15)
    L.orc_tonemap(2, 2.5, _d3([1, 2, 3]), out)
    assert list(out) == [2.5, 5.0, 7.5]
    assert L.orc_schlick(1.0, 1 / 1.5) == pytest.approx(0.04, rel=1e-14)
    assert L.orc_schlick(0.5, 1 / 1.5) == pytest.approx(s["schlick_cos05"], rel=1e-15) == pytest.approx(0.07, rel=1e-14)
    assert L.orc_quantize(float("nan")) == 0 and L.orc_quantize(-1.0) == 0 and L.orc_quantize(7.0) == 255 and L.orc_quantize(1.0) == 255
    assert L.orc_quantize(0.999999) == 254


def test_kat_c_lights(kat):
    """lights.js:22-47 — never called by the reference; these three outputs are all it pins."""
    sc = OracleScene()
    sc.add_point_light([0, 10, 0], [1, 0.5, 0.25], 3.0)
    sc.add_directional_light([-1, -1, -1], [1, 0.9, 0.7], 2.0)
    out = (C.c_double * 7)()
    sc.L.orc_illuminate(sc.h, 0, _d3([0, 0, 0]), out)
    att = kat["scalars"]["point_att_d10"]
    assert att == pytest.approx(1 / 3, rel=1e-15)
    np.testing.assert_allclose(out[:3], [0, 1, 0])
    np.testing.assert_allclose(out[3:6], [3 * att, 1.5 * att, 0.75 * att], rtol=1e-15)
    assert out[6] == 10.0
    sc.L.orc_illuminate(sc.h, 1, _d3([5, 5, 5]), out)
    np.testing.assert_allclose(out[:3], [1 / math.sqrt(3)] * 3, rtol=1e-15)
    np.testing.assert_allclose(out[3:6], [2.0, 1.8, 1.4], rtol=1e-15)
    assert out[6] == math.inf


def test_kat_c_backgrounds(kat):
    sc = OracleScene()
    s = kat["scalars"]
    sc.set_background("hdri", intensity=1.0)
    sun = np.array([-0.3, 0.6, -0.5]); sun /= np.linalg.norm(sun)
    for key, d in (("0,1,0", [0, 1, 0]), ("1,0,0", [1, 0, 0]), ("0,-1,0", [0, -1, 0]), ("sun", sun)):
        np.testing.assert_allclose(sc.background(d), s["hdri"][key], rtol=1e-13, err_msg=key)
    np.testing.assert_allclose(sc.background([0, 7, 0]), [0.6, 1.0, 1.6], rtol=1e-13)      # direction is normalised first
    sc.set_background("gradient", intensity=1.0)
    np.testing.assert_allclose(sc.background([0, 1, 0]), s["sky_up"], rtol=1e-15)
    np.testing.assert_allclose(sc.background([0, 1, 0]), [0.5, 0.7, 1.0], rtol=1e-15)
    sc.set_background("gradient", intensity=0.5)
    np.testing.assert_allclose(sc.background([1, 0, 0]), [0.375, 0.425, 0.5], rtol=1e-15)
    sc.set_background("solid", [0.2, 0.4, 0.6], 2.0)
    np.testing.assert_allclose(sc.background([1, 2, 3]), [0.4, 0.8, 1.2], rtol=1e-15)


def test_procedural_sky_terms():
    """world.js:46-72 at directions where each term can be isolated by hand."""
    sc = OracleScene()
    sc.set_perm(np.arange(256))
    sc.set_background("procedural_sky", intensity=1.0)
    # straight down: sky 0, sun 0, clouds 0 (max(0,y)=0); glow = exp(-4)*0.3; ground = 0.5
    g = math.exp(-4) * 0.3
    np.testing.assert_allclose(sc.background([0, -1, 0]), [g * 1.0 + 0.1 * 0.5, g * 0.8 + 0.15 * 0.5, g * 0.6 + 0.1 * 0.5], rtol=1e-14)
    # horizon along +x: y = 0 → sky/ground/cloud terms vanish, glow = 0.3, sun = (0.3/|s|)^512 * 10 ≈ 0
    sd = 0.3 / math.sqrt(0.09 + 0.36 + 0.64)
    sun = sd ** 512 * 10
    np.testing.assert_allclose(sc.background([1, 0, 0]), [0.3 + sun, 0.24 + 0.95 * sun, 0.18 + 0.8 * sun], rtol=1e-14)
    # Perlin noise is 0 on the integer lattice (noise.js:29-61) and bounded
    assert sc.L.orc_perlin(sc.h, _d3([3.0, -2.0, 7.0])) == 0.0
    rng = np.random.default_rng(0)
    vals = [sc.L.orc_perlin(sc.h, _d3(p)) for p in rng.uniform(-20, 20, size=(500, 3))]
    assert max(abs(v) for v in vals) <= 1.5 and np.std(vals) > 0.1
    assert sc.L.orc_turbulence(sc.h, _d3([0.3, 0.4, 0.5]), 7) >= 0


def test_perlin_hand_value():
    """One lattice cell worked by hand with the identity permutation (noise.js:20-61)."""
    sc = OracleScene()
    sc.set_perm(np.arange(256))
    p = np.arange(512) & 255
    x, y, z = 0.5, 0.25, 0.75
    fade = lambda t: t * t * t * (t * (t * 6 - 15) + 10)
    lerp = lambda t, a, b: a + t * (b - a)

    def grad(h, x, y, z):
        h &= 15
        u = x if h < 8 else y
        v = y if h < 4 else (x if h in (12, 14) else z)
        return (u if h & 1 == 0 else -u) + (v if h & 2 == 0 else -v)
    X = Y = Z = 0
    u, v, w = fade(x), fade(y), fade(z)
    A = p[X] + Y; AA = p[A] + Z; AB = p[A + 1] + Z; B = p[X + 1] + Y; BA = p[B] + Z; BB = p[B + 1] + Z
    want = lerp(w, lerp(v, lerp(u, grad(p[AA], x, y, z), grad(p[BA], x - 1, y, z)),
                        lerp(u, grad(p[AB], x, y - 1, z), grad(p[BB], x - 1, y - 1, z))),
                lerp(v, lerp(u, grad(p[AA + 1], x, y, z - 1), grad(p[BA + 1], x - 1, y, z - 1)),
                     lerp(u, grad(p[AB + 1], x, y - 1, z - 1), grad(p[BB + 1], x - 1, y - 1, z - 1))))
    assert sc.L.orc_perlin(sc.h, _d3([x, y, z])) == want


def test_kat_c_denoise_weights(kat):
    """post-processor.js:45-77: an impulse image reveals the 3x3 weights; σ=0.1 is the identity."""
    L = lib()
    W = H = 5
    img = np.zeros((H, W, 4), np.float32); img[2, 2, :3] = 1.0; img[..., 3] = 1.0
    out = np.empty_like(img)
    L.orc_denoise(img.ctypes.data_as(C.POINTER(C.c_float)), W, H, 0.5, out.ctypes.data_as(C.POINTER(C.c_float)))
    d = kat["scalars"]["denoise_sigma05"]
    assert d["edge"] == pytest.approx(0.1353352832366127, rel=1e-15) and d["total"] == pytest.approx(1.6146036885013875, rel=1e-15)
    assert out[2, 2, 0] == np.float32(1 / d["total"])
    assert out[2, 1, 0] == np.float32(d["edge"] / d["total"]) and out[1, 1, 0] == np.float32(d["corner"] / d["total"])
    assert out[0, 0, 0] == 0 and np.all(out[..., 3] == 1.0)
    L.orc_denoise(img.ctypes.data_as(C.POINTER(C.c_float)), W, H, 0.1, out.ctypes.data_as(C.POINTER(C.c_float)))
    np.testing.assert_allclose(out, img, atol=1e-20)
    # clamp-to-edge: a constant image stays constant, including at the borders
    const = np.full((4, 6, 4), 0.25, np.float32); o2 = np.empty_like(const)
    L.orc_denoise(const.ctypes.data_as(C.POINTER(C.c_float)), 6, 4, 1.0, o2.ctypes.data_as(C.POINTER(C.c_float)))
    np.testing.assert_allclose(o2, const, rtol=1e-7)
    assert kat["scalars"]["denoise_sigma1_total"] == pytest.approx(4.897640403536303, rel=1e-15)


def test_kat_d_deterministic_paths(kat):
    """Mirror metal + emissive plane + gradient sky: no RNG influence, so whole paths are known-answer."""
    D = kat["deterministic"]
    rt = OracleRayTracer(D["width"], D["height"], seed=123)
    assert rt.loadFromJSON(D["scene"])
    rt.updateRenderSettings(dict(maxBounces=D["depth"], samples=1, antiAliasing="none"))
    img = rt.render()
    for row in D["pixels"]:
        r, i = D["height"] - 1 - row["j"], row["i"]
        np.testing.assert_allclose(rt.linear[r, i, :3], row["linear"], rtol=1e-14)
        assert list(img[r, i]) == row["rgba8"], (row, img[r, i])
    # the survey's literal RGBA8 values
    assert list(img[400 - 1 - 200, 300]) == [155, 165, 176, 255]
    assert list(img[400 - 1 - 170, 300]) == [204, 176, 144, 255]
    # seed independence of a deterministic scene
    rt2 = OracleRayTracer(D["width"], D["height"], seed=999); rt2.loadFromJSON(D["scene"])
    rt2.updateRenderSettings(dict(maxBounces=D["depth"], samples=1, antiAliasing="none"))
    assert np.array_equal(rt2.render(rect=(250, 150, 350, 250))[150:250, 250:350], img[150:250, 250:350])


def test_depth_rule():
    """rayColor(depth <= 0) is black WITHOUT looking at the background (ray-tracer.js:103): with maxBounces = 1 a mirror
    pixel is black (emitted 0 + attenuation * 0) while a sky pixel still sees the background."""
    scene = dict(objects=[dict(type="sphere", center=[0, 0, -1], radius=0.5, material=dict(type="metal", color=[1, 1, 1], roughness=0))],
                 camera=dict(position=[0, 0, 1], lookAt=[0, 0, -1], fov=40, aspect=1.0, aperture=0, focusDist=2.0),
                 background=dict(type="gradient"))
    rt = OracleRayTracer(33, 33); assert rt.loadFromJSON(scene)
    rt.updateRenderSettings(dict(maxBounces=1, samples=1, antiAliasing="none"))
    rt.render()
    assert np.all(rt.linear[16, 16, :3] == 0) and np.all(rt.linear[0, 0, :3] > 0)
    rt.updateRenderSettings(dict(maxBounces=2, samples=1, antiAliasing="none"))
    rt.render()
    assert np.all(rt.linear[16, 16, :3] > 0)


def test_rng_stream_properties():
    L = lib()
    out = (C.c_double * 4096)()
    L.orc_rng_stream(42, 7, 3, 4096, out)
    a = np.array(out[:])
    assert a.min() >= 0 and a.max() < 1 and abs(a.mean() - 0.5) < 0.02 and abs(a.var() - 1 / 12) < 0.01
    assert np.all(a * 2 ** 24 == np.floor(a * 2 ** 24))                     # 24-bit uniforms: exact in fp32
    out2 = (C.c_double * 8)(); L.orc_rng_stream(42, 7, 4, 8, out2)
    assert list(out2) != list(out[:8])
    assert len(make_perm(1)) == 256 and sorted(make_perm(1)) == list(range(256))


def test_philox_random123_known_answers():
    """Philox4x32-10 against the published Random123 kat_vectors (Salmon et al., SC'11)."""
    L = lib()
    U4, U2 = C.c_uint32 * 4, C.c_uint32 * 2
    vectors = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in vectors:
        out = U4()
        L.orc_philox_raw(U4(*ctr), U2(*key), out)
        assert tuple(out) == want
    # the stream uniform is the top 24 bits of each word, consumed x, y, z, w
    out = U4(); L.orc_philox_raw(U4(5, 9, 0, 0x42525431), U2(1, 0), out)
    st = (C.c_double * 4)(); L.orc_rng_stream(1, 5, 9, 4, st)
    assert list(st) == [(w >> 8) / 2 ** 24 for w in out]


def test_threads_do_not_change_the_image(sample_scene):
    a = OracleRayTracer(60, 40, seed=5, threads=1); a.loadFromJSON(sample_scene); ia = a.render()
    b = OracleRayTracer(60, 40, seed=5, threads=4); b.loadFromJSON(sample_scene); ib = b.render()
    assert np.array_equal(ia, ib) and np.array_equal(a.linear, b.linear)
    c = OracleRayTracer(60, 40, seed=6, threads=4); c.loadFromJSON(sample_scene)
    assert not np.array_equal(c.render(), ia)


def test_oracle_matches_independent_port():
    """Cross-pin: the C++ oracle against tests/golden/independent_vectors.npz — whole stochastic images produced by a second,
    independently written restatement of the reference (tests/golden/independent_port.py, pure Python doubles) fed the same
    Philox stream.  Linear radiance must agree bit for bit (same IEEE operations in the same order), and so must the
    tone-mapped fp32 floatData and the RGBA8 bytes (both call the same libm for pow / tan / exp / sin)."""
    import json
    import os
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "independent_vectors.npz"))
    meta = json.loads(str(z["meta"]))
    assert len(meta) >= 13
    for case in meta:
        name, W, H = case["name"], case["W"], case["H"]
        rt = OracleRayTracer(W, H, seed=case["seed"], threads=2, perm_seed=case.get("perm_seed") or 0)
        if name.startswith("preset_"):
            rt.loadPreset(name[len("preset_"):])
        else:
            assert rt.loadFromJSON(case["scene"])
        rt.updateRenderSettings(dict(samples=case["spp"], maxBounces=case["depth"], antiAliasing=case.get("aa", "supersampling"),
                                     toneMapping=case.get("tonemap", "reinhard"), exposure=case.get("exposure", 1.0),
                                     gamma=case.get("gamma", 2.2), denoising=case.get("denoise", False),
                                     denoiseStrength=case.get("strength", 0.5)))
        img = rt.render()
        np.testing.assert_array_equal(rt.linear[..., :3], z[name + "_linear"], err_msg=name)
        want_f = z[name + "_float"]
        got_f = rt.denoised if case.get("denoise") else rt.floatData
        if case.get("denoise"):
            # the port keeps the un-denoised floatData (as ray-tracer.js does) and only the 8-bit image is replaced
            np.testing.assert_array_equal(rt.floatData, want_f, err_msg=name)
        else:
            np.testing.assert_array_equal(got_f, want_f, err_msg=name)
        np.testing.assert_array_equal(img, z[name + "_rgba"], err_msg=name)
