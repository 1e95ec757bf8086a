"""libbrt's native scene ingest and camera (csrc/scene_loader.cpp, host-only context: no GPU needed) against what the reference's
OWN loader makes of the same JSON — js/scene-loader.js, js/camera.js, js/geometry.js, js/materials.js, js/lights.js executed from
their unmodified source by baseline/minijs.py (tests/golden/reference_host_vectors.json, written by
baseline/make_host_fixtures_minijs.py).  Object order and kinds (skip rules shift IDs), every coordinate, `||` defaults, clamps,
the mesh index filter, lights, the derived camera vectors and the canvas size after a resolution override: exact doubles."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, HAVE_REFERENCE, REFERENCE, REFERENCE_JS
import blenderraytracer_b200 as brt
from blenderraytracer_b200 import _lib as L
from test_host_abi import _flat, _load, host  # noqa: F401  (host: fixture)

VECTORS = os.path.join(GOLDEN, "reference_host_vectors.json")
KIND = {"Sphere": L.OBJ_SPHERE, "Plane": L.OBJ_PLANE, "Box": L.OBJ_BOX, "Triangle": L.OBJ_TRIANGLE, "TriangleMesh": L.OBJ_MESH}
MAT = {"Lambertian": 0, "Metal": 1, "Dielectric": 2, "Emissive": 3}
BG = {"bound skyGradient": 0, "bound solidBackground": 1, "bound hdriBackground": 2, "bound proceduralSky": 3}


def eq(got, want, what):
    got, want = np.asarray(list(got), np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), (what, got.tolist(), want.tolist())


def check_state(lib, h, st, what, derived_only=False):
    objs, mats, lights, tris = _flat(lib, h)
    assert [o.type for o in objs] == [KIND[o["cls"]] for o in st["objects"]], what
    for i, (o, r) in enumerate(zip(objs, st["objects"])):
        w = f"{what} object {i} ({r['cls']})"
        m, rm = mats[o.material], r["material"]
        if rm is None:                                                   # a mesh whose every triangle was filtered out has no material to compare
            assert r["cls"] == "TriangleMesh" and o.tri_count == 0, w
            continue
        assert m.type == MAT[rm["cls"]], w
        if rm["cls"] in ("Lambertian", "Metal"): eq(m.color, rm["albedo"], w + " albedo")
        if rm["cls"] == "Metal": assert m.param == rm["roughness"], w
        if rm["cls"] == "Dielectric": assert m.param == rm["refractionIndex"], w
        if rm["cls"] == "Emissive":
            eq(m.color, rm["color"], w + " emission"); assert m.param == rm["intensity"], w
        if r["cls"] == "Sphere":
            eq(o.a, r["center"], w); assert o.b[0] == r["radius"], w
        elif r["cls"] == "Plane":
            eq(o.a, r["point"], w); eq(o.b, r["normal"], w + " normal (normalised in the constructor, geometry.js:52)")
        elif r["cls"] == "Box":
            eq(o.a, r["min"], w); eq(o.b, r["max"], w)
        elif r["cls"] == "Triangle":
            eq(o.a, r["v0"], w); eq(o.b, r["v1"], w); eq(o.c, r["v2"], w)
        else:
            assert o.tri_count == len(r["triangles"]), w + ": mesh triangle filtering (geometry.js:206-231)"
            for k, t in enumerate(r["triangles"]):
                eq(tris[o.first_tri + k], t[0] + t[1] + t[2], f"{w} triangle {k}")
    assert len(lights) == len(st["lights"]), what
    for l, r in zip(lights, st["lights"]):
        assert l.type == (L.LIGHT_POINT if r["cls"] == "PointLight" else L.LIGHT_DIRECTIONAL), what
        eq(l.v, r["position"] if r["cls"] == "PointLight" else r["direction"], what + " light vector")
        eq(l.color, r["color"], what + " light colour"); assert l.intensity == r["intensity"], what
    kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
    lib.brt_get_background(h, C.byref(kind), col, C.byref(inten))
    assert kind.value == BG[st["background"]], (what, st["background"])
    assert inten.value == st["skyIntensity"], what
    if st["camera"] is not None:
        c = L.brt_camera()
        assert lib.brt_get_camera(h, C.byref(c)) == L.BRT_OK
        rc = st["camera"]
        for key, got in (("origin", c.origin), ("lowerLeftCorner", c.lower_left_corner), ("horizontal", c.horizontal), ("vertical", c.vertical),
                         ("u", c.u), ("v", c.v), ("w", c.w)):
            eq(got, rc[key], f"{what} camera.{key}")
        assert c.lens_radius == rc["lensRadius"], what
        if not derived_only:       # a caller that hands over the Camera's derived members (the JS shim) keeps fov / aperture / focusDist on its side
            assert c.vfov == rc["fov"] and c.aperture == rc["aperture"] and c.focus_dist == rc["focusDist"], what
        assert c.type == {"perspective": L.CAM_PERSPECTIVE, "orthographic": L.CAM_ORTHOGRAPHIC}.get(rc["type"], L.CAM_OTHER), what


def resize_canvas_camera(lib, h, W, H):
    """What the caller of brt_scene_load_json does with the reported resolution (include/brt.h; the Python mirror's and the JS
    shim's resizeCanvas): RayTracer.resizeCanvas -> setupCamera (ray-tracer.js:598-614, 439-474) rebuilds the camera from its own
    derived vectors with aspect = width / height."""
    c = L.brt_camera()
    assert lib.brt_get_camera(h, C.byref(c)) == L.BRT_OK
    n = L.brt_camera()
    for k in range(3):
        n.look_from[k] = c.origin[k]
        n.look_at[k] = c.origin[k] - c.w[k] * c.focus_dist
        n.vup[k] = c.v[k]
    n.vfov = c.vfov if c.vfov else 45.0
    n.aspect = W / H
    n.aperture = c.aperture if c.aperture else 0.0
    n.focus_dist = c.focus_dist if c.focus_dist else 10.0
    n.type, n.use_derived = c.type, 0
    assert lib.brt_set_camera(h, C.byref(n)) == L.BRT_OK


def test_native_ingest_equals_the_reference_loader(host):
    lib, _ = host
    doc = json.load(open(VECTORS))
    assert "minijs" in doc["generator"] and len(doc["cases"]) >= 13
    for c in doc["cases"]:
        h = C.c_void_p()
        assert lib.brt_create(C.byref(h), -1) == L.BRT_OK                 # a fresh RayTracer per case, as the generator does
        try:
            W, H = c["W"], c["H"]
            cam_set = False
            for k, (scene, step) in enumerate(zip(c["scenes"], c["steps"])):
                what = f"{c['name']} step {k}"
                rc, has_cam, w, hh = _load(lib, h, scene, W, H)
                assert (rc == L.BRT_OK) == step["ok"], (what, lib.brt_last_error(h))
                if rc == L.BRT_OK and w and hh:
                    W, H = w, hh                                            # camera.resolution resized the canvas (ray-tracer.js:318-327)
                    if has_cam:
                        resize_canvas_camera(lib, h, W, H)
                st = step["state"]
                assert (W, H) == (st["width"], st["height"]), what
                if k == 0 and not step["ok"]:
                    continue
                cam_set = cam_set or bool(has_cam)
                if not cam_set:
                    # no camera in any JSON so far: the reference still holds the default camera its constructor made
                    # (ray-tracer.js:16-40 — the Python mirror and the JS shim do the same above the C ABI); the bare ctx has none
                    st = dict(st, camera=None)
                # the reference constructs its default scene in the RayTracer constructor; libbrt starts empty, so states are compared
                # from the first successful load on (a failed load must leave that state untouched in both)
                check_state(lib, h, st, what)
        finally:
            lib.brt_destroy(h)


def test_oracle_lights_match_the_reference():
    """lights.js illuminate() (js/lights.js:22-47; dead code in the reference's render loop, the formulas of the direct-lighting
    extension) executed from the reference's source, against the oracle's restatement: exact doubles."""
    from oracle.oracle import OracleRayTracer, _d3
    doc = json.load(open(VECTORS))
    assert len(doc["lights"]) >= 3
    for sp in doc["lights"]:
        sc = OracleRayTracer(8, 8).scene
        (sc.add_point_light if sp["kind"] == "point" else sc.add_directional_light)(sp["v"], sp["color"], sp["intensity"])
        idx = len(sc.lights_log) - 1                                       # after the default scene's own lights (ray-tracer.js:42-77)
        for p, want in zip(sp["points"], sp["illuminate"]):
            out = (C.c_double * 7)()
            sc.L.orc_illuminate(sc.h, idx, _d3(p), out)
            got = list(out)
            eq(got[0:3], want["direction"], "direction"); eq(got[3:6], want["color"], "colour")
            assert got[6] == (float("inf") if want["distance"] is None else want["distance"])


@pytest.mark.skipif(not HAVE_REFERENCE, reason="no reference checkout on this machine")
def test_committed_host_vectors_are_what_the_reference_source_computes(tmp_path):
    import subprocess, sys
    out = tmp_path / "host.json"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call([sys.executable, os.path.join(root, "baseline", "make_host_fixtures_minijs.py"), "--out", str(out), "--ref", REFERENCE], stdout=subprocess.DEVNULL)
    new, old = json.load(open(out)), json.load(open(VECTORS))
    assert new["cases"] == old["cases"] and new["lights"] == old["lights"] and new["controls"] == old["controls"]


def _controls_session(rt, cam_view, doc):
    """replays doc["controls"] on `rt` (an object with the reference's method names); cam_view(rt) -> dict of camera fields"""
    steps = doc["controls"]
    assert len(steps) >= 30 and steps[0]["call"] == ["constructor"]
    for k, st in enumerate(steps):
        call = st["call"]
        name, args = call[0], call[1:]
        ret = None
        if name != "constructor":
            if name == "setCameraPosition":
                args = [None if a is None else np.asarray(a, np.float64) for a in args]
            ret = getattr(rt, name)(*args)
        want, what = st["state"], f"controls step {k} {call}"
        assert (rt.width, rt.height) == (want["width"], want["height"]), what
        cam, rc = cam_view(rt), want["camera"]
        for key in ("origin", "lowerLeftCorner", "horizontal", "vertical", "u", "v", "w"):
            eq(cam[key], rc[key], f"{what} camera.{key}")
        assert (cam["lensRadius"], cam["fov"], cam["aperture"], cam["focusDist"], cam["type"]) == \
               (rc["lensRadius"], rc["fov"], rc["aperture"], rc["focusDist"], rc["type"]), what
        if name == "loadCameraPreset":
            assert ret is want["ret"], what
        if name == "getCameraPosition":
            for key in ("position", "lookAt", "up"):
                eq(ret[key], want["ret"][key], f"{what} {key}")
            assert (ret["fov"], ret["aperture"], ret["focusDist"], ret["type"]) == tuple(want["ret"][k2] for k2 in ("fov", "aperture", "focusDist", "type")), what
        yield st, what


def test_python_mirror_controls_equal_the_reference(host):
    """The reference's control surface above the C ABI — loadCameraPreset / updateCamera / setCameraPosition / getCameraPosition /
    resizeCanvas -> setupCamera / updateBackground / loadPreset (ray-tracer.js:282-299, 439-614, 627-680) — replayed on the Python
    mirror over a host-only libbrt context: after each of 31 calls the camera libbrt holds, the canvas size, the background kind
    and sky intensity are what the reference's own objects hold (exact doubles)."""
    lib, _ = host
    doc = json.load(open(VECTORS))
    rt = brt.RayTracer(600, 400, device=-1)
    try:
        def cam_view(r):
            c = r.camera
            return dict(c, lensRadius=c["lensRadius"])
        for st, what in _controls_session(rt, cam_view, doc):
            want = st["state"]
            kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
            lib.brt_get_background(rt._ctx, C.byref(kind), col, C.byref(inten))
            assert kind.value == L.BG[want["background"]] and inten.value == want["skyIntensity"], what
            assert len(_flat(lib, rt._ctx)[0]) == want["n_objects"], what
    finally:
        rt.close()


def test_oracle_controls_equal_the_reference():
    """the same session on the oracle's restatement of those host methods (oracle/oracle.py)"""
    from oracle.oracle import OracleRayTracer
    doc = json.load(open(VECTORS))
    rt = OracleRayTracer(600, 400)
    rt.setCloudPermutation(np.asarray(doc["probe_perm"], np.uint8))
    n = 0
    for st, what in _controls_session(rt, lambda r: r.scene.camera(), doc):
        for d, want in zip(doc["probe_dirs"], st["state"]["bgProbe"]):      # world.background on fixed rays: kind, colour and intensity
            if want is not None:
                eq(rt.scene.background(d), want, f"{what} background({d})"); n += 1
        assert rt.scene.object_count() == st["state"]["n_objects"], what
    assert n > 100


@pytest.mark.skipif(not HAVE_REFERENCE, reason="no reference checkout on this machine")
def test_native_ingest_equals_the_reference_loader_on_random_scenes():
    """Differential fuzz (tools/fuzz_ingest.py): 150 random scenes with missing fields, zeros, negatives, short / long arrays, odd
    capitalisation, unknown types, out-of-range mesh indices, cameras on top of their target and resolution overrides through the
    reference's own loader and through libbrt's: same objects, materials, lights, triangles, camera vectors, background, canvas.
    (1 900 scenes of five other seeds: 0 disagreements.)"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    import fuzz_ingest
    bad = fuzz_ingest.run(seed=7, n=150, ref=REFERENCE)
    assert not bad, bad[:3]
