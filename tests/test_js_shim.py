"""napi/raytracer_gpu.mjs — the JS shim a Node host puts on the reference's RayTracer — EXECUTED, unmodified, in an image without
Node.js: baseline/minijs.py runs the JavaScript, napi/napi_host.py plays Node's N-API (handles, typed-array backing stores, async
work on a worker thread, thread-safe functions, promises, setInterval ticking while the work runs) and loads the REAL
brt_addon.node, which calls the real libbrt.

  * where the reference checkout exists (the build container): the shim is installed on the reference's OWN RayTracer class; after
    loadFromJSON (the reference's loader) + render() the scene, camera and background that reached libbrt are compared with the
    reference's live objects for the 13 ingest cases of tests/golden/reference_host_vectors.json — exact doubles;
  * anywhere (the GPU box has no reference): the shim runs on tests/golden/mock_world.mjs, classes with the reference's member
    names, and on the GPU its image must equal the Python ctypes binding's byte for byte, with progress callbacks, the preview blit
    and cancellation through window.renderCancelled behaving as ray-tracer.js:166-281 prescribes."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, HAVE_REFERENCE, REFERENCE, REFERENCE_JS
import blenderraytracer_b200 as brt
from blenderraytracer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAPI = os.path.join(ROOT, "napi")
sys.path.insert(0, os.path.join(ROOT, "baseline"))
sys.path.insert(0, NAPI)
import minijs as J  # noqa: E402
import napi_host  # noqa: E402

HAVE_REF = HAVE_REFERENCE


@pytest.fixture(scope="module")
def addon():
    path = os.path.join(NAPI, "brt_addon.node")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), os.path.join(NAPI, "brt_addon.c"),
                           "-L", os.path.join(ROOT, "blenderraytracer_b200"), "-lbrt", "-Wl,-rpath," + os.path.join(ROOT, "blenderraytracer_b200"),
                           "-o", path])
    napi_host.make_forwarders()
    return path


def fake_canvas(W, H, blits):
    def create_image_data(this, a):
        w, h = int(a[0]), int(a[1])
        return J.JSObject(J.OBJECT_PROTO, {"width": float(w), "height": float(h), "data": J.JSTyped("u8c", w * h * 4)})
    def put_image_data(this, a):
        blits.append(np.asarray(a[0].get("data").items, np.float64).astype(np.uint8)); return J.UNDEF
    ctx = J.JSObject(J.OBJECT_PROTO, {"createImageData": J.native(create_image_data), "putImageData": J.native(put_image_data)})
    return J.JSObject(J.OBJECT_PROTO, {"width": float(W), "height": float(H), "style": J.JSObject(J.OBJECT_PROTO), "getContext": J.native(lambda t, a: ctx)})


def node_like(addon_path):
    """an interpreter with the N-API host installed and the shim loaded: (interp, host, shim exports)"""
    sys.setrecursionlimit(20000)
    interp = J.Interp()
    host = napi_host.NapiHost(interp, J)
    host.install_require(addon_path)
    return interp, host, interp.load_module(os.path.join(NAPI, "raytracer_gpu.mjs"))


def ctx_of(rt):
    return C.c_void_p(rt.get("_brt").props["%external"])


@pytest.mark.skipif(not HAVE_REF, reason="no reference checkout on this machine (the GPU box): the mock-world tests below cover the shim there")
def test_shim_on_the_reference_classes_hands_libbrt_the_reference_objects(addon):
    from test_reference_host_pin import VECTORS, check_state
    lib = brt.load()
    doc = json.load(open(VECTORS))
    n = 0
    for c in doc["cases"]:
        interp, host, shim = node_like(addon)
        RayTracer = interp.load_module(os.path.join(REFERENCE_JS, "ray-tracer.js"))["RayTracer"]
        interp.call(shim["installGpuRender"], J.UNDEF, [RayTracer, J.py_to_js({"device": -1})])     # host-only libbrt context: everything but the kernels
        rt = interp.construct(RayTracer, [fake_canvas(c["W"], c["H"], [])])
        oks = [J.truthy(interp.call(rt.get("loadFromJSON"), rt, [J.py_to_js(json.loads(json.dumps(s)))])) for s in c["scenes"]]
        assert oks == [s["ok"] for s in c["steps"]], c["name"]
        interp.globals.vars["window"].set("renderCancelled", False)
        with pytest.raises(J.JSThrow) as e:                          # scene, camera, background and params go in; the render itself needs a GPU
            interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
        assert "no CPU fallback" in J.to_str(e.value.value.get("message")) and not host.log, (c["name"], host.log)
        st = c["steps"][-1]["state"]
        if not any(oks):
            # nothing loaded: the reference still shows its constructor's default scene, and so does libbrt now — compare with the
            # default scene's own dump (the 'constructor' step of the controls session)
            assert len(brt_flat_objects(lib, ctx_of(rt))) == doc["controls"][0]["state"]["n_objects"], c["name"]
            continue
        check_state(lib, ctx_of(rt), st, "shim " + c["name"], derived_only=True)
        p = L.brt_render_params()
        lib.brt_get_render_params(ctx_of(rt), C.byref(p))
        assert (p.width, p.height, p.spp, p.max_depth, p.aa_mode, p.tonemap, p.preview) == (st["width"], st["height"], 4, 5, 1, 0, 1), c["name"]
        n += 1
    assert n >= 12


def brt_flat_objects(lib, h):
    d = L.brt_scene_desc()
    assert lib.brt_scene_get_flat(h, C.byref(d)) == L.BRT_OK
    return [d.objects[i] for i in range(d.n_objects)]


# ---- the shim on the mock world (runs anywhere) ---------------------------------------------------------------------------------
OBJ_KIND = {L.OBJ_SPHERE: "Sphere", L.OBJ_PLANE: "Plane", L.OBJ_BOX: "Box", L.OBJ_TRIANGLE: "Triangle", L.OBJ_MESH: "TriangleMesh"}
MAT_KIND = {0: "Lambertian", 1: "Metal", 2: "Dielectric", 3: "Emissive"}


def describe(rt):
    """the scene a ctypes-side RayTracer holds, as the plain description tests/golden/mock_world.mjs builds live objects from"""
    flat, cam = rt.sceneFlat(), rt.camera
    objs = []
    for o in flat["objects"]:
        m = flat["materials"][o["material"]]
        tris = flat["mesh_triangles"][o["first_tri"]:o["first_tri"] + o["tri_count"]].tolist() if o["type"] == L.OBJ_MESH else []
        objs.append(dict(kind=OBJ_KIND[o["type"]], a=list(o["a"]), b=list(o["b"]), c=list(o["c"]), tris=tris,
                         material=dict(kind=MAT_KIND[m["type"]], color=list(m["color"]), param=m["param"])))
    lights = [dict(kind="PointLight" if l["type"] == L.LIGHT_POINT else "DirectionalLight", v=list(l["v"]), color=list(l["color"]), intensity=l["intensity"])
              for l in flat["lights"]]
    camera = {k: [float(x) for x in cam[k]] for k in ("origin", "lowerLeftCorner", "horizontal", "vertical", "u", "v", "w")}
    camera.update(lensRadius=float(cam["lensRadius"]), type=cam["type"])
    return dict(objects=objs, lights=lights, camera=camera, perm=[float(x) for x in rt._perm])


def shim_raytracer(addon_path, desc, W, H, opts, blits):
    interp, host, shim = node_like(addon_path)
    mock = interp.load_module(os.path.join(GOLDEN, "mock_world.mjs"))
    interp.call(shim["installGpuRender"], J.UNDEF, [mock["RayTracer"], J.py_to_js(opts)])
    rt = interp.call(mock["buildRayTracer"], J.UNDEF, [fake_canvas(W, H, blits), J.py_to_js(desc)])
    interp.globals.vars["window"].set("renderCancelled", False)
    return interp, host, rt


def test_shim_on_the_mock_world_flattens_what_the_native_loader_ingested(addon):
    """host-only on both sides: scene -> native loader -> description -> mock live objects -> shim -> addon -> libbrt: the same flat scene"""
    py = brt.RayTracer(96, 64, device=-1)
    assert py.loadFromJSON(open(os.path.join(GOLDEN, "sample_mesh.json")).read())
    blits = []
    interp, host, rt = shim_raytracer(addon, describe(py), 96, 64, {"device": -1}, blits)
    with pytest.raises(J.JSThrow) as e:
        interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
    assert "no CPU fallback" in J.to_str(e.value.value.get("message")) and e.value.value.get("code") == "BRT_E" and not host.log
    assert blits == [] and not interp.intervals                         # no blit after a failed render; the cancel poll was cleared (finally)
    lib = brt.load()
    a, b = py.sceneFlat(), None
    d = L.brt_scene_desc()
    assert lib.brt_scene_get_flat(ctx_of(rt), C.byref(d)) == L.BRT_OK
    assert d.n_objects == len(a["objects"]) and d.n_mesh_triangles == len(a["mesh_triangles"]) and d.n_lights == len(a["lights"])
    for i, o in enumerate(a["objects"]):
        g = d.objects[i]
        assert (g.type, tuple(g.a), tuple(g.b), tuple(g.c), g.first_tri, g.tri_count) == (o["type"], o["a"], o["b"], o["c"], o["first_tri"], o["tri_count"])
        gm, m = d.materials[g.material], a["materials"][o["material"]]
        assert (gm.type, tuple(gm.color), gm.param) == (m["type"], m["color"], m["param"])
    assert np.array_equal(np.ctypeslib.as_array(d.mesh_triangles, shape=(int(d.n_mesh_triangles), 9)), a["mesh_triangles"])
    c = L.brt_camera()
    assert lib.brt_get_camera(ctx_of(rt), C.byref(c)) == L.BRT_OK
    for key, got in (("origin", c.origin), ("lowerLeftCorner", c.lower_left_corner), ("horizontal", c.horizontal), ("vertical", c.vertical), ("w", c.w)):
        assert np.array_equal(np.asarray(list(got)), py.camera[key]), key
    host.finalize_external(rt.get("_brt")); py.close()


@pytest.mark.gpu
def test_shim_render_on_the_gpu_equals_the_python_binding(addon):
    """RayTracer.render() of the shim, end to end on the device: JS shim (minijs) -> brt_addon.node -> libbrt -> B200.  Same bytes
    as the ctypes binding for the same scene / settings / seed; onProgress ends with 1; putImageData: one preview blit per batch
    callback + the final blit; the preview blits hold real pixels of fewer samples."""
    W, H, spp, depth, seed = 160, 100, 8, 6, 11
    py = brt.RayTracer(W, H, seed=seed + 1)                             # the shim renders with seed + frame number (frame 1)
    assert py.loadFromJSON(open(os.path.join(GOLDEN, "sample_mesh.json")).read())
    py.updateRenderSettings(dict(samples=spp, maxBounces=depth, toneMapping="aces"))
    py.updateBackground("procedural_sky", 0.8)
    py.sppBatch = 2
    want = py.render()
    blits, progress = [], []
    interp, host, rt = shim_raytracer(addon, describe(py), W, H, {"device": 0, "seed": seed, "sppBatch": 2}, blits)
    for k, val in (("samples", float(spp)), ("maxBounces", float(depth)), ("toneMapping", "aces")):
        rt.set(k, val)
    interp.call(rt.get("updateBackground"), rt, ["procedural_sky", 0.8])
    interp.call(rt.get("render"), rt, [J.native(lambda t, a: progress.append(a[0]))])
    assert not host.log, host.log
    got = np.asarray(rt.get("imageData").get("data").items, np.float64).astype(np.uint8).reshape(H, W, 4)
    assert np.array_equal(got, want), int((got != want).sum())
    assert got[..., 3].min() == 255 and got[..., :3].std() > 10
    assert progress[-1] == 1.0 and progress == sorted(progress) and len(progress) == spp // 2
    assert len(blits) == len([f for f in progress if f < 1]) + 1      # preview blits (:236-238) + the final one (:278)
    assert np.array_equal(blits[-1].reshape(H, W, 4), want)
    first = blits[0].reshape(H, W, 4)
    assert first[..., 3].min() == 255 and first[..., :3].std() > 10 and not np.array_equal(first, want)   # an image of fewer samples
    assert not interp.intervals                                         # clearInterval(poll) ran
    # a second render() is a new frame: another seed, another image of the same scene
    interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
    again = np.asarray(rt.get("imageData").get("data").items, np.float64).astype(np.uint8).reshape(H, W, 4)
    assert not np.array_equal(again, want) and np.abs(again[..., :3].astype(float).mean() - want[..., :3].astype(float).mean()) < 3
    host.finalize_external(rt.get("_brt")); py.close()


@pytest.mark.gpu
def test_shim_cancel_through_window_render_cancelled(addon):
    """window.renderCancelled = true (ui-controller.js:134-137) while the render is in flight: the shim's 50 ms poll — a setInterval
    callback, ticked by the host while the worker thread renders — calls addon.cancel; render() returns without the final blit
    (ray-tracer.js:264) and a later render() works."""
    W, H = 256, 160
    py = brt.RayTracer(W, H, device=-1)
    assert py.loadFromJSON(open(os.path.join(GOLDEN, "sample_mesh.json")).read())
    blits, progress = [], []
    interp, host, rt = shim_raytracer(addon, describe(py), W, H, {"device": 0, "seed": 3, "preview": False, "sppBatch": 100}, blits)
    rt.set("samples", 100000.0)                                         # 1000 batches, about a second of GPU work if nobody stops it
    window = interp.globals.vars["window"]
    def on_progress(t, a):
        progress.append(a[0])
        if len(progress) == 3: window.set("renderCancelled", True)
        return J.UNDEF
    r = interp.call(rt.get("render"), rt, [J.native(on_progress)])
    assert not host.log, host.log
    assert 3 <= len(progress) < 1000 and progress[-1] < 1.0 and blits == []   # stopped early, nothing blitted
    assert not interp.intervals
    window.set("renderCancelled", False)
    rt.set("samples", 4.0)
    interp.call(rt.get("render"), rt, [J.native(on_progress)])
    assert progress[-1] == 1.0 and len(blits) == 1
    host.finalize_external(rt.get("_brt")); py.close()


@pytest.mark.skipif(not HAVE_REF, reason="no reference checkout on this machine")
def test_shim_flattens_the_reference_textured_materials(addon):
    """TexturedLambertian / TexturedMetal over Checker / Noise / Marble / Wood / SolidColor textures built by the reference's own
    constructors (js/materials.js:99-126, js/textures.js): the texture table that reaches libbrt has the kind, colours, scale and
    the texture's own Perlin table (noise.p) of each live object, and the material rows point at it."""
    import make_texture_fixtures_minijs as T
    interp, host, shim = node_like(addon)
    js = REFERENCE_JS
    RayTracer = interp.load_module(js + "/ray-tracer.js")["RayTracer"]
    Vec3 = interp.load_module(js + "/math.js")["Vec3"]
    tex_ex, mat_ex = interp.load_module(js + "/textures.js"), interp.load_module(js + "/materials.js")
    interp.call(shim["installGpuRender"], J.UNDEF, [RayTracer, J.py_to_js({"device": -1})])
    rt = interp.construct(RayTracer, [fake_canvas(64, 48, [])])
    assert J.truthy(interp.call(rt.get("loadFromJSON"), rt, [J.py_to_js(json.load(open(os.path.join(GOLDEN, "sample_scene.json"))))]))
    objs = rt.get("world").get("objects").items
    by_name = {t["name"]: t for t in T.TEXTURES}
    assign = {0: "checker10", 1: "marble", 2: "wood", 3: "noise4", 4: "solid"}
    used = {}
    for i, name in assign.items():
        if i >= len(objs): continue
        m = objs[i].get("material"); cls = m.proto.get("constructor").name
        tex = T.make_texture(interp, tex_ex, Vec3, by_name[name])
        if cls == "Lambertian": objs[i].set("material", interp.construct(mat_ex["TexturedLambertian"], [tex])); used[i] = (name, 0, 0.0)
        elif cls == "Metal": objs[i].set("material", interp.construct(mat_ex["TexturedMetal"], [tex, m.get("roughness")])); used[i] = (name, 1, m.get("roughness"))
    assert len(used) >= 3
    interp.globals.vars["window"].set("renderCancelled", False)
    with pytest.raises(J.JSThrow):
        interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
    assert not host.log, host.log
    d = L.brt_scene_desc()
    assert brt.load().brt_scene_get_flat(ctx_of(rt), C.byref(d)) == L.BRT_OK
    assert d.n_textures == len(used) and d.n_objects == len(objs)
    for i, (name, mat_type, rough) in used.items():
        t, mat = by_name[name], d.materials[d.objects[i].material]
        assert mat.type == mat_type and mat.param == rough and 1 <= mat.texture <= d.n_textures, (i, name)
        tx = d.textures[mat.texture - 1]
        assert tx.kind == L.TEX[t["kind"]] and tx.scale == t["scale"], (i, name)
        assert tuple(tx.odd) == tuple(map(float, t["odd"])), (i, name)
        if t["kind"] == "checker": assert tuple(tx.even) == tuple(map(float, t["even"]))
        if t["perm_seed"] is not None: assert list(tx.perm) == T.perm_of(t["perm_seed"]), (i, name)
    for i in range(d.n_objects):
        if i not in used: assert d.materials[d.objects[i].material].texture == 0
    host.finalize_external(rt.get("_brt"))


@pytest.mark.skipif(not HAVE_REF, reason="no reference checkout on this machine")
def test_shim_flattens_the_reference_presets_and_ui_changes(addon):
    """The scenes a user of the reference actually builds — loadPreset('default' | 'glass' | 'metals' | 'cornell') (ray-tracer.js:282-435:
    hollow glass sphere with a negative radius, boxes, emissive panels), then UI changes (camera preset, background, resize) — reach
    libbrt as the live objects are: compared with a field-by-field dump of the SAME live World / Camera after each render() call."""
    from make_host_fixtures_minijs import dump_state
    from test_reference_host_pin import check_state
    lib = brt.load()
    interp, host, shim = node_like(addon)
    RayTracer = interp.load_module(os.path.join(REFERENCE_JS, "ray-tracer.js"))["RayTracer"]
    interp.call(shim["installGpuRender"], J.UNDEF, [RayTracer, J.py_to_js({"device": -1})])
    rt = interp.construct(RayTracer, [fake_canvas(600, 400, [])])
    interp.globals.vars["window"].set("renderCancelled", False)

    def render_and_compare(what, bg_kind, bg_color=None, bg_intensity=1.0):
        with pytest.raises(J.JSThrow):
            interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
        assert not host.log, (what, host.log)
        st = dump_state(rt)
        st["background"] = {0: "bound skyGradient", 1: "bound solidBackground", 2: "bound hdriBackground", 3: "bound proceduralSky"}[bg_kind]
        check_state(lib, ctx_of(rt), st, what, derived_only=True)
        kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
        lib.brt_get_background(ctx_of(rt), C.byref(kind), col, C.byref(inten))
        assert (kind.value, inten.value) == (bg_kind, bg_intensity), what
        if bg_color is not None: assert tuple(col) == bg_color, what
        p = L.brt_render_params()
        lib.brt_get_render_params(ctx_of(rt), C.byref(p))
        assert (p.width, p.height) == (st["width"], st["height"]), what

    render_and_compare("constructor's default scene", 0)
    for preset, kind, color in (("glass", 0, None), ("metals", 0, None), ("cornell", 1, (0.0, 0.0, 0.0)), ("default", 0, None)):
        interp.call(rt.get("loadPreset"), rt, [preset])
        render_and_compare("preset " + preset, kind, color)
    assert any(o["cls"] == "Sphere" and o["radius"] < 0 for o in (interp.call(rt.get("loadPreset"), rt, ["glass"]), dump_state(rt))[1]["objects"])
    render_and_compare("glass again (hollow sphere: negative radius)", 0)
    interp.call(rt.get("loadCameraPreset"), rt, ["close-up"]); render_and_compare("camera preset close-up", 0)
    interp.call(rt.get("updateBackground"), rt, ["procedural_sky", 0.5]); render_and_compare("procedural sky", 3, None, 0.5)
    interp.call(rt.get("updateBackground"), rt, ["solid", 2.0]); render_and_compare("solid background", 1, (0.1, 0.1, 0.1), 2.0)
    interp.call(rt.get("resizeCanvas"), rt, [320.0, 200.0]); render_and_compare("resized canvas", 1, (0.1, 0.1, 0.1), 2.0)
    interp.call(rt.get("updateRenderSettings"), rt, [J.py_to_js(dict(samples=9, maxBounces=7, toneMapping="aces", antiAliasing="stochastic", gamma=1.8, exposure=1.5))])
    render_and_compare("render settings", 1, (0.1, 0.1, 0.1), 2.0)
    p = L.brt_render_params()
    lib.brt_get_render_params(ctx_of(rt), C.byref(p))
    assert (p.spp, p.max_depth, p.tonemap, p.aa_mode, p.gamma, p.exposure) == (9, 7, 1, 2, 1.8, 1.5)
    host.finalize_external(rt.get("_brt"))


@pytest.mark.skipif(not HAVE_REF, reason="no reference checkout on this machine")
def test_shim_flattening_on_random_scenes(addon):
    """120 random scenes (tools/fuzz_ingest.py's generator: empty meshes, filtered triangles, unknown types, zero normals, every light
    and background kind ...) loaded by the reference's own loader, flattened by the shim, handed to libbrt through the real addon:
    the context then holds exactly what a dump of the live World / Camera says."""
    import random
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_ingest
    from make_host_fixtures_minijs import dump_state
    from test_reference_host_pin import check_state
    lib = brt.load()
    interp, host, shim = node_like(addon)
    RayTracer = interp.load_module(os.path.join(REFERENCE_JS, "ray-tracer.js"))["RayTracer"]
    interp.call(shim["installGpuRender"], J.UNDEF, [RayTracer, J.py_to_js({"device": -1})])
    interp.globals.vars["window"].set("renderCancelled", False)
    r = random.Random(17)
    n = 0
    for k in range(120):
        scene = fuzz_ingest.gen_scene(r)
        if scene.get("camera"): scene["camera"].pop("resolution", None)     # keeps imageData small: the host mirrors it into a C buffer per call
        rt = interp.construct(RayTracer, [fake_canvas(64, 40, [])])
        if not J.truthy(interp.call(rt.get("loadFromJSON"), rt, [J.py_to_js(json.loads(json.dumps(scene)))])):
            continue
        with pytest.raises(J.JSThrow) as e:
            interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
        assert "no CPU fallback" in J.to_str(e.value.value.get("message")) and not host.log, (k, J.to_str(e.value.value.get("message")), host.log, scene)
        st = dump_state(rt)
        bg = (scene.get("background") or {}).get("type")
        st["background"] = {"solid": "bound solidBackground", "hdri": "bound hdriBackground", "procedural_sky": "bound proceduralSky"}.get(bg, "bound skyGradient")
        check_state(lib, ctx_of(rt), st, f"shim fuzz {k}", derived_only=True)
        host.finalize_external(rt.get("_brt"))
        n += 1
    assert n >= 100


@pytest.mark.gpu
def test_shim_render_over_two_gpus(addon):
    """installGpuRender(RayTracer, { devices: [0, 1] }): ONE render() call of the JS shim, the samples split over two GPUs inside
    libbrt (brt_create_multi) — the image is the single-GPU image up to the fp32 order of the two partial sums (<= 1 LSB)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    W, H, spp, depth, seed = 160, 100, 16, 6, 5
    py = brt.RayTracer(W, H, seed=seed + 1)
    assert py.loadFromJSON(open(os.path.join(GOLDEN, "sample_mesh.json")).read())
    py.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    py.updateBackground("hdri", 1.2)                                    # the mock world starts from its own default background: set it on both sides
    want = py.render()
    blits, progress = [], []
    interp, host, rt = shim_raytracer(addon, describe(py), W, H, {"devices": [0, 1], "seed": seed}, blits)
    rt.set("samples", float(spp)); rt.set("maxBounces", float(depth))
    interp.call(rt.get("updateBackground"), rt, ["hdri", 1.2])
    interp.call(rt.get("render"), rt, [J.native(lambda t, a: progress.append(a[0]))])
    assert not host.log, host.log
    got = np.asarray(rt.get("imageData").get("data").items, np.float64).astype(np.uint8).reshape(H, W, 4)
    d = np.abs(got.astype(int) - want.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01 and progress[-1] == 1.0
    stats = interp.call(interp.builtin_modules["node:module"]["createRequire"].native(None, [""]).native(None, ["./brt_addon.node"]).get("stats"), J.UNDEF, [rt.get("_brt")])
    assert stats.get("devices") == 2.0
    host.finalize_external(rt.get("_brt")); py.close()
