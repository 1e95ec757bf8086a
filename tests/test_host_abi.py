"""CPU-side tests (-m "not gpu"): the C ABI library loads and exports every symbol include/brt.h declares, the host logic
(scene ingest, camera, parameter handling — scene-loader.js / camera.js / ray-tracer.js setters restated natively) agrees
with the oracle's independent Python restatement, and every compute entry point fails loudly without a GPU.
No compute calls are made here."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import blenderraytracer_b200 as brt
from blenderraytracer_b200 import _lib as L
from conftest import ROOT, load_scene
from oracle.oracle import OracleRayTracer, SceneLoader


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "brt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(brt_[a-z0-9_]+)\s*\(", src)) - {"brt_progress_cb"})


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(L.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/brt.h but not exported by libbrt.so"
    assert set(syms) == set(L.SIGNATURES), set(syms) ^ set(L.SIGNATURES)       # the ctypes binding covers the header exactly
    assert brt.load().brt_abi_version() == 3
    assert b"sm_100a" in brt.load().brt_version()


def test_struct_layouts_match_header(tmp_path):
    """sizeof() and every field offset of every ABI struct, as gcc lays them out from include/brt.h, equal the ctypes
    mirror in blenderraytracer_b200/_lib.py (a C compiler is the judge)."""
    import subprocess
    structs = ["brt_material", "brt_object", "brt_light", "brt_scene_desc", "brt_camera", "brt_render_params", "brt_scene_info", "brt_stats"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "brt.h"', "int main(void) {"]
    for sname in structs:
        lines.append(f'  printf("{sname} %zu\\n", sizeof({sname}));')
        for fname, _ in getattr(L, sname)._fields_:
            lines.append(f'  printf("{sname}.{fname} %zu\\n", offsetof({sname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for sname in structs:
        cls = getattr(L, sname)
        assert C.sizeof(cls) == int(out[sname]), sname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == int(out[f"{sname}.{fname}"]), f"{sname}.{fname}"


@pytest.fixture()
def host():
    """A host-only context (device_id = -1): ingest / camera / parameter logic, no device."""
    h = C.c_void_p()
    lib = brt.load()
    assert lib.brt_create(C.byref(h), -1) == L.BRT_OK
    yield lib, h
    lib.brt_destroy(h)


def _load(lib, h, scene, W=600, H=400):
    text = scene if isinstance(scene, bytes) else json.dumps(scene).encode()
    hc, w, hh = C.c_int(), C.c_int(), C.c_int()
    rc = lib.brt_scene_load_json(h, text, len(text), W, H, C.byref(hc), C.byref(w), C.byref(hh))
    return rc, hc.value, w.value, hh.value


def _flat(lib, h):
    d = L.brt_scene_desc()
    assert lib.brt_scene_get_flat(h, C.byref(d)) == L.BRT_OK
    objs = [d.objects[i] for i in range(d.n_objects)]
    mats = [d.materials[i] for i in range(d.n_materials)]
    lights = [d.lights[i] for i in range(d.n_lights)]
    n = int(d.n_mesh_triangles)
    tris = np.ctypeslib.as_array(d.mesh_triangles, shape=(n, 9)).copy() if n else np.zeros((0, 9))
    return objs, mats, lights, tris


def _compare_with_oracle_loader(lib, h, scene, W=600, H=400):
    """libbrt's native ingest vs the oracle's Python SceneLoader on the same JSON: object order / kinds / numbers,
    materials, lights, mesh triangle filtering, camera."""
    rc, has_cam, w, hh = _load(lib, h, scene, W, H)
    assert rc == L.BRT_OK, lib.brt_last_error(h)
    orc = OracleRayTracer(W, H)
    assert orc.loadFromJSON(scene)
    objs, mats, lights, tris = _flat(lib, h)
    log = orc.scene.objects_log
    assert len(objs) == len(log) == orc.scene.object_count()
    kinds = {"sphere": L.OBJ_SPHERE, "plane": L.OBJ_PLANE, "box": L.OBJ_BOX, "triangle": L.OBJ_TRIANGLE, "mesh": L.OBJ_MESH}
    mkind = {"lambertian": 0, "metal": 1, "dielectric": 2, "emissive": 3}
    for i, (o, rec) in enumerate(zip(objs, log)):
        assert o.type == kinds[rec[0]], (i, rec[0])
        m, (mt, mcol, mparam) = mats[o.material], rec[1]
        assert m.type == mkind[mt]
        if mt != "dielectric":
            np.testing.assert_array_equal(list(m.color), mcol)
        if mt == "metal":
            assert m.param == min(mparam, 1.0) or (m.param != m.param)
        elif mt in ("dielectric", "emissive"):
            assert m.param == mparam
        if rec[0] == "sphere":
            np.testing.assert_array_equal(list(o.a), rec[2]); assert o.b[0] == rec[3]
        elif rec[0] == "plane":
            np.testing.assert_array_equal(list(o.a), rec[2])
            n = np.array(rec[3], dtype=np.float64); ln = np.sqrt((n * n).sum())
            np.testing.assert_allclose(list(o.b), n / ln if ln > 0 else n * 0, rtol=1e-15)
        elif rec[0] == "box":
            np.testing.assert_array_equal(list(o.a), rec[2]); np.testing.assert_array_equal(list(o.b), rec[3])
        elif rec[0] == "triangle":
            np.testing.assert_array_equal(list(o.a), rec[2]); np.testing.assert_array_equal(list(o.b), rec[3]); np.testing.assert_array_equal(list(o.c), rec[4])
        else:
            assert o.tri_count == orc.scene.mesh_triangle_count(i), "mesh triangle filtering (geometry.js:206-231)"
    assert len(lights) == len(orc.scene.lights_log)
    for l, rec in zip(lights, orc.scene.lights_log):
        assert l.type == (L.LIGHT_POINT if rec[0] == "point" else L.LIGHT_DIRECTIONAL)
        v = np.array(rec[1], dtype=np.float64)
        if rec[0] == "directional":
            v = v / np.sqrt((v * v).sum())
        np.testing.assert_allclose(list(l.v), v, rtol=1e-15)
        np.testing.assert_array_equal(list(l.color), rec[2]); assert l.intensity == rec[3]
    if has_cam:
        c = L.brt_camera()
        assert lib.brt_get_camera(h, C.byref(c)) == L.BRT_OK
        oc = orc.scene.camera() if not (w and hh) else None
        if oc is not None:
            for key, got in (("origin", c.origin), ("lowerLeftCorner", c.lower_left_corner), ("horizontal", c.horizontal),
                             ("vertical", c.vertical), ("u", c.u), ("v", c.v), ("w", c.w)):
                np.testing.assert_allclose(list(got), oc[key], rtol=0, atol=1e-15, err_msg=key)
            assert c.lens_radius == oc["lensRadius"]
    kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
    lib.brt_get_background(h, C.byref(kind), col, C.byref(inten))
    assert kind.value == {"gradient": 0, "solid": 1, "hdri": 2, "procedural_sky": 3}[orc.scene.bg_kind]
    assert inten.value == orc.scene.sky_intensity
    return objs, mats, lights, tris, orc


def test_ingest_fixtures_match_oracle_loader(host, sample_scene, sample_mesh):
    lib, h = host
    objs, *_ = _compare_with_oracle_loader(lib, h, sample_scene)
    assert [o.type for o in objs] == [0, 0, 0, 1]
    objs, mats, lights, tris, _ = _compare_with_oracle_loader(lib, h, sample_mesh, 1280, 720)
    assert objs[0].type == L.OBJ_MESH and objs[0].tri_count == 12 and tris.shape == (12, 9)
    assert len(lights) == 2


def test_ingest_kat_camera(host, kat, sample_scene, sample_mesh):
    lib, h = host
    for name, scene, W, H in (("sample_scene", sample_scene, 600, 400), ("sample_mesh", sample_mesh, 1280, 720)):
        assert _load(lib, h, scene, W, H)[0] == L.BRT_OK
        c = L.brt_camera()
        lib.brt_get_camera(h, C.byref(c))
        k = kat["camera"][name]
        for key, got in (("w", c.w), ("u", c.u), ("v", c.v), ("horizontal", c.horizontal), ("vertical", c.vertical),
                         ("lowerLeftCorner", c.lower_left_corner), ("origin", c.origin)):
            np.testing.assert_allclose(list(got), k[key], rtol=0, atol=1e-15, err_msg=f"{name}.{key}")


def test_ingest_defaults_and_skip_rules(host):
    """scene-loader.js defaults: missing material -> lambertian 0.8; radius 0/absent -> 1.0; unknown / typeless objects are
    skipped (IDs shift); roughness clamp; ior / intensity defaults; bad vec3 -> (0,0,0); mesh index filtering; light
    defaults; camera lookAt-too-close push; focusDist default; resolution override."""
    lib, h = host
    scene = dict(
        objects=[
            dict(type="Sphere", center=[1, 2, 3]),                                       # upper case type, no radius, no material
            dict(type="sphere", center=[0, 0, 0], radius=0, material=dict(type="METAL", color=[1, 1, 1], roughness=7)),
            dict(type="torus", center=[0, 0, 0]),                                        # unknown -> skipped
            dict(center=[9, 9, 9]),                                                      # no type -> skipped
            dict(type="plane", point=[0, -1, 0], normal=[0, 5, 0], material=dict(type="dielectric")),
            dict(type="box", min=[0, 0], max=[1, 1, 1], material=dict(type="emissive", color=[1, 0.5, 0.25])),   # short vec3 -> 0
            dict(type="mesh", vertices=[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]],
                 indices=[0, 1, 2, 1, 3, 2, 0, 1, 9, 0, 1, -1, 0, 1, 2.5, 3, 2], material=dict(type="plastic")),
            dict(type="mesh", vertices=[[0, 0, 0]]),                                     # no indices -> skipped
            dict(type="triangle", v0=[0, 0, 0], v1=[1, 0, 0], v2=[0, 1, 0], material=dict(type="lambertian", color=[0.1, 0.2, 0.3])),
        ],
        lights=[dict(type="point", position=[1, 2, 3]), dict(type="DIRECTIONAL", direction=[0, -2, 0], color=[1, 0, 0], intensity=3),
                dict(type="spot"), dict(position=[0, 0, 0])],
        camera=dict(position=[0, 0, 0.5], lookAt=[0, 0, 0], fov=30),
        background=dict(type="procedural_sky", intensity=0.5),
    )
    objs, mats, lights, tris, orc = _compare_with_oracle_loader(lib, h, scene)
    assert [o.type for o in objs] == [0, 0, 1, 2, 4, 3]
    assert objs[0].b[0] == 1.0 and objs[1].b[0] == 1.0                                  # `radius || 1.0`
    assert mats[objs[0].material].type == 0 and list(mats[objs[0].material].color) == [0.8, 0.8, 0.8]
    assert mats[objs[1].material].param == 1.0                                          # Math.min(roughness, 1)
    assert mats[objs[2].material].param == 1.5                                          # ior default
    assert mats[objs[3].material].param == 1.0 and list(objs[3].a) == [0, 0, 0]         # emissive intensity default; bad vec3
    assert mats[objs[4].material].type == 0                                             # unknown material type -> lambertian
    # mesh: (0,1,2) (1,3,2) kept; (0,1,9) dropped (index >= n); (0,1,-1) and (0,1,2.5) kept with (0,0,0) vertices; tail (3,2) dropped
    assert objs[4].tri_count == 4
    np.testing.assert_array_equal(tris[2], [0, 0, 0, 1, 0, 0, 0, 0, 0])
    np.testing.assert_array_equal(tris[3], [0, 0, 0, 1, 0, 0, 0, 0, 0])
    assert len(lights) == 2 and list(lights[0].color) == [1, 1, 1] and lights[0].intensity == 1.0
    np.testing.assert_array_equal(list(lights[1].v), [0, -1, 0])
    c = L.brt_camera()
    lib.brt_get_camera(h, C.byref(c))
    np.testing.assert_allclose(list(c.look_at), [0, 0, -99.5])                          # pushed 100 units along the view direction
    assert c.focus_dist == pytest.approx(100.0) and c.vfov == 30 and c.aspect == 1.5 and c.type == L.CAM_PERSPECTIVE


def test_ingest_resolution_override_and_errors(host, sample_scene):
    lib, h = host
    sc = json.loads(json.dumps(sample_scene))
    sc["camera"]["resolution"] = [320, 200]
    del sc["camera"]["aspect"]
    rc, has_cam, w, hh = _load(lib, h, sc)
    assert (rc, has_cam, w, hh) == (L.BRT_OK, 1, 320, 200)
    c = L.brt_camera(); lib.brt_get_camera(h, C.byref(c))
    assert c.aspect == 320 / 200                                                        # `camData.aspect || width/height` with the override
    # failures: the reference throws inside loadFromJSON and returns false (ray-tracer.js:330-333); the old scene survives
    for bad in (b"{not json", b"[1,2,3]", json.dumps({"objects": [None]}).encode(), json.dumps({"objects": [{"type": 5}]}).encode()):
        assert _load(lib, h, bad)[0] == L.BRT_E_PARSE
        assert lib.brt_last_error(h)
    objs, *_ = _flat(lib, h)
    assert len(objs) == 4
    # a scene without a camera keeps the current camera (ray-tracer.js:315-317)
    rc, has_cam, *_ = _load(lib, h, dict(objects=[dict(type="sphere", center=[0, 0, 0], radius=2)]))
    assert rc == L.BRT_OK and has_cam == 0
    c2 = L.brt_camera(); assert lib.brt_get_camera(h, C.byref(c2)) == L.BRT_OK and list(c2.origin) == list(c.origin)
    assert lib.brt_scene_load_json(h, b"{}", 2, 0, 400, None, None, None) == L.BRT_E_INVALID


def test_background_loader_deviation_d1(host):
    """SURVEY F9 / deviation D1: JSON `solid` / `hdri` backgrounds are honoured as intended (ray-tracer.js:573-576)."""
    lib, h = host
    assert _load(lib, h, dict(objects=[], background=dict(type="solid", color=[0.2, 0.3, 0.4], intensity=2)))[0] == L.BRT_OK
    kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
    lib.brt_get_background(h, C.byref(kind), col, C.byref(inten))
    assert (kind.value, list(col), inten.value) == (L.BG["solid"], [0.2, 0.3, 0.4], 2.0)
    assert _load(lib, h, dict(objects=[], background=dict(type="nebula")))[0] == L.BRT_OK
    lib.brt_get_background(h, C.byref(kind), col, C.byref(inten))
    assert kind.value == L.BG["gradient"] and inten.value == 1.0


def test_flat_scene_roundtrip_and_validation(host):
    lib, h = host
    w = brt.World()
    w.add(brt.Plane((0, -0.5, 0), (0, 3, 0), brt.Lambertian((0.5, 0.5, 0.5))))
    w.add(brt.Sphere((0, 0, -1), -0.45, brt.Dielectric(1.5)))                          # negative radius kept (hollow-glass idiom)
    w.add(brt.Box((-1, -1, -1), (1, 1, 1), brt.Metal((0.8, 0.8, 0.9), 3.0)))
    w.add(brt.TriangleMesh([[0, 0, 0], [1, 0, 0], [0, 1, 0]], [0, 1, 2, 0, 1], brt.Emissive((1, 1, 1), 5)))
    w.addLight(brt.DirectionalLight((0, -4, 0), (1, 1, 1), 2))
    desc, keep = w.flatten()
    assert lib.brt_scene_set_flat(h, C.byref(desc)) == L.BRT_OK
    objs, mats, lights, tris = _flat(lib, h)
    assert [o.type for o in objs] == [1, 0, 2, 4]
    np.testing.assert_array_equal(list(objs[0].b), [0, 1, 0])                            # Plane ctor normalises (geometry.js:52)
    assert objs[1].b[0] == -0.45 and mats[2].param == 1.0 and objs[3].tri_count == 1
    np.testing.assert_array_equal(list(lights[0].v), [0, -1, 0])
    info = L.brt_scene_info()
    assert lib.brt_scene_info_get(h, C.byref(info)) == L.BRT_OK
    assert (info.n_objects, info.n_spheres, info.n_planes, info.n_boxes, info.n_triangles) == (4, 1, 1, 1, 1)
    # BRT_SCENE_CONSTRUCTED: rows read from constructed objects are stored as they are (a second normalisation moves the last bit of
    # a vector such as (1.179, 1e6, 0.001) / |.|); brt_scene_get_flat hands back post-constructor values with the flag set, so that
    # get -> set round-trips exactly
    w2 = brt.World()
    w2.add(brt.Plane((0, 0, 0), (1.179, 1000000.0, 0.001), brt.Metal((0.5, 0.5, 0.5), 7.0)))
    w2.addLight(brt.DirectionalLight((0.3, -1.7, 0.2), (1, 1, 1), 1))
    d2, keep2 = w2.flatten()
    assert d2.flags == 0 and lib.brt_scene_set_flat(h, C.byref(d2)) == L.BRT_OK
    objs, mats, lights, _ = _flat(lib, h)
    n1, l1 = list(objs[0].b), list(lights[0].v)
    assert abs(np.linalg.norm(n1) - 1) < 1e-15 and mats[0].param == 1.0
    got = L.brt_scene_desc()
    assert lib.brt_scene_get_flat(h, C.byref(got)) == L.BRT_OK and got.flags == L.SCENE_CONSTRUCTED
    for _ in range(3):                                                                   # any number of round trips: bit-stable
        rt_desc = L.brt_scene_desc()
        assert lib.brt_scene_get_flat(h, C.byref(rt_desc)) == L.BRT_OK
        o2 = (L.brt_object * 1)(rt_desc.objects[0]); m2 = (L.brt_material * 1)(rt_desc.materials[0]); li2 = (L.brt_light * 1)(rt_desc.lights[0])
        back = L.brt_scene_desc(); back.objects, back.n_objects, back.materials, back.n_materials, back.lights, back.n_lights = o2, 1, m2, 1, li2, 1
        back.flags = L.SCENE_CONSTRUCTED
        assert lib.brt_scene_set_flat(h, C.byref(back)) == L.BRT_OK
        objs, mats, lights, _ = _flat(lib, h)
        assert list(objs[0].b) == n1 and list(lights[0].v) == l1
    raw = (L.brt_object * 1)(objs[0]); raw[0].b[0], raw[0].b[1], raw[0].b[2] = 0.0, 3.0, 0.0
    back.objects, back.flags = raw, L.SCENE_CONSTRUCTED
    assert lib.brt_scene_set_flat(h, C.byref(back)) == L.BRT_OK and list(_flat(lib, h)[0][0].b) == [0.0, 3.0, 0.0]   # as is: the caller said so
    # validation
    bad = L.brt_scene_desc(); bad.n_objects = 1
    assert lib.brt_scene_set_flat(h, C.byref(bad)) == L.BRT_E_INVALID
    desc.objects[0].material = 99
    assert lib.brt_scene_set_flat(h, C.byref(desc)) == L.BRT_E_INVALID
    assert b"material" in lib.brt_last_error(h)


def test_render_params_validation(host):
    lib, h = host
    p = L.brt_render_params()
    assert lib.brt_get_render_params(h, C.byref(p)) == L.BRT_OK
    assert (p.width, p.height, p.spp, p.max_depth, p.gamma, p.exposure) == (600, 400, 4, 5, 2.2, 1.0)    # ray-tracer.js:19-30
    for field, val in (("width", 0), ("height", -3), ("spp", 0), ("max_depth", -1), ("aa_mode", 9), ("tonemap", 5), ("sampler", 2), ("accel", 3),
                       ("bvh_width", 3), ("bvh_width", 16), ("bvh_width", -2)):
        q = L.brt_render_params(); C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
        setattr(q, field, val)
        assert lib.brt_set_render_params(h, C.byref(q)) == L.BRT_E_INVALID, field
    for w in (0, 2, 4, 8):                                           # 0 = auto; 2 = binary LBVH; 4 / 8 = its wide collapse
        q = L.brt_render_params(); C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
        q.bvh_width = w
        assert lib.brt_set_render_params(h, C.byref(q)) == L.BRT_OK, w
    assert lib.brt_set_render_params(h, C.byref(p)) == L.BRT_OK
    assert lib.brt_set_background(h, 7, None, 1.0, None) == L.BRT_E_INVALID
    cam = L.brt_camera(); cam.type = 5
    assert lib.brt_set_camera(h, C.byref(cam)) == L.BRT_E_INVALID


def test_no_cpu_fallback(host, sample_scene):
    """Every compute entry point needs a CUDA device: on a host-only context each one fails with BRT_E_CUDA, and a RayTracer
    cannot even be constructed on a machine without a GPU."""
    lib, h = host
    assert _load(lib, h, sample_scene)[0] == L.BRT_OK
    buf = (C.c_uint8 * (600 * 400 * 4))()
    f = (C.c_float * 16)(); d = (C.c_double * 16)(); i32 = (C.c_int32 * 16)()
    vp = C.c_void_p()
    hd = C.create_string_buffer(64)
    calls = [
        lambda: lib.brt_render(h, buf, None, None, L.PROGRESS_CB(), None),
        lambda: lib.brt_render_accumulate(h, None, 0, 1),
        lambda: lib.brt_resolve_device(h, None, None, None, None),
        lambda: lib.brt_primary_aov_f32(h, i32, i32, f, f, buf),
        lambda: lib.brt_primary_aov_f64(h, i32, i32, d, d, buf),
        lambda: lib.brt_eval_background(h, d, 1, f),
        lambda: lib.brt_debug_rng_stream(h, 1, 0, 0, 4, f),
        lambda: lib.brt_postprocess_host(h, f, buf, None),
        lambda: lib.brt_measure_fp32_peak(h, d),
        lambda: lib.brt_stream_synchronize(h),
        lambda: lib.brt_set_stream(h, None),
        lambda: lib.brt_shared_alloc(h, 64, C.byref(vp), hd),
        lambda: lib.brt_device_memset(h, C.c_void_p(16), 0, 16),
        lambda: lib.brt_copy_to_host(h, buf, C.c_void_p(16), 16),
    ]
    for k, call in enumerate(calls):
        assert call() == L.BRT_E_CUDA, k
        assert b"no CPU fallback" in lib.brt_last_error(h) or b"host-only" in lib.brt_last_error(h)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(brt.BrtError):
            brt.RayTracer(64, 64)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under blenderraytracer_b200/ (Python or C/CUDA sources) may import, link,
    load or call it, and libbrt.so must not depend on liboracle.so."""
    import subprocess
    pkg = os.path.join(ROOT, "blenderraytracer_b200")
    needles = ("liboracle", "import oracle", "from oracle", "oracle/", "oracle.oracle", "orc_", "brt_oracle")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for n in needles:
                    assert n not in text, f"{os.path.join(dirpath, f)} references the oracle ({n})"
    deps = subprocess.check_output(["ldd", L.LIB_PATH], text=True)
    assert "oracle" not in deps
    syms = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH], text=True)
    assert "orc_" not in syms


def test_scene_generators_are_deterministic():
    from tools import gen_scenes
    a, b = gen_scenes.random_spheres(), gen_scenes.random_spheres()
    assert json.dumps(a) == json.dumps(b) and len(a["objects"]) == 486
    assert sum(o["type"] == "sphere" for o in a["objects"]) == 485
    c4 = gen_scenes.cornell()
    assert [o["type"] for o in c4["objects"]].count("mesh") == 3 and c4["background"]["type"] == "procedural_sky"
    V, I = gen_scenes.terrain_arrays(quads=16)
    assert V.shape == (17 * 17, 3) and I.shape == (16 * 16 * 6,) and I.max() == 17 * 17 - 1
    V5, I5 = gen_scenes.terrain_arrays()
    assert I5.shape[0] // 3 == 1002528
    # triangles of the C5 grid are large enough that the reference's absolute |a| < 1e-4 cull (geometry.js:157) stays negligible
    e = np.linalg.norm(V5[I5[1]] - V5[I5[0]])
    assert e > 0.25


def test_binary_container_ingest_equals_json(host, sample_mesh):
    """the BRTSCN01 container (tools/scene_binary.py) ingests to exactly the scene the JSON gives,
    including TriangleMesh's index filtering; malformed containers are rejected like malformed JSON."""
    from tools import gen_scenes, scene_binary
    lib, h = host
    ragged = json.loads(json.dumps(sample_mesh))
    ragged["objects"][0]["indices"] = ragged["objects"][0]["indices"][:-3] + [0, 1, 99, 2, 3]     # out-of-range triple + incomplete tail
    for scene in (sample_mesh, ragged, gen_scenes.cornell(), gen_scenes.terrain(quads=20)):
        snap = lambda objs: [(a.type, a.material, a.first_tri, a.tri_count, list(a.a), list(a.b), list(a.c)) for a in objs]
        assert _load(lib, h, scene)[0] == L.BRT_OK
        o1, m1, l1, t1 = _flat(lib, h)
        o1, m1, l1 = snap(o1), len(m1), len(l1)          # the pointers are borrowed from the ctx: copy before the next load
        data = scene_binary.pack(scene)
        hc, w, hh = C.c_int(), C.c_int(), C.c_int()
        assert lib.brt_scene_load_binary(h, data, len(data), 600, 400, C.byref(hc), C.byref(w), C.byref(hh)) == L.BRT_OK, lib.brt_last_error(h)
        o2, m2, l2, t2 = _flat(lib, h)
        assert np.array_equal(t1, t2) and o1 == snap(o2) and m1 == len(m2) and l1 == len(l2)
        assert hc.value == 1
    big = gen_scenes.terrain(quads=64)
    assert len(scene_binary.pack(big)) < 0.85 * len(json.dumps(big))                    # smaller, and no text parse
    data = scene_binary.pack(sample_mesh)
    for bad in (data[:12], b"BRTSCN02" + data[8:], data[:-40], b"BRTSCN01" + (2 ** 40).to_bytes(8, "little") + data[16:]):
        assert lib.brt_scene_load_binary(h, bad, len(bad), 600, 400, None, None, None) == L.BRT_E_PARSE
        assert lib.brt_last_error(h)
    # the Python mirror picks the loader from the magic
    import torch
    if torch.cuda.is_available():
        rt = brt.RayTracer(64, 48)
        assert rt.loadFromJSON(data) and rt.sceneInfo()["n_triangles"] == 12
