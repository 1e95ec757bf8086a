"""ORACLE — TEST INFRASTRUCTURE ONLY (see brt_oracle.cpp header).

ctypes binding of ``liboracle.so`` plus a Python restatement of the reference's host-side
logic that sits above the per-pixel loop:

* ``SceneLoader``  — js/scene-loader.js:20-284 (JSON → World/Camera, defaults, skip rules)
* ``OracleRayTracer`` — js/ray-tracer.js (constructor defaults :16-40, presets :42-77/:336-435,
  loadFromJSON :305-334, setupCamera :439-474, updateCamera :475-510, updateRenderSettings :554-566,
  updateBackground :568-585, resizeCanvas :598-614, loadCameraPreset :627-680, render :166-281)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  PARITY PINNED: tests/test_reference_pin.py compares it bit for bit with what the reference's own, unmodified
js/*.js computes when executed by baseline/minijs.py (tests/golden/reference_vectors.json, 20 cases); also held by hand-derived
float64 known-answer vectors and an independent second port (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MAT = {"lambertian": 0, "metal": 1, "dielectric": 2, "emissive": 3}
BG = {"gradient": 0, "solid": 1, "hdri": 2, "procedural_sky": 3}
AA = {"none": 0, "supersampling": 1, "stochastic": 2}
TONEMAP = {"reinhard": 0, "aces": 1, "linear": 2}
CAM_PERSPECTIVE, CAM_ORTHOGRAPHIC, CAM_OTHER = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "brt_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=2)       # keep our stdout clean (bench.py prints one JSON line)
    return _LIB_PATH


class _RenderParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("samples", C.c_int32), ("maxBounces", C.c_int32),
        ("antiAliasing", C.c_int32), ("toneMapping", C.c_int32),
        ("exposure", C.c_double), ("gamma", C.c_double),
        ("denoising", C.c_int32),
        ("denoiseStrength", C.c_double),
        ("seed", C.c_uint64),
        ("directLighting", C.c_int32),
        ("sampleBegin", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    L.orc_scene_new.restype = C.c_void_p
    L.orc_scene_free.argtypes = [C.c_void_p]
    for name in ("orc_add_sphere",):
        getattr(L, name).argtypes = [C.c_void_p, dp, C.c_double, C.c_int, dp, C.c_double]
    for name in ("orc_add_plane", "orc_add_box"):
        getattr(L, name).argtypes = [C.c_void_p, dp, dp, C.c_int, dp, C.c_double]
    L.orc_add_triangle.argtypes = [C.c_void_p, dp, dp, dp, C.c_int, dp, C.c_double]
    L.orc_add_mesh.argtypes = [C.c_void_p, dp, C.c_int, dp, C.c_int, C.c_int, dp, C.c_double]
    L.orc_mesh_triangle_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_object_count.argtypes = [C.c_void_p]
    L.orc_add_point_light.argtypes = [C.c_void_p, dp, dp, C.c_double]
    L.orc_add_directional_light.argtypes = [C.c_void_p, dp, dp, C.c_double]
    L.orc_illuminate.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.orc_set_camera.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.orc_get_camera.argtypes = [C.c_void_p, dp]
    L.orc_copy_camera.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_set_background.argtypes = [C.c_void_p, C.c_int, dp, C.c_double]
    L.orc_set_perm.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.orc_background.argtypes = [C.c_void_p, dp, dp]
    L.orc_perlin.argtypes = [C.c_void_p, dp]
    L.orc_perlin.restype = C.c_double
    L.orc_turbulence.argtypes = [C.c_void_p, dp, C.c_int]
    L.orc_turbulence.restype = C.c_double
    L.orc_tonemap.argtypes = [C.c_int, C.c_double, dp, dp]
    L.orc_gamma.argtypes = [C.c_double, dp, dp]
    L.orc_quantize.argtypes = [C.c_double]
    L.orc_schlick.argtypes = [C.c_double, C.c_double]
    L.orc_schlick.restype = C.c_double
    L.orc_refract.argtypes = [dp, dp, C.c_double, dp]
    L.orc_denoise.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float)]
    L.orc_quantize_image.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    L.orc_set_object_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, C.c_double, C.POINTER(C.c_int)]
    L.orc_texture_value.argtypes = [C.c_void_p, C.c_int, dp, dp]
    L.orc_philox_raw.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_rng_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, dp]
    L.orc_render_rect.argtypes = [C.c_void_p, C.POINTER(_RenderParams), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.POINTER(C.c_uint8), C.POINTER(C.c_float), dp]
    L.orc_render_rect.restype = C.c_longlong
    L.orc_primary_aov.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), dp, dp,
                                  C.POINTER(C.c_uint8)]
    _lib = L
    return L


def _d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


# ----------------------------------------------------------------------------- JS value semantics
def js_truthy(v) -> bool:
    """ECMAScript ToBoolean for JSON-decoded values ([] and {} are truthy, unlike Python)."""
    if v is None or v is False:
        return False
    if v is True:
        return True
    if isinstance(v, (int, float)):
        return not (v == 0 or v != v)
    if isinstance(v, str):
        return len(v) > 0
    return True


def js_num(v) -> float:
    """Numeric coercion of a JSON-decoded value as it behaves under JS arithmetic: null → 0, booleans → 0/1.
    Strings and containers are outside the documented format (docs/scene_format.md) and map to NaN."""
    if v is None:
        return 0.0
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, float)):
        return float(v)
    return math.nan


def make_perm(seed: int) -> np.ndarray:
    """noise.js:7-13 Fisher–Yates shuffle with a seeded generator standing in for Math.random."""
    rng = np.random.default_rng(seed)
    p = list(range(256))
    for i in range(255, -1, -1):
        j = int(math.floor(rng.random() * (i + 1)))
        p[i], p[j] = p[j], p[i]
    return np.asarray(p, dtype=np.int32)


# ----------------------------------------------------------------------------- scene container
class OracleScene:
    """A World + Camera pair living in liboracle (js/world.js + js/camera.js)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_scene_new())
        self.has_camera = False
        self.cam_type_str = "perspective"
        self.bg_kind = "gradient"
        self.sky_intensity = 1.0
        self.solid_color = (0.1, 0.1, 0.1)
        self.objects_log = []     # (type, material, data...) in world.objects order — lets tests compare ingest results
        self.lights_log = []

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_free(self.h)
                self.h = None
        except Exception:
            pass

    # materials are (type_name, color3, param)
    @staticmethod
    def _mat(m):
        t, col, p = m
        return MAT[t], _d3(col if col is not None else (0, 0, 0)), float(p)

    def add_sphere(self, center, radius, mat):
        t, col, p = self._mat(mat)
        self.objects_log.append(("sphere", mat, list(center), float(radius)))
        return self.L.orc_add_sphere(self.h, _d3(center), float(radius), t, col, p)

    def add_plane(self, point, normal, mat):
        t, col, p = self._mat(mat)
        self.objects_log.append(("plane", mat, list(point), list(normal)))
        return self.L.orc_add_plane(self.h, _d3(point), _d3(normal), t, col, p)

    def add_box(self, mn, mx, mat):
        t, col, p = self._mat(mat)
        self.objects_log.append(("box", mat, list(mn), list(mx)))
        return self.L.orc_add_box(self.h, _d3(mn), _d3(mx), t, col, p)

    def add_triangle(self, v0, v1, v2, mat):
        t, col, p = self._mat(mat)
        self.objects_log.append(("triangle", mat, list(v0), list(v1), list(v2)))
        return self.L.orc_add_triangle(self.h, _d3(v0), _d3(v1), _d3(v2), t, col, p)

    def add_mesh(self, vertices, indices, mat):
        t, col, p = self._mat(mat)
        v = np.ascontiguousarray(np.asarray(vertices, dtype=np.float64).reshape(-1, 3))
        i = np.ascontiguousarray(np.asarray(indices, dtype=np.float64).reshape(-1))
        self.objects_log.append(("mesh", mat, v.copy(), i.copy()))
        return self.L.orc_add_mesh(self.h, _ptr(v, C.c_double), v.shape[0], _ptr(i, C.c_double), i.shape[0], t, col, p)

    def add_point_light(self, pos, color, intensity):
        self.lights_log.append(("point", list(pos), list(color), float(intensity)))
        self.L.orc_add_point_light(self.h, _d3(pos), _d3(color), float(intensity))

    def add_directional_light(self, direction, color, intensity):
        self.lights_log.append(("directional", list(direction), list(color), float(intensity)))
        self.L.orc_add_directional_light(self.h, _d3(direction), _d3(color), float(intensity))

    def set_camera(self, look_from, look_at, vup, vfov, aspect, aperture, focus_dist, type_str="perspective"):
        code = CAM_PERSPECTIVE if type_str == "perspective" else CAM_ORTHOGRAPHIC if type_str == "orthographic" else CAM_OTHER
        self.L.orc_set_camera(self.h, _d3(look_from), _d3(look_at), _d3(vup), float(vfov), float(aspect),
                              float(aperture), float(focus_dist), code)
        self.has_camera = True
        self.cam_type_str = type_str

    def camera(self) -> dict:
        out = (C.c_double * 26)()
        self.L.orc_get_camera(self.h, out)
        a = np.array(out[:])
        names = ["origin", "lowerLeftCorner", "horizontal", "vertical", "u", "v", "w"]
        d = {n: a[3 * k:3 * k + 3].copy() for k, n in enumerate(names)}
        d.update(lensRadius=a[21], fov=a[22], aperture=a[23], focusDist=a[24], type=self.cam_type_str)
        return d

    def set_background(self, kind: str, color=(0.1, 0.1, 0.1), intensity=1.0):
        self.bg_kind = kind if kind in BG else "gradient"
        self.sky_intensity = float(intensity)
        self.solid_color = tuple(float(c) for c in color)
        self.L.orc_set_background(self.h, BG[self.bg_kind], _d3(color), float(intensity))

    TEX = {"solid": 0, "checker": 1, "noise": 2, "marble": 3, "wood": 4}

    def set_object_texture(self, obj, kind, odd=(1, 1, 1), even=(1, 1, 1), scale=1.0, perm256=None):
        """Replace object `obj`'s Lambertian / Metal by its Textured* variant (materials.js:99-126) with this texture."""
        perm = None
        if perm256 is not None:
            perm = (C.c_int * 256)(*[int(x) for x in perm256])
        rc = self.L.orc_set_object_texture(self.h, int(obj), self.TEX[kind], _d3(odd), _d3(even), float(scale), perm)
        assert rc == 0
        return rc

    def texture_value(self, obj, p):
        out = (C.c_double * 3)()
        self.L.orc_texture_value(self.h, int(obj), _d3(p), out)
        return np.array(out[:])

    def set_perm(self, perm256):
        p = np.ascontiguousarray(np.asarray(perm256, dtype=np.int32))
        assert p.shape == (256,)
        self.L.orc_set_perm(self.h, _ptr(p, C.c_int))

    def background(self, direction):
        out = (C.c_double * 3)()
        self.L.orc_background(self.h, _d3(direction), out)
        return np.array(out[:])

    def object_count(self):
        return self.L.orc_object_count(self.h)

    def mesh_triangle_count(self, obj):
        return self.L.orc_mesh_triangle_count(self.h, obj)

    def primary_aov(self, width, height):
        n = width * height
        obj = np.empty(n, np.int32); tri = np.empty(n, np.int32)
        t = np.empty(n, np.float64); nrm = np.empty((n, 3), np.float64); ff = np.empty(n, np.uint8)
        self.L.orc_primary_aov(self.h, width, height, _ptr(obj, C.c_int32), _ptr(tri, C.c_int32), _ptr(t, C.c_double),
                               _ptr(nrm, C.c_double), _ptr(ff, C.c_uint8))
        return dict(obj_id=obj.reshape(height, width), tri_id=tri.reshape(height, width), t=t.reshape(height, width),
                    normal=nrm.reshape(height, width, 3), front_face=ff.reshape(height, width))


# ----------------------------------------------------------------------------- js/scene-loader.js
class SceneLoadError(Exception):
    """Stands in for the TypeError the reference would throw (caught at ray-tracer.js:330-333 → load fails)."""


class SceneLoader:
    @staticmethod
    def _parse_vec3(arr):                                           # scene-loader.js:268-273
        if isinstance(arr, list) and len(arr) >= 3:
            return [js_num(arr[0]), js_num(arr[1]), js_num(arr[2])]
        return [0.0, 0.0, 0.0]

    @classmethod
    def _create_material(cls, m):                                   # scene-loader.js:143-173
        if not js_truthy(m) or not isinstance(m, dict) or not js_truthy(m.get("type")):
            return ("lambertian", [0.8, 0.8, 0.8], 0.0)
        if not isinstance(m["type"], str):
            raise SceneLoadError("material.type.toLowerCase is not a function")
        t = m["type"].lower()
        if t == "lambertian":
            return ("lambertian", cls._parse_vec3(m.get("color")), 0.0)
        if t == "metal":
            return ("metal", cls._parse_vec3(m.get("color")), js_num(m["roughness"]) if "roughness" in m else 0.0)
        if t == "dielectric":
            return ("dielectric", [0, 0, 0], js_num(m["ior"]) if "ior" in m else 1.5)
        if t == "emissive":
            return ("emissive", cls._parse_vec3(m.get("color")), js_num(m["intensity"]) if "intensity" in m else 1.0)
        return ("lambertian", [0.8, 0.8, 0.8], 0.0)

    @classmethod
    def _create_object(cls, scene: OracleScene, o) -> bool:         # scene-loader.js:90-137
        if o is None:
            raise SceneLoadError("cannot read properties of null (reading 'type')")
        if not isinstance(o, dict):
            return False                                               # 5 .type / "abc".type are undefined → skipped
        if not js_truthy(o.get("type")):
            return False
        mat_data = o.get("material")
        material = cls._create_material(mat_data if js_truthy(mat_data) else {"type": "lambertian", "color": [0.8, 0.8, 0.8]})
        if not isinstance(o["type"], str):
            raise SceneLoadError("type.toLowerCase is not a function")
        t = o["type"].lower()
        if t == "sphere":
            r = o.get("radius")
            radius = js_num(r) if js_truthy(r) else 1.0                # `objData.radius || 1.0` (:101)
            scene.add_sphere(cls._parse_vec3(o.get("center")), radius, material)
            return True
        if t == "plane":
            scene.add_plane(cls._parse_vec3(o.get("point")), cls._parse_vec3(o.get("normal")), material)
            return True
        if t == "box":
            scene.add_box(cls._parse_vec3(o.get("min")), cls._parse_vec3(o.get("max")), material)
            return True
        if t == "triangle":
            scene.add_triangle(cls._parse_vec3(o.get("v0")), cls._parse_vec3(o.get("v1")), cls._parse_vec3(o.get("v2")), material)
            return True
        if t == "mesh":
            if not js_truthy(o.get("vertices")) or not js_truthy(o.get("indices")):
                return False                                           # :120-123
            if not isinstance(o["vertices"], list):
                raise SceneLoadError("vertices.map is not a function")
            verts = [cls._parse_vec3(v) for v in o["vertices"]]
            idx = o["indices"]
            if not isinstance(idx, list):                              # geometry.js:199-202 → mesh with no triangles
                scene.add_mesh(np.zeros((0, 3)), np.zeros((0,)), material)
                return True
            # vertices[null] / vertices["x"] read `undefined` in JS → NaN here → (0,0,0) in the oracle's fetch
            idxf = [float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else math.nan for v in idx]
            scene.add_mesh(np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(idxf, dtype=np.float64), material)
            return True
        return False                                                   # unknown type (:133-135)

    @classmethod
    def _create_light(cls, scene: OracleScene, l) -> bool:          # scene-loader.js:179-200
        if not js_truthy(l) or not isinstance(l, dict) or not js_truthy(l.get("type")):
            return False
        color = cls._parse_vec3(l["color"] if js_truthy(l.get("color")) else [1, 1, 1])
        intensity = js_num(l["intensity"]) if "intensity" in l else 1.0
        if not isinstance(l["type"], str):
            raise SceneLoadError("type.toLowerCase is not a function")
        t = l["type"].lower()
        if t == "point":
            scene.add_point_light(cls._parse_vec3(l.get("position")), color, intensity)
            return True
        if t == "directional":
            scene.add_directional_light(cls._parse_vec3(l.get("direction")), color, intensity)
            return True
        return False

    @classmethod
    def _create_camera(cls, scene: OracleScene, cam, aspect):       # scene-loader.js:205-262
        position = cls._parse_vec3(cam["position"] if js_truthy(cam.get("position")) else [0, 0, 5])
        look_at = cls._parse_vec3(cam["lookAt"] if js_truthy(cam.get("lookAt")) else [0, 0, 0])
        up = cls._parse_vec3(cam["up"] if js_truthy(cam.get("up")) else [0, 1, 0])
        fov = js_num(cam["fov"]) if "fov" in cam else 45.0
        aperture = js_num(cam["aperture"]) if "aperture" in cam else 0.0
        d = [position[k] - look_at[k] for k in range(3)]
        dist = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        if dist < 1.0:                                                 # :213-224
            if dist > 0:
                direction = [-(d[k] / dist) for k in range(3)]
            else:
                direction = [-0.0, -0.0, -0.0]
            look_at = [position[k] + direction[k] * 100 for k in range(3)]
        if "focusDist" in cam:
            focus_dist = js_num(cam["focusDist"])
        else:                                                          # :228-233
            d = [position[k] - look_at[k] for k in range(3)]
            focus_dist = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        type_str = cam["type"] if js_truthy(cam.get("type")) else "perspective"   # not lower-cased (:235)
        final_aspect = js_num(cam["aspect"]) if js_truthy(cam.get("aspect")) else aspect
        scene.set_camera(position, look_at, up, fov, final_aspect, aperture, focus_dist, type_str)

    @classmethod
    def load_from_json(cls, data, width, height):                   # scene-loader.js:20-84
        """Returns (scene, has_camera, new_dimensions|None)."""
        if not isinstance(data, dict):
            raise SceneLoadError("scene root is not an object")
        new_dims = None
        cam = data.get("camera")
        if js_truthy(cam) and isinstance(cam, dict) and js_truthy(cam.get("resolution")):
            res = cam["resolution"]
            if not isinstance(res, list) or len(res) < 2:
                raise SceneLoadError("camera.resolution must be [w, h]")
            new_dims = (int(js_num(res[0])), int(js_num(res[1])))
            width, height = new_dims
        scene = OracleScene()
        bg = data.get("background")
        if js_truthy(bg) and isinstance(bg, dict):
            kind = bg.get("type")
            kind = kind if isinstance(kind, str) and kind in BG else "gradient"
            # D1: solid/hdri implement the INTENDED behaviour (ray-tracer.js:573-576), not the
            # loader's mis-bound factory that yields NaN → black.
            color = cls._parse_vec3(bg["color"] if js_truthy(bg.get("color")) else [0.1, 0.1, 0.1])
            intensity = js_num(bg["intensity"]) if "intensity" in bg else 1.0
            scene.set_background(kind, color, intensity)
        objs = data.get("objects")
        if js_truthy(objs) and isinstance(objs, list):
            for o in objs:
                cls._create_object(scene, o)
        lights = data.get("lights")
        if js_truthy(lights) and isinstance(lights, list):
            for l in lights:
                cls._create_light(scene, l)
        has_camera = False
        if js_truthy(cam):
            if not isinstance(cam, dict):
                cam = {}
            cls._create_camera(scene, cam, width / height)
            has_camera = True
        return scene, has_camera, new_dims


# ----------------------------------------------------------------------------- js/ray-tracer.js
class OracleRayTracer:
    """Mirror of `class RayTracer` (ray-tracer.js:15-681) driving the float64 oracle."""

    def __init__(self, width=600, height=400, seed=1, threads=1, perm_seed=0):
        self.width, self.height = int(width), int(height)
        self.maxBounces = 5                                            # :23-30
        self.samples = 4
        self.gamma = 2.2
        self.exposure = 1.0
        self.toneMapping = "reinhard"
        self.antiAliasing = "supersampling"
        self.denoising = False
        self.denoiseStrength = 0.5
        self.seed = int(seed)
        self.threads = int(threads)
        self.directLighting = False
        self.perm = make_perm(perm_seed)
        self.scene = OracleScene()
        self.scene.set_perm(self.perm)
        self.rays = 0
        self.floatData = None
        self.linear = None
        self.setupDefaultScene()

    def setCloudPermutation(self, perm256):
        """world.cloudNoise.p (noise.js:7-17): random per World in the reference, an explicit input here."""
        self.perm = np.ascontiguousarray(np.asarray(perm256, dtype=np.uint8).reshape(256))
        self.scene.set_perm(self.perm)

    # -- presets ------------------------------------------------------------------------------------
    def _new_world(self):
        self.scene = OracleScene()
        self.scene.set_perm(self.perm)

    def setupDefaultScene(self):                                       # :42-77
        s = self.scene
        s.add_plane([0, -0.5, 0], [0, 1, 0], ("lambertian", [0.5, 0.5, 0.5], 0))
        s.add_sphere([0, 0, -1], 0.5, ("lambertian", [0.7, 0.3, 0.3], 0))
        s.add_sphere([-1, 0, -1], 0.5, ("dielectric", None, 1.5))
        s.add_sphere([1, 0, -1], 0.5, ("metal", [0.8, 0.8, 0.9], 0.1))
        s.add_sphere([0, 1.5, -1], 0.3, ("emissive", [1, 1, 1], 5))
        s.add_point_light([2, 2, 0], [1, 1, 1], 10)
        s.add_directional_light([-1, -1, -1], [1, 0.9, 0.8], 2)
        s.set_camera([3, 2, 2], [0, 0, -1], [0, 1, 0], 45, self.width / self.height, 0.0, 10.0)

    def setupGlassScene(self):                                         # :336-364
        s = self.scene
        glass, glass2 = ("dielectric", None, 1.5), ("dielectric", None, 2.4)
        s.add_plane([0, -0.5, 0], [0, 1, 0], ("lambertian", [0.8, 0.8, 0.0], 0))
        s.add_sphere([0, 0, -1], 0.5, glass)
        s.add_sphere([0, 0, -1], -0.45, glass)
        s.add_sphere([-1, 0, -1], 0.5, glass2)
        s.add_sphere([1, 0, -1], 0.5, glass)
        s.add_sphere([0, 4, -1], 1, ("emissive", [1, 1, 1], 8))
        s.add_point_light([0, 4, -1], [1, 1, 1], 20)
        s.set_camera([3, 2, 2], [0, 0, -1], [0, 1, 0], 45, self.width / self.height, 0.02, math.sqrt(3 * 3 + 2 * 2 + 3 * 3))

    def setupMetalScene(self):                                         # :366-398
        s = self.scene
        metal2 = ("metal", [0.8, 0.6, 0.2], 0.1)
        s.add_plane([0, -0.5, 0], [0, 1, 0], ("lambertian", [0.5, 0.5, 0.5], 0))
        s.add_sphere([0, 0, -1], 0.5, ("metal", [0.8, 0.8, 0.9], 0.0))
        s.add_sphere([-1, 0, -1], 0.5, metal2)
        s.add_sphere([1, 0, -1], 0.5, ("metal", [0.7, 0.6, 0.5], 0.3))
        s.add_box([-0.3, -0.5, -2], [0.3, 0.3, -1.4], metal2)
        s.add_sphere([2, 3, 0], 0.5, ("emissive", [1, 0.8, 0.6], 10))
        s.add_point_light([2, 3, 0], [1, 0.8, 0.6], 15)
        s.add_directional_light([-1, -2, -1], [0.3, 0.4, 0.6], 1)
        s.set_camera([4, 2, 3], [0, 0, -1], [0, 1, 0], 45, self.width / self.height, 0.0, 10.0)

    def setupCornellBox(self):                                         # :400-435
        s = self.scene
        red = ("lambertian", [0.65, 0.05, 0.05], 0)
        white = ("lambertian", [0.73, 0.73, 0.73], 0)
        green = ("lambertian", [0.12, 0.45, 0.15], 0)
        s.add_plane([0, 0, -5], [0, 0, 1], white)
        s.add_plane([0, -2.5, 0], [0, 1, 0], white)
        s.add_plane([0, 2.5, 0], [0, -1, 0], white)
        s.add_plane([-2.5, 0, 0], [1, 0, 0], red)
        s.add_plane([2.5, 0, 0], [-1, 0, 0], green)
        s.add_box([-1, -2.5, -3.5], [-0.2, -1, -2.7], white)
        s.add_box([0.2, -2.5, -4], [1.2, -0.5, -3], white)
        s.add_sphere([-0.6, -1.8, -2.2], 0.7, ("dielectric", None, 1.5))
        s.add_sphere([0.7, -1.8, -3.5], 0.7, ("metal", [0.8, 0.85, 0.88], 0.0))
        s.add_box([-0.5, 2.45, -3.5], [0.5, 2.49, -2.5], ("emissive", [1, 1, 1], 15))
        s.set_background("solid", [0, 0, 0], s.sky_intensity)
        s.set_camera([0, 0, 2], [0, 0, -1], [0, 1, 0], 40, self.width / self.height, 0.0, 10.0)

    def loadPreset(self, name):                                        # :282-299
        self._new_world()
        {"glass": self.setupGlassScene, "metal": self.setupMetalScene,
         "cornell": self.setupCornellBox}.get(name, self.setupDefaultScene)()

    # -- JSON ---------------------------------------------------------------------------------------
    def loadFromJSON(self, data) -> bool:                              # :305-334
        try:
            scene, has_cam, new_dims = SceneLoader.load_from_json(data, self.width, self.height)
        except SceneLoadError:
            return False
        scene.set_perm(self.perm)
        if not has_cam and self.scene.has_camera:                      # :315-317 keep the previous camera
            scene.L.orc_copy_camera(scene.h, self.scene.h)
            scene.has_camera, scene.cam_type_str = True, self.scene.cam_type_str
        self.scene = scene
        if new_dims:
            self.resizeCanvas(*new_dims)
        return True

    def resizeCanvas(self, width, height):                             # :598-614
        self.width, self.height = int(width), int(height)
        if self.scene.has_camera:
            self.setupCamera()

    def setupCamera(self):                                             # :439-474
        if not self.scene.has_camera:
            return
        c = self.scene.camera()
        look_from = c["origin"]
        look_at = look_from - c["w"] * c["focusDist"]
        vup = c["v"]
        fov = c["fov"] if js_truthy(float(c["fov"])) else 45
        aperture = c["aperture"] if js_truthy(float(c["aperture"])) else 0.0
        focus = c["focusDist"] if js_truthy(float(c["focusDist"])) else 10.0
        type_str = c["type"] if js_truthy(c["type"]) else "perspective"
        self.scene.set_camera(look_from, look_at, vup, fov, self.width / self.height, aperture, focus, type_str)

    def updateCamera(self, params: dict):                              # :475-510
        look_from, look_at, vup = np.array([3., 2, 2]), np.array([0., 0, -1]), np.array([0., 1, 0])
        c = self.scene.camera() if self.scene.has_camera else None
        if c is not None:
            look_from = c["origin"]
            fd = c["focusDist"] if js_truthy(float(c["focusDist"])) else 10.0
            look_at = c["origin"] - c["w"] * fd
            vup = c["v"]
        if js_truthy(params.get("position")):
            look_from = np.array(params["position"][:3], dtype=np.float64)
        if js_truthy(params.get("lookAt")):
            look_at = np.array(params["lookAt"][:3], dtype=np.float64)
        if js_truthy(params.get("up")):
            vup = np.array(params["up"][:3], dtype=np.float64)

        def pick(key, cur_key, default):
            v = params.get(key)
            if js_truthy(v):
                return v
            cur = c[cur_key] if c is not None else None
            if isinstance(cur, np.floating):
                cur = float(cur)
            return cur if js_truthy(cur) else default

        self.scene.set_camera(look_from, look_at, vup, pick("fov", "fov", 45), self.width / self.height,
                              pick("aperture", "aperture", 0.0), pick("focusDist", "focusDist", 10.0),
                              pick("type", "type", "perspective"))

    def setCameraPosition(self, lookFrom=None, lookAt=None, vup=None):  # :515-535
        if not self.scene.has_camera:
            self.scene.set_camera(lookFrom if lookFrom is not None else (3, 2, 2), lookAt if lookAt is not None else (0, 0, -1),
                                  vup if vup is not None else (0, 1, 0), 45, self.width / self.height, 0.0, 10.0, "perspective")
        else:
            self.updateCamera({"position": None if lookFrom is None else list(lookFrom), "lookAt": None if lookAt is None else list(lookAt),
                               "up": None if vup is None else list(vup)})

    def getCameraPosition(self):                                       # :540-552
        if not self.scene.has_camera:
            return None
        c = self.scene.camera()
        return dict(position=c["origin"], lookAt=c["origin"] - c["w"] * c["focusDist"], up=c["v"], fov=c["fov"], aperture=c["aperture"],
                    focusDist=c["focusDist"], type=c["type"])

    def loadCameraPreset(self, name) -> bool:                          # :627-680
        presets = {
            "default": dict(position=[3, 2, 2], lookAt=[0, 0, -1], up=[0, 1, 0], fov=45, aperture=0.0, focusDist=10.0),
            "close-up": dict(position=[1, 1, 1], lookAt=[0, 0, -1], up=[0, 1, 0], fov=60, aperture=0.02, focusDist=2.0),
            "wide-angle": dict(position=[5, 3, 5], lookAt=[0, 0, 0], up=[0, 1, 0], fov=80, aperture=0.0, focusDist=15.0),
            "top-down": dict(position=[0, 5, 0], lookAt=[0, 0, -1], up=[0, 0, -1], fov=45, aperture=0.0, focusDist=5.0),
            "side-view": dict(position=[5, 0, 0], lookAt=[0, 0, -1], up=[0, 1, 0], fov=45, aperture=0.0, focusDist=5.0),
        }
        if name not in presets:
            return False
        self.updateCamera(presets[name])
        return True

    def updateRenderSettings(self, p: dict):                           # :554-566 (`||` defaults: 0 ⇒ default)
        def orr(k, d):
            v = p.get(k)
            return v if js_truthy(v) else d
        self.maxBounces = orr("maxBounces", 5)
        self.samples = orr("samples", 4)
        self.gamma = orr("gamma", 2.2)
        self.exposure = orr("exposure", 1.0)
        self.toneMapping = orr("toneMapping", "reinhard")
        self.antiAliasing = orr("antiAliasing", "supersampling")
        self.denoising = orr("denoising", False)
        self.denoiseStrength = orr("denoiseStrength", 0.5)

    def updateBackground(self, kind, intensity=1.0):                   # :568-585
        if kind == "solid":
            self.scene.set_background("solid", [0.1, 0.1, 0.1], intensity)
        elif kind in ("hdri", "procedural_sky"):
            self.scene.set_background(kind, [0.1, 0.1, 0.1], intensity)
        else:
            self.scene.set_background("gradient", [0.1, 0.1, 0.1], intensity)

    # -- render -------------------------------------------------------------------------------------
    def _params(self):
        rp = _RenderParams()
        rp.width, rp.height = self.width, self.height
        rp.samples, rp.maxBounces = int(self.samples), int(self.maxBounces)
        rp.antiAliasing = AA.get(self.antiAliasing, 3)      # any other string: pixel-centre arm (:142-148), still `samples` samples (:201)
        rp.toneMapping = TONEMAP.get(self.toneMapping, 0)   # default: reinhard (:157-159)
        rp.exposure, rp.gamma = float(self.exposure), float(self.gamma)
        rp.denoising = 1 if self.denoising else 0
        rp.denoiseStrength = float(self.denoiseStrength)
        rp.seed = self.seed
        rp.directLighting = 1 if self.directLighting else 0
        rp.sampleBegin = int(getattr(self, "sampleBegin", 0))
        return rp

    def render(self, onProgress=None, rect=None, reuse=False):         # :166-281
        """Returns RGBA8 (H, W, 4), row 0 = top.  `rect=(x0, y0, x1, y1)` renders a crop (rest stays 0).
        `reuse=True` keeps the output arrays between calls (timing loops: no per-call allocation of full frames)."""
        W, H = self.width, self.height
        if reuse and getattr(self, "_bufs", None) is not None and self._bufs[0].shape == (H, W, 4):
            rgba, fdat, lin = self._bufs
        else:
            rgba = np.zeros((H, W, 4), np.uint8)
            fdat = np.zeros((H, W, 4), np.float32)
            lin = np.zeros((H, W, 4), np.float64)
            self._bufs = (rgba, fdat, lin) if reuse else None
        rp = self._params()
        x0, y0, x1, y1 = rect if rect else (0, 0, W, H)
        self.rays = self.scene.L.orc_render_rect(self.scene.h, C.byref(rp), x0, y0, x1, y1, self.threads,
                                                 _ptr(rgba, C.c_uint8), _ptr(fdat, C.c_float), _ptr(lin, C.c_double))
        if self.denoising:                                             # :266-276
            out = np.empty_like(fdat)
            self.scene.L.orc_denoise(_ptr(fdat, C.c_float), W, H, float(self.denoiseStrength), _ptr(out, C.c_float))
            self.scene.L.orc_quantize_image(_ptr(out, C.c_float), W, H, _ptr(rgba, C.c_uint8))
            self.denoised = out
        self.floatData, self.linear = fdat, lin
        if onProgress:
            onProgress(1.0)
        return rgba

    def primary_aov(self):
        return self.scene.primary_aov(self.width, self.height)
