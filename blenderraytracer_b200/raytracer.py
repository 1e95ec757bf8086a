"""``RayTracer`` — host-side mirror of the reference's ``class RayTracer`` (js/ray-tracer.js:15-681) whose
``render()`` runs on the GPU through libbrt.  Method names, argument meaning, the ``||``-style defaults and the
error behaviour follow the reference so a caller of the JS class finds the same surface:

    rt = RayTracer(600, 400)            # new RayTracer(canvas)           ray-tracer.js:16-40
    rt.loadFromJSON(json_dict)           # RayTracer.loadFromJSON          :305-334 (+ scene-loader.js)
    rt.updateCamera({...}); rt.updateRenderSettings({...}); rt.updateBackground('gradient', 1.0)
    rgba = rt.render(onProgress)         # RayTracer.render                :166-281  -> (H, W, 4) uint8, row 0 = top

The canvas / DOM side (putImageData, per-row preview blits) stays with the caller; ``render`` returns the bytes that
the reference writes into ``imageData.data``.
"""
from __future__ import annotations

import ctypes as C
import json
import math

import numpy as np

from . import _lib as L
from .scene import (Box, Dielectric, DirectionalLight, Emissive, Lambertian, Metal, Plane, PointLight, Sphere, World)


def _truthy(v) -> bool:
    """ECMAScript ToBoolean for the values that reach the reference's `a || b` expressions."""
    if v is None or v is False:
        return False
    if v is True:
        return True
    if isinstance(v, (int, float, np.integer, np.floating)):
        return not (v == 0 or v != v)
    if isinstance(v, str):
        return len(v) > 0
    return True


_CAM_CODE = {"perspective": L.CAM_PERSPECTIVE, "orthographic": L.CAM_ORTHOGRAPHIC}
_CAM_NAME = {L.CAM_PERSPECTIVE: "perspective", L.CAM_ORTHOGRAPHIC: "orthographic"}


def make_perm(seed: int) -> np.ndarray:
    """PerlinNoise constructor shuffle (noise.js:7-13) with a seeded generator in place of Math.random."""
    rng = np.random.default_rng(seed)
    p = list(range(256))
    for i in range(255, -1, -1):
        j = int(math.floor(rng.random() * (i + 1)))
        p[i], p[j] = p[j], p[i]
    return np.asarray(p, dtype=np.uint8)


class RayTracer:
    def __init__(self, width=600, height=400, device=0, seed=1, perm_seed=0, stream=None, devices=None):
        """`devices=[0, 1, ...]`: one context spanning several GPUs of this process (brt_create_multi) — render() is still
        one call, every batch's samples are split over the devices and exchanged by the fused peer kernel."""
        self._L = L.load()
        h = C.c_void_p()
        if devices is not None and len(devices) > 0:
            ids = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self._L.brt_create_multi(C.byref(h), ids, len(devices))
        else:
            rc = self._L.brt_create(C.byref(h), int(device))
        if rc != L.BRT_OK:
            raise L.BrtError(rc, "brt_create failed — libbrt needs a CUDA device (sm_100a); there is no CPU fallback "
                                 "(device=-1 gives a host-only context for scene/camera logic, which cannot render)")
        self._ctx = h
        if stream is not None:
            L.check(self._ctx, self._L.brt_set_stream(self._ctx, C.c_void_p(int(stream) or 1)))
        self.width, self.height = int(width), int(height)
        # ray-tracer.js:23-30
        self.maxBounces = 5
        self.samples = 4
        self.gamma = 2.2
        self.exposure = 1.0
        self.toneMapping = "reinhard"
        self.antiAliasing = "supersampling"
        self.denoising = False
        self.denoiseStrength = 0.5
        # engine knobs (not in the reference)
        self.seed = int(seed)
        self.directLighting = False
        self.sampler = "fast"
        self.integrator = "auto"
        self.accel = "auto"
        self.sppBatch = 0
        self.countTests = False
        self.onResizeCallback = None
        self.floatData = None
        self.linearMean = None
        self._cam_type_str = "perspective"
        self._bg = ("gradient", (0.1, 0.1, 0.1), 1.0)
        self._perm = make_perm(perm_seed)
        self._push_background()
        self.world = World()
        self.setupDefaultScene()

    # -- lifecycle ------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.brt_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- scene ----------------------------------------------------------------------------------------------
    def _set_world(self, world: World):
        self.world = world
        desc, keep = world.flatten()
        L.check(self._ctx, self._L.brt_scene_set_flat(self._ctx, C.byref(desc)))
        del keep

    def _new_world(self):
        """`this.world = new World()` (ray-tracer.js:283): gradient background, skyIntensity 1 (world.js:12-13)."""
        self._bg = ("gradient", (0.1, 0.1, 0.1), 1.0)
        self._push_background()
        return World()

    def _set_camera_raw(self, look_from, look_at, vup, vfov, aspect, aperture, focus_dist, type_str="perspective"):
        c = L.brt_camera()
        c.look_from, c.look_at, c.vup = L.d3(*map(float, look_from)), L.d3(*map(float, look_at)), L.d3(*map(float, vup))
        c.vfov, c.aspect, c.aperture, c.focus_dist = float(vfov), float(aspect), float(aperture), float(focus_dist)
        c.type = _CAM_CODE.get(type_str, L.CAM_OTHER)
        c.use_derived = 0
        L.check(self._ctx, self._L.brt_set_camera(self._ctx, C.byref(c)))
        self._cam_type_str = type_str

    @property
    def camera(self):
        """Camera.debugReport-style view (camera.js:56-78) of the live camera, or None."""
        c = L.brt_camera()
        if self._L.brt_get_camera(self._ctx, C.byref(c)) != L.BRT_OK:
            return None
        g = lambda a: np.array(a[:], dtype=np.float64)
        return dict(origin=g(c.origin), lowerLeftCorner=g(c.lower_left_corner), horizontal=g(c.horizontal), vertical=g(c.vertical),
                    u=g(c.u), v=g(c.v), w=g(c.w), lensRadius=c.lens_radius, fov=c.vfov, aperture=c.aperture,
                    focusDist=c.focus_dist, aspect=c.aspect,
                    type=self._cam_type_str if c.type == L.CAM_OTHER else _CAM_NAME[c.type])

    def setupDefaultScene(self):                                       # ray-tracer.js:42-77
        w = self.world
        w.add(Plane((0, -0.5, 0), (0, 1, 0), Lambertian((0.5, 0.5, 0.5))))
        w.add(Sphere((0, 0, -1), 0.5, Lambertian((0.7, 0.3, 0.3))))
        w.add(Sphere((-1, 0, -1), 0.5, Dielectric(1.5)))
        w.add(Sphere((1, 0, -1), 0.5, Metal((0.8, 0.8, 0.9), 0.1)))
        w.add(Sphere((0, 1.5, -1), 0.3, Emissive((1, 1, 1), 5)))
        w.addLight(PointLight((2, 2, 0), (1, 1, 1), 10))
        w.addLight(DirectionalLight((-1, -1, -1), (1, 0.9, 0.8), 2))
        self._set_world(w)
        self._set_camera_raw((3, 2, 2), (0, 0, -1), (0, 1, 0), 45, self.width / self.height, 0.0, 10.0)

    def setupGlassScene(self):                                         # ray-tracer.js:336-364
        w = self.world
        glass, glass2 = Dielectric(1.5), Dielectric(2.4)
        w.add(Plane((0, -0.5, 0), (0, 1, 0), Lambertian((0.8, 0.8, 0.0))))
        w.add(Sphere((0, 0, -1), 0.5, glass))
        w.add(Sphere((0, 0, -1), -0.45, glass))
        w.add(Sphere((-1, 0, -1), 0.5, glass2))
        w.add(Sphere((1, 0, -1), 0.5, glass))
        w.add(Sphere((0, 4, -1), 1, Emissive((1, 1, 1), 8)))
        w.addLight(PointLight((0, 4, -1), (1, 1, 1), 20))
        self._set_world(w)
        self._set_camera_raw((3, 2, 2), (0, 0, -1), (0, 1, 0), 45, self.width / self.height, 0.02, math.sqrt(3 * 3 + 2 * 2 + 3 * 3))

    def setupMetalScene(self):                                         # ray-tracer.js:366-398
        w = self.world
        metal2 = Metal((0.8, 0.6, 0.2), 0.1)
        w.add(Plane((0, -0.5, 0), (0, 1, 0), Lambertian((0.5, 0.5, 0.5))))
        w.add(Sphere((0, 0, -1), 0.5, Metal((0.8, 0.8, 0.9), 0.0)))
        w.add(Sphere((-1, 0, -1), 0.5, metal2))
        w.add(Sphere((1, 0, -1), 0.5, Metal((0.7, 0.6, 0.5), 0.3)))
        w.add(Box((-0.3, -0.5, -2), (0.3, 0.3, -1.4), metal2))
        w.add(Sphere((2, 3, 0), 0.5, Emissive((1, 0.8, 0.6), 10)))
        w.addLight(PointLight((2, 3, 0), (1, 0.8, 0.6), 15))
        w.addLight(DirectionalLight((-1, -2, -1), (0.3, 0.4, 0.6), 1))
        self._set_world(w)
        self._set_camera_raw((4, 2, 3), (0, 0, -1), (0, 1, 0), 45, self.width / self.height, 0.0, 10.0)

    def setupCornellBox(self):                                         # ray-tracer.js:400-435
        w = self.world
        red, white, green = Lambertian((0.65, 0.05, 0.05)), Lambertian((0.73, 0.73, 0.73)), Lambertian((0.12, 0.45, 0.15))
        w.add(Plane((0, 0, -5), (0, 0, 1), white))
        w.add(Plane((0, -2.5, 0), (0, 1, 0), white))
        w.add(Plane((0, 2.5, 0), (0, -1, 0), white))
        w.add(Plane((-2.5, 0, 0), (1, 0, 0), red))
        w.add(Plane((2.5, 0, 0), (-1, 0, 0), green))
        w.add(Box((-1, -2.5, -3.5), (-0.2, -1, -2.7), white))
        w.add(Box((0.2, -2.5, -4), (1.2, -0.5, -3), white))
        w.add(Sphere((-0.6, -1.8, -2.2), 0.7, Dielectric(1.5)))
        w.add(Sphere((0.7, -1.8, -3.5), 0.7, Metal((0.8, 0.85, 0.88), 0.0)))
        w.add(Box((-0.5, 2.45, -3.5), (0.5, 2.49, -2.5), Emissive((1, 1, 1), 15)))
        self._set_world(w)
        self._bg = ("solid", (0.0, 0.0, 0.0), self._bg[2])             # :424
        self._push_background()
        self._set_camera_raw((0, 0, 2), (0, 0, -1), (0, 1, 0), 40, self.width / self.height, 0.0, 10.0)

    def loadPreset(self, presetName):                                  # ray-tracer.js:282-299
        self.world = self._new_world()
        {"glass": self.setupGlassScene, "metal": self.setupMetalScene, "cornell": self.setupCornellBox}.get(
            presetName, self.setupDefaultScene)()

    def refreshScene(self):                                            # ray-tracer.js:587-591
        self.world = self._new_world()
        self.setupDefaultScene()

    def loadFromJSON(self, jsonData) -> bool:                          # ray-tracer.js:305-334
        """`jsonData`: the parsed scene (dict) as the reference receives it, the JSON text itself (str / bytes), or a
        BRTSCN01 binary container (bytes)."""
        binary = False
        if isinstance(jsonData, (bytes, bytearray)):
            text = bytes(jsonData)
            if text[:8] == b"BRTSCN01":                                # binary container (tools/scene_binary.py)
                binary = True
        elif isinstance(jsonData, str):
            text = jsonData.encode("utf-8")
        else:
            try:
                text = json.dumps(jsonData).encode("utf-8")
            except (TypeError, ValueError):
                return False
        has_cam, w, h = C.c_int(0), C.c_int(0), C.c_int(0)
        load = self._L.brt_scene_load_binary if binary else self._L.brt_scene_load_json
        rc = load(self._ctx, text, len(text), self.width, self.height, C.byref(has_cam), C.byref(w), C.byref(h))
        if rc != L.BRT_OK:                                             # the reference logs and returns false (:330-333)
            self.lastError = self._L.brt_last_error(self._ctx).decode("utf-8", "replace")
            return False
        self.world = None                                              # the live world now exists only inside libbrt
        kind, col, inten = C.c_int(), (C.c_double * 3)(), C.c_double()
        self._L.brt_get_background(self._ctx, C.byref(kind), col, C.byref(inten))
        self._bg = ({v: k for k, v in L.BG.items()}[kind.value], tuple(col[:]), inten.value)
        if has_cam.value:
            c = L.brt_camera()
            self._L.brt_get_camera(self._ctx, C.byref(c))
            if c.type != L.CAM_OTHER:
                self._cam_type_str = _CAM_NAME[c.type]
            else:
                cam = jsonData.get("camera") if isinstance(jsonData, dict) else None
                self._cam_type_str = str(cam.get("type")) if isinstance(cam, dict) else "other"
        if w.value and h.value:                                        # :320-326
            self.resizeCanvas(w.value, h.value)
            if callable(self.onResizeCallback):
                self.onResizeCallback(w.value, h.value)
        return True

    # -- camera ---------------------------------------------------------------------------------------------
    def setupCamera(self):                                             # ray-tracer.js:439-474
        cam = self.camera
        if cam is None:
            return
        look_from = cam["origin"]
        look_at = look_from - cam["w"] * cam["focusDist"]
        self._set_camera_raw(look_from, look_at, cam["v"],
                             cam["fov"] if _truthy(cam["fov"]) else 45,
                             self.width / self.height,
                             cam["aperture"] if _truthy(cam["aperture"]) else 0.0,
                             cam["focusDist"] if _truthy(cam["focusDist"]) else 10.0,
                             cam["type"] if _truthy(cam["type"]) else "perspective")

    def updateCamera(self, params: dict):                              # ray-tracer.js:475-510
        look_from, look_at, vup = np.array([3.0, 2, 2]), np.array([0.0, 0, -1]), np.array([0.0, 1, 0])
        cam = self.camera
        if cam is not None:
            look_from = cam["origin"]
            look_at = cam["origin"] - cam["w"] * (cam["focusDist"] if _truthy(cam["focusDist"]) else 10.0)
            vup = cam["v"]
        if _truthy(params.get("position")):
            look_from = np.array(params["position"][:3], dtype=np.float64)
        if _truthy(params.get("lookAt")):
            look_at = np.array(params["lookAt"][:3], dtype=np.float64)
        if _truthy(params.get("up")):
            vup = np.array(params["up"][:3], dtype=np.float64)

        def pick(key, cam_key, default):
            v = params.get(key)
            if _truthy(v):
                return v
            cur = cam[cam_key] if cam is not None else None
            return cur if _truthy(cur) else default

        self._set_camera_raw(look_from, look_at, vup, pick("fov", "fov", 45), self.width / self.height,
                             pick("aperture", "aperture", 0.0), pick("focusDist", "focusDist", 10.0),
                             pick("type", "type", "perspective"))

    def setCameraPosition(self, lookFrom=None, lookAt=None, vup=None):  # ray-tracer.js:515-535
        if self.camera is None:
            self._set_camera_raw(lookFrom if lookFrom is not None else (3, 2, 2), lookAt if lookAt is not None else (0, 0, -1),
                                 vup if vup is not None else (0, 1, 0), 45, self.width / self.height, 0.0, 10.0)
        else:
            self.updateCamera({"position": list(lookFrom) if lookFrom is not None else None,
                               "lookAt": list(lookAt) if lookAt is not None else None,
                               "up": list(vup) if vup is not None else None})

    def getCameraPosition(self):                                       # ray-tracer.js:540-552
        cam = self.camera
        if cam is None:
            return None
        return dict(position=cam["origin"], lookAt=cam["origin"] - cam["w"] * cam["focusDist"], up=cam["v"], fov=cam["fov"],
                    aperture=cam["aperture"], focusDist=cam["focusDist"], type=cam["type"])

    def loadCameraPreset(self, presetName) -> bool:                    # ray-tracer.js:627-680
        presets = {
            "default": dict(position=[3, 2, 2], lookAt=[0, 0, -1], up=[0, 1, 0], fov=45, aperture=0.0, focusDist=10.0),
            "close-up": dict(position=[1, 1, 1], lookAt=[0, 0, -1], up=[0, 1, 0], fov=60, aperture=0.02, focusDist=2.0),
            "wide-angle": dict(position=[5, 3, 5], lookAt=[0, 0, 0], up=[0, 1, 0], fov=80, aperture=0.0, focusDist=15.0),
            "top-down": dict(position=[0, 5, 0], lookAt=[0, 0, -1], up=[0, 0, -1], fov=45, aperture=0.0, focusDist=5.0),
            "side-view": dict(position=[5, 0, 0], lookAt=[0, 0, -1], up=[0, 1, 0], fov=45, aperture=0.0, focusDist=5.0),
        }
        if presetName not in presets:
            return False
        self.updateCamera(presets[presetName])
        return True

    # -- settings -------------------------------------------------------------------------------------------
    def updateRenderSettings(self, params: dict):                      # ray-tracer.js:554-566 (`||`: 0 ⇒ default)
        def orr(k, d):
            v = params.get(k)
            return v if _truthy(v) else d
        self.maxBounces = orr("maxBounces", 5)
        self.samples = orr("samples", 4)
        self.gamma = orr("gamma", 2.2)
        self.exposure = orr("exposure", 1.0)
        self.toneMapping = orr("toneMapping", "reinhard")
        self.antiAliasing = orr("antiAliasing", "supersampling")
        self.denoising = orr("denoising", False)
        self.denoiseStrength = orr("denoiseStrength", 0.5)

    def updateBackground(self, type, intensity=1.0):                   # ray-tracer.js:568-585
        if type == "solid":
            self._bg = ("solid", (0.1, 0.1, 0.1), float(intensity))
        elif type in ("hdri", "procedural_sky"):
            self._bg = (type, self._bg[1], float(intensity))
        else:
            self._bg = ("gradient", self._bg[1], float(intensity))
        self._push_background()

    def setCloudPermutation(self, perm256):
        """world.cloudNoise.p (noise.js:7-17): random per World in the reference, an explicit input here."""
        self._perm = np.ascontiguousarray(np.asarray(perm256, dtype=np.uint8).reshape(256))
        self._push_background()

    def _push_background(self):
        kind, col, inten = self._bg
        L.check(self._ctx, self._L.brt_set_background(self._ctx, L.BG[kind], L.d3(*col), float(inten),
                                                      self._perm.ctypes.data_as(C.POINTER(C.c_uint8))))

    def resizeCanvas(self, width, height):                             # ray-tracer.js:598-614
        self.width, self.height = int(width), int(height)
        if self.camera is not None:
            self.setupCamera()

    def onResize(self, callback):                                      # ray-tracer.js:620-622
        self.onResizeCallback = callback

    # -- render ---------------------------------------------------------------------------------------------
    def _push_params(self):
        p = L.brt_render_params()
        p.width, p.height = self.width, self.height
        p.spp, p.max_depth = int(self.samples), int(self.maxBounces)
        # any other string takes the pixel-centre arm (:142-148) but `sampleCount` stays this.samples (:201)
        p.aa_mode = L.AA.get(self.antiAliasing, L.AA_CENTER)
        p.tonemap = L.TONEMAP.get(self.toneMapping, 0)      # default arm = reinhard (:157-159)
        p.exposure, p.gamma = float(self.exposure), float(self.gamma)
        p.denoise, p.denoise_strength = 1 if self.denoising else 0, float(self.denoiseStrength)
        p.seed = self.seed
        p.direct_lighting = 1 if self.directLighting else 0
        p.sampler, p.integrator, p.accel = L.SAMPLER[self.sampler], L.INTEGRATOR[self.integrator], L.ACCEL[self.accel]
        p.spp_batch, p.count_tests = int(self.sppBatch), 1 if self.countTests else 0
        p.refill_threshold = int(getattr(self, "refillThreshold", 0))
        p.paths_in_flight = int(getattr(self, "pathsInFlight", 0))
        p.preview = 1 if getattr(self, "preview", False) else 0
        p.bvh_width = int(getattr(self, "bvhWidth", 0))           # 0 = auto, 2 = binary LBVH, 4 / 8 = wide collapse (same images)
        L.check(self._ctx, self._L.brt_set_render_params(self._ctx, C.byref(p)))
        return p

    def render(self, onProgress=None, want_float=True, want_linear=False):   # ray-tracer.js:166-281
        """Blocking GPU render.  Returns imageData.data as (H, W, 4) uint8 (row 0 = top, alpha 255); keeps the
        reference's `floatData` in ``self.floatData`` and (optionally) the linear per-pixel mean in ``self.linearMean``.
        Raises BrtError(BRT_E_CANCELLED) if ``cancel()`` was called (window.renderCancelled)."""
        self._push_params()
        W, H = self.width, self.height
        rgba = np.empty((H, W, 4), np.uint8)
        fdat = np.empty((H, W, 4), np.float32) if want_float else None
        lin = np.empty((H, W, 4), np.float32) if want_linear else None
        # with self.preview = True, `rgba` already holds the image of the samples traced so far whenever onProgress fires
        # (pass a two-argument callable to receive it): the reference's progressive canvas blit (ray-tracer.js:236-238)
        def _cb(f, _u):
            try:
                onProgress(f, rgba)
            except TypeError:
                onProgress(f)
        cb = L.PROGRESS_CB(_cb) if onProgress else L.PROGRESS_CB()
        rc = self._L.brt_render(self._ctx, rgba.ctypes.data, fdat.ctypes.data if want_float else None,
                                lin.ctypes.data if want_linear else None, cb, None)
        L.check(self._ctx, rc)
        self.floatData, self.linearMean = fdat, lin
        return rgba

    def renderInto(self, rgba_ptr, float_ptr=None, linear_ptr=None, onProgress=None):
        """brt_render straight into caller-owned HOST buffers (addresses): the call the N-API addon makes with the backing
        store of imageData.data (ray-tracer.js:166-281).  No numpy allocation, no extra copy."""
        self._push_params()
        cb = L.PROGRESS_CB(lambda f, _u: onProgress(f)) if onProgress else L.PROGRESS_CB()
        vp = lambda p: C.c_void_p(int(p)) if p else None
        L.check(self._ctx, self._L.brt_render(self._ctx, vp(rgba_ptr), vp(float_ptr), vp(linear_ptr), cb, None))

    def cancel(self):                                                  # window.renderCancelled = true (ui-controller.js:134-137)
        self._L.brt_cancel(self._ctx)

    # -- engine extras ----------------------------------------------------------------------------------------
    def stats(self) -> dict:
        s = L.brt_stats()
        L.check(self._ctx, self._L.brt_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def sceneFlat(self) -> dict:
        """world.objects / materials / lights as libbrt holds them after ingest (numpy copies)."""
        d = L.brt_scene_desc()
        L.check(self._ctx, self._L.brt_scene_get_flat(self._ctx, C.byref(d)))
        objs = [dict(type=o.type, material=o.material, a=tuple(o.a), b=tuple(o.b), c=tuple(o.c), first_tri=o.first_tri,
                     tri_count=o.tri_count) for o in (d.objects[i] for i in range(d.n_objects))]
        mats = [dict(type=m.type, color=tuple(m.color), param=m.param) for m in (d.materials[i] for i in range(d.n_materials))]
        lights = [dict(type=l.type, v=tuple(l.v), color=tuple(l.color), intensity=l.intensity) for l in (d.lights[i] for i in range(d.n_lights))]
        n = int(d.n_mesh_triangles)
        tris = np.ctypeslib.as_array(d.mesh_triangles, shape=(n, 9)).copy() if n else np.zeros((0, 9))
        return dict(objects=objs, materials=mats, lights=lights, mesh_triangles=tris)

    def sceneFlatDesc(self):
        """-> (brt_scene_desc, keepalive): a caller-owned host copy of the ingested scene, in the form
        brt_scene_set_flat takes (what the JS shim builds from a live World)."""
        d = L.brt_scene_desc()
        L.check(self._ctx, self._L.brt_scene_get_flat(self._ctx, C.byref(d)))
        objs = (L.brt_object * max(1, d.n_objects))()
        mats = (L.brt_material * max(1, d.n_materials))()
        lights = (L.brt_light * max(1, d.n_lights))()
        C.memmove(objs, d.objects, C.sizeof(L.brt_object) * d.n_objects)
        C.memmove(mats, d.materials, C.sizeof(L.brt_material) * d.n_materials)
        C.memmove(lights, d.lights, C.sizeof(L.brt_light) * d.n_lights)
        n = int(d.n_mesh_triangles)
        tris = np.ctypeslib.as_array(d.mesh_triangles, shape=(n, 9)).copy() if n else np.zeros((0, 9))
        out = L.brt_scene_desc()
        out.objects, out.n_objects = objs, d.n_objects
        out.materials, out.n_materials = mats, d.n_materials
        out.lights, out.n_lights = lights, d.n_lights
        out.mesh_triangles, out.n_mesh_triangles = tris.ctypes.data_as(C.POINTER(C.c_double)), n
        texs = (L.brt_texture * max(1, d.n_textures))()
        if d.n_textures:
            C.memmove(texs, d.textures, C.sizeof(L.brt_texture) * d.n_textures)
        out.textures, out.n_textures = texs, d.n_textures
        out.flags = d.flags                                        # BRT_SCENE_CONSTRUCTED: post-constructor values, stored as they are on the way back
        return out, (objs, mats, lights, tris, texs)

    def setSceneFlat(self, desc_keep):
        """brt_scene_set_flat with a (desc, keepalive) pair from sceneFlatDesc() / World.flatten()."""
        L.check(self._ctx, self._L.brt_scene_set_flat(self._ctx, C.byref(desc_keep[0])))

    def sceneInfo(self) -> dict:
        s = L.brt_scene_info()
        L.check(self._ctx, self._L.brt_scene_info_get(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_ if not k.startswith("_")}

    def primaryAOV(self, precision=32) -> dict:
        """Primary-visibility AOVs at pixel centres, lens offset 0 (north-star parity contract)."""
        self._push_params()
        W, H = self.width, self.height
        obj, tri, ff = np.empty((H, W), np.int32), np.empty((H, W), np.int32), np.empty((H, W), np.uint8)
        ft = np.float64 if precision == 64 else np.float32
        t, n = np.empty((H, W), ft), np.empty((H, W, 3), ft)
        fn = self._L.brt_primary_aov_f64 if precision == 64 else self._L.brt_primary_aov_f32
        L.check(self._ctx, fn(self._ctx, obj.ctypes.data, tri.ctypes.data, t.ctypes.data, n.ctypes.data, ff.ctypes.data))
        return dict(obj_id=obj, tri_id=tri, t=t, normal=n, front_face=ff)

    def evalBackground(self, dirs) -> np.ndarray:
        d = np.ascontiguousarray(np.asarray(dirs, dtype=np.float64).reshape(-1, 3))
        out = np.empty((d.shape[0], 3), np.float32)
        L.check(self._ctx, self._L.brt_eval_background(self._ctx, d.ctypes.data_as(C.POINTER(C.c_double)), d.shape[0],
                                                       out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def evalTexture(self, tex_index, points) -> np.ndarray:
        """textures[tex_index].value(u, v, p) (js/textures.js) for an array of points, evaluated by the device code."""
        pts = np.ascontiguousarray(np.asarray(points, dtype=np.float64).reshape(-1, 3))
        out = np.empty((pts.shape[0], 3), np.float32)
        L.check(self._ctx, self._L.brt_eval_texture(self._ctx, int(tex_index), pts.ctypes.data_as(C.POINTER(C.c_double)), pts.shape[0],
                                                    out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def postprocess(self, linear_mean) -> np.ndarray:
        """Tone map / gamma / quantise (/ denoise) a host linear image with the current settings (GPU kernels)."""
        self._push_params()
        lin = np.ascontiguousarray(np.asarray(linear_mean, dtype=np.float32).reshape(self.height, self.width, 4))
        rgba = np.empty((self.height, self.width, 4), np.uint8)
        fdat = np.empty((self.height, self.width, 4), np.float32)
        L.check(self._ctx, self._L.brt_postprocess_host(self._ctx, lin.ctypes.data, rgba.ctypes.data, fdat.ctypes.data))
        self.floatData = fdat
        return rgba

    def rngStream(self, pixel, sample, n) -> np.ndarray:
        out = np.empty(n, np.float32)
        L.check(self._ctx, self._L.brt_debug_rng_stream(self._ctx, self.seed, pixel, sample, n, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def measureFp32Peak(self) -> float:
        v = C.c_double()
        L.check(self._ctx, self._L.brt_measure_fp32_peak(self._ctx, C.byref(v)))
        return v.value

    # device-resident path (multi-GPU spp split; bench.py)
    def setStream(self, cuda_stream):
        """`cuda_stream`: a cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream; 0 = the legacy default
        stream, passed on as cudaStreamLegacy = 0x1), or None to go back to the ctx-owned stream."""
        if cuda_stream is None:
            L.check(self._ctx, self._L.brt_set_stream(self._ctx, None))
        else:
            L.check(self._ctx, self._L.brt_set_stream(self._ctx, C.c_void_p(int(cuda_stream) or 1)))

    def renderAccumulate(self, d_accum_ptr, sample_begin, sample_count):
        L.check(self._ctx, self._L.brt_render_accumulate(self._ctx, C.c_void_p(int(d_accum_ptr)) if d_accum_ptr else None,
                                                         int(sample_begin), int(sample_count)))

    def resolveDevice(self, d_accum_ptr, d_rgba_ptr, d_float_ptr=None, d_linear_ptr=None):
        vp = lambda p: C.c_void_p(int(p)) if p else None
        L.check(self._ctx, self._L.brt_resolve_device(self._ctx, vp(d_accum_ptr), vp(d_rgba_ptr), vp(d_float_ptr), vp(d_linear_ptr)))

    def reduceResolvePeers(self, peer_ptrs, row_begin, row_end, d_rgba_root_ptr, d_float_root_ptr=None):
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        L.check(self._ctx, self._L.brt_reduce_resolve_peers(self._ctx, arr, len(peer_ptrs), int(row_begin), int(row_end),
                                                            C.c_void_p(int(d_rgba_root_ptr)),
                                                            C.c_void_p(int(d_float_root_ptr)) if d_float_root_ptr else None))

    def sharedAlloc(self, nbytes):
        """-> (device pointer, 64-byte CUDA IPC handle) of a library-owned buffer other ranks can map."""
        p, h = C.c_void_p(), C.create_string_buffer(64)
        L.check(self._ctx, self._L.brt_shared_alloc(self._ctx, int(nbytes), C.byref(p), h))
        return p.value, h.raw

    def sharedOpen(self, handle: bytes) -> int:
        p = C.c_void_p()
        L.check(self._ctx, self._L.brt_shared_open(self._ctx, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def sharedClose(self, ptr):
        L.check(self._ctx, self._L.brt_shared_close(self._ctx, C.c_void_p(int(ptr))))

    def sharedFree(self, ptr):
        L.check(self._ctx, self._L.brt_shared_free(self._ctx, C.c_void_p(int(ptr))))

    def deviceMemset(self, ptr, value, nbytes):
        L.check(self._ctx, self._L.brt_device_memset(self._ctx, C.c_void_p(int(ptr)), int(value), int(nbytes)))

    def copyToHost(self, host_ptr, dev_ptr, nbytes):
        L.check(self._ctx, self._L.brt_copy_to_host(self._ctx, C.c_void_p(int(host_ptr)), C.c_void_p(int(dev_ptr)), int(nbytes)))

    def synchronize(self):
        L.check(self._ctx, self._L.brt_stream_synchronize(self._ctx))

    # peer group: one process per GPU, exchange fused into the resolve kernel (include/brt.h, brt_peer_*)
    def deviceCount(self) -> int:
        return int(self._L.brt_device_count(self._ctx))

    def peerAlloc(self, rank, world) -> bytes:
        self._push_params()                                            # the block is sized from width x height
        h = C.create_string_buffer(64)
        L.check(self._ctx, self._L.brt_peer_alloc(self._ctx, int(rank), int(world), h))
        return h.raw

    def peerConnect(self, handles):
        blob = b"".join(bytes(h) for h in handles)
        L.check(self._ctx, self._L.brt_peer_connect(self._ctx, C.create_string_buffer(blob, len(blob))))

    def peerRender(self, sample_begin, sample_count, want_float=False, want_linear=False):
        L.check(self._ctx, self._L.brt_peer_render(self._ctx, int(sample_begin), int(sample_count), int(want_float), int(want_linear)))

    def peerFetch(self, rgba_ptr=None, float_ptr=None, linear_ptr=None):
        vp = lambda p: C.c_void_p(int(p)) if p else None
        L.check(self._ctx, self._L.brt_peer_fetch(self._ctx, vp(rgba_ptr), vp(float_ptr), vp(linear_ptr)))

    def peerImagePtr(self) -> int:
        p = C.c_void_p()
        L.check(self._ctx, self._L.brt_peer_image_ptr(self._ctx, C.byref(p)))
        return p.value

    def peerFree(self):
        L.check(self._ctx, self._L.brt_peer_free(self._ctx))
