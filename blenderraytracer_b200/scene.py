"""Host-side mirror of the reference's scene classes (same names, same constructor arguments) so that code which
builds a scene object by object — the presets in js/ray-tracer.js:42-77 and :336-435 — reads the same here.
These are plain data holders: all intersection / shading work happens on the GPU after ``World.flatten()``.

Reference: js/geometry.js (Sphere :8, Plane :49, Box :78, Triangle :136, TriangleMesh :192), js/materials.js
(Lambertian :14, Metal :28, Dielectric :44, Emissive :86), js/lights.js (PointLight :12, DirectionalLight :35),
js/world.js (World :8-18).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

from . import _lib as L


def _v(x):
    return (float(x[0]), float(x[1]), float(x[2]))


# ---- materials (js/materials.js) --------------------------------------------------------------------------
@dataclass
class Lambertian:
    albedo: Sequence[float]

    def _desc(self):
        return L.MAT_LAMBERTIAN, _v(self.albedo), 0.0


@dataclass
class Metal:
    albedo: Sequence[float]
    roughness: float = 0.0

    def _desc(self):
        return L.MAT_METAL, _v(self.albedo), float(self.roughness)


@dataclass
class Dielectric:
    refractionIndex: float

    def _desc(self):
        return L.MAT_DIELECTRIC, (1.0, 1.0, 1.0), float(self.refractionIndex)


@dataclass
class Emissive:
    color: Sequence[float]
    intensity: float = 1.0

    def _desc(self):
        return L.MAT_EMISSIVE, _v(self.color), float(self.intensity)


# ---- textures (js/textures.js) and textured materials (js/materials.js:99-126) --------------------------------
@dataclass
class SolidColor:
    color: Sequence[float]

    def _tex(self):
        return L.TEX["solid"], _v(self.color), (0.0, 0.0, 0.0), 1.0, None


@dataclass
class CheckerTexture:
    odd: Sequence[float]
    even: Sequence[float]
    scale: float = 10

    def _tex(self):
        return L.TEX["checker"], _v(self.odd), _v(self.even), float(self.scale), None


class _NoiseBased:
    """`perm256`: the texture's own PerlinNoise table (textures.js:44,58,74 draw it with Math.random; an input here)."""
    kind = "noise"

    def __init__(self, scale=1, perm256=None):
        self.scale = float(scale)
        self.perm = np.arange(256, dtype=np.uint8) if perm256 is None else np.asarray(perm256, dtype=np.uint8).reshape(256)

    def _tex(self):
        return L.TEX[self.kind], (1.0, 1.0, 1.0), (1.0, 1.0, 1.0), self.scale, self.perm


class NoiseTexture(_NoiseBased):
    kind = "noise"


class MarbleTexture(_NoiseBased):
    kind = "marble"


class WoodTexture(_NoiseBased):
    kind = "wood"


@dataclass
class TexturedLambertian:
    texture: object

    def _desc(self):
        return L.MAT_LAMBERTIAN, (1.0, 1.0, 1.0), 0.0


@dataclass
class TexturedMetal:
    texture: object
    roughness: float = 0.0

    def _desc(self):
        return L.MAT_METAL, (1.0, 1.0, 1.0), float(self.roughness)


# ---- geometry (js/geometry.js) ----------------------------------------------------------------------------
@dataclass
class Sphere:
    center: Sequence[float]
    radius: float
    material: object


@dataclass
class Plane:
    point: Sequence[float]
    normal: Sequence[float]
    material: object


@dataclass
class Box:
    min: Sequence[float]
    max: Sequence[float]
    material: object


@dataclass
class Triangle:
    v0: Sequence[float]
    v1: Sequence[float]
    v2: Sequence[float]
    material: object


class TriangleMesh:
    """geometry.js:193-237: triples of ``indices`` become triangles; an incomplete tail is dropped (:207-210), a triple
    with any index >= len(vertices) is dropped (:216-219), any other index that names no vertex reads (0,0,0) (:240-246).
    ``self.triangles`` is the surviving (n, 9) float64 array — its row numbers are the triangle IDs."""

    def __init__(self, vertices, indices, material):
        self.material = material
        v = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        idx = np.asarray(indices, dtype=np.float64).reshape(-1)
        idx = idx[: (idx.shape[0] // 3) * 3].reshape(-1, 3)
        nv = v.shape[0]
        keep = ~np.any(idx >= nv, axis=1)
        idx = idx[keep]
        ok = (idx >= 0) & (idx == np.floor(idx)) & (idx < nv)
        safe = np.where(ok, idx, 0).astype(np.int64)
        tri = v[safe] if nv else np.zeros(idx.shape + (3,))
        tri = np.where(ok[..., None], tri, 0.0)
        self.triangles = np.ascontiguousarray(tri.reshape(-1, 9))


# ---- lights (js/lights.js) --------------------------------------------------------------------------------
@dataclass
class PointLight:
    position: Sequence[float]
    color: Sequence[float]
    intensity: float = 1.0


@dataclass
class DirectionalLight:
    direction: Sequence[float]
    color: Sequence[float]
    intensity: float = 1.0


# ---- world (js/world.js:8-18) ------------------------------------------------------------------------------
@dataclass
class World:
    objects: List[object] = field(default_factory=list)
    lights: List[object] = field(default_factory=list)

    def add(self, obj):
        self.objects.append(obj)

    def addLight(self, light):
        self.lights.append(light)

    def flatten(self):
        """-> (brt_scene_desc, keepalive).  One brt_object per world.objects entry, in order (index = object ID)."""
        n = len(self.objects)
        objs = (L.brt_object * max(n, 1))()
        mats = (L.brt_material * max(n, 1))()
        meshes = []
        first = 0
        texs = [o.material.texture for o in self.objects if hasattr(o.material, "texture")]
        textures = (L.brt_texture * max(len(texs), 1))()
        n_tex = 0
        for i, o in enumerate(self.objects):
            t, col, p = o.material._desc()
            mats[i].type, mats[i].color, mats[i].param = t, L.d3(*col), p
            if hasattr(o.material, "texture"):
                kind, odd, even, scale, perm = o.material.texture._tex()
                tx = textures[n_tex]
                tx.kind, tx.odd, tx.even, tx.scale = kind, L.d3(*odd), L.d3(*even), scale
                if perm is not None:
                    C.memmove(tx.perm, np.ascontiguousarray(perm).ctypes.data, 256)
                n_tex += 1
                mats[i].texture = n_tex                      # 1-based
            ob = objs[i]
            ob.material = i
            if isinstance(o, Sphere):
                ob.type, ob.a, ob.b = L.OBJ_SPHERE, L.d3(*_v(o.center)), L.d3(float(o.radius), 0.0, 0.0)
            elif isinstance(o, Plane):
                ob.type, ob.a, ob.b = L.OBJ_PLANE, L.d3(*_v(o.point)), L.d3(*_v(o.normal))
            elif isinstance(o, Box):
                ob.type, ob.a, ob.b = L.OBJ_BOX, L.d3(*_v(o.min)), L.d3(*_v(o.max))
            elif isinstance(o, Triangle):
                ob.type, ob.a, ob.b, ob.c = L.OBJ_TRIANGLE, L.d3(*_v(o.v0)), L.d3(*_v(o.v1)), L.d3(*_v(o.v2))
            elif isinstance(o, TriangleMesh):
                ob.type, ob.first_tri, ob.tri_count = L.OBJ_MESH, first, o.triangles.shape[0]
                meshes.append(o.triangles)
                first += o.triangles.shape[0]
            else:
                raise TypeError(f"unsupported object {type(o).__name__}")
        tris = np.ascontiguousarray(np.concatenate(meshes, axis=0)) if meshes else np.zeros((0, 9))
        nl = len(self.lights)
        lights = (L.brt_light * max(nl, 1))()
        for i, l in enumerate(self.lights):
            if isinstance(l, PointLight):
                lights[i].type, lights[i].v = L.LIGHT_POINT, L.d3(*_v(l.position))
            elif isinstance(l, DirectionalLight):
                lights[i].type, lights[i].v = L.LIGHT_DIRECTIONAL, L.d3(*_v(l.direction))
            else:
                raise TypeError(f"unsupported light {type(l).__name__}")
            lights[i].color, lights[i].intensity = L.d3(*_v(l.color)), float(l.intensity)
        d = L.brt_scene_desc()
        d.objects, d.n_objects = objs, n
        d.materials, d.n_materials = mats, n
        d.mesh_triangles = tris.ctypes.data_as(C.POINTER(C.c_double))
        d.n_mesh_triangles = tris.shape[0]
        d.lights, d.n_lights = lights, nl
        d.textures, d.n_textures = textures, n_tex
        return d, (objs, mats, tris, lights, textures)
