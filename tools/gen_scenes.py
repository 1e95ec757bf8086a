"""Synthetic benchmark scenes of BASELINE.json (C3, C4, C5), emitted in the reference's own scene JSON format
(docs/scene_format.md) so the GPU engine and the oracle ingest the same bytes.  SURVEY.md §8(d) fixes the recipes.

    python tools/gen_scenes.py c3 > c3.json

Everything is seeded (numpy default_rng); no file of the reference is read.
"""
from __future__ import annotations

import json
import math
import sys

import numpy as np


def _r(x, nd=6):
    return [round(float(v), nd) for v in x]


def random_spheres(seed: int = 42, grid: int = 11, width: int = 1920, height: int = 1080, ground: str = "plane") -> dict:
    """C3 — RTiOW-style field: (2*grid)^2 candidate small spheres (r = 0.2) + three r = 1 spheres + a ground.
    Camera (13,2,3) -> origin, fov 20, aperture 0.1, focusDist 10; gradient sky.  ~485 spheres at grid = 11."""
    rng = np.random.default_rng(seed)
    objs = []
    if ground == "plane":
        objs.append(dict(type="plane", point=[0, 0, 0], normal=[0, 1, 0], material=dict(type="lambertian", color=[0.5, 0.5, 0.5])))
    else:                                     # robustness variant: the r = 1000 "ground sphere" idiom
        objs.append(dict(type="sphere", center=[0, -1000, 0], radius=1000, material=dict(type="lambertian", color=[0.5, 0.5, 0.5])))
    for a in range(-grid, grid):
        for b in range(-grid, grid):
            choose = rng.random()
            c = np.array([a + 0.9 * rng.random(), 0.2, b + 0.9 * rng.random()])
            mat_draw = rng.random(6)
            if np.linalg.norm(c - np.array([4, 0.2, 0])) <= 0.9:
                continue
            if choose < 0.8:
                m = dict(type="lambertian", color=_r(mat_draw[:3] * mat_draw[3:6]))
            elif choose < 0.95:
                m = dict(type="metal", color=_r(0.5 + 0.5 * mat_draw[:3]), roughness=round(float(0.5 * mat_draw[3]), 6))
            else:
                m = dict(type="dielectric", ior=1.5)
            objs.append(dict(type="sphere", center=_r(c), radius=0.2, material=m))
    objs.append(dict(type="sphere", center=[0, 1, 0], radius=1.0, material=dict(type="dielectric", ior=1.5)))
    objs.append(dict(type="sphere", center=[-4, 1, 0], radius=1.0, material=dict(type="lambertian", color=[0.4, 0.2, 0.1])))
    objs.append(dict(type="sphere", center=[4, 1, 0], radius=1.0, material=dict(type="metal", color=[0.7, 0.6, 0.5], roughness=0.0)))
    return dict(objects=objs, lights=[],
                camera=dict(position=[13, 2, 3], lookAt=[0, 0, 0], up=[0, 1, 0], fov=20, aspect=width / height,
                            aperture=0.1, focusDist=10.0, type="perspective"),
                background=dict(type="gradient", intensity=1.0))


def _quad(p0, p1, p2, p3, material):
    """Two triangles (p0,p1,p2), (p0,p2,p3) as one mesh object."""
    return dict(type="mesh", vertices=[_r(p0), _r(p1), _r(p2), _r(p3)], indices=[0, 1, 2, 0, 2, 3], material=material)


def cornell(background: str = "procedural_sky", width: int = 1920, height: int = 1080) -> dict:
    """C4 — the reference's own Cornell preset geometry (ray-tracer.js:400-435: 5 planes, 2 boxes, glass + mirror sphere)
    with the thin emissive box replaced by emissive QUADS (two triangles each), open front, and a procedural-sky or
    HDRI background seen through the open front via the mirror / glass."""
    white = dict(type="lambertian", color=[0.73, 0.73, 0.73])
    red = dict(type="lambertian", color=[0.65, 0.05, 0.05])
    green = dict(type="lambertian", color=[0.12, 0.45, 0.15])
    light = dict(type="emissive", color=[1, 1, 1], intensity=15)
    warm = dict(type="emissive", color=[1, 0.8, 0.6], intensity=6)
    objs = [
        dict(type="plane", point=[0, 0, -5], normal=[0, 0, 1], material=white),
        dict(type="plane", point=[0, -2.5, 0], normal=[0, 1, 0], material=white),
        dict(type="plane", point=[0, 2.5, 0], normal=[0, -1, 0], material=white),
        dict(type="plane", point=[-2.5, 0, 0], normal=[1, 0, 0], material=red),
        dict(type="plane", point=[2.5, 0, 0], normal=[-1, 0, 0], material=green),
        dict(type="box", min=[-1, -2.5, -3.5], max=[-0.2, -1, -2.7], material=white),
        dict(type="box", min=[0.2, -2.5, -4], max=[1.2, -0.5, -3], material=white),
        dict(type="sphere", center=[-0.6, -1.8, -2.2], radius=0.7, material=dict(type="dielectric", ior=1.5)),
        dict(type="sphere", center=[0.7, -1.8, -3.5], radius=0.7, material=dict(type="metal", color=[0.8, 0.85, 0.88], roughness=0.0)),
        _quad([-0.5, 2.45, -3.5], [0.5, 2.45, -3.5], [0.5, 2.45, -2.5], [-0.5, 2.45, -2.5], light),
        _quad([-2.45, 0.4, -4.2], [-2.45, 0.4, -3.4], [-2.45, 1.2, -3.4], [-2.45, 1.2, -4.2], warm),
        _quad([2.45, 0.4, -3.4], [2.45, 0.4, -4.2], [2.45, 1.2, -4.2], [2.45, 1.2, -3.4], warm),
    ]
    return dict(objects=objs, lights=[],
                camera=dict(position=[0, 0, 2], lookAt=[0, 0, -1], up=[0, 1, 0], fov=40, aspect=width / height,
                            aperture=0.0, focusDist=10.0, type="perspective"),
                background=dict(type=background, intensity=1.0))


def _value_noise(nx: int, nz: int, rng, octaves: int = 4) -> np.ndarray:
    """Seeded multi-octave value noise on an (nz, nx) vertex grid (heights for the C5 terrain)."""
    h = np.zeros((nz, nx))
    amp, cells = 1.0, 6
    zi, xi = np.meshgrid(np.linspace(0, 1, nz), np.linspace(0, 1, nx), indexing="ij")
    for _ in range(octaves):
        lat = rng.random((cells + 2, cells + 2))
        fx, fz = xi * cells, zi * cells
        ix, iz = np.minimum(fx.astype(int), cells - 1), np.minimum(fz.astype(int), cells - 1)
        tx, tz = fx - ix, fz - iz
        sx, sz = tx * tx * (3 - 2 * tx), tz * tz * (3 - 2 * tz)
        a = lat[iz, ix] * (1 - sx) + lat[iz, ix + 1] * sx
        b = lat[iz + 1, ix] * (1 - sx) + lat[iz + 1, ix + 1] * sx
        h += amp * (a * (1 - sz) + b * sz)
        amp *= 0.5
        cells *= 2
    return h


def terrain_arrays(quads: int = 708, extent: float = 200.0, seed: int = 42, height_scale: float = 12.0):
    """C5 geometry: displaced regular grid of quads x quads cells -> 2*quads^2 triangles (708 -> 1 002 528).
    Returns (vertices (n,3) float64 rounded to 1e-4 as they would print in JSON, indices (m,) int64)."""
    rng = np.random.default_rng(seed)
    n = quads + 1
    xs = np.linspace(-extent / 2, extent / 2, n)
    X, Z = np.meshgrid(xs, xs, indexing="xy")
    Y = (_value_noise(n, n, rng) - 0.9) * height_scale
    V = np.round(np.stack([X, Y, Z], axis=-1).reshape(-1, 3), 4)
    i0 = (np.arange(quads)[:, None] * n + np.arange(quads)[None, :]).reshape(-1)
    idx = np.stack([i0, i0 + n, i0 + 1, i0 + 1, i0 + n, i0 + n + 1], axis=1).reshape(-1)
    return V, idx.astype(np.int64)


def terrain(quads: int = 708, extent: float = 200.0, seed: int = 42, width: int = 3840, height: int = 2160, as_arrays: bool = False):
    """C5 — one lambertian terrain mesh + a few emissive quads floating above it; camera fd = |pos - lookAt|, aperture 0.
    With as_arrays=True the mesh's `vertices` / `indices` stay numpy arrays (fast path through World.flatten);
    otherwise they are nested lists, ready for json.dumps."""
    V, I = terrain_arrays(quads, extent, seed)
    s = extent / 200.0
    mesh = dict(type="mesh", vertices=V if as_arrays else V.tolist(), indices=I if as_arrays else I.tolist(),
                material=dict(type="lambertian", color=[0.55, 0.5, 0.42]))
    lights = []
    for k, (cx, cz) in enumerate([(-30, -20), (25, 10), (0, 45)]):
        y = 18.0 * s
        cx, cz, w = cx * s, cz * s, 8.0 * s
        lights.append(_quad([cx - w, y, cz - w], [cx + w, y, cz - w], [cx + w, y, cz + w], [cx - w, y, cz + w],
                            dict(type="emissive", color=[1.0, 0.9 - 0.1 * k, 0.7], intensity=6)))
    pos, at = [0.0, 30.0 * s, 95.0 * s], [0.0, -5.0 * s, 0.0]
    fd = math.dist(pos, at)
    return dict(objects=[mesh] + lights, lights=[],
                camera=dict(position=pos, lookAt=at, up=[0, 1, 0], fov=45, aspect=width / height, aperture=0.0, focusDist=fd,
                            type="perspective"),
                background=dict(type="gradient", intensity=1.0))


SCENES = {"c3": random_spheres, "c4": cornell, "c5": terrain}

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    json.dump(SCENES[name](), sys.stdout)
