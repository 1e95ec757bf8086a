"""Scratch GPU diagnostics (not part of the test suite)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blenderraytracer_b200 as brt
from oracle.oracle import OracleRayTracer
from tools import gen_scenes

def aov_diff(scene, W, H, accel="bvh"):
    rt = brt.RayTracer(W, H, seed=3); assert rt.loadFromJSON(scene)
    orc = OracleRayTracer(W, H, seed=3, threads=8); assert orc.loadFromJSON(scene)
    rt.accel = accel
    a, o = rt.primaryAOV(32), orc.primary_aov()
    mism = (a["obj_id"] != o["obj_id"]) | (a["tri_id"] != o["tri_id"])
    ok = ~mism & (o["obj_id"] >= 0)
    dn = np.abs(a["normal"].astype(np.float64) - o["normal"]).max(axis=-1)
    rel = np.abs(a["t"].astype(np.float64) - o["t"]) / np.where(ok, o["t"], 1)
    bad = ok & ((dn > 1e-5) | (rel > 1e-5))
    print(f"id mismatches {int(mism.sum())}/{mism.size}; bad t/normal pixels {int(bad.sum())}; max rel t {rel[ok].max():.2e} max dn {dn[ok].max():.2e}")
    ys, xs = np.nonzero(bad)
    for y, x in list(zip(ys, xs))[:12]:
        print(f"  px ({x},{y}) obj {o['obj_id'][y,x]} tri {o['tri_id'][y,x]} t ref {o['t'][y,x]:.9g} gpu {a['t'][y,x]:.9g} n ref {o['normal'][y,x]} gpu {a['normal'][y,x]} ff {o['front_face'][y,x]} {a['front_face'][y,x]}")

def bvh_vs_brute(scene, W, H):
    rt = brt.RayTracer(W, H, seed=5); assert rt.loadFromJSON(scene)
    aov = {}
    for accel in ("brute", "bvh"):
        rt.accel = accel
        aov[accel] = rt.primaryAOV(32)
    for key in ("obj_id", "tri_id", "t", "normal", "front_face"):
        d = aov["brute"][key] != aov["bvh"][key]
        if d.ndim == 3: d = d.any(axis=-1)
        print(f"  aov {key}: {int(d.sum())} differ")
        ys, xs = np.nonzero(d)
        for y, x in list(zip(ys, xs))[:5]:
            print(f"    px ({x},{y}) brute obj {aov['brute']['obj_id'][y,x]} tri {aov['brute']['tri_id'][y,x]} t {aov['brute']['t'][y,x]:.9g} | bvh obj {aov['bvh']['obj_id'][y,x]} tri {aov['bvh']['tri_id'][y,x]} t {aov['bvh']['t'][y,x]:.9g}")
    rt.updateRenderSettings(dict(samples=4, maxBounces=6)); rt.sampler = "reference"
    out = {}
    for accel in ("brute", "bvh"):
        rt.accel = accel
        rt.render(want_linear=True); out[accel] = rt.linearMean.copy()
    d = np.abs(out["brute"] - out["bvh"]).max(axis=-1)
    ys, xs = np.nonzero(d > 0)
    print(f"bvh vs brute: {len(ys)} pixels differ, max {d.max():.3e}")
    for y, x in list(zip(ys, xs))[:10]:
        print(f"  px ({x},{y}) brute {out['brute'][y,x]} bvh {out['bvh'][y,x]}")

if __name__ == "__main__":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T
    print("c4 cornell"); aov_diff(gen_scenes.cornell("hdri"), 480, 270)
    print("c5 small"); bvh_vs_brute(gen_scenes.terrain(quads=24, extent=200.0), 320, 180)
    print("ties"); bvh_vs_brute(T._tie_scene(), 360, 240); aov_diff(T._tie_scene(), 360, 240)
    print("sample_mesh"); bvh_vs_brute(json.load(open(os.path.join(ROOT, "tests/golden/sample_mesh.json"))), 1280, 720)
