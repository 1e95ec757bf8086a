"""Ad-hoc GPU diagnostics (first contact with the hardware); the real gates are tests/ -m gpu."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blenderraytracer_b200 as brt
from oracle.oracle import OracleRayTracer, make_perm

def load(name): return json.load(open(os.path.join(ROOT, "tests", "golden", name)))

def aov_check(name, W, H):
    sc = load(name)
    rt = brt.RayTracer(W, H, seed=3); assert rt.loadFromJSON(sc)
    orc = OracleRayTracer(W, H, seed=3, threads=8); assert orc.loadFromJSON(sc)
    ao = orc.primary_aov()
    a64 = rt.primaryAOV(64)
    print(name, "f64 ids equal:", np.array_equal(a64["obj_id"], ao["obj_id"]), np.array_equal(a64["tri_id"], ao["tri_id"]),
          "t bit-equal:", np.array_equal(a64["t"], ao["t"]), "n bit-equal:", np.array_equal(a64["normal"], ao["normal"]),
          "ff:", np.array_equal(a64["front_face"], ao["front_face"]))
    for accel in ("brute", "bvh"):
        rt.accel = accel
        a32 = rt.primaryAOV(32)
        mism = (a32["obj_id"] != ao["obj_id"]) | (a32["tri_id"] != ao["tri_id"])
        ok = ~mism & (ao["obj_id"] >= 0)
        rel = np.abs(a32["t"][ok] - ao["t"][ok]) / ao["t"][ok]
        nerr = np.abs(a32["normal"][ok] - ao["normal"][ok]).max()
        print(f"  f32 {accel}: id mismatches {int(mism.sum())}/{mism.size}, max rel t err {rel.max():.2e}, max |dn| {nerr:.2e}")
    return rt, orc

def render_check(name, W, H, spp, depth, sampler):
    sc = load(name)
    rt = brt.RayTracer(W, H, seed=11); assert rt.loadFromJSON(sc)
    rt.updateRenderSettings(dict(samples=spp, maxBounces=depth)); rt.sampler = sampler
    t0 = time.time(); img = rt.render(want_linear=True); t1 = time.time()
    st = rt.stats()
    orc = OracleRayTracer(W, H, seed=11, threads=8); assert orc.loadFromJSON(sc)
    orc.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    t2 = time.time(); ref = orc.render(); t3 = time.time()
    orc2 = OracleRayTracer(W, H, seed=12, threads=8); orc2.loadFromJSON(sc); orc2.updateRenderSettings(dict(samples=spp, maxBounces=depth))
    ref2 = orc2.render()
    f = lambda a: a[..., :3].astype(np.float64)
    rmse = np.sqrt(np.mean((rt.floatData[..., :3] - orc.floatData[..., :3]) ** 2))
    floor = np.sqrt(np.mean((orc2.floatData[..., :3] - orc.floatData[..., :3]) ** 2))
    bias = (rt.floatData[..., :3].astype(np.float64) - orc.floatData[..., :3]).mean(axis=(0, 1))
    d = np.abs(f(img) - f(ref)).max(axis=-1)
    print(f"{name} {W}x{H}x{spp} d{depth} {sampler}: gpu kernel {st['kernel_ms']:.2f} ms ({W*H*spp/st['kernel_ms']/1e3:.1f} Msamples/s), wall {t1-t0:.3f}s; "
          f"oracle 8T {t3-t2:.2f}s ({W*H*spp/(t3-t2)/1e6:.3f} Ms/s); RMSE {rmse:.5f} vs noise floor {floor:.5f}; bias {bias}; "
          f"px within 1 LSB {np.mean(d<=1):.4f}, 2 LSB {np.mean(d<=2):.4f}")

if __name__ == "__main__":
    rt = brt.RayTracer(64, 64)
    print("fp32 peak TFLOP/s:", rt.measureFp32Peak())
    from oracle.oracle import lib, C
    out = (C.c_double * 16)(); lib().orc_rng_stream(1, 5, 9, 16, out)
    print("rng equal:", np.array_equal(np.array(out[:], dtype=np.float32), rt.rngStream(5, 9, 16)))
    aov_check("sample_scene.json", 600, 400)
    aov_check("sample_mesh.json", 1280, 720)
    for sampler in ("reference", "fast"):
        render_check("sample_scene.json", 300, 200, 16, 10, sampler)
        render_check("sample_mesh.json", 320, 180, 16, 10, sampler)
    # full-size timing of C1/C2
    for name, W, H, spp in (("sample_scene.json", 600, 400, 16), ("sample_mesh.json", 1280, 720, 64)):
        rt = brt.RayTracer(W, H, seed=1); rt.loadFromJSON(load(name)); rt.updateRenderSettings(dict(samples=spp, maxBounces=10))
        for accel in ("brute", "bvh"):
            rt.accel = accel
            rt.render(); rt.render(); st = rt.stats()
            print(f"{name} {W}x{H}x{spp} {accel}: kernel {st['kernel_ms']:.2f} ms -> {W*H*spp/st['kernel_ms']/1e3:.1f} Msamples/s, total {st['total_ms']:.2f} ms")
