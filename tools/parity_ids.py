"""Primary-hit ID parity on every BASELINE config at its full image size (run on a B200): the render path's own
primary-hit code (brt_primary_aov_f32, hierarchy and linear loop) against the float64 brute-force kernel
(brt_primary_aov_f64, which tests/ pin bit-exactly to the oracle).  Writes gpurun_out/parity_ids.json (-> profiles/).
C5's float64 brute force is 8.3 Mpx x 1.0 M triangles: its frame is checked on a 768x432 rendering of the same camera."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import blenderraytracer_b200 as brt
from bench import load_workload

out = []
for name, size in (("c1", None), ("c2", None), ("c3", None), ("c4", None), ("c5", (768, 432))):
    w = load_workload(name, binary=True)
    W, H = size or (w["W"], w["H"])
    rt = brt.RayTracer(W, H, device=0, seed=1)
    assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode())
    rt.resizeCanvas(W, H)
    t0 = time.time()
    a64 = rt.primaryAOV(64)
    row = dict(config=name, workload=w["desc"], width=W, height=H, pixels=W * H, hit_pixels=int((a64["obj_id"] >= 0).sum()))
    for accel in ("bvh", "brute"):
        if accel == "brute" and name == "c5":
            continue                                        # the fp32 linear loop over 1 M triangles adds nothing the bvh row does not show
        rt.accel = accel
        a32 = rt.primaryAOV(32)
        mism = (a32["obj_id"] != a64["obj_id"]) | (a32["tri_id"] != a64["tri_id"])
        ok = ~mism & (a64["obj_id"] >= 0)
        rel = np.abs(a32["t"][ok].astype(np.float64) - a64["t"][ok]) / a64["t"][ok]
        row[accel] = dict(id_mismatches=int(mism.sum()), t_equals_float64_rounded=bool(np.array_equal(a32["t"][ok], a64["t"][ok].astype(np.float32))),
                          max_rel_t_err=float(rel.max()) if rel.size else 0.0,
                          max_normal_err=float(np.abs(a32["normal"][ok] - a64["normal"][ok]).max()) if ok.any() else 0.0,
                          front_face_mismatches=int((a32["front_face"][ok] != a64["front_face"][ok]).sum()))
    row["seconds"] = round(time.time() - t0, 2)
    out.append(row)
    print(json.dumps(row), flush=True)
    rt.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "parity_ids.json"), "w"), indent=1)
