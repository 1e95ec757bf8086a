"""Brute force vs hierarchy on a FULL-SIZE workload scene (fast sampler, megakernel): the O(N) loops of the reference
(world.js:24-30, geometry.js:253-259) and every hierarchy variant must give bit-identical sums for the same Philox stream.
    python tools/bvh_exact_check.py c5 960 540 4 [widths...]      # prints one JSON line: differing pixels per variant
The library under test is the one BRT_LIBBRT names (kernel A/B variants), default blenderraytracer_b200/libbrt.so."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import blenderraytracer_b200 as brt
from bench import load_workload


def main():
    name, W, H, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    widths = [int(x) for x in sys.argv[5:]] or [2]
    w = load_workload(name, binary=True)
    rt = brt.RayTracer(W, H, device=0, seed=7)
    assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode())
    rt.resizeCanvas(W, H)
    rt.updateRenderSettings(dict(samples=spp, maxBounces=w["depth"]))
    rt.sampler = "fast"
    rt.setStream(torch.cuda.current_stream().cuda_stream)
    out, res = {}, {"workload": name, "W": W, "H": H, "spp": spp, "lib": os.environ.get("BRT_LIBBRT", "default")}
    for key, accel, width in [("brute", "brute", 0)] + [(f"bvh{x}", "bvh", x) for x in widths]:
        rt.accel, rt.bvhWidth = accel, width
        rt._push_params()
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize(); t0 = time.time()
        rt.renderAccumulate(acc.data_ptr(), 0, spp)
        torch.cuda.synchronize()
        out[key] = acc.cpu().numpy()
        res[key + "_s"] = round(time.time() - t0, 3)
    for key in out:
        if key != "brute":
            d = (out[key] != out["brute"]).any(axis=2)
            res[key + "_differing_pixels"] = int(d.sum())
            if d.any():
                ys, xs = np.nonzero(d)
                res[key + "_first"] = [int(xs[0]), int(ys[0]), out[key][ys[0], xs[0]].tolist(), out["brute"][ys[0], xs[0]].tolist()]
    print("EXACT_CHECK " + json.dumps(res))


if __name__ == "__main__":
    main()
