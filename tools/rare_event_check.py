"""Which of two kernel variants is right when their full-size images differ in a handful of pixels?
Renders the workload at full size with both libraries (hierarchy, fast sampler), finds the image rows whose sums differ, then
re-renders ONLY those rows (BRT_DEBUG_ROW_WINDOW) with the reference's brute-force loops and with the hierarchy, in both
libraries, and reports which variant matches brute force on every pixel of those rows.
    python tools/rare_event_check.py c5:64 a=blenderraytracer_b200/libbrt_ch0.so b=blenderraytracer_b200/libbrt.so"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, sys, os, zlib
sys.path.insert(0, %r)
import torch, numpy as np
import blenderraytracer_b200 as brt
from bench import load_workload
name, spp = sys.argv[1].split(":"); spp = int(spp)
mode = sys.argv[2]
w = load_workload(name, binary=True)
W, H = w["W"], w["H"]
rt = brt.RayTracer(W, H, device=0, seed=1)
assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode())
rt.resizeCanvas(W, H)
rt.updateRenderSettings(dict(samples=spp, maxBounces=w["depth"]))
rt.sampler = "fast"
rt.setStream(torch.cuda.current_stream().cuda_stream)
def sums(accel):
    rt.accel = accel
    rt._push_params()
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    rt.renderAccumulate(acc.data_ptr(), 0, spp)
    torch.cuda.synchronize()
    return acc.cpu().numpy()
if mode == "full":
    a = sums("bvh")
    print("RESULT " + json.dumps([zlib.crc32(a[y].tobytes()) for y in range(H)]))
else:
    y0, y1 = [int(v) for v in os.environ["BRT_DEBUG_ROW_WINDOW"].split(",")]
    b, h = sums("brute")[y0:y1], sums("bvh")[y0:y1]
    d = (b != h).any(axis=2)
    ys, xs = np.nonzero(d)
    print("RESULT " + json.dumps(dict(differing_pixels=int(d.sum()), where=[[int(x), int(y0 + y)] for x, y in zip(xs[:8], ys[:8])],
                                      brute=[b[y, x].tolist() for x, y in zip(xs[:4], ys[:4])], bvh=[h[y, x].tolist() for x, y in zip(xs[:4], ys[:4])])))
''' % ROOT


def run(lib, spec, mode, env_extra=None):
    env = dict(os.environ, BRT_LIBBRT=os.path.abspath(lib))
    env.update(env_extra or {})
    p = subprocess.run([sys.executable, "-c", WORKER, spec, mode], capture_output=True, text=True, env=env, cwd=ROOT)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        raise SystemExit(f"{lib} {mode} failed:\n{p.stderr[-2000:]}")
    return json.loads(line[0][7:])


def main():
    spec = sys.argv[1]
    libs = dict(a.split("=", 1) for a in sys.argv[2:])
    rows = {k: run(v, spec, "full") for k, v in libs.items()}
    names = list(libs)
    diff = [y for y in range(len(rows[names[0]])) if len({rows[k][y] for k in names}) > 1]
    out = {"spec": spec, "differing_rows": diff[:32], "n_differing_rows": len(diff), "vs_brute": {}}
    for y in diff[:4]:
        for k, v in libs.items():
            out["vs_brute"][f"{k}@row{y}"] = run(v, spec, "rows", {"BRT_DEBUG_ROW_WINDOW": f"{y},{y + 1}"})
    print("RARE_EVENT_CHECK " + json.dumps(out))


if __name__ == "__main__":
    main()
