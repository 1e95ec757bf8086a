"""A/B timing of kernel-tuning variants of libbrt on a B200: every variant is a separate .so built with
`make -C blenderraytracer_b200/csrc B=build_x OUT=../libbrt_x.so EXTRA=-D...`, loaded through BRT_LIBBRT in its own process.

    python tools/ab.py base=blenderraytracer_b200/libbrt.so x=blenderraytracer_b200/libbrt_x.so -- c3:256 c5:64 c4:64
    python tools/ab.py w2=blenderraytracer_b200/libbrt.so,BRT_BVH_WIDTH=2 w8=blenderraytracer_b200/libbrt.so,BRT_BVH_WIDTH=8 -- c3:256
Prints best-of-5 kernel times (CUDA events around brt_render_accumulate) as Msamples/s, and the image checksum of each variant."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, sys, os, hashlib
sys.path.insert(0, %r)
import torch, numpy as np
import blenderraytracer_b200 as brt
from bench import load_workload
out = {}
for spec in sys.argv[1:]:
    name, spp = spec.split(":"); spp = int(spp)
    w = load_workload(name, binary=True)
    W, H = w["W"], w["H"]
    rt = brt.RayTracer(W, H, device=0, seed=1)
    assert rt.loadFromJSON(w.get("blob") or json.dumps(w["scene"]).encode())
    rt.resizeCanvas(W, H)
    rt.updateRenderSettings(dict(samples=spp, maxBounces=w["depth"]))
    rt.directLighting = bool(w.get("direct"))
    rt.setStream(torch.cuda.current_stream().cuda_stream)
    rt._push_params()
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    best = 1e30
    for k in range(6):
        acc.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); rt.renderAccumulate(acc.data_ptr(), 0, spp); e1.record(); torch.cuda.synchronize()
        if k: best = min(best, e0.elapsed_time(e1))
    host = acc.cpu().numpy()
    import zlib
    out[spec] = dict(ms=best, msamples_s=W * H * spp / best / 1e3, mean=float(acc[..., :3].mean().item() / spp),
                     sha=hashlib.sha1(host.tobytes()).hexdigest()[:12], rows=[zlib.crc32(host[y].tobytes()) for y in range(H)])
    rt.close()
print("AB_RESULT " + json.dumps(out))
''' % ROOT

def main():
    args = sys.argv[1:]
    cut = args.index("--")
    variants, specs = [a.split("=", 1) for a in args[:cut]], args[cut + 1:]
    res = {}
    for name, lib in variants:
        lib, *extra = lib.split(",")                      # name=path/to/lib.so[,ENV=VALUE ...]: environment of that variant's process
        env = dict(os.environ, BRT_LIBBRT=os.path.abspath(lib))
        env.update(kv.split("=", 1) for kv in extra)
        p = subprocess.run([sys.executable, "-c", WORKER] + specs, capture_output=True, text=True, env=env, cwd=ROOT)
        line = [l for l in p.stdout.splitlines() if l.startswith("AB_RESULT ")]
        if not line:
            print(name, "FAILED", p.stderr[-1500:]); continue
        res[name] = json.loads(line[0][len("AB_RESULT "):])
    base = variants[0][0]
    print(f"{'variant':<14}" + "".join(f"{s:>26}" for s in specs))
    for name, _ in variants:
        if name not in res: continue
        cells = []
        for s in specs:
            r, b = res[name][s], res[base][s]
            nrows = sum(1 for x, y in zip(r["rows"], b["rows"]) if x != y)   # image rows that differ from the base variant's
            cells.append(f"{r['msamples_s']:9.0f} {100 * (r['msamples_s'] / b['msamples_s'] - 1):+5.1f}% {'=' if r['sha'] == b['sha'] else '~'}{nrows}r")
        print(f"{name:<14}" + "".join(f"{c:>26}" for c in cells))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for v in res.values():
        for r in v.values():
            r.pop("rows", None)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
