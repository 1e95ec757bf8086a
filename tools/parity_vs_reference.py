"""Per-case numbers of the CUDA path (sampler = reference) against the reference's own vectors (tests/golden/reference_vectors.json,
written from the unmodified js/*.js by baseline/make_fixtures_minijs.py): median / max |linear error|, share of pixels within
1 and 2 LSB, and — for the primary-hit gate — nothing else is needed: the same file pins the oracle bit for bit.
    python tools/parity_vs_reference.py > gpurun_out/parity_vs_reference.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import blenderraytracer_b200 as brt

G = os.path.join(ROOT, "tests", "golden")
doc = json.load(open(os.path.join(G, "reference_vectors.json")))
cases = json.load(open(os.path.join(G, "reference_cases.json"))) + json.load(open(os.path.join(G, "reference_cases_extra.json")))
out = {"generator_of_the_vectors": doc["generator"], "cases": {}}
for c in cases:
    W, H, want = c["W"], c["H"], doc["cases"][c["name"]]
    rt = brt.RayTracer(W, H, seed=c["seed"])
    if "preset" in c: rt.loadPreset(c["preset"])
    else: assert rt.loadFromJSON(c["scene"])
    rt.setCloudPermutation(np.asarray(c["perm"], np.uint8))
    rt.updateRenderSettings(dict(samples=c["spp"], maxBounces=c["depth"], antiAliasing=c["aa"], toneMapping=c["tonemap"], exposure=c["exposure"],
                                 gamma=c["gamma"], denoising=c["denoise"], denoiseStrength=c["strength"]))
    rt.sampler = "reference"
    img = rt.render(want_linear=True)
    lin = np.asarray(want["linear"], np.float64).reshape(H, W, 3)
    rgba = np.asarray(want["rgba"], np.uint8).reshape(H, W, 4)
    err = np.abs(rt.linearMean[..., :3] - lin)
    d = np.abs(img[..., :3].astype(int) - rgba[..., :3].astype(int)).max(axis=-1)
    out["cases"][c["name"]] = dict(pixels=W * H, spp=c["spp"], depth=c["depth"], median_abs_linear_err=float(np.median(err)), p99_abs_linear_err=float(np.quantile(err, 0.99)),
                                   max_abs_linear_err=float(err.max()), within_1_lsb=float((d <= 1).mean()), within_2_lsb=float((d <= 2).mean()), identical_rgba8=float((d == 0).mean()))
    rt.close()
print(json.dumps(out, indent=1))
