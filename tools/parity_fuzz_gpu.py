"""The CUDA path (sampler = reference) on the random scenes of tests/golden/reference_fuzz_vectors.json against the reference's own
outputs: per case the share of pixels with identical RGBA8 and the median |linear error|.
    python tools/parity_fuzz_gpu.py > gpurun_out/parity_fuzz_gpu.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import blenderraytracer_b200 as brt


def compare(doc):
    rows = []
    for e in doc["cases"]:
        c = e["case"]
        W, H = c["W"], c["H"]
        rt = brt.RayTracer(W, H, seed=c["seed"])
        ok = rt.loadFromJSON(c["scene"])
        assert ok, c["name"]                                          # the reference accepted it
        rt.setCloudPermutation(np.asarray(c["perm"], np.uint8))
        rt.updateRenderSettings(dict(samples=c["spp"], maxBounces=c["depth"], antiAliasing=c["aa"], toneMapping=c["tonemap"], exposure=c["exposure"],
                                     gamma=c["gamma"], denoising=c["denoise"], denoiseStrength=c["strength"]))
        rt.sampler = "reference"
        img = rt.render(want_linear=True)
        lin = np.asarray(e["linear"], np.float64).reshape(H, W, 3)
        rgba = np.asarray(e["rgba"], np.uint8).reshape(H, W, 4)
        d = np.abs(img[..., :3].astype(int) - rgba[..., :3].astype(int)).max(axis=-1)
        both = np.isfinite(lin) & np.isfinite(rt.linearMean[..., :3])
        err = np.abs(rt.linearMean[..., :3] - lin)[both]
        rows.append(dict(name=c["name"], pixels=W * H, identical=float((d == 0).mean()), within_2=float((d <= 2).mean()),
                         median_err=float(np.median(err)) if err.size else 0.0, nonfinite_ref=int((~np.isfinite(lin)).sum()),
                         nonfinite_gpu=int((~np.isfinite(rt.linearMean[..., :3])).sum())))
        rt.close()
    return rows


if __name__ == "__main__":
    doc = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_fuzz_vectors.json")))
    rows = compare(doc)
    for r in rows: print(r, file=sys.stderr)
    tot = sum(r["pixels"] for r in rows)
    print(json.dumps(dict(cases=rows, identical_share=sum(r["identical"] * r["pixels"] for r in rows) / tot), indent=1))
