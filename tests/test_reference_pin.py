"""Pins the oracle to the reference's OWN output, when that output is available.

tests/golden/reference_vectors.json is produced by running the UNMODIFIED reference (js/ray-tracer.js RayTracer.render) under
Node with Math.random replaced by the oracle's Philox stream — `node baseline/make_fixtures.mjs` (see baseline/README.md).
The build image has no JavaScript engine, so the file cannot be generated there: without it this test SKIPS with an explicit
"parity unpinned" message and DESIGN.md says the same.  With it, every case must match the oracle: per-pixel mean radiance
to 1e-12 relative (JS and C++ doubles agree exactly on + - * / sqrt; Math.tan / pow / exp / sin / cos may differ from glibc
in the last ulp), RGBA8 within 1 LSB."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.oracle import OracleRayTracer

VECTORS = os.path.join(GOLDEN, "reference_vectors.json")
CASES = os.path.join(GOLDEN, "reference_cases.json")


def test_reference_cases_cover_the_second_port_cases():
    """The case list fed to the reference is the one the independent port was checked on (13 cases), seeds and Perlin tables explicit."""
    cases = json.load(open(CASES))
    z = np.load(os.path.join(GOLDEN, "independent_vectors.npz"))
    meta = json.loads(str(z["meta"]))
    assert [c["name"] for c in cases] == [m["name"] for m in meta] and len(cases) >= 13
    for c, m in zip(cases, meta):
        assert (c["W"], c["H"], c["spp"], c["depth"], c["seed"]) == (m["W"], m["H"], m["spp"], m["depth"], m["seed"])
        assert len(c["perm"]) == 256 and sorted(c["perm"]) == list(range(256))
        assert ("preset" in c) != ("scene" in c)


def test_oracle_matches_the_reference_itself():
    if not os.path.exists(VECTORS):
        pytest.skip("PARITY UNPINNED: tests/golden/reference_vectors.json is absent — no JavaScript engine in this image; "
                    "run `node baseline/make_fixtures.mjs` where Node.js exists (baseline/README.md)")
    ref = json.load(open(VECTORS))["cases"]
    for c in json.load(open(CASES)):
        name, W, H = c["name"], c["W"], c["H"]
        want = ref[name]
        rt = OracleRayTracer(W, H, seed=c["seed"], threads=2)
        if "preset" in c:
            rt.loadPreset(c["preset"])
        else:
            assert rt.loadFromJSON(c["scene"])
        rt.setCloudPermutation(np.asarray(c["perm"], np.uint8))
        rt.updateRenderSettings(dict(samples=c["spp"], maxBounces=c["depth"], antiAliasing=c["aa"], toneMapping=c["tonemap"],
                                     exposure=c["exposure"], gamma=c["gamma"], denoising=c["denoise"], denoiseStrength=c["strength"]))
        img = rt.render()
        lin = np.asarray(want["linear"], np.float64).reshape(H, W, 3)
        np.testing.assert_allclose(rt.linear[..., :3], lin, rtol=1e-12, atol=1e-15, err_msg=name)
        rgba = np.asarray(want["rgba"], np.uint8).reshape(H, W, 4)
        assert np.abs(img.astype(int) - rgba.astype(int)).max() <= 1, name
