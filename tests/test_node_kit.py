"""The Node kit (baseline/run_ref.mjs, baseline/make_fixtures.mjs: the harness for whoever has Node.js) EXECUTED here: the build image has
no Node, so its JavaScript — BigInt Philox, fake canvas, Math.random / toneMap / gammaCorrect hooks, dynamic import of the reference —
runs under baseline/minijs.py with the few Node built-ins it imports shimmed (baseline/node_shims.py), and must reproduce the
committed vectors that the Python harness (baseline/make_fixtures_minijs.py) wrote: two independent harnesses, one answer."""
import json
import os
import sys

import pytest

from conftest import GOLDEN, HAVE_REFERENCE, REFERENCE_JS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import minijs as J  # noqa: E402
import node_shims  # noqa: E402
import make_fixtures_minijs as M  # noqa: E402


def kit():
    sys.setrecursionlimit(20000)
    interp = node_shims.install(J.Interp())
    return interp, interp.load_module(os.path.join(ROOT, "baseline", "run_ref.mjs"))


def test_kit_philox_equals_the_python_harness():
    """the kit's BigInt Philox4x32-10 against the Python harness's (itself equal to the oracle's stream, tests/test_oracle_kat.py)"""
    interp, ex = kit()
    for pixel, sample, block, lo, hi in ((0, 0, 0, 1, 0), (5, 9, 2, 0x9ABCDEF0, 0x12345678), (2 ** 31 + 7, 4095, 11, 0xFFFFFFFF, 0xFFFFFFFF), (8294399, 3, 1, 505, 0)):
        got = J.js_to_py(interp.call(ex["philoxBlock"], J.UNDEF, [float(pixel), float(sample), float(block), float(lo), float(hi)]))
        assert [int(v) for v in got] == M.philox_block(pixel, sample, block, lo, hi), (pixel, sample, block)


def test_make_fixtures_mjs_parses():
    ast = J.Parser(open(os.path.join(ROOT, "baseline", "make_fixtures.mjs")).read(), "make_fixtures.mjs").program()
    assert len(ast) > 10


@pytest.mark.skipif(not HAVE_REFERENCE, reason="no reference checkout on this machine")
def test_kit_reproduces_the_committed_vectors():
    interp, ex = kit()
    ref = interp.call(ex["loadReference"], J.UNDEF, [REFERENCE_JS])
    doc = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))["cases"]
    cases = {c["name"]: c for fn in ("reference_cases.json", "reference_cases_extra.json") for c in json.load(open(os.path.join(GOLDEN, fn)))}
    for name in ("bg_procedural_sky", "preset_glass", "aa_none_linear", "denoise", "bg_solid", "mesh_orthographic_hdri_stochastic_linear"):
        got = J.js_to_py(interp.call(ex["renderSeeded"], J.UNDEF, [ref, J.py_to_js(json.loads(json.dumps(cases[name])))]))
        want = doc[name]
        assert got["rgba"] == want["rgba"] and got["linear"] == want["linear"] and got["float"] == want["float"], name
