#!/usr/bin/env python
"""make_fixtures_minijs.py — golden vectors for the oracle from the reference's OWN source, without Node.

Does exactly what baseline/run_ref.mjs + baseline/make_fixtures.mjs do under Node.js, but executes the reference's unmodified
js/*.js through baseline/minijs.py (a small interpreter for the JavaScript subset those files use), because the build image has
no JavaScript engine:

  * loads <ref>/js/ray-tracer.js as an ES module (and through its imports math.js, world.js, geometry.js, materials.js,
    camera.js, noise.js, post-processor.js, scene-loader.js, lights.js, textures.js) — nothing of the reference is edited, copied
    into the repo or re-implemented;
  * gives RayTracer a fake canvas ({width, height, getContext -> {createImageData, putImageData}}) and window.renderCancelled;
  * replaces Math.random (js/math.js:21-31, js/materials.js:62) by the oracle's Philox4x32-10 stream — counter (pixel, sample,
    block, 'BRT1'), key = seed, uniforms = top 24 bits / 2^24, restarted at every getAntiAliasSample(i, j, s) call
    (js/ray-tracer.js:203), pixel = (H-1-j)*W + i — so the reference draws the numbers oracle/brt_oracle.cpp draws, in ITS order;
  * calls RayTracer.render() (js/ray-tracer.js:166-281) and captures, per pixel, the mean radiance handed to toneMap (:209), the
    result of gammaCorrect (:210) and imageData.data;
  * writes tests/golden/reference_vectors.json for the 13 cases of tests/golden/reference_cases.json and the 7 BASELINE-shaped
    cases of tests/golden/reference_cases_extra.json — the file tests/test_reference_pin.py compares the oracle with.

    python baseline/make_fixtures_minijs.py [--ref /root/reference] [--only name,name]

Differences from a run under Node are confined to the C library: Math.pow / exp / sin / cos / tan / atan2 / acos come from glibc
here and from V8's fdlibm port there (last-ulp differences possible); + - * / and Math.sqrt are exact in both.
"""
import argparse
import json
import math
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import minijs as J  # noqa: E402

M0, M1, W0, W1, TAG = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0x42525431


def philox_block(pixel, sample, block, lo, hi):
    """Philox4x32-10 (Random123), as oracle/brt_oracle.cpp and baseline/run_ref.mjs"""
    c0, c1, c2, c3, k0, k1 = pixel & 0xFFFFFFFF, sample & 0xFFFFFFFF, block & 0xFFFFFFFF, TAG, lo, hi
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [c0, c1, c2, c3]


class PhiloxStream:
    def __init__(self, seed):
        self.lo, self.hi = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
        self.restart(0, 0)
    def restart(self, pixel, sample):
        self.pixel, self.sample, self.block, self.buf = pixel, sample, 0, []
    def next(self):
        if not self.buf:
            self.buf = philox_block(self.pixel, self.sample, self.block, self.lo, self.hi); self.block += 1
        return float(self.buf.pop(0) >> 8) / 16777216.0


def fake_canvas(interp, W, H):
    def create_image_data(this, a):
        w, h = int(a[0]), int(a[1])
        return J.JSObject(J.OBJECT_PROTO, {"width": float(w), "height": float(h), "data": J.JSTyped("u8c", w * h * 4)})
    ctx = J.JSObject(J.OBJECT_PROTO, {"createImageData": J.native(create_image_data), "putImageData": J.native(lambda t, a: J.UNDEF)})
    return J.JSObject(J.OBJECT_PROTO, {"width": float(W), "height": float(H), "style": J.JSObject(J.OBJECT_PROTO),
                                         "getContext": J.native(lambda t, a: ctx)})


def load_reference(ref_js_dir):
    interp = J.Interp()
    ex = interp.load_module(os.path.join(ref_js_dir, "ray-tracer.js"))
    math_ex = interp.load_module(os.path.join(ref_js_dir, "math.js"))
    return interp, ex["RayTracer"], math_ex["Vec3"]


def method(interp, obj, name, *args):
    return interp.call(obj.get(name), obj, list(args))


def build_case(interp, RayTracer, Vec3, c):
    rt = interp.construct(RayTracer, [fake_canvas(interp, c["W"], c["H"])])
    if "preset" in c:
        method(interp, rt, "loadPreset", c["preset"])
    elif not J.truthy(method(interp, rt, "loadFromJSON", J.py_to_js(json.loads(json.dumps(c["scene"]))))):
        raise RuntimeError("reference loadFromJSON returned false: " + "; ".join(interp.console_lines[-3:]))
    method(interp, rt, "updateRenderSettings", J.py_to_js(dict(
        maxBounces=c["depth"], samples=c["spp"], gamma=c.get("gamma", 2.2), exposure=c.get("exposure", 1.0), toneMapping=c.get("tonemap", "reinhard"),
        antiAliasing=c.get("aa", "supersampling"), denoising=bool(c.get("denoise")), denoiseStrength=c.get("strength", 0.5))))
    bg = (c.get("scene") or {}).get("background")
    world = rt.get("world")
    if bg and bg.get("type") in ("solid", "hdri"):
        # Deviation D1, handled exactly as baseline/run_ref.mjs does: js/scene-loader.js:43,45 bind the background
        # FACTORY instead of calling it; the harness installs the intended background the way js/ray-tracer.js:573-576 does.
        col = bg.get("color") or [0.1, 0.1, 0.1]
        if bg["type"] == "solid":
            world.set("background", method(interp, world, "solidBackground", interp.construct(Vec3, [float(col[0]), float(col[1]), float(col[2])])))
        else:
            world.set("background", method(interp, world, "hdriBackground"))
    noise = world.get("cloudNoise")
    if c.get("perm") and isinstance(noise, J.JSObject):                # world.cloudNoise.p is random per World (js/noise.js:7-17): an input here
        p = noise.get("p")
        for i in range(256):
            J.set_member(p, float(i), float(c["perm"][i])); J.set_member(p, float(256 + i), float(c["perm"][i]))
    return rt


def render_seeded(interp, RayTracer, Vec3, c):
    rt = build_case(interp, RayTracer, Vec3, c)
    W, H = c["W"], c["H"]
    stream = PhiloxStream(int(c["seed"]))
    proto = rt.proto
    orig_aa, orig_tm, orig_gc = proto.get("getAntiAliasSample"), proto.get("toneMap"), proto.get("gammaCorrect")
    linear, fdat = [0.0] * (W * H * 3), [0.0] * (W * H * 3)
    state = {"cur": 0}
    def aa(this, a):
        i, j, s = int(a[0]), int(a[1]), int(a[2])
        state["cur"] = (H - 1 - j) * W + i
        stream.restart(state["cur"], s)
        return interp.call(orig_aa, this, a)
    def tm(this, a):
        col = a[0]; k = state["cur"] * 3
        linear[k], linear[k + 1], linear[k + 2] = col.get("x"), col.get("y"), col.get("z")
        return interp.call(orig_tm, this, a)
    def gc(this, a):
        r = interp.call(orig_gc, this, a); k = state["cur"] * 3
        fdat[k], fdat[k + 1], fdat[k + 2] = r.get("x"), r.get("y"), r.get("z")
        return r
    rt.set("getAntiAliasSample", J.native(aa)); rt.set("toneMap", J.native(tm)); rt.set("gammaCorrect", J.native(gc))
    math_obj = interp.globals.vars["Math"]
    saved = math_obj.get("random")
    math_obj.set("random", J.native(lambda t, a: stream.next()))
    interp.globals.vars["window"].set("renderCancelled", False)
    try:
        if c.get("rect"):
            return render_rect(interp, rt, Vec3, c, linear, fdat)
        interp.call(rt.get("render"), rt, [J.native(lambda t, a: J.UNDEF)])
    finally:
        math_obj.set("random", saved)
    data = rt.get("imageData").get("data").items
    return {"rgba": [int(v) for v in data], "linear": linear, "float": fdat}


def render_rect(interp, rt, Vec3, c, linear, fdat):
    """A window of a frame too large to run whole under the interpreter (the BASELINE configs at 1920x1080): the body of the
    reference's pixel loop (js/ray-tracer.js:199-228) for the pixels of c["rect"] = [x0, y0, x1, y1] (row 0 = top), every step a call
    of the reference's OWN method — getAntiAliasSample, camera.getRay, rayColor, Vec3.add / div, toneMap, gammaCorrect — and the
    Math.floor(c * 255) store through a Uint8ClampedArray (:226-228).  Returns the crop only."""
    x0, y0, x1, y1 = c["rect"]
    W, H = c["W"], c["H"]
    cam = rt.get("camera")
    n = rt.get("samples") if rt.get("antiAliasing") != "none" else 1.0                # :201
    out_lin, out_f, out_rgba = [], [], J.JSTyped("u8c", (x1 - x0) * (y1 - y0) * 4)
    k = 0
    for y in range(y0, y1):
        j = H - 1 - y
        for i in range(x0, x1):
            color = interp.construct(Vec3, [0.0, 0.0, 0.0])
            for smp in range(int(n)):
                sample = method(interp, rt, "getAntiAliasSample", float(i), float(j), float(smp))
                ray = method(interp, cam, "getRay", sample.get("u"), sample.get("v"))
                color = method(interp, color, "add", method(interp, rt, "rayColor", ray, rt.get("maxBounces")))
            color = method(interp, color, "div", n)
            p = ((H - 1 - j) * W + i) * 3
            col = method(interp, rt, "toneMap", color)               # the wrappers record linear[] / fdat[] at the full-frame index
            col = method(interp, rt, "gammaCorrect", col)
            out_lin += linear[p:p + 3]; out_f += fdat[p:p + 3]
            for ch, key in enumerate(("x", "y", "z")):
                out_rgba.store(k + ch, math.floor(col.get(key) * 255) if col.get(key) == col.get(key) else col.get(key))
            out_rgba.store(k + 3, 255.0)
            k += 4
    return {"rgba": [int(v) for v in out_rgba.items], "linear": out_lin, "float": out_f}


def expand_case(c):
    """a case may name its scene by generator (tools/gen_scenes.py, deterministic) instead of carrying it: {"gen": [name, kwargs]}"""
    if "gen" in c and "scene" not in c:
        sys.path.insert(0, ROOT)
        from tools import gen_scenes
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            c = dict(c, scene=getattr(gen_scenes, c["gen"][0])(**c["gen"][1]))
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"), help="checkout of Shinzef/BlenderRayTracer (its js/ is executed, never copied)")
    ap.add_argument("--only", default="", help="comma-separated case names")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_vectors.json"))
    args = ap.parse_args()
    js_dir = os.path.join(args.ref, "js")
    if not os.path.isdir(js_dir):
        sys.exit(f"{js_dir}: no reference checkout (pass --ref)")
    sys.setrecursionlimit(20000)
    cases = []
    for fn in ("reference_cases.json", "reference_cases_extra.json", "reference_cases_fullsize.json"):      # 13 second-port cases, BASELINE-shaped extras, full-size windows
        path = os.path.join(ROOT, "tests", "golden", fn)
        if os.path.exists(path):
            cases += json.load(open(path))
    only = set(filter(None, args.only.split(",")))
    out = {"generator": "baseline/make_fixtures_minijs.py: the unmodified reference js/*.js executed by baseline/minijs.py (Python floats = IEEE doubles; libm = glibc)",
           "cases": {}}
    if only and os.path.exists(args.out):
        out["cases"] = json.load(open(args.out)).get("cases", {})
    for c in cases:
        if only and c["name"] not in only:
            continue
        c = expand_case(c)
        t0 = time.time()
        interp, RayTracer, Vec3 = load_reference(js_dir)                 # a fresh module graph per case, as a page load
        r = render_seeded(interp, RayTracer, Vec3, c)
        out["cases"][c["name"]] = {"W": c["W"], "H": c["H"], **r}
        print(f"{c['name']}: {c['W']}x{c['H']}, mean linear {sum(r['linear']) / len(r['linear']):.6f}, {time.time() - t0:.1f} s", flush=True)
    with open(args.out, "w") as f:
        json.dump(out, f)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
