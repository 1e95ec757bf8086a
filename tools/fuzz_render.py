"""Differential fuzz of the RENDER path: random small scenes (tools/fuzz_ingest.py's generator: every primitive and material
kind, degenerate values, any camera and background) rendered by the reference's OWN RayTracer.render() (js/*.js executed by
baseline/minijs.py, Math.random fed from the shared Philox stream) and by the oracle: per-pixel mean radiance, floatData and RGBA8
must be the same bits.  Needs a checkout of the reference (never copied).

    python tools/fuzz_render.py [--seed 1] [--n 40] [--ref /root/reference]"""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def gen_case(r, k):
    import fuzz_ingest
    scene = fuzz_ingest.gen_scene(r)
    cam = scene.get("camera") or {}
    cam.pop("resolution", None)                                       # the frame stays small
    if cam: scene["camera"] = cam
    perm = list(range(256)); r.shuffle(perm)
    return dict(name=f"fuzz{k}", W=r.choice([12, 16, 20]), H=r.choice([8, 12]), spp=r.choice([1, 2, 3]), depth=r.choice([1, 3, 6, 10]), seed=r.randrange(1, 1 << 30),
                aa=r.choice(["supersampling", "supersampling", "stochastic", "none", "whatever"]), tonemap=r.choice(["reinhard", "aces", "linear", "other"]),
                exposure=r.choice([1.0, 0.5, 2.0]), gamma=r.choice([2.2, 1.0, 1.8]), denoise=r.random() < 0.25, strength=r.choice([0.5, 1.0, 0.1]),
                scene=scene, perm=perm)


def run(seed, n, ref="/root/reference", verbose=False, emit=None):
    import numpy as np
    import make_fixtures_minijs as M
    from test_reference_pin import oracle_render
    sys.setrecursionlimit(20000)
    r = random.Random(seed)
    bad, done = [], 0
    for k in range(n):
        c = gen_case(r, k)
        interp, RayTracer, Vec3 = M.load_reference(os.path.join(ref, "js"))
        try:
            got = M.render_seeded(interp, RayTracer, Vec3, c)
        except RuntimeError:                                          # the reference refused the scene (loadFromJSON returned false)
            continue
        done += 1
        if emit is not None: emit.append(dict(case=c, rgba=got["rgba"], linear=got["linear"]))
        W, H = c["W"], c["H"]
        rt, img = oracle_render(c)
        lin = np.asarray(got["linear"], np.float64).reshape(H, W, 3)
        fdat = np.asarray(got["float"], np.float64).reshape(H, W, 3).astype(np.float32)
        rgba = np.asarray(got["rgba"], np.uint8).reshape(H, W, 4)
        why = None
        if not np.array_equal(rt.linear[..., :3], lin, equal_nan=True): why = f"linear differs in {int((~((rt.linear[..., :3] == lin) | ((rt.linear[..., :3] != rt.linear[..., :3]) & (lin != lin)))).sum())} values"
        elif not np.array_equal(rt.floatData[..., :3], fdat, equal_nan=True): why = "floatData differs"
        elif not np.array_equal(img, rgba): why = f"RGBA8 differs in {int((img != rgba).sum())} bytes"
        if why: bad.append((k, why, c))
    if verbose:
        for k, why, c in bad[:5]:
            print(f"--- case {k}: {why}\n{json.dumps({kk: v for kk, v in c.items() if kk != 'perm'})}")
    return bad, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    ap.add_argument("--emit", default="", help="write the cases and the reference's outputs to this JSON file (tests/golden/reference_fuzz_vectors.json: what the GPU test compares with)")
    args = ap.parse_args()
    emit = [] if args.emit else None
    bad, done = run(args.seed, args.n, args.ref, verbose=True, emit=emit)
    if args.emit and not bad:
        json.dump({"generator": f"tools/fuzz_render.py --seed {args.seed} --n {args.n}: the unmodified reference js/*.js executed by baseline/minijs.py", "cases": emit}, open(args.emit, "w"))
        print("wrote", args.emit)
    print(f"{done} scenes rendered by both, {len(bad)} disagreements")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
