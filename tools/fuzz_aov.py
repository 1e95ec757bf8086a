"""Differential fuzz of PRIMARY VISIBILITY: random scenes (tools/fuzz_ingest.py's generator) through the reference's own camera.getRay +
World.hit at every pixel centre (js/*.js executed by baseline/minijs.py) and through the oracle: object ID, triangle ID, t, normal and
frontFace must be the same (NaN where the reference has NaN).  Needs a checkout of the reference (never copied).

    python tools/fuzz_aov.py [--seed 1] [--n 60] [--emit tests/golden/reference_fuzz_aov_vectors.json]"""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)


def run(seed, n, ref="/root/reference", verbose=False, emit=None):
    import numpy as np
    import fuzz_ingest
    import make_aov_fixtures_minijs as A
    from oracle.oracle import OracleRayTracer
    from test_reference_pin import reference_aov
    sys.setrecursionlimit(20000)
    r = random.Random(seed)
    bad, done = [], 0
    for k in range(n):
        scene = fuzz_ingest.gen_scene(r)
        if scene.get("camera"): scene["camera"].pop("resolution", None)
        c = dict(name=f"aov{k}", W=r.choice([16, 20, 24]), H=r.choice([10, 12, 16]), scene=scene)
        try:
            rec = A.aov_of_case(os.path.join(ref, "js"), c)
        except RuntimeError:
            continue
        done += 1
        if emit is not None: emit.append(rec)
        rt = OracleRayTracer(c["W"], c["H"])
        assert rt.loadFromJSON(scene)
        got, want = rt.primary_aov(), reference_aov(rec)
        hit = want["obj_id"] >= 0
        why = None
        if not np.array_equal(got["obj_id"], want["obj_id"]): why = f"object IDs differ in {int((got['obj_id'] != want['obj_id']).sum())} pixels"
        elif not np.array_equal(got["tri_id"], want["tri_id"]): why = "triangle IDs differ"
        elif not np.array_equal(got["t"][hit], want["t"][hit], equal_nan=True): why = "t differs"
        elif not np.array_equal(got["normal"][hit], want["normal"][hit], equal_nan=True): why = "normal differs"
        elif not np.array_equal(got["front_face"][hit], want["front_face"][hit]): why = "frontFace differs"
        if why: bad.append((k, why, c))
    if verbose:
        for k, why, c in bad[:5]: print(f"--- case {k}: {why}\n{json.dumps(c)}")
    return bad, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--n", type=int, default=60)
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    ap.add_argument("--emit", default="")
    args = ap.parse_args()
    emit = [] if args.emit else None
    bad, done = run(args.seed, args.n, args.ref, verbose=True, emit=emit)
    if args.emit and not bad:
        json.dump({"generator": f"tools/fuzz_aov.py --seed {args.seed} --n {args.n}: camera.getRay + world.hit of the unmodified reference executed by baseline/minijs.py", "cases": emit}, open(args.emit, "w"))
        print("wrote", args.emit)
    print(f"{done} scenes, {len(bad)} disagreements")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
