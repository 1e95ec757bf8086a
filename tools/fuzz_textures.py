"""Differential fuzz of the procedural textures: random textures (js/textures.js: SolidColor / Checker / Noise / Marble / Wood with random
colours, scales and Perlin tables) evaluated at random points — incl. negative, huge and tiny coordinates — by the reference's own
classes (executed by baseline/minijs.py) and by the oracle's restatement: the same bits.
    python tools/fuzz_textures.py [--seed 1] [--n 200] [--ref /root/reference]"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "baseline")):
    sys.path.insert(0, p)


def run(seed, n, ref="/root/reference", points=40, verbose=False):
    import numpy as np
    import make_fixtures_minijs as M
    import make_texture_fixtures_minijs as T
    from oracle.oracle import OracleRayTracer
    sys.setrecursionlimit(20000)
    js = os.path.join(ref, "js")
    interp, RayTracer, Vec3 = M.load_reference(js)
    tex_ex = interp.load_module(os.path.join(js, "textures.js"))
    r = random.Random(seed)
    bad = []
    for k in range(n):
        kind = r.choice(["solid", "checker", "noise", "marble", "wood"])
        t = dict(name=f"t{k}", kind=kind, odd=[round(r.random(), 3) for _ in range(3)], even=[round(r.random(), 3) for _ in range(3)],
                 scale=r.choice([0.1, 0.5, 1.0, 1.5, 2.5, 4.0, 10.0, 37.0, -2.0, 0.0]), perm_seed=(r.randrange(1, 10 ** 6) if kind in ("noise", "marble", "wood") else None))
        if kind in ("noise", "marble", "wood"): t["odd"] = t["even"] = [1, 1, 1]
        tex = T.make_texture(interp, tex_ex, Vec3, t)
        pts = []
        for _ in range(points):
            m = r.choice([1.0, 1.0, 10.0, 1e3, 1e-4, 256.0])
            pts.append([r.uniform(-6, 6) * m, r.uniform(-6, 6) * m, r.uniform(-6, 6) * m])
        pts += [[0.0, 0.0, 0.0], [-0.0, 255.999999, -256.0], [1e9, -1e9, 0.5]]
        want = []
        for p in pts:
            v = interp.call(tex.get("value"), tex, [0.0, 0.0, interp.construct(Vec3, [float(p[0]), float(p[1]), float(p[2])])])
            want.append([v.get("x"), v.get("y"), v.get("z")])
        rt = OracleRayTracer(8, 8)
        sc = rt.scene
        sc.add_sphere((0, 0, 0), 1.0, ("lambertian", [1, 1, 1], 0.0))
        sc.set_object_texture(0, t["kind"], tuple(t["odd"]), tuple(t["even"]), t["scale"], perm256=(T.perm_of(t["perm_seed"]) if t["perm_seed"] is not None else None))
        got = np.array([sc.texture_value(0, p) for p in pts])
        w = np.asarray(want, np.float64)
        if not np.array_equal(got, w, equal_nan=True):
            i = int(np.argwhere(~((got == w) | (np.isnan(got) & np.isnan(w))))[0][0])
            bad.append((k, t, pts[i], got[i].tolist(), w[i].tolist()))
    if verbose:
        for b in bad[:5]: print("---", b)
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--ref", default=os.environ.get("BRT_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    bad = run(args.seed, args.n, args.ref, verbose=True)
    print(f"{args.n} textures, {len(bad)} disagreements")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
