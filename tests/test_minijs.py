"""baseline/minijs.py — the interpreter that executes the reference's js/*.js for the golden vectors — against known answers of
the ECMAScript semantics the reference relies on (values any JavaScript engine gives; each snippet's expectation is stated next
to it).  If the interpreter mis-evaluated the language, the reference pin (tests/test_reference_pin.py) would prove nothing."""
import math
import os
import sys

import pytest

from conftest import HAVE_REFERENCE, REFERENCE, REFERENCE_JS

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "baseline"))
import minijs as J  # noqa: E402


def run(src):
    """evaluates `src` as a module body; returns the Python value of its variable `out`"""
    interp = J.Interp()
    env = interp.run(src)
    return J.js_to_py(env.get("out")) if not isinstance(env.get("out"), float) else env.get("out")


CASES = [
    # numbers are IEEE doubles
    ("let out = 0.1 + 0.2;", 0.30000000000000004),
    ("let out = 1 / 3;", 1 / 3),
    ("let out = 7 % 3 + (-7) % 3;", 0.0),                          # remainder keeps the dividend's sign: 1 + (-1)
    ("let out = 5 / 0;", math.inf),
    ("let out = Math.sqrt(2) * Math.sqrt(2);", math.sqrt(2) * math.sqrt(2)),
    ("let out = Math.floor(-0.5) + Math.floor(2.7);", 1.0),
    ("let out = Math.max(1, 5, 3) - Math.min(4, -2);", 7.0),
    ("let out = Math.pow(2, 10) + 2 ** 3 ** 2;", 1024.0 + 512.0),   # ** is right-associative
    ("let out = (13 & 7) + (1 << 4) + (-16 >> 2) + (5 | 2) + (6 ^ 3);", 5 + 16 - 4 + 7 + 5),
    ("let out = 1e-6 + 0x10 + .5;", 1e-6 + 16 + 0.5),
    # truthiness, || and && return operands, ?? only skips null / undefined
    ("let out = [0 || 7, 3 || 7, '' || 'd', null ?? 4, 0 ?? 4, 5 && 6, 0 && 6];", [7, 3, "d", 4, 0, 6, 0]),
    ("let a; let out = [a === undefined, typeof a, typeof null, typeof 1, typeof 'x', typeof (() => 1), typeof {}];",
     [True, "undefined", "object", "number", "string", "function", "object"]),
    ("let out = [1 == '1', 1 === '1', null == undefined, null === undefined, NaN === NaN];", [True, False, True, False, False]),
    ("let out = [!!'0', !!0, !![], !!{}, !!NaN];", [True, False, True, True, False]),     # empty array / object are truthy
    ("let out = 'a' + 1 + 2 + ',' + (1 + 2) + `${3 * 4}x`;", "a12,312x"),
    # control flow
    ("let s = 0; for (let i = 0; i < 5; i++) { if (i === 1) continue; if (i === 4) break; s += i; } let out = s;", 5.0),
    ("let s = 0, i = 0; do { s += i; i++; } while (i < 4); let out = s;", 6.0),
    ("let s = ''; for (const x of [1, 2, 3]) s += x; let out = s;", "123"),
    ("function f(k) { switch (k) { case 'a': return 1; case 'b': case 'c': return 2; default: return 3; } } let out = [f('a'), f('b'), f('c'), f('z')];", [1, 2, 2, 3]),
    ("let r = []; switch (2) { case 1: r.push(1); case 2: r.push(2); case 3: r.push(3); break; case 4: r.push(4); } let out = r;", [2, 3]),   # fall-through
    ("let out; try { null.x; out = 'no'; } catch (e) { out = 'caught'; } finally { out += '!'; }", "caught!"),
    ("function f() { try { throw new Error('boom'); } catch (e) { return e.message; } } let out = f();", "boom"),
    ("let out = (function () { return\n 5; })();", None),               # ASI after return
    # functions, closures, arrows, defaults, rest / spread
    ("function mk() { let c = 0; return () => ++c; } const f = mk(); f(); f(); let out = f();", 3.0),
    ("const fs = []; for (let i = 0; i < 3; i++) fs.push(() => i); let out = fs.map(f => f());", [0, 1, 2]),       # per-iteration binding
    ("function f(a, b = a * 2, ...r) { return [a, b, r.length]; } let out = [f(1), f(1, 5, 6, 7)];", [[1, 2, 0], [1, 5, 2]]),
    ("const o = { x: 2, m() { return [1, 2].map(v => v * this.x); } }; let out = o.m();", [2, 4]),            # arrows take `this` lexically
    ("function f() { return this === undefined; } let out = f();", True),
    ("const g = function (a, b) { return this.k + a + b; }; let out = [g.call({ k: 1 }, 2, 3), g.apply({ k: 1 }, [2, 3]), g.bind({ k: 10 }, 1)(2)];", [6, 6, 13]),
    ("let [a, b] = [1, 2]; [a, b] = [b, a]; const { p, q: z } = { p: 5, q: 6 }; let out = [a, b, p, z];", [2, 1, 5, 6]),
    ("let out = Math.max(...[1, 9, 4]) + [...[1, 2], 3].length;", 12.0),
    # classes
    ("""class A { constructor(x) { this.x = x; } get2() { return this.x * 2; } static make(x) { return new A(x + 1); } }
        class B extends A { constructor(x) { super(x + 10); this.y = 1; } get2() { return super.get2() + this.y; } }
        class C extends B {}
        const c = new C(1);
        let out = [new A(2).get2(), A.make(2).x, new B(1).get2(), c.get2(), c instanceof A, c instanceof C, new A(1) instanceof B, C.make(0) instanceof A];""",
     [4, 3, 23, 23, True, True, False, True]),
    ("class V { constructor(x = 0, y = 0) { this.x = x; this.y = y; } add(v) { return new V(this.x + v.x, this.y + v.y); } } let v = new V(1).add(new V(2, 3)); let out = [v.x, v.y];", [3, 3]),
    # objects, arrays, optional chaining, in, Object.keys, Array.isArray
    ("const o = { a: 1, 'b c': 2, 3: 4 }; o.d = o.a + o['b c']; let out = [o.d, o[3], o.zz, 'a' in o, 'q' in o, Object.keys(o).length];", [3, 4, None, True, False, 4]),
    ("const o = { a: null }; let out = [o.a?.b, o.q?.r.s, o?.a, o.f?.()];", [None, None, None, None]),
    ("const a = []; a[3] = 1; a.push(2); let out = [a.length, a[0], a[3], a[4], Array.isArray(a), Array.isArray({})];", [5, None, 1, 2, True, False]),
    ("let out = ['Ab'.toLowerCase(), 'abc'.length, [3, 1].map(x => x * 2), [1, 2, 3].filter(x => x > 1), [1, 2, 3].reduce((s, x) => s + x, 0)];", ["ab", 3, [6, 2], [2, 3], 6]),
    # typed arrays: Float32Array rounds on store, Uint8ClampedArray clamps and rounds half to even
    ("const f = new Float32Array(2); f[0] = 0.1; f[1] = 16777217; let out = [f[0], f[1], f.length];", [0.10000000149011612, 16777216, 2]),
    ("const u = new Uint8ClampedArray(6); u[0] = -5; u[1] = 300; u[2] = 1.5; u[3] = 2.5; u[4] = 254.5; u[5] = NaN; let out = [u[0], u[1], u[2], u[3], u[4], u[5]];", [0, 255, 2, 2, 254, 0]),
    # what napi/raytracer_gpu.mjs needs beyond the reference's subset
    ("function f({ a = 1, b: c = 2, d } = {}) { return [a, c, d]; } let out = [f(), f({ a: 5, b: 6, d: 7 }), f({ a: undefined, b: null })];", [[1, 2, None], [5, 6, 7], [1, None, None]]),
    ("const u = Uint8Array.from([1, 255, 256, -1, 3.9]); const f = new Float64Array([0.1, 2]); let out = [u[0], u[1], u[2], u[3], u[4], u.length, f[0], Uint8Array.name];", [1, 255, 0, 255, 3, 5, 0.1, "Uint8Array"]),
    ("class K {} function g() {} const o = { m() {} }; let out = [new K().constructor.name, g.name, g.bind(null).name, (() => 1).name, o.m.name];", ["K", "g", "bound g", "", "m"]),
    ("let n = 0; const id = setInterval(() => { n++; }, 50); clearInterval(id); let out = [typeof id, n];", ["number", 0]),     # intervals are ticked by the host only
    ("const o = { solid: 1, hdri: 2 }; let out = [o['solid'] ?? 0, o['x'] ?? 0, { a: 1 }['a'], Array.from({ length: 3 }, (_, i) => i * 2)];", [1, 0, 1, [0, 2, 4]]),
    ("globalThis.zz = 5; let out = [zz, globalThis.Math === Math, globalThis.nope?.x, typeof globalThis.setInterval];", [5, True, None, "function"]),
    # BigInt (baseline/run_ref.mjs computes Philox's 32 x 32 -> 64-bit products with it), import(), promise reactions
    ("const p = 0xD2511F53n * BigInt(4000000000); let out = [Number(p >> 32n), Number(p & 0xFFFFFFFFn), typeof p, 7n === 7n, (Number(p >> 32n) ^ 5) >>> 0];", [3286201315, 4002805760, "bigint", True, 3286201318]),
    ("let out; try { out = 1n + 1; } catch (e) { out = e.name; }", "TypeError"),
    ("let out = []; Promise.resolve(1).then(v => { out.push(v); return Promise.resolve(2); }).then(v => out.push(v)).finally(() => out.push('f')); new Promise((_, rej) => rej(3)).catch(v => out.push(v));", [1, 2, "f", 3]),
    # async / await over immediately resolved promises (the reference yields with setTimeout between rows)
    ("let out = 0; async function r() { await new Promise(res => setTimeout(res, 1)); out = 7; return 3; } r();", 7.0),
]


@pytest.mark.parametrize("src,want", CASES, ids=[str(i) for i in range(len(CASES))])
def test_ecmascript_known_answers(src, want):
    got = run(src)
    if isinstance(want, float) or isinstance(got, float):
        assert got == want or (got != got and want != want), (src, got)
        assert isinstance(got, float) and math.copysign(1, got) == math.copysign(1, want)
    else:
        assert got == want, (src, got)


def test_unsupported_syntax_fails_loudly():
    """the interpreter never guesses: what it does not implement is an error"""
    for src in ("let out = /ab+/.test('x');", "label: for (;;) {}", "let out = 1 <=> 2;"):
        with pytest.raises((J.JSSyntaxError, J.JSRuntimeError, J.JSThrow)):
            run(src)


def test_modules_import_export(tmp_path):
    (tmp_path / "a.js").write_text("class P { constructor() { this.v = 41; } }\nconst K = 1;\nexport { P, K };\n")
    (tmp_path / "b.js").write_text("import { P, K as one } from './a.js';\nexport class Q extends P { get() { return this.v + one; } }\nexport function f() { return new Q().get(); }\n")
    interp = J.Interp()
    ex = interp.load_module(str(tmp_path / "b.js"))
    assert interp.call(ex["f"], J.UNDEF, []) == 42.0
    assert set(ex) == {"Q", "f"}


def test_host_modules_and_import_meta(tmp_path):
    """`import { createRequire } from 'node:module'` resolves to what the embedding host registered; import.meta.url names the file"""
    (tmp_path / "m.mjs").write_text("import { createRequire } from 'node:module';\nconst require = createRequire(import.meta.url);\n"
                                    "export const got = require('./x.node');\nexport const url = import.meta.url;\n")
    interp = J.Interp()
    seen = []
    interp.builtin_modules["node:module"] = {"createRequire": J.native(lambda t, a: (seen.append(a[0]), J.native(lambda t2, a2: "addon:" + a2[0]))[1])}
    ex = interp.load_module(str(tmp_path / "m.mjs"))
    assert ex["got"] == "addon:./x.node" and ex["url"] == "file://" + str(tmp_path / "m.mjs") and seen == [ex["url"]]


@pytest.mark.skipif(not HAVE_REFERENCE, reason="no reference checkout on this machine")
def test_reference_math_module_under_the_interpreter():
    """the reference's own Vec3 (js/math.js:6-31), executed from its source: a few identities with exact expectations"""
    interp = J.Interp()
    ex = interp.load_module(os.path.join(REFERENCE_JS, "math.js"))
    Vec3 = ex["Vec3"]
    v = interp.construct(Vec3, [1.0, 2.0, 2.0])
    assert interp.call(v.get("length"), v, []) == 3.0
    n = interp.call(v.get("normalize"), v, [])
    assert (n.get("x"), n.get("y"), n.get("z")) == (1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0)
    w = interp.call(v.get("cross"), v, [interp.construct(Vec3, [0.0, 0.0, 1.0])])
    assert (w.get("x"), w.get("y"), w.get("z")) == (2.0, -1.0, 0.0)
    z = interp.construct(Vec3, [])
    zn = interp.call(z.get("normalize"), z, [])
    assert (zn.get("x"), zn.get("y"), zn.get("z")) == (0.0, 0.0, 0.0)                  # zero vector stays zero (math.js:18)
