"""node_shims.py — the few Node.js built-ins baseline/run_ref.mjs and baseline/make_fixtures.mjs import (node:fs, node:path, node:url,
process), for executing those harnesses under baseline/minijs.py in an image without Node: tests/test_node_kit.py runs the Node
kit's own code that way and requires it to reproduce the committed vectors.  Only what the kit calls is provided."""
import os
import shutil

import minijs as J


def install(interp, argv=("node", "<minijs>")):
    S = J.to_str
    def n(f, name=""): return J.native(lambda t, a: f(*a), name)
    fs = J.JSObject(J.OBJECT_PROTO, {
        "readFileSync": n(lambda p, enc=J.UNDEF: open(S(p), encoding="utf-8").read(), "readFileSync"),
        "writeFileSync": n(lambda p, data: (open(S(p), "w", encoding="utf-8").write(S(data)), J.UNDEF)[1], "writeFileSync"),
        "existsSync": n(lambda p: os.path.exists(S(p)), "existsSync"),
        "mkdirSync": n(lambda p, opts=J.UNDEF: (os.makedirs(S(p), exist_ok=True), J.UNDEF)[1], "mkdirSync"),
        "readdirSync": n(lambda p: J.JSArray(sorted(os.listdir(S(p)))), "readdirSync"),
        "copyFileSync": n(lambda a, b: (shutil.copyfile(S(a), S(b)), J.UNDEF)[1], "copyFileSync"),
    })
    path = J.JSObject(J.OBJECT_PROTO, {
        "resolve": J.native(lambda t, a: os.path.abspath(os.path.join(*[S(x) for x in a])), "resolve"),
        "join": J.native(lambda t, a: os.path.normpath(os.path.join(*[S(x) for x in a])), "join"),
        "dirname": n(lambda p: os.path.dirname(S(p)), "dirname"),
        "basename": n(lambda p: os.path.basename(S(p)), "basename"),
    })
    url = {
        "pathToFileURL": n(lambda p: J.JSObject(J.OBJECT_PROTO, {"href": "file://" + os.path.abspath(S(p))}), "pathToFileURL"),
        "fileURLToPath": n(lambda u: S(u)[7:] if S(u).startswith("file://") else S(u), "fileURLToPath"),
    }
    interp.builtin_modules["node:fs"] = {"default": fs}
    interp.builtin_modules["node:path"] = {"default": path}
    interp.builtin_modules["node:url"] = url
    def exit_(t, a): raise SystemExit(int(J.to_num(a[0])) if a else 0)
    interp.globals.vars["process"] = J.JSObject(J.OBJECT_PROTO, {"argv": J.JSArray([str(x) for x in argv]), "version": "v0.0.0-minijs", "exit": J.native(exit_, "exit")})
    return interp
