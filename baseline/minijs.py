"""minijs — a small interpreter for the JavaScript subset the reference (Shinzef/BlenderRayTracer, js/*.js) is written in.

WHY: the build image has no JavaScript engine (no node / deno / quickjs / ...), so the reference cannot run there and the CPU
oracle (oracle/) could only be pinned by transcriptions of the reference.  This interpreter executes the reference's UNMODIFIED
source files — ES modules, classes with extends / super / static, closures and arrow functions, template literals, for-of,
switch, try / catch, optional chaining, async / await over immediately-resolved promises — with JavaScript's Number semantics
(IEEE doubles: Python floats), so `baseline/make_fixtures_minijs.py` can produce tests/golden/reference_vectors.json from the
reference's own text.  It is test infrastructure (never imported by the product) and deliberately small: only what those files
use is implemented, and anything else raises JSSyntaxError / JSRuntimeError instead of guessing.  Math.sqrt / + - * / are exact
in both worlds; Math.pow / exp / sin / cos / tan / atan2 / acos come from the C library here and from V8's fdlibm port under
Node, which may differ in the last ulp (the oracle, C++ on the same libm, sees what this interpreter sees).

Implementation: tokenizer -> Pratt parser -> AST compiled once into Python closures (`fn(env) -> value`).
"""
from __future__ import annotations

import math
import os
import re
import struct

# ------------------------------------------------------------------------------------------------ values
class _Undefined:
    __slots__ = ()
    def __repr__(self): return "undefined"
    def __bool__(self): return False
UNDEF = _Undefined()
class _Null:
    __slots__ = ()
    def __repr__(self): return "null"
    def __bool__(self): return False
NULL = _Null()
class _Short:                                               # an optional chain (a?.b.c) that met null / undefined: the whole chain is undefined
    __slots__ = ()
    def __bool__(self): return False
SHORT = _Short()


class JSSyntaxError(Exception): pass
class JSRuntimeError(Exception): pass
class JSThrow(Exception):
    def __init__(self, value): super().__init__(to_str(value)); self.value = value


class JSObject:
    __slots__ = ("props", "proto")
    def __init__(self, proto=None, props=None):
        self.props = props if props is not None else {}
        self.proto = proto
    def get(self, key):
        o = self
        while o is not None:
            p = o.props
            if key in p: return p[key]
            o = o.proto
        return UNDEF
    def set(self, key, value): self.props[key] = value
    def has(self, key):
        o = self
        while o is not None:
            if key in o.props: return True
            o = o.proto
        return False


class JSArray(JSObject):
    __slots__ = ("items",)
    def __init__(self, items=None):
        super().__init__(ARRAY_PROTO)
        self.items = items if items is not None else []


class JSTyped(JSObject):                                   # Float32Array / Float64Array / Uint8ClampedArray
    __slots__ = ("items", "kind")
    def __init__(self, kind, n):
        super().__init__(TYPED_PROTO)
        self.kind = kind
        self.items = [0.0] * int(n)
    def store(self, i, v):
        v = to_num(v)
        if self.kind == "f32":
            v = struct.unpack("f", struct.pack("f", v))[0] if math.isfinite(v) and abs(v) < 3.4028235677973366e38 else (v if not math.isfinite(v) else math.copysign(math.inf, v))
        elif self.kind == "u8c":                           # ToUint8Clamp: clamp to [0, 255], round half to even
            v = 0.0 if not (v > 0) else 255.0 if v >= 255 else float(round(v))
        elif self.kind == "u8":                            # ToUint8: truncate, modulo 256
            v = float(int(math.trunc(v)) % 256) if math.isfinite(v) else 0.0
        self.items[i] = v


class JSFunction(JSObject):
    __slots__ = ("params", "body", "env", "is_arrow", "this_val", "name", "home", "is_class", "ctor_kind", "fields", "native", "bound")
    def __init__(self, params=None, body=None, env=None, is_arrow=False, name="", native=None):
        super().__init__(FUNCTION_PROTO)
        self.params, self.body, self.env, self.is_arrow, self.name, self.native = params, body, env, is_arrow, name, native
        self.this_val = UNDEF; self.home = None; self.is_class = False; self.ctor_kind = "base"; self.bound = None


OBJECT_PROTO = JSObject()
FUNCTION_PROTO = JSObject(OBJECT_PROTO)
ARRAY_PROTO = JSObject(OBJECT_PROTO)
TYPED_PROTO = JSObject(OBJECT_PROTO)
STRING_PROTO = JSObject(OBJECT_PROTO)
PROMISE_PROTO = JSObject(OBJECT_PROTO)


def native(fn, name=""):
    return JSFunction(native=fn, name=name or getattr(fn, "__name__", ""))


def truthy(v):
    if v is True: return True
    if v is False or v is UNDEF or v is NULL: return False
    if isinstance(v, float): return not (v == 0.0 or v != v)
    if isinstance(v, str): return len(v) > 0
    return True


def to_num(v):
    if isinstance(v, float): return v
    if v is True: return 1.0
    if v is False or v is NULL: return 0.0
    if v is UNDEF: return math.nan
    if isinstance(v, str):
        s = v.strip()
        if not s: return 0.0
        try: return float(int(s, 16)) if s[:2].lower() == "0x" else float(s)
        except ValueError: return math.nan
    if isinstance(v, int): return float(v)
    return math.nan


def num_to_str(x):
    if x != x: return "NaN"
    if x in (math.inf, -math.inf): return "Infinity" if x > 0 else "-Infinity"
    if x == int(x) and abs(x) < 1e21: return str(int(x))
    return repr(x)


def to_str(v):
    if isinstance(v, str): return v
    if isinstance(v, float): return num_to_str(v)
    if v is True: return "true"
    if v is False: return "false"
    if v is UNDEF: return "undefined"
    if v is NULL: return "null"
    if isinstance(v, (JSArray, JSTyped)): return ",".join("" if x is UNDEF or x is NULL else to_str(x) for x in v.items)
    if isinstance(v, JSFunction): return "function " + v.name
    if isinstance(v, JSObject):
        m = v.get("message")
        return to_str(m) if m is not UNDEF and v.has("stack") else "[object Object]"
    return str(v)


def to_int32(v):
    x = to_num(v)
    if x != x or x in (math.inf, -math.inf): return 0
    n = int(x) & 0xFFFFFFFF
    return n - (1 << 32) if n & 0x80000000 else n


def typeof(v):
    if v is UNDEF: return "undefined"
    if v is NULL: return "object"
    if isinstance(v, bool): return "boolean"
    if isinstance(v, float): return "number"
    if isinstance(v, int): return "bigint"
    if isinstance(v, str): return "string"
    if isinstance(v, JSFunction): return "function"
    return "object"


def strict_eq(a, b):
    if isinstance(a, float) and isinstance(b, float): return a == b
    if isinstance(a, bool) or isinstance(b, bool): return a is b
    if isinstance(a, int) and isinstance(b, int): return a == b           # BigInt
    if isinstance(a, str) and isinstance(b, str): return a == b
    return a is b


def loose_eq(a, b):
    if (a is UNDEF or a is NULL) and (b is UNDEF or b is NULL): return True
    if type(a) is type(b) or (isinstance(a, JSObject) and isinstance(b, JSObject)): return strict_eq(a, b)
    if isinstance(a, JSObject) or isinstance(b, JSObject) or a is UNDEF or a is NULL or b is UNDEF or b is NULL: return False
    return to_num(a) == to_num(b)


def prop_key(k):
    if isinstance(k, str): return k
    if isinstance(k, float): return num_to_str(k)
    return to_str(k)


def get_member(obj, key):
    """obj[key] / obj.key for every value kind"""
    if isinstance(obj, (JSArray, JSTyped)):
        if isinstance(key, float):
            i = int(key)
            if i == key and 0 <= i < len(obj.items): return obj.items[i]
            if i == key: return UNDEF
        if key == "length": return float(len(obj.items))
        return obj.get(prop_key(key))
    if isinstance(obj, JSFunction) and key == "name" and "name" not in obj.props:
        return obj.name                                     # Function.prototype.name ("bound f" for f.bind(...))
    if isinstance(obj, JSObject):
        return obj.get(prop_key(key))
    if isinstance(obj, str):
        if key == "length": return float(len(obj))
        if isinstance(key, float): return obj[int(key)] if 0 <= int(key) < len(obj) else UNDEF
        return STRING_PROTO.get(prop_key(key))
    if obj is UNDEF or obj is NULL:
        raise JSThrow(make_error("TypeError", f"Cannot read properties of {obj!r} (reading '{prop_key(key)}')"))
    if isinstance(obj, float): return UNDEF
    return UNDEF


def set_member(obj, key, value):
    if isinstance(obj, JSArray) and isinstance(key, float) and int(key) == key and key >= 0:
        i = int(key); items = obj.items
        if i >= len(items): items.extend([UNDEF] * (i + 1 - len(items)))
        items[i] = value
        return value
    if isinstance(obj, JSTyped) and isinstance(key, float):
        i = int(key)
        if i == key and 0 <= i < len(obj.items): obj.store(i, value)
        return value
    if isinstance(obj, JSArray) and key == "length":
        n = int(to_num(value)); del obj.items[n:]; obj.items.extend([UNDEF] * (n - len(obj.items)))
        return value
    if isinstance(obj, JSObject):
        obj.props[prop_key(key)] = value
        return value
    raise JSThrow(make_error("TypeError", f"Cannot set properties of {obj!r} (setting '{prop_key(key)}')"))


def make_error(kind, message):
    e = JSObject(ERROR_PROTO)
    e.props["name"] = kind; e.props["message"] = message; e.props["stack"] = kind + ": " + message
    return e
ERROR_PROTO = JSObject(OBJECT_PROTO)


# ------------------------------------------------------------------------------------------------ tokenizer
KEYWORDS = {"var", "let", "const", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "new", "this", "class",
            "extends", "super", "static", "import", "export", "from", "default", "null", "undefined", "true", "false", "typeof", "instanceof",
            "in", "of", "switch", "case", "try", "catch", "finally", "throw", "async", "await", "void", "delete", "as", "get", "set"}
PUNCT = sorted(["===", "!==", ">>>", "...", "**=", "<<=", ">>=", "&&=", "||=", "??=", "==", "!=", "<=", ">=", "&&", "||", "??", "?.", "=>", "++", "--",
                "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<", ">>", "**", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/", "%",
                "&", "|", "^", "!", "~", "?", ":", "=", "."], key=len, reverse=True)
NUM_RE = re.compile(r"0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")
ID_RE = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*")
ESC = {"n": "\n", "t": "\t", "r": "\r", "0": "\0", "b": "\b", "f": "\f", "v": "\v"}


class Tok:
    __slots__ = ("kind", "val", "pos", "nl")
    def __init__(self, kind, val, pos, nl): self.kind, self.val, self.pos, self.nl = kind, val, pos, nl
    def __repr__(self): return f"{self.kind}:{self.val!r}"


def tokenize(src, fname="<js>"):
    toks, i, n, nl = [], 0, len(src), False
    def unescape(s):
        out, k = [], 0
        while k < len(s):
            c = s[k]
            if c == "\\" and k + 1 < len(s):
                d = s[k + 1]
                if d == "u": out.append(chr(int(s[k + 2:k + 6], 16))); k += 6; continue
                if d == "x": out.append(chr(int(s[k + 2:k + 4], 16))); k += 4; continue
                out.append(ESC.get(d, d)); k += 2; continue
            out.append(c); k += 1
        return "".join(out)
    while i < n:
        c = src[i]
        if c == "\n": nl = True; i += 1; continue
        if c in " \t\r﻿": i += 1; continue
        if src.startswith("//", i):
            j = src.find("\n", i); i = n if j < 0 else j; continue
        if src.startswith("/*", i):
            j = src.find("*/", i + 2)
            if j < 0: raise JSSyntaxError(f"{fname}: unterminated comment")
            if "\n" in src[i:j]: nl = True
            i = j + 2; continue
        if c in "'\"":
            j = i + 1
            while src[j] != c:
                j += 2 if src[j] == "\\" else 1
            toks.append(Tok("str", unescape(src[i + 1:j]), i, nl)); nl = False; i = j + 1; continue
        if c == "`":
            parts, j, cur = [], i + 1, []
            while src[j] != "`":
                if src[j] == "\\": cur.append(src[j:j + 2]); j += 2; continue
                if src.startswith("${", j):
                    parts.append(("s", unescape("".join(cur)))); cur = []
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == "{": depth += 1
                        elif src[k] == "}": depth -= 1
                        k += 1
                    parts.append(("e", src[j + 2:k - 1])); j = k; continue
                cur.append(src[j]); j += 1
            parts.append(("s", unescape("".join(cur))))
            toks.append(Tok("tpl", parts, i, nl)); nl = False; i = j + 1; continue
        m = NUM_RE.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            t = m.group(0)
            if m.end() < n and src[m.end()] == "n" and re.fullmatch(r"0[xX][0-9a-fA-F]+|\d+", t):      # BigInt literal: a Python int
                toks.append(Tok("num", int(t, 16) if t[:2].lower() == "0x" else int(t), i, nl)); nl = False; i = m.end() + 1; continue
            toks.append(Tok("num", float(int(t, 16)) if t[:2].lower() == "0x" else float(t), i, nl)); nl = False; i = m.end(); continue
        m = ID_RE.match(src, i)
        if m:
            w = m.group(0)
            toks.append(Tok("kw" if w in KEYWORDS else "id", w, i, nl)); nl = False; i = m.end(); continue
        for p in PUNCT:
            if src.startswith(p, i):
                toks.append(Tok("p", p, i, nl)); nl = False; i += len(p); break
        else:
            raise JSSyntaxError(f"{fname}: unexpected character {c!r} at {i}")
    toks.append(Tok("eof", None, n, nl))
    return toks


# ------------------------------------------------------------------------------------------------ parser (AST = tuples)
CONTEXTUAL = {"of", "from", "as", "get", "set", "static", "async", "default"}     # keywords usable as identifiers / property names
BINPREC = {"??": 1, "||": 2, "&&": 3, "|": 4, "^": 5, "&": 6, "==": 7, "!=": 7, "===": 7, "!==": 7, "<": 8, ">": 8, "<=": 8, ">=": 8, "instanceof": 8, "in": 8,
           "<<": 9, ">>": 9, ">>>": 9, "+": 10, "-": 10, "*": 11, "/": 11, "%": 11, "**": 12}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", "**=", "||=", "&&=", "??="}


class Parser:
    def __init__(self, src, fname="<js>"):
        self.src, self.fname = src, fname
        self.t = tokenize(src, fname)
        self.i = 0
    # -- helpers
    def peek(self, k=0): return self.t[self.i + k]
    def next(self): tok = self.t[self.i]; self.i += 1; return tok
    def is_p(self, v, k=0): tok = self.t[self.i + k]; return tok.kind == "p" and tok.val == v
    def is_kw(self, v, k=0): tok = self.t[self.i + k]; return tok.kind == "kw" and tok.val == v
    def err(self, msg):
        tok = self.peek(); line = self.src.count("\n", 0, tok.pos) + 1
        raise JSSyntaxError(f"{self.fname}:{line}: {msg} (at {tok!r})")
    def expect_p(self, v):
        if not self.is_p(v): self.err(f"expected '{v}'")
        return self.next()
    def eat_p(self, v):
        if self.is_p(v): self.i += 1; return True
        return False
    def eat_kw(self, v):
        if self.is_kw(v): self.i += 1; return True
        return False
    def semi(self):
        if self.eat_p(";"): return
        tok = self.peek()
        if tok.kind == "eof" or tok.nl or (tok.kind == "p" and tok.val == "}"): return
        self.err("expected ';'")
    def ident(self):
        tok = self.next()
        if tok.kind == "id" or (tok.kind == "kw" and tok.val in CONTEXTUAL): return tok.val
        self.i -= 1; self.err("expected identifier")
    def prop_name(self):
        tok = self.next()
        if tok.kind in ("id", "kw"): return ("lit", tok.val)
        if tok.kind == "str": return ("lit", tok.val)
        if tok.kind == "num": return ("lit", num_to_str(tok.val))
        if tok.kind == "p" and tok.val == "[":
            e = self.assign(); self.expect_p("]"); return ("computed", e)
        self.i -= 1; self.err("expected property name")
    # -- program / statements
    def program(self):
        body = []
        while self.peek().kind != "eof": body.append(self.statement())
        return body
    def block(self):
        self.expect_p("{"); body = []
        while not self.is_p("}"): body.append(self.statement())
        self.next(); return ("block", body)
    def statement(self):
        tok = self.peek()
        if tok.kind == "p":
            if tok.val == "{": return self.block()
            if tok.val == ";": self.next(); return ("empty",)
        if tok.kind == "kw":
            v = tok.val
            if v in ("var", "let", "const"): d = self.var_decl(); self.semi(); return d
            if v == "function" or (v == "async" and self.is_kw("function", 1)): return self.function_decl()
            if v == "class": return self.class_decl()
            if v == "return":
                self.next(); tok2 = self.peek()
                arg = None if (tok2.nl or tok2.kind == "eof" or (tok2.kind == "p" and tok2.val in (";", "}"))) else self.expression()
                self.semi(); return ("return", arg)
            if v == "if":
                self.next(); self.expect_p("("); c = self.expression(); self.expect_p(")")
                a = self.statement(); b = self.statement() if self.eat_kw("else") else None
                return ("if", c, a, b)
            if v == "for": return self.for_stmt()
            if v == "while":
                self.next(); self.expect_p("("); c = self.expression(); self.expect_p(")"); return ("while", c, self.statement())
            if v == "do":
                self.next(); body = self.statement()
                if not self.eat_kw("while"): self.err("expected while")
                self.expect_p("("); c = self.expression(); self.expect_p(")"); self.eat_p(";"); return ("dowhile", body, c)
            if v == "break": self.next(); self.semi(); return ("break",)
            if v == "continue": self.next(); self.semi(); return ("continue",)
            if v == "throw": self.next(); e = self.expression(); self.semi(); return ("throw", e)
            if v == "try":
                self.next(); blk = self.block(); param = None; handler = None; fin = None
                if self.eat_kw("catch"):
                    if self.eat_p("("): param = self.ident(); self.expect_p(")")
                    handler = self.block()
                if self.eat_kw("finally"): fin = self.block()
                return ("try", blk, param, handler, fin)
            if v == "switch": return self.switch_stmt()
            if v == "import" and not (self.is_p(".", 1) or self.is_p("(", 1)): return self.import_decl()
            if v == "export": return self.export_decl()
        e = self.expression(); self.semi(); return ("expr", e)
    def var_decl(self):
        kind = self.next().val; decls = []
        while True:
            target = self.binding_target()
            init = self.assign() if self.eat_p("=") else None
            decls.append((target, init))
            if not self.eat_p(","): break
        return ("var", kind, decls)
    def binding_target(self):
        if self.is_p("{"):                                  # const { a, b: c } = obj
            self.next(); names = []
            while not self.is_p("}"):
                k = self.ident(); alias = self.ident() if self.eat_p(":") else k
                dflt = self.assign() if self.eat_p("=") else None         # { a = 1, b: c = 2 }: used when the property is undefined
                names.append((k, alias, dflt)); self.eat_p(",")
            self.next(); return ("objpat", names)
        if self.is_p("["):
            self.next(); names = []
            while not self.is_p("]"):
                names.append(self.ident()); self.eat_p(",")
            self.next(); return ("arrpat", names)
        return ("name", self.ident())
    def for_stmt(self):
        self.next(); self.expect_p("(")
        init = None
        if self.is_kw("var") or self.is_kw("let") or self.is_kw("const"):
            save = self.i; kind = self.next().val; target = self.binding_target()
            if self.eat_kw("of"):
                it = self.assign(); self.expect_p(")"); return ("forof", kind, target, it, self.statement())
            if self.eat_kw("in"):
                it = self.expression(); self.expect_p(")"); return ("forin", kind, target, it, self.statement())
            self.i = save; init = self.var_decl()
        elif not self.is_p(";"):
            init = ("expr", self.expression())
        self.expect_p(";")
        test = None if self.is_p(";") else self.expression(); self.expect_p(";")
        upd = None if self.is_p(")") else self.expression(); self.expect_p(")")
        return ("for", init, test, upd, self.statement())
    def switch_stmt(self):
        self.next(); self.expect_p("("); d = self.expression(); self.expect_p(")"); self.expect_p("{")
        cases = []
        while not self.is_p("}"):
            if self.eat_kw("default"): test = None
            elif self.eat_kw("case"): test = self.expression()
            else: self.err("expected case")
            self.expect_p(":"); body = []
            while not (self.is_kw("case") or self.is_kw("default") or self.is_p("}")): body.append(self.statement())
            cases.append((test, body))
        self.next(); return ("switch", d, cases)
    def import_decl(self):
        self.next(); names = []; default = None
        if self.peek().kind == "str":
            src = self.next().val; self.semi(); return ("import", [], None, src)
        if not self.is_p("{"):
            default = self.ident(); self.eat_p(",")
        if self.eat_p("{"):
            while not self.is_p("}"):
                k = self.ident(); alias = self.ident() if self.eat_kw("as") else k
                names.append((k, alias)); self.eat_p(",")
            self.next()
        if not self.eat_kw("from"): self.err("expected from")
        src = self.next().val; self.semi()
        return ("import", names, default, src)
    def export_decl(self):
        self.next()
        if self.eat_kw("default"):
            if self.is_kw("class"): d = self.class_decl(); return ("export_default_decl", d)
            if self.is_kw("function"): d = self.function_decl(); return ("export_default_decl", d)
            e = self.assign(); self.semi(); return ("export_default", e)
        if self.eat_p("{"):
            names = []
            while not self.is_p("}"):
                k = self.ident(); alias = self.ident() if self.eat_kw("as") else k
                names.append((k, alias)); self.eat_p(",")
            self.next(); self.semi(); return ("export_names", names)
        d = self.statement(); return ("export_decl", d)
    def function_decl(self):
        is_async = self.eat_kw("async"); self.next()
        name = self.ident(); params = self.params(); body = self.block()
        return ("funcdecl", name, ("func", name, params, body, False, is_async))
    def params(self):
        self.expect_p("("); ps = []
        while not self.is_p(")"):
            rest = self.eat_p("...")
            target = self.binding_target()
            default = self.assign() if self.eat_p("=") else None
            ps.append((target, default, rest)); self.eat_p(",")
        self.next(); return ps
    def class_decl(self):
        c = self.class_expr()
        if c[1] is None: self.err("class declaration needs a name")
        return ("classdecl", c[1], c)
    def class_expr(self):
        self.next(); name = None
        if self.peek().kind == "id": name = self.next().val
        sup = self.unary() if self.eat_kw("extends") else None
        self.expect_p("{"); members = []
        while not self.is_p("}"):
            if self.eat_p(";"): continue
            static = False
            if self.is_kw("static") and not self.is_p("(", 1) and not self.is_p("=", 1): self.next(); static = True
            is_async = False
            if self.is_kw("async") and not self.is_p("(", 1) and not self.is_p("=", 1): self.next(); is_async = True
            accessor = None
            if (self.is_kw("get") or self.is_kw("set")) and not self.is_p("(", 1) and not self.is_p("=", 1): accessor = self.next().val
            key = self.prop_name()
            if self.is_p("("):
                params = self.params(); body = self.block()
                nm = key[1] if key[0] == "lit" else ""
                members.append(("method", static, key, ("func", nm, params, body, False, is_async), accessor))
            else:                                           # class field
                init = self.assign() if self.eat_p("=") else None
                self.semi(); members.append(("field", static, key, init, None))
        self.next()
        return ("class", name, sup, members)
    # -- expressions
    def expression(self):
        e = self.assign()
        while self.eat_p(","):
            e = ("seq", e, self.assign())
        return e
    def is_arrow_ahead(self):
        """at '(' : is this the parameter list of an arrow function?"""
        depth, k = 0, self.i
        while True:
            tok = self.t[k]
            if tok.kind == "eof": return False
            if tok.kind == "p":
                if tok.val in "([{": depth += 1
                elif tok.val in ")]}":
                    depth -= 1
                    if depth == 0: nxt = self.t[k + 1]; return nxt.kind == "p" and nxt.val == "=>"
            k += 1
    def arrow_body(self, params, is_async):
        if self.is_p("{"): body = self.block()
        else: body = ("block", [("return", self.assign())])
        return ("func", "", params, body, True, is_async)
    def assign(self):
        tok = self.peek()
        if tok.kind == "kw" and tok.val == "async" and not self.peek(1).nl:
            if self.is_p("(", 1):
                self.i += 1
                if self.is_arrow_ahead(): params = self.params(); self.expect_p("=>"); return self.arrow_body(params, True)
                self.i -= 1
            elif self.peek(1).kind == "id" and self.is_p("=>", 2):
                self.i += 1; name = self.next().val; self.next(); return self.arrow_body([(("name", name), None, False)], True)
        if tok.kind == "id" and self.is_p("=>", 1):
            self.i += 2; return self.arrow_body([(("name", tok.val), None, False)], False)
        if tok.kind == "p" and tok.val == "(" and self.is_arrow_ahead():
            params = self.params(); self.expect_p("=>"); return self.arrow_body(params, False)
        left = self.conditional()
        tok = self.peek()
        if tok.kind == "p" and tok.val in ASSIGN_OPS:
            if left[0] not in ("id", "member", "index") and not (left[0] == "array" and tok.val == "="): self.err("invalid assignment target")
            self.next(); right = self.assign()
            return ("assign", tok.val, left, right)
        return left
    def conditional(self):
        c = self.binary(0)
        if self.eat_p("?"):
            a = self.assign(); self.expect_p(":"); b = self.assign()
            return ("cond", c, a, b)
        return c
    def binary(self, minprec):
        left = self.unary()
        while True:
            tok = self.peek()
            op = tok.val if tok.kind == "p" or (tok.kind == "kw" and tok.val in ("instanceof", "in")) else None
            prec = BINPREC.get(op)
            if prec is None or prec <= minprec: return left
            self.next()
            right = self.binary(prec - 1 if op == "**" else prec)
            left = ("logical", op, left, right) if op in ("&&", "||", "??") else ("bin", op, left, right)
    def unary(self):
        tok = self.peek()
        if tok.kind == "p":
            if tok.val in ("!", "-", "+", "~"): self.next(); return ("unary", tok.val, self.unary())
            if tok.val in ("++", "--"): self.next(); return ("update", tok.val, True, self.unary())
        elif tok.kind == "kw":
            if tok.val in ("typeof", "void", "delete"): self.next(); return ("unary", tok.val, self.unary())
            if tok.val == "await": self.next(); return ("await", self.unary())
        e = self.postfix()
        if self.is_p("**"):
            self.next(); return ("bin", "**", e, self.unary())
        return e
    def postfix(self):
        e = self.call_member()
        tok = self.peek()
        if tok.kind == "p" and tok.val in ("++", "--") and not tok.nl:
            self.next(); return ("update", tok.val, False, e)
        return e
    def args(self):
        self.expect_p("("); a = []
        while not self.is_p(")"):
            if self.eat_p("..."): a.append(("spread", self.assign()))
            else: a.append(self.assign())
            self.eat_p(",")
        self.next(); return a
    def call_member(self):
        if self.is_kw("new"):
            self.next()
            if self.is_kw("new"): callee = self.call_member()
            else:
                callee = self.primary()
                while True:
                    if self.eat_p("."): callee = ("member", callee, self.next().val, False)
                    elif self.is_p("["): self.next(); k = self.expression(); self.expect_p("]"); callee = ("index", callee, k, False)
                    else: break
            e = ("new", callee, self.args() if self.is_p("(") else [])
        else:
            e = self.primary()
        has_opt = False
        while True:
            tok = self.peek()
            if tok.kind != "p": return ("optchain", e) if has_opt else e
            if tok.val == ".":
                self.next(); name = self.next()
                if name.kind not in ("id", "kw"): self.err("expected property name")
                e = ("member", e, name.val, False)
            elif tok.val == "?.":
                self.next(); has_opt = True
                if self.is_p("("): e = ("call", e, self.args(), True)
                elif self.is_p("["): self.next(); k = self.expression(); self.expect_p("]"); e = ("index", e, k, True)
                else: e = ("member", e, self.next().val, True)
            elif tok.val == "[":
                self.next(); k = self.expression(); self.expect_p("]"); e = ("index", e, k, False)
            elif tok.val == "(":
                e = ("call", e, self.args(), False)
            else:
                return ("optchain", e) if has_opt else e
    def primary(self):
        tok = self.next()
        k, v = tok.kind, tok.val
        if k == "num": return ("num", v)
        if k == "str": return ("str", v)
        if k == "tpl":
            parts = [("str", p[1]) if p[0] == "s" else Parser(p[1], self.fname).expression() for p in v]
            return ("template", parts)
        if k == "id": return ("id", v)
        if k == "kw":
            if v == "this": return ("this",)
            if v == "super": return ("super",)
            if v == "null": return ("const", NULL)
            if v == "undefined": return ("id", "undefined")
            if v == "true": return ("const", True)
            if v == "false": return ("const", False)
            if v == "function" or (v == "async" and self.is_kw("function")):
                is_async = v == "async"
                if is_async: self.next()
                name = self.next().val if self.peek().kind == "id" else ""
                params = self.params(); body = self.block()
                return ("func", name, params, body, False, is_async)
            if v == "class": self.i -= 1; return self.class_expr()
            if v == "import" and self.is_p(".") and self.peek(1).val == "meta":
                self.next(); self.next(); return ("import_meta",)
            if v == "import" and self.is_p("("):                                   # import(specifier): a promise of the module namespace
                self.next(); e = self.assign(); self.expect_p(")"); return ("dyn_import", e)
            if v in CONTEXTUAL: return ("id", v)
        if k == "p":
            if v == "(":
                e = self.expression(); self.expect_p(")"); return e
            if v == "[":
                items = []
                while not self.is_p("]"):
                    if self.eat_p("..."): items.append(("spread", self.assign()))
                    else: items.append(self.assign())
                    self.eat_p(",")
                self.next(); return ("array", items)
            if v == "{":
                props = []
                while not self.is_p("}"):
                    if self.eat_p("..."): props.append(("spread", self.assign())); self.eat_p(","); continue
                    is_async = False
                    if self.is_kw("async") and not (self.is_p(":", 1) or self.is_p("(", 1) or self.is_p(",", 1) or self.is_p("}", 1)): self.next(); is_async = True
                    key = self.prop_name()
                    if self.is_p("("):
                        params = self.params(); body = self.block()
                        props.append(("kv", key, ("func", key[1] if key[0] == "lit" else "", params, body, False, is_async)))
                    elif self.eat_p(":"): props.append(("kv", key, self.assign()))
                    else: props.append(("kv", key, ("id", key[1])))                 # shorthand {a, b}
                    self.eat_p(",")
                self.next(); return ("object", props)
        self.i -= 1; self.err("unexpected token")


# ------------------------------------------------------------------------------------------------ environments / completions
class Env:
    __slots__ = ("vars", "parent")
    def __init__(self, parent=None): self.vars = {}; self.parent = parent
    def lookup(self, name):
        e = self
        while e is not None:
            v = e.vars
            if name in v: return v
            e = e.parent
        return None
    def get(self, name):
        e = self
        while e is not None:
            v = e.vars
            if name in v: return v[name]
            e = e.parent
        raise JSThrow(make_error("ReferenceError", name + " is not defined"))
    def set(self, name, value):
        v = self.lookup(name)
        if v is None: raise JSThrow(make_error("ReferenceError", name + " is not defined"))
        v[name] = value

BREAK, CONTINUE = object(), object()
class Return:
    __slots__ = ("value",)
    def __init__(self, value): self.value = value


# ------------------------------------------------------------------------------------------------ compiler: AST -> closures
class Interp:
    def __init__(self, module_root="."):
        self.root = module_root
        self.modules = {}
        self._pat_defaults = {}
        self.builtin_modules = {}                            # specifier (e.g. 'node:module') -> exports dict, provided by the host
        self.globals = Env()
        install_globals(self)
    # -- modules
    def load_module(self, path):
        path = os.path.normpath(path)
        if path in self.modules: return self.modules[path]
        exports = {}
        self.modules[path] = exports
        with open(path) as f: src = f.read()
        ast = Parser(src, os.path.basename(path)).program()
        env = Env(self.globals)
        env.vars["%exports"] = exports; env.vars["%dir"] = os.path.dirname(path); env.vars["%file"] = path
        self.hoist(ast, env)
        for st in ast:
            r = self.stmt(st)(env)
            if r is not None: break
        return exports
    def run(self, src, env=None, fname="<js>"):
        env = env or Env(self.globals)
        ast = Parser(src, fname).program()
        self.hoist(ast, env)
        last = UNDEF
        for st in ast:
            last = self.stmt(st)(env)
        return env
    def hoist(self, body, env):
        for st in body:
            if st[0] == "export_decl" or st[0] == "export_default_decl": st = st[1]
            if st[0] == "funcdecl": env.vars[st[1]] = self.make_function(st[2], env)
            elif st[0] == "var" and st[1] == "var":
                for target, _ in st[2]:
                    if target[0] == "name": env.vars.setdefault(target[1], UNDEF)
    # -- functions
    def make_function(self, node, env, this_val=UNDEF, home=None):
        _, name, params, body, is_arrow, is_async = node
        f = JSFunction(params=[(t, self.expr(d) if d is not None else None, rest) for t, d, rest in params], body=self.compile_body(body[1]),
                       env=env, is_arrow=is_arrow, name=name)
        f.home = home
        if not is_arrow:
            f.props["prototype"] = JSObject(OBJECT_PROTO, {"constructor": f})
        return f
    def compile_body(self, stmts):
        hoisted = [st for st in stmts if st[0] == "funcdecl"]
        compiled = [self.stmt(st) for st in stmts]
        interp = self
        def run(env):
            for st in hoisted: env.vars[st[1]] = interp.make_function(st[2], env)
            for c in compiled:
                r = c(env)
                if r is not None: return r
            return None
        return run
    def bind_target(self, target, value, env):
        if target[0] == "name": env.vars[target[1]] = value
        elif target[0] == "objpat":
            for k, alias, dflt in target[1]:
                v = get_member(value, k)
                if v is UNDEF and dflt is not None:
                    c = self._pat_defaults.get(id(dflt))
                    if c is None: c = self._pat_defaults[id(dflt)] = self.expr(dflt)
                    v = c(env)
                env.vars[alias] = v
        else:
            for i, nm in enumerate(target[1]): env.vars[nm] = get_member(value, float(i))
    def call(self, f, this_val, args, new_target=None):
        if not isinstance(f, JSFunction):
            raise JSThrow(make_error("TypeError", f"{to_str(f)} is not a function"))
        if f.bound is not None:
            target, bthis, bargs = f.bound
            return self.call(target, bthis, list(bargs) + list(args))
        if f.native is not None: return f.native(this_val, args)
        env = Env(f.env)
        v = env.vars
        if not f.is_arrow:
            v["this"] = this_val; v["%home"] = f.home; v["%func"] = f; v["%newtarget"] = new_target
        n = len(args)
        for i, (target, default, rest) in enumerate(f.params):
            if rest: val = JSArray(list(args[i:]))
            else:
                val = args[i] if i < n else UNDEF
                if val is UNDEF and default is not None: val = default(env)
            if target[0] == "name": v[target[1]] = val
            else: self.bind_target(target, val, env)
        r = f.body(env)
        if isinstance(r, Return): return r.value
        return UNDEF
    def construct(self, f, args):
        if not isinstance(f, JSFunction): raise JSThrow(make_error("TypeError", "not a constructor"))
        if f.native is not None:
            return f.native(None, args)                     # native constructors build their own object
        proto = f.props.get("prototype", OBJECT_PROTO)
        if f.is_class and f.ctor_kind == "derived":
            env_this = UNDEF                                # set by super(...)
            obj_holder = {"obj": UNDEF, "proto": proto}
            return self.run_ctor(f, env_this, args, obj_holder)
        obj = JSObject(proto)
        self.init_fields(f, obj)
        r = self.call(f, obj, args, new_target=f) if f.body is not None else UNDEF
        return r if isinstance(r, JSObject) else obj
    def init_fields(self, cls, obj):
        for key, init, env in getattr_fields(cls):
            obj.props[key] = init(Env_with_this(env, obj)) if init is not None else UNDEF
    def run_ctor(self, f, this_val, args, holder):
        # derived-class constructor: `this` comes into being when super(...) returns
        env = Env(f.env); v = env.vars
        v["this"] = UNDEF; v["%home"] = f.home; v["%func"] = f; v["%holder"] = holder; v["%newtarget"] = f
        n = len(args)
        for i, (target, default, rest) in enumerate(f.params):
            if rest: val = JSArray(list(args[i:]))
            else:
                val = args[i] if i < n else UNDEF
                if val is UNDEF and default is not None: val = default(env)
            self.bind_target(target, val, env)
        r = f.body(env)
        if isinstance(r, Return) and isinstance(r.value, JSObject): return r.value
        return v["this"]
    # -- statements
    def stmt(self, node):
        k = node[0]
        if k == "expr":
            e = self.expr(node[1])
            def run(env): e(env); return None
            return run
        if k == "var":
            kind = node[1]; decls = [(t, self.expr(i) if i is not None else None) for t, i in node[2]]
            interp = self
            if len(decls) == 1 and decls[0][0][0] == "name":
                name, init = decls[0][0][1], decls[0][1]
                if init is None:
                    def run1(env): env.vars.setdefault(name, UNDEF) if kind == "var" else env.vars.__setitem__(name, UNDEF); return None
                    return run1
                def run2(env):
                    val = init(env)
                    if isinstance(val, JSFunction) and not val.name: val.name = name
                    env.vars[name] = val; return None
                return run2
            def run(env):
                for t, init in decls:
                    interp.bind_target(t, init(env) if init is not None else UNDEF, env)
                return None
            return run
        if k == "block":
            body = [self.stmt(s) for s in node[1]]
            hoisted = [st for st in node[1] if st[0] == "funcdecl"]
            interp = self
            def run(env):
                e2 = Env(env)
                for st in hoisted: e2.vars[st[1]] = interp.make_function(st[2], e2)
                for c in body:
                    r = c(e2)
                    if r is not None: return r
                return None
            return run
        if k == "return":
            e = self.expr(node[1]) if node[1] is not None else None
            if e is None: return lambda env: Return(UNDEF)
            return lambda env: Return(e(env))
        if k == "if":
            c = self.expr(node[1]); a = self.stmt(node[2]); b = self.stmt(node[3]) if node[3] is not None else None
            if b is None:
                return lambda env: a(env) if truthy(c(env)) else None
            return lambda env: a(env) if truthy(c(env)) else b(env)
        if k == "for":
            init = self.stmt(node[1]) if node[1] is not None else None
            test = self.expr(node[2]) if node[2] is not None else None
            upd = self.expr(node[3]) if node[3] is not None else None
            body = self.stmt(node[4])
            per_iter = node[1] is not None and node[1][0] == "var" and node[1][1] in ("let", "const")
            def run(env):
                e2 = Env(env)
                if init is not None: init(e2)
                while test is None or truthy(test(e2)):
                    r = body(e2)
                    if r is not None:
                        if r is BREAK: break
                        if r is not CONTINUE: return r
                    if per_iter:                            # fresh binding per iteration (closures capture the iteration's value)
                        e3 = Env(env); e3.vars.update(e2.vars); e2 = e3
                    if upd is not None: upd(e2)
                return None
            return run
        if k == "forof" or k == "forin":
            target = node[2]; it = self.expr(node[3]); body = self.stmt(node[4]); interp = self; keys = k == "forin"
            def run(env):
                seq = it(env)
                if keys:
                    items = [float(i) if False else num_to_str(float(i)) for i in range(len(seq.items))] if isinstance(seq, (JSArray, JSTyped)) else list(seq.props.keys())
                elif isinstance(seq, (JSArray, JSTyped)): items = seq.items
                elif isinstance(seq, str): items = list(seq)
                else: raise JSThrow(make_error("TypeError", "object is not iterable"))
                i = 0
                while i < len(items):
                    e2 = Env(env); interp.bind_target(target, items[i], e2)
                    r = body(e2)
                    if r is not None:
                        if r is BREAK: break
                        if r is not CONTINUE: return r
                    i += 1
                return None
            return run
        if k == "while":
            c = self.expr(node[1]); body = self.stmt(node[2])
            def run(env):
                while truthy(c(env)):
                    r = body(env)
                    if r is not None:
                        if r is BREAK: break
                        if r is not CONTINUE: return r
                return None
            return run
        if k == "dowhile":
            body = self.stmt(node[1]); c = self.expr(node[2])
            def run(env):
                while True:
                    r = body(env)
                    if r is not None:
                        if r is BREAK: break
                        if r is not CONTINUE: return r
                    if not truthy(c(env)): break
                return None
            return run
        if k == "break": return lambda env: BREAK
        if k == "continue": return lambda env: CONTINUE
        if k == "empty": return lambda env: None
        if k == "throw":
            e = self.expr(node[1])
            def run(env): raise JSThrow(e(env))
            return run
        if k == "try":
            blk = self.stmt(node[1]); param = node[2]; handler = self.stmt(node[3]) if node[3] is not None else None
            fin = self.stmt(node[4]) if node[4] is not None else None
            def run(env):
                try:
                    try:
                        return blk(env)
                    except JSThrow as ex:
                        if handler is None: raise
                        e2 = Env(env)
                        if param: e2.vars[param] = ex.value
                        return handler(e2)
                    except (JSRuntimeError, RecursionError, ZeroDivisionError, OverflowError, IndexError, KeyError, AttributeError, TypeError, ValueError) as ex:
                        if handler is None: raise
                        e2 = Env(env)
                        if param: e2.vars[param] = make_error("InternalError", f"{type(ex).__name__}: {ex}")
                        return handler(e2)
                finally:
                    if fin is not None:
                        r = fin(env)
                        if r is not None: return r
            return run
        if k == "switch":
            d = self.expr(node[1]); cases = [(self.expr(t) if t is not None else None, [self.stmt(s) for s in body]) for t, body in node[2]]
            def run(env):
                val = d(env); e2 = Env(env); start = None
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_eq(t(e2), val): start = i; break
                if start is None:
                    for i, (t, _) in enumerate(cases):
                        if t is None: start = i; break
                if start is None: return None
                for _, body in cases[start:]:
                    for c in body:
                        r = c(e2)
                        if r is not None:
                            if r is BREAK: return None
                            return r
                return None
            return run
        if k == "funcdecl":
            return lambda env: None                         # hoisted by the enclosing body
        if k == "classdecl":
            name = node[1]; c = self.expr(node[2])
            def run(env): env.vars[name] = c(env); return None
            return run
        if k == "import":
            names, default, src = node[1], node[2], node[3]; interp = self
            def run(env):
                exports = interp.builtin_modules[src] if src in interp.builtin_modules else interp.load_module(os.path.join(env.get("%dir"), src))
                for kname, alias in names:
                    if kname not in exports: raise JSRuntimeError(f"module {src} does not export {kname}")
                    env.vars[alias] = exports[kname]
                if default: env.vars[default] = exports.get("default", UNDEF)
                return None
            return run
        if k == "export_names":
            names = node[1]
            def run(env):
                ex = env.get("%exports")
                for kname, alias in names: ex[alias] = env.get(kname)
                return None
            return run
        if k == "export_decl" or k == "export_default_decl":
            inner = node[1]; c = self.stmt(inner); is_default = k == "export_default_decl"
            def run(env):
                if inner[0] == "funcdecl": pass
                c(env)
                ex = env.get("%exports")
                if inner[0] in ("funcdecl", "classdecl"): ex["default" if is_default else inner[1]] = env.get(inner[1])
                elif inner[0] == "var":
                    for t, _ in inner[2]: ex[t[1]] = env.get(t[1])
                return None
            return run
        if k == "export_default":
            e = self.expr(node[1])
            def run(env): env.get("%exports")["default"] = e(env); return None
            return run
        raise JSSyntaxError(f"unsupported statement {k}")
    # -- expressions
    def expr(self, node):
        k = node[0]
        interp = self
        if k == "num" or k == "str":
            v = node[1]; return lambda env: v
        if k == "const":
            v = node[1]; return lambda env: v
        if k == "id":
            name = node[1]
            if name == "undefined": return lambda env: UNDEF
            def get(env):
                e = env
                while e is not None:
                    v = e.vars
                    if name in v: return v[name]
                    e = e.parent
                raise JSThrow(make_error("ReferenceError", name + " is not defined"))
            return get
        if k == "dyn_import":
            spec = self.expr(node[1])
            def dyn(env):
                sp = to_str(spec(env))
                if sp in interp.builtin_modules: ex = interp.builtin_modules[sp]
                else:
                    path = sp[7:] if sp.startswith("file://") else (sp if os.path.isabs(sp) else os.path.join(env.get("%dir"), sp))
                    ex = interp.load_module(path)
                return JSObject(PROMISE_PROTO, {"%state": "fulfilled", "%value": JSObject(OBJECT_PROTO, dict(ex))})
            return dyn
        if k == "import_meta":
            return lambda env: JSObject(OBJECT_PROTO, {"url": "file://" + str(env.get("%file"))})
        if k == "this":
            def get_this(env):
                e = env
                while e is not None:
                    v = e.vars
                    if "this" in v: return v["this"]
                    e = e.parent
                return UNDEF
            return get_this
        if k == "template":
            parts = [self.expr(p) for p in node[1]]
            return lambda env: "".join(to_str(p(env)) for p in parts)
        if k == "array":
            items = [(True, self.expr(i[1])) if i[0] == "spread" else (False, self.expr(i)) for i in node[1]]
            def mk(env):
                out = []
                for sp, e in items:
                    if sp: out.extend(e(env).items)
                    else: out.append(e(env))
                return JSArray(out)
            return mk
        if k == "object":
            props = []
            for p in node[1]:
                if p[0] == "spread": props.append(("spread", self.expr(p[1])))
                else:
                    key = p[1]; val = p[2]
                    kf = (lambda kk: (lambda env: kk))(key[1]) if key[0] == "lit" else (lambda ke: (lambda env: prop_key(ke(env))))(self.expr(key[1]))
                    if val[0] == "func":
                        vf = (lambda vn: (lambda env: interp.make_function(vn, env)))(val)
                    else: vf = self.expr(val)
                    props.append(("kv", kf, vf))
            def mk(env):
                o = JSObject(OBJECT_PROTO)
                for p in props:
                    if p[0] == "spread":
                        s = p[1](env)
                        if isinstance(s, JSObject): o.props.update(s.props)
                    else: o.props[p[1](env)] = p[2](env)
                return o
            return mk
        if k == "func":
            return lambda env: interp.make_function(node, env)
        if k == "class": return self.compile_class(node)
        if k == "seq":
            a, b = self.expr(node[1]), self.expr(node[2])
            def seq(env): a(env); return b(env)
            return seq
        if k == "cond":
            c, a, b = self.expr(node[1]), self.expr(node[2]), self.expr(node[3])
            return lambda env: a(env) if truthy(c(env)) else b(env)
        if k == "logical":
            op, a, b = node[1], self.expr(node[2]), self.expr(node[3])
            if op == "&&":
                def land(env):
                    v = a(env); return b(env) if truthy(v) else v
                return land
            if op == "||":
                def lor(env):
                    v = a(env); return v if truthy(v) else b(env)
                return lor
            def nullish(env):
                v = a(env); return b(env) if (v is UNDEF or v is NULL) else v
            return nullish
        if k == "unary":
            op, a = node[1], self.expr(node[2])
            if op == "!": return lambda env: not truthy(a(env))
            if op == "-":
                def neg(env):
                    v = a(env); return -v if isinstance(v, float) else -to_num(v)
                return neg
            if op == "+": return lambda env: to_num(a(env))
            if op == "~": return lambda env: float(~to_int32(a(env)))
            if op == "typeof":
                if node[2][0] == "id":
                    nm = node[2][1]
                    def ty(env):
                        v = env.lookup(nm); return "undefined" if v is None else typeof(v[nm])
                    return ty
                return lambda env: typeof(a(env))
            if op == "void":
                def vd(env): a(env); return UNDEF
                return vd
            if op == "delete":
                tgt = node[2]
                if tgt[0] == "member":
                    o = self.expr(tgt[1]); nm = tgt[2]
                    def dl(env):
                        ob = o(env)
                        if isinstance(ob, JSObject): ob.props.pop(nm, None)
                        return True
                    return dl
                raise JSSyntaxError("unsupported delete target")
        if k == "await":
            a = self.expr(node[1])
            def aw(env):
                v = a(env)
                if isinstance(v, JSObject) and v.proto is PROMISE_PROTO:
                    if v.props.get("%state") == "rejected": raise JSThrow(v.props.get("%value", UNDEF))
                    return v.props.get("%value", UNDEF)
                return v
            return aw
        if k == "bin": return self.compile_bin(node)
        if k == "assign": return self.compile_assign(node)
        if k == "update":
            op, prefix, tgt = node[1], node[2], node[3]
            delta = 1.0 if op == "++" else -1.0
            getter = self.expr(tgt); setter = self.compile_store(tgt)
            def upd(env):
                old = to_num(getter(env)); new = old + delta
                setter(env, new)
                return new if prefix else old
            return upd
        if k == "member":
            o = self.expr(node[1]); name = node[2]; optional = node[3]
            def mem(env):
                ob = o(env)
                if ob.__class__ is JSObject:                # fast path
                    while ob is not None:
                        p = ob.props
                        if name in p: return p[name]
                        ob = ob.proto
                    return UNDEF
                if ob is SHORT or (optional and (ob is UNDEF or ob is NULL)): return SHORT
                return get_member(ob, name)
            return mem
        if k == "index":
            o = self.expr(node[1]); key = self.expr(node[2]); optional = node[3]
            def idx(env):
                ob = o(env)
                if ob is SHORT or (optional and (ob is UNDEF or ob is NULL)): return SHORT
                return get_member(ob, key(env))
            return idx
        if k == "call": return self.compile_call(node)
        if k == "optchain":
            inner = self.expr(node[1])
            def oc(env):
                v = inner(env); return UNDEF if v is SHORT else v
            return oc
        if k == "new":
            callee = self.expr(node[1]); args = self.compile_args(node[2])
            return lambda env: interp.construct(callee(env), args(env))
        if k == "super":
            raise JSSyntaxError("bare super")
        raise JSSyntaxError(f"unsupported expression {k}")
    def compile_args(self, args):
        if not any(a[0] == "spread" for a in args):
            cs = [self.expr(a) for a in args]
            if len(cs) == 0: return lambda env: []
            if len(cs) == 1:
                c0 = cs[0]; return lambda env: [c0(env)]
            if len(cs) == 2:
                c0, c1 = cs; return lambda env: [c0(env), c1(env)]
            if len(cs) == 3:
                c0, c1, c2 = cs; return lambda env: [c0(env), c1(env), c2(env)]
            return lambda env: [c(env) for c in cs]
        cs = [(True, self.expr(a[1])) if a[0] == "spread" else (False, self.expr(a)) for a in args]
        def mk(env):
            out = []
            for sp, c in cs:
                if sp: out.extend(c(env).items)
                else: out.append(c(env))
            return out
        return mk
    def compile_call(self, node):
        callee, args, optional = node[1], self.compile_args(node[2]), node[3]
        interp = self
        if callee[0] == "super":                            # super(...) in a derived constructor
            def sup(env):
                home = env.get("%home"); func = env.get("%func")
                parent = func.props.get("%super")
                a = args(env)
                holder = env.get("%holder")
                if parent.is_class and parent.ctor_kind == "derived":
                    obj = interp.run_ctor(parent, UNDEF, a, holder)
                else:
                    obj = JSObject(holder["proto"])
                    interp.init_fields(parent, obj)
                    if parent.native is not None: parent.native(obj, a)
                    elif parent.body is not None: interp.call(parent, obj, a, new_target=func)
                interp.init_fields(func, obj)
                e = env
                while "%holder" not in e.vars: e = e.parent
                e.vars["this"] = obj
                return UNDEF
            return sup
        if callee[0] == "member" and callee[1][0] == "super":        # super.method(...)
            name = callee[2]
            def supm(env):
                home = env.get("%home")
                m = home.proto.get(name) if home is not None and home.proto is not None else UNDEF
                return interp.call(m, env.get("this"), args(env))
            return supm
        if callee[0] == "member":
            o = self.expr(callee[1]); name = callee[2]; opt_member = callee[3]
            def mcall(env):
                ob = o(env)
                if ob.__class__ is JSObject:
                    f = UNDEF; q = ob
                    while q is not None:
                        p = q.props
                        if name in p: f = p[name]; break
                        q = q.proto
                else:
                    if ob is SHORT or (opt_member and (ob is UNDEF or ob is NULL)): return SHORT
                    f = get_member(ob, name)
                if f.__class__ is not JSFunction:
                    if optional and (f is UNDEF or f is NULL): return SHORT
                    raise JSThrow(make_error("TypeError", f"{name} is not a function"))
                if f.native is not None and f.bound is None: return f.native(ob, args(env))
                return interp.call(f, ob, args(env))
            return mcall
        if callee[0] == "index":
            o = self.expr(callee[1]); key = self.expr(callee[2])
            def icall(env):
                ob = o(env)
                if ob is SHORT: return SHORT
                f = get_member(ob, key(env))
                if optional and (f is UNDEF or f is NULL): return SHORT
                return interp.call(f, ob, args(env))
            return icall
        f = self.expr(callee)
        def call(env):
            fn = f(env)
            if fn is SHORT or (optional and (fn is UNDEF or fn is NULL)): return SHORT
            return interp.call(fn, UNDEF, args(env))
        return call
    def compile_store(self, tgt):
        if tgt[0] == "id":
            name = tgt[1]
            def st(env, v):
                e = env
                while e is not None:
                    d = e.vars
                    if name in d: d[name] = v; return
                    e = e.parent
                raise JSThrow(make_error("ReferenceError", name + " is not defined"))
            return st
        if tgt[0] == "member":
            o = self.expr(tgt[1]); name = tgt[2]
            def sm(env, v):
                ob = o(env)
                if ob.__class__ is JSObject: ob.props[name] = v
                else: set_member(ob, name, v)
            return sm
        if tgt[0] == "index":
            o = self.expr(tgt[1]); key = self.expr(tgt[2])
            return lambda env, v: set_member(o(env), key(env), v)
        raise JSSyntaxError("invalid assignment target")
    def compile_assign(self, node):
        op, tgt, rhs = node[1], node[2], self.expr(node[3])
        if tgt[0] == "array":                               # [a, b] = [c, d]
            stores = [self.compile_store(t) for t in tgt[1]]
            def da(env):
                v = rhs(env)
                for i, st in enumerate(stores): st(env, get_member(v, float(i)))
                return v
            return da
        store = self.compile_store(tgt)
        if op == "=":
            if tgt[0] == "member":
                o = self.expr(tgt[1]); name = tgt[2]
                def am(env):
                    ob = o(env); v = rhs(env)
                    if ob.__class__ is JSObject: ob.props[name] = v
                    else: set_member(ob, name, v)
                    return v
                return am
            def a(env):
                v = rhs(env); store(env, v); return v
            return a
        getter = self.expr(tgt)
        if op in ("||=", "&&=", "??="):
            def la(env):
                cur = getter(env)
                if (op == "||=" and truthy(cur)) or (op == "&&=" and not truthy(cur)) or (op == "??=" and not (cur is UNDEF or cur is NULL)): return cur
                v = rhs(env); store(env, v); return v
            return la
        bop = BINOPS[op[:-1]]
        def ca(env):
            v = bop(getter(env), rhs(env)); store(env, v); return v
        return ca
    def compile_bin(self, node):
        op, a, b = node[1], self.expr(node[2]), self.expr(node[3])
        if op == "+":
            def add(env):
                x = a(env); y = b(env)
                if x.__class__ is float and y.__class__ is float: return x + y
                if x.__class__ is int or y.__class__ is int:                         # BigInt
                    if x.__class__ is int and y.__class__ is int: return x + y
                    raise JSThrow(make_error("TypeError", "Cannot mix BigInt and other types, use explicit conversions"))
                return js_add(x, y)
            return add
        if op == "-":
            def sub(env):
                x = a(env); y = b(env)
                if x.__class__ is float and y.__class__ is float: return x - y
                if x.__class__ is int or y.__class__ is int:                         # BigInt
                    if x.__class__ is int and y.__class__ is int: return x - y
                    raise JSThrow(make_error("TypeError", "Cannot mix BigInt and other types, use explicit conversions"))
                return to_num(x) - to_num(y)
            return sub
        if op == "*":
            def mul(env):
                x = a(env); y = b(env)
                if x.__class__ is float and y.__class__ is float: return x * y
                if x.__class__ is int or y.__class__ is int:                         # BigInt
                    if x.__class__ is int and y.__class__ is int: return x * y
                    raise JSThrow(make_error("TypeError", "Cannot mix BigInt and other types, use explicit conversions"))
                return to_num(x) * to_num(y)
            return mul
        if op == "/":
            def div(env): return js_div(a(env), b(env))
            return div
        if op == "<":
            def lt(env):
                x = a(env); y = b(env)
                if x.__class__ is float and y.__class__ is float: return x < y
                return js_cmp(x, y, "<")
            return lt
        if op == ">":
            def gt(env):
                x = a(env); y = b(env)
                if x.__class__ is float and y.__class__ is float: return x > y
                return js_cmp(x, y, ">")
            return gt
        if op == "===": return lambda env: strict_eq(a(env), b(env))
        if op == "!==": return lambda env: not strict_eq(a(env), b(env))
        if op == "instanceof":
            def inst(env):
                ob = a(env); f = b(env)
                if not isinstance(ob, JSObject) or not isinstance(f, JSFunction): return False
                proto = f.props.get("prototype"); p = ob.proto
                while p is not None:
                    if p is proto: return True
                    p = p.proto
                return False
            return inst
        if op == "in":
            def isin(env):
                key = a(env); ob = b(env)
                if isinstance(ob, (JSArray, JSTyped)) and isinstance(key, float): return 0 <= key < len(ob.items)
                if not isinstance(ob, JSObject): raise JSThrow(make_error("TypeError", "right-hand side of 'in' is not an object"))
                return ob.has(prop_key(key))
            return isin
        f = BINOPS[op]
        g = BIGINT_OPS.get(op)
        def binop(env):
            x = a(env); y = b(env)
            if x.__class__ is int or y.__class__ is int:                  # BigInt (bool is its own class)
                if not (x.__class__ is int and y.__class__ is int) or g is None:
                    raise JSThrow(make_error("TypeError", "Cannot mix BigInt and other types, use explicit conversions"))
                return g(x, y)
            return f(x, y)
        return binop
    def compile_class(self, node):
        _, name, sup, members = node
        sup_e = self.expr(sup) if sup is not None else None
        interp = self
        ctor_node = None
        for m in members:
            if m[0] == "method" and not m[1] and m[2] == ("lit", "constructor"): ctor_node = m[3]
        def mk(env):
            parent = sup_e(env) if sup_e is not None else None
            cenv = Env(env)
            if ctor_node is not None: cls = interp.make_function(ctor_node, cenv)
            else:
                cls = JSFunction(params=[], body=None, env=cenv, name=name or "")
                cls.props["prototype"] = JSObject(OBJECT_PROTO, {"constructor": cls})
                if parent is not None:                      # default derived constructor: constructor(...args) { super(...args); }
                    def default_ctor_body(e):
                        args = list(e.vars["%args"].items)
                        holder = e.vars["%holder"]
                        if parent.is_class and parent.ctor_kind == "derived": obj = interp.run_ctor(parent, UNDEF, args, holder)
                        else:
                            obj = JSObject(holder["proto"]); interp.init_fields(parent, obj)
                            if parent.native is not None: parent.native(obj, args)
                            elif parent.body is not None: interp.call(parent, obj, args, new_target=cls)
                        interp.init_fields(cls, obj)
                        e.vars["this"] = obj
                        return None
                    cls.params = [(("name", "%args"), None, True)]; cls.body = default_ctor_body
            cls.name = name or cls.name; cls.is_class = True
            proto = cls.props["prototype"]
            cls.home = proto
            if parent is not None:
                if not isinstance(parent, JSFunction): raise JSThrow(make_error("TypeError", "class extends value is not a constructor"))
                cls.ctor_kind = "derived"; cls.props["%super"] = parent
                proto.proto = parent.props.get("prototype", OBJECT_PROTO)
                cls.proto = parent                          # static inheritance
            fields = []
            if name: cenv.vars[name] = cls
            for m in members:
                kind, static, key, val, accessor = m
                kname = key[1] if key[0] == "lit" else prop_key(interp.expr(key[1])(cenv))
                if kind == "method":
                    if not static and kname == "constructor": continue
                    f = interp.make_function(val, cenv, home=(cls if static else proto))
                    f.home = cls if static else proto
                    (cls if static else proto).props[kname] = f
                else:
                    init = interp.expr(val) if val is not None else None
                    if static: cls.props[kname] = init(Env_with_this(cenv, cls)) if init is not None else UNDEF
                    else: fields.append((kname, init, cenv))
            cls.props["%fields"] = fields
            return cls
        return mk


def getattr_fields(cls):
    return cls.props.get("%fields", []) if isinstance(cls, JSFunction) else []


def Env_with_this(env, this_val):
    e = Env(env); e.vars["this"] = this_val; return e


def js_add(x, y):
    if isinstance(x, JSObject) and not isinstance(x, JSFunction): x = to_str(x)
    if isinstance(y, JSObject) and not isinstance(y, JSFunction): y = to_str(y)
    if isinstance(x, str) or isinstance(y, str): return to_str(x) + to_str(y)
    return to_num(x) + to_num(y)


def js_div(x, y):
    x = x if x.__class__ is float else to_num(x); y = y if y.__class__ is float else to_num(y)
    if y == 0.0:
        if x != x or x == 0.0: return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)
    try: return x / y
    except OverflowError: return math.copysign(math.inf, x) * math.copysign(1.0, y)


def js_mod(x, y):
    x, y = to_num(x), to_num(y)
    if y == 0.0 or x != x or y != y or x in (math.inf, -math.inf): return math.nan
    if y in (math.inf, -math.inf): return x
    return math.fmod(x, y)


def js_pow(x, y):
    x, y = to_num(x), to_num(y)
    if y != y: return math.nan
    if y == 0.0: return 1.0
    # Math.pow is "implementation-approximated" (ECMA-262 21.3.2.26).  For an exponent of exactly 2 this interpreter returns x * x:
    # what V8's optimizing compiler emits for Math.pow(x, 2) (Float64Pow(x, 2.0) is reduced to Float64Mul(x, x)) and what GCC makes of
    # the oracle's std::pow(x, 2).  glibc's pow(x, 2.0) differs from the exactly rounded square in the last ulp for a few arguments
    # in 10^5 (found by tools/fuzz_render.py: one channel of one pixel in 440 random scenes, js/world.js:89 and :106).
    if y == 2.0: return x * x
    if (x == 1.0 or x == -1.0) and y in (math.inf, -math.inf): return math.nan
    try: return math.pow(x, y)
    except OverflowError: return math.inf if x > 0 or int(y) % 2 == 0 else -math.inf
    except ValueError:
        if x == 0.0 and y < 0: return math.inf
        return math.nan


def js_cmp(x, y, op):
    if isinstance(x, str) and isinstance(y, str): return {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]
    a, b = to_num(x), to_num(y)
    return {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]


BINOPS = {
    "+": js_add, "-": lambda x, y: to_num(x) - to_num(y), "*": lambda x, y: to_num(x) * to_num(y), "/": js_div, "%": js_mod, "**": js_pow,
    "<": lambda x, y: js_cmp(x, y, "<"), ">": lambda x, y: js_cmp(x, y, ">"), "<=": lambda x, y: js_cmp(x, y, "<="), ">=": lambda x, y: js_cmp(x, y, ">="),
    "==": loose_eq, "!=": lambda x, y: not loose_eq(x, y), "===": strict_eq, "!==": lambda x, y: not strict_eq(x, y),
    "&": lambda x, y: float(to_int32(x) & to_int32(y)), "|": lambda x, y: float(to_int32(x) | to_int32(y)), "^": lambda x, y: float(to_int32(x) ^ to_int32(y)),
    "<<": lambda x, y: float(to_int32(to_int32(x) << (to_int32(y) & 31))), ">>": lambda x, y: float(to_int32(x) >> (to_int32(y) & 31)),
    ">>>": lambda x, y: float((to_int32(x) & 0xFFFFFFFF) >> (to_int32(y) & 31)),
}


BIGINT_OPS = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "&": lambda x, y: x & y, "|": lambda x, y: x | y, "^": lambda x, y: x ^ y,
              "<<": lambda x, y: x << y, ">>": lambda x, y: x >> y, "==": lambda x, y: x == y, "!=": lambda x, y: x != y,
              "<": lambda x, y: x < y, ">": lambda x, y: x > y, "<=": lambda x, y: x <= y, ">=": lambda x, y: x >= y}


# ------------------------------------------------------------------------------------------------ globals / built-ins
def py_to_js(v):
    """JSON-like Python data -> JS values"""
    if v is None: return NULL
    if isinstance(v, bool): return v
    if isinstance(v, (int, float)): return float(v)
    if isinstance(v, str): return v
    if isinstance(v, (list, tuple)): return JSArray([py_to_js(x) for x in v])
    if isinstance(v, dict): return JSObject(OBJECT_PROTO, {str(k): py_to_js(x) for k, x in v.items()})
    raise TypeError(type(v))


def js_to_py(v):
    if v is UNDEF or v is NULL: return None
    if isinstance(v, (bool, str)): return v
    if isinstance(v, float): return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
    if isinstance(v, (JSArray, JSTyped)): return [js_to_py(x) for x in v.items]
    if isinstance(v, JSFunction): return None
    if isinstance(v, JSObject): return {k: js_to_py(x) for k, x in v.props.items() if not isinstance(x, JSFunction)}
    return v


def install_globals(interp):
    g = interp.globals.vars
    call = interp.call
    def num_fn(f):
        def w(this, a):
            try: return float(f(*[to_num(x) for x in a]))
            except (ValueError, OverflowError): return math.nan
        return native(w, f.__name__)
    def js_round(x):
        return x if x != x or x in (math.inf, -math.inf) else float(math.floor(x + 0.5))
    def js_max(this, a):
        r = -math.inf
        for x in a:
            x = to_num(x)
            if x != x: return math.nan
            if x > r or (x == 0.0 and r == 0.0 and math.copysign(1.0, r) < 0): r = x
        return r
    def js_min(this, a):
        r = math.inf
        for x in a:
            x = to_num(x)
            if x != x: return math.nan
            if x < r or (x == 0.0 and r == 0.0 and math.copysign(1.0, x) < 0): r = x
        return r
    def js_sqrt(this, a):
        x = to_num(a[0]) if a else math.nan
        return math.sqrt(x) if x >= 0 else (math.nan if x == x and x < 0 or x != x else x)
    def js_exp(x):
        try: return math.exp(x)
        except OverflowError: return math.inf
    def floor(x): return x if x != x or x in (math.inf, -math.inf) else float(math.floor(x))
    def ceil(x): return x if x != x or x in (math.inf, -math.inf) else float(math.ceil(x))
    def trunc(x): return x if x != x or x in (math.inf, -math.inf) else float(math.trunc(x))
    def safe(f):
        def w(x):
            if x != x: return math.nan
            try: return f(x)
            except (ValueError, OverflowError): return math.nan
        w.__name__ = f.__name__; return w
    Math = JSObject(OBJECT_PROTO, {
        "PI": math.pi, "E": math.e, "SQRT2": math.sqrt(2.0), "LN2": math.log(2.0),
        "sqrt": native(js_sqrt, "sqrt"), "abs": num_fn(abs), "floor": num_fn(floor), "ceil": num_fn(ceil), "round": num_fn(js_round), "trunc": num_fn(trunc),
        "max": native(js_max, "max"), "min": native(js_min, "min"), "pow": native(lambda t, a: js_pow(a[0], a[1]), "pow"),
        "exp": num_fn(js_exp), "log": num_fn(safe(lambda x: math.log(x) if x > 0 else (-math.inf if x == 0 else math.nan))),
        "sin": num_fn(safe(lambda x: math.sin(x) if x not in (math.inf, -math.inf) else math.nan)),
        "cos": num_fn(safe(lambda x: math.cos(x) if x not in (math.inf, -math.inf) else math.nan)),
        "tan": num_fn(safe(lambda x: math.tan(x) if x not in (math.inf, -math.inf) else math.nan)),
        "atan": num_fn(math.atan), "atan2": num_fn(math.atan2), "acos": num_fn(safe(math.acos)), "asin": num_fn(safe(math.asin)),
        "sign": num_fn(lambda x: x if x != x or x == 0 else math.copysign(1.0, x)), "hypot": num_fn(math.hypot),
        "random": native(lambda t, a: __import__("random").random(), "random"),
    })
    g["Math"] = Math
    g["Infinity"] = math.inf; g["NaN"] = math.nan; g["undefined"] = UNDEF
    g["globalThis"] = JSObject(OBJECT_PROTO, g)           # the global object: its properties ARE the global bindings
    quiet = lambda t, a: UNDEF
    interp.console_lines = []
    def log(t, a): interp.console_lines.append(" ".join(to_str(x) for x in a)); return UNDEF
    g["console"] = JSObject(OBJECT_PROTO, {k: native(log, k) for k in ("log", "warn", "error", "info", "debug", "table", "group", "groupEnd", "time", "timeEnd")})
    clock = {"t": 0.0}
    def now(t, a): clock["t"] += 0.001; return clock["t"]                       # a clock that barely moves: no 50 ms yields are taken
    g["performance"] = JSObject(OBJECT_PROTO, {"now": native(now, "now")})
    g["Date"] = JSObject(OBJECT_PROTO, {"now": native(now, "now")})
    g["window"] = JSObject(OBJECT_PROTO, {"renderCancelled": False})
    g["isNaN"] = native(lambda t, a: to_num(a[0] if a else UNDEF) != to_num(a[0] if a else UNDEF), "isNaN")
    g["isFinite"] = native(lambda t, a: math.isfinite(to_num(a[0] if a else UNDEF)), "isFinite")
    g["parseFloat"] = native(lambda t, a: to_num(to_str(a[0])), "parseFloat")
    def parse_int(t, a):
        m = re.match(r"\s*[+-]?\d+", to_str(a[0]))
        return float(int(m.group(0))) if m else math.nan
    g["parseInt"] = native(parse_int, "parseInt")
    g["Number"] = native(lambda t, a: to_num(a[0]) if a else 0.0, "Number")
    def bigint(t, a):
        v = a[0] if a else UNDEF
        if v.__class__ is int: return v
        x = to_num(v)
        if x != x or x in (math.inf, -math.inf) or x != math.floor(x): raise JSThrow(make_error("RangeError", "The number cannot be converted to a BigInt because it is not an integer"))
        return int(x)
    g["BigInt"] = native(bigint, "BigInt")
    g["Number"].props.update({"isFinite": native(lambda t, a: isinstance(a[0], float) and math.isfinite(a[0])), "isInteger": native(lambda t, a: isinstance(a[0], float) and math.isfinite(a[0]) and a[0] == int(a[0])),
                              "isNaN": native(lambda t, a: isinstance(a[0], float) and a[0] != a[0]), "EPSILON": 2.220446049250313e-16, "MAX_VALUE": 1.7976931348623157e308,
                              "MAX_SAFE_INTEGER": 9007199254740991.0, "POSITIVE_INFINITY": math.inf, "NEGATIVE_INFINITY": -math.inf})
    g["String"] = native(lambda t, a: to_str(a[0]) if a else "", "String")
    g["Boolean"] = native(lambda t, a: truthy(a[0]) if a else False, "Boolean")
    # Error
    def error_ctor(this, a):
        e = this if isinstance(this, JSObject) else JSObject(ERROR_PROTO)
        msg = to_str(a[0]) if a and a[0] is not UNDEF else ""
        e.props["message"] = msg; e.props["name"] = "Error"; e.props["stack"] = "Error: " + msg
        return e
    Error = native(error_ctor, "Error"); Error.props["prototype"] = ERROR_PROTO; ERROR_PROTO.props["constructor"] = Error
    ERROR_PROTO.props["toString"] = native(lambda t, a: to_str(t.get("name")) + ": " + to_str(t.get("message")))
    g["Error"] = Error; g["TypeError"] = Error; g["RangeError"] = Error
    # Object
    Obj = native(lambda t, a: JSObject(OBJECT_PROTO), "Object")
    def keys(t, a):
        o = a[0]
        if isinstance(o, (JSArray, JSTyped)): return JSArray([num_to_str(float(i)) for i in range(len(o.items))])
        return JSArray([k for k in o.props.keys() if not k.startswith("%")])
    def assign(t, a):
        for s in a[1:]:
            if isinstance(s, JSObject): a[0].props.update(s.props)
        return a[0]
    Obj.props.update({"keys": native(keys, "keys"), "assign": native(assign, "assign"),
                      "values": native(lambda t, a: JSArray([v for k, v in a[0].props.items() if not k.startswith("%")])),
                      "entries": native(lambda t, a: JSArray([JSArray([k, v]) for k, v in a[0].props.items() if not k.startswith("%")])),
                      "getPrototypeOf": native(lambda t, a: a[0].proto if a[0].proto is not None else NULL),
                      "freeze": native(lambda t, a: a[0]), "prototype": OBJECT_PROTO})
    OBJECT_PROTO.props["hasOwnProperty"] = native(lambda t, a: isinstance(t, JSObject) and prop_key(a[0]) in t.props, "hasOwnProperty")
    OBJECT_PROTO.props["toString"] = native(lambda t, a: to_str(t), "toString")
    g["Object"] = Obj
    # Function.prototype
    def fn_call(t, a): return call(t, a[0] if a else UNDEF, list(a[1:]))
    def fn_apply(t, a): return call(t, a[0] if a else UNDEF, list(a[1].items) if len(a) > 1 and isinstance(a[1], (JSArray, JSTyped)) else [])
    def fn_bind(t, a):
        b = JSFunction(name="bound " + t.name); b.bound = (t, a[0] if a else UNDEF, list(a[1:])); return b
    FUNCTION_PROTO.props.update({"call": native(fn_call, "call"), "apply": native(fn_apply, "apply"), "bind": native(fn_bind, "bind")})
    # Array
    def arr_ctor(t, a):
        if len(a) == 1 and isinstance(a[0], float): return JSArray([UNDEF] * int(a[0]))
        return JSArray(list(a))
    Arr = native(arr_ctor, "Array"); Arr.props["prototype"] = ARRAY_PROTO
    Arr.props["isArray"] = native(lambda t, a: isinstance(a[0], JSArray) if a else False, "isArray")
    def arr_from(t, a):
        src = a[0]; items = list(src.items) if isinstance(src, (JSArray, JSTyped)) else list(src) if isinstance(src, str) else [UNDEF] * int(to_num(get_member(src, "length")))
        if len(a) > 1: items = [call(a[1], UNDEF, [x, float(i)]) for i, x in enumerate(items)]
        return JSArray(items)
    Arr.props["from"] = native(arr_from, "from")
    g["Array"] = Arr
    def push(t, a): t.items.extend(a); return float(len(t.items))
    def arr_map(t, a): return JSArray([call(a[0], UNDEF, [x, float(i), t]) for i, x in enumerate(list(t.items))])
    def arr_foreach(t, a):
        for i, x in enumerate(list(t.items)): call(a[0], UNDEF, [x, float(i), t])
        return UNDEF
    def arr_filter(t, a): return JSArray([x for i, x in enumerate(list(t.items)) if truthy(call(a[0], UNDEF, [x, float(i), t]))])
    def arr_reduce(t, a):
        items = list(t.items); i = 0
        if len(a) > 1: acc = a[1]
        else: acc = items[0]; i = 1
        while i < len(items): acc = call(a[0], UNDEF, [acc, items[i], float(i), t]); i += 1
        return acc
    def arr_slice(t, a):
        n = len(t.items); s = int(to_num(a[0])) if a and a[0] is not UNDEF else 0; e = int(to_num(a[1])) if len(a) > 1 and a[1] is not UNDEF else n
        if s < 0: s += n
        if e < 0: e += n
        out = JSArray(list(t.items[max(s, 0):max(e, 0)])) if isinstance(t, JSArray) else None
        if out is None:
            out = JSTyped(t.kind, 0); out.items = list(t.items[max(s, 0):max(e, 0)])
        return out
    def arr_fill(t, a):
        for i in range(len(t.items)):
            if isinstance(t, JSTyped): t.store(i, a[0])
            else: t.items[i] = a[0]
        return t
    def arr_index_of(t, a):
        for i, x in enumerate(t.items):
            if strict_eq(x, a[0]): return float(i)
        return -1.0
    def arr_join(t, a):
        sep = to_str(a[0]) if a and a[0] is not UNDEF else ","
        return sep.join("" if x is UNDEF or x is NULL else to_str(x) for x in t.items)
    def arr_some(t, a): return any(truthy(call(a[0], UNDEF, [x, float(i), t])) for i, x in enumerate(list(t.items)))
    def arr_every(t, a): return all(truthy(call(a[0], UNDEF, [x, float(i), t])) for i, x in enumerate(list(t.items)))
    def arr_find(t, a):
        for i, x in enumerate(list(t.items)):
            if truthy(call(a[0], UNDEF, [x, float(i), t])): return x
        return UNDEF
    def arr_concat(t, a):
        out = list(t.items)
        for x in a: out.extend(x.items) if isinstance(x, JSArray) else out.append(x)
        return JSArray(out)
    def arr_set(t, a):
        off = int(to_num(a[1])) if len(a) > 1 else 0
        for i, x in enumerate(a[0].items): t.store(off + i, x)
        return UNDEF
    methods = {"push": push, "map": arr_map, "forEach": arr_foreach, "filter": arr_filter, "reduce": arr_reduce, "slice": arr_slice, "fill": arr_fill,
               "indexOf": arr_index_of, "includes": lambda t, a: arr_index_of(t, a) >= 0, "join": arr_join, "some": arr_some, "every": arr_every,
               "find": arr_find, "concat": arr_concat, "pop": lambda t, a: t.items.pop() if t.items else UNDEF,
               "shift": lambda t, a: t.items.pop(0) if t.items else UNDEF, "reverse": lambda t, a: (t.items.reverse(), t)[1]}
    for k, f in methods.items(): ARRAY_PROTO.props[k] = native(f, k)
    for k in ("map", "forEach", "slice", "fill", "indexOf", "join", "reduce", "some", "every"): TYPED_PROTO.props[k] = native(methods[k], k)
    TYPED_PROTO.props["set"] = native(arr_set, "set")
    def typed_ctor(kind):
        def ctor(t, a):
            src = a[0] if a else 0.0
            if isinstance(src, float): return JSTyped(kind, src)
            arr = JSTyped(kind, len(src.items))
            for i, x in enumerate(src.items): arr.store(i, x)
            return arr
        def typed_from(t, a):                              # %TypedArray%.from(arrayLike | iterable, mapFn)
            return ctor(t, [arr_from(t, a)])
        f = native(ctor, kind); f.props["prototype"] = TYPED_PROTO; f.props["from"] = native(typed_from, "from"); return f
    g["Float32Array"] = typed_ctor("f32"); g["Float64Array"] = typed_ctor("f64"); g["Uint8ClampedArray"] = typed_ctor("u8c")
    g["Uint8Array"] = typed_ctor("u8")
    for nm in ("Float32Array", "Float64Array", "Uint8ClampedArray", "Uint8Array"): g[nm].name = nm
    # String.prototype
    STRING_PROTO.props.update({
        "toLowerCase": native(lambda t, a: t.lower()), "toUpperCase": native(lambda t, a: t.upper()), "trim": native(lambda t, a: t.strip()),
        "includes": native(lambda t, a: to_str(a[0]) in t), "startsWith": native(lambda t, a: t.startswith(to_str(a[0]))),
        "endsWith": native(lambda t, a: t.endswith(to_str(a[0]))), "indexOf": native(lambda t, a: float(t.find(to_str(a[0])))),
        "split": native(lambda t, a: JSArray(list(t) if a and a[0] == "" else t.split(to_str(a[0])))),
        "slice": native(lambda t, a: t[int(to_num(a[0])):(int(to_num(a[1])) if len(a) > 1 and a[1] is not UNDEF else None)]),
        "replace": native(lambda t, a: t.replace(to_str(a[0]), to_str(a[1]), 1)), "charAt": native(lambda t, a: t[int(to_num(a[0]))] if 0 <= int(to_num(a[0])) < len(t) else ""),
        "toString": native(lambda t, a: t),
    })
    # Promise / timers: everything resolves at once, `await` unwraps synchronously
    def promise_ctor(t, a):
        p = JSObject(PROMISE_PROTO, {"%state": "pending", "%value": UNDEF})
        def resolve(_t, args): p.props["%state"] = "fulfilled"; p.props["%value"] = args[0] if args else UNDEF; return UNDEF
        def reject(_t, args): p.props["%state"] = "rejected"; p.props["%value"] = args[0] if args else UNDEF; return UNDEF
        call(a[0], UNDEF, [native(resolve, "resolve"), native(reject, "reject")])
        return p
    P = native(promise_ctor, "Promise"); P.props["prototype"] = PROMISE_PROTO
    P.props["resolve"] = native(lambda t, a: JSObject(PROMISE_PROTO, {"%state": "fulfilled", "%value": a[0] if a else UNDEF}))
    def settle(fn, args):
        """runs a reaction now (everything is already settled here): a returned promise is adopted, a throw rejects"""
        try: r = call(fn, UNDEF, args)
        except JSThrow as e: return JSObject(PROMISE_PROTO, {"%state": "rejected", "%value": e.value})
        if isinstance(r, JSObject) and r.proto is PROMISE_PROTO: return r
        return JSObject(PROMISE_PROTO, {"%state": "fulfilled", "%value": r})
    def then(t, a):
        if t.props.get("%state") == "rejected":
            return settle(a[1], [t.props.get("%value", UNDEF)]) if len(a) > 1 and isinstance(a[1], JSFunction) else t
        if a and isinstance(a[0], JSFunction): return settle(a[0], [t.props.get("%value", UNDEF)])
        return t
    def catch(t, a):
        if t.props.get("%state") == "rejected" and a and isinstance(a[0], JSFunction): return settle(a[0], [t.props.get("%value", UNDEF)])
        return t
    def fin(t, a):
        if a and isinstance(a[0], JSFunction):
            r = settle(a[0], [])
            if r.props.get("%state") == "rejected": return r
        return t
    PROMISE_PROTO.props["then"] = native(then, "then")
    PROMISE_PROTO.props["catch"] = native(catch, "catch")
    PROMISE_PROTO.props["finally"] = native(fin, "finally")
    g["Promise"] = P
    def set_timeout(t, a): call(a[0], UNDEF, list(a[2:])); return 0.0
    g["setTimeout"] = native(set_timeout, "setTimeout"); g["requestAnimationFrame"] = native(set_timeout, "requestAnimationFrame")
    g["clearTimeout"] = native(quiet)
    # intervals never fire by themselves (there is no event loop): the host ticks interp.intervals when it wants time to pass
    interp.intervals = {}
    def set_interval(t, a):
        k = float(len(interp.intervals) + 1 + getattr(interp, "_interval_base", 0)); interp._interval_base = k; interp.intervals[k] = a[0]; return k
    def clear_interval(t, a): interp.intervals.pop(a[0] if a else UNDEF, None); return UNDEF
    g["setInterval"] = native(set_interval, "setInterval"); g["clearInterval"] = native(clear_interval, "clearInterval")
    def json_parse(t, a): return py_to_js(__import__("json").loads(to_str(a[0])))
    def json_stringify(t, a): return __import__("json").dumps(js_to_py(a[0]))
    g["JSON"] = JSObject(OBJECT_PROTO, {"parse": native(json_parse, "parse"), "stringify": native(json_stringify, "stringify")})
