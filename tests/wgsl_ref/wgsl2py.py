"""A small WGSL -> Python transpiler for the subset the reference's LBM shaders use.

Purpose (test infrastructure only): the reference cannot be executed here (no Rust, no WebGPU
backend), so its shader SOURCE TEXT is the only executable ground truth.  This module parses the
unmodified .wgsl files of /root/reference (after the reference's own ``#include`` expansion,
simuverse/src/util/shader.rs:86-115) and emits Python that evaluates them with IEEE f32 scalars
(numpy.float32, one rounding per operation, no FMA).  tests/golden/make_wgsl_golden.py runs the
emitted code to produce golden vectors; nothing here is used by the product.

Supported: struct / const / global var / fn declarations, let / var, assignment (incl. swizzle and
indexed targets), if / else, for, return, continue, break, calls, constructors (vecN<T>, array<T,N>,
structs), the operators of the shaders and the builtins dot, clamp, floor, abs, min, max, length,
smoothstep, textureLoad, textureStore.
"""
import os
import re

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>(?:0[xX][0-9a-fA-F]+[iu]?)|(?:(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fhiu]?))
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>->|==|!=|<=|>=|&&|\|\||<<|>>|[-+*/%<>=!&|^~(){}\[\];:,.@])
""", re.X | re.S)

TYPE_CTORS = {"vec2", "vec3", "vec4", "array", "mat2x2", "mat3x3", "mat4x4"}


def preprocess(entry, wgsl_root):
    """The reference's include expansion: a line starting with ``#include `` is replaced by the file(s)."""
    out = []

    def expand(text):
        for line in text.splitlines():
            if line.startswith("#include "):
                for imp in line[len("#include "):].split(","):
                    path = os.path.join(wgsl_root, imp.strip().replace('"', ""))
                    expand(open(path, encoding="utf-8").read())
            else:
                out.append(line)

    expand(open(os.path.join(wgsl_root, entry), encoding="utf-8").read())
    return "\n".join(out)


def tokenize(src):
    toks, pos = [], 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"bad character {src[pos]!r} at {pos}")
        pos = m.end()
        if m.lastgroup != "ws":
            toks.append((m.lastgroup, m.group(m.lastgroup)))
    toks.append(("eof", ""))
    return toks


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    # ---- token helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] != "num":
            self.i += 1
            return True
        return False

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise SyntaxError(f"expected {val!r}, got {tok!r} near token {self.i}")
        return tok

    def ident(self):
        tok = self.next()
        if tok[0] != "id":
            raise SyntaxError(f"expected identifier, got {tok!r}")
        return tok[1]

    # ---- declarations
    def parse_module(self):
        decls = []
        while self.peek()[0] != "eof":
            attrs = self.attributes()
            kw = self.peek()[1]
            if kw == "struct":
                decls.append(self.struct())
            elif kw == "fn":
                decls.append(self.function(attrs))
            elif kw in ("const", "let", "override"):
                self.next()
                name = self.ident()
                ty = self.type() if self.accept(":") else None
                self.expect("=")
                decls.append(("const", name, ty, self.expr()))
                self.expect(";")
            elif kw == "var":
                self.next()
                if self.accept("<"):
                    while not self.accept(">"):
                        self.next()
                name = self.ident()
                ty = self.type() if self.accept(":") else None
                init = self.expr() if self.accept("=") else None
                self.expect(";")
                decls.append(("gvar", name, ty, init, attrs))
            elif kw == ";":
                self.next()
            else:
                raise SyntaxError(f"unexpected top-level token {self.peek()!r}")
        return decls

    def attributes(self):
        attrs = []
        while self.accept("@"):
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    args.append(self.next()[1])
            attrs.append((name, args))
        return attrs

    def type(self):
        name = self.ident()
        args = []
        if self.accept("<"):
            while True:
                if self.peek()[0] == "num":
                    args.append(("n", int(re.sub(r"[iu]$", "", self.next()[1]))))
                else:
                    args.append(self.type())
                if self.accept(">"):
                    break
                self.expect(",")
        return ("type", name, args)

    def struct(self):
        self.expect("struct")
        name = self.ident()
        self.expect("{")
        fields = []
        while not self.accept("}"):
            self.attributes()
            fname = self.ident()
            self.expect(":")
            fields.append((fname, self.type()))
            self.accept(",")
        self.accept(";")
        return ("struct", name, fields)

    def function(self, attrs):
        self.expect("fn")
        name = self.ident()
        self.expect("(")
        params = []
        while not self.accept(")"):
            self.attributes()
            pname = self.ident()
            self.expect(":")
            params.append((pname, self.type()))
            self.accept(",")
        ret = None
        if self.accept("->"):
            self.attributes()
            ret = self.type()
        return ("fn", name, params, ret, self.block(), attrs)

    # ---- statements
    def block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.statement())
        return stmts

    def statement(self):
        kw = self.peek()[1]
        if kw == "{":
            return ("block", self.block())
        if kw in ("var", "let", "const"):
            s = self.var_decl()
            self.expect(";")
            return s
        if kw == "if":
            return self.if_stmt()
        if kw == "for":
            self.next()
            self.expect("(")
            init = None if self.peek()[1] == ";" else (self.var_decl() if self.peek()[1] in ("var", "let") else self.simple())
            self.expect(";")
            cond = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            upd = None if self.peek()[1] == ")" else self.simple()
            self.expect(")")
            return ("for", init, cond, upd, self.block())
        if kw == "return":
            self.next()
            e = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            return ("return", e)
        if kw in ("continue", "break"):
            self.next()
            self.expect(";")
            return (kw,)
        s = self.simple()
        self.expect(";")
        return s

    def var_decl(self):
        self.next()
        name = self.ident()
        ty = self.type() if self.accept(":") else None
        init = self.expr() if self.accept("=") else None
        return ("var", name, ty, init)

    def if_stmt(self):
        self.expect("if")
        cond = self.expr()
        then = self.block()
        other = None
        if self.accept("else"):
            other = [self.if_stmt()] if self.peek()[1] == "if" else self.block()
        return ("if", cond, then, other)

    def simple(self):
        lhs = self.expr()
        if self.accept("="):
            return ("assign", lhs, self.expr())
        for op in ("+", "-", "*", "/"):
            if self.peek()[1] == op and self.peek(1)[1] == "=":
                self.next(); self.next()
                return ("assign", lhs, ("bin", op, lhs, self.expr()))
        return ("expr", lhs)

    # ---- expressions (precedence climbing)
    LEVELS = [["||"], ["&&"], ["|"], ["^"], ["&"], ["==", "!="], ["<", ">", "<=", ">="], ["<<", ">>"], ["+", "-"],
              ["*", "/", "%"]]

    def expr(self, level=0):
        if level == len(self.LEVELS):
            return self.unary()
        lhs = self.expr(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[level] and not (
                self.peek()[1] in ("+", "-", "*", "/") and self.peek(1) == ("op", "=")):
            op = self.next()[1]
            lhs = ("bin", op, lhs, self.expr(level + 1))
        return lhs

    def unary(self):
        if self.accept("-"):
            return ("neg", self.unary())
        if self.accept("!"):
            return ("not", self.unary())
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            if self.accept("."):
                e = ("member", e, self.ident())
            elif self.accept("["):
                e = ("index", e, self.expr())
                self.expect("]")
            else:
                return e

    def primary(self):
        kind, val = self.peek()
        if kind == "num":
            self.next()
            return ("num", val)
        if val == "(":
            self.next()
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if kind == "id":
            if val in ("true", "false"):
                self.next()
                return ("bool", val == "true")
            if val in TYPE_CTORS and self.peek(1)[1] == "<":
                ty = self.type()
                return ("ctor", ty, self.args())
            self.next()
            if self.peek()[1] == "(":
                return ("call", val, self.args())
            return ("name", val)
        raise SyntaxError(f"unexpected token {self.peek()!r} in expression (token {self.i})")

    def args(self):
        self.expect("(")
        out = []
        while not self.accept(")"):
            out.append(self.expr())
            self.accept(",")
        return out


# ------------------------------------------------------------------------------------------ emitter

SCALAR_CONV = {"f32": "_rt.f32", "i32": "_rt.i32", "u32": "_rt.u32", "bool": "bool"}
BUILTINS = {"dot", "clamp", "floor", "abs", "min", "max", "length", "smoothstep", "textureLoad", "textureStore", "sqrt",
            "atan2", "cos", "sin", "ceil", "round", "select", "fract", "mix", "textureSample"}
PY_KEYWORDS = {"in", "is", "from", "pass", "def", "class", "lambda", "with", "as", "not", "and", "or", "global", "del"}


def pyname(n):
    """WGSL identifiers that are Python keywords (`in: VertexOutput` of the fragment shaders) get a trailing underscore."""
    return n + "_" if n in PY_KEYWORDS else n


class Emitter:
    def __init__(self, decls):
        self.decls = decls
        self.structs = {d[1] for d in decls if d[0] == "struct"}
        self.lines = []
        self.loop_updates = []

    def emit(self):
        w = self.lines.append
        w("# generated by tests/wgsl_ref/wgsl2py.py from the reference's WGSL source")
        for d in self.decls:
            if d[0] == "struct":
                fields = [f for f, _ in d[2]]
                zeros = ", ".join(f"{f}={self.zero(t)}" for f, t in d[2])
                w(f"{d[1]} = _rt.make_struct({d[1]!r}, {fields!r}, lambda: dict({zeros}))")
            elif d[0] == "const":
                w(f"{d[1]} = {self.conv(d[2], self.ex(d[3]))}")
            elif d[0] == "gvar":
                w(f"# resource {d[1]} is bound by the harness")
            elif d[0] == "fn":
                self.function(d)
        return "\n".join(self.lines) + "\n"

    def conv(self, ty, code):
        if ty and ty[1] in SCALAR_CONV:
            return f"{SCALAR_CONV[ty[1]]}({code})"
        return code

    def zero(self, ty):
        name, args = ty[1], ty[2]
        if name == "f32":
            return "_rt.F(0)"
        if name in ("i32", "u32"):
            return "0"
        if name == "bool":
            return "False"
        if name in ("vec2", "vec3", "vec4"):
            n = int(name[3])
            return f"_rt.Vec([{self.zero(args[0])}] * {n})"
        if name == "array":
            if len(args) < 2:
                return "[]"
            return f"[{self.zero(args[0])} for _ in range({args[1][1]})]"
        if name in self.structs:
            return f"{name}()"
        return "None"

    def function(self, d):
        _, name, params, ret, body, attrs = d
        self.lines.append(f"def {name}({', '.join(pyname(p) for p, _ in params)}):")
        n0 = len(self.lines)
        self.block(body, 1)
        if len(self.lines) == n0:
            self.lines.append("    pass")
        self.lines.append("")

    def block(self, stmts, ind):
        pad = "    " * ind
        if not stmts:
            self.lines.append(pad + "pass")
        for s in stmts:
            k = s[0]
            if k == "block":
                self.block(s[1], ind)
            elif k == "var":
                _, name, ty, init = s
                if init is None:
                    self.lines.append(f"{pad}{pyname(name)} = {self.zero(ty)}")
                else:
                    self.lines.append(f"{pad}{pyname(name)} = _rt.copy({self.conv(ty, self.ex(init))})")
            elif k == "assign":
                self.lines.append(f"{pad}{self.lvalue(s[1])} = _rt.copy({self.ex(s[2])})")
            elif k == "expr":
                self.lines.append(f"{pad}{self.ex(s[1])}")
            elif k == "return":
                self.lines.append(f"{pad}return" + (f" {self.ex(s[1])}" if s[1] is not None else ""))
            elif k == "break":
                self.lines.append(f"{pad}break")
            elif k == "continue":
                if self.loop_updates and self.loop_updates[-1] is not None:
                    self.block([self.loop_updates[-1]], ind)
                self.lines.append(f"{pad}continue")
            elif k == "if":
                self.lines.append(f"{pad}if {self.ex(s[1])}:")
                self.block(s[2], ind + 1)
                if s[3] is not None:
                    self.lines.append(f"{pad}else:")
                    self.block(s[3], ind + 1)
            elif k == "for":
                _, init, cond, upd, body = s
                if init is not None:
                    self.block([init], ind)
                self.lines.append(f"{pad}while {self.ex(cond) if cond is not None else 'True'}:")
                self.loop_updates.append(upd)
                self.block(body, ind + 1)
                self.loop_updates.pop()
                if upd is not None:
                    self.block([upd], ind + 1)
            else:
                raise NotImplementedError(k)

    def lvalue(self, e):
        if e[0] == "name":
            return pyname(e[1])
        if e[0] == "member":
            return f"{self.ex(e[1])}.{e[2]}"
        if e[0] == "index":
            return f"{self.ex(e[1])}[{self.ex(e[2])}]"
        if e[0] == "paren":
            return self.lvalue(e[1])
        raise NotImplementedError(f"lvalue {e[0]}")

    def ex(self, e):
        k = e[0]
        if k == "num":
            v = e[1]
            if re.fullmatch(r"0[xX][0-9a-fA-F]+[iu]?|\d+[iu]?", v):
                return str(int(re.sub(r"[iu]$", "", v), 0))
            return f"_rt.F({float(re.sub(r'[fh]$', '', v))!r})"
        if k == "bool":
            return "True" if e[1] else "False"
        if k == "name":
            return pyname(e[1])
        if k == "paren":
            return f"({self.ex(e[1])})"
        if k == "neg":
            return f"(-{self.ex(e[1])})"
        if k == "not":
            return f"(not {self.ex(e[1])})"
        if k == "member":
            return f"{self.ex(e[1])}.{e[2]}"
        if k == "index":
            return f"{self.ex(e[1])}[{self.ex(e[2])}]"
        if k == "bin":
            op, a, b = e[1], self.ex(e[2]), self.ex(e[3])
            if op == "/":
                return f"_rt.div({a}, {b})"
            if op == "%":
                return f"_rt.mod({a}, {b})"
            if op == "&&":
                return f"({a} and {b})"
            if op == "||":
                return f"({a} or {b})"
            return f"({a} {op} {b})"
        if k == "ctor":
            ty, args = e[1], ", ".join(self.ex(a) for a in e[2])
            name, targs = ty[1], ty[2]
            if name in ("vec2", "vec3", "vec4"):
                return f"_rt.vec({int(name[3])}, {targs[0][1]!r}, [{args}])"
            if name == "array":
                return f"[{args}]" if e[2] else self.zero(ty)
            raise NotImplementedError(name)
        if k == "call":
            name, args = e[1], ", ".join(self.ex(a) for a in e[2])
            if name in SCALAR_CONV:
                return f"{SCALAR_CONV[name]}({args})"
            if name in BUILTINS:
                return f"_rt.{name}({args})"
            return f"{name}({args})"  # user function or struct constructor
        raise NotImplementedError(k)


def transpile(entry, wgsl_root):
    src = preprocess(entry, wgsl_root)
    return Emitter(Parser(tokenize(src)).parse_module()).emit()
