"""Rust-subset -> Python transpiler for the reference's HOST helpers on the LBM path.

Why: the image has no Rust toolchain, so the reference's Rust cannot be compiled or run here.  Its host-side arithmetic
on this path is a handful of small functions (mask generator, obstacle / force painters, click guards, uniform
constructor).  Like tests/wgsl_ref does for the shaders, this module EXECUTES THEIR SOURCE TEXT: the function bodies are
cut out of the files under /root/reference/simuverse/src/fluid/, parsed, and emitted as Python that performs the same
operations in the same order on IEEE f32 values (numpy.float32, one rounding per operation) and integers.

Subset: `let` (patterns: name, `mut name`, tuples), assignments and compound assignments, `for x in a..b`, `if` / `else if`
/ `else` (also as an expression), `match` on enum paths, `return` / `continue`, nested `fn`, trailing-expression returns,
struct literals (with field shorthand), array literals incl. `[v; n]`, `vec![..]`, paths (`a::b::c`), method calls,
field access, indexing, references (ignored), `as` casts, the usual operators with Rust's precedence.  Not modelled:
integer overflow (u32 arithmetic is done on unbounded ints; the tests stay inside the ranges the guards allow), borrow
rules, generics, traits.  Third-party pieces the functions call (glam 0.32.1 `Vec2`, `f32` math from std = the
platform libm) are stated in runtime.py.
"""
import re

TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*)
  | (?P<float>\d[\d_]*\.\d[\d_]*(?:[eE][+-]?\d+)?(?:_?f32|_?f64)?)
  | (?P<int>0x[0-9a-fA-F_]+|\d[\d_]*(?:_?(?:u8|u16|u32|u64|usize|i8|i16|i32|i64|isize|f32|f64))?)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<life>'[A-Za-z_]\w*(?!'))
  | (?P<id>[A-Za-z_]\w*)
  | (?P<op>::|->|=>|\.\.=|\.\.|==|!=|<=|>=|&&|\|\||\+=|-=|\*=|/=|<<|>>|[-+*/%<>=!&|^.,;:(){}\[\]#?])
""", re.X)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        k = m.lastgroup
        if k != "ws":
            out.append((k, m.group()))
    out.append(("eof", ""))
    return out


def extract_fn(text, name):
    """Source text of `fn name` (first definition) through its matching closing brace."""
    m = re.search(r"\bfn\s+" + re.escape(name) + r"\b", text)
    if not m:
        raise KeyError(name)
    i = text.index("{", m.end())
    depth, j = 0, i
    in_line_comment = False
    while j < len(text):
        c = text[j]
        if in_line_comment:
            in_line_comment = c != "\n"
        elif text.startswith("//", j):
            in_line_comment = True
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[m.start():j + 1]
        j += 1
    raise SyntaxError("unbalanced braces in " + name)


def extract_item(text, kind, name):
    """`enum Name {...}` / `const NAME: T = expr;`"""
    if kind == "enum":
        m = re.search(r"\benum\s+" + re.escape(name) + r"\s*\{", text)
        return text[m.start():text.index("}", m.end()) + 1]
    m = re.search(r"\bconst\s+" + re.escape(name) + r"\b[^;]*;", text)
    return m.group()


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, val):
        return self.peek()[1] == val and self.peek()[0] in ("op", "id")

    def accept(self, val):
        if self.at(val):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"expected {val!r}, got {self.peek()!r} (token {self.i})")

    def ident(self):
        k, v = self.next()
        if k != "id":
            raise SyntaxError(f"identifier expected, got {v!r}")
        return v

    # ---- types are skipped (Python is untyped); only their extent matters
    def skip_type(self, stops):
        depth = 0
        while True:
            k, v = self.peek()
            if k == "eof":
                return
            if depth == 0 and v in stops and k == "op":
                return
            if v in ("<", "(", "["):
                depth += 1
            elif v in (">", ")", "]"):
                if depth == 0:
                    return
                depth -= 1
            elif v == ">>":
                depth -= 2
            self.next()

    # ---- items
    def enum(self):
        self.expect("enum")
        name = self.ident()
        self.expect("{")
        variants, nxt = {}, 0
        while not self.accept("}"):
            v = self.ident()
            if self.accept("="):
                nxt = int(self.next()[1])
            variants[v] = nxt
            nxt += 1
            self.accept(",")
        return ("enum", name, variants)

    def const(self):
        self.expect("const")
        name = self.ident()
        self.expect(":")
        self.skip_type({"="})
        self.expect("=")
        e = self.expr()
        self.expect(";")
        return ("const", name, e)

    def function(self):
        self.accept("pub")
        self.expect("fn")
        name = self.ident()
        if self.accept("<"):  # generics / lifetimes
            depth = 1
            while depth:
                v = self.next()[1]
                depth += (v == "<") - (v == ">")
        self.expect("(")
        params = []
        while not self.accept(")"):
            while self.at("&") or self.at("mut") or self.peek()[0] == "life":
                self.next()
            p = self.ident()
            if self.accept(":"):
                self.skip_type({",", ")"})
            params.append(p)
            self.accept(",")
        if self.accept("->"):
            self.skip_type({"{"})
        return ("fn", name, params, self.block())

    # ---- statements
    def block(self):
        self.expect("{")
        stmts = []
        while not self.accept("}"):
            stmts.append(self.statement())
        return stmts

    def pattern(self):
        if self.accept("("):
            names = []
            while not self.accept(")"):
                names.append(self.pattern())
                self.accept(",")
            return ("tuple", names)
        self.accept("mut")
        return ("name", self.ident())

    def statement(self):
        k, v = self.peek()
        if v == "let" and k == "id":
            self.next()
            pat = self.pattern()
            if self.accept(":"):
                self.skip_type({"=", ";"})
            init = self.expr() if self.accept("=") else None
            self.expect(";")
            return ("let", pat, init)
        if v in ("fn", "pub") and k == "id":
            return self.function()
        if v == "for" and k == "id":
            self.next()
            pat = self.pattern()
            self.expect("in")
            it = self.expr(no_struct=True)
            return ("for", pat, it, self.block())
        if v == "continue" and k == "id":
            self.next()
            self.expect(";")
            return ("continue",)
        if v == "return" and k == "id":
            self.next()
            e = None if self.at(";") else self.expr()
            self.expect(";")
            return ("return", e)
        if v == "if" and k == "id":
            e = self.if_expr()
            self.accept(";")
            return ("expr", e)
        if v == "match" and k == "id":
            e = self.match_expr()
            self.accept(";")
            return ("expr", e)
        e = self.expr()
        for op in ("=", "+=", "-=", "*=", "/="):
            if self.accept(op):
                rhs = self.expr()
                self.expect(";")
                return ("assign", op, e, rhs)
        if self.accept(";"):
            return ("expr", e)
        if self.at("}"):
            return ("tail", e)  # the block's value
        raise SyntaxError(f"unexpected {self.peek()!r} after expression")

    def if_expr(self):
        self.expect("if")
        cond = self.expr(no_struct=True)
        then = self.block()
        other = None
        if self.accept("else"):
            other = [("expr", self.if_expr())] if self.at("if") else self.block()
        return ("if", cond, then, other)

    def match_expr(self):
        self.expect("match")
        subject = self.expr(no_struct=True)
        self.expect("{")
        arms = []
        while not self.accept("}"):
            if self.accept("_"):
                pat = None
            else:
                pat = self.expr(no_struct=True)
            self.expect("=>")
            body = self.block() if self.at("{") else [("tail", self.expr())]
            self.accept(",")
            arms.append((pat, body))
        return ("match", subject, arms)

    # ---- expressions (Rust precedence: unary > as > * / % > + - > << >> > & > ^ > | > comparisons > && > ||)
    LEVELS = [["||"], ["&&"], ["==", "!=", "<", ">", "<=", ">="], ["|"], ["^"], ["&"], ["<<", ">>"], ["+", "-"],
              ["*", "/", "%"]]

    def expr(self, level=0, no_struct=False, no_range=False):
        if level == 0 and not no_range:  # a..b binds loosest
            lo = self.expr(0, no_struct, True)
            if self.accept(".."):
                return ("range", lo, self.expr(0, no_struct, True))
            return lo
        if level == len(self.LEVELS):
            return self.cast(no_struct)
        lhs = self.expr(level + 1, no_struct, True)
        while self.peek()[0] == "op" and self.peek()[1] in self.LEVELS[level]:
            op = self.next()[1]
            lhs = ("bin", op, lhs, self.expr(level + 1, no_struct, True))
        return lhs

    def cast(self, no_struct):
        e = self.unary(no_struct)
        while self.accept("as"):
            e = ("as", e, self.ident())
        return e

    def unary(self, no_struct):
        if self.accept("-"):
            return ("neg", self.unary(no_struct))
        if self.accept("!"):
            return ("not", self.unary(no_struct))
        if self.accept("&"):
            self.accept("mut")
            return self.unary(no_struct)
        if self.accept("*"):
            return self.unary(no_struct)
        return self.postfix(no_struct)

    def postfix(self, no_struct):
        e = self.primary(no_struct)
        while True:
            if self.accept("."):
                k, v = self.next()
                if k == "int":  # tuple field
                    e = ("index", e, ("int", v))
                elif self.at("("):
                    e = ("method", e, v, self.args())
                else:
                    e = ("field", e, v)
            elif self.accept("["):
                idx = self.expr()
                self.expect("]")
                e = ("index", e, idx)
            elif self.at("(") and e[0] in ("path", "name"):
                e = ("call", e, self.args())
            elif self.accept("?"):
                pass  # Result / Option are modelled by their success value: errors do not occur on this path
            else:
                return e

    def args(self):
        self.expect("(")
        out = []
        while not self.accept(")"):
            out.append(self.expr())
            self.accept(",")
        return out

    def primary(self, no_struct):
        k, v = self.peek()
        if k == "float":
            self.next()
            return ("float", re.sub(r"_?f(32|64)$", "", v).replace("_", ""))
        if k == "int":
            self.next()
            m = re.fullmatch(r"(0x[0-9a-fA-F_]+|\d[\d_]*?)_?(u8|u16|u32|u64|usize|i8|i16|i32|i64|isize|f32|f64)?", v)
            if m.group(2) in ("f32", "f64"):
                return ("float", m.group(1).replace("_", "") + ".0")
            return ("int", m.group(1).replace("_", ""))
        if k == "str":
            self.next()
            return ("str", v)
        if v == "(":
            self.next()
            items = []
            trailing_comma = False
            while not self.accept(")"):
                items.append(self.expr())
                trailing_comma = self.accept(",")
            if len(items) == 1 and not trailing_comma:
                return ("paren", items[0])
            return ("tuple", items)
        if v == "[":
            self.next()
            first = None if self.at("]") else self.expr()
            if first is not None and self.accept(";"):
                n = self.expr()
                self.expect("]")
                return ("repeat", first, n)
            items = [] if first is None else [first]
            while not self.accept("]"):
                self.expect(",")
                if self.at("]"):
                    continue
                items.append(self.expr())
            return ("array", items)
        if v == "if" and k == "id":
            return self.if_expr()
        if v == "match" and k == "id":
            return self.match_expr()
        if v == "|" and k == "op":  # closure |a, b| expr
            self.next()
            params = []
            while not self.accept("|"):
                params.append(self.pattern())
                self.accept(",")
            return ("closure", params, self.expr(no_struct=no_struct))
        if v in ("true", "false") and k == "id":
            self.next()
            return ("bool", v == "true")
        if k == "id":
            segs = [self.ident()]
            while self.at("::"):
                self.next()
                segs.append(self.ident())
            if self.at("!") and self.peek(1)[1] == "[" and segs == ["vec"]:
                self.next()
                self.next()
                items = []
                while not self.accept("]"):
                    items.append(self.expr())
                    self.accept(",")
                return ("array", items)
            if self.at("!") and self.peek(1)[1] == "(":  # other macros (panic!, assert_eq!): name + arguments
                self.next()
                return ("macro", segs[-1], self.args())
            if self.at("{") and not no_struct and segs[-1][0].isupper():
                self.next()
                fields = []
                while not self.accept("}"):
                    f = self.ident()
                    val = self.expr() if self.accept(":") else ("name", f)
                    fields.append((f, val))
                    self.accept(",")
                return ("struct", "::".join(segs), fields)
            return ("name", segs[0]) if len(segs) == 1 else ("path", "::".join(segs))
        raise SyntaxError(f"unexpected token {self.peek()!r}")


PY_KEYWORDS = {"in", "is", "from", "pass", "def", "class", "lambda", "with", "as", "not", "and", "or", "global", "del"}


def pyname(n):
    return n + "_" if n in PY_KEYWORDS else n


class Emitter:
    def __init__(self):
        self.lines = []
        self.tmp = 0

    def fn(self, d, ind=0):
        _, name, params, body = d
        pad = "    " * ind
        self.lines.append(f"{pad}def {pyname(name)}({', '.join(pyname(p) for p in params)}):")
        n0 = len(self.lines)
        self.block(body, ind + 1, tail="return")
        if len(self.lines) == n0:
            self.lines.append(pad + "    pass")

    def pat(self, p):
        return pyname(p[1]) if p[0] == "name" else "(" + ", ".join(self.pat(q) for q in p[1]) + ",)"

    def block(self, stmts, ind, tail=None):
        """tail: what to do with the block's value — None (drop), 'return', or a variable name to assign."""
        pad = "    " * ind
        if not stmts:
            self.lines.append(pad + "pass")
        for s in stmts:
            k = s[0]
            if k == "let":
                if s[2] is None:
                    self.lines.append(f"{pad}{self.pat(s[1])} = None")
                elif s[2][0] == "if" and not self.is_simple_if(s[2]):   # let x = if c { stmts; v } else { stmts; w };
                    self.if_stmt(s[2], ind, self.pat(s[1]))
                elif s[2][0] == "match":
                    self.match_stmt(s[2], ind, self.pat(s[1]))
                else:
                    self.lines.append(f"{pad}{self.pat(s[1])} = {self.ex(s[2])}")
            elif k == "fn":
                self.fn(s, ind)
            elif k == "for":
                self.lines.append(f"{pad}for {self.pat(s[1])} in {self.ex(s[2])}:")
                self.block(s[3], ind + 1)
            elif k == "continue":
                self.lines.append(pad + "continue")
            elif k == "return":
                self.lines.append(pad + "return" + ("" if s[1] is None else " " + self.ex(s[1])))
            elif k == "assign":
                op = s[1]
                if op == "=":
                    self.lines.append(f"{pad}{self.ex(s[2])} = {self.ex(s[3])}")
                else:  # a op= b  ->  a = a op b, through the same operator helper as a plain binary expression
                    self.lines.append(f"{pad}{self.ex(s[2])} = {self.ex(('bin', op[0], s[2], s[3]))}")
            elif k in ("expr", "tail"):
                e = s[1]
                want = tail if k == "tail" else None
                if e[0] == "if" and not self.is_simple_if(e):
                    self.if_stmt(e, ind, want)
                elif e[0] == "match":
                    self.match_stmt(e, ind, want)
                elif want == "return":
                    self.lines.append(f"{pad}return {self.ex(e)}")
                elif want:
                    self.lines.append(f"{pad}{want} = {self.ex(e)}")
                else:
                    self.lines.append(pad + self.ex(e))
            else:
                raise NotImplementedError(k)

    @staticmethod
    def is_simple_if(e):
        """if c { v } else { w } with single-expression blocks: a Python conditional expression"""
        _, _, then, other = e
        return (other is not None and len(then) == 1 and then[0][0] == "tail" and len(other) == 1 and
                other[0][0] == "tail")

    def if_stmt(self, e, ind, tail):
        pad = "    " * ind
        _, cond, then, other = e
        self.lines.append(f"{pad}if {self.ex(cond)}:")
        self.block(then, ind + 1, tail)
        while other is not None:
            if len(other) == 1 and other[0][0] == "expr" and other[0][1][0] == "if":
                _, cond, then, other = other[0][1]
                self.lines.append(f"{pad}elif {self.ex(cond)}:")
                self.block(then, ind + 1, tail)
            else:
                self.lines.append(f"{pad}else:")
                self.block(other, ind + 1, tail)
                other = None

    def match_stmt(self, e, ind, tail):
        pad = "    " * ind
        _, subject, arms = e
        self.tmp += 1
        var = f"_m{self.tmp}"
        self.lines.append(f"{pad}{var} = {self.ex(subject)}")
        first = True
        for pat, body in arms:
            if pat is None:
                self.lines.append(f"{pad}else:" if not first else f"{pad}if True:")
            else:
                self.lines.append(f"{pad}{'if' if first else 'elif'} {var} == {self.ex(pat)}:")
            self.block(body, ind + 1, tail)
            first = False

    def ex(self, e):
        k = e[0]
        if k == "float":
            return f"_rt.F({float(e[1])!r})"
        if k == "int":
            return str(int(e[1], 0))
        if k == "str":
            return e[1]
        if k == "bool":
            return "True" if e[1] else "False"
        if k == "name":
            return pyname(e[1])
        if k == "path":
            return f"_rt.path({e[1]!r}, globals())"
        if k == "paren":
            return f"({self.ex(e[1])})"
        if k == "tuple":
            return "(" + ", ".join(self.ex(a) for a in e[1]) + ("," if len(e[1]) == 1 else "") + ")"
        if k == "array":
            return "[" + ", ".join(self.ex(a) for a in e[1]) + "]"
        if k == "repeat":
            return f"[{self.ex(e[1])}] * {self.ex(e[2])}"
        if k == "neg":
            return f"(-{self.ex(e[1])})"
        if k == "not":
            return f"(not {self.ex(e[1])})"
        if k == "as":
            return f"_rt.cast({self.ex(e[1])}, {e[2]!r})"
        if k == "field":
            return f"{self.ex(e[1])}.{pyname(e[2])}"
        if k == "index":
            if e[2][0] == "range":
                return f"{self.ex(e[1])}[{self.ex(e[2][1])}:{self.ex(e[2][2])}]"
            return f"{self.ex(e[1])}[{self.ex(e[2])}]"
        if k == "range":
            return f"range({self.ex(e[1])}, {self.ex(e[2])})"
        if k == "closure":
            return "(lambda " + ", ".join(self.pat(p) for p in e[1]) + ": " + self.ex(e[2]) + ")"
        if k == "macro":
            return f"_rt.macro({e[1]!r}" + "".join(", " + self.ex(a) for a in e[2]) + ")"
        if k == "method":
            return f"_rt.method({self.ex(e[1])}, {e[2]!r}" + "".join(", " + self.ex(a) for a in e[3]) + ")"
        if k == "call":
            args = ", ".join(self.ex(a) for a in e[2])
            if e[1][0] == "name":
                return f"{pyname(e[1][1])}({args})"
            return f"_rt.path({e[1][1]!r}, globals())({args})"
        if k == "struct":
            return f"_rt.struct({e[1]!r}" + "".join(f", {pyname(f)}={self.ex(v)}" for f, v in e[2]) + ")"
        if k == "bin":
            op, a, b = e[1], self.ex(e[2]), self.ex(e[3])
            if op == "&&":
                return f"({a} and {b})"
            if op == "||":
                return f"({a} or {b})"
            if op == "/":
                return f"_rt.div({a}, {b})"
            if op == "%":
                return f"_rt.rem({a}, {b})"
            return f"({a} {op} {b})"
        if k == "if":
            if not self.is_simple_if(e):
                raise NotImplementedError("if-expression with statements in its blocks")
            return f"({self.ex(e[2][0][1])} if {self.ex(e[1])} else {self.ex(e[3][0][1])})"
        raise NotImplementedError(k)


def transpile_fn(src):
    e = Emitter()
    e.fn(Parser(tokenize(src)).function())
    return "\n".join(e.lines) + "\n"


def parse_enum(src):
    return Parser(tokenize(src)).enum()


def transpile_const(src):
    _, name, expr = Parser(tokenize(src)).const()
    return f"{name} = {Emitter().ex(expr)}\n"


def extract_let(text, name):
    """Initialiser expression text of the first `let [mut] name [: T] = <expr>;` (balanced brackets up to the `;`)."""
    m = re.search(r"\blet\s+(?:mut\s+)?" + re.escape(name) + r"\b[^=;]*=", text)
    if not m:
        raise KeyError(name)
    depth, j = 0, m.end()
    while j < len(text):
        c = text[j]
        if c in "({[":
            depth += 1
        elif c in ")}]":
            depth -= 1
        elif c == ";" and depth == 0:
            return text[m.end():j]
        j += 1
    raise SyntaxError("unterminated let " + name)


def transpile_expr(src):
    p = Parser(tokenize(src))
    e = p.expr()
    return Emitter().ex(e)
