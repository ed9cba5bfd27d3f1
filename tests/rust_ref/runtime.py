"""Runtime for the Python emitted by rust2py: f32 = numpy.float32 (one IEEE rounding per operation), integers are
Python ints (overflow is not modelled), structs are immutable records (the reference's are `Copy`).

Third-party pieces the reference's functions call, restated here (not in /root/reference; pinned in its Cargo.lock):
  * glam 0.32.1 `Vec2` (scalar math on every target for Vec2): `new`, `ZERO`, `+`, `-`, `dot` = x*x' + y*y',
    `length` = sqrt(dot(self, self)), `distance(b)` = (self - b).length(), `round` component-wise;
  * Rust std f32 `atan2` / `cos` / `sin` / `round` / `ceil` = the platform libm (`atan2f`, `cosf`, `sinf`, `roundf`,
    `ceilf` on linux-gnu), called here through ctypes; `sqrt` is the IEEE operation;
  * `as` casts: float -> int truncates toward zero and saturates (NaN -> 0); int -> narrower int wraps; int -> f32 rounds
    to nearest even.
"""
import ctypes
import ctypes.util

import numpy as np

F = np.float32

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n, _argc in (("atan2f", 2), ("cosf", 1), ("sinf", 1), ("tanf", 1), ("roundf", 1), ("ceilf", 1), ("floorf", 1)):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float] * _argc


def _m(name, *xs):
    return F(getattr(_libm, name)(*[float(F(x)) for x in xs]))


class Vec2:
    __slots__ = ("x", "y")
    __array_ufunc__ = None

    def __init__(self, x, y):
        object.__setattr__(self, "x", F(x))
        object.__setattr__(self, "y", F(y))

    def __setattr__(self, k, v):
        raise AttributeError("Vec2 is a value type here")

    def __sub__(self, o): return Vec2(self.x - o.x, self.y - o.y)
    def __add__(self, o): return Vec2(self.x + o.x, self.y + o.y)
    def dot(self, o): return F(F(self.x * o.x) + F(self.y * o.y))
    def length(self): return F(np.sqrt(self.dot(self)))
    def distance(self, o): return (self - o).length()
    def round(self): return Vec2(_m("roundf", self.x), _m("roundf", self.y))
    def __eq__(self, o): return isinstance(o, Vec2) and self.x == o.x and self.y == o.y
    def __repr__(self): return f"Vec2({self.x!r}, {self.y!r})"


class Variant:
    """An enum variant: compares by identity of (enum, name); `as i32` gives its discriminant."""
    __slots__ = ("enum", "name", "value")

    def __init__(self, enum, name, value=None):
        self.enum, self.name, self.value = enum, name, value

    def __eq__(self, o): return isinstance(o, Variant) and (self.enum, self.name) == (o.enum, o.name)
    def __hash__(self): return hash((self.enum, self.name))
    def __repr__(self): return f"{self.enum}::{self.name}"


class Record:
    """Struct value (all the reference's structs on this path are Copy: immutable here, so sharing is copying)."""

    def __init__(self, ty, **fields):
        object.__setattr__(self, "_ty", ty)
        object.__setattr__(self, "_fields", tuple(fields))
        for k, v in fields.items():
            object.__setattr__(self, k, v)

    def __setattr__(self, k, v):
        raise AttributeError("struct fields are not assigned on this path")

    def __repr__(self):
        return self._ty + "{" + ", ".join(f"{k}: {getattr(self, k)!r}" for k in self._fields) + "}"


def struct(ty, **fields):
    return Record(ty.split("::")[-1], **fields)


ENUMS = {}  # "LatticeType" -> {"Bulk": 1, ...}, filled by the harness from the parsed enum items
PATHS = {
    "glam::Vec2::new": Vec2,
    "glam::Vec2::ZERO": Vec2(0.0, 0.0),
    "bytemuck::cast_slice": lambda x: x,
    "bytemuck::bytes_of": lambda x: x,
}


def path(p, ns):
    if p in PATHS:
        return PATHS[p]
    segs = p.split("::")
    if len(segs) == 2 and (segs[0] + "__" + segs[1]) in ns:       # associated function transpiled by the harness
        return ns[segs[0] + "__" + segs[1]]
    if len(segs) == 2:                                           # enum variant
        return Variant(segs[0], segs[1], ENUMS.get(segs[0], {}).get(segs[1]))
    raise KeyError(p)


_INT_BITS = {"u8": 8, "u16": 16, "u32": 32, "u64": 64, "usize": 64, "i8": 8, "i16": 16, "i32": 32, "i64": 64, "isize": 64}


def cast(v, ty):
    if isinstance(v, Variant):
        v = v.value
    if ty == "f32":
        return F(v)
    if ty == "f64":
        return float(v)
    bits, signed = _INT_BITS[ty], ty[0] == "i"
    lo, hi = (-(1 << (bits - 1)), (1 << (bits - 1)) - 1) if signed else (0, (1 << bits) - 1)
    if isinstance(v, (float, np.floating)):
        x = float(v)
        if x != x:
            return 0
        return int(min(max(np.trunc(x), lo), hi))
    v = int(v) & ((1 << bits) - 1)
    return v - (1 << bits) if signed and v > hi else v


def div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q
    with np.errstate(all="ignore"):
        return F(a) / F(b)


def rem(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(a) - int(b) * div(a, b)
    return F(np.fmod(F(a), F(b)))


_F32_METHODS = {
    "atan2": lambda y, x: _m("atan2f", y, x),
    "cos": lambda a: _m("cosf", a),
    "sin": lambda a: _m("sinf", a),
    "tan": lambda a: _m("tanf", a),
    "round": lambda a: _m("roundf", a),
    "ceil": lambda a: _m("ceilf", a),
    "floor": lambda a: _m("floorf", a),
    "sqrt": lambda a: F(np.sqrt(F(a))),
    "abs": lambda a: F(np.abs(F(a))),
}


def macro(name, *args):
    if name == "panic":
        raise RuntimeError("panic!: " + " ".join(str(a) for a in args))
    if name == "assert_eq":
        assert args[0] == args[1], args
        return None
    raise NotImplementedError(name + "!")


def method(obj, name, *args):
    if name in ("unwrap_or_else", "unwrap", "expect", "clone"):  # Result / Option = their success value here
        return obj
    if isinstance(obj, range):
        if name == "contains":
            return args[0] in obj
    if isinstance(obj, (np.floating, float)):
        if name == "min":
            return obj if obj < args[0] else F(args[0])  # f32::min (no NaN on this path)
        if name == "max":
            return obj if obj > args[0] else F(args[0])
        return _F32_METHODS[name](obj, *args)
    if isinstance(obj, list):
        if name == "push":
            return obj.append(args[0])
        if name == "len":
            return len(obj)
    if isinstance(obj, (int, np.integer)) and name == "div_ceil":
        return -(-int(obj) // int(args[0]))
    return getattr(obj, name)(*args)
