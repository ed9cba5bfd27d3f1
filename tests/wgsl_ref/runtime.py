"""Runtime for the Python emitted by wgsl2py: IEEE f32 scalars (numpy.float32, one rounding per
operation), i32 as Python ints, WGSL value semantics for vectors / structs / arrays."""
import numpy as np

F = np.float32


def f32(v):
    return F(v)


def i32(v):
    """WGSL f32 -> i32 conversion: truncate toward zero (saturating)."""
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, Vec):
        return Vec([i32(c) for c in v.v])
    x = float(v)
    if x != x:
        return 0
    return int(max(min(x, 2147483647.0), -2147483648.0))


def u32(v):
    return i32(v) & 0xFFFFFFFF if not isinstance(v, Vec) else Vec([u32(c) for c in v.v])


_SWZ = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


class Vec:
    __slots__ = ("v",)
    __array_ufunc__ = None  # numpy scalars must defer to Vec.__rmul__ etc.

    def __init__(self, comps):
        object.__setattr__(self, "v", list(comps))

    def __getattr__(self, name):
        idx = [_SWZ[c] for c in name]
        return self.v[idx[0]] if len(idx) == 1 else Vec([self.v[i] for i in idx])

    def __setattr__(self, name, value):
        idx = [_SWZ[c] for c in name]
        if len(idx) == 1:
            self.v[idx[0]] = value
        else:
            for i, c in zip(idx, value.v):
                self.v[i] = c

    def __getitem__(self, i):
        return self.v[i]

    def __setitem__(self, i, val):
        self.v[i] = val

    def _bin(self, o, fn, swap=False):
        ov = o.v if isinstance(o, Vec) else [o] * len(self.v)
        return Vec([fn(b, a) if swap else fn(a, b) for a, b in zip(self.v, ov)])

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __neg__(self): return Vec([-a for a in self.v])
    def __eq__(self, o): return all(a == b for a, b in zip(self.v, o.v))
    def __repr__(self): return f"Vec({self.v})"


def vec(n, elem, args):
    """vecN<elem>(...) constructor: scalars and vectors are flattened, one scalar is splatted."""
    flat = []
    for a in args:
        flat.extend(a.v if isinstance(a, Vec) else [a])
    if len(flat) == 1:
        flat = flat * n
    assert len(flat) == n, (n, flat)
    conv = {"f32": f32, "i32": i32, "u32": u32}[elem]
    return Vec([conv(c) for c in flat])


def div(a, b):
    if isinstance(a, Vec) or isinstance(b, Vec):
        av = a.v if isinstance(a, Vec) else [a] * len(b.v)
        bv = b.v if isinstance(b, Vec) else [b] * len(av)
        return Vec([div(x, y) for x, y in zip(av, bv)])
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        q = abs(int(a)) // abs(int(b))  # WGSL integer division truncates toward zero
        return q if (a >= 0) == (b >= 0) else -q
    with np.errstate(all="ignore"):
        return F(a) / F(b)


def mod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(a) - int(b) * div(a, b)
    return F(np.fmod(F(a), F(b)))


def copy(v):
    if isinstance(v, Vec):
        return Vec(v.v)
    if isinstance(v, list):
        return [copy(x) for x in v]
    if hasattr(v, "_wgsl_copy"):
        return v._wgsl_copy()
    return v


def make_struct(name, fields, zeros):
    class S:
        __slots__ = tuple(fields)

        def __init__(self, *args, **kw):
            init = zeros()
            for f, a in zip(fields, args):
                init[f] = a
            init.update(kw)
            for f in fields:
                object.__setattr__(self, f, copy(init[f]))

        def _wgsl_copy(self):
            return S(*[getattr(self, f) for f in fields])

        def __repr__(self):
            return name + "(" + ", ".join(f"{f}={getattr(self, f)!r}" for f in fields) + ")"

    S.__name__ = name
    return S


# ---------------------------------------------------------------- builtins
def dot(a, b):
    s = a.v[0] * b.v[0]
    for x, y in zip(a.v[1:], b.v[1:]):
        s = s + x * y
    return s


def _map(fn, *xs):
    if any(isinstance(x, Vec) for x in xs):
        n = len(next(x for x in xs if isinstance(x, Vec)).v)
        cols = [x.v if isinstance(x, Vec) else [x] * n for x in xs]
        return Vec([fn(*c) for c in zip(*cols)])
    return fn(*xs)


def clamp(x, lo, hi):
    return _map(lambda a, b, c: min(max(a, b), c) if isinstance(a, (int, np.integer)) else F(min(max(a, F(b)), F(c))), x, lo, hi)


def floor(x): return _map(lambda a: F(np.floor(a)), x)
def ceil(x): return _map(lambda a: F(np.ceil(a)), x)
def abs(x): return _map(lambda a: a.__abs__() if isinstance(a, (int, np.integer)) else F(np.abs(a)), x)  # noqa: A001
def sqrt(x): return _map(lambda a: F(np.sqrt(a)), x)


def min(a, b):  # noqa: A001
    import builtins
    return _map(lambda x, y: builtins.min(x, y), a, b)


def max(a, b):  # noqa: A001
    import builtins
    return _map(lambda x, y: builtins.max(x, y), a, b)


def length(v): return F(np.sqrt(dot(v, v)))


def smoothstep(lo, hi, x):
    t = clamp(div(x - lo, hi - lo), F(0), F(1))
    return t * t * (F(3) - F(2) * t)


class StorageF32:
    """``struct StoreFloat { data: array<f32> }`` bound to a numpy float32 array."""

    def __init__(self, arr):
        self.data = arr


class StructArray:
    """``array<SomeStruct>`` over a numpy structured array; reads copy out, writes copy in."""

    def __init__(self, arr, cls, to_struct, from_struct):
        self.arr, self.cls, self.to_struct, self.from_struct = arr, cls, to_struct, from_struct

    def __getitem__(self, i):
        return self.to_struct(self.cls, self.arr[i])

    def __setitem__(self, i, s):
        self.from_struct(self.arr, i, s)


class Texture16F:
    """rgba16float texture (storage write and sampled read): array of shape (ny, nx, 4) float16."""

    def __init__(self, arr):
        self.arr = arr


def textureStore(tex, uv, value):
    x, y = int(uv.v[0]), int(uv.v[1])
    with np.errstate(over="ignore"):
        tex.arr[y, x, :] = np.array([F(c) for c in value.v], np.float32).astype(np.float16)


def textureLoad(tex, uv, level):
    x, y = int(uv.v[0]), int(uv.v[1])
    ny, nx = tex.arr.shape[:2]
    if not (0 <= x < nx and 0 <= y < ny):
        # WGSL leaves an out-of-bounds textureLoad to the implementation (zero, (0,0,0,1) or some in-bounds texel);
        # wgpu compiles shaders with naga's BoundsCheckPolicy::ReadZeroSkipWrite for image loads, i.e. zeros.  Only
        # lbm/curl_update.wgsl gets here (its `min(.., lattice_size)` clamp is one past the last texel).
        return Vec([F(0.0)] * 4)
    return Vec([F(c) for c in tex.arr[y, x, :].astype(np.float32)])


# ---------------------------------------------------------------- builtins of the render shaders (lbm/present.wgsl)
def fract(x): return _map(lambda a: F(a - F(np.floor(a))), x)   # WGSL: e - floor(e)


def mix(a, b, t):
    """WGSL linear blend, the spec's own expression e1 * (1 - e3) + e2 * e3, one f32 rounding per operation."""
    return _map(lambda x, y, w: F(F(x * F(F(1.0) - w)) + F(y * w)), a, b, t)


def atan2(y, x): return _map(lambda p, q: F(np.arctan2(F(p), F(q))), y, x)


class BilinearClampSampler:
    """util/load_texture.rs:229-243 `bilinear_sampler`: ClampToEdge, linear mag / min filter (one mip level)."""


def textureSample(tex, sampler, uv):
    """textureSample of a 2D rgba16float texture through `bilinear_sampler`.  WebGPU (like Vulkan / Metal / D3D) leaves
    the precision of the filter weights to the hardware (fixed point, >= 4 fractional bits); this restatement — and the
    oracle and the CUDA pass pinned on it — evaluate the spec's formula in f32, one rounding per operation:
        c = uv * size - 0.5;  i0 = floor(c);  f = c - i0;  taps clamped to the edge;
        (1-fx)(1-fy) t00 + fx(1-fy) t10 + (1-fx) fy t01 + fx fy t11, summed in that order."""
    assert isinstance(sampler, BilinearClampSampler)
    ny, nx = tex.arr.shape[:2]
    cx = F(F(F(uv.v[0]) * F(nx)) - F(0.5))
    cy = F(F(F(uv.v[1]) * F(ny)) - F(0.5))
    fx0, fy0 = F(np.floor(cx)), F(np.floor(cy))
    fx, fy = F(cx - fx0), F(cy - fy0)
    import builtins
    x0 = builtins.min(builtins.max(int(fx0), 0), nx - 1)
    x1 = builtins.min(builtins.max(int(fx0) + 1, 0), nx - 1)
    y0 = builtins.min(builtins.max(int(fy0), 0), ny - 1)
    y1 = builtins.min(builtins.max(int(fy0) + 1, 0), ny - 1)
    gx, gy = F(F(1.0) - fx), F(F(1.0) - fy)
    w00, w10, w01, w11 = F(gx * gy), F(fx * gy), F(gx * fy), F(fx * fy)

    def t(y, x):
        return [F(c) for c in tex.arr[y, x, :].astype(np.float32)]

    t00, t10, t01, t11 = t(y0, x0), t(y0, x1), t(y1, x0), t(y1, x1)
    out = []
    for k in range(4):
        s = F(w00 * t00[k])
        s = F(s + F(w10 * t10[k]))
        s = F(s + F(w01 * t01[k]))
        s = F(s + F(w11 * t11[k]))
        out.append(s)
    return Vec(out)
