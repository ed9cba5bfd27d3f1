"""ctypes front-end for the CPU oracle (``oracle/liblbm_oracle.so``).

Test infrastructure: imported only from tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wire():
    """The wire structs (byte images of the reference's Pod structs).  Inside a process that already uses the
    product package they are its ``simuverse_b200.wire`` classes (so structs can be handed back and forth);
    otherwise — the CPU-only reference arm of bench.py — wire.py is loaded on its own, WITHOUT importing the
    package, so that the product's shared library is never mapped into a process that only runs the oracle."""
    m = sys.modules.get("simuverse_b200.wire")
    if m is None:
        spec = importlib.util.spec_from_file_location("_lbm_wire_standalone",
                                                      os.path.join(_ROOT, "simuverse_b200", "wire.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    return m


_w = _wire()
LATTICE_INFO_DTYPE, PARTICLE_DTYPE, PIXEL_DTYPE = _w.LATTICE_INFO_DTYPE, _w.PARTICLE_DTYPE, _w.PIXEL_DTYPE
FieldUniform, LatticeInfo, LbmUniform, ParticleUniform, ptr = (_w.FieldUniform, _w.LatticeInfo, _w.LbmUniform,
                                                               _w.ParticleUniform, _w.ptr)

_ORACLE_DIR = os.path.join(_ROOT, "oracle")
_SO = os.path.join(_ORACLE_DIR, "liblbm_oracle.so")


def build_oracle(force=False):
    src = [os.path.join(_ORACLE_DIR, n) for n in ("lbm_oracle.c", "lbm_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None


def _f(v):
    """Python float carrying exactly the f32 value of v (ctypes c_float argument)."""
    return float(np.float32(v))


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        vp, i32, u32, u64, f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
        L.orc_f32_to_f16.restype = C.c_uint16
        L.orc_f32_to_f16.argtypes = [f32]
        L.orc_f16_to_f32.restype = f32
        L.orc_f16_to_f32.argtypes = [C.c_uint16]
        L.orc_set_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_get_max_threads.restype = C.c_int
        L.orc_lbm_uniform_new.argtypes = [f32, i32, i32, C.POINTER(LbmUniform)]
        L.orc_tau_from_viscosity.restype = f32
        L.orc_tau_from_viscosity.argtypes = [f32]
        L.orc_init_lattice_material.argtypes = [i32, i32, i32, vp]
        L.orc_init_porous_material.argtypes = [i32, i32, u64, f32, vp]
        L.orc_init.argtypes = [C.POINTER(LbmUniform), i32, i32, vp, vp, vp, vp]
        L.orc_collide_stream.argtypes = [C.POINTER(LbmUniform), i32, i32, vp, vp, vp, vp, vp]
        L.orc_boundary.argtypes = [C.POINTER(LbmUniform), i32, i32, vp, vp]
        L.orc_step.argtypes = [C.POINTER(LbmUniform), i32, i32, vp, vp, vp, vp, vp]
        L.orc_step_n.restype = C.c_int
        L.orc_step_n.argtypes = [C.POINTER(LbmUniform), i32, i32, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        L.orc_particle_update.argtypes = [
            C.POINTER(LbmUniform), C.POINTER(FieldUniform), C.POINTER(ParticleUniform), vp, vp, vp]
        L.orc_canvas_fade.argtypes = [C.POINTER(FieldUniform), C.POINTER(ParticleUniform), vp]
        L.orc_add_obstacle.restype = C.c_size_t
        L.orc_add_obstacle.argtypes = [i32, i32, vp, u32, u32, vp, C.POINTER(u64)]
        L.orc_on_click_guard.restype = C.c_int
        L.orc_on_click_guard.argtypes = [i32, i32, u32, f32, f32, C.POINTER(u32), C.POINTER(u32)]
        L.orc_add_external_force.restype = C.c_size_t
        L.orc_add_external_force.argtypes = [i32, i32, u32, f32, f32, f32, f32, vp, vp, C.c_size_t]
        L.orc_field_uniform_new.argtypes = [i32, i32, u32, i32, i32, C.POINTER(FieldUniform)]
        L.orc_particle_grid.argtypes = [u32, u32, i32, C.POINTER(i32), C.POINTER(i32)]
        L.orc_init_trajectory_particles.argtypes = [u32, u32, i32, i32, f32, u64, vp]
        L.orc_curl_update.argtypes = [i32, i32, vp, vp]
        L.orc_present.restype = None
        L.orc_present.argtypes = [vp, vp, vp, i32, i32, vp]
        L.orc_total_mass.restype = C.c_double
        L.orc_total_mass.argtypes = [i32, i32, vp]
        _lib = L
    return _lib


def uniform_new(tau, fluid_ty, soa_offset):
    u = LbmUniform()
    lib().orc_lbm_uniform_new(_f(tau), fluid_ty, soa_offset, C.byref(u))
    return u


def tau_from_viscosity(v):
    return float(lib().orc_tau_from_viscosity(_f(v)))


def init_lattice_material(nx, ny, ty):
    out = np.zeros(nx * ny, dtype=LATTICE_INFO_DTYPE)
    lib().orc_init_lattice_material(nx, ny, ty, ptr(out))
    return out


def init_porous_material(nx, ny, seed=0x5EED, solid_fraction=0.30):
    out = np.zeros(nx * ny, dtype=LATTICE_INFO_DTYPE)
    lib().orc_init_porous_material(nx, ny, seed, _f(solid_fraction), ptr(out))
    return out


class OracleSim:
    """The reference's D2Q9Node state on the host: two ping-pong distribution buffers in the
    reference layout ``buf[dir*N + y*nx + x]``, the LatticeInfo buffer, the RGBA16F macro
    texture (plus its f32 pre-quantisation image)."""

    def __init__(self, nx, ny, info, uniform, threads=1):
        self.nx, self.ny = nx, ny
        self.N = nx * ny
        self.u = uniform
        self.info = np.array(info, dtype=LATTICE_INFO_DTYPE, copy=True).reshape(-1)
        assert self.info.size == self.N
        self.buf = [np.zeros(9 * self.N, np.float32), np.zeros(9 * self.N, np.float32)]
        self.macro_f16 = np.zeros(4 * self.N, np.uint16)
        self.macro_f32 = np.zeros(4 * self.N, np.float32)
        self.swap = 0
        self.threads = threads
        self.reset()

    def reset(self):
        lib().orc_init(C.byref(self.u), self.nx, self.ny, ptr(self.buf[0]), ptr(self.buf[1]),
                       ptr(self.info), ptr(self.macro_f16))
        self.macro_f32[:] = 0
        self.macro_f32[3::4] = 1.0
        self.swap = 0

    def step(self, n=1):
        lib().orc_set_num_threads(self.threads)
        self.swap = lib().orc_step_n(C.byref(self.u), self.nx, self.ny, ptr(self.buf[0]), ptr(self.buf[1]),
                                     ptr(self.info), ptr(self.macro_f16), ptr(self.macro_f32), self.swap, n)

    @property
    def current(self):
        """Buffer holding the latest post-step state (the one the next step reads)."""
        return self.buf[self.swap]

    def distributions(self, which):
        return self.buf[which].reshape(9, self.ny, self.nx)

    def macro(self):
        """(ux, uy, rho) f32 planes before the f16 store, shape (3, ny, nx)."""
        m = self.macro_f32.reshape(self.ny, self.nx, 4)
        return np.ascontiguousarray(np.moveaxis(m[..., :3], 2, 0))

    def total_mass(self):
        return float(lib().orc_total_mass(self.nx, self.ny, ptr(self.current)))

    def write_lattice_info(self, byte_offset, cells):
        cells = np.ascontiguousarray(cells, dtype=LATTICE_INFO_DTYPE).reshape(-1)
        assert byte_offset % 16 == 0
        i0 = byte_offset // 16
        self.info[i0:i0 + cells.size] = cells

    def particle_update(self, field, pu, particles, canvas):
        lib().orc_particle_update(C.byref(self.u), C.byref(field), C.byref(pu), ptr(particles),
                                  ptr(canvas) if canvas is not None else None, ptr(self.macro_f16))


def curl_update(nx, ny, macro_f16):
    """curl_update.wgsl over an RGBA16F macro texture given as uint16 bits; returns (ny, nx, 4) uint16."""
    macro_f16 = np.ascontiguousarray(macro_f16, np.uint16).reshape(-1)
    assert macro_f16.size == 4 * nx * ny
    out = np.zeros(4 * nx * ny, np.uint16)
    lib().orc_curl_update(nx, ny, ptr(macro_f16), ptr(out))
    return out.reshape(ny, nx, 4)


def present(field, macro_f16, curl_f16, row0=0, rows=None):
    """lbm/present.wgsl (colour present of the field) for canvas rows [row0, row0 + rows): (rows, W, 4) float32."""
    nx, ny = field.lattice_size[0], field.lattice_size[1]
    W, H = field.canvas_size[0], field.canvas_size[1]
    rows = H - row0 if rows is None else rows
    macro_f16 = np.ascontiguousarray(macro_f16, np.uint16).reshape(-1)
    curl_f16 = np.ascontiguousarray(curl_f16, np.uint16).reshape(-1)
    assert macro_f16.size == 4 * nx * ny and curl_f16.size == 4 * nx * ny and 0 <= row0 and row0 + rows <= H
    out = np.zeros((rows, W, 4), np.float32)
    lib().orc_present(C.byref(field), ptr(macro_f16), ptr(curl_f16), row0, rows, ptr(out))
    return out


def canvas_fade(field, pu, canvas):
    lib().orc_canvas_fade(C.byref(field), C.byref(pu), ptr(canvas))


def add_obstacle(nx, ny, mirror, x, y):
    patch = np.zeros(56 * nx, dtype=LATTICE_INFO_DTYPE)
    off = C.c_uint64(0)
    n = lib().orc_add_obstacle(nx, ny, ptr(mirror), x, y, ptr(patch), C.byref(off))
    return int(off.value), patch[:n]


def on_click_guard(nx, ny, lps, px, py):
    x, y = C.c_uint32(0), C.c_uint32(0)
    ok = lib().orc_on_click_guard(nx, ny, lps, _f(px), _f(py), C.byref(x), C.byref(y))
    return (int(x.value), int(y.value)) if ok else None


def add_external_force(nx, ny, lps, pos, pre_pos, cap=4096):
    offs = np.zeros(cap, np.uint64)
    cells = np.zeros(cap, dtype=LATTICE_INFO_DTYPE)
    n = lib().orc_add_external_force(nx, ny, lps, _f(pos[0]), _f(pos[1]),
                                     _f(pre_pos[0]), _f(pre_pos[1]), ptr(offs), ptr(cells), cap)
    assert n <= cap
    return offs[:n].copy(), cells[:n].copy()


def field_uniform_new(nx, ny, lps, canvas_w, canvas_h):
    f = FieldUniform()
    lib().orc_field_uniform_new(nx, ny, lps, canvas_w, canvas_h, C.byref(f))
    return f


def particle_grid(canvas_w, canvas_h, count):
    a, b = C.c_int32(0), C.c_int32(0)
    lib().orc_particle_grid(canvas_w, canvas_h, count, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


def init_trajectory_particles(canvas_w, canvas_h, num_x, num_y, life_time, seed):
    out = np.zeros(num_x * num_y, dtype=PARTICLE_DTYPE)
    lib().orc_init_trajectory_particles(canvas_w, canvas_h, num_x, num_y, _f(life_time), seed, ptr(out))
    return out


def f32_to_f16_bits(a):
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    f = lib().orc_f32_to_f16
    return np.array([f(float(v)) for v in a], np.uint16)


__all__ = [n for n in dir() if not n.startswith("_")]
_ = (LatticeInfo, PIXEL_DTYPE)
