"""Loads the reference's Rust host helpers on the LBM path from their source files, transpiles them (rust2py) and binds
them to stand-ins for the objects they touch (`self`, `queue`, `app`, `setting`), so they can be executed like
`D2Q9Node` / `FluidSimulator` methods.  What is executed is the reference's own text:

    fluid/mod.rs             const OBSTACLE_RADIUS, LbmUniform::new (:32-55), is_sd_sphere (:57-59)
    fluid/lattice.rs         enum LatticeType (:15-24), init_lattice_material (:26-98)
    fluid/d2q9_node.rs       add_obstacle (:215-245), add_external_force (:263-300)
    fluid/fluid_simulator.rs on_click (:137-152), touch_begin (:154-156), touch_move (:158-173),
                             update_uniforms (:175-193)
    util/matrix_helper.rs    fullscreen_factor (:19-41), and the FieldUniform literal of D2Q9Node::new (d2q9_node.rs:61-76)
    lib.rs                   get_particles_data (:247-273): the tracer grid extent and workgroup count (its call of
                             init_trajectory_particles, which draws from an unseeded rand::rng(), is stubbed out)
"""
import os
import struct
import types

import numpy as np

from . import runtime as rt
from . import rust2py

SRC = "/root/reference/simuverse/src/fluid"
F = np.float32

LATTICE_INFO_DTYPE = np.dtype([("material", "<i4"), ("block_iter", "<i4"), ("vx", "<f4"), ("vy", "<f4")])


def available():
    return os.path.isfile(os.path.join(SRC, "d2q9_node.rs"))


_ns = None


def namespace():
    """One Python namespace holding every transpiled item."""
    global _ns
    if _ns is not None:
        return _ns
    read = lambda f: open(os.path.join(SRC, f), encoding="utf-8").read()  # noqa: E731
    mod, lattice, node, sim = read("mod.rs"), read("lattice.rs"), read("d2q9_node.rs"), read("fluid_simulator.rs")
    code = [rust2py.transpile_const(rust2py.extract_item(mod, "const", "OBSTACLE_RADIUS"))]
    _, name, variants = rust2py.parse_enum(rust2py.extract_item(lattice, "enum", "LatticeType"))
    rt.ENUMS[name] = variants
    uniform_new = rust2py.transpile_fn(rust2py.extract_fn(mod, "new"))
    code.append(uniform_new.replace("def new(", "def LbmUniform__new(", 1))
    code.append(rust2py.transpile_fn(rust2py.extract_fn(mod, "is_sd_sphere")))
    code.append(rust2py.transpile_fn(rust2py.extract_fn(lattice, "init_lattice_material")))
    for fn in ("add_obstacle", "add_external_force"):
        code.append(rust2py.transpile_fn(rust2py.extract_fn(node, fn)))
    for fn in ("on_click", "touch_begin", "touch_move", "update_uniforms"):
        code.append(rust2py.transpile_fn(rust2py.extract_fn(sim, fn)))
    util = open(os.path.join(os.path.dirname(SRC), "util", "matrix_helper.rs"), encoding="utf-8").read()
    code.append(rust2py.transpile_fn(rust2py.extract_fn(util, "fullscreen_factor")))
    # the FieldUniform literal of D2Q9Node::new (d2q9_node.rs:65-76), as an expression of lattice / lattice_pixel_size /
    # canvas_size / sx / sy
    code.append("def field_uniform_literal(lattice, lattice_pixel_size, canvas_size, sx, sy):\n    return "
                + rust2py.transpile_expr(rust2py.extract_let(node, "field_uniform_data")) + "\n")
    lib_rs = open(os.path.join(os.path.dirname(SRC), "lib.rs"), encoding="utf-8").read()
    code.append(rust2py.transpile_const(rust2py.extract_item(lib_rs, "const", "MAX_PARTICLE_COUNT")))
    code.append(rust2py.transpile_fn(rust2py.extract_fn(lib_rs, "get_particles_data")))
    code.append("def init_trajectory_particles(canvas_size, num, life_time):\n    return []\n")
    rt.PATHS["TrajectoryParticle::zero"] = lambda: None
    ns = {"_rt": rt}
    src = "\n".join(code)
    exec(compile(src, "<reference Rust host helpers>", "exec"), ns)
    ns["__source__"] = src
    _ns = ns
    return ns


def animation(ty):
    """discriminant of lib.rs:119-127 (0 Basic .. 4 Poiseuille, 5 LidDrivenCavity, 6 Custom) -> FieldAnimationType variant"""
    names = ["Basic", "JuliaSet", "Spirl", "BlackHole", "Poiseuille", "LidDrivenCavity", "Custom"]
    return rt.Variant("FieldAnimationType", names[int(ty)], int(ty))


def info_to_array(records):
    out = np.zeros(len(records), LATTICE_INFO_DTYPE)
    for i, r in enumerate(records):
        out[i] = (r.material, r.block_iter, r.vx, r.vy)
    return out


def array_to_info(arr):
    return [rt.Record("LatticeInfo", material=int(a["material"]), block_iter=int(a["block_iter"]), vx=F(a["vx"]),
                      vy=F(a["vy"])) for a in arr]


def uniform_bytes(u):
    """#[repr(C)] LbmUniform (fluid/mod.rs:12-29): 304 bytes"""
    b = struct.pack("<ffii", float(u.tau), float(u.omega), int(u.fluid_ty), int(u.soa_offset))
    for row in u.e_w_max:
        b += struct.pack("<4f", *[float(F(c)) for c in row])
    for row in u.inversed_direction:
        b += struct.pack("<4i", *[int(c) for c in row])
    assert len(b) == 304
    return b


class Queue:
    """wgpu::Queue::write_buffer: records (buffer label, byte offset, bytes)."""

    def __init__(self):
        self.writes = []

    def write_buffer(self, buffer, offset, data):
        if isinstance(data, list):
            payload = info_to_array(data).tobytes()
        else:
            payload = uniform_bytes(data)
        self.writes.append((buffer, int(offset), payload))


class Node:
    """The fields of D2Q9Node (d2q9_node.rs:13-27) the executed methods touch."""

    def __init__(self, nx, ny, lattice_pixel_size, ty, info=None):
        ns = namespace()
        self.lattice = types.SimpleNamespace(width=nx, height=ny, depth_or_array_layers=1)
        self.lattice_pixel_size = lattice_pixel_size
        self.animation_ty = animation(ty)
        self.info_buf = types.SimpleNamespace(buffer="info_buf")
        self.lbm_uniform_buf = types.SimpleNamespace(buffer="lbm_uniform_buf")
        self.lattice_info_data = (ns["init_lattice_material"](self.lattice, self.animation_ty) if info is None
                                  else array_to_info(info))

    def add_obstacle(self, queue, x, y):
        return namespace()["add_obstacle"](self, queue, x, y)

    def add_external_force(self, queue, pos, pre_pos):
        return namespace()["add_external_force"](self, queue, pos, pre_pos)


class Simulator:
    """The fields of FluidSimulator (fluid_simulator.rs:14-28) the executed methods touch."""

    def __init__(self, nx, ny, lattice_pixel_size, ty, info=None):
        self.fluid_compute_node = Node(nx, ny, lattice_pixel_size, ty, info)
        self.lattice = self.fluid_compute_node.lattice
        self.lattice_pixel_size = lattice_pixel_size
        self.pre_pos = rt.Vec2(0.0, 0.0)
        self.app = types.SimpleNamespace(queue=Queue())

    @property
    def writes(self):
        return self.app.queue.writes

    def on_click(self, x, y):
        namespace()["on_click"](self, self.app, rt.Vec2(x, y))

    def touch_begin(self):
        namespace()["touch_begin"](self, self.app)

    def touch_move(self, x, y):
        namespace()["touch_move"](self, self.app, rt.Vec2(x, y))

    def update_uniforms(self, viscosity, ty):
        setting = types.SimpleNamespace(fluid_viscosity=F(viscosity), animation_type=animation(ty))
        namespace()["update_uniforms"](self, self.app, setting)


def init_lattice_material(nx, ny, ty):
    ns = namespace()
    lattice = types.SimpleNamespace(width=nx, height=ny, depth_or_array_layers=1)
    return info_to_array(ns["init_lattice_material"](lattice, animation(ty)))


def uniform_new(tau, fluid_ty, soa_offset):
    return uniform_bytes(namespace()["LbmUniform__new"](F(tau), int(fluid_ty), int(soa_offset)))


def particle_grid(canvas_w, canvas_h, count):
    """(width, height) of the tracer grid and the (16, 16) workgroup count, lib.rs:247-264; also checks the padding of
    the particle buffer to MAX_PARTICLE_COUNT records (:266-271)."""
    size, groups, particles = namespace()["get_particles_data"](types.SimpleNamespace(x=canvas_w, y=canvas_h), int(count),
                                                               F(60.0))
    assert len(particles) == namespace()["MAX_PARTICLE_COUNT"]
    return (size.width, size.height), tuple(groups)


# ---------------------------------------------------------------- the repo's own Rust shim, executed the same way
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "rust",
                    "simuverse-cuda-lbm", "src")
_shim_ns = None


def shim_namespace():
    """rust/simuverse-cuda-lbm/src/{cuda_d2q9_node,cuda_fluid_simulator}.rs cannot be compiled in this image either.  Their
    host-logic bodies are transpiled like the reference's and run on top of the reference's own items they import
    (`OBSTACLE_RADIUS`, `is_sd_sphere`, `LbmUniform::new`, `LatticeType`), against a stand-in for the FFI calls."""
    global _shim_ns
    if _shim_ns is not None:
        return _shim_ns
    ns = dict(namespace())
    read = lambda f: open(os.path.join(SHIM, f), encoding="utf-8").read()  # noqa: E731
    node, sim = read("cuda_d2q9_node.rs"), read("cuda_fluid_simulator.rs")
    code = ["def Ok(v=None):\n    return v\n"]
    for fn in ("add_obstacle", "add_external_force"):
        code.append(rust2py.transpile_fn(rust2py.extract_fn(node, fn)))
    for fn in ("on_click", "touch_begin", "touch_move", "update_uniforms"):
        code.append(rust2py.transpile_fn(rust2py.extract_fn(sim, fn)))
    src = "\n".join(code)
    exec(compile(src, "<rust/simuverse-cuda-lbm host logic>", "exec"), ns)
    ns["__source__"] = src
    _shim_ns = ns
    return ns


class ShimNode:
    """CudaD2Q9Node with the FFI behind `upload_info` / `write_uniform` replaced by recorders."""

    def __init__(self, nx, ny, lattice_pixel_size, ty, info):
        self.lattice = types.SimpleNamespace(width=nx, height=ny, depth_or_array_layers=1)
        self.lattice_pixel_size = lattice_pixel_size
        self.animation_ty = animation(ty)
        self.lattice_info_data = array_to_info(info)
        self.writes = []

    def upload_info(self, byte_offset, cells):
        self.writes.append(("info_buf", int(byte_offset), info_to_array(list(cells)).tobytes()))

    def write_uniform(self, uniform):
        self.writes.append(("lbm_uniform_buf", 0, uniform_bytes(uniform)))

    def add_obstacle(self, x, y):
        return shim_namespace()["add_obstacle"](self, x, y)

    def add_external_force(self, pos, pre_pos):
        return shim_namespace()["add_external_force"](self, pos, pre_pos)


class ShimSimulator:
    """CudaFluidSimulator: the fields its `impl Simulator` host logic touches."""

    def __init__(self, nx, ny, lattice_pixel_size, ty, info):
        self.fluid_compute_node = ShimNode(nx, ny, lattice_pixel_size, ty, info)
        self.lattice = self.fluid_compute_node.lattice
        self.lattice_pixel_size = lattice_pixel_size
        self.pre_pos = rt.Vec2(0.0, 0.0)

    @property
    def writes(self):
        return self.fluid_compute_node.writes

    def on_click(self, x, y):
        shim_namespace()["on_click"](self, None, rt.Vec2(x, y))

    def touch_begin(self):
        shim_namespace()["touch_begin"](self, None)

    def touch_move(self, x, y):
        shim_namespace()["touch_move"](self, None, rt.Vec2(x, y))

    def update_uniforms(self, viscosity, ty):
        setting = types.SimpleNamespace(fluid_viscosity=F(viscosity), animation_type=animation(ty))
        shim_namespace()["update_uniforms"](self, None, setting)


def field_uniform(canvas_w, canvas_h, lattice_pixel_size):
    """The 48 bytes of FieldUniform as D2Q9Node::new builds them (d2q9_node.rs:38-76): lattice = canvas / lattice_pixel_size,
    (sx, sy) from util::matrix_helper::fullscreen_factor(canvas, 75 deg)."""
    ns = namespace()
    canvas = types.SimpleNamespace(x=int(canvas_w), y=int(canvas_h))
    lattice = types.SimpleNamespace(width=canvas.x // lattice_pixel_size, height=canvas.y // lattice_pixel_size,
                                    depth_or_array_layers=1)
    pi = F(np.pi)  # core::f32::consts::PI
    fovy = F(F(F(75.0) / F(180.0)) * pi)
    _, sx, sy = ns["fullscreen_factor"](rt.Vec2(F(canvas.x), F(canvas.y)), fovy)
    u = ns["field_uniform_literal"](lattice, int(lattice_pixel_size), canvas, sx, sy)
    b = struct.pack("<2i", *[int(v) for v in u.lattice_size]) + struct.pack("<2f", *[float(F(v)) for v in u.lattice_pixel_size])
    b += struct.pack("<2i", *[int(v) for v in u.canvas_size]) + struct.pack("<2f", *[float(F(v)) for v in u.proj_ratio])
    b += struct.pack("<2f", *[float(F(v)) for v in u.ndc_pixel]) + struct.pack("<if", int(u.speed_ty), float(F(u._padding)))
    assert len(b) == 48
    return b
