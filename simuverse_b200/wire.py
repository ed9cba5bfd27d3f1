"""Wire formats at the drop-in boundary (mirror of ``include/lbm_wire.h``).

Each ctypes structure / numpy dtype is the byte image of the reference's
``#[repr(C)]`` Pod struct of the same name:

* ``LbmUniform``          simuverse/src/fluid/mod.rs:12-29
* ``FieldUniform``        simuverse/src/lib.rs:161-179
* ``LatticeInfo``         simuverse/src/fluid/lattice.rs:5-13
* ``ParticleUniform``     simuverse/src/lib.rs:180-196
* ``TrajectoryParticle``  simuverse/src/lib.rs:198-205
* ``Pixel``               simuverse/src/lib.rs:235-243
"""
import ctypes as C

import numpy as np

Q = 9

# fluid/lattice.rs:15-24
BULK, BOUNDARY, INLET, OBSTACLE, OUTLET, EXTERNAL_FORCE, GHOST = 1, 2, 3, 4, 5, 6, 7
# lib.rs:119-141 (the three presets the LBM node understands)
POISEUILLE, LID_DRIVEN_CAVITY, CUSTOM = 4, 5, 6


class LbmUniform(C.Structure):
    _fields_ = [
        ("tau", C.c_float),
        ("omega", C.c_float),
        ("fluid_ty", C.c_int32),
        ("soa_offset", C.c_int32),
        ("e_w_max", (C.c_float * 4) * Q),
        ("inversed_direction", (C.c_int32 * 4) * Q),
    ]


class FieldUniform(C.Structure):
    _fields_ = [
        ("lattice_size", C.c_int32 * 2),
        ("lattice_pixel_size", C.c_float * 2),
        ("canvas_size", C.c_int32 * 2),
        ("proj_ratio", C.c_float * 2),
        ("ndc_pixel", C.c_float * 2),
        ("speed_ty", C.c_int32),
        ("_padding", C.c_float),
    ]


class LatticeInfo(C.Structure):
    _fields_ = [
        ("material", C.c_int32),
        ("block_iter", C.c_int32),
        ("vx", C.c_float),
        ("vy", C.c_float),
    ]


class ParticleUniform(C.Structure):
    _fields_ = [
        ("color", C.c_float * 4),
        ("num", C.c_int32 * 2),
        ("point_size", C.c_int32),
        ("life_time", C.c_float),
        ("fade_out_factor", C.c_float),
        ("speed_factor", C.c_float),
        ("color_ty", C.c_int32),
        ("is_only_update_pos", C.c_int32),
    ]


class TrajectoryParticle(C.Structure):
    _fields_ = [
        ("pos", C.c_float * 2),
        ("pos_initial", C.c_float * 2),
        ("life_time", C.c_float),
        ("fade", C.c_float),
    ]


class Pixel(C.Structure):
    _fields_ = [("alpha", C.c_float), ("velocity_x", C.c_float), ("velocity_y", C.c_float)]


assert C.sizeof(LbmUniform) == 304
assert C.sizeof(FieldUniform) == 48
assert C.sizeof(LatticeInfo) == 16
assert C.sizeof(ParticleUniform) == 48
assert C.sizeof(TrajectoryParticle) == 24
assert C.sizeof(Pixel) == 12

# numpy views of the array-typed buffers
LATTICE_INFO_DTYPE = np.dtype(
    [("material", "<i4"), ("block_iter", "<i4"), ("vx", "<f4"), ("vy", "<f4")]
)
PARTICLE_DTYPE = np.dtype(
    [("pos", "<f4", 2), ("pos_initial", "<f4", 2), ("life_time", "<f4"), ("fade", "<f4")]
)
PIXEL_DTYPE = np.dtype([("alpha", "<f4"), ("velocity_x", "<f4"), ("velocity_y", "<f4")])
assert LATTICE_INFO_DTYPE.itemsize == 16
assert PARTICLE_DTYPE.itemsize == 24
assert PIXEL_DTYPE.itemsize == 12


def ptr(a, ty=C.c_void_p):
    """ctypes pointer to a C-contiguous numpy array's data."""
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ty)
