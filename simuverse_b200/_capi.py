"""ctypes binding of ``include/lbm_b200.h`` (the C ABI a Rust ``-sys`` crate would bind).

The shared library is built in-tree by ``simuverse_b200/csrc/build.sh`` into
``simuverse_b200/_native/liblbm_b200.so``.  There is no fallback of any kind: if the library
is missing the import of this module raises, and every compute entry point fails with
``LbmError`` when no CUDA device is present.
"""
import ctypes as C
import os

from .wire import FieldUniform, LbmUniform, ParticleUniform

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBM_B200_LIB: tuning experiments load an alternative in-tree build of the SAME sources (never a fallback)
LIB_PATH = os.environ.get("LBM_B200_LIB") or os.path.join(_HERE, "_native", "liblbm_b200.so")

OK, ERR_INVALID_ARG, ERR_CUDA, ERR_NO_DEVICE, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED, ERR_STATE = range(7)
FLAG_MACRO_EVERY_STEP = 0x1
FLAG_KERNEL_GENERIC = 0x2
FLAG_NO_GRAPH = 0x4
FLAG_AA = 0x8
FLAG_NO_FUSE = 0x10
MACRO_F32_PLANES, MACRO_RGBA16F = 0, 1
PRESET_POROUS = 100


class LbmError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"lbm_b200 status {status}: {message}")
        self.status = status


class LbmDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("lattice_pixel_size", C.c_int32),
        ("canvas_w", C.c_int32),
        ("canvas_h", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("flags", C.c_uint32),
        ("max_particles", C.c_int32),
    ]


class LbmIpcBlob(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 256)]


# name -> (restype, argtypes); one entry per function declared in include/lbm_b200.h
_vp, _i32, _u32, _u64, _f32 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
_H = C.c_void_p
PROTOTYPES = {
    "lbm_abi_version": (C.c_int, []),
    "lbm_device_count": (C.c_int, []),
    "lbm_create": (C.c_int, [C.POINTER(LbmDesc), C.POINTER(_H)]),
    "lbm_destroy": (None, [_H]),
    "lbm_last_error": (C.c_char_p, [_H]),
    "lbm_status_string": (C.c_char_p, [C.c_int]),
    "lbm_write_uniform": (C.c_int, [_H, C.POINTER(LbmUniform)]),
    "lbm_write_field_uniform": (C.c_int, [_H, C.POINTER(FieldUniform)]),
    "lbm_write_lattice_info": (C.c_int, [_H, _u64, _vp, _u64]),
    "lbm_generate_lattice_info": (C.c_int, [_H, _i32, _u64, _f32]),
    "lbm_reset": (C.c_int, [_H]),
    "lbm_step": (C.c_int, [_H, _i32]),
    "lbm_step_n": (C.c_int, [_H, _i32]),
    "lbm_compute_frames": (C.c_int, [_H, _i32]),
    "lbm_swap_index": (C.c_int, [_H]),
    "lbm_sync": (C.c_int, [_H]),
    "lbm_slab_rows": (C.c_int, [_H, C.POINTER(_i32), C.POINTER(_i32)]),
    "lbm_read_distributions": (C.c_int, [_H, _i32, _vp]),
    "lbm_write_distributions": (C.c_int, [_H, _i32, _vp]),
    "lbm_read_macro": (C.c_int, [_H, _i32, _vp]),
    "lbm_read_macro_async": (C.c_int, [_H, _vp]),
    "lbm_read_curl": (C.c_int, [_H, _vp]),
    "lbm_read_present": (C.c_int, [_H, _i32, _i32, _vp]),
    "lbm_read_lattice_info": (C.c_int, [_H, _vp]),
    "lbm_total_mass": (C.c_int, [_H, _i32, C.POINTER(C.c_double)]),
    "lbm_write_particle_uniform": (C.c_int, [_H, C.POINTER(ParticleUniform)]),
    "lbm_particles_write": (C.c_int, [_H, _vp, _u64]),
    "lbm_particles_update": (C.c_int, [_H]),
    "lbm_particles_read": (C.c_int, [_H, _vp, _u64]),
    "lbm_canvas_clear": (C.c_int, [_H]),
    "lbm_canvas_fade": (C.c_int, [_H]),
    "lbm_canvas_read": (C.c_int, [_H, _vp]),
    "lbm_ipc_export": (C.c_int, [_H, C.POINTER(LbmIpcBlob)]),
    "lbm_ipc_attach": (C.c_int, [_H, C.POINTER(LbmIpcBlob), C.POINTER(LbmIpcBlob)]),
    "lbm_refresh_previous": (C.c_int, [_H]),
    "lbm_launch_count": (_u64, [_H]),
    "lbm_fused_sweep_count": (_u64, [_H]),
    "lbm_sweep_uses_masked_path": (C.c_int, [_H]),
    "lbm_last_step_n_ms": (C.c_int, [_H, C.POINTER(_f32)]),
    "lbm_edge_wait_stats": (C.c_int, [_H, C.POINTER(_u64), C.POINTER(_u64)]),
    "lbm_stream": (_vp, [_H]),
    # host-side mirrors
    "lbm_sweep_blocks": (_i32, [_i32, _i32, _vp, _i32, C.POINTER(_i32)]),
    "lbm_sweep_blocks_tail": (_i32, [_i32, _i32, _i32, _i32, _vp, _i32, C.POINTER(_i32)]),
    "lbm_scan_lattice_info_write": (_i32, [_i32, _i32, _u64, _vp, _u64, C.POINTER(_i32), C.POINTER(_i32)]),
    "lbm_uniform_new": (None, [_f32, _i32, _i32, C.POINTER(LbmUniform)]),
    "lbm_tau_from_viscosity": (_f32, [_f32]),
    "lbm_field_uniform_new": (None, [_i32, _i32, _u32, _i32, _i32, C.POINTER(FieldUniform)]),
    "lbm_init_lattice_material": (C.c_int, [_i32, _i32, _i32, _vp]),
    "lbm_init_porous_material": (C.c_int, [_i32, _i32, _u64, _f32, _vp]),
    "lbm_on_click_guard": (C.c_int, [_i32, _i32, _u32, _f32, _f32, C.POINTER(_u32), C.POINTER(_u32)]),
    "lbm_obstacle_patch": (_u64, [_i32, _i32, _vp, _u32, _u32, _vp, C.POINTER(_u64)]),
    "lbm_external_force_cells": (_u64, [_i32, _i32, _u32, _f32, _f32, _f32, _f32, _vp, _vp, _u64]),
    "lbm_particle_grid": (None, [_u32, _u32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "lbm_init_trajectory_particles": (None, [_u32, _u32, _i32, _i32, _f32, _u64, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with simuverse_b200/csrc/build.sh "
            "(or __graft_entry__.build()). simuverse_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.lbm_abi_version() != 1:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.lbm_abi_version()} != 1")
    return lib


lib = _load()


def check(status, handle=None):
    if status != OK:
        msg = lib.lbm_last_error(handle)
        text = msg.decode() if msg else ""
        raise LbmError(status, f"{lib.lbm_status_string(status).decode()}: {text}")
