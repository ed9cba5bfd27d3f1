"""Host-side mirror of the reference's ``D2Q9Node`` (simuverse/src/fluid/d2q9_node.rs:30-313).

Same entry points and argument meaning — ``new``, ``reset``, ``add_obstacle``,
``reset_lattice_info``, ``add_external_force``, ``compute_by_pass(swap_index)`` — but every wgpu
object the reference owns (two ping-pong storage buffers, info buffer, uniforms, RGBA16F macro
texture, pipelines) is replaced by one ``LbmSim`` handle of the C ABI (include/lbm_b200.h).
The ``wgpu::Queue`` / ``CommandEncoder`` / ``ComputePass`` arguments of the reference have no
counterpart: calls are stream-ordered on the handle.
"""
import ctypes as C
import math

import numpy as np

from . import _capi
from ._capi import LbmDesc, LbmIpcBlob, check, lib
from .wire import (
    LATTICE_INFO_DTYPE,
    LID_DRIVEN_CAVITY,
    PARTICLE_DTYPE,
    PIXEL_DTYPE,
    POISEUILLE,
    FieldUniform,
    LbmUniform,
    ParticleUniform,
    ptr,
)


def _f(v):
    return float(np.float32(v))


class SettingObj:
    """The slice of ``SettingObj`` (simuverse/src/setting/setting_obj.rs:4-56) the LBM player reads:
    preset, viscosity and the particle uniform with the app's defaults
    (control_panel.rs:29-31: 10000 particles, lifetime 90, point size 2)."""

    def __init__(self, animation_type=POISEUILLE, particles_count=10000, particle_lifetime=90.0, point_size=2,
                 color_ty=0):
        self.animation_type = animation_type
        self.fluid_viscosity = 0.02  # setting_obj.rs:32 (no slider in the app)
        self.particles_count = particles_count
        pu = ParticleUniform()
        pu.color[:] = [1.0] * 4
        pu.num[:] = [0, 0]
        pu.point_size = point_size
        pu.life_time = particle_lifetime
        pu.fade_out_factor = 0.96
        pu.speed_factor = 4.15  # setting_obj.rs:49-53, SimuType::Fluid
        pu.color_ty = color_ty
        pu.is_only_update_pos = 1
        self.particles_uniform_data = pu


def lbm_uniform_new(tau, fluid_ty, soa_offset):
    """``LbmUniform::new`` (fluid/mod.rs:31-55)."""
    u = LbmUniform()
    lib.lbm_uniform_new(_f(tau), fluid_ty, soa_offset, C.byref(u))
    return u


def init_lattice_material(nx, ny, ty):
    """``init_lattice_material`` (fluid/lattice.rs:26-98) -> structured array of nx*ny LatticeInfo."""
    out = np.zeros(nx * ny, dtype=LATTICE_INFO_DTYPE)
    rc = lib.lbm_init_lattice_material(nx, ny, ty, ptr(out))
    if rc:
        raise ValueError("lbm_init_lattice_material: bad arguments")
    return out


def init_porous_material(nx, ny, seed=0x5EED, solid_fraction=0.30):
    out = np.zeros(nx * ny, dtype=LATTICE_INFO_DTYPE)
    rc = lib.lbm_init_porous_material(nx, ny, seed, _f(solid_fraction), ptr(out))
    if rc:
        raise ValueError("lbm_init_porous_material: bad arguments")
    return out


class D2Q9Node:
    def __init__(self, canvas_size, setting, scale_factor=1.0, *, lattice=None, lattice_info=None,
                 device_preset=None, preset_seed=0x5EED, preset_solid_fraction=0.30, device=-1, rank=0, world=1,
                 flags=0, max_particles=0):
        """``D2Q9Node::new`` (d2q9_node.rs:31-209).

        canvas_size: (width, height) in physical pixels.  ``lattice=(nx, ny)`` overrides the
        reference's ``canvas / ceil(2*scale_factor)`` (d2q9_node.rs:38-43) for headless runs.
        ``lattice_info`` supplies a caller-made mask instead of ``init_lattice_material``;
        ``device_preset`` (a FieldAnimationType or PRESET_POROUS) generates the same mask on the
        device without a host array (no CPU mirror is kept until one is needed).
        """
        self.lattice_pixel_size = int(math.ceil(2.0 * scale_factor))
        if lattice is None:
            lattice = (canvas_size[0] // self.lattice_pixel_size, canvas_size[1] // self.lattice_pixel_size)
        self.lattice = (int(lattice[0]), int(lattice[1]))
        nx, ny = self.lattice
        self.canvas_size = (int(canvas_size[0]), int(canvas_size[1]))
        self.animation_ty = setting.animation_type
        self.workgroup_count = (-(-nx // 64), -(-ny // 4), 1)  # d2q9_node.rs:45 (informational)
        self.rank, self.world = rank, world

        desc = LbmDesc()
        desc.struct_size = C.sizeof(LbmDesc)
        desc.nx, desc.ny = nx, ny
        desc.lattice_pixel_size = self.lattice_pixel_size
        desc.canvas_w, desc.canvas_h = self.canvas_size
        desc.device = device
        desc.rank, desc.world = rank, world
        desc.flags = flags
        desc.max_particles = max_particles
        h = C.c_void_p()
        check(lib.lbm_create(C.byref(desc), C.byref(h)))
        self._h = h
        y0, rows = C.c_int32(), C.c_int32()
        check(lib.lbm_slab_rows(h, C.byref(y0), C.byref(rows)), h)
        self.y0, self.rows = y0.value, rows.value

        tau = lib.lbm_tau_from_viscosity(_f(setting.fluid_viscosity))  # d2q9_node.rs:50
        fluid_ty = 1 if setting.animation_type == LID_DRIVEN_CAVITY else 0  # d2q9_node.rs:53-57
        self.lbm_uniform_data = lbm_uniform_new(tau, fluid_ty, (nx * ny) & 0x7FFFFFFF)
        check(lib.lbm_write_uniform(h, C.byref(self.lbm_uniform_data)), h)
        self.field_uniform_data = FieldUniform()
        lib.lbm_field_uniform_new(nx, ny, self.lattice_pixel_size, self.canvas_size[0], self.canvas_size[1],
                                  C.byref(self.field_uniform_data))
        check(lib.lbm_write_field_uniform(h, C.byref(self.field_uniform_data)), h)

        self._device_preset = device_preset
        self._preset_args = (int(preset_seed), float(preset_solid_fraction))
        self.lattice_info_data = None
        if device_preset is not None:
            self._generate_on_device()
        else:
            if lattice_info is None:
                lattice_info = init_lattice_material(nx, ny, setting.animation_type)  # d2q9_node.rs:106
            self.lattice_info_data = np.array(lattice_info, dtype=LATTICE_INFO_DTYPE, copy=True).reshape(-1)
            assert self.lattice_info_data.size == nx * ny
            self.write_lattice_info(0, self.lattice_info_data)
        if world == 1:
            self.reset()  # d2q9_node.rs:206 reset_lattice_info -> init.wgsl; slabs reset after attach

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            lib.lbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference entry points
    def reset(self):
        """``D2Q9Node::reset`` (d2q9_node.rs:211-213): dispatch init.wgsl."""
        check(lib.lbm_reset(self._h), self._h)

    def add_obstacle(self, x, y):
        """``add_obstacle`` (d2q9_node.rs:215-245): disc R=28 at (x+.5, y+.5), 56 full rows re-uploaded."""
        nx, ny = self.lattice
        mirror = self._mirror()
        patch = np.zeros(56 * nx, dtype=LATTICE_INFO_DTYPE)
        off = C.c_uint64()
        n = lib.lbm_obstacle_patch(nx, ny, ptr(mirror), x, y, ptr(patch), C.byref(off))
        self.write_lattice_info(off.value, patch[:n])

    def reset_lattice_info(self):
        """``reset_lattice_info`` (d2q9_node.rs:247-261): Poiseuille re-uploads the preset mask, then init."""
        if self.animation_ty == POISEUILLE:
            if self._device_preset is not None:
                self._generate_on_device()
            else:
                nx, ny = self.lattice
                self.lattice_info_data = init_lattice_material(nx, ny, self.animation_ty)
                self.write_lattice_info(0, self.lattice_info_data)
        self.reset()

    def add_external_force(self, pos, pre_pos):
        """``add_external_force`` (d2q9_node.rs:263-300): one 16-byte write per sample point."""
        nx, ny = self.lattice
        cap = 4096
        offs = np.zeros(cap, np.uint64)
        cells = np.zeros(cap, dtype=LATTICE_INFO_DTYPE)
        n = int(lib.lbm_external_force_cells(nx, ny, self.lattice_pixel_size, _f(pos[0]), _f(pos[1]),
                                             _f(pre_pos[0]), _f(pre_pos[1]), ptr(offs), ptr(cells), cap))
        for k in range(min(n, cap)):
            # like the reference, the CPU mirror is NOT updated by force writes
            self.write_lattice_info(int(offs[k]), cells[k:k + 1])
        return n

    def compute_by_pass(self, swap_index):
        """``compute_by_pass`` (d2q9_node.rs:302-312): collide_stream + boundary with bind group
        ``swap_index`` — here one fused kernel reading buffer ``swap_index``."""
        check(lib.lbm_step(self._h, swap_index), self._h)

    # ------------------------------------------------------------------ buffer access
    def write_lattice_info(self, byte_offset, cells):
        """``queue.write_buffer(&info_buf, offset, bytes)`` (d2q9_node.rs:244,250-254,298)."""
        cells = np.ascontiguousarray(cells, dtype=LATTICE_INFO_DTYPE)
        check(lib.lbm_write_lattice_info(self._h, byte_offset, ptr(cells), cells.nbytes), self._h)

    def write_uniform(self, u):
        self.lbm_uniform_data = u
        check(lib.lbm_write_uniform(self._h, C.byref(u)), self._h)

    def step_n(self, n):
        check(lib.lbm_step_n(self._h, n), self._h)

    def compute_frames(self, n_frames):
        """``FluidSimulator::compute`` n times inside the library (fluid_simulator.rs:217-232)."""
        check(lib.lbm_compute_frames(self._h, n_frames), self._h)

    @property
    def swap_index(self):
        return lib.lbm_swap_index(self._h)

    def sync(self):
        check(lib.lbm_sync(self._h), self._h)

    def read_distributions(self, which):
        """Owned rows of ping-pong buffer ``which`` in the reference layout: (9, rows, nx) f32."""
        out = np.empty((9, self.rows, self.lattice[0]), np.float32)
        check(lib.lbm_read_distributions(self._h, which, ptr(out)), self._h)
        return out

    def write_distributions(self, which, data):
        data = np.ascontiguousarray(data, np.float32)
        assert data.size == 9 * self.rows * self.lattice[0]
        check(lib.lbm_write_distributions(self._h, which, ptr(data)), self._h)

    def read_macro(self):
        """(u.x, u.y, rho) of the last step as f32 planes, shape (3, rows, nx)."""
        out = np.empty((3, self.rows, self.lattice[0]), np.float32)
        check(lib.lbm_read_macro(self._h, _capi.MACRO_F32_PLANES, ptr(out)), self._h)
        return out

    def read_macro_tex(self):
        """The RGBA16F macro texture of the reference (d2q9_node.rs:91-104): (rows, nx, 4) float16."""
        out = np.empty((self.rows, self.lattice[0], 4), np.float16)
        check(lib.lbm_read_macro(self._h, _capi.MACRO_RGBA16F, ptr(out)), self._h)
        return out

    def read_macro_tex_async(self, out):
        """Enqueue the texture read-back into ``out`` ((rows, nx, 4) float16, ideally pinned); complete after sync()."""
        assert out.dtype == np.float16 and out.size == self.rows * self.lattice[0] * 4 and out.flags["C_CONTIGUOUS"]
        check(lib.lbm_read_macro_async(self._h, ptr(out)), self._h)

    def read_curl_tex(self):
        """``curl_tex`` of the reference's ``_curl_cal_node`` (fluid_simulator.rs:36-71, curl_update.wgsl:12-33) for
        the newest macro field: (rows, nx, 4) float16 texels (curl * 3.5 + 0.5, 0, 0, 0)."""
        out = np.empty((self.rows, self.lattice[0], 4), np.float16)
        check(lib.lbm_read_curl(self._h, ptr(out)), self._h)
        return out

    def read_present(self, row0=0, rows=None):
        """Fragment outputs of the reference's ``render_node`` (fluid_simulator.rs:69-87, lbm/present.wgsl:21-46) for
        canvas rows [row0, row0 + rows): (rows, canvas_w, 4) float32 (r, g, b, a = rho) from the newest macro field."""
        w, h = self.field_uniform_data.canvas_size[0], self.field_uniform_data.canvas_size[1]
        rows = h - row0 if rows is None else rows
        out = np.empty((rows, w, 4), np.float32)
        check(lib.lbm_read_present(self._h, row0, rows, ptr(out)), self._h)
        return out

    def read_lattice_info(self):
        out = np.empty(self.rows * self.lattice[0], dtype=LATTICE_INFO_DTYPE)
        check(lib.lbm_read_lattice_info(self._h, ptr(out)), self._h)
        return out

    def total_mass(self, which=None):
        if which is None:
            which = self.swap_index
        m = C.c_double()
        check(lib.lbm_total_mass(self._h, which, C.byref(m)), self._h)
        return m.value

    def last_step_n_ms(self):
        ms = C.c_float()
        check(lib.lbm_last_step_n_ms(self._h, C.byref(ms)), self._h)
        return ms.value

    def edge_wait_stats(self):
        """(nanoseconds summed over edge CTAs, number of waits) spent waiting for the neighbour slabs so far."""
        ns, n = C.c_uint64(), C.c_uint64()
        check(lib.lbm_edge_wait_stats(self._h, C.byref(ns), C.byref(n)), self._h)
        return ns.value, n.value

    def refresh_previous(self):
        """Recompute the buffer a two-update sweep left two updates behind (collective on multi-slab lattices)."""
        check(lib.lbm_refresh_previous(self._h), self._h)

    @property
    def launch_count(self):
        return int(lib.lbm_launch_count(self._h))

    @property
    def fused_sweep_count(self):
        """Launches that advanced the lattice by two updates at once (csrc/lbm_fused.cuh)."""
        return int(lib.lbm_fused_sweep_count(self._h))

    @property
    def sweep_uses_masked_path(self):
        """Whether sweeps run the kernel instance with the inline masked path (lattices with many solid cells)."""
        return bool(lib.lbm_sweep_uses_masked_path(self._h))

    # ------------------------------------------------------------------ particles
    def write_particle_uniform(self, pu):
        check(lib.lbm_write_particle_uniform(self._h, C.byref(pu)), self._h)

    def write_particles(self, particles):
        particles = np.ascontiguousarray(particles, dtype=PARTICLE_DTYPE)
        check(lib.lbm_particles_write(self._h, ptr(particles), particles.size), self._h)

    def particles_update(self):
        check(lib.lbm_particles_update(self._h), self._h)

    def read_particles(self, count):
        out = np.empty(count, dtype=PARTICLE_DTYPE)
        check(lib.lbm_particles_read(self._h, ptr(out), count), self._h)
        return out

    def canvas_clear(self):
        check(lib.lbm_canvas_clear(self._h), self._h)

    def canvas_fade(self):
        check(lib.lbm_canvas_fade(self._h), self._h)

    def read_canvas(self):
        out = np.empty(self.canvas_size[0] * self.canvas_size[1], dtype=PIXEL_DTYPE)
        check(lib.lbm_canvas_read(self._h, ptr(out)), self._h)
        return out.reshape(self.canvas_size[1], self.canvas_size[0])

    # ------------------------------------------------------------------ multi-slab wiring
    def ipc_export(self):
        blob = LbmIpcBlob()
        check(lib.lbm_ipc_export(self._h, C.byref(blob)), self._h)
        return bytes(blob.bytes)

    def ipc_attach(self, up_bytes, down_bytes):
        up, dn = LbmIpcBlob(), LbmIpcBlob()
        C.memmove(up.bytes, up_bytes, 256)
        C.memmove(dn.bytes, down_bytes, 256)
        check(lib.lbm_ipc_attach(self._h, C.byref(up), C.byref(dn)), self._h)

    # ------------------------------------------------------------------ internals
    def _generate_on_device(self):
        seed, frac = self._preset_args
        check(lib.lbm_generate_lattice_info(self._h, self._device_preset, seed, _f(frac)), self._h)

    def _mirror(self):
        if self.lattice_info_data is None:
            if self.world != 1:
                raise RuntimeError("no CPU mirror of the lattice info on a slab created with device_preset")
            self.lattice_info_data = self.read_lattice_info()
        return self.lattice_info_data
