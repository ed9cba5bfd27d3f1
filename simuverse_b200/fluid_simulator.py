"""Host-side mirror of the reference's ``FluidSimulator`` plugin
(simuverse/src/fluid/fluid_simulator.rs:14-249), the ``impl Simulator`` the app drives
(simuverse/src/lib.rs:72-106).  Same entry points: ``on_click``, ``touch_begin``, ``touch_move``,
``update_uniforms``, ``reset``, ``compute``.  Rendering entry points (``draw_by_rpass``) are out of
scope: the canvas / particle / field buffers are read back instead.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib
from .d2q9_node import D2Q9Node, SettingObj, _f, lbm_uniform_new
from .wire import LID_DRIVEN_CAVITY, PARTICLE_DTYPE, ptr

OBSTACLE_RADIUS = 28  # fluid/mod.rs:1


def particle_grid(canvas_size, count):
    """Particle grid extent of ``get_particles_data`` (lib.rs:247-264)."""
    a, b = C.c_int32(), C.c_int32()
    lib.lbm_particle_grid(canvas_size[0], canvas_size[1], count, C.byref(a), C.byref(b))
    return a.value, b.value


def init_trajectory_particles(canvas_size, num, life_time, seed):
    """``init_trajectory_particles`` (lib.rs:275-316) on a seeded stream (the reference draws from
    an unseeded ``rand::rng()``, so its particle positions are not reproducible)."""
    out = np.zeros(num[0] * num[1], dtype=PARTICLE_DTYPE)
    lib.lbm_init_trajectory_particles(canvas_size[0], canvas_size[1], num[0], num[1], _f(life_time), seed, ptr(out))
    return out


class FluidSimulator:
    def __init__(self, canvas_size, setting=None, scale_factor=1.0, *, particles=True, particle_seed=0x5EED,
                 **node_kwargs):
        """``FluidSimulator::new`` (fluid_simulator.rs:26-133).  ``particles=False`` skips the tracer
        buffers (and the per-step macro texture they read)."""
        self.setting = setting or SettingObj()
        flags = node_kwargs.pop("flags", 0)
        self.particles_num = (0, 0)
        max_particles = 0
        if particles:
            self.particles_num = particle_grid(canvas_size, self.setting.particles_count)
            max_particles = self.particles_num[0] * self.particles_num[1]
            flags |= _capi.FLAG_MACRO_EVERY_STEP
        self.fluid_compute_node = D2Q9Node(canvas_size, self.setting, scale_factor, flags=flags,
                                           max_particles=max_particles, **node_kwargs)
        self.lattice = self.fluid_compute_node.lattice
        self.lattice_pixel_size = self.fluid_compute_node.lattice_pixel_size
        self.pre_pos = (0.0, 0.0)
        if particles:
            pu = self.setting.particles_uniform_data
            pu.num[:] = list(self.particles_num)
            self.fluid_compute_node.write_particle_uniform(pu)
            data = init_trajectory_particles(canvas_size, self.particles_num, pu.life_time, particle_seed)
            self.fluid_compute_node.write_particles(data)
        self._particles = particles

    # ------------------------------------------------------------------ impl Simulator
    def on_click(self, pos):
        """fluid_simulator.rs:137-152"""
        x, y = C.c_uint32(), C.c_uint32()
        nx, ny = self.lattice
        if not lib.lbm_on_click_guard(nx, ny, self.lattice_pixel_size, _f(pos[0]), _f(pos[1]), C.byref(x), C.byref(y)):
            return False
        self.fluid_compute_node.add_obstacle(x.value, y.value)
        return True

    def touch_begin(self):
        """fluid_simulator.rs:154-156"""
        self.pre_pos = (0.0, 0.0)

    def touch_move(self, pos):
        """fluid_simulator.rs:158-173"""
        if pos[0] <= 0.0 or pos[1] <= 0.0:
            self.pre_pos = (0.0, 0.0)
            return 0
        dx = np.float32(pos[0]) - np.float32(self.pre_pos[0])
        dy = np.float32(pos[1]) - np.float32(self.pre_pos[1])
        dis = np.sqrt(dx * dx + dy * dy, dtype=np.float32)
        if (self.pre_pos[0] == 0.0 and self.pre_pos[1] == 0.0) or dis > 300.0:
            self.pre_pos = (float(pos[0]), float(pos[1]))
            return 0
        n = self.fluid_compute_node.add_external_force(pos, self.pre_pos)
        self.pre_pos = (float(pos[0]), float(pos[1]))
        return n

    def update_uniforms(self, setting):
        """fluid_simulator.rs:175-193: tau = 3*viscosity + 0.5, uniform re-uploaded."""
        tau = lib.lbm_tau_from_viscosity(_f(setting.fluid_viscosity))
        fluid_ty = 1 if setting.animation_type == LID_DRIVEN_CAVITY else 0
        nx, ny = self.lattice
        self.fluid_compute_node.write_uniform(lbm_uniform_new(tau, fluid_ty, (nx * ny) & 0x7FFFFFFF))

    def reset(self):
        """fluid_simulator.rs:210-215"""
        self.fluid_compute_node.reset_lattice_info()
        self.pre_pos = (0.0, 0.0)

    def draw_by_rpass(self):
        """fluid_simulator.rs:234-248 draws the canvas with present.wgsl; the part of that pass that changes
        state — the in-place alpha fade of the canvas (present.wgsl:43-49) — is all that is reproduced."""
        if self._particles:
            self.fluid_compute_node.canvas_fade()

    def compute(self, n_frames=1):
        """fluid_simulator.rs:217-232: one frame = step(0), particles, step(1), particles."""
        self.fluid_compute_node.compute_frames(n_frames)
