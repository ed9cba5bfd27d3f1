"""Binds numpy state to the transpiled reference shaders and dispatches them like
D2Q9Node::compute_by_pass / FluidSimulator::compute do (d2q9_node.rs:302-312, fluid_simulator.rs:217-232)."""
import os

import numpy as np

from . import runtime as rt
from . import wgsl2py

WGSL_ROOT = "/root/reference/assets/wgsl"
F = np.float32


def available():
    return os.path.isdir(os.path.join(WGSL_ROOT, "lbm"))


def _load(entry):
    ns = {"_rt": rt}
    code = wgsl2py.transpile(entry, WGSL_ROOT)
    exec(compile(code, entry, "exec"), ns)
    return ns


class WgslLbm:
    """The reference's LBM state driven by its own WGSL source."""

    def __init__(self, nx, ny, info, uniform, lattice_pixel_size=2, canvas=None):
        self.nx, self.ny, self.N = nx, ny, nx * ny
        self.info = np.array(info, copy=True).reshape(-1)
        self.buf = [np.zeros(9 * self.N, np.float32), np.zeros(9 * self.N, np.float32)]
        self.macro = np.zeros((ny, nx, 4), np.float16)
        self.swap = 0
        self.canvas_size = canvas or (nx * lattice_pixel_size, ny * lattice_pixel_size)
        self.mods = {k: _load(f"lbm/{k}.wgsl") for k in ("init", "collide_stream", "boundary", "particle_update")}
        self.u = uniform
        self.lps = lattice_pixel_size
        for ns in self.mods.values():
            self._bind_common(ns)
        self.dispatch("init", self.buf[0], self.buf[1])

    # ---- resource binding
    def _bind_common(self, ns):
        u = self.u
        if "LbmUniform" in ns:
            ns["fluid"] = ns["LbmUniform"](
                F(u.tau), F(u.omega), int(u.fluid_ty), int(u.soa_offset),
                [rt.Vec([F(c) for c in u.e_w_max[i]]) for i in range(9)],
                [rt.Vec([int(c) for c in u.inversed_direction[i]]) for i in range(9)])
        ns["field"] = ns["FieldUniform"](
            rt.Vec([self.nx, self.ny]), rt.Vec([F(self.lps), F(self.lps)]), rt.Vec(list(self.canvas_size)),
            rt.Vec([F(0), F(0)]), rt.Vec([F(0), F(0)]), 1)
        if "LatticeInfo" in ns:
            def to_s(cls, rec):
                return cls(int(rec["material"]), int(rec["block_iter"]), F(rec["vx"]), F(rec["vy"]))

            def from_s(arr, i, s):
                arr[i] = (s.material, s.block_iter, s.vx, s.vy)

            ns["lattice_info"] = rt.StructArray(self.info, ns["LatticeInfo"], to_s, from_s)
        ns["macro_info"] = rt.Texture16F(self.macro)
        ns["fb"] = rt.Texture16F(self.macro)

    def dispatch(self, which, collide, stream):
        ns = self.mods[which]
        ns["collide_cell"] = rt.StorageF32(collide)
        ns["stream_cell"] = rt.StorageF32(stream)
        main = ns["cs_main"]
        # workgroup (64,4), grid (ceil(nx/64), ceil(ny/4)) (d2q9_node.rs:45): cover the padded grid so the
        # shaders' own bounds checks are exercised
        for gy in range(-(-self.ny // 4) * 4):
            for gx in range(-(-self.nx // 64) * 64):
                if gx >= self.nx + 2:
                    break
                main(rt.Vec([gx, gy, 0]))

    def step(self, n=1):
        for _ in range(n):
            rd, wr = self.buf[self.swap], self.buf[1 - self.swap]
            self.dispatch("collide_stream", rd, wr)
            self.dispatch("boundary", rd, wr)
            self.swap ^= 1

    # ---- derived field (curl_update.wgsl; never dispatched by the reference: fluid_simulator.rs:226,230)
    def curl_update(self):
        """Dispatches lbm/curl_update.wgsl over the current macro texture; returns the (ny, nx, 4) f16 curl texture."""
        if "curl_update" not in self.mods:
            self.mods["curl_update"] = _load("lbm/curl_update.wgsl")
            self._bind_common(self.mods["curl_update"])
        ns = self.mods["curl_update"]
        curl = np.zeros((self.ny, self.nx, 4), np.float16)
        ns["fb"] = rt.Texture16F(self.macro)
        ns["curl_info"] = rt.Texture16F(curl)
        main = ns["cs_main"]
        for gy in range(-(-self.ny // 4) * 4):
            for gx in range(-(-self.nx // 64) * 64):
                main(rt.Vec([gx, gy, 0]))
        return curl

    # ---- colour present of the field (lbm/present.wgsl; `render_node`, built by the reference but its draw call is
    # commented out: fluid_simulator.rs:69-87,243-244)
    def present(self, curl, canvas=None, rows=None):
        """Runs the fragment shader lbm/present.wgsl once per pixel of a canvas_size render target and returns the
        (H, W, 4) float32 fragment outputs (before the surface-format conversion).  Fragment inputs as the rasteriser
        defines them for the bufferless full-screen triangle (bufferless.vs.wgsl): position = pixel centre, uv =
        position.xy / target size (f32 division).  `curl`: the (ny, nx, 4) f16 texture of curl_update()."""
        if "present" not in self.mods:
            self.mods["present"] = _load("lbm/present.wgsl")
            self._bind_common(self.mods["present"])
        ns = self.mods["present"]
        W, H = self.canvas_size
        if canvas is None:
            canvas = np.zeros(W * H, np.dtype([("alpha", "<f4"), ("velocity_x", "<f4"), ("velocity_y", "<f4")]))
        ns["particle_uniform"] = ns["ParticleUniform"]()
        ns["canvas"] = rt.StructArray(canvas, ns["Pixel"],
                                      lambda cls, rec: cls(F(rec["alpha"]), F(rec["velocity_x"]), F(rec["velocity_y"])),
                                      None)
        ns["macro_info"] = rt.Texture16F(self.macro)
        ns["cur_info"] = rt.Texture16F(curl)
        ns["tex_sampler"] = rt.BilinearClampSampler()
        out = np.zeros((H, W, 4), np.float32)
        main, VO = ns["fs_main"], ns["VertexOutput"]
        for py in (range(H) if rows is None else rows):
            for px in range(W):
                pos = rt.Vec([F(px + 0.5), F(py + 0.5), F(0.1), F(1.0)])
                uv = rt.Vec([F(pos.v[0] / F(W)), F(pos.v[1] / F(H))])
                out[py, px, :] = main(VO(uv, pos)).v
        return out

    # ---- particles (particle_update.wgsl)
    def bind_particles(self, pu, particles, canvas):
        ns = self.mods["particle_update"]
        ns["particle_uniform"] = ns["ParticleUniform"](
            rt.Vec([F(c) for c in pu.color]), rt.Vec([int(pu.num[0]), int(pu.num[1])]), int(pu.point_size),
            F(pu.life_time), F(pu.fade_out_factor), F(pu.speed_factor), int(pu.color_ty), int(pu.is_only_update_pos))

        def p_to(cls, rec):
            return cls(rt.Vec([F(rec["pos"][0]), F(rec["pos"][1])]),
                       rt.Vec([F(rec["pos_initial"][0]), F(rec["pos_initial"][1])]), F(rec["life_time"]), F(rec["fade"]))

        def p_from(arr, i, s):
            arr[i] = ((s.pos.x, s.pos.y), (s.pos_initial.x, s.pos_initial.y), s.life_time, s.fade)

        def c_to(cls, rec):
            return cls(F(rec["alpha"]), F(rec["velocity_x"]), F(rec["velocity_y"]))

        def c_from(arr, i, s):
            arr[i] = (s.alpha, s.velocity_x, s.velocity_y)

        ns["particle_buf"] = rt.StructArray(particles, ns["TrajectoryParticle"], p_to, p_from)
        ns["canvas"] = rt.StructArray(canvas, ns["Pixel"], c_to, c_from)
        self._pnum = (int(pu.num[0]), int(pu.num[1]))

    def particle_update(self):
        main = self.mods["particle_update"]["cs_main"]
        for gy in range(-(-self._pnum[1] // 16) * 16):
            for gx in range(-(-self._pnum[0] // 16) * 16):
                main(rt.Vec([gx, gy, 0]))
