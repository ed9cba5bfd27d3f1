"""Generates tests/golden/wgsl_*.npz by EXECUTING THE REFERENCE'S OWN WGSL SOURCE.

The reference's arithmetic for this path lives in assets/wgsl/lbm/{init,collide_stream,boundary,
particle_update}.wgsl.  It cannot be run through wgpu here (no Rust toolchain, no WebGPU backend), so
tests/wgsl_ref transpiles the unmodified shader text (after the reference's own #include expansion) to
Python and evaluates it with IEEE f32 scalars, dispatching the passes in the order of
D2Q9Node::compute_by_pass / FluidSimulator::compute.  The resulting vectors pin both the CPU oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py).

Needs /root/reference (present in the build container only); the outputs are committed.
Run from the repo root:  python tests/golden/make_wgsl_golden.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from simuverse_b200 import wire as W  # noqa: E402
from simuverse_b200.d2q9_node import SettingObj, init_lattice_material, lbm_uniform_new  # noqa: E402
from simuverse_b200.fluid_simulator import init_trajectory_particles  # noqa: E402
from wgsl_ref import harness as H  # noqa: E402

TAU = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))


def channel(nx, ny):
    """Poiseuille frame (lattice.rs:65-76) with small obstacles instead of the R=28 discs."""
    info = np.zeros(nx * ny, W.LATTICE_INFO_DTYPE)
    g = info.reshape(ny, nx)
    g["material"], g["block_iter"] = W.BULK, -1
    g["material"][0, :] = g["material"][ny - 1, :] = W.BOUNDARY
    g["material"][1:ny - 1, 0] = g["material"][1:ny - 1, nx - 1] = W.GHOST
    g["material"][1:ny - 1, 1] = W.INLET
    g["vx"][1:ny - 1, 1] = 0.12
    g["material"][1:ny - 1, nx - 2] = W.OUTLET
    yy, xx = np.mgrid[0:ny, 0:nx]
    g["material"][(xx - nx // 4) ** 2 + (yy - ny // 2) ** 2 <= 16] = W.OBSTACLE
    g["material"][ny // 4:ny // 4 + 3, nx // 2:nx // 2 + 5] = W.OBSTACLE
    return info


def case_channel():
    nx, ny = 72, 48
    return dict(nx=nx, ny=ny, fluid_ty=0, info=channel(nx, ny), steps=40, post=[])


def case_channel100():
    """north_star's bar is stated "after 100 steps": the channel with obstacles (inlet, outlet, ghost columns, walls,
    a disc and a bar) for exactly 100 updates of the reference's WGSL."""
    nx, ny = 64, 48
    return dict(nx=nx, ny=ny, fluid_ty=0, info=channel(nx, ny), steps=100, post=[])


def case_cavity():
    nx, ny = 40, 32
    return dict(nx=nx, ny=ny, fluid_ty=1, info=init_lattice_material(nx, ny, W.LID_DRIVEN_CAVITY), steps=40, post=[])


def case_cavity100():
    """The lid-driven cavity (fluid_ty = 1: the other init branch, ghost row wrapping onto a solid row, a whole row of force
    cells) carried to north_star's "after 100 steps" too, on an odd-sized lattice."""
    nx, ny = 46, 38
    return dict(nx=nx, ny=ny, fluid_ty=1, info=init_lattice_material(nx, ny, W.LID_DRIVEN_CAVITY), steps=100, post=[])


def case_force():
    nx, ny = 36, 30
    info = init_lattice_material(nx, ny, W.CUSTOM)
    g = info.reshape(ny, nx)
    g["material"][12:15, 20:24] = W.OBSTACLE
    g[25, 5] = (W.EXTERNAL_FORCE, 9, 0.07, 0.0)  # armed before init: init.wgsl:51-59 disarms it
    post = [(10, 10, 7, 0.05, 0.02), (11, 10, 90, -0.03, 0.04), (20, 18, 1, 0.1, 0.0), (12, 22, 100, 0.0, -0.11),
            (1, 1, 30, 0.06, 0.06)]
    return dict(nx=nx, ny=ny, fluid_ty=0, info=info, steps=100, post=post)


def case_periodic():
    nx, ny = 22, 14
    info = np.zeros(nx * ny, W.LATTICE_INFO_DTYPE)
    g = info.reshape(ny, nx)
    g["material"], g["block_iter"] = W.BULK, -1
    g[0, 0] = (W.EXTERNAL_FORCE, -1, 0.06, 0.05)
    g["material"][6:8, 9:12] = W.OBSTACLE
    g["material"][0, 15:18] = W.OBSTACLE       # solids on the ring and across the periodic wrap
    g["material"][4:6, nx - 1] = W.OBSTACLE
    return dict(nx=nx, ny=ny, fluid_ty=0, info=info, steps=30, post=[])


def case_particles():
    nx, ny = 60, 40
    return dict(nx=nx, ny=ny, fluid_ty=0, info=channel(nx, ny), steps=40, post=[], particles=(9, 6))


def case_midrun():
    """add_obstacle-style mask change in the middle of a run: the new solid cells keep their last
    distributions, which their neighbours pull for exactly one more step."""
    nx, ny = 48, 36
    c = dict(nx=nx, ny=ny, fluid_ty=0, info=channel(nx, ny), steps=44, post=[])
    c["midrun"] = dict(after=21, x0=30, x1=36, y0=14, y1=21)
    return c


CASES = {"wgsl_midrun_48x36_s44": case_midrun, "wgsl_channel_72x48_s40": case_channel, "wgsl_cavity_40x32_s40": case_cavity,
         "wgsl_force_36x30_s100": case_force, "wgsl_periodic_22x14_s30": case_periodic,
         "wgsl_particles_60x40_f20": case_particles, "wgsl_channel100_64x48_s100": case_channel100,
         "wgsl_cavity100_46x38_s100": case_cavity100}


def main():
    if not H.available():
        raise SystemExit("needs the reference tree at /root/reference")
    only = sys.argv[1:]
    for name, make in CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        c = make()
        nx, ny = c["nx"], c["ny"]
        u = lbm_uniform_new(TAU, c["fluid_ty"], nx * ny)
        sim = H.WgslLbm(nx, ny, c["info"], u)
        post_off = np.array([(y * nx + x) * 16 for (x, y, *_rest) in c["post"]], np.uint64)
        post = np.array([(W.EXTERNAL_FORCE, it, vx, vy) for (_, _, it, vx, vy) in c["post"]], W.LATTICE_INFO_DTYPE)
        for off, cell in zip(post_off, post):  # queue.write_buffer(info_buf, off, cell) after init (d2q9_node.rs:298)
            sim.info[int(off) // 16] = cell
        extra = {}
        if "particles" in c:
            num = c["particles"]
            canvas_size = (nx * 2, ny * 2)
            pu = SettingObj().particles_uniform_data
            pu.num[:] = list(num)
            parts = init_trajectory_particles(canvas_size, num, pu.life_time, 0x5EED)
            extra["particles_init"] = parts.copy()
            canvas = np.zeros(canvas_size[0] * canvas_size[1], W.PIXEL_DTYPE)
            sim.bind_particles(pu, parts, canvas)
            for _ in range(c["steps"] // 2):  # FluidSimulator::compute: step(0), particles, step(1), particles
                sim.step(1)
                sim.particle_update()
                sim.step(1)
                sim.particle_update()
            extra.update(particles=parts, canvas=canvas, particle_num=np.array(num))
        elif "midrun" in c:
            m = c["midrun"]
            sim.step(m["after"])
            g2 = sim.info.reshape(ny, nx)
            g2[m["y0"]:m["y1"], m["x0"]:m["x1"]] = (W.OBSTACLE, -1, 0.0, 0.0)  # queue.write_buffer(info_buf, ..)
            sim.step(c["steps"] - m["after"])
            extra.update(midrun=np.array([m["after"], m["x0"], m["x1"], m["y0"], m["y1"]]))
        else:
            sim.step(c["steps"])
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            out, nx=nx, ny=ny, fluid_ty=c["fluid_ty"], steps=c["steps"], info=c["info"], post_cells=post,
            post_offsets=post_off, swap=sim.swap, buf_cur=sim.buf[sim.swap].reshape(9, ny, nx),
            buf_prev=sim.buf[1 - sim.swap].reshape(9, ny, nx), macro_f16=sim.macro.view(np.uint16),
            info_after=sim.info, **extra)
        print(f"{name}: {os.path.getsize(out)} bytes, {time.time() - t0:.1f} s")


if __name__ == "__main__":
    main()
