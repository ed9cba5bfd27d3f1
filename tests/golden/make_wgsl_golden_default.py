"""tests/golden/wgsl_default_600x375_f2.npz — the reference's DEFAULT configuration (BASELINE configs[0]:
600x375 lattice, Poiseuille preset with its three R=28 discs, tau 0.56, 127x80 tracer particles), two
frames of FluidSimulator::compute executed from the reference's own WGSL source (see make_wgsl_golden.py).
The state is large, so the fixture keeps SHA-256 digests of every buffer plus three full rows.
Takes ~10 minutes of pure Python.  Run from the repo root:  python tests/golden/make_wgsl_golden_default.py
"""
import hashlib
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
from simuverse_b200.fluid_simulator import init_trajectory_particles, particle_grid  # noqa: E402
from wgsl_ref import harness as H  # noqa: E402

ROWS = (1, 187, 373)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    nx, ny, frames = 600, 375, 2
    canvas_size = (1200, 750)
    info = init_lattice_material(nx, ny, W.POISEUILLE)
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    t0 = time.time()
    sim = H.WgslLbm(nx, ny, info, lbm_uniform_new(tau, 0, nx * ny), canvas=canvas_size)
    setting = SettingObj()
    pu = setting.particles_uniform_data
    num = particle_grid(canvas_size, setting.particles_count)
    assert num == (127, 80)
    pu.num[:] = list(num)
    parts = init_trajectory_particles(canvas_size, num, pu.life_time, 0x5EED)
    canvas = np.zeros(canvas_size[0] * canvas_size[1], W.PIXEL_DTYPE)
    sim.bind_particles(pu, parts, canvas)
    for f in range(frames):
        for _ in range(2):
            sim.step(1)
            sim.particle_update()
        print(f"frame {f + 1}/{frames}: {time.time() - t0:.0f} s", flush=True)
    cur = sim.buf[sim.swap].reshape(9, ny, nx)
    prev = sim.buf[1 - sim.swap].reshape(9, ny, nx)
    out = os.path.join(HERE, "wgsl_default_600x375_f2.npz")
    np.savez_compressed(
        out, nx=nx, ny=ny, frames=frames, swap=sim.swap, rows=np.array(ROWS),
        sha_cur=digest(cur), sha_prev=digest(prev), sha_macro=digest(sim.macro.view(np.uint16)),
        sha_info=digest(sim.info), sha_particles=digest(parts), sha_canvas=digest(canvas),
        cur_rows=cur[:, ROWS, :], prev_rows=prev[:, ROWS, :], macro_rows=sim.macro.view(np.uint16)[list(ROWS)],
        total_mass=float(np.sum(cur.astype(np.float64))), particles=parts)
    print(out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
