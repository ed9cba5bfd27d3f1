"""Generates tests/golden/*.npz.

The reference (Rust + WGSL on wgpu) cannot be executed in this environment and ships no
golden vectors, so these fixtures are produced by tests/np_restatement.py — the numpy
restatement of the WGSL step that is written independently of the C oracle — and pin BOTH
the oracle (tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py) to a committed
answer.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import np_restatement as R  # noqa: E402
from simuverse_b200 import wire as W  # noqa: E402
from simuverse_b200.d2q9_node import init_lattice_material, lbm_uniform_new  # noqa: E402

CASES = {
    # name: (nx, ny, preset, steps)
    "poiseuille_96x64_s60": (96, 64, W.POISEUILLE, 60),
    "cavity_48x40_s80": (48, 40, W.LID_DRIVEN_CAVITY, 80),
    "custom_force_40x36_s100": (40, 36, W.CUSTOM, 100),
}


def make_info(name, nx, ny, preset):
    info = init_lattice_material(nx, ny, preset)
    if name.startswith("poiseuille"):
        # the preset discs (R=28) do not fit a 96x64 lattice sensibly: add a small block obstacle
        g = info.reshape(ny, nx)
        g["material"][20:30, 40:47] = W.OBSTACLE
    if name.startswith("custom"):
        g = info.reshape(ny, nx)
        g["material"][15:19, 25:28] = W.OBSTACLE
        g[30, 5] = (W.EXTERNAL_FORCE, 9, 0.07, 0.0)  # armed before init: init.wgsl:51-59 disarms it
    return info


# transient force cells written AFTER init, as add_external_force does (d2q9_node.rs:288-298):
# (x, y, block_iter, vx, vy)
POST_INIT_FORCE = [(10, 10, 7, 0.05, 0.02), (11, 10, 90, -0.03, 0.04), (20, 18, 1, 0.1, 0.0), (12, 22, 100, 0.0, -0.11)]


def main():
    for name, (nx, ny, preset, steps) in CASES.items():
        info = make_info(name, nx, ny, preset)
        fluid_ty = 1 if preset == W.LID_DRIVEN_CAVITY else 0
        u = lbm_uniform_new(np.float32(3.0) * np.float32(0.02) + np.float32(0.5), fluid_ty, nx * ny)
        sim = R.NpSim(nx, ny, info, u)
        post = np.zeros(0, W.LATTICE_INFO_DTYPE)
        post_off = np.zeros(0, np.uint64)
        if name.startswith("custom"):
            post = np.array([(W.EXTERNAL_FORCE, it, vx, vy) for (_, _, it, vx, vy) in POST_INIT_FORCE], W.LATTICE_INFO_DTYPE)
            post_off = np.array([(y * nx + x) * 16 for (x, y, _, _, _) in POST_INIT_FORCE], np.uint64)
            for (x, y, _, _, _), c in zip(POST_INIT_FORCE, post):
                sim.material[y, x], sim.block_iter[y, x] = c["material"], c["block_iter"]
                sim.vx[y, x], sim.vy[y, x] = c["vx"], c["vy"]
        sim.step(steps)
        ux, uy, rho = sim.macro
        out = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            out, nx=nx, ny=ny, preset=preset, steps=steps, info=info, post_cells=post, post_offsets=post_off,
            buf_cur=sim.buf[sim.swap], buf_prev=sim.buf[1 - sim.swap], swap=sim.swap,
            ux=ux, uy=uy, rho=rho, material=sim.material, block_iter=sim.block_iter)
        print(name, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
