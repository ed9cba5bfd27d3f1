#!/usr/bin/env python
"""Headless "LBM player": drives simuverse_b200.FluidSimulator the way SimuverseApp drives the reference's
FluidSimulator (simuverse/src/simuverse_app.rs:182-254 — compute, then draw, once per frame), with a
scripted click (add_obstacle) and drag (add_external_force), and writes the speed field as PGM images.

    python examples/headless_player.py --frames 600 --every 200 --out /tmp/lbm

Needs a CUDA device (there is no CPU fallback).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simuverse_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--canvas", type=int, nargs=2, default=(1200, 750), help="window size in physical pixels")
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--every", type=int, default=200, help="dump the field every N frames (0 = never)")
    ap.add_argument("--out", default="lbm_out")
    ap.add_argument("--preset", choices=["poiseuille", "cavity", "custom"], default="poiseuille")
    args = ap.parse_args()
    preset = {"poiseuille": sb.wire.POISEUILLE, "cavity": sb.wire.LID_DRIVEN_CAVITY, "custom": sb.wire.CUSTOM}[args.preset]
    sim = sb.FluidSimulator(tuple(args.canvas), sb.SettingObj(animation_type=preset), particles=True)
    node = sim.fluid_compute_node
    nx, ny = sim.lattice
    print(f"lattice {nx}x{ny}, {sim.particles_num[0]}x{sim.particles_num[1]} tracer particles")
    if args.every:
        os.makedirs(args.out, exist_ok=True)
    t0 = time.perf_counter()
    for frame in range(1, args.frames + 1):
        if frame == args.frames // 3:  # a click in the right half of the channel
            print("click ->", sim.on_click((args.canvas[0] * 0.6, args.canvas[1] * 0.5)))
        if args.frames // 2 <= frame < args.frames // 2 + 20:  # a short drag
            sim.touch_move((args.canvas[0] * 0.3 + 6.0 * (frame - args.frames // 2), args.canvas[1] * 0.3))
        sim.compute()          # FluidSimulator::compute: step(0), particles, step(1), particles
        sim.draw_by_rpass()    # canvas fade of the present pass
        if args.every and frame % args.every == 0:
            tex = node.read_macro_tex().astype(np.float32)     # (ny, nx, 4): u.x, u.y, rho, 1
            speed = np.hypot(tex[..., 0], tex[..., 1])
            img = np.clip(speed / 0.25 * 255.0, 0, 255).astype(np.uint8)
            path = os.path.join(args.out, f"speed_{frame:05d}.pgm")
            with open(path, "wb") as f:
                f.write(f"P5 {nx} {ny} 255\n".encode() + img.tobytes())
            print(f"frame {frame}: mass {node.total_mass():.3f}, max speed {speed.max():.4f} -> {path}")
    node.sync()
    dt = time.perf_counter() - t0
    print(f"{args.frames} frames ({2 * args.frames} lattice updates) in {dt:.2f} s = "
          f"{2 * args.frames * nx * ny / dt / 1e6:.0f} MLUPS including the host loop")
    node.close()


if __name__ == "__main__":
    main()
