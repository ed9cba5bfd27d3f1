"""Device time of the tracer-particle pass alone on BASELINE configs[4] (8192^2 porous, 1000x1000 particles)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simuverse_b200 as sb
from simuverse_b200 import wire as W

nx = ny = 8192
s = sb.SettingObj(animation_type=W.POISEUILLE, particles_count=1000000)
fs = sb.FluidSimulator((nx * 2, ny * 2), s, particles=True, lattice=(nx, ny), device_preset=sb.PRESET_POROUS)
node = fs.fluid_compute_node
fs.compute(20)
node.sync()
t0 = time.perf_counter()
fs.compute(50)
node.sync()
frame_ms = (time.perf_counter() - t0) / 50 * 1e3
for _ in range(20):
    node.particles_update()
node.sync()
t0 = time.perf_counter()
for _ in range(200):
    node.particles_update()
node.sync()
pass_us = (time.perf_counter() - t0) / 200 * 1e6
print(f"block_x={os.environ.get('LBM_PARTICLE_BLOCK_X', 'default')}: frame {frame_ms:.3f} ms = {nx * ny * 2 / frame_ms / 1e6:.1f} GLUPS, "
      f"particle pass {pass_us:.1f} us ({fs.particles_num[0]}x{fs.particles_num[1]} particles)")
node.close()
