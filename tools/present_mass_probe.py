#!/usr/bin/env python
"""Runs the derived-field passes on a 4096^2 channel once (for ncu / timing): lbm_read_curl, lbm_read_present for a
canvas of 8192 x 2048 pixels per call (rows streamed in 4 calls), lbm_total_mass.  Prints CUDA-event-free wall times (the
calls synchronise) — kernel times come from ncu."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simuverse_b200 as sb  # noqa: E402
from simuverse_b200 import wire as W  # noqa: E402

nx = ny = 4096
s = sb.SettingObj()
s.animation_type = W.POISEUILLE
node = sb.D2Q9Node((2 * nx, 2 * ny), s, lattice=(nx, ny), device_preset=W.POISEUILLE, flags=sb.FLAG_MACRO_EVERY_STEP)
node.step_n(40)
node.sync()
for name, fn in (("total_mass", lambda: node.total_mass()), ("read_curl_tex", lambda: node.read_curl_tex()),
                 ("read_present 8192x2048", lambda: [node.read_present(r, 2048) for r in range(0, 8192, 2048)][-1])):
    fn()
    t = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t
    print(f"{name}: {dt * 1e3:.2f} ms wall (incl. device-to-host copy)", np.asarray(out).shape if name != "total_mass" else out)
node.close()
