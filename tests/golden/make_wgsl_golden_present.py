"""Golden vector of the reference's colour present of the field, assets/wgsl/lbm/present.wgsl:21-46 (fragment shader of
`render_node`, fluid_simulator.rs:69-87), produced by EXECUTING that shader text (tests/wgsl_ref) once per pixel of a
canvas_size target, on the macro texture of wgsl_channel100_64x48_s100.npz (executed collide_stream.wgsl, 100 updates)
and the curl texture of wgsl_curl_64x48.npz (executed curl_update.wgsl).  Two targets: the reference's own ratio
(lattice_pixel_size 2: 128 x 96 pixels) and a non-integer one (150 x 101) that exercises every filter weight.
textureSample is the restatement in tests/wgsl_ref/runtime.py (f32 weights, WebGPU's formula, ClampToEdge).

Needs /root/reference (build container only); the output is committed.
Run from the repo root:  python tests/golden/make_wgsl_golden_present.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from simuverse_b200.d2q9_node import lbm_uniform_new  # noqa: E402
from wgsl_ref import harness as H  # noqa: E402


def main():
    if not H.available():
        raise SystemExit("needs the reference tree at /root/reference")
    g = np.load(os.path.join(HERE, "wgsl_channel100_64x48_s100.npz"))
    c = np.load(os.path.join(HERE, "wgsl_curl_64x48.npz"))
    nx, ny = int(g["nx"]), int(g["ny"])
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    out = {}
    for tag, canvas in (("a", (2 * nx, 2 * ny)), ("b", (150, 101))):
        sim = H.WgslLbm(nx, ny, g["info"], lbm_uniform_new(tau, 0, nx * ny), canvas=canvas)
        sim.macro[...] = g["macro_f16"].view(np.float16).reshape(ny, nx, 4)
        rgba = sim.present(c["curl_f16"].view(np.float16).reshape(ny, nx, 4))
        out[f"canvas_{tag}"] = np.array(canvas, np.int32)
        out[f"rgba_{tag}"] = rgba
        print(tag, canvas, "distinct colours:", len(np.unique(rgba.reshape(-1, 4), axis=0)))
    path = os.path.join(HERE, "wgsl_present_64x48.npz")
    np.savez_compressed(path, nx=nx, ny=ny, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
