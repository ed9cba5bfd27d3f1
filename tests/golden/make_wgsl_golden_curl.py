"""Golden vector of the reference's derived-field pass, assets/wgsl/lbm/curl_update.wgsl:12-33, produced by EXECUTING
that shader text (tests/wgsl_ref) on a macro texture that itself came out of the executed collide_stream.wgsl: the
texture stored in wgsl_channel100_64x48_s100.npz (channel with obstacles, 100 updates).

Needs /root/reference (build container only); the output is committed.
Run from the repo root:  python tests/golden/make_wgsl_golden_curl.py
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
    nx, ny = int(g["nx"]), int(g["ny"])
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    sim = H.WgslLbm(nx, ny, g["info"], lbm_uniform_new(tau, 0, nx * ny))
    sim.macro[...] = g["macro_f16"].view(np.float16).reshape(ny, nx, 4)
    curl = sim.curl_update()
    out = os.path.join(HERE, "wgsl_curl_64x48.npz")
    np.savez_compressed(out, nx=nx, ny=ny, macro_f16=g["macro_f16"], curl_f16=curl.view(np.uint16))
    print(out, os.path.getsize(out), "bytes; distinct curl values:", len(np.unique(curl[..., 0])))


if __name__ == "__main__":
    main()
