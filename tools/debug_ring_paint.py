"""Debug helper: paint solids next to the ghost column mid-run and list where the CUDA path and the oracle differ."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import simuverse_b200 as sb
from simuverse_b200 import wire as W
import oracle as orc
from helpers import tau_default

nx, ny = 240, 160
for flags, name in ((sb.FLAG_NO_FUSE, "single"), (0, "sweeps")):
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = sb.D2Q9Node((nx * 2, ny * 2), sb.SettingObj(animation_type=W.POISEUILLE), lattice=(nx, ny), lattice_info=info, flags=flags)
    sim = orc.OracleSim(nx, ny, info, orc.uniform_new(tau_default(), 0, nx * ny), threads=4)
    a.step_n(50); sim.step(50)
    cells = np.zeros(2, W.LATTICE_INFO_DTYPE); cells["material"] = W.OBSTACLE; cells["block_iter"] = -1
    for y in range(70, 74):
        o = (y * nx + 1) * 16
        a.write_lattice_info(o, cells); sim.write_lattice_info(o, cells)
    for k in (0, 1, 1, 2, 26):
        a.step_n(k); sim.step(k)
        for which in (0, 1):
            got, want = a.read_distributions(which), sim.distributions(which)
            bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
            print(name, "after +%d" % k, "buf", which, "diffs", len(bad), [tuple(int(v) for v in b) + (float(got[tuple(b)]), float(want[tuple(b)])) for b in bad[:12]])
    print(name, "sweeps", a.fused_sweep_count)
    a.close()
