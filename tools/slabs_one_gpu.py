"""All slabs of one lattice in ONE process on ONE GPU (SlabGroup): the multi-slab kernel instances (k_frame2<.., SLABS>,
edge row blocks waiting on / signalling the neighbour slabs through device flags) for single-GPU profiling.
usage: slabs_one_gpu.py NX NY N_SLABS UPDATES"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import simuverse_b200 as sb
from simuverse_b200 import wire as W
from simuverse_b200.slabs import SlabGroup

nx, ny, n_slabs, updates = (int(a) for a in sys.argv[1:5])
grp = SlabGroup((nx * 2, ny * 2), sb.SettingObj(animation_type=W.POISEUILLE), lattice=(nx, ny), n_slabs=n_slabs,
                device_preset=W.POISEUILLE)
grp.step_n(updates)
grp.sync()
t0 = time.perf_counter()
grp.step_n(updates)
grp.sync()
dt = time.perf_counter() - t0
print(f"{n_slabs} slabs of {nx}x{ny // n_slabs} on one GPU: {updates} updates in {dt * 1e3:.2f} ms wall, sweeps {grp.nodes[0].fused_sweep_count}")
grp.close()
