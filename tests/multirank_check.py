"""Run under torch.distributed.run with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multirank_check.py [nx ny steps]

Every rank owns one y-slab (one process per GPU, IPC-mapped neighbour memory over NVLink); rank 0
gathers the slabs and checks them bit for bit against a single-slab run on its own GPU and against
the CPU oracle.  Prints MULTIRANK_OK on success.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import simuverse_b200 as sb  # noqa: E402
from simuverse_b200 import wire as W  # noqa: E402
from simuverse_b200.slabs import SlabRank  # noqa: E402


def main():
    nx, ny, steps = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (520, 384, 120)
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    info = sb.init_lattice_material(nx, ny, W.POISEUILLE)
    g = info.reshape(ny, nx)
    for r in range(1, world):  # obstacles straddling every slab cut, plus a force patch on a cut row
        cut = ny * r // world
        g["material"][cut - 4:cut + 5, 100 + 40 * r:130 + 40 * r] = W.OBSTACLE
        g[cut, 300:310] = (W.EXTERNAL_FORCE, -1, 0.03, 0.05)
        g[cut - 1, 320:330] = (W.EXTERNAL_FORCE, -1, -0.02, -0.06)
    setting = sb.SettingObj(animation_type=W.POISEUILLE)
    slab = SlabRank((nx * 2, ny * 2), setting, lattice=(nx, ny), dist=dist, device=local, lattice_info=info)
    slab.step_n(steps)
    slab.barrier()
    mass = slab.total_mass()
    got = slab.gather_distributions()  # global (9, ny, nx) on every rank
    ok = True
    if rank == 0:
        one = sb.D2Q9Node((nx * 2, ny * 2), setting, lattice=(nx, ny), lattice_info=info, device=local)
        one.step_n(steps)
        want = one.read_distributions(one.swap_index)
        same_gpu = np.array_equal(got.view(np.uint32), want.view(np.uint32))
        import oracle as orc

        tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
        sim = orc.OracleSim(nx, ny, info, orc.uniform_new(tau, 0, nx * ny), threads=orc.lib().orc_get_max_threads())
        sim.step(steps)
        same_orc = np.array_equal(got.view(np.uint32), sim.distributions(sim.swap).view(np.uint32))
        mass_ok = abs(mass - sim.total_mass()) / sim.total_mass() < 1e-9
        ok = same_gpu and same_orc and mass_ok
        print(f"world={world} lattice={nx}x{ny} steps={steps}: slabs==single-GPU {same_gpu}, slabs==oracle {same_orc}, "
              f"mass {mass:.6f} vs {sim.total_mass():.6f}")
        one.close()
    slab.barrier()
    slab.node.close()
    dist.barrier()
    if rank == 0:
        print("MULTIRANK_OK" if ok else "MULTIRANK_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
