#!/usr/bin/env python
"""bench.py — MLUPS of the D2Q9 lattice-Boltzmann step on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One bench "step" = one lattice update (pull-stream + BGK collide + bounce-back, the reference's
compute_by_pass) of the whole lattice.  Workload at every N: BASELINE configs[2], ONE 16384x16384
channel split into N y-slabs (strong scaling, N = 1 included, so that the 1 -> 8 curve is one
workload); under torch.distributed.run one rank per GPU, the slabs exchange edge rows through peer
memory inside the step kernel, there is no collective on the data path.  At N = 1 the line carries
BASELINE configs[1] (4096x4096, the single-GPU roofline target) as a second block, `secondary`.
At N > 1 a cross-process parity check (slabs == one GPU == CPU oracle, bitwise) runs before timing.

Prints ONE JSON line (rank 0).  `value` = lattice sites updated per second / 1e6 with the state
resident in HBM, timed with CUDA events on the library's stream, max over ranks.  `e2e` = the same
metric through the public host API with HOST buffers inside the timed region (per step: upload of a
56-row LatticeInfo patch from pinned memory — the reference's add_obstacle write — one step, and a
read-back of the RGBA16F macro field into pinned memory).  `roofline` and `cpu_baseline` as
specified in DESIGN.md.  `--impl reference` times the CPU oracle (the reference cannot be built
here: no Rust / WebGPU) on all host cores instead.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BYTES_PER_SITE = 72  # 9 f32 read + 9 f32 written
METRIC = "MLUPS (D2Q9 f32)"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(config, overridden, fused=False, world=1):
    """DRAM bytes per launch of the dominant kernel from the committed ncu captures (profiles/traffic.json).  Multi-slab
    runs of config 3: the capture of one 16384x2048 slab at 8 slabs; for 2 and 4 slabs the one-slab figure / N (the slab
    instance moves the same bytes per site, profiles/r02_k_frame2_slab_ncu_full.md)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = str(config) + ("_fused" if fused else "")
        if overridden:
            return None, None
        if world > 1:
            if config != 3 or not fused:
                return None, None
            if world == 8:
                return t["3_fused_slab8"]["bytes_per_launch"], t["3_fused_slab8"]["source"]
            return t[key]["bytes_per_launch"] // world, t[key]["source"] + f" / {world} slabs"
        if key in t:
            return t[key]["bytes_per_launch"], t[key]["source"]
    except Exception:
        pass
    return None, None


def workload_for(n_gpus, override, config=None):
    """(nx, ny, name, kind) of the BASELINE.json config being run. kind: 'steps' or 'frames'."""
    if config is None:
        config = 3  # BASELINE configs[2]: ONE 16384x16384 lattice on 1/2/4/8 GPUs (N = 1 included: 19.3 + 4.3 GB)
    if override:
        nx, ny = override
        return nx, ny, f"D2Q9 BGK {nx}x{ny} channel + cylinder obstacles (Poiseuille preset), f32", "steps", config
    if config == 1:
        return 600, 375, ("simuverse default lattice 600x375, channel flow past 3 cylinders, 127x80 tracer particles, "
                          "frame loop of FluidSimulator::compute (BASELINE configs[0])"), "frames", config
    if config == 2:
        return 4096, 4096, "D2Q9 BGK 4096x4096 channel + cylinder obstacle, f32, single B200 (BASELINE configs[1])", "steps", config
    if config == 3:
        return 16384, 16384, (f"D2Q9 16384x16384 strong scaling, {n_gpus} y-slab(s) with peer-memory edge rows "
                              "(BASELINE configs[2])"), "steps", config
    if config == 4:
        return 16384, 16384 * n_gpus, (f"D2Q9 weak scaling, 16384x16384 per GPU, {n_gpus} y-slab(s) = 16384x{16384 * n_gpus} "
                                       "(BASELINE configs[3])"), "steps", config
    if config == 5:
        return 8192, 8192, ("D2Q9 8192x8192 porous-media mask (30% solid) + 1000x1000 tracer particles updated after "
                            "every step (BASELINE configs[4])"), "frames", config
    raise SystemExit(f"unknown --config {config}")


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._mark = None
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, bits))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self, t0=None, t1=None):
        self._stop.set()
        if self._t:
            self._t.join()
        win = [s for s in self.samples if t0 is None or t0 <= s[0] <= t1] or self.samples
        if not win:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        bits = 0
        for s in win:
            bits |= s[2]
        reasons = [name for b, name in self.REASONS.items() if bits & b]
        return {"sm_mhz": statistics.median(s[1] for s in win), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(win)}


def bind_to_gpu_numa_node(index):
    """Pins this process (and, by first touch, the pinned host buffers it allocates afterwards) to the CPUs of the NUMA
    node GPU `index` hangs off: with N ranks reading their fields back at once, buffers that all land on one socket
    share that socket's memory and PCIe root.  Returns a short description for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
        return f"numa node {node}, {len(ids)} cpus"
    except Exception as e:  # not fatal: the run is only slower
        return f"unbound ({type(e).__name__})"


def shared_config(name, config, nx, ny, world):
    """The `config` object of the JSON line — identical in the b200 arm and the reference arm."""
    return {"workload": name, "baseline_config": config, "lattice": [nx, ny], "tau": 0.56,
            "l2": "inputs_larger_than_l2" if nx * ny * 72 / world > 2.6e8 else "lattice fits in L2 (the reference's own size)"}


CPU_BAND_ROWS = 2048  # largest band of a big lattice the CPU legs update per step (memory: 112 B per site)


def oracle_band(nx, ny, rows, porous):
    """OracleSim over the first `rows` rows of the workload's lattice (the whole lattice when it is small)."""
    import numpy as np

    import oracle as orc

    rows = min(rows, ny)
    cores = orc.lib().orc_get_max_threads()
    info = orc.init_porous_material(nx, rows) if porous else orc.init_lattice_material(nx, rows, 4)
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    return orc.OracleSim(nx, rows, info, orc.uniform_new(tau, 0, (nx * rows) & 0x7FFFFFFF), threads=cores), rows, cores


def cpu_leg(nx, ny, budget_s, porous=False, min_steps=3):
    """cpu_baseline: the CPU oracle (OpenMP, all host cores) on a bounded band of the same workload."""
    sim, rows, cores = oracle_band(nx, ny, CPU_BAND_ROWS, porous)
    sim.step(1)
    t0 = time.perf_counter()
    n = 0
    while n < min_steps or time.perf_counter() - t0 < budget_s:
        sim.step(1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": nx * rows * n / dt / 1e6, "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"{nx}x{rows} band of the {nx}x{ny} lattice, {n} steps in {dt:.1f} s (CPU restatement of the reference WGSL, OpenMP)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own step on the host cores.  The reference (Rust + WGSL via
    wgpu) cannot be built in this image, so this is the oracle port; each step updates a bounded
    band of the workload's lattice so that K+W steps end within a few minutes."""
    if rank != 0:
        return
    nx, ny, name, _, config = workload_for(args.gpus, args.lattice, args.config)
    # calibrate on a thin band, then size the band for ~150 s total (capped by memory)
    sim, rows, cores = oracle_band(nx, ny, 64, config == 5)
    sim.step(1)
    t0 = time.perf_counter()
    sim.step(3)
    per_row = (time.perf_counter() - t0) / 3 / rows
    total = args.steps + args.warmup
    rows = int(max(64, min(ny, CPU_BAND_ROWS, 150.0 / total / per_row)))
    sim, rows, cores = oracle_band(nx, ny, rows, config == 5)
    sim.step(args.warmup)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    v = nx * rows * args.steps / dt / 1e6
    sample = f"{nx}x{rows} band of the {nx}x{ny} lattice per step, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": scaling_of(config), "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": shared_config(name, config, nx, ny, args.gpus),
        "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def scaling_of(config):
    """configs[2] cuts ONE 16384x16384 lattice into N slabs (strong, also at N = 1); configs[3] stacks one per GPU."""
    return "strong" if config == 3 else "weak"


def multirank_parity(sb, W, SlabRank, dist, rank, world, local_rank):
    """N > 1, before anything is timed: the comparison of tests/multirank_check.py — a 520x384 channel with obstacles and
    force patches straddling every slab cut, 120 updates on `world` slabs (one process per GPU, peer-memory edge rows),
    gathered and compared bit for bit with ONE GPU and with the CPU oracle (the checker).  Returns "ok" or raises."""
    import numpy as np
    import torch

    nx, ny, steps = 520, 384, 120
    info = sb.init_lattice_material(nx, ny, W.POISEUILLE)
    g = info.reshape(ny, nx)
    for r in range(1, world):
        cut = ny * r // world
        g["material"][cut - 4:cut + 5, 100 + 40 * r:130 + 40 * r] = W.OBSTACLE
        g[cut, 300:310] = (W.EXTERNAL_FORCE, -1, 0.03, 0.05)
        g[cut - 1, 320:330] = (W.EXTERNAL_FORCE, -1, -0.02, -0.06)
    setting = sb.SettingObj(animation_type=W.POISEUILLE)
    slab = SlabRank((nx * 2, ny * 2), setting, lattice=(nx, ny), dist=dist, device=local_rank, lattice_info=info)
    slab.step_n(steps)
    slab.barrier()
    sweeps = slab.node.fused_sweep_count
    got = slab.gather_distributions()
    verdict = [1]
    if rank == 0:
        one = sb.D2Q9Node((nx * 2, ny * 2), setting, lattice=(nx, ny), lattice_info=info, device=local_rank)
        one.step_n(steps)
        same_gpu = np.array_equal(got.view(np.uint32), one.read_distributions(one.swap_index).view(np.uint32))
        one.close()
        import oracle as orc

        tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
        sim = orc.OracleSim(nx, ny, info, orc.uniform_new(tau, 0, nx * ny), threads=orc.lib().orc_get_max_threads())
        sim.step(steps)
        same_orc = np.array_equal(got.view(np.uint32), sim.distributions(sim.swap).view(np.uint32))
        verdict = [1 if (same_gpu and same_orc and sweeps > 0) else 0]
        if not verdict[0]:
            sys.stderr.write(f"multirank parity FAILED: slabs==single-GPU {same_gpu}, slabs==oracle {same_orc}, sweeps {sweeps}\n")
    slab.barrier()
    slab.node.close()
    t = torch.tensor(verdict, dtype=torch.int64, device="cuda")
    dist.broadcast(t, 0)
    if int(t.item()) != 1:
        raise SystemExit("bench.py: cross-process multi-GPU parity check failed; nothing was timed")
    return "ok"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=None, help="BASELINE.json config 1..5 (default 3: 16384x16384 in N slabs)")
    ap.add_argument("--aa", action="store_true", help="AA-pattern in-place variant (single copy of the distributions)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of CUDA graphs")
    ap.add_argument("--lattice", type=int, nargs=2, default=None, metavar=("NX", "NY"), help="override the workload")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--generic", action="store_true", help="time the one-thread-per-cell kernel instead")
    ap.add_argument("--no-fuse", action="store_true", help="one lattice update per launch (k_step_vec) instead of two (k_frame2)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the 4096x4096 (BASELINE configs[1]) block at N=1")
    ap.add_argument("--no-parity", action="store_true", help="skip the cross-process parity check at N>1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run with {args.gpus} ranks (see docstring)")
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    import numpy as np
    import torch

    import simuverse_b200 as sb
    from simuverse_b200 import wire as W
    from simuverse_b200.slabs import SlabRank

    if sb.lib.lbm_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: simuverse_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    parity = None
    if world > 1 and not args.no_parity:
        parity = multirank_parity(sb, W, SlabRank, dist, rank, world, local_rank)

    base_flags = ((sb.FLAG_KERNEL_GENERIC if args.generic else 0) | (sb.FLAG_NO_GRAPH if args.no_graph else 0)
                  | (sb.FLAG_AA if args.aa else 0) | (sb.FLAG_NO_FUSE if args.no_fuse else 0))

    def reduce_max(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Workload:
        def __init__(self, lattice, config):
            self.nx, self.ny, self.name, self.kind, self.config = workload_for(args.gpus, lattice, config)
            if self.kind == "frames" and world > 1:
                raise SystemExit("configs 1 and 5 (tracer particles) are single-GPU")
            self.porous = self.config == 5
            self.setting = sb.SettingObj(animation_type=W.POISEUILLE, particles_count=1000000 if self.porous else 10000)
            self.canvas = (self.nx * 2, self.ny * 2)
            self.preset = sb.PRESET_POROUS if self.porous else W.POISEUILLE
            self.sites = self.nx * self.ny

        def make_sim(self, flags):
            """(slab or None, node, FluidSimulator or None)"""
            if self.kind == "frames":
                fs = sb.FluidSimulator(self.canvas, self.setting, particles=True, lattice=(self.nx, self.ny),
                                       device_preset=self.preset, device=local_rank, flags=flags)
                return None, fs.fluid_compute_node, fs
            if world == 1:
                return None, sb.D2Q9Node(self.canvas, self.setting, lattice=(self.nx, self.ny), device_preset=self.preset,
                                         device=local_rank, flags=flags), None
            sl = SlabRank(self.canvas, self.setting, lattice=(self.nx, self.ny), dist=dist, device=local_rank,
                          device_preset=self.preset, flags=flags)
            return sl, sl.node, None

        def advance(self, n, steps):
            if self.kind == "frames":
                n.compute_frames(steps // 2)   # one frame = 2 lattice updates + 2 particle updates
            else:
                n.step_n(steps)

    def barrier(n):
        n.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def measure_device(wl, flags, with_details):
        """W warm-up + K timed updates, state resident in HBM; CUDA events on the library's stream, max over ranks."""
        steps = args.steps - (args.steps % 2) if wl.kind == "frames" else args.steps
        slab, node, fs = wl.make_sim(flags)
        sampler = ClockSampler(local_rank)
        sampler.start()
        wl.advance(node, args.warmup + (args.warmup % 2))
        barrier(node)
        launches0, sweeps0 = node.launch_count, node.fused_sweep_count
        wait0 = node.edge_wait_stats() if world > 1 else (0, 0)
        t0 = time.perf_counter()
        wl.advance(node, steps)          # CUDA events recorded around the launches on the library's stream
        ms = node.last_step_n_ms()       # synchronises on the end event
        barrier(node)
        t1 = time.perf_counter()
        r = {"steps": steps, "launches": node.launch_count - launches0, "sweeps": node.fused_sweep_count - sweeps0,
             "clocks": sampler.stop(t0, t1), "ms": reduce_max(ms), "host_wall_ms": (t1 - t0) * 1e3}
        if world > 1:
            wait_ns, waits = node.edge_wait_stats()
            r["edge_wait"] = {"cta_ms_summed_max_over_ranks": reduce_max((wait_ns - wait0[0]) * 1e-6),
                              "waits": waits - wait0[1],
                              "mean_us_per_wait": ((wait_ns - wait0[0]) / max(waits - wait0[1], 1)) * 1e-3}
        r["value"] = wl.sites * steps / (r["ms"] * 1e-3) / 1e6
        if with_details:
            r["mass"] = slab.total_mass() if slab is not None else node.total_mass()
            # fluid-only rate for masked lattices: non-solid sites from the owned rows' LatticeInfo
            mat = node.read_lattice_info()["material"]
            fluid_sites = int(((mat != W.BOUNDARY) & (mat != W.OBSTACLE)).sum())
            del mat
            if dist is not None:
                t = torch.tensor([fluid_sites], dtype=torch.int64, device="cuda")
                dist.all_reduce(t)
                fluid_sites = int(t.item())
            r["fluid_sites"] = fluid_sites
        barrier(node)  # no slab may unmap memory a neighbour still reads
        node.close()
        return r

    def roofline_of(wl, r, overridden):
        """Roofline of the dominant kernel: ALGORITHMIC bytes (72 B per site per update) of one launch over
        its average duration.  k_frame2 performs two updates per launch while moving the distributions once, so its
        algorithmic figure can exceed the HBM peak; `traffic` (ncu DRAM bytes per launch) and `dram_frac` (those bytes
        over the same duration against the same peak) show what the memory system really carried."""
        peak, peak_src = measured_hbm_peak()
        steps_kind = wl.kind == "steps"
        fused = r["sweeps"] > 0
        lattice_launches = r["launches"] if steps_kind else (r["sweeps"] if fused else r["steps"])
        updates_per_launch = r["steps"] / lattice_launches if lattice_launches else 1.0
        per_launch_s = r["ms"] * 1e-3 / max(lattice_launches, 1)
        alg = BYTES_PER_SITE * (wl.sites / world) * updates_per_launch
        achieved = alg / per_launch_s / 1e9
        traffic, traffic_src = measured_traffic(wl.config, overridden, fused, world)
        return fused, updates_per_launch, {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": int(alg), "launch_us": per_launch_s * 1e6,
            "dram_frac": (traffic / per_launch_s / 1e9 / peak) if traffic else None,
            "frac_of_nominal_8TBs": achieved / 8000.0}

    def kernel_name(fused):
        return "k_step_generic" if args.generic else ("k_aa_pull/k_aa_local" if args.aa else
                                                     ("k_frame2 (two updates per launch)" if fused else "k_step_vec"))

    wl = Workload(args.lattice, args.config)
    nx, ny, sites, kind, config = wl.nx, wl.ny, wl.sites, wl.kind, wl.config
    main_r = measure_device(wl, base_flags, True)
    overridden = bool(args.lattice) or args.aa or args.generic
    # §8d "macro output off and on": the same updates on a handle that writes the RGBA16F macro texture (the renderer's
    # configuration, collide_stream.wgsl:74), device-timed like `value`
    macro_on = None
    if kind == "steps" and not args.aa:
        m = measure_device(wl, base_flags | sb.FLAG_MACRO_EVERY_STEP, False)
        macro_on = {"value": m["value"], "unit": "MLUPS", "ms_per_step": m["ms"] / m["steps"], "gpu_launches": m["launches"],
                    "two_update_sweeps": m["sweeps"],
                    "what": "same workload, LBM_FLAG_MACRO_EVERY_STEP: the macro texture (u.x, u.y, rho, 1) as RGBA16F is "
                            "stored by the step kernel (+8 B per site per stored texture)"}

    # ---- e2e: the same updates driven through the host API with HOST buffers in the timed region.
    # The handle is the tracer/renderer configuration (macro texture written by every step, like
    # collide_stream.wgsl:74); every rank uploads a patch of its slab and reads its slab's field.
    e2e = None
    if args.e2e_steps > 0:
        from simuverse_b200._capi import check, lib
        from simuverse_b200.wire import ptr

        slab2, node2, fs2 = wl.make_sim(base_flags | sb.FLAG_MACRO_EVERY_STEP)
        rows = min(56, node2.rows)
        patch_t = torch.empty(rows * nx * 16, dtype=torch.uint8).pin_memory()
        macro_t = torch.empty(node2.rows * nx * 8, dtype=torch.uint8).pin_memory()
        macro_t2 = torch.empty(node2.rows * nx * 8, dtype=torch.uint8).pin_memory()
        patch = patch_t.numpy().view(W.LATTICE_INFO_DTYPE)
        macro = macro_t.numpy()
        macros = [macro, macro_t2.numpy()]
        l_lo = (node2.rows - rows) // 2
        patch[:] = node2.read_lattice_info()[l_lo * nx:(l_lo + rows) * nx]  # unchanged rows: same mask, real traffic
        off = (node2.y0 + l_lo) * nx * 16
        per_call = 2  # one host call = one FluidSimulator::compute frame = 2 lattice updates (fluid_simulator.rs:223-231)
        n_part = fs2.particles_num[0] * fs2.particles_num[1] if fs2 is not None else 0
        parts_t = torch.empty(max(n_part, 1) * 24, dtype=torch.uint8).pin_memory()
        parts = parts_t.numpy()

        def e2e_step(k=0):
            check(lib.lbm_write_lattice_info(node2._h, off, ptr(patch), patch.nbytes), node2._h)
            check(lib.lbm_compute_frames(node2._h, 1), node2._h)
            if kind == "frames":
                check(lib.lbm_particles_read(node2._h, ptr(parts), n_part), node2._h)
            # pipelined read-back: frame k's field travels to host buffer k%2 while frame k+1 is computed
            check(lib.lbm_read_macro_async(node2._h, ptr(macros[k % 2])), node2._h)

        for k in range(3):
            e2e_step(k)
        barrier(node2)
        sweeps0 = node2.fused_sweep_count
        ta = time.perf_counter()
        for k in range(args.e2e_steps):
            e2e_step(k)
        barrier(node2)  # lbm_sync: every enqueued copy has landed
        dt = reduce_max(time.perf_counter() - ta)
        d2h = int(macro.nbytes) + (n_part * 24 if kind == "frames" else 0)
        e2e = {"value": sites * per_call * args.e2e_steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(patch.nbytes) * world // per_call,
               "d2h_bytes_per_step": d2h * world // per_call,
               "steps": args.e2e_steps * per_call,
               "two_update_sweeps": node2.fused_sweep_count - sweeps0,
               "d2h_GBps_per_rank": d2h * args.e2e_steps / dt / 1e9, "host_binding": numa,
               "what": ("per host call, per rank: lbm_write_lattice_info(56-row LatticeInfo patch from pinned host memory) + "
                        + ("lbm_compute_frames(1) [= FluidSimulator::compute: 2 updates + 2 particle updates] + "
                           "lbm_particles_read + " if kind == "frames" else
                           "lbm_compute_frames(1) [= FluidSimulator::compute: 2 updates, macro texture written] + ")
                        + "lbm_read_macro_async(RGBA16F field of the slab -> pinned host memory, double-buffered: the "
                          "copy of frame k overlaps the computation of frame k+1); wall clock incl. the final "
                          "lbm_sync, max over ranks")}
        # informational second flavour: the per-frame result read back is one scalar metric (the f64 total
        # mass, like reading a loss) instead of the whole field — not PCIe-bound, but pays a reduction pass
        import ctypes as C

        mass_out = C.c_double()

        def e2e_metric_step():
            check(lib.lbm_write_lattice_info(node2._h, off, ptr(patch), patch.nbytes), node2._h)
            check(lib.lbm_compute_frames(node2._h, 1), node2._h)
            check(lib.lbm_total_mass(node2._h, lib.lbm_swap_index(node2._h), C.byref(mass_out)), node2._h)

        for _ in range(3):
            e2e_metric_step()
        barrier(node2)
        ta = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_metric_step()
        barrier(node2)
        dt2 = reduce_max(time.perf_counter() - ta)
        e2e["metric_only_variant"] = {
            "value": sites * per_call * args.e2e_steps / dt2 / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": int(patch.nbytes) * world // per_call, "d2h_bytes_per_step": 8 * world // per_call,
            "what": "same host calls, but the result read back per frame is lbm_total_mass (one f64) instead of the field"}
        barrier(node2)
        node2.close()
        del patch_t, macro_t, macro_t2, parts_t

    # ---- N = 1 only: BASELINE configs[1] (4096x4096, the single-GPU roofline target) as a second block of the same line
    secondary = None
    if world == 1 and config == 3 and not args.lattice and not args.no_secondary:
        wl2 = Workload(None, 2)
        r2 = measure_device(wl2, base_flags, True)
        fused2, upl2, roof2 = roofline_of(wl2, r2, args.aa or args.generic)
        m2 = measure_device(wl2, base_flags | sb.FLAG_MACRO_EVERY_STEP, False) if not args.aa else None
        secondary = {"config": shared_config(wl2.name, 2, wl2.nx, wl2.ny, 1), "value": r2["value"], "unit": "MLUPS",
                     "ms_per_step": r2["ms"] / r2["steps"], "steps": r2["steps"], "gpu_launches": r2["launches"],
                     "kernel": kernel_name(fused2), "updates_per_launch": upl2, "roofline": roof2, "clocks": r2["clocks"],
                     "total_mass_after": r2["mass"],
                     "macro_on": ({"value": m2["value"], "unit": "MLUPS", "two_update_sweeps": m2["sweeps"]} if m2 else None)}

    if rank == 0:
        fused, updates_per_launch, roof = roofline_of(wl, main_r, overridden)
        out = {
            "metric": METRIC, "value": main_r["value"], "unit": "MLUPS", "n_gpus": args.gpus, "steps": main_r["steps"],
            "warmup": args.warmup, "ms_per_step": main_r["ms"] / main_r["steps"], "higher_is_better": True,
            "scaling": scaling_of(config), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(wl.name, config, nx, ny, world),
            "detail": {"cuda_graphs": not args.no_graph, "kernel": kernel_name(fused),
                       "updates_per_launch": updates_per_launch,
                       "state": "AA in-place, one copy of the SoA planes" if args.aa else "A/B ping-pong SoA planes",
                       "total_mass_after": main_r["mass"], "fluid_sites": main_r["fluid_sites"],
                       "mflups_fluid_only": main_r["value"] * main_r["fluid_sites"] / sites},
            "clocks": main_r["clocks"],
            "gpu_launches": main_r["launches"],
            "roofline": roof,
            "host_wall_ms_per_step": main_r["host_wall_ms"] / main_r["steps"],
        }
        if macro_on is not None:
            out["macro_on"] = macro_on
        if parity is not None:
            out["multirank_parity"] = parity
        if "edge_wait" in main_r:
            out["edge_wait"] = main_r["edge_wait"]
        if e2e is not None:
            out["e2e"] = e2e
        if secondary is not None:
            out["secondary"] = secondary
        if world == 1 and args.cpu_seconds > 0:
            out["cpu_baseline"] = cpu_leg(nx, ny, args.cpu_seconds, wl.porous)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
