#!/usr/bin/env python
"""bench.py — MLUPS of the D2Q9 lattice-Boltzmann step on N B200s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One bench "step" = one lattice update (pull-stream + BGK collide + bounce-back, the reference's
compute_by_pass) of the whole lattice.  N=1: BASELINE configs[1], 4096x4096 channel + cylinders.
N>1 (launched by torch.distributed.run, one rank per GPU): configs[2], 16384x16384 split into N
y-slabs (strong scaling); the slabs exchange edge rows through peer memory inside the step kernel,
there is no collective on the data path.

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


def measured_traffic(config, overridden, fused=False):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = str(config) + ("_fused" if fused else "")
        if not overridden and key in t:
            return t[key]["bytes_per_launch"], t[key]["source"]
    except Exception:
        pass
    return None, None


def workload_for(n_gpus, override, config=None):
    """(nx, ny, name, kind) of the BASELINE.json config being run. kind: 'steps' or 'frames'."""
    if config is None:
        config = 2 if n_gpus == 1 else 3
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


def cpu_leg(nx, ny, budget_s, porous=False, min_steps=3):
    """Times the CPU oracle (OpenMP, all host cores) on rows of the same workload. Returns a dict."""
    import numpy as np

    import oracle as orc

    cores = orc.lib().orc_get_max_threads()
    info = orc.init_porous_material(nx, ny) if porous else orc.init_lattice_material(nx, ny, 4)
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    sim = orc.OracleSim(nx, ny, info, orc.uniform_new(tau, 0, (nx * ny) & 0x7FFFFFFF), threads=cores)
    sim.step(1)
    t0 = time.perf_counter()
    n = 0
    while n < min_steps or time.perf_counter() - t0 < budget_s:
        sim.step(1)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": nx * ny * n / dt / 1e6, "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"{nx}x{ny} lattice, {n} steps in {dt:.1f} s (CPU restatement of the reference WGSL, OpenMP)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own step on the host cores.  The reference (Rust + WGSL via
    wgpu) cannot be built in this image, so this is the oracle port; each step updates a bounded
    band of the workload's lattice so that K+W steps end within a few minutes."""
    if rank != 0:
        return
    import numpy as np

    import oracle as orc

    nx, ny, name, _, _ = workload_for(args.gpus, args.lattice, args.config)
    cores = orc.lib().orc_get_max_threads()
    tau = float(np.float32(3.0) * np.float32(0.02) + np.float32(0.5))
    # calibrate on a thin band, then size the band for ~150 s total
    rows = 64
    make_info = (lambda r: orc.init_porous_material(nx, r)) if args.config == 5 else (lambda r: orc.init_lattice_material(nx, r, 4))
    info = make_info(rows)
    sim = orc.OracleSim(nx, rows, info, orc.uniform_new(tau, 0, nx * rows), threads=cores)
    sim.step(1)
    t0 = time.perf_counter()
    sim.step(3)
    per_row = (time.perf_counter() - t0) / 3 / rows
    total = args.steps + args.warmup
    rows = int(max(64, min(ny, 150.0 / total / per_row)))
    info = make_info(rows)
    sim = orc.OracleSim(nx, rows, info, orc.uniform_new(tau, 0, nx * rows), threads=cores)
    sim.step(args.warmup)
    t0 = time.perf_counter()
    sim.step(args.steps)
    dt = time.perf_counter() - t0
    v = nx * rows * args.steps / dt / 1e6
    sample = f"{nx}x{rows} band of the {nx}x{ny} lattice per step, {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MLUPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak" if (args.gpus == 1 or args.config == 4) else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=None, help="BASELINE.json config 1..5 (default 2 at N=1, 3 at N>1)")
    ap.add_argument("--aa", action="store_true", help="AA-pattern in-place variant (single copy of the distributions)")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of CUDA graphs")
    ap.add_argument("--lattice", type=int, nargs=2, default=None, metavar=("NX", "NY"), help="override the workload")
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--generic", action="store_true", help="time the one-thread-per-cell kernel instead")
    ap.add_argument("--no-fuse", action="store_true", help="one lattice update per launch (k_step_vec) instead of two (k_frame2)")
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
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nx, ny, name, kind, config = workload_for(args.gpus, args.lattice, args.config)
    if kind == "frames" and world > 1:
        raise SystemExit("configs 1 and 5 (tracer particles) are single-GPU")
    porous = config == 5
    setting = sb.SettingObj(animation_type=W.POISEUILLE, particles_count=1000000 if porous else 10000)
    canvas = (nx * 2, ny * 2)
    preset = sb.PRESET_POROUS if porous else W.POISEUILLE
    base_flags = ((sb.FLAG_KERNEL_GENERIC if args.generic else 0) | (sb.FLAG_NO_GRAPH if args.no_graph else 0)
                  | (sb.FLAG_AA if args.aa else 0) | (sb.FLAG_NO_FUSE if args.no_fuse else 0))

    def make_sim(flags):
        """(slab or None, node, FluidSimulator or None)"""
        if kind == "frames":
            fs = sb.FluidSimulator(canvas, setting, particles=True, lattice=(nx, ny), device_preset=preset,
                                   device=local_rank, flags=flags)
            return None, fs.fluid_compute_node, fs
        if world == 1:
            return None, sb.D2Q9Node(canvas, setting, lattice=(nx, ny), device_preset=preset, device=local_rank,
                                     flags=flags), None
        sl = SlabRank(canvas, setting, lattice=(nx, ny), dist=dist, device=local_rank, device_preset=preset, flags=flags)
        return sl, sl.node, None

    def barrier(n):
        n.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def advance(n, steps):
        if kind == "frames":
            n.compute_frames(steps // 2)   # one frame = 2 lattice updates + 2 particle updates
        else:
            n.step_n(steps)

    steps = args.steps - (args.steps % 2) if kind == "frames" else args.steps
    slab, node, fs = make_sim(base_flags)
    sampler = ClockSampler(local_rank)
    sampler.start()
    advance(node, args.warmup + (args.warmup % 2))
    barrier(node)
    launches0 = node.launch_count
    sweeps0 = node.fused_sweep_count
    t0 = time.perf_counter()
    advance(node, steps)             # CUDA events recorded around the launches on the library's stream
    ms = node.last_step_n_ms()       # synchronises on the end event
    barrier(node)
    t1 = time.perf_counter()
    launches = node.launch_count - launches0
    sweeps = node.fused_sweep_count - sweeps0   # launches that advanced the lattice by two updates (k_frame2)
    clocks = sampler.stop(t0, t1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    sites = nx * ny
    value = sites * steps / (ms * 1e-3) / 1e6
    mass = slab.total_mass() if slab is not None else node.total_mass()
    # fluid-only rate for masked lattices: non-solid sites from the owned rows' LatticeInfo
    mat = node.read_lattice_info()["material"]
    fluid_sites = int(((mat != W.BOUNDARY) & (mat != W.OBSTACLE)).sum())
    del mat
    if dist is not None:
        t = torch.tensor([fluid_sites], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        fluid_sites = int(t.item())
    barrier(node)  # no slab may unmap memory a neighbour still reads
    node.close()

    # ---- e2e: the same updates driven through the host API with HOST buffers in the timed region.
    # The handle is the tracer/renderer configuration (macro texture written by every step, like
    # collide_stream.wgsl:74); every rank uploads a patch of its slab and reads its slab's field.
    e2e = None
    if args.e2e_steps > 0:
        from simuverse_b200._capi import check, lib
        from simuverse_b200.wire import ptr

        slab2, node2, fs2 = make_sim(base_flags | sb.FLAG_MACRO_EVERY_STEP)
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
        ta = time.perf_counter()
        for k in range(args.e2e_steps):
            e2e_step(k)
        barrier(node2)  # lbm_sync: every enqueued copy has landed
        dt = time.perf_counter() - ta
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": sites * per_call * args.e2e_steps / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(patch.nbytes) * world // per_call,
               "d2h_bytes_per_step": (int(macro.nbytes) + (n_part * 24 if kind == "frames" else 0)) * world // per_call,
               "steps": args.e2e_steps * per_call,
               "what": ("per host call, per rank: lbm_write_lattice_info(56-row LatticeInfo patch from pinned host memory) + "
                        + ("lbm_compute_frames(1) [= FluidSimulator::compute: 2 updates + 2 particle updates] + "
                           "lbm_particles_read + " if kind == "frames" else
                           "lbm_compute_frames(1) [= FluidSimulator::compute: 2 updates, macro texture written by "
                           "each] + ")
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
        dt2 = time.perf_counter() - ta
        if dist is not None:
            t = torch.tensor([dt2], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt2 = float(t.item())
        e2e["metric_only_variant"] = {
            "value": sites * per_call * args.e2e_steps / dt2 / 1e6, "unit": "MLUPS",
            "h2d_bytes_per_step": int(patch.nbytes) * world // per_call, "d2h_bytes_per_step": 8 * world // per_call,
            "what": "same host calls, but the result read back per frame is lbm_total_mass (one f64) instead of the field"}
        barrier(node2)
        node2.close()

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        # Roofline of the dominant kernel: ALGORITHMIC bytes (72 B per site per update) of one launch over
        # its average duration.  k_frame2 performs two updates per launch while moving the distributions once, so
        # its algorithmic figure can exceed the HBM peak; `traffic` (ncu DRAM bytes per launch) and `dram_frac`
        # (those bytes over the same duration against the same peak) show what the memory system really carried.
        fused = sweeps > 0 and kind == "steps"
        updates_per_launch = steps / launches if (launches and kind == "steps") else 1.0
        per_launch_s = ms * 1e-3 / (launches if (launches and kind == "steps") else steps)
        alg_bytes_per_launch = BYTES_PER_SITE * (sites / world) * updates_per_launch
        achieved = alg_bytes_per_launch / per_launch_s / 1e9
        traffic, traffic_src = measured_traffic(config, bool(args.lattice) or world > 1 or args.aa or args.generic, fused)
        out = {
            "metric": METRIC, "value": value, "unit": "MLUPS", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak" if (world == 1 or config == 4) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "baseline_config": config, "lattice": [nx, ny], "tau": 0.56,
                       "l2": "inputs_larger_than_l2" if sites * 72 / world > 2.6e8 else "lattice fits in L2 (the reference's own size)",
                       "cuda_graphs": not args.no_graph,
                       "kernel": "k_step_generic" if args.generic else ("k_aa_pull/k_aa_local" if args.aa else
                                                                        ("k_frame2 (two updates per launch)" if fused else "k_step_vec")),
                       "updates_per_launch": updates_per_launch,
                       "state": "AA in-place, one copy of the SoA planes" if args.aa else "A/B ping-pong SoA planes",
                       "total_mass_after": mass, "fluid_sites": fluid_sites,
                       "mflups_fluid_only": value * fluid_sites / sites},
            "clocks": clocks,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes_per_launch),
                         "launch_us": per_launch_s * 1e6,
                         "dram_frac": (traffic / per_launch_s / 1e9 / peak) if traffic else None,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "host_wall_ms_per_step": (t1 - t0) * 1e3 / steps,
        }
        if e2e is not None:
            out["e2e"] = e2e
        if world == 1 and args.cpu_seconds > 0:
            out["cpu_baseline"] = cpu_leg(nx, ny, args.cpu_seconds, porous)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
