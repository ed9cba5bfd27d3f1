"""The reference's real frame loop on the two-update sweep (round 2).

`FluidSimulator::compute` (fluid_simulator.rs:217-232) runs `particle_update` after EACH lattice update and
`collide_stream.wgsl:74` stores the macro texture in every update.  The sweep kernel (csrc/lbm_fused.cuh) now stores
the texture of t+2 — and, when tracer particles exist, of t+1 into a second texture — so that a frame is ONE launch of
k_frame2 followed by the two particle passes.  Everything here is bit equality against the single-update kernels
(LBM_FLAG_NO_FUSE) and the CPU oracle, and every test checks through lbm_fused_sweep_count that sweeps really ran.
Also here: the interactive path (lbm_write_lattice_info is O(bytes written), asynchronous) and the multi-slab rules
for resuming sweeps after mask edits.
"""
import time

import numpy as np
import pytest

import simuverse_b200 as sb
from helpers import assert_bits_equal, tau_default
from simuverse_b200 import wire as W
from test_gpu_fused import live_slots, node_for, oracle_for, random_mask, setting

pytestmark = pytest.mark.gpu


def fluid_sim(nx, ny, info, flags, num=(40, 25), seed=7):
    s = setting(W.POISEUILLE)
    fs = sb.FluidSimulator((nx * 2, ny * 2), s, particles=True, lattice=(nx, ny), lattice_info=info, flags=flags,
                           particle_seed=seed)
    node = fs.fluid_compute_node
    pu = s.particles_uniform_data
    pu.num[:] = list(num)
    node.write_particle_uniform(pu)
    node.write_particles(sb.init_trajectory_particles((nx * 2, ny * 2), num, pu.life_time, seed))
    return fs, node, num[0] * num[1]


def mask_with_forces(orc, nx, ny, seed, solid):
    info = random_mask(orc, nx, ny, W.POISEUILLE, seed=seed, solid=solid, forces=8)
    g = info.reshape(ny, nx)
    ys, xs = np.nonzero(g["material"] == W.BULK)
    # force components that exercise the sign of a zero numerator in u = force * 0.5 / rho (the texel stores u)
    g[ys[10], xs[10]] = (W.EXTERNAL_FORCE, -1, -0.0, 0.05)
    g[ys[200], xs[200]] = (W.EXTERNAL_FORCE, -1, 0.03, -0.0)
    g[ys[400], xs[400]] = (W.EXTERNAL_FORCE, -1, -0.04, -1e-39)  # denormal numerator: the exact-division fallback
    return info


@pytest.mark.parametrize("nx,ny,solid", [(600, 375, 0.0), (244, 122, 0.06), (120, 40, 0.30)])
def test_particle_frames_on_sweeps_equal_single_update_frames(orc, nx, ny, solid):
    info = mask_with_forces(orc, nx, ny, seed=nx, solid=solid)
    fa, a, n = fluid_sim(nx, ny, info, 0)
    fb, b, _ = fluid_sim(nx, ny, info, sb.FLAG_NO_FUSE)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    frames = 0
    # single calls; captured runs of 4 (4 <= k < 8) and 8 frames, whose particle passes run on a side stream beside the
    # next sweep (two alternating texture sets), plus the frames left over after the last run
    for k in (1, 2, 5, 9, 12):
        fa.compute(k)
        fb.compute(k)
        frames += k
        sim.step(2 * k)
        assert a.read_particles(n).tobytes() == b.read_particles(n).tobytes(), f"particles after {frames} frames"
        np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16), b.read_macro_tex().view(np.uint16))
        np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
        for which in (0, 1):
            assert_bits_equal(a.read_distributions(which), sim.distributions(which), f"{frames} frames buf{which}")
        ca, cb = a.read_canvas().reshape(-1), b.read_canvas().reshape(-1)
        assert ((ca["alpha"] != 0) == (cb["alpha"] != 0)).all()
    assert a.fused_sweep_count == frames, "frames with tracer particles did not run as sweeps"
    assert b.fused_sweep_count == 0
    a.close()
    b.close()


@pytest.mark.parametrize("nx,ny,preset,solid", [(600, 375, W.POISEUILLE, 0.0), (256, 96, W.CUSTOM, 0.10),
                                                (128, 128, W.LID_DRIVEN_CAVITY, 0.02), (244, 61, W.POISEUILLE, 0.30)])
def test_macro_texture_written_by_sweeps(orc, nx, ny, preset, solid):
    """LBM_FLAG_MACRO_EVERY_STEP without particles (a renderer): sweeps store the texture of every second update —
    the one in between is overwritten before anything could read it (collide_stream.wgsl:74)."""
    info = random_mask(orc, nx, ny, preset, seed=3, solid=solid) if solid else orc.init_lattice_material(nx, ny, preset)
    a = node_for(nx, ny, preset, info, flags=sb.FLAG_MACRO_EVERY_STEP)
    sim = oracle_for(orc, nx, ny, preset, info)
    total = 0
    for n in (2, 7, 40, 64, 3):
        a.step_n(n)
        sim.step(n)
        total += n
        np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16, err_msg=f"{total} updates")
        assert_bits_equal(a.read_distributions(sim.swap), sim.distributions(sim.swap), f"{total} updates")
    assert a.fused_sweep_count >= total // 2 - 3
    a.compute_frames(5)
    sim.step(10)
    np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16).reshape(-1), sim.macro_f16)
    assert_bits_equal(a.read_macro(), sim.macro(), "f32 field on demand from a texture handle")
    a.close()


def test_pipelined_field_readback_from_sweeps(orc):
    """bench.py's e2e loop: mask patch, one frame, asynchronous read-back of the texture — with sweeps."""
    import torch

    nx, ny = 512, 256
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info, flags=sb.FLAG_MACRO_EVERY_STEP)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    bufs = [torch.empty((ny, nx, 4), dtype=torch.float16).pin_memory().numpy() for _ in range(2)]
    patch = info[100 * nx:156 * nx].copy()
    for k in range(6):
        a.write_lattice_info(100 * nx * 16, patch)
        a.compute_frames(1)
        a.read_macro_tex_async(bufs[k % 2])
        sim.step(2)
        if k >= 1:
            a.sync()
            np.testing.assert_array_equal(bufs[k % 2].view(np.uint16).reshape(-1), sim.macro_f16, err_msg=f"frame {k}")
    assert a.fused_sweep_count == 6
    a.close()


def test_drag_of_force_writes_is_cheap_and_exact(orc):
    """add_external_force issues one 16-byte write per sample point (d2q9_node.rs:298): each costs the host a memcpy
    and two asynchronous enqueues, no device round trip; afterwards the lattice counts the cells down with single
    updates and returns to sweeps."""
    nx, ny = 4096, 4096
    a = sb.D2Q9Node((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), device_preset=W.POISEUILLE)
    a.step_n(8)
    a.sync()
    cells = np.zeros(1, W.LATTICE_INFO_DTYPE)
    cells[0] = (W.EXTERNAL_FORCE, 6, 0.02, -0.03)
    offs = [((2000 + k) * nx + 1000 + 3 * k) * 16 for k in range(50)]
    a.write_lattice_info(offs[0], cells)  # first use allocates the staging ring
    a.sync()
    t0 = time.perf_counter()
    for off in offs:
        a.write_lattice_info(off, cells)
    dt = time.perf_counter() - t0
    a.sync()
    print(f"50 16-byte lbm_write_lattice_info calls at 4096^2: {dt * 1e3:.3f} ms host time")
    assert dt < 5e-3, f"50 force-cell writes took {dt * 1e3:.2f} ms on the host"
    before = a.fused_sweep_count
    a.step_n(30)   # 7 single updates (countdown + retire), then sweeps again
    assert a.fused_sweep_count - before >= 10
    info = a.read_lattice_info()
    assert (info["material"] == W.EXTERNAL_FORCE).sum() == 0, "armed cells did not retire"
    a.close()
    # the same sequence, small, against the oracle
    nx, ny = 256, 128
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    a.step_n(8)
    sim.step(8)
    for k in range(20):
        off = ((40 + k) * nx + 60 + 2 * k) * 16
        cells[0] = (W.EXTERNAL_FORCE, 3 + k % 4, 0.02, -0.03)
        a.write_lattice_info(off, cells)
        sim.write_lattice_info(off, cells)
    for n in (3, 4, 21):
        a.step_n(n)
        sim.step(n)
        for which in (0, 1):
            assert_bits_equal(a.read_distributions(which), sim.distributions(which), f"drag + {n}")
        assert a.read_lattice_info().tobytes() == sim.info.tobytes()
    assert a.fused_sweep_count >= 4 + 10
    a.close()


def test_update_uniforms_after_sweeps_keeps_the_previous_buffer(orc):
    """lbm_write_uniform while the non-current buffer is stale: it is recomputed with the OLD coefficients first."""
    nx, ny = 200, 120
    info = orc.init_lattice_material(nx, ny, W.POISEUILLE)
    a = node_for(nx, ny, W.POISEUILLE, info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)
    a.step_n(20)
    sim.step(20)
    assert a.fused_sweep_count == 10
    tau = orc.tau_from_viscosity(0.1)
    a.write_uniform(sb.lbm_uniform_new(tau, 0, nx * ny))
    sim.u = orc.uniform_new(tau, 0, nx * ny)
    for which in (0, 1):
        assert_bits_equal(a.read_distributions(which), sim.distributions(which), f"after the uniform write buf{which}")
    assert_bits_equal(a.read_macro(), sim.macro(), "field of the last update (old tau)")
    a.step_n(11)
    sim.step(11)
    for which in (0, 1):
        assert_bits_equal(a.read_distributions(which), sim.distributions(which), f"new tau buf{which}")
    a.close()


@pytest.mark.parametrize("n_slabs", [2, 3])
def test_slabs_resume_sweeps_after_interior_edits_and_armed_cells_on_cuts(orc, n_slabs):
    from simuverse_b200.slabs import SlabGroup

    nx, ny = 248, 96
    info = random_mask(orc, nx, ny, W.POISEUILLE, seed=11, solid=0.04)
    grp = SlabGroup((nx * 2, ny * 2), setting(W.POISEUILLE), lattice=(nx, ny), n_slabs=n_slabs, lattice_info=info)
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, info)

    def check(what, live=None):
        assert grp.swap_index == sim.swap
        for which in (0, 1):
            got, want = grp.read_distributions(which), sim.distributions(which)
            if live is not None:
                got, want = got[live], want[live]
            assert_bits_equal(got, want, f"{what} buf{which}")
        assert grp.read_lattice_info().tobytes() == sim.info.tobytes(), what

    grp.step_n(12)
    sim.step(12)
    check("start")
    sweeps = grp.nodes[0].fused_sweep_count
    assert sweeps == 6
    # (1) an obstacle painted in the interior, across a cut: sweeps go on, on every slab
    cut = ny // n_slabs
    cells = np.zeros(5, W.LATTICE_INFO_DTYPE)
    cells["material"] = W.OBSTACLE
    cells["block_iter"] = -1
    for y in (cut - 1, cut, cut + 1):
        off = (y * nx + 120) * 16
        grp.write_lattice_info(off, cells)
        sim.write_lattice_info(off, cells)
    grp.step_n(10)
    sim.step(10)
    live = live_slots(sim.info["material"].reshape(ny, nx))
    check("interior edit", live)
    assert all(n.fused_sweep_count == sweeps + 5 for n in grp.nodes), [n.fused_sweep_count for n in grp.nodes]
    sweeps += 5
    # (2) armed force cells ON the cut rows: single updates while they count down, then the halo copies retire too
    force = np.zeros(1, W.LATTICE_INFO_DTYPE)
    for y, k in ((cut, 4), (cut - 1, 6)):
        force[0] = (W.EXTERNAL_FORCE, k, 0.05, -0.04)
        off = (y * nx + 40) * 16
        grp.write_lattice_info(off, force)
        sim.write_lattice_info(off, force)
    grp.step_n(30)
    sim.step(30)
    check("armed cells on a cut, after the countdown", live)
    assert all(n.fused_sweep_count > sweeps for n in grp.nodes)
    # (3) armed again, then reset before the countdown ends (init.wgsl:51-59 retires them), then sweeps
    force[0] = (W.EXTERNAL_FORCE, 50, -0.03, 0.02)
    for y in (cut, cut - 1):
        off = (y * nx + 44) * 16
        grp.write_lattice_info(off, force)
        sim.write_lattice_info(off, force)
    grp.step_n(3)
    sim.step(3)
    grp.reset()
    sim.reset()
    sweeps = grp.nodes[0].fused_sweep_count
    grp.step_n(16)
    sim.step(16)
    check("armed on a cut, reset, sweeps")
    assert all(n.fused_sweep_count == sweeps + 8 for n in grp.nodes)
    # (4) a solid painted next to the outer ring: single updates from here on, on every slab
    cells = np.zeros(2, W.LATTICE_INFO_DTYPE)
    cells["material"] = W.OBSTACLE
    cells["block_iter"] = -1
    off = (1 * nx + 30) * 16
    grp.write_lattice_info(off, cells)
    sim.write_lattice_info(off, cells)
    sweeps = [n.fused_sweep_count for n in grp.nodes]
    grp.step_n(10)
    sim.step(10)
    live = live_slots(sim.info["material"].reshape(ny, nx))
    check("solid next to the ring", live)
    assert [n.fused_sweep_count for n in grp.nodes] == sweeps
    grp.close()


def test_curl_pass_matches_executed_wgsl_and_oracle(orc):
    """lbm_read_curl (curl_update.wgsl:12-33, the reference's never-dispatched `_curl_cal_node`)."""
    from helpers import WGSL_CURL

    g = np.load(WGSL_CURL)
    nx, ny = int(g["nx"]), int(g["ny"])
    c100 = np.load(WGSL_CURL.replace("wgsl_curl_64x48", "wgsl_channel100_64x48_s100"))
    # the channel case of the golden, 100 updates: texture handle (sweeps) and on-demand handle
    for flags in (sb.FLAG_MACRO_EVERY_STEP, 0):
        a = node_for(nx, ny, W.CUSTOM, c100["info"], flags=flags)
        a.step_n(100)
        np.testing.assert_array_equal(a.read_macro_tex().view(np.uint16), g["macro_f16"].reshape(ny, nx, 4))
        np.testing.assert_array_equal(a.read_curl_tex().view(np.uint16), g["curl_f16"], err_msg=f"flags {flags}")
        a.close()
    # a larger lattice against the oracle, odd sizes included
    for nx, ny, preset in [(600, 375, W.POISEUILLE), (131, 77, W.LID_DRIVEN_CAVITY)]:
        info = orc.init_lattice_material(nx, ny, preset)
        a = node_for(nx, ny, preset, info, flags=sb.FLAG_MACRO_EVERY_STEP)
        sim = oracle_for(orc, nx, ny, preset, info)
        a.step_n(60)
        sim.step(60)
        np.testing.assert_array_equal(a.read_curl_tex().view(np.uint16), orc.curl_update(nx, ny, sim.macro_f16))
        a.close()


def test_present_pass_matches_executed_wgsl_and_oracle(orc):
    """lbm_read_present (lbm/present.wgsl:21-46, the fragment shader of the reference's `render_node`): fragment outputs
    bit-equal to the executed shader text (golden, two target sizes) and to the oracle on larger fields, whole targets
    and row windows, from the per-step texture (sweeps) and from the on-demand field."""
    from helpers import WGSL_CURL, WGSL_PRESENT

    g = np.load(WGSL_PRESENT)
    nx, ny = int(g["nx"]), int(g["ny"])
    c100 = np.load(WGSL_CURL.replace("wgsl_curl_64x48", "wgsl_channel100_64x48_s100"))
    for tag in "ab":
        canvas = tuple(int(v) for v in g[f"canvas_{tag}"])
        for flags in (sb.FLAG_MACRO_EVERY_STEP, 0):
            a = sb.D2Q9Node(canvas, setting(W.CUSTOM), lattice=(nx, ny), lattice_info=c100["info"], flags=flags)
            a.step_n(100)
            assert_bits_equal(a.read_present(), g[f"rgba_{tag}"], f"target {canvas}, flags {flags}")
            assert_bits_equal(a.read_present(5, 9), g[f"rgba_{tag}"][5:14], "row window")
            a.close()
    for nx, ny, preset, canvas in [(600, 375, W.POISEUILLE, (1200, 750)), (131, 77, W.LID_DRIVEN_CAVITY, (1000, 333))]:
        info = orc.init_lattice_material(nx, ny, preset)
        a = sb.D2Q9Node(canvas, setting(preset), lattice=(nx, ny), lattice_info=info, flags=sb.FLAG_MACRO_EVERY_STEP)
        sim = oracle_for(orc, nx, ny, preset, info)
        a.step_n(60)
        sim.step(60)
        field = orc.field_uniform_new(nx, ny, 2, *canvas)
        want = orc.present(field, sim.macro_f16, orc.curl_update(nx, ny, sim.macro_f16))
        got = a.read_present()
        assert_bits_equal(got, want, f"{nx}x{ny} -> {canvas}")
        assert len(np.unique(got.reshape(-1, 4), axis=0)) > 1000  # a real picture, not a constant
        with pytest.raises(sb.LbmError):
            a.read_present(canvas[1] - 2, 3)
        a.close()


def test_clicks_and_drag_of_the_executed_rust_golden_through_the_device(orc):
    """The click / drag sequence of tests/golden/rust_host_helpers.npz (produced by executing the reference's Rust text)
    driven through FluidSimulator.on_click / touch_move on a real handle: the device's info buffer ends up with exactly
    the bytes the reference's write_buffer calls would have left, and the lattice then steps like the oracle's."""
    import os

    from helpers import GOLDEN_DIR

    g = np.load(os.path.join(GOLDEN_DIR, "rust_host_helpers.npz"))
    nx, ny, lps = (int(v) for v in g["lattice"])
    fs = sb.FluidSimulator((nx * lps, ny * lps), setting(W.POISEUILLE), particles=False, lattice=(nx, ny),
                           lattice_info=g["mask_0"])
    node = fs.fluid_compute_node
    want = g["mask_0"].copy()
    for pos, wrote in zip(g["clicks"], g["click_wrote"]):
        assert fs.on_click((float(pos[0]), float(pos[1]))) == bool(wrote)
    for off, patch in zip(g["click_offsets"], g["click_patches"]):
        want[int(off) // 16:int(off) // 16 + patch.size] = patch
    fs.touch_begin()
    for pos, count in zip(g["drag"], g["drag_write_counts"]):
        assert fs.touch_move((float(pos[0]), float(pos[1]))) == int(count)
    for off, cell in zip(g["drag_offsets"], g["drag_cells"]):
        want[int(off) // 16] = cell
    assert node.read_lattice_info().tobytes() == want.tobytes()
    sim = oracle_for(orc, nx, ny, W.POISEUILLE, g["mask_0"])
    sim.write_lattice_info(0, want)
    node.step_n(100)   # the armed force cells (block_iter 90) count down and retire on the way
    sim.step(100)
    # (slots of the painted discs that only another solid would read keep leftovers that are racy in the reference itself)
    live = live_slots(want["material"].reshape(ny, nx))
    for which in (0, 1):
        assert_bits_equal(node.read_distributions(which)[live], sim.distributions(which)[live], f"buf{which} after the drag")
    assert node.read_lattice_info().tobytes() == sim.info.tobytes()
    node.close()
